// uniform_launch.cuh -- what a translation unit needs to LAUNCH a fused stage kernel on the uniform path:
// the path's host state, the launch configuration helpers and the launch itself.  The stage kernels are
// compiled one translation unit per (kernel form, stage) -- stage_tu.cu, see the Makefile -- so that the
// ~200 instantiations build in parallel; uniform_path.cuh reaches them through stage_launch_table().
#pragma once

#include "mmf_common.cuh"
#include "uniform_device.cuh"

#include <cstdlib>
#include <utility>
#include <vector>

namespace mmf {

// which stage-kernel form runs a stage and with how many warps per CTA:
//   'p' ping-pong low-face kernel (uniform_stage_v5.cuh), 'r' its rotate form (uniform_stage_v5r.cuh),
//   'd' the rotate form with the y exchange decoupled by one plane (uniform_stage_v6.cuh; opt-in until it
//       has been measured on the GPU), 'h' the same with ONE warp serving both halo rows (a CTA updates
//       nw-1 rows instead of nw-2; opt-in likewise), 'w' the merged-halo decoupled kernel with TWO y rows per
//       warp (uniform_stage_v7.cuh: 2 (nw-1) rows per CTA, 8 warps; opt-in likewise), '3' the older
//       high-face kernel (uniform_stage_v3.cuh, 12 warps), kept as an independent cross-check, 'b' the rotate
//       form for a box WITH BODIES (uniform_stage_v5rb.cuh, 12 warps; chosen by the path itself, never by
//       MMF_STAGE_CFG), 'c' the same with the wall cells recomputed by a small pass around the stage kernel
//       instead of a slow path inside it
struct StageShape {
    char form = 'p';
    int nw = 16;
    int lz = 0; // planes per CTA
    // y rows a CTA updates: all warps but the two halo rows; forms 'h' and 'w' serve both halo rows with
    // one warp, and every other warp of form 'w' owns two rows
    int rows() const { return form == 'w' ? 2 * (nw - 1) : form == 'h' ? nw - 1 : nw - 2; }
    StageShape() = default;
    StageShape(char f, int n) : form(f), nw(n) {}
};

struct UniformPath {
    UniformGeom g{};
    int cell_numbering = NUM_MORTON;  // raw id <-> lattice (ignored when cell_off is set)
    int iface_numbering = NUM_MORTON; // accumulation order
    int order_exact = 1;
    int *cell_off = nullptr;          // optional explicit raw id -> padded offset
    double *arr[4] = { nullptr, nullptr, nullptr, nullptr }; // U, Wa, Wb, RHS (lazy)
    int w_cur = 1;                    // which array currently holds field W
    StageShape shape[4];              // kernel form, CTA size and z chunk per stage (0 = RHS only, 1..3)
    float *cta_est = nullptr;         // per stage-3 tile: FP32 estimate of the max eigenvalue of what it wrote
    int *eig_cand = nullptr;          // [0] = number of listed tiles, then their indices
    int n_tiles3 = 0;
    bool eig_candidate = false;       // ctl->eig_next holds the max eigenvalue of the current U
    bool clamp_ff = true;             // no stage runs the v3 form: free-flow ghosts are never read
    // bodies (opt-in, MMF_UNIFORM_BODIES=1): one flag per padded cell, 1 = not solved (src/main.cpp:221-237);
    // the ghost shell repeats the flag of the cell it touches
    unsigned char *solid = nullptr;
    bool bodies = false;              // set before the arrays are laid out: selects kernel form 'b' for every stage
    // MMF_UNIFORM_BODIES=2: form 'c' instead -- flag 2 marks the fluid cells with a wall interface, which the stage
    // kernel does not store and wall_cell_update recomputes (list of their padded offsets, compact result buffer)
    bool bodies_fixup = false;
    int *wall_list = nullptr;
    int n_wall = 0;
    double *wall_compact = nullptr;
    int nbr_rank[6] = { -1, -1, -1, -1, -1, -1 };
    double *send_buf[6] = {}, *recv_buf[6] = {};
    // direct peer stores over NVLink (comm.cuh): the neighbours' state arrays and arrival flags,
    // mapped through CUDA IPC
    bool p2p = false;
    double *peer_arr[6][3] = {};
    unsigned long long *flags = nullptr;            // [6] arrival counters, written by the neighbours
    unsigned long long *peer_flags[6] = {};
    unsigned int *push_count = nullptr;             // blocks of the running push kernel that are done
    // compact ghost columns for x partition sides (XGhost): one allocation, [side 0/1][array U/Wa/Wb]
    double *xghost = nullptr;
    double *peer_xghost[6] = {};
    long long xg_fs = 0;
    unsigned long long xchg_seq = 0;
    unsigned long long arr_seq[3] = { 0, 0, 0 }; // exchange that last refreshed the ghosts of U / Wa / Wb
    bool halo_inkernel = false;       // boundary CTAs of the stage kernels wait for the neighbours themselves
    int *tile_order[4] = {};          // per stage shape: interior tiles first, tiles on a partition side last
    // the push of a stage's output runs on the communication stream, next to the interior tiles of the
    // following stage; allowed only when every stage has at least one full wave of interior tiles, so
    // that waiting boundary CTAs can never hold all SMs before the push has been scheduled
    bool push_async = false;
    cudaEvent_t ev_stage = nullptr;           // the stage whose output is to be pushed has finished
    cudaEvent_t ev_push[3] = {};              // the last push that read U / Wa / Wb has finished
    bool push_pending[3] = { false, false, false };
    void *ipc_opened[6][5] = {};
};

// ---- launch helpers -----------------------------------------------------------------------------

// compact x ghost columns are read by the XG = true instantiations, built for the default CTA shapes only
static bool uniform_use_xghost(const mmf_ctx *ctx)
{
    const UniformPath *u = ctx->uni;
    if (!(ctx->comm && u->p2p && u->halo_inkernel && u->xghost)) return false;
    if (u->nbr_rank[0] < 0 && u->nbr_rank[1] < 0) return false;
    for (int st = 0; st < 4; ++st) {
        const StageShape &sh = u->shape[st];
        if (sh.form == 'w' ? sh.nw != 8 : (sh.nw != 12 && sh.nw != 16)) return false;
    }
    return true;
}

// free-flow sides: the v5 stage kernels re-read the boundary cell instead of a ghost cell
static LoadClamp uniform_load_clamp(const UniformPath *u)
{
    const UniformGeom &g = u->g;
    const bool on = u->clamp_ff;
    LoadClamp lc;
    lc.ilo = (on && g.bc[0] == BC_FREE_FLOW) ? 0 : -1;
    lc.ihi = (on && g.bc[1] == BC_FREE_FLOW) ? g.nx - 1 : g.nx;
    lc.jlo = (on && g.bc[2] == BC_FREE_FLOW) ? 0 : -1;
    lc.jhi = (on && g.bc[3] == BC_FREE_FLOW) ? g.ny - 1 : g.ny;
    lc.klo = (on && g.bc[4] == BC_FREE_FLOW) ? 0 : -1;
    lc.khi = (on && g.bc[5] == BC_FREE_FLOW) ? g.nz - 1 : g.nz;
    return lc;
}

// opt-in to more than 48 KB of dynamic shared memory, once per kernel and device
template <typename K>
static cudaError_t stage_smem_attribute(K kern, size_t smem)
{
    static std::vector<std::pair<const void *, int>> done;
    int dev = 0;
    cudaGetDevice(&dev);
    for (const auto &d : done) if (d.first == (const void *) kern && d.second == dev) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e == cudaSuccess) done.emplace_back((const void *) kern, dev);
    return e;
}

template <typename K>
static int launch_stage_v3(mmf_ctx *ctx, K kern, int stage, const double *Sin, const double *Un, double *Out, double *d_max)
{
    UniformPath *u = ctx->uni;
    const UniformGeom &g = u->g;
    const int nw = 12, lz = u->shape[stage].lz;
    dim3 grid((g.nx + XW - 1) / XW, (g.ny + (nw - 2) - 1) / (nw - 2), (g.nz + lz - 1) / lz);
    const size_t smem = (size_t) nw * 16 * 32 * sizeof(double) + 2 * nw * sizeof(unsigned long long);
    MMF_CUDA(ctx, stage_smem_attribute(kern, smem));
    {
        ScopedLaunchTimer timer(ctx, stage);
        kern<<<grid, nw * 32, smem, ctx->stream>>>(g, Sin, Un, Out, ctx->d_ctl, d_max, lz);
    }
    MMF_LAUNCH_CHECK(ctx);
    return MMF_OK;
}

template <typename K>
static int launch_stage_k(mmf_ctx *ctx, K kern, int stage, int nw, size_t smem, const double *Sin, const double *Un, double *Out,
                          double *d_max)
{
    UniformPath *u = ctx->uni;
    const UniformGeom &g = u->g;
    const int lz = u->shape[stage].lz, rows = u->shape[stage].rows();
    dim3 grid((g.nx + XW - 1) / XW, (g.ny + rows - 1) / rows, (g.nz + lz - 1) / lz);
    MMF_CUDA(ctx, stage_smem_attribute(kern, smem));
    for (int q = 0; q < 3; ++q) {
        if (Out == u->arr[q] && u->push_pending[q]) { // the array about to be overwritten is still being pushed
            MMF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, u->ev_push[q], 0));
            u->push_pending[q] = false;
        }
    }
    HaloWait hw{};
    hw.tx = (int) grid.x; hw.ty = (int) grid.y; hw.tz = (int) grid.z;
    if (ctx->comm && u->p2p && u->halo_inkernel) {
        int a = -1;
        for (int q = 0; q < 3; ++q) if (Sin == u->arr[q]) a = q;
        for (int s = 0; s < 6; ++s) hw.mask |= (u->nbr_rank[s] >= 0) ? (1u << s) : 0u;
        if (a >= 0 && hw.mask) {
            hw.flags = u->flags;
            hw.seq = u->arr_seq[a];
            hw.tile_order = u->tile_order[stage];
            if (hw.tile_order) grid = dim3(grid.x * grid.y * grid.z, 1, 1);
        }
    }
    XGhost xg{};
    if (hw.flags && uniform_use_xghost(ctx)) { // x ghosts of partition sides come from the compact columns
        int a = 0;
        for (int q = 0; q < 3; ++q) if (Sin == u->arr[q]) a = q;
        xg.fs = u->xg_fs;
        xg.pitch = g.ny + 2;
        if (u->nbr_rank[0] >= 0) xg.lo = u->xghost + (size_t) (0 * 3 + a) * NF * u->xg_fs;
        if (u->nbr_rank[1] >= 0) xg.hi = u->xghost + (size_t) (1 * 3 + a) * NF * u->xg_fs;
    }
    {
        ScopedLaunchTimer timer(ctx, stage);
        kern<<<grid, nw * 32, smem, ctx->stream>>>(g, Sin, Un, Out, ctx->d_ctl, d_max, lz, (stage == 3) ? u->cta_est : nullptr,
                                                   uniform_load_clamp(u), hw, xg);
    }
    MMF_LAUNCH_CHECK(ctx);
    return MMF_OK;
}

// form 'b' (a box with bodies, single GPU): the 12-warp rotate-form geometry plus the flag array
template <typename K>
static int launch_stage_body(mmf_ctx *ctx, K kern, int stage, const double *Sin, const double *Un, double *Out, double *d_max)
{
    UniformPath *u = ctx->uni;
    const UniformGeom &g = u->g;
    const int nw = 12, lz = u->shape[stage].lz, rows = u->shape[stage].rows();
    if (!u->solid) return fail(ctx, MMF_ERR_INVALID, "stage-kernel form 'b' needs the flag array of a box with bodies");
    dim3 grid((g.nx + XW - 1) / XW, (g.ny + rows - 1) / rows, (g.nz + lz - 1) / lz);
    const size_t smem = (size_t) nw * 16 * 32 * sizeof(double) + 2 * nw * sizeof(unsigned long long);
    MMF_CUDA(ctx, stage_smem_attribute(kern, smem));
    HaloWait hw{};
    hw.tx = (int) grid.x; hw.ty = (int) grid.y; hw.tz = (int) grid.z;
    {
        ScopedLaunchTimer timer(ctx, stage);
        kern<<<grid, nw * 32, smem, ctx->stream>>>(g, Sin, Un, Out, ctx->d_ctl, d_max, lz, (stage == 3) ? u->cta_est : nullptr,
                                                   uniform_load_clamp(u), hw, u->solid);
    }
    MMF_LAUNCH_CHECK(ctx);
    return MMF_OK;
}

// ---- per (form, stage) launchers, defined in stage_tu.cu ----------------------------------------------
// order = NUM_MORTON / NUM_LEXI / NUM_AXIS; CTA shape and z chunk come from ctx->uni->shape[stage]
typedef int (*StageLauncher)(mmf_ctx *ctx, int order, const double *Sin, const double *Un, double *Out, double *d_max);

#define MMF_STAGE_TU_NAME_(F, S) launch_stage_##F##_##S
#define MMF_STAGE_TU_NAME(F, S) MMF_STAGE_TU_NAME_(F, S)
#define MMF_DECLARE_STAGE_TUS(F)                                                                       \
    int MMF_STAGE_TU_NAME(F, 0)(mmf_ctx *, int, const double *, const double *, double *, double *);   \
    int MMF_STAGE_TU_NAME(F, 1)(mmf_ctx *, int, const double *, const double *, double *, double *);   \
    int MMF_STAGE_TU_NAME(F, 2)(mmf_ctx *, int, const double *, const double *, double *, double *);   \
    int MMF_STAGE_TU_NAME(F, 3)(mmf_ctx *, int, const double *, const double *, double *, double *);
MMF_DECLARE_STAGE_TUS(p)  // uniform_stage_v5.cuh
MMF_DECLARE_STAGE_TUS(r)  // uniform_stage_v5r.cuh
MMF_DECLARE_STAGE_TUS(d)  // uniform_stage_v6.cuh
MMF_DECLARE_STAGE_TUS(h)  // uniform_stage_v6.cuh, one warp for both halo rows
MMF_DECLARE_STAGE_TUS(t)  // uniform_stage_v3.cuh (form '3')
MMF_DECLARE_STAGE_TUS(w)  // uniform_stage_v7.cuh, two y rows per warp
MMF_DECLARE_STAGE_TUS(b)  // uniform_stage_v5rb.cuh, a box with bodies
MMF_DECLARE_STAGE_TUS(c)  // uniform_stage_v5rb.cuh, a box with bodies, wall cells by a fix-up pass
#undef MMF_DECLARE_STAGE_TUS

// the launcher of a kernel form ('p', 'r', 'd', 'h', 'w', '3', 'b', 'c') for a stage
inline StageLauncher stage_launcher(char form, int stage)
{
    static const StageLauncher tab[8][4] = {
        { launch_stage_p_0, launch_stage_p_1, launch_stage_p_2, launch_stage_p_3 },
        { launch_stage_r_0, launch_stage_r_1, launch_stage_r_2, launch_stage_r_3 },
        { launch_stage_d_0, launch_stage_d_1, launch_stage_d_2, launch_stage_d_3 },
        { launch_stage_t_0, launch_stage_t_1, launch_stage_t_2, launch_stage_t_3 },
        { launch_stage_h_0, launch_stage_h_1, launch_stage_h_2, launch_stage_h_3 },
        { launch_stage_w_0, launch_stage_w_1, launch_stage_w_2, launch_stage_w_3 },
        { launch_stage_b_0, launch_stage_b_1, launch_stage_b_2, launch_stage_b_3 },
        { launch_stage_c_0, launch_stage_c_1, launch_stage_c_2, launch_stage_c_3 },
    };
    const int f = (form == 'r') ? 1 : (form == 'd') ? 2 : (form == '3') ? 3 : (form == 'h') ? 4 : (form == 'w') ? 5 : (form == 'b') ? 6 : (form == 'c') ? 7 : 0;
    return tab[f][stage];
}

} // namespace mmf
