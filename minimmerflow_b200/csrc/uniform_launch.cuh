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
//   'r' the rotate form of the low-face streaming kernel (uniform_stage_v5r.cuh), the default: 12 warps
//   't' the same scheme with the input staged in shared memory by bulk tensor (TMA) loads (uniform_stage_t.cuh)
//   'h' form 't' with ONE warp serving both halo rows of the tile: nw - 1 update rows per CTA
//   'b' form 'h' for a box WITH BODIES (12 warps; chosen by the path itself, never by MMF_STAGE_CFG): one flag byte per
//       cell, the wall cells recomputed by a small pass around the stage kernel
struct StageShape {
    char form = 'r';
    int nw = 12;
    int lz = 0; // planes per CTA
    int rows() const { return (form == 'h' || form == 'b') ? nw - 1 : nw - 2; } // y rows a CTA updates: all warps but the halo warp(s)
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
    // kernel form 't': bulk tensor loads; descriptors of the padded arrays per (array, rows of the box)
    struct InMap { const double *arr; int rows; TmaDesc map; };
    std::vector<InMap> in_maps;
    bool eig_candidate = false;       // ctl->eig_next holds the max eigenvalue of the current U
    cudaGraphExec_t step_graph = nullptr; // the launches of a steady-state step, captured once (uniform_path.cuh: uniform_step)
    int step_graph_launches = 0;
    bool clamp_ff = true;             // free-flow ghosts are never read by the stage kernels (LoadClamp)
    // bodies: one flag per padded cell, 1 = not solved (src/main.cpp:221-237), 2 = fluid cell with a wall interface,
    // which the stage kernel does not store and wall_cell_update recomputes (list of their padded offsets, compact
    // result buffer); the ghost shell repeats the flag of the cell it touches
    unsigned char *solid = nullptr;
    bool bodies = false;              // set before the arrays are laid out: selects kernel form 'b' for every stage
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
    // exchange by the copy engines (comm.cuh: comm_uniform_dma_push): y and z layers travel as strided peer copies, the
    // arrival counters as 8-byte copies out of a pinned ring of sequence numbers -- no SM involved
    bool dma_push = false;
    unsigned long long *seq_ring = nullptr;   // pinned host memory, DMA_SEQ_RING entries
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
        if (sh.nw != 12 && sh.nw != 16) return false;
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

// what every launch of a stage kernel starts with: the grid, the wait for a push that still reads the output array,
// the in-kernel halo wait of a partitioned run and the compact x ghost columns
struct StageLaunchArgs {
    dim3 grid;
    HaloWait hw{};
    XGhost xg{};
};

static int stage_launch_prelude(mmf_ctx *ctx, int stage, const double *Sin, double *Out, StageLaunchArgs &a)
{
    UniformPath *u = ctx->uni;
    const UniformGeom &g = u->g;
    const int lz = u->shape[stage].lz, rows = u->shape[stage].rows();
    a.grid = dim3((g.nx + XW - 1) / XW, (g.ny + rows - 1) / rows, (g.nz + lz - 1) / lz);
    for (int q = 0; q < 3; ++q) {
        if (Out == u->arr[q] && u->push_pending[q]) { // the array about to be overwritten is still being pushed
            MMF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, u->ev_push[q], 0));
            u->push_pending[q] = false;
        }
    }
    HaloWait &hw = a.hw;
    hw.tx = (int) a.grid.x; hw.ty = (int) a.grid.y; hw.tz = (int) a.grid.z;
    int ain = -1;
    for (int q = 0; q < 3; ++q) if (Sin == u->arr[q]) ain = q;
    if (ctx->comm && u->p2p && u->halo_inkernel) {
        for (int s = 0; s < 6; ++s) hw.mask |= (u->nbr_rank[s] >= 0) ? (1u << s) : 0u;
        if (ain >= 0 && hw.mask) {
            hw.flags = u->flags;
            hw.seq = u->arr_seq[ain];
            hw.tile_order = u->tile_order[stage];
            hw.timeouts = &ctx->d_ctl->halo_timeouts;
            hw.timeout_ns = halo_timeout_ns();
            if (hw.tile_order) a.grid = dim3(a.grid.x * a.grid.y * a.grid.z, 1, 1);
        }
    }
    if (hw.flags && uniform_use_xghost(ctx)) { // x ghosts of partition sides come from the compact columns
        const int q = ain < 0 ? 0 : ain;
        a.xg.fs = u->xg_fs;
        a.xg.pitch = g.ny + 2;
        if (u->nbr_rank[0] >= 0) a.xg.lo = u->xghost + (size_t) (0 * 3 + q) * NF * u->xg_fs;
        if (u->nbr_rank[1] >= 0) a.xg.hi = u->xghost + (size_t) (1 * 3 + q) * NF * u->xg_fs;
    }
    return MMF_OK;
}

template <typename K>
static int launch_stage_k(mmf_ctx *ctx, K kern, int stage, int nw, size_t smem, const double *Sin, const double *Un, double *Out,
                          double *d_max)
{
    UniformPath *u = ctx->uni;
    StageLaunchArgs a;
    MMF_CUDA(ctx, stage_smem_attribute(kern, smem));
    if (int rc = stage_launch_prelude(ctx, stage, Sin, Out, a)) return rc;
    {
        ScopedLaunchTimer timer(ctx, stage);
        kern<<<a.grid, nw * 32, smem, ctx->stream>>>(u->g, Sin, Un, Out, ctx->d_ctl, d_max, u->shape[stage].lz,
                                                     (stage == 3) ? u->cta_est : nullptr, uniform_load_clamp(u), a.hw, a.xg);
    }
    MMF_LAUNCH_CHECK(ctx);
    return MMF_OK;
}

// ---- bulk tensor loads (kernel form 't') -------------------------------------------------------------------------
// cuTensorMapEncodeTiled comes from the driver through the runtime (cudaGetDriverEntryPoint): the library does not
// link libcuda.
typedef CUresult (*TmaEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int uniform_tma_encoder(mmf_ctx *ctx, TmaEncodeFn *out)
{
    static TmaEncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        MMF_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return fail(ctx, MMF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        encode = reinterpret_cast<TmaEncodeFn>(fn);
    }
    *out = encode;
    return MMF_OK;
}

// The descriptor of one padded state array for LOADS: dimensions (x, y, z, field) = (px, py, pz, NF), i.e. the whole
// padded array with its ghost shell, element (0, 0, 0, f) = the first padded element of field f (cudaMalloc alignment;
// rows are multiples of 32 bytes, field strides of 128), and a box of 32 x `rows` x 1 x NF: one z plane of everything
// a CTA of the stage kernel reads.  Box coordinates are element coordinates and need no alignment; what sticks out
// of the array is zero filled.  Descriptors are cached per (array, rows).
static int uniform_in_map(mmf_ctx *ctx, const double *arr, int rows, const TmaDesc **out)
{
    UniformPath *u = ctx->uni;
    for (const auto &m : u->in_maps) if (m.arr == arr && m.rows == rows) { *out = &m.map; return MMF_OK; }
    TmaEncodeFn encode = nullptr;
    if (int rc = uniform_tma_encoder(ctx, &encode)) return rc;
    const UniformGeom &g = u->g;
    const cuuint64_t dims[4] = { (cuuint64_t) g.px, (cuuint64_t) g.py, (cuuint64_t) g.pz, (cuuint64_t) NF };
    const cuuint64_t strides[3] = { (cuuint64_t) g.px * 8, (cuuint64_t) g.px * g.py * 8, (cuuint64_t) g.fs * 8 };
    const cuuint32_t box[4] = { 32, (cuuint32_t) rows, 1, (cuuint32_t) NF };
    const cuuint32_t estr[4] = { 1, 1, 1, 1 };
    UniformPath::InMap m;
    m.arr = arr;
    m.rows = rows;
    const CUresult r = encode(&m.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double *>(arr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, MMF_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for loads of a %d x %d x %d box", (int) r, g.nx, g.ny, g.nz);
    u->in_maps.push_back(m);
    *out = &u->in_maps.back().map;
    return MMF_OK;
}

template <typename K>
static int launch_stage_tl(mmf_ctx *ctx, K kern, int stage, int nw, bool merged_halo, size_t smem, const double *Sin, const double *Un,
                           double *Out, double *d_max, const unsigned char *solid = nullptr)
{
    UniformPath *u = ctx->uni;
    const int nu = merged_halo ? nw - 1 : nw - 2; // update rows, tile rows = nu + 2
    const TmaDesc *smap = nullptr, *umap = nullptr;
    if (int rc = uniform_in_map(ctx, Sin, nu + 2, &smap)) return rc;
    const TmaDesc sm = *smap; // (the cache may grow below)
    if (stage >= 2) { if (int rc = uniform_in_map(ctx, Un, nu, &umap)) return rc; }
    const TmaDesc um = umap ? *umap : sm;
    StageLaunchArgs a;
    MMF_CUDA(ctx, stage_smem_attribute(kern, smem));
    if (int rc = stage_launch_prelude(ctx, stage, Sin, Out, a)) return rc;
    if (a.xg.lo || a.xg.hi) return fail(ctx, MMF_ERR_INVALID, "stage-kernel forms 't' / 'h' do not read compact x ghost columns");
    {
        ScopedLaunchTimer timer(ctx, stage);
        kern<<<a.grid, nw * 32, smem, ctx->stream>>>(u->g, Out, ctx->d_ctl, d_max, u->shape[stage].lz,
                                                     (stage == 3) ? u->cta_est : nullptr, uniform_load_clamp(u), a.hw, sm, um, solid);
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
MMF_DECLARE_STAGE_TUS(r)  // uniform_stage_v5r.cuh
MMF_DECLARE_STAGE_TUS(t)  // uniform_stage_t.cuh: input staged by bulk tensor loads
#undef MMF_DECLARE_STAGE_TUS

// the launcher of a kernel form ('r', 't' / 'h' / 'b') for a stage
inline StageLauncher stage_launcher(char form, int stage)
{
    static const StageLauncher tab[2][4] = {
        { launch_stage_r_0, launch_stage_r_1, launch_stage_r_2, launch_stage_r_3 },
        { launch_stage_t_0, launch_stage_t_1, launch_stage_t_2, launch_stage_t_3 },
    };
    return tab[(form == 't' || form == 'h' || form == 'b') ? 1 : 0][stage];
}

} // namespace mmf
