// uniform_path.cuh -- host side of MMF_PATH_UNIFORM: eligibility check of a host mesh description,
// padded SoA allocation, launch configuration and the per-step kernel sequence.
#pragma once

#include "uniform_launch.cuh"
#include "uniform_kernels.cuh"
#include "uniform_eligibility.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <utility>

namespace mmf {

inline int uniform_order_exact(const mmf_ctx *ctx) { return ctx->uni ? ctx->uni->order_exact : 0; }
inline void uniform_invalidate_eig(mmf_ctx *ctx) { if (ctx->uni) ctx->uni->eig_candidate = false; }

inline void uniform_destroy(mmf_ctx *ctx)
{
    if (ctx->uni) {
        if (ctx->uni->step_graph) cudaGraphExecDestroy(ctx->uni->step_graph);
        if (ctx->uni->ev_stage) cudaEventDestroy(ctx->uni->ev_stage);
        for (cudaEvent_t e : ctx->uni->ev_push) if (e) cudaEventDestroy(e);
    }
    delete ctx->uni;
    ctx->uni = nullptr;
}

static inline double *uniform_field_ptr(mmf_ctx *ctx, int field)
{
    UniformPath *u = ctx->uni;
    if (field == MMF_FIELD_U) return u->arr[0];
    if (field == MMF_FIELD_W) return u->arr[u->w_cur];
    return u->arr[3];
}

static int uniform_ensure_rhs(mmf_ctx *ctx)
{
    UniformPath *u = ctx->uni;
    if (u->arr[3]) return MMF_OK;
    int rc = dev_alloc(ctx, &u->arr[3], (size_t) NF * u->g.fs);
    if (rc) return rc;
    MMF_CUDA(ctx, cudaMemsetAsync(u->arr[3], 0, sizeof(double) * NF * u->g.fs, ctx->stream));
    return MMF_OK;
}

// visiting order of the tiles of every stage shape: tiles that touch no partition side first
static int uniform_build_tile_orders(mmf_ctx *ctx)
{
    UniformPath *u = ctx->uni;
    const UniformGeom &g = u->g;
    u->push_async = u->halo_inkernel && !(getenv("MMF_PUSH_SYNC") && atoi(getenv("MMF_PUSH_SYNC")));
    for (int st = 0; st < 4; ++st) {
        const StageShape &sh = u->shape[st];
        const int tx = (g.nx + XW - 1) / XW, ty = (g.ny + sh.rows() - 1) / sh.rows(), tz = (g.nz + sh.lz - 1) / sh.lz;
        std::vector<int> inner, outer;
        for (int t = 0; t < tx * ty * tz; ++t) {
            const int bx = t % tx, by = (t / tx) % ty, bz = t / (tx * ty);
            const bool side[6] = { bx == 0, bx == tx - 1, by == 0, by == ty - 1, bz == 0, bz == tz - 1 };
            bool waits = false;
            for (int s = 0; s < 6; ++s) waits = waits || (side[s] && u->nbr_rank[s] >= 0);
            (waits ? outer : inner).push_back(t);
        }
        // (a push KERNEL needs an SM: waiting boundary CTAs must never hold all of them before it has been scheduled;
        //  the copy engines do not care)
        if (st >= 1 && !u->dma_push) u->push_async = u->push_async && (int) inner.size() >= ctx->prop.multiProcessorCount;
        inner.insert(inner.end(), outer.begin(), outer.end());
        int rc = dev_upload(ctx, &u->tile_order[st], inner);
        if (rc) return rc;
    }
    if (u->push_async && !u->ev_stage) {
        MMF_CUDA(ctx, cudaEventCreateWithFlags(&u->ev_stage, cudaEventDisableTiming));
        for (int a = 0; a < 3; ++a) MMF_CUDA(ctx, cudaEventCreateWithFlags(&u->ev_push[a], cudaEventDisableTiming));
    }
    return MMF_OK;
}

template <int STAGE>
static int launch_stage(mmf_ctx *ctx, const double *Sin, const double *Un, double *Out, double *d_max)
{
    const UniformPath *u = ctx->uni;
    return stage_launcher(u->shape[STAGE].form, STAGE)(ctx, u->iface_numbering, Sin, Un, Out, d_max);
}

int comm_uniform_exchange_enqueue(mmf_ctx *ctx, double *S, bool defer_wait); // comm.cuh

// refresh the ghost shell of a padded array: physical sides from the BC, partition sides by exchange
static int uniform_refresh_ghosts(mmf_ctx *ctx, double *S, int check_active, double *eig_next = nullptr)
{
    const UniformGeom &g = ctx->uni->g;
    // in a fused step (check_active) free-flow sides need no ghosts: the stage kernels clamp their loads
    const int skip_ff = (check_active && ctx->uni->clamp_ff) ? 1 : 0;
    bool any = false; // a physical side that needs the boundary-condition pass
    for (int s = 0; s < 6; ++s) any = any || (g.bc[s] >= 0 && !(skip_ff && g.bc[s] == BC_FREE_FLOW));
    if (any) {
        const int na = std::max(g.nx, g.ny), nb = std::max(g.ny, g.nz);
        dim3 grid((na + 255) / 256, nb, 6);
        uniform_ghost_kernel<<<grid, 256, 0, ctx->stream>>>(g, S, ctx->d_ctl, check_active, eig_next, skip_ff, ctx->uni->solid);
        MMF_LAUNCH_CHECK(ctx);
    }
    trace_point(ctx, "bc");
    // inside a fused step the consumer (the next stage kernel) waits for the neighbours itself
    if (ctx->comm) return comm_uniform_exchange_enqueue(ctx, S, check_active && ctx->uni->halo_inkernel);
    return MMF_OK;
}

// ---- creation -----------------------------------------------------------------------------------

// Everything that follows from the stages' CTA shapes: the z chunk per CTA, the number of stage-3 tiles and the buffers
// of the per-tile eigenvalue estimates.  Called at creation and again if the shapes change (comm_ipc_import: a box
// with an x partition side runs the rotate form).
static int uniform_setup_shapes(mmf_ctx *ctx)
{
    UniformPath *u = ctx->uni;
    const UniformGeom &g = u->g;
    int rc;
    // z chunk per CTA: every CTA holds one SM (1 CTA/SM), so the grid runs in ceil(CTAs/SMs) rounds.
    // Pick the chunk count whose last round is fullest, charging each chunk the extra plane it
    // derives for its first z interface; chunks stay between 4 and 96 planes (a small box -- the reference's own 32^3
    // cases -- fills the SMs only with short chunks: a CTA marches through its planes one after the other, 1.8 us each).
    const char *env_lz = getenv("MMF_STAGE_LZ");
    for (int st = 0; st < 4; ++st) {
        StageShape &sh = u->shape[st];
        if (env_lz && atoi(env_lz) > 0) { sh.lz = atoi(env_lz); continue; }
        const long long tiles_xy = (long long) ((g.nx + XW - 1) / XW) * ((g.ny + sh.rows() - 1) / sh.rows());
        const double sms = (double) ctx->prop.multiProcessorCount;
        double best = -1.;
        for (int chunks = std::max(1, (g.nz + 95) / 96); chunks <= std::max(1, g.nz / 4); ++chunks) {
            const int lz = (g.nz + chunks - 1) / chunks;
            const long long n_chunks = (g.nz + lz - 1) / lz;
            const double waves = (double) (tiles_xy * n_chunks) / sms;
            const double score = waves / std::ceil(waves) * (double) lz / (lz + 1.0);
            if (score > best + 1e-9) { best = score; sh.lz = lz; }
        }
        if (sh.lz <= 0) sh.lz = g.nz;
    }
    {
        const StageShape &s3 = u->shape[3];
        u->n_tiles3 = ((g.nx + XW - 1) / XW) * ((g.ny + s3.rows() - 1) / s3.rows()) * ((g.nz + s3.lz - 1) / s3.lz);
        if ((rc = dev_alloc(ctx, &u->cta_est, (size_t) u->n_tiles3))) return rc;
        MMF_CUDA(ctx, cudaMemsetAsync(u->cta_est, 0, sizeof(float) * u->n_tiles3, ctx->stream));
        if ((rc = dev_alloc(ctx, &u->eig_cand, (size_t) u->n_tiles3 + 1))) return rc;
    }
    return MMF_OK;
}

static int uniform_alloc(mmf_ctx *ctx, UniformPath *u)
{
    UniformGeom &g = u->g;
    g.px = padded_row(g.nx);
    g.py = g.ny + 2;
    g.pz = g.nz + 2;
    g.fs = ((long long) g.px * g.py * g.pz + 15) / 16 * 16;
    // the stage kernels index one field with 32-bit element offsets (4 field strides must stay below 2^31)
    if (g.fs * (NF - 1) >= ((long long) 1 << 31)) return fail(ctx, MMF_ERR_INVALID, "uniform box too large for one GPU (> 536 M padded cells)");
    int rc;
    for (int a = 0; a < 3; ++a) {
        if ((rc = dev_alloc(ctx, &u->arr[a], (size_t) NF * g.fs))) return rc;
        fill_benign_kernel<<<grid_for(g.fs, 256), 256, 0, ctx->stream>>>(u->arr[a], g.fs);
        MMF_LAUNCH_CHECK(ctx);
    }
    // Launch shapes, measured at 256^3 on B200 (profiles/r02d_experiments.md): the form whose input is staged by bulk
    // tensor loads with ONE warp for both halo rows, at 12 warps (168 registers: previous plane carried, loop unrolled
    // by the ring depth), is the fastest kernel of every stage.  MMF_STAGE_CFG overrides per stage, e.g.
    // "r12:t16:h16:h12" (stage 0:1:2:3; 't' = two halo warps, 'r' = the rotate form with per-thread global loads,
    // which is also what runs when an x side is a partition side).
    for (int st = 0; st < 4; ++st) u->shape[st] = StageShape{ 'h', 12 };
    if (const char *cfg = getenv("MMF_STAGE_CFG")) {
        int st = 0;
        for (const char *p = cfg; *p && st < 4;) {
            StageShape sh;
            sh.form = *p++;
            sh.nw = atoi(p);
            while (*p && *p != ':') ++p;
            const bool last = (*p == 0);
            if (*p == ':') ++p;
            if ((sh.form != 'r' && sh.form != 't' && sh.form != 'h') || (sh.nw != 8 && sh.nw != 12 && sh.nw != 16)) break;
            if (sh.form == 'h' && sh.nw == 8) sh.nw = 12;
            u->shape[st++] = sh;
            if (last) { for (; st < 4; ++st) u->shape[st] = sh; } // one entry = all stages
        }
    }
    if (u->bodies) {
        // a box with bodies: the kernel form that knows about them, whatever MMF_STAGE_CFG says -- 'b', the TMA-fed kernel
        // with one warp for both halo rows plus the flag array
        for (int st = 0; st < 4; ++st) u->shape[st] = StageShape{ 'b', 12 };
    }
    u->clamp_ff = true;
    u->halo_inkernel = u->clamp_ff && !(getenv("MMF_HALO_WAIT_KERNEL") && atoi(getenv("MMF_HALO_WAIT_KERNEL")));
    if ((rc = uniform_setup_shapes(ctx))) return rc;
    MMF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMF_OK;
}

static int uniform_create(mmf_ctx *ctx, const mmf_uniform_desc *d)
{
    for (int e = 0; e < 3; ++e) {
        if (d->box_dims[e] <= 0 || d->global_dims[e] < d->box_dims[e] || d->box_offset[e] < 0 ||
            d->box_offset[e] + d->box_dims[e] > d->global_dims[e]) {
            return fail(ctx, MMF_ERR_INVALID, "mmf_create_uniform: inconsistent box along axis %d", e);
        }
    }
    if (!(d->h > 0.)) return fail(ctx, MMF_ERR_INVALID, "mmf_create_uniform: h must be positive");
    if (d->cell_numbering != MMF_NUMBERING_MORTON && d->cell_numbering != MMF_NUMBERING_LEXICOGRAPHIC) {
        return fail(ctx, MMF_ERR_INVALID, "mmf_create_uniform: cell_numbering must be MORTON or LEXICOGRAPHIC");
    }
    if (d->interface_numbering < 0 || d->interface_numbering > 2) {
        return fail(ctx, MMF_ERR_INVALID, "mmf_create_uniform: unknown interface_numbering");
    }
    if (d->cell_numbering == MMF_NUMBERING_MORTON) {
        const int n = d->box_dims[0];
        if (n != d->box_dims[1] || n != d->box_dims[2] || (n & (n - 1))) {
            return fail(ctx, MMF_ERR_INVALID, "mmf_create_uniform: Morton cell numbering needs a power-of-two cube");
        }
    }
    UniformPath *u = new UniformPath();
    ctx->uni = u;
    UniformGeom &g = u->g;
    g.nx = d->box_dims[0]; g.ny = d->box_dims[1]; g.nz = d->box_dims[2];
    g.gx0 = d->box_offset[0]; g.gy0 = d->box_offset[1]; g.gz0 = d->box_offset[2];
    g.gnx = d->global_dims[0]; g.gny = d->global_dims[1]; g.gnz = d->global_dims[2];
    g.h = d->h;
    const bool has_geom = d->struct_size >= sizeof(mmf_uniform_desc);
    if (has_geom && (d->area < 0. || d->volume < 0. || (d->area > 0.) != (d->volume > 0.))) {
        return fail(ctx, MMF_ERR_INVALID, "mmf_create_uniform: area and volume must both be given (> 0) or both be 0");
    }
    g.area = (has_geom && d->area > 0.) ? d->area : g.h * g.h;           // the host's own values verbatim, if it passes them
    g.volume = (has_geom && d->volume > 0.) ? d->volume : g.h * g.h * g.h;
    for (int s = 0; s < 6; ++s) {
        const int axis = s >> 1;
        const bool hi = s & 1;
        const bool at_border = hi ? (d->box_offset[axis] + d->box_dims[axis] == d->global_dims[axis])
                                  : (d->box_offset[axis] == 0);
        if (at_border) {
            if (d->bc_side[s] < MMF_BC_FREE_FLOW || d->bc_side[s] > MMF_BC_DIRICHLET) {
                return fail(ctx, MMF_ERR_INVALID, "mmf_create_uniform: side %d needs a boundary condition", s);
            }
            g.bc[s] = d->bc_side[s];
        } else {
            g.bc[s] = -2;
        }
    }
    memcpy(g.dirichlet, d->dirichlet_info, sizeof g.dirichlet);
    u->cell_numbering = d->cell_numbering;
    u->iface_numbering = d->interface_numbering;
    u->order_exact = (d->interface_numbering != MMF_NUMBERING_AXIS) && !(d->flags & MMF_FLAG_ORDER_AXIS);
    if (d->flags & MMF_FLAG_ORDER_AXIS) u->iface_numbering = NUM_AXIS;
    ctx->path = MMF_PATH_UNIFORM;
    ctx->n_cells = (int64_t) g.nx * g.ny * g.nz;
    ctx->n_ifaces = (int64_t) (g.nx + 1) * g.ny * g.nz + (int64_t) g.nx * (g.ny + 1) * g.nz + (int64_t) g.nx * g.ny * (g.nz + 1);
    return uniform_alloc(ctx, u);
}

// Decide whether a host mesh description is a full, conforming, uniform 3-D box whose interface numbering
// matches a known convention; if so build the uniform path from it.  All cells solved, or a box with bodies:
// cells that are not solved, and BC_WALL on exactly the interfaces between a solved and an unsolved cell
// (src/main.cpp:221-237, 251-277).  MMF_UNIFORM_BODIES=0 keeps a mesh with bodies on the generic path.
static int uniform_try_create(mmf_ctx *ctx, const mmf_mesh_desc *d, bool *used)
{
    *used = false;
    // (a handle is created before it can join a communicator: comm_set_box_neighbours rejects a box with bodies)
    const bool allow_bodies = !(getenv("MMF_UNIFORM_BODIES") && atoi(getenv("MMF_UNIFORM_BODIES")) == 0);
    const UniformBoxAnalysis an = analyze_uniform_box(d, allow_bodies);
    if (!an.eligible) return MMF_OK;
    const int nx = d->box_dims[0], ny = d->box_dims[1], nz = d->box_dims[2];
    const int64_t nc = d->n_cells;
    const bool bodies = an.bodies;
    const int numbering = an.numbering, order_exact = an.order_exact;
    const int *bc_side = an.bc_side;
    const double A = an.area, V = an.volume, h = an.h;

    UniformPath *u = new UniformPath();
    ctx->uni = u;
    UniformGeom &g = u->g;
    g.nx = nx; g.ny = ny; g.nz = nz;
    g.gx0 = g.gy0 = g.gz0 = 0;
    g.gnx = nx; g.gny = ny; g.gnz = nz;
    g.h = h;
    g.area = A;   // taken verbatim from the host tables, never rebuilt from h
    g.volume = V;
    for (int s = 0; s < 6; ++s) g.bc[s] = bc_side[s];
    memcpy(g.dirichlet, d->dirichlet_info, sizeof g.dirichlet);
    u->iface_numbering = numbering;
    u->order_exact = order_exact;
    u->bodies = bodies;
    ctx->path = MMF_PATH_UNIFORM;
    int rc = uniform_alloc(ctx, u);
    if (rc) return rc;
    if (bodies) {
        std::vector<unsigned char> flag;
        std::vector<int> walls;
        body_flags(g, nc, d->cell_ijk, d->solved, true, flag, walls);
        u->n_wall = (int) walls.size();
        if (walls.empty()) walls.push_back(0);
        if ((rc = dev_upload(ctx, &u->wall_list, walls))) return rc;
        if ((rc = dev_alloc(ctx, &u->wall_compact, (size_t) NF * walls.size()))) return rc;
        if ((rc = dev_upload(ctx, &u->solid, flag))) return rc;
    }
    std::vector<int> off((size_t) nc);
    for (int64_t c = 0; c < nc; ++c) {
        off[c] = (int) uoff(g, d->cell_ijk[3 * c], d->cell_ijk[3 * c + 1], d->cell_ijk[3 * c + 2]);
    }
    if ((rc = dev_upload(ctx, &u->cell_off, off))) return rc;
    *used = true;
    return MMF_OK;
}

// ---- state transfer -----------------------------------------------------------------------------

static int uniform_scatter_state(mmf_ctx *ctx, int field, const double *staging)
{
    UniformPath *u = ctx->uni;
    int rc;
    if (field == MMF_FIELD_RHS && (rc = uniform_ensure_rhs(ctx))) return rc;
    double *S = uniform_field_ptr(ctx, field);
    for (int q = 0; q < 3; ++q) {
        if (S == u->arr[q] && u->push_pending[q]) {
            MMF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, u->ev_push[q], 0));
            u->push_pending[q] = false;
        }
    }
    uniform_scatter_kernel<<<grid_for(ctx->n_cells, 256), 256, 0, ctx->stream>>>(
        u->g, u->cell_numbering, u->cell_off, staging, S, ctx->n_cells);
    MMF_LAUNCH_CHECK(ctx);
    if (u->bodies && field == MMF_FIELD_W) {
        // the fused stages alternate between two work arrays and never write a cell that is not solved: such
        // cells must hold the host's value in both
        uniform_scatter_kernel<<<grid_for(ctx->n_cells, 256), 256, 0, ctx->stream>>>(
            u->g, u->cell_numbering, u->cell_off, staging, u->arr[3 - u->w_cur], ctx->n_cells);
        MMF_LAUNCH_CHECK(ctx);
    }
    if (field == MMF_FIELD_U) u->eig_candidate = false;
    if (field != MMF_FIELD_RHS) return uniform_refresh_ghosts(ctx, S, 0);
    return MMF_OK;
}

static int uniform_gather_state(mmf_ctx *ctx, int field, double *staging)
{
    UniformPath *u = ctx->uni;
    int rc;
    if (field == MMF_FIELD_RHS && (rc = uniform_ensure_rhs(ctx))) return rc;
    const double *S = uniform_field_ptr(ctx, field);
    uniform_gather_kernel<<<grid_for(ctx->n_cells, 256), 256, 0, ctx->stream>>>(
        u->g, u->cell_numbering, u->cell_off, S, staging, ctx->n_cells);
    MMF_LAUNCH_CHECK(ctx);
    return MMF_OK;
}

// ---- operators ----------------------------------------------------------------------------------

// euler::computeRHS on the uniform path: the stage kernel with STAGE 0 (RHS materialised)
static int uniform_rhs(mmf_ctx *ctx, int field, double *d_max)
{
    int rc = uniform_ensure_rhs(ctx);
    if (rc) return rc;
    const double *S = uniform_field_ptr(ctx, field);
    return launch_stage<0>(ctx, S, S, ctx->uni->arr[3], d_max);
}

static int uniform_rk(mmf_ctx *ctx, int stage)
{
    UniformPath *u = ctx->uni;
    int rc = uniform_ensure_rhs(ctx);
    if (rc) return rc;
    const UniformGeom &g = u->g;
    dim3 grid((g.nx + 255) / 256, g.ny, g.nz);
    double *U = u->arr[0], *W = u->arr[u->w_cur], *R = u->arr[3];
    switch (stage) {
    case 1: uniform_rk_kernel<1><<<grid, 256, 0, ctx->stream>>>(g, ctx->d_ctl, U, W, R, u->solid); break;
    case 2: uniform_rk_kernel<2><<<grid, 256, 0, ctx->stream>>>(g, ctx->d_ctl, U, W, R, u->solid); break;
    default: uniform_rk_kernel<3><<<grid, 256, 0, ctx->stream>>>(g, ctx->d_ctl, U, W, R, u->solid); break;
    }
    MMF_LAUNCH_CHECK(ctx);
    if (stage == 3) u->eig_candidate = false;
    return uniform_refresh_ghosts(ctx, stage == 3 ? U : W, 0);
}

int comm_allreduce_max_enqueue(mmf_ctx *ctx, double *d_value, int count); // comm.cuh

// One fused SSP-RK3 step (replaces src/main.cpp:383-506):
//   max eigenvalue of U -> dt on the device -> three fused residual+update kernels, each followed by
//   the ghost refresh of its output; behind stage 3 the max eigenvalue of the new U is found from the
//   stage's per-tile estimates, so the next step starts without a pass over U.
static int uniform_step_enqueue(mmf_ctx *ctx)
{
    UniformPath *u = ctx->uni;
    const UniformGeom &g = u->g;
    StepControl *c = ctx->d_ctl;
    int rc;
    double *U = u->arr[0], *Wa = u->arr[1], *Wb = u->arr[2];

    trace_point(ctx, "gap");
    if (u->eig_candidate) {
        // steady state: eig_next was found (and reduced over the ranks) at the end of the previous step
        begin_step_choose_dt_kernel<<<1, 1, 0, ctx->stream>>>(c);
        MMF_LAUNCH_CHECK(ctx);
    } else {
        begin_step_kernel<<<1, 1, 0, ctx->stream>>>(c, 0);
        MMF_LAUNCH_CHECK(ctx);
        if (u->bodies) { // a box with bodies: fluid cells, border ghosts behind them, wall images
            dim3 grid((g.nx + 255) / 256, g.ny, g.nz);
            uniform_eig_body_kernel<<<grid, 256, 0, ctx->stream>>>(g, U, u->solid, &c->max_eig[0]);
            MMF_LAUNCH_CHECK(ctx);
        } else { // first step after the host (re)wrote U: full pass (+ all-reduce)
            dim3 grid((g.nx + 2 + 255) / 256, g.ny + 2, (g.nz + 2 + EIG_ZCHUNK - 1) / EIG_ZCHUNK);
            uniform_eig_kernel<<<grid, 256, 0, ctx->stream>>>(g, U, &c->max_eig[0]);
            MMF_LAUNCH_CHECK(ctx);
            if (ctx->comm && (rc = comm_allreduce_max_enqueue(ctx, &c->max_eig[0], 1))) return rc;
        }
        choose_dt_kernel<<<1, 1, 0, ctx->stream>>>(c);
        MMF_LAUNCH_CHECK(ctx);
    }
    u->eig_candidate = false;
    trace_point(ctx, "begin+dt");

    // the stage-1 kernel re-derives the same face maximum as a by-product; advance_time_kernel
    // compares the two and mmf_step / mmf_run fail loudly if they ever disagree
    if ((rc = launch_stage<1>(ctx, U, U, Wa, &c->max_eig_chk))) return rc;
    trace_point(ctx, "stage1");
    if ((rc = uniform_refresh_ghosts(ctx, Wa, 1))) return rc;
    trace_point(ctx, "halo1");
    if ((rc = launch_stage<2>(ctx, Wa, U, Wb, &c->max_eig[1]))) return rc;
    trace_point(ctx, "stage2");
    if ((rc = uniform_refresh_ghosts(ctx, Wb, 1))) return rc;
    trace_point(ctx, "halo2");
    if ((rc = launch_stage<3>(ctx, Wb, U, U, &c->max_eig[2]))) return rc;
    trace_point(ctx, "stage3");
    u->w_cur = 2;
    const StageShape &s3 = u->shape[3];
    {
        // ghost cells of a non-copy boundary condition add their own eigenvalue, then the listed tiles
        if ((rc = uniform_refresh_ghosts(ctx, U, 1, &c->eig_next))) return rc;
        const int tx = (g.nx + XW - 1) / XW, ty = (g.ny + s3.rows() - 1) / s3.rows();
        if (ctx->comm) {
            // the largest estimate is reduced over the ranks first: with its own maximum as the yardstick a rank that
            // holds nothing but free stream -- a plateau of equal estimates -- would list every tile and pay a full pass
            // over U (measured at 8 GPUs: 0.27 ms per step on those ranks against 0.03 ms for the all-reduce)
            uniform_eig_estmax_kernel<<<1, 1024, 0, ctx->stream>>>(u->cta_est, u->n_tiles3, &c->est_max);
            MMF_LAUNCH_CHECK(ctx);
            if ((rc = comm_allreduce_max_enqueue(ctx, &c->est_max, 1))) return rc;
            uniform_eig_select_kernel<<<1, 1024, 0, ctx->stream>>>(u->cta_est, u->n_tiles3, &c->est_max, u->eig_cand);
            MMF_LAUNCH_CHECK(ctx);
        } else {
            uniform_eig_estmax_select_kernel<<<1, 1024, 0, ctx->stream>>>(u->cta_est, u->n_tiles3, &c->est_max, u->eig_cand);
            MMF_LAUNCH_CHECK(ctx);
        }
        uniform_eig_tiles_kernel<<<4 * ctx->prop.multiProcessorCount, 320, 0, ctx->stream>>>(g, U, u->eig_cand, tx, ty, s3.rows(),
                                                                                             s3.lz, &c->eig_next, u->solid);
        MMF_LAUNCH_CHECK(ctx);
        if (u->bodies && u->n_wall > 0) {
            // the cells that touch a wall are neither stored nor estimated by the stage kernel, and a wall's mirror image
            // has its own eigenvalue: a pass over the wall-cell list (a surface)
            uniform_eig_wall_kernel<<<(u->n_wall + 127) / 128, 128, 0, ctx->stream>>>(g, U, u->solid, u->wall_list, u->n_wall, c, &c->eig_next);
            MMF_LAUNCH_CHECK(ctx);
        }
        u->eig_candidate = true;
        trace_point(ctx, "halo3+eig");
    }
    // one message per step: the two values main.cpp only logs (:436, :472), the stage-1 check value
    // and the next step's max eigenvalue (max_eig[1], max_eig[2], max_eig_chk, eig_next are contiguous)
    if (ctx->comm && (rc = comm_allreduce_max_enqueue(ctx, &c->max_eig[1], 4))) return rc;
    trace_point(ctx, "allreduce");
    advance_time_kernel<<<1, 1, 0, ctx->stream>>>(c, 1);
    MMF_LAUNCH_CHECK(ctx);
    trace_point(ctx, "advance");
    return MMF_OK;
}

// A step of the steady state (one GPU, the eigenvalue candidate at hand, nothing being timed per launch) is the same
// seven launches with the same arguments every time: they are captured ONCE into a CUDA graph and replayed -- one
// host call per step instead of seven, which is what a step costs on the reference's own 32^3 / 64^2 sized cases.
// MMF_STEP_GRAPH=0 turns it off.
static int uniform_step(mmf_ctx *ctx)
{
    UniformPath *u = ctx->uni;
    static const bool graphs_on = !(getenv("MMF_STEP_GRAPH") && atoi(getenv("MMF_STEP_GRAPH")) == 0);
    const bool steady = graphs_on && u->eig_candidate && !ctx->comm && !ctx->profiling && !ctx->tracing && u->w_cur == 2;
    if (!steady) return uniform_step_enqueue(ctx);
    if (!u->step_graph) {
        const int64_t launches0 = ctx->kernel_launches;
        MMF_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = uniform_step_enqueue(ctx);
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess || !graph) return fail(ctx, MMF_ERR_CUDA, "capture of a step failed: %s", cudaGetErrorString(e));
        const cudaError_t ei = cudaGraphInstantiate(&u->step_graph, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) return fail(ctx, MMF_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ei));
        u->step_graph_launches = (int) (ctx->kernel_launches - launches0);
        ctx->kernel_launches = launches0; // captured, not run
    }
    MMF_CUDA(ctx, cudaGraphLaunch(u->step_graph, ctx->stream));
    ctx->kernel_launches += u->step_graph_launches;
    u->eig_candidate = true;
    u->w_cur = 2;
    return MMF_OK;
}

} // namespace mmf
