// stage_tu.cu -- one translation unit per (kernel form, stage): the Makefile compiles this file with
// -DMMF_TU_FORM=<r|t> -DMMF_TU_FORM_ID=<0|2> -DMMF_TU_STAGE=<0..3>, so that the stage-kernel instantiations
// (3 accumulation orders x CTA shapes x padded / compact x ghosts per form and stage) build in parallel.
//   r = the rotate form (uniform_stage_v5r.cuh), the default of every stage
//   t = the same scheme with its input staged in shared memory by bulk tensor (TMA) loads (uniform_stage_t.cuh);
//       also serves form 'h' (one warp for both halo rows) and form 'b' (form 'h' for a box with bodies, the wall
//       cells by a small pass around it: uniform_body_cells.cuh)
#include "uniform_launch.cuh"

#ifndef MMF_TU_FORM_ID // a bare `nvcc -c stage_tu.cu` (no Makefile): the rotate form, RHS only
#define MMF_TU_FORM r
#define MMF_TU_FORM_ID 0
#define MMF_TU_STAGE 0
#endif

#if MMF_TU_FORM_ID == 0
#include "uniform_stage_v5r.cuh"
#elif MMF_TU_FORM_ID == 2
#include "uniform_stage_v5r.cuh"
#include "uniform_body_cells.cuh"
#include "uniform_stage_t.cuh"
#else
#error "MMF_TU_FORM_ID must be 0 (r) or 2 (t)"
#endif

namespace mmf {

#if MMF_TU_FORM_ID == 2
// a box with bodies (form 'b'): the wall cells' results into the compact buffer BEFORE the stage kernel runs
// (it does not store them, and stage 3 updates U in place), and from the buffer into the output array behind it
template <int STAGE, int ORDER>
static int launch_wall_cells(mmf_ctx *ctx, const double *Sin, const double *Un, double *d_max)
{
    UniformPath *u = ctx->uni;
    if (!u->solid || !u->wall_list || !u->wall_compact) return fail(ctx, MMF_ERR_INVALID, "the stage kernels for a box with bodies need the flag array and the wall-cell list");
    if (u->n_wall > 0) {
        uniform_wall_cells_kernel<STAGE, ORDER><<<(u->n_wall + 127) / 128, 128, 0, ctx->stream>>>(
            u->g, uniform_load_clamp(u), Sin, Un, u->solid, u->wall_list, u->n_wall, ctx->d_ctl, u->wall_compact, d_max);
        MMF_LAUNCH_CHECK(ctx);
    }
    return MMF_OK;
}

template <int STAGE>
static int launch_wall_scatter(mmf_ctx *ctx, double *Out)
{
    UniformPath *u = ctx->uni;
    if (u->n_wall > 0) {
        uniform_wall_scatter_kernel<<<(u->n_wall + 127) / 128, 128, 0, ctx->stream>>>(u->g.fs, u->wall_list, u->n_wall,
                                                                                      u->wall_compact, Out, ctx->d_ctl, STAGE >= 1);
        MMF_LAUNCH_CHECK(ctx);
    }
    return MMF_OK;
}
#endif

template <int STAGE, int ORDER>
static int launch_stage_o(mmf_ctx *ctx, const double *Sin, const double *Un, double *Out, double *d_max)
{
    UniformPath *u = ctx->uni;
    const StageShape sh = u->shape[STAGE];
#if MMF_TU_FORM_ID == 2
    // (compact x ghost columns -- an x partition side -- are only read by the rotate form: comm_ipc_import switches the
    //  stages' shapes to it)
    if (uniform_use_xghost(ctx)) return fail(ctx, MMF_ERR_STATE, "stage-kernel forms 't' / 'h' cannot read compact x ghost columns");
    if (sh.form == 'b') {
        if (int rc = launch_wall_cells<STAGE, ORDER>(ctx, Sin, Un, d_max)) return rc;
        if (int rc = launch_stage_tl(ctx, uniform_stage_kernel_t<STAGE, ORDER, 12, T_DEPTH, true, true>, STAGE, 12, true,
                                     stage_t_smem_bytes(12, STAGE, T_DEPTH, true), Sin, Un, Out, d_max, u->solid)) return rc;
        return launch_wall_scatter<STAGE>(ctx, Out);
    }
    if (u->bodies) return fail(ctx, MMF_ERR_STATE, "a box with bodies runs kernel form 'b'");
#define MMF_LAUNCH_T(NWV, DV, MHV) return launch_stage_tl(ctx, uniform_stage_kernel_t<STAGE, ORDER, NWV, DV, MHV>, STAGE, NWV, MHV, stage_t_smem_bytes(NWV, STAGE, DV, MHV), Sin, Un, Out, d_max)
    if (sh.form == 'h') { if (sh.nw == 12) MMF_LAUNCH_T(12, T_DEPTH, true); MMF_LAUNCH_T(16, T_DEPTH, true); }
    if (sh.nw == 16) MMF_LAUNCH_T(16, T_DEPTH, false);
    if (sh.nw == 8) MMF_LAUNCH_T(8, T_DEPTH, false);
    MMF_LAUNCH_T(12, T_DEPTH, false);
#undef MMF_LAUNCH_T
#else
    const bool xgk = uniform_use_xghost(ctx);
    // record (11) + flux (5) doubles per lane and row, two mbarriers per row
#define MMF_LAUNCH(NWV, XGV) return launch_stage_k(ctx, uniform_stage_kernel_v5r<STAGE, ORDER, NWV, XGV>, STAGE, NWV, stage_v5_smem_bytes(NWV), Sin, Un, Out, d_max)
    if (sh.nw == 16) { if (xgk) MMF_LAUNCH(16, true); MMF_LAUNCH(16, false); }
    if (sh.nw == 8) MMF_LAUNCH(8, false);
    if (xgk) MMF_LAUNCH(12, true);
    MMF_LAUNCH(12, false);
#undef MMF_LAUNCH
#endif
}

int MMF_STAGE_TU_NAME(MMF_TU_FORM, MMF_TU_STAGE)(mmf_ctx *ctx, int order, const double *Sin, const double *Un, double *Out, double *d_max)
{
    switch (order) {
    case NUM_MORTON: return launch_stage_o<MMF_TU_STAGE, NUM_MORTON>(ctx, Sin, Un, Out, d_max);
    case NUM_LEXI:   return launch_stage_o<MMF_TU_STAGE, NUM_LEXI>(ctx, Sin, Un, Out, d_max);
    default:         return launch_stage_o<MMF_TU_STAGE, NUM_AXIS>(ctx, Sin, Un, Out, d_max);
    }
}

} // namespace mmf
