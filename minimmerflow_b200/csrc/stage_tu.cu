// stage_tu.cu -- one translation unit per (kernel form, stage): the Makefile compiles this file with
// -DMMF_TU_FORM=<p|r|d|t|h|w|b|c> -DMMF_TU_FORM_ID=<0..7> -DMMF_TU_STAGE=<0..3>, so that the stage-kernel instantiations
// (3 accumulation orders x CTA shapes x padded / compact x ghosts per form and stage) build in parallel.
#include "uniform_launch.cuh"

#ifndef MMF_TU_FORM_ID // a bare `nvcc -c stage_tu.cu` (no Makefile): the ping-pong form, RHS only
#define MMF_TU_FORM p
#define MMF_TU_FORM_ID 0
#define MMF_TU_STAGE 0
#endif

#if MMF_TU_FORM_ID == 0
#include "uniform_stage_v5.cuh"
#elif MMF_TU_FORM_ID == 1
#include "uniform_stage_v5r.cuh"
#elif MMF_TU_FORM_ID == 2 || MMF_TU_FORM_ID == 4
#include "uniform_stage_v6.cuh"
#elif MMF_TU_FORM_ID == 3
#include "uniform_stage_v3.cuh"
#elif MMF_TU_FORM_ID == 5
#include "uniform_stage_v7.cuh"
#elif MMF_TU_FORM_ID == 6 || MMF_TU_FORM_ID == 7
#include "uniform_stage_v5rb.cuh"
#else
#error "MMF_TU_FORM_ID must be 0 (p), 1 (r), 2 (d), 3 (t), 4 (h), 5 (w), 6 (b) or 7 (c)"
#endif

namespace mmf {

template <int STAGE, int ORDER>
static int launch_stage_o(mmf_ctx *ctx, const double *Sin, const double *Un, double *Out, double *d_max)
{
    UniformPath *u = ctx->uni;
    const StageShape sh = u->shape[STAGE];
#if MMF_TU_FORM_ID == 3
    (void) sh;
    return launch_stage_v3(ctx, uniform_stage_kernel_v3<STAGE, ORDER, 12>, STAGE, Sin, Un, Out, d_max);
#elif MMF_TU_FORM_ID == 6
    (void) sh;
    return launch_stage_body(ctx, uniform_stage_kernel_v5rb<STAGE, ORDER, 12, false>, STAGE, Sin, Un, Out, d_max);
#elif MMF_TU_FORM_ID == 7
    (void) sh;
    // wall cells into the compact buffer, the stage kernel (which does not store them), the buffer into the output
    UniformPath *uw = ctx->uni;
    if (!uw->wall_list || !uw->wall_compact) return fail(ctx, MMF_ERR_INVALID, "stage-kernel form 'c' needs the wall-cell list");
    if (uw->n_wall > 0) {
        uniform_wall_cells_kernel<STAGE, ORDER><<<(uw->n_wall + 127) / 128, 128, 0, ctx->stream>>>(
            uw->g, uniform_load_clamp(uw), Sin, Un, uw->solid, uw->wall_list, uw->n_wall, ctx->d_ctl, uw->wall_compact, d_max);
        MMF_LAUNCH_CHECK(ctx);
    }
    if (int rc = launch_stage_body(ctx, uniform_stage_kernel_v5rb<STAGE, ORDER, 12, true>, STAGE, Sin, Un, Out, d_max)) return rc;
    if (uw->n_wall > 0) {
        uniform_wall_scatter_kernel<<<(uw->n_wall + 127) / 128, 128, 0, ctx->stream>>>(uw->g.fs, uw->wall_list, uw->n_wall,
                                                                                       uw->wall_compact, Out, ctx->d_ctl, STAGE >= 1);
        MMF_LAUNCH_CHECK(ctx);
    }
    return MMF_OK;
#else
    const bool xgk = uniform_use_xghost(ctx);
#if MMF_TU_FORM_ID == 5
    (void) sh;
    // two rows per warp: ports = warps + 1, the shared-memory layout of the merged-halo v6 kernel
#define MMF_LAUNCH(NWV, XGV) return launch_stage_k(ctx, uniform_stage_kernel_v7<STAGE, ORDER, NWV, XGV>, STAGE, NWV, stage_v6_smem_bytes(NWV, true), Sin, Un, Out, d_max)
    // 8 warps only: at 12 warps (168 registers) the two-cell body spills about 1 KB per thread
    if (xgk) MMF_LAUNCH(8, true);
    MMF_LAUNCH(8, false);
#undef MMF_LAUNCH
#else
    // v5 forms: record (11) + flux (5) doubles per lane and row, two mbarriers per row; v6: twice that
#if MMF_TU_FORM_ID == 2 || MMF_TU_FORM_ID == 4
#define MMF_MH (MMF_TU_FORM_ID == 4)
#define MMF_LAUNCH(NWV, XGV) return launch_stage_k(ctx, uniform_stage_kernel_v6<STAGE, ORDER, NWV, XGV, MMF_MH>, STAGE, NWV, stage_v6_smem_bytes(NWV, MMF_MH), Sin, Un, Out, d_max)
#else
#if MMF_TU_FORM_ID == 1
#define MMF_KERN uniform_stage_kernel_v5r
#else
#define MMF_KERN uniform_stage_kernel_v5
#endif
#define MMF_SMEM_V5(NWV) ((size_t) (NWV) * 16 * 32 * sizeof(double) + 2 * (NWV) * sizeof(unsigned long long))
#define MMF_LAUNCH(NWV, XGV) return launch_stage_k(ctx, MMF_KERN<STAGE, ORDER, NWV, XGV>, STAGE, NWV, MMF_SMEM_V5(NWV), Sin, Un, Out, d_max)
#endif
    if (sh.nw == 16) { if (xgk) MMF_LAUNCH(16, true); MMF_LAUNCH(16, false); }
    if (sh.nw == 8) MMF_LAUNCH(8, false);
    if (xgk) MMF_LAUNCH(12, true);
    MMF_LAUNCH(12, false);
#undef MMF_LAUNCH
#endif
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
