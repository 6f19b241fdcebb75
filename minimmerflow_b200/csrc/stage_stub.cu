// stage_stub.cu -- experiment builds only (make FORMS="r" ...): the launchers of the kernel forms that were left out,
// so that the library still links and loads; asking for such a form fails loudly.
#include "uniform_launch.cuh"

namespace mmf {
#define MMF_STUB(F, S)                                                                                             \
    int MMF_STAGE_TU_NAME(F, S)(mmf_ctx *ctx, int, const double *, const double *, double *, double *)             \
    {                                                                                                              \
        return fail(ctx, MMF_ERR_INVALID, "this experiment build of the library leaves out a stage-kernel form");  \
    }
MMF_STUB(MMF_STUB_FORM, 0)
MMF_STUB(MMF_STUB_FORM, 1)
MMF_STUB(MMF_STUB_FORM, 2)
MMF_STUB(MMF_STUB_FORM, 3)
} // namespace mmf
