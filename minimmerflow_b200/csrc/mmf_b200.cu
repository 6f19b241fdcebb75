// mmf_b200.cu -- C-ABI of libmmf_b200.so (see include/mmf_b200.h) and the host side of the
// generic (connectivity-driven) path.  Host code is C++17; all numerical work happens in the
// sm_100a kernels of generic_kernels.cuh / uniform_path.cuh.  There is no CPU fallback.
#include "mmf_common.cuh"
#include "generic_tables.h"
#include "uniform_path.cuh"
#include "comm.cuh"

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>

namespace mmf {
thread_local std::string g_last_error;
}

using namespace mmf;

// ------------------------------------------------------------------------------------------------
// device discovery
// ------------------------------------------------------------------------------------------------

static int usable_device_count()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int usable = 0;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp p{};
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) usable++;
    }
    return usable;
}

extern "C" int mmf_device_count(void) { return usable_device_count(); }

extern "C" const char *mmf_last_error(const mmf_ctx *ctx)
{
    if (ctx && !ctx->err.empty()) return ctx->err.c_str();
    return g_last_error.c_str();
}

// ------------------------------------------------------------------------------------------------
// creation: generic path
// ------------------------------------------------------------------------------------------------

static int validate_desc(const mmf_mesh_desc *d)
{
    if (!d) return fail(nullptr, MMF_ERR_INVALID, "mmf_create: null mesh description");
    if (d->struct_size != sizeof(mmf_mesh_desc)) {
        return fail(nullptr, MMF_ERR_INVALID, "mmf_create: struct_size %zu != %zu (ABI mismatch)",
                    d->struct_size, sizeof(mmf_mesh_desc));
    }
    if (d->dim != 2 && d->dim != 3) return fail(nullptr, MMF_ERR_INVALID, "mmf_create: dim must be 2 or 3");
    if (d->n_cells <= 0 || d->n_interfaces < 0) return fail(nullptr, MMF_ERR_INVALID, "mmf_create: empty mesh");
    if (d->n_cells >= (int64_t) 1 << 31 || d->n_interfaces >= (int64_t) 1 << 30) {
        return fail(nullptr, MMF_ERR_INVALID, "mmf_create: mesh too large for 32-bit device indices "
                                              "(cells < 2^31, interfaces < 2^30 per GPU)");
    }
    if (!d->owner || !d->neigh || !d->bc || !d->area || !d->normal || !d->volume || !d->solved) {
        return fail(nullptr, MMF_ERR_INVALID, "mmf_create: a required array is NULL");
    }
    return MMF_OK;
}

static int create_generic(mmf_ctx *ctx, const mmf_mesh_desc *d)
{
    // the tables themselves are host logic without a device: generic_tables.h (unit-tested on the CPU)
    GenericTables t;
    {
        std::string err;
        const int trc = build_generic_tables(d, t, err);
        if (trc) return fail(ctx, trc, "%s", err.c_str());
    }
    const int64_t nc = t.n_cells, nf = t.n_ifaces;
    const std::vector<int32_t> &owner = t.owner, &neigh = t.neigh, &ent = t.ent;
    const std::vector<int8_t> &bc = t.bc;
    const std::vector<double> &normal = t.normal, &area = t.area, &volume = t.volume;
    const std::vector<uint8_t> &solved = t.solved, &update = t.update;
    const std::vector<int64_t> &ptr = t.ptr;

    GenericMesh &g = ctx->gm;
    g.n_cells  = nc;
    g.n_ifaces = nf;
    g.stride   = t.stride;
    memcpy(g.dirichlet_info, d->dirichlet_info, sizeof g.dirichlet_info);

    int rc;
    int64_t *d_ptr; int32_t *d_ent, *d_owner, *d_neigh; int8_t *d_bc; double *d_area, *d_normal, *d_vol;
    uint8_t *d_solved, *d_update;
    if ((rc = dev_upload(ctx, &d_ptr, ptr))) return rc;
    if ((rc = dev_upload(ctx, &d_ent, ent))) return rc;
    if ((rc = dev_upload(ctx, &d_owner, owner))) return rc;
    if ((rc = dev_upload(ctx, &d_neigh, neigh))) return rc;
    if ((rc = dev_upload(ctx, &d_bc, bc))) return rc;
    if ((rc = dev_upload(ctx, &d_area, area))) return rc;
    if ((rc = dev_upload(ctx, &d_normal, normal))) return rc;
    if ((rc = dev_upload(ctx, &d_vol, volume))) return rc;
    if ((rc = dev_upload(ctx, &d_solved, solved))) return rc;
    if ((rc = dev_upload(ctx, &d_update, update))) return rc;
    g.cf_ptr = d_ptr; g.cf_ent = d_ent; g.f_owner = d_owner; g.f_neigh = d_neigh; g.f_bc = d_bc;
    g.f_area = d_area; g.f_normal = d_normal; g.c_volume = d_vol; g.c_solved = d_solved; g.c_update = d_update;

    for (int i = 0; i < 3; ++i) {
        if ((rc = dev_alloc(ctx, &ctx->fields[i], (size_t) NF * g.stride))) return rc;
        MMF_CUDA(ctx, cudaMemset(ctx->fields[i], 0, sizeof(double) * NF * g.stride));
    }
    // (not for a description with a boundary condition between two solved cells, see generic_tables.h)
    // stages 2 and 3 as one kernel each (bit-exact on the GPU, 10 % faster per step at 128^3: profiles/r02c_experiments.md);
    // MMF_GENERIC_FUSED=0 restores the unfused sequence
    ctx->generic_fused = !(getenv("MMF_GENERIC_FUSED") && atoi(getenv("MMF_GENERIC_FUSED")) == 0) && !t.bc_between_solved;
    if (ctx->generic_fused) {
        if ((rc = dev_alloc(ctx, &ctx->w_alt, (size_t) NF * g.stride))) return rc;
        MMF_CUDA(ctx, cudaMemset(ctx->w_alt, 0, sizeof(double) * NF * g.stride));
    }
    ctx->path = MMF_PATH_GENERIC;
    return MMF_OK;
}

static int create_common(mmf_ctx *ctx, int device)
{
    if (usable_device_count() == 0) {
        return fail(ctx, MMF_ERR_NO_DEVICE, "no usable sm_100 (B200) device: libmmf_b200 has no CPU fallback");
    }
    MMF_CUDA(ctx, cudaSetDevice(device));
    MMF_CUDA(ctx, cudaGetDeviceProperties(&ctx->prop, device));
    if (ctx->prop.major != 10) {
        return fail(ctx, MMF_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                    device, ctx->prop.major, ctx->prop.minor);
    }
    ctx->device = device;
    ctx->tracing = getenv("MMF_TRACE") && atoi(getenv("MMF_TRACE")) != 0;
    MMF_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0; // the communication stream gets the highest priority: its (small) kernels are
        MMF_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi)); // dispatched ahead of pending stage CTAs
        MMF_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi));
    }
    MMF_CUDA(ctx, cudaEventCreate(&ctx->ev_start));
    MMF_CUDA(ctx, cudaEventCreate(&ctx->ev_stop));
    int rc = dev_alloc(ctx, &ctx->d_ctl, 1);
    if (rc) return rc;
    MMF_CUDA(ctx, cudaMemset(ctx->d_ctl, 0, sizeof(StepControl)));
    MMF_CUDA(ctx, cudaHostAlloc((void **) &ctx->h_ctl, sizeof(StepControl), cudaHostAllocDefault));
    memset(ctx->h_ctl, 0, sizeof(StepControl));
    return MMF_OK;
}

extern "C" int mmf_create(const mmf_mesh_desc *desc, int device, mmf_ctx **out)
{
    if (!out) return fail(nullptr, MMF_ERR_INVALID, "mmf_create: out is NULL");
    *out = nullptr;
    int rc = validate_desc(desc);
    if (rc) return rc;
    mmf_ctx *ctx = new (std::nothrow) mmf_ctx();
    if (!ctx) return fail(nullptr, MMF_ERR_INVALID, "mmf_create: out of host memory");
    ctx->dim = desc->dim;
    ctx->problem_type = desc->problem_type;
    ctx->n_cells = desc->n_cells;
    ctx->n_ifaces = desc->n_interfaces;
    rc = create_common(ctx, device);
    if (rc == MMF_OK) {
        bool use_uniform = false;
        rc = uniform_try_create(ctx, desc, &use_uniform);
        if (rc == MMF_OK && !use_uniform) rc = create_generic(ctx, desc);
    }
    if (rc != MMF_OK) {
        g_last_error = ctx->err;
        mmf_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return MMF_OK;
}

extern "C" int mmf_create_uniform(const mmf_uniform_desc *desc, int device, mmf_ctx **out)
{
    if (!out) return fail(nullptr, MMF_ERR_INVALID, "mmf_create_uniform: out is NULL");
    *out = nullptr;
    // (the description grew by `area` and `volume`: a caller built against the shorter struct is still served)
    if (!desc || (desc->struct_size != sizeof(mmf_uniform_desc) && desc->struct_size != offsetof(mmf_uniform_desc, area))) {
        return fail(nullptr, MMF_ERR_INVALID, "mmf_create_uniform: bad description / ABI mismatch");
    }
    mmf_ctx *ctx = new (std::nothrow) mmf_ctx();
    if (!ctx) return fail(nullptr, MMF_ERR_INVALID, "mmf_create_uniform: out of host memory");
    ctx->dim = 3;
    ctx->problem_type = desc->problem_type;
    int rc = create_common(ctx, device);
    if (rc == MMF_OK) rc = uniform_create(ctx, desc);
    if (rc != MMF_OK) {
        g_last_error = ctx->err;
        mmf_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return MMF_OK;
}

extern "C" int mmf_destroy(mmf_ctx *ctx)
{
    if (!ctx) return MMF_OK;
    if (ctx->device >= 0) cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
    trace_report(ctx);
    for (cudaGraphExec_t &ge : ctx->gen_graph) if (ge) { cudaGraphExecDestroy(ge); ge = nullptr; }
    comm_destroy(ctx);
    uniform_destroy(ctx);
    for (void *p : ctx->owned) cudaFree(p);
    if (ctx->h_ctl) cudaFreeHost(ctx->h_ctl);
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    delete ctx;
    return MMF_OK;
}

extern "C" int mmf_get_info(const mmf_ctx *ctx, mmf_info *info)
{
    if (!ctx || !info) return fail(nullptr, MMF_ERR_INVALID, "mmf_get_info: null argument");
    info->path = ctx->path;
    info->device = ctx->device;
    info->sm_count = ctx->prop.multiProcessorCount;
    info->cc_major = ctx->prop.major;
    info->cc_minor = ctx->prop.minor;
    info->order_exact = (ctx->path == MMF_PATH_GENERIC) ? 1 : uniform_order_exact(ctx);
    info->n_cells = ctx->n_cells;
    info->n_interfaces = ctx->n_ifaces;
    info->kernel_launches = ctx->kernel_launches;
    info->device_bytes = ctx->device_bytes;
    return MMF_OK;
}

// ------------------------------------------------------------------------------------------------
// state transfer
// ------------------------------------------------------------------------------------------------

static int check_field(mmf_ctx *ctx, int field, const char *who)
{
    if (!ctx) return fail(nullptr, MMF_ERR_INVALID, "%s: null handle", who);
    if (field < 0 || field > 2) return fail(ctx, MMF_ERR_INVALID, "%s: unknown field %d", who, field);
    MMF_CUDA(ctx, cudaSetDevice(ctx->device));
    return MMF_OK;
}

static int ensure_staging(mmf_ctx *ctx)
{
    if (ctx->staging) return MMF_OK;
    return dev_alloc(ctx, &ctx->staging, (size_t) NF * ctx->n_cells);
}

static int set_state_enqueue(mmf_ctx *ctx, int field, const double *host_aos)
{
    int rc = ensure_staging(ctx);
    if (rc) return rc;
    const size_t bytes = sizeof(double) * NF * (size_t) ctx->n_cells;
    MMF_CUDA(ctx, cudaMemcpyAsync(ctx->staging, host_aos, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->path == MMF_PATH_UNIFORM) {
        rc = uniform_scatter_state(ctx, field, ctx->staging);
        if (rc) return rc;
    } else {
        aos_to_soa_kernel<<<grid_for(ctx->n_cells, 256), 256, 0, ctx->stream>>>(
            ctx->staging, ctx->fields[field], ctx->n_cells, ctx->gm.stride);
        MMF_LAUNCH_CHECK(ctx);
    }
    ctx->state_valid[field] = true;
    return MMF_OK;
}

static int get_state_enqueue(mmf_ctx *ctx, int field, double *host_aos, bool primitives = false)
{
    int rc = ensure_staging(ctx);
    if (rc) return rc;
    if (ctx->path == MMF_PATH_UNIFORM) {
        rc = uniform_gather_state(ctx, field, ctx->staging);
        if (rc) return rc;
    } else {
        soa_to_aos_kernel<<<grid_for(ctx->n_cells, 256), 256, 0, ctx->stream>>>(
            ctx->fields[field], ctx->staging, ctx->n_cells, ctx->gm.stride);
        MMF_LAUNCH_CHECK(ctx);
    }
    if (primitives) { // cons -> prim on the staged AoS stream, before it leaves the device
        aos_cons_to_prim_kernel<<<grid_for(ctx->n_cells, 256), 256, 0, ctx->stream>>>(ctx->staging, ctx->n_cells);
        MMF_LAUNCH_CHECK(ctx);
    }
    const size_t bytes = sizeof(double) * NF * (size_t) ctx->n_cells;
    MMF_CUDA(ctx, cudaMemcpyAsync(host_aos, ctx->staging, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return MMF_OK;
}

extern "C" int mmf_set_state(mmf_ctx *ctx, int field, const double *host_aos)
{
    int rc = check_field(ctx, field, "mmf_set_state");
    if (rc) return rc;
    if (!host_aos) return fail(ctx, MMF_ERR_INVALID, "mmf_set_state: null buffer");
    if ((rc = set_state_enqueue(ctx, field, host_aos))) return rc;
    MMF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMF_OK;
}

extern "C" int mmf_get_state(mmf_ctx *ctx, int field, double *host_aos)
{
    int rc = check_field(ctx, field, "mmf_get_state");
    if (rc) return rc;
    if (!host_aos) return fail(ctx, MMF_ERR_INVALID, "mmf_get_state: null buffer");
    if ((rc = get_state_enqueue(ctx, field, host_aos))) return rc;
    MMF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMF_OK;
}

extern "C" int mmf_get_primitives(mmf_ctx *ctx, int field, double *host_aos)
{
    int rc = check_field(ctx, field, "mmf_get_primitives");
    if (rc) return rc;
    if (!host_aos) return fail(ctx, MMF_ERR_INVALID, "mmf_get_primitives: null buffer");
    if (field == MMF_FIELD_RHS) return fail(ctx, MMF_ERR_INVALID, "mmf_get_primitives: the residual is not a conservative state");
    if ((rc = get_state_enqueue(ctx, field, host_aos, true))) return rc;
    MMF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMF_OK;
}

// ------------------------------------------------------------------------------------------------
// operators
// ------------------------------------------------------------------------------------------------

extern "C" int mmf_compute_polynomials(mmf_ctx *ctx, int field)
{
    // reconstruction::computePolynomials is an empty function at order 1
    // (src/reconstruction.cpp:47-55); nothing to launch.
    return check_field(ctx, field, "mmf_compute_polynomials");
}

// residual of `field` into RHS, face-max eigenvalue into ctl->max_eig[slot]; no synchronisation
// derived: the fused sequence's variant of the generic residual kernel (the default sequence of step_enqueue; MMF_GENERIC_FUSED=0 turns it off)
static int rhs_enqueue(mmf_ctx *ctx, int field, int slot, bool derived = false)
{
    double *d_max = &ctx->d_ctl->max_eig[slot];
    MMF_CUDA(ctx, cudaMemsetAsync(d_max, 0, sizeof(double), ctx->stream));
    if (ctx->path == MMF_PATH_UNIFORM) return uniform_rhs(ctx, field, d_max);
    {
        ScopedLaunchTimer timer(ctx, 0);
        if (derived && ctx->dim == 2) { // (the register budget that is fastest for four entries per cell)
            generic_rhs_derived_kernel<MMF_GEN_MINBLOCKS_2D><<<grid_for(ctx->n_cells, 128), 128, 0, ctx->stream>>>(
                ctx->gm, ctx->fields[field], ctx->fields[MMF_FIELD_RHS], d_max);
        } else if (derived) {
            generic_rhs_derived_kernel<><<<grid_for(ctx->n_cells, 128), 128, 0, ctx->stream>>>(
                ctx->gm, ctx->fields[field], ctx->fields[MMF_FIELD_RHS], d_max);
        } else {
            generic_rhs_kernel<<<grid_for(ctx->n_cells, 128), 128, 0, ctx->stream>>>(
                ctx->gm, ctx->fields[field], ctx->fields[MMF_FIELD_RHS], d_max);
        }
    }
    MMF_LAUNCH_CHECK(ctx);
    return MMF_OK;
}

static int rk_enqueue(mmf_ctx *ctx, int stage)
{
    if (ctx->path == MMF_PATH_UNIFORM) return uniform_rk(ctx, stage);
    const GenericMesh &g = ctx->gm;
    const unsigned grid = grid_for(g.n_cells, 256);
    double *U = ctx->fields[MMF_FIELD_U], *W = ctx->fields[MMF_FIELD_W], *R = ctx->fields[MMF_FIELD_RHS];
    switch (stage) {
    case 1: generic_rk_kernel<1><<<grid, 256, 0, ctx->stream>>>(g.n_cells, g.stride, g.c_update, g.c_volume, ctx->d_ctl, U, W, R); break;
    case 2: generic_rk_kernel<2><<<grid, 256, 0, ctx->stream>>>(g.n_cells, g.stride, g.c_update, g.c_volume, ctx->d_ctl, U, W, R); break;
    default: generic_rk_kernel<3><<<grid, 256, 0, ctx->stream>>>(g.n_cells, g.stride, g.c_update, g.c_volume, ctx->d_ctl, U, W, R); break;
    }
    MMF_LAUNCH_CHECK(ctx);
    return MMF_OK;
}

// stages 2 and 3 of the generic path as one kernel each (default; MMF_GENERIC_FUSED=0: the unfused sequence)
static int generic_stage_enqueue(mmf_ctx *ctx, int stage)
{
    double *d_max = &ctx->d_ctl->max_eig[stage - 1];
    MMF_CUDA(ctx, cudaMemsetAsync(d_max, 0, sizeof(double), ctx->stream));
    double *U = ctx->fields[MMF_FIELD_U], *W = ctx->fields[MMF_FIELD_W], *R = ctx->fields[MMF_FIELD_RHS];
    const unsigned grid = grid_for(ctx->n_cells, 128);
    {
        ScopedLaunchTimer timer(ctx, stage);
        if (ctx->dim == 2) {
            if (stage == 2) generic_stage_kernel<2, MMF_GEN_MINBLOCKS_2D><<<grid, 128, 0, ctx->stream>>>(ctx->gm, W, U, ctx->w_alt, R, ctx->d_ctl, d_max);
            else            generic_stage_kernel<3, MMF_GEN_MINBLOCKS_2D><<<grid, 128, 0, ctx->stream>>>(ctx->gm, W, U, U, R, ctx->d_ctl, d_max);
        } else {
            if (stage == 2) generic_stage_kernel<2><<<grid, 128, 0, ctx->stream>>>(ctx->gm, W, U, ctx->w_alt, R, ctx->d_ctl, d_max);
            else            generic_stage_kernel<3><<<grid, 128, 0, ctx->stream>>>(ctx->gm, W, U, U, R, ctx->d_ctl, d_max);
        }
    }
    MMF_LAUNCH_CHECK(ctx);
    if (stage == 2) std::swap(ctx->fields[MMF_FIELD_W], ctx->w_alt); // field W is what stage 2 wrote
    return MMF_OK;
}

extern "C" int mmf_compute_rhs(mmf_ctx *ctx, int field, int order, double *max_eig)
{
    int rc = check_field(ctx, field, "mmf_compute_rhs");
    if (rc) return rc;
    if (order != 1) {
        return fail(ctx, MMF_ERR_UNSUPPORTED_ORDER,
                    "mmf_compute_rhs: reconstruction order %d is not supported (the reference exits with "
                    "status 2, src/reconstruction.cpp:76)", order);
    }
    if (field == MMF_FIELD_RHS) return fail(ctx, MMF_ERR_INVALID, "mmf_compute_rhs: input field cannot be RHS");
    if (!ctx->state_valid[field]) return fail(ctx, MMF_ERR_STATE, "mmf_compute_rhs: field %d was never set", field);
    if ((rc = rhs_enqueue(ctx, field, 0))) return rc;
    ctx->state_valid[MMF_FIELD_RHS] = true;
    MMF_CUDA(ctx, cudaMemcpyAsync(&ctx->h_ctl->max_eig[0], &ctx->d_ctl->max_eig[0], sizeof(double),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    MMF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (max_eig) *max_eig = ctx->h_ctl->max_eig[0];
    return MMF_OK;
}

extern "C" int mmf_compute_rhs_host(mmf_ctx *ctx, const double *cons_aos, int order, double *rhs_aos, double *max_eig)
{
    int rc = check_field(ctx, MMF_FIELD_U, "mmf_compute_rhs_host");
    if (rc) return rc;
    if (!cons_aos || !rhs_aos) return fail(ctx, MMF_ERR_INVALID, "mmf_compute_rhs_host: null buffer");
    if (order != 1) return mmf_compute_rhs(ctx, MMF_FIELD_U, order, max_eig);
    // Strict mode: the HOST owns every storage (main.cpp's loops stay on the host), the device U
    // slot is just the mirror of whichever host storage is passed in this call.
    if ((rc = set_state_enqueue(ctx, MMF_FIELD_U, cons_aos))) return rc;
    if ((rc = rhs_enqueue(ctx, MMF_FIELD_U, 0))) return rc;
    ctx->state_valid[MMF_FIELD_RHS] = true;
    if ((rc = get_state_enqueue(ctx, MMF_FIELD_RHS, rhs_aos))) return rc;
    MMF_CUDA(ctx, cudaMemcpyAsync(&ctx->h_ctl->max_eig[0], &ctx->d_ctl->max_eig[0], sizeof(double),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    MMF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (max_eig) *max_eig = ctx->h_ctl->max_eig[0];
    return MMF_OK;
}

extern "C" int mmf_rk_stage(mmf_ctx *ctx, int stage, double dt)
{
    int rc = check_field(ctx, MMF_FIELD_U, "mmf_rk_stage");
    if (rc) return rc;
    if (stage < 1 || stage > 3) return fail(ctx, MMF_ERR_INVALID, "mmf_rk_stage: stage must be 1, 2 or 3");
    if (!ctx->state_valid[MMF_FIELD_U] || !ctx->state_valid[MMF_FIELD_RHS] ||
        (stage > 1 && !ctx->state_valid[MMF_FIELD_W])) {
        return fail(ctx, MMF_ERR_STATE, "mmf_rk_stage: U, RHS (and W for stages 2,3) must be set first");
    }
    set_dt_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_ctl, dt);
    MMF_LAUNCH_CHECK(ctx);
    if ((rc = rk_enqueue(ctx, stage))) return rc;
    if (stage < 3) ctx->state_valid[MMF_FIELD_W] = true;
    MMF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMF_OK;
}

static int generic_step_enqueue(mmf_ctx *ctx);

// one RK3 step on the stream; the control block already holds t, t_max, cfl, min_h
static int step_enqueue(mmf_ctx *ctx)
{
    if (ctx->path == MMF_PATH_UNIFORM) return uniform_step(ctx);
    // One GPU, nothing timed per launch: the launches of a step are the same every time, up to the two work arrays
    // of the fused stages swapping roles from one step to the next -- one captured CUDA graph per role assignment,
    // replayed with one host call per step (the reference's own 64^2 / 32^3 sized cases are launch bound).
    static const bool graphs_on = !(getenv("MMF_STEP_GRAPH") && atoi(getenv("MMF_STEP_GRAPH")) == 0);
    if (!graphs_on || ctx->comm || ctx->profiling || ctx->tracing) return generic_step_enqueue(ctx);
    const int role = (ctx->w_alt && ctx->fields[MMF_FIELD_W] > ctx->w_alt) ? 1 : 0;
    double *w0 = ctx->fields[MMF_FIELD_W], *a0 = ctx->w_alt;
    if (!ctx->gen_graph[role]) {
        const int64_t launches0 = ctx->kernel_launches;
        MMF_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = generic_step_enqueue(ctx);
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        // (captured, not run: the role swap the enqueue did is undone and redone by the launch below)
        ctx->fields[MMF_FIELD_W] = w0; ctx->w_alt = a0;
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess || !graph) return fail(ctx, MMF_ERR_CUDA, "capture of a step failed: %s", cudaGetErrorString(e));
        const cudaError_t ei = cudaGraphInstantiate(&ctx->gen_graph[role], graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) return fail(ctx, MMF_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ei));
        ctx->gen_graph_launches = (int) (ctx->kernel_launches - launches0);
        ctx->kernel_launches = launches0;
    }
    MMF_CUDA(ctx, cudaGraphLaunch(ctx->gen_graph[role], ctx->stream));
    ctx->kernel_launches += ctx->gen_graph_launches;
    if (ctx->generic_fused) std::swap(ctx->fields[MMF_FIELD_W], ctx->w_alt); // what stage 2 of the replayed step did
    ctx->state_valid[MMF_FIELD_W] = true;
    return MMF_OK;
}

static int generic_step_enqueue(mmf_ctx *ctx)
{
    int rc;
    // unfused reference-shaped sequence (src/main.cpp:383-506)
    if ((rc = rhs_enqueue(ctx, MMF_FIELD_U, 0, ctx->generic_fused))) return rc;
    if (ctx->comm && (rc = comm_allreduce_max_enqueue(ctx, &ctx->d_ctl->max_eig[0], 1))) return rc;
    choose_dt_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_ctl);
    MMF_LAUNCH_CHECK(ctx);
    if ((rc = rk_enqueue(ctx, 1))) return rc;
    if (ctx->comm && (rc = comm_exchange_enqueue(ctx, MMF_FIELD_W))) return rc;
    if (ctx->generic_fused) { // residual + stage update in one kernel for stages 2 and 3 (generic_stage_kernel)
        if ((rc = generic_stage_enqueue(ctx, 2))) return rc;
        if (ctx->comm && (rc = comm_exchange_enqueue(ctx, MMF_FIELD_W))) return rc;
        if ((rc = generic_stage_enqueue(ctx, 3))) return rc;
    } else {
        if ((rc = rhs_enqueue(ctx, MMF_FIELD_W, 1))) return rc;
        if ((rc = rk_enqueue(ctx, 2))) return rc;
        if (ctx->comm && (rc = comm_exchange_enqueue(ctx, MMF_FIELD_W))) return rc;
        if ((rc = rhs_enqueue(ctx, MMF_FIELD_W, 2))) return rc;
        if ((rc = rk_enqueue(ctx, 3))) return rc;
    }
    if (ctx->comm && (rc = comm_exchange_enqueue(ctx, MMF_FIELD_U))) return rc;
    if (ctx->comm && (rc = comm_allreduce_max_enqueue(ctx, &ctx->d_ctl->max_eig[1], 2))) return rc; // logged only (:436, :472)
    advance_time_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_ctl, 0);
    MMF_LAUNCH_CHECK(ctx);
    return MMF_OK;
}

static int upload_control(mmf_ctx *ctx, double cfl, double min_h, double t, double t_max)
{
    StepControl *h = ctx->h_ctl;
    // the pinned mirror may still be in flight from a previous async copy
    MMF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memset(h, 0, sizeof *h);
    h->t = t; h->t_max = t_max; h->cfl = cfl; h->min_h = min_h; h->steps = 0.0; h->active = 0.0;
    // only the host-owned head of the block: the eigenvalue by-product of the previous step's stage 3
    // (eig_next / eig_seed) stays valid across calls as long as nothing else touched field U
    MMF_CUDA(ctx, cudaMemcpyAsync(ctx->d_ctl, h, STEP_CONTROL_HOST_FIELDS * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    MMF_CUDA(ctx, cudaMemsetAsync(&ctx->d_ctl->mismatches, 0, sizeof(double), ctx->stream));
    // (halo_timeouts stays: once a neighbour rank failed to deliver, the handle's state is invalid for good)
    return MMF_OK;
}

static int halo_timeout_error(mmf_ctx *ctx, const char *who)
{
    return fail(ctx, MMF_ERR_NCCL, "%s: %d halo wait(s) gave up after %.1f s without a neighbour rank's layer (a rank died or "
                "fell behind; MMF_HALO_TIMEOUT_MS, 0 = wait for ever): the state of this rank is invalid", who,
                (int) ctx->h_ctl->halo_timeouts, 1e-9 * (double) halo_timeout_ns());
}

static int download_control(mmf_ctx *ctx)
{
    MMF_CUDA(ctx, cudaMemcpyAsync(ctx->h_ctl, ctx->d_ctl, sizeof(StepControl), cudaMemcpyDeviceToHost, ctx->stream));
    MMF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMF_OK;
}

extern "C" int mmf_step(mmf_ctx *ctx, double cfl, double min_cell_size, double t, double t_max,
                        double *dt_out, double max_eig_out[3])
{
    int rc = check_field(ctx, MMF_FIELD_U, "mmf_step");
    if (rc) return rc;
    if (!ctx->state_valid[MMF_FIELD_U]) return fail(ctx, MMF_ERR_STATE, "mmf_step: field U was never set");
    if ((rc = upload_control(ctx, cfl, min_cell_size, t, t_max))) return rc;
    if ((rc = step_enqueue(ctx))) return rc;
    ctx->state_valid[MMF_FIELD_W] = ctx->state_valid[MMF_FIELD_RHS] = true;
    if ((rc = download_control(ctx))) return rc;
    if (ctx->h_ctl->halo_timeouts != 0.0) return halo_timeout_error(ctx, "mmf_step");
    if (ctx->h_ctl->mismatches != 0.0) {
        return fail(ctx, MMF_ERR_STATE, "mmf_step: internal check failed: the max eigenvalue that chose dt (%.17g) is not "
                    "the face maximum of the stage-1 residual (%.17g)", ctx->h_ctl->max_eig[0], ctx->h_ctl->max_eig_chk);
    }
    if (dt_out) *dt_out = ctx->h_ctl->dt;
    if (max_eig_out) {
        for (int i = 0; i < 3; ++i) max_eig_out[i] = ctx->h_ctl->max_eig[i];
    }
    return MMF_OK;
}

extern "C" int mmf_run(mmf_ctx *ctx, double cfl, double min_cell_size, double *t, double t_max,
                       int max_steps, int *steps_out)
{
    int rc = check_field(ctx, MMF_FIELD_U, "mmf_run");
    if (rc) return rc;
    if (!t) return fail(ctx, MMF_ERR_INVALID, "mmf_run: t is NULL");
    if (!ctx->state_valid[MMF_FIELD_U]) return fail(ctx, MMF_ERR_STATE, "mmf_run: field U was never set");
    if ((rc = upload_control(ctx, cfl, min_cell_size, *t, t_max))) return rc;
    const bool bounded_time = std::isfinite(t_max);
    if (!bounded_time && max_steps < 0) return fail(ctx, MMF_ERR_INVALID, "mmf_run: neither t_max nor max_steps bounds the loop");

    int enqueued = 0;
    // Steps past t_max switch themselves off on the device, so the host only needs to look at
    // the clock every few steps; with an unbounded t_max it never needs to.
    const int batch = bounded_time ? 8 : std::numeric_limits<int>::max();
    for (;;) {
        int n = batch;
        if (max_steps >= 0) n = std::min(n, max_steps - enqueued);
        for (int i = 0; i < n; ++i) {
            if ((rc = step_enqueue(ctx))) return rc;
        }
        enqueued += n;
        if ((rc = download_control(ctx))) return rc;
        if (max_steps >= 0 && enqueued >= max_steps) break;
        if (bounded_time && !(ctx->h_ctl->t < t_max)) break;
    }
    ctx->state_valid[MMF_FIELD_W] = ctx->state_valid[MMF_FIELD_RHS] = true;
    if (ctx->h_ctl->halo_timeouts != 0.0) return halo_timeout_error(ctx, "mmf_run");
    if (ctx->h_ctl->mismatches != 0.0) {
        return fail(ctx, MMF_ERR_STATE, "mmf_run: internal check failed in %d step(s): the max eigenvalue that chose dt "
                    "is not the face maximum of the stage-1 residual", (int) ctx->h_ctl->mismatches);
    }
    *t = ctx->h_ctl->t;
    if (steps_out) *steps_out = (int) ctx->h_ctl->steps;
    return MMF_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU (comm.cuh)
// ------------------------------------------------------------------------------------------------

extern "C" int mmf_comm_unique_id(void *id_out_128) { return comm_unique_id(id_out_128); }

extern "C" int mmf_comm_init(mmf_ctx *ctx, int rank, int n_ranks, const void *nccl_unique_id)
{
    int rc = check_field(ctx, MMF_FIELD_U, "mmf_comm_init");
    if (rc) return rc;
    return comm_init(ctx, rank, n_ranks, nccl_unique_id);
}

extern "C" int mmf_comm_set_ghost_lists(mmf_ctx *ctx, int n_neighbours, const int32_t *neighbour_ranks,
                                        const int64_t *send_offsets, const int64_t *send_ids,
                                        const int64_t *recv_offsets, const int64_t *recv_ids)
{
    int rc = check_field(ctx, MMF_FIELD_U, "mmf_comm_set_ghost_lists");
    if (rc) return rc;
    return comm_set_ghost_lists(ctx, n_neighbours, neighbour_ranks, send_offsets, send_ids, recv_offsets, recv_ids);
}

extern "C" int mmf_comm_set_box_neighbours(mmf_ctx *ctx, const int32_t neighbour_ranks[6])
{
    int rc = check_field(ctx, MMF_FIELD_U, "mmf_comm_set_box_neighbours");
    if (rc) return rc;
    if (!neighbour_ranks) return fail(ctx, MMF_ERR_INVALID, "mmf_comm_set_box_neighbours: null argument");
    return comm_set_box_neighbours(ctx, neighbour_ranks);
}

extern "C" int mmf_comm_ipc_export(mmf_ctx *ctx, void *blob_out)
{
    int rc = check_field(ctx, MMF_FIELD_U, "mmf_comm_ipc_export");
    if (rc) return rc;
    if (!blob_out) return fail(ctx, MMF_ERR_INVALID, "mmf_comm_ipc_export: null buffer");
    return comm_ipc_export(ctx, blob_out);
}

extern "C" int mmf_comm_ipc_import(mmf_ctx *ctx, const void *all_ranks_blobs)
{
    int rc = check_field(ctx, MMF_FIELD_U, "mmf_comm_ipc_import");
    if (rc) return rc;
    if (!all_ranks_blobs) return fail(ctx, MMF_ERR_INVALID, "mmf_comm_ipc_import: null buffer");
    return comm_ipc_import(ctx, all_ranks_blobs);
}

extern "C" int mmf_exchange(mmf_ctx *ctx, int field)
{
    int rc = check_field(ctx, field, "mmf_exchange");
    if (rc) return rc;
    if (!ctx->comm) return MMF_OK; // not partitioned: main.cpp guards with mesh.isPartitioned()
    if ((rc = comm_exchange_enqueue(ctx, field))) return rc;
    MMF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMF_OK;
}

extern "C" int mmf_allreduce_max(mmf_ctx *ctx, double *value)
{
    int rc = check_field(ctx, MMF_FIELD_U, "mmf_allreduce_max");
    if (rc) return rc;
    if (!value) return fail(ctx, MMF_ERR_INVALID, "mmf_allreduce_max: null value");
    if (!ctx->comm) return MMF_OK;
    MMF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->h_ctl->max_eig[0] = *value;
    MMF_CUDA(ctx, cudaMemcpyAsync(&ctx->d_ctl->max_eig[0], &ctx->h_ctl->max_eig[0], sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = comm_allreduce_max_enqueue(ctx, &ctx->d_ctl->max_eig[0], 1))) return rc;
    MMF_CUDA(ctx, cudaMemcpyAsync(&ctx->h_ctl->max_eig[0], &ctx->d_ctl->max_eig[0], sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MMF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *value = ctx->h_ctl->max_eig[0];
    return MMF_OK;
}

// ------------------------------------------------------------------------------------------------
// measurement helpers
// ------------------------------------------------------------------------------------------------

extern "C" int mmf_timer_start(mmf_ctx *ctx)
{
    if (!ctx) return fail(nullptr, MMF_ERR_INVALID, "mmf_timer_start: null handle");
    MMF_CUDA(ctx, cudaSetDevice(ctx->device));
    for (auto &t : ctx->trace) cudaEventDestroy(t.ev); // MMF_TRACE covers the timed region only
    ctx->trace.clear();
    MMF_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    return MMF_OK;
}

extern "C" int mmf_timer_stop(mmf_ctx *ctx, float *milliseconds)
{
    if (!ctx || !milliseconds) return fail(ctx, MMF_ERR_INVALID, "mmf_timer_stop: null argument");
    MMF_CUDA(ctx, cudaSetDevice(ctx->device));
    MMF_CUDA(ctx, cudaEventRecord(ctx->ev_stop, ctx->stream));
    MMF_CUDA(ctx, cudaEventSynchronize(ctx->ev_stop));
    MMF_CUDA(ctx, cudaEventElapsedTime(milliseconds, ctx->ev_start, ctx->ev_stop));
    trace_report(ctx);
    ctx->tracing = false; // one report per process
    return MMF_OK;
}

extern "C" int mmf_profile_begin(mmf_ctx *ctx)
{
    if (!ctx) return fail(nullptr, MMF_ERR_INVALID, "mmf_profile_begin: null handle");
    MMF_CUDA(ctx, cudaSetDevice(ctx->device));
    MMF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &t : ctx->timed) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
    ctx->timed.clear();
    ctx->profiling = true;
    return MMF_OK;
}

extern "C" int mmf_profile_end(mmf_ctx *ctx, double total_ms[4], int64_t launches[4])
{
    if (!ctx || !total_ms || !launches) return fail(ctx, MMF_ERR_INVALID, "mmf_profile_end: null argument");
    MMF_CUDA(ctx, cudaSetDevice(ctx->device));
    MMF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->profiling = false;
    for (int k = 0; k < 4; ++k) { total_ms[k] = 0.0; launches[k] = 0; }
    for (auto &t : ctx->timed) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess && t.kind >= 0 && t.kind < 4) {
            total_ms[t.kind] += ms;
            launches[t.kind]++;
        }
        cudaEventDestroy(t.e0);
        cudaEventDestroy(t.e1);
    }
    ctx->timed.clear();
    return MMF_OK;
}

extern "C" int mmf_synchronize(mmf_ctx *ctx)
{
    if (!ctx) return fail(nullptr, MMF_ERR_INVALID, "mmf_synchronize: null handle");
    MMF_CUDA(ctx, cudaSetDevice(ctx->device));
    MMF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MMF_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
    return MMF_OK;
}

extern "C" int mmf_flush_l2(mmf_ctx *ctx)
{
    if (!ctx) return fail(nullptr, MMF_ERR_INVALID, "mmf_flush_l2: null handle");
    MMF_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->flush_buf) {
        ctx->flush_bytes = std::max<size_t>((size_t) ctx->prop.l2CacheSize * 2, (size_t) 256 << 20);
        int rc = dev_alloc(ctx, (char **) &ctx->flush_buf, ctx->flush_bytes);
        if (rc) return rc;
    }
    MMF_CUDA(ctx, cudaMemsetAsync(ctx->flush_buf, 0, ctx->flush_bytes, ctx->stream));
    return MMF_OK;
}

extern "C" int mmf_selftest_division(int device, long long n_samples, unsigned long long seed, unsigned long long *mismatches)
{
    if (!mismatches) return fail(nullptr, MMF_ERR_INVALID, "mmf_selftest_division: null argument");
    if (usable_device_count() == 0) return fail(nullptr, MMF_ERR_NO_DEVICE, "no usable sm_100 device");
    MMF_CUDA(nullptr, cudaSetDevice(device));
    unsigned long long *d = nullptr;
    MMF_CUDA(nullptr, cudaMalloc(&d, sizeof *d));
    MMF_CUDA(nullptr, cudaMemset(d, 0, sizeof *d));
    const int blocks = 148 * 8, threads = 256;
    const long long per_thread = std::max<long long>(1, n_samples / ((long long) blocks * threads));
    division_selftest_kernel<<<blocks, threads>>>(seed, per_thread, d);
    division_directed_kernel<<<64, 64>>>(d); // structured mantissas x exponent pairs, +-a, products, a = +0
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(mismatches, d, sizeof *d, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(nullptr, MMF_ERR_CUDA, "division self-test failed: %s", cudaGetErrorString(e));
    return MMF_OK;
}

extern "C" int mmf_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return fail(nullptr, MMF_ERR_INVALID, "mmf_host_alloc: null pointer");
    cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return fail(nullptr, MMF_ERR_CUDA, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return MMF_OK;
}

extern "C" int mmf_host_free(void *ptr)
{
    if (ptr) cudaFreeHost(ptr);
    return MMF_OK;
}
