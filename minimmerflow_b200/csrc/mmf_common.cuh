// mmf_common.cuh -- handle layout and error plumbing shared by the translation units of
// libmmf_b200.so (host side, C++17).
#pragma once

#include "../../include/mmf_b200.h"
#include "generic_types.cuh"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace mmf {

struct UniformPath; // uniform_path.cuh
struct Comm;        // comm.cuh

} // namespace mmf

// how long a halo wait spins for a neighbour rank's layer before it gives up (halo_spin, uniform_device.cuh)
static inline unsigned long long halo_timeout_ns()
{
    static const unsigned long long ns = [] {
        const char *e = getenv("MMF_HALO_TIMEOUT_MS");
        const double ms = e ? atof(e) : 30000.0;
        return (unsigned long long) (ms > 0. ? ms * 1e6 : 0.);
    }();
    return ns;
}

struct mmf_ctx {
    int device = -1;
    cudaDeviceProp prop{};
    cudaStream_t stream = nullptr;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    std::string err;

    int path = MMF_PATH_GENERIC;
    int dim = 3;
    int problem_type = 0;
    int64_t n_cells = 0, n_ifaces = 0;
    int64_t kernel_launches = 0;
    int64_t device_bytes = 0;
    bool state_valid[3] = { false, false, false };

    // generic path
    mmf::GenericMesh gm{};
    double *fields[3] = { nullptr, nullptr, nullptr }; // SoA, 5*stride each
    // stages 2 and 3 of a step as one kernel each (default; MMF_GENERIC_FUSED=0 turns it off)
    // (generic_stage_kernel); stage 2 then writes the second work array and the two swap roles
    bool generic_fused = false;
    double *w_alt = nullptr;
    cudaGraphExec_t gen_graph[2] = { nullptr, nullptr }; // the launches of a step, captured per role of the two work arrays
    int gen_graph_launches = 0;
    std::vector<void *> owned;                         // every cudaMalloc'd pointer of the handle

    // shared
    double *staging = nullptr;          // device AoS staging buffer, 5*n_cells
    mmf::StepControl *d_ctl = nullptr;  // device control block
    mmf::StepControl *h_ctl = nullptr;  // pinned host mirror
    void *flush_buf = nullptr;
    size_t flush_bytes = 0;

    mmf::UniformPath *uni = nullptr;
    mmf::Comm *comm = nullptr;

    // MMF_TRACE=1: events between the phases of a uniform step, summarised on stderr at destroy
    bool tracing = false;
    struct TracePoint { const char *label; cudaEvent_t ev; };
    std::vector<TracePoint> trace;

    // optional per-launch timing of the residual kernels (CUDA events on the launching stream)
    bool profiling = false;
    struct TimedLaunch { int kind; cudaEvent_t e0, e1; };
    std::vector<TimedLaunch> timed;
};

namespace mmf {

extern thread_local std::string g_last_error;

inline int fail(mmf_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (ctx) ctx->err = buf;
    return code;
}

#define MMF_CUDA(ctx, call)                                                                        \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            return mmf::fail((ctx), MMF_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__,       \
                             __LINE__, cudaGetErrorString(e__));                                   \
        }                                                                                          \
    } while (0)

#define MMF_LAUNCH_CHECK(ctx)                                                                      \
    do {                                                                                           \
        (ctx)->kernel_launches++;                                                                  \
        MMF_CUDA((ctx), cudaGetLastError());                                                       \
    } while (0)

template <typename T>
int dev_alloc(mmf_ctx *ctx, T **ptr, size_t count)
{
    void *p = nullptr;
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = sizeof(T);
    MMF_CUDA(ctx, cudaMalloc(&p, bytes));
    ctx->owned.push_back(p);
    ctx->device_bytes += (int64_t) bytes;
    *ptr = static_cast<T *>(p);
    return MMF_OK;
}

template <typename T>
int dev_upload(mmf_ctx *ctx, T **ptr, const std::vector<T> &host)
{
    int rc = dev_alloc(ctx, ptr, host.size());
    if (rc != MMF_OK) return rc;
    if (!host.empty()) {
        MMF_CUDA(ctx, cudaMemcpy(*ptr, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    }
    return MMF_OK;
}

// brackets one launch with events when profiling is on: kind 0 = RHS only, 1..3 = fused stage
struct ScopedLaunchTimer {
    mmf_ctx *ctx;
    cudaEvent_t e1 = nullptr;
    ScopedLaunchTimer(mmf_ctx *c, int kind) : ctx(c)
    {
        if (!ctx->profiling) return;
        cudaEvent_t e0;
        if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { e1 = nullptr; return; }
        cudaEventRecord(e0, ctx->stream);
        ctx->timed.push_back({ kind, e0, e1 });
    }
    ~ScopedLaunchTimer() { if (e1) cudaEventRecord(e1, ctx->stream); }
};

inline void trace_point(mmf_ctx *ctx, const char *label)
{
    if (!ctx->tracing || ctx->trace.size() > 20000) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, ctx->stream);
    ctx->trace.push_back({ label, e });
}

inline void trace_report(mmf_ctx *ctx)
{
    if (ctx->trace.size() < 2) return;
    cudaStreamSynchronize(ctx->stream);
    struct Acc { const char *label; double ms; int n; };
    std::vector<Acc> acc;
    // each interval is attributed to the label at its END
    const size_t first = 0;
    for (size_t i = std::max<size_t>(first, 1); i < ctx->trace.size(); ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->trace[i - 1].ev, ctx->trace[i].ev) != cudaSuccess) continue;
        bool found = false;
        for (auto &a : acc) if (a.label == ctx->trace[i].label) { a.ms += ms; a.n++; found = true; }
        if (!found) acc.push_back({ ctx->trace[i].label, ms, 1 });
    }
    fprintf(stderr, "[mmf trace, device %d] average time up to each point (ms):", ctx->device);
    for (auto &a : acc) fprintf(stderr, "  %s=%.4f(x%d)", a.label, a.ms / a.n, a.n);
    fprintf(stderr, "\n");
    for (auto &t : ctx->trace) cudaEventDestroy(t.ev);
    ctx->trace.clear();
}

inline unsigned grid_for(int64_t n, int block) { return (unsigned) ((n + block - 1) / block); }

} // namespace mmf
