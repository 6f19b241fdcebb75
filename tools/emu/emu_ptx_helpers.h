// tools/emu/emu_ptx_helpers.h -- CPU stand-ins for minimmerflow_b200/csrc/ptx_helpers.cuh
// (DEVELOPMENT TOOL, see shim/cuda_runtime.h).
//
//  * mbarrier: one 64-bit word, [phase : 32 | expected : 16 | pending : 16], with the PTX semantics the
//    kernels rely on: an arrival decrements `pending`, the last one starts the next phase;
//    try_wait.parity(P) succeeds once the CURRENT phase's parity differs from P.  A waiter that is
//    overtaken by two phase completions therefore never wakes up -- on the GPU and here.
//  * division: on the GPU a/b == div_nr(a, b, rcp_nr(b)) bit for bit (mmf_selftest_division checks
//    that on the device); the reciprocal seed instruction does not exist on the CPU, so here div_nr IS
//    the IEEE division.  The emulator checks protocol and indexing, not that identity.
//  * the approximate FP32 operations only feed the eigenvalue ESTIMATE (never compared bitwise).
#pragma once

#include <cuda_runtime.h>

namespace mmf {

inline void mbar_init(unsigned long long *bar, unsigned count)
{
    __atomic_store_n(bar, ((unsigned long long) count << 16) | count, __ATOMIC_SEQ_CST);
}

inline void mbar_arrive(unsigned long long *bar)
{
    emu::chaos_delay();
    unsigned long long old = __atomic_load_n(bar, __ATOMIC_RELAXED), upd;
    do {
        const unsigned long long phase = old >> 32, expected = (old >> 16) & 0xffffu, pending = old & 0xffffu;
        if (pending == 0) { fprintf(stderr, "emu: arrival on an uninitialised mbarrier\n"); abort(); }
        upd = (pending == 1) ? (((phase + 1) << 32) | (expected << 16) | expected)
                             : ((phase << 32) | (expected << 16) | (pending - 1));
    } while (!__atomic_compare_exchange_n(bar, &old, upd, false, __ATOMIC_ACQ_REL, __ATOMIC_RELAXED));
}

inline void mbar_wait(unsigned long long *bar, unsigned parity)
{
    emu::chaos_delay();
    long long spins = 0;
    while (((__atomic_load_n(bar, __ATOMIC_ACQUIRE) >> 32) & 1u) == (parity & 1u)) {
        emu::yield_lane();
        emu::spin_pause(spins, "mbar_wait");
    }
}

#define MMF_EXP_NOSYNC 0
#define MMF_ARRIVE_PRED 1
inline void mbar_arrive_if(unsigned long long *bar, bool on) { if (on) mbar_arrive(bar); }

// TMA loads: the descriptor is a plain struct here and the copy is synchronous (what lies outside the tensor extents
// is zero filled like the hardware does).  The producer's protocol is mbar_expect_tx -> tma_load_4d ... -> mbar_arrive:
// with synchronous copies the announced byte count has nothing left to do and the one arrival completes the phase
// after the data is in place, which is the ordering a consumer may rely on.
struct TmaDesc {
    const double *base;      // element (0,0,0,0)
    long long stride[4];     // in elements
    int dim[4], box[4];
};
inline unsigned long long global_timer_ns() { return 0; }
inline void fence_barrier_init() {}
inline void fence_proxy_async_global() {}
inline void mbar_expect_tx(unsigned long long *, unsigned) {}
inline void tma_load_4d(const TmaDesc *d, void *smem_dst, unsigned long long *, int c0, int c1, int c2, int c3)
{
    emu::chaos_delay();
    if ((c0 * 8) % 16) { fprintf(stderr, "emu: bulk tensor load from a box row that does not start on a 16-byte boundary\n"); abort(); }
    double *dst = static_cast<double *>(smem_dst);
    const int c[4] = { c0, c1, c2, c3 };
    for (int q3 = 0; q3 < d->box[3]; ++q3)
        for (int q2 = 0; q2 < d->box[2]; ++q2)
            for (int q1 = 0; q1 < d->box[1]; ++q1)
                for (int q0 = 0; q0 < d->box[0]; ++q0) {
                    const int x[4] = { c[0] + q0, c[1] + q1, c[2] + q2, c[3] + q3 };
                    bool in = true;
                    for (int e = 0; e < 4; ++e) in = in && x[e] >= 0 && x[e] < d->dim[e];
                    dst[((q3 * d->box[2] + q2) * d->box[1] + q1) * d->box[0] + q0] =
                        in ? d->base[x[0] * d->stride[0] + x[1] * d->stride[1] + x[2] * d->stride[2] + x[3] * d->stride[3]] : 0.0;
                }
}

inline double rcp_nr(double b) { return 1.0 / b; }
inline double div_nr(double a, double b, double) { return a / b; }

inline float rcp_approx_f32(float x) { return 1.0f / x; }
inline float sqrt_approx_f32(float x) { return sqrtf(x); }

} // namespace mmf
