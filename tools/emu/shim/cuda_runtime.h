// tools/emu/shim/cuda_runtime.h -- DEVELOPMENT TOOL, not product code, not a fallback.
//
// A minimal SIMT emulation that lets g++ compile the stage kernels' SOURCE (minimmerflow_b200/csrc/
// uniform_stage_*.cuh) and run it on the CPU: one OS thread per warp, the 32 lanes of a warp are
// ucontext fibers that switch at warp collectives (shuffles, __syncwarp) and inside spin loops,
// mbarriers keep the PTX phase/parity semantics (a producer that gets two phases ahead of a parity
// wait hangs here exactly as it would on the GPU -- and the watchdog reports it instead of costing a
// GPU box).  Used by tools/emu/run_emu.py to check a kernel variant's synchronisation protocol and
// index arithmetic against the CPU oracle BEFORE GPU time is spent on it.  Nothing in
// minimmerflow_b200/, bench.py or the C-ABI can reach this code; numbers are never measured here.
//
// This header shadows <cuda_runtime.h> for the emulator build only (g++ -I tools/emu/shim).
#pragma once

#include <ucontext.h>
#include <sched.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __align__(n)
#define __grid_constant__

namespace emu {

struct uint3e { unsigned x, y, z; };

struct Warp;

struct Lane {
    ucontext_t ctx;
    uint3e tid;
    Warp *warp;
    int lane;
    unsigned gen;   // warp collectives this lane has entered
    bool done;
    char *stack;
};

struct Cta; // emu_runtime.h

struct Warp {
    Lane lanes[32];
    ucontext_t main;
    uint64_t buf[2][32];
    int arrived[2], departed[2];
    Cta *cta;
    int index;
    unsigned rng;   // chaos mode: per-warp delay generator
};

extern thread_local Lane *tl_cur;
extern uint3e g_blockIdx, g_blockDim, g_gridDim;
extern int g_chaos;             // > 0: random delays around mbarrier operations (widens the warp skew)
extern long long g_spin_limit;  // watchdog: yields inside one wait before the run is declared hung

void yield_lane();              // hand the OS thread to the next live lane of this warp
void spin_pause(long long &spins, const char *what);
void chaos_delay();
void cta_barrier();

inline uint64_t exchange(uint64_t bits, int src)
{
    Lane *me = tl_cur;
    Warp *w = me->warp;
    const int slot = (int) (me->gen++ & 1u);
    w->buf[slot][me->lane] = bits;
    w->arrived[slot]++;
    long long spins = 0;
    while (w->arrived[slot] < 32) { yield_lane(); if (w->arrived[slot] < 32) spin_pause(spins, "warp collective (divergent?)"); }
    const uint64_t r = w->buf[slot][src];
    if (++w->departed[slot] == 32) { w->arrived[slot] = 0; w->departed[slot] = 0; }
    return r;
}

template <typename T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

} // namespace emu

#define threadIdx (emu::tl_cur->tid)
#define blockIdx  (emu::g_blockIdx)
#define blockDim  (emu::g_blockDim)
#define gridDim   (emu::g_gridDim)

template <typename T> inline T __shfl_xor_sync(unsigned, T v, int o)
{
    const int lane = emu::tl_cur->lane;
    return emu::from_bits<T>(emu::exchange(emu::to_bits(v), (lane ^ o) & 31));
}
template <typename T> inline T __shfl_up_sync(unsigned, T v, int d)
{
    const int lane = emu::tl_cur->lane;
    return emu::from_bits<T>(emu::exchange(emu::to_bits(v), lane - d >= 0 ? lane - d : lane));
}
template <typename T> inline T __shfl_down_sync(unsigned, T v, int d)
{
    const int lane = emu::tl_cur->lane;
    return emu::from_bits<T>(emu::exchange(emu::to_bits(v), lane + d < 32 ? lane + d : lane));
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu::exchange(0, emu::tl_cur->lane); }
inline void __syncthreads() { emu::cta_barrier(); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

inline int __ffs(int x) { return __builtin_ffs(x); }
template <typename T> inline T __ldcg(const T *p) { return *p; }
template <typename T> inline T __ldg(const T *p) { return *p; }
inline long long __double_as_longlong(double v) { return emu::from_bits<long long>(emu::to_bits(v)); }
inline double __longlong_as_double(long long v) { return emu::from_bits<double>(emu::to_bits(v)); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline double __dmul_rn(double a, double b) { return a * b; }
inline int __double2hiint(double v) { return (int) (emu::to_bits(v) >> 32); }
inline int __double2loint(double v) { return (int) (emu::to_bits(v) & 0xffffffffu); }
inline double __hiloint2double(int hi, int lo) { return emu::from_bits<double>(((uint64_t) (unsigned) hi << 32) | (unsigned) lo); }

inline unsigned long long atomicMax(unsigned long long *a, unsigned long long v)
{
    unsigned long long old = __atomic_load_n(a, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) { }
    return old;
}
template <typename T> inline T atomicAdd(T *a, T v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
inline double atomicAdd(double *a, double v)
{
    static std::mutex m;
    std::lock_guard<std::mutex> lock(m);
    const double old = *a;
    *a = old + v;
    return old;
}

using std::min;
using std::max;
inline long long min(long long a, int64_t b) { return a < (long long) b ? a : (long long) b; }
