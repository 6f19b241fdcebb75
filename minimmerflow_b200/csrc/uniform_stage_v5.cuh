// uniform_stage_v5.cuh -- fused residual + RK-stage kernels ("low-face streaming"): the scheme and the pieces every
// form shares (the kernels themselves: uniform_stage_v5r.cuh and uniform_stage_t.cuh; uniform_body_cells.cuh for a box with bodies).
//
// Tiling (uniform_kernels.cuh): a warp owns a 32-cell x window (30 updated, +-x data by warp shuffle), the
// CTA's NW warps are consecutive y rows (rows 0 and NW-1 are halo rows), the CTA marches along z with plane
// k-1 in registers.
//   * every cell evaluates its three LOW interfaces (i-1|i), (j-1|j), (k-1|k).  In the reference's
//     accumulation order (src/euler.cpp:153, 237-247) the low faces of a cell come first and the high
//     faces last, so the running sum is formed as the data arrives: S = ordered sum of the three locally
//     computed low faces, then -x_hi (one shuffle from lane+1), then -y_hi (row+1's low face, through
//     shared memory), then -z_hi one plane later.  No flux array waits in registers for a neighbour.
//   * rows hand their record (U, Fy, lam_y) up and the low y flux down through shared memory, ordered by
//     mbarriers that take ONE arrival (elected lane after __syncwarp).
//   * halo rows prefetch their loads one plane ahead.
// Arithmetic per value is the same sequence of IEEE operations as the reference's (same helpers in every
// form), so the results stay bit-identical to the CPU restatement (tests/test_uniform_gpu.py).
// History (measured, profiles/): a high-face form (v3), a ping-pong register allocation of this scheme, a form
// with the y exchange decoupled by one plane (with and without one warp serving both halo rows), two y rows
// per warp, warp-specialised register budgets (setmaxnreg: 12 update warps at 144 registers + a light
// warpgroup) all lost to the 12-warp rotate form of uniform_stage_v5r.cuh and were removed in round 2
// (profiles/r02c_experiments.md).
#pragma once

#include "uniform_device.cuh"

namespace mmf {

__device__ __forceinline__ void mbar_arrive_elect(unsigned long long *bar, int lane)
{
    __syncwarp();
#if MMF_ARRIVE_PRED
    mbar_arrive_if(bar, lane == 0); // one predicated instruction: no branch, the warp stays converged for the compiler
#else
    if (lane == 0) mbar_arrive(bar);
#endif
}

// interface-order key of a cell's low face along `axis`: Morton 3*ctz(coord)+axis (the creator of
// the face is the lower cell; larger key = created earlier), lexicographic: axis; -1 = the face lies
// on the domain border and belongs to the cell's own group
template <int ORDER>
__device__ __forceinline__ int order_key(int coord, int axis)
{
    if (coord == 0) return -1;
    return (ORDER == NUM_LEXI) ? axis : 3 * (__ffs(coord) - 1) + axis;
}

// FP32 ESTIMATE of a cell's largest side eigenvalue max_d |u_d| + a (src/euler.cpp:59-66).  Stage 3
// evaluates it for the state it writes and keeps the maximum per CTA; the next step's exact max
// eigenvalue (src/main.cpp:398) is then found by re-evaluating in FP64 only the tiles whose estimate
// comes within EIG_EST_MARGIN of the largest estimate (uniform_eig_tiles_kernel): the cell holding
// the true maximum is necessarily in one of them, because an estimate is off by far less than the
// margin (relative error ~1e-6: two approximate SFU operations, FP32 rounding, and the cancellation
// in E - kinetic energy).
constexpr float EIG_EST_MARGIN = 0.999f;

__device__ __forceinline__ float eig_estimate(const double *c)
{
    const float r = (float) c[FID_RHO], mx = (float) c[FID_RHO_U], my = (float) c[FID_RHO_V], mz = (float) c[FID_RHO_W];
    const float inv = rcp_approx_f32(r);
    const float u = mx * inv, v = my * inv, w = mz * inv;
    const float p = 0.4f * ((float) c[FID_RHO_E] - 0.5f * (mx * u + my * v + mz * w));
    return fmaxf(fmaxf(fabsf(u), fabsf(v)), fabsf(w)) + sqrt_approx_f32(1.4f * p * inv);
}

// RK stage applied to one finished residual (src/main.cpp:409-423, 445-459, 481-495)
template <int STAGE>
__device__ __forceinline__ void finish_plane(const double *S, const double *AFz, const double *pU, const double *pUn,
                                             double dt, double volume, double y_vol, double *op, long long fs, bool pred,
                                             float &est_max)
{
    double out[NF];
#pragma unroll
    for (int k = 0; k < NF; ++k) {
        const double rhs = S[k] - AFz[k];
        if (STAGE == 0) {
            out[k] = rhs;
        } else {
            const double dq = div_nr(dt * rhs, volume, y_vol); // dt * RHS[k] / cellVolume
            if (STAGE == 1)      out[k] = pU[k] + dq;
            else if (STAGE == 2) out[k] = 0.75 * pUn[k] + 0.25 * (pU[k] + dq);
            else                 out[k] = (1. / 3) * pUn[k] + (2. / 3) * (pU[k] + dq);
        }
        if (pred) op[k * fs] = out[k];
    }
    if (STAGE == 3) {
        const float est = eig_estimate(out);
        est_max = fmaxf(est_max, pred ? est : 0.f); // fmaxf drops a NaN from a never-stored halo lane
    }
}

// which tile of the box a CTA works on: the block index, or -- multi-GPU with peer stores -- the entry
// of the interior-first visiting order
struct TileId {
    int bx, by, bz, tile;
};

__device__ __forceinline__ TileId stage_tile(const HaloWait &hw)
{
    TileId t;
    if (hw.tile_order) {
        t.tile = hw.tile_order[blockIdx.x];
        t.bx = t.tile % hw.tx;
        t.by = (t.tile / hw.tx) % hw.ty;
        t.bz = t.tile / (hw.tx * hw.ty);
    } else {
        t.bx = blockIdx.x; t.by = blockIdx.y; t.bz = blockIdx.z;
        t.tile = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    }
    return t;
}

// CTAs on a partition side wait for the neighbours' layers of the array they are about to read
// (the counters are raised by uniform_push_kernel on the neighbour GPUs, comm.cuh)
__device__ __forceinline__ void halo_wait(const HaloWait &hw, const TileId &t)
{
    if (!hw.flags) return;
    unsigned int touched = 0;
    if (t.bx == 0) touched |= 1u;
    if (t.bx == hw.tx - 1) touched |= 2u;
    if (t.by == 0) touched |= 4u;
    if (t.by == hw.ty - 1) touched |= 8u;
    if (t.bz == 0) touched |= 16u;
    if (t.bz == hw.tz - 1) touched |= 32u;
    const unsigned int need = touched & hw.mask;
    if (need && threadIdx.x < 6 && ((need >> threadIdx.x) & 1u)) {
        if (!halo_spin(hw.flags + threadIdx.x, hw.seq, hw.timeout_ns)) atomicAdd(hw.timeouts, 1.0);
        __threadfence_system();
    }
}

// block-wide maxima at the end of a stage kernel: the face eigenvalue (warp shuffle, one value per
// warp through shared memory, one integer atomic per CTA) and, for stage 3, the CTA's eigenvalue
// estimate, stored per CTA (no atomic)
template <int NW>
__device__ __forceinline__ void block_maxima(double v, double *out, float est, float *cta_est, int tile, double *scratch)
{
    const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
    __syncthreads(); // every warp is done with the exchange buffers
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double a = __shfl_xor_sync(0xffffffffu, v, o);
        v = (v < a) ? a : v;
        est = fmaxf(est, __shfl_xor_sync(0xffffffffu, est, o));
    }
    if (lane == 0) { scratch[row] = v; scratch[NW + row] = (double) est; }
    __syncthreads();
    if (row == 0) {
        double a = (lane < NW) ? scratch[lane] : 0.0, b = (lane < NW) ? scratch[NW + lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double x = __shfl_xor_sync(0xffffffffu, a, o), y = __shfl_xor_sync(0xffffffffu, b, o);
            a = (a < x) ? x : a;
            b = (b < y) ? y : b;
        }
        if (lane == 0) {
            atomic_max_nonneg(out, a);
            if (cta_est) cta_est[tile] = (float) b;
        }
    }
}

// Registers per thread for a CTA of NW warps that owns a whole SM.  The register file is handed out
// in units of four warps, so 13..16 warps all get 128 registers per thread (a 14-warp CTA at 144
// fails to launch) and 9..12 warps get 168; only the 4-warp steps are worth instantiating.
__host__ __device__ constexpr int stage_regs(int nw)
{
    return (65536 / ((nw + 3) / 4 * 4 * 32) / 8 * 8 > 255) ? 248 : 65536 / ((nw + 3) / 4 * 4 * 32) / 8 * 8;
}

} // namespace mmf
