// uniform_stage_v5.cuh -- fused residual + RK-stage kernel, fifth generation ("low-face streaming").
//
// Tiling as before (uniform_kernels.cuh): a warp owns a 32-cell x window (30 updated, +-x data by
// warp shuffle), the CTA's NW warps are consecutive y rows (rows 0 and NW-1 are halo rows), the CTA
// marches along z with plane k-1 in registers.
//
// What changed against v3, driven by the ncu captures in profiles/ (the kernel is latency bound:
// one warp needs ~2700 cycles per plane whether 2 or 3 warps share a scheduler, the FP64 pipe sits
// at 45 %, and 168 registers per thread cap the SM at 12 warps):
//   * every cell evaluates its three LOW interfaces (i-1|i), (j-1|j), (k-1|k) instead of the high
//     ones.  In the reference's accumulation order (src/euler.cpp:153, 237-247) the low faces of a
//     cell come first and the high faces last, so the running sum can now be formed as the data
//     arrives: S = ordered sum of the three locally computed low faces, then -x_hi (one shuffle from
//     lane+1), then -y_hi (row+1's low face, through shared memory), then -z_hi one plane later.
//     No flux array waits in registers for a neighbour any more (v3 held 25 doubles across both
//     waits), and what follows the last wait is one subtraction.
//   * that brings the kernel under 128 registers: 16 warps per SM instead of 12.
//   * mbarriers take ONE arrival (elected lane after __syncwarp) instead of 32 same-address ones.
//   * halo rows prefetch their loads one plane ahead.
//   * the first and last plane are peeled out of the steady-state loop, which is unrolled by two
//     over a ping-pong pair of plane states: no register rotation (~65 moves per plane in v3, and
//     ptxas liked to place the copy of the prefetched plane right behind its own load, so that
//     every plane stalled on DRAM latency).  The single-body "rotate" form is kept in
//     uniform_stage_v5r.cuh because it schedules better for stages 2 and 3 (see there).
// Arithmetic per value is the same sequence of IEEE operations as in v1/v3 (same helpers), so the
// results stay bit-identical to the CPU restatement of the reference (tests/test_uniform_gpu.py).
#pragma once

#include "uniform_device.cuh"

namespace mmf {

__device__ __forceinline__ void mbar_arrive_elect(unsigned long long *bar, int lane)
{
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
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
        const volatile unsigned long long *f = hw.flags + threadIdx.x;
        while (*f < hw.seq) { }
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

// state of one plane of one cell as it travels through two iterations
struct PlaneState {
    double U[NF];   // residual input of the plane
    double Fz[NF];  // z flux of the cell, lam_z below
    double lz;
    double S[NF];   // running residual: everything except -z_hi
    double Un[NF];  // U^n of the plane (stages 2, 3)
};

// everything an update row needs per plane that does not change from plane to plane
template <int STAGE, int ORDER>
struct RowCtx {
    bool upd;
    int lane, key_x, key_y, gz0, khi;
    double dt, Ah, volume;
    DivConsts dc;
    long long fs, plane;   // own cells: U^n loads and stores
    long long sfs, splane; // this lane's column of the residual input (padded array or compact x ghosts)
    double *d_own, *f_own;
    const double *d_dn, *f_up;
    unsigned long long *barD_own, *barD_dn, *barF_own, *barF_up;
    const double *sp, *unp;
    double *op;
    double lmx, lmy, lmz;
    float est_max;

    // One plane.  On entry P is the finished state of plane kz-1 (S lacks -z_hi) and C.U holds
    // plane kz; on exit C is the finished state of plane kz and P.U holds plane kz+1.
    __device__ __forceinline__ void body(PlaneState &P, PlaneState &C, int kz, int z0)
    {
        const unsigned par = (unsigned) ((kz - z0) & 1);
        if (STAGE >= 2 && upd) {
#pragma unroll
            for (int k = 0; k < NF; ++k) C.Un[k] = unp[k * fs];
        }
        unp += plane;

        CellPrim q;
        derive_cell(C.U, dc, q);

        // ---- y record for row+1 (the earlier it is out, the less row+1 waits) -------------------------
        double cFy[NF], cly;
        axis_flux<1>(q, cFy, cly);
#pragma unroll
        for (int k = 0; k < NF; ++k) { d_own[k * 32] = C.U[k]; d_own[(NF + k) * 32] = cFy[k]; }
        d_own[10 * 32] = cly;
        mbar_arrive_elect(barD_own, lane);

        // ---- z interface (kz-1 | kz): completes plane kz-1 --------------------------------------------
        double AFz[NF];
        axis_flux<2>(q, C.Fz, C.lz);
        {
            const double lam = llf_area_flux(P.U, P.Fz, P.lz, C.U, C.Fz, C.lz, Ah, AFz);
            lmz = (lam < lmz) ? lmz : lam;
        }
        finish_plane<STAGE>(P.S, AFz, P.U, P.Un, dt, volume, dc.y_vol, op, fs, upd && kz > z0, est_max);
        op += plane;

        // ---- x interface (i-1 | i): lane-1's state by warp shuffle -------------------------------------
        double AFx[NF];
        {
            double cFx[NF], clx, lU[NF], lF[NF];
            axis_flux<0>(q, cFx, clx);
#pragma unroll
            for (int k = 0; k < NF; ++k) { lU[k] = shfl_up_d(C.U[k]); lF[k] = shfl_up_d(cFx[k]); }
            const double ll  = shfl_up_d(clx);
            const double lam = llf_area_flux(lU, lF, ll, C.U, cFx, clx, Ah, AFx);
            lmx = (lam < lmx) ? lmx : lam;
        }

        // ---- y interface (j-1 | j): row-1's record through shared memory ------------------------------
        double AFy[NF];
        mbar_wait(barD_dn, par);
        // ---- plane kz+1 (the ghost plane nz, or plane nz-1 again on a free-flow side) into the registers
        //      plane kz-1 left in finish_plane.  Issued BEHIND the acquire on purpose: ptxas cannot
        //      hoist the loads above it into the live range of the old values, which made it park the
        //      results in other registers and copy them right behind the loads (a DRAM round trip
        //      exposed in every plane, 15 % of all stall samples in profiles/r01b).
        if (kz + 1 <= khi) sp += splane;
#pragma unroll
        for (int k = 0; k < NF; ++k) P.U[k] = ldsin(sp + k * sfs);
        {
            double lU[NF], lF[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) { lU[k] = d_dn[k * 32]; lF[k] = d_dn[(NF + k) * 32]; }
            const double ll  = d_dn[10 * 32];
            const double lam = llf_area_flux(lU, lF, ll, C.U, cFy, cly, Ah, AFy);
            lmy = (lam < lmy) ? lmy : lam;
            // row-1 published record `it` only after it had read this row's flux `it-1`
#pragma unroll
            for (int k = 0; k < NF; ++k) f_own[k * 32] = AFy[k];
            mbar_arrive_elect(barF_own, lane);
        }

        // ---- ordered accumulation (src/euler.cpp:153, 237-247) ----------------------------------------
        // interior low faces first, sorted by their creator (largest key first; a+b commutes, so only
        // the LAST one matters); then the cell's own faces in the order it created them:
        // (-x if border) +x (-y if border) +y (-z if border) +z; low faces `+=`, high faces `-=`.
        const int key_z = order_key<ORDER>(gz0 + kz, 2);
        const bool edge = (key_y < 0) | (key_z < 0); // warp-uniform: one row, one plane per warp
        if (ORDER == NUM_AXIS) {
#pragma unroll
            for (int k = 0; k < NF; ++k) C.S[k] = 0.0 + AFx[k];
        } else if (!edge) {
            // a border low face in x alone is simply "last" (key -1), directly followed by -x_hi
            const int last = (key_x < key_y) ? ((key_x < key_z) ? 0 : 2) : ((key_y < key_z) ? 1 : 2);
#pragma unroll
            for (int k = 0; k < NF; ++k) {
                const double p = (last == 0) ? AFy[k] : AFx[k];
                const double t = (last == 2) ? AFy[k] : AFz[k];
                const double r = (last == 0) ? AFx[k] : (last == 1) ? AFy[k] : AFz[k];
                C.S[k] = (p + t) + r;
            }
        } else {
            // low y / low z side of the domain: those faces enter after -x_hi, see below
            const bool bx = key_x < 0;
#pragma unroll
            for (int k = 0; k < NF; ++k) {
                double s = 0.0; // at most two interior low faces remain: their order is immaterial
                if (!bx) s += AFx[k];
                if (key_y >= 0) s += AFy[k];
                if (key_z >= 0) s += AFz[k];
                if (bx) s += AFx[k];
                C.S[k] = s;
            }
        }
        // -x_hi: the low x face of lane+1
#pragma unroll
        for (int k = 0; k < NF; ++k) C.S[k] -= shfl_down_d(AFx[k]);
        if (ORDER == NUM_AXIS || (edge && key_y < 0)) {
#pragma unroll
            for (int k = 0; k < NF; ++k) C.S[k] += AFy[k];
        }
        // -y_hi: the low y face of row+1
        mbar_wait(barF_up, par);
#pragma unroll
        for (int k = 0; k < NF; ++k) C.S[k] -= f_up[k * 32];
        if (ORDER == NUM_AXIS || (edge && key_z < 0)) {
#pragma unroll
            for (int k = 0; k < NF; ++k) C.S[k] += AFz[k];
        }
        // -z_hi follows in the next iteration (or the epilogue)
    }

    // plane z1 only closes the last z interface
    __device__ __forceinline__ void epilogue(PlaneState &P, const double *nU)
    {
        CellPrim q;
        derive_cell(nU, dc, q);
        double cFz[NF], clz, AFz[NF];
        axis_flux<2>(q, cFz, clz);
        const double lam = llf_area_flux(P.U, P.Fz, P.lz, nU, cFz, clz, Ah, AFz);
        lmz = (lam < lmz) ? lmz : lam;
        finish_plane<STAGE>(P.S, AFz, P.U, P.Un, dt, volume, dc.y_vol, op, fs, upd, est_max);
    }
};

template <int STAGE, int ORDER, int NW, bool XG>
__global__ void __maxnreg__(stage_regs(NW))
uniform_stage_kernel_v5(const UniformGeom g, const double *__restrict__ Sin, const double *Un, double *Out,
                        const StepControl *__restrict__ ctl, double *__restrict__ max_eig, const int lz,
                        float *__restrict__ cta_est, const LoadClamp lc, const HaloWait hw, const XGhost xg)
{
    extern __shared__ double smem[];
    // sm_d[row][q][lane], q = U0..U4, Fy0..Fy4, lam_y ; sm_f[row][k][lane] = area * flux of (j-1 | j)
    double *sm_d = smem;
    double *sm_f = smem + NW * 11 * 32;
    unsigned long long *barD = reinterpret_cast<unsigned long long *>(sm_f + NW * NF * 32); // record of row r published
    unsigned long long *barF = barD + NW;                                                    // low y flux of row r published

    if (STAGE >= 1 && ctl->active == 0.0) return;

    const int lane = threadIdx.x & 31;
    const int row  = threadIdx.x >> 5;
    const TileId tid = stage_tile(hw);
    if (threadIdx.x < NW) {
        mbar_init(&barD[threadIdx.x], 1);
        mbar_init(&barF[threadIdx.x], 1);
    }
    halo_wait(hw, tid);
    __syncthreads();

    const int i  = tid.bx * XW - 1 + lane;
    const int j  = tid.by * (NW - 2) - 1 + row;
    const int z0 = tid.bz * lz;
    const int z1 = min(z0 + lz, g.nz);
    const int ic = min(max(i, lc.ilo), lc.ihi); // load coordinates (free-flow sides re-read the boundary cell)
    const int jc = min(max(j, lc.jlo), lc.jhi);
    const bool in_x = (i >= 0 && i < g.nx);
    const bool in_y = (j >= 0 && j < g.ny);

    const double Ah = 0.5 * g.area;
    DivConsts dc;
    dc.y_gm1 = rcp_nr(GM1);
    dc.y_c1  = rcp_nr(TWO_OVER_GM1);
    dc.y_vol = rcp_nr(g.volume);

    const long long plane = (long long) g.py * g.px;
    const long long fs    = g.fs;
    const long long col   = (long long) (jc + 1) * g.px + (ic + 1);
    // where this lane's column of the residual input lives: the padded array, or -- halo lanes across an
    // x partition side -- the compact ghost columns
    const double *scol = Sin + col;
    // (XG = false instantiations keep the strides uniform: per-lane strides cost registers the
    //  single-GPU kernels do not have to spare)
    int sfs_lane = (int) fs, splane_lane = (int) plane; // element counts: < 2^31 for any box that fits one GPU
    if (XG) {
        if (xg.lo && i < 0)     { scol = xg.lo + (jc + 1); sfs_lane = (int) xg.fs; splane_lane = xg.pitch; }
        if (xg.hi && i >= g.nx) { scol = xg.hi + (jc + 1); sfs_lane = (int) xg.fs; splane_lane = xg.pitch; }
    }
    const long long sfs = XG ? (long long) sfs_lane : fs, splane = XG ? (long long) splane_lane : plane;
    double lmax = 0.0;
    float emax = 0.f;

    if (row == 0) {
        // ================= low halo row: publishes (U, Fy, lam_y) of row j for row 1 =================
        const double *sp = scol + (long long) (z0 + 1) * splane; // plane z0
        double *d = sm_d + lane;
        double nxt[NF];
#pragma unroll
        for (int k = 0; k < NF; ++k) nxt[k] = ldsin(sp + k * sfs);
        for (int kz = z0; kz < z1; ++kz) {
            const int it = kz - z0;
            double cU[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) cU[k] = nxt[k];
            sp += splane;
            if (kz + 1 < z1) {
#pragma unroll
                for (int k = 0; k < NF; ++k) nxt[k] = ldsin(sp + k * sfs);
            }
            CellPrim q;
            derive_cell(cU, dc, q);
            double cFy[NF], cly;
            axis_flux<1>(q, cFy, cly);
            if (it > 0) mbar_wait(&barF[1], (unsigned) ((it - 1) & 1)); // row 1 is done with the previous record
#pragma unroll
            for (int k = 0; k < NF; ++k) { d[k * 32] = cU[k]; d[(NF + k) * 32] = cFy[k]; }
            d[10 * 32] = cly;
            mbar_arrive_elect(&barD[0], lane);
        }
    } else if (row == NW - 1) {
        // ================= high halo row: computes the y face (j-1 | j) for row NW-2 ================
        const bool yf_ok = in_x && lane >= 1 && lane <= XW && j >= 0 && j <= g.ny;
        const double *sp = scol + (long long) (z0 + 1) * splane; // plane z0
        const double *d_dn = sm_d + (NW - 2) * 11 * 32 + lane;
        double *f = sm_f + (NW - 1) * NF * 32 + lane;
        double lmy = 0.0;
        double nxt[NF];
#pragma unroll
        for (int k = 0; k < NF; ++k) nxt[k] = ldsin(sp + k * sfs);
        for (int kz = z0; kz < z1; ++kz) {
            const int it = kz - z0;
            double cU[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) cU[k] = nxt[k];
            sp += splane;
            if (kz + 1 < z1) {
#pragma unroll
                for (int k = 0; k < NF; ++k) nxt[k] = ldsin(sp + k * sfs);
            }
            CellPrim q;
            derive_cell(cU, dc, q);
            double cFy[NF], cly;
            axis_flux<1>(q, cFy, cly);
            mbar_wait(&barD[NW - 2], (unsigned) (it & 1));
            double lU[NF], lF[NF], AFy[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) { lU[k] = d_dn[k * 32]; lF[k] = d_dn[(NF + k) * 32]; }
            const double ll  = d_dn[10 * 32];
            const double lam = llf_area_flux(lU, lF, ll, cU, cFy, cly, Ah, AFy);
            lmy = (lam < lmy) ? lmy : lam;
            // row NW-2 published record `it` only after it had read flux `it-1`: the slot is free
#pragma unroll
            for (int k = 0; k < NF; ++k) f[k * 32] = AFy[k];
            mbar_arrive_elect(&barF[NW - 1], lane);
        }
        lmax = yf_ok ? lmy : 0.0;
    } else {
        // ================= update rows ==============================================================
        RowCtx<STAGE, ORDER> c;
        c.upd   = lane >= 1 && lane <= XW && in_x && in_y;
        c.lane  = lane;
        c.dt    = (STAGE >= 1) ? ctl->dt : 0.0;
        c.Ah    = Ah;
        c.volume = g.volume;
        c.dc    = dc;
        c.key_x = order_key<ORDER>(g.gx0 + i, 0);
        c.key_y = order_key<ORDER>(g.gy0 + j, 1);
        c.gz0   = g.gz0;
        c.khi   = lc.khi;
        c.fs    = fs;
        c.plane = plane;
        c.sfs   = sfs;
        c.splane = splane;
        c.d_own = sm_d + row * 11 * 32 + lane;
        c.d_dn  = sm_d + (row - 1) * 11 * 32 + lane;
        c.f_own = sm_f + row * NF * 32 + lane;
        c.f_up  = sm_f + (row + 1) * NF * 32 + lane;
        c.barD_own = &barD[row];
        c.barD_dn  = &barD[row - 1];
        c.barF_own = &barF[row];
        c.barF_up  = &barF[row + 1];
        c.sp  = scol + (long long) (max(z0 - 1, lc.klo) + 1) * splane; // plane z0-1 (clamped)
        c.unp = Un + col + (long long) (z0 + 1) * plane;   // plane z0
        c.op  = Out + col + (long long) z0 * plane;        // plane z0-1 (the first store goes to plane z0)
        c.lmx = c.lmy = c.lmz = 0.0;
        c.est_max = 0.f;

        // Two plane states in ping-pong: after body(P, C) the roles swap, so nothing is ever copied
        // from "current" to "previous" registers (the rotate form spends ~65 moves per plane on that,
        // and ptxas may place the copy of the prefetched plane right behind its own load).
        PlaneState A, B;
        // ---- prologue: plane z0-1 only provides the low side of the first z interface --------------
        {
#pragma unroll
            for (int k = 0; k < NF; ++k) A.U[k] = ldsin(c.sp + k * sfs);
            c.sp = scol + (long long) (z0 + 1) * splane; // plane z0
#pragma unroll
            for (int k = 0; k < NF; ++k) B.U[k] = ldsin(c.sp + k * sfs);
            CellPrim q;
            derive_cell(A.U, dc, q);
            axis_flux<2>(q, A.Fz, A.lz);
#pragma unroll
            for (int k = 0; k < NF; ++k) { A.S[k] = 0.0; A.Un[k] = 0.0; }
        }
        int kz = z0;
        for (; kz + 1 < z1; kz += 2) {
            c.body(A, B, kz, z0);
            c.body(B, A, kz + 1, z0);
        }
        if (kz < z1) {
            c.body(A, B, kz, z0);
            c.epilogue(B, A.U);
        } else {
            c.epilogue(A, B.U);
        }
        const bool xf_ok = in_y && lane >= 1 && i >= 0 && i <= g.nx;                  // face (i-1 | i)
        const bool yf_ok = in_x && lane >= 1 && lane <= XW && j >= 0 && j <= g.ny;   // face (j-1 | j)
        const bool zf_ok = in_x && in_y;                                              // face (k-1 | k)
        lmax = xf_ok ? c.lmx : 0.0;
        if (yf_ok) lmax = (c.lmy < lmax) ? lmax : c.lmy;
        if (zf_ok) lmax = (c.lmz < lmax) ? lmax : c.lmz;
        emax = c.est_max;
    }

    block_maxima<NW>(lmax, max_eig, emax, (STAGE == 3) ? cta_est : nullptr, tid.tile, smem);
}

} // namespace mmf
