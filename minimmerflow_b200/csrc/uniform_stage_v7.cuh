// uniform_stage_v7.cuh -- the plane-decoupled stage kernel (uniform_stage_v6.cuh) with TWO y rows per warp
// (kernel form 'w', opt-in until measured on the GPU).
//
// The v5/v6 kernels are bound by latency, not by a pipe: profiles/r01d has the FP64 pipe at 50 %, the issue
// slots at 57 %, and fixed-latency dependencies as the largest stall class -- three warps per scheduler do
// not cover the 11-cycle FP64 chains of a cell (reciprocals -> quotients -> sqrt -> fluxes), and a fourth
// warp does not fit the register file.  Here every thread owns two cells, (i, j) and (i, j+1):
//   * two independent dependency chains per thread -- instruction-level parallelism the scheduler gets
//     without another warp (CTAs are 8 warps at up to 248 registers: 7 update warps x 2 rows = 14 rows per
//     CTA against 10 or 11);
//   * the y interface between the two cells never leaves the thread: half the records, fluxes, mbarrier
//     operations and shared-memory traffic per cell;
//   * one halo warp serves both halo rows of the tile, as in form 'h'.
// Per value the sequence of IEEE operations is the one of v5/v6 (same helpers, same accumulation order):
// bit-identical results, checked on the CPU emulator (tools/emu) before any GPU time.
//
// Exchange between warps ("port" p = warp index; port 0 / NW belong to the halo warp), double buffered by
// plane parity with one mbarrier per slot and port exactly as in v6:
//   record[slot][p] = (U, Fy, lam_y) of warp p's UPPER row     (read by warp p+1 for its lower row's low y face)
//   flux[slot][p]   = area * flux of the low y face of warp p's LOWER row (warp p-1 subtracts it, one plane late)
#pragma once

#include "uniform_stage_v6.cuh"

namespace mmf {

__host__ __device__ constexpr int stage_v7_rows(int nw) { return 2 * (nw - 1); }

// ordered sum of a cell's three low faces, then -x_hi, then +y_lo where that face lies on the domain border
// (src/euler.cpp:153, 237-247; see uniform_stage_v5.cuh for the order)
template <int ORDER>
__device__ __forceinline__ void lows_minus_xhi(const double *AFx, const double *AFy, const double *AFz, const int key_x,
                                               const int key_y, const int key_z, double *S)
{
    const bool edge = (key_y < 0) | (key_z < 0); // warp-uniform
    if (ORDER == NUM_AXIS) {
#pragma unroll
        for (int k = 0; k < NF; ++k) S[k] = 0.0 + AFx[k];
    } else if (!edge) {
        const int last = (key_x < key_y) ? ((key_x < key_z) ? 0 : 2) : ((key_y < key_z) ? 1 : 2);
#pragma unroll
        for (int k = 0; k < NF; ++k) {
            const double p = (last == 0) ? AFy[k] : AFx[k];
            const double t = (last == 2) ? AFy[k] : AFz[k];
            const double r = (last == 0) ? AFx[k] : (last == 1) ? AFy[k] : AFz[k];
            S[k] = (p + t) + r;
        }
    } else {
        const bool bx = key_x < 0;
#pragma unroll
        for (int k = 0; k < NF; ++k) {
            double s = 0.0;
            if (!bx) s += AFx[k];
            if (key_y >= 0) s += AFy[k];
            if (key_z >= 0) s += AFz[k];
            if (bx) s += AFx[k];
            S[k] = s;
        }
    }
#pragma unroll
    for (int k = 0; k < NF; ++k) S[k] -= shfl_down_d(AFx[k]);
    if (ORDER == NUM_AXIS || (edge && key_y < 0)) {
#pragma unroll
        for (int k = 0; k < NF; ++k) S[k] += AFy[k];
    }
}

// derive_cell for two independent cells with the statements of the two dependency chains interleaved in
// program order (ptxas keeps two equally long chains apart otherwise, and the point of two cells per thread
// is that one chain's 11-cycle FP64 latencies are filled by the other).  Same operations per cell as
// derive_cell, hence the same bits.
__device__ __forceinline__ void derive_cell2(const double *c0, const double *c1, const DivConsts &dc, CellPrim &q0, CellPrim &q1)
{
    const double rho0 = c0[FID_RHO], rho1 = c1[FID_RHO];
    const double rr0 = rho0 * rho0, rr1 = rho1 * rho1;
    double y0, y1, yrr0, yrr1;
    rcp_nr2(rho0, rho1, y0, y1);
    rcp_nr2(rr0, rr1, yrr0, yrr1);
    const double m0 = c0[FID_RHO_U] * c0[FID_RHO_U] + c0[FID_RHO_V] * c0[FID_RHO_V] + c0[FID_RHO_W] * c0[FID_RHO_W];
    const double m1 = c1[FID_RHO_U] * c1[FID_RHO_U] + c1[FID_RHO_V] * c1[FID_RHO_V] + c1[FID_RHO_W] * c1[FID_RHO_W];
    const double K0 = div_nr(m0, rr0, yrr0);
    const double K1 = div_nr(m1, rr1, yrr1);
    const double e0 = div_nr(2.0 * c0[FID_RHO_E], rho0, y0);
    const double e1 = div_nr(2.0 * c1[FID_RHO_E], rho1, y1);
    const double T0 = div_nr(e0 - K0, TWO_OVER_GM1, dc.y_c1);
    const double T1 = div_nr(e1 - K1, TWO_OVER_GM1, dc.y_c1);
    q0.rho = rho0; q1.rho = rho1;
    q0.u = div_nr(c0[FID_RHO_U], rho0, y0); q1.u = div_nr(c1[FID_RHO_U], rho1, y1);
    q0.v = div_nr(c0[FID_RHO_V], rho0, y0); q1.v = div_nr(c1[FID_RHO_V], rho1, y1);
    q0.w = div_nr(c0[FID_RHO_W], rho0, y0); q1.w = div_nr(c1[FID_RHO_W], rho1, y1);
    q0.p = rho0 * T0; q1.p = rho1 * T1;
    const double v0 = q0.u * q0.u + q0.v * q0.v + q0.w * q0.w;
    const double v1 = q1.u * q1.u + q1.v * q1.v + q1.w * q1.w;
    const double t0 = div_nr(q0.p, GM1, dc.y_gm1) + 0.5 * rho0 * v0;
    const double t1 = div_nr(q1.p, GM1, dc.y_gm1) + 0.5 * rho1 * v1;
    q0.H = t0 + q0.p; q1.H = t1 + q1.p;
    q0.a = sqrt(GAMMA * T0); q1.a = sqrt(GAMMA * T1);
}

template <int STAGE, int ORDER, int NW, bool XG>
__global__ void __maxnreg__(stage_regs(NW))
uniform_stage_kernel_v7(const UniformGeom g, const double *__restrict__ Sin, const double *Un, double *Out,
                        const StepControl *__restrict__ ctl, double *__restrict__ max_eig, const int lz,
                        float *__restrict__ cta_est, const LoadClamp lc, const HaloWait hw, const XGhost xg)
{
    extern __shared__ double smem[];
    constexpr int NR = NW + 1;                           // ports
    constexpr int RT = stage_v7_rows(NW);                // rows updated per tile
    constexpr int DS = NR * 11 * 32, FS = NR * NF * 32;  // doubles per slot
    double *sm_d = smem;
    double *sm_f = smem + 2 * DS;
    unsigned long long *barD = reinterpret_cast<unsigned long long *>(sm_f + 2 * FS); // [slot][port]: record published
    unsigned long long *barF = barD + 2 * NR;                                          // [slot][port]: flux published

    if (STAGE >= 1 && ctl->active == 0.0) return;

    const int lane = threadIdx.x & 31;
    const int w    = threadIdx.x >> 5;
    const TileId tid = stage_tile(hw);
    if (threadIdx.x < 2 * NR) {
        mbar_init(&barD[threadIdx.x], 1);
        mbar_init(&barF[threadIdx.x], 1);
    }
    halo_wait(hw, tid);
    __syncthreads();

    const int i  = tid.bx * XW - 1 + lane;
    const int j0 = tid.by * RT;                          // first row of the tile
    const int z0 = tid.bz * lz;
    const int z1 = min(z0 + lz, g.nz);
    const int n  = z1 - z0;
    const int ic = min(max(i, lc.ilo), lc.ihi);          // load coordinates (free-flow sides re-read the boundary cell)
    const bool in_x = (i >= 0 && i < g.nx);

    const double Ah = 0.5 * g.area;
    DivConsts dc;
    dc.y_gm1 = rcp_nr(GM1);
    dc.y_c1  = rcp_nr(TWO_OVER_GM1);
    dc.y_vol = rcp_nr(g.volume);

    const long long plane = (long long) g.py * g.px;
    const long long fs    = g.fs;
    // this lane's columns of the residual input: the padded array, or -- halo lanes across an x partition
    // side -- the compact ghost columns, where consecutive rows are consecutive elements
    const double *sbase = Sin + (ic + 1);
    int srow = g.px, sfs_lane = (int) fs, splane_lane = (int) plane;
    if (XG) {
        if (xg.lo && i < 0)     { sbase = xg.lo; srow = 1; sfs_lane = (int) xg.fs; splane_lane = xg.pitch; }
        if (xg.hi && i >= g.nx) { sbase = xg.hi; srow = 1; sfs_lane = (int) xg.fs; splane_lane = xg.pitch; }
    }
    const long long sfs = XG ? (long long) sfs_lane : fs, splane = XG ? (long long) splane_lane : plane;
    double lmax = 0.0;
    float emax = 0.f;

    if (w == 0) {
        // ================= halo warp: row j0-1 (record for warp 1) and row j0+RT (flux for warp NW-1) ===
        const int jb  = j0 - 1, jt = j0 + RT;
        const int jbc = min(max(jb, lc.jlo), lc.jhi), jtc = min(max(jt, lc.jlo), lc.jhi);
        const bool yf_ok = in_x && lane >= 1 && lane <= XW && jt >= 0 && jt <= g.ny;
        const double *sp_b = sbase + (long long) (jbc + 1) * srow + (long long) (z0 + 1) * splane; // plane z0
        const double *sp_t = sbase + (long long) (jtc + 1) * srow + (long long) (z0 + 1) * splane;
        double lmy = 0.0;
        double nb[NF], nt[NF];             // prefetched plane `it` of the bottom / top row
        double tU[NF], tFy[NF], tly = 0.0; // top row, plane it-1
#pragma unroll
        for (int k = 0; k < NF; ++k) { nb[k] = ldsin(sp_b + k * sfs); nt[k] = ldsin(sp_t + k * sfs); tU[k] = 0.0; tFy[k] = 0.0; }
#pragma unroll 1
        for (int it = 0; it <= n; ++it) {
            const int slot = it & 1;
            if (it < n) {
                double bU[NF], cT[NF], bFy[NF], bly, cFy[NF], cly;
#pragma unroll
                for (int k = 0; k < NF; ++k) { bU[k] = nb[k]; cT[k] = nt[k]; }
                sp_b += splane;
                sp_t += splane;
                if (it + 1 < n) {
#pragma unroll
                    for (int k = 0; k < NF; ++k) { nb[k] = ldsin(sp_b + k * sfs); nt[k] = ldsin(sp_t + k * sfs); }
                }
                CellPrim qb, qt; // two independent chains
                derive_cell(bU, dc, qb);
                derive_cell(cT, dc, qt);
                axis_flux<1>(qb, bFy, bly);
                axis_flux<1>(qt, cFy, cly);
                // bottom record `it`.  The slot still holds record it-2: warp 1 has read it once its flux it-2 is
                // out, and it cannot complete that barrier again before it has seen record `it`.
                if (it >= 2) mbar_wait(&barF[slot * NR + 1], (unsigned) (((it - 2) >> 1) & 1));
                double *d = sm_d + slot * DS + lane;
#pragma unroll
                for (int k = 0; k < NF; ++k) { d[k * 32] = bU[k]; d[(NF + k) * 32] = bFy[k]; }
                d[10 * 32] = bly;
                mbar_arrive_elect(&barD[slot * NR + 0], lane);
                if (it >= 1) {
                    // top flux of plane it-1 (see below), before the top state is replaced
                    const int ps = (it - 1) & 1;
                    mbar_wait(&barD[ps * NR + NW - 1], (unsigned) (((it - 1) >> 1) & 1));
                    const double *d_dn = sm_d + ps * DS + (NW - 1) * 11 * 32 + lane;
                    double lU[NF], lF[NF], AFy[NF];
#pragma unroll
                    for (int k = 0; k < NF; ++k) { lU[k] = d_dn[k * 32]; lF[k] = d_dn[(NF + k) * 32]; }
                    const double ll  = d_dn[10 * 32];
                    const double lam = llf_area_flux(lU, lF, ll, tU, tFy, tly, Ah, AFy);
                    lmy = (lam < lmy) ? lmy : lam;
                    double *f = sm_f + ps * FS + NW * NF * 32 + lane;
#pragma unroll
                    for (int k = 0; k < NF; ++k) f[k * 32] = AFy[k];
                    mbar_arrive_elect(&barF[ps * NR + NW], lane);
                }
#pragma unroll
                for (int k = 0; k < NF; ++k) { tU[k] = cT[k]; tFy[k] = cFy[k]; }
                tly = cly;
            } else {
                // top flux of the last plane.  Warp NW-1 publishes record it+1 on this barrier only at the top of
                // its iteration it+1, after its iteration `it` has taken the flux written here; the slot held
                // flux it-3, which warp NW-1 took during its plane it-2, before it published record it-1.
                const int ps = (it - 1) & 1;
                mbar_wait(&barD[ps * NR + NW - 1], (unsigned) (((it - 1) >> 1) & 1));
                const double *d_dn = sm_d + ps * DS + (NW - 1) * 11 * 32 + lane;
                double lU[NF], lF[NF], AFy[NF];
#pragma unroll
                for (int k = 0; k < NF; ++k) { lU[k] = d_dn[k * 32]; lF[k] = d_dn[(NF + k) * 32]; }
                const double ll  = d_dn[10 * 32];
                const double lam = llf_area_flux(lU, lF, ll, tU, tFy, tly, Ah, AFy);
                lmy = (lam < lmy) ? lmy : lam;
                double *f = sm_f + ps * FS + NW * NF * 32 + lane;
#pragma unroll
                for (int k = 0; k < NF; ++k) f[k * 32] = AFy[k];
                mbar_arrive_elect(&barF[ps * NR + NW], lane);
            }
        }
        lmax = yf_ok ? lmy : 0.0;
    } else {
        // ================= update warps: rows jA (cell 0) and jA + 1 (cell 1) ===========================
        const int jA = j0 + 2 * (w - 1);
        int  jc[2], key_y[2];
        bool in_y[2], upd[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int j = jA + c;
            jc[c]    = min(max(j, lc.jlo), lc.jhi);
            in_y[c]  = (j >= 0 && j < g.ny);
            upd[c]   = lane >= 1 && lane <= XW && in_x && in_y[c];
            key_y[c] = order_key<ORDER>(g.gy0 + j, 1);
        }
        const double dt = (STAGE >= 1) ? ctl->dt : 0.0;
        const int key_x = order_key<ORDER>(g.gx0 + i, 0);
        float est_max = 0.f;

        double *d_own = sm_d + w * 11 * 32 + lane;              // + slot * DS   (record of cell 1)
        const double *d_dn = sm_d + (w - 1) * 11 * 32 + lane;   //               (record below cell 0)
        double *f_own = sm_f + w * NF * 32 + lane;              // + slot * FS   (low y face of cell 0)
        const double *f_up = sm_f + (w + 1) * NF * 32 + lane;   //               (-y_hi of cell 1)

        // cell 1's columns relative to cell 0's: one row up unless the load clamp folds them together
        const int drow_s = (jc[1] - jc[0]) * srow;               // residual input
        const int drow   = (jc[1] - jc[0]) * g.px;               // own cells: U^n loads and stores
        const long long colA = (long long) (jc[0] + 1) * g.px + (ic + 1);
        const double *scolA = sbase + (long long) (jc[0] + 1) * srow;
        const double *sp  = scolA + (long long) (max(z0 - 1, lc.klo) + 1) * splane; // plane z0-1 (clamped)
        const double *unp = Un + colA + (long long) z0 * plane;  // plane z0-1: U^n of a plane is loaded in the iteration that finishes it
        double *op = Out + colA + (long long) z0 * plane;        // plane z0-1 (the first store goes to plane z0)

        double pU[2][NF], pFz[2][NF], plz[2], pS[2][NF], pUn[2][NF], nxt[2][NF];
        double lmx[2] = { 0.0, 0.0 }, lmy[2] = { 0.0, 0.0 }, lmz[2] = { 0.0, 0.0 };
        // ---- prologue: plane z0-1 only provides the low side of the first z interfaces ----------------
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int k = 0; k < NF; ++k) pU[c][k] = ldsin(sp + c * drow_s + k * sfs);
        }
        sp = scolA + (long long) (z0 + 1) * splane; // plane z0
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int k = 0; k < NF; ++k) nxt[c][k] = ldsin(sp + c * drow_s + k * sfs);
            CellPrim q;
            derive_cell(pU[c], dc, q);
            axis_flux<2>(q, pFz[c], plz[c]);
#pragma unroll
            for (int k = 0; k < NF; ++k) { pS[c][k] = 0.0; pUn[c][k] = 0.0; }
        }

        auto body = [&](auto slot_tag, const int it) {
            constexpr int slot = decltype(slot_tag)::value;
            const int kz = z0 + it;
            const unsigned par = (unsigned) ((it >> 1) & 1);
            double cU[2][NF];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int k = 0; k < NF; ++k) cU[c][k] = nxt[c][k];
            }
            if (kz + 1 <= lc.khi) sp += splane; // plane kz+1 (the ghost plane nz, or plane nz-1 again on a free-flow side)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int k = 0; k < NF; ++k) nxt[c][k] = ldsin(sp + c * drow_s + k * sfs);
                if (STAGE >= 2 && upd[c] && it > 0) {
#pragma unroll
                    for (int k = 0; k < NF; ++k) pUn[c][k] = unp[c * drow + k * fs];
                }
            }
            unp += plane;

            CellPrim q[2];
            double cFy[2][NF], cly[2];
            derive_cell2(cU[0], cU[1], dc, q[0], q[1]);
#pragma unroll
            for (int c = 0; c < 2; ++c) axis_flux<1>(q[c], cFy[c], cly[c]);
            // ---- record of cell 1 for the warp above.  The slot held record it-2; that warp read it before it
            //      published flux it-2, and this warp took that flux during plane it-1 (or it-2): free. ---------
            {
                double *d = d_own + slot * DS;
#pragma unroll
                for (int k = 0; k < NF; ++k) { d[k * 32] = cU[1][k]; d[(NF + k) * 32] = cFy[1][k]; }
                d[10 * 32] = cly[1];
            }
            mbar_arrive_elect(&barD[slot * NR + w], lane);

            // ---- z interfaces (kz-1 | kz) ---------------------------------------------------------------------
            double cFz[2][NF], clz[2], AFz[2][NF];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                axis_flux<2>(q[c], cFz[c], clz[c]);
                const double lam = llf_area_flux(pU[c], pFz[c], plz[c], cU[c], cFz[c], clz[c], Ah, AFz[c]);
                lmz[c] = (lam < lmz[c]) ? lmz[c] : lam;
            }
            // ---- plane kz-1 of cell 1: -y_hi (low y face of the warp above, published one plane ago) ----------
            // that warp completes this barrier again (flux it+1) only after it has seen record it+1 of this
            // warp, which is published in the next iteration, after this wait.
            if (ORDER != NUM_AXIS && it > 0 && g.gz0 + kz - 1 != 0) {
                mbar_wait(&barF[(slot ^ 1) * NR + w + 1], (unsigned) (((it - 1) >> 1) & 1));
                const double *f = f_up + (slot ^ 1) * FS;
#pragma unroll
                for (int k = 0; k < NF; ++k) pS[1][k] -= f[k * 32];
            }
            // ---- -z_hi, RK update and store of plane kz-1 -----------------------------------------------------
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                finish_plane<STAGE>(pS[c], AFz[c], pU[c], pUn[c], dt, g.volume, dc.y_vol, op + c * drow, fs, upd[c] && it > 0, est_max);
            }
            op += plane;

            // ---- x interfaces (i-1 | i): lane-1's states by warp shuffle ---------------------------------------
            double AFx[2][NF];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                double cFx[NF], clx, lU[NF], lF[NF];
                axis_flux<0>(q[c], cFx, clx);
#pragma unroll
                for (int k = 0; k < NF; ++k) { lU[k] = shfl_up_d(cU[c][k]); lF[k] = shfl_up_d(cFx[k]); }
                const double ll  = shfl_up_d(clx);
                const double lam = llf_area_flux(lU, lF, ll, cU[c], cFx, clx, Ah, AFx[c]);
                lmx[c] = (lam < lmx[c]) ? lmx[c] : lam;
            }

            // ---- y interfaces: (cell 0 | cell 1) inside the thread, (row below | cell 0) through shared memory
            double AFy[2][NF];
            {
                const double lam = llf_area_flux(cU[0], cFy[0], cly[0], cU[1], cFy[1], cly[1], Ah, AFy[1]);
                lmy[1] = (lam < lmy[1]) ? lmy[1] : lam;
            }
            // the warp below publishes record it+2 on this barrier only after it has taken this warp's flux `it`
            mbar_wait(&barD[slot * NR + w - 1], par);
            {
                const double *d = d_dn + slot * DS;
                double lU[NF], lF[NF];
#pragma unroll
                for (int k = 0; k < NF; ++k) { lU[k] = d[k * 32]; lF[k] = d[(NF + k) * 32]; }
                const double ll  = d[10 * 32];
                const double lam = llf_area_flux(lU, lF, ll, cU[0], cFy[0], cly[0], Ah, AFy[0]);
                lmy[0] = (lam < lmy[0]) ? lmy[0] : lam;
                // the slot held flux it-2: the warp below took it during its plane it-1, before it published the
                // record `it` this warp has just waited for
                double *f = f_own + slot * FS;
#pragma unroll
                for (int k = 0; k < NF; ++k) f[k * 32] = AFy[0][k];
                mbar_arrive_elect(&barF[slot * NR + w], lane);
            }

            // ---- ordered accumulation of plane kz (its registers are free: plane kz-1 is finished) ------------
            const int key_z = order_key<ORDER>(g.gz0 + kz, 2);
            // cell 0: its -y_hi is the thread's own face (cell 0 | cell 1): complete up to -z_hi
            lows_minus_xhi<ORDER>(AFx[0], AFy[0], AFz[0], key_x, key_y[0], key_z, pS[0]);
#pragma unroll
            for (int k = 0; k < NF; ++k) pS[0][k] -= AFy[1][k];
            if (ORDER == NUM_AXIS || key_z < 0) {
#pragma unroll
                for (int k = 0; k < NF; ++k) pS[0][k] += AFz[0][k];
            }
            // cell 1: -y_hi follows one plane later, except where a +z_lo term has to come behind it
            lows_minus_xhi<ORDER>(AFx[1], AFy[1], AFz[1], key_x, key_y[1], key_z, pS[1]);
            if (ORDER == NUM_AXIS || key_z < 0) {
                mbar_wait(&barF[slot * NR + w + 1], par);
                const double *f = f_up + slot * FS;
#pragma unroll
                for (int k = 0; k < NF; ++k) pS[1][k] -= f[k * 32];
#pragma unroll
                for (int k = 0; k < NF; ++k) pS[1][k] += AFz[1][k];
            }

            // ---- plane kz becomes the previous plane -----------------------------------------------------------
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int k = 0; k < NF; ++k) { pU[c][k] = cU[c][k]; pFz[c][k] = cFz[c][k]; }
                plz[c] = clz[c];
            }
        };
        {
            // two planes per trip, no test between them: ptxas renames the rotating state instead of copying it
            int it = 0;
#pragma unroll 1
            for (; it + 1 < n; it += 2) {
                body(std::integral_constant<int, 0>{}, it);
                body(std::integral_constant<int, 1>{}, it + 1);
            }
            if (it < n) body(std::integral_constant<int, 0>{}, it);
        }

        // ---- epilogue: plane z1 only closes the last z interfaces; cell 1 of plane z1-1 still lacks -y_hi ----
        {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (STAGE >= 2 && upd[c]) {
#pragma unroll
                    for (int k = 0; k < NF; ++k) pUn[c][k] = unp[c * drow + k * fs];
                }
            }
            if (ORDER != NUM_AXIS && g.gz0 + z1 - 1 != 0) {
                mbar_wait(&barF[((n - 1) & 1) * NR + w + 1], (unsigned) (((n - 1) >> 1) & 1));
                const double *f = f_up + ((n - 1) & 1) * FS;
#pragma unroll
                for (int k = 0; k < NF; ++k) pS[1][k] -= f[k * 32];
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                CellPrim q;
                derive_cell(nxt[c], dc, q);
                double cFz[NF], clz, AFz[NF];
                axis_flux<2>(q, cFz, clz);
                const double lam = llf_area_flux(pU[c], pFz[c], plz[c], nxt[c], cFz, clz, Ah, AFz);
                lmz[c] = (lam < lmz[c]) ? lmz[c] : lam;
                finish_plane<STAGE>(pS[c], AFz, pU[c], pUn[c], dt, g.volume, dc.y_vol, op + c * drow, fs, upd[c], est_max);
            }
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int j = jA + c;
            const bool xf_ok = in_y[c] && lane >= 1 && i >= 0 && i <= g.nx;                // face (i-1 | i)
            const bool yf_ok = in_x && lane >= 1 && lane <= XW && j >= 0 && j <= g.ny;     // face (j-1 | j)
            const bool zf_ok = in_x && in_y[c];                                            // face (k-1 | k)
            if (xf_ok) lmax = (lmx[c] < lmax) ? lmax : lmx[c];
            if (yf_ok) lmax = (lmy[c] < lmax) ? lmax : lmy[c];
            if (zf_ok) lmax = (lmz[c] < lmax) ? lmax : lmz[c];
        }
        emax = est_max;
    }

    block_maxima<NW>(lmax, max_eig, emax, (STAGE == 3) ? cta_est : nullptr, tid.tile, smem);
}

} // namespace mmf
