// uniform_stage_v5r.cuh -- the "rotate" form of the low-face streaming stage kernel
// (uniform_stage_v5.cuh explains the scheme): one loop body per plane, the current plane's state is
// copied into the previous plane's registers at the end of the iteration, U^n is loaded one plane
// ahead.  Same arithmetic, different register allocation / instruction schedule: measured on B200
// (profiles/), this form is the faster one for stages 2 and 3 at 12 warps (164-166 registers, no
// spills, 0.54 / 0.55 ms at 256^3), the ping-pong form of uniform_stage_v5.cuh for stage 1 at 16
// warps (128 registers, 0.50 ms).  uniform_path.cuh picks per stage.
#pragma once

#include "uniform_stage_v5.cuh"

namespace mmf {

// planes per trip of the steady-state loop: 2 lets ptxas rename instead of copying part of the rotated
// state and overlap the tail of one plane with the head of the next (measured: stage 1 0.65 -> 0.50 ms,
// stages 2/3 -1..3 %)
#ifndef MMF_R_UNROLL
#define MMF_R_UNROLL 2
#endif
constexpr int R_UNROLL = MMF_R_UNROLL;

__host__ __device__ constexpr size_t stage_v5_smem_bytes(int nw)
{
    return (size_t) nw * 16 * 32 * sizeof(double) + 2 * nw * sizeof(unsigned long long);
}

template <int STAGE, int ORDER, int NW, bool XG>
__device__ __forceinline__ void
stage_v5r_body(const UniformGeom &g, const double *__restrict__ Sin, const double *Un, double *Out,
               const StepControl *__restrict__ ctl, double *__restrict__ max_eig, const int lz,
               float *__restrict__ cta_est, const LoadClamp &lc, const HaloWait &hw, const XGhost &xg)
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
    const long long col   = (long long) (jc + 1) * g.px + (ic + XOFF);
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
        const bool upd   = lane >= 1 && lane <= XW && in_x && in_y;
        const bool xf_ok = in_y && lane >= 1 && i >= 0 && i <= g.nx;                  // face (i-1 | i)
        const bool yf_ok = in_x && lane >= 1 && lane <= XW && j >= 0 && j <= g.ny;   // face (j-1 | j)
        const bool zf_ok = in_x && in_y;                                              // face (k-1 | k)
        const double dt = (STAGE >= 1) ? ctl->dt : 0.0;
        const int key_x = order_key<ORDER>(g.gx0 + i, 0);
        const int key_y = order_key<ORDER>(g.gy0 + j, 1);
        float est_max = 0.f;

        double *d_own = sm_d + row * 11 * 32 + lane;
        // (MMF_EXP_NOSYNC, ptx_helpers.cuh; experiment builds: a row takes its OWN record and flux for the neighbours', so that the
        //  numbers stay sane while nothing waits -- timing experiment, wrong results)
        const double *d_dn = sm_d + (row - (MMF_EXP_NOSYNC ? 0 : 1)) * 11 * 32 + lane;
        double *f_own = sm_f + row * NF * 32 + lane;
        const double *f_up = sm_f + (row + (MMF_EXP_NOSYNC ? 0 : 1)) * NF * 32 + lane;

        const double *sp  = scol + (long long) (max(z0 - 1, lc.klo) + 1) * splane; // plane z0-1 (clamped)
        const double *unp = Un + col + (long long) (z0 + 1) * splane; // plane z0
        double *op = Out + col + (long long) z0 * plane;        // plane z0-1 (first store goes to plane z0)

        double pU[NF], pFz[NF], plz, pS[NF], pUn[NF], nxt[NF];
        double lmx = 0.0, lmy = 0.0, lmz = 0.0;
        // ---- prologue: plane z0-1 only provides the low side of the first z interface --------------
        {
#pragma unroll
            for (int k = 0; k < NF; ++k) pU[k] = ldsin(sp + k * sfs);
            sp = scol + (long long) (z0 + 1) * splane; // plane z0
#pragma unroll
            for (int k = 0; k < NF; ++k) nxt[k] = ldsin(sp + k * sfs);
            CellPrim q;
            derive_cell(pU, dc, q);
            axis_flux<2>(q, pFz, plz);
#pragma unroll
            for (int k = 0; k < NF; ++k) { pS[k] = 0.0; pUn[k] = 0.0; }
        }

#pragma unroll R_UNROLL
        for (int kz = z0; kz < z1; ++kz) {
            const unsigned par = (unsigned) ((kz - z0) & 1);
            double cU[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) cU[k] = nxt[k];
            if (kz + 1 <= lc.khi) sp += splane; // plane kz+1 (the ghost plane nz, or plane nz-1 again on a free-flow side)
#pragma unroll
            for (int k = 0; k < NF; ++k) nxt[k] = ldsin(sp + k * sfs);
            double cUn[NF];
            if (STAGE >= 2 && upd) {
#pragma unroll
                for (int k = 0; k < NF; ++k) cUn[k] = unp[k * fs];
            }
            unp += plane;

            CellPrim q;
            derive_cell(cU, dc, q);

            // ---- y record for row+1 (the earlier it is out, the less row+1 waits) ---------------------
            double cFy[NF], cly;
            axis_flux<1>(q, cFy, cly);
#pragma unroll
            for (int k = 0; k < NF; ++k) { d_own[k * 32] = cU[k]; d_own[(NF + k) * 32] = cFy[k]; }
            d_own[10 * 32] = cly;
            mbar_arrive_elect(&barD[row], lane);

            // ---- z interface (kz-1 | kz): completes plane kz-1 ----------------------------------------
            double cFz[NF], clz, AFz[NF];
            axis_flux<2>(q, cFz, clz);
            {
                const double lam = llf_area_flux(pU, pFz, plz, cU, cFz, clz, Ah, AFz);
                lmz = (lam < lmz) ? lmz : lam;
            }
            finish_plane<STAGE>(pS, AFz, pU, pUn, dt, g.volume, dc.y_vol, op, fs, upd && kz > z0, est_max);
            op += plane;

            // ---- x interface (i-1 | i): lane-1's state by warp shuffle ---------------------------------
            double AFx[NF];
            {
                double cFx[NF], clx, lU[NF], lF[NF];
                axis_flux<0>(q, cFx, clx);
#pragma unroll
                for (int k = 0; k < NF; ++k) { lU[k] = shfl_up_d(cU[k]); lF[k] = shfl_up_d(cFx[k]); }
                const double ll  = shfl_up_d(clx);
                const double lam = llf_area_flux(lU, lF, ll, cU, cFx, clx, Ah, AFx);
                lmx = (lam < lmx) ? lmx : lam;
            }

            // ---- y interface (j-1 | j): row-1's record through shared memory --------------------------
            double AFy[NF];
            mbar_wait(&barD[row - 1], par);
            {
                double lU[NF], lF[NF];
#pragma unroll
                for (int k = 0; k < NF; ++k) { lU[k] = d_dn[k * 32]; lF[k] = d_dn[(NF + k) * 32]; }
                const double ll  = d_dn[10 * 32];
                const double lam = llf_area_flux(lU, lF, ll, cU, cFy, cly, Ah, AFy);
                lmy = (lam < lmy) ? lmy : lam;
                // row-1 published record `it` only after it had read this row's flux `it-1`
#pragma unroll
                for (int k = 0; k < NF; ++k) f_own[k * 32] = AFy[k];
                mbar_arrive_elect(&barF[row], lane);
            }

            // ---- ordered accumulation (src/euler.cpp:153, 237-247) ------------------------------------
            // interior low faces first, sorted by their creator (largest key first; a+b commutes, so
            // only the LAST one matters); then the cell's own faces in the order it created them:
            // (-x if border) +x (-y if border) +y (-z if border) +z; low faces `+=`, high faces `-=`.
            const int key_z = order_key<ORDER>(g.gz0 + kz, 2);
            double S[NF];
            const bool edge = (key_y < 0) | (key_z < 0); // warp-uniform: one row, one plane per warp
            if (ORDER == NUM_AXIS) {
#pragma unroll
                for (int k = 0; k < NF; ++k) S[k] = 0.0 + AFx[k];
            } else if (!edge) {
                // a border low face in x alone is simply "last" (key -1), directly followed by -x_hi
                const int last = (key_x < key_y) ? ((key_x < key_z) ? 0 : 2) : ((key_y < key_z) ? 1 : 2);
#pragma unroll
                for (int k = 0; k < NF; ++k) {
                    const double p = (last == 0) ? AFy[k] : AFx[k];
                    const double t = (last == 2) ? AFy[k] : AFz[k];
                    const double r = (last == 0) ? AFx[k] : (last == 1) ? AFy[k] : AFz[k];
                    S[k] = (p + t) + r;
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
                    S[k] = s;
                }
            }
            // -x_hi: the low x face of lane+1
#pragma unroll
            for (int k = 0; k < NF; ++k) S[k] -= shfl_down_d(AFx[k]);
            if (ORDER == NUM_AXIS) {
#pragma unroll
                for (int k = 0; k < NF; ++k) S[k] += AFy[k];
            } else if (edge && key_y < 0) {
#pragma unroll
                for (int k = 0; k < NF; ++k) S[k] += AFy[k];
            }
            // -y_hi: the low y face of row+1
            mbar_wait(&barF[row + 1], par);
#pragma unroll
            for (int k = 0; k < NF; ++k) S[k] -= f_up[k * 32];
            if (ORDER == NUM_AXIS) {
#pragma unroll
                for (int k = 0; k < NF; ++k) S[k] += AFz[k];
            } else if (edge && key_z < 0) {
#pragma unroll
                for (int k = 0; k < NF; ++k) S[k] += AFz[k];
            }

            // ---- plane kz becomes the previous plane; -z_hi follows in the next iteration -------------
#pragma unroll
            for (int k = 0; k < NF; ++k) { pS[k] = S[k]; pU[k] = cU[k]; pFz[k] = cFz[k]; }
            plz = clz;
            if (STAGE >= 2) {
#pragma unroll
                for (int k = 0; k < NF; ++k) pUn[k] = cUn[k];
            }
        }

        // ---- epilogue: plane z1 only closes the last z interface -----------------------------------
        {
            CellPrim q;
            derive_cell(nxt, dc, q);
            double cFz[NF], clz, AFz[NF];
            axis_flux<2>(q, cFz, clz);
            const double lam = llf_area_flux(pU, pFz, plz, nxt, cFz, clz, Ah, AFz);
            lmz = (lam < lmz) ? lmz : lam;
            finish_plane<STAGE>(pS, AFz, pU, pUn, dt, g.volume, dc.y_vol, op, fs, upd, est_max);
        }
        lmax = xf_ok ? lmx : 0.0;
        if (yf_ok) lmax = (lmy < lmax) ? lmax : lmy;
        if (zf_ok) lmax = (lmz < lmax) ? lmax : lmz;
        emax = est_max;
    }

    block_maxima<NW>(lmax, max_eig, emax, (STAGE == 3) ? cta_est : nullptr, tid.tile, smem);
}

template <int STAGE, int ORDER, int NW, bool XG>
__global__ void __maxnreg__(stage_regs(NW))
uniform_stage_kernel_v5r(const UniformGeom g, const double *__restrict__ Sin, const double *Un, double *Out,
                         const StepControl *__restrict__ ctl, double *__restrict__ max_eig, const int lz,
                         float *__restrict__ cta_est, const LoadClamp lc, const HaloWait hw, const XGhost xg)
{
    stage_v5r_body<STAGE, ORDER, NW, XG>(g, Sin, Un, Out, ctl, max_eig, lz, cta_est, lc, hw, xg);
}

} // namespace mmf
