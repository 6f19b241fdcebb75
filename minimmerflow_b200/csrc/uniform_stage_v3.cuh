// uniform_stage_v3.cuh -- fused residual + RK-stage kernel, third generation.
//
// Same tiling and arithmetic as uniform_stage_kernel (uniform_kernels.cuh): a warp owns a 32-cell
// x window (30 updated, +-x neighbours by warp shuffle), the CTA's NW warps are consecutive y rows
// (rows 0 and NW-1 are halo rows), the CTA marches along z with plane k-1 in registers.
//
// What changed, driven by the ncu captures in profiles/ (20 % of warp stalls on the two CTA-wide
// barriers per plane, which also phase-lock all warps into the same low-ILP phase):
//   * rows synchronise PAIRWISE through shared-memory mbarriers -- row r only ever waits for rows
//     r-1 and r+1 -- and split-phase: a row arrives, does independent work (finishing plane k-1,
//     the x faces), and only then waits;
//   * the two halo rows run their own lean loops (state, y flux, nothing else);
//   * masks for the max-eigenvalue reduction are loop-invariant and applied once at the end.
// Bit-identical to v1 and to the CPU restatement of the reference (tests/test_uniform_gpu.py runs
// every kernel version).
#pragma once

#include "uniform_device.cuh"

namespace mmf {

template <int STAGE, int ORDER, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
uniform_stage_kernel_v3(const UniformGeom g, const double *__restrict__ Sin, const double *Un, double *Out,
                        const StepControl *__restrict__ ctl, double *__restrict__ max_eig, const int lz)
{
    extern __shared__ double smem[];
    // sm_d[row][q][lane], q = U0..U4, Fy0..Fy4, lam_y ; sm_f[row][k][lane] = area * y-flux
    double *sm_d = smem;
    double *sm_f = smem + NW * 11 * 32;
    unsigned long long *barD = reinterpret_cast<unsigned long long *>(sm_f + NW * NF * 32); // record of row r published
    unsigned long long *barF = barD + NW;                                                    // y flux of row r published

    if (STAGE >= 1 && ctl->active == 0.0) return;

    const int lane = threadIdx.x & 31;
    const int row  = threadIdx.x >> 5;
    if (threadIdx.x < NW) {
        mbar_init(&barD[threadIdx.x], 32);
        mbar_init(&barF[threadIdx.x], 32);
    }
    __syncthreads();

    const int i  = blockIdx.x * XW - 1 + lane;
    const int j  = blockIdx.y * (NW - 2) - 1 + row;
    const int z0 = blockIdx.z * lz;
    const int z1 = min(z0 + lz, g.nz);
    const int ic = min(max(i, -1), g.nx);
    const int jc = min(max(j, -1), g.ny);
    const bool in_x = (i >= 0 && i < g.nx);
    const bool in_y = (j >= 0 && j < g.ny);

    const double Ah = 0.5 * g.area;
    DivConsts dc;
    dc.y_gm1 = rcp_nr(GM1);
    dc.y_c1  = rcp_nr(TWO_OVER_GM1);
    dc.y_vol = rcp_nr(g.volume);

    const long long plane = (long long) g.py * g.px;
    const long long col   = (long long) (jc + 1) * g.px + (ic + 1);
    const long long fs    = g.fs;
    double lmax = 0.0;

    if (row == NW - 1) {
        // ================= high halo row: publishes (U, Fy, lam_y) of row j for row NW-2 ============
        double *d = sm_d + row * 11 * 32 + lane;
        for (int kz = z0; kz < z1; ++kz) {
            const int it = kz - z0;
            const double *sp = Sin + col + (long long) (kz + 1) * plane;
            double cU[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) cU[k] = sp[k * fs];
            CellPrim q;
            derive_cell(cU, dc, q);
            double cFy[NF], cly;
            axis_flux<1>(q, cFy, cly);
            if (it > 0) mbar_wait(&barF[NW - 2], (unsigned) ((it - 1) & 1)); // row NW-2 is done with the previous record
#pragma unroll
            for (int k = 0; k < NF; ++k) { d[k * 32] = cU[k]; d[(NF + k) * 32] = cFy[k]; }
            d[10 * 32] = cly;
            mbar_arrive(&barD[row]);
        }
    } else if (row == 0) {
        // ================= low halo row: computes the y face (j | j+1) for row 1 ===================
        const bool yf_ok = in_x && lane >= 1 && lane <= XW && j >= -1 && j < g.ny;
        double lmy = 0.0;
        for (int kz = z0; kz < z1; ++kz) {
            const int it = kz - z0;
            const double *sp = Sin + col + (long long) (kz + 1) * plane;
            double cU[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) cU[k] = sp[k * fs];
            CellPrim q;
            derive_cell(cU, dc, q);
            double cFy[NF], cly;
            axis_flux<1>(q, cFy, cly);
            mbar_wait(&barD[1], (unsigned) (it & 1));
            const double *d = sm_d + 1 * 11 * 32 + lane;
            double nU[NF], nF[NF], AFyhi[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) { nU[k] = d[k * 32]; nF[k] = d[(NF + k) * 32]; }
            const double nl  = d[10 * 32];
            const double lam = llf_area_flux(cU, cFy, cly, nU, nF, nl, Ah, AFyhi);
            lmy = (lam < lmy) ? lmy : lam;
            // row 1 published record `it` only after it had read flux `it-1`: f[0] is free
            double *f = sm_f + lane;
#pragma unroll
            for (int k = 0; k < NF; ++k) f[k * 32] = AFyhi[k];
            mbar_arrive(&barF[0]);
        }
        lmax = yf_ok ? lmy : 0.0;
    } else {
        // ================= update rows ==============================================================
        const bool upd   = lane >= 1 && lane <= XW && in_x && in_y;
        const bool xf_ok = in_y && lane <= XW && i >= -1 && i < g.nx; // face (i | i+1)
        const bool yf_ok = in_x && lane >= 1 && lane <= XW && j >= -1 && j < g.ny;
        const bool zf_ok = in_x && in_y;
        const double dt = (STAGE >= 1) ? ctl->dt : 0.0;
        const int gi = g.gx0 + i, gj = g.gy0 + j;
        const bool blo_x = (gi == 0), blo_y = (gj == 0);
        const int key_x = blo_x ? -1 : 3 * (__ffs(gi) - 1);
        const int key_y = blo_y ? -1 : 3 * (__ffs(gj) - 1) + 1;

        const double *sp = Sin + col + (long long) z0 * plane; // plane z0-1
        double nxt[NF];
#pragma unroll
        for (int k = 0; k < NF; ++k) nxt[k] = sp[k * fs];

        double pU[NF], pFz[NF], plz = 0.0, S[NF], pUn[NF];
        double lmx = 0.0, lmy = 0.0, lmz = 0.0;
#pragma unroll
        for (int k = 0; k < NF; ++k) { pU[k] = 0.0; pFz[k] = 0.0; S[k] = 0.0; pUn[k] = 0.0; }

        double *d_own = sm_d + row * 11 * 32 + lane;
        const double *d_up = sm_d + (row + 1) * 11 * 32 + lane;
        double *f_own = sm_f + row * NF * 32 + lane;
        const double *f_dn = sm_f + (row - 1) * NF * 32 + lane;

        for (int kz = z0 - 1; kz <= z1; ++kz) {
            const int it = kz - z0;
            double cU[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) cU[k] = nxt[k];
            if (kz < z1) {
                const double *np = Sin + col + (long long) (kz + 2) * plane;
#pragma unroll
                for (int k = 0; k < NF; ++k) nxt[k] = np[k * fs];
            }
            double cUn[NF];
            if (STAGE >= 2 && upd && kz >= z0 && kz < z1) {
                const double *up = Un + col + (long long) (kz + 1) * plane;
#pragma unroll
                for (int k = 0; k < NF; ++k) cUn[k] = up[k * fs];
            }

            CellPrim q;
            derive_cell(cU, dc, q);

            // ---- z interface (kz-1 | kz) -----------------------------------------------------------
            double cFz[NF], clz, AFz[NF];
            axis_flux<2>(q, cFz, clz);
            if (kz >= z0) {
                const double lam = llf_area_flux(pU, pFz, plz, cU, cFz, clz, Ah, AFz);
                lmz = (lam < lmz) ? lmz : lam;
            }

            // ---- publish (U, Fy, lam_y) for row-1; row-1 finished with the previous record before
            //      it published the flux this row waited for at the end of the previous plane --------
            double cFy[NF], cly;
            if (kz < z1) {
                axis_flux<1>(q, cFy, cly);
                if (kz >= z0) {
#pragma unroll
                    for (int k = 0; k < NF; ++k) { d_own[k * 32] = cU[k]; d_own[(NF + k) * 32] = cFy[k]; }
                    d_own[10 * 32] = cly;
                    mbar_arrive(&barD[row]);
                }
            }

            // ---- finish cell (i,j,kz-1) while the neighbour rows catch up -----------------------------
            if (upd && kz > z0) {
                double *op = Out + col + (long long) kz * plane; // plane kz-1
#pragma unroll
                for (int k = 0; k < NF; ++k) {
                    const double rhs = S[k] - AFz[k];
                    double out;
                    if (STAGE == 0) {
                        out = rhs;
                    } else {
                        const double dq = div_nr(dt * rhs, g.volume, dc.y_vol); // dt * RHS[k] / cellVolume
                        if (STAGE == 1)      out = pU[k] + dq;
                        else if (STAGE == 2) out = 0.75 * pUn[k] + 0.25 * (pU[k] + dq);
                        else                 out = (1. / 3) * pUn[k] + (2. / 3) * (pU[k] + dq);
                    }
                    op[k * fs] = out;
                }
            }
            if (kz == z1) break;

            if (kz >= z0) {
                // ---- y face (j | j+1): needs row+1's record ------------------------------------------
                double AFyhi[NF], AFylo[NF];
                mbar_wait(&barD[row + 1], (unsigned) (it & 1));
                {
                    double nU[NF], nF[NF];
#pragma unroll
                    for (int k = 0; k < NF; ++k) { nU[k] = d_up[k * 32]; nF[k] = d_up[(NF + k) * 32]; }
                    const double nl  = d_up[10 * 32];
                    const double lam = llf_area_flux(cU, cFy, cly, nU, nF, nl, Ah, AFyhi);
                    lmy = (lam < lmy) ? lmy : lam;
                    // row+1 published record `it` only after reading this row's flux `it-1`
#pragma unroll
                    for (int k = 0; k < NF; ++k) f_own[k * 32] = AFyhi[k];
                    mbar_arrive(&barF[row]);
                }

                // ---- x faces by warp shuffle while row-1 finishes its y face ---------------------------
                double cFx[NF], clx, AFxhi[NF], AFxlo[NF];
                axis_flux<0>(q, cFx, clx);
                {
                    double nU[NF], nF[NF];
#pragma unroll
                    for (int k = 0; k < NF; ++k) { nU[k] = shfl_down_d(cU[k]); nF[k] = shfl_down_d(cFx[k]); }
                    const double nl  = shfl_down_d(clx);
                    const double lam = llf_area_flux(cU, cFx, clx, nU, nF, nl, Ah, AFxhi);
                    lmx = (lam < lmx) ? lmx : lam;
#pragma unroll
                    for (int k = 0; k < NF; ++k) AFxlo[k] = shfl_up_d(AFxhi[k]);
                }

                mbar_wait(&barF[row - 1], (unsigned) (it & 1));
#pragma unroll
                for (int k = 0; k < NF; ++k) AFylo[k] = f_dn[k * 32];

                // ---- ordered accumulation (src/euler.cpp:153, 237-247; derivation in uniform_kernels.cuh)
                const int gk = g.gz0 + kz;
                const bool blo_z = (gk == 0);
                if (ORDER == NUM_AXIS) {
#pragma unroll
                    for (int k = 0; k < NF; ++k) S[k] = ((((0.0 + AFxlo[k]) - AFxhi[k]) + AFylo[k]) - AFyhi[k]) + AFz[k];
                } else if (blo_x | blo_y | blo_z) {
                    int kx = key_x, ky = key_y, kzz = blo_z ? -1 : 3 * (__ffs(gk) - 1) + 2;
                    if (ORDER == NUM_LEXI) { kx = blo_x ? -1 : 0; ky = blo_y ? -1 : 1; kzz = blo_z ? -1 : 2; }
                    const int last = (kx < ky) ? ((kx < kzz) ? 0 : 2) : ((ky < kzz) ? 1 : 2);
#pragma unroll
                    for (int k = 0; k < NF; ++k) {
                        const double cx = blo_x ? 0.0 : AFxlo[k];
                        const double cy = blo_y ? 0.0 : AFylo[k];
                        const double cz = blo_z ? 0.0 : AFz[k];
                        const double p = (last == 0) ? cy : cx;
                        const double r = (last == 0) ? cx : (last == 1) ? cy : cz;
                        const double t = (last == 2) ? cy : cz;
                        double s = ((0.0 + p) + t) + r;
                        if (blo_x) s += AFxlo[k];
                        s -= AFxhi[k];
                        if (blo_y) s += AFylo[k];
                        s -= AFyhi[k];
                        if (blo_z) s += AFz[k];
                        S[k] = s;
                    }
                } else if (ORDER == NUM_LEXI) {
#pragma unroll
                    for (int k = 0; k < NF; ++k) S[k] = (((AFz[k] + AFylo[k]) + AFxlo[k]) - AFxhi[k]) - AFyhi[k];
                } else {
                    const int key_z = 3 * (__ffs(gk) - 1) + 2;
                    if (key_y < key_z) {
                        const bool xl = key_x < key_y;
#pragma unroll
                        for (int k = 0; k < NF; ++k) {
                            const double a = xl ? AFylo[k] : AFxlo[k];
                            const double r = xl ? AFxlo[k] : AFylo[k];
                            S[k] = (((a + AFz[k]) + r) - AFxhi[k]) - AFyhi[k];
                        }
                    } else {
                        const bool xl = key_x < key_z;
#pragma unroll
                        for (int k = 0; k < NF; ++k) {
                            const double a = xl ? AFz[k] : AFxlo[k];
                            const double r = xl ? AFxlo[k] : AFz[k];
                            S[k] = (((a + AFylo[k]) + r) - AFxhi[k]) - AFyhi[k];
                        }
                    }
                }
            }

#pragma unroll
            for (int k = 0; k < NF; ++k) { pU[k] = cU[k]; pFz[k] = cFz[k]; }
            plz = clz;
            if (STAGE >= 2) {
#pragma unroll
                for (int k = 0; k < NF; ++k) pUn[k] = cUn[k];
            }
        }
        lmax = xf_ok ? lmx : 0.0;
        if (yf_ok) lmax = (lmy < lmax) ? lmax : lmy;
        if (zf_ok) lmax = (lmz < lmax) ? lmax : lmz;
    }

    // ---- max eigenvalue: warp shuffle, block reduction, one atomic per CTA ----------------------
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, lmax, o);
        lmax = (lmax < other) ? other : lmax;
    }
    if (lane == 0) smem[row] = lmax;
    __syncthreads();
    if (row == 0) {
        double v = (lane < NW) ? smem[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, v, o);
            v = (v < other) ? other : v;
        }
        if (lane == 0) atomic_max_nonneg(max_eig, v);
    }
}

} // namespace mmf
