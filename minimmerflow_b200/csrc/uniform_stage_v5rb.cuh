// uniform_stage_v5rb.cuh -- the rotate-form stage kernel (uniform_stage_v5r.cuh) for a uniform box WITH
// BODIES (kernel form 'c'; the path picks it by itself for such a mesh, MMF_UNIFORM_BODIES=0 sends the mesh to the
// generic path instead).
//
// The reference marks the cells whose centroid lies inside a body box as not solved (src/main.cpp:221-237,
// src/body.cpp:80-95), gives every interface between a fluid and a solid cell BC_WALL (src/main.cpp:251-277)
// and, in euler::computeRHS, (a) skips interfaces whose two cells are both not solved (src/euler.cpp:181-183),
// (b) evaluates a wall interface between the fluid cell's state and its mirror image about the interface
// normal as seen from the fluid side (:198-225, :352-362, :322-339), (c) accumulates into solved cells only
// (:237-247); the RK loops skip the cells that are not solved (src/main.cpp:409-423).  The interface ids,
// hence the accumulation order of a fluid cell, do not depend on any of this.
//
// The flag travels with the data every face evaluation already receives: a solid cell reports a NEGATIVE max
// eigenvalue (lam = -1; a real one is |u_n| + a > 0) in the x shuffle, in the y record and in the carried z
// state (lam = -1 never raises a maximum).  This kernel evaluates a wall like any interface and does not store
// the cells that touch one (flag 2) nor the solid ones (flag 1): the wall cells -- a surface -- are recomputed
// reference-shaped by wall_cell_update (uniform_device.cuh; uniform_wall_cells_kernel before the stage kernel,
// uniform_wall_scatter_kernel behind it), so the hot path has no slow path, no call and no divergent branch.
// One byte per cell of extra traffic (the padded flag array).  Measured at 128^3 with two bodies (round 2): 0.43 ms
// per step against 0.29 ms without bodies and 1.60 ms on the generic path; the variant with the wall evaluation
// inside the kernel (a call at each face whose sides differ: spills around three call sites per plane) took 0.86 ms
// and was removed.
#pragma once

#include "uniform_stage_v5r.cuh"

namespace mmf {

// the interface (low | high) along AXIS; a negative lam marks a solid side.  A wall is evaluated like any interface
// here -- the fluid cell it belongs to is not stored by this kernel but recomputed by wall_cell_update -- and the value
// max(lam_fluid, -1) = lam_fluid it contributes to the face maximum is one the true maximum contains anyway.
template <int AXIS, bool FIXUP>
__device__ __forceinline__ double body_face_flux(const double *LU, const double *LF, const double ll, const double *HU,
                                                 const double *HF, const double hl, const double Ah, const DivConsts &,
                                                 double *AF)
{
    static_assert(FIXUP, "the in-kernel wall evaluation was removed: wall cells are recomputed by wall_cell_update");
    return llf_area_flux(LU, LF, ll, HU, HF, hl, Ah, AF);
}

template <int STAGE, int ORDER, int NW, bool FIXUP>
__global__ void __maxnreg__(stage_regs(NW))
uniform_stage_kernel_v5rb(const UniformGeom g, const double *__restrict__ Sin, const double *Un, double *Out,
                          const StepControl *__restrict__ ctl, double *__restrict__ max_eig, const int lz,
                          float *__restrict__ cta_est, const LoadClamp lc, const HaloWait hw,
                          const unsigned char *__restrict__ solid)
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
    const double *scol = Sin + col;
    const unsigned char *mcol = solid + col; // the flag array has the layout of one field
    double lmax = 0.0;
    float emax = 0.f;

    if (row == 0) {
        // ================= low halo row: publishes (U, Fy, lam_y) of row j for row 1 =================
        const double *sp = scol + (long long) (z0 + 1) * plane; // plane z0
        const unsigned char *mp = mcol + (long long) (z0 + 1) * plane;
        double *d = sm_d + lane;
        double nxt[NF];
        unsigned nsol = *mp;
#pragma unroll
        for (int k = 0; k < NF; ++k) nxt[k] = ldsin(sp + k * fs);
        for (int kz = z0; kz < z1; ++kz) {
            const int it = kz - z0;
            double cU[NF];
            const unsigned csol = nsol;
#pragma unroll
            for (int k = 0; k < NF; ++k) cU[k] = nxt[k];
            sp += plane;
            mp += plane;
            if (kz + 1 < z1) {
                nsol = *mp;
#pragma unroll
                for (int k = 0; k < NF; ++k) nxt[k] = ldsin(sp + k * fs);
            }
            CellPrim q;
            derive_cell(cU, dc, q);
            double cFy[NF], cly;
            axis_flux<1>(q, cFy, cly);
            if (csol == 1) cly = -1.0;
            if (it > 0) mbar_wait(&barF[1], (unsigned) ((it - 1) & 1)); // row 1 is done with the previous record
#pragma unroll
            for (int k = 0; k < NF; ++k) { d[k * 32] = cU[k]; d[(NF + k) * 32] = cFy[k]; }
            d[10 * 32] = cly;
            mbar_arrive_elect(&barD[0], lane);
        }
    } else if (row == NW - 1) {
        // ================= high halo row: computes the y face (j-1 | j) for row NW-2 ================
        const bool yf_ok = in_x && lane >= 1 && lane <= XW && j >= 0 && j <= g.ny;
        const double *sp = scol + (long long) (z0 + 1) * plane; // plane z0
        const unsigned char *mp = mcol + (long long) (z0 + 1) * plane;
        const double *d_dn = sm_d + (NW - 2) * 11 * 32 + lane;
        double *f = sm_f + (NW - 1) * NF * 32 + lane;
        double lmy = 0.0;
        double nxt[NF];
        unsigned nsol = *mp;
#pragma unroll
        for (int k = 0; k < NF; ++k) nxt[k] = ldsin(sp + k * fs);
        for (int kz = z0; kz < z1; ++kz) {
            const int it = kz - z0;
            double cU[NF];
            const unsigned csol = nsol;
#pragma unroll
            for (int k = 0; k < NF; ++k) cU[k] = nxt[k];
            sp += plane;
            mp += plane;
            if (kz + 1 < z1) {
                nsol = *mp;
#pragma unroll
                for (int k = 0; k < NF; ++k) nxt[k] = ldsin(sp + k * fs);
            }
            CellPrim q;
            derive_cell(cU, dc, q);
            double cFy[NF], cly;
            axis_flux<1>(q, cFy, cly);
            if (csol == 1) cly = -1.0;
            mbar_wait(&barD[NW - 2], (unsigned) (it & 1));
            double lU[NF], lF[NF], AFy[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) { lU[k] = d_dn[k * 32]; lF[k] = d_dn[(NF + k) * 32]; }
            const double ll  = d_dn[10 * 32];
            const double lam = body_face_flux<1, FIXUP>(lU, lF, ll, cU, cFy, cly, Ah, dc, AFy);
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
        const double *d_dn = sm_d + (row - 1) * 11 * 32 + lane;
        double *f_own = sm_f + row * NF * 32 + lane;
        const double *f_up = sm_f + (row + 1) * NF * 32 + lane;

        const long long zoff = (long long) (max(z0 - 1, lc.klo) + 1) * plane; // plane z0-1 (clamped)
        const double *sp  = scol + zoff;
        const unsigned char *mp = mcol + zoff;
        const double *unp = Un + col + (long long) (z0 + 1) * plane; // plane z0
        double *op = Out + col + (long long) z0 * plane;        // plane z0-1 (first store goes to plane z0)

        double pU[NF], pFz[NF], plz, pS[NF], pUn[NF], nxt[NF];
        unsigned nsol, pflag; // pflag: flag of the previous plane's cell (a cell is stored only if it is 0)
        double lmx = 0.0, lmy = 0.0, lmz = 0.0;
        // ---- prologue: plane z0-1 only provides the low side of the first z interface --------------
        {
            const unsigned psol = *mp;
#pragma unroll
            for (int k = 0; k < NF; ++k) pU[k] = ldsin(sp + k * fs);
            sp = scol + (long long) (z0 + 1) * plane; // plane z0
            mp = mcol + (long long) (z0 + 1) * plane;
            nsol = *mp;
#pragma unroll
            for (int k = 0; k < NF; ++k) nxt[k] = ldsin(sp + k * fs);
            CellPrim q;
            derive_cell(pU, dc, q);
            axis_flux<2>(q, pFz, plz);
            if (psol == 1) plz = -1.0;
            pflag = psol;
#pragma unroll
            for (int k = 0; k < NF; ++k) { pS[k] = 0.0; pUn[k] = 0.0; }
        }

#pragma unroll R_UNROLL
        for (int kz = z0; kz < z1; ++kz) {
            const unsigned par = (unsigned) ((kz - z0) & 1);
            double cU[NF];
            const unsigned csol = nsol;
#pragma unroll
            for (int k = 0; k < NF; ++k) cU[k] = nxt[k];
            if (kz + 1 <= lc.khi) { sp += plane; mp += plane; } // plane kz+1 (the ghost plane nz, or plane nz-1 again on a free-flow side)
            nsol = *mp;
#pragma unroll
            for (int k = 0; k < NF; ++k) nxt[k] = ldsin(sp + k * fs);
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
            if (csol == 1) cly = -1.0;
#pragma unroll
            for (int k = 0; k < NF; ++k) { d_own[k * 32] = cU[k]; d_own[(NF + k) * 32] = cFy[k]; }
            d_own[10 * 32] = cly;
            mbar_arrive_elect(&barD[row], lane);

            // ---- z interface (kz-1 | kz): completes plane kz-1 (never stored if that cell is solid) ----
            double cFz[NF], clz, AFz[NF];
            axis_flux<2>(q, cFz, clz);
            if (csol == 1) clz = -1.0;
            {
                const double lam = body_face_flux<2, FIXUP>(pU, pFz, plz, cU, cFz, clz, Ah, dc, AFz);
                lmz = (lam < lmz) ? lmz : lam;
            }
            finish_plane<STAGE>(pS, AFz, pU, pUn, dt, g.volume, dc.y_vol, op, fs, upd && kz > z0 && pflag == 0, est_max);
            op += plane;

            // ---- x interface (i-1 | i): lane-1's state by warp shuffle ---------------------------------
            double AFx[NF];
            {
                double cFx[NF], clx, lU[NF], lF[NF];
                axis_flux<0>(q, cFx, clx);
                if (csol == 1) clx = -1.0;
#pragma unroll
                for (int k = 0; k < NF; ++k) { lU[k] = shfl_up_d(cU[k]); lF[k] = shfl_up_d(cFx[k]); }
                const double ll  = shfl_up_d(clx);
                const double lam = body_face_flux<0, FIXUP>(lU, lF, ll, cU, cFx, clx, Ah, dc, AFx);
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
                const double lam = body_face_flux<1, FIXUP>(lU, lF, ll, cU, cFy, cly, Ah, dc, AFy);
                lmy = (lam < lmy) ? lmy : lam;
                // row-1 published record `it` only after it had read this row's flux `it-1`
#pragma unroll
                for (int k = 0; k < NF; ++k) f_own[k * 32] = AFy[k];
                mbar_arrive_elect(&barF[row], lane);
            }

            // ---- ordered accumulation (src/euler.cpp:153, 237-247), as in uniform_stage_v5r.cuh --------
            const int key_z = order_key<ORDER>(g.gz0 + kz, 2);
            double S[NF];
            const bool edge = (key_y < 0) | (key_z < 0); // warp-uniform: one row, one plane per warp
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
            pflag = csol;
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
            if (nsol == 1) clz = -1.0;
            const double lam = body_face_flux<2, FIXUP>(pU, pFz, plz, nxt, cFz, clz, Ah, dc, AFz);
            lmz = (lam < lmz) ? lmz : lam;
            finish_plane<STAGE>(pS, AFz, pU, pUn, dt, g.volume, dc.y_vol, op, fs, upd && pflag == 0, est_max);
        }
        lmax = xf_ok ? lmx : 0.0;
        if (yf_ok) lmax = (lmy < lmax) ? lmax : lmy;
        if (zf_ok) lmax = (lmz < lmax) ? lmax : lmz;
        emax = est_max;
    }

    block_maxima<NW>(lmax, max_eig, emax, (STAGE == 3) ? cta_est : nullptr, tid.tile, smem);
}

// the two small kernels around the stage kernel of form 'c': the wall cells' results into a compact buffer
// BEFORE the stage kernel runs (stage 3 updates U in place: U^n of a wall cell must still be there), and from
// the buffer into the output array behind it
template <int STAGE, int ORDER>
__global__ void __launch_bounds__(128) uniform_wall_cells_kernel(const UniformGeom g, const LoadClamp lc,
                                                                 const double *__restrict__ Sin, const double *__restrict__ Un,
                                                                 const unsigned char *__restrict__ flag,
                                                                 const int *__restrict__ list, const int n_list,
                                                                 const StepControl *__restrict__ ctl,
                                                                 double *__restrict__ compact, double *__restrict__ max_eig)
{
    if (STAGE >= 1 && ctl->active == 0.0) return;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    double lmax = 0.0;
    if (q < n_list) {
        double out[NF];
        lmax = wall_cell_update<STAGE, ORDER>(g, lc, Sin, Un, flag, (long long) list[q], (STAGE >= 1) ? ctl->dt : 0.0, out);
#pragma unroll
        for (int f = 0; f < NF; ++f) compact[(size_t) f * n_list + q] = out[f];
    }
    block_max_to_global(lmax, max_eig);
}

static __global__ void __launch_bounds__(128) uniform_wall_scatter_kernel(const long long fs, const int *__restrict__ list,
                                                                   const int n_list, const double *__restrict__ compact,
                                                                   double *__restrict__ Out,
                                                                   const StepControl *__restrict__ ctl, const int check_active)
{
    if (check_active && ctl->active == 0.0) return;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_list) return;
    const long long o = list[q];
#pragma unroll
    for (int f = 0; f < NF; ++f) Out[f * fs + o] = compact[(size_t) f * n_list + q];
}

} // namespace mmf
