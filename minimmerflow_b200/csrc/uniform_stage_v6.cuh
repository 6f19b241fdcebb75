// uniform_stage_v6.cuh -- the low-face streaming stage kernel (uniform_stage_v5.cuh explains the scheme)
// with the y exchange DECOUPLED by one plane.
//
// In v5 a row publishes its record (U, Fy, lam_y) of plane k, the row above turns it into the low y flux
// of its own cell and publishes that, and the first row waits for it IN THE SAME PLANE to subtract it as
// its -y_hi: record(r,k) -> flux(r+1,k) -> S(r,k).  That round trip ties every row to both neighbours
// within a plane, so the rows of a CTA move in lock step and any delay of one warp (a DRAM load, a lost
// issue slot) is handed to all of them: profiles/r01d shows 22-30 % of the warp time in those waits.
//
// Here -y_hi is subtracted ONE PLANE LATER, together with -z_hi which has to wait for the next plane
// anyway (both are the last two terms of the reference's accumulation order, src/euler.cpp:237-247, so
// the order of operations per value is unchanged and the result stays bit-identical):
//   iteration k of row r:  publish record(r,k) | z face (k-1|k)
//                          | wait flux(r+1,k-1): S(k-1) -= it; S(k-1) -= AFz; update + store plane k-1
//                          | x face | wait record(r-1,k) -> y face, publish flux(r,k)
//                          | S(k) = ordered low faces - x_hi (+ y_lo on the border)
// The only same-plane dependency left points downwards (record of row r-1), so the rows settle into a
// skew instead of a lock step, and the flux a row waits for was published about one plane earlier.
// Records and fluxes are double buffered (slot = plane parity) with one mbarrier per slot and row: with a
// single barrier the producer could complete two phases before the consumer's parity wait looks at the
// first one (try_wait.parity cannot tell phase k from phase k+2), which is a hang.  Why no slot is
// overwritten early and no barrier is overtaken is argued next to each wait below; tools/emu runs this
// source on a CPU SIMT emulator with the same mbarrier semantics and random warp delays to check it.
//
// Two situations keep the same-plane wait: the plane on the low z side of the DOMAIN, whose border face
// enters the sum after -y_hi (one plane of one z chunk), and the NUM_AXIS accumulation order
// (x_lo - x_hi + y_lo - y_hi + z_lo - z_hi for every cell).
#pragma once

#include "uniform_stage_v5r.cuh"

#include <type_traits>

#ifndef MMF_V6_LATE_UN
#define MMF_V6_LATE_UN 1
#endif
#ifndef MMF_V6_EARLY_RCP
#define MMF_V6_EARLY_RCP 0
#endif

namespace mmf {

// dynamic shared memory of the v6 kernels: records (11) + fluxes (5) doubles per lane, row and slot,
// four mbarriers per row
// (merged halo: one more virtual row, see MH below)
__host__ __device__ constexpr size_t stage_v6_smem_bytes(int nw, bool merged_halo = false)
{
    return (size_t) 2 * (nw + (merged_halo ? 1 : 0)) * 16 * 32 * sizeof(double) +
           (size_t) 4 * (nw + (merged_halo ? 1 : 0)) * sizeof(unsigned long long);
}

// cells a CTA of nw warps updates along y
__host__ __device__ constexpr int stage_v6_rows(int nw, bool merged_halo) { return merged_halo ? nw - 1 : nw - 2; }

// The plane loop, two planes per trip with the slot (= plane parity) a compile-time constant in each
// copy of the body, so that slot offsets fold into the shared-memory addresses.
#define MMF_V6_PLANE_PAIRS(IT0, SLOT)                                                                        \
    _Pragma("unroll 1") for (int IT0 = 0; IT0 < z1 - z0; IT0 += 2)                                          \
        _Pragma("unroll") for (int SLOT = 0; SLOT < 2; ++SLOT)                                              \
            if (IT0 + SLOT < z1 - z0)

// MH ("merged halo", kernel form 'h'): ONE warp serves both halo rows of the tile -- it publishes the record of
// the row below the tile and turns the record of the tile's top row into that row's -y_hi flux -- so a CTA
// of NW warps updates NW-1 rows instead of NW-2 (11 of 12 warps do full work instead of 10; with three
// warps per scheduler the two schedulers that hosted a halo row were under-used anyway).  The halo warp
// handles its top part one plane late (plane it-1 in iteration it): the flux is only consumed one plane
// later by the decoupled scheme, and the halo warp then practically never waits.
template <int STAGE, int ORDER, int NW, bool XG, bool MH = false>
__global__ void __maxnreg__(stage_regs(NW))
uniform_stage_kernel_v6(const UniformGeom g, const double *__restrict__ Sin, const double *Un, double *Out,
                        const StepControl *__restrict__ ctl, double *__restrict__ max_eig, const int lz,
                        float *__restrict__ cta_est, const LoadClamp lc, const HaloWait hw, const XGhost xg)
{
    extern __shared__ double smem[];
    // sm_d[slot][row][q][lane], q = U0..U4, Fy0..Fy4, lam_y ; sm_f[slot][row][k][lane] = area * flux of (j-1 | j)
    constexpr int NR = MH ? NW + 1 : NW;                 // rows of the exchange buffers (MH: virtual row NW on top)
    constexpr int RT = stage_v6_rows(NW, MH);            // rows updated per tile
    constexpr int DS = NR * 11 * 32, FS = NR * NF * 32;  // doubles per slot
    double *sm_d = smem;
    double *sm_f = smem + 2 * DS;
    unsigned long long *barD = reinterpret_cast<unsigned long long *>(sm_f + 2 * FS); // [slot][row]: record published
    unsigned long long *barF = barD + 2 * NR;                                          // [slot][row]: low y flux published

    if (STAGE >= 1 && ctl->active == 0.0) return;

    const int lane = threadIdx.x & 31;
    const int row  = threadIdx.x >> 5;
    const TileId tid = stage_tile(hw);
    if (threadIdx.x < 2 * NR) {
        mbar_init(&barD[threadIdx.x], 1);
        mbar_init(&barF[threadIdx.x], 1);
    }
    halo_wait(hw, tid);
    __syncthreads();

    const int i  = tid.bx * XW - 1 + lane;
    const int j  = tid.by * RT - 1 + row;
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
    const double *scol = Sin + col;
    int sfs_lane = (int) fs, splane_lane = (int) plane; // element counts: < 2^31 for any box that fits one GPU
    if (XG) {
        if (xg.lo && i < 0)     { scol = xg.lo + (jc + 1); sfs_lane = (int) xg.fs; splane_lane = xg.pitch; }
        if (xg.hi && i >= g.nx) { scol = xg.hi + (jc + 1); sfs_lane = (int) xg.fs; splane_lane = xg.pitch; }
    }
    const long long sfs = XG ? (long long) sfs_lane : fs, splane = XG ? (long long) splane_lane : plane;
    double lmax = 0.0;
    float emax = 0.f;

    if (MH && row == 0) {
        // ================= merged halo warp ==========================================================
        // bottom: row j = tile_j0 - 1, publishes (U, Fy, lam_y) for row 1;  top: virtual row NW (j = tile_j0 - 1
        // + NW), turns the record of row NW-1 into the flux of the face between them, one plane late.
        const int jt  = j + NW;
        const int jtc = min(max(jt, lc.jlo), lc.jhi);
        const bool yf_ok = in_x && lane >= 1 && lane <= XW && jt >= 0 && jt <= g.ny;
        const double *scol_t = Sin + (long long) (jtc + 1) * g.px + (ic + 1);
        if (XG) {
            if (xg.lo && i < 0)     scol_t = xg.lo + (jtc + 1);
            if (xg.hi && i >= g.nx) scol_t = xg.hi + (jtc + 1);
        }
        const double *sp_b = scol + (long long) (z0 + 1) * splane;   // plane z0
        const double *sp_t = scol_t + (long long) (z0 + 1) * splane;
        const int n = z1 - z0;
        double lmy = 0.0;
        double nb[NF], nt[NF];     // prefetched plane `it` of the bottom / top row
        double tU[NF], tFy[NF], tly = 0.0; // top row, plane it-1
#pragma unroll
        for (int k = 0; k < NF; ++k) { nb[k] = ldsin(sp_b + k * sfs); nt[k] = ldsin(sp_t + k * sfs); tU[k] = 0.0; tFy[k] = 0.0; }
#pragma unroll 1
        for (int it = 0; it <= n; ++it) {
            const int slot = it & 1;
            double bU[NF], cT[NF], bFy[NF], bly = 0.0, cFy[NF], cly = 0.0;
            if (it < n) {
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
                // bottom record `it`.  The slot still holds record it-2: row 1 has read it once its flux it-2 is
                // out, and it cannot complete that barrier again before it has seen record `it`.
                if (it >= 2) mbar_wait(&barF[slot * NR + 1], (unsigned) (((it - 2) >> 1) & 1));
                double *d = sm_d + slot * DS + lane;
#pragma unroll
                for (int k = 0; k < NF; ++k) { d[k * 32] = bU[k]; d[(NF + k) * 32] = bFy[k]; }
                d[10 * 32] = bly;
                mbar_arrive_elect(&barD[slot * NR + 0], lane);
            }
            if (it >= 1) {
                // top flux of plane it-1.  Row NW-1 publishes record it+1 on this barrier only at the top of
                // its iteration it+1, after its iteration `it` has taken the flux written here.
                const int ps = (it - 1) & 1;
                mbar_wait(&barD[ps * NR + NW - 1], (unsigned) (((it - 1) >> 1) & 1));
                const double *d_dn = sm_d + ps * DS + (NW - 1) * 11 * 32 + lane;
                double lU[NF], lF[NF], AFy[NF];
#pragma unroll
                for (int k = 0; k < NF; ++k) { lU[k] = d_dn[k * 32]; lF[k] = d_dn[(NF + k) * 32]; }
                const double ll  = d_dn[10 * 32];
                const double lam = llf_area_flux(lU, lF, ll, tU, tFy, tly, Ah, AFy);
                lmy = (lam < lmy) ? lmy : lam;
                // the slot held flux it-3: row NW-1 took it during its plane it-2, before it published the
                // record it-1 this warp has just waited for
                double *f = sm_f + ps * FS + NW * NF * 32 + lane;
#pragma unroll
                for (int k = 0; k < NF; ++k) f[k * 32] = AFy[k];
                mbar_arrive_elect(&barF[ps * NR + NW], lane);
            }
            if (it < n) {
#pragma unroll
                for (int k = 0; k < NF; ++k) { tU[k] = cT[k]; tFy[k] = cFy[k]; }
                tly = cly;
            }
        }
        lmax = yf_ok ? lmy : 0.0;
    } else if (!MH && row == 0) {
        // ================= low halo row: publishes (U, Fy, lam_y) of row j for row 1 =================
        const double *sp = scol + (long long) (z0 + 1) * splane; // plane z0
        double nxt[NF];
#pragma unroll
        for (int k = 0; k < NF; ++k) nxt[k] = ldsin(sp + k * sfs);
        MMF_V6_PLANE_PAIRS(it0, slot) {
            const int it = it0 + slot, kz = z0 + it;
            double *d = sm_d + slot * DS + lane;
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
            // the slot still holds record it-2: row 1 has read it once its flux it-2 is out.  Row 1 cannot
            // complete this barrier again before it has seen record `it`, which follows this wait.
            if (it >= 2) mbar_wait(&barF[slot * NR + 1], (unsigned) (((it - 2) >> 1) & 1));
#pragma unroll
            for (int k = 0; k < NF; ++k) { d[k * 32] = cU[k]; d[(NF + k) * 32] = cFy[k]; }
            d[10 * 32] = cly;
            mbar_arrive_elect(&barD[slot * NR + 0], lane);
        }
    } else if (!MH && row == NW - 1) {
        // ================= high halo row: computes the y face (j-1 | j) for row NW-2 ================
        const bool yf_ok = in_x && lane >= 1 && lane <= XW && j >= 0 && j <= g.ny;
        const double *sp = scol + (long long) (z0 + 1) * splane; // plane z0
        double lmy = 0.0;
        double nxt[NF];
#pragma unroll
        for (int k = 0; k < NF; ++k) nxt[k] = ldsin(sp + k * sfs);
        MMF_V6_PLANE_PAIRS(it0, slot) {
            const int it = it0 + slot, kz = z0 + it;
            const double *d_dn = sm_d + slot * DS + (NW - 2) * 11 * 32 + lane;
            double *f = sm_f + slot * FS + (NW - 1) * NF * 32 + lane;
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
            // row NW-2 publishes record it+2 on this barrier only after it has taken flux `it`, below
            mbar_wait(&barD[slot * NR + NW - 2], (unsigned) ((it >> 1) & 1));
            double lU[NF], lF[NF], AFy[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) { lU[k] = d_dn[k * 32]; lF[k] = d_dn[(NF + k) * 32]; }
            const double ll  = d_dn[10 * 32];
            const double lam = llf_area_flux(lU, lF, ll, cU, cFy, cly, Ah, AFy);
            lmy = (lam < lmy) ? lmy : lam;
            // the slot held flux it-2: row NW-2 took it during its plane it-1, before it published record `it`
#pragma unroll
            for (int k = 0; k < NF; ++k) f[k * 32] = AFy[k];
            mbar_arrive_elect(&barF[slot * NR + NW - 1], lane);
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

        double *d_own = sm_d + row * 11 * 32 + lane;             // + slot * DS
        const double *d_dn = sm_d + (row - 1) * 11 * 32 + lane;
        double *f_own = sm_f + row * NF * 32 + lane;             // + slot * FS
        const double *f_up = sm_f + (row + 1) * NF * 32 + lane;

        const double *sp  = scol + (long long) (max(z0 - 1, lc.klo) + 1) * splane; // plane z0-1 (clamped)
        // U^n of a plane is needed when the plane is finished, one iteration after its residual input.
        // LATE_UN loads it at the top of THAT iteration (consumed a third of an iteration later: enough to
        // cover the DRAM latency) instead of carrying it across the loop edge: 10 registers less at the
        // pressure peak, which is what the 16-warp (128-register) builds of stages 2 and 3 lack.
        constexpr bool LATE_UN = MMF_V6_LATE_UN;
        const double *unp = Un + col + (long long) (z0 + (LATE_UN ? 0 : 1)) * plane; // plane z0 (LATE_UN: z0-1)
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
        // EARLY_RCP: 1/rho and 1/rho^2 of a plane -- the root of all its dependency chains -- are started one
        // plane early, as soon as the prefetched rho is there (4 more registers across the loop edge)
        constexpr bool EARLY_RCP = MMF_V6_EARLY_RCP;
        double ny = 0.0, nyrr = 0.0;
        if (EARLY_RCP) { ny = rcp_nr(nxt[FID_RHO]); nyrr = rcp_nr(nxt[FID_RHO] * nxt[FID_RHO]); }

        // one plane; the slot (= plane parity) arrives as a compile-time constant so that the slot offsets
        // fold into the shared-memory addresses
        auto body = [&](auto slot_tag, const int it) {
            constexpr int slot = decltype(slot_tag)::value;
            const int kz = z0 + it;
            const unsigned par = (unsigned) ((it >> 1) & 1);
            double cU[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) cU[k] = nxt[k];
            if (kz + 1 <= lc.khi) sp += splane; // plane kz+1 (the ghost plane nz, or plane nz-1 again on a free-flow side)
#pragma unroll
            for (int k = 0; k < NF; ++k) nxt[k] = ldsin(sp + k * sfs);
            double cUn[NF];
            if (STAGE >= 2 && upd && !LATE_UN) {
#pragma unroll
                for (int k = 0; k < NF; ++k) cUn[k] = unp[k * fs];
            }
            if (STAGE >= 2 && upd && LATE_UN && it > 0) {
#pragma unroll
                for (int k = 0; k < NF; ++k) pUn[k] = unp[k * fs];
            }
            unp += plane;

            CellPrim q;
            if (EARLY_RCP) {
                derive_cell_pre(cU, ny, nyrr, dc, q);
                // the reciprocals the NEXT plane starts from, off its critical path (its rho was prefetched above)
                ny   = rcp_nr(nxt[FID_RHO]);
                nyrr = rcp_nr(nxt[FID_RHO] * nxt[FID_RHO]);
            } else {
                derive_cell(cU, dc, q);
            }

            // ---- y record for row+1.  The slot held record it-2; row+1 read it before it published flux
            //      it-2, and this row took that flux during plane it-1 (or it-2): free. -------------------
            double cFy[NF], cly;
            axis_flux<1>(q, cFy, cly);
            {
                double *d = d_own + slot * DS;
#pragma unroll
                for (int k = 0; k < NF; ++k) { d[k * 32] = cU[k]; d[(NF + k) * 32] = cFy[k]; }
                d[10 * 32] = cly;
            }
            mbar_arrive_elect(&barD[slot * NR + row], lane);

            // ---- z interface (kz-1 | kz) ---------------------------------------------------------------
            double cFz[NF], clz, AFz[NF];
            axis_flux<2>(q, cFz, clz);
            {
                const double lam = llf_area_flux(pU, pFz, plz, cU, cFz, clz, Ah, AFz);
                lmz = (lam < lmz) ? lmz : lam;
            }
            // ---- plane kz-1: -y_hi (the low y face of row+1, published one plane ago), -z_hi, update ----
            // row+1 completes this barrier again (flux it+1) only after it has seen record it+1 of this
            // row, which is published in the next iteration, after this wait.
            if (ORDER != NUM_AXIS && it > 0 && g.gz0 + kz - 1 != 0) {
                mbar_wait(&barF[(slot ^ 1) * NR + row + 1], (unsigned) (((it - 1) >> 1) & 1));
                const double *f = f_up + (slot ^ 1) * FS;
#pragma unroll
                for (int k = 0; k < NF; ++k) pS[k] -= f[k * 32];
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
            // row-1 publishes record it+2 on this barrier only after it has taken this row's flux `it`
            double AFy[NF];
            mbar_wait(&barD[slot * NR + row - 1], par);
            {
                const double *d = d_dn + slot * DS;
                double lU[NF], lF[NF];
#pragma unroll
                for (int k = 0; k < NF; ++k) { lU[k] = d[k * 32]; lF[k] = d[(NF + k) * 32]; }
                const double ll  = d[10 * 32];
                const double lam = llf_area_flux(lU, lF, ll, cU, cFy, cly, Ah, AFy);
                lmy = (lam < lmy) ? lmy : lam;
                // the slot held flux it-2: row-1 took it during its plane it-1, before it published the
                // record `it` this row has just waited for
                double *f = f_own + slot * FS;
#pragma unroll
                for (int k = 0; k < NF; ++k) f[k * 32] = AFy[k];
                mbar_arrive_elect(&barF[slot * NR + row], lane);
            }

            // ---- ordered accumulation (src/euler.cpp:153, 237-247) ------------------------------------
            // interior low faces first, sorted by their creator (largest key first; a+b commutes, so
            // only the LAST one matters); then the cell's own faces in the order it created them:
            // (-x if border) +x (-y if border) +y (-z if border) +z; low faces `+=`, high faces `-=`.
            // (plane kz-1 is finished: its registers take the sum of plane kz)
            const int key_z = order_key<ORDER>(g.gz0 + kz, 2);
            const bool edge = (key_y < 0) | (key_z < 0); // warp-uniform: one row, one plane per warp
            if (ORDER == NUM_AXIS) {
#pragma unroll
                for (int k = 0; k < NF; ++k) pS[k] = 0.0 + AFx[k];
            } else if (!edge) {
                // a border low face in x alone is simply "last" (key -1), directly followed by -x_hi
                const int last = (key_x < key_y) ? ((key_x < key_z) ? 0 : 2) : ((key_y < key_z) ? 1 : 2);
#pragma unroll
                for (int k = 0; k < NF; ++k) {
                    const double p = (last == 0) ? AFy[k] : AFx[k];
                    const double t = (last == 2) ? AFy[k] : AFz[k];
                    const double r = (last == 0) ? AFx[k] : (last == 1) ? AFy[k] : AFz[k];
                    pS[k] = (p + t) + r;
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
                    pS[k] = s;
                }
            }
            // -x_hi: the low x face of lane+1
#pragma unroll
            for (int k = 0; k < NF; ++k) pS[k] -= shfl_down_d(AFx[k]);
            if (ORDER == NUM_AXIS || (edge && key_y < 0)) {
#pragma unroll
                for (int k = 0; k < NF; ++k) pS[k] += AFy[k];
            }
            // -y_hi follows one plane later, except where a +z_lo term has to come behind it
            if (ORDER == NUM_AXIS || key_z < 0) {
                mbar_wait(&barF[slot * NR + row + 1], par);
                const double *f = f_up + slot * FS;
#pragma unroll
                for (int k = 0; k < NF; ++k) pS[k] -= f[k * 32];
#pragma unroll
                for (int k = 0; k < NF; ++k) pS[k] += AFz[k];
            }

            // ---- plane kz becomes the previous plane ---------------------------------------------------
#pragma unroll
            for (int k = 0; k < NF; ++k) { pU[k] = cU[k]; pFz[k] = cFz[k]; }
            plz = clz;
            if (STAGE >= 2 && !LATE_UN) {
#pragma unroll
                for (int k = 0; k < NF; ++k) pUn[k] = cUn[k];
            }
        };
        {
            // two planes per trip, no test between them: ptxas renames the rotating state instead of copying it
            const int n = z1 - z0;
            int it = 0;
#pragma unroll 1
            for (; it + 1 < n; it += 2) {
                body(std::integral_constant<int, 0>{}, it);
                body(std::integral_constant<int, 1>{}, it + 1);
            }
            if (it < n) body(std::integral_constant<int, 0>{}, it);
        }

        // ---- epilogue: plane z1 only closes the last z interface; plane z1-1 still lacks -y_hi ------
        {
            if (STAGE >= 2 && upd && LATE_UN) {
#pragma unroll
                for (int k = 0; k < NF; ++k) pUn[k] = unp[k * fs];
            }
            CellPrim q;
            if (EARLY_RCP) derive_cell_pre(nxt, ny, nyrr, dc, q);
            else           derive_cell(nxt, dc, q);
            double cFz[NF], clz, AFz[NF];
            axis_flux<2>(q, cFz, clz);
            const double lam = llf_area_flux(pU, pFz, plz, nxt, cFz, clz, Ah, AFz);
            lmz = (lam < lmz) ? lmz : lam;
            const int it = z1 - z0; // the iteration that would follow
            if (ORDER != NUM_AXIS && g.gz0 + z1 - 1 != 0) {
                mbar_wait(&barF[((it - 1) & 1) * NR + row + 1], (unsigned) (((it - 1) >> 1) & 1));
                const double *f = f_up + ((it - 1) & 1) * FS;
#pragma unroll
                for (int k = 0; k < NF; ++k) pS[k] -= f[k * 32];
            }
            finish_plane<STAGE>(pS, AFz, pU, pUn, dt, g.volume, dc.y_vol, op, fs, upd, est_max);
        }
        lmax = xf_ok ? lmx : 0.0;
        if (yf_ok) lmax = (lmy < lmax) ? lmax : lmy;
        if (zf_ok) lmax = (lmz < lmax) ? lmax : lmz;
        emax = est_max;
    }

    block_maxima<NW>(lmax, max_eig, emax, (STAGE == 3) ? cta_est : nullptr, tid.tile, smem);
}

#undef MMF_V6_PLANE_PAIRS

} // namespace mmf
