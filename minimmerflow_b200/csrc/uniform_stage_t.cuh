// uniform_stage_t.cuh -- kernel form 't': the low-face streaming stage kernel (uniform_stage_v5.cuh explains the
// scheme, uniform_stage_v5r.cuh is the form it grew from) with its input staged by the TMA unit.
//
// What changes against the rotate form:
//   * no thread loads the residual input or U^n from global memory.  One elected lane issues, per z plane, ONE bulk
//     tensor copy of the CTA's whole (32-cell x window) x (NW rows) x (5 fields) box of the residual input and one
//     of the (NW-2 update rows) box of U^n into a ring of D shared-memory slots, D-1 planes ahead of the plane
//     being worked on; the bytes are counted on the slot's `full` mbarrier (ptx_helpers.cuh: tma_load_4d).  The
//     update warps read their cell with LDS when they need it: the two prefetch register sets of the rotate form
//     (next plane's state, U^n one plane ahead: 20 registers), their LDGs with address arithmetic and the
//     long-scoreboard stalls behind them are gone.
//   * free-flow sides: the rotate form clamps its LOAD coordinates (LoadClamp); here the copy brings the raw box
//     and a lane reads its slot at the clamped column / row instead -- the same cell, the same bits.  The z clamp is
//     applied to the plane coordinate of the copy.  Boxes that stick out of the padded array are zero filled by the
//     hardware and never read (the clamped coordinates always lie inside).
//   * the y record a row hands up shrinks from (U, Fy, lam_y) to (Fy, lam_y): the row above reads the lower
//     cell's U from the slot itself.  The previous plane's state (low side of the z interface, and the W term of
//     the RK update) is read back from the previous slot instead of being carried in registers.
//   * a slot is handed back (`empty` mbarrier, one elected arrival per warp) when a warp has finished the plane
//     AFTER it, because that is when the previous plane's U and U^n are last read.
// Arithmetic per value: the same helpers in the same order as every other form, so the results are bit-identical.
//
// Not here: compact x ghost columns (XGhost: partition sides across x).  Such launches keep the rotate form.
//
// BODY = true (kernel form 'b'): the same kernel for a uniform box WITH BODIES, by the scheme uniform_body_cells.cuh
// explains -- one flag byte per padded cell (1 = not solved, 2 = fluid cell with a wall interface); a cell that is not
// solved reports a NEGATIVE max eigenvalue in its x / y / z record (the sign bit is set: a negative value never raises a
// maximum, and max(lam_fluid, negative) = lam_fluid is in the true maximum anyway), a wall is evaluated like any
// interface, flagged cells are not stored and the wall cells -- a surface -- are recomputed reference-shaped around the
// kernel (uniform_wall_cells_kernel / uniform_wall_scatter_kernel).  The flag is a per-thread one-byte global load,
// one plane ahead: a bulk tensor copy of bytes would need x windows that start on 16-byte boundaries.
#pragma once

#include "uniform_stage_v5.cuh"

namespace mmf {

// planes per trip of the update rows' loop.  0 = by CTA size (measured, profiles/r02d_experiments.md): 4 = the ring depth
// at 12 warps -- the slot of every unrolled copy is then a compile-time constant and 168 registers leave the compiler
// room to overlap the copies --, 2 at 16 warps (128 registers)
#ifndef MMF_T_UNROLL
#define MMF_T_UNROLL 0
#endif
// slots of the ring: two planes are being read (the current and the previous one), the others are in flight
#ifndef MMF_T_DEPTH
#define MMF_T_DEPTH 4
#endif
constexpr int T_DEPTH = MMF_T_DEPTH;
__host__ __device__ constexpr int t_unroll(int nw) { return MMF_T_UNROLL > 0 ? MMF_T_UNROLL : (nw <= 12 ? 4 : 2); }
// experiment switches (profiles/r02d_experiments.md): the left x neighbour's U from the slot instead of five shuffles;
// the ordered sum with a warp-uniform branch and two selects per field instead of three; the previous plane's U
// carried in registers instead of read back from its slot
#ifndef MMF_T_XLDS
#define MMF_T_XLDS 1
#endif
#ifndef MMF_T_SELBR
#define MMF_T_SELBR 1
#endif
// (-1 = by CTA size: carried at 12 warps, where the registers are there, re-read at 16)
#ifndef MMF_T_CARRY
#define MMF_T_CARRY -1
#endif
// the previous plane's slot handed back right after its last read (the z interface) instead of at the end of the step
#ifndef MMF_T_EARLY_RELEASE
#define MMF_T_EARLY_RELEASE 1
#endif

// Rows of a CTA's tile.  Two halo warps (MH = false): warps 0 and NW-1 serve the low and the high halo row, NW-2 warps
// update.  Merged halo warp (MH = true): warp 0 serves BOTH halo rows -- it publishes the record of the row below the
// tile and turns the record of the tile's top row into that row's -y_hi; together that is 175 FP64 instructions per
// plane against the 224 of an update row -- and NW-1 warps update: 15 rows per 16-warp CTA instead of 14.
__host__ __device__ constexpr int t_update_rows(int nw, bool mh) { return mh ? nw - 1 : nw - 2; }
__host__ __device__ constexpr int t_tile_rows(int nw, bool mh) { return t_update_rows(nw, mh) + 2; }

// ring geometry (doubles): a slot = the residual-input box [NF][tile rows][32] and, for stages 2 and 3, the U^n box
// [NF][update rows][32]; both are multiples of 128 bytes, which the destination of a bulk tensor copy must be aligned to
__host__ __device__ constexpr int t_sin_doubles(int nw, bool mh) { return NF * t_tile_rows(nw, mh) * 32; }
__host__ __device__ constexpr int t_un_doubles(int nw, int stage, bool mh) { return stage >= 2 ? NF * t_update_rows(nw, mh) * 32 : 0; }
__host__ __device__ constexpr int t_slot_doubles(int nw, int stage, bool mh) { return t_sin_doubles(nw, mh) + t_un_doubles(nw, stage, mh); }
// behind the ring: records [tile rows][6][32] (Fy, lam_y), low y fluxes [tile rows][NF][32], then the mbarriers
__host__ __device__ constexpr size_t stage_t_smem_bytes(int nw, int stage, int depth, bool mh)
{
    return (size_t) (depth * t_slot_doubles(nw, stage, mh) + t_tile_rows(nw, mh) * (6 + NF) * 32) * sizeof(double) +
           (size_t) (2 * depth + 2 * t_tile_rows(nw, mh)) * sizeof(unsigned long long);
}

// (BODY) the sign bit of a max eigenvalue marks a cell that is not solved: flag 1 -> bit 31 of the high word
__device__ __forceinline__ double body_mark(const double lam, const unsigned flag)
{
    return __hiloint2double(__double2hiint(lam) | (int) (flag << 31), __double2loint(lam));
}

template <int STAGE, int ORDER, int NW, int D, bool MH, bool BODY = false>
__global__ void __maxnreg__(stage_regs(NW))
uniform_stage_kernel_t(const UniformGeom g, double *Out, const StepControl *__restrict__ ctl, double *__restrict__ max_eig,
                       const int lz, float *__restrict__ cta_est, const LoadClamp lc, const HaloWait hw,
                       const __grid_constant__ TmaDesc smap, const __grid_constant__ TmaDesc umap,
                       const unsigned char *__restrict__ solid = nullptr)
{
    extern __shared__ double smem[]; // (no static shared memory: the dynamic window starts 1 KB aligned)
    constexpr int NU = t_update_rows(NW, MH), NR = t_tile_rows(NW, MH); // rows the CTA updates / rows of its tile
    constexpr bool CARRY = (MMF_T_CARRY < 0) ? (NW <= 12) : (MMF_T_CARRY != 0);
    constexpr int UNROLL = t_unroll(NW);
    constexpr int SLOT = t_slot_doubles(NW, STAGE, MH);
    constexpr int SIN = t_sin_doubles(NW, MH), UN = t_un_doubles(NW, STAGE, MH);
    constexpr int FSTR = NR * 32;        // field stride inside the residual-input box
    constexpr int UFSTR = NU * 32;       // ... inside the U^n box
    double *ring = smem;
    double *sm_r = smem + D * SLOT;            // records: sm_r[tile row][q][lane], q = Fy0..Fy4, lam_y
    double *sm_f = sm_r + NR * 6 * 32;         // sm_f[tile row][k][lane] = area * flux of (j-1 | j)
    unsigned long long *full  = reinterpret_cast<unsigned long long *>(sm_f + NR * NF * 32); // slot filled
    unsigned long long *empty = full + D;      // slot handed back by all NW warps
    unsigned long long *barD  = empty + D;     // record of tile row r published
    unsigned long long *barF  = barD + NR;     // low y flux of tile row r published

    if (STAGE >= 1 && ctl->active == 0.0) return;

    const int lane = threadIdx.x & 31;
    const int row  = threadIdx.x >> 5;
    const TileId tid = stage_tile(hw);
    if (threadIdx.x < NR) {
        mbar_init(&barD[threadIdx.x], 1);
        mbar_init(&barF[threadIdx.x], 1);
    }
    if (threadIdx.x < D) {
        mbar_init(&full[threadIdx.x], 1);
        mbar_init(&empty[threadIdx.x], NW);
    }
    fence_barrier_init();
    halo_wait(hw, tid);
    __syncthreads();

    const int i0 = tid.bx * XW - 1;        // cell coordinates of the box's first column / row
    const int j0 = tid.by * NU - 1;
    const int i  = i0 + lane;
    const int j  = j0 + row;               // warp w works on tile row w (the merged halo warp: on rows 0 and NR-1)
    const int z0 = tid.bz * lz;
    const int z1 = min(z0 + lz, g.nz);
    const int nsteps = z1 - z0;            // planes this CTA updates; ring steps 0 .. nsteps+1 = planes z0-1 .. z1
    const bool in_x = (i >= 0 && i < g.nx);
    const bool in_y = (j >= 0 && j < g.ny);
    // where this lane finds its (clamped) cell and its low y neighbour inside a slot
    const int cx    = min(max(i, lc.ilo), lc.ihi) - i0;
    const int ry    = min(max(j, lc.jlo), lc.jhi) - j0;
    const int ry_dn = min(max(j - 1, lc.jlo), lc.jhi) - j0;
    const int own = ry * 32 + cx;
    const int dn  = ry_dn * 32 + cx;
    // the high halo row of the tile (tile row NR-1), for the warp that serves it
    const int j_hi   = j0 + NR - 1;
    const int own_hi = (min(max(j_hi, lc.jlo), lc.jhi) - j0) * 32 + cx;
    const int dn_hi  = (min(max(j_hi - 1, lc.jlo), lc.jhi) - j0) * 32 + cx;

    // (BODY) flag of this lane's (clamped) cell: the flag array has the layout of one field, its ghost shell repeats the
    // flag of the cell it touches, so the plane coordinate needs no clamp.  Offsets of plane z0-1 (ring step 0).
    const int mplane = g.py * g.px;
    int moff    = BODY ? z0 * mplane + (j0 + ry + 1) * g.px + (i0 + cx + XOFF) : 0;
    int moff_hi = BODY ? z0 * mplane + (own_hi / 32 + j0 + 1) * g.px + (i0 + cx + XOFF) : 0; // (the high halo row's cell)

    const double Ah = 0.5 * g.area;
    DivConsts dc;
    dc.y_gm1 = rcp_nr(GM1);
    dc.y_c1  = rcp_nr(TWO_OVER_GM1);
    dc.y_vol = rcp_nr(g.volume);

    double lmax = 0.0;
    float emax = 0.f;

    // ---- the high halo row's step: the y face (j-1 | j) between the tile's top row and the row above it -----------
    double lmy_hi = 0.0;
    auto high_halo_step = [&](const int s, const unsigned csol_hi) {
        const double *ts = ring + (s % D) * SLOT;
        const double *r_dn = sm_r + (NR - 2) * 6 * 32 + lane;
        double *f = sm_f + (NR - 1) * NF * 32 + lane;
        double cU[NF];
#pragma unroll
        for (int k = 0; k < NF; ++k) cU[k] = ts[k * FSTR + own_hi];
        CellPrim q;
        derive_cell(cU, dc, q);
        double cFy[NF], cly;
        axis_flux<1>(q, cFy, cly);
        if (BODY) cly = body_mark(cly, csol_hi);
        double lU[NF], lF[NF], AFy[NF];
#pragma unroll
        for (int k = 0; k < NF; ++k) lU[k] = ts[k * FSTR + dn_hi];
        mbar_wait(&barD[NR - 2], (unsigned) ((s - 1) & 1));
#pragma unroll
        for (int k = 0; k < NF; ++k) lF[k] = r_dn[k * 32];
        const double ll  = r_dn[NF * 32];
        const double lam = llf_area_flux(lU, lF, ll, cU, cFy, cly, Ah, AFy);
        lmy_hi = (lam < lmy_hi) ? lmy_hi : lam;
        // the row below published this record only after it had read the previous flux: the slot is free
#pragma unroll
        for (int k = 0; k < NF; ++k) f[k * 32] = AFy[k];
        mbar_arrive_elect(&barF[NR - 1], lane);
    };
    const bool yf_ok_hi = in_x && lane >= 1 && lane <= XW && j_hi >= 0 && j_hi <= g.ny;

    if (row == 0) {
        // ================= low halo row: publishes (Fy, lam_y) of row j0 for row 1; lane 0 feeds the ring ==========
        // (MH: the same warp then serves the high halo row)
        // step p: residual-input plane clamp(z0-1+p) and, for 1 <= p <= nsteps, U^n of plane z0-1+p
        auto produce = [&](const int p) {
            const int slot = p % D;
            if (p >= D) mbar_wait(&empty[slot], (unsigned) ((p / D - 1) & 1));
            const bool with_un = STAGE >= 2 && p >= 1 && p <= nsteps;
            double *dst = ring + slot * SLOT;
            mbar_expect_tx(&full[slot], (unsigned) ((SIN + (with_un ? UN : 0)) * sizeof(double)));
            const int kp = min(max(z0 - 1 + p, lc.klo), lc.khi);
            tma_load_4d(&smap, dst, &full[slot], i0 + XOFF, j0 + 1, kp + 1, 0);
            if (with_un) tma_load_4d(&umap, dst + SIN, &full[slot], i0 + XOFF, j0 + 2, z0 + p, 0);
            mbar_arrive(&full[slot]);
        };
        // all D slots start out free: steps 0 .. D-1 go out at once; from then on the copy of step s+D-1 is issued at
        // the END of this row's step s -- this row is the lightest of the CTA and would otherwise sit waiting for
        // row 1 to consume its record -- and waits for the slot of step s-1, which the update rows hand back right
        // behind the z interface of step s
        if (lane == 0) {
            if (hw.flags) fence_proxy_async_global(); // the neighbours' layers arrived through the generic proxy
            for (int p = 0; p <= min(D - 1, nsteps + 1); ++p) produce(p);
        }
        double *r = sm_r + lane;
        unsigned nsol = 0, nsol_hi = 0; // (BODY) flags of the next step's cells
        if (BODY) { moff += mplane; moff_hi += mplane; nsol = solid[moff]; if (MH) nsol_hi = solid[moff_hi]; }
        for (int s = 1; s <= nsteps; ++s) {
            mbar_arrive_elect(&empty[(s - 1) % D], lane); // this row never reads a slot after its own step
            mbar_wait(&full[s % D], (unsigned) ((s / D) & 1));
            const double *ts = ring + (s % D) * SLOT + own;
            double cU[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) cU[k] = ts[k * FSTR];
            const unsigned csol = nsol, csol_hi = nsol_hi;
            if (BODY && s < nsteps) { moff += mplane; moff_hi += mplane; nsol = solid[moff]; if (MH) nsol_hi = solid[moff_hi]; }
            CellPrim q;
            derive_cell(cU, dc, q);
            double cFy[NF], cly;
            axis_flux<1>(q, cFy, cly);
            if (BODY) cly = body_mark(cly, csol);
            if (s > 1) mbar_wait(&barF[1], (unsigned) ((s - 2) & 1)); // row 1 is done with the previous record
#pragma unroll
            for (int k = 0; k < NF; ++k) r[k * 32] = cFy[k];
            r[NF * 32] = cly;
            mbar_arrive_elect(&barD[0], lane);
            if (MH) high_halo_step(s, csol_hi);
            if (lane == 0 && s + D - 1 <= nsteps + 1) produce(s + D - 1);
        }
        if (MH) lmax = yf_ok_hi ? lmy_hi : 0.0;
    } else if (!MH && row == NW - 1) {
        // ================= high halo row: computes the y face (j-1 | j) for the row below ==========
        for (int s = 1; s <= nsteps; ++s) {
            mbar_arrive_elect(&empty[(s - 1) % D], lane);
            mbar_wait(&full[s % D], (unsigned) ((s / D) & 1));
            unsigned csol_hi = 0;
            if (BODY) { moff_hi += mplane; csol_hi = solid[moff_hi]; }
            high_halo_step(s, csol_hi);
        }
        lmax = yf_ok_hi ? lmy_hi : 0.0;
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

        double *r_own = sm_r + row * 6 * 32 + lane;
        const double *r_dn = sm_r + (row - 1) * 6 * 32 + lane;
        double *f_own = sm_f + row * NF * 32 + lane;
        const double *f_up = sm_f + (row + 1) * NF * 32 + lane;
        const int un_own = (row - 1) * 32 + lane; // this lane's own cell in the U^n box (no clamp: stored cells only)

        const long long plane = (long long) g.py * g.px;
        const long long fs    = g.fs;
        double *op = Out + ((long long) (j + 1) * g.px + (i + XOFF)) + (long long) z0 * plane; // plane z0-1: the first store goes to z0

        double pFz[NF], plz, pS[NF];
        double pUc[NF]; // (CARRY only)
        double lmx = 0.0, lmy = 0.0, lmz = 0.0;
        const int xl_idx = min(max(min(max(i - 1, lc.ilo), lc.ihi) - i0, 0), 31) + ry * 32; // the left x neighbour's cell in a slot
        // (BODY) nsol: flag of the next step's cell, loaded a step ahead; pflag: of the previous step's (it is stored
        // only if that is 0)
        unsigned nsol = 0, pflag = 0;
        if (BODY) { pflag = solid[moff]; moff += mplane; nsol = solid[moff]; }
        // ---- step 0: plane z0-1 only provides the low side of the first z interface ------------------
        {
            mbar_wait(&full[0], 0u);
            const double *ts = ring + own;
            double c0[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) c0[k] = ts[k * FSTR];
            CellPrim q;
            derive_cell(c0, dc, q);
            axis_flux<2>(q, pFz, plz);
            if (BODY) plz = body_mark(plz, pflag);
#pragma unroll
            for (int k = 0; k < NF; ++k) pS[k] = 0.0;
#pragma unroll
            for (int k = 0; k < NF; ++k) pUc[k] = c0[k];
        }

#pragma unroll UNROLL
        for (int s = 1; s <= nsteps; ++s) {
            const int kz = z0 + s - 1;
            const unsigned par = (unsigned) ((s - 1) & 1);
            const double *ts = ring + (s % D) * SLOT;       // this plane
            const double *tp = ring + ((s - 1) % D) * SLOT; // the previous plane
            mbar_wait(&full[s % D], (unsigned) ((s / D) & 1));
            double cU[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) cU[k] = ts[k * FSTR + own];
            const unsigned csol = nsol;
            if (BODY) { moff += mplane; nsol = solid[moff]; } // plane z0+s: at most the ghost plane nz

            CellPrim q;
            derive_cell(cU, dc, q);

            // ---- y record for row+1 (the earlier it is out, the less row+1 waits) ---------------------
            double cFy[NF], cly;
            axis_flux<1>(q, cFy, cly);
            if (BODY) cly = body_mark(cly, csol);
#pragma unroll
            for (int k = 0; k < NF; ++k) r_own[k * 32] = cFy[k];
            r_own[NF * 32] = cly;
            mbar_arrive_elect(&barD[row], lane);

            // ---- z interface (kz-1 | kz): completes plane kz-1 ----------------------------------------
            double cFz[NF], clz, AFz[NF];
            axis_flux<2>(q, cFz, clz);
            if (BODY) clz = body_mark(clz, csol);
            {
                double pU[NF], pUn[NF];
#pragma unroll
                for (int k = 0; k < NF; ++k) pU[k] = CARRY ? pUc[k] : tp[k * FSTR + own];
                const double lam = llf_area_flux(pU, pFz, plz, cU, cFz, clz, Ah, AFz);
                lmz = (lam < lmz) ? lmz : lam;
                if (STAGE >= 2) {
#pragma unroll
                    for (int k = 0; k < NF; ++k) pUn[k] = tp[SIN + k * UFSTR + un_own];
                }
                finish_plane<STAGE>(pS, AFz, pU, pUn, dt, g.volume, dc.y_vol, op, fs, upd && s > 1 && (!BODY || pflag == 0), est_max);
                op += plane;
                if (MMF_T_EARLY_RELEASE) mbar_arrive_elect(&empty[(s - 1) % D], lane); // the previous plane's slot is no longer read by this warp
            }

            // ---- x interface (i-1 | i): lane-1's state by warp shuffle ---------------------------------
            double AFx[NF];
            {
                double cFx[NF], clx, lU[NF], lF[NF];
                axis_flux<0>(q, cFx, clx);
                if (BODY) clx = body_mark(clx, csol);
#pragma unroll
                for (int k = 0; k < NF; ++k) { lU[k] = MMF_T_XLDS ? ts[k * FSTR + xl_idx] : shfl_up_d(cU[k]); lF[k] = shfl_up_d(cFx[k]); }
                const double ll  = shfl_up_d(clx);
                const double lam = llf_area_flux(lU, lF, ll, cU, cFx, clx, Ah, AFx);
                lmx = (lam < lmx) ? lmx : lam;
            }

            // ---- y interface (j-1 | j): the lower cell's U from the slot, its flux from row-1's record ---
            double AFy[NF];
            {
                double lU[NF], lF[NF];
#pragma unroll
                for (int k = 0; k < NF; ++k) lU[k] = ts[k * FSTR + dn];
                mbar_wait(&barD[row - 1], par);
#pragma unroll
                for (int k = 0; k < NF; ++k) lF[k] = r_dn[k * 32];
                const double ll  = r_dn[NF * 32];
                const double lam = llf_area_flux(lU, lF, ll, cU, cFy, cly, Ah, AFy);
                lmy = (lam < lmy) ? lmy : lam;
                // row-1 published this record only after it had read this row's previous flux
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
            } else if (!edge && MMF_T_SELBR) {
                // the interface created last enters last; which of y and z that would be is warp-uniform (one row, one
                // plane per warp), whether x beats it differs from lane to lane: a uniform branch and a swap (a + b == b + a)
                const bool xl = key_x < min(key_y, key_z);
                if (key_y < key_z) { // (x + z) + y, or (y + z) + x
#pragma unroll
                    for (int k = 0; k < NF; ++k) S[k] = ((xl ? AFy[k] : AFx[k]) + AFz[k]) + (xl ? AFx[k] : AFy[k]);
                } else {             // (y + x) + z, or (y + z) + x
#pragma unroll
                    for (int k = 0; k < NF; ++k) S[k] = (AFy[k] + (xl ? AFz[k] : AFx[k])) + (xl ? AFx[k] : AFz[k]);
                }
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
                    double sacc = 0.0;
                    if (!bx) sacc += AFx[k];
                    if (key_y >= 0) sacc += AFy[k];
                    if (key_z >= 0) sacc += AFz[k];
                    if (bx) sacc += AFx[k];
                    S[k] = sacc;
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

            // ---- plane kz becomes the previous plane; -z_hi follows in the next step ------------------
#pragma unroll
            for (int k = 0; k < NF; ++k) { pS[k] = S[k]; pFz[k] = cFz[k]; }
            plz = clz;
            if (BODY) pflag = csol;
            if (CARRY) {
#pragma unroll
                for (int k = 0; k < NF; ++k) pUc[k] = cU[k];
            }
            if (!MMF_T_EARLY_RELEASE) mbar_arrive_elect(&empty[(s - 1) % D], lane); // the previous plane's slot is no longer read by this warp
        }

        // ---- last step: plane z1 only closes the last z interface -----------------------------------
        {
            const int s = nsteps + 1;
            const double *ts = ring + (s % D) * SLOT;
            const double *tp = ring + ((s - 1) % D) * SLOT;
            mbar_wait(&full[s % D], (unsigned) ((s / D) & 1));
            double cU[NF], pU[NF], pUn[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) { cU[k] = ts[k * FSTR + own]; pU[k] = CARRY ? pUc[k] : tp[k * FSTR + own]; }
            CellPrim q;
            derive_cell(cU, dc, q);
            double cFz[NF], clz, AFz[NF];
            axis_flux<2>(q, cFz, clz);
            if (BODY) clz = body_mark(clz, nsol);
            const double lam = llf_area_flux(pU, pFz, plz, cU, cFz, clz, Ah, AFz);
            lmz = (lam < lmz) ? lmz : lam;
            if (STAGE >= 2) {
#pragma unroll
                for (int k = 0; k < NF; ++k) pUn[k] = tp[SIN + k * UFSTR + un_own];
            }
            finish_plane<STAGE>(pS, AFz, pU, pUn, dt, g.volume, dc.y_vol, op, fs, upd && (!BODY || pflag == 0), est_max);
        }
        lmax = xf_ok ? lmx : 0.0;
        if (yf_ok) lmax = (lmy < lmax) ? lmax : lmy;
        if (zf_ok) lmax = (lmz < lmax) ? lmax : lmz;
        emax = est_max;
    }

    // (every copy that was issued has been waited for by the update rows: nothing is in flight into the ring)
    block_maxima<NW>(lmax, max_eig, emax, (STAGE == 3) ? cta_est : nullptr, tid.tile, sm_r);
}

} // namespace mmf
