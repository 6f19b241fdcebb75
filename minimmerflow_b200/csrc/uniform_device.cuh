// uniform_device.cuh -- device-side vocabulary of the uniform path: geometry and launch-argument structs,
// per-cell derived quantities, axis-specialised fluxes, the LLF interface flux.  Templates and inline
// functions only (no kernels), so that every stage-kernel translation unit can include it.
// uniform_kernels.cuh describes the layout and the work decomposition.
#pragma once

#include "generic_types.cuh"
#include "ptx_helpers.cuh"

#include <vector>

namespace mmf {

enum { NUM_MORTON = 0, NUM_LEXI = 1, NUM_AXIS = 2 };

struct UniformGeom {
    int nx, ny, nz;        // local box (cells)
    int gx0, gy0, gz0;     // lattice coordinate of the first local cell
    int gnx, gny, gnz;     // global lattice
    int px, py, pz;        // padded extents (px = row pitch)
    long long fs;          // field stride in doubles
    double h, area, volume;
    int bc[6];             // physical BC per side (-x,+x,-y,+y,-z,+z); -2 = partition boundary
    double dirichlet[NF];
};

// Bounds the stage kernels clamp their LOAD coordinates to.  The virtual state of BC_FREE_FLOW is a
// copy of the inner cell (src/euler.cpp:298-310), so on such a side the kernels simply read the
// boundary cell again instead of a ghost cell (ilo = 0 instead of -1, ...): no ghost pass between
// the stages, bitwise the same fluxes.
struct LoadClamp {
    int ilo, ihi, jlo, jhi, klo, khi;
};

// Multi-GPU, direct peer stores: instead of a wait kernel in front of a stage, the CTAs whose tile
// touches a partition side wait themselves (one thread per side spins on this rank's arrival
// counter), and the tiles are visited interior first (tile_order), so that by the time the first
// boundary tile is scheduled the neighbours' layers have long arrived: the exchange and the rank skew
// hide behind the interior work of the same kernel.
struct HaloWait {
    const unsigned long long *flags; // [6] arrival counters of this rank (nullptr: nothing to wait for)
    unsigned long long seq;          // value the counters must have reached for the array being read
    unsigned int mask;               // sides that have a neighbour rank
    const int *tile_order;           // [tx*ty*tz] tile visited by CTA b (1-D grid); nullptr: 3-D grid
    int tx, ty, tz;
    // a neighbour that died never raises its counter: a waiting thread gives up after timeout_ns (0 = never) and
    // counts itself in StepControl::halo_timeouts, which the host turns into an error when the call returns
    double *timeouts;
    unsigned long long timeout_ns;
};

// spin until the arrival counter reaches seq; false = gave up
__device__ __forceinline__ bool halo_spin(const volatile unsigned long long *f, unsigned long long seq, unsigned long long timeout_ns)
{
    if (*f >= seq) return true;
    const unsigned long long t0 = global_timer_ns();
    for (;;) {
#pragma unroll 1
        for (int n = 0; n < 256; ++n) {
            if (*f >= seq) return true;
        }
        if (timeout_ns && global_timer_ns() - t0 > timeout_ns) return false;
    }
}

// Multi-GPU, direct peer stores: ghost cells across an x partition side live in a COMPACT array
// [field][k+1][j+1] instead of the padded state array.  In the padded array the x ghost column is one
// 8-byte element per row, so a neighbour could only fill it with 8-byte stores 2 KB apart over NVLink
// (measured: 50 us per push against 20 us for a y or z layer, and it slowed the stage kernel running
// next to it).  The stage kernels' halo lanes (i = -1 / i = nx) simply take their column from here.
struct XGhost {
    const double *lo, *hi; // ghost columns of the -x / +x side for the array being read; nullptr: padded array
    long long fs;          // field stride
    int pitch;             // ny + 2
};

// Residual input loads go through L2 only: ghost layers are written by the neighbour GPUs while this
// kernel runs, and a non-coherent L1 line fetched earlier on the same SM could hold the old values
__device__ __forceinline__ double ldsin(const double *p) { return __ldcg(p); }

// Position of cell 0 in a padded row: the x ghost sits at column 0.  The x windows of the stage kernels start at
// cell bx*XW - 1 with XW even, i.e. at an even padded column, and rows are multiples of 32 bytes: the first byte of
// every box row a bulk tensor (TMA) load fetches lies on a 16-byte boundary, which the hardware demands.
constexpr int XOFF = 1;
__host__ __device__ constexpr int padded_row(int nx) { return (nx + XOFF + 1 + 3) / 4 * 4; }

__host__ __device__ __forceinline__ long long uoff(const UniformGeom &g, int i, int j, int k)
{
    return ((long long) (k + 1) * g.py + (j + 1)) * g.px + (i + XOFF);
}

// ---- per-cell derived quantities ---------------------------------------------------------------

struct CellPrim {
    double rho, u, v, w, p, H, a; // H = eto + p (src/euler.cpp:103,112), a = sqrt(GAMMA*T) (:61)
};

struct DivConsts {
    double y_gm1, y_c1, y_vol; // reciprocals of GAMMA-1, 2/(GAMMA-1) and the cell volume
};

// conservative2primitive (src/utils.cpp:48-63) + the per-side part of evalSplitting/evalFluxes
// (src/euler.cpp:45-63, 85-103), evaluated once per cell
__device__ __forceinline__ void derive_cell(const double *c, const DivConsts &dc, CellPrim &q)
{
    const double rho = c[FID_RHO];
    const double y   = rcp_nr(rho);
    const double rr  = rho * rho;
    const double yrr = rcp_nr(rr);
    const double K = div_nr(c[FID_RHO_U] * c[FID_RHO_U] + c[FID_RHO_V] * c[FID_RHO_V] + c[FID_RHO_W] * c[FID_RHO_W], rr, yrr);
    const double T = div_nr(div_nr(2.0 * c[FID_RHO_E], rho, y) - K, TWO_OVER_GM1, dc.y_c1);
    q.rho = rho;
    q.u = div_nr(c[FID_RHO_U], rho, y);
    q.v = div_nr(c[FID_RHO_V], rho, y);
    q.w = div_nr(c[FID_RHO_W], rho, y);
    q.p = rho * T;
    const double vel2 = q.u * q.u + q.v * q.v + q.w * q.w;
    const double eto  = div_nr(q.p, GM1, dc.y_gm1) + 0.5 * rho * vel2;
    q.H = eto + q.p;
    q.a = sqrt(GAMMA * T);
}

// derive_cell with the reciprocals of rho and rho*rho -- the root of every dependency chain of a cell --
// supplied by the caller, who can start them as soon as rho is known (same operations, same bits)
__device__ __forceinline__ void derive_cell_pre(const double *c, const double y, const double yrr, const DivConsts &dc, CellPrim &q)
{
    const double rho = c[FID_RHO];
    const double rr  = rho * rho;
    const double K = div_nr(c[FID_RHO_U] * c[FID_RHO_U] + c[FID_RHO_V] * c[FID_RHO_V] + c[FID_RHO_W] * c[FID_RHO_W], rr, yrr);
    const double T = div_nr(div_nr(2.0 * c[FID_RHO_E], rho, y) - K, TWO_OVER_GM1, dc.y_c1);
    q.rho = rho;
    q.u = div_nr(c[FID_RHO_U], rho, y);
    q.v = div_nr(c[FID_RHO_V], rho, y);
    q.w = div_nr(c[FID_RHO_W], rho, y);
    q.p = rho * T;
    const double vel2 = q.u * q.u + q.v * q.v + q.w * q.w;
    const double eto  = div_nr(q.p, GM1, dc.y_gm1) + 0.5 * rho * vel2;
    q.H = eto + q.p;
    q.a = sqrt(GAMMA * T);
}

// evalFluxes with n = +e_AXIS (src/euler.cpp:105-112): u*1 + v*0 + w*0 == u and p*0 == +0 exactly
template <int AXIS>
__device__ __forceinline__ void axis_flux(const CellPrim &q, double *F, double &lam)
{
    const double un = (AXIS == 0) ? q.u : (AXIS == 1) ? q.v : q.w;
    const double m  = q.rho * un;
    F[0] = m;
    F[1] = (AXIS == 0) ? m * q.u + q.p : m * q.u;
    F[2] = (AXIS == 1) ? m * q.v + q.p : m * q.v;
    F[3] = (AXIS == 2) ? m * q.w + q.p : m * q.w;
    F[4] = un * q.H;
    lam  = fabs(un) + q.a; // src/euler.cpp:60-66
}

// LLF splitting (src/euler.cpp:68-72) times the interface area (:239, :245).  Ah = 0.5*area:
// A*(0.5*x) == (0.5*A)*x bit for bit because scaling by a power of two commutes with rounding.
__device__ __forceinline__ double llf_area_flux(const double *UL, const double *FL, double lamL,
                                                const double *UR, const double *FR, double lamR,
                                                double Ah, double *AF)
{
    const double lam = (lamR < lamL) ? lamL : lamR; // std::max(lambdaR, lambdaL)
#pragma unroll
    for (int k = 0; k < NF; ++k) {
        AF[k] = Ah * ((FR[k] + FL[k]) - lam * (UR[k] - UL[k]));
    }
    return lam;
}

__device__ __forceinline__ double shfl_down_d(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_up_d(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }

// ---- the fused stage kernels live in uniform_stage_v3.cuh / uniform_stage_v5.cuh -----------------
// STAGE 0: RHS only (euler::computeRHS).   Out = RHS array.
// STAGE 1: W  = U + dt*R(U)/V                         Sin = U,  Out = Wa
// STAGE 2: W' = 0.75*U + 0.25*(W + dt*R(W)/V)         Sin = Wa, Un = U, Out = Wb
// STAGE 3: U' = (1./3)*U + (2./3)*(W' + dt*R(W')/V)   Sin = Wb, Un = U, Out = U (in place, pointwise)
// ORDER: interface numbering convention deciding the per-cell accumulation order (NUM_*).
constexpr int XW = 30; // cells updated per warp row (32-lane window, 2 overlap)

// ---- a box with bodies: what one cell contributes to the max eigenvalue over the processed interfaces --------
// (uniform_eig_body_kernel in uniform_kernels.cuh explains; solid = one flag per padded cell, 1 = not solved)
__device__ __forceinline__ double eig_body_cell(const UniformGeom &g, const double *__restrict__ Sin,
                                                const unsigned char *__restrict__ solid, const int i, const int j, const int k)
{
    const long long o = uoff(g, i, j, k);
    if (solid[o] == 1) return 0.0; // (2 = a fluid cell next to a wall)
    DivConsts dc;
    dc.y_gm1 = rcp_nr(GM1); dc.y_c1 = rcp_nr(TWO_OVER_GM1); dc.y_vol = 0.0;
    double c[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) c[f] = Sin[f * g.fs + o];
    CellPrim pr;
    derive_cell(c, dc, pr);
    double lmax = fmax(fmax(fabs(pr.u), fabs(pr.v)), fabs(pr.w)) + pr.a;
    const int ijk[3] = { i, j, k }, ext[3] = { g.nx, g.ny, g.nz };
    const long long step[3] = { 1, g.px, (long long) g.py * g.px };
    for (int side = 0; side < 6; ++side) {
        const int axis = side >> 1;
        const bool hi = side & 1;
        const long long on = hi ? o + step[axis] : o - step[axis];
        const bool border = hi ? (ijk[axis] == ext[axis] - 1) : (ijk[axis] == 0);
        double v[NF];
        if (border) {
            // the ghost cell holds the boundary condition's virtual state; a free-flow ghost is a copy
            if (g.bc[side] == BC_FREE_FLOW || g.bc[side] < 0) continue;
#pragma unroll
            for (int f = 0; f < NF; ++f) v[f] = Sin[f * g.fs + on];
        } else {
            if (solid[on] != 1) continue;
            // wall: interface normal +e_axis with the low cell as owner; the BC sees it from the fluid side
            double bn[3] = { 0.0, 0.0, 0.0 };
            bn[axis] = 1.0;
            if (!hi) { bn[0] = -1. * bn[0]; bn[1] = -1. * bn[1]; bn[2] = -1. * bn[2]; }
            interface_bc_values(BC_WALL, bn, nullptr, c, v);
        }
        CellPrim pv;
        derive_cell(v, dc, pv);
        const double lam = fabs(axis == 0 ? pv.u : axis == 1 ? pv.v : pv.w) + pv.a;
        lmax = (lam < lmax) ? lmax : lam;
    }
    return lmax;
}

// ---- a box with bodies, fix-up formulation (kernel form 'b') ------------------------------------------------
// Flag values: 0 = fluid, 1 = not solved (solid), 2 = fluid cell with at least one wall interface.  The stage
// kernel proper treats a wall like an ordinary interface (its result for a flag-2 cell is meaningless and is
// not stored); those cells -- a surface -- are recomputed here, one thread per cell, reference-shaped: the
// cell's six interfaces in the reference's processing order (src/euler.cpp:153-248; interior low faces by
// descending creator key, then (-a if border) +a per axis, uniform_stage_v5.cuh), each with
// euler::evalSplitting's own operations, wall and border sides from euler::evalInterfaceBCValues, then the RK
// stage of src/main.cpp:409-495.  Returns the largest interface eigenvalue met.
template <int STAGE, int ORDER>
__device__ __forceinline__ double wall_cell_update(const UniformGeom &g, const LoadClamp &lc, const double *__restrict__ Sin,
                                                   const double *__restrict__ Un, const unsigned char *__restrict__ flag,
                                                   const long long o, const double dt, double *out)
{
    const long long plane = (long long) g.py * g.px;
    const int k = (int) (o / plane) - 1, j = (int) ((o % plane) / g.px) - 1, i = (int) (o % g.px) - XOFF;
    const int ijk[3] = { i, j, k }, gijk[3] = { g.gx0 + i, g.gy0 + j, g.gz0 + k }, ext[3] = { g.nx, g.ny, g.nz };
    const long long step[3] = { 1, g.px, plane };
    // processing order of the six face slots (0 -x, 1 +x, 2 -y, 3 +y, 4 -z, 5 +z)
    int order[6], n = 0;
    if (ORDER == NUM_AXIS) {
        for (int s = 0; s < 6; ++s) order[n++] = s;
    } else {
        int lows[3], keys[3], nl = 0;
        for (int a = 0; a < 3; ++a) {
            if (gijk[a] == 0) continue;
            lows[nl] = a;
            keys[nl] = (ORDER == NUM_LEXI) ? a : 3 * (__ffs(gijk[a]) - 1) + a;
            nl++;
        }
        for (int a = 0; a < nl; ++a)
            for (int b = a + 1; b < nl; ++b)
                if (keys[b] > keys[a]) { const int tk = keys[a]; keys[a] = keys[b]; keys[b] = tk; const int tl = lows[a]; lows[a] = lows[b]; lows[b] = tl; }
        for (int a = 0; a < nl; ++a) order[n++] = 2 * lows[a];
        for (int a = 0; a < 3; ++a) {
            if (gijk[a] == 0) order[n++] = 2 * a;
            order[n++] = 2 * a + 1;
        }
    }
    double c[NF], acc[NF] = { 0., 0., 0., 0., 0. };
#pragma unroll
    for (int f = 0; f < NF; ++f) c[f] = Sin[f * g.fs + o];
    double lmax = 0.0;
    for (int q = 0; q < 6; ++q) {
        const int side = order[q], axis = side >> 1;
        const bool hi = side & 1;
        const long long on = hi ? o + step[axis] : o - step[axis];
        const bool border = hi ? (ijk[axis] == ext[axis] - 1) : (ijk[axis] == 0);
        double nrm[3] = { 0.0, 0.0, 0.0 };
        double other[NF], flux[NF], lambda;
        bool cell_is_owner;
        if (border) {
            // the cell owns its border interface, outward normal; the ghost cell holds the virtual state, except
            // on a free-flow side inside a fused step, where it is the copy the stage kernels never read
            nrm[axis] = hi ? 1.0 : -1.0;
            cell_is_owner = true;
            if (g.bc[side] == BC_FREE_FLOW) {
#pragma unroll
                for (int f = 0; f < NF; ++f) other[f] = c[f];
            } else {
#pragma unroll
                for (int f = 0; f < NF; ++f) other[f] = Sin[f * g.fs + on];
            }
        } else {
            nrm[0] = 0.0; nrm[axis] = 1.0;          // interior: owner = the low cell, normal +e_axis
            cell_is_owner = hi;
            if (flag[on] == 1) {
                double bn[3] = { nrm[0], nrm[1], nrm[2] };
                if (!cell_is_owner) { bn[0] = -1. * nrm[0]; bn[1] = -1. * nrm[1]; bn[2] = -1. * nrm[2]; }
                interface_bc_values(BC_WALL, bn, nullptr, c, other);
            } else {
#pragma unroll
                for (int f = 0; f < NF; ++f) other[f] = Sin[f * g.fs + on];
            }
        }
        if (cell_is_owner) eval_splitting(c, other, nrm, flux, &lambda);
        else               eval_splitting(other, c, nrm, flux, &lambda);
        lmax = (lambda < lmax) ? lmax : lambda;
        if (cell_is_owner) {
#pragma unroll
            for (int f = 0; f < NF; ++f) acc[f] -= g.area * flux[f];
        } else {
#pragma unroll
            for (int f = 0; f < NF; ++f) acc[f] += g.area * flux[f];
        }
    }
    (void) lc;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        if (STAGE == 0) {
            out[f] = acc[f];
        } else {
            const double qd = dt * acc[f] / g.volume;
            const double un = (STAGE >= 2) ? Un[f * g.fs + o] : 0.0;
            if (STAGE == 1)      out[f] = c[f] + qd;
            else if (STAGE == 2) out[f] = 0.75 * un + 0.25 * (c[f] + qd);
            else                 out[f] = (1. / 3) * un + (2. / 3) * (c[f] + qd);
        }
    }
    return lmax;
}

// ---- a box with bodies: the flag array and the wall-cell list, host side --------------------------------------
// One flag per padded cell (layout of one field): 1 = not solved (src/main.cpp:221-237).  The ghost shell repeats
// the flag of the cell it touches, so that the border interface of an unsolved cell is skipped like the
// reference skips it (src/euler.cpp:181-183).  mark_walls (kernel form 'b'): fluid cells with at least one wall
// interface get flag 2 and are listed by padded offset, ascending; their ghost cells keep 0.
// Plain host code (no CUDA calls), shared with tools/emu so that it is unit-tested on the CPU.
inline void body_flags(const UniformGeom &g, const long long n_cells, const int *cell_ijk, const unsigned char *solved,
                       const bool mark_walls, std::vector<unsigned char> &flag, std::vector<int> &walls)
{
    const int nx = g.nx, ny = g.ny, nz = g.nz;
    flag.assign((size_t) g.fs, 0);
    walls.clear();
    for (long long c = 0; c < n_cells; ++c) {
        if (!solved[c]) flag[(size_t) uoff(g, cell_ijk[3 * c], cell_ijk[3 * c + 1], cell_ijk[3 * c + 2])] = 1;
    }
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j) {
            flag[(size_t) uoff(g, -1, j, k)] = flag[(size_t) uoff(g, 0, j, k)];
            flag[(size_t) uoff(g, nx, j, k)] = flag[(size_t) uoff(g, nx - 1, j, k)];
        }
    for (int k = 0; k < nz; ++k)
        for (int i = 0; i < nx; ++i) {
            flag[(size_t) uoff(g, i, -1, k)] = flag[(size_t) uoff(g, i, 0, k)];
            flag[(size_t) uoff(g, i, ny, k)] = flag[(size_t) uoff(g, i, ny - 1, k)];
        }
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            flag[(size_t) uoff(g, i, j, -1)] = flag[(size_t) uoff(g, i, j, 0)];
            flag[(size_t) uoff(g, i, j, nz)] = flag[(size_t) uoff(g, i, j, nz - 1)];
        }
    if (!mark_walls) return;
    const long long step[3] = { 1, g.px, (long long) g.py * g.px };
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const long long o = uoff(g, i, j, k);
                if (flag[(size_t) o] == 1) continue;
                const int ijk[3] = { i, j, k }, ext[3] = { nx, ny, nz };
                bool wall = false;
                for (int a = 0; a < 3; ++a) {
                    if (ijk[a] > 0 && flag[(size_t) (o - step[a])] == 1) wall = true;
                    if (ijk[a] < ext[a] - 1 && flag[(size_t) (o + step[a])] == 1) wall = true;
                }
                if (wall) walls.push_back((int) o);
            }
    for (int o : walls) flag[(size_t) o] = 2;
}

} // namespace mmf
