// generic_kernels.cuh -- connectivity-driven kernels that work on ANY mesh the host describes
// (non-uniform octree levels, bodies / BC_WALL, BC_DIRICHLET, 2-D).
//
// Residual assembly is a CELL GATHER: one thread owns one cell and walks that cell's interfaces in
// the host's interface processing order, applying `-=` on the owner side and `+=` on the neighbour
// side.  That is the order in which the reference's scatter loop (src/euler.cpp:153-248) touches
// the cell, so the RHS is reproduced bit for bit without atomics.
#pragma once

#include "generic_types.cuh"
#include "ptx_helpers.cuh"

namespace mmf {

// resident blocks of 128 threads per SM the fused generic kernels are compiled for (5: up to 102 registers; measured: 4 = 0.463, 5 = 0.434, 6 = 0.399 ms per step on the 2-D 1024^2 vortex, but 6 loses 6 % on a two-level octree)
#ifndef MMF_GEN_MINBLOCKS
#define MMF_GEN_MINBLOCKS 5
#endif
// ... and for a 2-D mesh (four entries per cell instead of six and more): 6 blocks, 85 registers, the faster shape there
#ifndef MMF_GEN_MINBLOCKS_2D
#define MMF_GEN_MINBLOCKS_2D 6
#endif

// ---- host AoS (raw order, [c*5+k]) <-> device SoA ([k*stride+c]) -------------------------------

__global__ void __launch_bounds__(256) aos_to_soa_kernel(const double *__restrict__ aos, double *__restrict__ soa,
                                                         int64_t n_cells, int64_t stride)
{
    // coalesced read of the AoS stream through shared memory, coalesced write per field
    __shared__ double tile[256 * NF];
    const int64_t c0 = (int64_t) blockIdx.x * 256;
    const int64_t n  = min((int64_t) 256, n_cells - c0);
    for (int i = threadIdx.x; i < n * NF; i += 256) tile[i] = aos[c0 * NF + i];
    __syncthreads();
    if (threadIdx.x < n) {
#pragma unroll
        for (int k = 0; k < NF; ++k) soa[k * stride + c0 + threadIdx.x] = tile[threadIdx.x * NF + k];
    }
}

__global__ void __launch_bounds__(256) soa_to_aos_kernel(const double *__restrict__ soa, double *__restrict__ aos,
                                                         int64_t n_cells, int64_t stride)
{
    __shared__ double tile[256 * NF];
    const int64_t c0 = (int64_t) blockIdx.x * 256;
    const int64_t n  = min((int64_t) 256, n_cells - c0);
    if (threadIdx.x < n) {
#pragma unroll
        for (int k = 0; k < NF; ++k) tile[threadIdx.x * NF + k] = soa[k * stride + c0 + threadIdx.x];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * NF; i += 256) aos[c0 * NF + i] = tile[i];
}

// utils::conservative2primitive (src/utils.cpp:48-63) over an AoS stream, in place: what main.cpp does for
// every cell before mesh.write() (src/main.cpp:511-518).  One cell per thread through a shared-memory tile,
// so that the 40-byte records are read and written as one contiguous stream per block.
__global__ void __launch_bounds__(256) aos_cons_to_prim_kernel(double *__restrict__ aos, int64_t n_cells)
{
    __shared__ double tile[256 * NF];
    const int64_t c0 = (int64_t) blockIdx.x * 256;
    const int64_t n  = min((int64_t) 256, n_cells - c0);
    for (int i = threadIdx.x; i < n * NF; i += 256) tile[i] = aos[c0 * NF + i];
    __syncthreads();
    if (threadIdx.x < n) {
        double c[NF], p[NF];
#pragma unroll
        for (int k = 0; k < NF; ++k) c[k] = tile[threadIdx.x * NF + k];
        conservative2primitive(c, p);
#pragma unroll
        for (int k = 0; k < NF; ++k) tile[threadIdx.x * NF + k] = p[k];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * NF; i += 256) aos[c0 * NF + i] = tile[i];
}

// ---- euler::computeRHS (src/euler.cpp:127-249) --------------------------------------------------

__device__ __forceinline__ void load_cell(const double *__restrict__ S, int64_t stride, int64_t c, double *u)
{
#pragma unroll
    for (int k = 0; k < NF; ++k) u[k] = S[k * stride + c];
}

// ---- one cell's residual with the cell's own derived state computed ONCE ------------------------------------
// In generic_rhs_kernel below every interface converts both of its cells to primitives (euler::evalSplitting,
// src/euler.cpp:42-73): a cell is converted once per interface it touches.  The cell a thread gathers for is a
// side of every interface in its list, so the fused stage kernel (generic_stage_kernel) derives it once --
// primitives (src/utils.cpp:48-63), sound speed, |v|^2 and total energy as src/euler.cpp:52-58, 83-92 form them
// -- and per interface derives the other side only.  Same inputs, same operations, same order => the same bits
// (checked bit for bit on the CPU by tests/test_emu_generic.py); about a third fewer FP64
// instructions per cell.  generic_rhs_kernel keeps the reference-shaped text: its SASS is the one the GPU parity
// runs validated.
struct DerivedCell {
    double cons[NF], prim[NF];
    double a;    // sqrt(GAMMA * T)                   (src/euler.cpp:54, :58)
    double vel2; // u*u + v*v + w*w                   (src/euler.cpp:89)
    double eto;  // p / (GAMMA-1) + 0.5 * rho * vel2   (src/euler.cpp:103)
};

// reciprocals of the two constants conservative2primitive / evalFluxes divide by, once per thread
struct GenericDivConsts {
    double y_gm1, y_c1; // 1 / (GAMMA-1), 1 / (2/(GAMMA-1))
};

__device__ __forceinline__ GenericDivConsts generic_div_consts()
{
    GenericDivConsts k;
    k.y_gm1 = rcp_nr(GM1);
    k.y_c1  = rcp_nr(TWO_OVER_GM1);
    return k;
}

// conservative2primitive + sound speed + |v|^2 + total energy of one cell.  The seven IEEE divisions share their
// reciprocals (ptx_helpers.cuh: a / b == div_nr(a, b, rcp_nr(b)) bit for bit, mmf_selftest_division): four quotients
// by rho cost one reciprocal, the two constants' reciprocals come from the caller -- the same construction as
// derive_cell of the uniform path, a third of the FP64 instructions of seven stand-alone divisions.
__device__ __forceinline__ void derive_cell_generic(DerivedCell &d, const GenericDivConsts &k)
{
    const double rho = d.cons[FID_RHO];
    const double y   = rcp_nr(rho);
    const double rr  = rho * rho;
    const double yrr = rcp_nr(rr);
    // src/utils.cpp:48-63
    const double K = div_nr(d.cons[FID_RHO_U] * d.cons[FID_RHO_U] + d.cons[FID_RHO_V] * d.cons[FID_RHO_V] +
                            d.cons[FID_RHO_W] * d.cons[FID_RHO_W], rr, yrr);
    d.prim[FID_T] = div_nr(div_nr(2.0 * d.cons[FID_RHO_E], rho, y) - K, TWO_OVER_GM1, k.y_c1);
    d.prim[FID_U] = div_nr(d.cons[FID_RHO_U], rho, y);
    d.prim[FID_V] = div_nr(d.cons[FID_RHO_V], rho, y);
    d.prim[FID_W] = div_nr(d.cons[FID_RHO_W], rho, y);
    d.prim[FID_P] = rho * d.prim[FID_T];
    d.a = sqrt(GAMMA * d.prim[FID_T]);
    const double u = d.prim[FID_U], v = d.prim[FID_V], w = d.prim[FID_W];
    d.vel2 = u * u + v * v + w * w;
    d.eto  = div_nr(d.prim[FID_P], GM1, k.y_gm1) + 0.5 * rho * d.vel2;
}

// euler::evalFluxes (src/euler.cpp:83-113) from a derived cell; un is the normal velocity evalSplitting
// evaluates a second time with the same operands (src/euler.cpp:52, :56)
__device__ __forceinline__ void fluxes_of_derived(const DerivedCell &d, const double *n, double *flux, double *un_out)
{
    const double un = normal_velocity(d.prim, n);
    const double p  = d.prim[FID_P];
    const double massFlux = d.cons[FID_RHO] * un;
    flux[0] = massFlux;
    flux[1] = massFlux * d.prim[FID_U] + p * n[0];
    flux[2] = massFlux * d.prim[FID_V] + p * n[1];
    flux[3] = massFlux * d.prim[FID_W] + p * n[2];
    flux[4] = un * (d.eto + p);
    *un_out = un;
}

// euler::evalInterfaceBCValues (src/euler.cpp:261-376) from the derived inner cell: the reflecting / wall branch
// starts from the inner cell's primitives (:326), which are the ones already held
__device__ __forceinline__ void bc_values_of_derived(int bc, const double *normal, const double *dirichlet_info,
                                                     const DerivedCell &in, double *cons_bc)
{
    if (bc == BC_FREE_FLOW) {
#pragma unroll
        for (int k = 0; k < NF; ++k) cons_bc[k] = in.cons[k];
    } else if (bc == BC_REFLECTING || bc == BC_WALL) {
        double prim[NF];
        const double un = normal_velocity(in.prim, normal);
        const double n0 = un * normal[0], n1 = un * normal[1], n2 = un * normal[2];
        prim[FID_P] = in.prim[FID_P];
        prim[FID_T] = in.prim[FID_T];
        prim[FID_U] = in.prim[FID_U] - 2 * n0;
        prim[FID_V] = in.prim[FID_V] - 2 * n1;
        prim[FID_W] = in.prim[FID_W] - 2 * n2;
        primitive2conservative(prim, cons_bc);
    } else if (bc == BC_DIRICHLET) {
        primitive2conservative(dirichlet_info, cons_bc);
    }
}

// One cell's residual: its interfaces in the reference's processing order, `-=` on the owner side and `+=` on
// the neighbour side (src/euler.cpp:150-247); lmax = running maximum of the interface eigenvalues (:234).
// Only solved cells have a list, so on a boundary / wall interface the gathering cell is the fluid side
// (src/euler.cpp:198-225): as the neighbour it evaluates the BC with the flipped normal.
__device__ __forceinline__ void generic_cell_residual(const GenericMesh &m, const double *__restrict__ S, int64_t c,
                                                      double *acc, double &lmax)
{
    const int64_t e0 = m.cf_ptr[c], e1 = m.cf_ptr[c + 1];
    if (e0 == e1) return;
    const GenericDivConsts dk = generic_div_consts();
    DerivedCell own;
    load_cell(S, m.stride, c, own.cons);
    derive_cell_generic(own, dk);
    for (int64_t e = e0; e < e1; ++e) {
        const int32_t ent  = m.cf_ent[e];
        const int32_t f    = ent >> 1;
        const int     side = ent & 1;
        const int     bc   = m.f_bc[f];
        const double  A    = m.f_area[f];
        const double  nrm[3] = { m.f_normal[f], m.f_normal[m.n_ifaces + f], m.f_normal[2 * m.n_ifaces + f] };

        DerivedCell other;
        if (bc == BC_NONE) {
            // order-1 reconstruction: face state = cell mean (src/reconstruction.cpp:90-97)
            load_cell(S, m.stride, side == 0 ? m.f_neigh[f] : m.f_owner[f], other.cons);
        } else if (side == 0) {
            bc_values_of_derived(bc, nrm, m.dirichlet_info, own, other.cons);
        } else {
            const double flipped[3] = { -1. * nrm[0], -1. * nrm[1], -1. * nrm[2] };
            bc_values_of_derived(bc, flipped, m.dirichlet_info, own, other.cons);
        }
        derive_cell_generic(other, dk);

        // euler::evalSplitting (src/euler.cpp:42-73) with L = owner, R = neighbour and the un-flipped normal (:232)
        double fOwn[NF], fOth[NF], unOwn, unOth;
        fluxes_of_derived(own, nrm, fOwn, &unOwn);
        fluxes_of_derived(other, nrm, fOth, &unOth);
        const double lamOwn = fabs(unOwn) + own.a;
        const double lamOth = fabs(unOth) + other.a;
        double lam;
        if (side == 0) lam = (lamOth < lamOwn) ? lamOwn : lamOth; // std::max(lambdaR, lambdaL), L = own
        else           lam = (lamOwn < lamOth) ? lamOth : lamOwn; // L = other
        lmax = (lam < lmax) ? lmax : lam;                         // :234

        if (side == 0) {
#pragma unroll
            for (int k = 0; k < NF; ++k) {
                const double flux = 0.5 * ((fOth[k] + fOwn[k]) - lam * (other.cons[k] - own.cons[k]));
                acc[k] -= A * flux;                               // :237-241
            }
        } else {
#pragma unroll
            for (int k = 0; k < NF; ++k) {
                const double flux = 0.5 * ((fOwn[k] + fOth[k]) - lam * (own.cons[k] - other.cons[k]));
                acc[k] += A * flux;                               // :243-247
            }
        }
    }
}

__global__ void __launch_bounds__(128) generic_rhs_kernel(GenericMesh m, const double *__restrict__ S,
                                                          double *__restrict__ RHS, double *__restrict__ max_eig)
{
    const int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    double lmax = 0.0;
    if (c < m.n_cells) {
        // RHS starts at zero for every cell, solved or not (src/euler.cpp:135-148)
        double acc[NF] = { 0., 0., 0., 0., 0. };
        const int64_t e0 = m.cf_ptr[c], e1 = m.cf_ptr[c + 1];
        for (int64_t e = e0; e < e1; ++e) {
            const int32_t ent  = m.cf_ent[e];
            const int32_t f    = ent >> 1;
            const int     side = ent & 1;
            const int32_t o    = m.f_owner[f];
            const int32_t nb   = m.f_neigh[f];
            const int     bc   = m.f_bc[f];
            const double  A    = m.f_area[f];
            const double  nrm[3] = { m.f_normal[f], m.f_normal[m.n_ifaces + f], m.f_normal[2 * m.n_ifaces + f] };

            double ownerRec[NF], neighRec[NF];
            if (bc == BC_NONE) {
                // order-1 reconstruction: face state = cell mean (src/reconstruction.cpp:90-97)
                load_cell(S, m.stride, o, ownerRec);
                load_cell(S, m.stride, nb, neighRec);
            } else {
                // src/euler.cpp:198-225: the fluid side is the owner when it is solved, otherwise
                // the neighbour; the flipped normal is used for the BC evaluation only
                const bool ownerSolved = m.c_solved[o] != 0;
                if (ownerSolved) {
                    load_cell(S, m.stride, o, ownerRec);
                    interface_bc_values(bc, nrm, m.dirichlet_info, ownerRec, neighRec);
                } else {
                    const double flipped[3] = { -1. * nrm[0], -1. * nrm[1], -1. * nrm[2] };
                    load_cell(S, m.stride, nb, neighRec);
                    interface_bc_values(bc, flipped, m.dirichlet_info, neighRec, ownerRec);
                }
            }

            double flux[NF], lambda;
            eval_splitting(ownerRec, neighRec, nrm, flux, &lambda); // un-flipped normal (:232)
            lmax = (lambda < lmax) ? lmax : lambda;                  // :234

            if (side == 0) {
#pragma unroll
                for (int k = 0; k < NF; ++k) acc[k] -= A * flux[k];  // :237-241
            } else {
#pragma unroll
                for (int k = 0; k < NF; ++k) acc[k] += A * flux[k];  // :243-247
            }
        }
#pragma unroll
        for (int k = 0; k < NF; ++k) RHS[k * m.stride + c] = acc[k];
    }
    block_max_to_global(lmax, max_eig);
}

// ---- time-step control block, device resident --------------------------------------------------
// Lets a whole RK3 step (and a batch of steps) be enqueued without any host round trip:
// dt is chosen on the device from the stage-1 max eigenvalue exactly like src/main.cpp:398-402,
// and steps enqueued past t_max switch themselves off through `active`.

// ---- RK stage loops (src/main.cpp:409-423, 445-459, 481-495) -----------------------------------

template <int STAGE>
__global__ void __launch_bounds__(256) generic_rk_kernel(int64_t n_cells, int64_t stride,
                                                         const uint8_t *__restrict__ update,
                                                         const double *__restrict__ volume,
                                                         const StepControl *__restrict__ ctl,
                                                         double *__restrict__ U, double *__restrict__ W,
                                                         const double *__restrict__ RHS)
{
    const int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (ctl->active == 0.0) return;
    if (c >= n_cells || !update[c]) return;
    const double dt = ctl->dt;
    const double V  = volume[c];
#pragma unroll
    for (int k = 0; k < NF; ++k) {
        const int64_t i = k * stride + c;
        const double  q = dt * RHS[i] / V;
        if (STAGE == 1) {
            W[i] = U[i] + q;
        } else if (STAGE == 2) {
            W[i] = 0.75 * U[i] + 0.25 * (W[i] + q);
        } else {
            U[i] = (1. / 3) * U[i] + (2. / 3) * (W[i] + q);
        }
    }
}

// Residual and RK stage in one pass for stages 2 and 3 (the default of a fused step; MMF_GENERIC_FUSED=0 turns it off): every thread forms its
// cell's residual exactly like generic_rhs_kernel and applies the stage loop of src/main.cpp:445-459 / 481-495
// to it right away, so the residual is neither written nor read back between the two.  Stage 1 stays unfused:
// its dt is chosen from its own residual's face maximum (src/main.cpp:398-402).
//   stage 2: Out = the other work array (neighbours still read Sin); cells the loop skips carry Sin over
//   stage 3: Out = Un = U, in place (a cell's U^n is read by that cell only); the residual is ALSO stored,
//            because cellRHS of the reference holds the stage-3 residual when a step ends (it is written
//            out with the solution, src/main.cpp:284-298)
// A step switched off on the device (ctl->active == 0) updates nothing, like generic_rk_kernel.
template <int STAGE, int MINBLOCKS = MMF_GEN_MINBLOCKS>
__global__ void __launch_bounds__(128, MINBLOCKS) generic_stage_kernel(GenericMesh m, const double *__restrict__ Sin,
                                                            const double *Un, double *Out, double *__restrict__ RHS,
                                                            const StepControl *__restrict__ ctl,
                                                            double *__restrict__ max_eig)
{
    static_assert(STAGE == 2 || STAGE == 3, "fused generic stages are 2 and 3");
    const int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    double lmax = 0.0;
    if (c < m.n_cells) {
        double acc[NF] = { 0., 0., 0., 0., 0. };
        generic_cell_residual(m, Sin, c, acc, lmax);
        const bool upd = ctl->active != 0.0 && m.c_update[c];
        const double dt = ctl->dt;
        const double V  = m.c_volume[c];
#pragma unroll
        for (int k = 0; k < NF; ++k) {
            const int64_t i = k * m.stride + c;
            if (STAGE == 3) RHS[i] = acc[k];
            if (upd) {
                const double q = dt * acc[k] / V;
                if (STAGE == 2) Out[i] = 0.75 * Un[i] + 0.25 * (Sin[i] + q);
                else            Out[i] = (1. / 3) * Un[i] + (2. / 3) * (Sin[i] + q);
            } else if (STAGE == 2) {
                Out[i] = Sin[i];
            }
        }
    }
    block_max_to_global(lmax, max_eig);
}

// The stage-1 residual of the fused sequence: generic_rhs_kernel's result from generic_cell_residual (the
// gathering cell derived once).  Stage 1 itself stays two kernels, its dt comes out of this one's maximum.
template <int MINBLOCKS = MMF_GEN_MINBLOCKS>
__global__ void __launch_bounds__(128, MINBLOCKS) generic_rhs_derived_kernel(GenericMesh m, const double *__restrict__ S,
                                                                  double *__restrict__ RHS, double *__restrict__ max_eig)
{
    const int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    double lmax = 0.0;
    if (c < m.n_cells) {
        double acc[NF] = { 0., 0., 0., 0., 0. };
        generic_cell_residual(m, S, c, acc, lmax);
#pragma unroll
        for (int k = 0; k < NF; ++k) RHS[k * m.stride + c] = acc[k];
    }
    block_max_to_global(lmax, max_eig);
}

// Runs after the stage-1 residual: `while (t < tMax)` test (src/main.cpp:377) and the dt choice
// dt = 0.9*cfl*minCellSize/maxEig; if (t+dt > tMax) dt = tMax-t  (src/main.cpp:398-402).
__global__ void choose_dt_kernel(StepControl *ctl)
{
    if (ctl->t < ctl->t_max) {
        double dt = 0.9 * ctl->cfl * ctl->min_h / ctl->max_eig[0];
        if (ctl->t + dt > ctl->t_max) dt = ctl->t_max - ctl->t;
        ctl->dt     = dt;
        ctl->active = 1.0;
    } else {
        ctl->dt     = 0.0;
        ctl->active = 0.0;
    }
}

// Uniform path, start of a step: the max eigenvalue of U is either what the previous step left in
// eig_next (have_candidate) or is computed next by the full eigenvalue pass.
__global__ void begin_step_kernel(StepControl *ctl, int have_candidate)
{
    ctl->max_eig[0]  = have_candidate ? ctl->eig_next : 0.0;
    ctl->max_eig[1]  = 0.0;
    ctl->max_eig[2]  = 0.0;
    ctl->max_eig_chk = 0.0;
    ctl->eig_next    = 0.0;
}

// begin_step_kernel with a candidate followed by choose_dt_kernel, as one launch: the steady state of the uniform path
__global__ void begin_step_choose_dt_kernel(StepControl *ctl)
{
    const double eig = ctl->eig_next;
    ctl->max_eig[0]  = eig;
    ctl->max_eig[1]  = 0.0;
    ctl->max_eig[2]  = 0.0;
    ctl->max_eig_chk = 0.0;
    ctl->eig_next    = 0.0;
    if (ctl->t < ctl->t_max) {
        double dt = 0.9 * ctl->cfl * ctl->min_h / eig;
        if (ctl->t + dt > ctl->t_max) dt = ctl->t_max - ctl->t;
        ctl->dt     = dt;
        ctl->active = 1.0;
    } else {
        ctl->dt     = 0.0;
        ctl->active = 0.0;
    }
}

// host-chosen dt (mmf_rk_stage keeps main.cpp's own dt logic on the host)
__global__ void set_dt_kernel(StepControl *ctl, double dt)
{
    ctl->dt     = dt;
    ctl->active = 1.0;
}

// src/main.cpp:505-506; check_eig: the uniform path compares the eigenvalue that chose dt with the
// face maximum the fused stage-1 kernel re-derived.
// A step enqueued at t >= tMax switched itself off: U is unchanged, so the eigenvalue this step started from
// (max_eig[0]) is still the one of U and stays the next step's candidate -- the stage kernels returned early, and
// whatever the eigenvalue passes behind stage 3 left in eig_next comes from stale tile estimates without the ghost
// cells' own values.
__global__ void advance_time_kernel(StepControl *ctl, int check_eig)
{
    if (ctl->active != 0.0) {
        ctl->t += ctl->dt;
        ctl->steps += 1.0;
        if (check_eig && ctl->max_eig_chk != ctl->max_eig[0]) ctl->mismatches += 1.0;
    } else if (check_eig) {
        ctl->eig_next = ctl->max_eig[0];
    }
}

} // namespace mmf
