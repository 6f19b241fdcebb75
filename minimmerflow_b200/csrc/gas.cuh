// gas.cuh -- perfect-gas relations and the Local Lax-Friedrichs interface flux, FP64, device side.
//
// Everything here is compiled with -fmad=false and without fast-math: a*b+c is two roundings,
// `/` and sqrt() are IEEE round-to-nearest, exactly as in the reference's x86-64 -O2 build.
// Expression shapes follow the cited reference lines (paths relative to the reference tree).
#pragma once

#include <cuda_runtime.h>

namespace mmf {

constexpr int NF = 5;

// src/constants.hpp:37-47
constexpr int FID_P = 0, FID_U = 1, FID_V = 2, FID_W = 3, FID_T = 4;
constexpr int FID_RHO = 0, FID_RHO_U = 1, FID_RHO_V = 2, FID_RHO_W = 3, FID_RHO_E = 4;

// src/constants.hpp:55; the two derived constants are folded at compile time by the reference's
// compiler as well (IEEE double): GAMMA-1 and 2/(GAMMA-1).
constexpr double GAMMA        = 1.4;
constexpr double GM1          = GAMMA - 1.0;
constexpr double TWO_OVER_GM1 = 2.0 / (GAMMA - 1.0);

constexpr int BC_NONE = -1, BC_FREE_FLOW = 0, BC_REFLECTING = 1, BC_WALL = 2, BC_DIRICHLET = 3;

// src/utils.cpp:37-40
__device__ __forceinline__ double normal_velocity(const double *f, const double *n)
{
    return (f[FID_U] * n[0] + f[FID_V] * n[1] + f[FID_W] * n[2]);
}

// src/utils.cpp:48-63
__device__ __forceinline__ void conservative2primitive(const double *c, double *p)
{
    double K = (c[FID_RHO_U] * c[FID_RHO_U] + c[FID_RHO_V] * c[FID_RHO_V] + c[FID_RHO_W] * c[FID_RHO_W])
             / (c[FID_RHO] * c[FID_RHO]);
    p[FID_T] = (2.0 * c[FID_RHO_E] / c[FID_RHO] - K) / TWO_OVER_GM1;
    p[FID_U] = c[FID_RHO_U] / c[FID_RHO];
    p[FID_V] = c[FID_RHO_V] / c[FID_RHO];
    p[FID_W] = c[FID_RHO_W] / c[FID_RHO];
    p[FID_P] = c[FID_RHO] * p[FID_T];
}

// src/utils.cpp:71-83
__device__ __forceinline__ void primitive2conservative(const double *p, double *c)
{
    c[FID_RHO]   = p[FID_P] / p[FID_T];
    c[FID_RHO_U] = c[FID_RHO] * p[FID_U];
    c[FID_RHO_V] = c[FID_RHO] * p[FID_V];
    c[FID_RHO_W] = c[FID_RHO] * p[FID_W];
    c[FID_RHO_E] = c[FID_RHO] * p[FID_T] / GM1
                 + 0.5 * c[FID_RHO] * (p[FID_U] * p[FID_U] + p[FID_V] * p[FID_V] + p[FID_W] * p[FID_W]);
}

// src/euler.cpp:83-113 (the two log-only branches :94-101 have no numerical effect)
__device__ __forceinline__ void eval_fluxes(const double *cons, const double *prim, const double *n, double *flux)
{
    double u = prim[FID_U];
    double v = prim[FID_V];
    double w = prim[FID_W];

    double vel2 = u * u + v * v + w * w;
    double un   = normal_velocity(prim, n);

    double p   = prim[FID_P];
    double rho = cons[FID_RHO];

    double eto = p / GM1 + 0.5 * rho * vel2;

    double massFlux = rho * un;

    flux[0] = massFlux;
    flux[1] = massFlux * u + p * n[0];
    flux[2] = massFlux * v + p * n[1];
    flux[3] = massFlux * w + p * n[2];
    flux[4] = un * (eto + p);
}

// src/euler.cpp:42-73
__device__ __forceinline__ void eval_splitting(const double *consL, const double *consR, const double *n,
                                               double *flux, double *lambda)
{
    double primL[NF], primR[NF];
    conservative2primitive(consL, primL);
    conservative2primitive(consR, primR);

    double fL[NF], fR[NF];
    eval_fluxes(consL, primL, n, fL);
    eval_fluxes(consR, primR, n, fR);

    double unL     = normal_velocity(primL, n);
    double aL      = sqrt(GAMMA * primL[FID_T]);
    double lambdaL = fabs(unL) + aL;

    double unR     = normal_velocity(primR, n);
    double aR      = sqrt(GAMMA * primR[FID_T]);
    double lambdaR = fabs(unR) + aR;

    double lam = (lambdaR < lambdaL) ? lambdaL : lambdaR; // std::max(lambdaR, lambdaL)
    *lambda = lam;

#pragma unroll
    for (int k = 0; k < NF; ++k) {
        flux[k] = 0.5 * ((fR[k] + fL[k]) - lam * (consR[k] - consL[k]));
    }
}

// src/euler.cpp:322-339 (reflecting) and :352-362 (wall forwards to reflecting)
__device__ __forceinline__ void reflecting_bc_values(const double *normal, const double *cons, double *cons_bc)
{
    double prim[NF];
    conservative2primitive(cons, prim);

    double u0 = prim[FID_U], u1 = prim[FID_V], u2 = prim[FID_W];
    double un = normal_velocity(prim, normal);
    double n0 = un * normal[0], n1 = un * normal[1], n2 = un * normal[2];

    prim[FID_U] = u0 - 2 * n0;
    prim[FID_V] = u1 - 2 * n1;
    prim[FID_W] = u2 - 2 * n2;

    primitive2conservative(prim, cons_bc);
}

// src/euler.cpp:261-288: virtual (outer) state for a boundary interface.
// dirichlet_info = problem::getBorderBCInfo data in primitive order (src/problem.cpp:450-477).
__device__ __forceinline__ void interface_bc_values(int bc, const double *normal, const double *dirichlet_info,
                                                    const double *cons, double *cons_bc)
{
    if (bc == BC_FREE_FLOW) {
#pragma unroll
        for (int k = 0; k < NF; ++k) cons_bc[k] = cons[k];
    } else if (bc == BC_REFLECTING || bc == BC_WALL) {
        reflecting_bc_values(normal, cons, cons_bc);
    } else if (bc == BC_DIRICHLET) {
        primitive2conservative(dirichlet_info, cons_bc);
    }
    // any other code: the reference leaves the virtual state untouched (uninitialised there);
    // mmf_create rejects such codes up front.
}

// Positive doubles order like their bit patterns, so max over faces can use an integer atomic:
// max is exact and order-independent, hence deterministic.
__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v)
{
    atomicMax(reinterpret_cast<unsigned long long *>(addr),
              static_cast<unsigned long long>(__double_as_longlong(v)));
}

// warp-shuffle max, then one value per warp to shared memory, then one atomic per block
__device__ __forceinline__ void block_max_to_global(double v, double *gmax)
{
    __shared__ double warp_max[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double other = __shfl_xor_sync(0xffffffffu, v, o);
        v = (v < other) ? other : v;
    }
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (lane == 0) warp_max[warp] = v;
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x * blockDim.y * blockDim.z + 31) >> 5;
        v = (lane < nw) ? warp_max[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double other = __shfl_xor_sync(0xffffffffu, v, o);
            v = (v < other) ? other : v;
        }
        if (lane == 0) atomic_max_nonneg(gmax, v);
    }
}

} // namespace mmf
