/*
 * mmf_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see mmf_oracle.h).
 *
 * Plain-C restatement of the reference's explicit FV Euler residual-and-update
 * path.  Arithmetic expression shapes (operand order, parenthesisation, true
 * divisions) follow the cited reference lines exactly, because the acceptance
 * criterion is the 13 printed digits of the "Final error" line.
 *
 * Build: gcc -O2 -ffp-contract=off (oracle/Makefile).
 */
#include "mmf_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define NF ORC_N_FIELDS

/* src/constants.hpp:35-56 */
enum { P_ = 0, U_ = 1, V_ = 2, W_ = 3, T_ = 4 };
enum { RHO = 0, RHO_U = 1, RHO_V = 2, RHO_W = 3, RHO_E = 4 };
static const double GAMMA = 1.4;

/* ======================================================================== */
/* Gas law: src/utils.cpp                                                    */
/* ======================================================================== */

/* src/utils.cpp:37-40 */
double orc_normal_velocity(const double *f, const double n[3])
{
    return (f[U_] * n[0] + f[V_] * n[1] + f[W_] * n[2]);
}

/* src/utils.cpp:48-63 */
void orc_conservative2primitive(const double *c, double *p)
{
    double K = (c[RHO_U] * c[RHO_U] + c[RHO_V] * c[RHO_V] + c[RHO_W] * c[RHO_W]) / (c[RHO] * c[RHO]);

    p[T_] = (2.0 * c[RHO_E] / c[RHO] - K) / (2.0 / (GAMMA - 1.0));

    p[U_] = c[RHO_U] / c[RHO];
    p[V_] = c[RHO_V] / c[RHO];
    p[W_] = c[RHO_W] / c[RHO];

    p[P_] = c[RHO] * p[T_];
}

/* src/utils.cpp:71-83 */
void orc_primitive2conservative(const double *p, double *c)
{
    c[RHO] = p[P_] / p[T_];

    c[RHO_U] = c[RHO] * p[U_];
    c[RHO_V] = c[RHO] * p[V_];
    c[RHO_W] = c[RHO] * p[W_];

    c[RHO_E] = c[RHO] * p[T_] / (GAMMA - 1.0)
             + 0.5 * c[RHO] * (p[U_] * p[U_] + p[V_] * p[V_] + p[W_] * p[W_]);
}

/* ======================================================================== */
/* Flux: src/euler.cpp                                                       */
/* ======================================================================== */

/* src/euler.cpp:83-113 (the negative p / rho log lines :94-101 have no effect on results) */
void orc_eval_fluxes(const double *cons, const double *prim, const double n[3], double flux[5])
{
    double u = prim[U_];
    double v = prim[V_];
    double w = prim[W_];

    double vel2 = u * u + v * v + w * w;
    double un   = orc_normal_velocity(prim, n);

    double p   = prim[P_];
    double rho = cons[RHO];

    double eto = p / (GAMMA - 1.) + 0.5 * rho * vel2;

    double massFlux = rho * un;

    flux[0] = massFlux;
    flux[1] = massFlux * u + p * n[0];
    flux[2] = massFlux * v + p * n[1];
    flux[3] = massFlux * w + p * n[2];
    flux[4] = un * (eto + p);
}

/* src/euler.cpp:42-73 -- Local Lax-Friedrichs */
void orc_eval_splitting(const double *consL, const double *consR, const double n[3],
                        double flux[5], double *lambda)
{
    double primL[NF], primR[NF];
    orc_conservative2primitive(consL, primL);
    orc_conservative2primitive(consR, primR);

    double fL[NF], fR[NF];
    orc_eval_fluxes(consL, primL, n, fL);
    orc_eval_fluxes(consR, primR, n, fR);

    double unL     = orc_normal_velocity(primL, n);
    double aL      = sqrt(GAMMA * primL[T_]);
    double lambdaL = fabs(unL) + aL;

    double unR     = orc_normal_velocity(primR, n);
    double aR      = sqrt(GAMMA * primR[T_]);
    double lambdaR = fabs(unR) + aR;

    /* std::max(lambdaR, lambdaL): returns lambdaR unless lambdaR < lambdaL */
    *lambda = (lambdaR < lambdaL) ? lambdaL : lambdaR;

    for (int k = 0; k < NF; ++k) {
        flux[k] = 0.5 * ((fR[k] + fL[k]) - (*lambda) * (consR[k] - consL[k]));
    }
}

/* src/problem.cpp:450-477 -- only FFSTEP/DIRICHLET fills data */
static void border_bc_info(int problem, int bc, double info[NF])
{
    if (problem == ORC_PROBLEM_FFSTEP && bc == ORC_BC_DIRICHLET) {
        info[U_] = 3.0;
        info[V_] = 0.;
        info[W_] = 0.;
        info[P_] = 1.;
        info[T_] = 1. / 1.4;
    }
}

/* src/euler.cpp:347-356 (reflecting) ; :369-376 (wall forwards to reflecting) */
static void reflecting_bc(const double normal[3], const double *cons, double *cons_bc)
{
    double prim[NF];
    orc_conservative2primitive(cons, prim);

    double u[3]  = { prim[U_], prim[V_], prim[W_] };
    double un    = orc_normal_velocity(prim, normal);
    double u_n[3] = { un * normal[0], un * normal[1], un * normal[2] };

    prim[U_] = u[0] - 2 * u_n[0];
    prim[V_] = u[1] - 2 * u_n[1];
    prim[W_] = u[2] - 2 * u_n[2];

    orc_primitive2conservative(prim, cons_bc);
}

/* src/euler.cpp:261-288 dispatch; :298-309 free flow; :386-397 dirichlet */
void orc_eval_interface_bc_values(int problem, int bc, const double point[3], const double normal[3],
                                  const double *cons, double *cons_bc)
{
    (void) point;
    double info[NF] = { 0, 0, 0, 0, 0 };
    border_bc_info(problem, bc, info);

    switch (bc) {
    case ORC_BC_FREE_FLOW:
        memcpy(cons_bc, cons, NF * sizeof(double));
        break;
    case ORC_BC_REFLECTING:
    case ORC_BC_WALL:
        reflecting_bc(normal, cons, cons_bc);
        break;
    case ORC_BC_DIRICHLET:
        orc_primitive2conservative(info, cons_bc);
        break;
    default:
        break; /* reference leaves cons_bc untouched for unknown codes */
    }
}

/* ======================================================================== */
/* Uniform mesh in bitpit VolOctree conventions                              */
/*                                                                          */
/* bitpit is a third-party dependency that is absent from the reference     */
/* tree (CMakeLists.txt:103 find_package(BITPIT), no version pin; API usage  */
/* implies 1.7.x).  What is restated here is its published behaviour for a   */
/* uniform octree: level = ceil(log2(L/dh)) (VolOctree ctor called at        */
/* src/main.cpp:146-150), cells in Morton (Z-order) sequence with x the      */
/* lowest interleaved bit, interfaces created while visiting cells in order  */
/* and faces in order -x,+x,-y,+y,-z,+z when no interface exists yet, border */
/* faces owned by their only cell with the outward normal, interior faces    */
/* owned by the lower cell with the normal pointing owner->neigh.  Geometry  */
/* values (h, h^(d-1), h^d, centroids; z-centroid = origin_z in 2-D) are     */
/* pinned by the five golden strings; the id ORDER is not observable through */
/* any reference test (results are order-insensitive at printed precision).  */
/* ======================================================================== */

static inline uint64_t morton_encode(int dim, uint32_t i, uint32_t j, uint32_t k)
{
    uint64_t m = 0;
    for (int b = 0; b < 21; ++b) {
        m |= (uint64_t) ((i >> b) & 1u) << (dim * b);
        m |= (uint64_t) ((j >> b) & 1u) << (dim * b + 1);
        if (dim == 3) {
            m |= (uint64_t) ((k >> b) & 1u) << (dim * b + 2);
        }
    }
    return m;
}

static inline void morton_decode(int dim, uint64_t m, uint32_t *i, uint32_t *j, uint32_t *k)
{
    uint32_t x = 0, y = 0, z = 0;
    for (int b = 0; b < 21; ++b) {
        x |= (uint32_t) ((m >> (dim * b)) & 1u) << b;
        y |= (uint32_t) ((m >> (dim * b + 1)) & 1u) << b;
        if (dim == 3) {
            z |= (uint32_t) ((m >> (dim * b + 2)) & 1u) << b;
        }
    }
    *i = x; *j = y; *k = z;
}

/* level = ceil(log2(max(1, L/dh))), dh = L / nCells  (src/main.cpp:146-150 + bitpit VolOctree) */
int orc_level_for(double length, long n_cells_per_dir)
{
    double dh = length / (double) n_cells_per_dir;
    double ratio = length / dh;
    if (ratio < 1.) ratio = 1.;
    return (int) ceil(log2(ratio));
}

void orc_uniform_counts(int dim, int level, long *n_cells, long *n_ifaces)
{
    long N = 1L << level;
    if (dim == 3) {
        *n_cells  = N * N * N;
        *n_ifaces = 3 * N * N * (N + 1);
    } else {
        *n_cells  = N * N;
        *n_ifaces = 2 * N * (N + 1);
    }
}

void orc_uniform_mesh(int dim, const double origin[3], double length, int level,
                      int64_t *owner, int64_t *neigh, double *area, double *normal, double *icentroid,
                      double *volume, double *size, double *ccentroid, int32_t *cell_ijk)
{
    const long N = 1L << level;
    const double h = length / (double) N;
    const double V = (dim == 3) ? h * h * h : h * h;
    const double A = (dim == 3) ? h * h : h;
    long n_cells, n_ifaces;
    orc_uniform_counts(dim, level, &n_cells, &n_ifaces);

    long f = 0;
    for (long c = 0; c < n_cells; ++c) {
        uint32_t ijk[3];
        morton_decode(dim, (uint64_t) c, &ijk[0], &ijk[1], &ijk[2]);

        double cc[3];
        for (int d = 0; d < 3; ++d) {
            cc[d] = (d < dim) ? origin[d] + ((double) ijk[d] + 0.5) * h : origin[d];
            ccentroid[3 * c + d] = cc[d];
            if (cell_ijk) cell_ijk[3 * c + d] = (int32_t) ijk[d];
        }
        volume[c] = V;
        size[c]   = h;

        for (int face = 0; face < 2 * dim; ++face) {
            int d    = face / 2;
            int sign = (face % 2) ? +1 : -1;
            long nb_coord = (long) ijk[d] + sign;
            int border = (nb_coord < 0 || nb_coord >= N);
            int64_t nb = -1;
            if (!border) {
                uint32_t q[3] = { ijk[0], ijk[1], ijk[2] };
                q[d] = (uint32_t) nb_coord;
                nb = (int64_t) morton_encode(dim, q[0], q[1], q[2]);
                if (nb < c) continue; /* interface already created by the lower cell */
            }
            owner[f] = c;
            neigh[f] = nb;
            area[f]  = A;
            for (int e = 0; e < 3; ++e) {
                normal[3 * f + e]    = (e == d) ? (double) sign : 0.;
                icentroid[3 * f + e] = (e == d) ? cc[e] + 0.5 * sign * h : cc[e];
            }
            ++f;
        }
    }
    if (f != n_ifaces) {
        fprintf(stderr, "orc_uniform_mesh: interface count mismatch %ld != %ld\n", f, n_ifaces);
        abort();
    }
}

/* ======================================================================== */
/* Problem catalogue: src/problem.cpp                                        */
/* ======================================================================== */

/* src/problem.cpp:70-149 (defaults only; custom <domain> handled by the caller) */
void orc_domain_defaults(int problem, int dim_in, int *dim, double origin[3], double *length)
{
    switch (problem) {
    case ORC_PROBLEM_VORTEX_ZX:
    case ORC_PROBLEM_VORTEX_YZ:
        *dim = 3;
        break;
    default:
        *dim = (dim_in == 3) ? 3 : 2; /* default 2: problem.cpp:81 */
        break;
    }

    switch (problem) {
    case ORC_PROBLEM_VORTEX_XY:
    case ORC_PROBLEM_VORTEX_ZX:
    case ORC_PROBLEM_VORTEX_YZ:
        origin[0] = -5; origin[1] = -5; origin[2] = -5.;
        *length = 10.;
        break;
    case ORC_PROBLEM_SOD_X:
    case ORC_PROBLEM_SOD_Y:
    case ORC_PROBLEM_SOD_Z:
        origin[0] = -1; origin[1] = -1; origin[2] = -1.;
        *length = 2.;
        break;
    case ORC_PROBLEM_RADSOD:
        origin[0] = 0; origin[1] = 0; origin[2] = 0;
        *length = 8.;
        break;
    default: /* ffstep: length is mandatory in the reference (problem.cpp:145) */
        origin[0] = 0; origin[1] = 0; origin[2] = 0;
        *length = -1.;
        break;
    }
}

/* src/problem.cpp:174-211 */
double orc_end_time_default(int problem, int dim)
{
    switch (problem) {
    case ORC_PROBLEM_VORTEX_XY:
    case ORC_PROBLEM_VORTEX_ZX:
    case ORC_PROBLEM_VORTEX_YZ:
        return (dim == 2) ? 2. : 1.;
    case ORC_PROBLEM_SOD_X:
    case ORC_PROBLEM_SOD_Y:
    case ORC_PROBLEM_SOD_Z:
        return 0.3;
    case ORC_PROBLEM_RADSOD:
        return 4.;
    case ORC_PROBLEM_FFSTEP:
        return 4.;
    }
    return 0.;
}

/* src/problem.cpp:247-392 */
void orc_exact_conservatives(int problem, int dim, const double point_in[3], double t, double *cons)
{
    double point[3] = { point_in[0], point_in[1], point_in[2] };
    double prim[NF] = { 0, 0, 0, 0, 0 };

    switch (problem) {
    case ORC_PROBLEM_VORTEX_XY:
    case ORC_PROBLEM_VORTEX_ZX:
    case ORC_PROBLEM_VORTEX_YZ:
    {
        const double BETA  = 5.0;
        const double P_INF = 1.0;
        const double T_INF = 1.0;
        double U_INF = 1.0, V_INF = 1.0, W_INF = 1.0;
        if (problem == ORC_PROBLEM_VORTEX_XY)      W_INF = 0.;
        else if (problem == ORC_PROBLEM_VORTEX_ZX) V_INF = 0.;
        else                                       U_INF = 0.;

        point[0] -= t * U_INF;
        point[1] -= t * V_INF;
        point[2] -= t * W_INF;

        double csi = 0., eta = 0.;
        if (problem == ORC_PROBLEM_VORTEX_XY)      { csi = point[0]; eta = point[1]; }
        else if (problem == ORC_PROBLEM_VORTEX_ZX) { csi = point[2]; eta = point[0]; }
        else                                       { csi = point[1]; eta = point[2]; }

        double r     = sqrt(csi * csi + eta * eta);
        double shape = BETA / (2 * M_PI) * exp(0.5 * (1 - r * r));

        double du = 0., dv = 0., dw = 0.;
        if (problem == ORC_PROBLEM_VORTEX_XY)      { du = -eta * shape; dv = csi * shape; dw = 0.; }
        else if (problem == ORC_PROBLEM_VORTEX_ZX) { dw = -eta * shape; du = csi * shape; dv = 0.; }
        else                                       { dv = -eta * shape; dw = csi * shape; du = 0.; }

        const double dT = -(GAMMA - 1.0) / (2 * GAMMA) * shape * shape;
        const double dp = pow(T_INF + dT, GAMMA / (GAMMA - 1.0)) - 1.0;

        prim[P_] = P_INF + dp;
        prim[U_] = U_INF + du;
        prim[V_] = V_INF + dv;
        prim[W_] = (dim == 3) ? W_INF + dw : 0;
        prim[T_] = T_INF + dT;
        break;
    }
    case ORC_PROBLEM_RADSOD:
    {
        double r = sqrt(point[0] * point[0] + point[1] * point[1] + point[2] * point[2]);
        prim[U_] = 0.0; prim[V_] = 0.0; prim[W_] = 0.0;
        if (r < 1.) { prim[P_] = 1.0; prim[T_] = 1.0 / 1.0; }
        else        { prim[P_] = 0.1; prim[T_] = 0.1 / 0.125; }
        break;
    }
    case ORC_PROBLEM_SOD_X:
    case ORC_PROBLEM_SOD_Y:
    case ORC_PROBLEM_SOD_Z:
    {
        prim[U_] = 0.0; prim[V_] = 0.0; prim[W_] = 0.0;
        double r = (problem == ORC_PROBLEM_SOD_X) ? point[0]
                 : (problem == ORC_PROBLEM_SOD_Y) ? point[1] : point[2];
        if (r < 0.) { prim[P_] = 1.0; prim[T_] = 1.0 / 1.0; }
        else        { prim[P_] = 0.1; prim[T_] = 0.1 / 0.125; }
        break;
    }
    case ORC_PROBLEM_FFSTEP:
        prim[U_] = 3.0; prim[V_] = 0.0; prim[W_] = 0.0;
        prim[P_] = 1.0; prim[T_] = 1.0 / 1.4;
        break;
    }

    orc_primitive2conservative(prim, cons);
}

/* src/problem.cpp:404-438 */
int orc_border_bc_type(int problem, const double fc[3])
{
    switch (problem) {
    case ORC_PROBLEM_RADSOD:
        return ORC_BC_REFLECTING;
    case ORC_PROBLEM_FFSTEP:
        if (fc[0] < 1e-10)            return ORC_BC_DIRICHLET;
        else if (fc[0] > 3.2 - 1e-10) return ORC_BC_FREE_FLOW;
        else                          return ORC_BC_REFLECTING;
    default:
        return ORC_BC_FREE_FLOW;
    }
}

/* src/main.cpp:335-343 (initial conditions = exact solution at t, problem.cpp:232-235) */
void orc_init_state(int problem, int dim, long n_cells, const double *ccentroid, double t, double *U)
{
    for (long c = 0; c < n_cells; ++c) {
        orc_exact_conservatives(problem, dim, &ccentroid[3 * c], t, &U[NF * c]);
    }
}

/* ======================================================================== */
/* Flags and BC table                                                        */
/* ======================================================================== */

/* src/body.cpp:80-95 (closed-interval point-in-box on the cell centroid) + src/main.cpp:221-228 */
void orc_fluid_flags(long n_cells, const double *ccentroid, int n_boxes, const double *boxes, uint8_t *fluid)
{
    for (long c = 0; c < n_cells; ++c) {
        const double *pt = &ccentroid[3 * c];
        uint8_t is_fluid = 1;
        for (int b = 0; b < n_boxes; ++b) {
            const double *bx = &boxes[6 * b]; /* xMin,yMin,zMin,xMax,yMax,zMax */
            if (pt[0] < bx[0] || pt[0] > bx[3]) continue;
            else if (pt[1] < bx[1] || pt[1] > bx[4]) continue;
            else if (pt[2] < bx[2] || pt[2] > bx[5]) continue;
            is_fluid = 0;
            break;
        }
        fluid[c] = is_fluid;
    }
}

/* src/main.cpp:251-277 */
void orc_interface_bcs(int problem, long n_ifaces, const int64_t *owner, const int64_t *neigh,
                       const double *icentroid, const uint8_t *fluid, int32_t *bc)
{
    for (long f = 0; f < n_ifaces; ++f) {
        if (neigh[f] < 0) {
            bc[f] = orc_border_bc_type(problem, &icentroid[3 * f]);
        } else if (fluid[owner[f]] ^ fluid[neigh[f]]) {
            bc[f] = ORC_BC_WALL;
        } else {
            bc[f] = ORC_BC_NONE;
        }
    }
}

/* ======================================================================== */
/* Residual assembly: src/euler.cpp:127-249                                  */
/* ======================================================================== */

/* Per-interface body shared by the serial and threaded loops.  Returns 0 if the
 * interface is skipped (:181-183), else 1 with flux[] and *lambda filled. */
static inline int face_flux(int problem, long f,
                            const int64_t *owner, const int64_t *neigh, const int32_t *bc,
                            const double *normal, const double *icentroid,
                            const uint8_t *solved, const double *U,
                            double flux[NF], double *lambda, int *owner_solved, int *neigh_solved)
{
    const int64_t o = owner[f];
    const int64_t n = neigh[f];
    const int oS = solved[o] != 0;
    const int nS = (n >= 0) ? (solved[n] != 0) : 0;
    *owner_solved = oS;
    *neigh_solved = nS;
    if (!oS && !nS) return 0;

    const double *nrm = &normal[3 * f];
    double ownerRec[NF], neighRec[NF];
    if (bc[f] == ORC_BC_NONE) {
        /* order-1 reconstruction = copy of the means: src/reconstruction.cpp:90-97 */
        memcpy(ownerRec, &U[NF * o], sizeof ownerRec);
        memcpy(neighRec, &U[NF * n], sizeof neighRec);
    } else {
        /* :198-225 -- fluid side is the owner if it is solved, else the neighbour with a
         * flipped normal that is used for the BC evaluation only */
        double bcNormal[3] = { nrm[0], nrm[1], nrm[2] };
        double *fluidRec, *virtualRec;
        const double *fluidMean;
        if (oS) {
            fluidMean = &U[NF * o]; fluidRec = ownerRec; virtualRec = neighRec;
        } else {
            fluidMean = &U[NF * n]; fluidRec = neighRec; virtualRec = ownerRec;
            bcNormal[0] = -1. * nrm[0]; bcNormal[1] = -1. * nrm[1]; bcNormal[2] = -1. * nrm[2];
        }
        memcpy(fluidRec, fluidMean, NF * sizeof(double));
        orc_eval_interface_bc_values(problem, bc[f], &icentroid[3 * f], bcNormal, fluidRec, virtualRec);
    }

    /* :232 -- the splitting always uses the un-flipped interface normal */
    orc_eval_splitting(ownerRec, neighRec, nrm, flux, lambda);
    return 1;
}

void orc_compute_rhs(int problem, long n_cells, long n_ifaces,
                     const int64_t *owner, const int64_t *neigh, const int32_t *bc,
                     const double *area, const double *normal, const double *icentroid,
                     const uint8_t *solved, const double *U, double *RHS, double *max_eig)
{
    /* :135-148 -- reset for ALL cells (the list is getCellRawIds()) */
    for (long i = 0; i < n_cells * NF; ++i) RHS[i] = 0.;

    *max_eig = 0.0;
    for (long f = 0; f < n_ifaces; ++f) {
        double flux[NF], lambda;
        int oS, nS;
        if (!face_flux(problem, f, owner, neigh, bc, normal, icentroid, solved, U, flux, &lambda, &oS, &nS)) {
            continue;
        }
        /* :234 std::max(faceMaxEig, *maxEig) */
        *max_eig = (lambda < *max_eig) ? *max_eig : lambda;

        const double A = area[f];
        if (oS) {
            double *r = &RHS[NF * owner[f]];
            for (int k = 0; k < NF; ++k) r[k] -= A * flux[k];
        }
        if (nS) {
            double *r = &RHS[NF * neigh[f]];
            for (int k = 0; k < NF; ++k) r[k] += A * flux[k];
        }
    }
}

/* ======================================================================== */
/* RK3 stages and dt: src/main.cpp:391-495                                   */
/* ======================================================================== */

static inline void rk_cell(int stage, double dt, double V, double *u, double *w, const double *r)
{
    switch (stage) {
    case 1: /* :409-423 */
        for (int k = 0; k < NF; ++k) w[k] = u[k] + dt * r[k] / V;
        break;
    case 2: /* :445-459 */
        for (int k = 0; k < NF; ++k) w[k] = 0.75 * u[k] + 0.25 * (w[k] + dt * r[k] / V);
        break;
    default: /* :481-495 */
        for (int k = 0; k < NF; ++k) u[k] = (1. / 3) * u[k] + (2. / 3) * (w[k] + dt * r[k] / V);
        break;
    }
}

void orc_rk_stage(int stage, long n_cells, const uint8_t *update_mask, const double *volume,
                  double dt, double *U, double *W, const double *RHS)
{
    for (long c = 0; c < n_cells; ++c) {
        if (!update_mask[c]) continue;
        rk_cell(stage, dt, volume[c], &U[NF * c], &W[NF * c], &RHS[NF * c]);
    }
}

/* src/main.cpp:398-402 */
double orc_choose_dt(double cfl, double min_cell_size, double max_eig, double t, double t_max)
{
    double dt = 0.9 * cfl * min_cell_size / max_eig;
    if (t + dt > t_max) dt = t_max - t;
    return dt;
}

/* src/main.cpp:383-502, serial build */
double orc_step(int problem, long n_cells, long n_ifaces,
                const int64_t *owner, const int64_t *neigh, const int32_t *bc,
                const double *area, const double *normal, const double *icentroid,
                const uint8_t *solved, const uint8_t *update_mask, const double *volume,
                double cfl, double min_cell_size, double t, double t_max,
                double *U, double *W, double *RHS, double max_eig3[3])
{
    double me;
    orc_compute_rhs(problem, n_cells, n_ifaces, owner, neigh, bc, area, normal, icentroid, solved, U, RHS, &me);
    max_eig3[0] = me;
    double dt = orc_choose_dt(cfl, min_cell_size, me, t, t_max);
    orc_rk_stage(1, n_cells, update_mask, volume, dt, U, W, RHS);

    orc_compute_rhs(problem, n_cells, n_ifaces, owner, neigh, bc, area, normal, icentroid, solved, W, RHS, &me);
    max_eig3[1] = me;
    orc_rk_stage(2, n_cells, update_mask, volume, dt, U, W, RHS);

    orc_compute_rhs(problem, n_cells, n_ifaces, owner, neigh, bc, area, normal, icentroid, solved, W, RHS, &me);
    max_eig3[2] = me;
    orc_rk_stage(3, n_cells, update_mask, volume, dt, U, W, RHS);
    return dt;
}

/* src/main.cpp:550-561 -- FID_P == FID_RHO == 0, sequential accumulation over internal cells */
double orc_error_norm(int problem, int dim, long n_cells, const double *ccentroid, const double *volume,
                      const uint8_t *internal, const double *U, double t_max)
{
    double error = 0.;
    for (long c = 0; c < n_cells; ++c) {
        if (internal && !internal[c]) continue;
        double exact[NF];
        orc_exact_conservatives(problem, dim, &ccentroid[3 * c], t_max, exact);
        error += fabs(U[NF * c + 0] - exact[0]) * volume[c];
    }
    return error;
}

/* ======================================================================== */
/* Whole serial flow of src/main.cpp:46-573 on a uniform mesh                */
/* ======================================================================== */

typedef struct {
    int dim, level;
    long n_cells, n_ifaces;
    double h;
    int64_t *owner, *neigh;
    int32_t *bc, *ijk;
    double *area, *normal, *icentroid, *volume, *size, *ccentroid;
    uint8_t *fluid;
} mesh_t;

static void mesh_alloc(mesh_t *m, int dim, const double origin[3], double length, int level)
{
    m->dim = dim;
    m->level = level;
    orc_uniform_counts(dim, level, &m->n_cells, &m->n_ifaces);
    m->h = length / (double) (1L << level);
    m->owner     = malloc(sizeof(int64_t) * m->n_ifaces);
    m->neigh     = malloc(sizeof(int64_t) * m->n_ifaces);
    m->bc        = malloc(sizeof(int32_t) * m->n_ifaces);
    m->area      = malloc(sizeof(double) * m->n_ifaces);
    m->normal    = malloc(sizeof(double) * 3 * m->n_ifaces);
    m->icentroid = malloc(sizeof(double) * 3 * m->n_ifaces);
    m->volume    = malloc(sizeof(double) * m->n_cells);
    m->size      = malloc(sizeof(double) * m->n_cells);
    m->ccentroid = malloc(sizeof(double) * 3 * m->n_cells);
    m->ijk       = malloc(sizeof(int32_t) * 3 * m->n_cells);
    m->fluid     = malloc(m->n_cells);
    orc_uniform_mesh(dim, origin, length, level, m->owner, m->neigh, m->area, m->normal, m->icentroid,
                     m->volume, m->size, m->ccentroid, m->ijk);
}

static void mesh_free(mesh_t *m)
{
    free(m->owner); free(m->neigh); free(m->bc); free(m->area); free(m->normal); free(m->icentroid);
    free(m->volume); free(m->size); free(m->ccentroid); free(m->ijk); free(m->fluid);
}

int orc_run(int problem, int dim_in, long n_cells_per_dir, double t_end, double cfl,
            int n_boxes, const double *boxes, int max_steps,
            double *error_out, double *t_out, double *U_out)
{
    int dim;
    double origin[3], length;
    orc_domain_defaults(problem, dim_in, &dim, origin, &length);
    const double tMin = 0.;
    const double tMax = (t_end >= 0.) ? t_end : orc_end_time_default(problem, dim);

    mesh_t m;
    mesh_alloc(&m, dim, origin, length, orc_level_for(length, n_cells_per_dir));

    /* flags: serial build -> solved == fluid, every cell internal (main.cpp:221-237) */
    orc_fluid_flags(m.n_cells, m.ccentroid, n_boxes, boxes, m.fluid);
    orc_interface_bcs(problem, m.n_ifaces, m.owner, m.neigh, m.icentroid, m.fluid, m.bc);

    double *U   = malloc(sizeof(double) * NF * m.n_cells);
    double *W   = calloc(NF * m.n_cells, sizeof(double));
    double *RHS = calloc(NF * m.n_cells, sizeof(double));
    orc_init_state(problem, dim, m.n_cells, m.ccentroid, 0.0, U);

    /* main.cpp:351-356 */
    double minCellSize = 1.7976931348623157e308;
    for (long c = 0; c < m.n_cells; ++c) {
        if (m.size[c] < minCellSize) minCellSize = m.size[c];
    }

    int step = 0;
    double t = tMin;
    while (t < tMax && (max_steps < 0 || step < max_steps)) {
        double me3[3];
        double dt = orc_step(problem, m.n_cells, m.n_ifaces, m.owner, m.neigh, m.bc, m.area, m.normal,
                             m.icentroid, m.fluid, m.fluid, m.volume, cfl, minCellSize, t, tMax,
                             U, W, RHS, me3);
        t += dt;
        step++;
    }

    if (error_out) *error_out = orc_error_norm(problem, dim, m.n_cells, m.ccentroid, m.volume, NULL, U, tMax);
    if (t_out) *t_out = t;
    if (U_out) memcpy(U_out, U, sizeof(double) * NF * m.n_cells);

    free(U); free(W); free(RHS);
    mesh_free(&m);
    return step;
}

/* ======================================================================== */
/* Threaded CPU baseline                                                     */
/*                                                                          */
/* The reference's only parallelism is MPI domain decomposition into         */
/* contiguous Morton chunks with one ghost layer (src/main.cpp:157-186,      */
/* 231-235).  MPI is not in this image, so the same decomposition is run on  */
/* threads: each thread owns a contiguous chunk of cells, walks every        */
/* interface that touches its chunk in ascending interface order and         */
/* accumulates only into its own cells -- i.e. faces on a chunk boundary are  */
/* computed by both sides, exactly like the reference's interior/ghost rule. */
/* Per-cell accumulation order is unchanged, so the result is bitwise equal  */
/* to the serial loop.                                                       */
/* ======================================================================== */

typedef struct {
    int tid, n_threads, problem, n_steps_total, n_warmup;
    mesh_t *m;
    long c0, c1;          /* owned cell range */
    long *iface; long n_iface;
    double *U, *W, *RHS;
    double cfl, min_h;
    double *thread_max;   /* [n_threads] */
    pthread_barrier_t *bar;
    double *dt_shared;
    struct timespec *t_start, *t_end;
} worker_t;

static void worker_rhs(worker_t *w, const double *S)
{
    mesh_t *m = w->m;
    for (long i = NF * w->c0; i < NF * w->c1; ++i) w->RHS[i] = 0.;
    double me = 0.;
    for (long q = 0; q < w->n_iface; ++q) {
        long f = w->iface[q];
        double flux[NF], lambda;
        int oS, nS;
        if (!face_flux(w->problem, f, m->owner, m->neigh, m->bc, m->normal, m->icentroid, m->fluid, S,
                       flux, &lambda, &oS, &nS)) continue;
        me = (lambda < me) ? me : lambda;
        const double A = m->area[f];
        int64_t o = m->owner[f], n = m->neigh[f];
        if (oS && o >= w->c0 && o < w->c1) {
            double *r = &w->RHS[NF * o];
            for (int k = 0; k < NF; ++k) r[k] -= A * flux[k];
        }
        if (nS && n >= w->c0 && n < w->c1) {
            double *r = &w->RHS[NF * n];
            for (int k = 0; k < NF; ++k) r[k] += A * flux[k];
        }
    }
    w->thread_max[w->tid] = me;
}

static void worker_update(worker_t *w, int stage, double dt)
{
    mesh_t *m = w->m;
    for (long c = w->c0; c < w->c1; ++c) {
        if (!m->fluid[c]) continue;
        rk_cell(stage, dt, m->volume[c], &w->U[NF * c], &w->W[NF * c], &w->RHS[NF * c]);
    }
}

static void *worker_main(void *arg)
{
    worker_t *w = arg;
    for (int step = 0; step < w->n_steps_total; ++step) {
        if (step == w->n_warmup) {
            pthread_barrier_wait(w->bar);
            if (w->tid == 0) clock_gettime(CLOCK_MONOTONIC, w->t_start);
        }
        /* stage 1 */
        worker_rhs(w, w->U);
        pthread_barrier_wait(w->bar);
        if (w->tid == 0) {
            double me = 0.;
            for (int i = 0; i < w->n_threads; ++i) me = (w->thread_max[i] < me) ? me : w->thread_max[i];
            *w->dt_shared = 0.9 * w->cfl * w->min_h / me;
        }
        pthread_barrier_wait(w->bar);
        double dt = *w->dt_shared;
        worker_update(w, 1, dt);
        pthread_barrier_wait(w->bar);   /* "ghost exchange": neighbours' W must be complete */
        /* stage 2 */
        worker_rhs(w, w->W);
        pthread_barrier_wait(w->bar);   /* everyone has read W before anyone overwrites it */
        worker_update(w, 2, dt);
        pthread_barrier_wait(w->bar);
        /* stage 3 */
        worker_rhs(w, w->W);
        pthread_barrier_wait(w->bar);
        worker_update(w, 3, dt);
        pthread_barrier_wait(w->bar);
    }
    if (w->tid == 0) clock_gettime(CLOCK_MONOTONIC, w->t_end);
    return NULL;
}

static double run_threads(int problem, int dim_in, int level, int n_warmup, int n_steps, int n_threads,
                          double cfl, uint64_t *state_hash, double *U0_out, double *U_out)
{
    int dim;
    double origin[3], length;
    orc_domain_defaults(problem, dim_in, &dim, origin, &length);

    mesh_t m;
    mesh_alloc(&m, dim, origin, length, level);
    orc_fluid_flags(m.n_cells, m.ccentroid, 0, NULL, m.fluid);
    orc_interface_bcs(problem, m.n_ifaces, m.owner, m.neigh, m.icentroid, m.fluid, m.bc);

    double *U   = malloc(sizeof(double) * NF * m.n_cells);
    double *W   = calloc(NF * m.n_cells, sizeof(double));
    double *RHS = calloc(NF * m.n_cells, sizeof(double));
    orc_init_state(problem, dim, m.n_cells, m.ccentroid, 0.0, U);
    if (U0_out) memcpy(U0_out, U, sizeof(double) * NF * m.n_cells);

    if (n_threads < 1) n_threads = 1;
    if (n_threads > m.n_cells) n_threads = (int) m.n_cells;

    worker_t *ws = calloc(n_threads, sizeof(worker_t));
    pthread_t *th = malloc(sizeof(pthread_t) * n_threads);
    double *thread_max = calloc(n_threads, sizeof(double));
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, n_threads);
    double dt_shared = 0.;
    struct timespec ts0, ts1;

    /* equal contiguous chunks, remainder to the first ranks */
    long base = m.n_cells / n_threads, rem = m.n_cells % n_threads, c = 0;
    for (int t = 0; t < n_threads; ++t) {
        worker_t *w = &ws[t];
        w->tid = t; w->n_threads = n_threads; w->problem = problem;
        w->n_steps_total = n_warmup + n_steps; w->n_warmup = n_warmup;
        w->m = &m; w->c0 = c; w->c1 = c + base + (t < rem ? 1 : 0); c = w->c1;
        w->U = U; w->W = W; w->RHS = RHS; w->cfl = cfl; w->min_h = m.h;
        w->thread_max = thread_max; w->bar = &bar; w->dt_shared = &dt_shared;
        w->t_start = &ts0; w->t_end = &ts1;
        long cnt = 0;
        for (long f = 0; f < m.n_ifaces; ++f) {
            int64_t o = m.owner[f], n = m.neigh[f];
            if ((o >= w->c0 && o < w->c1) || (n >= w->c0 && n < w->c1)) cnt++;
        }
        w->iface = malloc(sizeof(long) * (cnt ? cnt : 1));
        w->n_iface = 0;
        for (long f = 0; f < m.n_ifaces; ++f) {
            int64_t o = m.owner[f], n = m.neigh[f];
            if ((o >= w->c0 && o < w->c1) || (n >= w->c0 && n < w->c1)) w->iface[w->n_iface++] = f;
        }
    }

    for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, worker_main, &ws[t]);
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);

    double secs = (double) (ts1.tv_sec - ts0.tv_sec) + 1e-9 * (double) (ts1.tv_nsec - ts0.tv_nsec);

    if (state_hash) {
        uint64_t hsh = 1469598103934665603ULL;
        const unsigned char *bytes = (const unsigned char *) U;
        for (size_t i = 0; i < sizeof(double) * NF * (size_t) m.n_cells; ++i) {
            hsh ^= bytes[i];
            hsh *= 1099511628211ULL;
        }
        *state_hash = hsh;
    }
    if (U_out) memcpy(U_out, U, sizeof(double) * NF * m.n_cells);

    for (int t = 0; t < n_threads; ++t) free(ws[t].iface);
    pthread_barrier_destroy(&bar);
    free(ws); free(th); free(thread_max);
    free(U); free(W); free(RHS);
    mesh_free(&m);
    return secs;
}

double orc_bench_threads(int problem, int dim_in, int level, int n_warmup, int n_steps, int n_threads,
                         double cfl, uint64_t *state_hash)
{
    return run_threads(problem, dim_in, level, n_warmup, n_steps, n_threads, cfl, state_hash, NULL, NULL);
}

/* The same threaded loop as a CHECKER for large meshes: n_steps RK3 steps (no tMax clamp, like
 * main.cpp:398-402 while t + dt stays below tMax) from the problem's initial state; returns the initial and the
 * final conservative fields in Morton cell order.  Bitwise equal to the serial loop (see above). */
void orc_run_threads(int problem, int dim_in, int level, int n_steps, int n_threads, double cfl,
                     double *U0_out, double *U_out)
{
    run_threads(problem, dim_in, level, 0, n_steps, n_threads, cfl, NULL, U0_out, U_out);
}
