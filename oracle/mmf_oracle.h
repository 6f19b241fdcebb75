/*
 * mmf_oracle.h -- CPU oracle for the minimmerflow explicit finite-volume Euler
 * residual-and-update path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check in
 * __graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py may
 * load it.  The product path (minimmerflow_b200/, libmmf_b200.so) never links,
 * imports or calls anything in this directory.
 *
 * It is a from-scratch plain-C restatement of the reference algorithm (the
 * reference itself cannot be built here: every translation unit includes bitpit,
 * which is neither installed nor vendored -- see DESIGN.md).  Each function cites
 * the reference file:line it follows (paths relative to the reference tree).
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks that orc_run()
 * reproduces the five "Final error" strings that the reference's own regression
 * tests hard-code (test/<case>/CMakeLists.txt:33) to all 13 printed digits.
 *
 * Conventions (validated by those five strings):
 *   - fields: conservative {rho, rho*u, rho*v, rho*w, rho*E}, primitive {p,u,v,w,T}
 *     (src/constants.hpp:35-52), AoS, value (cell c, field k) at [c*5+k]
 *     (src/storage.hpp:31-41).
 *   - IEEE double, no FMA contraction (build with -O2 -ffp-contract=off; the
 *     reference Release build is plain -O2 on x86-64: CMakeLists.txt:152,157).
 */
#ifndef MMF_ORACLE_H
#define MMF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_N_FIELDS 5

/* src/problem.hpp:33-42 */
enum {
    ORC_PROBLEM_VORTEX_XY = 0,
    ORC_PROBLEM_VORTEX_ZX = 1,
    ORC_PROBLEM_VORTEX_YZ = 2,
    ORC_PROBLEM_RADSOD    = 3,
    ORC_PROBLEM_SOD_X     = 4,
    ORC_PROBLEM_SOD_Y     = 5,
    ORC_PROBLEM_SOD_Z     = 6,
    ORC_PROBLEM_FFSTEP    = 7
};

/* src/constants.hpp:58-62 */
enum {
    ORC_BC_NONE       = -1,
    ORC_BC_FREE_FLOW  =  0,
    ORC_BC_REFLECTING =  1,
    ORC_BC_WALL       =  2,
    ORC_BC_DIRICHLET  =  3
};

/* ---- gas law (src/utils.cpp) ------------------------------------------- */
double orc_normal_velocity(const double *prim, const double n[3]);
void   orc_conservative2primitive(const double *c, double *p);
void   orc_primitive2conservative(const double *p, double *c);

/* ---- flux (src/euler.cpp) ---------------------------------------------- */
void orc_eval_fluxes(const double *cons, const double *prim, const double n[3], double flux[5]);
void orc_eval_splitting(const double *consL, const double *consR, const double n[3],
                        double flux[5], double *lambda);
void orc_eval_interface_bc_values(int problem, int bc, const double point[3], const double normal[3],
                                  const double *cons, double *cons_bc);

/* ---- uniform VolOctree-convention mesh --------------------------------- */
int  orc_level_for(double length, long n_cells_per_dir);
void orc_uniform_counts(int dim, int level, long *n_cells, long *n_ifaces);
/* All arrays caller-allocated: owner/neigh [n_ifaces]; area [n_ifaces]; normal,
 * icentroid [n_ifaces*3]; volume,size [n_cells]; ccentroid [n_cells*3];
 * cell_ijk [n_cells*3] (integer lattice coordinates of each cell). */
void orc_uniform_mesh(int dim, const double origin[3], double length, int level,
                      int64_t *owner, int64_t *neigh, double *area, double *normal, double *icentroid,
                      double *volume, double *size, double *ccentroid, int32_t *cell_ijk);

/* ---- problem catalogue (src/problem.cpp) ------------------------------- */
void   orc_domain_defaults(int problem, int dim_in, int *dim, double origin[3], double *length);
double orc_end_time_default(int problem, int dim);
void   orc_exact_conservatives(int problem, int dim, const double point[3], double t, double *cons);
int    orc_border_bc_type(int problem, const double face_centroid[3]);
void   orc_init_state(int problem, int dim, long n_cells, const double *ccentroid, double t, double *U);

/* ---- flags / BC table (src/main.cpp:221-277, src/body.cpp:80-95) ------- */
void orc_fluid_flags(long n_cells, const double *ccentroid, int n_boxes, const double *boxes, uint8_t *fluid);
void orc_interface_bcs(int problem, long n_ifaces, const int64_t *owner, const int64_t *neigh,
                       const double *icentroid, const uint8_t *fluid, int32_t *bc);

/* ---- residual assembly (src/euler.cpp:127-249) ------------------------- */
void orc_compute_rhs(int problem, long n_cells, long n_ifaces,
                     const int64_t *owner, const int64_t *neigh, const int32_t *bc,
                     const double *area, const double *normal, const double *icentroid,
                     const uint8_t *solved, const double *U, double *RHS, double *max_eig);

/* ---- RK3 stage updates and dt (src/main.cpp:391-495) ------------------- */
/* update_mask[c] != 0  <=>  cell is internal AND solved. stage in {1,2,3}. */
void   orc_rk_stage(int stage, long n_cells, const uint8_t *update_mask, const double *volume,
                    double dt, double *U, double *W, const double *RHS);
double orc_choose_dt(double cfl, double min_cell_size, double max_eig, double t, double t_max);

/* One full SSP-RK3 step, serial. Returns dt used; max_eig3[3] = per-stage max eigenvalue. */
double orc_step(int problem, long n_cells, long n_ifaces,
                const int64_t *owner, const int64_t *neigh, const int32_t *bc,
                const double *area, const double *normal, const double *icentroid,
                const uint8_t *solved, const uint8_t *update_mask, const double *volume,
                double cfl, double min_cell_size, double t, double t_max,
                double *U, double *W, double *RHS, double max_eig3[3]);

/* ---- error norm (src/main.cpp:550-573) --------------------------------- */
double orc_error_norm(int problem, int dim, long n_cells, const double *ccentroid, const double *volume,
                      const uint8_t *internal, const double *U, double t_max);

/* ---- whole main.cpp flow on a uniform mesh (serial) -------------------- */
/* t_end < 0 -> problem default.  boxes: n_boxes * {xMin,yMin,zMin,xMax,yMax,zMax}.
 * max_steps < 0 -> run to t_end.  U_out (may be NULL): final state [n_cells*5].
 * Returns the number of steps taken; *error_out = Sum |rho - rho_exact(tMax)| V. */
int orc_run(int problem, int dim_in, long n_cells_per_dir, double t_end, double cfl,
            int n_boxes, const double *boxes, int max_steps,
            double *error_out, double *t_out, double *U_out);

/* ---- threaded CPU baseline (reference-faithful face loop + RK loops) --- */
/* Runs n_warmup + n_steps RK3 steps of `problem` on a uniform dim-D mesh with 2^level
 * cells per side on n_threads threads (contiguous Morton chunks, the partition PABLO
 * gives the reference: src/main.cpp:159,183); returns wall seconds for the n_steps
 * timed steps and writes a FNV-1a hash of the final state bits to *state_hash. */
double orc_bench_threads(int problem, int dim, int level, int n_warmup, int n_steps, int n_threads,
                         double cfl, uint64_t *state_hash);

#ifdef __cplusplus
}
#endif
#endif
