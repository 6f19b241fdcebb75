// generic_types.cuh -- plain data shared by host code and kernels: the generic mesh description on the
// device and the step control block.  No kernels here, so that every translation unit can include it.
#pragma once

#include "gas.cuh"

#include <stdint.h>

namespace mmf {

struct GenericMesh {
    int64_t n_cells;
    int64_t n_ifaces;
    int64_t stride;           // SoA field stride (>= n_cells)
    // cell -> ordered interface entries, entry = (interface raw id << 1) | side (0 owner, 1 neigh);
    // only processed interfaces of solved cells are listed
    const int64_t *cf_ptr;    // [n_cells + 1]
    const int32_t *cf_ent;
    // per interface, raw id indexed
    const int32_t *f_owner;
    const int32_t *f_neigh;   // -1 border
    const int8_t  *f_bc;
    const double  *f_area;
    const double  *f_normal;  // SoA [e * n_ifaces + f]
    // per cell
    const uint8_t *c_solved;
    const uint8_t *c_update;  // internal AND solved
    const double  *c_volume;
    double dirichlet_info[NF];
};

struct StepControl {
    double t, dt, t_max, cfl, min_h, steps, active; // the first seven are (re)set by the host per call
    double max_eig[3];   // per-stage max face eigenvalue (src/main.cpp:399, :440, :476)
    double max_eig_chk;  // uniform path: stage-1 face maximum re-derived by the fused kernel
    // uniform path: max eigenvalue of the state stage 3 wrote, found right behind stage 3 from its
    // per-tile FP32 estimates (uniform_eig_select_kernel / uniform_eig_tiles_kernel)
    double eig_next;
    double est_max;      // largest FP32 eigenvalue estimate over the tiles (all ranks after the all-reduce)
    double mismatches;   // sticky: steps whose dt eigenvalue differed from stage 1's own face maximum
    double halo_timeouts; // sticky: halo waits that gave up on a neighbour rank's layer (results are then invalid)
};
constexpr int STEP_CONTROL_HOST_FIELDS = 7;

} // namespace mmf
