// uniform_body_cells.cuh -- a uniform box WITH BODIES on the fused path: the pass over the wall cells that runs around the
// stage kernel (kernel form 'b': uniform_stage_t.cuh with BODY = true).
//
// The reference marks the cells whose centroid lies inside a body box as not solved (src/main.cpp:221-237,
// src/body.cpp:80-95), gives every interface between a fluid and a solid cell BC_WALL (src/main.cpp:251-277)
// and, in euler::computeRHS, (a) skips interfaces whose two cells are both not solved (src/euler.cpp:181-183),
// (b) evaluates a wall interface between the fluid cell's state and its mirror image about the interface
// normal as seen from the fluid side (:198-225, :352-362, :322-339), (c) accumulates into solved cells only
// (:237-247); the RK loops skip the cells that are not solved (src/main.cpp:409-423).  The interface ids,
// hence the accumulation order of a fluid cell, do not depend on any of this.
//
// The stage kernel evaluates a wall like any interface and does not store the cells that touch one (flag 2) nor the
// solid ones (flag 1): the wall cells -- a surface -- are recomputed reference-shaped by wall_cell_update
// (uniform_device.cuh; uniform_wall_cells_kernel before the stage kernel, uniform_wall_scatter_kernel behind it), so
// the hot path has no slow path, no call and no divergent branch.  (A variant with the wall evaluation inside the
// stage kernel -- a call at each face whose sides differ: spills around three call sites per plane -- took twice as
// long and was removed in round 1; the rotate-form body kernel, form 'c', lost to form 'b' in round 2:
// profiles/r02k_bodies.md.)
#pragma once

#include "uniform_stage_v5.cuh"

namespace mmf {

// the two small kernels around the stage kernel: the wall cells' results into a compact buffer
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
