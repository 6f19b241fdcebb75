// uniform_eligibility.h -- is a host mesh description a full, conforming, uniform 3-D box whose interface
// numbering matches a known convention?  Plain host code (no CUDA calls, no library state): uniform_path.cuh
// builds the fused path from the answer, and tools/emu compiles the same function so that the dispatcher's
// decision is unit-tested on the CPU (tests/test_emu_protocol.py).
#pragma once

#include "../../include/mmf_b200.h"

#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

namespace mmf {

// (NUM_MORTON / NUM_LEXI / NUM_AXIS come from uniform_device.cuh, which every includer has seen before)

struct UniformBoxAnalysis {
    bool eligible = false;
    bool bodies = false;          // some cells are not solved (only ever true when allow_bodies)
    int numbering = -1;           // NUM_MORTON / NUM_LEXI / NUM_AXIS: the order the stage kernels accumulate in
    int order_exact = 1;          // 0: MMF_FLAG_ORDER_AXIS asked for the axis order (not the reference's bits)
    int bc_side[6] = { -9, -9, -9, -9, -9, -9 };
    double area = 0., volume = 0., h = 0.;
};

// Per-cell order in which the reference's interface loop touches the six faces, predicted from a
// numbering convention; slots: 0 -x, 1 +x, 2 -y, 3 +y, 4 -z, 5 +z.
inline void predicted_face_order(int numbering, const int ijk[3], int order[6])
{
    int n = 0;
    int lows[3], keys[3], nl = 0;
    for (int a = 0; a < 3; ++a) {
        if (ijk[a] == 0) continue;
        lows[nl] = a;
        if (numbering == NUM_MORTON) keys[nl] = 3 * __builtin_ctz((unsigned) ijk[a]) + a;
        else                         keys[nl] = a == 2 ? 2 : a == 1 ? 1 : 0; // lexicographic: z, y, x
        nl++;
    }
    // descending key first
    for (int a = 0; a < nl; ++a)
        for (int b = a + 1; b < nl; ++b)
            if (keys[b] > keys[a]) { std::swap(keys[a], keys[b]); std::swap(lows[a], lows[b]); }
    for (int a = 0; a < nl; ++a) order[n++] = 2 * lows[a];
    for (int a = 0; a < 3; ++a) {
        if (ijk[a] == 0) order[n++] = 2 * a;
        order[n++] = 2 * a + 1;
    }
}

// All cells solved -- or, with allow_bodies, a box with bodies: cells that are not solved, and BC_WALL on
// exactly the interfaces between a solved and an unsolved cell (src/main.cpp:221-237, 251-277).
inline UniformBoxAnalysis analyze_uniform_box(const mmf_mesh_desc *d, const bool allow_bodies)
{
    UniformBoxAnalysis r;
    if (!d->cell_ijk || d->dim != 3 || (d->flags & MMF_FLAG_FORCE_GENERIC)) return r;
    const int nx = d->box_dims[0], ny = d->box_dims[1], nz = d->box_dims[2];
    if (nx <= 0 || ny <= 0 || nz <= 0) return r;
    const int64_t nc = d->n_cells, nf = d->n_interfaces;
    if ((int64_t) nx * ny * nz != nc) return r;
    for (int e = 0; e < 3; ++e) {
        if (d->global_dims[e] != d->box_dims[e] || d->box_offset[e] != 0) return r; // single-box only
    }
    const int64_t nf_expected = (int64_t) (nx + 1) * ny * nz + (int64_t) nx * (ny + 1) * nz + (int64_t) nx * ny * (nz + 1);
    if (nf != nf_expected) return r;
    if (d->interface_order && d->n_interfaces_listed != nf) return r;

    // cells: a bijection onto the lattice, all internal, one volume; all solved unless bodies are allowed
    std::vector<int64_t> lattice_to_raw((size_t) nc, -1);
    const double V = d->volume[0];
    for (int64_t c = 0; c < nc; ++c) {
        const int i = d->cell_ijk[3 * c], j = d->cell_ijk[3 * c + 1], k = d->cell_ijk[3 * c + 2];
        if (i < 0 || i >= nx || j < 0 || j >= ny || k < 0 || k >= nz) return r;
        const int64_t l = ((int64_t) k * ny + j) * nx + i;
        if (lattice_to_raw[l] >= 0) return r;
        lattice_to_raw[l] = c;
        if ((d->internal && !d->internal[c]) || d->volume[c] != V) return r;
        if (!d->solved[c]) {
            if (!allow_bodies) return r;
            r.bodies = true;
        }
    }
    const double A = d->area[0];
    const double h = std::sqrt(A);

    // interfaces: axis-aligned unit normals owner->neigh between lattice neighbours, one area,
    // one BC per side; record for every cell the position of each of its six faces
    std::vector<int64_t> face_pos((size_t) nc * 6, -1);
    int bc_side[6] = { -9, -9, -9, -9, -9, -9 };
    for (int64_t q = 0; q < nf; ++q) {
        const int64_t f = d->interface_order ? d->interface_order[q] : q;
        if (f < 0 || f >= nf) return r;
        const int64_t o = d->owner[f], n = d->neigh[f];
        if (o < 0 || o >= nc || n >= nc || d->area[f] != A) return r;
        int axis = -1, sgn = 0;
        for (int e = 0; e < 3; ++e) {
            const double v = d->normal[3 * f + e];
            if (v == 1.0 || v == -1.0) { if (axis >= 0) return r; axis = e; sgn = (int) v; }
            else if (v != 0.0) return r;
        }
        if (axis < 0) return r;
        const int *oc = &d->cell_ijk[3 * o];
        if (n >= 0) {
            const int *ncell = &d->cell_ijk[3 * n];
            for (int e = 0; e < 3; ++e) {
                if (ncell[e] - oc[e] != (e == axis ? sgn : 0)) return r;
            }
            const bool wall = (d->solved[o] != 0) != (d->solved[n] != 0);
            if (d->bc[f] != (wall ? MMF_BC_WALL : MMF_BC_NONE)) return r;
            const int so = 2 * axis + (sgn > 0 ? 1 : 0), sn = 2 * axis + (sgn > 0 ? 0 : 1);
            if (face_pos[o * 6 + so] >= 0 || face_pos[n * 6 + sn] >= 0) return r;
            face_pos[o * 6 + so] = q;
            face_pos[n * 6 + sn] = q;
        } else {
            const int side = 2 * axis + (sgn > 0 ? 1 : 0);
            const int lim = (axis == 0 ? nx : axis == 1 ? ny : nz) - 1;
            if (oc[axis] != (sgn > 0 ? lim : 0)) return r; // outward normal on the matching side
            if (d->bc[f] < MMF_BC_FREE_FLOW || d->bc[f] > MMF_BC_DIRICHLET) return r;
            if (bc_side[side] == -9) bc_side[side] = d->bc[f];
            else if (bc_side[side] != d->bc[f]) return r;
            if (face_pos[o * 6 + side] >= 0) return r;
            face_pos[o * 6 + side] = q;
        }
    }
    for (size_t x = 0; x < face_pos.size(); ++x) if (face_pos[x] < 0) return r;

    // which numbering convention reproduces the host's per-cell interface order?
    int numbering = -1;
    for (int cand = 0; cand < 2 && numbering < 0; ++cand) {
        bool ok = true;
        for (int64_t c = 0; c < nc && ok; ++c) {
            int order[6];
            predicted_face_order(cand, &d->cell_ijk[3 * c], order);
            for (int s = 0; s + 1 < 6; ++s) {
                if (face_pos[c * 6 + order[s]] >= face_pos[c * 6 + order[s + 1]]) { ok = false; break; }
            }
        }
        if (ok) numbering = cand;
    }
    int order_exact = 1;
    if (numbering < 0) {
        if (!(d->flags & MMF_FLAG_ORDER_AXIS)) return r; // unknown order: stay on the exact generic path
        numbering = NUM_AXIS;
        order_exact = 0;
    }
    if (d->flags & MMF_FLAG_ORDER_AXIS) { numbering = NUM_AXIS; order_exact = 0; }
    r.eligible = true;
    r.numbering = numbering;
    r.order_exact = order_exact;
    for (int s = 0; s < 6; ++s) r.bc_side[s] = bc_side[s];
    r.area = A;
    r.volume = V;
    r.h = h;
    return r;
}

} // namespace mmf
