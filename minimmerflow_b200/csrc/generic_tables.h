// generic_tables.h -- host side of the generic path, no device code: the tables mmf_create uploads for a mesh
// the host describes interface by interface.  A plain function, so that it is unit-tested on the CPU
// (tools/emu builds it into its library) and shared with the tool that runs the generic kernels' source there.
//
// Replaces the id -> raw-index lookups of the reference's interface loop (src/euler.cpp:150-179) by per-cell
// lists: for every solved cell the interfaces that touch it, in the order the reference processes them
// (interface_order = MeshGeometricalInfo::getInterfaceRawIds(), src/mesh_info.cpp:86-118), skipping the
// interfaces with no solved side exactly like src/euler.cpp:181-183.  Walking a cell's list reproduces the
// sequence of `+=` / `-=` the reference applies to that cell (src/euler.cpp:237-247) without atomics.
#pragma once

#include "../../include/mmf_b200.h"

#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

namespace mmf {

struct GenericTables {
    int64_t n_cells = 0, n_ifaces = 0;
    int64_t stride = 0;                 // SoA field stride: n_cells rounded up to a multiple of 32
    std::vector<int32_t> owner, neigh;  // per interface, raw id indexed; neigh = -1 on the border
    std::vector<int8_t> bc;
    std::vector<double> normal;         // SoA [e * n_ifaces + f]
    std::vector<double> area, volume;
    std::vector<uint8_t> solved, update; // update = solved AND internal (the cells the RK loops touch)
    std::vector<int64_t> ptr;           // [n_cells + 1]
    std::vector<int32_t> ent;           // (interface raw id << 1) | side, side 0 = owner, 1 = neigh
    // a boundary condition on an interface between two SOLVED cells: the reference never builds one (BC_WALL sits
    // between a solved and an unsolved cell, src/main.cpp:251-277) but evaluates it from the owner's state for both
    // sides (src/euler.cpp:198-225); the fused stage kernel assumes the gathering cell is the fluid side
    bool bc_between_solved = false;
};

// returns MMF_OK or MMF_ERR_INVALID with a message in err
inline int build_generic_tables(const mmf_mesh_desc *d, GenericTables &t, std::string &err)
{
    auto bad = [&err](const char *fmt, long long a, int b) {
        char buf[256];
        snprintf(buf, sizeof buf, fmt, a, b);
        err = buf;
        return (int) MMF_ERR_INVALID;
    };
    const int64_t nc = d->n_cells, nf = d->n_interfaces;
    const int64_t n_listed = d->interface_order ? d->n_interfaces_listed : nf;
    t.n_cells = nc;
    t.n_ifaces = nf;
    t.stride = (nc + 31) / 32 * 32;

    t.owner.resize(nf); t.neigh.resize(nf); t.bc.resize(nf);
    t.normal.resize(3 * (size_t) nf);
    for (int64_t f = 0; f < nf; ++f) {
        if (d->owner[f] < 0 || d->owner[f] >= nc || d->neigh[f] >= nc) {
            return bad("mmf_create: interface %lld has owner/neigh out of range", (long long) f, 0);
        }
        t.owner[f] = (int32_t) d->owner[f];
        t.neigh[f] = d->neigh[f] < 0 ? -1 : (int32_t) d->neigh[f];
        if (d->bc[f] < MMF_BC_NONE || d->bc[f] > MMF_BC_DIRICHLET) {
            return bad("mmf_create: interface %lld has unknown BC code %d", (long long) f, d->bc[f]);
        }
        if (t.neigh[f] < 0 && d->bc[f] == MMF_BC_NONE) {
            return bad("mmf_create: border interface %lld has BC_NONE", (long long) f, 0);
        }
        t.bc[f] = (int8_t) d->bc[f];
        for (int e = 0; e < 3; ++e) t.normal[(size_t) e * nf + f] = d->normal[3 * f + e];
    }
    t.area.assign(d->area, d->area + nf);
    t.volume.assign(d->volume, d->volume + nc);

    t.solved.resize(nc); t.update.resize(nc);
    for (int64_t c = 0; c < nc; ++c) {
        t.solved[c] = d->solved[c] ? 1 : 0;
        t.update[c] = (t.solved[c] && (!d->internal || d->internal[c])) ? 1 : 0;
    }
    // cell -> interface lists in processing order (a counting sort keeps the order)
    auto processed = [&t](int64_t f, bool &oS, bool &nS) {
        oS = t.solved[t.owner[f]] != 0;
        nS = t.neigh[f] >= 0 && t.solved[t.neigh[f]] != 0;
        return oS || nS; // src/euler.cpp:181-183
    };
    t.ptr.assign(nc + 1, 0);
    for (int64_t q = 0; q < n_listed; ++q) {
        const int64_t f = d->interface_order ? d->interface_order[q] : q;
        if (f < 0 || f >= nf) return bad("mmf_create: interface_order[%lld] out of range", (long long) q, 0);
        bool oS, nS;
        if (!processed(f, oS, nS)) continue;
        if (oS && nS && t.bc[f] != MMF_BC_NONE) t.bc_between_solved = true;
        if (oS) t.ptr[t.owner[f] + 1]++;
        if (nS) t.ptr[t.neigh[f] + 1]++;
    }
    for (int64_t c = 0; c < nc; ++c) t.ptr[c + 1] += t.ptr[c];
    t.ent.resize((size_t) t.ptr[nc]);
    std::vector<int64_t> cursor(t.ptr.begin(), t.ptr.end() - 1);
    for (int64_t q = 0; q < n_listed; ++q) {
        const int64_t f = d->interface_order ? d->interface_order[q] : q;
        bool oS, nS;
        if (!processed(f, oS, nS)) continue;
        if (oS) t.ent[(size_t) cursor[t.owner[f]]++] = (int32_t) (f << 1);
        if (nS) t.ent[(size_t) cursor[t.neigh[f]]++] = (int32_t) ((f << 1) | 1);
    }
    return MMF_OK;
}

} // namespace mmf
