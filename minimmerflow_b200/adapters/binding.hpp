// binding.hpp -- describes a bitpit-hosted minimmerflow mesh to libmmf_b200.so (include/mmf_b200.h).
// Shared by the strict drop-in adapters (solver_b200.cpp) and the device-resident driver
// (driver_b200.cpp).  Only reference HEADERS are used (from the reference tree on the include path).
#ifndef MMF_B200_ADAPTERS_BINDING_HPP
#define MMF_B200_ADAPTERS_BINDING_HPP

#include "euler.hpp"
#include "reconstruction.hpp"

#include "mmf_b200.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>

namespace mmf_b200 {

[[noreturn]] inline void fail(const char *what, mmf_ctx *ctx)
{
    throw std::runtime_error(std::string(what) + ": " + mmf_last_error(ctx));
}

// Describe the host mesh to the library: what MeshGeometricalInfo caches plus the flag / BC tables
// main.cpp builds (src/main.cpp:221-237, 251-277).  Everything is addressed by RAW id.
inline mmf_ctx *createContext(problem::ProblemType problemType, const MeshGeometricalInfo &meshInfo,
                       const CellStorageBool &cellSolvedFlag, const InterfaceStorageInt &interfaceBCs)
{
    const bitpit::VolumeKernel &mesh = meshInfo.getPatch();
    const std::vector<std::size_t> &cellRawIds = meshInfo.getCellRawIds();
    const std::vector<std::size_t> &interfaceRawIds = meshInfo.getInterfaceRawIds();

    std::size_t nCellSlots = 0, nInterfaceSlots = 0;
    for (std::size_t raw : cellRawIds) nCellSlots = std::max(nCellSlots, raw + 1);
    for (std::size_t raw : interfaceRawIds) nInterfaceSlots = std::max(nInterfaceSlots, raw + 1);

    std::vector<double> volume(nCellSlots, 1.);
    std::vector<std::uint8_t> solved(nCellSlots, 0), internal(nCellSlots, 0);
    std::vector<std::int32_t> ijk(3 * nCellSlots, 0);

    // structured hint: integer lattice coordinates when every cell has the same size
    bool uniform = mesh.getDimension() == 3 && !cellRawIds.empty();
    const double h = cellRawIds.empty() ? 1. : meshInfo.rawGetCellSize(cellRawIds[0]);
    std::array<double, 3> lo = { { 0., 0., 0. } };
    if (uniform) {
        lo = meshInfo.rawGetCellCentroid(cellRawIds[0]);
        for (std::size_t raw : cellRawIds) {
            const std::array<double, 3> &c = meshInfo.rawGetCellCentroid(raw);
            for (int d = 0; d < 3; ++d) lo[d] = std::min(lo[d], c[d]);
            uniform = uniform && meshInfo.rawGetCellSize(raw) == h;
        }
    }
    std::int32_t dims[3] = { 0, 0, 0 };
    for (std::size_t raw : cellRawIds) {
        const bitpit::Cell &cell = mesh.getCells().rawAt(raw);
        volume[raw]   = meshInfo.rawGetCellVolume(raw);
        solved[raw]   = cellSolvedFlag.rawAt(raw) ? 1 : 0; // element-wise: the bool storage has no raw pointer
        internal[raw] = cell.isInterior() ? 1 : 0;
        if (uniform) {
            const std::array<double, 3> &c = meshInfo.rawGetCellCentroid(raw);
            for (int d = 0; d < 3; ++d) {
                ijk[3 * raw + d] = (std::int32_t) std::llround((c[d] - lo[d]) / h);
                dims[d] = std::max(dims[d], ijk[3 * raw + d] + 1);
            }
        }
    }

    std::vector<std::int64_t> owner(nInterfaceSlots, 0), neigh(nInterfaceSlots, -1), order(interfaceRawIds.size());
    std::vector<std::int32_t> bc(nInterfaceSlots, BC_FREE_FLOW);
    std::vector<double> area(nInterfaceSlots, 0.), normal(3 * nInterfaceSlots, 0.);
    for (std::size_t q = 0; q < interfaceRawIds.size(); ++q) {
        const std::size_t raw = interfaceRawIds[q];
        const bitpit::Interface &interface = mesh.getInterfaces().rawAt(raw);
        order[q] = (std::int64_t) raw;
        owner[raw] = (std::int64_t) mesh.getCellConstIterator(interface.getOwner()).getRawIndex();
        const long neighId = interface.getNeigh();
        neigh[raw] = (neighId >= 0) ? (std::int64_t) mesh.getCellConstIterator(neighId).getRawIndex() : -1;
        bc[raw]    = interfaceBCs.rawAt(raw);
        area[raw]  = meshInfo.rawGetInterfaceArea(raw);
        const std::array<double, 3> &n = meshInfo.rawGetInterfaceNormal(raw);
        for (int d = 0; d < 3; ++d) normal[3 * raw + d] = n[d];
    }

    mmf_mesh_desc desc = {};
    desc.struct_size  = sizeof desc;
    desc.dim          = mesh.getDimension();
    desc.problem_type = (std::int32_t) problemType;
    desc.n_cells      = (std::int64_t) nCellSlots;
    desc.n_interfaces = (std::int64_t) nInterfaceSlots;
    desc.interface_order     = order.data();
    desc.n_interfaces_listed = (std::int64_t) order.size();
    desc.owner = owner.data();  desc.neigh = neigh.data();  desc.bc = bc.data();
    desc.area  = area.data();   desc.normal = normal.data();
    desc.volume = volume.data(); desc.solved = solved.data(); desc.internal = internal.data();
    // BC_DIRICHLET data (src/problem.cpp:450-477; only filled for the forward-facing step)
    std::array<double, BC_INFO_SIZE> info;
    info.fill(0.);
    problem::getBorderBCInfo(problemType, BC_DIRICHLET, { { 0., 0., 0. } }, { { 1., 0., 0. } }, info);
    for (int k = 0; k < N_FIELDS; ++k) desc.dirichlet_info[k] = info[k];
    if (uniform && (std::size_t) dims[0] * dims[1] * dims[2] == cellRawIds.size() && cellRawIds.size() == nCellSlots) {
        desc.cell_ijk = ijk.data();
        for (int d = 0; d < 3; ++d) { desc.box_dims[d] = desc.global_dims[d] = dims[d]; desc.box_offset[d] = 0; }
    }

    const char *device = std::getenv("MMF_DEVICE");
    mmf_ctx *ctx = nullptr;
    if (mmf_create(&desc, device ? std::atoi(device) : 0, &ctx) != MMF_OK) fail("mmf_create", nullptr);
    return ctx;
}

// solver_b200.cpp: drops the device context of the strict adapters; the next computeRHS / computePolynomials call
// describes the mesh, the solved flags and the BC table again.  For hosts that edit those tables in place.
void invalidateContext();

} // namespace mmf_b200

#endif
