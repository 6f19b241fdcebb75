// solver_b200.cpp -- drop-in definitions of the reference's solver entry points
//
//     reconstruction::initialize / computePolynomials   (src/reconstruction.hpp:39-42)
//     euler::computeRHS                                 (src/euler.hpp:41-43)
//
// that forward to libmmf_b200.so through its C-ABI (include/mmf_b200.h).  A maintainer compiles
// this file INSTEAD OF src/euler.cpp and src/reconstruction.cpp; main.cpp, problem.cpp, body.cpp,
// mesh_info.cpp, solver_writer.cpp, ... stay untouched.  Only reference HEADERS
// are included, from the reference tree on the include path; nothing of the reference is copied.
//
// This is the STRICT mode of the boundary: the caller owns every storage on the host (bitpit
// PiercedStorage, AoS), so each computeRHS call uploads the conservative field it is handed and
// downloads the residual (PCIe bound by construction).  The device-resident mode that replaces the
// whole `while (t < tMax)` body by one call is driver_b200.cpp.
//
// Error behaviour mirrors the reference: an unsupported reconstruction order ends the process with
// exit(2) like reconstruction::eval (src/reconstruction.cpp:76); any other library failure throws
// std::runtime_error (the reference's own failure mode for bad input, e.g. src/problem.cpp:57).
#include "euler.hpp"
#include "reconstruction.hpp"

#include "mmf_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>

namespace {

struct Binding {
    const MeshGeometricalInfo *meshInfo = nullptr;
    std::size_t nCells = 0, nInterfaces = 0;
    mmf_ctx *ctx = nullptr;
    ~Binding() { if (ctx) mmf_destroy(ctx); }
};

Binding g_binding;

[[noreturn]] void fail(const char *what, mmf_ctx *ctx)
{
    throw std::runtime_error(std::string(what) + ": " + mmf_last_error(ctx));
}

// Describe the host mesh to the library: what MeshGeometricalInfo caches plus the flag / BC tables
// main.cpp builds (src/main.cpp:221-237, 251-277).  Everything is addressed by RAW id.
mmf_ctx *createContext(problem::ProblemType problemType, const MeshGeometricalInfo &meshInfo,
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

mmf_ctx *context(problem::ProblemType problemType, const MeshGeometricalInfo &meshInfo,
                 const CellStorageBool &cellSolvedFlag, const InterfaceStorageInt &interfaceBCs)
{
    Binding &b = g_binding;
    const std::size_t nCells = meshInfo.getCellRawIds().size(), nInterfaces = meshInfo.getInterfaceRawIds().size();
    if (b.ctx && (b.meshInfo != &meshInfo || b.nCells != nCells || b.nInterfaces != nInterfaces)) {
        mmf_destroy(b.ctx); // the mesh changed: describe it again
        b.ctx = nullptr;
    }
    if (!b.ctx) {
        b.ctx = createContext(problemType, meshInfo, cellSolvedFlag, interfaceBCs);
        b.meshInfo = &meshInfo;
        b.nCells = nCells;
        b.nInterfaces = nInterfaces;
    }
    return b.ctx;
}

} // namespace

namespace reconstruction {

void initialize()
{
}

void computePolynomials(problem::ProblemType problemType, const MeshGeometricalInfo &meshInfo, const CellStorageBool &cellSolved,
                        const CellStorageDouble &conservativeFields, const InterfaceStorageInt &interfaceBCs)
{
    // order 1: face state = cell mean, nothing to prepare (the reference's function is empty too,
    // src/reconstruction.cpp:47-55); the C-ABI twin is called to keep the call sequence observable
    BITPIT_UNUSED(conservativeFields);
    mmf_ctx *ctx = context(problemType, meshInfo, cellSolved, interfaceBCs);
    if (mmf_compute_polynomials(ctx, MMF_FIELD_U) != MMF_OK) fail("mmf_compute_polynomials", ctx);
}

} // namespace reconstruction

namespace euler {

void computeRHS(problem::ProblemType problemType, const MeshGeometricalInfo &meshInfo, const CellStorageBool &cellSolvedFlag,
                const int order, const CellStorageDouble &cellConservatives, const InterfaceStorageInt &interfaceBCs,
                CellStorageDouble *cellsRHS, double *maxEig)
{
    mmf_ctx *ctx = context(problemType, meshInfo, cellSolvedFlag, interfaceBCs);
    // PiercedStorage<double> keeps (raw position p, field k) at [p*nFields + k]: one contiguous block
    const int status = mmf_compute_rhs_host(ctx, cellConservatives.rawData(0), order, cellsRHS->rawData(0), maxEig);
    if (status == MMF_ERR_UNSUPPORTED_ORDER) {
        bitpit::log::cout() << "Reconstruction order " << order << " is not supported." << std::endl;
        std::exit(2); // what reconstruction::eval does (src/reconstruction.cpp:76)
    }
    if (status != MMF_OK) fail("mmf_compute_rhs_host", ctx);
}

} // namespace euler
