// driver_b200.cpp -- the device-resident host driver: minimmerflow's set-up, error check and output
// run through the reference's OWN translation units (problem.cpp, body.cpp, mesh_info.cpp,
// solver_writer.cpp, utils.cpp, memory.cpp -- linked unchanged), while the whole time loop
// (src/main.cpp:377-524) is one stream of mmf_step calls with the state living in HBM.
//
// It is what INTEGRATION.md section 2 describes as the patch to main.cpp, as a stand-alone program:
// same command line (argv[1] cells per direction, argv[2] number of saves), same ./settings.xml, same
// log lines for dt / eigenvalues / final error, same background_<N> / final_background_<N> output.
// Written from scratch; only reference HEADERS are included.
#include "binding.hpp"

#include "body.hpp"
#include "memory.hpp"
#include "solver_writer.hpp"

#include <cstdlib>
#include "utils.hpp"

#include <bitpit_IO.hpp>
#include <bitpit_voloctree.hpp>

#include <chrono>
#include <limits>

using namespace bitpit;

namespace {

struct RunSetup {
    problem::ProblemType problemType;
    int dimensions, order;
    std::array<double, 3> origin;
    double length, tMin, tMax, cfl;
    long cellsPerDirection;
    int nSaves;
};

RunSetup readSetup(int argc, char *argv[])
{
    RunSetup s;
    config::reset("minimmerflow", 1);
    config::read("settings.xml");
    s.problemType = problem::getProblemType();
    problem::getDomainData(s.problemType, s.dimensions, &s.origin, &s.length);
    s.tMin  = problem::getStartTime(s.problemType, s.dimensions);
    s.tMax  = problem::getEndTime(s.problemType, s.dimensions);
    s.order = config::root["discretization"]["space"].get<int>("order");
    s.cfl   = config::root["discretization"]["time"].get<double>("CFL");
    s.cellsPerDirection = (argc > 1) ? std::atol(argv[1]) : config::root["discretization"]["space"].get<long>("nCells");
    s.nSaves = (argc > 2) ? std::atoi(argv[2]) : std::numeric_limits<int>::max();
    return s;
}

// Bring the fields the writer streams back to the host.  The fused stages never materialise the
// residual; the reference's cellRHS at output time is the residual of the last stage's input
// (cellConservativesWork, src/main.cpp:468-469), which is still on the device as field W.
void downloadForOutput(mmf_ctx *ctx, int order, bool haveStep, CellStorageDouble *cons, CellStorageDouble *rhs)
{
    if (mmf_get_state(ctx, MMF_FIELD_U, cons->rawData(0)) != MMF_OK) mmf_b200::fail("mmf_get_state", ctx);
    if (!haveStep) return;
    double eig;
    if (mmf_compute_rhs(ctx, MMF_FIELD_W, order, &eig) != MMF_OK) mmf_b200::fail("mmf_compute_rhs", ctx);
    if (mmf_get_state(ctx, MMF_FIELD_RHS, rhs->rawData(0)) != MMF_OK) mmf_b200::fail("mmf_get_state", ctx);
}

// cons -> prim for the writer (what main.cpp does before every mesh.write(), src/main.cpp:511-518): evaluated
// on the device from the resident state; before the first upload there is nothing resident yet
void refreshPrimitives(mmf_ctx *ctx, const std::vector<std::size_t> &cellRawIds, const CellStorageDouble &cons,
                       CellStorageDouble *prim)
{
    // (MMF_DEVICE_PRIMITIVES=0: the reference's host loop instead)
    static const bool onDevice = !(std::getenv("MMF_DEVICE_PRIMITIVES") && std::atoi(std::getenv("MMF_DEVICE_PRIMITIVES")) == 0);
    if (ctx && onDevice) {
        if (mmf_get_primitives(ctx, MMF_FIELD_U, prim->rawData(0)) != MMF_OK) mmf_b200::fail("mmf_get_primitives", ctx);
        return;
    }
    for (std::size_t raw : cellRawIds) ::utils::conservative2primitive(cons.rawData(raw), prim->rawData(raw));
}

} // namespace

int main(int argc, char *argv[])
{
    log::manager().initialize(log::COMBINED, "minimmerflow", true, ".", 1, 0);
    log::cout() << "minimmerflow -- B200 device-resident driver" << std::endl;

    const RunSetup setup = readSetup(argc, argv);
    log::cout() << "Domain: origin " << setup.origin << ", length " << setup.length << "; time " << setup.tMin << " -> "
                << setup.tMax << "; order " << setup.order << ", CFL " << setup.cfl << ", cells per direction "
                << setup.cellsPerDirection << std::endl;

    // ---- mesh, geometry cache, flags, boundary conditions: host side, as in the reference ---------
    VolOctree mesh(setup.dimensions, setup.origin, setup.length, setup.length / setup.cellsPerDirection);
    mesh.initializeAdjacencies();
    mesh.initializeInterfaces();
    mesh.update();
    {
        std::stringstream name;
        name << "background_" << setup.cellsPerDirection;
        mesh.getVTK().setName(name.str());
    }
    body::initialize();
    MeshGeometricalInfo meshInfo(&mesh);
    const std::vector<std::size_t> &cellRawIds = meshInfo.getCellRawIds();
    const std::vector<std::size_t> &internalCellRawIds = meshInfo.getInternalCellRawIds();
    const std::vector<std::size_t> &interfaceRawIds = meshInfo.getInterfaceRawIds();

    CellStorageBool cellSolvedFlag(1, &mesh.getCells()), cellFluidFlag(1, &mesh.getCells());
    CellStorageDouble cellPrimitives(N_FIELDS, &mesh.getCells()), cellConservatives(N_FIELDS, &mesh.getCells());
    CellStorageDouble cellRHS(N_FIELDS, &mesh.getCells());
    for (std::size_t raw : cellRawIds) {
        const bool fluid = body::isPointFluid(meshInfo.rawGetCellCentroid(raw));
        cellFluidFlag.rawSet(raw, fluid);
        cellSolvedFlag.rawSet(raw, fluid && mesh.getCells().rawAt(raw).isInterior());
    }
    InterfaceStorageInt interfaceBCs(1, &mesh.getInterfaces());
    for (std::size_t raw : interfaceRawIds) {
        const Interface &face = mesh.getInterfaces().rawAt(raw);
        if (face.isBorder()) {
            interfaceBCs.rawAt(raw) = problem::getBorderBCType(setup.problemType, face.getId(), meshInfo);
        } else {
            const bool ownerFluid = cellFluidFlag.rawAt(mesh.getCellConstIterator(face.getOwner()).getRawIndex());
            const bool neighFluid = cellFluidFlag.rawAt(mesh.getCellConstIterator(face.getNeigh()).getRawIndex());
            interfaceBCs.rawAt(raw) = (ownerFluid != neighFluid) ? BC_WALL : BC_NONE;
        }
    }

    SolverWriter writer(&mesh, &cellPrimitives, &cellConservatives, &cellRHS, &cellSolvedFlag);
    VTKUnstructuredGrid &vtk = mesh.getVTK();
    vtk.setCounter(0);
    vtk.addData<int>("solved", VTKFieldType::SCALAR, VTKLocation::CELL, &writer);
    vtk.addData<double>("velocity", VTKFieldType::VECTOR, VTKLocation::CELL, &writer);
    for (const char *name : { "pressure", "temperature", "density", "residualC", "residualMX", "residualMY", "residualMZ", "residualE" }) {
        vtk.addData<double>(name, VTKFieldType::SCALAR, VTKLocation::CELL, &writer);
    }

    for (std::size_t raw : cellRawIds) {
        problem::evalCellInitalConservatives(setup.problemType, mesh.getCells().rawAt(raw), meshInfo, cellConservatives.rawData(raw));
    }
    refreshPrimitives(nullptr, cellRawIds, cellConservatives, &cellPrimitives); // nothing is resident yet
    mesh.write();

    double minCellSize = std::numeric_limits<double>::max();
    for (std::size_t raw : internalCellRawIds) minCellSize = std::min(minCellSize, meshInfo.rawGetCellSize(raw));
    if (setup.order != 1) {
        log::cout() << "Reconstruction order " << setup.order << " is not supported." << std::endl;
        return 2; // what reconstruction::eval does in the reference (src/reconstruction.cpp:76)
    }

    // ---- the time loop, device resident -----------------------------------------------------------
    mmf_ctx *ctx = mmf_b200::createContext(setup.problemType, meshInfo, cellSolvedFlag, interfaceBCs);
    if (mmf_set_state(ctx, MMF_FIELD_U, cellConservatives.rawData(0)) != MMF_OK) mmf_b200::fail("mmf_set_state", ctx);
    log_memory_status();

    const auto wallStart = std::chrono::steady_clock::now();
    double outputSeconds = 0.;
    int step = 0;
    double t = setup.tMin, nextSave = setup.tMin;
    while (t < setup.tMax) {
        double dt, maxEig[3];
        if (mmf_step(ctx, setup.cfl, minCellSize, t, setup.tMax, &dt, maxEig) != MMF_OK) mmf_b200::fail("mmf_step", ctx);
        log::cout() << "Step n. " << step << std::endl;
        log::cout() << "Using dt= " << dt << " (maxEig = " << maxEig[0] << ", " << maxEig[1] << ", " << maxEig[2] << ")" << std::endl;
        t += dt;
        ++step;
        if (t > nextSave) { // output only when due; the state comes back to the host for it
            const auto t0 = std::chrono::steady_clock::now();
            downloadForOutput(ctx, setup.order, true, &cellConservatives, &cellRHS);
            refreshPrimitives(ctx, cellRawIds, cellConservatives, &cellPrimitives);
            mesh.write();
            outputSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            nextSave += (setup.tMax - setup.tMin) / setup.nSaves;
        }
    }
    downloadForOutput(ctx, setup.order, step > 0, &cellConservatives, &cellRHS);
    refreshPrimitives(ctx, cellRawIds, cellConservatives, &cellPrimitives); // for the final write below
    const double wallSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - wallStart).count();
    mmf_info info;
    mmf_get_info(ctx, &info);
    mmf_destroy(ctx);

    {
        std::stringstream name;
        name << "final_background_" << setup.cellsPerDirection;
        mesh.write(name.str());
    }
    log::cout() << "Computation time (without disk saving time) is " << wallSeconds - outputSeconds << std::endl;
    log::cout() << "Disk time " << outputSeconds << std::endl;
    log::cout() << "Device path " << (info.path == MMF_PATH_UNIFORM ? "uniform" : "generic") << ", " << info.kernel_launches
                << " kernel launches" << std::endl;

    // ---- error check (src/main.cpp:550-573) ---------------------------------------------------------
    std::array<double, N_FIELDS> exact;
    double error = 0.;
    for (std::size_t raw : internalCellRawIds) {
        const Cell &cell = mesh.getCells().rawAt(raw);
        problem::evalCellExactConservatives(setup.problemType, cell, meshInfo, setup.tMax, exact.data());
        error += std::abs(cellConservatives.rawData(raw)[FID_RHO] - exact[FID_RHO]) * meshInfo.getCellVolume(cell.getId());
    }
    log::cout() << std::endl << " ::::::::: Error check :::::::::" << std::endl << std::endl;
    log::cout() << " Final error:  " << std::setprecision(12) << std::scientific << error << std::endl << std::endl;
    return 0;
}
