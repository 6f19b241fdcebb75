// bitpit_patchkernel.hpp -- minimal stand-in for bitpit's patch kernel layer (see README.md):
// Cell / Interface, PatchKernel / VolumeKernel, PatchInfo and the VTK streaming hooks.
#ifndef MMF_COMPAT_BITPIT_PATCHKERNEL_HPP
#define MMF_COMPAT_BITPIT_PATCHKERNEL_HPP

#include "bitpit_common.hpp"
#include "bitpit_containers.hpp"
#include "bitpit_IO.hpp"

namespace bitpit {

class Cell {
public:
    Cell() = default;
    Cell(long id, bool interior) : m_id(id), m_interior(interior) {}
    long getId() const { return m_id; }
    bool isInterior() const { return m_interior; }

private:
    long m_id = -1;
    bool m_interior = true;
};

class Interface {
public:
    Interface() = default;
    Interface(long id, long owner, int ownerFace, long neigh, int neighFace)
        : m_id(id), m_owner(owner), m_neigh(neigh), m_ownerFace(ownerFace), m_neighFace(neighFace) {}
    long getId() const { return m_id; }
    long getOwner() const { return m_owner; }
    long getNeigh() const { return m_neigh; } // negative for a border interface
    int getOwnerFace() const { return m_ownerFace; }
    int getNeighFace() const { return m_neighFace; }
    bool isBorder() const { return m_neigh < 0; }

private:
    long m_id = -1, m_owner = -1, m_neigh = -1;
    int m_ownerFace = -1, m_neighFace = -1;
};

// ---- VTK streaming hooks ------------------------------------------------------------------------
enum class VTKFormat { ASCII, APPENDED };
enum class VTKFieldType { SCALAR = 1, VECTOR = 3 };
enum class VTKLocation { CELL, POINT };

class VTKBaseStreamer {
public:
    virtual ~VTKBaseStreamer() = default;
    virtual void flushData(std::fstream &stream, const std::string &name, VTKFormat format) = 0;
};

namespace genericIO {
template <typename T>
void flushBINARY(std::fstream &stream, const T &value)
{
    stream.write(reinterpret_cast<const char *>(&value), sizeof(T));
}
} // namespace genericIO

template <typename T> struct VTKTypeName;
template <> struct VTKTypeName<int> { static const char *name() { return "Int32"; } };
template <> struct VTKTypeName<long> { static const char *name() { return "Int64"; } };
template <> struct VTKTypeName<float> { static const char *name() { return "Float32"; } };
template <> struct VTKTypeName<double> { static const char *name() { return "Float64"; } };

class VTKUnstructuredGrid {
public:
    struct Field {
        std::string name, type;
        int components;
        std::size_t bytesPerValue;
        VTKLocation location;
        VTKBaseStreamer *streamer;
    };

    void setName(const std::string &name) { m_name = name; }
    const std::string &getName() const { return m_name; }
    void setDirectory(const std::string &dir) { m_directory = dir; }
    void setCounter(int counter = 0) { m_counter = counter; }
    int unsetCounter() { int c = m_counter; m_counter = -1; return c; }
    int getCounter() const { return m_counter; }

    template <typename T>
    void addData(const std::string &name, VTKFieldType fieldType, VTKLocation location, VTKBaseStreamer *streamer)
    {
        m_fields.push_back(Field{ name, VTKTypeName<T>::name(), (int) fieldType, sizeof(T), location, streamer });
    }
    const std::vector<Field> &getFields() const { return m_fields; }

    // "<dir>/<name>.<counter, 4 digits>.vtu" when a counter is set, "<dir>/<name>.vtu" otherwise
    std::string nextFileName()
    {
        std::ostringstream s;
        s << m_directory << "/" << m_name;
        if (m_counter >= 0) s << "." << std::setfill('0') << std::setw(4) << m_counter++;
        s << ".vtu";
        return s.str();
    }

private:
    std::string m_name = "mesh", m_directory = ".";
    int m_counter = -1;
    std::vector<Field> m_fields;
};

// ---- patches ------------------------------------------------------------------------------------
class PatchKernel {
public:
    typedef PiercedVector<Cell>::const_iterator CellConstIterator;
    typedef PiercedVector<Cell>::iterator CellIterator;
    typedef PiercedVector<Interface>::const_iterator InterfaceConstIterator;

    virtual ~PatchKernel() = default;

    int getDimension() const { return m_dimension; }
    long getCellCount() const { return (long) m_cells.size(); }
    long getInternalCellCount() const { return (long) m_cells.size(); } // serial stand-in: no ghosts
    long getInterfaceCount() const { return (long) m_interfaces.size(); }

    PiercedVector<Cell> &getCells() { return m_cells; }
    const PiercedVector<Cell> &getCells() const { return m_cells; }
    PiercedVector<Interface> &getInterfaces() { return m_interfaces; }
    const PiercedVector<Interface> &getInterfaces() const { return m_interfaces; }

    CellConstIterator cellConstBegin() const { return m_cells.begin(); }
    CellConstIterator cellConstEnd() const { return m_cells.end(); }
    CellConstIterator getCellConstIterator(long id) const { return m_cells.find(id); }
    InterfaceConstIterator interfaceConstBegin() const { return m_interfaces.begin(); }
    InterfaceConstIterator interfaceConstEnd() const { return m_interfaces.end(); }
    InterfaceConstIterator getInterfaceConstIterator(long id) const { return m_interfaces.find(id); }

    bool isPartitioned() const { return false; }
    int getRank() const { return 0; }
    int getProcessorCount() const { return 1; }

    virtual void initializeAdjacencies() {}
    virtual void initializeInterfaces() {}
    virtual void update() {}
    virtual void markCellForRefinement(long id)
    {
        BITPIT_UNUSED(id);
        throw std::runtime_error("bitpit stand-in: mesh adaption is not supported");
    }

    VTKUnstructuredGrid &getVTK() { return m_vtk; }
    void write();
    void write(const std::string &name);

protected:
    int m_dimension = 3;
    PiercedVector<Cell> m_cells;
    PiercedVector<Interface> m_interfaces;
    VTKUnstructuredGrid m_vtk;

    // geometry the VTK writer needs: vertex coordinates and cell -> vertex connectivity
    virtual void _vtkGeometry(std::vector<double> *points, std::vector<long> *connectivity, int *verticesPerCell, int *vtkCellType) const = 0;

private:
    void _writeVTU(const std::string &fileName);
};

class VolumeKernel : public PatchKernel {
public:
    virtual double evalCellVolume(long id) const = 0;
    virtual double evalCellSize(long id) const = 0;
    virtual std::array<double, 3> evalCellCentroid(long id) const = 0;
    virtual double evalInterfaceArea(long id) const = 0;
    virtual std::array<double, 3> evalInterfaceCentroid(long id) const = 0;
    virtual std::array<double, 3> evalInterfaceNormal(long id) const = 0;
};

class PatchInfo {
public:
    virtual ~PatchInfo() = default;
    PatchKernel const &getPatch() const { return *m_patch; }
    void setPatch(PatchKernel const *patch) { m_patch = patch; }
    void reset() { _reset(); }
    void extract() { if (m_patch) _extract(); }
    void update() { reset(); extract(); }

protected:
    PatchKernel const *m_patch;
    PatchInfo(PatchKernel const *patch) : m_patch(patch) {}
    virtual void _init() = 0;
    virtual void _reset() = 0;
    virtual void _extract() = 0;
};

// ---- VTU output: XML header + appended raw binary, every field streamed by its VTKBaseStreamer --
inline void PatchKernel::write()
{
    _writeVTU(m_vtk.nextFileName());
}

inline void PatchKernel::write(const std::string &name)
{
    const std::string old = m_vtk.getName();
    const int counter = m_vtk.unsetCounter();
    m_vtk.setName(name);
    _writeVTU(m_vtk.nextFileName());
    m_vtk.setName(old);
    m_vtk.setCounter(counter);
}

inline void PatchKernel::_writeVTU(const std::string &fileName)
{
    const char *env = std::getenv("BITPIT_SHIM_VTK");
    if (env && std::string(env) == "0") return; // output switched off (timing runs, tests)

    std::vector<double> points;
    std::vector<long> conn;
    int vpc = 0, cellType = 0;
    _vtkGeometry(&points, &conn, &vpc, &cellType);
    const std::size_t nPoints = points.size() / 3, nCells = m_cells.size();

    std::fstream out(fileName, std::ios::out | std::ios::binary | std::ios::trunc);
    if (!out) throw std::runtime_error("cannot open " + fileName);
    std::size_t offset = 0;
    auto header = [&](const std::string &type, const std::string &name, int comps, std::size_t bytes) {
        out << "        <DataArray type=\"" << type << "\" Name=\"" << name << "\" NumberOfComponents=\"" << comps
            << "\" format=\"appended\" offset=\"" << offset << "\"/>\n";
        offset += sizeof(std::uint64_t) + bytes;
    };
    out << "<?xml version=\"1.0\"?>\n"
        << "<VTKFile type=\"UnstructuredGrid\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n"
        << "  <UnstructuredGrid>\n    <Piece NumberOfPoints=\"" << nPoints << "\" NumberOfCells=\"" << nCells << "\">\n";
    out << "      <CellData>\n";
    for (const auto &f : m_vtk.getFields()) {
        if (f.location == VTKLocation::CELL) header(f.type, f.name, f.components, nCells * f.components * f.bytesPerValue);
    }
    out << "      </CellData>\n      <Points>\n";
    header("Float64", "Points", 3, points.size() * sizeof(double));
    out << "      </Points>\n      <Cells>\n";
    header("Int64", "connectivity", 1, conn.size() * sizeof(long));
    header("Int64", "offsets", 1, nCells * sizeof(long));
    header("UInt8", "types", 1, nCells);
    out << "      </Cells>\n    </Piece>\n  </UnstructuredGrid>\n  <AppendedData encoding=\"raw\">\n_";
    auto blockSize = [&](std::size_t bytes) {
        const std::uint64_t n = bytes;
        out.write(reinterpret_cast<const char *>(&n), sizeof n);
    };
    for (const auto &f : m_vtk.getFields()) {
        if (f.location != VTKLocation::CELL) continue;
        blockSize(nCells * f.components * f.bytesPerValue);
        f.streamer->flushData(out, f.name, VTKFormat::APPENDED);
    }
    blockSize(points.size() * sizeof(double));
    out.write(reinterpret_cast<const char *>(points.data()), points.size() * sizeof(double));
    blockSize(conn.size() * sizeof(long));
    out.write(reinterpret_cast<const char *>(conn.data()), conn.size() * sizeof(long));
    blockSize(nCells * sizeof(long));
    for (std::size_t c = 0; c < nCells; ++c) {
        const long end = (long) ((c + 1) * vpc);
        out.write(reinterpret_cast<const char *>(&end), sizeof end);
    }
    blockSize(nCells);
    const std::vector<unsigned char> types(nCells, (unsigned char) cellType);
    out.write(reinterpret_cast<const char *>(types.data()), nCells);
    out << "\n  </AppendedData>\n</VTKFile>\n";
}

} // namespace bitpit

#endif
