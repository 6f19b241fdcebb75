// bitpit_voloctree.hpp -- minimal stand-in for bitpit::VolOctree (see README.md): a UNIFORM octree
// (all octants on one level) of a square / cubic domain, which is everything minimmerflow's serial
// run creates (src/main.cpp:146-155).
//
// Conventions (the ones bitpit's PABLO-based VolOctree is understood to follow; the five reference
// regression strings pin the geometry, nothing in the reference pins the numbering):
//   * level = ceil(log2(max(1, length/dh))), N = 2^level cells per side, h = length/N;
//   * cells are stored in Morton order (x is the lowest interleaved bit), id == raw position;
//   * interfaces are created while visiting the cells in that order, faces in the order
//     -x,+x,-y,+y,(-z,+z); an interior face is created by (and owned by) the lower cell with the
//     normal pointing to the neighbour, a border face has the outward normal and neigh = -1;
//   * cell centroid = origin + (i+1/2) h (the unused z coordinate of a 2-D patch stays at origin_z),
//     volume h^d, size h, interface area h^(d-1).
#ifndef MMF_COMPAT_BITPIT_VOLOCTREE_HPP
#define MMF_COMPAT_BITPIT_VOLOCTREE_HPP

#include "bitpit_patchkernel.hpp"

namespace bitpit {

class VolOctree : public VolumeKernel {
public:
    VolOctree(int dimension, const std::array<double, 3> &origin, double length, double dh)
        : m_origin(origin), m_length(length)
    {
        m_dimension = dimension;
        if (dimension != 2 && dimension != 3) throw std::runtime_error("VolOctree: dimension must be 2 or 3");
        double ratio = length / dh;
        if (ratio < 1.) ratio = 1.;
        m_level = (int) std::ceil(std::log2(ratio));
        m_n = 1L << m_level;
        m_h = length / (double) m_n;
        _buildCells();
    }

    int getLevel() const { return m_level; }
    long getCellsPerDirection() const { return m_n; }
    double getLength() const { return m_length; }
    const std::array<double, 3> &getOrigin() const { return m_origin; }

    void initializeAdjacencies() override {}
    void initializeInterfaces() override { m_wantInterfaces = true; }
    void update() override
    {
        if (m_wantInterfaces && m_interfaces.size() == 0) _buildInterfaces();
    }

    // lattice coordinates of a cell (not part of bitpit: used by adapters that hand the structure of
    // a uniform patch to an accelerator)
    std::array<int, 3> cellLattice(long id) const
    {
        std::array<int, 3> ijk = { { 0, 0, 0 } };
        for (int b = 0; b < 21; ++b) {
            for (int d = 0; d < m_dimension; ++d) ijk[d] |= (int) (((unsigned long) id >> (m_dimension * b + d)) & 1ul) << b;
        }
        return ijk;
    }

    double evalCellVolume(long) const override { return (m_dimension == 3) ? m_h * m_h * m_h : m_h * m_h; }
    double evalCellSize(long) const override { return m_h; }
    std::array<double, 3> evalCellCentroid(long id) const override
    {
        const std::array<int, 3> ijk = cellLattice(id);
        std::array<double, 3> c = m_origin;
        for (int d = 0; d < m_dimension; ++d) c[d] = m_origin[d] + ((double) ijk[d] + 0.5) * m_h;
        return c;
    }
    double evalInterfaceArea(long) const override { return (m_dimension == 3) ? m_h * m_h : m_h; }
    std::array<double, 3> evalInterfaceCentroid(long id) const override
    {
        const Interface &f = m_interfaces.at(id);
        std::array<double, 3> c = evalCellCentroid(f.getOwner());
        const int d = f.getOwnerFace() / 2, sign = (f.getOwnerFace() % 2) ? +1 : -1;
        c[d] = c[d] + 0.5 * sign * m_h;
        return c;
    }
    std::array<double, 3> evalInterfaceNormal(long id) const override
    {
        const Interface &f = m_interfaces.at(id);
        std::array<double, 3> n = { { 0., 0., 0. } };
        n[f.getOwnerFace() / 2] = (f.getOwnerFace() % 2) ? 1. : -1.;
        return n;
    }

protected:
    void _vtkGeometry(std::vector<double> *points, std::vector<long> *conn, int *verticesPerCell, int *vtkCellType) const override
    {
        const long np = m_n + 1, nz = (m_dimension == 3) ? np : 1;
        points->resize((std::size_t) (np * np * nz) * 3);
        for (long k = 0; k < nz; ++k)
            for (long j = 0; j < np; ++j)
                for (long i = 0; i < np; ++i) {
                    double *p = points->data() + 3 * ((k * np + j) * np + i);
                    p[0] = m_origin[0] + i * m_h;
                    p[1] = m_origin[1] + j * m_h;
                    p[2] = (m_dimension == 3) ? m_origin[2] + k * m_h : m_origin[2];
                }
        *verticesPerCell = (m_dimension == 3) ? 8 : 4;
        *vtkCellType = (m_dimension == 3) ? 11 /* VTK_VOXEL */ : 8 /* VTK_PIXEL */;
        conn->clear();
        conn->reserve(m_cells.size() * (std::size_t) *verticesPerCell);
        for (std::size_t c = 0; c < m_cells.size(); ++c) {
            const std::array<int, 3> q = cellLattice((long) c);
            for (int v = 0; v < *verticesPerCell; ++v) {
                const long i = q[0] + (v & 1), j = q[1] + ((v >> 1) & 1), k = q[2] + ((v >> 2) & 1);
                conn->push_back((k * np + j) * np + i);
            }
        }
    }

private:
    std::array<double, 3> m_origin;
    double m_length, m_h = 0.;
    int m_level = 0;
    long m_n = 1;
    bool m_wantInterfaces = false;

    long _morton(const std::array<int, 3> &ijk) const
    {
        unsigned long m = 0;
        for (int b = 0; b < 21; ++b) {
            for (int d = 0; d < m_dimension; ++d) m |= (unsigned long) ((ijk[d] >> b) & 1) << (m_dimension * b + d);
        }
        return (long) m;
    }

    void _buildCells()
    {
        long nCells = 1;
        for (int d = 0; d < m_dimension; ++d) nCells *= m_n;
        m_cells.reserve((std::size_t) nCells);
        for (long c = 0; c < nCells; ++c) m_cells.emplaceBack(c, true);
    }

    void _buildInterfaces()
    {
        const long nCells = (long) m_cells.size();
        long nFaces = m_dimension * (m_n + 1);
        for (int d = 1; d < m_dimension; ++d) nFaces *= m_n;
        m_interfaces.reserve((std::size_t) nFaces);
        long id = 0;
        for (long c = 0; c < nCells; ++c) {
            const std::array<int, 3> ijk = cellLattice(c);
            for (int face = 0; face < 2 * m_dimension; ++face) {
                const int d = face / 2, sign = (face % 2) ? +1 : -1;
                const long coord = (long) ijk[d] + sign;
                long neigh = -1;
                if (coord >= 0 && coord < m_n) {
                    std::array<int, 3> q = ijk;
                    q[d] = (int) coord;
                    neigh = _morton(q);
                    if (neigh < c) continue; // already created by the lower cell
                }
                m_interfaces.emplaceBack(id, c, face, neigh, neigh < 0 ? -1 : (face ^ 1));
                ++id;
            }
        }
    }
};

} // namespace bitpit

#endif
