import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def golden_cases():
    with open(os.path.join(HERE, "golden", "final_errors.json")) as f:
        return json.load(f)["cases"]


def bits_equal(a, b):
    """Bitwise equality of two float64 arrays, treating +0 and -0 as equal (the sign of an exact
    zero is the one thing axis-specialised arithmetic does not preserve)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def max_rel_diff(a, b):
    """max |a-b| per field normalised by the field's max-abs (SURVEY 8d parity metric)."""
    a = np.asarray(a, dtype=np.float64).reshape(-1, 5)
    b = np.asarray(b, dtype=np.float64).reshape(-1, 5)
    scale = np.maximum(np.abs(b).max(axis=0), 1e-300)
    return float((np.abs(a - b).max(axis=0) / scale).max())


def lexicographic_box_mesh(nx, ny, nz, h, bc_code, origin=(0.0, 0.0, 0.0)):
    """A full box mesh numbered lexicographically (x fastest), interfaces created while visiting the
    cells in that order and the faces in order -x,+x,-y,+y,-z,+z (test-side generator for the
    non-Morton numbering; pure numpy/python, small sizes only)."""
    nc = nx * ny * nz
    idx = lambda i, j, k: (k * ny + j) * nx + i
    owner, neigh, normal, bc = [], [], [], []
    dims = (nx, ny, nz)
    ijk = np.empty((nc, 3), np.int32)
    cc = np.empty((nc, 3))
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                c = idx(i, j, k)
                ijk[c] = (i, j, k)
                cc[c] = (origin[0] + (i + 0.5) * h, origin[1] + (j + 0.5) * h, origin[2] + (k + 0.5) * h)
                for face in range(6):
                    d, sgn = face // 2, (1 if face % 2 else -1)
                    q = [i, j, k]
                    q[d] += sgn
                    n = [0.0, 0.0, 0.0]
                    n[d] = float(sgn)
                    if q[d] < 0 or q[d] >= dims[d]:
                        owner.append(c); neigh.append(-1); normal.append(n); bc.append(bc_code)
                    else:
                        nb = idx(*q)
                        if nb < c:
                            continue
                        owner.append(c); neigh.append(nb); normal.append(n); bc.append(-1)
    nf = len(owner)
    return dict(dim=3, owner=np.array(owner, np.int64), neigh=np.array(neigh, np.int64),
                bc=np.array(bc, np.int32), area=np.full(nf, h * h), normal=np.array(normal),
                icentroid=np.zeros((nf, 3)), volume=np.full(nc, h * h * h), size=np.full(nc, h),
                ccentroid=cc, cell_ijk=ijk, box_dims=(nx, ny, nz), solved=np.ones(nc, np.uint8),
                internal=np.ones(nc, np.uint8), fluid=np.ones(nc, np.uint8), h=h)


def with_bodies(m, boxes):
    """Flags and BC table of a mesh with body boxes (xMin,yMin,zMin,xMax,yMax,zMax), set up the way
    src/main.cpp:221-237, 251-277 does: cells whose centroid lies in a closed box are not solved, interfaces
    between a solved and an unsolved cell get BC_WALL (2).  Border interfaces keep their code."""
    m = dict(m)
    cc = m["ccentroid"]
    solid = np.zeros(cc.shape[0], bool)
    for b in boxes:
        solid |= np.all((cc >= np.array(b[:3])) & (cc <= np.array(b[3:])), axis=1)
    fluid = (~solid).astype(np.uint8)
    bc = m["bc"].copy()
    inner = m["neigh"] >= 0
    o, n = m["owner"][inner], m["neigh"][inner]
    bc[inner] = np.where(fluid[o] != fluid[n], 2, -1)
    m.update(fluid=fluid, solved=fluid.copy(), bc=bc)
    return m


def reference_cases():
    with open(os.path.join(HERE, "golden", "reference_cases.json")) as f:
        return json.load(f)["cases"]


def reference_fields():
    """Final fields of reference_cases() as produced by the unmodified reference
    (tests/golden/make_reference_fields.py)."""
    return np.load(os.path.join(HERE, "golden", "reference_fields.npz"))


def case_mesh(oracle, case):
    """Mesh + flags + BC table of a reference case (custom domain and bodies included), set up the
    way src/main.cpp does."""
    import ctypes as C
    import oracle_lib as O
    dim, origin, length = oracle.domain(case["problem"], case["dim"])
    if case.get("origin") is not None:
        origin = np.array(case["origin"], dtype=np.float64)
    if case.get("length") is not None:
        length = float(case["length"])
    m = oracle.uniform_mesh(dim, origin, length, case["n_cells"])
    nc, nf = m["volume"].shape[0], m["owner"].shape[0]
    boxes = np.ascontiguousarray(case.get("bodies") or np.zeros((0, 6)), dtype=np.float64).reshape(-1, 6)
    fluid = np.empty(nc, np.uint8)
    oracle.lib.orc_fluid_flags(nc, O._p(m["ccentroid"], O._D), boxes.shape[0], O._p(boxes, O._D), O._p(fluid, O._U8))
    bc = np.empty(nf, np.int32)
    oracle.lib.orc_interface_bcs(O.PROBLEMS[case["problem"]], nf, O._p(m["owner"], O._I64), O._p(m["neigh"], O._I64),
                                 O._p(m["icentroid"], O._D), O._p(fluid, O._U8), O._p(bc, O._I32))
    m.update(problem=case["problem"], fluid=fluid, solved=fluid.copy(), internal=np.ones(nc, np.uint8), bc=bc)
    return m


def primitives(oracle, U):
    """utils::conservative2primitive (src/utils.cpp:48-63) applied cell by cell -> {p,u,v,w,T}."""
    import oracle_lib as O
    U = np.ascontiguousarray(U, dtype=np.float64)
    P = np.empty_like(U)
    for c in range(U.shape[0]):
        oracle.lib.orc_conservative2primitive(U[c].ctypes.data_as(O._D), P[c].ctypes.data_as(O._D))
    return P


def dirichlet_info(case):
    """problem::getBorderBCInfo (src/problem.cpp:450-477): only the forward-facing step sets data."""
    return [1.0, 3.0, 0.0, 0.0, 1.0 / 1.4] if case["problem"] == "ffstep" else None


def two_level_mesh(dim, n0, refine, H=1.0, bc_code=1, origin=(0.0, 0.0, 0.0)):
    """A 2:1-balanced two-level octree/quadtree mesh: n0^dim coarse cells of size H; those for which
    refine(i,j,k) is true are split into 2^dim children (hanging faces towards coarse neighbours).
    Cells are numbered along the Morton curve of their lower corner on the fine lattice; one interface
    per pair of face neighbours, sized by the finer side, OWNER = THE FINER CELL (else the low-side
    cell), normal from owner to neighbour (so interior normals of both signs occur); interfaces are
    created cell by cell in the order -x,+x,-y,+y,(-z,+z). Test-side generator, small sizes only."""
    nz0 = n0 if dim == 3 else 1
    nf = 2 * n0                                             # fine lattice cells per side
    nfz = 2 * nz0 if dim == 3 else 1
    h = 0.5 * H

    def morton(i, j, k):
        key = 0
        for b in range(16):
            key |= ((i >> b) & 1) << (3 * b) | ((j >> b) & 1) << (3 * b + 1) | ((k >> b) & 1) << (3 * b + 2)
        return key

    cells = []                                              # (morton key, fine i, j, k, level)
    for k0 in range(nz0):
        for j0 in range(n0):
            for i0 in range(n0):
                if refine(i0, j0, k0):
                    for dk in range(2 if dim == 3 else 1):
                        for dj in range(2):
                            for di in range(2):
                                i, j, k = 2 * i0 + di, 2 * j0 + dj, (2 * k0 + dk) if dim == 3 else 0
                                cells.append((morton(i, j, k), i, j, k, 1))
                else:
                    i, j, k = 2 * i0, 2 * j0, (2 * k0) if dim == 3 else 0
                    cells.append((morton(i, j, k), i, j, k, 0))
    cells.sort()
    nc = len(cells)
    lattice = -np.ones((nfz, nf, nf), np.int64)             # fine lattice -> cell id
    for c, (_, i, j, k, lev) in enumerate(cells):
        w = 1 if lev else 2
        wz = w if dim == 3 else 1
        lattice[k:k + wz, j:j + w, i:i + w] = c
    assert (lattice >= 0).all()

    size = np.array([h if lev else H for (_, _, _, _, lev) in cells])
    cc = np.empty((nc, 3))
    for c, (_, i, j, k, lev) in enumerate(cells):
        w = size[c]
        cc[c] = (origin[0] + i * h + 0.5 * w, origin[1] + j * h + 0.5 * w,
                 (origin[2] + k * h + 0.5 * w) if dim == 3 else origin[2])
    owner, neigh, normal, area, bc, icent = [], [], [], [], [], []
    seen = set()
    ext = (nf, nf, nfz)
    for c, (_, i, j, k, lev) in enumerate(cells):
        w = 1 if lev else 2
        lo = (i, j, k)
        for face in range(2 * dim):
            d, sgn = face // 2, (1 if face % 2 else -1)
            q = lo[d] + (w if sgn > 0 else -1)              # fine-lattice coordinate just across the face
            n = [0.0, 0.0, 0.0]
            n[d] = float(sgn)
            t_axes = [a for a in range(dim) if a != d]
            if q < 0 or q >= ext[d]:
                fc = list(cc[c]); fc[d] += sgn * 0.5 * size[c]
                owner.append(c); neigh.append(-1); normal.append(n); area.append(size[c] ** (dim - 1))
                bc.append(bc_code); icent.append(fc)
                continue
            # distinct neighbours across this face (1, or 2^(dim-1) finer ones)
            nbs = []
            for s1 in range(w):
                for s2 in range(w if dim == 3 else 1):
                    p = list(lo)
                    p[d] = q
                    p[t_axes[0]] += s1
                    if dim == 3:
                        p[t_axes[1]] += s2
                    nb = int(lattice[p[2], p[1], p[0]])
                    if nb not in nbs:
                        nbs.append(nb)
            for nb in nbs:
                pair = (min(c, nb), max(c, nb), d)
                if pair in seen:
                    continue
                seen.add(pair)
                fs = min(size[c], size[nb])
                finer = c if size[c] < size[nb] else (nb if size[nb] < size[c] else (c if sgn > 0 else nb))
                other = nb if finer == c else c
                nn = [0.0, 0.0, 0.0]
                nn[d] = float(sgn) if finer == c else float(-sgn)
                fc = list(cc[finer]); fc[d] += nn[d] * 0.5 * fs
                owner.append(finer); neigh.append(other); normal.append(nn); area.append(fs ** (dim - 1))
                bc.append(-1); icent.append(fc)
    nfaces = len(owner)
    return dict(dim=dim, owner=np.array(owner, np.int64), neigh=np.array(neigh, np.int64),
                bc=np.array(bc, np.int32), area=np.array(area), normal=np.array(normal),
                icentroid=np.array(icent), volume=size ** dim, size=size, ccentroid=cc,
                solved=np.ones(nc, np.uint8), internal=np.ones(nc, np.uint8), fluid=np.ones(nc, np.uint8),
                h=float(size.min()), n_faces=nfaces)
