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
