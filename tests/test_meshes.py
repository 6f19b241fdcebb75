"""minimmerflow_b200.meshes (vectorised host descriptions of uniform boxes for the C-ABI, used by the development
tools at benchmark sizes) against the independent loop construction of tests/common.py and the checker's vortex."""
import ctypes as C
import os
import sys

import numpy as np

from common import lexicographic_box_mesh, with_bodies as with_bodies_ref
from minimmerflow_b200.meshes import box_mesh, vortex_state, with_bodies

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_box_description_matches_the_loop_construction():
    for dims, h, bc, origin in (((7, 5, 3), 0.5, 1, (-1.0, 0.0, 2.0)), ((1, 4, 2), 0.25, 0, (0.0, 0.0, 0.0)), ((3, 1, 1), 1.0, 3, (0.0, 0.0, 0.0))):
        a, b = box_mesh(*dims, h, bc, origin=origin), lexicographic_box_mesh(*dims, h, bc, origin=origin)
        assert set(a) == set(b)
        for key in b:
            assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), key
    boxes = [[0.1, 0.1, 2.1, 1.4, 1.4, 3.2]]
    a = with_bodies(box_mesh(7, 5, 3, 0.5, 1, origin=(-1.0, 0.0, 2.0)), boxes)
    b = with_bodies_ref(lexicographic_box_mesh(7, 5, 3, 0.5, 1, origin=(-1.0, 0.0, 2.0)), boxes)
    assert (a["solved"] == 0).any() and (a["bc"] == 2).any()
    for key in b:
        assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), key


def test_vortex_state_is_the_reference_initial_condition(oracle):
    mo = oracle.problem_mesh("vortex_xy", 3, 16)
    Uo = oracle.init_state(mo)
    ml = box_mesh(16, 16, 16, 10.0 / 16, 0, origin=(-5.0, -5.0, -5.0))
    Ul = vortex_state(ml)
    key = lambda m: (np.asarray(m["cell_ijk"])[:, 2] * 16 + np.asarray(m["cell_ijk"])[:, 1]) * 16 + np.asarray(m["cell_ijk"])[:, 0]  # noqa: E731
    A, B = Uo[np.argsort(key(mo))], Ul[np.argsort(key(ml))]
    assert np.abs(A - B).max() < 1e-14          # numpy's exp / pow against glibc's


def test_benchmark_boxes_take_the_fused_path():
    """The dispatcher's decision (uniform_eligibility.h, through the tool library) on what tools/generic_bench.py
    builds: lexicographic numbering, free flow; with bodies only when allowed."""
    sys.path.insert(0, os.path.join(ROOT, "tools", "emu"))
    import run_emu
    from minimmerflow_b200.solver import mesh_desc
    emu = run_emu.load()
    origin, length, n = (-5.0, -5.0, -5.0), 10.0, 12
    m = box_mesh(n, n, n, length / n, 0, origin=origin)
    lo = lambda f: [origin[e] + f[e] * length for e in range(3)]  # noqa: E731
    mb = with_bodies(m, [lo((0.30, 0.35, 0.25)) + lo((0.45, 0.60, 0.55)), lo((0.70, 0.10, 0.60)) + lo((0.85, 0.30, 0.95))])
    out = np.zeros(9, np.int32)
    po = out.ctypes.data_as(C.POINTER(C.c_int))
    d, keep = mesh_desc(m)
    assert emu.emu_analyze_box(C.addressof(d), 0, po) == 1 and list(out[:3]) == [1, 1, 0] and list(out[3:]) == [0] * 6
    d, keep = mesh_desc(mb)
    assert emu.emu_analyze_box(C.addressof(d), 0, po) == 0
    assert emu.emu_analyze_box(C.addressof(d), 1, po) == 1 and out[2] == 1
