"""The fused stage kernels' SOURCE, run on the CPU SIMT emulator of tools/emu (fibers per lane, PTX
mbarrier phase/parity semantics, random warp delays) and compared with the oracle bit for bit.

This is a check of the kernels' synchronisation protocol and index arithmetic -- a slot overwritten early,
a barrier overtaken by two phases (a hang on the GPU), a ragged-tile slip -- that needs no GPU, so a new
kernel form is debugged here before GPU minutes are spent on it.  It is NOT a product path (nothing in
minimmerflow_b200/ can reach the emulator) and it proves nothing about SASS, registers or speed: the
`-m gpu` tests through the C-ABI remain the parity proof."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "emu"))

import run_emu  # noqa: E402
from common import (bits_equal, case_mesh, lexicographic_box_mesh, primitives, reference_cases, reference_fields,  # noqa: E402
                    with_bodies)


@pytest.fixture(scope="module")
def emu():
    return run_emu.load()


@pytest.mark.parametrize("form", ["r", "t", "h"])
@pytest.mark.parametrize("chaos", [0, 300])
def test_stage_kernel_source_matches_oracle_on_the_emulator(emu, oracle, form, chaos):
    # Morton cube (the reference's numbering), z chunks of 6 planes: general and steady-state bodies
    m = oracle.problem_mesh("vortex_xy", 3, 16)
    assert run_emu.check_case(emu, oracle, "vortex 16^3", dict(m), 0, form, 8, 6, 2, chaos, 1)
    # ragged lexicographic box with reflecting borders (ghost pass), two-plane chunks, 12-warp CTAs
    m = lexicographic_box_mesh(33, 8, 5, 0.5, 1)
    assert run_emu.check_case(emu, oracle, "box 33x8x5", dict(m), 1, form, 12, 2, 1, chaos, 2)
    # one-plane chunks on free-flow borders (clamped loads)
    m = lexicographic_box_mesh(31, 7, 2, 0.5, 0)
    assert run_emu.check_case(emu, oracle, "box 31x7x2", dict(m), 1, form, 8, 1, 1, chaos, 3)


def test_emulated_mbarrier_keeps_ptx_phase_semantics(emu):
    """The hazard the emulator exists to catch: a parity wait overtaken by two phase completions never
    returns.  Checked on the barrier word itself (no kernel): after two completions the parity a waiter of
    the first phase polls for is current again."""
    import ctypes as C
    emu.emu_mbar_selftest.restype = C.c_int
    assert emu.emu_mbar_selftest() == 0


def test_axis_order_forms_agree_on_the_emulator(emu, oracle):
    """NUM_AXIS accumulation (x_lo - x_hi + y_lo - y_hi + z_lo - z_hi) is not the reference's order, so
    there is no oracle for its bits; the form fed by bulk tensor loads must give exactly what the rotate form gives, and
    both stay within rounding of the oracle."""
    m = oracle.problem_mesh("radsod", 3, 16)
    U0 = oracle.init_state(m)
    ref, ref_eig = oracle.compute_rhs(m, U0)
    out = {}
    for form in ("r", "t", "h"):
        box = run_emu.Box(emu, oracle, dict(m), 2)
        U, R = box.new_array(), box.new_array()
        box.scatter(U, U0)
        box.fill_ghosts(U)
        eig, _ = box.stage(form, 0, 8, 5, U, U, R, 0.0, 200, 5)
        assert eig == ref_eig
        out[form] = box.gather(R)
    assert np.array_equal(out["r"], out["t"]) and np.array_equal(out["r"], out["h"])
    assert np.abs(out["r"] - ref).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("form", ["r"])
def test_compact_x_ghost_columns_on_the_emulator(emu, oracle, form):
    """Multi-GPU layout of an x partition side (XGhost): the halo lanes i = -1 / i = nx take their column
    from compact arrays [field][k+1][j+1] instead of the padded array.  Here the columns hold the
    reflecting-wall virtual states, and the padded array's own x ghost columns are poisoned with NaNs."""
    import ctypes as C
    D = C.POINTER(C.c_double)
    m = lexicographic_box_mesh(33, 12, 5, 0.5, 1)
    m["problem"] = "radsod"
    nc = m["volume"].shape[0]
    rng = np.random.default_rng(3)
    rho = rng.uniform(0.5, 1.5, nc); vel = rng.uniform(-0.4, 0.4, (nc, 3)); p = rng.uniform(0.6, 1.4, nc)
    U0 = np.column_stack([rho, rho * vel[:, 0], rho * vel[:, 1], rho * vel[:, 2], p / 0.4 + 0.5 * rho * (vel ** 2).sum(1)])
    ref, ref_eig = oracle.compute_rhs(m, U0)
    box = run_emu.Box(emu, oracle, dict(m), 1)
    U, R = box.new_array(), box.new_array()
    box.scatter(U, U0)
    box.fill_ghosts(U)
    lo, hi, fs, pitch = box.compact_x_ghosts(U)
    emu.emu_set_xghost(lo.ctypes.data_as(D), hi.ctypes.data_as(D), fs, pitch)
    try:
        eig, _ = box.stage(form, 0, 12, 3, U, U, R, 0.0, 100, 9)
    finally:
        emu.emu_set_xghost(None, None, 0, 0)
    assert eig == ref_eig
    assert np.array_equal(box.gather(R), ref)


@pytest.mark.parametrize("form", ["b"])
@pytest.mark.parametrize("chaos", [0, 300])
def test_body_stage_kernel_source_matches_oracle_on_the_emulator(emu, oracle, chaos, form):
    """Kernel form 'b' (uniform_stage_t.cuh with BODY: the TMA-fed kernel plus one flag byte per cell), wall cells
    recomputed by wall_cell_update (uniform_body_cells.cuh) around a stage kernel without a slow path: a uniform box
    with bodies -- unsolved cells, wall interfaces
    evaluated against the fluid cell's mirror image, solid | solid interfaces skipped -- and the eigenvalue
    pass that chooses dt there (eig_body_cell), bit for bit against the oracle."""
    # the reference's set-up: Morton cube, reflecting borders, a box body inside
    m = oracle.problem_mesh("radsod", 3, 16, boxes=[[2.1, 3.2, 1.3, 4.9, 5.4, 3.6]])
    assert (m["solved"] == 0).sum() > 50 and (m["bc"] == 2).sum() > 100
    assert run_emu.check_case(emu, oracle, "radsod 16^3 + body", dict(m), 0, form, 12, 6, 2, chaos, 1)
    # ragged lexicographic box, free-flow borders (clamped loads), bodies touching the border, a one-cell
    # body and a one-cell gap between two bodies; two-plane z chunks
    m = with_bodies(lexicographic_box_mesh(33, 9, 5, 0.5, 0), [[-1, -1, -1, 1.2, 1.2, 1.2], [7.1, 2.1, 1.1, 7.4, 2.4, 1.4],
                                                               [10.1, 0.0, 0.0, 11.9, 9.0, 1.4], [12.6, 0.0, 0.0, 16.4, 2.4, 9.0]])
    m["problem"] = "vortex_xy"
    assert (m["solved"] == 0).sum() > 30
    assert run_emu.check_case(emu, oracle, "box 33x9x5 + bodies", dict(m), 1, form, 12, 2, 2, chaos, 2)
    # reflecting borders with a body on them
    m = with_bodies(lexicographic_box_mesh(7, 23, 4, 0.5, 1), [[-1, 4.1, -1, 1.4, 6.4, 9.0], [2.1, 10.1, 0.6, 2.9, 11.4, 1.4]])
    m["problem"] = "radsod"
    assert run_emu.check_case(emu, oracle, "box 7x23x4 + bodies", dict(m), 1, form, 12, 3, 2, chaos, 3)


BODY_CASES_3D = [c for c in reference_cases() if c["dim"] == 3 and c.get("bodies")]


@pytest.mark.parametrize("form", ["b"])
@pytest.mark.parametrize("case", BODY_CASES_3D, ids=lambda c: c["name"])
def test_body_kernel_source_reproduces_the_reference_fields_on_the_emulator(emu, oracle, case, form):
    """The whole run of a reference case with bodies -- dt from the eigenvalue pass (eig_body_cell), three fused
    stages of a body kernel form per step, `while (t < tMax)` with the clamp -- driven from the emulator alone and
    compared with the final fields the UNMODIFIED reference wrote (tests/golden/reference_fields.npz)."""
    ref = reference_fields()
    n = case["name"]
    m = case_mesh(oracle, case)
    box = run_emu.Box(emu, oracle, dict(m), 0)
    assert box.solid is not None
    U, Wa, Wb = box.new_array(), box.new_array(), box.new_array()
    box.scatter(U, oracle.init_state(m))
    box.fill_ghosts(U)
    t, steps, h = 0.0, 0, float(m["size"].min())
    while t < case["t_end"]:
        eig = box.eig_body(U)
        dt = oracle.choose_dt(case["cfl"], h, eig, t, case["t_end"])
        e1, _ = box.stage(form, 1, 12, 6, U, U, Wa, dt)
        assert e1 == eig                      # the check mmf_step makes every step
        box.fill_ghosts(Wa)
        box.stage(form, 2, 12, 6, Wa, U, Wb, dt)
        box.fill_ghosts(Wb)
        box.stage(form, 3, 12, 6, Wb, U, U, dt)
        box.fill_ghosts(U)
        t += dt
        steps += 1
    Uf = box.gather(U)
    assert steps == int(ref[n + "/steps"]) and t == case["t_end"]
    P = primitives(oracle, Uf)
    assert bits_equal(ref[n + "/density"], Uf[:, 0])
    assert bits_equal(ref[n + "/velocity"], P[:, 1:4])
    assert bits_equal(ref[n + "/pressure"], P[:, 0])
    assert bits_equal(ref[n + "/temperature"], P[:, 4])


PLAIN_CASES_3D = [c for c in reference_cases() if c["dim"] == 3 and not c.get("bodies")]


@pytest.mark.parametrize("form,nw", [("r", 12), ("t", 16), ("h", 12), ("h", 16)])
@pytest.mark.parametrize("case", PLAIN_CASES_3D, ids=lambda c: c["name"])
def test_stage_kernel_source_reproduces_the_reference_fields_on_the_emulator(emu, oracle, case, form, nw):
    """Every stage-kernel form through
    the whole run of the plain 3-D reference cases (dt from the face maximum of an RHS-only launch), against
    the final fields the UNMODIFIED reference wrote."""
    ref = reference_fields()
    n = case["name"]
    m = case_mesh(oracle, case)
    box = run_emu.Box(emu, oracle, dict(m), 0)
    U, Wa, Wb, R = box.new_array(), box.new_array(), box.new_array(), box.new_array()
    box.scatter(U, oracle.init_state(m))
    box.fill_ghosts(U)
    t, steps, h = 0.0, 0, float(m["size"].min())
    while t < case["t_end"]:
        eig, _ = box.stage(form, 0, nw, 7, U, U, R, 0.0)
        dt = oracle.choose_dt(case["cfl"], h, eig, t, case["t_end"])
        e1, _ = box.stage(form, 1, nw, 7, U, U, Wa, dt)
        assert e1 == eig
        box.fill_ghosts(Wa)
        box.stage(form, 2, nw, 7, Wa, U, Wb, dt)
        box.fill_ghosts(Wb)
        box.stage(form, 3, nw, 7, Wb, U, U, dt)
        box.fill_ghosts(U)
        t += dt
        steps += 1
    Uf = box.gather(U)
    assert steps == int(ref[n + "/steps"]) and t == case["t_end"]
    P = primitives(oracle, Uf)
    assert bits_equal(ref[n + "/density"], Uf[:, 0])
    assert bits_equal(ref[n + "/velocity"], P[:, 1:4])
    assert bits_equal(ref[n + "/pressure"], P[:, 0])
    assert bits_equal(ref[n + "/temperature"], P[:, 4])


def test_host_side_body_flags_and_wall_list(emu, oracle):
    """body_flags (uniform_device.cuh), the host code uniform_try_create runs for a box with bodies: flag array
    with the replicated ghost shell and -- kernel form 'b' -- flag 2 plus the ascending list of fluid cells that
    touch a wall, against the independent construction of the test harness (run_emu.Box)."""
    import ctypes as C
    I = C.POINTER(C.c_int)
    meshes = [oracle.problem_mesh("radsod", 3, 16, boxes=[[2.1, 3.2, 1.3, 4.9, 5.4, 3.6], [5.6, 0.0, 5.1, 8.0, 1.4, 5.9]]),
              with_bodies(lexicographic_box_mesh(33, 9, 5, 0.5, 0), [[-1, -1, -1, 1.2, 1.2, 1.2], [7.1, 2.1, 1.1, 7.4, 2.4, 1.4],
                                                                     [10.1, 0.0, 0.0, 11.9, 9.0, 1.4], [12.6, 0.0, 0.0, 16.4, 2.4, 9.0]])]
    for m in meshes:
        m.setdefault("problem", "radsod")
        box = run_emu.Box(emu, oracle, dict(m), 0)
        nc = m["volume"].shape[0]
        ijk = np.ascontiguousarray(m["cell_ijk"], np.int32)
        solved = np.ascontiguousarray(m["solved"], np.uint8)
        # inner cells and the ghosts behind a face; edge and corner ghosts touch no interface
        px, py, pz = (int(v) for v in box.pad)
        nx, ny, nz = (int(v) for v in box.dims)
        kk, jj, ii = np.meshgrid(np.arange(pz) - 1, np.arange(py) - 1, np.arange(px) - run_emu.XOFF, indexing="ij")
        outside = ((ii < 0) | (ii >= nx)).astype(int) + ((jj < 0) | (jj >= ny)) + ((kk < 0) | (kk >= nz))
        used = np.zeros(box.fs, bool)
        used[:px * py * pz] = ((outside <= 1) & (ii >= -1) & (ii <= nx) & (jj <= ny) & (kk <= nz)).reshape(-1)
        for mark in (0, 1):
            flag = np.full(box.fs, 255, np.uint8)
            walls = np.zeros(nc, np.int32)
            n = emu.emu_body_flags(box.dims.ctypes.data_as(I), nc, ijk.ctypes.data_as(I), solved.ctypes.data, mark,
                                   flag.ctypes.data, walls.ctypes.data_as(I))
            want = box.flag_c if mark else box.solid
            assert np.array_equal(flag[used], want[used]) and flag.max() <= 2
            assert n == (len(box.walls) if mark else 0)
            if mark:
                assert n > 20 and np.array_equal(walls[:n], box.walls)


def test_dispatcher_decision_uniform_or_generic(emu, oracle):
    """analyze_uniform_box (uniform_eligibility.h) -- what mmf_create decides between the fused uniform path and
    the generic one -- on the CPU: numbering conventions, one BC per side, bodies only when allowed and only
    with BC_WALL on exactly the fluid | solid interfaces, everything irregular to the generic path."""
    import ctypes as C
    from minimmerflow_b200.solver import mesh_desc
    from common import two_level_mesh

    def analyze(m, allow_bodies=0, flags=0):
        d, keep = mesh_desc(m, flags=flags)
        out = np.zeros(9, np.int32)
        ok = emu.emu_analyze_box(C.addressof(d), allow_bodies, out.ctypes.data_as(C.POINTER(C.c_int)))
        return ok, out

    ok, out = analyze(oracle.problem_mesh("vortex_xy", 3, 8))                    # the reference's numbering
    assert ok and list(out[:3]) == [0, 1, 0] and list(out[3:]) == [0] * 6         # Morton, exact, free flow
    ok, out = analyze(oracle.problem_mesh("radsod", 3, 8))
    assert ok and list(out[3:]) == [1] * 6                                       # reflecting
    ok, out = analyze(lexicographic_box_mesh(7, 5, 3, 0.5, 1))
    assert ok and list(out[:3]) == [1, 1, 0]                                     # lexicographic
    assert not analyze(oracle.problem_mesh("vortex_xy", 2, 8))[0]                # 2-D
    assert not analyze(oracle.problem_mesh("vortex_xy", 3, 8), flags=1)[0]        # MMF_FLAG_FORCE_GENERIC
    ok, out = analyze(oracle.problem_mesh("vortex_xy", 3, 8), flags=2)             # MMF_FLAG_ORDER_AXIS
    assert ok and list(out[:2]) == [2, 0]
    assert not analyze(two_level_mesh(3, 2, lambda i, j, k: i == 0))[0]          # hanging faces
    # bodies
    mb = oracle.problem_mesh("radsod", 3, 8, boxes=[[2.1, 2.1, 2.1, 5.9, 4.9, 3.9]])
    assert not analyze(mb)[0]                                                    # not asked for: generic
    ok, out = analyze(mb, allow_bodies=1)
    assert ok and out[2] == 1
    bad = dict(mb); bad["bc"] = mb["bc"].copy()
    wall = np.nonzero(bad["bc"] == 2)[0][0]
    bad["bc"][wall] = -1                                                         # a wall interface without BC_WALL
    assert not analyze(bad, allow_bodies=1)[0]
    bad["bc"][wall] = 2
    inner = np.nonzero((mb["bc"] == -1) & (mb["neigh"] >= 0))[0][0]
    bad["bc"][inner] = 2                                                         # BC_WALL between two fluid cells
    assert not analyze(bad, allow_bodies=1)[0]
    # an irregularity in an otherwise perfect box
    m = oracle.problem_mesh("vortex_xy", 3, 8)
    bad = dict(m); bad["volume"] = m["volume"].copy(); bad["volume"][5] *= 2
    assert not analyze(bad)[0]
    bad = dict(m); bad["bc"] = m["bc"].copy(); bad["bc"][np.nonzero(m["neigh"] < 0)[0][0]] = 1   # two BCs on one side
    assert not analyze(bad)[0]
    bad = dict(m); bad["interface_order"] = np.arange(m["owner"].shape[0])[::-1].copy()           # unknown order
    assert not analyze(bad)[0]
