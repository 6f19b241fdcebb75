"""GPU parity, uniform fused path (the benchmark path), through the C-ABI.

The fused kernels compute every interface once, share per-cell primitives, divide with a shared
reciprocal and accumulate in the host's interface-id order: the results must still be BITWISE equal
to the reference-shaped oracle, not just within north_star's 1e-12 relative."""
import os

import numpy as np
import pytest

import oracle_lib
from common import bits_equal, golden_cases, lexicographic_box_mesh, max_rel_diff, with_bodies

pytestmark = pytest.mark.gpu

GOLDEN_3D = [c for c in golden_cases() if c["dim"] == 3]


def test_shared_reciprocal_division_is_ieee(mmf):
    assert mmf.selftest_division(n_samples=1 << 30, seed=2024) == 0


def _bc_sides(m):
    border = m["neigh"] < 0
    code = int(m["bc"][border][0])
    assert np.all(m["bc"][border] == code)
    return [code] * 6


@pytest.mark.parametrize("problem,n", [("vortex_xy", 32), ("vortex_yz", 16), ("radsod", 32), ("sod3d_z", 8)])
def test_uniform_path_selected_and_rhs_bit_exact(mmf, oracle, problem, n):
    m = oracle.problem_mesh(problem, 3, n)
    U = oracle.init_state(m)
    ref, ref_eig = oracle.compute_rhs(m, U)
    with mmf.EulerSolver.from_mesh(m) as s:
        info = s.info()
        assert info["path"] == mmf.PATH_UNIFORM and info["order_exact"] == 1
        s.set_state(mmf.FIELD_U, U)
        assert bits_equal(s.get_state(mmf.FIELD_U), U)
        eig = s.compute_rhs(mmf.FIELD_U)
        got = s.get_state(mmf.FIELD_RHS)
    assert eig == ref_eig
    assert bits_equal(got, ref)


def test_random_state_rhs_bit_exact_all_bcs(mmf, oracle):
    """Random (non-smooth) states exercise every rounding path; all three side BC kinds."""
    rng = np.random.default_rng(7)
    for problem in ("radsod", "vortex_xy"):
        m = oracle.problem_mesh(problem, 3, 16)
        nc = m["volume"].shape[0]
        rho = rng.uniform(0.3, 2.0, nc); vel = rng.uniform(-2, 2, (nc, 3)); p = rng.uniform(0.2, 3, nc)
        U = np.column_stack([rho, rho * vel[:, 0], rho * vel[:, 1], rho * vel[:, 2], p / 0.4 + 0.5 * rho * (vel ** 2).sum(1)])
        ref, ref_eig = oracle.compute_rhs(m, U)
        with mmf.EulerSolver.from_mesh(m) as s:
            s.set_state(mmf.FIELD_U, U)
            eig = s.compute_rhs(mmf.FIELD_U)
            got = s.get_state(mmf.FIELD_RHS)
        assert eig == ref_eig and bits_equal(got, ref), problem


@pytest.mark.parametrize("lz", ["5", "1", "2"])
@pytest.mark.parametrize("cfg", ["", "r12", "r16", "r8", "t12", "t16", "t8", "h12", "h16", "r8:t12:h16:r12"])
def test_fused_steps_bit_exact(mmf, oracle, monkeypatch, cfg, lz):
    """Every stage-kernel form and CTA shape (default first), ragged z chunks down to one- and two-plane chunks."""
    if cfg:
        monkeypatch.setenv("MMF_STAGE_CFG", cfg)
    monkeypatch.setenv("MMF_STAGE_LZ", lz)           # ragged z chunks on purpose
    m = oracle.problem_mesh("vortex_xy", 3, 32)
    U = oracle.init_state(m)
    Uo, Wo, Ro = U.copy(), np.zeros_like(U), np.zeros_like(U)
    with mmf.EulerSolver.from_mesh(m) as s:
        assert s.info()["path"] == mmf.PATH_UNIFORM
        s.set_state(mmf.FIELD_U, U)
        t = 0.0
        for _ in range(10):
            dto, me3 = oracle.step(m, 0.45, t, 2.0, Uo, Wo, Ro)
            dtg, meg = s.step(0.45, m["h"], t, 2.0)
            assert dtg == dto and list(me3) == meg
            t += dto
        assert bits_equal(s.get_state(mmf.FIELD_U), Uo)
        assert bits_equal(s.get_state(mmf.FIELD_W), Wo)


@pytest.mark.parametrize("generic", [False, True])
def test_get_primitives_is_conservative2primitive(mmf, oracle, generic):
    """mmf_get_primitives = the utils::conservative2primitive loop src/main.cpp:511-518 runs before every
    mesh.write(), evaluated on the device: bitwise the oracle's (src/utils.cpp:48-63), on both paths, for
    U and for the work field."""
    import ctypes as C
    D = C.POINTER(C.c_double)
    m = oracle.problem_mesh("vortex_xy", 3, 16)
    rng = np.random.default_rng(5)
    nc = m["volume"].shape[0]
    rho = rng.uniform(0.3, 2.0, nc); vel = rng.uniform(-2, 2, (nc, 3)); p = rng.uniform(0.2, 3, nc)
    U = np.column_stack([rho, rho * vel[:, 0], rho * vel[:, 1], rho * vel[:, 2], p / 0.4 + 0.5 * rho * (vel ** 2).sum(1)])
    W = oracle.init_state(m)

    def c2p(S):
        out = np.empty_like(S)
        for c in range(nc):
            oracle.lib.orc_conservative2primitive(S[c].ctypes.data_as(D), out[c].ctypes.data_as(D))
        return out

    with mmf.EulerSolver.from_mesh(m, flags=mmf.FLAG_FORCE_GENERIC if generic else 0) as s:
        assert s.info()["path"] == (mmf.PATH_GENERIC if generic else mmf.PATH_UNIFORM)
        s.set_state(mmf.FIELD_U, U)
        s.set_state(mmf.FIELD_W, W)
        assert bits_equal(s.get_primitives(mmf.FIELD_U), c2p(U))
        assert bits_equal(s.get_primitives(mmf.FIELD_W), c2p(W))
        assert bits_equal(s.get_state(mmf.FIELD_U), U)          # the resident state is untouched
        with pytest.raises(mmf.MmfError):
            s.get_primitives(mmf.FIELD_RHS)


def _body_meshes(oracle):
    yield "radsod 32^3 + body", oracle.problem_mesh("radsod", 3, 32, boxes=[[2.1, 3.2, 1.3, 4.9, 5.4, 3.6]])
    yield "sod3d_x 16^3 + bodies at the border", oracle.problem_mesh(
        "sod3d_x", 3, 16, boxes=[[-1.1, -1.1, -1.1, -0.6, -0.3, -0.8], [0.3, 0.1, 0.4, 1.1, 1.1, 1.1]])
    m = with_bodies(lexicographic_box_mesh(37, 29, 11, 0.5, 1), [[-1, 4.1, -1, 1.4, 6.4, 9.0], [7.1, 2.1, 1.1, 7.4, 2.4, 1.4],
                                                                [10.1, 0.0, 0.0, 11.9, 9.0, 1.4], [12.6, 0.0, 0.0, 16.4, 2.4, 9.0]])
    m["problem"] = "radsod"
    yield "box 37x29x11 + bodies", m


def test_uniform_path_with_bodies_bit_exact(mmf, oracle, monkeypatch):
    """Kernel form 'b' (the TMA-fed kernel with the flag array, uniform_stage_t.cuh BODY) + wall_cell_update: a uniform
    box with bodies takes the fused path by itself -- RHS, the dt eigenvalue (full pass at the first step, then from the stage-3
    tile estimates, the wall cells and the border ghosts), ten fused steps and the unfused operator sequence, bitwise
    against the oracle; cells that are not solved keep the host's values."""
    monkeypatch.delenv("MMF_UNIFORM_BODIES", raising=False)
    rng = np.random.default_rng(11)
    for name, m in _body_meshes(oracle):
        nc = m["volume"].shape[0]
        if "origin" in m:
            U = oracle.init_state(m)
        else:
            rho = rng.uniform(0.5, 1.5, nc); vel = rng.uniform(-0.4, 0.4, (nc, 3)); p = rng.uniform(0.6, 1.4, nc)
            U = np.column_stack([rho, rho * vel[:, 0], rho * vel[:, 1], rho * vel[:, 2], p / 0.4 + 0.5 * rho * (vel ** 2).sum(1)])
        ref, ref_eig = oracle.compute_rhs(m, U)
        W0 = U[::-1].copy()                      # what the host holds in W: must survive in the unsolved cells
        Uo, Wo, Ro = U.copy(), W0.copy(), np.zeros_like(U)
        with mmf.EulerSolver.from_mesh(m) as s:
            assert s.info()["path"] == mmf.PATH_UNIFORM, name
            s.set_state(mmf.FIELD_U, U)
            s.set_state(mmf.FIELD_W, W0)
            assert s.compute_rhs(mmf.FIELD_U) == ref_eig, name
            assert bits_equal(s.get_state(mmf.FIELD_RHS), ref), name
            t = 0.0
            for _ in range(10):
                dto, me3 = oracle.step(m, 0.45, t, 1e30, Uo, Wo, Ro)
                dtg, meg = s.step(0.45, float(m["size"].min()), t, 1e30)
                assert dtg == dto and list(me3) == meg, name
                t += dto
            assert bits_equal(s.get_state(mmf.FIELD_U), Uo), name
            assert bits_equal(s.get_state(mmf.FIELD_W), Wo), name
            # unfused operators on the same handle (src/main.cpp's own sequence)
            R, me = oracle.compute_rhs(m, Uo)
            assert s.compute_rhs(mmf.FIELD_U) == me, name
            dt = oracle.choose_dt(0.45, float(m["size"].min()), me, 0.0, 4.0)
            oracle.rk_stage(m, 1, dt, Uo, Wo, R); s.rk_stage(1, dt)
            R, me = oracle.compute_rhs(m, Wo)
            assert s.compute_rhs(mmf.FIELD_W) == me, name
            oracle.rk_stage(m, 2, dt, Uo, Wo, R); s.rk_stage(2, dt)
            assert bits_equal(s.get_state(mmf.FIELD_W), Wo), name


def test_bodies_run_to_tmax_then_continue(mmf, oracle, monkeypatch):
    """A box with bodies through mmf_run (CUDA graph, the next step's eigenvalue from the tile estimates + wall cells):
    run to a first tMax -- the trailing steps of the last batch switch themselves off and must keep the eigenvalue
    candidate --, then continue on the same handle; every dt and the final state bitwise the oracle's."""
    monkeypatch.delenv("MMF_UNIFORM_BODIES", raising=False)
    m = oracle.problem_mesh("radsod", 3, 32, boxes=[[2.1, 3.2, 1.3, 4.9, 5.4, 3.6], [6.2, 6.2, 6.2, 8.5, 7.5, 9.0]])
    h = float(m["size"].min())
    U = oracle.init_state(m)
    Uo, Wo, Ro = U.copy(), np.zeros_like(U), np.zeros_like(U)
    t, n = 0.0, 0
    marks = []
    for t_max in (0.05, 0.11):
        while t < t_max:
            dto, _ = oracle.step(m, 0.45, t, t_max, Uo, Wo, Ro)
            t += dto
            n += 1
        marks.append((t, n))
    with mmf.EulerSolver.from_mesh(m) as s:
        assert s.info()["path"] == mmf.PATH_UNIFORM
        s.set_state(mmf.FIELD_U, U)
        t1, n1 = s.run(0.45, h, 0.0, 0.05)
        assert (t1, n1) == marks[0]
        t2, n2 = s.run(0.45, h, t1, 0.11)
        assert (t2, n1 + n2) == marks[1]
        assert bits_equal(s.get_state(mmf.FIELD_U), Uo)


def test_unfused_operator_sequence_on_uniform_path(mmf, oracle):
    m = oracle.problem_mesh("radsod", 3, 16)
    U = oracle.init_state(m)
    with mmf.EulerSolver.from_mesh(m) as s:
        s.set_state(mmf.FIELD_U, U)
        Uo, Wo = U.copy(), np.zeros_like(U)
        for _ in range(2):
            R, me = oracle.compute_rhs(m, Uo)
            assert s.compute_rhs(mmf.FIELD_U) == me
            dt = oracle.choose_dt(0.45, m["h"], me, 0.0, 4.0)
            oracle.rk_stage(m, 1, dt, Uo, Wo, R); s.rk_stage(1, dt)
            R, me = oracle.compute_rhs(m, Wo)
            assert s.compute_rhs(mmf.FIELD_W) == me
            oracle.rk_stage(m, 2, dt, Uo, Wo, R); s.rk_stage(2, dt)
            R, me = oracle.compute_rhs(m, Wo)
            assert s.compute_rhs(mmf.FIELD_W) == me
            oracle.rk_stage(m, 3, dt, Uo, Wo, R); s.rk_stage(3, dt)
        assert bits_equal(s.get_state(mmf.FIELD_U), Uo)


@pytest.mark.parametrize("case", GOLDEN_3D, ids=lambda c: c["name"])
def test_reference_golden_strings_uniform_path(mmf, oracle, case):
    m = oracle.problem_mesh(case["problem"], 3, case["n_cells"])
    t_end = case["t_end"] if case["t_end"] >= 0 else oracle.end_time(case["problem"], 3)
    with mmf.EulerSolver.from_mesh(m) as s:
        assert s.info()["path"] == mmf.PATH_UNIFORM
        s.set_state(mmf.FIELD_U, oracle.init_state(m))
        t, steps = s.run(case["cfl"], m["h"], 0.0, t_end)
        U = s.get_state(mmf.FIELD_U)
    assert steps == case["steps"] and t == t_end
    assert oracle_lib.format_error(oracle.error_norm(m, U, t_end)) == case["expected"]
    ref = oracle.run(case["problem"], 3, case["n_cells"], t_end=case["t_end"], cfl=case["cfl"], want_state=True)
    assert bits_equal(U, ref["U"])


def test_compact_descriptor_equals_full_descriptor(mmf, oracle):
    """mmf_create_uniform (no connectivity arrays) == mmf_create(full description)."""
    m = oracle.problem_mesh("radsod", 3, 32)
    U = oracle.init_state(m)
    with mmf.EulerSolver.from_mesh(m) as a, mmf.EulerSolver.uniform((32, 32, 32), m["h"], _bc_sides(m)) as b:
        for s in (a, b):
            s.set_state(mmf.FIELD_U, U)
        ta, na = a.run(0.45, m["h"], 0.0, 1e30, max_steps=6)
        tb, nb = b.run(0.45, m["h"], 0.0, 1e30, max_steps=6)
        assert (ta, na) == (tb, nb)
        assert bits_equal(a.get_state(mmf.FIELD_U), b.get_state(mmf.FIELD_U))


def test_compact_descriptor_takes_host_area_and_volume(mmf, oracle):
    """The compact description builds area = h*h and volume = h*h*h unless the host passes ITS values
    (mmf_uniform_desc.area / .volume): values one ulp away from the products must be used verbatim, like the
    tables of a full description (src/mesh_info.cpp:86-118 caches whatever the mesh evaluates)."""
    m = oracle.problem_mesh("radsod", 3, 16)
    h = m["h"]
    A, V = np.nextafter(h * h, 1.0), np.nextafter(h * h * h, 0.0)
    m["area"] = np.full_like(m["area"], A)
    m["volume"] = np.full_like(m["volume"], V)
    U = oracle.init_state(m)
    Uo, Wo, Ro = U.copy(), np.zeros_like(U), np.zeros_like(U)
    t = 0.0
    for _ in range(4):
        dto, _ = oracle.step(m, 0.45, t, 1e30, Uo, Wo, Ro)
        t += dto
    with mmf.EulerSolver.uniform((16, 16, 16), h, _bc_sides(m), area=A, volume=V) as b, \
            mmf.EulerSolver.uniform((16, 16, 16), h, _bc_sides(m)) as c:
        b.set_state(mmf.FIELD_U, U)
        c.set_state(mmf.FIELD_U, U)
        assert b.run(0.45, h, 0.0, 1e30, max_steps=4) == (t, 4)
        c.run(0.45, h, 0.0, 1e30, max_steps=4)
        assert bits_equal(b.get_state(mmf.FIELD_U), Uo)
        assert not bits_equal(c.get_state(mmf.FIELD_U), Uo)  # the products differ by an ulp, and it shows
    with pytest.raises(mmf.MmfError):
        mmf.EulerSolver.uniform((16, 16, 16), h, _bc_sides(m), area=A)  # one of the two only


def test_lexicographic_numbering_non_cubic_box(mmf, oracle):
    """Non-cubic, non-power-of-two box numbered lexicographically: uniform path must detect the
    numbering and stay bit-exact; ragged tiles in x (37 = 30 + 7) and y."""
    m = lexicographic_box_mesh(37, 19, 11, 0.125, 1)
    m["problem"] = "radsod"
    rng = np.random.default_rng(11)
    nc = m["volume"].shape[0]
    rho = rng.uniform(0.5, 1.5, nc); vel = rng.uniform(-1, 1, (nc, 3)); p = rng.uniform(0.5, 1.5, nc)
    U = np.column_stack([rho, rho * vel[:, 0], rho * vel[:, 1], rho * vel[:, 2], p / 0.4 + 0.5 * rho * (vel ** 2).sum(1)])
    Uo, Wo, Ro = U.copy(), np.zeros_like(U), np.zeros_like(U)
    with mmf.EulerSolver.from_mesh(m) as s:
        info = s.info()
        assert info["path"] == mmf.PATH_UNIFORM and info["order_exact"] == 1
        s.set_state(mmf.FIELD_U, U)
        ref, ref_eig = oracle.compute_rhs(m, U)
        assert s.compute_rhs(mmf.FIELD_U) == ref_eig
        assert bits_equal(s.get_state(mmf.FIELD_RHS), ref)
        t = 0.0
        for _ in range(4):
            dto, _ = oracle.step(m, 0.45, t, 1e30, Uo, Wo, Ro)
            dtg, _ = s.step(0.45, 0.125, t, 1e30)
            assert dtg == dto
            t += dto
        assert bits_equal(s.get_state(mmf.FIELD_U), Uo)


def test_axis_order_flag_is_within_tolerance_not_exact(mmf, oracle):
    """MMF_FLAG_ORDER_AXIS trades the host's accumulation order for a fixed one: results move by
    rounding only (north_star tolerance 1e-12 relative)."""
    m = oracle.problem_mesh("vortex_xy", 3, 32)
    U = oracle.init_state(m)
    Uo, Wo, Ro = U.copy(), np.zeros_like(U), np.zeros_like(U)
    with mmf.EulerSolver.from_mesh(m, flags=mmf.FLAG_ORDER_AXIS) as s:
        assert s.info()["order_exact"] == 0
        s.set_state(mmf.FIELD_U, U)
        t = 0.0
        for _ in range(10):
            dto, _ = oracle.step(m, 0.45, t, 2.0, Uo, Wo, Ro)
            s.step(0.45, m["h"], t, 2.0)
            t += dto
        assert max_rel_diff(s.get_state(mmf.FIELD_U), Uo) <= 1e-12


def test_mesh_with_bodies_path_selection(mmf, oracle, monkeypatch):
    """A uniform box with bodies takes the fused path (kernel form 'b'); MMF_UNIFORM_BODIES=0 and MMF_FLAG_FORCE_GENERIC
    keep it on the generic one."""
    boxes = np.array([[3.0, 3.0, 3.0, 5.0, 5.0, 5.0]])
    m = oracle.problem_mesh("radsod", 3, 16, boxes=boxes)
    monkeypatch.delenv("MMF_UNIFORM_BODIES", raising=False)
    with mmf.EulerSolver.from_mesh(m) as s:
        assert s.info()["path"] == mmf.PATH_UNIFORM
    with mmf.EulerSolver.from_mesh(m, flags=mmf.FLAG_FORCE_GENERIC) as s:
        assert s.info()["path"] == mmf.PATH_GENERIC
    monkeypatch.setenv("MMF_UNIFORM_BODIES", "0")
    with mmf.EulerSolver.from_mesh(m) as s:
        assert s.info()["path"] == mmf.PATH_GENERIC


@pytest.mark.parametrize("n,steps", [(64, 10), (128, 10), (256, 3)])
def test_bench_configuration_bit_exact(mmf, oracle, n, steps):
    """The configuration bench.py times -- compact descriptor, LEXICOGRAPHIC cell numbering with the MORTON
    interface order, free-flow load clamps, default kernel mix -- against the oracle at 64^3 / 128^3 (10 steps,
    SURVEY 8d) and at the full benchmark size 256^3 (3 steps, the threaded loop of the oracle, which is bitwise the
    serial one: tests/test_oracle_units.py).  The oracle's cells are Morton numbered: the per-cell values do not
    depend on the CELL numbering (the accumulation order follows the interface numbering), so they are compared
    cell by cell after the permutation.  Bitwise."""
    level = n.bit_length() - 1
    h = 10.0 / n
    U0m, Urefm = oracle.run_threads("vortex_xy", 3, level, steps)
    perm = oracle_lib.morton_to_lexicographic(n)
    U0 = np.empty_like(U0m); U0[perm] = U0m
    with mmf.EulerSolver.uniform((n, n, n), h, [0] * 6, cell_numbering=mmf.NUMBERING_LEXICOGRAPHIC,
                                 interface_numbering=mmf.NUMBERING_MORTON) as s:
        info = s.info()
        assert info["path"] == mmf.PATH_UNIFORM and info["order_exact"] == 1
        s.set_state(mmf.FIELD_U, U0)
        t, done = s.run(0.45, h, 0.0, 1e30, max_steps=steps)
        assert done == steps
        got = s.get_state(mmf.FIELD_U)
    assert bits_equal(got[perm], Urefm), max_rel_diff(got[perm], Urefm)


def test_large_mesh_properties(mmf):
    """Full benchmark size (256^3): properties that need no oracle.
    (1) a uniform free-stream state is a fixed point of the scheme (every interface flux cancels);
    (2) the total of each conserved variable changes only by the fluxes through the free-flow borders:
        with the vortex far from them, mass and energy move by < 1e-10 and momentum by < 1e-7 relative;
    (3) the run is deterministic (no atomics on data): two runs give identical bits."""
    n, L = 256, 10.0
    h = L / n
    with mmf.EulerSolver.uniform((n, n, n), h, [0] * 6, cell_numbering=mmf.NUMBERING_LEXICOGRAPHIC,
                                 interface_numbering=mmf.NUMBERING_MORTON) as s:
        rho, u, v, w, p = 1.0, 1.0, 1.0, 0.0, 1.0
        U = np.empty((n ** 3, 5))
        U[:] = [rho, rho * u, rho * v, rho * w, p / 0.4 + 0.5 * rho * (u * u + v * v + w * w)]
        s.set_state(mmf.FIELD_U, U)
        s.run(0.45, h, 0.0, 1e30, max_steps=2)
        out = s.get_state(mmf.FIELD_U)
        assert max_rel_diff(out, U) <= 1e-14
        # isentropic vortex centred in the box
        x = (np.arange(n) + 0.5) * h - 5.0
        X, Y = np.meshgrid(x, x, indexing="xy")
        r2 = X * X + Y * Y
        shape = 5.0 / (2 * np.pi) * np.exp(0.5 * (1 - r2))
        T = 1.0 - 0.4 / 2.8 * shape * shape
        pp = T ** 3.5
        rr = pp / T
        uu, vv = 1.0 - Y * shape, 1.0 + X * shape
        plane = np.stack([rr, rr * uu, rr * vv, 0 * rr, pp / 0.4 + 0.5 * rr * (uu * uu + vv * vv)], axis=-1)
        U = np.broadcast_to(plane[None], (n, n, n, 5)).reshape(-1, 5).copy()
        s.set_state(mmf.FIELD_U, U)
        s.run(0.45, h, 0.0, 1e30, max_steps=3)
        a = s.get_state(mmf.FIELD_U).copy()
        tot0, tot1 = U.sum(0), a.sum(0)
        assert np.all(np.abs(tot1 - tot0)[[0, 4]] <= 1e-10 * np.abs(tot0).max())
        assert np.all(np.abs(tot1 - tot0) <= 1e-7 * np.abs(tot0).max())
        assert np.isfinite(a).all() and a[:, 0].min() > 0.3
        s.set_state(mmf.FIELD_U, U)
        s.run(0.45, h, 0.0, 1e30, max_steps=3)
        assert bits_equal(s.get_state(mmf.FIELD_U), a)


@pytest.mark.parametrize("problem", ["vortex_xy", "radsod"])
def test_next_step_eigenvalue_from_stage3_estimates(mmf, oracle, problem):
    """Behind stage 3 the next step's max eigenvalue is found by re-evaluating exactly only the tiles
    whose FP32 estimate comes close to the largest estimate.  Two hard cases: a single hot cell whose
    maximum collapses within a step (the maximum jumps between tiles), and the radial Sod problem whose
    maximum sits on a plateau of identical cells (many tiles listed).  dt and the three logged
    eigenvalues must equal the oracle's bit for bit at every step, on ragged tiles."""
    m = oracle.problem_mesh(problem, 3, 16)
    U = oracle.init_state(m)
    hot = 16 * 16 * 8 + 16 * 8 + 8
    rho = U[hot, 0]
    U[hot, 1:4] = rho * np.array([6.0, -5.0, 4.0])
    U[hot, 4] += 0.5 * rho * (36 + 25 + 16)
    Uo, Wo, Ro = U.copy(), np.zeros_like(U), np.zeros_like(U)
    with mmf.EulerSolver.from_mesh(m) as s:
        assert s.info()["path"] == mmf.PATH_UNIFORM
        s.set_state(mmf.FIELD_U, U)
        t = 0.0
        for _ in range(12):
            dto, me3 = oracle.step(m, 0.45, t, 1e30, Uo, Wo, Ro)
            dtg, meg = s.step(0.45, m["h"], t, 1e30)
            assert dtg == dto and list(me3) == meg
            t += dto
        assert bits_equal(s.get_state(mmf.FIELD_U), Uo)
        # and through the batched loop (no host round trip between the steps)
        s.set_state(mmf.FIELD_U, U)
        tb, nb = s.run(0.45, m["h"], 0.0, 1e30, max_steps=12)
        assert nb == 12 and tb == t
        assert bits_equal(s.get_state(mmf.FIELD_U), Uo)


@pytest.mark.parametrize("kind", ["radsod", "dirichlet"])
def test_run_to_tmax_then_continue(mmf, oracle, kind):
    """mmf_run enqueues steps in batches; those past tMax switch themselves off on the device.  A second run on the
    same handle (a later tMax) must start from the right max eigenvalue -- including the boundary ghosts' own
    (reflecting images, a Dirichlet inflow that holds the maximum) -- and a run called at t >= tMax must change
    nothing.  dt sequence and state bitwise against the oracle stepping with the same two end times."""
    if kind == "radsod":
        m = oracle.problem_mesh("radsod", 3, 16)
        U = oracle.init_state(m)
        make = lambda: mmf.EulerSolver.from_mesh(m)
    else:
        m = lexicographic_box_mesh(14, 9, 5, 0.25, 1)
        m["problem"] = "ffstep"
        border = m["neigh"] < 0
        m["bc"][border & (m["normal"][:, 0] < 0)] = 3       # -x Dirichlet (u = 3: it holds the largest eigenvalue)
        m["bc"][border & (m["normal"][:, 0] > 0)] = 0
        rng = np.random.default_rng(5)
        nc = m["volume"].shape[0]
        rho = rng.uniform(0.9, 1.1, nc); vel = rng.uniform(-0.1, 0.1, (nc, 3)); p = rng.uniform(0.9, 1.1, nc)
        U = np.column_stack([rho, rho * vel[:, 0], rho * vel[:, 1], rho * vel[:, 2], p / 0.4 + 0.5 * rho * (vel ** 2).sum(1)])
        make = lambda: mmf.EulerSolver.from_mesh(m, problem_type=7, dirichlet_info=[1.0, 3.0, 0.0, 0.0, 1.0 / 1.4])
    h = float(m["size"].min())
    # oracle: main.cpp's loop with tMax = T1, then continued with tMax = T2
    dt0, _ = oracle.step(m, 0.45, 0.0, 1e30, U.copy(), np.zeros_like(U), np.zeros_like(U))
    T1, T2 = 3.4 * dt0, 7.7 * dt0
    Uo, Wo, Ro = U.copy(), np.zeros_like(U), np.zeros_like(U)
    t, n1, n2 = 0.0, 0, 0
    while t < T1:
        t += oracle.step(m, 0.45, t, T1, Uo, Wo, Ro)[0]; n1 += 1
    U1 = Uo.copy()
    while t < T2:
        t += oracle.step(m, 0.45, t, T2, Uo, Wo, Ro)[0]; n2 += 1
    with make() as s:
        assert s.info()["path"] == mmf.PATH_UNIFORM
        s.set_state(mmf.FIELD_U, U)
        assert s.run(0.45, h, T1, T1) == (T1, 0)             # first call at t >= tMax: nothing happens
        assert bits_equal(s.get_state(mmf.FIELD_U), U)
        ta, na = s.run(0.45, h, 0.0, T1)
        assert (ta, na) == (T1, n1) and bits_equal(s.get_state(mmf.FIELD_U), U1)
        assert s.run(0.45, h, ta, T1) == (T1, 0)
        tb, nb = s.run(0.45, h, ta, T2)
        assert (tb, nb) == (T2, n2)
        assert bits_equal(s.get_state(mmf.FIELD_U), Uo)
        # and step by step through mmf_step, with a step at t >= tMax in between
        s.set_state(mmf.FIELD_U, U)
        t = 0.0
        while t < T1:
            t += s.step(0.45, h, t, T1)[0]
        assert s.step(0.45, h, t, T1)[0] == 0.0
        while t < T2:
            t += s.step(0.45, h, t, T2)[0]
        assert t == T2 and bits_equal(s.get_state(mmf.FIELD_U), Uo)


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 3), (31, 11, 1), (30, 10, 4), (61, 21, 5)])
@pytest.mark.parametrize("bc_code", [0, 1])
def test_degenerate_and_ragged_boxes(mmf, oracle, dims, bc_code):
    """One-cell-thick boxes, a single cell, exact tile multiples and one-past multiples (x window = 30
    cells, y tile = rows - 2), free-flow (clamped loads) and reflecting (ghost pass) borders: residual,
    eigenvalue and a few fused steps must equal the oracle bit for bit."""
    m = lexicographic_box_mesh(dims[0], dims[1], dims[2], 0.25, bc_code)
    m["problem"] = "radsod" if bc_code == 1 else "vortex_xy"
    rng = np.random.default_rng(dims[0] * 100 + dims[1] * 10 + dims[2] + bc_code)
    nc = m["volume"].shape[0]
    rho = rng.uniform(0.5, 1.5, nc); vel = rng.uniform(-1, 1, (nc, 3)); p = rng.uniform(0.5, 1.5, nc)
    U = np.column_stack([rho, rho * vel[:, 0], rho * vel[:, 1], rho * vel[:, 2], p / 0.4 + 0.5 * rho * (vel ** 2).sum(1)])
    Uo, Wo, Ro = U.copy(), np.zeros_like(U), np.zeros_like(U)
    with mmf.EulerSolver.from_mesh(m) as s:
        assert s.info()["path"] == mmf.PATH_UNIFORM
        s.set_state(mmf.FIELD_U, U)
        ref, ref_eig = oracle.compute_rhs(m, U)
        assert s.compute_rhs(mmf.FIELD_U) == ref_eig
        assert bits_equal(s.get_state(mmf.FIELD_RHS), ref)
        t = 0.0
        for _ in range(3):
            dto, me3 = oracle.step(m, 0.45, t, 1e30, Uo, Wo, Ro)
            dtg, meg = s.step(0.45, 0.25, t, 1e30)
            assert dtg == dto and list(me3) == meg
            t += dto
        assert bits_equal(s.get_state(mmf.FIELD_U), Uo)
