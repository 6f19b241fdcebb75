"""GPU parity, generic (connectivity-driven) path, through the C-ABI: bit-exact against the oracle.

Integer/index work is exact by construction (the host's raw ids and interface order are inputs);
FP64 results are compared BITWISE (+0 == -0), which is stronger than north_star's 1e-12 relative."""
import numpy as np
import pytest

import oracle_lib
from common import bits_equal, golden_cases, lexicographic_box_mesh, max_rel_diff, two_level_mesh

import os

pytestmark = pytest.mark.gpu


GENERIC = 1  # MMF_FLAG_FORCE_GENERIC


def _solver(mmf, m, flags=GENERIC, **kw):
    s = mmf.EulerSolver.from_mesh(m, flags=flags, **kw)
    return s


@pytest.mark.parametrize("problem,dim,n", [("vortex_xy", 2, 64), ("radsod", 2, 64), ("vortex_xy", 3, 32),
                                           ("vortex_yz", 3, 16), ("radsod", 3, 32), ("sod3d_x", 3, 16)])
def test_compute_rhs_bit_exact(mmf, oracle, problem, dim, n):
    m = oracle.problem_mesh(problem, dim, n)
    U = oracle.init_state(m)
    ref, ref_eig = oracle.compute_rhs(m, U)
    with _solver(mmf, m) as s:
        assert s.info()["path"] == mmf.PATH_GENERIC
        s.set_state(mmf.FIELD_U, U)
        assert bits_equal(s.get_state(mmf.FIELD_U), U)          # AoS <-> SoA round trip
        s.compute_polynomials(mmf.FIELD_U)                       # no-op, call shape only
        eig = s.compute_rhs(mmf.FIELD_U, order=1)
        got = s.get_state(mmf.FIELD_RHS)
    assert eig == ref_eig
    assert bits_equal(got, ref)


def test_strict_host_adapter_matches(mmf, oracle):
    m = oracle.problem_mesh("radsod", 3, 16)
    U = oracle.init_state(m)
    ref, ref_eig = oracle.compute_rhs(m, U)
    with _solver(mmf, m) as s:
        got, eig = s.compute_rhs_host(U)
    assert eig == ref_eig and bits_equal(got, ref)


def test_unsupported_order_is_an_error(mmf, oracle):
    m = oracle.problem_mesh("vortex_xy", 2, 8)
    with _solver(mmf, m) as s:
        s.set_state(mmf.FIELD_U, oracle.init_state(m))
        with pytest.raises(mmf.MmfError) as e:
            s.compute_rhs(mmf.FIELD_U, order=2)
        assert e.value.code == 3  # MMF_ERR_UNSUPPORTED_ORDER (reference: exit(2))


def test_call_before_set_state_is_an_error(mmf, oracle):
    m = oracle.problem_mesh("vortex_xy", 2, 8)
    with _solver(mmf, m) as s:
        with pytest.raises(mmf.MmfError) as e:
            s.compute_rhs(mmf.FIELD_U)
        assert e.value.code == 6


def test_bad_mesh_is_rejected(mmf, oracle):
    m = oracle.problem_mesh("vortex_xy", 2, 8)
    bad = dict(m); bad["bc"] = m["bc"].copy(); bad["bc"][0] = 7
    with pytest.raises(mmf.MmfError):
        _solver(mmf, bad)
    bad = dict(m); bad["owner"] = m["owner"].copy(); bad["owner"][3] = 10 ** 6
    with pytest.raises(mmf.MmfError):
        _solver(mmf, bad)


def test_bodies_wall_faces_bit_exact(mmf, oracle):
    """Immersed boxes -> BC_WALL staircase faces, incl. the flipped-normal branch
    (src/euler.cpp:198-225): solid cell is the owner of some faces and the neighbour of others."""
    boxes = np.array([[2.9, 2.9, 2.9, 5.1, 5.1, 5.1], [0.0, 6.0, 0.0, 1.2, 8.0, 8.0]])
    m = oracle.problem_mesh("radsod", 3, 16, boxes=boxes)
    solid = m["fluid"] == 0
    inter = m["neigh"] >= 0
    assert (solid[m["owner"]] & inter & ~solid[np.maximum(m["neigh"], 0)]).any()   # flipNormal case
    assert (~solid[m["owner"]] & inter & solid[np.maximum(m["neigh"], 0)]).any()
    U = oracle.init_state(m)
    rng = np.random.default_rng(0)
    U[:, 1:4] = rng.uniform(-0.3, 0.3, (U.shape[0], 3)) * U[:, :1]
    U[:, 4] += 0.5 * (U[:, 1:4] ** 2).sum(1) / U[:, 0]
    ref, ref_eig = oracle.compute_rhs(m, U)
    with _solver(mmf, m) as s:
        s.set_state(mmf.FIELD_U, U)
        eig = s.compute_rhs(mmf.FIELD_U)
        got = s.get_state(mmf.FIELD_RHS)
        # 10 full steps with bodies
        Uo, Wo, Ro = U.copy(), np.zeros_like(U), np.zeros_like(U)
        t = 0.0
        for _ in range(10):
            dt, _ = oracle.step(m, 0.45, t, 4.0, Uo, Wo, Ro)
            dtg, _ = s.step(0.45, m["h"], t, 4.0)
            assert dtg == dt
            t += dt
        Ug = s.get_state(mmf.FIELD_U)
    assert eig == ref_eig and bits_equal(got, ref)
    assert np.all(got[solid] == 0.0)
    assert bits_equal(Ug, Uo)


def test_dirichlet_ffstep_bit_exact(mmf, oracle):
    """BC_DIRICHLET (only ffstep fills data, src/problem.cpp:450-477) on a hand-built box."""
    m = lexicographic_box_mesh(12, 6, 4, 0.25, 1)
    m["problem"] = "ffstep"
    # -x side Dirichlet, +x side free flow, others reflecting (src/problem.cpp:424-433 semantics)
    border = m["neigh"] < 0
    nx_ = m["normal"][:, 0]
    m["bc"][border & (nx_ < 0)] = 3
    m["bc"][border & (nx_ > 0)] = 0
    rng = np.random.default_rng(3)
    nc = m["volume"].shape[0]
    rho = rng.uniform(0.8, 1.2, nc); vel = rng.uniform(-0.5, 3.0, (nc, 3)); p = rng.uniform(0.8, 1.2, nc)
    U = np.column_stack([rho, rho * vel[:, 0], rho * vel[:, 1], rho * vel[:, 2], p / 0.4 + 0.5 * rho * (vel ** 2).sum(1)])
    ref, ref_eig = oracle.compute_rhs(m, U)
    info = [1.0, 3.0, 0.0, 0.0, 1.0 / 1.4]
    with _solver(mmf, m, problem_type=7, dirichlet_info=info) as s:
        s.set_state(mmf.FIELD_U, U)
        eig = s.compute_rhs(mmf.FIELD_U)
        got = s.get_state(mmf.FIELD_RHS)
    assert eig == ref_eig and bits_equal(got, ref)


@pytest.mark.parametrize("dim", [2, 3])
def test_hanging_faces_bit_exact(mmf, oracle, dim):
    """Non-uniform (2:1) octree: coarse cells meet 2^(dim-1) interfaces on a refined side, interfaces
    are owned by the finer cell so interior normals of both signs occur, and cell sizes differ (volume
    in the RK update, min size in dt). Residual, stage updates and fused steps are bit-exact."""
    m = two_level_mesh(dim, 6, lambda i, j, k: (1 <= i < 4 and 2 <= j < 5 and k < 3) or (i, j, k) == (5, 5, 0) or (i + j + k) % 5 == 0)
    m["problem"] = "radsod"
    nc = m["volume"].shape[0]
    rng = np.random.default_rng(11)
    rho = rng.uniform(0.5, 1.5, nc); vel = rng.uniform(-0.4, 0.4, (nc, 3)); p = rng.uniform(0.6, 1.4, nc)
    if dim == 2:
        vel[:, 2] = 0.0
    U = np.column_stack([rho, rho * vel[:, 0], rho * vel[:, 1], rho * vel[:, 2], p / 0.4 + 0.5 * rho * (vel ** 2).sum(1)])
    ref, ref_eig = oracle.compute_rhs(m, U)
    with _solver(mmf, m) as s:
        assert s.info()["path"] == mmf.PATH_GENERIC
        s.set_state(mmf.FIELD_U, U)
        eig = s.compute_rhs(mmf.FIELD_U)
        got = s.get_state(mmf.FIELD_RHS)
        assert eig == ref_eig and bits_equal(got, ref)
        Uo, Wo, Ro = U.copy(), np.zeros_like(U), np.zeros_like(U)
        t = 0.0
        for _ in range(6):
            dt, me3 = oracle.step(m, 0.45, t, 10.0, Uo, Wo, Ro)
            dtg, meg = s.step(0.45, m["h"], t, 10.0)
            assert dtg == dt and list(me3) == meg
            t += dt
        assert bits_equal(s.get_state(mmf.FIELD_U), Uo)


def test_rk_stages_and_step_bit_exact(mmf, oracle):
    m = oracle.problem_mesh("vortex_xy", 3, 16)
    U = oracle.init_state(m)
    with _solver(mmf, m) as s:
        s.set_state(mmf.FIELD_U, U)
        # reference-shaped host loop: computeRHS -> dt on the host -> RK loop, three times
        Uo, Wo = U.copy(), np.zeros_like(U)
        R, me = oracle.compute_rhs(m, Uo)
        assert s.compute_rhs(mmf.FIELD_U) == me
        dt = oracle.choose_dt(0.45, m["h"], me, 0.0, 2.0)
        oracle.rk_stage(m, 1, dt, Uo, Wo, R); s.rk_stage(1, dt)
        assert bits_equal(s.get_state(mmf.FIELD_W), Wo)
        R, me = oracle.compute_rhs(m, Wo)
        assert s.compute_rhs(mmf.FIELD_W) == me
        oracle.rk_stage(m, 2, dt, Uo, Wo, R); s.rk_stage(2, dt)
        assert bits_equal(s.get_state(mmf.FIELD_W), Wo)
        R, me = oracle.compute_rhs(m, Wo)
        assert s.compute_rhs(mmf.FIELD_W) == me
        oracle.rk_stage(m, 3, dt, Uo, Wo, R); s.rk_stage(3, dt)
        assert bits_equal(s.get_state(mmf.FIELD_U), Uo)
        # resident fused-call step
        Ro = np.zeros_like(U)
        t = dt
        for _ in range(5):
            dto, me3 = oracle.step(m, 0.45, t, 2.0, Uo, Wo, Ro)
            dtg, meg = s.step(0.45, m["h"], t, 2.0)
            assert dtg == dto and list(me3) == meg
            t += dto
        assert bits_equal(s.get_state(mmf.FIELD_U), Uo)


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_reference_golden_strings_on_gpu(mmf, oracle, case):
    """The whole main.cpp loop on the GPU reproduces the reference's printed 'Final error' digits."""
    m = oracle.problem_mesh(case["problem"], case["dim"], case["n_cells"])
    t_end = case["t_end"] if case["t_end"] >= 0 else oracle.end_time(case["problem"], m["dim"])
    with _solver(mmf, m) as s:
        s.set_state(mmf.FIELD_U, oracle.init_state(m))
        t, steps = s.run(case["cfl"], float(m["size"].min()), 0.0, t_end)
        U = s.get_state(mmf.FIELD_U)
    assert steps == case["steps"] and t == t_end
    assert oracle_lib.format_error(oracle.error_norm(m, U, t_end)) == case["expected"]
    ref = oracle.run(case["problem"], case["dim"], case["n_cells"], t_end=case["t_end"], cfl=case["cfl"], want_state=True)
    assert bits_equal(U, ref["U"])


def test_run_respects_max_steps_and_tmax_clamp(mmf, oracle):
    m = oracle.problem_mesh("radsod", 2, 32)
    with _solver(mmf, m) as s:
        s.set_state(mmf.FIELD_U, oracle.init_state(m))
        t, steps = s.run(0.45, m["h"], 0.0, 1.0e30, max_steps=7)
        assert steps == 7 and 0 < t < 1.0
        t2, steps2 = s.run(0.45, m["h"], t, t + 1e-3)      # one clamped step lands exactly on t_max
        assert steps2 == 1 and t2 == t + 1e-3
        ref = oracle.run("radsod", 2, 32, t_end=1.0e30, max_steps=7, want_state=True)
    assert ref["t"] == t


@pytest.mark.parametrize("fused", ["1", "0"])
@pytest.mark.parametrize("kind", ["vortex2d", "bodies3d", "hanging3d"])
def test_generic_fused_stages_bit_exact(mmf, oracle, monkeypatch, kind, fused):
    """The generic path's step: residual + RK update of stages 2 and 3 in one kernel each, two work arrays swapping
    roles (the default) and the unfused sequence (MMF_GENERIC_FUSED=0); dt, the three logged eigenvalues, U, W and
    the stage-3 residual left in RHS are those of the oracle."""
    monkeypatch.setenv("MMF_GENERIC_FUSED", fused)
    if kind == "vortex2d":
        m = oracle.problem_mesh("vortex_xy", 2, 64)
        U = oracle.init_state(m)
    elif kind == "bodies3d":
        m = oracle.problem_mesh("radsod", 3, 16, boxes=np.array([[2.9, 2.9, 2.9, 5.1, 5.1, 5.1], [0.0, 6.0, 0.0, 1.2, 8.0, 8.0]]))
        U = oracle.init_state(m)
    else:
        m = two_level_mesh(3, 6, lambda i, j, k: (1 <= i < 4 and 2 <= j < 5 and k < 3) or (i + j + k) % 5 == 0)
        m["problem"] = "radsod"
        nc = m["volume"].shape[0]
        rng = np.random.default_rng(11)
        rho = rng.uniform(0.5, 1.5, nc); vel = rng.uniform(-0.4, 0.4, (nc, 3)); p = rng.uniform(0.6, 1.4, nc)
        U = np.column_stack([rho, rho * vel[:, 0], rho * vel[:, 1], rho * vel[:, 2], p / 0.4 + 0.5 * rho * (vel ** 2).sum(1)])
    with _solver(mmf, m, flags=mmf.FLAG_FORCE_GENERIC) as s:
        assert s.info()["path"] == mmf.PATH_GENERIC
        s.set_state(mmf.FIELD_U, U)
        s.set_state(mmf.FIELD_W, U)
        Uo, Wo, Ro = U.copy(), U.copy(), np.zeros_like(U)
        t = 0.0
        for _ in range(6):
            dt, me3 = oracle.step(m, 0.45, t, 10.0, Uo, Wo, Ro)
            dtg, meg = s.step(0.45, float(m["size"].min()), t, 10.0)
            assert dtg == dt and list(me3) == meg
            t += dt
        assert bits_equal(s.get_state(mmf.FIELD_U), Uo)
        assert bits_equal(s.get_state(mmf.FIELD_W), Wo)
        assert bits_equal(s.get_state(mmf.FIELD_RHS), Ro)
