"""Unit checks of the oracle's building blocks and of properties the GPU design relies on."""
import ctypes as C

import numpy as np
import pytest

from common import bits_equal, lexicographic_box_mesh, two_level_mesh
from oracle_lib import PROBLEMS

_D = C.POINTER(C.c_double)


def _splitting(oracle, L, R, n):
    L, R, n = (np.ascontiguousarray(x, dtype=np.float64) for x in (L, R, n))
    f = np.zeros(5)
    lam = C.c_double(0)
    oracle.lib.orc_eval_splitting(L.ctypes.data_as(_D), R.ctypes.data_as(_D), n.ctypes.data_as(_D),
                                  f.ctypes.data_as(_D), C.byref(lam))
    return f, lam.value


def _random_states(rng, n):
    rho = rng.uniform(0.3, 2.0, n)
    vel = rng.uniform(-2.0, 2.0, (n, 3))
    p = rng.uniform(0.2, 3.0, n)
    E = p / 0.4 + 0.5 * rho * (vel ** 2).sum(1)
    return np.column_stack([rho, rho * vel[:, 0], rho * vel[:, 1], rho * vel[:, 2], E])


def test_mesh_counts_and_conventions(oracle):
    m = oracle.uniform_mesh(3, (-5, -5, -5), 10.0, 8)
    assert m["volume"].shape[0] == 512 and m["owner"].shape[0] == 3 * 64 * 9
    assert np.all(m["volume"] == 1.25 ** 3) and np.all(m["area"] == 1.25 ** 2)
    # Morton order, x lowest bit
    assert m["cell_ijk"][1].tolist() == [1, 0, 0] and m["cell_ijk"][2].tolist() == [0, 1, 0]
    assert m["cell_ijk"][4].tolist() == [0, 0, 1] and m["cell_ijk"][7].tolist() == [1, 1, 1]
    interior = m["neigh"] >= 0
    # interior faces: owner is the lower cell, normal is +e_d from owner to neigh
    assert np.all(m["owner"][interior] < m["neigh"][interior])
    d = m["cell_ijk"][m["neigh"][interior]] - m["cell_ijk"][m["owner"][interior]]
    assert np.array_equal(d.astype(float), m["normal"][interior])
    # border faces: outward normal
    b = ~interior
    ijk = m["cell_ijk"][m["owner"][b]]
    nrm = m["normal"][b]
    axis = np.abs(nrm).argmax(1)
    coord = ijk[np.arange(len(axis)), axis]
    sign = nrm[np.arange(len(axis)), axis]
    assert np.all(np.where(sign > 0, coord == 7, coord == 0))
    # 2-D: z centroid stays at the origin's z
    m2 = oracle.uniform_mesh(2, (0, 0, 0), 8.0, 64)
    assert m2["volume"].shape[0] == 4096 and m2["owner"].shape[0] == 8320
    assert np.all(m2["ccentroid"][:, 2] == 0.0) and np.all(m2["area"] == 0.125)


def test_level_rounds_up_to_power_of_two(oracle):
    assert oracle.lib.orc_level_for(10.0, 32) == 5
    assert oracle.lib.orc_level_for(10.0, 33) == 6
    assert oracle.lib.orc_level_for(10.0, 1) == 0


def test_llf_antisymmetry_is_exact(oracle):
    """F(L,R,n) == -F(R,L,-n) bit for bit: lets the GPU compute each interface once."""
    rng = np.random.default_rng(1)
    S = _random_states(rng, 400)
    for q in range(0, 400, 2):
        for d in range(3):
            n = np.zeros(3); n[d] = 1.0
            f1, l1 = _splitting(oracle, S[q], S[q + 1], n)
            f2, l2 = _splitting(oracle, S[q + 1], S[q], -n)
            assert bits_equal(f1, -f2) and l1 == l2


def test_reflecting_state_is_normal_sign_independent(oracle):
    rng = np.random.default_rng(2)
    S = _random_states(rng, 100)
    pt = np.zeros(3)
    for q in range(100):
        for d in range(3):
            n = np.zeros(3); n[d] = 1.0
            a, b = np.zeros(5), np.zeros(5)
            for out, nn in ((a, n), (b, -n)):
                nn = np.ascontiguousarray(nn)
                oracle.lib.orc_eval_interface_bc_values(PROBLEMS["radsod"], 1, pt.ctypes.data_as(_D),
                                                        nn.ctypes.data_as(_D), S[q].ctypes.data_as(_D),
                                                        out.ctypes.data_as(_D))
            assert bits_equal(a, b)


def test_threaded_baseline_is_bitwise_serial(oracle):
    _, h1 = oracle.bench_threads("vortex_xy", 3, 4, 0, 2, 1)
    _, h3 = oracle.bench_threads("vortex_xy", 3, 4, 0, 2, 3)
    _, h8 = oracle.bench_threads("vortex_xy", 3, 4, 0, 2, 8)
    assert h1 == h3 == h8


def test_threaded_checker_equals_serial_run(oracle):
    """orc_run_threads (the checker of the 128^3 / 256^3 GPU parity tests) == the serial main.cpp loop, bit for
    bit, on any thread count; its initial state is the problem's; the Morton -> lexicographic permutation is the
    mesh's own cell_ijk."""
    import oracle_lib
    ref = oracle.run("vortex_xy", 3, 32, t_end=1e30, max_steps=4, want_state=True)
    m = oracle.problem_mesh("vortex_xy", 3, 32)
    for threads in (1, 3, 8):
        U0, U = oracle.run_threads("vortex_xy", 3, 5, 4, threads)
        assert bits_equal(U, ref["U"]) and bits_equal(U0, oracle.init_state(m))
    ijk = m["cell_ijk"].astype(np.int64)
    assert np.array_equal(oracle_lib.morton_to_lexicographic(32), (ijk[:, 2] * 32 + ijk[:, 1]) * 32 + ijk[:, 0])


def test_body_boxes_make_wall_faces(oracle):
    boxes = np.array([[3.0, 3.0, 3.0, 5.0, 5.0, 5.0]])
    m = oracle.problem_mesh("radsod", 3, 16, boxes=boxes)
    assert 0 < int((m["fluid"] == 0).sum()) < m["fluid"].size
    inter = m["neigh"] >= 0
    mixed = inter & (m["fluid"][m["owner"]] != m["fluid"][np.maximum(m["neigh"], 0)])
    assert np.all(m["bc"][mixed] == 2) and mixed.sum() > 0
    assert np.all(m["bc"][~inter] == 1)
    assert np.all(m["bc"][inter & ~mixed] == -1)
    # solid cells keep a zero residual, wall faces contribute to their fluid side only
    U = oracle.init_state(m)
    RHS, me = oracle.compute_rhs(m, U)
    assert np.all(RHS[m["fluid"] == 0] == 0.0) and me > 0


def test_step_equals_manual_sequence(oracle):
    m = oracle.problem_mesh("vortex_xy", 3, 8)
    U = oracle.init_state(m)
    U1, W1, R1 = U.copy(), np.zeros_like(U), np.zeros_like(U)
    dt, me3 = oracle.step(m, 0.45, 0.0, 1.0, U1, W1, R1)
    U2, W2 = U.copy(), np.zeros_like(U)
    R, me = oracle.compute_rhs(m, U2)
    dt2 = oracle.choose_dt(0.45, m["h"], me, 0.0, 1.0)
    oracle.rk_stage(m, 1, dt2, U2, W2, R)
    R, _ = oracle.compute_rhs(m, W2)
    oracle.rk_stage(m, 2, dt2, U2, W2, R)
    R, _ = oracle.compute_rhs(m, W2)
    oracle.rk_stage(m, 3, dt2, U2, W2, R)
    assert dt == dt2 and me3[0] == me and bits_equal(U1, U2)


def test_lexicographic_generator_matches_morton_physics(oracle):
    """Same lattice, different numbering: RHS per lattice cell agrees to rounding (order differs)."""
    m = oracle.problem_mesh("radsod", 3, 8)
    lx = lexicographic_box_mesh(8, 8, 8, m["h"], 1)
    lx["problem"] = "radsod"
    U = oracle.init_state(m)
    perm = (m["cell_ijk"][:, 2] * 8 + m["cell_ijk"][:, 1]) * 8 + m["cell_ijk"][:, 0]
    Ul = np.empty_like(U); Ul[perm] = U
    R1, e1 = oracle.compute_rhs(m, U)
    R2, e2 = oracle.compute_rhs(lx, Ul)
    assert e1 == e2 and np.allclose(R2[perm], R1, rtol=0, atol=1e-13)


@pytest.mark.parametrize("dim", [2, 3])
def test_two_level_mesh_freestream_and_conservation(oracle, dim):
    """Hanging (2:1) faces: the face loop (src/euler.cpp:150-245) is connectivity-driven, so a coarse
    cell simply meets 2^(dim-1) interfaces on a refined side. Free stream stays free stream and, with
    reflecting borders, the residuals of mass and energy sum to rounding (every interior flux enters
    two cells with opposite signs; the mirror state gives exactly zero mass/energy flux)."""
    m = two_level_mesh(dim, 4, lambda i, j, k: (i in (1, 2) and j in (1, 2) and k in (0, 1, 2)) or (i, j, k) == (3, 3, 0))
    m["problem"] = "radsod"
    nc = m["volume"].shape[0]
    inter = m["neigh"] >= 0
    assert (m["size"][m["owner"][inter]] != m["size"][np.maximum(m["neigh"], 0)[inter]]).any()
    assert np.isclose(m["volume"].sum(), 4.0 ** dim)
    # closed cells: sum of A*n over the faces of every cell vanishes exactly (power-of-two sizes)
    acc = np.zeros((nc, 3))
    np.add.at(acc, m["owner"], m["area"][:, None] * m["normal"])
    np.add.at(acc, m["neigh"][inter], -(m["area"][:, None] * m["normal"])[inter])
    assert np.all(acc == 0.0)
    # free stream (free-flow borders)
    ff = dict(m, bc=np.where(inter, -1, 0).astype(np.int32))
    vel = (0.3, -0.2, 0.1 if dim == 3 else 0.0)
    U = np.tile([1.0, vel[0], vel[1], vel[2], 1.0 / 0.4 + 0.5 * sum(v * v for v in vel)], (nc, 1))
    R, me = oracle.compute_rhs(ff, U)
    assert np.abs(R).max() < 1e-13 and me > 0
    # conservation with reflecting borders
    rng = np.random.default_rng(5)
    U = _random_states(rng, nc)
    if dim == 2:
        U[:, 4] -= 0.5 * U[:, 3] ** 2 / U[:, 0]
        U[:, 3] = 0.0
    R, _ = oracle.compute_rhs(m, U)
    scale = np.abs(R).max()
    assert abs(R[:, 0].sum()) < 1e-12 * scale * nc and abs(R[:, 4].sum()) < 1e-12 * scale * nc
