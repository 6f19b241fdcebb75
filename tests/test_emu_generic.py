"""The generic (connectivity-driven) path WITHOUT a GPU: its host logic and its kernels' source on the CPU.

* build_generic_tables (minimmerflow_b200/csrc/generic_tables.h) -- the per-cell interface lists mmf_create
  uploads -- against an independent construction from the host description;
* the generic kernels' source (generic_kernels.cuh: residual, RK stages, the fused stage kernel that
  MMF_GENERIC_FUSED=1 selects, the device-side dt choice), compiled by g++ against the SIMT shim of tools/emu and
  run in the kernel sequence of step_enqueue, bit for bit against the oracle on the reference's mesh kinds:
  2-D, bodies (BC_WALL with flipped normals), Dirichlet, 2:1 hanging faces, ghost-like cells that are solved but
  not internal, and the tMax clamp.

tools/emu is a development tool, never a product path; the GPU tests (tests/test_generic_gpu.py) remain the
parity proof of the compiled kernels."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "emu"))

import oracle_lib  # noqa: E402
from common import bits_equal, golden_cases, lexicographic_box_mesh, two_level_mesh  # noqa: E402

_D = C.POINTER(C.c_double)
NF = 5


@pytest.fixture(scope="module")
def emu():
    import run_emu
    return run_emu.load()


def _desc(m, dirichlet_info=None):
    from minimmerflow_b200.solver import mesh_desc
    return mesh_desc(m, flags=1, dirichlet_info=dirichlet_info)


def _soa(emu, aos):
    nc = aos.shape[0]
    stride = emu.emu_generic_stride(nc)
    out = np.zeros((NF, stride))
    out[:, :nc] = aos.T
    return out


def emu_run(emu, m, U, W, fused, cfl, t, t_max, max_steps, dirichlet_info=None):
    """Steps of the generic path on the emulator; returns (U, W, RHS) AoS, t, dt, the three eigenvalues, steps."""
    d, keep = _desc(m, dirichlet_info)
    nc = U.shape[0]
    Us, Ws, Rs = _soa(emu, U), _soa(emu, W), _soa(emu, np.zeros_like(U))
    tt = C.c_double(t)
    dt = C.c_double(0.0)
    eig = np.zeros(3)
    steps = emu.emu_generic_run(C.addressof(d), fused, Us.ctypes.data_as(_D), Ws.ctypes.data_as(_D), Rs.ctypes.data_as(_D),
                                cfl, float(m["size"].min()), C.byref(tt), t_max, max_steps, C.byref(dt),
                                eig.ctypes.data_as(_D))
    assert steps >= 0
    back = lambda S: np.ascontiguousarray(S[:, :nc].T)  # noqa: E731
    return back(Us), back(Ws), back(Rs), tt.value, dt.value, eig, steps


def random_state(nc, dim, seed):
    rng = np.random.default_rng(seed)
    rho = rng.uniform(0.5, 1.5, nc); vel = rng.uniform(-0.4, 0.4, (nc, 3)); p = rng.uniform(0.6, 1.4, nc)
    if dim == 2:
        vel[:, 2] = 0.0
    return np.column_stack([rho, rho * vel[:, 0], rho * vel[:, 1], rho * vel[:, 2], p / 0.4 + 0.5 * rho * (vel ** 2).sum(1)])


def oracle_steps(oracle, m, U, W, cfl, t, t_max, max_steps):
    Uo, Wo, Ro = U.copy(), W.copy(), np.zeros_like(U)
    dt, me3, steps = 0.0, np.zeros(3), 0
    while t < t_max and steps < max_steps:
        dt, me3 = oracle.step(m, cfl, t, t_max, Uo, Wo, Ro)
        t += dt
        steps += 1
    return Uo, Wo, Ro, t, dt, me3, steps


def check_against_oracle(emu, oracle, m, U, t_max=10.0, max_steps=3, dirichlet_info=None, solid=None, W=None):
    if W is None:
        W = np.full_like(U, 7.25)   # the host's W before the first step: cells no stage updates must keep it
    want = oracle_steps(oracle, m, U, W, 0.45, 0.0, t_max, max_steps)
    for fused in (0, 1):
        got = emu_run(emu, m, U, W, fused, 0.45, 0.0, t_max, max_steps, dirichlet_info)
        assert got[6] == want[6] and got[3] == want[3] and got[4] == want[4], (fused, got[3:], want[3:])
        assert list(got[5]) == list(want[5]), (fused, got[5], want[5])
        for name, a, b in zip("UWR", got[:3], want[:3]):
            assert bits_equal(a, b), (fused, name, int((a != b).any(axis=1).sum()))
        if solid is not None:
            assert np.all(got[2][solid] == 0.0) and bits_equal(got[0][solid], U[solid]) and np.all(got[1][solid] == 7.25)
    return want


def test_cell_interface_lists(emu, oracle):
    """Per solved cell: the interfaces that touch it in processing order, owner side 0 / neighbour side 1,
    interfaces without a solved side skipped (src/euler.cpp:181-183); update = solved AND internal."""
    meshes = [oracle.problem_mesh("radsod", 3, 8, boxes=[[2.1, 2.1, 2.1, 5.9, 4.9, 3.9]]),
              two_level_mesh(3, 3, lambda i, j, k: (i + j + k) % 3 == 0),
              oracle.problem_mesh("vortex_xy", 2, 8)]
    rng = np.random.default_rng(5)
    shuffled = dict(meshes[2])
    shuffled["interface_order"] = rng.permutation(shuffled["owner"].shape[0])[:-7].copy()   # a partial, permuted order
    shuffled["internal"] = (rng.uniform(size=shuffled["volume"].shape[0]) < 0.8).astype(np.uint8)
    meshes.append(shuffled)
    for m in meshes:
        d, keep = _desc(m)
        nc, nf = m["volume"].shape[0], m["owner"].shape[0]
        ptr = np.zeros(nc + 1, np.int64)
        ent = np.zeros(2 * nf, np.int32)
        upd = np.zeros(nc, np.uint8)
        err = C.create_string_buffer(256)
        n = emu.emu_generic_tables(C.addressof(d), ptr.ctypes.data_as(C.POINTER(C.c_longlong)),
                                   ent.ctypes.data_as(C.POINTER(C.c_int)), upd.ctypes.data, err, 256)
        assert n >= 0, err.value
        solved = np.asarray(m["solved"]).astype(bool)
        order = m.get("interface_order")
        order = np.arange(nf) if order is None else np.asarray(order)
        lists = [[] for _ in range(nc)]
        for f in order:
            o, nb = int(m["owner"][f]), int(m["neigh"][f])
            if not (solved[o] or (nb >= 0 and solved[nb])):
                continue
            if solved[o]:
                lists[o].append(2 * int(f))
            if nb >= 0 and solved[nb]:
                lists[nb].append(2 * int(f) + 1)
        assert n == sum(len(x) for x in lists) == ptr[nc]
        for c in range(nc):
            assert list(ent[ptr[c]:ptr[c + 1]]) == lists[c], c
        internal = m.get("internal")
        internal = np.ones(nc, bool) if internal is None else np.asarray(internal).astype(bool)
        assert np.array_equal(upd.astype(bool), solved & internal)


def test_bad_descriptions_are_rejected(emu, oracle):
    m = oracle.problem_mesh("vortex_xy", 2, 8)
    err = C.create_string_buffer(256)
    for key, idx, val, text in (("bc", 0, 7, b"unknown BC"), ("owner", 3, 10 ** 6, b"out of range")):
        bad = dict(m); bad[key] = m[key].copy(); bad[key][idx] = val
        d, keep = _desc(bad)
        assert emu.emu_generic_tables(C.addressof(d), None, None, None, err, 256) == -1 and text in err.value
    bad = dict(m); bad["bc"] = m["bc"].copy(); bad["bc"][np.nonzero(m["neigh"] < 0)[0][0]] = -1
    d, keep = _desc(bad)
    assert emu.emu_generic_tables(C.addressof(d), None, None, None, err, 256) == -1 and b"BC_NONE" in err.value


@pytest.mark.parametrize("problem,dim,n", [("vortex_xy", 2, 16), ("radsod", 2, 16), ("vortex_yz", 3, 8)])
def test_reference_problems_bit_exact(emu, oracle, problem, dim, n):
    m = oracle.problem_mesh(problem, dim, n)
    check_against_oracle(emu, oracle, m, oracle.init_state(m))


def test_tmax_clamp_and_loop_end(emu, oracle):
    """dt = tMax - t on the last step and `while (t < tMax)` (src/main.cpp:377, :401-402): both arms stop at
    the same step with t == tMax exactly."""
    m = oracle.problem_mesh("vortex_xy", 2, 16)
    U = oracle.init_state(m)
    want = check_against_oracle(emu, oracle, m, U, t_max=0.2, max_steps=50)
    assert want[3] == 0.2 and 1 < want[6] < 50


def test_bodies_wall_faces_bit_exact(emu, oracle):
    boxes = np.array([[2.9, 2.9, 2.9, 5.1, 5.1, 5.1], [0.0, 6.0, 0.0, 1.2, 8.0, 8.0]])
    m = oracle.problem_mesh("radsod", 3, 8, boxes=boxes)
    solid = m["fluid"] == 0
    inter = m["neigh"] >= 0
    assert (solid[m["owner"]] & inter & ~solid[np.maximum(m["neigh"], 0)]).any()   # the flipped-normal branch
    U = random_state(m["volume"].shape[0], 3, 0)
    check_against_oracle(emu, oracle, m, U, solid=solid)


def test_dirichlet_bit_exact(emu, oracle):
    m = lexicographic_box_mesh(8, 5, 3, 0.25, 1)
    m["problem"] = "ffstep"
    border = m["neigh"] < 0
    m["bc"][border & (m["normal"][:, 0] < 0)] = 3
    m["bc"][border & (m["normal"][:, 0] > 0)] = 0
    U = random_state(m["volume"].shape[0], 3, 3)
    check_against_oracle(emu, oracle, m, U, dirichlet_info=[1.0, 3.0, 0.0, 0.0, 1.0 / 1.4])


@pytest.mark.parametrize("dim", [2, 3])
def test_hanging_faces_bit_exact(emu, oracle, dim):
    m = two_level_mesh(dim, 4, lambda i, j, k: (1 <= i < 3 and j < 2 and k < 3) or (i + j + k) % 5 == 0)
    m["problem"] = "radsod"
    U = random_state(m["volume"].shape[0], dim, 11)
    check_against_oracle(emu, oracle, m, U)


def test_cells_solved_but_not_internal(emu, oracle):
    """Ghost-like cells: solved for the residual of their neighbours, never updated by the RK loops
    (src/main.cpp:231-235, :409-423).  The fused stage 2 must carry their W across the swap of the work arrays."""
    m = dict(oracle.problem_mesh("vortex_xy", 2, 16))
    nc = m["volume"].shape[0]
    internal = np.ones(nc, np.uint8)
    internal[np.random.default_rng(2).choice(nc, nc // 6, replace=False)] = 0
    m["internal"] = internal
    U = oracle.init_state(m)
    W = U * np.array([1.0, 0.9, 1.1, 1.0, 1.0])   # such cells enter the residual of stages 2 and 3 with the host's W
    want = check_against_oracle(emu, oracle, m, U, W=W)
    frozen = internal == 0
    assert bits_equal(want[0][frozen], U[frozen]) and bits_equal(want[1][frozen], W[frozen])
    assert not np.isnan(want[0]).any()


@pytest.mark.parametrize("case", [c for c in golden_cases() if c["dim"] == 2], ids=lambda c: c["name"])
def test_golden_strings_from_the_fused_sequence(emu, oracle, case):
    """The whole `while (t < tMax)` loop of the reference's two 2-D test cases through the fused generic kernel
    sequence: step count, end time and the printed 'Final error' digits (test/*/CMakeLists.txt:33).  (The three
    3-D cases reproduce their strings the same way; they take 20-40 s each on the emulator and are left to the GPU.)"""
    m = oracle.problem_mesh(case["problem"], case["dim"], case["n_cells"])
    t_end = case["t_end"] if case["t_end"] >= 0 else oracle.end_time(case["problem"], m["dim"])
    U = oracle.init_state(m)
    got = emu_run(emu, m, U, U.copy(), 1, case["cfl"], 0.0, t_end, -1)
    assert got[6] == case["steps"] and got[3] == t_end
    assert oracle_lib.format_error(oracle.error_norm(m, got[0], t_end)) == case["expected"]
