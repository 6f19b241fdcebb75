"""Pins the CPU oracle against the REFERENCE ITSELF at field level.

tests/golden/reference_fields.npz holds the final cell fields the unmodified reference
(/root/reference/src/*.cpp compiled against compat/bitpit, oracle/_ref/minimmerflow_ref) wrote for
the small cases of reference_cases.json -- including the paths its own five regression strings never
touch: bodies / BC_WALL with flipped normals, BC_DIRICHLET (ffstep), sod3d, vortex_zx.  The oracle
must reproduce every stored field BITWISE (the sign of an exact zero aside), the printed error string
and the step count."""
import os

import numpy as np
import pytest

import oracle_lib
import reference_runner as R
from common import bits_equal, case_mesh, golden_cases, primitives, reference_cases, reference_fields


def run_oracle(oracle, case):
    m = case_mesh(oracle, case)
    U = oracle.init_state(m)
    W, RHS = np.zeros_like(U), np.zeros_like(U)
    t, steps = 0.0, 0
    while t < case["t_end"]:
        dt, _ = oracle.step(m, case["cfl"], t, case["t_end"], U, W, RHS)
        t += dt
        steps += 1
    return m, U, RHS, steps


@pytest.mark.parametrize("case", reference_cases(), ids=lambda c: c["name"])
def test_oracle_matches_reference_fields_bitwise(oracle, case):
    ref = reference_fields()
    n = case["name"]
    m, U, RHS, steps = run_oracle(oracle, case)
    assert steps == int(ref[n + "/steps"])
    assert oracle_lib.format_error(oracle.error_norm(m, U, case["t_end"])) == str(ref[n + "/final_error"])
    P = primitives(oracle, U)
    assert np.array_equal(ref[n + "/solved"], m["solved"].astype(np.int32))
    assert bits_equal(ref[n + "/density"], U[:, 0])
    assert bits_equal(ref[n + "/velocity"], P[:, 1:4])
    assert bits_equal(ref[n + "/pressure"], P[:, 0])
    assert bits_equal(ref[n + "/temperature"], P[:, 4])
    assert bits_equal(ref[n + "/residual"], RHS)          # residual of the last RK stage (cellRHS)


@pytest.mark.skipif(not os.path.exists(R.REF_EXE), reason="oracle/_ref/minimmerflow_ref not built (make -C oracle ref)")
@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_compiled_reference_reproduces_its_golden_strings(case):
    """The unmodified reference sources on the bitpit stand-in print the strings their own CMake
    tests expect (test/<case>/CMakeLists.txt:33): validates compat/bitpit, not the oracle."""
    r = R.run_case(R.REF_EXE, case)
    assert r["final_error"] == case["expected"]
    assert r["steps"] == case["steps"]


@pytest.mark.skipif(not os.path.exists(R.REF_EXE), reason="oracle/_ref/minimmerflow_ref not built (make -C oracle ref)")
def test_fixture_is_current(oracle):
    """Re-running the compiled reference reproduces the committed fixture (one case, fields)."""
    case = [c for c in reference_cases() if c["name"] == "radsod_3d_16_body"][0]
    ref = reference_fields()
    r = R.run_case(R.REF_EXE, case, want_fields=True)
    assert r["final_error"] == str(ref[case["name"] + "/final_error"])
    assert np.array_equal(r["fields"]["density"], ref[case["name"] + "/density"])


@pytest.mark.skipif(not os.path.exists(R.REF_EXE), reason="oracle/_ref/minimmerflow_ref not built (make -C oracle ref)")
@pytest.mark.parametrize("dim,n", [(2, 8), (3, 4)])
def test_standin_vtu_is_wellformed(dim, n):
    """The .vtu the bitpit stand-in writes for the reference's SolverWriter: every cell field has one
    value (three for velocity) per cell, the connectivity indexes existing points, cells are VTK pixels /
    voxels of size h, and the initial density of the radial Sod problem takes its two states."""
    case = dict(problem="radsod", dim=dim, n_cells=n, cfl=0.45, t_end=0.0)
    r = R.run_case(R.REF_EXE, case, want_fields=True)
    f = r["fields"]
    nc = n ** dim
    assert f["_n_cells"] == nc and r["steps"] == 0
    for name in ("solved", "pressure", "temperature", "density", "residualC", "residualE"):
        assert f[name].shape == (nc,)
    assert f["velocity"].shape == (nc, 3) and f["Points"].shape == ((n + 1) ** dim, 3)
    vpc = 2 ** dim
    conn = f["connectivity"].reshape(nc, vpc)
    assert conn.min() == 0 and conn.max() == (n + 1) ** dim - 1
    assert np.array_equal(f["offsets"], vpc * np.arange(1, nc + 1))
    assert set(f["types"]) == {11 if dim == 3 else 8}
    h = 8.0 / n
    corners = f["Points"][conn]                                   # [cell, vertex, xyz]
    assert np.allclose(corners.max(1)[:, :dim] - corners.min(1)[:, :dim], h)
    assert set(np.round(f["density"], 12)) <= {1.0, 0.125} and f["solved"].min() == 1
