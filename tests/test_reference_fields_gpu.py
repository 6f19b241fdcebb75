"""GPU parity against the REFERENCE's own final fields (tests/golden/reference_fields.npz, written by
the unmodified reference sources): the device-resident time loop (mmf_run, what replaces
src/main.cpp:377-506) must end on bitwise the same state, after the same number of steps, on every
case -- 2-D (generic path) and the 3-D boxes, plain, with bodies / BC_WALL and with BC_DIRICHLET (fused
uniform path)."""
import os

import numpy as np
import pytest

import oracle_lib
from common import bits_equal, case_mesh, dirichlet_info, primitives, reference_cases, reference_fields

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", reference_cases(), ids=lambda c: c["name"])
def test_resident_run_matches_reference_fields(mmf, oracle, case):
    ref = reference_fields()
    n = case["name"]
    m = case_mesh(oracle, case)
    with mmf.EulerSolver.from_mesh(m, dirichlet_info=dirichlet_info(case)) as s:
        # every 3-D box takes the fused path, with bodies (kernel form 'b') or without, unless MMF_UNIFORM_BODIES=0
        bodies_fused = os.environ.get("MMF_UNIFORM_BODIES", "1") not in ("", "0")
        fused = m["dim"] == 3 and (bodies_fused or not case.get("bodies"))
        assert s.info()["path"] == (mmf.PATH_UNIFORM if fused else mmf.PATH_GENERIC)
        s.set_state(mmf.FIELD_U, oracle.init_state(m))
        t, steps = s.run(case["cfl"], float(m["size"].min()), 0.0, case["t_end"])
        U = s.get_state(mmf.FIELD_U)
    assert steps == int(ref[n + "/steps"]) and t == case["t_end"]
    assert oracle_lib.format_error(oracle.error_norm(m, U, case["t_end"])) == str(ref[n + "/final_error"])
    P = primitives(oracle, U)
    assert bits_equal(ref[n + "/density"], U[:, 0])
    assert bits_equal(ref[n + "/velocity"], P[:, 1:4])
    assert bits_equal(ref[n + "/pressure"], P[:, 0])
    assert bits_equal(ref[n + "/temperature"], P[:, 4])
