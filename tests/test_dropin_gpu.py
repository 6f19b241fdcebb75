"""The drop-in boundary, end to end, in its two modes:

* strict   -- the reference's UNMODIFIED main.cpp / problem.cpp / body.cpp / mesh_info.cpp /
              solver_writer.cpp linked with minimmerflow_b200/adapters/solver_b200.cpp (which defines
              euler::computeRHS and reconstruction::computePolynomials on top of the C-ABI) instead of
              euler.cpp / reconstruction.cpp;
* resident -- the same set-up / output units with minimmerflow_b200/adapters/driver_b200.cpp as
              main(): the time loop is a stream of mmf_step calls, the state stays on the device.

Built in the dev container (oracle/Makefile `dropin`), run here on the GPU.  Both executables must
print the reference's golden strings and write the same final fields as the reference did
(tests/golden/reference_fields.npz) -- bitwise."""
import os

import numpy as np
import pytest

import reference_runner as R
from common import bits_equal, golden_cases, reference_cases, reference_fields

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (os.path.exists(R.DROPIN_EXE) and os.path.exists(R.RESIDENT_EXE)),
                                 reason="oracle/_ref drop-in executables not built (make -C oracle dropin)")]


EXES = [pytest.param(R.DROPIN_EXE, id="strict"), pytest.param(R.RESIDENT_EXE, id="resident")]


@pytest.mark.parametrize("exe", EXES)
@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_dropin_prints_reference_golden_strings(exe, case):
    r = R.run_case(exe, case)
    assert r["final_error"] == case["expected"]          # test/test_driver.py:63: string equality
    assert r["steps"] == case["steps"]


@pytest.mark.parametrize("exe", EXES)
@pytest.mark.parametrize("case", reference_cases(), ids=lambda c: c["name"])
def test_dropin_fields_equal_reference_fields(exe, case):
    ref = reference_fields()
    n = case["name"]
    r = R.run_case(exe, case, want_fields=True)
    f = r["fields"]
    assert r["final_error"] == str(ref[n + "/final_error"]) and r["steps"] == int(ref[n + "/steps"])
    assert np.array_equal(f["solved"], ref[n + "/solved"])
    assert bits_equal(f["density"], ref[n + "/density"])
    assert bits_equal(f["velocity"], ref[n + "/velocity"])
    assert bits_equal(f["pressure"], ref[n + "/pressure"])
    assert bits_equal(f["temperature"], ref[n + "/temperature"])
    res = np.stack([f["residualC"], f["residualMX"], f["residualMY"], f["residualMZ"], f["residualE"]], axis=1)
    assert bits_equal(res, ref[n + "/residual"])


@pytest.mark.parametrize("exe", EXES)
def test_dropin_unsupported_order_exits_like_the_reference(exe):
    """order != 1: reconstruction::eval calls exit(2) (src/reconstruction.cpp:76); so does the adapter."""
    import subprocess
    import tempfile
    case = dict(golden_cases()[4], order=2)
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "settings.xml"), "w").write(R.settings_xml(case))
        env = dict(os.environ, BITPIT_SHIM_VTK="0")
        out = subprocess.run([exe, "8"], cwd=tmp, env=env, capture_output=True, timeout=300)
    assert out.returncode == 2
