"""Pins the CPU oracle: it must reproduce the reference's five golden 'Final error' strings
(test/<case>/CMakeLists.txt:33) to all 13 printed digits, with the reference's own comparison rule
(string equality, test/test_driver.py:63)."""
import os
import subprocess

import pytest

import oracle_lib
from common import golden_cases


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_oracle_reproduces_reference_golden(oracle, case):
    res = oracle.run(case["problem"], case["dim"], case["n_cells"], t_end=case["t_end"], cfl=case["cfl"])
    assert oracle_lib.format_error(res["error"]) == case["expected"]
    assert res["steps"] == case["steps"]


def test_oracle_cli_prints_reference_line(oracle):
    """The CLI prints the same ' Final error:  <%.12e>' line the reference's test driver parses."""
    exe = os.path.join(oracle_lib.ORACLE_DIR, "build", "oracle_run")
    out = subprocess.check_output([exe, "radsod", "2", "64", "1"]).decode()
    line = [l for l in out.splitlines() if "Final error:" in l][0]
    assert line.split("Final error:")[1].strip() == "1.091296206337e+00"
