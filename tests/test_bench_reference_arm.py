"""`bench.py --impl reference` on the CPU: the contract of the reference arm (one JSON line, the keys the driver
reads, no GPU work) on a small sample.  The arm times the unmodified reference compiled into oracle/_ref on all
host cores (one serial replica per core, the image has no MPI) and, beside it, the C restatement on all threads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--cpu-size", "32", "--cpu-level", "5"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [ln for ln in out.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["higher_is_better"] is True
    assert d["metric"].startswith("cell-updates/sec per RK stage") and d["unit"] == "cell-updates/s" and d["dtype"] == "f64"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if cb["kind"] == "reference" and cb["cores"] > 1:
        # all-core figure from concurrent serial replicas; it cannot be far below one replica's rate
        assert cb["serial"]["cores"] == 1 and cb["value"] > 0.5 * cb["serial"]["value"]
    assert d["cpu_port_all_cores"]["kind"] == "port" and d["cpu_port_all_cores"]["value"] > 0


def test_other_ranks_of_the_reference_arm_do_nothing():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
