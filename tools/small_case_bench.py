#!/usr/bin/env python
"""ms per RK3 step on the reference's own small cases (launch-bound on a GPU): the 3-D 32^3 vortex (uniform fused path)
and the 2-D 64^2 vortex (generic path), resident mode (mmf_run), CUDA-event timed by the library.  Development tool:
    MMF_STEP_GRAPH=0 python tools/small_case_bench.py      # one stream launch per kernel
    python tools/small_case_bench.py                       # the step replayed from a CUDA graph (default)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import minimmerflow_b200 as mmf  # noqa: E402
import oracle_lib  # noqa: E402  (mesh and initial state of the reference's cases only; nothing is checked or timed there)


def main():
    orc = oracle_lib.load()
    for name, dim, n in (("3dIsentropicVortex 32^3", 3, 32), ("2dIsentropicVortex 64^2", 2, 64), ("3d 64^3", 3, 64)):
        m = orc.problem_mesh("vortex_xy", dim, n)
        U = orc.init_state(m)
        with mmf.EulerSolver.from_mesh(m) as s:
            s.set_state(mmf.FIELD_U, U)
            s.run(0.45, m["h"], 0.0, 1e30, max_steps=20)
            l0 = s.info()["kernel_launches"]
            s.timer_start()
            s.run(0.45, m["h"], 0.0, 1e30, max_steps=400)
            ms = s.timer_stop()
            print(json.dumps({"case": name, "path": s.info()["path"], "graph": os.environ.get("MMF_STEP_GRAPH", "1"),
                              "us_per_step": 1e3 * ms / 400, "kernels_per_step": (s.info()["kernel_launches"] - l0) / 400}), flush=True)


if __name__ == "__main__":
    main()
