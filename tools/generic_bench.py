#!/usr/bin/env python
"""Times the GENERIC (connectivity-driven, any-mesh) path next to the uniform fused path on the same
problem, through the C-ABI: cell-updates/s per RK stage of K fused steps, CUDA-event timed by the library.
Development tool (the number that tells how far the generic path is from the fused one
before it is optimised).  Needs a GPU:

    python tools/generic_bench.py --size 128 --steps 6
    python tools/generic_bench.py --size 128 --steps 6 --bodies     # a box with bodies: form 'b' against the generic path
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import minimmerflow_b200 as mmf  # noqa: E402
from minimmerflow_b200.meshes import box_mesh, vortex_state, with_bodies  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--problem", default="vortex_xy")
    ap.add_argument("--bodies", action="store_true", help="two body boxes in the domain; the fused path then needs MMF_UNIFORM_BODIES=1")
    args = ap.parse_args()
    if args.problem != "vortex_xy":
        raise SystemExit("generic_bench.py builds the isentropic vortex (vortex_xy) only")
    # the benchmark domain (src/problem.cpp:70-149: origin -5, length 10), free-flow borders, lexicographic numbering
    origin, length = (-5.0, -5.0, -5.0), 10.0
    n = args.size
    m = box_mesh(n, n, n, length / n, 0, origin=origin)
    if args.bodies:
        os.environ.setdefault("MMF_UNIFORM_BODIES", "1")   # 1 = form b, 2 = form c
        lo = lambda f: [origin[e] + f[e] * length for e in range(3)]
        m = with_bodies(m, [lo((0.30, 0.35, 0.25)) + lo((0.45, 0.60, 0.55)), lo((0.70, 0.10, 0.60)) + lo((0.85, 0.30, 0.95))])
    U = vortex_state(m)
    cells = m["volume"].shape[0]
    out = {}
    for name, flags in (("uniform", 0), ("generic", mmf.FLAG_FORCE_GENERIC)):
        with mmf.EulerSolver.from_mesh(m, flags=flags) as s:
            path = s.info()["path"]
            s.set_state(mmf.FIELD_U, U)
            s.run(0.45, m["h"], 0.0, 1e30, max_steps=3)
            s.timer_start()
            s.run(0.45, m["h"], 0.0, 1e30, max_steps=args.steps)
            ms = s.timer_stop()
            out[name] = s.get_state(mmf.FIELD_U)
            print(json.dumps({"path": name, "path_code": path, "bodies_mode": os.environ.get("MMF_UNIFORM_BODIES", "") if args.bodies else "",
                              "generic_fused": os.environ.get("MMF_GENERIC_FUSED", "0"), "cells": cells, "solved_cells": int(m["solved"].sum()),
                              "ms_per_step": ms / args.steps,
                              "cell_updates_per_s": int(m["solved"].sum()) * 3 * args.steps / (ms * 1e-3)}), flush=True)
    print(json.dumps({"bitwise_equal": bool(np.array_equal(out["uniform"], out["generic"]))}))


if __name__ == "__main__":
    main()
