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
sys.path.insert(0, os.path.join(ROOT, "tests"))
import minimmerflow_b200 as mmf  # noqa: E402
import oracle_lib  # noqa: E402  (mesh generator only: the oracle computes nothing here)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--problem", default="vortex_xy")
    ap.add_argument("--bodies", action="store_true", help="two body boxes in the domain; the fused path then needs MMF_UNIFORM_BODIES=1")
    args = ap.parse_args()
    orc = oracle_lib.load()
    boxes = None
    if args.bodies:
        os.environ.setdefault("MMF_UNIFORM_BODIES", "1")   # 1 = form b, 2 = form c
        _, origin, length = orc.domain(args.problem, 3)
        lo = lambda f: [origin[e] + f[e] * length for e in range(3)]
        boxes = [lo((0.30, 0.35, 0.25)) + lo((0.45, 0.60, 0.55)), lo((0.70, 0.10, 0.60)) + lo((0.85, 0.30, 0.95))]
    m = orc.problem_mesh(args.problem, 3, args.size, boxes=boxes)
    U = orc.init_state(m)
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
