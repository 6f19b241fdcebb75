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


def other_meshes(args):
    """Meshes only the generic path takes: the reference's 2-D vortex, or a two-level octree with hanging faces."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    if args.two_level:
        from common import two_level_mesh
        m = two_level_mesh(3, args.two_level, lambda i, j, k: (i + j + k) % 2 == 0, H=1.0, bc_code=0)
        nc = m["volume"].shape[0]
        rng = np.random.default_rng(7)
        rho = rng.uniform(0.9, 1.1, nc); vel = rng.uniform(-0.2, 0.2, (nc, 3)); p = rng.uniform(0.9, 1.1, nc)
        U = np.column_stack([rho, rho * vel[:, 0], rho * vel[:, 1], rho * vel[:, 2], p / 0.4 + 0.5 * rho * (vel ** 2).sum(1)])
        h, name = float(m["size"].min()), f"two-level octree, {args.two_level}^3 coarse cells, every other one refined"
    else:
        import oracle_lib   # (mesh and initial state of the reference's 2-D case; nothing is checked or timed there)
        orc = oracle_lib.load()
        m = orc.problem_mesh("vortex_xy", 2, args.size)
        U = orc.init_state(m)
        h, name = m["h"], f"2-D vortex_xy {args.size}^2"
    cells = m["volume"].shape[0]
    with mmf.EulerSolver.from_mesh(m) as s:
        s.set_state(mmf.FIELD_U, U)
        s.run(0.45, h, 0.0, 1e30, max_steps=3)
        s.timer_start()
        s.run(0.45, h, 0.0, 1e30, max_steps=args.steps)
        ms = s.timer_stop()
        s.profile_begin()
        s.run(0.45, h, 0.0, 1e30, max_steps=args.steps)
        kms, kn = s.profile_end()
        print(json.dumps({"mesh": name, "path_code": s.info()["path"], "generic_fused": os.environ.get("MMF_GENERIC_FUSED", "1"),
                          "cells": cells, "interfaces": int(m["owner"].shape[0]), "ms_per_step": ms / args.steps,
                          "cell_updates_per_s": cells * 3 * args.steps / (ms * 1e-3),
                          "kernel_ms": {"rhs": kms[0] / max(kn[0], 1), "stage2": kms[2] / max(kn[2], 1), "stage3": kms[3] / max(kn[3], 1)},
                          # bytes a fused generic step must move: stage 1 unfused (RHS: 40 in, 40 out; RK: 120 in, 40 out), stage 2
                          # 120, stage 3 160 + connectivity
                          "algorithmic_GBps": 520.0 * cells / (ms / args.steps * 1e-3) / 1e9}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--problem", default="vortex_xy")
    ap.add_argument("--bodies", action="store_true", help="two body boxes in the domain (the fused path then runs kernel form 'b')")
    ap.add_argument("--no-generic", action="store_true", help="time the fused path only")
    ap.add_argument("--generic-only", action="store_true", help="time the generic path only")
    ap.add_argument("--dim", type=int, default=3, help="2: the 2-D vortex on size^2 cells (generic path only: the fused path is 3-D)")
    ap.add_argument("--two-level", type=int, default=0, metavar="N0", help="a 2:1 two-level octree of N0^3 coarse cells, every "
                    "other coarse cell refined (hanging faces; generic path only); random admissible state")
    args = ap.parse_args()
    if args.problem != "vortex_xy":
        raise SystemExit("generic_bench.py builds the isentropic vortex (vortex_xy) only")
    # the benchmark domain (src/problem.cpp:70-149: origin -5, length 10), free-flow borders, lexicographic numbering
    origin, length = (-5.0, -5.0, -5.0), 10.0
    n = args.size
    if args.dim == 2 or args.two_level:
        return other_meshes(args)
    m = box_mesh(n, n, n, length / n, 0, origin=origin)
    if args.bodies:
        os.environ.setdefault("MMF_UNIFORM_BODIES", "1")
        lo = lambda f: [origin[e] + f[e] * length for e in range(3)]
        m = with_bodies(m, [lo((0.30, 0.35, 0.25)) + lo((0.45, 0.60, 0.55)), lo((0.70, 0.10, 0.60)) + lo((0.85, 0.30, 0.95))])
    U = vortex_state(m)
    cells = m["volume"].shape[0]
    out = {}
    for name, flags in (("uniform", 0), ("generic", mmf.FLAG_FORCE_GENERIC)):
        if (name == "generic" and args.no_generic) or (name == "uniform" and args.generic_only):
            continue
        with mmf.EulerSolver.from_mesh(m, flags=flags) as s:
            path = s.info()["path"]
            s.set_state(mmf.FIELD_U, U)
            s.run(0.45, m["h"], 0.0, 1e30, max_steps=3)
            s.timer_start()
            s.run(0.45, m["h"], 0.0, 1e30, max_steps=args.steps)
            ms = s.timer_stop()
            out[name] = s.get_state(mmf.FIELD_U)
            kernel_ms = None
            if name == "uniform":   # per-launch times of the stage kernels (events around each launch: no graph, no overlap)
                s.profile_begin()
                s.run(0.45, m["h"], 0.0, 1e30, max_steps=args.steps)
                kms, kn = s.profile_end()
                kernel_ms = [kms[q] / max(kn[q], 1) for q in (1, 2, 3)]
            print(json.dumps({"path": name, "stage_kernel_ms": kernel_ms, "path_code": path, "bodies_mode": os.environ.get("MMF_UNIFORM_BODIES", "") if args.bodies else "",
                              "generic_fused": os.environ.get("MMF_GENERIC_FUSED", "0"), "cells": cells, "solved_cells": int(m["solved"].sum()),
                              "ms_per_step": ms / args.steps,
                              "cell_updates_per_s": int(m["solved"].sum()) * 3 * args.steps / (ms * 1e-3)}), flush=True)
    if "generic" in out and "uniform" in out:
        print(json.dumps({"bitwise_equal": bool(np.array_equal(out["uniform"], out["generic"]))}))


if __name__ == "__main__":
    main()
