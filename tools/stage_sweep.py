#!/usr/bin/env python
"""Times the fused stage kernels (CUDA events around each launch, mmf_profile_*) for a set of
kernel variants selected through the MMF_STAGE_* environment knobs.  Development tool."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minimmerflow_b200 as mmf  # noqa: E402


def vortex(n):
    h = 10.0 / n
    x = (np.arange(n) + 0.5) * h - 5.0
    X, Y = np.meshgrid(x, x, indexing="xy")
    shape = 5.0 / (2 * np.pi) * np.exp(0.5 * (1 - (X * X + Y * Y)))
    T = 1.0 - 0.4 / 2.8 * shape * shape
    p = T ** 3.5
    r = p / T
    u, v = 1.0 - Y * shape, 1.0 + X * shape
    plane = np.stack([r, r * u, r * v, 0 * r, p / 0.4 + 0.5 * r * (u * u + v * v)], axis=-1)
    return np.broadcast_to(plane[None], (n, n, n, 5)).reshape(-1, 5).copy(), h


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--variants", default="p16:p16:r12:r12,p16:p16:d12:d12,p16:p16:h12:h12,p16:h16:h12:h12,p16:p16:w8:w8,p16:w8:w8:w8,w8,h12,h16,d12,d16,p12,r12,p16,r16,312")
    args = ap.parse_args()
    U, h = vortex(args.size)
    cells = args.size ** 3
    ref = None
    for var in args.variants.split(","):
        cfg, lz = var.split("@") if "@" in var else (var, "0")   # e.g. p16:p16:r12:r12@43
        os.environ["MMF_STAGE_CFG"] = cfg
        if int(lz) > 0:
            os.environ["MMF_STAGE_LZ"] = lz
        else:
            os.environ.pop("MMF_STAGE_LZ", None)
        with mmf.EulerSolver.uniform((args.size,) * 3, h, [0] * 6, cell_numbering=mmf.NUMBERING_LEXICOGRAPHIC) as s:
            s.set_state(mmf.FIELD_U, U)
            s.run(0.45, h, 0.0, 1e30, max_steps=3)
            s.timer_start()
            s.run(0.45, h, 0.0, 1e30, max_steps=args.steps)
            ms = s.timer_stop()
            s.profile_begin()
            s.run(0.45, h, 0.0, 1e30, max_steps=args.steps)
            kms, kn = s.profile_end()
            out = s.get_state(mmf.FIELD_U)
        if ref is None:
            ref = out
        same = bool(np.array_equal(out, ref))
        st = [kms[i] / max(kn[i], 1) for i in (1, 2, 3)]
        print(json.dumps({"variant": var, "ms_per_step": ms / args.steps, "stage_ms": st,
                          "cell_updates_per_s": cells * 3 * args.steps / (ms * 1e-3),
                          "stage_GBps": [b * cells / (t * 1e-3) / 1e9 for b, t in zip((80, 120, 120), st)],
                          "bitwise_equal_to_first": same}), flush=True)


if __name__ == "__main__":
    main()
