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


def run_one(var, size, steps):
    """One variant in THIS process: per-stage kernel times, whole-step time and a checksum of the final state."""
    import hashlib
    U, h = vortex(size)
    cells = size ** 3
    cfg, lz = var.split("@") if "@" in var else (var, "0")   # e.g. r16:r16:r12:r12@43
    os.environ["MMF_STAGE_CFG"] = cfg
    if int(lz) > 0:
        os.environ["MMF_STAGE_LZ"] = lz
    else:
        os.environ.pop("MMF_STAGE_LZ", None)
    # MMF_SWEEP_TOLERANT=1: experiment builds whose results are wrong on purpose (e.g. -DMMF_EXP_NOSYNC=1) trip the
    # library's internal eigenvalue check at the end of mmf_run; the launches have run and are timed all the same
    tolerant = os.environ.get("MMF_SWEEP_TOLERANT", "0") not in ("", "0")

    def run(sol, n):
        try:
            sol.run(0.45, h, 0.0, 1e30, max_steps=n)
        except mmf.MmfError:
            if not tolerant:
                raise
    with mmf.EulerSolver.uniform((size,) * 3, h, [0] * 6, cell_numbering=mmf.NUMBERING_LEXICOGRAPHIC) as s:
        s.set_state(mmf.FIELD_U, U)
        run(s, 3)
        s.timer_start()
        run(s, steps)
        ms = s.timer_stop()
        s.profile_begin()
        run(s, steps)
        kms, kn = s.profile_end()
        out = s.get_state(mmf.FIELD_U)
    st = [kms[i] / max(kn[i], 1) for i in (1, 2, 3)]
    return {"variant": var, "ms_per_step": ms / steps, "stage_ms": st,
            "cell_updates_per_s": cells * 3 * steps / (ms * 1e-3),
            "stage_GBps": [b * cells / (t * 1e-3) / 1e9 for b, t in zip((80, 120, 120), st)],
            "state_sha256": hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest()}


def summarise(rows):
    """What the sweep says: the fastest whole step, and per stage the fastest kernel among the bit-identical variants
    (candidate_mix = stage 0 : 1 : 2 : 3 shapes for MMF_STAGE_CFG; stage 0, the RHS-only launch, follows stage 1)."""
    good = [r for r in rows if r.get("bitwise_equal_to_first")]
    if not good:
        return None

    def stage_shape(r, st):   # the shape a variant runs stage st (1..3) with; one entry = all stages
        parts = r["variant"].split("@")[0].split(":")
        return parts[st] if len(parts) > st else parts[-1]
    best_step = min(good, key=lambda r: r["ms_per_step"])
    per_stage = [min(good, key=lambda r, i=i: r["stage_ms"][i]) for i in range(3)]
    mix = [stage_shape(per_stage[0], 1)] + [stage_shape(per_stage[i], i + 1) for i in range(3)]
    return {"summary": True, "fastest_step": best_step["variant"], "fastest_step_ms": best_step["ms_per_step"],
            "fastest_per_stage": [{"stage": i + 1, "variant": per_stage[i]["variant"], "shape": stage_shape(per_stage[i], i + 1),
                                   "ms": per_stage[i]["stage_ms"][i]} for i in range(3)],
            "candidate_mix": ":".join(mix),
            "not_bit_identical": [r["variant"] for r in rows if not r.get("bitwise_equal_to_first")]}


def main():
    import subprocess
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--variants", default="h12,h16,t16,t12,r12")
    ap.add_argument("--variant-timeout", type=float, default=90.0,
                    help="seconds per variant: every variant runs in its own process, so a kernel form that hangs on real "
                         "hardware costs its own slot only")
    ap.add_argument("--one", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.one is not None:
        print(json.dumps(run_one(args.one, args.size, args.steps)), flush=True)
        return
    rows, ref = [], None
    for var in args.variants.split(","):
        cmd = [sys.executable, os.path.abspath(__file__), "--size", str(args.size), "--steps", str(args.steps), "--one", var]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=args.variant_timeout)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            row = json.loads(lines[-1]) if r.returncode == 0 and lines else {"variant": var, "error": (r.stderr or "no output")[-400:]}
        except subprocess.TimeoutExpired:
            row = {"variant": var, "error": f"timeout after {args.variant_timeout:.0f} s"}
        if "error" not in row:
            if ref is None:
                ref = row["state_sha256"]
            row["bitwise_equal_to_first"] = row["state_sha256"] == ref
            rows.append(row)
        print(json.dumps(row), flush=True)
    summary = summarise(rows)
    if summary:
        print(json.dumps(summary), flush=True)


if __name__ == "__main__":
    main()
