#!/usr/bin/env python
"""bench.py -- cell-updates/s per RK stage (FP64) of the explicit FV Euler residual-and-update path
on a synthetic uniform 3-D isentropic vortex (BASELINE.json metric), on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl ours|reference]

One "step" = one SSP-RK3 step = 3 RK stages over every cell of the mesh.  value = cells * 3 * K /
time, the WHOLE-JOB aggregate.  Inputs are resident in HBM when the timed region starts; the state
arrays (3 x 5 fields x 8 B x cells, 2 GB at 256^3) are far larger than the 126 MB L2, so no extra L2
flush is needed between timed steps.  `e2e` is the same metric through the C-ABI with HOST buffers:
upload (pinned AoS) + K steps + download inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cell-updates/sec per RK stage (FP64)"
UNIT = "cell-updates/s"
ALG_BYTES_PER_CELL_STAGE = (80.0, 120.0, 120.0)  # SURVEY 8d / DESIGN.md: stage 1, 2, 3


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--size", type=int, default=256, help="cells per side: per GPU (weak scaling) or of the whole cube (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: one size^3 box per GPU (default, the contract's line); strong: one size^3 cube cut "
                         "into the N contiguous Morton chunks (SURVEY 8e: 512^3 -> 512x512x256 / 512x256x256 / 256^3)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-level", type=int, default=0, help="oracle-port mesh level for the CPU legs (0 = 7)")
    ap.add_argument("--cpu-size", type=int, default=0, help="cells per side for the compiled-reference CPU leg (0 = 128)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strict", action="store_true", help="skip the timing of the strict drop-in call (mmf_compute_rhs_host)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check that precedes the timed region")
    ap.add_argument("--no-bodies", action="store_true", help="skip the secondary measurement of a box with bodies (N = 1)")
    ap.add_argument("--bodies-size", type=int, default=192, help="cells per side of the secondary box with bodies")
    ap.add_argument("--repeats", type=int, default=5,
                    help="the timed region (exactly --steps steps between barriers) is run this many times; the line reports "
                         "the MEDIAN repeat and lists all of them (SURVEY 8d)")
    ap.add_argument("--nccl-halo", action="store_true", help="exchange halos with NCCL send/recv instead of direct peer stores")
    return ap.parse_args()


# ---- synthetic input: isentropic vortex of src/problem.cpp:252-327 at cell centroids ------------
def vortex_state(dims, offset, global_n, length=10.0, origin=-5.0):
    """Lexicographic (x fastest) AoS state of the local box; numpy, host side."""
    h = length / global_n
    x = origin + (np.arange(dims[0]) + offset[0] + 0.5) * h
    y = origin + (np.arange(dims[1]) + offset[1] + 0.5) * h
    X, Y = np.meshgrid(x, y, indexing="xy")            # [ny, nx]
    r2 = X * X + Y * Y
    shape = 5.0 / (2 * np.pi) * np.exp(0.5 * (1 - r2))
    dT = -(1.4 - 1.0) / (2 * 1.4) * shape * shape
    T = 1.0 + dT
    p = 1.0 + (np.power(T, 1.4 / (1.4 - 1.0)) - 1.0)
    u, v = 1.0 - Y * shape, 1.0 + X * shape
    rho = p / T
    plane = np.stack([rho, rho * u, rho * v, 0.0 * rho, rho * T / 0.4 + 0.5 * rho * (u * u + v * v)], axis=-1)
    return plane, h


def box_decomposition(n_ranks):
    """Process grid (px,py,pz) for N = 1,2,4,8 in the order a Morton partition of a cube splits:
    z first, then y, then x (x is the lowest Morton bit, src/main.cpp:159,183 + PABLO)."""
    grid = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (1, 2, 4)}[n_ranks]
    override = os.environ.get("MMF_BENCH_GRID")  # development: e.g. "2x1x1" to put the partition side on x
    if override:
        g = tuple(int(v) for v in override.split("x"))
        if len(g) == 3 and g[0] * g[1] * g[2] == n_ranks:
            grid = g
    return grid


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.maxes = [], set(), []
        self.gpu = gpu_index
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.maxes.append(float(out[1]))
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(max(self.maxes)),
                "reasons": sorted(self.reasons)}


def stage_kernel_label():
    """Name of the kernel that runs stages 2 and 3 (the dominant one), from the same knob the library reads."""
    names = {"r": "uniform_stage_kernel_v5r", "t": "uniform_stage_kernel_t (input staged by bulk tensor loads)",
             "h": "uniform_stage_kernel_t (input staged by bulk tensor loads, one warp for both halo rows)"}
    shapes = ["h12"] * 4                                       # library defaults (uniform_path.cuh)
    cfg = [c for c in os.environ.get("MMF_STAGE_CFG", "").split(":") if c]
    if len(cfg) == 1:
        shapes = cfg * 4
    else:
        shapes[:len(cfg)] = cfg[:4]
    s2, s3 = shapes[2], shapes[3]
    label = f"{names.get(s2[0], s2[0])}<2>, {s2[1:] or '12'} warps"
    if s3 != s2:
        label += f" / {names.get(s3[0], s3[0])}<3>, {s3[1:] or '12'} warps"
    else:
        label = label.replace("<2>", "<2|3>")
    return label + " (stages 2 and 3: two of the three launches of an RK3 step)"


# FP64 instructions per cell and stage in the stage kernels' SASS (profiles/r02d_experiments.md: 224 of 526 instructions
# per updated plane of a warp at stage 2; stage 1 has the shorter RK update) and the measured FP64 issue rate
FP64_INSTR_PER_CELL_STAGE = (214, 224, 224)
FP64_LANES_PER_CLK_SM = 61.2


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---- CPU legs (oracle/ and oracle/_ref, timed; never on the product path) -------------------------
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "minimmerflow_ref")


def cpu_port_run(level, warmup, steps, threads):
    """The oracle port (oracle/mmf_oracle.c) on `threads` host threads, contiguous Morton chunks."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    orc = oracle_lib.load()
    secs, _ = orc.bench_threads("vortex_xy", 3, level, warmup, steps, threads)
    cells = (1 << level) ** 3
    return cells * 3.0 * steps / secs, secs, cells


def cpu_reference_run(n_side, steps):
    """The UNMODIFIED reference (oracle/_ref/minimmerflow_ref: /root/reference/src/*.cpp compiled in the
    dev container against compat/bitpit; serial -- the reference's only parallelism is MPI, which this
    image does not have) on the benchmark problem at n_side^3 cells for `steps` RK3 steps.  The time is
    the reference's own "Computation time (without disk saving time)" line (src/main.cpp:371-372, 525,
    542-544: clock() around the time loop), so mesh set-up and output are excluded like on the GPU."""
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import reference_runner as R
    h = 10.0 / n_side
    # dt = 0.9*CFL*h/maxEig with maxEig ~ 2.94 for this vortex: end time placed mid-way through step `steps`
    t_end = (steps - 0.5) * 0.9 * 0.45 * h / 2.94
    case = dict(problem="vortex_xy", dim=3, n_cells=n_side, cfl=0.45, t_end=t_end)
    r = R.run_case(REF_EXE, case, timeout=1800)
    line = [ln for ln in r["output"].splitlines() if "Computation time" in ln][0]
    secs = float(line.split(" is ")[1])
    cells = n_side ** 3
    return cells * 3.0 * r["steps"] / secs, secs, cells, r["steps"]


def cpu_reference_replicas(n_side, steps, procs):
    """`procs` concurrent processes of the unmodified serial reference, one per host core, each on its own
    n_side^3 mesh: what the reference's MPI decomposition would give with free communication (the image has no
    MPI, so the ranks cannot talk; every replica does the work of one rank of a weak-scaled run).  Aggregate
    rate = sum over the replicas of their own loop rate while all of them run; also the slowest loop clock."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=procs) as pool:
        runs = list(pool.map(lambda _: cpu_reference_run(n_side, steps), range(procs)))
    return sum(r[0] for r in runs), max(r[1] for r in runs), runs[0][2], runs[0][3]


def cpu_baseline_entry(args, cores):
    """Bounded CPU sample of the same workload for the `cpu_baseline` object (and the reference arm)."""
    if os.path.exists(REF_EXE):
        n_side = args.cpu_size or 128
        v1, secs1, cells, steps = cpu_reference_run(n_side, 8)
        serial = {"value": v1, "unit": UNIT, "cores": 1, "kind": "reference",
                  "sample": f"{steps} RK3 steps of the same problem on a {n_side}^3 mesh by the unmodified reference "
                            f"(serial build: no MPI in the image), {secs1:.1f} s of its own loop clock"}
        # all the host threads the reference can use without MPI: one serial replica per core (at most 64, on
        # 64^3 meshes, to bound memory and time)
        procs = max(1, min(cores, 64))
        if procs > 1:
            v, secs, cells, steps = cpu_reference_replicas(64, 8, procs)
            entry = {"value": v, "unit": UNIT, "cores": procs, "kind": "reference",
                     "sample": f"{procs} concurrent processes of the unmodified serial reference (no MPI in the image: "
                               f"one replica per core stands in for one rank, communication free), each {steps} RK3 steps "
                               f"of the same problem on a 64^3 mesh; sum of the replicas' own loop rates, slowest loop {secs:.1f} s",
                     "serial": serial}
            timing = (secs, steps)
        else:
            entry, timing = serial, (secs1, steps)
    else:
        entry = None
    level = args.cpu_level or 7
    pv, psecs, _ = cpu_port_run(level, 1, 20, cores)
    port = {"value": pv, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"20 RK3 steps on a {1 << level}^3 mesh by the oracle port on {cores} threads "
                      f"(contiguous Morton chunks, the reference's MPI decomposition), {psecs:.1f} s"}
    if entry is None:
        return port, None, (psecs, 20)
    return entry, port, timing


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    entry, port, (secs, steps) = cpu_baseline_entry(args, cores)
    value = entry["value"]
    px, py, pz = box_decomposition(args.gpus)
    dims, gdims = local_dims(args, (px, py, pz))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": 0 if entry["kind"] == "reference" else 1,
        "ms_per_step": 1e3 * secs / steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(dims, gdims, (px, py, pz)),
                       sample_run=entry["sample"],
                       sample_note="the workload / decomposition keys name the GPU arm's configuration (the driver pairs the two "
                                   "lines by them); what THIS line timed is sample_run: a bounded CPU sample of the same problem"),
        "cpu_baseline": entry,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU arm: the reference's own implementation of the path (src/euler.cpp, src/main.cpp time loop, unmodified) on "
                "the host cores, on a bounded sample of the workload (config.sample_run).  It is built against compat/bitpit, "
                "this repository's stand-in for the bitpit containers and octree (bitpit itself is not available offline): "
                "id -> position look-ups and iteration cost whatever the stand-in makes them cost, and the replicas do not "
                "communicate -- a stated baseline, not a tuned competitor",
    }
    if port is not None:
        line["cpu_port_all_cores"] = port
    print(json.dumps(line), flush=True)


def local_dims(args, grid):
    """Cells per axis of one rank's box, and of the whole mesh."""
    S = args.size
    if args.scaling == "strong":
        if any(S % g for g in grid):
            raise SystemExit(f"--size {S} is not divisible by the process grid {grid}")
        return (S // grid[0], S // grid[1], S // grid[2]), (S, S, S)
    return (S, S, S), (S * grid[0], S * grid[1], S * grid[2])


def workload_config(dims, gdims, grid):
    return {"workload": f"3-D isentropic vortex (vortex_xy), uniform {gdims[0]}x{gdims[1]}x{gdims[2]} cells "
                        f"({dims[0]}x{dims[1]}x{dims[2]} per GPU), order 1, CFL 0.45, free-flow borders, fixed step count",
            "decomposition": f"{grid[0]}x{grid[1]}x{grid[2]} boxes (halo: peer copies over NVLink by the copy engines next to the following "
                             f"stage, boundary tiles wait in-kernel and run last; scalar all-reduces: NCCL)",
            "path": "uniform fused stage kernels",
            "l2": "state arrays (2.0 GB at 256^3) >> 126 MB L2, no flush needed"}


# ---- parity inside the bench run (the oracle as the checker, never as the thing measured) ----------
PARITY_STEPS = 3


def ulp_distance(a, b):
    """Largest distance in units in the last place between two float64 arrays (+0 == -0)."""
    a = np.ascontiguousarray(a, dtype=np.float64).ravel()
    b = np.ascontiguousarray(b, dtype=np.float64).ravel()
    bad = a != b
    if not bad.any():
        return 0, 0
    ia, ib = a[bad].view(np.int64), b[bad].view(np.int64)
    # map the sign-magnitude bit patterns onto a monotone integer line
    ia = np.where(ia < 0, np.int64(-2 ** 63) - ia, ia)
    ib = np.where(ib < 0, np.int64(-2 ** 63) - ib, ib)
    with np.errstate(over="ignore"):
        d = np.abs(ia - ib)                                  # int64; wraps only for values of opposite huge magnitude
    far = (d < 0) | (np.sign(ia) * np.sign(ib) < 0) & (np.abs(ia.astype(np.float64) - ib.astype(np.float64)) > 2.0 ** 62)
    d[far] = np.iinfo(np.int64).max
    return int(d.max()), int(bad.sum())


def parity_check(args, make_solver, dist, rank, world, grid, coords):
    """Before anything is timed: PARITY_STEPS RK3 steps of the benchmark problem on a size^3 cube cut by the SAME
    process grid, kernels, numbering conventions and halo exchange as the timed run, every rank's box compared
    bit for bit with the CPU oracle (oracle/mmf_oracle.c, threaded loop; rank 0 runs it and broadcasts).  At N = 1
    this is the timed configuration itself.  Exits non-zero on any difference."""
    S = args.size
    if args.no_parity or S & (S - 1) or any(S % g for g in grid):
        return {"checked": False, "why": "disabled (--no-parity)" if args.no_parity else
                f"the oracle's octree mesh needs a power-of-two cube divisible by the process grid, size is {S}"}
    import minimmerflow_b200 as mmf
    t0 = time.perf_counter()
    level = S.bit_length() - 1
    n = S ** 3
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        orc = oracle_lib.load()
        U0m, Urm = orc.run_threads("vortex_xy", 3, level, PARITY_STEPS)
        perm = oracle_lib.morton_to_lexicographic(S)
        both = np.empty((2, n, 5))
        both[0][perm] = U0m
        both[1][perm] = Urm
        del U0m, Urm, perm
    else:
        both = np.empty((2, n, 5))
    if world > 1:
        import torch
        for q in range(2):   # one array at a time: 671 MB each at 256^3
            tt = torch.from_numpy(both[q]).cuda()
            dist.broadcast(tt, src=0)
            both[q] = tt.cpu().numpy()
            del tt
        torch.cuda.empty_cache()
    box = (S // grid[0], S // grid[1], S // grid[2])
    off = (coords[0] * box[0], coords[1] * box[1], coords[2] * box[2])
    sl = (slice(off[2], off[2] + box[2]), slice(off[1], off[1] + box[1]), slice(off[0], off[0] + box[0]))
    G = both.reshape(2, S, S, S, 5)
    mine0 = np.ascontiguousarray(G[0][sl]).reshape(-1, 5)
    mine1 = np.ascontiguousarray(G[1][sl]).reshape(-1, 5)
    del both, G
    h = 10.0 / S
    with make_solver(box, (S, S, S), off, h) as sol:
        sol.set_state(mmf.FIELD_U, mine0)
        _, done = sol.run(0.45, h, 0.0, 1.0e30, max_steps=PARITY_STEPS)
        got = sol.get_state(mmf.FIELD_U)
    ulp, n_bad = ulp_distance(got, mine1)
    if world > 1:
        import torch
        tt = torch.tensor([float(ulp), float(n_bad)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ulp, n_bad = int(tt[0]), int(tt[1])
    out = {"checked": True, "max_ulp": ulp, "values_differing": n_bad, "steps": PARITY_STEPS,
           "what": f"{S}^3 vortex_xy cube cut into {grid[0]}x{grid[1]}x{grid[2]} boxes of {box[0]}x{box[1]}x{box[2]}, "
                   f"{PARITY_STEPS} RK3 steps, every rank's box against the CPU oracle (threaded loop, bitwise the serial "
                   f"one), same kernels / numbering / halo exchange as the timed run",
           "seconds": round(time.perf_counter() - t0, 1)}
    if ulp != 0 or done != PARITY_STEPS:
        if rank == 0:
            print(json.dumps({"parity": out, "error": "GPU result differs from the oracle"}), flush=True)
        sys.exit(3)
    return out


# ---- our arm --------------------------------------------------------------------------------------
def bodies_secondary(args, mmf, device):
    """The reference's own use case next to the headline: a uniform box with two body boxes inside (cells not solved,
    BC_WALL around them, src/main.cpp:221-277), described interface by interface through mmf_create like the adapter
    does, which picks the fused path with kernel form 'b' by itself; and the same box without bodies.  Device-timed
    (CUDA events on the library's stream), --steps steps after 3, one GPU."""
    from minimmerflow_b200.meshes import box_mesh, vortex_state as mesh_vortex_state, with_bodies
    n, length, origin = args.bodies_size, 10.0, (-5.0, -5.0, -5.0)
    lo = lambda f: [origin[e] + f[e] * length for e in range(3)]
    plain = box_mesh(n, n, n, length / n, 0, origin=origin)
    boxes = [lo((0.30, 0.35, 0.25)) + lo((0.45, 0.60, 0.55)), lo((0.70, 0.10, 0.60)) + lo((0.85, 0.30, 0.95))]
    out = {}
    for name, m in (("bodies", with_bodies(plain, boxes)), ("plain", plain)):
        U = mesh_vortex_state(m)
        solved = int(m["solved"].sum())
        with mmf.EulerSolver.from_mesh(m, device=device) as sol:
            if sol.info()["path"] != mmf.PATH_UNIFORM:
                raise RuntimeError("the box with bodies did not take the fused path")
            sol.set_state(mmf.FIELD_U, U)
            sol.run(0.45, m["h"], 0.0, 1.0e30, max_steps=3)
            sol.timer_start()
            sol.run(0.45, m["h"], 0.0, 1.0e30, max_steps=args.steps)
            ms = sol.timer_stop() / args.steps
        out[name] = {"ms_per_step": ms, "value": solved * 3.0 / (ms * 1e-3), "cells": int(m["volume"].shape[0]), "solved_cells": solved}
    return {"workload": f"3-D isentropic vortex on {n}^3 cells with two body boxes inside ({out['bodies']['cells'] - out['bodies']['solved_cells']} "
                        f"cells not solved, BC_WALL around them), full mesh description through mmf_create; fused path, kernel form 'b'",
            "unit": UNIT, "value": out["bodies"]["value"], "ms_per_step": out["bodies"]["ms_per_step"],
            "same_box_without_bodies": out["plain"], "ratio_to_plain": out["bodies"]["value"] / out["plain"]["value"]}


def run_ours(args):
    import ctypes as C
    import minimmerflow_b200 as mmf

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = args.gpus
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    px, py, pz = box_decomposition(world)
    cz, cy, cx = rank // (px * py), (rank // px) % py, rank % px
    dims, gdims = local_dims(args, (px, py, pz))
    offset = (cx * dims[0], cy * dims[1], cz * dims[2])
    # h from the x extent: the weak-scaling slabs keep the cell size of the 10-wide cube they grow from
    plane, h = vortex_state(dims, offset, gdims[0])
    cells_local = dims[0] * dims[1] * dims[2]
    cells_total = cells_local * world

    lib = mmf.load_library()

    def make_solver(box, gbox, off, hh):
        """One rank's handle for the box `box` at `off` of the global lattice `gbox`, the benchmark's numbering
        conventions, halo exchange set up for the process grid."""
        sol = mmf.EulerSolver.uniform(box, hh, [mmf.BC_FREE_FLOW] * 6, device=local_rank,
                                      cell_numbering=mmf.NUMBERING_LEXICOGRAPHIC, interface_numbering=mmf.NUMBERING_MORTON,
                                      global_dims=gbox, box_offset=off)
        if world > 1:
            uid = [mmf.EulerSolver.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            sol.comm_init(rank, world, uid[0])
            def nb(dx, dy, dz):
                x, y, z = cx + dx, cy + dy, cz + dz
                if not (0 <= x < px and 0 <= y < py and 0 <= z < pz):
                    return -1
                return (z * py + y) * px + x
            sol.comm_set_box_neighbours([nb(-1, 0, 0), nb(1, 0, 0), nb(0, -1, 0), nb(0, 1, 0), nb(0, 0, -1), nb(0, 0, 1)])
            if not args.nccl_halo:
                # halo exchange by direct peer stores over NVLink: gather every rank's CUDA IPC handles
                blobs = [None] * world
                dist.all_gather_object(blobs, sol.comm_ipc_export())
                sol.comm_ipc_import(blobs)
        return sol

    # parity first: the numbers below are only worth reading next to a green check of the same code path
    parity = parity_check(args, make_solver, dist, rank, world, (px, py, pz), (cx, cy, cz))

    # pinned host AoS buffer (the reference's storage layout), filled plane by plane
    nbytes = cells_local * 5 * 8
    hptr = C.c_void_p()
    mmf._cabi.check(lib.mmf_host_alloc(C.byref(hptr), nbytes))
    host = np.ctypeslib.as_array((C.c_double * (cells_local * 5)).from_address(hptr.value)).reshape(
        dims[2], dims[1] * dims[0], 5)
    flat_plane = plane.reshape(-1, 5)
    for k in range(dims[2]):
        host[k] = flat_plane

    s = make_solver(dims, gdims, offset, h)

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()
        s.synchronize()

    cfl, t_inf = 0.45, 1.0e30
    s.set_state_ptr(mmf.FIELD_U, hptr.value)
    s.run(cfl, h, 0.0, t_inf, max_steps=max(args.warmup, 3))         # warm-up (>= 3 steps)

    sampler = ClockSampler(local_rank)
    sampler.start()
    repeats_ms, launches = [], 0
    for _ in range(max(1, args.repeats)):
        launches0 = s.info()["kernel_launches"]
        barrier()
        s.timer_start()
        s.run(cfl, h, 0.0, t_inf, max_steps=args.steps)              # exactly K timed steps
        rep_ms = s.timer_stop()
        barrier()
        launches = s.info()["kernel_launches"] - launches0
        if dist is not None:                                         # a repeat takes as long as its slowest rank
            import torch
            tt = torch.tensor([rep_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            rep_ms = float(tt[0])
        repeats_ms.append(rep_ms)
    clocks = sampler.stop()
    ms = float(np.median(repeats_ms))

    # per-kernel device time of the fused stage kernels, CUDA events on the launching stream around every launch.  With
    # several ranks a boundary CTA waits IN the kernel for its neighbours' layers, so a launch also measures how far the
    # ranks drifted apart at that moment: the leg runs as five segments between barriers and reports the per-stage MEDIAN
    # of the segments' averages (at 8 GPUs a single late neighbour once put 27 ms into one stage-2 launch)
    seg_steps = max(args.steps // 5, 2)
    seg_ms = []
    for _ in range(5):
        barrier()
        s.profile_begin()
        s.run(cfl, h, 0.0, t_inf, max_steps=seg_steps)
        kms, kn = s.profile_end()
        seg_ms.append([kms[i] / kn[i] if kn[i] else None for i in (1, 2, 3)])

    # e2e: the call sequence a host like main.cpp makes, with HOST buffers and wall-clock time:
    # upload the AoS state (pinned), K x mmf_step -- each returns dt and the three max eigenvalues
    # main.cpp logs per step (src/main.cpp:399, 440, 476), i.e. one device->host read and one host
    # sync per step -- then download the state.
    barrier()
    t0 = time.perf_counter()
    s.set_state_ptr(mmf.FIELD_U, hptr.value)
    t_now = 0.0
    for _ in range(args.steps):
        dt_k, _eig = s.step(cfl, h, t_now, t_inf)
        t_now += dt_k
    s.get_state_ptr(mmf.FIELD_U, hptr.value)
    barrier()
    e2e_s = time.perf_counter() - t0

    # the STRICT drop-in call: euler::computeRHS with caller-owned host storages (mmf_compute_rhs_host: upload U,
    # residual kernel, download RHS and the max eigenvalue on every call) -- what the adapter of INTEGRATION.md costs
    # when main.cpp stays entirely unchanged; PCIe bound by construction.  One GPU only, pageable numpy buffers.
    strict = None
    if world == 1 and not args.no_strict:
        Uh = np.ascontiguousarray(host.reshape(-1, 5))
        Rh = np.empty_like(Uh)
        s.compute_rhs_host(Uh, out=Rh)                                # warm-up (allocates the RHS array)
        t0 = time.perf_counter()
        n_calls = 3
        for _ in range(n_calls):
            s.compute_rhs_host(Uh, out=Rh)
        strict_s = (time.perf_counter() - t0) / n_calls
        strict = {"value": cells_local / strict_s, "unit": "cell residuals/s", "ms_per_call": strict_s * 1e3,
                  "h2d_bytes_per_call": nbytes, "d2h_bytes_per_call": nbytes + 8,
                  "what": "mmf_compute_rhs_host = euler::computeRHS with host storages: upload, residual kernel, download, "
                          "wall clock; one call is one RK stage's residual (a third of a cell-update's work, none of its update)"}
        del Uh, Rh

    # secondary (one GPU): a box with bodies on the fused path, after the main handle's arrays are gone
    bodies = None
    if world == 1 and not args.no_bodies:
        s.close()
        try:
            bodies = bodies_secondary(args, mmf, local_rank)
        except Exception as e:  # the headline stands on its own: report, do not fail the line
            bodies = {"error": str(e)[-300:]}

    if dist is not None:
        import torch
        tt = torch.tensor([e2e_s * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt[0])
    else:
        e2e_ms = e2e_s * 1e3

    if rank == 0:
        value = cells_total * 3.0 * args.steps / (ms * 1e-3)
        peak, peak_src = measured_peak()
        stage_ms = [float(np.median([seg[i] for seg in seg_ms])) if all(seg[i] is not None for seg in seg_ms) else None
                    for i in range(3)]
        stage_gbs = [ALG_BYTES_PER_CELL_STAGE[i] * cells_local / (stage_ms[i] * 1e-3) / 1e9 if stage_ms[i] else None
                     for i in range(3)]
        # dominant kernel: the stage-2/3 kernel (one template, two of the three launches of a step)
        dom_ms = 0.5 * (stage_ms[1] + stage_ms[2]) if stage_ms[1] and stage_ms[2] else None
        achieved = ALG_BYTES_PER_CELL_STAGE[1] * cells_local / (dom_ms * 1e-3) / 1e9 if dom_ms else None
        step_gbs = sum(ALG_BYTES_PER_CELL_STAGE) * cells_local / (ms / args.steps * 1e-3) / 1e9
        # DRAM bytes of one launch of the dominant kernel: from the round's own `ncu --set full` capture of this
        # configuration (tools/make_profiles.py writes the file and the commit it was taken at), never under the timer
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "stage_kernel_traffic.json")) as f:
                tj = json.load(f)
            if tj.get("cells_per_launch") == cells_local:
                traffic = tj["stage23_dram_bytes_per_launch"]
                traffic_src = {k: tj.get(k) for k in ("source", "commit", "kernel")}
        except Exception:
            pass
        # secondary limiter: the FP64 pipe.  Algorithmic FP64 instructions per cell and stage (the
        # reference's arithmetic without FMA contraction, primitives once per cell) over the measured issue rate of the
        # pipe (profiles/r02_fp64_peak.md: tools/ubench_fp64 on this pool's B200)
        fp64_peak = FP64_LANES_PER_CLK_SM * 148 * 1.965e9
        fp64_ach = FP64_INSTR_PER_CELL_STAGE[1] * cells_local / (dom_ms * 1e-3) if dom_ms else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "repeats": {"n": len(repeats_ms), "ms_per_step": [r / args.steps for r in repeats_ms],
                        "spread": (max(repeats_ms) - min(repeats_ms)) / ms, "reported": "median",
                        "timed_seconds_total": sum(repeats_ms) * 1e-3},
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(dims, gdims, (px, py, pz)),
            "clocks": clocks,
            "parity": parity,
            "gpu_launches": int(launches),
            "e2e": {"value": cells_total * 3.0 * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": nbytes / args.steps, "d2h_bytes_per_step": nbytes / args.steps + 88,
                    "what": f"pinned host AoS upload, {args.steps} x mmf_step (dt and max eigenvalues read back and the host "
                            f"synchronised every step), download; wall clock through the C-ABI.  The state is resident between "
                            f"the steps, so the upload and the download are amortised over --steps",
                    "strict_dropin": strict},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "secondary": {"bound": "fp64", "achieved": fp64_ach, "peak": fp64_peak, "unit": "FP64 lane-instructions/s",
                                       "frac": (fp64_ach / fp64_peak) if fp64_ach else None,
                                       "what": f"{FP64_INSTR_PER_CELL_STAGE[1]} FP64 instructions per cell of stage 2/3 (algorithmic: "
                                               f"halo and padding work not counted) over {FP64_LANES_PER_CLK_SM} lanes/clk/SM x 148 SMs "
                                               f"x 1.965 GHz measured"},
                         "kernel": stage_kernel_label(),
                         "algorithmic_bytes_per_launch": ALG_BYTES_PER_CELL_STAGE[1] * cells_local,
                         "avg_launch_ms": dom_ms,
                         "stage_ms": stage_ms, "stage_GBps": stage_gbs,
                         "stage_ms_how": f"CUDA events around every launch of the stage kernels, 5 segments of {seg_steps} steps "
                                         f"between barriers, per-stage median of the segments' averages",
                         "algorithmic_bytes_per_cell": list(ALG_BYTES_PER_CELL_STAGE),
                         "whole_step": {"achieved": step_gbs, "frac": step_gbs / peak,
                                        "what": "320 B/cell per RK3 step over the timed step time (all kernels, launch gaps included)"}},
        }
        if bodies is not None:
            line["secondary"] = {"bodies": bodies}
        if not args.no_cpu_baseline and world == 1:
            entry, port, _ = cpu_baseline_entry(args, os.cpu_count() or 1)
            line["cpu_baseline"] = entry
            if port is not None:
                line["cpu_port_all_cores"] = port
        print(json.dumps(line), flush=True)

    s.close()
    lib.mmf_host_free(hptr)
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
