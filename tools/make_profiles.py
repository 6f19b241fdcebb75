#!/usr/bin/env python
"""Turns what a gpurun profiling call brought back (gpurun_out/*.ncu-rep, launch-list csv) into the
tracked summaries under profiles/:  <tag>_launches.csv (verbatim), <tag>_launch_shares.md,
<tag>_ncu_stage_kernels.md, stage_kernel_traffic.json (read by bench.py for roofline.traffic) and
profiles/sass/*.sass (cuobjdump listings of the hot kernels, encodings stripped).

  python tools/make_profiles.py --tag r01c --rep gpurun_out/x.ncu-rep --launches gpurun_out/x.csv [--cells 16777216]
"""
import argparse
import collections
import csv
import json
import os
import re
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")

METRICS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", required=True)
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    ap.add_argument("--cells", type=int, default=256 ** 3)
    ap.add_argument("--command", default="")
    args = ap.parse_args()
    os.makedirs(PROF, exist_ok=True)

    if args.launches:
        shutil.copy(args.launches, os.path.join(PROF, f"{args.tag}_launches.csv"))
        rows = [r for r in csv.reader(open(args.launches)) if len(r) > 10 and r[0].isdigit()]
        tot = collections.OrderedDict()
        for r in rows:
            tot.setdefault(r[4].split("(")[0].replace("void ", ""), []).append(float(r[-1]) / 1e3)
        total = sum(sum(v) for v in tot.values())
        with open(os.path.join(PROF, f"{args.tag}_launch_shares.md"), "w") as f:
            f.write(f"# Launch list shares ({args.tag})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over "
                    f"{len(rows)} launches of `{args.command or 'python bench.py'}` (cold-cache, serialised: compare shares).\n\n"
                    "| kernel | launches | avg us | total us | share |\n|---|---|---|---|---|\n")
            for k, v in sorted(tot.items(), key=lambda kv: -sum(kv[1])):
                f.write(f"| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {sum(v):.1f} | {100 * sum(v) / total:.1f} % |\n")

    if args.rep:
        hdr, units, rows = ncu_raw(args.rep)
        ix = {h: i for i, h in enumerate(hdr)}
        names = [r[ix["Kernel Name"]] for r in rows]
        stalls = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
        with open(os.path.join(PROF, f"{args.tag}_ncu_stage_kernels.md"), "w") as f:
            f.write(f"# ncu --set full, fused stage kernels ({args.tag})\n\nCommand: `{args.command}`\n\n")
            f.write("| metric | unit | " + " | ".join(f"launch {i + 1}" for i in range(len(rows))) + " |\n")
            f.write("|---|---|" + "---|" * len(rows) + "\n")
            f.write("| kernel | | " + " | ".join(re.sub(r"\(mmf::UniformGeom.*", "", n).replace("void mmf::", "") for n in names) + " |\n")
            for m in METRICS + stalls:
                if m in ix:
                    short = m.replace("smsp__average_warps_issue_stalled_", "stall: ").replace("_per_issue_active.ratio", "")
                    f.write(f"| {short} | {units[ix[m]]} | " + " | ".join(r[ix[m]] for r in rows) + " |\n")
        # traffic of the dominant kernel (stages 2/3) for bench.py
        def gb(r, key):
            v, u = float(r[ix[key]]), units[ix[key]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
        def stage_of(n):
            m = re.search(r"uniform_stage_kernel\w*<(?:\(int\))?(\d)", n)
            return int(m.group(1)) if m else -1
        t23 = [gb(r, "dram__bytes_read.sum") + gb(r, "dram__bytes_write.sum") for r, n in zip(rows, names) if stage_of(n) in (2, 3)]
        t1 = [gb(r, "dram__bytes_read.sum") + gb(r, "dram__bytes_write.sum") for r, n in zip(rows, names) if stage_of(n) == 1]
        commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
        k23 = [re.sub(r"\(mmf::UniformGeom.*", "", n).replace("void mmf::", "") for n in names if stage_of(n) in (2, 3)]
        json.dump({"source": f"profiles/{args.tag}_ncu_stage_kernels.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)",
                   "commit": commit, "kernel": sorted(set(k23)),
                   "cells_per_launch": args.cells,
                   "stage23_dram_bytes_per_launch": sum(t23) / max(len(t23), 1),
                   "stage1_dram_bytes_per_launch": sum(t1) / max(len(t1), 1),
                   "stage23_algorithmic_bytes_per_launch": 120.0 * args.cells,
                   "stage1_algorithmic_bytes_per_launch": 80.0 * args.cells},
                  open(os.path.join(PROF, "stage_kernel_traffic.json"), "w"), indent=1)

    # SASS listings of the hot kernels (Morton order = the reference's numbering)
    lib = os.path.join(ROOT, "minimmerflow_b200", "lib", "libmmf_b200.so")
    if os.path.exists(lib):
        sass_dir = os.path.join(PROF, "sass")
        os.makedirs(sass_dir, exist_ok=True)
        text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
        chunks = re.split(r"\n(?=\s*Function : )", text)
        # (ELb0 = padded x ghost columns: the single-GPU instantiation the bench runs)
        want = {"uniform_stage_kernel_tILi0ELi0ELi12ELi4ELb1ELb0E": "stage0_rhs_only_h_12warps",
                "uniform_stage_kernel_tILi1ELi0ELi12ELi4ELb1ELb0E": "stage1_h_12warps",
                "uniform_stage_kernel_tILi2ELi0ELi12ELi4ELb1ELb0E": "stage2_h_12warps",
                "uniform_stage_kernel_tILi3ELi0ELi12ELi4ELb1ELb0E": "stage3_h_12warps",
                "uniform_stage_kernel_tILi2ELi0ELi16ELi4ELb0ELb0E": "stage2_t_16warps",
                # a box with bodies (kernel form 'b'): the same kernel with the flag array
                "uniform_stage_kernel_tILi2ELi0ELi12ELi4ELb1ELb1E": "stage2_b_bodies_12warps",
                "uniform_eig_wall_kernel": "uniform_eig_wall",
                # the rotate form (per-thread global loads): what runs when an x side is a partition side
                "uniform_stage_kernel_v5rILi2ELi0ELi12ELb0": "stage2_v5r_12warps",
                "uniform_stage_kernel_v5rILi2ELi0ELi12ELb1": "stage2_v5r_12warps_xghost",
                # ... and the pass over the wall cells around it
                "uniform_wall_cells_kernelILi2ELi0": "stage2_wall_cells",
                "uniform_eig_body_kernel": "uniform_eig_body",
                "uniform_eig_kernel": "uniform_eig", "uniform_ghost_kernel": "uniform_ghost",
                "generic_rhs_kernel": "generic_rhs", "generic_rhs_derived_kernel": "generic_rhs_derived",
                "generic_stage_kernelILi2": "generic_stage2_fused", "generic_rk_kernelILi1": "generic_rk_stage1",
                "uniform_layer_kernel": "uniform_layer_pack"}
        for c in chunks:
            m = re.match(r"\s*Function : (\S+)", c)
            if not m:
                continue
            for key, fname in want.items():
                if key in m.group(1):
                    body = re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/", "", c)
                    body = body.split("\nFatbin elf code:")[0]
                    body = "\n".join(ln.rstrip() for ln in body.splitlines() if ln.strip())
                    open(os.path.join(sass_dir, fname + ".sass"), "w").write(body + "\n")


if __name__ == "__main__":
    main()
