#!/usr/bin/env python
"""Per-instruction stall samples of one profiled launch (ncu --page source --csv): totals per stall class, the
instructions with the most samples and the samples per instruction class.  Development tool.
  python tools/ncu_source_top.py gpurun_out/x.ncu-rep --launch 1 [--top 40]"""
import argparse, collections, csv, io, re, subprocess
ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("--launch", type=int, default=0); ap.add_argument("--top", type=int, default=40)
ap.add_argument("--hot-only", action="store_true", help="only instructions executed at least half as often as the most executed one")
a = ap.parse_args()
out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--launch-skip", str(a.launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
print(lines[0][:160])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]; rows = rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); n = 0
execs = [int(r[ix["Instructions Executed"]]) for r in rows]
emax = max(execs)
for r in rows:
    for s in stalls: tot[s] += int(r[ix[s]])
    n += int(r[ix["# Samples"]])
print("samples", n, " instructions", len(rows), " max executed", emax)
print("by class:", ", ".join(f"{k[6:]} {v} ({100*v/n:.1f}%)" for k, v in tot.most_common(10)))
cls = collections.Counter(); cnt = collections.Counter(); exe = collections.Counter()
for r, e in zip(rows, execs):
    op = r[ix["Source"]].split()
    op = [o for o in op if not o.startswith("@")]
    m = op[0].split(".")[0] if op else "?"
    if m == "SYNCS": m = ".".join(op[0].split(".")[:2])
    cls[m] += int(r[ix["# Samples"]]); cnt[m] += 1; exe[m] += e
print("by opcode (samples, static count, executed/maxexec):")
for k, v in cls.most_common(25): print(f"  {k:18s} {v:7d} {100*v/n:5.1f}%  n={cnt[k]:4d} exec={exe[k]/emax:7.1f}")
print("top instructions:")
srt = sorted(range(len(rows)), key=lambda i: -int(rows[i][ix["# Samples"]]))
for i in srt[:a.top]:
    r = rows[i]
    top = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:3]
    print(f"  {i:5d} {int(r[ix['# Samples']]):6d} exec {execs[i]/emax:5.2f}  {r[ix['Source']].strip()[:70]:70s} " + " ".join(f"{s}={v}" for v, s in top if v))
print("dynamic instruction mix (warp instructions executed):")
tote = sum(execs)
mix = collections.Counter()
for r, e in zip(rows, execs):
    op = [o for o in r[ix["Source"]].split() if not o.startswith("@")]
    m = op[0].split(".")[0] if op else "?"
    mix[m] += e
for k, v in mix.most_common(30): print(f"  {k:10s} {v:12d} {100*v/tote:5.1f}%")
print("  total", tote)
