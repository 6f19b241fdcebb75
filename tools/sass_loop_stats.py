#!/usr/bin/env python
"""DEVELOPMENT TOOL: static instruction mix of the stage kernels' steady-state loops, read from the SASS of
the built library (no GPU needed).  For every kernel whose name matches, the largest backward-branch loop
(= the plane loop; the rotate / decoupled forms hold two planes per trip) is cut out of the listing and its
instructions are counted by pipe.

    python tools/sass_loop_stats.py                          # default: the forms bench.py can select, Morton order
    python tools/sass_loop_stats.py --match 'kernel_v7ILi2ELi0ELi8ELb0' --rows 14

Columns: planes per trip, instructions per plane and per updated cell row (a CTA-plane updates `rows` rows of
30 cells; `rows` as in StageShape::rows), FP64 / shuffle / shared / global / barrier-try-wait instructions per
plane of ONE warp.  A static count says nothing about stalls: it ranks variants before GPU time is spent.
"""
import argparse
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "minimmerflow_b200", "lib", "libmmf_b200.so")

PIPES = [
    ("fp64", re.compile(r"^(DADD|DMUL|DFMA|DSETP|DMNMX|MUFU\.RCP64H|MUFU\.RSQ64H)")),
    ("shfl", re.compile(r"^SHFL")),
    ("lds/sts", re.compile(r"^(LDS|STS)")),
    ("ldg/stg", re.compile(r"^(LDG|STG|LD\.|ST\.)")),
    ("local", re.compile(r"^(LDL|STL)")),
    ("sync", re.compile(r"^(SYNCS|BAR|WARPSYNC|ELECT|VOTE)")),
    ("mov", re.compile(r"^(MOV|IMAD\.MOV|UMOV)")),
]

INSN = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);")


def kernels(lib, match):
    out = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name and re.search(match, name):
                yield name, body
            name, body = m.group(1), []
        elif name:
            body.append(line)
    if name and re.search(match, name):
        yield name, body


def loop_stats(body):
    insns = []
    for line in body:
        m = INSN.match(line)
        if m:
            insns.append((int(m.group(1), 16), m.group(3), m.group(4), bool(m.group(2))))
    addr_index = {a: n for n, (a, _, _, _) in enumerate(insns)}
    best = None
    for n, (a, op, rest, pred) in enumerate(insns):
        # a loop closes with a PREDICATED backward branch; the unconditional backward branches are the
        # returns of out-of-line slow paths (barrier spin loops, divergent shuffles) at the end of the kernel
        if op == "BRA" and pred:
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t and int(t.group(1), 16) in addr_index and int(t.group(1), 16) < a:
                lo = addr_index[int(t.group(1), 16)]
                if best is None or n - lo > best[1] - best[0]:
                    best = (lo, n)
    if best is None:
        return None
    loop = insns[best[0]:best[1] + 1]
    counts = {k: 0 for k, _ in PIPES}
    for _, op, _, _ in loop:
        for k, rx in PIPES:
            if rx.match(op):
                counts[k] += 1
                break
    return len(insns), len(loop), counts


def describe(name):
    m = re.search(r"kernel_(v\d+[a-z]*)ILi(\d)ELi(\d)ELi(\d+)E(?:Lb([01])E)?(?:Lb([01])E)?", name)
    if not m:
        return name[:60], None, None
    form, stage, order, nw, xg, mh = m.groups()
    nw = int(nw)
    rows = {"v5": nw - 2, "v5r": nw - 2, "v5rb": nw - 2, "v3": nw - 2}.get(form)
    planes = 2
    if form == "v6":
        rows = nw - 1 if mh == "1" else nw - 2
        form = "v6h" if mh == "1" else "v6"
    if form == "v7":
        rows = 2 * (nw - 1)
    if form == "v5rb":   # the bool is FIXUP (form 'c'), not the compact x ghosts
        form, xg = ("v5rc" if xg == "1" else "v5rb"), "0"
    if form == "v3":
        planes = 1
    return "%-4s stage %s order %s nw %2d%s" % (form, stage, order, nw, " xg" if xg == "1" else ""), rows, planes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=LIB)
    ap.add_argument("--match", default=r"uniform_stage_kernel_v(5|5r|6|7)ILi[0-3]ELi0ELi(8|12|16)ELb0|uniform_stage_kernel_v5rbILi[0-3]ELi0ELi12")
    ap.add_argument("--rows", type=int, default=0, help="override the rows a CTA updates")
    ap.add_argument("--planes", type=int, default=0, help="override planes per loop trip")
    args = ap.parse_args()
    if not os.path.exists(args.lib):
        sys.exit("build the library first: make -j -C minimmerflow_b200/csrc")
    print("%-34s %6s %6s %9s %9s | per warp-plane: %s" % ("kernel", "total", "loop", "ins/plane", "ins/row",
                                                          " ".join("%7s" % k for k, _ in PIPES)))
    rows_out = []
    for name, body in kernels(args.lib, args.match):
        st = loop_stats(body)
        if st is None:
            continue
        label, rows, planes = describe(name)
        rows = args.rows or rows or 1
        planes = args.planes or planes or 1
        total, loop, counts = st
        nw = int(re.search(r"nw +(\d+)", label).group(1)) if "nw" in label else 1
        per_plane = loop / planes
        # instructions the whole CTA issues per plane, divided by the rows it updates
        per_row = per_plane * nw / rows
        rows_out.append((label, "%-34s %6d %6d %9.0f %9.0f | %s %s" % (
            label, total, loop, per_plane, per_row, " " * 15, " ".join("%7.0f" % (counts[k] / planes) for k, _ in PIPES))))
    for _, line in sorted(rows_out):
        print(line)
    print("ins/row = loop instructions per plane x warps per CTA / rows the CTA updates: all warps (halo warps\n"
          "execute the same loop length at most) charged to the updated rows")


if __name__ == "__main__":
    main()
