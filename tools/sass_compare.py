#!/usr/bin/env python
"""Compare the SASS of two builds of libmmf_b200.so kernel by kernel (whitespace-normalised instruction text and
encodings from `cuobjdump -sass`).  Used to show that a refactoring left every kernel that was validated on the GPU
untouched: a kernel whose SASS is identical needs no new parity run.

    python tools/sass_compare.py OLD.so NEW.so        # exit code 1 if a kernel present in both differs
"""
import re
import subprocess
import sys


def kernels(path):
    text = subprocess.run(["cuobjdump", "-sass", path], check=True, capture_output=True, text=True).stdout
    out, name = {}, None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = []
        elif name and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
            out[name].append(" ".join(line.split()))
    return out


def demangle(names):
    if not names:
        return []
    return subprocess.run(["c++filt"] + list(names), check=True, capture_output=True, text=True).stdout.splitlines()


def main():
    old, new = kernels(sys.argv[1]), kernels(sys.argv[2])
    both = [k for k in old if k in new]
    differ = [k for k in both if old[k] != new[k]]
    print(f"kernels: old {len(old)}, new {len(new)}, in both {len(both)}, identical {len(both) - len(differ)}, differ {len(differ)}")
    for label, names in (("DIFFERS", differ), ("ONLY IN OLD", [k for k in old if k not in new]),
                         ("ONLY IN NEW", [k for k in new if k not in old])):
        for k, d in zip(names, demangle(names)):
            print(f"{label}: {d.split('(')[0]}  [{len(old.get(k, []))} -> {len(new.get(k, []))} lines]")
    return 1 if differ else 0


if __name__ == "__main__":
    sys.exit(main())
