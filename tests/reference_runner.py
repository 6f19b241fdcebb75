"""Helpers that drive the reference-shaped executables of oracle/_ref/ (TEST INFRASTRUCTURE):

* ``minimmerflow_ref``          -- the UNMODIFIED reference sources compiled against compat/bitpit
                                   (CPU); built in the dev container by ``make -C oracle ref``.
* ``minimmerflow_b200_dropin``  -- the same unmodified main.cpp & co. linked with the GPU adapters
                                   (minimmerflow_b200/adapters/solver_b200.cpp) and libmmf_b200.so.
* ``minimmerflow_b200_resident`` -- the reference's set-up / output units unchanged, with
                                   minimmerflow_b200/adapters/driver_b200.cpp as main(): the time loop
                                   is a stream of mmf_step calls, the state stays in HBM.

Both read ./settings.xml and print the reference's " Final error:  %.12e" line; this module writes
the settings file from a case description, runs the binary in a scratch directory and parses the
line and, optionally, the final .vtu the reference's SolverWriter streams (src/solver_writer.cpp)."""
import os
import re
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_EXE = os.path.join(REF_DIR, "minimmerflow_ref")
DROPIN_EXE = os.path.join(REF_DIR, "minimmerflow_b200_dropin")
RESIDENT_EXE = os.path.join(REF_DIR, "minimmerflow_b200_resident")


def settings_xml(case):
    """settings.xml text for a case dict: problem, dim, n_cells, cfl and optionally t_end (>= 0),
    order, origin (3), length, bodies (list of [xMin,yMin,zMin,xMax,yMax,zMax])."""
    problem = ["<type>%s</type>" % case["problem"], "<dimensions>%d</dimensions>" % case["dim"]]
    if case.get("t_end", -1.0) >= 0:
        problem.append("<time><end>%r</end></time>" % float(case["t_end"]))
    domain = []
    if case.get("origin") is not None:
        o = case["origin"]
        domain.append("<origin><x>%r</x><y>%r</y><z>%r</z></origin>" % (float(o[0]), float(o[1]), float(o[2])))
    if case.get("length") is not None:
        domain.append("<length>%r</length>" % float(case["length"]))
    if domain:
        problem.append("<domain>%s</domain>" % "".join(domain))
    bodies = ""
    if case.get("bodies"):
        items = []
        for q, b in enumerate(case["bodies"]):
            items.append("<body%d><type>box</type><xMin>%r</xMin><yMin>%r</yMin><zMin>%r</zMin>"
                         "<xMax>%r</xMax><yMax>%r</yMax><zMax>%r</zMax></body%d>" % ((q,) + tuple(float(x) for x in b) + (q,)))
        bodies = "<bodies>%s</bodies>" % "".join(items)
    return ("<?xml version=\"1.0\"?>\n<minimmerflow version=\"1\">\n<problem>%s</problem>\n%s\n"
            "<discretization><space><order>%d</order><nCells>%d</nCells></space>"
            "<time><CFL>%r</CFL></time></discretization>\n</minimmerflow>\n"
            % ("".join(problem), bodies, int(case.get("order", 1)), int(case["n_cells"]), float(case["cfl"])))


def read_vtu(path):
    """Cell fields of a .vtu written with appended raw data and UInt64 block headers."""
    blob = open(path, "rb").read()
    marker = blob.index(b"<AppendedData")
    start = blob.index(b"_", marker) + 1
    header = blob[:marker].decode()
    n_cells = int(re.search(r'NumberOfCells="(\d+)"', header).group(1))
    dtypes = {"Float64": np.float64, "Int32": np.int32, "Int64": np.int64, "UInt8": np.uint8}
    fields = {}
    for m in re.finditer(r'<DataArray type="(\w+)" Name="(\w+)" NumberOfComponents="(\d+)" format="appended" offset="(\d+)"/>', header):
        typ, name, comps, off = m.group(1), m.group(2), int(m.group(3)), int(m.group(4))
        nbytes = int(np.frombuffer(blob, np.uint64, 1, start + off)[0])
        arr = np.frombuffer(blob, dtypes[typ], nbytes // np.dtype(dtypes[typ]).itemsize, start + off + 8)
        fields[name] = arr.reshape(-1, comps).copy() if comps > 1 else arr.copy()
    fields["_n_cells"] = n_cells
    return fields


def run_case(exe, case, want_fields=False, extra_env=None, timeout=600):
    """Runs `exe` on `case` in a scratch directory.  Returns dict(final_error=<string>, steps=<int>,
    fields=<dict or None>, output=<stdout>)."""
    env = dict(os.environ)
    env["BITPIT_SHIM_VTK"] = "1" if want_fields else "0"
    if extra_env:
        env.update(extra_env)
    with tempfile.TemporaryDirectory(prefix="mmf_ref_") as tmp:
        with open(os.path.join(tmp, "settings.xml"), "w") as f:
            f.write(settings_xml(case))
        # argv[2] = nSaves: 1 -> intermediate output once instead of after every step (src/main.cpp:119-122, 509-523)
        out = subprocess.run([exe, str(int(case["n_cells"])), "1"], cwd=tmp, env=env, capture_output=True, timeout=timeout)
        text = out.stdout.decode("ascii", "ignore")
        if out.returncode != 0:
            raise RuntimeError("%s failed (rc=%d): %s\n%s" % (exe, out.returncode, out.stderr.decode("ascii", "ignore")[-2000:], text[-500:]))
        line = [ln for ln in text.splitlines() if "Final error:" in ln][0]          # test/test_driver.py:42-49
        steps = [int(ln.split("Step n.")[1]) for ln in text.splitlines() if ln.startswith("Step n.")]
        fields = None
        if want_fields:
            name = [p for p in os.listdir(tmp) if p.startswith("final_background_") and p.endswith(".vtu")][0]
            fields = read_vtu(os.path.join(tmp, name))
        return dict(final_error=line.split("Final error:")[1].strip(), steps=(steps[-1] + 1) if steps else 0,
                    fields=fields, output=text)
