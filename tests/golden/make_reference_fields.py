#!/usr/bin/env python
"""Generates tests/golden/reference_fields.npz by running the UNMODIFIED reference
(oracle/_ref/minimmerflow_ref, built by `make -C oracle ref` from /root/reference/src compiled
against compat/bitpit) on the cases of reference_cases.json.  Runs in the dev container only (the
GPU box has no reference tree); the fixture it writes is committed and travels.

Per case it stores the printed "Final error" string, the step count and the final cell fields the
reference's SolverWriter streams into final_background_<N>.vtu (src/solver_writer.cpp:56-119):
density, velocity[3], pressure, temperature, solved, and the last stage's residuals."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import reference_runner as R  # noqa: E402


def main():
    cases = json.load(open(os.path.join(HERE, "reference_cases.json")))["cases"]
    out = {}
    for case in cases:
        r = R.run_case(R.REF_EXE, case, want_fields=True)
        f = r["fields"]
        n = case["name"]
        out[n + "/final_error"] = np.array(r["final_error"])
        out[n + "/steps"] = np.array(r["steps"])
        for key in ("density", "velocity", "pressure", "temperature", "solved"):
            out[n + "/" + key] = f[key]
        out[n + "/residual"] = np.stack([f["residualC"], f["residualMX"], f["residualMY"], f["residualMZ"], f["residualE"]], axis=1)
        print(n, r["final_error"], r["steps"], "steps,", f["_n_cells"], "cells")
    np.savez_compressed(os.path.join(HERE, "reference_fields.npz"), **out)


if __name__ == "__main__":
    main()
