#!/usr/bin/env python
"""DEVELOPMENT TOOL: run the fused stage kernels' SOURCE on the CPU SIMT emulator (tools/emu) and compare
with the CPU oracle, bit for bit.  Purpose: catch protocol errors of a new kernel variant (mbarrier phase
hazards, buffer reuse races, index slips) before GPU time is spent on it.  Not a product path, not a
fallback, never timed; the GPU tests (tests/test_uniform_gpu.py) remain the parity proof.

    python tools/emu/run_emu.py                 # default matrix
    python tools/emu/run_emu.py --forms r,t --nw 8,12 --chaos 300 --repeat 5
"""
import argparse
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

XOFF = 1   # position of cell 0 in a padded row (uniform_device.cuh)
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
from common import bits_equal, lexicographic_box_mesh  # noqa: E402

CSRC = os.path.join(ROOT, "minimmerflow_b200", "csrc")
LIB = os.environ.get("MMF_EMU_LIB") or os.path.join(HERE, "build", "libmmf_emu.so")
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("emu_stage.cpp", "emu_ptx_helpers.h", "shim/cuda_runtime.h")]
    srcs += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh") or f.endswith(".h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-ffp-contract=off", "-Wno-unknown-pragmas",
           "-I", os.path.join(HERE, "shim"), "-I", HERE, "-I", CSRC, '-DMMF_EMU_PTX_HELPERS="emu_ptx_helpers.h"',
           "-o", LIB, os.path.join(HERE, "emu_stage.cpp")]
    # kernel build options under test, e.g. MMF_EMU_CXXFLAGS="-DMMF_R_UNROLL=1" (use with MMF_EMU_LIB=<other file>)
    cmd[1:1] = os.environ.get("MMF_EMU_CXXFLAGS", "").split()
    subprocess.run(cmd, check=True)


def load():
    build()
    lib = C.CDLL(LIB)
    lib.emu_padded.argtypes = [_I, _I, C.POINTER(C.c_longlong)]
    lib.emu_stage.restype = C.c_int
    lib.emu_stage.argtypes = [C.c_int] * 5 + [_I, _I, _I, _I, C.c_double, C.c_double, C.c_double, _D, _I,
                                              _D, _D, _D, C.c_double, _D, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_uint]
    lib.emu_set_spin_limit.argtypes = [C.c_longlong]
    lib.emu_set_xghost.argtypes = [_D, _D, C.c_longlong, C.c_int]
    lib.emu_set_solid.argtypes = [C.c_void_p]
    lib.emu_analyze_box.restype = C.c_int
    lib.emu_analyze_box.argtypes = [C.c_void_p, C.c_int, _I]
    lib.emu_body_flags.restype = C.c_int
    lib.emu_body_flags.argtypes = [_I, C.c_longlong, _I, C.c_void_p, C.c_int, C.c_void_p, _I]
    lib.emu_wall_cells.restype = C.c_double
    lib.emu_wall_cells.argtypes = [C.c_int, C.c_int, _I, _I, C.c_double, C.c_double, _I, _D, _D, C.c_void_p, _I, C.c_int,
                                   C.c_double, _D]
    lib.emu_eig_body.restype = C.c_double
    lib.emu_eig_body.argtypes = [_I, _I, _D, C.c_void_p]
    lib.emu_generic_tables.restype = C.c_longlong
    lib.emu_generic_tables.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), _I, C.c_void_p, C.c_char_p, C.c_int]
    lib.emu_generic_stride.restype = C.c_longlong
    lib.emu_generic_stride.argtypes = [C.c_longlong]
    lib.emu_generic_run.restype = C.c_int
    lib.emu_generic_run.argtypes = [C.c_void_p, C.c_int, _D, _D, _D, C.c_double, C.c_double, _D, C.c_double, C.c_int, _D, _D]
    return lib


def tile_rows(form, nw):
    """y rows a CTA of nw warps updates (StageShape::rows in uniform_launch.cuh)."""
    return nw - 1 if form in ("h", "b") else nw - 2


def ia(v):
    return np.ascontiguousarray(v, dtype=np.int32)


class Box:
    """Padded SoA arrays of one box, laid out like uniform_alloc (uniform_path.cuh)."""

    def __init__(self, lib, oracle, m, order):
        self.lib, self.oracle, self.m, self.order = lib, oracle, m, order
        self.dims = ia(m["box_dims"])
        pad = ia([0, 0, 0])
        fs = C.c_longlong(0)
        lib.emu_padded(self.dims.ctypes.data_as(_I), pad.ctypes.data_as(_I), C.byref(fs))
        self.pad, self.fs = pad, fs.value
        self.ijk = m["cell_ijk"].astype(np.int64)
        self.off = ((self.ijk[:, 2] + 1) * pad[1] + self.ijk[:, 1] + 1) * pad[0] + self.ijk[:, 0] + XOFF
        # one BC code per side, from the border interfaces of the host description
        self.bc = [-9] * 6
        for f in np.nonzero(m["neigh"] < 0)[0]:
            n = m["normal"][f]
            axis = int(np.argmax(np.abs(n)))
            self.bc[2 * axis + (1 if n[axis] > 0 else 0)] = int(m["bc"][f])
        nx, ny, nz = self.dims
        ff = [b == 0 for b in self.bc]
        self.clamp = ia([0 if ff[0] else -1, nx - 1 if ff[1] else nx, 0 if ff[2] else -1, ny - 1 if ff[3] else ny,
                         0 if ff[4] else -1, nz - 1 if ff[5] else nz])
        # bodies: one flag per padded cell (1 = not solved), the ghost shell repeats the cell it touches
        # (what uniform_try_create builds for kernel form 'b')
        self.solid = None
        if "solved" in m and not np.all(m["solved"]):
            px, py, pz = (int(v) for v in pad)
            inner = np.zeros((nz, ny, nx), np.uint8)
            inner[self.ijk[:, 2], self.ijk[:, 1], self.ijk[:, 0]] = 1 - m["solved"].astype(np.uint8)
            vol = np.zeros((pz, py, px), np.uint8)
            vol[1:nz + 1, 1:ny + 1, XOFF:nx + XOFF] = inner
            vol[0, :, :] = vol[1, :, :]; vol[nz + 1, :, :] = vol[nz, :, :]
            vol[:, 0, :] = vol[:, 1, :]; vol[:, ny + 1, :] = vol[:, ny, :]
            vol[:, :, XOFF - 1] = vol[:, :, XOFF]; vol[:, :, nx + XOFF] = vol[:, :, nx + XOFF - 1]
            self.solid = np.zeros(self.fs, np.uint8)
            self.solid[:px * py * pz] = vol.reshape(-1)
            # kernel form 'b': fluid cells with a wall interface get flag 2 and are listed by padded offset
            # (what uniform_try_create builds)
            step = (1, px, px * py)
            walls = []
            for o, (i, j, k) in sorted(zip(self.off.tolist(), self.ijk.tolist())):
                if self.solid[o] == 1:
                    continue
                ext = (nx, ny, nz)
                c = (i, j, k)
                if any((c[a] > 0 and self.solid[o - step[a]] == 1) or (c[a] < ext[a] - 1 and self.solid[o + step[a]] == 1)
                       for a in range(3)):
                    walls.append(o)
            self.flag_c = self.solid.copy()
            self.flag_c[walls] = 2
            self.walls = ia(walls)

    def new_array(self):
        a = np.empty((5, self.fs))
        a[:] = np.array([1.0, 0.0, 0.0, 0.0, 2.5])[:, None]  # fill_benign_kernel
        return a

    def scatter(self, arr, U):
        for f in range(5):
            arr[f, self.off] = U[:, f]

    def gather(self, arr):
        return np.stack([arr[f, self.off] for f in range(5)], axis=1)

    def fill_ghosts(self, arr):
        """uniform_ghost_kernel: virtual states of the non-free-flow sides (free-flow sides are clamped)."""
        nx, ny, nz = (int(v) for v in self.dims)
        px, py = int(self.pad[0]), int(self.pad[1])
        prob = oracle_lib.PROBLEMS.get(self.m.get("problem", "radsod"), 3)
        point = np.zeros(3)
        for side in range(6):
            bc = self.bc[side]
            if bc == 0:
                continue
            axis, hi = side >> 1, side & 1
            n = np.zeros(3)
            n[axis] = 1.0 if hi else -1.0
            ext = [nx, ny, nz]
            ra = range(ext[(axis + 1) % 3])
            rb = range(ext[(axis + 2) % 3])
            for a in ra:
                for b in rb:
                    c = [0, 0, 0]
                    c[axis] = ext[axis] - 1 if hi else 0
                    c[(axis + 1) % 3], c[(axis + 2) % 3] = a, b
                    gcell = list(c)
                    gcell[axis] = ext[axis] if hi else -1
                    oc = ((c[2] + 1) * py + c[1] + 1) * px + c[0] + XOFF
                    og = ((gcell[2] + 1) * py + gcell[1] + 1) * px + gcell[0] + XOFF
                    cons = np.ascontiguousarray(arr[:, oc])
                    out = np.empty(5)
                    self.oracle.lib.orc_eval_interface_bc_values(prob, bc, point.ctypes.data_as(_D), n.ctypes.data_as(_D),
                                                                 cons.ctypes.data_as(_D), out.ctypes.data_as(_D))
                    arr[:, og] = out

    def compact_x_ghosts(self, arr):
        """The XGhost layout of the multi-GPU path: the x ghost columns of `arr` as compact arrays
        [field][k+1][j+1]; the padded array's own x ghost columns are then poisoned, so a kernel that still
        read them would produce NaNs."""
        nx, ny, nz = (int(v) for v in self.dims)
        px, py, pz = (int(v) for v in self.pad)
        pitch = ny + 2
        fs = (pitch * (nz + 2) + 15) // 16 * 16
        lo = np.empty((5, fs)); hi = np.empty((5, fs))
        lo[:] = np.array([1.0, 0.0, 0.0, 0.0, 2.5])[:, None]
        hi[:] = lo
        vol = arr[:, :px * py * pz].reshape(5, pz, py, px)
        lo[:, :pitch * (nz + 2)] = vol[:, :, :, XOFF - 1].reshape(5, -1)
        hi[:, :pitch * (nz + 2)] = vol[:, :, :, nx + XOFF].reshape(5, -1)
        vol[:, :, :, XOFF - 1] = np.nan
        vol[:, :, :, nx + XOFF] = np.nan
        return lo, hi, fs, pitch

    def smem_doubles(self, form, nw):
        base = nw * 16 * 32 + 2 * nw                 # records and fluxes, two mbarriers per row
        if form in ("t", "h", "b"):                  # the ring of bulk tensor loads + records, fluxes, mbarriers (stage_t_smem_bytes)
            depth = 4
            nu = tile_rows(form, nw)
            return depth * 5 * 32 * (2 * nu + 2) + (nu + 2) * 11 * 32 + 2 * depth + 2 * (nu + 2)
        return base

    def eig_body(self, arr):
        """uniform_eig_body_kernel: the max eigenvalue that chooses dt on a box with bodies."""
        return self.lib.emu_eig_body(self.dims.ctypes.data_as(_I), ia(self.bc).ctypes.data_as(_I), arr.ctypes.data_as(_D),
                                     self.solid.ctypes.data)

    def stage(self, form, stage, nw, lz, Sin, Un, Out, dt, chaos=0, seed=1):
        m = self.m
        me = np.zeros(1)
        nx, ny, nz = (int(v) for v in self.dims)
        rows = tile_rows(form, nw)
        ntiles = ((nx + 29) // 30) * ((ny + rows - 1) // rows) * ((nz + lz - 1) // lz)
        est = np.zeros(ntiles, np.float32)
        zero3 = ia([0, 0, 0])
        dirichlet = np.zeros(5)
        body_form = form == "b"
        if (self.solid is not None) != body_form:
            raise RuntimeError("kernel form 'b' is the one for a box with bodies, and only that one")
        flags = self.flag_c if body_form else self.solid
        self.lib.emu_set_solid(flags.ctypes.data if flags is not None else None)
        e_wall, compact = 0.0, None
        if body_form:     # the wall cells BEFORE the stage kernel (stage 3 updates U in place)
            compact = np.empty((5, max(len(self.walls), 1)))
            e_wall = self.lib.emu_wall_cells(stage, self.order, self.dims.ctypes.data_as(_I), ia(self.bc).ctypes.data_as(_I),
                                             float(m["area"][0]), float(m["volume"][0]), self.clamp.ctypes.data_as(_I),
                                             Sin.ctypes.data_as(_D), Un.ctypes.data_as(_D), flags.ctypes.data,
                                             self.walls.ctypes.data_as(_I), len(self.walls), dt, compact.ctypes.data_as(_D))
        rc = self.lib.emu_stage(ord(form), stage, self.order, nw, lz, self.dims.ctypes.data_as(_I), zero3.ctypes.data_as(_I),
                                self.dims.ctypes.data_as(_I), ia(self.bc).ctypes.data_as(_I), float(m["h"]),
                                float(m["area"][0]), float(m["volume"][0]), dirichlet.ctypes.data_as(_D),
                                self.clamp.ctypes.data_as(_I), Sin.ctypes.data_as(_D), Un.ctypes.data_as(_D),
                                Out.ctypes.data_as(_D), dt, me.ctypes.data_as(_D), est.ctypes.data_as(C.POINTER(C.c_float)),
                                self.smem_doubles(form, nw), chaos, seed)
        if rc:
            raise RuntimeError(f"emu_stage: configuration not built (rc={rc})")
        if body_form and len(self.walls):     # uniform_wall_scatter_kernel
            Out[:, self.walls] = compact
        return max(float(me[0]), e_wall), est


def check_case(lib, oracle, name, m, order, form, nw, lz, steps, chaos, seed):
    box = Box(lib, oracle, m, order)
    if "problem" in m and m["problem"] in oracle_lib.PROBLEMS and "origin" in m:
        U0 = oracle.init_state(m)
    else:
        nc = m["volume"].shape[0]
        rng = np.random.default_rng(7)
        rho = rng.uniform(0.5, 1.5, nc); vel = rng.uniform(-0.4, 0.4, (nc, 3)); p = rng.uniform(0.6, 1.4, nc)
        U0 = np.column_stack([rho, rho * vel[:, 0], rho * vel[:, 1], rho * vel[:, 2], p / 0.4 + 0.5 * rho * (vel ** 2).sum(1)])
        m.setdefault("problem", "radsod")
    t0 = time.time()
    ok = True
    # RHS only
    ref_rhs, ref_eig = oracle.compute_rhs(m, U0)
    U = box.new_array(); Wa = box.new_array(); Wb = box.new_array(); R = box.new_array()
    if box.solid is not None:
        R[:] = 0.0   # uniform_ensure_rhs: solid cells keep the zero euler::computeRHS starts from
    box.scatter(U, U0)
    box.fill_ghosts(U)
    eig, _ = box.stage(form, 0, nw, lz, U, U, R, 0.0, chaos, seed)
    got = box.gather(R)
    if not (bits_equal(got, ref_rhs) and eig == ref_eig):
        ok = False
        bad = int((got != ref_rhs).any(axis=1).sum())
        print(f"   RHS mismatch: {bad} cells differ, eig {eig!r} vs {ref_eig!r}")
    # fused steps
    Uo, Wo, Ro = U0.copy(), np.zeros_like(U0), np.zeros_like(U0)
    t = 0.0
    for s in range(steps):
        dt, me3 = oracle.step(m, 0.45, t, 1e30, Uo, Wo, Ro)
        if box.solid is not None and box.eig_body(U) != me3[0]:
            ok = False
            print(f"   step {s}: eigenvalue pass {box.eig_body(U)!r} vs stage-1 face maximum {me3[0]!r}")
        e1, _ = box.stage(form, 1, nw, lz, U, U, Wa, dt, chaos, seed + 10 * s + 1)
        box.fill_ghosts(Wa)
        e2, _ = box.stage(form, 2, nw, lz, Wa, U, Wb, dt, chaos, seed + 10 * s + 2)
        box.fill_ghosts(Wb)
        e3, est = box.stage(form, 3, nw, lz, Wb, U, U, dt, chaos, seed + 10 * s + 3)
        box.fill_ghosts(U)
        t += dt
        got = box.gather(U)
        if not bits_equal(got, Uo) or [e1, e2, e3] != list(me3):
            ok = False
            bad = int((got != Uo).any(axis=1).sum())
            print(f"   step {s}: {bad} cells differ; eig {[e1, e2, e3]} vs {list(me3)}")
            break
        # stage 3's FP32 estimate per TILE (what uniform_eig_select/tiles_kernel work from) must match the
        # largest cell eigenvalue max_d|u_d| + a of exactly that tile's cells
        rows = tile_rows(form, nw)
        nx, ny, nz = (int(v) for v in box.dims)
        tx, ty = (nx + 29) // 30, (ny + rows - 1) // rows
        rho = Uo[:, 0]; vel = Uo[:, 1:4] / rho[:, None]
        pr = 0.4 * (Uo[:, 4] - 0.5 * rho * (vel ** 2).sum(1))
        lam = np.abs(vel).max(1) + np.sqrt(1.4 * pr / rho)
        tile = ((box.ijk[:, 2] // lz) * ty + box.ijk[:, 1] // rows) * tx + box.ijk[:, 0] // 30
        want = np.zeros(est.shape[0])
        if box.solid is not None:
            lam = np.where(m["solved"] != 0, lam, 0.0)   # solid cells are never written: no estimate
            if form == "b":                              # ... nor are the wall cells, by the stage kernel itself
                lam = np.where(box.flag_c[box.off] == 0, lam, 0.0)
        np.maximum.at(want, tile, lam)
        if not np.allclose(est, want, rtol=2e-5, atol=0):
            ok = False
            bad = int((~np.isclose(est, want, rtol=2e-5, atol=0)).sum())
            print(f"   step {s}: {bad} of {est.shape[0]} tile estimates off (max rel {np.abs(est / np.maximum(want, 1e-300) - 1).max():.2e})")
    print(f"{'ok  ' if ok else 'FAIL'} {name:28s} form {form} nw {nw:2d} lz {lz:2d} order {order} chaos {chaos:3d}  ({time.time() - t0:.1f} s)")
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--forms", default="r,t,h,b")
    ap.add_argument("--nw", default="8,12,16")
    ap.add_argument("--chaos", type=int, default=0, help="max random delay (us) around mbarrier operations")
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    lib = load()
    oracle = oracle_lib.load()
    forms = [f for f in args.forms.split(",") if f]
    cases = []
    m = oracle.problem_mesh("vortex_xy", 3, 16)
    cases.append(("vortex 16^3 morton", m, 0, 6))
    if not args.quick:
        cases.append(("radsod 16^3 morton reflect", oracle.problem_mesh("radsod", 3, 16), 0, 16))
        cases.append(("box 37x9x5 lexi free-flow", lexicographic_box_mesh(37, 9, 5, 0.25, 0), 1, 3))
        cases.append(("box 5x23x7 lexi reflecting", lexicographic_box_mesh(5, 23, 7, 0.5, 1), 1, 7))
        cases.append(("box 31x7x2 lexi free-flow", lexicographic_box_mesh(31, 7, 2, 0.5, 0), 1, 1))
        cases.append(("box 33x8x5 lexi reflecting", lexicographic_box_mesh(33, 8, 5, 0.5, 1), 1, 2))
    # boxes with bodies (kernel form 'b' only): a box body off the Morton cube's centre, two bodies touching
    # the border, a one-cell body
    body_cases = []
    body_forms = [f for f in forms if f == "b"]
    if body_forms:
        forms = [f for f in forms if f != "b"]
        body_cases.append(("radsod 16^3 + box body", oracle.problem_mesh("radsod", 3, 16, boxes=[[2.1, 3.2, 1.3, 4.9, 5.4, 3.6]]), 0, 6))
        if not args.quick:
            body_cases.append(("sod3d_x 16^3 + 2 bodies at the border", oracle.problem_mesh(
                "sod3d_x", 3, 16, boxes=[[-1.1, -1.1, -1.1, -0.6, -0.3, -0.8], [0.3, 0.1, 0.4, 1.1, 1.1, 1.1]]), 0, 5))
            body_cases.append(("vortex 16^3 + one-cell body", oracle.problem_mesh(
                "vortex_xy", 3, 16, boxes=[[0.0, 0.0, 0.0, 0.6, 0.6, 0.6]]), 0, 16))
    all_ok = True
    for rep in range(args.repeat):
        for name, mesh, order, lz in body_cases:
            for form in body_forms:
                for nw in (int(x) for x in args.nw.split(",")):
                    if form == "b" and nw == 8:
                        continue   # (the merged halo warp exists at 12 and 16 warps)
                    all_ok &= check_case(lib, oracle, name, dict(mesh), order, form, nw, lz, args.steps, args.chaos, 1 + 100 * rep)
        for name, mesh, order, lz in cases:
            for form in forms:
                for nw in (int(x) for x in args.nw.split(",")):
                    all_ok &= check_case(lib, oracle, name, dict(mesh), order, form, nw, lz, args.steps, args.chaos, 1 + 100 * rep)
    print("ALL OK" if all_ok else "FAILURES")
    return 0 if all_ok else 1


if __name__ == "__main__":
    sys.exit(main())
