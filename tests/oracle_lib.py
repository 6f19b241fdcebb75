"""ctypes front end of the CPU oracle (oracle/mmf_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "build", "libmmf_oracle.so")

PROBLEMS = {"vortex_xy": 0, "vortex_zx": 1, "vortex_yz": 2, "radsod": 3,
            "sod3d_x": 4, "sod3d_y": 5, "sod3d_z": 6, "ffstep": 7}

_D = C.POINTER(C.c_double)
_I64 = C.POINTER(C.c_int64)
_I32 = C.POINTER(C.c_int32)
_U8 = C.POINTER(C.c_uint8)


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        L = lib
        L.orc_level_for.restype = C.c_int
        L.orc_level_for.argtypes = [C.c_double, C.c_long]
        L.orc_uniform_counts.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.orc_uniform_mesh.argtypes = [C.c_int, _D, C.c_double, C.c_int, _I64, _I64, _D, _D, _D, _D, _D, _D, _I32]
        L.orc_domain_defaults.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), _D, _D]
        L.orc_end_time_default.restype = C.c_double
        L.orc_end_time_default.argtypes = [C.c_int, C.c_int]
        L.orc_exact_conservatives.argtypes = [C.c_int, C.c_int, _D, C.c_double, _D]
        L.orc_init_state.argtypes = [C.c_int, C.c_int, C.c_long, _D, C.c_double, _D]
        L.orc_fluid_flags.argtypes = [C.c_long, _D, C.c_int, _D, _U8]
        L.orc_interface_bcs.argtypes = [C.c_int, C.c_long, _I64, _I64, _D, _U8, _I32]
        L.orc_compute_rhs.argtypes = [C.c_int, C.c_long, C.c_long, _I64, _I64, _I32, _D, _D, _D, _U8, _D, _D, _D]
        L.orc_rk_stage.argtypes = [C.c_int, C.c_long, _U8, _D, C.c_double, _D, _D, _D]
        L.orc_choose_dt.restype = C.c_double
        L.orc_choose_dt.argtypes = [C.c_double] * 5
        L.orc_step.restype = C.c_double
        L.orc_step.argtypes = [C.c_int, C.c_long, C.c_long, _I64, _I64, _I32, _D, _D, _D, _U8, _U8, _D,
                               C.c_double, C.c_double, C.c_double, C.c_double, _D, _D, _D, _D]
        L.orc_error_norm.restype = C.c_double
        L.orc_error_norm.argtypes = [C.c_int, C.c_int, C.c_long, _D, _D, _U8, _D, C.c_double]
        L.orc_run.restype = C.c_int
        L.orc_run.argtypes = [C.c_int, C.c_int, C.c_long, C.c_double, C.c_double, C.c_int, _D, C.c_int, _D, _D, _D]
        L.orc_bench_threads.restype = C.c_double
        L.orc_bench_threads.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                        C.POINTER(C.c_uint64)]
        L.orc_run_threads.restype = None
        L.orc_run_threads.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _D, _D]
        L.orc_conservative2primitive.argtypes = [_D, _D]
        L.orc_primitive2conservative.argtypes = [_D, _D]
        L.orc_eval_splitting.argtypes = [_D, _D, _D, _D, _D]
        L.orc_eval_interface_bc_values.argtypes = [C.c_int, C.c_int, _D, _D, _D, _D]

    # ---- mesh -----------------------------------------------------------------------------
    def domain(self, problem, dim_in):
        dim = C.c_int(0)
        origin = np.zeros(3)
        length = C.c_double(0)
        self.lib.orc_domain_defaults(PROBLEMS[problem], dim_in, C.byref(dim), _p(origin, _D), C.byref(length))
        return dim.value, origin, length.value

    def uniform_mesh(self, dim, origin, length, n_cells_per_dir):
        level = self.lib.orc_level_for(length, n_cells_per_dir)
        nc, nf = C.c_long(0), C.c_long(0)
        self.lib.orc_uniform_counts(dim, level, C.byref(nc), C.byref(nf))
        nc, nf = nc.value, nf.value
        m = dict(dim=dim, level=level, n_side=1 << level, h=length / (1 << level),
                 origin=np.array(origin, dtype=np.float64), length=length,
                 owner=np.empty(nf, np.int64), neigh=np.empty(nf, np.int64), area=np.empty(nf),
                 normal=np.empty((nf, 3)), icentroid=np.empty((nf, 3)), volume=np.empty(nc),
                 size=np.empty(nc), ccentroid=np.empty((nc, 3)), cell_ijk=np.empty((nc, 3), np.int32))
        self.lib.orc_uniform_mesh(dim, _p(m["origin"], _D), length, level, _p(m["owner"], _I64), _p(m["neigh"], _I64),
                                  _p(m["area"], _D), _p(m["normal"], _D), _p(m["icentroid"], _D), _p(m["volume"], _D),
                                  _p(m["size"], _D), _p(m["ccentroid"], _D), _p(m["cell_ijk"], _I32))
        n = 1 << level
        m["box_dims"] = (n, n, n if dim == 3 else 1)
        return m

    def problem_mesh(self, problem, dim_in, n_cells_per_dir, boxes=None):
        """Mesh + flags + BC table exactly as src/main.cpp sets them up (serial build)."""
        dim, origin, length = self.domain(problem, dim_in)
        m = self.uniform_mesh(dim, origin, length, n_cells_per_dir)
        nc, nf = m["volume"].shape[0], m["owner"].shape[0]
        boxes = np.ascontiguousarray(boxes if boxes is not None else np.zeros((0, 6)), dtype=np.float64)
        fluid = np.empty(nc, np.uint8)
        self.lib.orc_fluid_flags(nc, _p(m["ccentroid"], _D), boxes.shape[0], _p(boxes, _D), _p(fluid, _U8))
        bc = np.empty(nf, np.int32)
        self.lib.orc_interface_bcs(PROBLEMS[problem], nf, _p(m["owner"], _I64), _p(m["neigh"], _I64),
                                   _p(m["icentroid"], _D), _p(fluid, _U8), _p(bc, _I32))
        m.update(problem=problem, fluid=fluid, solved=fluid.copy(), internal=np.ones(nc, np.uint8), bc=bc)
        return m

    def init_state(self, m, t=0.0):
        nc = m["volume"].shape[0]
        U = np.empty((nc, 5))
        self.lib.orc_init_state(PROBLEMS[m["problem"]], m["dim"], nc, _p(m["ccentroid"], _D), t, _p(U, _D))
        return U

    # ---- operators ------------------------------------------------------------------------
    def compute_rhs(self, m, U, solved=None):
        nc, nf = m["volume"].shape[0], m["owner"].shape[0]
        solved = m["solved"] if solved is None else solved
        U = np.ascontiguousarray(U)
        RHS = np.empty((nc, 5))
        me = C.c_double(0)
        self.lib.orc_compute_rhs(PROBLEMS[m["problem"]], nc, nf, _p(m["owner"], _I64), _p(m["neigh"], _I64),
                                 _p(m["bc"], _I32), _p(m["area"], _D), _p(m["normal"], _D), _p(m["icentroid"], _D),
                                 _p(solved, _U8), _p(U, _D), _p(RHS, _D), C.byref(me))
        return RHS, me.value

    def rk_stage(self, m, stage, dt, U, W, RHS, mask=None):
        nc = m["volume"].shape[0]
        mask = (m["solved"] & m["internal"]).astype(np.uint8) if mask is None else mask
        self.lib.orc_rk_stage(stage, nc, _p(mask, _U8), _p(m["volume"], _D), dt, _p(U, _D), _p(W, _D), _p(RHS, _D))

    def choose_dt(self, cfl, min_h, max_eig, t, t_max):
        return self.lib.orc_choose_dt(cfl, min_h, max_eig, t, t_max)

    def step(self, m, cfl, t, t_max, U, W, RHS):
        nc, nf = m["volume"].shape[0], m["owner"].shape[0]
        mask = (m["solved"] & m["internal"]).astype(np.uint8)
        me3 = np.zeros(3)
        dt = self.lib.orc_step(PROBLEMS[m["problem"]], nc, nf, _p(m["owner"], _I64), _p(m["neigh"], _I64),
                               _p(m["bc"], _I32), _p(m["area"], _D), _p(m["normal"], _D), _p(m["icentroid"], _D),
                               _p(m["solved"], _U8), _p(mask, _U8), _p(m["volume"], _D), cfl, float(m["size"].min()),
                               t, t_max, _p(U, _D), _p(W, _D), _p(RHS, _D), _p(me3, _D))
        return dt, me3

    def error_norm(self, m, U, t_max):
        nc = m["volume"].shape[0]
        U = np.ascontiguousarray(U)
        return self.lib.orc_error_norm(PROBLEMS[m["problem"]], m["dim"], nc, _p(m["ccentroid"], _D),
                                       _p(m["volume"], _D), None, _p(U, _D), t_max)

    def end_time(self, problem, dim):
        return self.lib.orc_end_time_default(PROBLEMS[problem], dim)

    def run(self, problem, dim_in, n, t_end=-1.0, cfl=0.45, boxes=None, max_steps=-1, want_state=False):
        dim, origin, length = self.domain(problem, dim_in)
        level = self.lib.orc_level_for(length, n)
        nc = (1 << level) ** dim
        boxes = np.ascontiguousarray(boxes if boxes is not None else np.zeros((0, 6)), dtype=np.float64)
        err, t = C.c_double(0), C.c_double(0)
        U = np.empty((nc, 5)) if want_state else None
        steps = self.lib.orc_run(PROBLEMS[problem], dim_in, n, t_end, cfl, boxes.shape[0], _p(boxes, _D), max_steps,
                                 C.byref(err), C.byref(t), _p(U, _D))
        return dict(steps=steps, error=err.value, t=t.value, U=U)

    def bench_threads(self, problem, dim, level, n_warmup, n_steps, n_threads, cfl=0.45):
        h = C.c_uint64(0)
        secs = self.lib.orc_bench_threads(PROBLEMS[problem], dim, level, n_warmup, n_steps, n_threads, cfl, C.byref(h))
        return secs, h.value


    def run_threads(self, problem, dim, level, n_steps, n_threads=0, cfl=0.45):
        """n_steps RK3 steps (no tMax clamp) on 2^level cells per side with the threaded loop (bitwise the serial
        one); returns (U0, U) in Morton cell order.  The checker for meshes the serial loop is too slow for."""
        n_threads = n_threads or (os.cpu_count() or 1)
        nc = (1 << level) ** dim
        U0, U = np.empty((nc, 5)), np.empty((nc, 5))
        self.lib.orc_run_threads(PROBLEMS[problem], dim, level, n_steps, n_threads, cfl, _p(U0, _D), _p(U, _D))
        return U0, U


def morton_to_lexicographic(n, dim=3):
    """perm with lexi_state[perm] = morton_state ... i.e. perm[m] = lexicographic index of the cell whose Morton id is m
    (x = lowest bit; n a power of two)."""
    m = np.arange(n ** dim, dtype=np.int64)
    bits = int(np.log2(n))
    coords = [np.zeros_like(m) for _ in range(3)]
    for b in range(bits):
        for a in range(dim):
            coords[a] |= ((m >> (dim * b + a)) & 1) << b
    return (coords[2] * n + coords[1]) * n + coords[0]


_oracle = None


def load():
    global _oracle
    if _oracle is None:
        if not os.path.exists(LIB_PATH):
            build()
        _oracle = Oracle(C.CDLL(LIB_PATH))
    return _oracle


def format_error(err):
    """Same text as src/main.cpp:572 (setprecision(12) << scientific)."""
    return "%.12e" % err
