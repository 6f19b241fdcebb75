"""Host-side mirror of the reference's operator interface for the residual-and-update path.

``EulerSolver`` wraps one ``mmf_ctx`` handle.  Method names follow the reference's entry points:
``compute_polynomials`` / ``compute_rhs`` (reconstruction::computePolynomials, euler::computeRHS),
``rk_stage`` (the three inline loops of src/main.cpp:409-495), ``step`` / ``run`` (the time loop
src/main.cpp:377-506), ``exchange`` (GhostCommunicator::start/completeAllExchanges).
All numerical work happens inside libmmf_b200.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import _cabi as A


def _ptr(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


def _as(arr, dtype):
    a = np.ascontiguousarray(arr, dtype=dtype)
    return a


def selftest_division(n_samples=1 << 28, seed=12345, device=0):
    """Bitwise check of the kernels' shared-reciprocal division against IEEE `/` on the GPU."""
    lib = A.load_library()
    bad = C.c_ulonglong(0)
    A.check(lib.mmf_selftest_division(device, n_samples, seed, C.byref(bad)))
    return bad.value


def mesh_desc(mesh, flags=0, problem_type=0, dirichlet_info=None):
    """mmf_mesh_desc of a host mesh mapping (see EulerSolver.from_mesh) plus the arrays it points into, which
    the caller keeps alive for as long as the description is used."""
    d = A.MeshDesc()
    d.struct_size = C.sizeof(A.MeshDesc)
    d.dim = int(mesh["dim"])
    d.problem_type = int(problem_type)
    d.flags = int(flags)
    keep = {}
    keep["owner"] = _as(mesh["owner"], np.int64)
    keep["neigh"] = _as(mesh["neigh"], np.int64)
    keep["bc"] = _as(mesh["bc"], np.int32)
    keep["area"] = _as(mesh["area"], np.float64)
    keep["normal"] = _as(mesh["normal"], np.float64)
    keep["volume"] = _as(mesh["volume"], np.float64)
    keep["solved"] = _as(mesh["solved"], np.uint8)
    d.n_cells = keep["volume"].shape[0]
    d.n_interfaces = keep["owner"].shape[0]
    d.owner = _ptr(keep["owner"], C.c_int64)
    d.neigh = _ptr(keep["neigh"], C.c_int64)
    d.bc = _ptr(keep["bc"], C.c_int32)
    d.area = _ptr(keep["area"], C.c_double)
    d.normal = _ptr(keep["normal"], C.c_double)
    d.volume = _ptr(keep["volume"], C.c_double)
    d.solved = _ptr(keep["solved"], C.c_uint8)
    if mesh.get("internal") is not None:
        keep["internal"] = _as(mesh["internal"], np.uint8)
        d.internal = _ptr(keep["internal"], C.c_uint8)
    if mesh.get("interface_order") is not None:
        keep["order"] = _as(mesh["interface_order"], np.int64)
        d.interface_order = _ptr(keep["order"], C.c_int64)
        d.n_interfaces_listed = keep["order"].shape[0]
    if mesh.get("cell_ijk") is not None:
        keep["ijk"] = _as(mesh["cell_ijk"], np.int32)
        d.cell_ijk = _ptr(keep["ijk"], C.c_int32)
        for e in range(3):
            d.box_dims[e] = int(mesh["box_dims"][e])
            d.global_dims[e] = int(mesh["box_dims"][e])
            d.box_offset[e] = 0
    if dirichlet_info is not None:
        for k in range(A.N_FIELDS):
            d.dirichlet_info[k] = float(dirichlet_info[k])
    return d, keep


class EulerSolver:
    """One GPU, one mesh (or one rank's partition)."""

    def __init__(self, handle, n_cells, keepalive=None):
        self._lib = A.load_library()
        self._h = handle
        self.n_cells = int(n_cells)
        self._keepalive = keepalive

    # ---- construction ---------------------------------------------------------------------
    @classmethod
    def from_mesh(cls, mesh, device=0, flags=0, problem_type=0, dirichlet_info=None):
        """mesh: mapping with the arrays MeshGeometricalInfo caches (raw-id indexed):
        dim, owner, neigh, bc, area, normal[nf,3], volume, solved and optionally internal,
        interface_order, cell_ijk + box_dims (structured hint)."""
        lib = A.load_library()
        d, keep = mesh_desc(mesh, flags, problem_type, dirichlet_info)
        h = C.c_void_p()
        A.check(lib.mmf_create(C.byref(d), device, C.byref(h)))
        return cls(h, d.n_cells)

    @classmethod
    def uniform(cls, box_dims, h, bc_side, device=0, flags=0, problem_type=0,
                cell_numbering=A.NUMBERING_MORTON, interface_numbering=A.NUMBERING_MORTON,
                global_dims=None, box_offset=(0, 0, 0), dirichlet_info=None, area=0.0, volume=0.0):
        """Full uniform 3-D box from the compact description (mmf_create_uniform).  area / volume: the host's own
        interface area and cell volume (0 = h*h and h*h*h)."""
        lib = A.load_library()
        d = A.UniformDesc()
        d.struct_size = C.sizeof(A.UniformDesc)
        d.problem_type = int(problem_type)
        d.flags = int(flags)
        gd = global_dims if global_dims is not None else box_dims
        for e in range(3):
            d.box_dims[e] = int(box_dims[e])
            d.global_dims[e] = int(gd[e])
            d.box_offset[e] = int(box_offset[e])
        d.cell_numbering = int(cell_numbering)
        d.interface_numbering = int(interface_numbering)
        for s in range(6):
            d.bc_side[s] = int(bc_side[s])
        d.h = float(h)
        d.area = float(area)
        d.volume = float(volume)
        if dirichlet_info is not None:
            for k in range(A.N_FIELDS):
                d.dirichlet_info[k] = float(dirichlet_info[k])
        hnd = C.c_void_p()
        A.check(lib.mmf_create_uniform(C.byref(d), device, C.byref(hnd)))
        return cls(hnd, int(box_dims[0]) * int(box_dims[1]) * int(box_dims[2]))

    def close(self):
        if self._h is not None:
            self._lib.mmf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        A.check(rc, self._h)

    # ---- info -----------------------------------------------------------------------------
    def info(self):
        i = A.Info()
        self._ck(self._lib.mmf_get_info(self._h, C.byref(i)))
        return {f[0]: getattr(i, f[0]) for f in A.Info._fields_}

    # ---- state ----------------------------------------------------------------------------
    def set_state(self, field, aos):
        a = np.ascontiguousarray(aos, dtype=np.float64)
        if a.size != self.n_cells * A.N_FIELDS:
            raise ValueError("state must hold n_cells*5 doubles (AoS, raw cell order)")
        self._ck(self._lib.mmf_set_state(self._h, field, a.ctypes.data))

    def get_state(self, field, out=None):
        if out is None:
            out = np.empty((self.n_cells, A.N_FIELDS), dtype=np.float64)
        self._ck(self._lib.mmf_get_state(self._h, field, out.ctypes.data))
        return out

    def get_primitives(self, field=A.FIELD_U, out=None):
        """{p,u,v,w,T} per cell: utils::conservative2primitive (src/utils.cpp:48-63) evaluated on the device,
        what src/main.cpp:511-518 computes for the writer before every mesh.write()."""
        if out is None:
            out = np.empty((self.n_cells, A.N_FIELDS), dtype=np.float64)
        self._ck(self._lib.mmf_get_primitives(self._h, field, out.ctypes.data))
        return out

    def set_state_ptr(self, field, ptr):
        self._ck(self._lib.mmf_set_state(self._h, field, ptr))

    def get_state_ptr(self, field, ptr):
        self._ck(self._lib.mmf_get_state(self._h, field, ptr))

    # ---- operators (reference call shapes) ------------------------------------------------
    def compute_polynomials(self, field=A.FIELD_U):
        self._ck(self._lib.mmf_compute_polynomials(self._h, field))

    def compute_rhs(self, field=A.FIELD_U, order=1):
        m = C.c_double(0.0)
        self._ck(self._lib.mmf_compute_rhs(self._h, field, order, C.byref(m)))
        return m.value

    def compute_rhs_host(self, cons_aos, order=1, out=None):
        a = np.ascontiguousarray(cons_aos, dtype=np.float64)
        if out is None:
            out = np.empty((self.n_cells, A.N_FIELDS), dtype=np.float64)
        m = C.c_double(0.0)
        self._ck(self._lib.mmf_compute_rhs_host(self._h, a.ctypes.data, order, out.ctypes.data, C.byref(m)))
        return out, m.value

    def rk_stage(self, stage, dt):
        self._ck(self._lib.mmf_rk_stage(self._h, stage, dt))

    def step(self, cfl, min_cell_size, t, t_max):
        dt = C.c_double(0.0)
        me = (C.c_double * 3)()
        self._ck(self._lib.mmf_step(self._h, cfl, min_cell_size, t, t_max, C.byref(dt), me))
        return dt.value, [me[0], me[1], me[2]]

    def run(self, cfl, min_cell_size, t, t_max, max_steps=-1):
        tt = C.c_double(t)
        steps = C.c_int(0)
        self._ck(self._lib.mmf_run(self._h, cfl, min_cell_size, C.byref(tt), t_max, max_steps, C.byref(steps)))
        return tt.value, steps.value

    # ---- multi-GPU ------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        A.check(A.load_library().mmf_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, rank, n_ranks, unique_id):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        self._ck(self._lib.mmf_comm_init(self._h, rank, n_ranks, buf))

    def comm_set_box_neighbours(self, ranks6):
        arr = (C.c_int32 * 6)(*[int(r) for r in ranks6])
        self._ck(self._lib.mmf_comm_set_box_neighbours(self._h, arr))

    def comm_set_ghost_lists(self, neighbour_ranks, send_lists, recv_lists):
        n = len(neighbour_ranks)
        ranks = np.asarray(neighbour_ranks, dtype=np.int32)
        so = np.zeros(n + 1, dtype=np.int64)
        ro = np.zeros(n + 1, dtype=np.int64)
        for q in range(n):
            so[q + 1] = so[q] + len(send_lists[q])
            ro[q + 1] = ro[q] + len(recv_lists[q])
        si = np.concatenate([np.asarray(x, dtype=np.int64) for x in send_lists]) if n else np.zeros(0, np.int64)
        ri = np.concatenate([np.asarray(x, dtype=np.int64) for x in recv_lists]) if n else np.zeros(0, np.int64)
        self._ck(self._lib.mmf_comm_set_ghost_lists(self._h, n, _ptr(ranks, C.c_int32), _ptr(so, C.c_int64),
                                                    _ptr(si, C.c_int64), _ptr(ro, C.c_int64), _ptr(ri, C.c_int64)))

    IPC_BLOB_BYTES = 512

    def comm_ipc_export(self):
        buf = (C.c_char * self.IPC_BLOB_BYTES)()
        self._ck(self._lib.mmf_comm_ipc_export(self._h, buf))
        return bytes(buf)

    def comm_ipc_import(self, blobs):
        """blobs: the exported blobs of ALL ranks, in rank order."""
        data = b"".join(blobs)
        buf = (C.c_char * len(data)).from_buffer_copy(data)
        self._ck(self._lib.mmf_comm_ipc_import(self._h, buf))

    def exchange(self, field):
        self._ck(self._lib.mmf_exchange(self._h, field))

    def allreduce_max(self, value):
        v = C.c_double(value)
        self._ck(self._lib.mmf_allreduce_max(self._h, C.byref(v)))
        return v.value

    # ---- measurement ----------------------------------------------------------------------
    def timer_start(self):
        self._ck(self._lib.mmf_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float(0.0)
        self._ck(self._lib.mmf_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def profile_begin(self):
        self._ck(self._lib.mmf_profile_begin(self._h))

    def profile_end(self):
        ms = (C.c_double * 4)()
        n = (C.c_int64 * 4)()
        self._ck(self._lib.mmf_profile_end(self._h, ms, n))
        return list(ms), list(n)

    def synchronize(self):
        self._ck(self._lib.mmf_synchronize(self._h))

    def flush_l2(self):
        self._ck(self._lib.mmf_flush_l2(self._h))
