"""ctypes declarations of the C-ABI in include/mmf_b200.h (one-to-one, no logic)."""
import ctypes as C
import os

N_FIELDS = 5

# status codes
OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED_ORDER, ERR_NO_DEVICE, ERR_NCCL, ERR_STATE = range(7)
# src/constants.hpp:58-62
BC_NONE, BC_FREE_FLOW, BC_REFLECTING, BC_WALL, BC_DIRICHLET = -1, 0, 1, 2, 3
FIELD_U, FIELD_W, FIELD_RHS = 0, 1, 2
PATH_GENERIC, PATH_UNIFORM = 0, 1
FLAG_FORCE_GENERIC, FLAG_ORDER_AXIS = 1, 2
NUMBERING_MORTON, NUMBERING_LEXICOGRAPHIC, NUMBERING_AXIS = 0, 1, 2


class MmfError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libmmf_b200 error {code}: {message}")
        self.code = code


class MeshDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("dim", C.c_int32),
        ("problem_type", C.c_int32),
        ("flags", C.c_uint32),
        ("reserved0", C.c_int32),
        ("n_cells", C.c_int64),
        ("n_interfaces", C.c_int64),
        ("interface_order", C.POINTER(C.c_int64)),
        ("n_interfaces_listed", C.c_int64),
        ("owner", C.POINTER(C.c_int64)),
        ("neigh", C.POINTER(C.c_int64)),
        ("bc", C.POINTER(C.c_int32)),
        ("area", C.POINTER(C.c_double)),
        ("normal", C.POINTER(C.c_double)),
        ("volume", C.POINTER(C.c_double)),
        ("solved", C.POINTER(C.c_uint8)),
        ("internal", C.POINTER(C.c_uint8)),
        ("dirichlet_info", C.c_double * N_FIELDS),
        ("cell_ijk", C.POINTER(C.c_int32)),
        ("box_dims", C.c_int32 * 3),
        ("global_dims", C.c_int32 * 3),
        ("box_offset", C.c_int32 * 3),
        ("reserved1", C.c_int32),
    ]


class UniformDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("problem_type", C.c_int32),
        ("flags", C.c_uint32),
        ("box_dims", C.c_int32 * 3),
        ("global_dims", C.c_int32 * 3),
        ("box_offset", C.c_int32 * 3),
        ("cell_numbering", C.c_int32),
        ("interface_numbering", C.c_int32),
        ("bc_side", C.c_int32 * 6),
        ("h", C.c_double),
        ("dirichlet_info", C.c_double * N_FIELDS),
        ("area", C.c_double),
        ("volume", C.c_double),
    ]


class Info(C.Structure):
    _fields_ = [
        ("path", C.c_int32),
        ("device", C.c_int32),
        ("sm_count", C.c_int32),
        ("cc_major", C.c_int32),
        ("cc_minor", C.c_int32),
        ("order_exact", C.c_int32),
        ("n_cells", C.c_int64),
        ("n_interfaces", C.c_int64),
        ("kernel_launches", C.c_int64),
        ("device_bytes", C.c_int64),
    ]


_P = C.c_void_p
_D = C.POINTER(C.c_double)

# name -> (restype, argtypes); must list every function declared in include/mmf_b200.h
SIGNATURES = {
    "mmf_create": (C.c_int, [C.POINTER(MeshDesc), C.c_int, C.POINTER(_P)]),
    "mmf_create_uniform": (C.c_int, [C.POINTER(UniformDesc), C.c_int, C.POINTER(_P)]),
    "mmf_destroy": (C.c_int, [_P]),
    "mmf_last_error": (C.c_char_p, [_P]),
    "mmf_get_info": (C.c_int, [_P, C.POINTER(Info)]),
    "mmf_device_count": (C.c_int, []),
    "mmf_set_state": (C.c_int, [_P, C.c_int, _P]),
    "mmf_get_state": (C.c_int, [_P, C.c_int, _P]),
    "mmf_get_primitives": (C.c_int, [_P, C.c_int, _P]),
    "mmf_compute_polynomials": (C.c_int, [_P, C.c_int]),
    "mmf_compute_rhs": (C.c_int, [_P, C.c_int, C.c_int, _D]),
    "mmf_compute_rhs_host": (C.c_int, [_P, _P, C.c_int, _P, _D]),
    "mmf_rk_stage": (C.c_int, [_P, C.c_int, C.c_double]),
    "mmf_step": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_double, _D, _D]),
    "mmf_run": (C.c_int, [_P, C.c_double, C.c_double, _D, C.c_double, C.c_int, C.POINTER(C.c_int)]),
    "mmf_comm_unique_id": (C.c_int, [_P]),
    "mmf_comm_init": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "mmf_comm_set_ghost_lists": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                           C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mmf_comm_set_box_neighbours": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "mmf_comm_ipc_export": (C.c_int, [_P, _P]),
    "mmf_comm_ipc_import": (C.c_int, [_P, _P]),
    "mmf_exchange": (C.c_int, [_P, C.c_int]),
    "mmf_allreduce_max": (C.c_int, [_P, _D]),
    "mmf_timer_start": (C.c_int, [_P]),
    "mmf_timer_stop": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "mmf_synchronize": (C.c_int, [_P]),
    "mmf_profile_begin": (C.c_int, [_P]),
    "mmf_profile_end": (C.c_int, [_P, _D, C.POINTER(C.c_int64)]),
    "mmf_flush_l2": (C.c_int, [_P]),
    "mmf_selftest_division": (C.c_int, [C.c_int, C.c_longlong, C.c_ulonglong, C.POINTER(C.c_ulonglong)]),
    "mmf_host_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "mmf_host_free": (C.c_int, [_P]),
}

_lib = None


def library_path():
    # MMF_LIB_PATH: development override (e.g. a build with different compiler flags)
    return os.environ.get("MMF_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libmmf_b200.so")


def load_library():
    """Load libmmf_b200.so (built in-tree by __graft_entry__.build()). Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise MmfError(-1, f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU or pure-Python fallback)")
    lib = C.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def device_count():
    return load_library().mmf_device_count()


def check(rc, ctx=None):
    if rc != OK:
        msg = load_library().mmf_last_error(ctx)
        raise MmfError(rc, msg.decode("utf-8", "replace") if msg else "")
