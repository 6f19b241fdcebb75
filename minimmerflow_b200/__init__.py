"""minimmerflow_b200 -- B200 (sm_100a) implementation of minimmerflow's explicit finite-volume
Euler residual-and-update path.

The product is the C-ABI shared library ``lib/libmmf_b200.so`` (``include/mmf_b200.h``) built from
``csrc/``.  This Python package is only the host-side mirror of that interface (ctypes), used by the
tests and the benchmark; it contains no numerical fallback: importing works anywhere, but every
compute call raises :class:`MmfError` when the library or a B200 is missing.
"""
from ._cabi import (  # noqa: F401
    MmfError,
    library_path,
    load_library,
    device_count,
    BC_NONE, BC_FREE_FLOW, BC_REFLECTING, BC_WALL, BC_DIRICHLET,
    FIELD_U, FIELD_W, FIELD_RHS,
    PATH_GENERIC, PATH_UNIFORM,
    FLAG_FORCE_GENERIC, FLAG_ORDER_AXIS,
    NUMBERING_MORTON, NUMBERING_LEXICOGRAPHIC, NUMBERING_AXIS,
)
from .solver import EulerSolver, selftest_division  # noqa: F401

__all__ = [
    "EulerSolver", "MmfError", "library_path", "load_library", "device_count", "selftest_division",
]
