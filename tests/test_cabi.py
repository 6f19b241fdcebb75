"""The C-ABI library loads, exports every symbol include/mmf_b200.h declares, and fails loudly
(no CPU fallback) when there is no B200."""
import ctypes as C
import os
import re

import pytest

from common import HERE

ROOT = os.path.dirname(HERE)


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "mmf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mmf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(mmf):
    from minimmerflow_b200 import _cabi
    lib = mmf.load_library()
    names = _declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _cabi.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_cabi.SIGNATURES) == names


def test_struct_sizes_match_header_layout(mmf):
    from minimmerflow_b200 import _cabi
    # 8-byte aligned C layout of the two descriptors (guards against silent ABI drift)
    assert C.sizeof(_cabi.MeshDesc) == 8 + 16 + 16 + 16 + 8 * 5 + 8 * 3 + 40 + 8 + 12 * 3 + 4
    assert C.sizeof(_cabi.UniformDesc) % 8 == 0


def test_uniform_descriptor_sizes_accepted(mmf):
    """mmf_uniform_desc grew by the optional `area` / `volume` fields: a caller built against the shorter struct
    (struct_size = offset of `area`) is still served, any other size is an ABI mismatch -- decided before a device is
    looked for, so the check runs without a GPU (where the accepted sizes then fail with MMF_ERR_NO_DEVICE)."""
    from minimmerflow_b200 import _cabi
    lib = mmf.load_library()
    full, legacy = C.sizeof(_cabi.UniformDesc), _cabi.UniformDesc.area.offset
    assert legacy == full - 16
    for size, ok in ((full, True), (legacy, True), (full - 8, False), (full + 8, False), (0, False)):
        d = _cabi.UniformDesc()
        d.struct_size = size
        for e in range(3):
            d.box_dims[e] = d.global_dims[e] = 4
        d.h = 1.0
        h = C.c_void_p()
        rc = lib.mmf_create_uniform(C.byref(d), 0, C.byref(h))
        if h.value:
            lib.mmf_destroy(h)
        if ok:
            assert rc in (0, 4), (size, rc)        # MMF_OK on a B200, MMF_ERR_NO_DEVICE without one
        else:
            assert rc == 1, (size, rc)             # MMF_ERR_INVALID
            assert b"ABI mismatch" in lib.mmf_last_error(None)


def test_no_gpu_means_loud_failure(mmf, oracle):
    if mmf.device_count() > 0:
        pytest.skip("a B200 is present")
    m = oracle.problem_mesh("vortex_xy", 2, 8)
    with pytest.raises(mmf.MmfError) as e:
        mmf.EulerSolver.from_mesh(m)
    assert e.value.code == 4 and "no CPU fallback" in str(e.value)
    with pytest.raises(mmf.MmfError):
        mmf.EulerSolver.uniform((8, 8, 8), 1.0, [0] * 6)
    with pytest.raises(mmf.MmfError):
        mmf.selftest_division(1000)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "minimmerflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower().replace("test-only cpu oracle", ""), f
