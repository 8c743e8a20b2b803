"""CPU-only checks of the drop-in boundary: libllacuda.so loads without a GPU, exports every
symbol include/llacuda.h declares, and refuses to compute (loudly) when no device is present."""
import ctypes
import os
import re

import pytest

from lla_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "llacuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([a-z][a-z0-9_]*)\s*\(", src)
    return sorted({n for n in names if n.startswith("llacuda_") or n.endswith("_")})


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    return ctypes.CDLL(_lib.SO_PATH)


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_fortran_symbol_set(lib):
    for routine in ("gemm", "getrf", "getrs", "gesv", "potrf", "potrs"):
        assert hasattr(lib, f"d{routine}_"), routine
        assert hasattr(lib, f"llacuda_d{routine}"), routine


def test_ilp64_variant_exports_the_same_fortran_symbols(lib):
    """libllacuda64.so: the same symbols with 64-bit integer cells (LLA :int64, reference src/types.lisp:7-8)."""
    assert os.path.exists(_lib.SO64_PATH)
    lib64 = ctypes.CDLL(_lib.SO64_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib64, s)]
    assert not missing, missing
    import torch
    if torch.cuda.is_available():
        return
    # argument checking happens before any device work and uses the 64-bit cells: n = -1 in an int64 cell
    import numpy as np
    a = np.eye(4)
    ipiv = np.zeros(4, dtype=np.int64)
    info = ctypes.c_int64(7)
    four, minus = ctypes.c_int64(4), ctypes.c_int64(-1)
    lib64.dgetrf_(ctypes.byref(four), ctypes.byref(minus), a.ctypes.data_as(ctypes.c_void_p), ctypes.byref(four),
                  ipiv.ctypes.data_as(ctypes.c_void_p), ctypes.byref(info))
    assert info.value == -2


def test_no_silent_cpu_fallback(lib):
    """Without a CUDA device the library must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    rc = lib.llacuda_init()
    assert rc != 0
    lib.llacuda_last_error.restype = ctypes.c_char_p
    assert b"no CUDA device" in lib.llacuda_last_error() or b"sm_100a" in lib.llacuda_last_error()
    # LAPACK-style entry point: INFO carries the out-of-band device failure code, not a LAPACK code
    import numpy as np
    a = np.eye(4)
    ipiv = np.zeros(4, dtype=np.int32)
    info = ctypes.c_int(0)
    four = ctypes.c_int(4)
    lib.dgetrf_(ctypes.byref(four), ctypes.byref(four), a.ctypes.data_as(ctypes.c_void_p), ctypes.byref(four),
                ipiv.ctypes.data_as(ctypes.c_void_p), ctypes.byref(info))
    assert info.value <= -10000
    assert np.array_equal(a, np.eye(4))


def test_argument_checks_match_lapack(lib):
    """INFO = -k for the k-th illegal argument is decided before any device work (netlib order)."""
    import numpy as np
    a = np.eye(4)
    ipiv = np.zeros(4, dtype=np.int32)
    info = ctypes.c_int(0)
    I = lambda v: ctypes.byref(ctypes.c_int(v))
    P = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    lib.dgetrf_(I(-1), I(4), P(a), I(4), P(ipiv), ctypes.byref(info)); assert info.value == -1
    lib.dgetrf_(I(4), I(-1), P(a), I(4), P(ipiv), ctypes.byref(info)); assert info.value == -2
    lib.dgetrf_(I(4), I(4), P(a), I(3), P(ipiv), ctypes.byref(info)); assert info.value == -4
    c = lambda s: ctypes.byref(ctypes.c_char(s))
    lib.dpotrf_(c(b"X"), I(4), P(a), I(4), ctypes.byref(info)); assert info.value == -1
    lib.dpotrf_(c(b"U"), I(-2), P(a), I(4), ctypes.byref(info)); assert info.value == -2
    lib.dpotrf_(c(b"U"), I(4), P(a), I(2), ctypes.byref(info)); assert info.value == -4
    lib.dgetrs_(c(b"Q"), I(4), I(1), P(a), I(4), P(ipiv), P(a), I(4), ctypes.byref(info)); assert info.value == -1
    lib.dgetrs_(c(b"N"), I(4), I(1), P(a), I(4), P(ipiv), P(a), I(3), ctypes.byref(info)); assert info.value == -8
    lib.dgesv_(I(4), I(-1), P(a), I(4), P(ipiv), P(a), I(4), ctypes.byref(info)); assert info.value == -2
    lib.dgesv_(I(4), I(1), P(a), I(4), P(ipiv), P(a), I(1), ctypes.byref(info)); assert info.value == -7
    lib.dpotrs_(c(b"U"), I(4), I(1), P(a), I(1), P(a), I(4), ctypes.byref(info)); assert info.value == -5
    # zero-size problems return INFO = 0 without touching the device
    lib.dgetrf_(I(0), I(0), P(a), I(1), P(ipiv), ctypes.byref(info)); assert info.value == 0
    lib.dpotrf_(c(b"U"), I(0), P(a), I(1), ctypes.byref(info)); assert info.value == 0


def test_host_mirror_validates_like_lla():
    """mm / solve dimension checks happen in the host layer, before the FFI call
    (reference src/linear-algebra.lisp:44-47, :268)."""
    import numpy as np
    from lla_b200 import lla as L
    with pytest.raises(AssertionError):
        L.mm(np.ones((3, 2)), np.ones(3))
    with pytest.raises(ValueError):
        L.mm(np.ones(2), np.ones(3))
    with pytest.raises(L.LlaIncompatibleDimensions):
        L.solve(np.eye(3), np.ones(4))
    with pytest.raises(TypeError):
        L._as_memory(np.array([1j]), L.DOUBLE)   # complex -> real raises, tests/pinned-array.lisp:67
    assert L.common_float_type(np.zeros(1, np.float32), np.zeros(1, np.complex64)) == L.COMPLEX_SINGLE
    assert L.common_float_type(np.zeros(1, np.float32), np.zeros(1)) == L.DOUBLE
    assert L.common_float_type(np.zeros(1, dtype=np.int64)) == L.DOUBLE
    assert list(L.ipiv(np.array([2, 2], dtype=np.int32))) == [1, 0]
    assert L.permutations(np.array([2, 2, 3], dtype=np.int32)) == 1
