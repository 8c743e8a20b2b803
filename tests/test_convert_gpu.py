"""llacuda_convert_dev: element-wise coercion (+ transpose) on the device against NumPy's astype / .T -- the device-side
replacement of LLA's copy-array-to-memory / transpose-matrix-to-memory (reference src/foreign-memory.lisp:168-255;
marshalling rules tested by the reference in tests/pinned-array.lisp).  Conversions are exact or correctly rounded,
so the comparison is bit for bit."""
import ctypes

import numpy as np
import pytest

from lla_b200 import _lib, device as D, lla as L

pytestmark = pytest.mark.gpu

IN = [np.float32, np.float64, np.complex64, np.complex128, np.int32, np.int64]
OUT = [np.float32, np.float64, np.complex64, np.complex128]


def _roundtrip(a, out_dt, transpose):
    lib = _lib.lib()
    _lib.init(0)
    lib.llacuda_reset_stream()
    rows, cols = a.shape[1], a.shape[0]        # the row-major (r, c) array is the column-major c x r matrix
    src, dst = L._DevBuf(a.nbytes), L._DevBuf(a.size * np.dtype(out_dt).itemsize)
    _lib.check(lib.llacuda_memcpy_h2d(src.p, a.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(a.nbytes)), "h2d")
    D.convert(a.dtype, np.dtype(out_dt), transpose, rows, cols, src.p, rows, dst.p, cols if transpose else rows)
    out = np.empty((rows, cols) if transpose else (cols, rows), dtype=out_dt)
    _lib.check(lib.llacuda_memcpy_d2h(out.ctypes.data_as(ctypes.c_void_p), dst.p, ctypes.c_size_t(out.nbytes)), "d2h")
    return out


@pytest.mark.parametrize("in_dt", IN)
@pytest.mark.parametrize("out_dt", OUT)
@pytest.mark.parametrize("shape", [(1, 1), (5, 3), (33, 70), (257, 129)])
def test_convert_and_transpose(in_dt, out_dt, shape):
    rng = np.random.default_rng(shape[0] + shape[1])
    if np.issubdtype(in_dt, np.integer):
        a = rng.integers(-1000, 1000, shape).astype(in_dt)
    elif np.issubdtype(in_dt, np.complexfloating):
        a = (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(in_dt)
    else:
        a = rng.normal(size=shape).astype(in_dt)
    if np.issubdtype(in_dt, np.complexfloating) and not np.issubdtype(out_dt, np.complexfloating):
        with pytest.raises(_lib.LlacudaError):      # the reference refuses too (tests/pinned-array.lisp:67)
            _roundtrip(a, out_dt, False)
        _lib.lib().llacuda_clear_error()
        return
    assert np.array_equal(_roundtrip(a, out_dt, False), a.astype(out_dt))
    assert np.array_equal(_roundtrip(a, out_dt, True), a.astype(out_dt).T)


def test_device_resident_factorisations_take_integer_arrays():
    """integer arrays classify as double (reference src/types.lisp:69-80): the coercion runs on the device"""
    from oracle import lla_oracle as O
    rng = np.random.default_rng(1)
    a = rng.integers(0, 100, (300, 300))
    f = L.lu_device(a)
    r = O.lu(a, O.netlib())
    assert np.array_equal(f.ipiv_internal, r.ipiv_internal)
    assert np.max(np.abs(f.lu - r.lu)) <= 50 * 300 * np.finfo(float).eps * np.abs(r.lu).max()
    s = (a @ a.T + 300 * 100 * 100 * np.eye(300, dtype=np.int64))
    c = L.cholesky_device(s)
    assert np.max(np.abs(c.left @ c.left.T - s)) <= 20 * 300 * np.finfo(float).eps * np.abs(s).max()
