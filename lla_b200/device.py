"""Thin typed bindings of the llacuda_*_dev entry points (device pointers, column-major).

Torch is used only as the owner of device memory and streams: a column-major m x n matrix with
leading dimension ld is a torch tensor of shape (n, ld) whose row j is column j.
"""
from __future__ import annotations

import ctypes

from . import _lib

_i64 = ctypes.c_int64
_dbl = ctypes.c_double
_vp = ctypes.c_void_p
_chr = ctypes.c_char


def _sig(name, argtypes):
    f = getattr(_lib.lib(), name)
    f.argtypes = argtypes
    f.restype = ctypes.c_int
    return f


def use_torch_stream():
    """Enqueue every *_dev call on torch's current stream so torch.cuda.Event timing sees it."""
    import torch
    _lib.check(_sig("llacuda_set_stream", [_vp])(torch.cuda.current_stream().cuda_stream), "set_stream")


def colmajor(t):
    """(data_ptr, ld) of a torch tensor interpreted as column-major storage (shape (n, ld))."""
    assert t.is_cuda and t.is_contiguous()
    return _vp(t.data_ptr()), t.shape[-1]


def dgemm(opa, opb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri=0):
    f = _sig("llacuda_dsyrk_like_dev", [_chr, _chr, _i64, _i64, _i64, _dbl, _vp, _i64, _vp, _i64, _dbl, _vp, _i64, ctypes.c_int])
    _lib.check(f(opa.encode(), opb.encode(), m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri), "dgemm_dev")


def probe_fp64_peak(kind, iters=20000):
    tf, ms = ctypes.c_double(0), ctypes.c_float(0)
    f = _sig("llacuda_probe_fp64_peak", [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)])
    _lib.check(f(kind, iters, ctypes.byref(tf), ctypes.byref(ms)), "probe")
    return tf.value, ms.value


_ip = ctypes.c_void_p


def dgetrf(m, n, A, lda, ipiv, info):
    _lib.check(_sig("llacuda_dgetrf_dev", [_i64, _i64, _vp, _i64, _ip, _ip])(m, n, A, lda, ipiv, info), "dgetrf_dev")


def dgetrs(trans, n, nrhs, A, lda, ipiv, B, ldb):
    f = _sig("llacuda_dgetrs_dev", [_chr, _i64, _i64, _vp, _i64, _ip, _vp, _i64])
    _lib.check(f(trans.encode(), n, nrhs, A, lda, ipiv, B, ldb), "dgetrs_dev")


def dgesv(n, nrhs, A, lda, ipiv, B, ldb, info):
    f = _sig("llacuda_dgesv_dev", [_i64, _i64, _vp, _i64, _ip, _vp, _i64, _ip])
    _lib.check(f(n, nrhs, A, lda, ipiv, B, ldb, info), "dgesv_dev")


def dpotrf(uplo, n, A, lda, info):
    _lib.check(_sig("llacuda_dpotrf_dev", [_chr, _i64, _vp, _i64, _ip])(uplo.encode(), n, A, lda, info), "dpotrf_dev")


def dpotrs(uplo, n, nrhs, A, lda, B, ldb):
    f = _sig("llacuda_dpotrs_dev", [_chr, _i64, _i64, _vp, _i64, _vp, _i64])
    _lib.check(f(uplo.encode(), n, nrhs, A, lda, B, ldb), "dpotrs_dev")


def dtrsm(uplo, trans, diag, m, n, T, ldt, B, ldb):
    f = _sig("llacuda_dtrsm_dev", [_chr, _chr, _chr, _i64, _i64, _vp, _i64, _vp, _i64])
    _lib.check(f(uplo.encode(), trans.encode(), diag.encode(), m, n, T, ldt, B, ldb), "dtrsm_dev")


def dlaswp(ncols, A, lda, k1, k2, ipiv, forward=True):
    f = _sig("llacuda_dlaswp_dev", [_i64, _vp, _i64, _i64, _i64, _ip, ctypes.c_int])
    _lib.check(f(ncols, A, lda, k1, k2, ipiv, 1 if forward else 0), "dlaswp_dev")


def dtranspose(rows, cols, src, ldin, dst, ldout):
    f = _sig("llacuda_dtranspose_dev", [_i64, _i64, _vp, _i64, _vp, _i64])
    _lib.check(f(rows, cols, src, ldin, dst, ldout), "dtranspose_dev")


def zgemm(opa, opb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc):
    f = _sig("llacuda_zgemm_dev", [_chr, _chr, _i64, _i64, _i64, ctypes.POINTER(_dbl), _vp, _i64, _vp, _i64,
                                   ctypes.POINTER(_dbl), _vp, _i64])
    al = (_dbl * 2)(alpha.real, alpha.imag)
    be = (_dbl * 2)(beta.real, beta.imag)
    _lib.check(f(opa.encode(), opb.encode(), m, n, k, al, A, lda, B, ldb, be, C, ldc), "zgemm_dev")


_TYPE_CODE = {"float32": 0, "float64": 1, "complex64": 2, "complex128": 3, "int32": 4, "int64": 5}


def convert(in_dtype, out_dtype, transpose, rows, cols, src, ldin, dst, ldout):
    """llacuda_convert_dev: element-wise coercion (+ transpose) on the device, the replacement of LLA's
    copy-array-to-memory / transpose-matrix-to-memory (reference src/foreign-memory.lisp:168-255)."""
    f = _sig("llacuda_convert_dev", [ctypes.c_int, ctypes.c_int, ctypes.c_int, _i64, _i64, _vp, _i64, _vp, _i64])
    _lib.check(f(_TYPE_CODE[str(in_dtype)], _TYPE_CODE[str(out_dtype)], 1 if transpose else 0, rows, cols, src, ldin, dst, ldout),
               "convert_dev")


def dgesv_batched(n, nrhs, batch, A, strideA, ipiv, B, strideB, info, keep_lu=0):
    f = _sig("llacuda_dgesv_batched_dev", [ctypes.c_int, ctypes.c_int, _i64, _vp, _i64, _ip, _vp, _i64, _ip, ctypes.c_int])
    _lib.check(f(n, nrhs, batch, A, strideA, ipiv, B, strideB, info, keep_lu), "dgesv_batched_dev")


_flt = ctypes.c_float


def sgemm(opa, opb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc):
    f = _sig("llacuda_sgemm_dev", [_chr, _chr, _i64, _i64, _i64, _flt, _vp, _i64, _vp, _i64, _flt, _vp, _i64])
    _lib.check(f(opa.encode(), opb.encode(), m, n, k, alpha, A, lda, B, ldb, beta, C, ldc), "sgemm_dev")
