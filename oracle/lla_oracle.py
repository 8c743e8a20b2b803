"""oracle/lla_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of LLA's hot path *as LLA executes it*: Lisp row-major arrays
are marshalled exactly as src/pinned-array.lisp / src/foreign-memory.lisp do
(share untransposed, or copy transposed to column-major), ONE Fortran-ABI call
is issued with LLA's argument list, the result is marshalled back and INFO is
mapped to a condition (src/fortran-call.lisp:336-347).

The Fortran call goes to a *backend* library:
  * ``netlib()``   -- oracle/_build/liboracle_netlib.so, the scalar C restatement
                      of the netlib algorithms (symbols ``oracle_dgemm_`` ...).
  * ``openblas()`` -- the host OpenBLAS bundled with SciPy in this image
                      (symbols ``scipy_dgemm_`` ..., LP64, pthreads).  This is the
                      library a user of the reference would load through
                      ``:libraries`` (src/libraries.lisp:10-12); it is what
                      ``bench.py --impl reference`` and ``cpu_baseline`` time.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this
module.  Parity status: pinned against the reference's golden vectors
(tests/linear-algebra.lisp:9-29,89-115,366-385, tests/blas.lisp:19-46) in
tests/test_oracle.py; ipiv *values* for n>10, S/C/Z arithmetic, nrhs>1 and
rectangular LU are not pinned by any reference test and
are defined by agreement between the two backends.
"""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# LLA internal type codes, src/types.lisp:22-26
SINGLE, DOUBLE, COMPLEX_SINGLE, COMPLEX_DOUBLE = 0, 1, 2, 3
_DTYPES = {SINGLE: np.float32, DOUBLE: np.float64,
           COMPLEX_SINGLE: np.complex64, COMPLEX_DOUBLE: np.complex128}
_LETTER = {SINGLE: "s", DOUBLE: "d", COMPLEX_SINGLE: "c", COMPLEX_DOUBLE: "z"}


class LapackInvalidArgument(Exception):
    """src/conditions.lisp:27-33 -- INFO<0, slot POSITION."""
    def __init__(self, position):
        super().__init__(f"LAPACK: argument {position} illegal")
        self.position = position


class LapackFailure(Exception):
    """src/conditions.lisp:35-43 -- INFO>0, slot INFO."""
    def __init__(self, info):
        super().__init__(f"LAPACK failure, INFO={info}")
        self.info = info


class LlaIncompatibleDimensions(Exception):
    """src/conditions.lisp:45-47."""


def array_float_type(a) -> int:
    """src/types.lisp:69-80: element type -> internal code; integers/rationals
    classify as double."""
    dt = np.asarray(a).dtype
    if dt == np.float32:
        return SINGLE
    if dt == np.complex64:
        return COMPLEX_SINGLE
    if dt == np.complex128:
        return COMPLEX_DOUBLE
    return DOUBLE


def common_float_type(*arrays) -> int:
    """src/types.lisp:82-89: LOGIOR of the codes."""
    code = 0
    for a in arrays:
        code |= array_float_type(a)
    return code


class Backend:
    """A library exporting LLA's Fortran symbols under prefix+name."""

    def __init__(self, lib: ctypes.CDLL, prefix: str, name: str):
        self.lib, self.prefix, self.name = lib, prefix, name

    def fn(self, tcode: int, name: str):
        # src/fortran-call.lisp:405-416: letter + name + "_"
        f = getattr(self.lib, f"{self.prefix}{_LETTER[tcode]}{name}_")
        f.restype = None
        return f


def build_netlib() -> str:
    so = os.path.join(_HERE, "_build", "liboracle_netlib.so")
    src = [os.path.join(_HERE, f) for f in ("netlib_ref.c", "netlib_templ.h")]
    if (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


_NETLIB = None
_OPENBLAS = None


def netlib() -> Backend:
    global _NETLIB
    if _NETLIB is None:
        _NETLIB = Backend(ctypes.CDLL(build_netlib()), "oracle_", "netlib-restatement")
    return _NETLIB


def openblas(threads: int | None = None) -> Backend:
    """Host OpenBLAS (scipy wheel, LP64).  BASELINE.md section 2."""
    global _OPENBLAS
    if _OPENBLAS is None:
        import scipy
        pat = os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas-*.so")
        hits = glob.glob(pat)
        if not hits:
            raise RuntimeError("host OpenBLAS (scipy.libs/libscipy_openblas-*.so) not found")
        _OPENBLAS = Backend(ctypes.CDLL(hits[0]), "scipy_", "openblas-scipy-lp64")
    if threads is not None:
        _OPENBLAS.lib.scipy_openblas_set_num_threads(int(threads))
    return _OPENBLAS


def openblas_info() -> dict:
    be = openblas()
    be.lib.scipy_openblas_get_corename.restype = ctypes.c_char_p
    be.lib.scipy_openblas_get_config.restype = ctypes.c_char_p
    return {"corename": be.lib.scipy_openblas_get_corename().decode(),
            "threads": int(be.lib.scipy_openblas_get_num_threads()),
            "config": be.lib.scipy_openblas_get_config().decode()}


# ---------------------------------------------------------------- marshalling
def _int(v):
    return ctypes.byref(ctypes.c_int(int(v)))


def _chr(c):
    # src/foreign-memory.lisp:104-111: one byte holding the character code
    return ctypes.byref(ctypes.c_char(c.encode()))


def _atom(v, tcode):
    # src/foreign-memory.lisp:94-102, &atom src/fortran-call.lisp:160-162
    a = np.array([v], dtype=_DTYPES[tcode])
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def copy_array_to_memory(a, tcode) -> np.ndarray:
    """src/foreign-memory.lisp:168-182: element-wise coercing copy, row-major order."""
    a = np.asarray(a)
    if np.iscomplexobj(a) and tcode in (SINGLE, DOUBLE):
        raise TypeError("cannot coerce complex to real")  # tests/pinned-array.lisp:67
    return np.array(a, dtype=_DTYPES[tcode], order="C", copy=True)


def transpose_matrix_to_memory(a, tcode) -> np.ndarray:
    """src/foreign-memory.lisp:211-230: row-major matrix -> column-major image.
    Returned as a C-contiguous array of shape (ncol, nrow) holding the same bytes."""
    a = np.asarray(a)
    if a.ndim == 1:
        return copy_array_to_memory(a, tcode)
    if np.iscomplexobj(a) and tcode in (SINGLE, DOUBLE):
        raise TypeError("cannot coerce complex to real")
    return np.array(a.T, dtype=_DTYPES[tcode], order="C", copy=True)


def transpose_matrix_from_memory(mem: np.ndarray, shape) -> np.ndarray:
    """src/foreign-memory.lisp:232-255: column-major image -> new row-major matrix."""
    if len(shape) == 1:
        return mem.reshape(shape).copy()
    return np.ascontiguousarray(mem.reshape(shape[1], shape[0]).T)


def dimensions_as_matrix(a, orientation):
    """src/linear-algebra.lisp:18-26."""
    a = np.asarray(a)
    if a.ndim == 1:
        return (1, a.shape[0], True) if orientation == "row" else (a.shape[0], 1, True)
    if a.ndim == 2:
        return a.shape[0], a.shape[1], False
    raise ValueError("rank must be 1 or 2")


def _check_info(info, condition=LapackFailure):
    """src/fortran-call.lisp:336-347."""
    if info < 0:
        raise LapackInvalidArgument(-info)
    if info > 0 and condition is not None:
        raise condition(info)


# ------------------------------------------------------------------- hot path
def mm(a, b, be: Backend):
    """src/linear-algebra.lisp:39-54.  C = A B on row-major arrays through
    GEMM('N','N', b1, a0, b0, 1, b, b1, a, a1, 0, c, b1)."""
    a, b = np.asarray(a), np.asarray(b)
    t = common_float_type(a, b)
    a0, a1, av = dimensions_as_matrix(a, "row")
    b0, b1, bv = dimensions_as_matrix(b, "column")
    if av and bv:
        raise ValueError("MM called with two vectors.")
    if a1 != b0:
        raise AssertionError("Incompatible dimensions.")
    am, bm = copy_array_to_memory(a, t), copy_array_to_memory(b, t)  # shared in place when types match
    c = np.zeros(a0 * b1 if (av or bv) else (a0, b1), dtype=_DTYPES[t])
    one, pone = _atom(1, t)
    zero, pzero = _atom(0, t)
    be.fn(t, "gemm")(_chr("N"), _chr("N"), _int(b1), _int(a0), _int(b0), pone, _ptr(bm), _int(b1),
                     _ptr(am), _int(a1), pzero, _ptr(c), _int(b1))
    return c


def gemm(alpha, a, b, beta, c, be: Backend, transpose_a=False, transpose_b=False,
         m=None, n=None, k=None, lda=None, ldb=None, ldc=None):
    """src/blas.lisp:7-63 (gemm!).  C is updated in place and returned."""
    a, b = np.asarray(a), np.asarray(b)
    t = common_float_type(a, b)
    m = c.shape[0] if m is None else m
    n = c.shape[1] if n is None else n
    k = (a.shape[0] if transpose_a else a.shape[1]) if k is None else k
    lda = a.shape[1] if lda is None else lda
    ldb = b.shape[1] if ldb is None else ldb
    ldc = c.shape[1] if ldc is None else ldc
    if transpose_a:
        assert m <= lda and lda * k <= a.size
    else:
        assert k <= lda and m * lda <= a.size
    if transpose_b:
        assert k <= ldb and ldb * n <= b.size
    else:
        assert n <= ldb and k * ldb <= b.size
    assert n <= ldc and m * ldc <= c.size
    am, bm = copy_array_to_memory(a, t), copy_array_to_memory(b, t)
    cm = c if (c.dtype == _DTYPES[t] and c.flags.c_contiguous) else copy_array_to_memory(c, t)
    al, pal = _atom(alpha, t)
    bt, pbt = _atom(beta, t)
    be.fn(t, "gemm")(_chr("C" if transpose_b else "N"), _chr("C" if transpose_a else "N"),
                     _int(n), _int(m), _int(k), pal, _ptr(bm), _int(ldb), _ptr(am), _int(lda),
                     pbt, _ptr(cm), _int(ldc))
    if cm is not c:
        c[...] = cm  # src/pinned-array.lisp:90-103 copy-back
    return c


class LU:
    """src/factorizations.lisp:47-52 + ipiv-mixin :7-13."""

    def __init__(self, lu, ipiv_internal):
        self.lu, self.ipiv_internal = lu, ipiv_internal


def lu(a, be: Backend) -> LU:
    """src/linear-algebra.lisp:225-234: GETRF on a transposed copy, INFO>0 ignored."""
    a = np.asarray(a)
    t = common_float_type(a)
    a0, a1 = a.shape
    mem = transpose_matrix_to_memory(a, t)
    ipiv = np.zeros(min(a0, a1), dtype=np.int32)
    info = ctypes.c_int(0)
    be.fn(t, "getrf")(_int(a0), _int(a1), _ptr(mem), _int(a0), _ptr(ipiv), ctypes.byref(info))
    _check_info(info.value, None)
    return LU(transpose_matrix_from_memory(mem, (a0, a1)), ipiv)


def solve_lu(f: LU, b, be: Backend):
    """src/linear-algebra.lisp:264-276: GETRS('N', ...)."""
    b = np.asarray(b)
    lu0, lu1 = f.lu.shape
    b0, b1, _ = dimensions_as_matrix(b, "column")
    if not (lu0 == lu1 == b0):
        raise LlaIncompatibleDimensions()
    t = common_float_type(f.lu, b)
    amem = transpose_matrix_to_memory(f.lu, t)
    bmem = transpose_matrix_to_memory(b, t)
    info = ctypes.c_int(0)
    be.fn(t, "getrs")(_chr("N"), _int(lu0), _int(b1), _ptr(amem), _int(lu0), _ptr(f.ipiv_internal),
                      _ptr(bmem), _int(lu0), ctypes.byref(info))
    _check_info(info.value)
    return transpose_matrix_from_memory(bmem, b.shape)


def solve(a, b, be: Backend):
    """src/linear-algebra.lisp:278-289: GESV, ipiv scratch discarded."""
    a, b = np.asarray(a), np.asarray(b)
    a0, a1 = a.shape
    b0, b1, _ = dimensions_as_matrix(b, "column")
    if not (a0 == a1 == b0):
        raise LlaIncompatibleDimensions()
    t = common_float_type(a, b)
    amem = transpose_matrix_to_memory(a, t)
    bmem = transpose_matrix_to_memory(b, t)
    ipiv = np.zeros(a0, dtype=np.int32)
    info = ctypes.c_int(0)
    be.fn(t, "gesv")(_int(a0), _int(b1), _ptr(amem), _int(a0), _ptr(ipiv), _ptr(bmem), _int(a0),
                     ctypes.byref(info))
    _check_info(info.value)
    return transpose_matrix_from_memory(bmem, b.shape)


class Cholesky:
    """src/factorizations.lisp:130-136: left (lower-triangular) square root."""

    def __init__(self, left):
        self.left = left


def cholesky(a, be: Backend) -> Cholesky:
    """src/linear-algebra.lisp:666-674: POTRF('U') on an UNtransposed copy of the
    row-major elements => row-major lower factor; strict upper masked by the
    lower-triangular-matrix wrapper."""
    a = np.asarray(a)
    a0, a1 = a.shape
    assert a0 == a1
    t = common_float_type(a)
    mem = copy_array_to_memory(a, t)
    info = ctypes.c_int(0)
    be.fn(t, "potrf")(_chr("U"), _int(a0), _ptr(mem), _int(a0), ctypes.byref(info))
    _check_info(info.value)
    return Cholesky(np.tril(mem))


def solve_cholesky(c: Cholesky, b, be: Backend):
    """src/linear-algebra.lisp:291-301: POTRS('U') with L shared untransposed."""
    b = np.asarray(b)
    a = c.left
    a0, a1 = a.shape
    b0, b1, _ = dimensions_as_matrix(b, "column")
    if not (a0 == a1 == b0):
        raise LlaIncompatibleDimensions()
    t = common_float_type(a, b)
    amem = copy_array_to_memory(a, t)
    bmem = transpose_matrix_to_memory(b, t)
    info = ctypes.c_int(0)
    be.fn(t, "potrs")(_chr("U"), _int(a0), _int(b1), _ptr(amem), _int(a0), _ptr(bmem), _int(b0),
                      ctypes.byref(info))
    _check_info(info.value)
    return transpose_matrix_from_memory(bmem, b.shape)


# ------------------------------------------------ "next" rows
def mm_hermitian(a, transpose_left: bool, be: Backend):
    """src/linear-algebra.lisp:62-76: (mm a t) = A A^T, (mm t a) = A^T A through SYRK('U') on the untransposed
    row-major buffer (its column-major view is A^T, hence the swapped N / C); the result is wrapped as a
    hermitian-matrix, which reads the row-major lower triangle = the column-major upper one SYRK wrote."""
    a = np.asarray(a)
    a0, a1 = a.shape
    dim_c, other = (a1, a0) if transpose_left else (a0, a1)
    t = common_float_type(a)
    amem = copy_array_to_memory(a, t)
    c = np.zeros((dim_c, dim_c), dtype=_DTYPES[t])
    (_k1, one), (_k0, zero) = _atom(1, t), _atom(0, t)
    # ("syrk" "herk"): complex arrays go to HERK, which reads the leading real part of the alpha / beta cells
    name = "herk" if t in (COMPLEX_SINGLE, COMPLEX_DOUBLE) else "syrk"
    be.fn(t, name)(_chr("U"), _chr("N" if transpose_left else "C"), _int(dim_c), _int(other), one,
                   _ptr(amem), _int(a1), zero, _ptr(c), _int(dim_c))
    low = np.tril(c)
    return low + np.tril(c, -1).conj().T   # hermitian-matrix: the lower triangle defines the matrix


def solve_hermitian(a, b, be: Backend):
    """src/linear-algebra.lisp:303-315: POSV('U') on a forced copy of the row-major elements (only the row-major
    lower triangle is read), B transposed in and out."""
    a, b = np.asarray(a), np.asarray(b)
    a0, a1 = a.shape
    b0, b1, _ = dimensions_as_matrix(b, "column")
    if not (a0 == a1 == b0):
        raise LlaIncompatibleDimensions()
    t = common_float_type(a, b)
    amem = copy_array_to_memory(a, t)
    bmem = transpose_matrix_to_memory(b, t)
    info = ctypes.c_int(0)
    be.fn(t, "posv")(_chr("U"), _int(a0), _int(b1), _ptr(amem), _int(a0), _ptr(bmem), _int(b0), ctypes.byref(info))
    _check_info(info.value)
    return transpose_matrix_from_memory(bmem, b.shape)


def trsm(a, a_upper: bool, b, be: Backend):
    """src/linear-algebra.lisp:333-347 (trsm%): solve A X = B for triangular A with TRSM('R', L/U swapped, 'N', 'N')
    on the untransposed row-major buffers (X^T A^T = B^T in the column-major view)."""
    a, b = np.asarray(a), np.asarray(b)
    a0, a1 = a.shape
    b0, b1, _ = dimensions_as_matrix(b, "column")
    if not (a0 == a1 == b0):
        raise LlaIncompatibleDimensions()
    t = common_float_type(a, b)
    amem = copy_array_to_memory(a, t)
    xmem = copy_array_to_memory(b, t).reshape(b0, b1)
    _k1, one = _atom(1, t)
    be.fn(t, "trsm")(_chr("R"), _chr("L" if a_upper else "U"), _chr("N"), _chr("N"), _int(b1), _int(b0), one,
                     _ptr(amem), _int(a1), _ptr(xmem), _int(b1))
    return xmem.reshape(b.shape)


def invert(a, be: Backend, kind="array"):
    """src/linear-algebra.lisp:352-409 (both back ends export the LAPACK inverses): kind = "array"
    ((invert (lu a)) = GETRF + GETRI with workspace query), "cholesky" (a = row-major lower factor, POTRI 'U'),
    "upper" / "lower" (TRTRI with L / U swapped)."""
    a = np.asarray(a)
    n = a.shape[0]
    t = common_float_type(a)
    info = ctypes.c_int(0)
    if kind in ("upper", "lower"):
        mem = copy_array_to_memory(a, t)
        be.fn(t, "trtri")(_chr("L" if kind == "upper" else "U"), _chr("N"), _int(n), _ptr(mem), _int(n), ctypes.byref(info))
        _check_info(info.value)
        return np.triu(mem) if kind == "upper" else np.tril(mem)
    if kind == "cholesky":
        mem = copy_array_to_memory(a, t)
        be.fn(t, "potri")(_chr("U"), _int(n), _ptr(mem), _int(n), ctypes.byref(info))
        _check_info(info.value)
        low = np.tril(mem)
        return low + np.tril(mem, -1).T
    f = lu(a, be)
    mem = transpose_matrix_to_memory(f.lu, t)
    query = np.zeros(1, dtype=_DTYPES[t])
    be.fn(t, "getri")(_int(n), _ptr(mem), _int(n), _ptr(f.ipiv_internal), _ptr(query), _int(-1), ctypes.byref(info))
    work = np.zeros(max(1, int(query[0].real)), dtype=_DTYPES[t])
    be.fn(t, "getri")(_int(n), _ptr(mem), _int(n), _ptr(f.ipiv_internal), _ptr(work), _int(work.size), ctypes.byref(info))
    _check_info(info.value)
    return transpose_matrix_from_memory(mem, (n, n))


# ------------------------------------------------ factorization post-processing
def ipiv(ipiv_internal) -> np.ndarray:
    """src/factorizations.lisp:15-28: swap sequence -> 0-based permutation."""
    n = len(ipiv_internal)
    it = np.arange(n, dtype=np.int64)
    for index in range(n):
        p = int(ipiv_internal[index]) - 1
        if p != index:
            it[index], it[p] = it[p], it[index]
    return it


def ipiv_inverse(ipiv_internal) -> np.ndarray:
    """src/factorizations.lisp:30-43."""
    p = ipiv(ipiv_internal)
    inv = np.zeros(len(p), dtype=np.int64)
    inv[p] = np.arange(len(p))
    return inv


def lu_u(f: LU) -> np.ndarray:
    """src/factorizations.lisp:54-56."""
    return np.triu(f.lu)


def lu_l(f: LU) -> np.ndarray:
    """src/factorizations.lisp:58-65: unit diagonal forced."""
    l = np.tril(f.lu).copy()
    for i in range(min(l.shape)):
        l[i, i] = 1
    return l


def permutations(ipiv_internal) -> int:
    """src/factorizations.lisp:143-151."""
    return int(sum(1 for i, p in enumerate(ipiv_internal, start=1) if p != i))
