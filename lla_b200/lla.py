"""Host-side mirror of LLA's user API for the hot path, on NumPy row-major arrays.

There is no Common Lisp in the build image, so this module plays the role of the Lisp layers
L7..L3 (reference src/linear-algebra.lisp, src/blas.lisp, src/factorizations.lisp,
src/fortran-call.lisp, src/pinned-array.lisp, src/foreign-memory.lisp) with the same names,
argument meaning and error behaviour, and issues *exactly* the Fortran-ABI calls LLA issues --
but into libllacuda.so.  It is what the parity tests drive; the Lisp glue that does the same from
SBCL lives under lisp/ (see INTEGRATION.md).

Lisp arrays are row-major; "share" means the NumPy buffer is passed untouched (what SBCL does by
pinning, src/pinned-array.lisp:124-141), "transpose" means a column-major copy
(src/foreign-memory.lisp:211-255).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

# internal type codes, reference src/types.lisp:22-26
SINGLE, DOUBLE, COMPLEX_SINGLE, COMPLEX_DOUBLE = 0, 1, 2, 3
_DTYPES = {SINGLE: np.float32, DOUBLE: np.float64, COMPLEX_SINGLE: np.complex64, COMPLEX_DOUBLE: np.complex128}
_LETTER = {SINGLE: "s", DOUBLE: "d", COMPLEX_SINGLE: "c", COMPLEX_DOUBLE: "z"}


# ------------------------------------------------------------------ conditions (src/conditions.lisp)
class LapackInvalidArgument(Exception):
    """reference src/conditions.lisp:27-33 (INFO < 0), slot POSITION."""

    def __init__(self, position):
        super().__init__(f"LAPACK: argument {position} illegal")
        self.position = position


class LapackFailure(Exception):
    """reference src/conditions.lisp:35-43 (INFO > 0), slot INFO."""

    def __init__(self, info):
        super().__init__(f"LAPACK failure, INFO={info}")
        self.info = info


class LapackSingularMatrix(LapackFailure):
    """reference src/conditions.lisp:41-43."""


class LlaIncompatibleDimensions(Exception):
    """reference src/conditions.lisp:45-47."""


# ------------------------------------------------------------------ types (src/types.lisp)
def array_float_type(a) -> int:
    """reference src/types.lisp:69-80; integer / rational element types classify as double."""
    dt = np.asarray(a).dtype
    if dt == np.float32:
        return SINGLE
    if dt == np.complex64:
        return COMPLEX_SINGLE
    if dt == np.complex128:
        return COMPLEX_DOUBLE
    return DOUBLE


def common_float_type(*arrays) -> int:
    """reference src/types.lisp:82-89: LOGIOR of the codes."""
    code = 0
    for a in arrays:
        code |= array_float_type(a)
    return code


# ------------------------------------------------------------------ marshalling (L3/L4)
# ------------------------------------------------------------------ efficiency warnings (reference src/conditions.lisp:49-74)
class LlaEfficiencyWarning(UserWarning):
    """lla-efficiency-warning"""


class LlaEfficiencyWarningArrayType(LlaEfficiencyWarning):
    """lla-efficiency-warning-array-type: an array without a specialised element type had to be checked elementwise."""


class LlaEfficiencyWarningArrayConversion(LlaEfficiencyWarning):
    """lla-efficiency-warning-array-conversion: an array had to be copied / converted to the call's element type --
    with the device back end that also means its upload could not be shared with the caller's buffer."""


# the toggles of src/conditions.lisp:52-66 (off by default, as in the reference)
efficiency_warning_array_type = False
efficiency_warning_array_conversion = False
# below this matrix order a call stays on the host LAPACK in the Lisp glue (lisp/llacuda-glue.lisp *llacuda-min-dimension*,
# SURVEY.md section 5 :cuda-min-n): the PCIe round trip costs more than the host BLAS call.  The mirror reports it.
min_dimension = 256


def device_worthwhile(*dims) -> bool:
    """True when every given dimension reaches `min_dimension` (the Lisp glue routes smaller calls to the host library)."""
    return all(int(d) >= min_dimension for d in dims)


def _efficiency_warning_type(a):
    if efficiency_warning_array_type:
        import warnings
        warnings.warn(f"Efficiency warning: had to check a {type(a).__name__} of dtype {getattr(a, 'dtype', None)} elementwise.",
                      LlaEfficiencyWarningArrayType, stacklevel=3)


def _efficiency_warning_conversion(a, dtype):
    if efficiency_warning_array_conversion:
        import warnings
        warnings.warn(f"Efficiency warning: array of dtype {a.dtype} had to be converted to {np.dtype(dtype)} before the upload.",
                      LlaEfficiencyWarningArrayConversion, stacklevel=3)


def _as_memory(a, tcode) -> np.ndarray:
    """share when the array already has the target element type and is a simple (contiguous)
    array, else element-wise coercing copy (src/pinned-array.lisp:124-128, src/foreign-memory.lisp:168-182)."""
    a = np.asarray(a)
    if np.iscomplexobj(a) and tcode in (SINGLE, DOUBLE):
        raise TypeError("cannot coerce a complex array to a real type")
    if a.dtype == _DTYPES[tcode] and a.flags.c_contiguous:
        return a
    if a.dtype == object:
        _efficiency_warning_type(a)
    if a.dtype != _DTYPES[tcode]:
        _efficiency_warning_conversion(a, _DTYPES[tcode])
    return np.array(a, dtype=_DTYPES[tcode], order="C", copy=True)


def _transposed_to_memory(a, tcode) -> np.ndarray:
    """reference src/foreign-memory.lisp:211-230: row-major matrix -> column-major image."""
    a = np.asarray(a)
    if np.iscomplexobj(a) and tcode in (SINGLE, DOUBLE):
        raise TypeError("cannot coerce a complex array to a real type")
    if a.dtype == object:
        _efficiency_warning_type(a)
    if a.dtype != _DTYPES[tcode]:
        _efficiency_warning_conversion(a, _DTYPES[tcode])
    if a.ndim == 1:
        return np.array(a, dtype=_DTYPES[tcode], order="C", copy=True)
    return np.array(a.T, dtype=_DTYPES[tcode], order="C", copy=True)


def _transposed_from_memory(mem, shape) -> np.ndarray:
    """reference src/foreign-memory.lisp:232-255."""
    if len(shape) == 1:
        return mem.reshape(shape).copy()
    return np.ascontiguousarray(mem.reshape(shape[1], shape[0]).T)


def dimensions_as_matrix(a, orientation):
    """reference src/linear-algebra.lisp:18-26."""
    a = np.asarray(a)
    if a.ndim == 1:
        return (1, a.shape[0], True) if orientation == "row" else (a.shape[0], 1, True)
    if a.ndim == 2:
        return a.shape[0], a.shape[1], False
    raise ValueError("array rank must be 1 or 2")


def _int(v):
    return ctypes.byref(ctypes.c_int(int(v)))


def _chr(c):
    return ctypes.byref(ctypes.c_char(c.encode()))


def _atom(v, tcode):
    a = np.array([v], dtype=_DTYPES[tcode])
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _fn(tcode, name):
    """reference src/fortran-call.lisp:405-416: type letter + routine name + underscore."""
    f = getattr(_lib.lib(), f"{_LETTER[tcode]}{name}_")
    f.restype = None
    return f


def _check_device_error():
    if _lib.lib().llacuda_last_error_code() != 0:
        msg = _lib.last_error()
        _lib.lib().llacuda_clear_error()
        raise _lib.LlacudaError(msg)


def _check_info(info, condition=LapackFailure):
    """reference src/fortran-call.lisp:336-347."""
    if info <= -10000:
        msg = _lib.last_error()
        _lib.lib().llacuda_clear_error()
        raise _lib.LlacudaError(f"device failure (info={info}): {msg}")
    if info < 0:
        raise LapackInvalidArgument(-info)
    if info > 0 and condition is not None:
        raise condition(info)


# ------------------------------------------------------------------ mm / gemm!
def mm(a, b):
    """reference src/linear-algebra.lisp:39-54; (mm a t) / (mm t a) are the hermitian products of :62-84."""
    if b is True:
        return mm_hermitian(a, False)
    if a is True:
        return mm_hermitian(b, True)
    a, b = np.asarray(a), np.asarray(b)
    t = common_float_type(a, b)
    a0, a1, av = dimensions_as_matrix(a, "row")
    b0, b1, bv = dimensions_as_matrix(b, "column")
    if av and bv:
        raise ValueError("MM called with two vectors.  Use OUTER or DOT instead.")
    if a1 != b0:
        raise AssertionError("Incompatible dimensions.")
    am, bm = _as_memory(a, t), _as_memory(b, t)
    c = np.zeros(a0 * b1 if (av or bv) else (a0, b1), dtype=_DTYPES[t])
    _one, pone = _atom(1, t)
    _zero, pzero = _atom(0, t)
    _fn(t, "gemm")(_chr("N"), _chr("N"), _int(b1), _int(a0), _int(b0), pone, _ptr(bm), _int(b1),
                   _ptr(am), _int(a1), pzero, _ptr(c), _int(b1))
    _check_device_error()
    return c


def gemm(alpha, a, b, beta, c, transpose_a=False, transpose_b=False, m=None, n=None, k=None,
         lda=None, ldb=None, ldc=None):
    """gemm!, reference src/blas.lisp:7-63.  C is updated in place and returned."""
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
    am, bm = _as_memory(a, t), _as_memory(b, t)
    cm = _as_memory(c, t)
    _al, pal = _atom(alpha, t)
    _bt, pbt = _atom(beta, t)
    _fn(t, "gemm")(_chr("C" if transpose_b else "N"), _chr("C" if transpose_a else "N"),
                   _int(n), _int(m), _int(k), pal, _ptr(bm), _int(ldb), _ptr(am), _int(lda),
                   pbt, _ptr(cm), _int(ldc))
    _check_device_error()
    if cm is not c:
        c[...] = cm  # copy-back rule, src/pinned-array.lisp:90-103
    return c


# ------------------------------------------------------------------ lu / solve
class LU:
    """reference src/factorizations.lisp:47-52 with the ipiv-mixin of :7-13."""

    def __init__(self, lu, ipiv_internal):
        self.lu, self.ipiv_internal = lu, ipiv_internal


def _native_fn(name):
    f = getattr(_lib.lib(), name)
    f.restype = None
    return f


def lu(a, native=False) -> LU:
    """reference src/linear-algebra.lisp:225-234: GETRF on a transposed copy; INFO>0 ignored.
    native=True: the row-major array goes to llacuda_dgetrf_rm as it is (layout conversion on the device)."""
    a = np.asarray(a)
    t = common_float_type(a)
    a0, a1 = a.shape
    if native:
        if t != DOUBLE:
            raise NotImplementedError("row-major-native entry points are double precision only")
        mem = np.array(a, dtype=np.float64, order="C", copy=True)
        ipiv_internal = np.zeros(min(a0, a1), dtype=np.int32)
        info = ctypes.c_int(0)
        _native_fn("llacuda_dgetrf_rm")(_int(a0), _int(a1), _ptr(mem), _int(a1), _ptr(ipiv_internal), ctypes.byref(info))
        _check_info(info.value, None)
        return LU(mem, ipiv_internal)
    mem = _transposed_to_memory(a, t)
    ipiv_internal = np.zeros(min(a0, a1), dtype=np.int32)
    info = ctypes.c_int(0)
    _fn(t, "getrf")(_int(a0), _int(a1), _ptr(mem), _int(a0), _ptr(ipiv_internal), ctypes.byref(info))
    _check_info(info.value, None)
    return LU(_transposed_from_memory(mem, (a0, a1)), ipiv_internal)


class Cholesky:
    """reference src/factorizations.lisp:130-136: LEFT is the lower-triangular left square root."""

    def __init__(self, left):
        self.left = left


def _solve_native(a, b):
    """solve through the row-major-native entry points: no host-side transposes, input-only operands stay put."""
    b = np.asarray(b)
    b0, b1, _ = dimensions_as_matrix(b, "column")
    x = np.array(b, dtype=np.float64, order="C", copy=True).reshape(b0, b1)
    info = ctypes.c_int(0)
    if isinstance(a, LU):
        n0, n1 = a.lu.shape
        if not (n0 == n1 == b0):
            raise LlaIncompatibleDimensions()
        lu_ = np.ascontiguousarray(a.lu, dtype=np.float64)
        _native_fn("llacuda_dgetrs_rm")(_int(n0), _int(b1), _ptr(lu_), _int(n0), _ptr(a.ipiv_internal), _ptr(x), _int(b1),
                                        ctypes.byref(info))
    elif isinstance(a, Cholesky):
        l = np.ascontiguousarray(a.left, dtype=np.float64)
        n0, n1 = l.shape
        if not (n0 == n1 == b0):
            raise LlaIncompatibleDimensions()
        _native_fn("llacuda_dpotrs_rm")(_int(n0), _int(b1), _ptr(l), _int(n0), _ptr(x), _int(b1), ctypes.byref(info))
    else:
        am = np.ascontiguousarray(a, dtype=np.float64)
        n0, n1 = am.shape
        if not (n0 == n1 == b0):
            raise LlaIncompatibleDimensions()
        _native_fn("llacuda_dgesv_rm")(_int(n0), _int(b1), _ptr(am), _int(n0), _ptr(x), _int(b1), ctypes.byref(info))
    _check_info(info.value)
    return x.reshape(b.shape)


def solve(a, b, native=False):
    """reference src/linear-algebra.lisp:257-301, dispatching on the class of A like the generic function.
    native=True (double precision): the row-major-native entry points of include/llacuda.h."""
    if native and not isinstance(a, (Hermitian, LowerTriangular, UpperTriangular, DeviceLU, DeviceCholesky)):
        return _solve_native(a, b)
    if isinstance(a, Hermitian):
        return solve_hermitian(a, b)
    if isinstance(a, LowerTriangular):
        return trsm(a.elements, False, b)
    if isinstance(a, UpperTriangular):
        return trsm(a.elements, True, b)
    b = np.asarray(b)
    if isinstance(a, (DeviceLU, DeviceCholesky)):
        return _solve_device(a, b)
    if isinstance(a, LU):
        lu0, lu1 = a.lu.shape
        b0, b1, _ = dimensions_as_matrix(b, "column")
        if not (lu0 == lu1 == b0):
            raise LlaIncompatibleDimensions()
        t = common_float_type(a.lu, b)
        amem = _transposed_to_memory(a.lu, t)
        bmem = _transposed_to_memory(b, t)
        info = ctypes.c_int(0)
        _fn(t, "getrs")(_chr("N"), _int(lu0), _int(b1), _ptr(amem), _int(lu0), _ptr(a.ipiv_internal),
                        _ptr(bmem), _int(lu0), ctypes.byref(info))
        _check_info(info.value)
        return _transposed_from_memory(bmem, b.shape)
    if isinstance(a, Cholesky):
        l = a.left
        a0, a1 = l.shape
        b0, b1, _ = dimensions_as_matrix(b, "column")
        if not (a0 == a1 == b0):
            raise LlaIncompatibleDimensions()
        t = common_float_type(l, b)
        amem = _as_memory(l, t)  # shared, read-only
        bmem = _transposed_to_memory(b, t)
        info = ctypes.c_int(0)
        _fn(t, "potrs")(_chr("U"), _int(a0), _int(b1), _ptr(amem), _int(a0), _ptr(bmem), _int(b0),
                        ctypes.byref(info))
        _check_info(info.value)
        return _transposed_from_memory(bmem, b.shape)
    a = np.asarray(a)
    a0, a1 = a.shape
    b0, b1, _ = dimensions_as_matrix(b, "column")
    if not (a0 == a1 == b0):
        raise LlaIncompatibleDimensions()
    t = common_float_type(a, b)
    amem = _transposed_to_memory(a, t)
    bmem = _transposed_to_memory(b, t)
    ipiv_scratch = np.zeros(a0, dtype=np.int32)  # (&work a0 +integer+), discarded
    info = ctypes.c_int(0)
    _fn(t, "gesv")(_int(a0), _int(b1), _ptr(amem), _int(a0), _ptr(ipiv_scratch), _ptr(bmem), _int(a0),
                   ctypes.byref(info))
    _check_info(info.value)
    return _transposed_from_memory(bmem, b.shape)


def cholesky(a) -> Cholesky:
    """reference src/linear-algebra.lisp:666-674: POTRF('U') on an untransposed copy of the
    row-major elements; only the row-major lower triangle is read, and the result is wrapped as a
    lower-triangular matrix (the strict upper triangle is masked)."""
    a = np.asarray(a)
    a0, a1 = a.shape
    if a0 != a1:
        raise LlaIncompatibleDimensions()
    t = common_float_type(a)
    mem = np.array(a, dtype=_DTYPES[t], order="C", copy=True)  # force-copy: input is not output
    info = ctypes.c_int(0)
    _fn(t, "potrf")(_chr("U"), _int(a0), _ptr(mem), _int(a0), ctypes.byref(info))
    _check_info(info.value)
    return Cholesky(np.tril(mem))


# ------------------------------------------------------------------ "next" rows: mm-hermitian%, posv, trsm%
class Hermitian:
    """cl-num-utils hermitian-matrix: the row-major LOWER triangle of `elements` defines the matrix."""

    def __init__(self, elements):
        self.elements = np.asarray(elements)

    def as_array(self):
        low = np.tril(self.elements)
        return low + np.tril(self.elements, -1).conj().T


class LowerTriangular:
    def __init__(self, elements):
        self.elements = np.asarray(elements)


class UpperTriangular:
    def __init__(self, elements):
        self.elements = np.asarray(elements)


def mm_hermitian(a, transpose_left: bool) -> Hermitian:
    """reference src/linear-algebra.lisp:62-76: (mm a t) = A A^T, (mm t a) = A^T A through SYRK('U') on the untransposed
    row-major buffer (its column-major view is A^T, hence 'C' for A A^T and 'N' for A^T A)."""
    a = np.asarray(a)
    a0, a1 = a.shape
    dim_c, other = (a1, a0) if transpose_left else (a0, a1)
    t = common_float_type(a)
    am = _as_memory(a, t)
    c = np.zeros((dim_c, dim_c), dtype=_DTYPES[t])
    (_k1, pone), (_k0, pzero) = _atom(1, t), _atom(0, t)
    # ("syrk" "herk"): SYRK for real arrays, HERK for complex ones; alpha / beta travel as cells of the matrix type
    name = "herk" if t in (COMPLEX_SINGLE, COMPLEX_DOUBLE) else "syrk"
    _fn(t, name)(_chr("U"), _chr("N" if transpose_left else "C"), _int(dim_c), _int(other), pone, _ptr(am),
                 _int(a1), pzero, _ptr(c), _int(dim_c))
    _check_device_error()
    return Hermitian(c)


def solve_hermitian(a: Hermitian, b):
    """reference src/linear-algebra.lisp:303-315: POSV('U') on a forced copy of the elements, B transposed in and out."""
    el, b = a.elements, np.asarray(b)
    a0, a1 = el.shape
    b0, b1, _ = dimensions_as_matrix(b, "column")
    if not (a0 == a1 == b0):
        raise LlaIncompatibleDimensions()
    t = common_float_type(el, b)
    amem = np.array(el, dtype=_DTYPES[t], order="C", copy=True)
    bmem = _transposed_to_memory(b, t)
    info = ctypes.c_int(0)
    _fn(t, "posv")(_chr("U"), _int(a0), _int(b1), _ptr(amem), _int(a0), _ptr(bmem), _int(b0), ctypes.byref(info))
    _check_info(info.value)
    return _transposed_from_memory(bmem, b.shape)


def trsm(a, a_upper: bool, b):
    """reference src/linear-algebra.lisp:333-347 (trsm%): A X = B for triangular A as TRSM('R', L/U swapped, 'N', 'N')
    on the untransposed row-major buffers."""
    a, b = np.asarray(a), np.asarray(b)
    a0, a1 = a.shape
    b0, b1, _ = dimensions_as_matrix(b, "column")
    if not (a0 == a1 == b0):
        raise LlaIncompatibleDimensions()
    t = common_float_type(a, b)
    am = _as_memory(a, t)
    x = np.array(b, dtype=_DTYPES[t], order="C", copy=True).reshape(b0, b1)
    _k1, pone = _atom(1, t)
    _fn(t, "trsm")(_chr("R"), _chr("L" if a_upper else "U"), _chr("N"), _chr("N"), _int(b1), _int(b0), pone,
                   _ptr(am), _int(a1), _ptr(x), _int(b1))
    _check_device_error()
    return x.reshape(b.shape)


# ------------------------------------------------------------------ the invert family
def invert(a):
    """reference src/linear-algebra.lisp:352-409: arrays through (invert (lu a)) = GETRI with a workspace query,
    cholesky objects through POTRI('U') -> Hermitian, triangular wrappers through TRTRI (L/U swapped: row-major)."""
    if isinstance(a, (UpperTriangular, LowerTriangular)):
        upper = isinstance(a, UpperTriangular)
        el = np.asarray(a.elements)
        a0, a1 = el.shape
        assert a0 == a1
        t = common_float_type(el)
        mem = np.array(el, dtype=_DTYPES[t], order="C", copy=True)
        info = ctypes.c_int(0)
        _fn(t, "trtri")(_chr("L" if upper else "U"), _chr("N"), _int(a0), _ptr(mem), _int(a0), ctypes.byref(info))
        _check_info(info.value)
        return UpperTriangular(np.triu(mem)) if upper else LowerTriangular(np.tril(mem))
    if isinstance(a, Cholesky):
        l = np.asarray(a.left)
        a0, a1 = l.shape
        assert a0 == a1
        t = common_float_type(l)
        mem = np.array(l, dtype=_DTYPES[t], order="C", copy=True)
        info = ctypes.c_int(0)
        _fn(t, "potri")(_chr("U"), _int(a0), _ptr(mem), _int(a0), ctypes.byref(info))
        _check_info(info.value)
        return Hermitian(mem)
    if not isinstance(a, LU):
        a = lu(a)
    lu0, lu1 = a.lu.shape
    assert lu0 == lu1 == len(a.ipiv_internal)
    t = common_float_type(a.lu)
    mem = _transposed_to_memory(a.lu, t)
    info = ctypes.c_int(0)
    query = np.zeros(1, dtype=_DTYPES[t])
    _fn(t, "getri")(_int(lu0), _ptr(mem), _int(lu0), _ptr(a.ipiv_internal), _ptr(query), _int(-1), ctypes.byref(info))
    _check_info(info.value)                      # lapack-call-w/query: first call sizes the work area
    lwork = max(1, int(query[0].real))
    work = np.zeros(lwork, dtype=_DTYPES[t])
    _fn(t, "getri")(_int(lu0), _ptr(mem), _int(lu0), _ptr(a.ipiv_internal), _ptr(work), _int(lwork), ctypes.byref(info))
    _check_info(info.value)
    return _transposed_from_memory(mem, (lu0, lu0))


# ------------------------------------------------------------------ device-resident factorization objects
# the reference re-marshals the factors on every call -- `solve ((lu lu) b)` transposes the
# stored n x n LU into foreign memory each time (src/linear-algebra.lisp:264-276, src/foreign-memory.lisp:211-255).
# These objects keep the factors in HBM in LAPACK's column-major layout; the row-major <-> column-major conversion of
# the operands runs on the device (llacuda_dtranspose_dev, HBM-bound) instead of the element-wise host loops, and a
# solve moves only B and X over PCIe.  Double precision only (the device factorisations are D).
class _DevBuf:
    def __init__(self, nbytes):
        self.p = ctypes.c_void_p()
        self.nbytes = nbytes
        _lib.check(_lib.lib().llacuda_malloc(ctypes.byref(self.p), ctypes.c_size_t(max(nbytes, 16))), "malloc")

    def __del__(self):
        try:
            if self.p:
                _lib.lib().llacuda_free(self.p)
                self.p = ctypes.c_void_p()
        except Exception:
            pass


def _dev():
    from . import device as D
    _lib.lib().llacuda_reset_stream()      # the library's own stream: the blocking memcpy helpers synchronise on it
    return D


def _upload_colmajor(a):
    """Row-major (a0, a1) array of any supported element type -> device buffer holding the column-major a0 x a1 double
    matrix (ld = a0).  The raw bytes travel as they are; coercion to double and the transpose happen in one device
    kernel (llacuda_convert_dev) instead of LLA's element-wise transpose-matrix-to-memory (src/foreign-memory.lisp:211-230)."""
    D = _dev()
    a = np.asarray(a)
    if str(a.dtype) not in ("float32", "float64", "int32", "int64"):
        _efficiency_warning_conversion(a, np.float64)
        a = a.astype(np.float64)            # exotic element types (object arrays, small ints, bools): host walk, as the reference
    elif a.dtype != np.float64:
        _efficiency_warning_conversion(a, np.float64)
    a = np.ascontiguousarray(a)
    a0, a1 = a.shape
    raw, cm = _DevBuf(a.nbytes), _DevBuf(a0 * a1 * 8)
    _lib.check(_lib.lib().llacuda_memcpy_h2d(raw.p, _ptr(a), ctypes.c_size_t(a.nbytes)), "h2d")
    # the row-major bytes are the column-major image of a^T (a1 x a0, ld = a1): convert + transpose it on the device
    D.convert(a.dtype, np.dtype(np.float64), True, a1, a0, raw.p, a1, cm.p, a0)
    return cm


def _download_rowmajor(cm, a0, a1):
    """Device column-major a0 x a1 matrix (ld = a0) -> row-major (a0, a1) numpy array."""
    D = _dev()
    raw = _DevBuf(a0 * a1 * 8)
    D.dtranspose(a0, a1, cm.p, a0, raw.p, a1)
    out = np.empty((a0, a1), dtype=np.float64)
    _lib.check(_lib.lib().llacuda_memcpy_d2h(_ptr(out), raw.p, ctypes.c_size_t(out.nbytes)), "d2h")
    return out


class DeviceLU(LU):
    """`lu` object whose factors and pivots live on the device; `.lu` / `.ipiv_internal` download on first use."""

    def __init__(self, shape, d_lu, d_ipiv):
        self.shape, self._d_lu, self._d_ipiv = shape, d_lu, d_ipiv
        self._lu = self._ipiv = None

    @property
    def lu(self):
        if self._lu is None:
            self._lu = _download_rowmajor(self._d_lu, *self.shape)
        return self._lu

    @property
    def ipiv_internal(self):
        if self._ipiv is None:
            k = min(self.shape)
            self._ipiv = np.zeros(k, dtype=np.int32)
            _lib.check(_lib.lib().llacuda_memcpy_d2h(_ptr(self._ipiv), self._d_ipiv.p, ctypes.c_size_t(4 * k)), "d2h")
        return self._ipiv


class DeviceCholesky(Cholesky):
    """`cholesky` object whose factor lives on the device (POTRF 'U' of the untransposed row-major buffer = the
    row-major lower factor, reference notes.org:32-37); `.left` downloads on first use."""

    def __init__(self, n, d_u):
        self.n, self._d_u = n, d_u
        self._left = None

    @property
    def left(self):
        if self._left is None:
            out = np.empty((self.n, self.n), dtype=np.float64)
            _lib.check(_lib.lib().llacuda_memcpy_d2h(_ptr(out), self._d_u.p, ctypes.c_size_t(out.nbytes)), "d2h")
            self._left = np.tril(out)
        return self._left


def _device_info(d_info, condition):
    info = ctypes.c_int(0)
    _lib.check(_lib.lib().llacuda_memcpy_d2h(ctypes.byref(info), d_info.p, ctypes.c_size_t(4)), "d2h")
    _check_info(info.value, condition)


def lu_device(a) -> DeviceLU:
    """`lu` (reference src/linear-algebra.lisp:225-234) keeping the result on the device."""
    a = np.asarray(a)
    if common_float_type(a) != DOUBLE:
        raise NotImplementedError("device-resident factorisations are double precision only")
    a0, a1 = a.shape
    D = _dev()
    d_lu = _upload_colmajor(a)
    d_ipiv, d_info = _DevBuf(4 * max(min(a0, a1), 1)), _DevBuf(4)
    D.dgetrf(a0, a1, d_lu.p, a0, d_ipiv.p, d_info.p)
    _device_info(d_info, None)   # INFO > 0 is ignored by lu (src/linear-algebra.lisp:234)
    return DeviceLU((a0, a1), d_lu, d_ipiv)


def cholesky_device(a) -> DeviceCholesky:
    """`cholesky` (reference src/linear-algebra.lisp:666-674) keeping the factor on the device."""
    a = np.ascontiguousarray(a)
    if common_float_type(a) != DOUBLE:
        raise NotImplementedError("device-resident factorisations are double precision only")
    a0, a1 = a.shape
    if a0 != a1:
        raise LlaIncompatibleDimensions()
    D = _dev()
    if str(a.dtype) in ("float32", "int32", "int64"):       # classifies as double (integer arrays): coerce on the device
        _efficiency_warning_conversion(a, np.float64)
        raw, d_u, d_info = _DevBuf(a.nbytes), _DevBuf(a0 * a1 * 8), _DevBuf(4)
        _lib.check(_lib.lib().llacuda_memcpy_h2d(raw.p, _ptr(a), ctypes.c_size_t(a.nbytes)), "h2d")
        D.convert(a.dtype, np.dtype(np.float64), False, a1, a0, raw.p, a1, d_u.p, a1)
        a = None
    else:
        a = a.astype(np.float64, copy=False)
        d_u, d_info = _DevBuf(a.nbytes), _DevBuf(4)
        _lib.check(_lib.lib().llacuda_memcpy_h2d(d_u.p, _ptr(a), ctypes.c_size_t(a.nbytes)), "h2d")
    D.dpotrf("U", a0, d_u.p, a0, d_info.p)
    _device_info(d_info, LapackFailure)
    return DeviceCholesky(a0, d_u)


def _solve_device(f, b):
    b = np.asarray(b)
    b0, b1, vec = dimensions_as_matrix(b, "column")
    n = f.shape[0] if isinstance(f, DeviceLU) else f.n
    if isinstance(f, DeviceLU) and f.shape[0] != f.shape[1]:
        raise LlaIncompatibleDimensions()
    if n != b0:
        raise LlaIncompatibleDimensions()
    D = _dev()
    d_b = _upload_colmajor(np.asarray(b, dtype=np.float64).reshape(b0, b1))
    if isinstance(f, DeviceLU):
        D.dgetrs("N", n, b1, f._d_lu.p, n, f._d_ipiv.p, d_b.p, n)
    else:
        D.dpotrs("U", n, b1, f._d_u.p, n, d_b.p, n)
    x = _download_rowmajor(d_b, b0, b1)
    _check_device_error()
    return x.reshape(b.shape)


# ------------------------------------------------------------------ factorization post-processing
def ipiv(ipiv_internal) -> np.ndarray:
    """reference src/factorizations.lisp:15-28: LAPACK swap sequence -> 0-based permutation."""
    n = len(ipiv_internal)
    it = np.arange(n, dtype=np.int64)
    for index in range(n):
        p = int(ipiv_internal[index]) - 1
        if p != index:
            it[index], it[p] = it[p], it[index]
    return it


def ipiv_inverse(ipiv_internal) -> np.ndarray:
    """reference src/factorizations.lisp:30-43."""
    p = ipiv(ipiv_internal)
    inv = np.zeros(len(p), dtype=np.int64)
    inv[p] = np.arange(len(p))
    return inv


def lu_u(f: LU) -> np.ndarray:
    """reference src/factorizations.lisp:54-56."""
    return np.triu(f.lu)


def lu_l(f: LU) -> np.ndarray:
    """reference src/factorizations.lisp:58-65 (unit diagonal forced)."""
    l = np.tril(f.lu).copy()
    for i in range(min(l.shape)):
        l[i, i] = 1
    return l


def permutations(ipiv_internal) -> int:
    """reference src/factorizations.lisp:143-151."""
    return int(sum(1 for i, p in enumerate(ipiv_internal, start=1) if p != i))
