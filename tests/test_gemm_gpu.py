"""GPU parity: DGEMM through the Fortran drop-in (`dgemm_`, host pointers) against the CPU oracle.

Cases restated from the reference's own tests:
  mm goldens           tests/linear-algebra.lisp:9-29
  gemm! randomized     tests/blas.lisp:19-46   (integer-valued operands in padded "sloppy" arrays)
plus seeded random cases at sizes the oracle finishes in seconds and BASELINE configs[0] (n=1024).
Tolerance (north_star "elementwise error <= k * eps * ||A|| ||B||"): with a_i the i-th row of
op(A) and b_j the j-th column of op(B),  |C_ij - Cref_ij| <= k * eps * ||a_i||_2 ||b_j||_2  with
k = 4*sqrt(K) (the rigorous worst case is k = K; both sides of the comparison carry rounding).
"""
import numpy as np
import pytest

from lla_b200 import lla as L
from oracle import lla_oracle as O

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def gemm_tol(opa_a, opb_b, K, scale=1.0):
    """k*eps*||a_i|| ||b_j|| matrix for op(A) (m x K) and op(B) (K x n)."""
    ra = np.linalg.norm(opa_a, axis=1)[:, None]
    cb = np.linalg.norm(opb_b, axis=0)[None, :]
    return 4 * np.sqrt(max(K, 1)) * EPS * scale * ra * cb + 1e-300


def test_mm_goldens():
    a = np.array([[1, 2], [3, 4], [5, 6]], dtype=np.float64)
    b2 = np.array([1, 2], dtype=np.float64)
    b3 = np.array([1, 2, 3], dtype=np.float64)
    assert np.array_equal(L.mm(a, np.eye(2)), a)
    assert np.array_equal(L.mm(np.eye(3), a), a)
    r = L.mm(a, b2)
    assert r.shape == (3,) and np.array_equal(r, [5, 11, 17])
    r = L.mm(b3, a)
    assert r.shape == (2,) and np.array_equal(r, [22, 28])
    with pytest.raises(Exception):
        L.mm(a, b3)
    with pytest.raises(Exception):
        L.mm(b2, b3)


def sloppy(rng, h, w):
    full = rng.integers(0, 100, (h + rng.integers(0, 3), w + rng.integers(0, 3))).astype(np.float64)
    return full


def test_gemm_randomized_like_reference():
    rng = np.random.default_rng(11)
    be = O.netlib()
    for _ in range(100):
        m, n, k = (int(v) for v in rng.integers(1, 11, 3))
        ta, tb = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        a_ = sloppy(rng, *((k, m) if ta else (m, k)))
        b_ = sloppy(rng, *((n, k) if tb else (k, n)))
        c_ = sloppy(rng, m, n)
        alpha, beta = rng.uniform(0, 10), rng.uniform(0, 10)
        want = O.gemm(alpha, a_, b_, beta, c_.copy(), be, transpose_a=ta, transpose_b=tb, m=m, n=n, k=k)
        got = L.gemm(alpha, a_, b_, beta, c_.copy(), transpose_a=ta, transpose_b=tb, m=m, n=n, k=k)
        # integer operands: the products are exact, only alpha/beta scaling rounds -> 2 ulp
        assert np.allclose(got, want, rtol=4 * EPS, atol=0), (m, n, k, ta, tb)
        # cells outside the logical m x n block must be untouched
        mask = np.ones_like(c_, dtype=bool)
        mask[:m, :n] = False
        assert np.array_equal(got[mask], c_[mask])


@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (7, 5, 3), (64, 64, 64), (129, 67, 33), (200, 31, 250), (33, 300, 17), (257, 255, 130)])
def test_gemm_shapes_vs_oracle(ta, tb, m, n, k):
    rng = np.random.default_rng(m * 1000 + n * 10 + k)
    a = rng.uniform(-1, 1, (k, m) if ta else (m, k))
    b = rng.uniform(-1, 1, (n, k) if tb else (k, n))
    c = rng.uniform(-1, 1, (m, n))
    for alpha, beta in ((1.0, 0.0), (0.7, 1.3), (0.0, 0.5), (2.0, 1.0)):
        want = O.gemm(alpha, a, b, beta, c.copy(), O.openblas(), transpose_a=ta, transpose_b=tb)
        got = L.gemm(alpha, a, b, beta, c.copy(), transpose_a=ta, transpose_b=tb)
        tol = gemm_tol(a.T if ta else a, b.T if tb else b, k, abs(alpha)) + 4 * EPS * abs(beta) * np.abs(c)
        assert np.all(np.abs(got - want) <= tol)


def test_beta_zero_does_not_read_c():
    rng = np.random.default_rng(1)
    a, b = rng.uniform(-1, 1, (40, 30)), rng.uniform(-1, 1, (30, 20))
    c = np.full((40, 20), np.nan)
    got = L.gemm(1.0, a, b, 0.0, c)
    assert np.all(np.isfinite(got))
    assert np.allclose(got, a @ b, atol=30 * 4 * EPS)


def test_mm_config0_1024():
    """BASELINE configs[0]: (lla:mm a b) 1024x1024 double, (i) integer operands => exact, (ii) U[0,1)."""
    rng = np.random.default_rng(0)
    a = rng.integers(0, 100, (1024, 1024)).astype(np.float64)
    b = rng.integers(0, 100, (1024, 1024)).astype(np.float64)
    assert np.array_equal(L.mm(a, b), O.mm(a, b, O.openblas()))
    a, b = rng.uniform(0, 1, (1024, 1024)), rng.uniform(0, 1, (1024, 1024))
    got, want = L.mm(a, b), O.mm(a, b, O.openblas())
    assert np.all(np.abs(got - want) <= gemm_tol(a, b, 1024))


def test_mm_linearity_large():
    """size-independent property at a size the oracle would not finish quickly:
    (A1 + A2) B == A1 B + A2 B up to rounding, and mm(A, I) == A exactly."""
    rng = np.random.default_rng(5)
    n = 3000
    a1, a2, b = rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, n))
    lhs = L.mm(a1 + a2, b)
    rhs = L.mm(a1, b) + L.mm(a2, b)
    assert np.all(np.abs(lhs - rhs) <= 2 * gemm_tol(np.abs(a1) + np.abs(a2), b, n))
    assert np.array_equal(L.mm(a1, np.eye(n)), a1)


# ------------------------------------------------------------------ complex double (ZGEMM)
EPSZ = np.finfo(np.float64).eps


def crand(rng, *shape):
    return rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)


@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (5, 7, 3), (64, 64, 64), (70, 130, 33), (129, 65, 100), (200, 31, 257)])
def test_zgemm_vs_oracle(ta, tb, m, n, k):
    """gemm! on complex-double arrays: a transposed operand is CONJUGATE-transposed (src/blas.lisp:57-58)."""
    rng = np.random.default_rng(m * 100 + n * 10 + k)
    a = crand(rng, *((k, m) if ta else (m, k)))
    b = crand(rng, *((n, k) if tb else (k, n)))
    c = crand(rng, m, n)
    for alpha, beta in ((1.0 + 0j, 0j), (0.7 - 0.2j, 1.3 + 0.5j), (0j, 0.5 + 0.1j)):
        want = O.gemm(alpha, a, b, beta, c.copy(), O.openblas(), transpose_a=ta, transpose_b=tb)
        got = L.gemm(alpha, a, b, beta, c.copy(), transpose_a=ta, transpose_b=tb)
        opa = a.conj().T if ta else a
        opb = b.conj().T if tb else b
        tol = 2 * gemm_tol(np.abs(opa), np.abs(opb), k, abs(alpha)) + 8 * EPSZ * abs(beta) * np.abs(c)
        assert np.all(np.abs(got - want) <= tol)


def test_zmm_mixed_types_and_identity():
    rng = np.random.default_rng(8)
    a = crand(rng, 90, 40)
    b = rng.uniform(-1, 1, (40, 70))              # real * complex -> complex-double (LOGIOR of the codes)
    got = L.mm(a, b)
    assert got.dtype == np.complex128
    assert np.allclose(got, a @ b, atol=40 * 8 * EPSZ)
    assert np.array_equal(L.mm(a, np.eye(40)), a)


def test_zgemm_large_property():
    """n = 1500: conj-transpose identity (A B)^H = B^H A^H through the two transposed code paths."""
    rng = np.random.default_rng(4)
    n = 1500
    a, b = crand(rng, n, n), crand(rng, n, n)
    ab = L.mm(a, b)
    c = np.zeros((n, n), dtype=np.complex128)
    bhah = L.gemm(1.0, b, a, 0.0, c, transpose_a=True, transpose_b=True)
    assert np.all(np.abs(ab.conj().T - bhah) <= 4 * gemm_tol(np.abs(a), np.abs(b), n).T)


# ------------------------------------------------------------------ single precision (tcgen05, 3xTF32)
EPS32 = float(np.finfo(np.float32).eps)


def sgemm_tol(opa, opb, K, scale=1.0):
    """k * eps32 * ||a_i|| ||b_j|| with k = 16 + 4 sqrt(K): the 3xTF32 split drops terms of relative
    size 2^-21 per product (8 eps32) on top of the FP32 accumulation error."""
    ra = np.linalg.norm(opa.astype(np.float64), axis=1)[:, None]
    cb = np.linalg.norm(opb.astype(np.float64), axis=0)[None, :]
    return (16 + 4 * np.sqrt(max(K, 1))) * EPS32 * scale * ra * cb + 1e-30


@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (5, 7, 3), (128, 128, 32), (128, 128, 64), (130, 70, 33), (257, 129, 100), (64, 300, 257)])
def test_sgemm_vs_oracle(ta, tb, m, n, k):
    rng = np.random.default_rng(m * 100 + n * 10 + k)
    a = rng.uniform(-1, 1, (k, m) if ta else (m, k)).astype(np.float32)
    b = rng.uniform(-1, 1, (n, k) if tb else (k, n)).astype(np.float32)
    c = rng.uniform(-1, 1, (m, n)).astype(np.float32)
    for alpha, beta in ((1.0, 0.0), (0.7, 1.3), (0.0, 0.5)):
        got = L.gemm(alpha, a, b, beta, c.copy(), transpose_a=ta, transpose_b=tb)
        assert got.dtype == np.float32
        opa = (a.T if ta else a).astype(np.float64)
        opb = (b.T if tb else b).astype(np.float64)
        want = alpha * (opa @ opb) + beta * c.astype(np.float64)
        tol = sgemm_tol(opa, opb, k, abs(alpha)) + 4 * EPS32 * (abs(beta) * np.abs(c) + np.abs(want))
        assert np.all(np.abs(got.astype(np.float64) - want) <= tol), np.max(np.abs(got - want) / tol)


def test_smm_accuracy_is_fp32_grade_not_tf32():
    """plain TF32 would be wrong at the 1e-3 level; the 3xTF32 split with chunked accumulation must be at
    the 1e-6 level even for all-positive data (the tensor core's truncating FP32 accumulation alone
    gave 2.3e-5 here; with 128-deep TMEM chunks the residual bias is ~128 * 2^-24 * 0.4 = 3e-6)."""
    rng = np.random.default_rng(6)
    n = 1024
    a = rng.uniform(0, 1, (n, n)).astype(np.float32)
    b = rng.uniform(0, 1, (n, n)).astype(np.float32)
    got = L.mm(a, b).astype(np.float64)
    want = a.astype(np.float64) @ b.astype(np.float64)
    rel = np.max(np.abs(got - want) / np.abs(want))
    assert rel < 4e-6, rel
    ref32 = O.mm(a, b, O.openblas()).astype(np.float64)        # host SGEMM for comparison
    assert rel <= 4 * np.max(np.abs(ref32 - want) / np.abs(want)) + 1e-6


# ------------------------------------------------------------------ degenerate shapes (netlib quick returns)
def test_gemm_degenerate_shapes():
    rng = np.random.default_rng(3)
    c = rng.uniform(-1, 1, (4, 5))
    # k = 0: C = beta * C, A and B are never read
    # (padded operands with k = 0 logical columns: gemm!'s own asserts, src/blas.lisp:40-53, need lda >= 1 rows to exist)
    got = L.gemm(1.0, np.zeros((4, 1)), np.zeros((1, 5)), 0.5, c.copy(), k=0, lda=1, ldb=5)
    assert np.allclose(got, 0.5 * c, rtol=0, atol=1e-16)
    # m = 0 / n = 0: nothing happens
    e = np.zeros((0, 5))
    assert L.gemm(1.0, np.zeros((0, 3)), np.zeros((3, 5)), 0.0, e, lda=3).shape == (0, 5)
    # vectors through mm: (1 x n) and (n x 1) results
    a = rng.uniform(-1, 1, (1, 200))
    b = rng.uniform(-1, 1, (200, 1))
    assert np.allclose(L.mm(a, b), a @ b, atol=1e-13)
    assert np.allclose(L.mm(b, a), b @ a, atol=1e-13)


# ------------------------------------------------------------------ complex single (planar split over the tcgen05 SGEMM)
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (5, 7, 3), (130, 70, 33), (257, 129, 100)])
def test_cgemm_vs_oracle(ta, tb, m, n, k):
    """gemm! on complex-single arrays (type letter c): a transposed operand is CONJUGATE-transposed."""
    rng = np.random.default_rng(m * 100 + n * 10 + k + 5)
    mk = lambda r, c: (rng.uniform(-1, 1, (r, c)) + 1j * rng.uniform(-1, 1, (r, c))).astype(np.complex64)
    a = mk(*((k, m) if ta else (m, k)))
    b = mk(*((n, k) if tb else (k, n)))
    c = mk(m, n)
    for alpha, beta in ((1.0 + 0j, 0j), (0.7 - 0.2j, 1.3 + 0.5j), (0j, 0.5 + 0.1j)):
        got = L.gemm(np.complex64(alpha), a, b, np.complex64(beta), c.copy(), transpose_a=ta, transpose_b=tb)
        assert got.dtype == np.complex64
        opa = (a.conj().T if ta else a).astype(np.complex128)
        opb = (b.conj().T if tb else b).astype(np.complex128)
        want = alpha * (opa @ opb) + beta * c.astype(np.complex128)
        host = O.gemm(np.complex64(alpha), a, b, np.complex64(beta), c.copy(), O.openblas(), transpose_a=ta, transpose_b=tb)
        tol = 4 * sgemm_tol(np.abs(opa), np.abs(opb), k, abs(alpha)) + 8 * EPS32 * (abs(beta) * np.abs(c) + np.abs(want))
        assert np.all(np.abs(got.astype(np.complex128) - want) <= tol), np.max(np.abs(got - want) / tol)
        assert np.all(np.abs(got.astype(np.complex128) - host.astype(np.complex128)) <= 2 * tol)


def test_cmm_types():
    rng = np.random.default_rng(12)
    a = (rng.uniform(-1, 1, (60, 40)) + 1j * rng.uniform(-1, 1, (60, 40))).astype(np.complex64)
    b = rng.uniform(-1, 1, (40, 50)).astype(np.float32)      # single * complex-single -> complex-single
    got = L.mm(a, b)
    assert got.dtype == np.complex64
    assert np.allclose(got, a.astype(np.complex128) @ b.astype(np.float64), atol=1e-4)


def test_dgemm_host_two_dimensional_pipeline():
    """The host-pointer dgemm_ cuts big products into 2 row blocks x 8 column blocks that are uploaded, multiplied and
    downloaded as a pipeline (csrc/fortran_abi.cu); Fortran m >= 8192 and 2 m n k > 4e10 select that path.  With
    LLA's row-major convention (C^T = B^T A^T) Fortran's m is the row-major number of columns."""
    rng = np.random.default_rng(21)
    m, n, k = 2176, 8192, 1200
    a = rng.uniform(-1, 1, (m, k))
    b = rng.uniform(-1, 1, (k, n))
    c = rng.uniform(-1, 1, (m, n))
    want0 = a @ b
    got = L.mm(a, b)
    tol = gemm_tol(np.abs(a), np.abs(b), k)
    assert np.all(np.abs(got - want0) <= tol)
    got = L.gemm(0.7, a, b, 1.3, c.copy())
    assert np.all(np.abs(got - (0.7 * want0 + 1.3 * c)) <= 0.7 * tol + 8 * EPS * 1.3 * np.abs(c))


def test_sgemm_dev_takes_unaligned_operands():
    """llacuda_sgemm_dev with leading dimensions that TMA cannot take (odd lda / ldb, LLA's sloppy arrays): served by the
    FP32-FMA tile kernel instead of an error."""
    import ctypes
    from lla_b200 import _lib, device as D
    lib = _lib.lib()
    _lib.init(0)
    lib.llacuda_reset_stream()
    rng = np.random.default_rng(3)
    m, n, k, lda, ldb, ldc = 70, 45, 130, 71, 131, 73
    a = rng.normal(size=(k, lda)).astype(np.float32)       # column-major m x k with leading dimension lda (image rows = columns)
    b = rng.normal(size=(n, ldb)).astype(np.float32)
    c = rng.normal(size=(n, ldc)).astype(np.float32)
    bufs = []
    for h in (a, b, c):
        d = L._DevBuf(h.nbytes)
        _lib.check(lib.llacuda_memcpy_h2d(d.p, h.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(h.nbytes)), "h2d")
        bufs.append(d)
    D.sgemm("N", "N", m, n, k, 1.5, bufs[0].p, lda, bufs[1].p, ldb, 0.5, bufs[2].p, ldc)
    out = np.empty_like(c)
    _lib.check(lib.llacuda_memcpy_d2h(out.ctypes.data_as(ctypes.c_void_p), bufs[2].p, ctypes.c_size_t(out.nbytes)), "d2h")
    am, bm = a[:, :m].T.astype(np.float64), b[:, :k].T.astype(np.float64)
    ref = 1.5 * (am @ bm) + 0.5 * c[:, :m].T
    assert np.max(np.abs(out[:, :m].T - ref)) <= 64 * k * np.finfo(np.float32).eps * np.abs(ref).max()
    assert np.array_equal(out[:, m:], c[:, m:])             # the padding rows of C are untouched
