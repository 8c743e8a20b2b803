"""Pins the CPU oracle (oracle/) against the reference's own golden vectors and
against host OpenBLAS issuing LLA's exact Fortran calls.  CPU only.

Goldens restated from /root/reference/tests (inputs and expected values only):
  mm            tests/linear-algebra.lisp:9-29
  mm-solve-lu   tests/linear-algebra.lisp:89-104
  lu-singular   tests/linear-algebra.lisp:106-107
  lu-random     tests/linear-algebra.lisp:109-115
  cholesky      tests/linear-algebra.lisp:366-385
  gemm          tests/blas.lisp:19-46
"""
import numpy as np
import pytest

from oracle import lla_oracle as O

BACKENDS = [O.netlib, O.openblas]
IDS = ["netlib", "openblas"]


def num_eq(a, b, tol=1e-5):
    """cl-num-utils num= : relative tolerance (tests/linear-algebra.lisp:330)."""
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b))))


@pytest.mark.parametrize("mk", BACKENDS, ids=IDS)
def test_mm_goldens(mk):
    be = mk()
    a = np.array([[1, 2], [3, 4], [5, 6]], dtype=np.float64)
    b2 = np.array([1, 2], dtype=np.float64)
    b3 = np.array([1, 2, 3], dtype=np.float64)
    i2, i3 = np.eye(2), np.eye(3)
    assert num_eq(O.mm(a, i2, be), a)
    assert num_eq(O.mm(i3, a, be), a)
    r = O.mm(a, b2, be)
    assert r.shape == (3,) and num_eq(r, [5, 11, 17])
    r = O.mm(b3, a, be)
    assert r.shape == (2,) and num_eq(r, [22, 28])
    with pytest.raises(Exception):
        O.mm(a, b3, be)
    with pytest.raises(Exception):
        O.mm(b2, b3, be)


@pytest.mark.parametrize("mk", BACKENDS, ids=IDS)
def test_mm_solve_lu(mk):
    be = mk()
    a = np.array([[1, 2], [3, 4]], dtype=object)  # (mx t ...): generic T array of integers
    a = np.array(a, dtype=np.int64)
    x = np.array([5.0, 6.0])
    b = O.mm(a, x, be)
    assert num_eq(b, [17.0, 39.0])
    assert num_eq(O.solve(a, b, be), x)
    f = O.lu(a, be)
    assert list(f.ipiv_internal) == [2, 2]
    assert num_eq(O.solve_lu(f, b, be), x)


@pytest.mark.parametrize("mk", BACKENDS, ids=IDS)
def test_lu_singular_does_not_signal(mk):
    f = O.lu(np.full((2, 2), 0.5), mk())
    assert f.lu.shape == (2, 2)
    with pytest.raises(O.LapackFailure) as e:
        O.solve(np.full((2, 2), 0.5), np.ones(2), mk())
    assert e.value.info == 2


@pytest.mark.parametrize("mk", BACKENDS, ids=IDS)
def test_lu_random_reconstruction(mk):
    be = mk()
    rng = np.random.default_rng(7)
    for n in range(2, 11):
        for _ in range(20):
            a = rng.uniform(0, 10, (n, n))
            f = O.lu(a, be)
            p, pinv = O.ipiv(f.ipiv_internal), O.ipiv_inverse(f.ipiv_internal)
            lu_prod = O.lu_l(f) @ O.lu_u(f)
            assert num_eq(a[p, :], lu_prod)
            assert num_eq(a, lu_prod[pinv, :])


@pytest.mark.parametrize("mk", BACKENDS, ids=IDS)
def test_cholesky_golden(mk):
    be = mk()
    a = np.array([[2, -1, 0], [-1, 2, -1], [0, -1, 2]], dtype=np.float64)
    l = np.array([[1.414214, 0, 0], [-0.7071068, 1.2247449, 0], [0.0, -0.8164966, 1.1547005]])
    c = O.cholesky(a, be)
    assert num_eq(c.left, l)
    b = np.array([5.0, 7.0, 13.0])
    ab = O.solve(a, b, be)
    assert num_eq(O.solve_cholesky(c, b, be), ab)
    assert num_eq(a @ ab, b)
    # only the row-major LOWER triangle is read
    a2 = a.copy(); a2[0, 2] = 99.0; a2[0, 1] = -55.0
    assert np.array_equal(O.cholesky(a2, be).left, c.left)
    with pytest.raises(O.LapackFailure) as e:
        O.cholesky(np.array([[1.0, 2.0], [2.0, 1.0]]), be)
    assert e.value.info == 2


def sloppy(rng, h, w):
    full = rng.integers(0, 100, (h + rng.integers(0, 3), w + rng.integers(0, 3))).astype(np.float64)
    return full[:h, :w].copy(), full


@pytest.mark.parametrize("mk", BACKENDS, ids=IDS)
def test_gemm_randomized(mk):
    be = mk()
    rng = np.random.default_rng(11)
    for _ in range(100):
        m, n, k = (int(v) for v in rng.integers(1, 11, 3))
        ta, tb = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        a, a_ = sloppy(rng, *((k, m) if ta else (m, k)))
        b, b_ = sloppy(rng, *((n, k) if tb else (k, n)))
        c, c_ = sloppy(rng, m, n)
        alpha, beta = rng.uniform(0, 10), rng.uniform(0, 10)
        want = alpha * ((a.T if ta else a) @ (b.T if tb else b)) + beta * c
        O.gemm(alpha, a_, b_, beta, c_, be, transpose_a=ta, transpose_b=tb, m=m, n=n, k=k)
        assert num_eq(c_[:m, :n], want, 1e-12)


@pytest.mark.parametrize("tcode,dt", [(0, np.float32), (1, np.float64), (2, np.complex64), (3, np.complex128)])
def test_backends_agree_all_types(tcode, dt):
    """netlib restatement vs host OpenBLAS on every routine and type, incl.
    complex conjugate-transpose gemm, nrhs>1, bit-exact ipiv."""
    rng = np.random.default_rng(100 + tcode)
    cplx = np.issubdtype(dt, np.complexfloating)
    eps = np.finfo(dt).eps
    def rnd(*s):
        x = rng.uniform(-1, 1, s)
        if cplx:
            x = x + 1j * rng.uniform(-1, 1, s)
        return x.astype(dt)
    n, nrhs = 37, 5
    a, b = rnd(n, n), rnd(n, nrhs)
    nl, ob = O.netlib(), O.openblas()
    f1, f2 = O.lu(a, nl), O.lu(a, ob)
    assert np.array_equal(f1.ipiv_internal, f2.ipiv_internal)
    assert np.allclose(f1.lu, f2.lu, rtol=0, atol=200 * eps * n)
    x1, x2 = O.solve(a, b, nl), O.solve(a, b, ob)
    for x in (x1, x2, O.solve_lu(f1, b, nl), O.solve_lu(f2, b, ob)):
        res = np.linalg.norm(a.astype(np.complex128) @ x - b) / (n * eps * np.linalg.norm(a) * np.linalg.norm(x))
        assert res <= 10
    g = rnd(n, n)
    spd = (g @ g.conj().T + n * np.eye(n)).astype(dt)
    c1, c2 = O.cholesky(spd, nl), O.cholesky(spd, ob)
    assert np.allclose(c1.left, c2.left, rtol=0, atol=100 * eps * n)
    assert np.allclose(c1.left @ c1.left.conj().T, spd, atol=100 * eps * n * n)
    y1, y2 = O.solve_cholesky(c1, b, nl), O.solve_cholesky(c2, b, ob)
    assert np.allclose(y1, y2, atol=1e3 * eps)
    for y in (y1, y2):
        # LLA shares the row-major L untransposed, so LAPACK sees A^T = conj(A):
        # for complex types the reference path solves conj(A) X = B (src/linear-algebra.lisp:296-301).
        res = np.linalg.norm(spd.conj().astype(np.complex128) @ y - b) / (n * eps * np.linalg.norm(spd) * np.linalg.norm(y))
        assert res <= 10
    m, k = 13, 21
    for ta in (False, True):
        for tb in (False, True):
            A = rnd(*((k, m) if ta else (m, k)))
            B = rnd(*((n, k) if tb else (k, n)))
            C = rnd(m, n)
            al, bt = (0.7 - 0.2j, 1.3 + 0.5j) if cplx else (0.7, 1.3)
            want = al * ((A.conj().T if ta else A) @ (B.conj().T if tb else B)) + bt * C
            for be in (nl, ob):
                got = O.gemm(al, A, B, bt, C.copy(), be, transpose_a=ta, transpose_b=tb)
                assert np.allclose(got, want, atol=50 * eps * k)


def test_rectangular_lu_backends_agree():
    rng = np.random.default_rng(5)
    for shape in [(9, 5), (5, 9), (64, 17)]:
        a = rng.uniform(0, 10, shape)
        f1, f2 = O.lu(a, O.netlib()), O.lu(a, O.openblas())
        assert np.array_equal(f1.ipiv_internal, f2.ipiv_internal)
        assert np.allclose(f1.lu, f2.lu, atol=1e-12)


def test_ipiv_bit_exact_midsize():
    """SURVEY hard part 3: pivots are robust for continuous random input."""
    rng = np.random.default_rng(3)
    a = rng.uniform(0, 10, (300, 300))
    assert np.array_equal(O.lu(a, O.netlib()).ipiv_internal, O.lu(a, O.openblas()).ipiv_internal)


# ------------------------------------------------------------------ "next" rows: mm-hermitian%, trsm%, posv
@pytest.mark.parametrize("be", ["netlib", "openblas"])
def test_next_rows_reference_goldens(be):
    """tests/linear-algebra.lisp:117-124 (mm-hermitian), :167-181 (mm-solve-triangular), :366-385 ((solve a b) on the
    hermitian matrix of the cholesky test)."""
    be = getattr(O, be)()
    a = np.array([[1.0, 2.0], [3.0, 4.0]])
    assert np.array_equal(O.mm_hermitian(a, False, be), a @ a.T)
    assert np.array_equal(O.mm_hermitian(a, True, be), a.T @ a)
    u, l, b = np.array([[1.0, 2.0], [0.0, 3.0]]), np.array([[1.0, 0.0], [2.0, 3.0]]), np.array([[5.0, 6.0], [7.0, 8.0]])
    assert np.allclose(u @ O.trsm(u, True, b, be), b, atol=1e-14)
    assert np.allclose(l @ O.trsm(l, False, b, be), b, atol=1e-14)
    spd = np.array([[2.0, -1, 0], [-1, 2, -1], [0, -1, 2]])
    rhs = np.array([5.0, 7.0, 13.0])
    assert np.allclose(O.solve_hermitian(spd, rhs, be), np.linalg.solve(spd, rhs), atol=1e-13)
    # only the row-major lower triangle of a hermitian-matrix is read
    junk = spd.copy(); junk[np.triu_indices(3, 1)] = 99.0
    assert np.allclose(O.solve_hermitian(junk, rhs, be), np.linalg.solve(spd, rhs), atol=1e-13)


def test_next_rows_backends_agree():
    rng = np.random.default_rng(31)
    r = rng.uniform(-1, 1, (40, 23))
    for tl in (False, True):
        assert np.allclose(O.mm_hermitian(r, tl, O.netlib()), O.mm_hermitian(r, tl, O.openblas()), atol=1e-13)
    t = np.triu(rng.uniform(1, 2, (30, 30)))
    b = rng.uniform(-1, 1, (30, 4))
    assert np.allclose(O.trsm(t, True, b, O.netlib()), O.trsm(t, True, b, O.openblas()), rtol=1e-10)
    assert np.allclose(O.trsm(t.T.copy(), False, b, O.netlib()), O.trsm(t.T.copy(), False, b, O.openblas()), rtol=1e-10)


@pytest.mark.parametrize("be", ["netlib", "openblas"])
def test_invert_reference_goldens(be):
    """tests/linear-algebra.lisp:229-244 (invert) and :366-385 ((invert c) of the cholesky test), on the netlib
    restatement (xTRTI2, xLAUUM, unblocked xGETRI with the LWORK = -1 query) and on host OpenBLAS."""
    be = getattr(O, be)()
    m = np.array([[1.0, 2.0], [3.0, 4.0]])
    assert np.allclose(O.invert(m, be), [[-2, 1], [1.5, -0.5]], atol=1e-14)
    assert np.allclose(O.invert(np.triu(m), be, "upper"), [[1, -0.5], [0, 0.25]], atol=1e-14)
    assert np.allclose(O.invert(np.tril(m), be, "lower"), [[1, 0], [-0.75, 0.25]], atol=1e-14)
    spd = np.array([[2.0, -1, 0], [-1, 2, -1], [0, -1, 2]])
    c = O.cholesky(spd, be)
    assert np.allclose(O.invert(c.left, be, "cholesky"), np.linalg.inv(spd), atol=1e-14)


def test_invert_backends_agree():
    rng = np.random.default_rng(77)
    a = rng.uniform(0, 10, (60, 60))
    assert np.allclose(O.invert(a, O.netlib()), O.invert(a, O.openblas()), rtol=1e-9, atol=1e-12)
    spd = a @ a.T + 60 * np.eye(60)
    l = O.cholesky(spd, O.netlib()).left
    assert np.allclose(O.invert(l, O.netlib(), "cholesky"), O.invert(l, O.openblas(), "cholesky"), rtol=1e-9, atol=1e-14)
    t = np.triu(rng.uniform(-1, 1, (50, 50))) + 4 * np.eye(50)
    for kind, mat in (("upper", t), ("lower", t.T.copy())):
        assert np.allclose(O.invert(mat, O.netlib(), kind), O.invert(mat, O.openblas(), kind), rtol=1e-10, atol=1e-14)
    with pytest.raises(O.LapackFailure) as e:
        O.invert(np.array([[1.0, 2.0], [2.0, 4.0]]), O.netlib())
    assert e.value.info == 2
