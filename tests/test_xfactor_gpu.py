"""GPU parity of the single, complex-single and complex-double factorisations, solves and hermitian products
({s,c,z}getrf_ / getrs_ / gesv_ / potrf_ / potrs_ / posv_, ssyrk_ / cherk_ / zherk_: csrc/xfactor.cu) through the
host mirror of LLA's generic functions, against the oracle (LLA's calls on the netlib restatement and on host OpenBLAS).

LLA dispatches lu / solve / cholesky / mm on the common float type of the operands (reference
src/linear-algebra.lisp:62-76,225-234,264-301,666-674).  Bars: ipiv bit-exact with both oracle back ends (complex
pivot rule |re| + |im|, first maximum), factors within k n eps of the oracle's, scaled backward residual <= 10 in the
operand precision.
"""
import numpy as np
import pytest

from lla_b200 import lla as L
from oracle import lla_oracle as O

pytestmark = pytest.mark.gpu

TYPES = [np.float32, np.complex64, np.complex128]


def eps_of(dt):
    return np.finfo(np.float32 if dt in (np.float32, np.complex64) else np.float64).eps


def rand(rng, shape, dt, lo=0.0, hi=10.0):
    a = rng.uniform(lo, hi, shape)
    if np.issubdtype(dt, np.complexfloating):
        a = a + 1j * rng.uniform(lo, hi, shape)
    return a.astype(dt)


def residual(a, x, b, eps):
    a64, x64, b64 = (np.asarray(v, dtype=np.complex128) for v in (a, x, b))
    n = a.shape[0]
    return np.linalg.norm(a64 @ x64 - b64) / (n * eps * np.linalg.norm(a64) * np.linalg.norm(x64))


@pytest.mark.parametrize("dt", TYPES)
@pytest.mark.parametrize("shape", [(1, 1), (7, 7), (33, 33), (64, 64), (100, 100), (257, 257), (300, 300), (9, 5), (5, 9),
                                   (200, 70), (70, 200), (520, 520)])
def test_lu_vs_oracle(dt, shape):
    rng = np.random.default_rng(shape[0] * 31 + shape[1] + (0 if dt == np.float32 else 5))
    a = rand(rng, shape, dt)
    f = L.lu(a)
    r1, r2 = O.lu(a, O.netlib()), O.lu(a, O.openblas())
    assert f.lu.dtype == dt
    if np.array_equal(r1.ipiv_internal, r2.ipiv_internal):      # the reference LAPACKs agree: so must we
        assert np.array_equal(f.ipiv_internal, r1.ipiv_internal), np.flatnonzero(f.ipiv_internal != r1.ipiv_internal)[:5]
        assert np.max(np.abs(f.lu - r1.lu)) <= 100 * max(shape) * eps_of(dt) * np.abs(r1.lu).max()
    # reconstruction P A = L U (lu-random-test, reference tests/linear-algebra.lisp:109-115) holds in any case
    m, n = shape
    p = np.arange(m)
    for i, pi in enumerate(f.ipiv_internal):             # LAPACK's swap sequence replayed on the m row indices
        p[i], p[pi - 1] = p[pi - 1], p[i]
    k = min(m, n)
    lo = np.tril(f.lu.astype(np.complex128), -1)[:, :k] + np.eye(m, k)
    up = np.triu(f.lu.astype(np.complex128))[:k, :]
    assert np.max(np.abs(a.astype(np.complex128)[p, :] - lo @ up)) <= 30 * max(shape) * eps_of(dt) * np.abs(a).max() * max(1, np.abs(up).max() / np.abs(a).max())


@pytest.mark.parametrize("dt", TYPES)
def test_complex_pivot_rule_and_ties(dt):
    """IZAMAX compares |re| + |im|, not the modulus: 3+3i (6, modulus 4.24) beats 5 (5, modulus 5); ties take the first."""
    if dt == np.float32:
        a = np.array([[1, 2, 3], [-4, 1, 0], [4, 5, 6]], dtype=dt)      # |-4| == |4|: the first wins
    else:
        a = np.array([[5, 1, 2], [3 + 3j, 2, 1], [0, 1j, 4]], dtype=dt)
    f, r = L.lu(a), O.lu(a, O.netlib())
    assert np.array_equal(f.ipiv_internal, r.ipiv_internal)
    assert f.ipiv_internal[0] == 2
    assert np.allclose(f.lu, r.lu, rtol=50 * eps_of(dt))


@pytest.mark.parametrize("dt", TYPES)
@pytest.mark.parametrize("n,nrhs", [(5, 1), (64, 3), (300, 7), (1100, 4)])
def test_solve_paths(dt, n, nrhs):
    rng = np.random.default_rng(n + nrhs)
    a = rand(rng, (n, n), dt)
    b = rand(rng, (n, nrhs), dt) if nrhs > 1 else rand(rng, (n,), dt)
    eps = eps_of(dt)
    x = L.solve(a, b)                                   # xGESV
    assert x.dtype == dt and x.shape == b.shape
    bb, xx = (b if nrhs > 1 else b[:, None]), (x if nrhs > 1 else x[:, None])
    assert residual(a, xx, bb, eps) <= 10
    f = L.lu(a)
    x2 = L.solve(f, b)                                  # xGETRS on the stored factors
    assert residual(a, x2 if nrhs > 1 else x2[:, None], bb, eps) <= 10
    xr = O.solve(a, b, O.openblas())
    assert np.max(np.abs(x - xr)) <= 1e4 * n * eps * np.abs(xr).max()


@pytest.mark.parametrize("dt", TYPES)
def test_singular_matrix_reports_info_and_leaves_b(dt):
    a = np.full((2, 2), 0.5, dtype=dt)                  # lu-singular, reference tests/linear-algebra.lisp:106-107
    f, r = L.lu(a), O.lu(a, O.netlib())
    assert np.array_equal(f.ipiv_internal, r.ipiv_internal)
    assert np.array_equal(f.lu, r.lu)
    with pytest.raises(L.LapackFailure) as e:
        L.solve(a, np.ones(2, dtype=dt))
    assert e.value.info == 2


def spd(rng, n, dt):
    g = rand(rng, (n, n), dt, -1, 1)
    a = g @ g.conj().T + n * np.eye(n)
    return ((a + a.conj().T) / 2).astype(dt)


@pytest.mark.parametrize("dt", TYPES)
@pytest.mark.parametrize("n", [1, 3, 31, 32, 33, 100, 257, 700])
def test_cholesky_vs_oracle(dt, n):
    rng = np.random.default_rng(n)
    a = spd(rng, n, dt)
    c = L.cholesky(a)
    r = O.cholesky(a, O.netlib())
    eps = eps_of(dt)
    assert c.left.dtype == dt
    assert np.max(np.abs(np.triu(c.left, 1))) == 0
    assert np.max(np.abs(c.left - r.left)) <= 40 * n * eps * np.abs(r.left).max()
    l = c.left.astype(np.complex128)
    assert np.max(np.abs(l @ l.conj().T - a)) <= 20 * n * eps * np.abs(a).max()
    assert np.all(np.imag(np.diag(c.left)) == 0)
    b = rand(rng, (n, 3), dt)
    # The reference shares the row-major buffers untransposed, so LAPACK sees the column-major view A^T = conj(A):
    # for complex arrays (solve cholesky b) and (solve hermitian b) return the solution of conj(A) X = B (checked on the
    # oracle with both back ends).  Parity means reproducing exactly that.
    ac = a.conj()
    x = L.solve(c, b)                                   # xPOTRS
    assert residual(ac, x, b, eps) <= 10
    xr = O.solve_cholesky(r, b, O.netlib())
    assert np.max(np.abs(x - xr)) <= 1e3 * n * eps * np.abs(xr).max()
    x2 = L.solve(L.Hermitian(a), b)                     # xPOSV
    assert residual(ac, x2, b, eps) <= 10
    assert np.max(np.abs(x2 - O.solve_hermitian(a, b, O.netlib()))) <= 1e3 * n * eps * np.abs(xr).max()


@pytest.mark.parametrize("dt", TYPES)
def test_cholesky_not_positive_definite(dt):
    rng = np.random.default_rng(2)
    a = spd(rng, 90, dt)
    a[50, 50] = -1
    with pytest.raises(L.LapackFailure) as e:
        L.cholesky(a)
    try:
        O.cholesky(a, O.netlib())
        want = 0
    except O.LapackFailure as ee:
        want = ee.info
    assert e.value.info == want == 51
    # POSV: INFO and an untouched right-hand side
    b = rand(rng, (90, 2), dt)
    with pytest.raises(L.LapackFailure):
        L.solve(L.Hermitian(a), b)


@pytest.mark.parametrize("dt", TYPES)
@pytest.mark.parametrize("shape", [(5, 3), (3, 5), (130, 70), (64, 300)])
def test_mm_hermitian(dt, shape):
    """(mm a t) = A A^H and (mm t a) = A^H A: SYRK for single, HERK for the complex types (reference
    src/linear-algebra.lisp:62-76); the diagonal of a HERK result is exactly real."""
    rng = np.random.default_rng(shape[0] + 7 * shape[1])
    a = rand(rng, shape, dt, -1, 1)
    eps = eps_of(dt)
    for left in (False, True):
        h = L.mm(True, a) if left else L.mm(a, True)
        got = h.as_array()
        ref = O.mm_hermitian(a, left, O.netlib())
        a64 = a.astype(np.complex128)
        exact = a64.conj().T @ a64 if left else a64 @ a64.conj().T
        k = shape[0] if left else shape[1]
        assert got.dtype == dt
        assert np.max(np.abs(got - exact)) <= 8 * k * eps * np.abs(exact).max()
        assert np.max(np.abs(got - ref)) <= 8 * k * eps * np.abs(exact).max()
        assert np.all(np.imag(np.diag(h.elements)) == 0)
