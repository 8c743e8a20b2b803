"""Graded-condition inputs: the triangular solves of this library run through explicit inverses of the 128 x 128 diagonal
blocks (DMMA GEMMs instead of substitution), which is not backward stable in general.  These tests pin what the drop-ins
deliver on ill-conditioned systems against the same bound the well-conditioned tests use (scaled backward residual
<= 10, north_star) and against the host reference (oracle, OpenBLAS back end)."""
import numpy as np
import pytest

from lla_b200 import lla as L
from oracle import lla_oracle as O

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def backward_residual(a, x, b):
    n = a.shape[0]
    return np.linalg.norm(a @ x - b) / (n * EPS * np.linalg.norm(a) * np.linalg.norm(x))


def graded(n, kappa, rng):
    """random orthogonal factors around a geometrically graded spectrum: condition number kappa"""
    q1, _ = np.linalg.qr(rng.normal(size=(n, n)))
    q2, _ = np.linalg.qr(rng.normal(size=(n, n)))
    s = np.logspace(0, -np.log10(kappa), n)
    return q1, s, q2


@pytest.mark.parametrize("n,kappa", [(300, 1e8), (700, 1e12), (1500, 1e10)])
def test_solve_general_ill_conditioned(n, kappa):
    rng = np.random.default_rng(n)
    q1, s, q2 = graded(n, kappa, rng)
    a = (q1 * s) @ q2.T
    b = rng.normal(size=(n, 3))
    x = L.solve(a, b)
    xr = O.solve(a, b, O.openblas())
    r, rr = backward_residual(a, x, b), backward_residual(a, xr, b)
    assert r <= 10, (r, rr)
    f = L.lu(a)
    assert backward_residual(a, L.solve(f, b), b) <= 10


@pytest.mark.parametrize("n,kappa", [(300, 1e8), (900, 1e12)])
def test_cholesky_ill_conditioned(n, kappa):
    rng = np.random.default_rng(n + 1)
    q1, s, _ = graded(n, kappa, rng)
    a = (q1 * s) @ q1.T
    a = (a + a.T) / 2
    b = rng.normal(size=(n, 2))
    c = L.cholesky(a)
    l = c.left
    assert np.linalg.norm(l @ l.T - a) / (n * EPS * np.linalg.norm(a)) <= 10
    x = L.solve(c, b)
    assert backward_residual(a, x, b) <= 10


@pytest.mark.parametrize("grading", ["rows", "columns"])
@pytest.mark.parametrize("n,off", [(200, 1.0), (1000, 1.0), (1000, 3.0)])
def test_triangular_solve_graded_diagonal(n, off, grading):
    """upper-triangular systems whose diagonal is graded over 12 orders of magnitude by a row or a column scaling of a
    triangle T = +-I + off * N(0,1)/sqrt(n) (trsm%, reference src/linear-algebra.lisp:333-347); off = 3 puts cond(T) near
    2e3.  (An UNSCALED random triangle with such a diagonal has a solution that overflows double precision - substitution
    on the host returns inf for it too - so it pins nothing.)"""
    rng = np.random.default_rng(n + 2)
    t = np.triu(rng.normal(size=(n, n)), 1) * (off / np.sqrt(n))
    t[np.diag_indices(n)] = np.sign(rng.normal(size=n))
    d = np.logspace(0, -12, n)
    u = d[:, None] * t if grading == "rows" else t * d[None, :]
    b = rng.normal(size=(n, 2))
    x = L.solve(L.UpperTriangular(u), b)
    xr = O.trsm(u, True, b, O.openblas())
    assert np.all(np.isfinite(x))
    # row-wise backward error of a triangular solve: |U x - b|_i <= c n eps (|U| |x|)_i
    num = np.abs(u @ x - b)
    den = (np.abs(u) @ np.abs(x)) * n * EPS
    assert np.max(num / den) <= 10
    numr = np.abs(u @ xr - b)                           # and within a constant of what substitution on the host delivers
    assert np.max(num / den) <= 50 * max(np.max(numr / ((np.abs(u) @ np.abs(xr)) * n * EPS)), 1e-3)
