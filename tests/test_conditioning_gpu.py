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


@pytest.mark.parametrize("n", [200, 1000])
def test_triangular_solve_graded_diagonal(n):
    """upper-triangular system with a diagonal graded over 12 orders of magnitude (trsm%, reference src/linear-algebra.lisp:333-347)"""
    rng = np.random.default_rng(n + 2)
    u = np.triu(rng.normal(size=(n, n)))
    u[np.diag_indices(n)] = np.logspace(0, -12, n) * np.sign(rng.normal(size=n))
    b = rng.normal(size=(n, 2))
    x = L.solve(L.UpperTriangular(u), b)
    # row-wise backward error of a triangular solve: |U x - b|_i <= c n eps (|U| |x|)_i
    num = np.abs(u @ x - b)
    den = (np.abs(u) @ np.abs(x)) * n * EPS
    assert np.max(num / den) <= 10
