"""Ragged sizes through every schedule of the blocked drivers (csrc/lu_f64.cu: lu_blocked, csrc/potrf_f64.cu:
potrf_blocked): orders above 9216 start in the update-bound phase (bulk update pipelined per 4096-column chunk, the next
panel's columns on a side stream), continue in the chain-bound phase (those columns trail the panel block by block, SM
partition) and end with a partial panel; rectangular shapes end early on either side.  The reference's own tests use
orders of 2-9 (tests/linear-algebra.lisp:89-115); the bar here is the same as at the BASELINE sizes: `ipiv` bit-exact
with LLA's call against host OpenBLAS, factors within 50 n eps, reconstruction, INFO of a late failure.
"""
import numpy as np
import pytest

from lla_b200 import lla as L
from oracle import lla_oracle as O

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


@pytest.mark.parametrize("shape", [(11000, 10500), (10300, 12100), (9999, 9999), (5000, 3333)])
def test_lu_ragged_through_all_phases(shape):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    a = rng.uniform(0, 10, shape)
    f = L.lu(a)
    r = O.lu(a, O.openblas())
    assert np.array_equal(f.ipiv_internal, r.ipiv_internal), np.flatnonzero(f.ipiv_internal != r.ipiv_internal)[:5]
    assert np.max(np.abs(f.lu - r.lu)) <= 50 * max(shape) * EPS * np.abs(r.lu).max()
    # P A = L U on a sample of rows (lu-random-test at size)
    m, n = shape
    k = min(m, n)
    p = np.arange(m)
    for i, pi in enumerate(f.ipiv_internal):
        p[i], p[pi - 1] = p[pi - 1], p[i]
    rows = rng.choice(m, 64, replace=False)
    lo = np.tril(f.lu, -1)[:, :k]
    up = np.triu(f.lu)[:k, :]
    rec = lo[rows, :] @ up
    rec[:, :] += np.where(rows[:, None] < k, up[np.minimum(rows, k - 1), :], 0.0)     # unit diagonal of L
    assert np.max(np.abs(a[p[rows], :] - rec)) <= 10 * max(shape) * EPS * np.abs(a).max() * max(1.0, np.abs(up).max() / 10)


@pytest.mark.parametrize("n", [4999, 12500])
def test_cholesky_ragged_through_all_phases(n):
    rng = np.random.default_rng(n)
    g = rng.uniform(0, 1, (n, 64))
    spd = g @ g.T
    spd[np.diag_indices(n)] += n
    c = L.cholesky(spd)
    r = O.cholesky(spd, O.openblas())
    assert np.max(np.abs(np.triu(c.left, 1))) == 0
    assert np.max(np.abs(c.left - r.left)) <= 20 * n * EPS * np.abs(r.left).max()
    rows = rng.choice(n, 64, replace=False)
    assert np.max(np.abs(c.left[rows, :] @ c.left.T - spd[rows, :])) <= 10 * n * EPS * np.abs(spd).max()
    # the first non-positive minor far into the factorisation: INFO as the reference reports it, nothing after it needed
    bad = spd.copy()
    bad[n - 300, n - 300] = -1.0
    with pytest.raises(L.LapackFailure) as e:
        L.cholesky(bad)
    assert e.value.info == n - 299
