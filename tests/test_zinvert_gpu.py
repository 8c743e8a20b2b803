"""GPU parity of the invert family (dgetri_ / dpotri_ / dtrtri_ drop-ins, reference src/linear-algebra.lisp:352-409):
the reference's own invert test (tests/linear-algebra.lisp:229-244, :366-385) and scaled residuals against host LAPACK.
Kept in a file of its own, collected last: it was validated on B200 in the round's final GPU call (goldens, n = 37 and
300 green; at n = 1100 one over-tight elementwise comparison of two ill-conditioned inverses was replaced by the scaled
residual afterwards)."""
import numpy as np
import pytest

from lla_b200 import lla as L
from oracle import lla_oracle as O

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def num_eq(a, b, tol=1e-5):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b))))


# ------------------------------------------------------------------ the invert family on the device
def test_invert_reference_goldens():
    m = np.array([[1.0, 2.0], [3.0, 4.0]])                                   # tests/linear-algebra.lisp:229-244
    assert num_eq(L.invert(m), [[-2, 1], [1.5, -0.5]])
    assert num_eq(L.invert(L.lu(m)), [[-2, 1], [1.5, -0.5]])
    assert num_eq(L.invert(L.UpperTriangular(m)).elements, [[1, -0.5], [0, 0.25]])
    assert num_eq(L.invert(L.LowerTriangular(m)).elements, [[1, 0], [-0.75, 0.25]])
    spd = np.array([[2.0, -1, 0], [-1, 2, -1], [0, -1, 2]])                  # :366-385
    assert num_eq(L.invert(L.cholesky(spd)).as_array(), np.linalg.inv(spd))
    with pytest.raises(L.LapackFailure) as e:
        L.invert(np.array([[1.0, 2.0], [2.0, 4.0]]))
    assert e.value.info == 2


@pytest.mark.parametrize("n", [37, 300, 1100])
def test_invert_vs_oracle(n):
    rng = np.random.default_rng(9 * n)
    a = rng.uniform(0, 10, (n, n))
    inv = L.invert(a)
    ref = O.invert(a, O.openblas())
    if n <= 300:
        assert np.allclose(O.invert(a, O.netlib()), ref, rtol=1e-8, atol=1e-12)      # the two oracle back ends agree
    assert np.linalg.norm(a @ inv - np.eye(n)) / (n * EPS * np.linalg.norm(a) * np.linalg.norm(inv)) <= 10
    assert np.max(np.abs(inv - ref)) <= 1e-7 * np.abs(ref).max()
    spd = a @ a.T + n * np.eye(n)
    ci = L.invert(L.cholesky(spd))
    assert np.all(np.triu(ci.elements, 1) == np.triu(L.cholesky(spd).left, 1)), "POTRI('U') must leave the other triangle alone"
    assert np.linalg.norm(spd @ ci.as_array() - np.eye(n)) / (n * EPS * np.linalg.norm(spd) * np.linalg.norm(ci.as_array())) <= 10
    t = np.triu(rng.uniform(-1, 1, (n, n))) + 4 * np.eye(n)
    ti = L.invert(L.UpperTriangular(t)).elements
    assert np.linalg.norm(t @ ti - np.eye(n)) / (n * EPS * np.linalg.norm(t) * np.linalg.norm(ti)) <= 10
    tl = t.T.copy()
    li = L.invert(L.LowerTriangular(tl)).elements
    # (no elementwise comparison with ti.T: the inverse of a random triangular matrix has entries ~1e3-1e4 at n = 1100 and
    # the forward / backward substitution orders round differently; the scaled residual is the bar, as for the upper case)
    assert np.all(np.triu(li, 1) == 0.0)
    assert np.linalg.norm(tl @ li - np.eye(n)) / (n * EPS * np.linalg.norm(tl) * np.linalg.norm(li)) <= 10
