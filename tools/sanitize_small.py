"""Small invocations of the shared-memory-heavy kernels for compute-sanitizer (racecheck / memcheck): Cholesky leaf and
panel, triangular inverses, batched solver, LU leaf."""
import sys
import numpy as np
sys.path.insert(0, ".")
from lla_b200 import lla as L
rng = np.random.default_rng(0)
for n in (100, 128, 300, 1100):
    g = rng.uniform(0, 1, (n, n)); spd = g @ g.T + n * np.eye(n)
    c = L.cholesky(spd)
    assert np.linalg.norm(c.left @ c.left.T - spd) / (n * 2.2e-16 * np.linalg.norm(spd)) <= 10
    b = rng.uniform(0, 1, (n, 2))
    x = L.solve(c, b)
    assert np.linalg.norm(spd @ x - b) / (n * 2.2e-16 * np.linalg.norm(spd) * np.linalg.norm(x)) <= 10
a = rng.uniform(0, 10, (300, 300))
f = L.lu(a)
x = L.solve(a, rng.uniform(0, 1, (300, 3)))
from lla_b200 import batched as Bt
A = rng.uniform(0, 10, (64, 32, 32)); B = rng.uniform(0, 1, (64, 32))
X, piv, info = Bt.solve_batched(A, B)
assert np.all(info == 0) and np.max(np.abs(np.einsum("sij,sj->si", A, X) - B)) < 1e-9
print("sanitize_small ok")
