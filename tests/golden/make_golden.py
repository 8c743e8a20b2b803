"""Generates tests/golden/hotpath_golden.npz: seeded inputs and the ORACLE's outputs for the hot path.

The reference (tpapp/lla) is Common Lisp and cannot run in the build image, so these vectors come
from the oracle's netlib restatement (oracle/netlib_ref.c through oracle/lla_oracle.py), which is
itself pinned against the reference's own goldens in tests/test_oracle.py; ipiv additionally had to
agree with host OpenBLAS when the file was made.  Run from the repository root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lla_oracle as O  # noqa: E402

out = {}
nl, ob = O.netlib(), O.openblas()
rng = np.random.default_rng(20261019)

# the reference's own literal goldens (tests/linear-algebra.lisp:9-29, :89-104, :366-385)
out["ref_mm_a"] = np.array([[1, 2], [3, 4], [5, 6]], dtype=np.float64)
out["ref_chol_a"] = np.array([[2, -1, 0], [-1, 2, -1], [0, -1, 2]], dtype=np.float64)
out["ref_chol_l"] = np.array([[1.414214, 0, 0], [-0.7071068, 1.2247449, 0], [0.0, -0.8164966, 1.1547005]])

for i, (m, n, k, ta, tb) in enumerate([(7, 5, 3, False, False), (33, 17, 29, True, False), (64, 40, 50, False, True), (20, 31, 9, True, True)]):
    for tag, dt in (("d", np.float64), ("s", np.float32), ("z", np.complex128)):
        def r(*s):
            x = rng.uniform(-1, 1, s)
            if dt == np.complex128:
                x = x + 1j * rng.uniform(-1, 1, s)
            return x.astype(dt)
        a = r(*((k, m) if ta else (m, k)))
        b = r(*((n, k) if tb else (k, n)))
        c = r(m, n)
        alpha, beta = (0.7 - 0.2j, 1.3 + 0.5j) if dt == np.complex128 else (0.7, 1.3)
        out[f"gemm{i}{tag}_a"], out[f"gemm{i}{tag}_b"], out[f"gemm{i}{tag}_c"] = a, b, c
        out[f"gemm{i}{tag}_flags"] = np.array([ta, tb])
        out[f"gemm{i}{tag}_alphabeta"] = np.array([alpha, beta])
        out[f"gemm{i}{tag}_out"] = O.gemm(alpha, a, b, beta, c.copy(), nl, transpose_a=ta, transpose_b=tb)

for i, shape in enumerate([(6, 6), (32, 32), (45, 45), (130, 130), (40, 25), (25, 40), (160, 160)]):
    a = rng.uniform(0, 10, shape)
    f1, f2 = O.lu(a, nl), O.lu(a, ob)
    assert np.array_equal(f1.ipiv_internal, f2.ipiv_internal)
    out[f"lu{i}_a"], out[f"lu{i}_lu"], out[f"lu{i}_ipiv"] = a, f1.lu, f1.ipiv_internal
    if shape[0] == shape[1]:
        b = rng.uniform(0, 10, (shape[0], 3))
        out[f"lu{i}_b"], out[f"lu{i}_x"] = b, O.solve(a, b, nl)

for i, n in enumerate([5, 32, 97, 140]):
    g = rng.uniform(0, 1, (n, n))
    spd = g @ g.T + n * np.eye(n)
    b = rng.uniform(0, 1, (n, 2))
    c = O.cholesky(spd, nl)
    out[f"chol{i}_a"], out[f"chol{i}_l"], out[f"chol{i}_b"] = spd, c.left, b
    out[f"chol{i}_x"] = O.solve_cholesky(c, b, nl)

np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz"), **out)
print("wrote", len(out), "arrays")

# ---------------------------------------------------------------------------------------------------------------
# second file (added later in round 1, own seed so that the file above stays byte-identical): complex-single GEMM and
# the "next" rows of SURVEY section 8f -- (mm a t) / (mm t a), solve on triangular and hermitian matrices.
out2 = {}
rng2 = np.random.default_rng(20261020)
for i, (m, n, k, ta, tb) in enumerate([(7, 5, 3, False, False), (33, 17, 29, True, False), (64, 40, 50, False, True), (20, 31, 9, True, True)]):
    def rc(*s):
        return (rng2.uniform(-1, 1, s) + 1j * rng2.uniform(-1, 1, s)).astype(np.complex64)
    a, b, c = rc(*((k, m) if ta else (m, k))), rc(*((n, k) if tb else (k, n))), rc(m, n)
    alpha, beta = np.complex64(0.7 - 0.2j), np.complex64(1.3 + 0.5j)
    out2[f"cgemm{i}_a"], out2[f"cgemm{i}_b"], out2[f"cgemm{i}_c"] = a, b, c
    out2[f"cgemm{i}_flags"] = np.array([ta, tb])
    out2[f"cgemm{i}_out"] = O.gemm(alpha, a, b, beta, c.copy(), nl, transpose_a=ta, transpose_b=tb)
for i, (m, k) in enumerate([(6, 4), (40, 23), (150, 70)]):
    r = rng2.uniform(-1, 1, (m, k))
    out2[f"syrk{i}_a"] = r
    out2[f"syrk{i}_aat"], out2[f"syrk{i}_ata"] = O.mm_hermitian(r, False, nl), O.mm_hermitian(r, True, nl)
    t = np.triu(rng2.uniform(-1, 1, (m, m))) + 4 * np.eye(m)
    b = rng2.uniform(-1, 1, (m, 3))
    out2[f"trsm{i}_u"], out2[f"trsm{i}_b"] = t, b
    out2[f"trsm{i}_xu"], out2[f"trsm{i}_xl"] = O.trsm(t, True, b, nl), O.trsm(t.T.copy(), False, b, nl)
    g = rng2.uniform(0, 1, (m, m))
    spd = g @ g.T + m * np.eye(m)
    out2[f"posv{i}_a"], out2[f"posv{i}_x"] = spd, O.solve_hermitian(spd, b, nl)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nextrows_golden.npz"), **out2)
print("wrote", len(out2), "arrays (next rows)")
