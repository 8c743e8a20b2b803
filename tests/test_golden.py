"""Committed golden vectors (tests/golden/hotpath_golden.npz, made by tests/golden/make_golden.py).
CPU: both oracle back ends reproduce them.  GPU: the CUDA path through the Fortran drop-ins matches them."""
import os

import numpy as np
import pytest

from oracle import lla_oracle as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_golden.npz"))
EPS = {"d": np.finfo(np.float64).eps, "z": np.finfo(np.float64).eps, "s": float(np.finfo(np.float32).eps)}


def gemm_cases():
    return sorted({k.split("_")[0] for k in G.files if k.startswith("gemm")})


def tol_gemm(case):
    a, b = G[f"{case}_a"], G[f"{case}_b"]
    ta, tb = (bool(v) for v in G[f"{case}_flags"])
    alpha, beta = G[f"{case}_alphabeta"]
    opa = np.abs(a.conj().T if ta else a).astype(np.float64)
    opb = np.abs(b.conj().T if tb else b).astype(np.float64)
    K = opa.shape[1]
    k = (16 + 4 * np.sqrt(K)) if case[-1] == "s" else 8 * np.sqrt(K)
    return (k * EPS[case[-1]] * abs(alpha) * np.linalg.norm(opa, axis=1)[:, None] * np.linalg.norm(opb, axis=0)[None, :]
            + 8 * EPS[case[-1]] * (abs(beta) * np.abs(G[f"{case}_c"]) + np.abs(G[f"{case}_out"])))


@pytest.mark.parametrize("case", gemm_cases())
def test_oracle_reproduces_gemm_golden(case):
    ta, tb = (bool(v) for v in G[f"{case}_flags"])
    alpha, beta = G[f"{case}_alphabeta"]
    if case[-1] != "z":
        alpha, beta = float(alpha.real), float(beta.real)
    for be in (O.netlib(), O.openblas()):
        got = O.gemm(alpha, G[f"{case}_a"], G[f"{case}_b"], beta, G[f"{case}_c"].copy(), be, transpose_a=ta, transpose_b=tb)
        assert np.all(np.abs(got - G[f"{case}_out"]) <= tol_gemm(case))


def test_oracle_reproduces_factor_goldens():
    for i in range(7):
        a = G[f"lu{i}_a"]
        for be in (O.netlib(), O.openblas()):
            f = O.lu(a, be)
            assert np.array_equal(f.ipiv_internal, G[f"lu{i}_ipiv"])
            assert np.allclose(f.lu, G[f"lu{i}_lu"], atol=1e-11)
    for i in range(4):
        for be in (O.netlib(), O.openblas()):
            assert np.allclose(O.cholesky(G[f"chol{i}_a"], be).left, G[f"chol{i}_l"], atol=1e-11)
    assert np.allclose(O.cholesky(G["ref_chol_a"], O.netlib()).left, G["ref_chol_l"], atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("case", gemm_cases())
def test_cuda_gemm_matches_golden(case):
    from lla_b200 import lla as L
    ta, tb = (bool(v) for v in G[f"{case}_flags"])
    alpha, beta = G[f"{case}_alphabeta"]
    if case[-1] != "z":
        alpha, beta = float(alpha.real), float(beta.real)
    got = L.gemm(alpha, G[f"{case}_a"], G[f"{case}_b"], beta, G[f"{case}_c"].copy(), transpose_a=ta, transpose_b=tb)
    assert got.dtype == G[f"{case}_out"].dtype
    assert np.all(np.abs(got.astype(np.complex128) - G[f"{case}_out"]) <= tol_gemm(case))


@pytest.mark.gpu
def test_cuda_factorisations_match_golden():
    from lla_b200 import lla as L
    eps = EPS["d"]
    for i in range(7):
        a = G[f"lu{i}_a"]
        f = L.lu(a)
        assert np.array_equal(f.ipiv_internal, G[f"lu{i}_ipiv"]), i          # bit-exact pivots
        if max(a.shape) <= 32:
            assert np.array_equal(f.lu, G[f"lu{i}_lu"])                        # single leaf: bit-identical LU
        assert np.max(np.abs(f.lu - G[f"lu{i}_lu"])) <= 50 * max(a.shape) * eps * np.abs(G[f"lu{i}_lu"]).max()
        if f"lu{i}_b" in G.files:
            x = L.solve(a, G[f"lu{i}_b"])
            n = a.shape[0]
            assert np.linalg.norm(a @ x - G[f"lu{i}_b"]) / (n * eps * np.linalg.norm(a) * np.linalg.norm(x)) <= 10
            assert np.allclose(x, G[f"lu{i}_x"], rtol=1e-8, atol=1e-10)
    for i in range(4):
        a = G[f"chol{i}_a"]
        c = L.cholesky(a)
        n = a.shape[0]
        assert np.max(np.abs(c.left - G[f"chol{i}_l"])) <= 20 * n * eps * np.abs(G[f"chol{i}_l"]).max()
        assert np.allclose(L.solve(c, G[f"chol{i}_b"]), G[f"chol{i}_x"], rtol=1e-9, atol=1e-12)
    assert np.allclose(L.cholesky(G["ref_chol_a"]).left, G["ref_chol_l"], atol=1e-6)
    assert np.array_equal(L.mm(G["ref_mm_a"], np.array([1.0, 2.0])), [5, 11, 17])
