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


# ------------------------------------------------------------------ second fixture file: CGEMM and the "next" rows
G2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nextrows_golden.npz"))
CAB = (np.complex64(0.7 - 0.2j), np.complex64(1.3 + 0.5j))


def _cgemm_tol(i):
    a, b = G2[f"cgemm{i}_a"], G2[f"cgemm{i}_b"]
    ta, tb = (bool(v) for v in G2[f"cgemm{i}_flags"])
    opa = np.abs(a.conj().T if ta else a).astype(np.float64)
    opb = np.abs(b.conj().T if tb else b).astype(np.float64)
    k = 4 * (16 + 4 * np.sqrt(opa.shape[1]))
    return (k * EPS["s"] * abs(CAB[0]) * np.linalg.norm(opa, axis=1)[:, None] * np.linalg.norm(opb, axis=0)[None, :]
            + 8 * EPS["s"] * (abs(CAB[1]) * np.abs(G2[f"cgemm{i}_c"]) + np.abs(G2[f"cgemm{i}_out"])))


def test_oracle_reproduces_nextrows_goldens():
    for be in (O.netlib(), O.openblas()):
        for i in range(4):
            ta, tb = (bool(v) for v in G2[f"cgemm{i}_flags"])
            got = O.gemm(CAB[0], G2[f"cgemm{i}_a"], G2[f"cgemm{i}_b"], CAB[1], G2[f"cgemm{i}_c"].copy(), be, transpose_a=ta, transpose_b=tb)
            assert np.all(np.abs(got - G2[f"cgemm{i}_out"]) <= _cgemm_tol(i))
        for i in range(3):
            a = G2[f"syrk{i}_a"]
            assert np.allclose(O.mm_hermitian(a, False, be), G2[f"syrk{i}_aat"], atol=1e-12)
            assert np.allclose(O.mm_hermitian(a, True, be), G2[f"syrk{i}_ata"], atol=1e-12)
            u, b = G2[f"trsm{i}_u"], G2[f"trsm{i}_b"]
            assert np.allclose(O.trsm(u, True, b, be), G2[f"trsm{i}_xu"], rtol=1e-9, atol=1e-12)
            assert np.allclose(O.trsm(u.T.copy(), False, b, be), G2[f"trsm{i}_xl"], rtol=1e-9, atol=1e-12)
            assert np.allclose(O.solve_hermitian(G2[f"posv{i}_a"], b, be), G2[f"posv{i}_x"], rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
def test_cuda_nextrows_match_golden():
    from lla_b200 import lla as L
    for i in range(4):
        ta, tb = (bool(v) for v in G2[f"cgemm{i}_flags"])
        got = L.gemm(CAB[0], G2[f"cgemm{i}_a"], G2[f"cgemm{i}_b"], CAB[1], G2[f"cgemm{i}_c"].copy(), transpose_a=ta, transpose_b=tb)
        assert got.dtype == np.complex64
        assert np.all(np.abs(got.astype(np.complex128) - G2[f"cgemm{i}_out"]) <= _cgemm_tol(i))
    for i in range(3):
        a = G2[f"syrk{i}_a"]
        assert np.allclose(L.mm(a, True).as_array(), G2[f"syrk{i}_aat"], atol=1e-12)
        assert np.allclose(L.mm(True, a).as_array(), G2[f"syrk{i}_ata"], atol=1e-12)
        u, b = G2[f"trsm{i}_u"], G2[f"trsm{i}_b"]
        assert np.allclose(L.solve(L.UpperTriangular(u), b), G2[f"trsm{i}_xu"], rtol=1e-9, atol=1e-12)
        assert np.allclose(L.solve(L.LowerTriangular(u.T.copy()), b), G2[f"trsm{i}_xl"], rtol=1e-9, atol=1e-12)
        assert np.allclose(L.solve(L.Hermitian(G2[f"posv{i}_a"]), b), G2[f"posv{i}_x"], rtol=1e-9, atol=1e-12)
