"""GPU parity of the multi-GPU engine (csrc/dist*.cu) through its C entry points, on however many GPUs are visible.

One process / all GPUs (`llacuda_pd*`, host pointers -- what LLA's foreign-funcall can reach) and one rank per process
(`llacuda_*_local` under torchrun).  With one GPU the grid is 1 x 1: the block-cyclic ownership arithmetic, padding,
look-ahead and the rank functions still run, only the collectives are skipped; `gpurun --gpus 2/4/8` exercises the
NCCL paths with the same tests.  Oracle: LLA's calls against the netlib restatement / host OpenBLAS (oracle/), ipiv
bit-exact, residual bounds as north_star states them.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from lla_b200 import _lib, lla as L, pdist as PD
from oracle import lla_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPS = np.finfo(np.float64).eps


def _ndev():
    import torch
    return torch.cuda.device_count()


def _grids(n):
    return [(p, n // p) for p in range(1, n + 1) if n % p == 0]


@pytest.fixture(scope="module")
def engine():
    _lib.init(0)
    n, P, Q = PD.init_all(0)
    assert n == _ndev() and P * Q == n
    yield n
    PD.finalize()
    for k in ("LLACUDA_LU_GRID", "LLACUDA_CHOL_GRID", "LLACUDA_GEMM_GRID", "LLACUDA_DIST_NB", "LLACUDA_GEMM_NB"):
        os.environ.pop(k, None)


def _set(grid_var, P, Q, nb=None, gemm_nb=None):
    os.environ[grid_var] = f"{P}x{Q}"
    if nb:
        os.environ["LLACUDA_DIST_NB"] = str(nb)
    if gemm_nb:
        os.environ["LLACUDA_GEMM_NB"] = str(gemm_nb)


def cm(a):
    """column-major buffer of a matrix given as a NumPy (row-major) array"""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def test_pdgemm_all_transposes_and_padding(engine):
    rng = np.random.default_rng(0)
    for P, Q in _grids(engine):
        _set("LLACUDA_GEMM_GRID", P, Q, gemm_nb=128)
        for (m, n, k) in [(700, 500, 900), (1024, 1024, 1024), (130, 257, 1000)]:
            for ta in "NT":
                for tb in "NC":
                    a = rng.uniform(-1, 1, (k, m) if ta != "N" else (m, k))
                    b = rng.uniform(-1, 1, (n, k) if tb != "N" else (k, n))
                    c0 = rng.uniform(-1, 1, (m, n))
                    af, bf, cf = cm(a), cm(b), cm(c0)
                    PD.pdgemm(ta, tb, m, n, k, 1.5, af, af.shape[0], bf, bf.shape[0], -0.5, cf, m)
                    opa = a.T if ta != "N" else a
                    opb = b.T if tb != "N" else b
                    ref = 1.5 * (opa @ opb) - 0.5 * c0
                    bound = 8 * np.sqrt(k) * EPS * (np.linalg.norm(opa, axis=1)[:, None] * np.linalg.norm(opb, axis=0)[None, :] + np.abs(c0))
                    assert np.all(np.abs(cf - ref) <= bound + 1e-300), (P, Q, m, n, k, ta, tb)
    # beta = 0 never reads C
    m = n = k = 384
    a, b = cm(rng.uniform(-1, 1, (m, k))), cm(rng.uniform(-1, 1, (k, n)))
    c = np.full((m, n), np.nan, order="F")
    PD.pdgemm("N", "N", m, n, k, 1.0, a, m, b, k, 0.0, c, m)
    assert np.isfinite(c).all()


def test_pzgemm(engine):
    rng = np.random.default_rng(1)
    for P, Q in _grids(engine):
        _set("LLACUDA_GEMM_GRID", P, Q, gemm_nb=128)
        m, n, k = 515, 389, 640
        a = rng.uniform(-1, 1, (m, k)) + 1j * rng.uniform(-1, 1, (m, k))
        b = rng.uniform(-1, 1, (k, n)) + 1j * rng.uniform(-1, 1, (k, n))
        c0 = rng.uniform(-1, 1, (m, n)) + 1j * rng.uniform(-1, 1, (m, n))
        af, bf, cf = (np.asfortranarray(x) for x in (a, b, c0.copy()))
        PD.pzgemm("N", "N", m, n, k, 1 - 2j, af, m, bf, k, 0.5j, cf, m)
        ref = (1 - 2j) * (a @ b) + 0.5j * c0
        assert np.max(np.abs(cf - ref)) <= 64 * np.sqrt(k) * EPS * np.abs(ref).max()


@pytest.mark.parametrize("n,nb", [(97, 128), (1000, 128), (1537, 256), (2111, 512)])
def test_pdgetrf_ipiv_bit_exact_and_solve(engine, n, nb):
    rng = np.random.default_rng(n)
    a = rng.uniform(0, 10, (n, n))
    b = rng.uniform(0, 10, (n, 7))
    ref = O.lu(a, O.netlib())
    ref2 = O.lu(a, O.openblas())
    assert np.array_equal(ref.ipiv_internal, ref2.ipiv_internal)
    one = L.lu(a)                                     # the single-GPU factorisation of the same matrix
    for P, Q in _grids(engine):
        _set("LLACUDA_LU_GRID", P, Q, nb=nb)
        # LLA's lu: transposed copy in (column-major), GETRF, transposed copy out (reference src/linear-algebra.lisp:227-234)
        af = np.asfortranarray(a.copy())
        ipiv = np.zeros(n, dtype=np.int32)
        info = PD.pdgetrf(n, n, af, n, ipiv)
        assert info == 0
        assert np.array_equal(ipiv, ref.ipiv_internal), (P, Q, np.flatnonzero(ipiv != ref.ipiv_internal)[:5])
        assert np.array_equal(ipiv, one.ipiv_internal)
        assert np.max(np.abs(af - ref.lu)) <= 50 * n * EPS * np.abs(ref.lu).max()
        af = np.asfortranarray(a.copy())
        bf = np.asfortranarray(b.copy())
        info = PD.pdgesv(n, 7, af, n, ipiv, bf, n)
        assert info == 0
        assert np.array_equal(ipiv, ref.ipiv_internal)
        res = np.linalg.norm(a @ bf - b) / (n * EPS * np.linalg.norm(a) * np.linalg.norm(bf))
        assert res <= 10, (P, Q, res)


def test_pdgesv_singular_and_ties(engine):
    """lu-singular at size: INFO = first zero pivot, the factorisation completes, B is left alone (DGESV)."""
    rng = np.random.default_rng(5)
    n = 300
    a = rng.integers(-3, 4, (n, n)).astype(np.float64)      # many exact ties
    a[:, 41] = 0.0                                          # exactly singular
    b = rng.uniform(0, 1, (n, 2))
    ref = O.lu(a, O.netlib())
    for P, Q in _grids(engine):
        _set("LLACUDA_LU_GRID", P, Q, nb=128)
        af, bf = np.asfortranarray(a.copy()), np.asfortranarray(b.copy())
        ipiv = np.zeros(n, dtype=np.int32)
        info = PD.pdgesv(n, 2, af, n, ipiv, bf, n)
        assert info > 0
        assert np.array_equal(ipiv, ref.ipiv_internal)
        assert np.array_equal(bf, b)


@pytest.mark.parametrize("n,nb", [(130, 128), (1100, 128), (2304, 512)])
def test_pdpotrf_and_solves(engine, n, nb):
    rng = np.random.default_rng(n + 1)
    g = rng.uniform(0, 1, (n, n))
    a = g @ g.T + n * np.eye(n)
    b = rng.uniform(0, 1, (n, 5))
    ref = O.cholesky(a, O.netlib()).left                   # row-major lower factor = column-major upper factor
    for P, Q in _grids(engine):
        _set("LLACUDA_CHOL_GRID", P, Q, nb=nb)
        # LLA's call site: POTRF('U') on the untransposed row-major buffer (reference src/linear-algebra.lisp:670-674)
        buf = a.copy()
        buf[np.triu_indices(n, 1)] = np.nan                 # the other triangle is never read ...
        info = PD.pdpotrf("U", n, buf, n)
        assert info == 0
        low = np.tril(buf)
        assert np.isnan(buf[np.triu_indices(n, 1)]).all()   # ... nor written
        assert np.max(np.abs(low - ref)) <= 20 * n * EPS * np.abs(ref).max(), (P, Q)
        # solve with the stored factor (POTRS 'U', reference :296-301): B transposed in / out
        bt = np.ascontiguousarray(b.T)
        assert PD.pdpotrs("U", n, 5, low, n, bt, n) == 0
        x = bt.T
        assert np.linalg.norm(a @ x - b) / (n * EPS * np.linalg.norm(a) * np.linalg.norm(x)) <= 10
        # one-shot POSV (reference :303-315)
        buf, bt = a.copy(), np.ascontiguousarray(b.T)
        assert PD.pdposv("U", n, 5, buf, n, bt, n) == 0
        x = bt.T
        assert np.linalg.norm(a @ x - b) / (n * EPS * np.linalg.norm(a) * np.linalg.norm(x)) <= 10


def test_pdpotrf_not_positive_definite(engine):
    rng = np.random.default_rng(3)
    n = 700
    g = rng.uniform(0, 1, (n, n))
    a = g @ g.T + n * np.eye(n)
    a[400, 400] = -1.0
    try:
        O.cholesky(a, O.netlib())
        want = 0
    except O.LapackFailure as e:
        want = e.info
    assert want == 401
    for P, Q in _grids(engine):
        _set("LLACUDA_CHOL_GRID", P, Q, nb=128)
        assert PD.pdpotrf("U", n, a.copy(), n) == want


def test_pdgesv_batched(engine):
    rng = np.random.default_rng(8)
    a = rng.uniform(0, 10, (5001, 32, 32))
    b = rng.uniform(0, 10, (5001, 32))
    x, ipiv, info, lu = PD.pdgesv_batched(a, b, keep_lu=True)
    assert np.all(info == 0)
    be = O.netlib()
    for i in range(0, 5001, 97):
        f = O.lu(a[i], be)
        assert np.array_equal(ipiv[i], f.ipiv_internal)
        assert np.array_equal(lu[i], f.lu)
    xx = x[:, :, None]
    res = np.linalg.norm(a @ xx - b[:, :, None], axis=(1, 2)) / (32 * EPS * np.linalg.norm(a, axis=(1, 2)) * np.linalg.norm(xx, axis=(1, 2)))
    assert res.max() <= 10


def test_fortran_symbols_route_to_all_gpus():
    """LLACUDA_GPUS=N: LLA's unchanged dgesv_ / dpotrf_ / dgemm_ run over every GPU of the box (subprocess: the
    switch is read once per process)."""
    code = f"""
import sys, numpy as np
sys.path.insert(0, {ROOT!r})
from lla_b200 import lla as L, pdist as PD
from oracle import lla_oracle as O
rng = np.random.default_rng(0)
n = 1400
a, b = rng.uniform(0, 10, (n, n)), rng.uniform(0, 10, (n, 3))
f = L.lu(a)
assert PD.info()[0] >= 1 and PD.info()[3] == 1, PD.info()
assert np.array_equal(f.ipiv_internal, O.lu(a, O.netlib()).ipiv_internal)
x = L.solve(a, b)
eps = np.finfo(float).eps
assert np.linalg.norm(a @ x - b) / (n * eps * np.linalg.norm(a) * np.linalg.norm(x)) <= 10
c = L.mm(a, a.T)
assert np.allclose(c, a @ a.T, rtol=1e-12)
s = a @ a.T + n * np.eye(n)
l = L.cholesky(s).left
assert np.max(np.abs(l @ l.T - s)) <= 50 * n * eps * np.abs(s).max()
print("routed ok", PD.info())
"""
    env = dict(os.environ, LLACUDA_GPUS=str(max(2, _ndev())), LLACUDA_DIST_MIN_N="1000", LLACUDA_DIST_NB="256")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]


def test_launcher_mode_local_entry_points(tmp_path):
    """One rank per process under torchrun (how bench.py --gpus N drives the engine): block-cyclic operands
    built per rank, *_local rank functions, results gathered with torch.distributed and checked on rank 0."""
    script = os.path.join(ROOT, "tools", "pdist_local_check.py")
    n = _ndev()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", script],
                       capture_output=True, text=True, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "pdist_local_check ok" in r.stdout
