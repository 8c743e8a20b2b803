"""GPU leg of the block-cyclic layer: the product path (DeviceOps = llacuda kernels on device pointers, two CUDA
streams with look-ahead) on one GPU, grid 1 x 1, against numpy.  The multi-rank logic is covered on CPU by
tests/test_dist_cpu.py (gloo) and measured on 2-8 GPUs by tools/dist_chol_bench.py / tools/dist_lu_bench.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def _setup():
    import torch
    from lla_b200 import _lib, device as D, dist as LD
    _lib.init(0)
    D.use_torch_stream()
    return torch, D, LD


@pytest.mark.parametrize("lookahead", [True, False])
def test_pdpotrf_pdpotrs_device_ops(lookahead):
    torch, D, LD = _setup()
    n, nb = 1536, 256
    rng = np.random.default_rng(5)
    g = rng.uniform(0, 1, (n, n))
    spd = g @ g.T + n * np.eye(n)
    img = np.ascontiguousarray(np.triu(spd).T)          # column-major image; the other triangle holds a sentinel
    img[np.triu_indices(n, 1)] = -777.0
    A = torch.from_numpy(img).cuda()
    grid = LD.Grid(1, 1, 0)
    info = LD.pdpotrf(grid, n, nb, A, lookahead=lookahead)
    assert info == 0
    out = A.cpu().numpy()
    u = np.triu(out.T)
    assert np.all(out[np.triu_indices(n, 1)] == -777.0), "only the upper triangle may be written"
    assert np.linalg.norm(u.T @ u - spd) / (n * EPS * np.linalg.norm(spd)) <= 10
    b = rng.uniform(0, 1, (n, 7))
    B = torch.from_numpy(np.ascontiguousarray(b.T)).cuda()
    LD.pdpotrs(grid, n, nb, A, B)
    x = B.cpu().numpy().T
    assert np.linalg.norm(spd @ x - b) / (n * EPS * np.linalg.norm(spd) * np.linalg.norm(x)) <= 10
    # not positive definite: LAPACK's INFO
    bad = spd.copy(); bad[700, 700] = -1.0
    A = torch.from_numpy(np.ascontiguousarray(np.triu(bad).T)).cuda()
    assert LD.pdpotrf(grid, n, nb, A, lookahead=lookahead) == 701


@pytest.mark.parametrize("lookahead", [True, False])
def test_pdgetrf_pdgetrs_device_ops_ipiv_bit_exact(lookahead):
    torch, D, LD = _setup()
    from lla_b200 import lla as L
    n, nb = 1536, 256
    rng = np.random.default_rng(6)
    a = rng.uniform(0, 10, (n, n))
    A = torch.from_numpy(np.ascontiguousarray(a.T)).cuda()
    ipiv = torch.zeros(n, dtype=torch.int32, device="cuda")
    grid = LD.Grid(1, 1, 0)
    info = LD.pdgetrf(grid, n, nb, A, ipiv, lookahead=lookahead)
    assert info == 0
    ref = L.lu(a)                                        # the single-GPU DGETRF through the Fortran drop-in
    assert np.array_equal(ipiv.cpu().numpy(), ref.ipiv_internal), "ipiv must not depend on the blocking"
    lu = A.cpu().numpy().T
    assert np.max(np.abs(lu - ref.lu)) <= 50 * n * EPS * np.abs(ref.lu).max()
    b = rng.uniform(0, 10, (n, 5))
    B = torch.from_numpy(np.ascontiguousarray(b.T)).cuda()
    LD.pdgetrs(grid, n, nb, A, ipiv, B)
    x = B.cpu().numpy().T
    assert np.linalg.norm(a @ x - b) / (n * EPS * np.linalg.norm(a) * np.linalg.norm(x)) <= 10


def test_dtrsm_and_dlaswp_device_entry_points():
    torch, D, LD = _setup()
    rng = np.random.default_rng(8)
    m, n = 300, 37
    t = np.triu(rng.uniform(-1, 1, (m, m))) + 4 * np.eye(m)
    b = rng.uniform(-1, 1, (m, n))
    for uplo, trans, diag in (("U", "N", "N"), ("U", "T", "N"), ("L", "N", "U"), ("L", "T", "N")):
        tm = t if uplo == "U" else t.T.copy()
        eff = tm.copy()
        if diag == "U":
            np.fill_diagonal(eff, 1.0)
        T = torch.from_numpy(np.ascontiguousarray(tm.T)).cuda()
        B = torch.from_numpy(np.ascontiguousarray(b.T)).cuda()
        D.dtrsm(uplo, trans, diag, m, n, T.data_ptr(), m, B.data_ptr(), m)
        x = B.cpu().numpy().T
        op = eff.T if trans == "T" else eff
        assert np.linalg.norm(op @ x - b) / (m * EPS * np.linalg.norm(op) * np.linalg.norm(x)) <= 10
    # DLASWP, forward and back: 200 interchanges (the dense-block kernel) and 40 (the chunked kernel)
    for k in (200, 40):
        a = rng.uniform(-1, 1, (m, n))
        piv = np.array([rng.integers(i, m) + 1 for i in range(k)], dtype=np.int32)
        want = a.copy()
        for i in range(k):
            p = piv[i] - 1
            want[[i, p]] = want[[p, i]]
        A = torch.from_numpy(np.ascontiguousarray(a.T)).cuda()
        P = torch.from_numpy(piv).cuda()
        D.dlaswp(n, A.data_ptr(), m, 0, k, P.data_ptr(), True)
        assert np.array_equal(A.cpu().numpy().T, want)
        D.dlaswp(n, A.data_ptr(), m, 0, k, P.data_ptr(), False)
        assert np.array_equal(A.cpu().numpy().T, a)
