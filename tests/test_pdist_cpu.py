"""CPU tests of the multi-GPU engine's host logic: ownership arithmetic of the block-cyclic layout (the C functions
against the Python mirror and against brute force), that scatter parts tile the matrix exactly once for every grid,
and that the entry points fail loudly without a device (no CPU fallback, no silent zero result)."""
import ctypes
import itertools

import numpy as np
import pytest

from lla_b200 import _lib, pdist as PD


@pytest.fixture(scope="module")
def lib():
    return _lib.lib()


def test_numroc_padded_l2g_match_brute_force(lib):
    lib.llacuda_dist_numroc.restype = ctypes.c_int64
    lib.llacuda_dist_padded.restype = ctypes.c_int64
    lib.llacuda_dist_l2g.restype = ctypes.c_int64
    i64 = ctypes.c_int64
    for n, nb, nprocs in itertools.product((0, 1, 127, 128, 1000, 4096, 5001), (1, 32, 128, 512), (1, 2, 3, 4, 8)):
        owner = (np.arange(n) // nb) % nprocs
        total = 0
        for ip in range(nprocs):
            want = int((owner == ip).sum())
            assert lib.llacuda_dist_numroc(i64(n), i64(nb), ip, nprocs) == want == PD.numroc(n, nb, ip, nprocs)
            total += want
            mine = np.flatnonzero(owner == ip)
            for l in (0, len(mine) // 2, len(mine) - 1):
                if len(mine):
                    assert lib.llacuda_dist_l2g(i64(l), i64(nb), ip, nprocs) == mine[l]
        assert total == n
    for n, nb, P, Q in itertools.product((1, 511, 512, 16384, 20000), (128, 512), (1, 2, 3), (1, 2, 4)):
        np_ = lib.llacuda_dist_padded(i64(n), i64(nb), P, Q)
        assert np_ == PD.padded(n, nb, P, Q) and np_ >= n and np_ % (nb * P) == 0 and np_ % (nb * Q) == 0
        assert np_ - n < nb * PD.lcm(P, Q)


def test_default_grid(lib):
    for w, want in ((1, (1, 1)), (2, (1, 2)), (4, (2, 2)), (6, (2, 3)), (8, (2, 4)), (7, (1, 7))):
        p, q = ctypes.c_int(0), ctypes.c_int(0)
        lib.llacuda_dist_default_grid(w, ctypes.byref(p), ctypes.byref(q))
        assert (p.value, q.value) == want == PD.default_grid(w)


@pytest.mark.parametrize("P,Q", [(1, 1), (1, 2), (2, 1), (2, 2), (2, 4), (1, 8), (4, 2)])
def test_scatter_parts_tile_the_matrix(P, Q):
    nb = 4
    n = nb * int(PD.lcm(P, Q)) * 2
    img = np.arange(n * n, dtype=np.float64).reshape(n, n)
    seen = np.zeros((n, n), dtype=int)
    for r in range(P * Q):
        bc = PD.BlockCyclic(n, n, nb, P, Q, r)
        part = bc.scatter(img)
        assert part.shape == (bc.nloc, bc.mloc) == (n // Q, n // P)
        rows, cols = bc.local_rows(), bc.local_cols()
        assert np.array_equal(part, img[np.ix_(cols, rows)])
        seen[np.ix_(cols, rows)] += 1
        # block (li, lj) of the part is the global block (li P + p, lj Q + q)
        p, q = r // Q, r % Q
        for li, lj in ((0, 0), (bc.mloc // nb - 1, bc.nloc // nb - 1)):
            I, J = li * P + p, lj * Q + q
            assert np.array_equal(part[lj * nb:(lj + 1) * nb, li * nb:(li + 1) * nb], img[J * nb:(J + 1) * nb, I * nb:(I + 1) * nb])
    assert np.all(seen == 1)


def test_engine_entry_points_fail_loudly_without_engine_or_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("no-device behaviour")
    assert lib.llacuda_dist_init_all(0, 0, 0) != 0          # no device (or no NCCL): an error code, never a fallback
    n = 8
    a = np.asfortranarray(np.eye(n))
    b = np.asfortranarray(np.ones((n, 1)))
    ipiv = np.zeros(n, dtype=np.int32)
    assert PD.pdgesv(n, 1, a.copy(order="F"), n, ipiv, b, n) <= -10000
    assert PD.pdpotrf("U", n, a.copy(order="F"), n) <= -10000
    c = np.zeros((n, n), order="F")
    with pytest.raises(_lib.LlacudaError):
        PD.pdgemm("N", "N", n, n, n, 1.0, a, n, a, n, 0.0, c, n)
    assert np.isnan(c).all()                                 # poisoned, never a silent all-zero product
    # argument checks come first, in LAPACK order
    assert PD.pdgesv(-1, 1, a, n, ipiv, b, n) == -1
    assert PD.pdpotrf("X", n, a, n) == -1
