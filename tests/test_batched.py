"""Batched small solves (BASELINE configs[3]).  GPU parity against the oracle's DGESV per system;
CPU (gloo, world_size 2) test of the one-slice-per-GPU sharding."""
import os
import subprocess
import sys

import numpy as np
import pytest

from lla_b200 import batched as BT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPS = np.finfo(np.float64).eps


def test_shard_range_partitions():
    for batch in (0, 1, 7, 200000, 200003):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                lo, hi = BT.shard_range(batch, r, world)
                assert 0 <= lo <= hi <= batch
                cover.extend(range(lo, hi)) if batch < 100 else cover.append((lo, hi))
            if batch < 100:
                assert cover == list(range(batch))
            else:
                assert cover[0][0] == 0 and cover[-1][1] == batch
                assert all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
                sizes = [hi - lo for lo, hi in cover]
                assert max(sizes) - min(sizes) <= 1


def test_sharding_two_ranks_gloo(tmp_path):
    """world_size = 2 on CPU: each rank takes its slice, results are gathered with gloo and must tile
    the batch exactly (host-side logic of the multi-GPU leg; the slices themselves need no collective)."""
    script = tmp_path / "shard.py"
    script.write_text(f"""
import os, sys
sys.path.insert(0, {ROOT!r})
import torch, torch.distributed as dist
from lla_b200.batched import shard_range
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
batch = 1001
lo, hi = shard_range(batch, r, w)
mine = torch.zeros(batch, dtype=torch.int64)
mine[lo:hi] = 1
dist.all_reduce(mine)
assert int(mine.min()) == 1 and int(mine.max()) == 1, "slices must tile the batch exactly once"
t = torch.tensor([float(hi - lo)])
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() - (hi - lo) <= 1
dist.destroy_process_group()
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("n,nrhs,batch", [(32, 1, 3000), (32, 4, 500), (17, 1, 400), (1, 1, 10), (5, 8, 64), (31, 2, 100)])
def test_batched_vs_oracle(n, nrhs, batch):
    from oracle import lla_oracle as O
    rng = np.random.default_rng(n * 100 + nrhs)
    a = rng.uniform(0, 10, (batch, n, n))
    b = rng.uniform(0, 10, (batch, n)) if nrhs == 1 else rng.uniform(0, 10, (batch, n, nrhs))
    x, ipiv, info, lu = BT.solve_batched(a, b, keep_lu=True)
    assert np.all(info == 0)
    be = O.netlib()
    for i in range(0, batch, max(1, batch // 200)):
        f = O.lu(a[i], be)
        assert np.array_equal(ipiv[i], f.ipiv_internal), i
        assert np.max(np.abs(lu[i] - f.lu)) <= 50 * n * EPS * np.abs(f.lu).max()
    xr = np.linalg.solve(a, b[:, :, None] if nrhs == 1 else b)
    xr = xr[:, :, 0] if nrhs == 1 else xr
    bb = b[:, :, None] if nrhs == 1 else b
    xx = x[:, :, None] if nrhs == 1 else x
    res = np.linalg.norm(a @ xx - bb, axis=(1, 2)) / (n * EPS * np.linalg.norm(a, axis=(1, 2)) * np.linalg.norm(xx, axis=(1, 2)))
    assert res.max() <= 10
    assert np.allclose(x, xr, rtol=1e-7, atol=1e-9)


@pytest.mark.gpu
def test_batched_singular_and_ties():
    from oracle import lla_oracle as O
    rng = np.random.default_rng(1)
    a = rng.integers(-2, 3, (300, 32, 32)).astype(np.float64)     # many exact ties, some singular
    a[7] = 0.5                                                     # lu-singular analogue: INFO = 2
    b = rng.uniform(0, 1, (300, 32))
    x, ipiv, info = BT.solve_batched(a, b)
    be = O.netlib()
    for i in range(300):
        f = O.lu(a[i], be)
        assert np.array_equal(ipiv[i], f.ipiv_internal), i
        try:
            O.solve(a[i], b[i], be)
            want = 0
        except O.LapackFailure as e:
            want = e.info
        assert info[i] == want, (i, info[i], want)
        if want != 0:
            assert np.array_equal(x[i], b[i])                      # DGESV leaves B alone when singular
    assert info[7] == 2
