"""Batched small solves and their sharding across GPUs (BASELINE.json configs[3]).

`solve_batched` is the batched form of `(lla:solve a b)` (reference src/linear-algebra.lisp:278-289)
for a stack of row-major systems; `shard_range` is the contiguous one-slice-per-GPU partition
(independent units, no data-path collective).
"""
from __future__ import annotations

import numpy as np


def shard_range(batch: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of the batch owned by `rank`; slices differ by at most one system."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def solve_batched(a, b, device=None, keep_lu=False):
    """a: (batch, n, n) row-major systems, b: (batch, n) or (batch, n, nrhs).  Returns (x, ipiv, info)
    with x shaped like b, ipiv the 1-based LAPACK swap sequences, info the per-system DGESV INFO."""
    import torch
    from . import _lib, device as D
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    batch, n, n2 = a.shape
    assert n == n2 and b.shape[0] == batch and b.shape[1] == n
    vec = b.ndim == 2
    nrhs = 1 if vec else b.shape[2]
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    _lib.init(dev.index)
    D.use_torch_stream()
    # column-major images, as LLA's transposing copy produces them (src/foreign-memory.lisp:211-230)
    A = torch.from_numpy(np.ascontiguousarray(np.swapaxes(a, 1, 2))).to(dev)
    bb = b[:, :, None] if vec else b
    B = torch.from_numpy(np.ascontiguousarray(np.swapaxes(bb, 1, 2))).to(dev)
    ipiv = torch.zeros((batch, n), dtype=torch.int32, device=dev)
    info = torch.zeros(batch, dtype=torch.int32, device=dev)
    D.dgesv_batched(n, nrhs, batch, A.data_ptr(), n * n, ipiv.data_ptr(), B.data_ptr(), n * nrhs, info.data_ptr(),
                    1 if keep_lu else 0)
    x = np.swapaxes(B.cpu().numpy(), 1, 2)
    x = np.ascontiguousarray(x[:, :, 0] if vec else x)
    out = (x, ipiv.cpu().numpy(), info.cpu().numpy())
    if keep_lu:
        out = out + (np.ascontiguousarray(np.swapaxes(A.cpu().numpy(), 1, 2)),)
    return out
