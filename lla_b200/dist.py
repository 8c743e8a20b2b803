"""2-D block-cyclic multi-GPU layer (one process per GPU, torch.distributed for the plumbing).

The reference has no distributed code at all; this is the net-new layer
BASELINE.json configs[2]/[4] ask for: a P x Q process grid over the GPUs of one NVSwitch box,
ScaLAPACK-style block-cyclic ownership (block (I, J) lives on process (I mod P, J mod Q)), panels
broadcast along grid rows / columns with NCCL over NVLink, local work on the llacuda kernels.

`Grid`, `BlockCyclic` and the SUMMA driver take the local GEMM as a callable so that the index
logic can be exercised on CPU (gloo, world_size 2) with a torch stand-in; the product path passes
`device_gemm`, which calls the DMMA kernel on device pointers and has no fallback.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass
class Grid:
    P: int
    Q: int
    rank: int
    row_group: object = None   # the Q processes of my grid row
    col_group: object = None   # the P processes of my grid column

    @property
    def p(self):
        return self.rank // self.Q

    @property
    def q(self):
        return self.rank % self.Q

    def rank_of(self, p, q):
        return p * self.Q + q


def choose_grid(world: int):
    """Most-square P x Q with P <= Q (8 GPUs -> 2 x 4, the shape SURVEY.md section 8e plans for)."""
    p = int(world ** 0.5)
    while world % p:
        p -= 1
    return p, world // p


def make_grid(P: int | None = None, Q: int | None = None) -> Grid:
    world, rank = dist.get_world_size(), dist.get_rank()
    if P is None:
        P, Q = choose_grid(world)
    assert P * Q == world
    g = Grid(P, Q, rank)
    # every process must create every group, in the same order
    for p in range(P):
        grp = dist.new_group([p * Q + q for q in range(Q)])
        if p == g.p:
            g.row_group = grp
    for q in range(Q):
        grp = dist.new_group([p * Q + q for p in range(P)])
        if q == g.q:
            g.col_group = grp
    return g


class BlockCyclic:
    """Ownership arithmetic of an (m x n) matrix in nb x nb blocks on a P x Q grid."""

    def __init__(self, m, n, nb, grid: Grid):
        if m % (nb * grid.P) or n % (nb * grid.Q):
            raise ValueError("matrix dimensions must be multiples of nb*P and nb*Q")
        self.m, self.n, self.nb, self.grid = m, n, nb, grid
        self.mloc, self.nloc = m // grid.P, n // grid.Q

    def local_block_rows(self):
        return [I for I in range(self.m // self.nb) if I % self.grid.P == self.grid.p]

    def local_block_cols(self):
        return [J for J in range(self.n // self.nb) if J % self.grid.Q == self.grid.q]

    def scatter_from_global(self, full: torch.Tensor) -> torch.Tensor:
        """full: (n, m) tensor = column-major image of the m x n matrix (row j = column j).
        Returns the local column-major image (nloc, mloc)."""
        nb = self.nb
        rows = torch.cat([torch.arange(I * nb, (I + 1) * nb) for I in self.local_block_rows()])
        cols = torch.cat([torch.arange(J * nb, (J + 1) * nb) for J in self.local_block_cols()])
        return full[cols.to(full.device)][:, rows.to(full.device)].contiguous()

    def gather_to_global(self, local: torch.Tensor) -> torch.Tensor:
        """Inverse of scatter_from_global, on every rank (all_gather; test / verification helper)."""
        g = self.grid
        parts = [torch.empty_like(local) for _ in range(g.P * g.Q)]
        dist.all_gather(parts, local)
        full = torch.empty((self.n, self.m), dtype=local.dtype, device=local.device)
        nb = self.nb
        for r, part in enumerate(parts):
            p, q = r // g.Q, r % g.Q
            rows = torch.cat([torch.arange(I * nb, (I + 1) * nb) for I in range(self.m // nb) if I % g.P == p])
            cols = torch.cat([torch.arange(J * nb, (J + 1) * nb) for J in range(self.n // nb) if J % g.Q == q])
            full[cols.to(full.device)[:, None], rows.to(full.device)[None, :]] = part
        return full


def device_gemm(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc):
    """Local C += A B on the DMMA kernel; A, B, C are torch CUDA tensors holding column-major images."""
    from . import device as D
    D.dgemm("N", "N", m, n, k, alpha, A.data_ptr(), lda, B.data_ptr(), ldb, beta, C.data_ptr(), ldc)


def torch_gemm(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc):
    """CPU stand-in with the same signature (tests of the distribution logic only)."""
    a = A.reshape(-1)[: lda * k].view(k, lda)[:, :m].t()       # m x k
    b = B.reshape(-1)[: ldb * n].view(n, ldb)[:, :k].t()       # k x n
    c = C.reshape(-1)[: ldc * n].view(n, ldc)[:, :m]           # (n, m) image of the m x n block
    c.copy_((alpha * (a @ b) + beta * c.t()).t())


def summa(grid: Grid, m, n, k, nb, A_loc, B_loc, C_loc, local_gemm=device_gemm, lookahead=True):
    """C = A B for block-cyclic operands (SUMMA).  *_loc are local column-major images:
    A_loc (k/Q, m/P), B_loc (n/Q, k/P), C_loc (n/Q, m/P).  One k-panel per step: the owning grid
    column broadcasts its slice of A along the grid rows, the owning grid row its slice of B along
    the grid columns (NCCL over NVLink); with `lookahead` the broadcasts of step s+1 are in flight
    while the local DMMA GEMM of step s runs."""
    P, Q = grid.P, grid.Q
    mloc, nloc = m // P, n // Q
    steps = k // nb
    dev, dt = C_loc.device, C_loc.dtype
    abuf = [torch.empty((nb, mloc), dtype=dt, device=dev) for _ in range(2)]   # A panel: mloc x nb, ld = mloc
    bbuf = [torch.empty((nloc, nb), dtype=dt, device=dev) for _ in range(2)]   # B panel: nb x nloc, ld = nb

    def post(s):
        slot = s % 2
        qk, pk = s % Q, s % P
        if grid.q == qk:
            lk = s // Q
            abuf[slot].copy_(A_loc[lk * nb:(lk + 1) * nb, :])
        if grid.p == pk:
            lk = s // P
            bbuf[slot].copy_(B_loc[:, lk * nb:(lk + 1) * nb])
        wa = dist.broadcast(abuf[slot], src=grid.rank_of(grid.p, qk), group=grid.row_group, async_op=True) if Q > 1 else None
        wb = dist.broadcast(bbuf[slot], src=grid.rank_of(pk, grid.q), group=grid.col_group, async_op=True) if P > 1 else None
        return wa, wb

    pending = post(0)
    for s in range(steps):
        nxt = post(s + 1) if (lookahead and s + 1 < steps) else None
        for w in pending:
            if w is not None:
                w.wait()
        slot = s % 2
        local_gemm(mloc, nloc, nb, 1.0, abuf[slot], mloc, bbuf[slot], nb, 0.0 if s == 0 else 1.0, C_loc, mloc)
        if nxt is None and s + 1 < steps:
            nxt = post(s + 1)
        pending = nxt
    return C_loc
