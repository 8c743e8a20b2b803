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


def device_zgemm(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc):
    """Local complex-double C = alpha A B + beta C on the DMMA ZGEMM kernel (torch complex128 CUDA tensors)."""
    from . import device as D
    D.zgemm("N", "N", m, n, k, complex(alpha), A.data_ptr(), lda, B.data_ptr(), ldb, complex(beta), C.data_ptr(), ldc)


def torch_gemm(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc):
    """CPU stand-in with the same signature (tests of the distribution logic only)."""
    a = A.reshape(-1)[: lda * k].view(k, lda)[:, :m].t()       # m x k
    b = B.reshape(-1)[: ldb * n].view(n, ldb)[:, :k].t()       # k x n
    c = C.reshape(-1)[: ldc * n].view(n, ldc)[:, :m]           # (n, m) image of the m x n block
    c.copy_((alpha * (a @ b) + beta * c.t()).t())


def summa(grid: Grid, m, n, k, nb, A_loc, B_loc, C_loc, local_gemm=None, lookahead=True):
    """C = A B for block-cyclic operands (SUMMA), double or complex double (the local kernel follows the dtype).  *_loc are local column-major images:
    A_loc (k/Q, m/P), B_loc (n/Q, k/P), C_loc (n/Q, m/P).  One k-panel per step: the owning grid
    column broadcasts its slice of A along the grid rows, the owning grid row its slice of B along
    the grid columns (NCCL over NVLink); with `lookahead` the broadcasts of step s+1 are in flight
    while the local DMMA GEMM of step s runs."""
    P, Q = grid.P, grid.Q
    mloc, nloc = m // P, n // Q
    steps = k // nb
    dev, dt = C_loc.device, C_loc.dtype
    if local_gemm is None:
        local_gemm = device_zgemm if dt.is_complex else device_gemm
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


# --------------------------------------------------------------------------------------------------
# Block-cyclic Cholesky (BASELINE.json configs[2]).  The reference has one call
# site, POTRF('U') on the untransposed row-major buffer (src/linear-algebra.lisp:670-674) followed by
# POTRS('U') (:296-301); the distributed form keeps that convention: A = U^T U, only the upper
# triangle of the column-major view is read or written.
#
# Local work goes through an `ops` object so that the index / communication logic runs on CPU (gloo)
# with `TorchOps`; the product path is `DeviceOps` (llacuda kernels on device pointers, no fallback).
# Every operand is an *image view*: a torch view of shape (cols, rows) whose element [c, r] is the
# matrix element (r, c) and whose stride(0) is the leading dimension (column-major storage).

class DeviceOps:
    """llacuda kernels on device pointers: DPOTRF leaf/blocked driver, inverse-based DTRSM, DMMA GEMM."""

    def __init__(self):
        from . import device as D
        self.D = D
        self._info = None

    def _ld(self, img):
        assert img.stride(1) == 1
        return img.stride(0) if img.shape[0] > 1 else max(img.shape[1], 1)

    def potrf_upper(self, img):
        n = img.shape[0]
        if self._info is None:
            self._info = torch.zeros(1, dtype=torch.int32, device=img.device)
        self.D.dpotrf("U", n, img.data_ptr(), self._ld(img), self._info.data_ptr())
        return self._info   # device scalar; the caller decides when to read it

    def trsm(self, uplo, trans, timg, bimg, diag="N"):
        m, n = bimg.shape[1], bimg.shape[0]
        self.D.dtrsm(uplo, trans, diag, m, n, timg.data_ptr(), self._ld(timg), bimg.data_ptr(), self._ld(bimg))

    def getrf(self, img, ipiv):
        """LU with partial pivoting of the (rows x cols) panel behind the (cols, rows) image; ipiv: int32 device
        tensor (1-based, relative to the panel's first row).  Returns the device INFO scalar."""
        if self._info is None:
            self._info = torch.zeros(1, dtype=torch.int32, device=img.device)
        self.D.dgetrf(img.shape[1], img.shape[0], img.data_ptr(), self._ld(img), ipiv.data_ptr(), self._info.data_ptr())
        return self._info

    def laswp(self, img, ipiv, forward=True):
        """rows i <-> ipiv[i]-1, i in [0, len(ipiv)), on all columns of the image."""
        if img.shape[0] > 0:
            self.D.dlaswp(img.shape[0], img.data_ptr(), self._ld(img), 0, ipiv.numel(), ipiv.data_ptr(), forward)

    def gemm(self, opa, m, n, k, alpha, aimg, bimg, beta, cimg, tri=0):
        """C (m x n) = alpha op(A) B + beta C, op = 'T' or 'N'; tri: 0 full, 1 upper, 2 lower."""
        self.D.dgemm(opa, "N", m, n, k, alpha, aimg.data_ptr(), self._ld(aimg), bimg.data_ptr(), self._ld(bimg),
                     beta, cimg.data_ptr(), self._ld(cimg), tri)


class TorchOps:
    """CPU stand-in with the same interface (tests of the distribution logic only)."""

    def potrf_upper(self, img):
        a = img.t()                                    # matrix view
        full = torch.triu(a) + torch.triu(a, 1).t()
        l, info = torch.linalg.cholesky_ex(full)
        a.copy_(torch.triu(l.t()) + torch.tril(a, -1))  # only the upper triangle is written
        return info.to(torch.int32).reshape(1)

    def getrf(self, img, ipiv):
        a = img.t()
        lu, piv, info = torch.linalg.lu_factor_ex(a.contiguous())
        a.copy_(lu)
        ipiv.copy_(piv[:ipiv.numel()].to(torch.int32))
        return info.to(torch.int32).reshape(1)

    def laswp(self, img, ipiv, forward=True):
        a = img.t()
        order = range(ipiv.numel()) if forward else range(ipiv.numel() - 1, -1, -1)
        for i in order:
            pi = int(ipiv[i]) - 1
            if pi != i:
                tmp = a[i].clone(); a[i] = a[pi]; a[pi] = tmp

    def trsm(self, uplo, trans, timg, bimg, diag="N"):
        t = timg.t()
        t = torch.triu(t) if uplo == "U" else torch.tril(t)
        if diag == "U":
            t = t - torch.diag(torch.diagonal(t)) + torch.eye(t.shape[0], dtype=t.dtype)
        if trans == "T":
            x = torch.linalg.solve_triangular(t.t(), bimg.t(), upper=(uplo != "U"))
        else:
            x = torch.linalg.solve_triangular(t, bimg.t(), upper=(uplo == "U"))
        bimg.copy_(x.t())

    def gemm(self, opa, m, n, k, alpha, aimg, bimg, beta, cimg, tri=0):
        a = aimg.t() if opa == "N" else aimg          # op(A): m x k
        a = a[:m, :k]
        b = bimg.t()[:k, :n]
        c = cimg.t()
        upd = alpha * (a @ b) + beta * c[:m, :n]
        if tri == 1:
            upd = torch.triu(upd) + torch.tril(c[:m, :n], -1)
        elif tri == 2:
            upd = torch.tril(upd) + torch.triu(c[:m, :n], 1)
        c[:m, :n] = upd


def _count_le(k, r, R):
    """Number of block indices I <= k with I mod R == r."""
    return (k - r) // R + 1 if k >= r else 0


def pdpotrf(grid: Grid, n, nb, A_loc, ops=None, lookahead=True):
    """In-place block-cyclic Cholesky A = U^T U of the SPD matrix whose upper triangle is stored in
    A_loc ((n/Q, n/P) local column-major image, block (I, J) on process (I mod P, J mod Q)).
    Right-looking; the panel of block step k is
      1. the owner of the diagonal block factors it (DPOTRF 'U') and broadcasts U_kk along its grid row;
      2. the processes of that grid row solve U_kk^T X = A(k, J > k) on their blocks of block row k;
      3. every owner broadcasts its slice of the finished block row to all processes (NCCL over NVLink:
         nb x (n - k nb) doubles per step, far below the update's work);
    and the update of step k is one DMMA GEMM per local block column over the blocks (I, J), k < I <= J
    (the triangular staircase of the block-cyclic layout; diagonal blocks through the masked GEMM).
    With `lookahead` the update is split: block row k+1 first, then panel k+1 runs on a second stream (its
    kernels and its broadcasts) while the rest of update k keeps the GPU busy.
    Returns LAPACK's INFO (0, or the index of the first leading minor that is not positive definite),
    identical on every rank."""
    ops = ops or DeviceOps()
    P, Q, p, q = grid.P, grid.Q, grid.p, grid.q
    mloc, nloc = n // P, n // Q
    assert A_loc.shape == (nloc, mloc) and n % (nb * P) == 0 and n % (nb * Q) == 0
    nblk = n // nb
    dev, dt = A_loc.device, A_loc.dtype
    info_acc = torch.zeros(1, dtype=torch.int32, device=dev)
    multi = P * Q > 1
    cuda = dev.type == "cuda"
    # persistent double buffers: the two streams hand them over through events, nothing is freed in the loop
    pbuf = [[torch.empty((nloc, nb), dtype=dt, device=dev) for _ in range(Q)] for _ in range(2)]
    wrbuf = [torch.empty((mloc, nb), dtype=dt, device=dev) for _ in range(2)]
    dblk = torch.empty((nb, nb), dtype=dt, device=dev)
    main = torch.cuda.current_stream() if cuda else None
    side = torch.cuda.Stream(priority=-1) if (cuda and lookahead) else main
    ev_panel = [torch.cuda.Event() for _ in range(2)] if cuda else None
    ev_rowdone = torch.cuda.Event() if cuda else None

    class _on:  # run a section on `stream`: torch's current stream and llacuda's stream move together
        def __init__(self, stream): self.stream = stream
        def __enter__(self):
            if cuda:
                self.ctx = torch.cuda.stream(self.stream); self.ctx.__enter__()
                if hasattr(ops, "D"): ops.D.use_torch_stream()
        def __exit__(self, *a):
            if cuda:
                self.ctx.__exit__(*a)
                if hasattr(ops, "D"): ops.D.use_torch_stream()

    def panel(k):
        """Steps 1-3 for block step k; fills pbuf[k % 2] and returns (pieces, first)."""
        nonlocal info_acc
        pk, qk = k % P, k % Q
        lkr = k // P
        if p == pk:
            if q == qk:
                lkc = k // Q
                diag = A_loc[lkc * nb:(lkc + 1) * nb, lkr * nb:(lkr + 1) * nb]
                inf = ops.potrf_upper(diag)
                info_acc += torch.where((info_acc == 0) & (inf != 0), inf + k * nb, torch.zeros_like(inf))
                dblk.copy_(diag)
            if Q > 1:
                dist.broadcast(dblk, src=grid.rank_of(pk, qk), group=grid.row_group)
            lj0 = _count_le(k, q, Q)
            if lj0 * nb < nloc:
                ops.trsm("U", "T", dblk, A_loc[lj0 * nb:, lkr * nb:(lkr + 1) * nb])
        pieces, first = [], []
        for qq in range(Q):
            lj0q = _count_le(k, qq, Q)
            ncq = nloc - lj0q * nb
            first.append(lj0q)
            if ncq <= 0:
                pieces.append(None)
                continue
            buf = pbuf[k % 2][qq][:ncq]
            if p == pk and q == qq:
                buf.copy_(A_loc[lj0q * nb:, lkr * nb:(lkr + 1) * nb])
            if multi:
                dist.broadcast(buf, src=grid.rank_of(pk, qq))
            pieces.append(buf)
        return pieces, first

    def gather_rows(k, pieces, first):
        """U(k, I) for my block rows I = li P + p > k, stacked in local row order."""
        li0 = _count_le(k, p, P)
        nrows = mloc - li0 * nb
        if nrows <= 0:
            return None, li0
        wr = wrbuf[k % 2][:nrows]
        for li in range(li0, mloc // nb):
            I = li * P + p
            idx = I // Q - first[I % Q]
            wr[(li - li0) * nb:(li - li0 + 1) * nb].copy_(pieces[I % Q][idx * nb:(idx + 1) * nb])
        return wr, li0

    def update(k, wr, li0, wc, lj0, lo, hi):
        """A(I, J) -= U(k, I)^T U(k, J) for my blocks with local block row in [lo, hi), k < I <= J."""
        for lj in range(lj0, nloc // nb):
            J = lj * Q + q
            nfull = _count_le(J - 1, p, P)                  # local block rows with I < J are [0, nfull)
            r0, r1 = max(lo, li0), min(hi, nfull)
            bj = wc[(lj - lj0) * nb:(lj - lj0 + 1) * nb]
            if r1 > r0:
                ops.gemm("T", (r1 - r0) * nb, nb, nb, -1.0, wr[(r0 - li0) * nb:(r1 - li0) * nb], bj, 1.0,
                         A_loc[lj * nb:(lj + 1) * nb, r0 * nb:r1 * nb])
            if J % P == p and lo <= J // P < hi:            # the diagonal block (J, J) is mine
                li = J // P
                ops.gemm("T", nb, nb, nb, -1.0, wr[(li - li0) * nb:(li - li0 + 1) * nb], bj, 1.0,
                         A_loc[lj * nb:(lj + 1) * nb, li * nb:(li + 1) * nb], tri=1)

    with _on(side):
        nxt = panel(0)
        if cuda: ev_panel[0].record(side)
    for k in range(nblk):
        pieces, first = nxt
        if cuda and side is not main: main.wait_event(ev_panel[k % 2])
        wc = pieces[q]
        wr, li0 = gather_rows(k, pieces, first)
        lj0 = first[q]
        have = wc is not None and wr is not None
        # block row k+1 first: it is all that panel k+1 needs
        own_next = (k + 1 < nblk) and ((k + 1) % P == p)
        split = li0 + 1 if own_next else li0
        if have and own_next:
            update(k, wr, li0, wc, lj0, li0, split)
        if k + 1 < nblk:
            if cuda and side is not main:
                ev_rowdone.record(main); side.wait_event(ev_rowdone)
            with _on(side):
                nxt = panel(k + 1)
                if cuda: ev_panel[(k + 1) % 2].record(side)
        if have:
            update(k, wr, li0, wc, lj0, split, mloc // nb)
    if cuda and side is not main: main.wait_stream(side)
    if multi:
        # INFO of the first failing diagonal block wins (smallest positive)
        big = torch.where(info_acc > 0, info_acc, torch.full_like(info_acc, 2 ** 31 - 1))
        dist.all_reduce(big, op=dist.ReduceOp.MIN)
        info_acc = torch.where(big == 2 ** 31 - 1, torch.zeros_like(big), big)
    return int(info_acc.item())


def pdpotrs(grid: Grid, n, nb, U_loc, B, ops=None):
    """Solves A X = B in place with the block-cyclic factor of pdpotrf (A = U^T U); B is replicated on
    every rank as the (nrhs, n) column-major image of the n x nrhs right-hand sides (SURVEY.md section 8e:
    few right-hand sides do not shard; the chain is one small reduction + one broadcast per block)."""
    ops = ops or DeviceOps()
    P, Q, p, q = grid.P, grid.Q, grid.p, grid.q
    mloc, nloc = n // P, n // Q
    nblk = n // nb
    nrhs = B.shape[0]
    dev, dt = B.device, B.dtype
    multi = P * Q > 1
    rows_of = lambda cnt: torch.cat([torch.arange((li * P + p) * nb, (li * P + p + 1) * nb) for li in range(cnt)]).to(dev) \
        if cnt > 0 else None
    # ---- U^T y = b, block forward substitution down the block columns
    for k in range(nblk):
        pk, qk = k % P, k % Q
        s = torch.zeros((nrhs, nb), dtype=dt, device=dev)
        if q == qk:
            cnt = _count_le(k - 1, p, P)                   # my block rows I < k of block column k
            if cnt > 0:
                lkc = k // Q
                yr = B[:, rows_of(cnt)].contiguous()       # (nrhs, cnt nb) image of the known part of y
                ops.gemm("T", nb, nrhs, cnt * nb, 1.0, U_loc[lkc * nb:(lkc + 1) * nb, :cnt * nb], yr, 0.0, s)
        if multi:
            dist.all_reduce(s)
        yk = (B[:, k * nb:(k + 1) * nb] - s).contiguous()
        if p == pk and q == qk:
            ops.trsm("U", "T", U_loc[(k // Q) * nb:(k // Q + 1) * nb, (k // P) * nb:(k // P + 1) * nb], yk)
        if multi:
            dist.broadcast(yk, src=grid.rank_of(pk, qk))
        B[:, k * nb:(k + 1) * nb] = yk
    # ---- U x = y, block backward substitution along the block rows
    for k in range(nblk - 1, -1, -1):
        pk, qk = k % P, k % Q
        s = torch.zeros((nrhs, nb), dtype=dt, device=dev)
        if p == pk:
            lj0 = _count_le(k, q, Q)                       # my block columns J > k of block row k
            if lj0 * nb < nloc:
                cols = torch.cat([torch.arange((lj * Q + q) * nb, (lj * Q + q + 1) * nb) for lj in range(lj0, nloc // nb)]).to(dev)
                xr = B[:, cols].contiguous()
                lkr = k // P
                ops.gemm("N", nb, nrhs, cols.numel(), 1.0, U_loc[lj0 * nb:, lkr * nb:(lkr + 1) * nb], xr, 0.0, s)
        if multi:
            dist.all_reduce(s)
        xk = (B[:, k * nb:(k + 1) * nb] - s).contiguous()
        if p == pk and q == qk:
            ops.trsm("U", "N", U_loc[(k // Q) * nb:(k // Q + 1) * nb, (k // P) * nb:(k // P + 1) * nb], xk)
        if multi:
            dist.broadcast(xk, src=grid.rank_of(pk, qk))
        B[:, k * nb:(k + 1) * nb] = xk
    return B


# --------------------------------------------------------------------------------------------------
# Block-cyclic LU with partial pivoting on a 1 x Q grid: block column J lives on
# process J mod Q with all of its rows, so a panel's pivot search never leaves the GPU that owns it --
# the panel is factored by the very kernel the single-GPU DGETRF uses, on the same numbers in the same
# order, which is what keeps ipiv bit-exact for every GPU count.  Per block step k:
#   1. the owner factors its m_k x nb panel (DGETRF) and broadcasts the panel and its pivots (NCCL over
#      NVLink: m_k x nb doubles per step);
#   2. every process applies the interchanges to its own block columns (left and right of the panel),
#      solves U_kJ = L_kk^-1 A_kJ and updates A_IJ -= L_Ik U_kJ with the DMMA GEMM;
# with `lookahead` the owner of panel k+1 updates that block column first and factors it on a second
# stream while everybody's remaining update of step k is still running.

def _one_row_grid(grid):
    if grid.P != 1:
        raise ValueError("pdgetrf distributes block columns: use a 1 x Q grid (make_grid(1, world))")


def pdgetrf(grid: Grid, n, nb, A_loc, ipiv, ops=None, lookahead=True):
    """In-place LU of the n x n matrix held as local block columns A_loc ((n/Q, n) column-major image);
    ipiv: int32 tensor of length n on every rank (LAPACK's 1-based swap sequence on return).
    Returns LAPACK's INFO (0 or the index of the first exactly-zero pivot; the factorisation completes)."""
    _one_row_grid(grid)
    ops = ops or DeviceOps()
    Q, q = grid.Q, grid.q
    nloc = n // Q
    assert A_loc.shape == (nloc, n) and n % (nb * Q) == 0
    nblk = n // nb
    dev, dt = A_loc.device, A_loc.dtype
    cuda = dev.type == "cuda"
    multi = Q > 1
    info_acc = torch.zeros(1, dtype=torch.int32, device=dev)
    lbuf = [torch.empty(nb * n, dtype=dt, device=dev) for _ in range(2)]         # panel k: contiguous (nb, n - k nb) image
    pbuf = [torch.empty(nb, dtype=torch.int32, device=dev) for _ in range(2)]
    main = torch.cuda.current_stream() if cuda else None
    side = torch.cuda.Stream(priority=-1) if (cuda and lookahead) else main
    ev_panel = [torch.cuda.Event() for _ in range(2)] if cuda else None
    ev_col = torch.cuda.Event() if cuda else None

    class _on:
        def __init__(self, stream): self.stream = stream
        def __enter__(self):
            if cuda:
                self.ctx = torch.cuda.stream(self.stream); self.ctx.__enter__()
                if hasattr(ops, "D"): ops.D.use_torch_stream()
        def __exit__(self, *a):
            if cuda:
                self.ctx.__exit__(*a)
                if hasattr(ops, "D"): ops.D.use_torch_stream()

    def panel(k):
        nonlocal info_acc
        r0 = k * nb
        pan = lbuf[k % 2][:nb * (n - r0)].view(nb, n - r0)
        piv = pbuf[k % 2]
        if q == k % Q:
            lk = k // Q
            mine = A_loc[lk * nb:(lk + 1) * nb, r0:]
            inf = ops.getrf(mine, piv)
            info_acc += torch.where((info_acc == 0) & (inf != 0), inf + r0, torch.zeros_like(inf))
            pan.copy_(mine)
        if multi:
            dist.broadcast(pan, src=k % Q)
            dist.broadcast(piv, src=k % Q)
        return pan, piv

    def update(k, pan, piv, lj_lo, lj_hi):
        """interchanges + U_kJ + trailing update on my local block columns [lj_lo, lj_hi) (all right of panel k)."""
        if lj_hi <= lj_lo:
            return
        r0 = k * nb
        cols = A_loc[lj_lo * nb:lj_hi * nb, r0:]                  # image of the rows >= r0 of those columns
        ops.laswp(cols, piv, True)
        ops.trsm("L", "N", pan[:, :nb], cols[:, :nb], diag="U")
        m2 = n - r0 - nb
        if m2 > 0:
            ops.gemm("N", m2, cols.shape[0], nb, -1.0, pan[:, nb:], cols[:, :nb], 1.0, cols[:, nb:])

    with _on(side):
        nxt = panel(0)
        if cuda: ev_panel[0].record(side)
    for k in range(nblk):
        pan, piv = nxt
        if cuda and side is not main: main.wait_event(ev_panel[k % 2])
        ipiv[k * nb:(k + 1) * nb] = piv + k * nb
        lj0 = _count_le(k, q, Q)                                  # my first block column right of panel k
        # interchanges of my block columns left of the panel (nothing depends on them until the end)
        nleft = _count_le(k - 1, q, Q)
        if nleft > 0:
            ops.laswp(A_loc[:nleft * nb, k * nb:], piv, True)
        own_next = (k + 1 < nblk) and ((k + 1) % Q == q)
        split = lj0 + 1 if own_next else lj0
        if own_next:
            update(k, pan, piv, lj0, split)                        # block column k+1 first
        if k + 1 < nblk:
            if cuda and side is not main:
                ev_col.record(main); side.wait_event(ev_col)
            with _on(side):
                nxt = panel(k + 1)
                if cuda: ev_panel[(k + 1) % 2].record(side)
        update(k, pan, piv, split, nloc // nb)
    if cuda and side is not main: main.wait_stream(side)
    if multi:
        big = torch.where(info_acc > 0, info_acc, torch.full_like(info_acc, 2 ** 31 - 1))
        dist.all_reduce(big, op=dist.ReduceOp.MIN)
        info_acc = torch.where(big == 2 ** 31 - 1, torch.zeros_like(big), big)
    return int(info_acc.item())


def pdgetrs(grid: Grid, n, nb, LU_loc, ipiv, B, ops=None):
    """Solves A X = B in place with the factors of pdgetrf; B is replicated on every rank as the (nrhs, n)
    column-major image of the right-hand sides.  Column-oriented substitution: the owner of block column k
    holds all of L(:, k) and U(:, k), so it solves its block and folds it into the rest, then broadcasts."""
    _one_row_grid(grid)
    ops = ops or DeviceOps()
    Q, q = grid.Q, grid.q
    nblk = n // nb
    nrhs = B.shape[0]
    multi = Q > 1
    ops.laswp(B, ipiv, True)
    for k in range(nblk):                                          # L y = P b
        r0 = k * nb
        tail = B[:, r0:].contiguous()
        if q == k % Q:
            lk = k // Q
            col = LU_loc[lk * nb:(lk + 1) * nb, r0:]
            ops.trsm("L", "N", col[:, :nb], tail[:, :nb], diag="U")
            if n - r0 - nb > 0:
                ops.gemm("N", n - r0 - nb, nrhs, nb, -1.0, col[:, nb:], tail[:, :nb], 1.0, tail[:, nb:])
        if multi:
            dist.broadcast(tail, src=k % Q)
        B[:, r0:] = tail
    for k in range(nblk - 1, -1, -1):                              # U x = y
        r1 = (k + 1) * nb
        head = B[:, :r1].contiguous()
        if q == k % Q:
            lk = k // Q
            col = LU_loc[lk * nb:(lk + 1) * nb, :r1]
            ops.trsm("U", "N", col[:, r1 - nb:], head[:, r1 - nb:])
            if r1 - nb > 0:
                ops.gemm("N", r1 - nb, nrhs, nb, -1.0, col[:, :r1 - nb], head[:, r1 - nb:], 1.0, head[:, :r1 - nb])
        if multi:
            dist.broadcast(head, src=k % Q)
        B[:, :r1] = head
    return B
