#!/usr/bin/env python
"""bench.py -- headline benchmark of the llacuda hot path (BASELINE.json metric).

Metric: FP64 GFLOP/s for mm, LU-solve and Cholesky at n = 16384 (double), whole job.
One "step" = one pass of the three hot-path calls on ONE set of synthetic operands:

    mm        C = A B                         (dgemm,  2 n^3 flop)
    solve     X = A^-1 B, nrhs = 64           (dgesv,  2/3 n^3 + 2 n^2 nrhs flop)
    cholesky  A = U^T U                       (dpotrf, 1/3 n^3 flop)

N = 1: the single-GPU library on device-resident operands.
N > 1 (torchrun, one rank per GPU): the SAME three problems, each ONE matrix partitioned block-cyclically over the N
GPUs (strong scaling): SUMMA DGEMM on the most-square grid, LU and Cholesky on their grids, panels broadcast with NCCL
over NVLink by the C++ engine (csrc/dist*.cu) through its `llacuda_*_local` entry points.  The operands are generated
block by block from seeds that do not depend on N, so every N factors the same matrices: the printed ipiv CRC must
agree across N (and with host OpenBLAS, checked at N = 1 in the cpu_baseline leg).

`value`  : operands already in HBM, every op bracketed by CUDA events on the stream its kernels run on, flop over the
           summed event time (max over ranks).
`e2e`    : the same step through the reference-facing C ABI with HOST buffers that are ordinary PAGEABLE memory (what
           LLA passes: reference src/pinned-array.lisp:124-141), copies inside the timed region.  N = 1: dgemm_ / dgesv_ /
           dpotrf_; N > 1: rank 0 alone calls llacuda_pdgemm / pdgesv / pdpotrf, which drive all N GPUs from one process
           (the way a Lisp image would), the other ranks stand by.  `e2e.pinned` repeats it with page-locked buffers.
`roofline`: the dominant kernel (DMMA DGEMM, also the trailing update of both factorisations) against the FP64 tensor
           roof measured in the same run by a register-resident DMMA loop (MEASURED_PEAKS.json has no FP64 figure).
`cpu_baseline` / `--impl reference`: the calls LLA would issue, against the host OpenBLAS a user of the reference loads
           (oracle/lla_oracle.py back end), on the box's host cores, at the workload's own n.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_DEFAULT = 16384
NRHS = 64
GEN_NB = 512          # block size of the seeded generator (independent of the process grid)
DIST_NB = 512         # block size of the distributed factorisations
EPS = 2.220446049250313e-16


def flops(n, nrhs=NRHS):
    return {"mm": 2.0 * n ** 3, "solve": (2.0 / 3.0) * n ** 3 + 2.0 * n * n * nrhs, "cholesky": n ** 3 / 3.0}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region (a step at N = 8 is 160 ms)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); pw.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- synthetic operands
def _seed(kind, I, J):
    return (kind * 1_000_003 + I) * 1_000_003 + J + 12345


def local_image(kind, n, P, Q, rank, dev, scale=1.0, spd=False, nb=None):
    """Block-cyclic local part (image (n/Q, n/P): row j = local column j) of the seeded global n x n matrix `kind`,
    distributed in nb x nb blocks (default GEN_NB).  The matrix itself is defined GEN_NB block by GEN_NB block: global
    block (I, J) = scale * U[0,1) from a generator seeded by (kind, I, J) only, so every process grid and every layout
    block size sees the same matrix.  spd: only blocks I <= J are drawn and n is added to the diagonal -- the symmetric
    matrix named by that upper triangle is strictly diagonally dominant, hence SPD (POTRF 'U' reads nothing else)."""
    import torch
    g0 = GEN_NB
    nb = nb or g0
    assert nb % g0 == 0 and n % (nb * P) == 0 and n % (nb * Q) == 0
    sub = nb // g0
    p, q = rank // Q, rank % Q
    nblk = n // nb
    rows = [I * sub + s for I in range(nblk) if I % P == p for s in range(sub)]      # global GEN_NB block rows, local order
    cols = [J * sub + s for J in range(nblk) if J % Q == q for s in range(sub)]
    out = torch.zeros((len(cols) * g0, len(rows) * g0), dtype=torch.float64, device=dev)
    g = torch.Generator(device=dev)
    for lj, J in enumerate(cols):
        for li, I in enumerate(rows):
            if spd and I > J:
                continue
            g.manual_seed(_seed(kind, I, J))
            blk = torch.rand((g0, g0), dtype=torch.float64, device=dev, generator=g)
            if scale != 1.0:
                blk *= scale
            if spd and I == J:
                blk.diagonal().add_(float(n))
            out[lj * g0:(lj + 1) * g0, li * g0:(li + 1) * g0] = blk
    return out


def rhs_image(n, nrhs, dev):
    import torch
    g = torch.Generator(device=dev).manual_seed(777)
    return torch.rand((nrhs, n), dtype=torch.float64, device=dev, generator=g) * 10


def crc(arr):
    return f"{zlib.crc32(arr.tobytes()) & 0xffffffff:08x}"


# --------------------------------------------------------------------------- CPU (reference) leg
def cpu_leg(n, steps, warmup, nrhs=NRHS, seed=0, operands=None):
    """LLA's calls against host OpenBLAS through the oracle back end (test infrastructure; this leg is the only
    place bench.py executes oracle/).  operands = (a, b, rhs, spd) column-major images to run on instead of drawing
    its own (the GPU leg's matrices: their pivots are then compared bit for bit)."""
    import numpy as np
    from oracle import lla_oracle as O
    cores = os.cpu_count() or 1
    be = O.openblas(threads=cores)
    info_ob = O.openblas_info()
    dbl = O.DOUBLE
    I = lambda v: ctypes.byref(ctypes.c_int(int(v)))
    ch = lambda s: ctypes.byref(ctypes.c_char(s.encode()))
    one, zero = ctypes.byref(ctypes.c_double(1.0)), ctypes.byref(ctypes.c_double(0.0))
    P = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    if operands is None:
        rng = np.random.default_rng(seed)
        a = rng.uniform(0, 10, (n, n))
        b = rng.uniform(0, 10, (n, n))
        rhs = rng.uniform(0, 10, (nrhs, n))      # column-major n x nrhs image
        g = rng.uniform(0, 1, (n, n))
        spd = np.zeros((n, n))
        be.fn(dbl, "gemm")(ch("T"), ch("N"), I(n), I(n), I(n), one, P(g), I(n), P(g), I(n), zero, P(spd), I(n))
        spd[np.diag_indices(n)] += n
        del g
    else:
        a, b, rhs, spd = operands
    c = np.zeros((n, n))
    work = np.empty((n, n))
    xw = np.empty((nrhs, n))
    ipiv = np.zeros(n, dtype=np.int32)
    info = ctypes.c_int(0)
    t = {"mm": 0.0, "solve": 0.0, "cholesky": 0.0}
    keep = {}
    for it in range(warmup + steps):
        timed = it >= warmup
        t0 = time.perf_counter()
        be.fn(dbl, "gemm")(ch("N"), ch("N"), I(n), I(n), I(n), one, P(b), I(n), P(a), I(n), zero, P(c), I(n))
        t1 = time.perf_counter()
        np.copyto(work, a); np.copyto(xw, rhs)
        t2 = time.perf_counter()
        be.fn(dbl, "gesv")(I(n), I(nrhs), P(work), I(n), P(ipiv), P(xw), I(n), ctypes.byref(info))
        t3 = time.perf_counter()
        assert info.value == 0
        if it == 0:
            keep = {"ipiv": ipiv.copy(), "x": xw.copy(), "c_rows": c[:4].copy(), "lu_rows": work[:4].copy()}
        np.copyto(work, spd)
        t4 = time.perf_counter()
        be.fn(dbl, "potrf")(ch("U"), I(n), P(work), I(n), ctypes.byref(info))
        t5 = time.perf_counter()
        assert info.value == 0
        if it == 0:
            keep["chol_cols"] = work[-4:].copy()     # last 4 columns of U (image rows)
        if timed:
            t["mm"] += t1 - t0; t["solve"] += t3 - t2; t["cholesky"] += t5 - t4
    f = flops(n, nrhs)
    total_t = sum(t.values())
    total_f = sum(f.values()) * steps
    return {"value": total_f / total_t / 1e9, "ms_per_step": total_t / steps * 1e3,
            "ops": {k: f[k] * steps / t[k] / 1e9 for k in t}, "cores": cores, "n": n,
            "blas": f"OpenBLAS {info_ob['corename']} threads={info_ob['threads']}", "keep": keep}


def cpu_n_for_budget(n, passes, budget_s):
    """the workload's n unless `passes` passes at ~50 GFLOP/s per core (measured on the GPU box's host: 80) would exceed the budget"""
    cores = os.cpu_count() or 1
    while n > 2048 and sum(flops(n).values()) * passes / (50e9 * cores) > budget_s:
        n //= 2
    return n


# --------------------------------------------------------------------------- GPU: device-resident leg
def gpu_value_leg(args, rank, world, local_rank, ctl):
    import numpy as np
    import torch
    from lla_b200 import _lib, device as D, pdist as PD
    lib = _lib.lib()
    lib.llacuda_launch_count.restype = ctypes.c_ulonglong
    n, nrhs = args.n, NRHS
    dev = torch.device("cuda", local_rank)
    dist_mode = world > 1
    if dist_mode:
        import torch.distributed as dist
        PD.init_from_torch()
        dstream = torch.cuda.ExternalStream(PD.stream_ptr(), device=dev)
    else:
        D.use_torch_stream()

    dmma_tf, _ = D.probe_fp64_peak(0, 20000)
    dfma_tf, _ = D.probe_fp64_peak(1, 20000)

    gemm_grid = PD.default_grid(world)
    lu_grid = _grid_env("LLACUDA_LU_GRID", world, (1, world))
    chol_grid = _grid_env("LLACUDA_CHOL_GRID", world, (1, world))
    gemm_nb = 2048 if (n % (2048 * int(PD.lcm(*gemm_grid))) == 0) else GEN_NB

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # ---- operands, block-cyclic parts (N = 1: the whole matrix)
    Pg, Qg = gemm_grid
    A_g = local_image(1, n, Pg, Qg, rank, dev, 10.0, nb=gemm_nb)
    B_g = local_image(2, n, Pg, Qg, rank, dev, 10.0, nb=gemm_nb)
    C_g = torch.zeros_like(A_g)
    Pl, Ql = lu_grid
    A_l = A_g if world == 1 else local_image(1, n, Pl, Ql, rank, dev, 10.0, nb=DIST_NB)
    W_l = torch.empty_like(A_l)
    Pc, Qc = chol_grid
    S_c = local_image(3, n, Pc, Qc, rank, dev, 1.0, spd=True, nb=DIST_NB)
    W_c = torch.empty_like(S_c)
    RHS = rhs_image(n, nrhs, dev)
    X = torch.empty_like(RHS)
    ipiv = torch.zeros(n, dtype=torch.int32, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    f = flops(n, nrhs)
    t_ms = {"mm": 0.0, "solve": 0.0, "cholesky": 0.0}
    gemm_ms, infos = [], []
    sampler = ClockSampler(local_rank)
    launches0 = 0

    def timed_call(fn, stream):
        e0, e1 = ev(), ev()
        if dist_mode:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    cur = torch.cuda.current_stream()
    for it in range(args.warmup + args.steps):
        timed = it >= args.warmup
        if it == args.warmup:
            torch.cuda.synchronize()
            if dist_mode:
                dist.barrier()
            sampler.start()
            launches0 = lib.llacuda_launch_count()
        if dist_mode:
            PD.set_grid(Pg, Qg)
            ms_mm = timed_call(lambda: PD.pdgemm_local(n, n, n, gemm_nb, 1.0, A_g.data_ptr(), n // Pg, B_g.data_ptr(), n // Pg,
                                                       0.0, C_g.data_ptr(), n // Pg), dstream)
            W_l.copy_(A_l); X.copy_(RHS)
            PD.set_grid(Pl, Ql)

            def solve():
                infos.append(PD.pdgesv_local(n, DIST_NB, W_l.data_ptr(), n // Pl, ipiv.data_ptr(), X.data_ptr(), n, nrhs))
            ms_solve = timed_call(solve, dstream)
            W_c.copy_(S_c)
            PD.set_grid(Pc, Qc)
            ms_chol = timed_call(lambda: infos.append(PD.pdpotrf_local(n, DIST_NB, W_c.data_ptr(), n // Pc)), dstream)
        else:
            # LLA's mm call: C^T = B^T A^T, i.e. dgemm(N, N) on the row-major buffers taken as column-major
            ms_mm = timed_call(lambda: D.dgemm("N", "N", n, n, n, 1.0, B_g.data_ptr(), n, A_g.data_ptr(), n, 0.0, C_g.data_ptr(), n), cur)
            W_l.copy_(A_l); X.copy_(RHS)                 # restore in-place inputs (untimed)
            ms_solve = timed_call(lambda: D.dgesv(n, nrhs, W_l.data_ptr(), n, ipiv.data_ptr(), X.data_ptr(), n, info.data_ptr()), cur)
            infos.append(int(info.item()))
            W_c.copy_(S_c)
            ms_chol = timed_call(lambda: D.dpotrf("U", n, W_c.data_ptr(), n, info.data_ptr()), cur)
            infos.append(int(info.item()))
        if timed:
            t_ms["mm"] += ms_mm; t_ms["solve"] += ms_solve; t_ms["cholesky"] += ms_chol
            gemm_ms.append(ms_mm)
    torch.cuda.synchronize()
    launches = lib.llacuda_launch_count() - launches0
    if dist_mode:
        dist.barrier()
    clocks = sampler.stop()
    total_ms = sum(t_ms.values())
    if dist_mode:
        tt = torch.tensor([total_ms, t_ms["mm"], t_ms["solve"], t_ms["cholesky"], float(launches)], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms, t_ms["mm"], t_ms["solve"], t_ms["cholesky"] = (float(v) for v in tt.tolist()[:4])
        lt = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())

    # ---- parity of what was just timed (rank 0; the checker is torch fp64 on the device, no oracle here)
    parity = {"info": [int(i) for i in infos if i != 0][:4], "bound": 10}
    ipiv_h = ipiv.cpu().numpy()
    parity["ipiv_crc32"] = crc(ipiv_h)
    host_ops = None
    if rank == 0:
        fullA = A_g if world == 1 else local_image(1, n, 1, 1, 0, dev, 10.0)          # image of the whole A
        Xm, Am, Bm = X.t(), fullA.t(), RHS.t()
        parity["solve_scaled_residual"] = (torch.linalg.norm(Am @ Xm - Bm) /
                                           (n * EPS * torch.linalg.norm(Am) * torch.linalg.norm(Xm))).item()
        # GEMM: 8 local rows of my part of C against fp64 dot products of the whole operands
        fullB = B_g if world == 1 else local_image(2, n, 1, 1, 0, dev, 10.0)
        sel = torch.randint(0, C_g.shape[1], (8,), generator=torch.Generator().manual_seed(5)).tolist()
        if world == 1:
            # LLA's mm: dgemm(N, N) on the untransposed row-major buffers, C_cm = B_cm A_cm, i.e. C_img = A_img B_img
            ref = fullA @ fullB[:, sel]
            got = C_g[:, sel]
            scale = torch.linalg.norm(fullA, dim=1)[:, None] * torch.linalg.norm(fullB[:, sel], dim=0)[None, :]
        else:
            # C = A B; images X_img[j, i] = X(i, j): C(rows, cols)^T = B_img[cols, :] A_img[:, rows]; 8 of my local rows
            bcg = PD.BlockCyclic(n, n, gemm_nb, Pg, Qg, 0)
            rows_g = torch.as_tensor(bcg.local_rows(), device=dev)[sel]
            cols_g = torch.as_tensor(bcg.local_cols(), device=dev)
            Arow = fullA[:, rows_g]
            Bcol = fullB[cols_g, :]
            ref = Bcol @ Arow
            got = C_g[:, sel]
            scale = torch.linalg.norm(Bcol, dim=1)[:, None] * torch.linalg.norm(Arow, dim=0)[None, :]
        parity["mm_max_err_over_bound"] = (torch.abs(got - ref) / (4 * (n ** 0.5) * EPS * scale)).max().item()
        del fullB, ref, got, scale
        # Cholesky: solve with the factor and check the residual against the symmetric matrix the upper triangle names
        torch.cuda.empty_cache()
    # every rank takes part in the distributed triangular solves
    Y = RHS.clone()
    if dist_mode:
        PD.set_grid(Pc, Qc)
        PD.pdpotrs_local(n, DIST_NB, W_c.data_ptr(), n // Pc, Y.data_ptr(), n, nrhs)
    else:
        D.dpotrs("U", n, nrhs, W_c.data_ptr(), n, Y.data_ptr(), n)
    torch.cuda.synchronize()
    if rank == 0:
        Su = S_c if world == 1 else local_image(3, n, 1, 1, 0, dev, 1.0, spd=True)
        Su = torch.triu(Su.t())                          # matrix upper triangle
        Sfull = Su + torch.triu(Su, 1).t()
        del Su
        Ym, Bm = Y.t(), RHS.t()
        parity["cholesky_solve_scaled_residual"] = (torch.linalg.norm(Sfull @ Ym - Bm) /
                                                    (n * EPS * torch.linalg.norm(Sfull) * torch.linalg.norm(Ym))).item()
        if world > 1:
            # ipiv against the single-GPU factorisation of the same matrix (bit for bit)
            D.use_torch_stream()
            W1 = fullA.clone()
            ip1 = torch.zeros(n, dtype=torch.int32, device=dev)
            D.dgetrf(n, n, W1.data_ptr(), n, ip1.data_ptr(), info.data_ptr())
            torch.cuda.synchronize()
            parity["ipiv_equal_single_gpu"] = bool(np.array_equal(ip1.cpu().numpy(), ipiv_h))
            del W1
        elif not args.no_cpu_baseline and not args.profile:
            host_ops = (fullA.cpu().numpy(), B_g.cpu().numpy(), RHS.cpu().numpy(), Sfull.t().contiguous().cpu().numpy())
            parity["_gpu"] = {"ipiv": ipiv_h, "x": X.cpu().numpy(), "lu_rows": W_l[:4].cpu().numpy(), "c_rows": C_g[:4].cpu().numpy(),
                              "chol_cols": W_c[-4:].cpu().numpy()}
        del Sfull, fullA
    res = {"t_ms": t_ms, "total_ms": total_ms, "launches": int(launches), "clocks": clocks, "dmma_tf": dmma_tf, "dfma_tf": dfma_tf,
           "gemm_ms": statistics.mean(gemm_ms), "parity": parity, "host_ops": host_ops,
           "grids": {"gemm": list(gemm_grid), "lu": list(lu_grid), "cholesky": list(chol_grid), "gemm_nb": gemm_nb, "dist_nb": DIST_NB}}
    del A_g, B_g, C_g, A_l, W_l, S_c, W_c, X, Y
    torch.cuda.empty_cache()
    return res


def _grid_env(name, world, default):
    e = os.environ.get(name)
    if e:
        try:
            p, q = (int(v) for v in e.lower().split("x"))
            if p * q == world:
                return (p, q)
        except ValueError:
            pass
    return default


# --------------------------------------------------------------------------- secondary legs (device-resident)
def batched_leg(rank, world, local_rank):
    """BASELINE configs[3]: 200 000 independent 32 x 32 double systems, one contiguous slice per GPU, no collective.
    HBM-bound by SURVEY.md section 8(d): 8 704 algorithmic bytes per system (A 8 192 + b 256 in, x 256 out)."""
    import torch
    from lla_b200 import device as D, batched as BT
    dev = torch.device("cuda", local_rank)
    D.use_torch_stream()
    batch, n = 200_000, 32
    lo, hi = BT.shard_range(batch, rank, world)
    cnt = hi - lo
    g = torch.Generator(device=dev).manual_seed(4 + rank)
    A0 = torch.rand((cnt, n, n), dtype=torch.float64, device=dev, generator=g) * 10
    B0 = torch.rand((cnt, n), dtype=torch.float64, device=dev, generator=g) * 10
    A, B = A0.clone(), B0.clone()
    ipiv = torch.zeros((cnt, n), dtype=torch.int32, device=dev)
    info = torch.zeros(cnt, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # larger than the 126 MB L2
    best, times = None, []
    for it in range(6):
        B.copy_(B0)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        D.dgesv_batched(n, 1, cnt, A.data_ptr(), n * n, ipiv.data_ptr(), B.data_ptr(), n, info.data_ptr(), 0)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            times.append(e0.elapsed_time(e1))
    ms = statistics.mean(times)
    Am = A0.transpose(1, 2)                                             # systems are column-major images
    r = torch.linalg.norm(Am @ B[:, :, None] - B0[:, :, None], dim=(1, 2)) / (
        n * EPS * torch.linalg.norm(Am, dim=(1, 2)) * torch.linalg.norm(B, dim=1))
    out = {"ms": ms, "systems": cnt, "max_scaled_residual": r.max().item(), "info_nonzero": int((info != 0).sum().item())}
    del A0, B0, A, B, flush
    torch.cuda.empty_cache()
    return out


def sgemm_leg(local_rank):
    """single precision on tcgen05 (3xTF32, TMA + TMEM): n = 8192, NN, device-resident, rank 0's GPU only"""
    import torch
    from lla_b200 import device as D
    dev = torch.device("cuda", local_rank)
    D.use_torch_stream()
    n = 8192
    g = torch.Generator(device=dev).manual_seed(21)
    A = torch.rand((n, n), dtype=torch.float32, device=dev, generator=g) - 0.5
    B = torch.rand((n, n), dtype=torch.float32, device=dev, generator=g) - 0.5
    C = torch.empty_like(A)
    times = []
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        D.sgemm("N", "N", n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, 0.0, C.data_ptr(), n)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            times.append(e0.elapsed_time(e1))
    ms = statistics.mean(times)
    # C_img = B_img A_img in torch terms (column-major product A B): sampled columns against float64
    sel = [5, 4097, 8191]
    ref = (B.double()[sel, :] @ A.double()).float()
    err = ((C[sel, :] - ref).abs().max() / ref.abs().max()).item()
    del A, B, C
    torch.cuda.empty_cache()
    return {"ms": ms, "tflops": 2.0 * n ** 3 / ms / 1e9, "max_err_over_max_abs": err}


def n32768_leg(args, rank, world, local_rank):
    """BASELINE configs[4] and the north_star scaling target: DGEMM / ZGEMM / Cholesky / LU at n = 32768, one run each
    after one warm-up, same engine, same seeded matrices for every N."""
    import torch
    from lla_b200 import device as D, pdist as PD
    n = 32768
    dev = torch.device("cuda", local_rank)
    dist_mode = world > 1
    out = {}
    if dist_mode:
        import torch.distributed as dist
        stream = torch.cuda.ExternalStream(PD.stream_ptr(), device=dev)
    else:
        D.use_torch_stream()
        stream = torch.cuda.current_stream()
    info = torch.zeros(1, dtype=torch.int32, device=dev)

    def run(fn, reps=2):
        ms = None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if dist_mode:
                dist.barrier()
            torch.cuda.synchronize()
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        if dist_mode:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    Pg, Qg = PD.default_grid(world)
    nbg = 2048
    A = local_image(11, n, Pg, Qg, rank, dev, 2.0, nb=nbg).sub_(1.0)
    B = local_image(12, n, Pg, Qg, rank, dev, 2.0, nb=nbg).sub_(1.0)
    C = torch.empty_like(A)
    if dist_mode:
        PD.set_grid(Pg, Qg)
        ms = run(lambda: PD.pdgemm_local(n, n, n, nbg, 1.0, A.data_ptr(), n // Pg, B.data_ptr(), n // Pg, 0.0, C.data_ptr(), n // Pg))
    else:
        ms = run(lambda: D.dgemm("N", "N", n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, 0.0, C.data_ptr(), n))
    out["dgemm"] = {"ms": ms, "gflops": 2.0 * n ** 3 / ms / 1e6}
    del C
    if not args.skip_zgemm:
        # complex operands: re = the real matrices above, im = a second draw
        Az = torch.complex(A, local_image(13, n, Pg, Qg, rank, dev, 2.0, nb=nbg).sub_(1.0))
        del A
        Bz = torch.complex(B, local_image(14, n, Pg, Qg, rank, dev, 2.0, nb=nbg).sub_(1.0))
        del B
        Cz = torch.empty_like(Az)
        if dist_mode:
            ms = run(lambda: PD.pzgemm_local(n, n, n, nbg, 1.0, Az.data_ptr(), n // Pg, Bz.data_ptr(), n // Pg, 0.0, Cz.data_ptr(), n // Pg), reps=2)
        else:
            ms = run(lambda: D.zgemm("N", "N", n, n, n, 1 + 0j, Az.data_ptr(), n, Bz.data_ptr(), n, 0j, Cz.data_ptr(), n), reps=2)
        out["zgemm"] = {"ms": ms, "gflops": 8.0 * n ** 3 / ms / 1e6}
        del Az, Bz, Cz
    else:
        del A, B
    torch.cuda.empty_cache()
    Pc, Qc = _grid_env("LLACUDA_CHOL_GRID", world, (1, world))
    S = local_image(15, n, Pc, Qc, rank, dev, 1.0, spd=True, nb=DIST_NB)
    W = torch.empty_like(S)
    infos = []

    def chol():
        W.copy_(S)
        torch.cuda.synchronize()
    if dist_mode:
        PD.set_grid(Pc, Qc)
    ms = None
    for _ in range(2):
        chol()
        ms = run((lambda: infos.append(PD.pdpotrf_local(n, DIST_NB, W.data_ptr(), n // Pc))) if dist_mode
                 else (lambda: D.dpotrf("U", n, W.data_ptr(), n, info.data_ptr())), reps=1)
    out["cholesky"] = {"ms": ms, "gflops": n ** 3 / 3.0 / ms / 1e6, "grid": [Pc, Qc]}
    del S, W
    torch.cuda.empty_cache()
    Pl, Ql = _grid_env("LLACUDA_LU_GRID", world, (1, world))
    A = local_image(11, n, Pl, Ql, rank, dev, 10.0, nb=DIST_NB)
    W = torch.empty_like(A)
    ipiv = torch.zeros(n, dtype=torch.int32, device=dev)
    if dist_mode:
        PD.set_grid(Pl, Ql)
    for _ in range(2):
        W.copy_(A)
        torch.cuda.synchronize()
        ms = run((lambda: infos.append(PD.pdgetrf_local(n, DIST_NB, W.data_ptr(), n // Pl, ipiv.data_ptr()))) if dist_mode
                 else (lambda: D.dgetrf(n, n, W.data_ptr(), n, ipiv.data_ptr(), info.data_ptr())), reps=1)
    out["lu"] = {"ms": ms, "gflops": 2.0 * n ** 3 / 3.0 / ms / 1e6, "grid": [Pl, Ql], "ipiv_crc32": crc(ipiv.cpu().numpy())}
    del A, W
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------- e2e legs (host buffers through the C ABI)
def e2e_leg(args, world, pageable, steps):
    """One process, HOST buffers, copies inside the timed region.  world == 1: the Fortran drop-ins LLA binds;
    world > 1: the single-process multi-GPU entry points (llacuda_pd*), all `world` GPUs driven from this process."""
    import numpy as np
    from lla_b200 import _lib, pdist as PD
    lib = _lib.lib()
    n, nrhs = args.n, NRHS
    vp = ctypes.c_void_p
    hbuf = []

    def hostmat(rows, cols):
        if pageable:
            return np.empty((rows, cols))
        p = vp()
        _lib.check(lib.llacuda_host_alloc(ctypes.byref(p), ctypes.c_size_t(rows * cols * 8)), "host_alloc")
        hbuf.append(p)
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_double)), shape=(rows, cols))
    rng = np.random.default_rng(99)
    hA, hB, hC = hostmat(n, n), hostmat(n, n), hostmat(n, n)
    hX = hostmat(nrhs, n)
    hA[...] = rng.uniform(0, 10, (n, n)); hB[...] = rng.uniform(0, 10, (n, n))
    hR = rng.uniform(0, 10, (nrhs, n))
    # SPD master copy: symmetric, entries in [0, 1), n on the diagonal (strictly diagonally dominant)
    hS0 = rng.uniform(0, 1, (n, n))
    hS0 = np.triu(hS0) + np.triu(hS0, 1).T
    hS0[np.diag_indices(n)] += n
    I = lambda v: ctypes.byref(ctypes.c_int(int(v)))
    ch = lambda s: ctypes.byref(ctypes.c_char(s.encode()))
    one, zero = ctypes.byref(ctypes.c_double(1.0)), ctypes.byref(ctypes.c_double(0.0))
    P = lambda x: x.ctypes.data_as(vp)
    hipiv = np.zeros(n, dtype=np.int32)
    hinfo = ctypes.c_int(0)
    names = ("llacuda_pdgemm", "llacuda_pdgesv", "llacuda_pdpotrf") if world > 1 else ("dgemm_", "dgesv_", "dpotrf_")
    gemm, gesv, potrf = (getattr(lib, nm) for nm in names)
    for fn in (gemm, gesv, potrf):
        fn.restype = None
    te = {"mm": 0.0, "solve": 0.0, "cholesky": 0.0}
    per_step_ms = []
    resid = None
    E2E_WARMUP = 2      # untimed passes: the first one grows the device pool and the pinned ring, the second runs warm
    for it in range(E2E_WARMUP + steps):
        timed = it >= E2E_WARMUP
        t0 = time.perf_counter()
        gemm(ch("N"), ch("N"), I(n), I(n), I(n), one, P(hB), I(n), P(hA), I(n), zero, P(hC), I(n))
        t1 = time.perf_counter()
        _lib.check(lib.llacuda_last_error_code(), names[0])
        np.copyto(hC, hA); np.copyto(hX, hR)          # hC doubles as the in-place work matrix
        t2 = time.perf_counter()
        gesv(I(n), I(nrhs), P(hC), I(n), P(hipiv), P(hX), I(n), ctypes.byref(hinfo))
        t3 = time.perf_counter()
        assert hinfo.value == 0, (names[1], hinfo.value, _lib.last_error())
        if it == 0:
            resid = float(np.linalg.norm(hX @ hA - hR) / (n * EPS * np.linalg.norm(hA) * np.linalg.norm(hX)))   # images: X^T A^T = B^T
        np.copyto(hC, hS0)
        t4 = time.perf_counter()
        potrf(ch("U"), I(n), P(hC), I(n), ctypes.byref(hinfo))
        t5 = time.perf_counter()
        assert hinfo.value == 0, (names[2], hinfo.value, _lib.last_error())
        if timed:
            te["mm"] += t1 - t0; te["solve"] += t3 - t2; te["cholesky"] += t5 - t4
            per_step_ms.append([round((t1 - t0) * 1e3, 1), round((t3 - t2) * 1e3, 1), round((t5 - t4) * 1e3, 1)])
    f = flops(n, nrhs)
    total = sum(te.values())
    for p in hbuf:
        lib.llacuda_host_free(p)
    return {"value": sum(f.values()) * steps / total / 1e9, "ops": {k: f[k] * steps / te[k] / 1e9 for k in te},
            "steps": steps, "warmup": E2E_WARMUP, "per_step_ms_mm_solve_cholesky": per_step_ms, "solve_scaled_residual": resid,
            "entry_points": list(names)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="llacuda", choices=["llacuda", "reference"])
    ap.add_argument("--n", type=int, default=N_DEFAULT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the batched and n = 32768 legs")
    ap.add_argument("--skip-zgemm", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="profiler pass: device-resident leg only, warm-up not forced to >= 3")
    ap.add_argument("--cpu-n", type=int, default=0, help="override the CPU sample size (testing)")
    args = ap.parse_args()
    if args.impl == "llacuda" and not args.profile:
        args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.n
    if n % GEN_NB:
        raise SystemExit(f"--n must be a multiple of {GEN_NB}")
    config = {"workload": f"mm + solve(nrhs={NRHS}) + cholesky, double-float, n={n} (BASELINE.json metric shapes; "
                          "solve = configs[1] call at the metric size, cholesky = configs[2])",
              "n": n, "nrhs": NRHS,
              "parallelism": (f"one problem of each kind partitioned 2-D block-cyclic over {world} GPUs (NCCL panel broadcasts)"
                              if world > 1 else "single GPU"),
              "l2": "operands are 2 GiB each at n=16384, far larger than the 126 MB L2; no flush needed",
              "in_place_inputs": "restored between steps outside the event pairs"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        passes = args.steps + args.warmup
        ncpu = args.cpu_n or cpu_n_for_budget(n, passes, 900.0)
        r = cpu_leg(ncpu, args.steps, args.warmup)
        if ncpu != n:   # never label a sample as the workload
            config = dict(config, n=ncpu, workload=config["workload"].replace(f"n={n}", f"n={ncpu}"),
                          workload_n=n, note=f"{passes} passes at n={n} exceed the time budget on {r['cores']} cores; measured at n={ncpu}")
        sample = (f"the same three calls (dgemm/dgesv/dpotrf as LLA issues them) at n={ncpu}, nrhs={NRHS}, "
                  f"{args.steps} timed steps after {args.warmup} warm-up; {r['blas']}")
        line = {"impl": "reference", "metric": "fp64_gflops_mm_solve_cholesky", "value": r["value"], "unit": "GFLOP/s",
                "n_gpus": 0, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "n_measured": ncpu, "same_config": ncpu == n, "ops_gflops": r["ops"],
                "cpu_baseline": {"value": r["value"], "unit": "GFLOP/s", "cores": r["cores"], "kind": "port", "sample": sample},
                "e2e": {"value": r["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    import torch
    from lla_b200 import _lib, pdist as PD
    torch.cuda.set_device(local_rank)
    ctl = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ctl = dist.new_group(backend="gloo")      # control-plane barriers that launch nothing on the GPUs
    _lib.init(local_rank)
    g = gpu_value_leg(args, rank, world, local_rank, ctl)
    secondary = {}
    if not args.no_secondary and not args.profile:
        b = batched_leg(rank, world, local_rank)
        if world > 1:
            t = torch.tensor([b["ms"], b["max_scaled_residual"], float(b["info_nonzero"])], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            b["ms"], b["max_scaled_residual"], b["info_nonzero"] = float(t[0]), float(t[1]), int(t[2])
        hbm = 6539.5
        try:
            hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
        gbs = 200_000 * 8704 / (b["ms"] * 1e-3) / 1e9
        secondary["batched_32x32"] = {
            "workload": "configs[3]: 200 000 independent 32 x 32 double systems (dgesv, 1 rhs), one slice per GPU, no collective",
            "ms": b["ms"], "systems_per_s": 200_000 / (b["ms"] * 1e-3), "max_scaled_residual": b["max_scaled_residual"],
            "info_nonzero": b["info_nonzero"],
            "roofline": {"kernel": "lla_dgesv_batched_smem_kernel", "bound": "hbm", "achieved": gbs, "peak": hbm * world, "unit": "GB/s",
                         "frac": gbs / (hbm * world), "algorithmic_bytes_per_system": 8704,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured) x GPUs; L2 flushed between launches",
                         "note": "DRAM traffic is the algorithmic minimum (ncu: 1.74 GB per launch) but the kernel is issue-bound, not "
                                 "HBM-bound: DGETF2's separately rounded multiply and subtract cost 1 800 FP64-pipe instructions per "
                                 "system at 2 issue cycles each = 0.62 ms per 200 000 systems on 148 SMs (profiles/r02_batched.md)"}}
        if rank == 0:
            sg = sgemm_leg(local_rank)
            bf16 = 1663.5
            try:
                bf16 = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
            except Exception:
                pass
            roof = bf16 / 2 / 3      # dense TF32 = half the measured bf16 rate; three TF32 products per FP32 product
            secondary["sgemm_8192"] = {
                "workload": "SGEMM n = 8192 (lla:mm on single-float arrays), one GPU, device-resident",
                "ms": sg["ms"], "tflops_fp32_equivalent": sg["tflops"], "max_err_over_max_abs": sg["max_err_over_max_abs"],
                "roofline": {"kernel": "lla_sgemm_tcgen05_kernel (tcgen05.mma kind::tf32 x3, TMA, TMEM)", "bound": "tensor",
                             "achieved": sg["tflops"], "peak": roof, "unit": "TFLOP/s", "frac": sg["tflops"] / roof,
                             "peak_source": "MEASURED_PEAKS.json bf16_tflops (of measured) / 2 (TF32) / 3 (3xTF32 split)"}}
        secondary["n32768"] = n32768_leg(args, rank, world, local_rank)
    # ---- e2e: one process, host buffers
    e2e = None
    if not args.no_e2e and not args.profile:
        if world > 1:
            PD.finalize()
            _lib.lib().llacuda_trim()
            torch.cuda.empty_cache()
            torch.cuda.synchronize()
            dist.barrier(group=ctl)
        if rank == 0:
            if world > 1:
                PD.init_all(world)
            steps = max(1, min(args.steps, 2))
            ep = e2e_leg(args, world, True, steps)
            pin = e2e_leg(args, world, False, steps)
            tri = sum(min(c0 + 512, n) * (min(c0 + 512, n) - c0) for c0 in range(0, n, 512)) * 8
            if world > 1:   # block columns travel whole; the right-hand sides go to every GPU
                tri = sum(min((J + 1) * DIST_NB, n) * DIST_NB for J in range(n // DIST_NB)) * 8
            h2d = 2 * n * n * 8 + (n * n * 8 + world * n * NRHS * 8) + tri
            d2h = n * n * 8 + (n * n * 8 + n * NRHS * 8 + n * 4 + 4) + (tri + 4)
            e2e = {"value": ep["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "steps": steps, "warmup": ep["warmup"], "per_step_ms_mm_solve_cholesky": ep["per_step_ms_mm_solve_cholesky"],
                   "host_memory": "pageable (plain NumPy arrays, what LLA's with-pinned-objects yields), staged by "
                   "the library's pinned ring (csrc/hostpipe.cu)", "ops": ep["ops"], "entry_points": ep["entry_points"],
                   "solve_scaled_residual": ep["solve_scaled_residual"],
                   "pinned": {"value": pin["value"], "ops": pin["ops"], "host_memory": "page-locked (llacuda_host_alloc)",
                              "per_step_ms_mm_solve_cholesky": pin["per_step_ms_mm_solve_cholesky"]},
                   "pageable_over_pinned": ep["value"] / pin["value"]}
            if world > 1:
                PD.finalize()
        if world > 1:
            dist.barrier(group=ctl)
    f = flops(n)
    total_f = sum(f.values()) * args.steps
    value = total_f / (g["total_ms"] * 1e-3) / 1e9
    cpu = None
    parity = g["parity"]
    gpu_side = parity.pop("_gpu", None)
    if rank == 0 and world == 1 and g["host_ops"] is not None:
        import numpy as np
        ncpu = args.cpu_n or cpu_n_for_budget(n, 2, 60.0)
        same = ncpu == n
        r = cpu_leg(ncpu, 1, 1, operands=g["host_ops"] if same else None)
        cpu = {"value": r["value"], "unit": "GFLOP/s", "cores": r["cores"], "kind": "port", "n": ncpu,
               "sample": f"dgemm/dgesv/dpotrf as LLA issues them at n={ncpu}, nrhs={NRHS}, 1 timed pass after 1 warm-up, on the GPU leg's "
                         f"own matrices; {r['blas']}" if same else
                         f"dgemm/dgesv/dpotrf as LLA issues them at n={ncpu} (n={n} does not fit the time budget on {r['cores']} cores), "
                         f"nrhs={NRHS}, 1 timed pass after 1 warm-up; {r['blas']}",
               "ops_gflops": r["ops"]}
        if same and gpu_side is not None:
            k = r["keep"]
            parity["ipiv_equal_host_openblas"] = bool(np.array_equal(k["ipiv"], gpu_side["ipiv"]))
            parity["ipiv_crc32_host_openblas"] = crc(k["ipiv"])
            parity["lu_rows_max_diff_over_50_n_eps"] = float(np.max(np.abs(k["lu_rows"] - gpu_side["lu_rows"])) /
                                                             (50 * n * EPS * np.abs(k["lu_rows"]).max()))
            parity["x_max_rel_diff_vs_openblas"] = float(np.max(np.abs(k["x"] - gpu_side["x"])) / np.abs(k["x"]).max())
            parity["cholesky_cols_max_diff_over_20_n_eps"] = float(np.max(np.abs(np.triu(k["chol_cols"].T).T - np.triu(gpu_side["chol_cols"].T).T))
                                                                   / (20 * n * EPS * np.abs(k["chol_cols"]).max()))
    if rank == 0:
        peak = g["dmma_tf"]
        ach = f["mm"] / (g["gemm_ms"] * 1e-3) / 1e12 / world
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "dgemm_traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {"metric": "fp64_gflops_mm_solve_cholesky", "value": value, "unit": "GFLOP/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": g["total_ms"] / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(config, grids=g["grids"]),
                "ops_gflops": {k: f[k] * args.steps / (g["t_ms"][k] * 1e-3) / 1e9 for k in f},
                "ops_frac_of_fp64_tensor_roof": {k: f[k] * args.steps / (g["t_ms"][k] * 1e-3) / 1e12 / (peak * world) for k in f},
                "parity": parity, "clocks": g["clocks"], "e2e": e2e, "gpu_launches": g["launches"],
                "roofline": {"kernel": "lla_dgemm_kernel<128,64,16,...> (DMMA m8n8k4, two CTAs/SM): the mm of the step"
                                       + (" (per GPU: the local SUMMA products)" if world > 1 else "") + "; algorithmic flop = 2 n^3",
                             "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                             "traffic": traffic if world == 1 else None,
                             "peak_source": "in-run llacuda_probe_fp64_peak DMMA loop; MEASURED_PEAKS.json has no FP64 figure "
                                            f"(DFMA loop: {g['dfma_tf']:.2f} TFLOP/s)"},
                "cpu_baseline": cpu, "secondary": secondary}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(group=ctl)
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
