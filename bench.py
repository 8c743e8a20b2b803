#!/usr/bin/env python
"""bench.py -- headline benchmark of the llacuda hot path (BASELINE.json metric).

Metric: FP64 GFLOP/s for mm, LU-solve and Cholesky at n = 16384 (double), whole job.
One "step" = one pass of the three hot-path calls on one batch of synthetic operands:

    mm        C = A B                         (dgemm,  2 n^3 flop)
    solve     X = A^-1 B, nrhs = 64           (dgesv,  2/3 n^3 + 2 n^2 nrhs flop)
    cholesky  A = L L^T                       (dpotrf, 1/3 n^3 flop)

`value`  : device-resident leg -- operands already in HBM, every op bracketed by CUDA events on
           the stream the kernels run on, flop / summed event time.
`e2e`    : the same step through the Fortran drop-ins (`dgemm_`, `dgesv_`, `dpotrf_`, the symbols
           LLA binds) with HOST buffers: host->device and device->host copies are inside the timed
           region (wall clock around each synchronous call).
`roofline`: the dominant kernel (DMMA DGEMM, also the trailing update of both factorisations)
           against the FP64 tensor roof measured in the same run by a register-resident DMMA loop
           (MEASURED_PEAKS.json carries no FP64 figure).
`cpu_baseline` / `--impl reference`: the calls LLA would issue, against the host OpenBLAS that a
           user of the reference loads (oracle/lla_oracle.py back end), on the box's host cores.

Launch: `python bench.py --gpus N --steps K --warmup W` (N=1), or under torchrun for N>1
(one rank per GPU; every rank runs its own replica of the step -- see DESIGN.md "multi-GPU").
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

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_DEFAULT = 16384
NRHS = 64


def flops(n, nrhs=NRHS):
    return {"mm": 2.0 * n ** 3, "solve": (2.0 / 3.0) * n ** 3 + 2.0 * n * n * nrhs, "cholesky": n ** 3 / 3.0}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)],
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


# --------------------------------------------------------------------------- CPU (reference) leg
def cpu_leg(n, steps, warmup, nrhs=NRHS, seed=0):
    """LLA's calls against host OpenBLAS through the oracle back end (test infrastructure; this
    leg is the only place bench.py executes oracle/)."""
    import numpy as np
    from oracle import lla_oracle as O
    cores = os.cpu_count() or 1
    be = O.openblas(threads=cores)
    info_ob = O.openblas_info()
    rng = np.random.default_rng(seed)
    a = rng.uniform(0, 10, (n, n))
    b = rng.uniform(0, 10, (n, n))
    c = np.zeros((n, n))
    rhs = rng.uniform(0, 10, (nrhs, n))      # column-major n x nrhs image
    spd = None
    dbl = O.DOUBLE
    I = lambda v: ctypes.byref(ctypes.c_int(int(v)))
    ch = lambda s: ctypes.byref(ctypes.c_char(s.encode()))
    one, zero = ctypes.byref(ctypes.c_double(1.0)), ctypes.byref(ctypes.c_double(0.0))
    P = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    # SPD input: G G^T + n I built with the same BLAS (outside the timed region)
    g = rng.uniform(0, 1, (n, n))
    spd = np.zeros((n, n))
    be.fn(dbl, "gemm")(ch("T"), ch("N"), I(n), I(n), I(n), one, P(g), I(n), P(g), I(n), zero, P(spd), I(n))
    spd[np.diag_indices(n)] += n
    del g
    work = np.empty((n, n))
    xw = np.empty((nrhs, n))
    ipiv = np.zeros(n, dtype=np.int32)
    info = ctypes.c_int(0)
    t = {"mm": 0.0, "solve": 0.0, "cholesky": 0.0}
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
        np.copyto(work, spd)
        t4 = time.perf_counter()
        be.fn(dbl, "potrf")(ch("U"), I(n), P(work), I(n), ctypes.byref(info))
        t5 = time.perf_counter()
        assert info.value == 0
        if timed:
            t["mm"] += t1 - t0; t["solve"] += t3 - t2; t["cholesky"] += t5 - t4
    f = flops(n, nrhs)
    total_t = sum(t.values())
    total_f = sum(f.values()) * steps
    return {"value": total_f / total_t / 1e9, "ms_per_step": total_t / steps * 1e3,
            "ops": {k: f[k] * steps / t[k] / 1e9 for k in t}, "cores": cores, "n": n,
            "blas": f"OpenBLAS {info_ob['corename']} threads={info_ob['threads']}"}


def pick_cpu_n(budget_s, steps_total):
    """largest n in {16384, 8192, 4096, 2048} whose (steps_total) passes fit the budget, assuming
    ~25 GFLOP/s per core sustained (calibrated in BASELINE.md section 3)."""
    cores = os.cpu_count() or 1
    rate = 25e9 * cores
    for n in (16384, 8192, 4096, 2048):
        if sum(flops(n).values()) * steps_total / rate <= budget_s:
            return n
    return 2048


# --------------------------------------------------------------------------- GPU legs
def gpu_legs(args, rank, world, local_rank):
    import numpy as np
    import torch
    from lla_b200 import _lib, device as D
    torch.cuda.set_device(local_rank)
    _lib.init(local_rank)
    D.use_torch_stream()
    lib = _lib.lib()
    lib.llacuda_launch_count.restype = ctypes.c_ulonglong
    n, nrhs = args.n, NRHS
    dev = torch.device("cuda", local_rank)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    # ---- FP64 roofs, measured in this run
    dmma_tf, _ = D.probe_fp64_peak(0, 20000)
    dfma_tf, _ = D.probe_fp64_peak(1, 20000)

    # ---- operands resident in HBM (column-major images)
    A = torch.rand((n, n), dtype=torch.float64, device=dev, generator=gen) * 10
    B = torch.rand((n, n), dtype=torch.float64, device=dev, generator=gen) * 10
    C = torch.zeros((n, n), dtype=torch.float64, device=dev)
    RHS = torch.rand((nrhs, n), dtype=torch.float64, device=dev, generator=gen) * 10
    W = torch.empty((n, n), dtype=torch.float64, device=dev)      # in-place work copy
    X = torch.empty((nrhs, n), dtype=torch.float64, device=dev)
    ipiv = torch.zeros(n, dtype=torch.int32, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    pA, pB, pC, pW, pX = (D.colmajor(t)[0] for t in (A, B, C, W, X))
    # SPD = G^T G + n I with the library's own DGEMM (G = A/10 in [0,1))
    S = torch.empty((n, n), dtype=torch.float64, device=dev)
    pS = D.colmajor(S)[0]
    D.dgemm("T", "N", n, n, n, 0.01, pA, n, pA, n, 0.0, pS, n)
    S.diagonal().add_(float(n))
    torch.cuda.synchronize()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    f = flops(n, nrhs)
    t_ms = {"mm": 0.0, "solve": 0.0, "cholesky": 0.0}
    gemm_ms = []
    if world > 1:
        import torch.distributed as dist
    sampler = ClockSampler(local_rank)
    launches0 = 0
    for it in range(args.warmup + args.steps):
        timed = it >= args.warmup
        if it == args.warmup:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            sampler.start()
            launches0 = lib.llacuda_launch_count()
        e = [ev() for _ in range(6)]
        e[0].record()
        D.dgemm("N", "N", n, n, n, 1.0, pB, n, pA, n, 0.0, pC, n)      # LLA's mm call: C^T = B^T A^T
        e[1].record()
        W.copy_(A); X.copy_(RHS)                                        # restore in-place inputs (untimed)
        e[2].record()
        D.dgesv(n, nrhs, pW, n, ipiv.data_ptr(), pX, n, info.data_ptr())
        e[3].record()
        W.copy_(S)
        e[4].record()
        D.dpotrf("U", n, pW, n, info.data_ptr())
        e[5].record()
        torch.cuda.synchronize()
        if timed:
            t_ms["mm"] += e[0].elapsed_time(e[1])
            t_ms["solve"] += e[2].elapsed_time(e[3])
            t_ms["cholesky"] += e[4].elapsed_time(e[5])
            gemm_ms.append(e[0].elapsed_time(e[1]))
    torch.cuda.synchronize()
    launches = lib.llacuda_launch_count() - launches0
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    total_ms = sum(t_ms.values())
    if world > 1:
        tt = torch.tensor([total_ms, t_ms["mm"], t_ms["solve"], t_ms["cholesky"]], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms, t_ms["mm"], t_ms["solve"], t_ms["cholesky"] = (float(v) for v in tt.tolist())
    # quick parity guard on what was just timed (oracle-free: residuals on device)
    Xm, Am, Bm = X.t(), A.t(), RHS.t()
    eps = 2.220446049250313e-16
    resid = (torch.linalg.norm(Am @ Xm - Bm) / (n * eps * torch.linalg.norm(Am) * torch.linalg.norm(Xm))).item()
    del Xm, Am, Bm

    res = {"t_ms": t_ms, "total_ms": total_ms, "launches": int(launches), "clocks": clocks,
           "dmma_tf": dmma_tf, "dfma_tf": dfma_tf, "gemm_ms": statistics.mean(gemm_ms), "solve_resid": resid,
           "info": int(info.item())}

    if args.profile:
        res["e2e"] = None
        return res
    # ---- end-to-end leg: Fortran drop-ins on pinned HOST buffers
    del B, C, S, W, X, A, RHS
    torch.cuda.empty_cache()
    e2e_steps = max(1, min(args.steps, 3))
    vp = ctypes.c_void_p
    hbuf = []

    def hostmat(rows, cols):
        p = vp()
        _lib.check(lib.llacuda_host_alloc(ctypes.byref(p), ctypes.c_size_t(rows * cols * 8)), "host_alloc")
        hbuf.append(p)
        arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_double)), shape=(rows, cols))
        return arr
    rng = np.random.default_rng(99 + rank)
    hA, hB, hC = hostmat(n, n), hostmat(n, n), hostmat(n, n)
    hX = hostmat(nrhs, n)
    hA[...] = rng.uniform(0, 10, (n, n)); hB[...] = rng.uniform(0, 10, (n, n))
    hR = rng.uniform(0, 10, (nrhs, n))
    # SPD master copy built on the device, kept in ordinary host memory
    Gd = torch.from_numpy(hA).to(dev)
    Sd = torch.empty_like(Gd)
    D.dgemm("T", "N", n, n, n, 0.01, D.colmajor(Gd)[0], n, D.colmajor(Gd)[0], n, 0.0, D.colmajor(Sd)[0], n)
    Sd.diagonal().add_(float(n))
    hS0 = Sd.cpu().numpy()
    del Gd, Sd
    torch.cuda.empty_cache()
    I = lambda v: ctypes.byref(ctypes.c_int(int(v)))
    ch = lambda s: ctypes.byref(ctypes.c_char(s.encode()))
    one, zero = ctypes.byref(ctypes.c_double(1.0)), ctypes.byref(ctypes.c_double(0.0))
    P = lambda x: x.ctypes.data_as(vp)
    hipiv = np.zeros(n, dtype=np.int32)
    hinfo = ctypes.c_int(0)
    for name in ("dgemm_", "dgesv_", "dpotrf_"):
        getattr(lib, name).restype = None
    te = {"mm": 0.0, "solve": 0.0, "cholesky": 0.0}
    for it in range(1 + e2e_steps):
        timed = it >= 1
        t0 = time.perf_counter()
        lib.dgemm_(ch("N"), ch("N"), I(n), I(n), I(n), one, P(hB), I(n), P(hA), I(n), zero, P(hC), I(n))
        t1 = time.perf_counter()
        np.copyto(hC, hA); np.copyto(hX, hR)          # hC doubles as the in-place work matrix
        t2 = time.perf_counter()
        lib.dgesv_(I(n), I(nrhs), P(hC), I(n), P(hipiv), P(hX), I(n), ctypes.byref(hinfo))
        t3 = time.perf_counter()
        assert hinfo.value == 0, hinfo.value
        np.copyto(hC, hS0)
        t4 = time.perf_counter()
        lib.dpotrf_(ch("U"), I(n), P(hC), I(n), ctypes.byref(hinfo))
        t5 = time.perf_counter()
        assert hinfo.value == 0, hinfo.value
        if timed:
            te["mm"] += t1 - t0; te["solve"] += t3 - t2; te["cholesky"] += t5 - t4
    e2e_total = sum(te.values())
    if world > 1:
        tt = torch.tensor([e2e_total], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_total = float(tt.item())
    # dpotrf_ moves only the upper triangle, as 512-column trapezoids (csrc/staging.cu: stage_upload_tri)
    tri = sum(min(c0 + 512, n) * (min(c0 + 512, n) - c0) for c0 in range(0, n, 512)) * 8
    h2d = 2 * n * n * 8 + (n * n * 8 + n * nrhs * 8) + tri
    d2h = n * n * 8 + (n * n * 8 + n * nrhs * 8 + n * 4 + 4) + (tri + 4)
    res["e2e"] = {"value": world * sum(f.values()) * e2e_steps / e2e_total / 1e9, "unit": "GFLOP/s",
                  "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                  "ops": {k: f[k] * e2e_steps / te[k] / 1e9 for k in te}, "host_memory": "pinned (llacuda_host_alloc)"}
    for p in hbuf:
        lib.llacuda_host_free(p)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="llacuda", choices=["llacuda", "reference"])
    ap.add_argument("--n", type=int, default=N_DEFAULT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    config = {"workload": f"mm + solve(nrhs={NRHS}) + cholesky, double-float, n={n} (BASELINE.json metric shapes; "
                          "solve = configs[1] call at the metric size, cholesky = configs[2])",
              "n": n, "nrhs": NRHS, "parallelism": f"replicas x{world}" if world > 1 else "single GPU",
              "l2": "operands are 2 GiB each at n=16384, far larger than the 126 MB L2; no flush needed",
              "in_place_inputs": "restored between steps outside the event pairs"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        ncpu = args.cpu_n or min(n, pick_cpu_n(150.0, args.steps + args.warmup))   # a bounded sample, never above the workload
        r = cpu_leg(ncpu, args.steps, args.warmup)
        sample = (f"the same three calls (dgemm/dgesv/dpotrf as LLA issues them) at n={ncpu}, nrhs={NRHS}, "
                  f"{args.steps} timed steps after {args.warmup} warm-up; {r['blas']}")
        line = {"impl": "reference", "metric": "fp64_gflops_mm_solve_cholesky", "value": r["value"], "unit": "GFLOP/s",
                "n_gpus": 0, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "ops_gflops": r["ops"],
                "cpu_baseline": {"value": r["value"], "unit": "GFLOP/s", "cores": r["cores"], "kind": "port", "sample": sample},
                "e2e": {"value": r["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g = gpu_legs(args, rank, world, local_rank)
    f = flops(n)
    total_f = sum(f.values()) * args.steps * world
    value = total_f / (g["total_ms"] * 1e-3) / 1e9
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.profile:
        ncpu = args.cpu_n or min(n, pick_cpu_n(25.0, 2))
        r = cpu_leg(ncpu, 1, 1)
        cpu = {"value": r["value"], "unit": "GFLOP/s", "cores": r["cores"], "kind": "port",
               "sample": f"dgemm/dgesv/dpotrf as LLA issues them at n={ncpu}, nrhs={NRHS}, 1 timed pass after 1 warm-up; {r['blas']}",
               "ops_gflops": r["ops"]}
    if rank == 0:
        peak = g["dmma_tf"]
        ach = f["mm"] / (g["gemm_ms"] * 1e-3) / 1e12
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "dgemm_traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {"metric": "fp64_gflops_mm_solve_cholesky", "value": value, "unit": "GFLOP/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": g["total_ms"] / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "ops_gflops": {k: f[k] * args.steps / (g["t_ms"][k] * 1e-3) / 1e9 for k in f},
                "ops_frac_of_fp64_tensor_roof": {k: f[k] * args.steps / (g["t_ms"][k] * 1e-3) / 1e12 / peak for k in f},
                "parity": {"solve_scaled_residual": g["solve_resid"], "bound": 10, "info": g["info"]},
                "clocks": g["clocks"], "e2e": g["e2e"], "gpu_launches": g["launches"],
                "roofline": {"kernel": "lla_dgemm_kernel<128,64,16,...> (DMMA m8n8k4, two CTAs/SM), the mm launch of the step; algorithmic flop = 2 n^3",
                             "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                             "traffic": traffic,
                             "peak_source": "in-run llacuda_probe_fp64_peak DMMA loop; MEASURED_PEAKS.json has no FP64 figure "
                                            f"(DFMA loop: {g['dfma_tf']:.2f} TFLOP/s)"},
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
