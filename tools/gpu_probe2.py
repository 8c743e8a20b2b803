"""Scratch GPU probe: device-resident DGEMM / DGETRF / DGESV / DPOTRF throughput + residuals."""
import json, sys, os
import torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D

_lib.init()
D.use_torch_stream()
EPS = 2.220446049250313e-16
out = {}

def timeit(fn, reps=3, setup=None):
    best = 1e30
    for _ in range(reps):
        if setup: setup()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

def bench_gemm(m, n, k, opa="N", opb="N"):
    ar, ac = (k, m) if opa != "N" else (m, k)
    br, bc = (n, k) if opb != "N" else (k, n)
    A = torch.rand((ac, ar), dtype=torch.float64, device="cuda")
    B = torch.rand((bc, br), dtype=torch.float64, device="cuda")
    C = torch.zeros((n, m), dtype=torch.float64, device="cuda")
    pa, lda = D.colmajor(A); pb, ldb = D.colmajor(B); pc, ldc = D.colmajor(C)
    f = lambda: D.dgemm(opa, opb, m, n, k, 1.0, pa, lda, pb, ldb, 0.0, pc, ldc)
    f()
    ms = timeit(f)
    tf = 2.0 * m * n * k / (ms * 1e-3) / 1e12
    print(f"dgemm {opa}{opb} m={m} n={n} k={k}: {ms:.3f} ms {tf:.2f} TFLOP/s", flush=True)
    return tf

for n in (1024, 2048, 4096, 8192, 16384):
    out[f"dgemm_nn_{n}"] = bench_gemm(n, n, n)
for oa, ob in (("T", "N"), ("N", "T"), ("T", "T")):
    out[f"dgemm_{oa}{ob}_8192"] = bench_gemm(8192, 8192, 8192, oa, ob)
for k in (32, 64, 128, 256, 512, 1024):
    out[f"dgemm_trail_k{k}"] = bench_gemm(16384, 16384, k)
out["dgemm_tn_trail_k256"] = bench_gemm(16384, 16384, 256, "T", "N")

def bench_lu(n, nrhs=64):
    A0 = torch.rand((n, n), dtype=torch.float64, device="cuda") * 10
    B0 = torch.rand((nrhs, n), dtype=torch.float64, device="cuda") * 10
    A = A0.clone(); B = B0.clone()
    ipiv = torch.zeros(n, dtype=torch.int32, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    pa, lda = D.colmajor(A); pb, ldb = D.colmajor(B)
    def setup():
        A.copy_(A0); B.copy_(B0)
    f = lambda: D.dgetrf(n, n, pa, lda, ipiv.data_ptr(), info.data_ptr())
    setup(); f()
    ms = timeit(f, setup=setup)
    tf = (2.0 / 3.0) * n ** 3 / (ms * 1e-3) / 1e12
    print(f"dgetrf n={n}: {ms:.3f} ms {tf:.2f} TFLOP/s info={info.item()}", flush=True)
    g = lambda: D.dgesv(n, nrhs, pa, lda, ipiv.data_ptr(), pb, ldb, info.data_ptr())
    ms2 = timeit(g, setup=setup)
    tf2 = ((2.0 / 3.0) * n ** 3 + 2.0 * n * n * nrhs) / (ms2 * 1e-3) / 1e12
    # residual: column-major A0 is A0.t() as a math matrix
    Am = A0.t(); X = B.t(); Bm = B0.t()
    res = (torch.linalg.norm(Am @ X - Bm) / (n * EPS * torch.linalg.norm(Am) * torch.linalg.norm(X))).item()
    print(f"dgesv n={n} nrhs={nrhs}: {ms2:.3f} ms {tf2:.2f} TFLOP/s resid={res:.3f} info={info.item()}", flush=True)
    return tf, tf2, res

def bench_chol(n, nrhs=64):
    G = torch.rand((n, n), dtype=torch.float64, device="cuda")
    S0 = G @ G.t() + n * torch.eye(n, dtype=torch.float64, device="cuda")
    del G
    B0 = torch.rand((nrhs, n), dtype=torch.float64, device="cuda")
    A = S0.clone(); B = B0.clone()
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    pa, lda = D.colmajor(A); pb, ldb = D.colmajor(B)
    def setup():
        A.copy_(S0); B.copy_(B0)
    f = lambda: D.dpotrf("U", n, pa, lda, info.data_ptr())
    setup(); f()
    ms = timeit(f, setup=setup)
    tf = (1.0 / 3.0) * n ** 3 / (ms * 1e-3) / 1e12
    g = lambda: D.dpotrs("U", n, nrhs, pa, lda, pb, ldb)
    ms2 = timeit(g, reps=1, setup=lambda: B.copy_(B0))
    X = B.t(); Bm = B0.t()
    res = (torch.linalg.norm(S0 @ X - Bm) / (n * EPS * torch.linalg.norm(S0) * torch.linalg.norm(X))).item()
    print(f"dpotrf n={n}: {ms:.3f} ms {tf:.2f} TFLOP/s info={info.item()} | dpotrs nrhs={nrhs}: {ms2:.3f} ms resid={res:.3f}", flush=True)
    return tf, res

for n in (2048, 4096, 8192, 16384):
    out[f"lu_{n}"] = bench_lu(n)
for n in (2048, 4096, 8192, 16384):
    out[f"chol_{n}"] = bench_chol(n)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe2.json", "w"), indent=1)
