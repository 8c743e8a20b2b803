"""Scratch GPU probe (run under gpurun): FP64 roofs, DGEMM device-resident throughput."""
import json, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D

_lib.init()
D.use_torch_stream()
out = {}
for kind, name in ((0, "dmma"), (1, "dfma")):
    tf, ms = D.probe_fp64_peak(kind, 20000)
    out[f"peak_{name}_tflops"] = tf
    print(name, tf, "TFLOP/s", ms, "ms", flush=True)

def bench_gemm(n, opa="N", opb="N", reps=3, m=None, k=None):
    m = m or n; k = k or n
    g = torch.Generator(device="cuda").manual_seed(0)
    ar, ac = (k, m) if opa != "N" else (m, k)
    br, bc = (n, k) if opb != "N" else (k, n)
    A = torch.rand((ac, ar), dtype=torch.float64, device="cuda", generator=g)
    B = torch.rand((bc, br), dtype=torch.float64, device="cuda", generator=g)
    C = torch.zeros((n, m), dtype=torch.float64, device="cuda")
    pa, lda = D.colmajor(A); pb, ldb = D.colmajor(B); pc, ldc = D.colmajor(C)
    D.dgemm(opa, opb, m, n, k, 1.0, pa, lda, pb, ldb, 0.0, pc, ldc)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        D.dgemm(opa, opb, m, n, k, 1.0, pa, lda, pb, ldb, 0.0, pc, ldc)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    tf = 2.0 * m * n * k / (best * 1e-3) / 1e12
    # check against torch (cuBLAS) on a sample
    At = A.t() if opa == "N" else A
    Bt = B.t() if opb == "N" else B
    ref = (At[:64] @ Bt)  # rows 0..63 of op(A) op(B)
    err = (C.t()[:64] - ref).abs().max().item() / (k * 2.2e-16)
    print(f"dgemm {opa}{opb} m={m} n={n} k={k}: {best:.3f} ms {tf:.2f} TFLOP/s err/(k eps)={err:.3f}", flush=True)
    return tf

for n in (1024, 2048, 4096, 8192, 16384):
    out[f"dgemm_nn_{n}"] = bench_gemm(n)
for oa, ob in (("T", "N"), ("N", "T"), ("T", "T")):
    out[f"dgemm_{oa}{ob}_8192"] = bench_gemm(8192, oa, ob)
out["dgemm_trail_k256"] = bench_gemm(16384, "N", "N", k=256)
out["dgemm_trail_k512"] = bench_gemm(16384, "N", "N", k=512)
out["dgemm_trail_tn_k256"] = bench_gemm(16384, "T", "N", k=256)
import os
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
