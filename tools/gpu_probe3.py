"""Scratch GPU probe: ZGEMM throughput and the batched 32x32 solver (configs[3] shape)."""
import sys, json, os
import torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D
_lib.init(0); D.use_torch_stream()
def timeit(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for n in (2048, 4096, 8192):
    for oa, ob in (("N", "N"), ("C", "N"), ("N", "C")):
        A = torch.randn((n, n), dtype=torch.complex128, device="cuda"); B = torch.randn((n, n), dtype=torch.complex128, device="cuda")
        C = torch.zeros((n, n), dtype=torch.complex128, device="cuda")
        f = lambda: D.zgemm(oa, ob, n, n, n, 1.0 + 0j, A.data_ptr(), n, B.data_ptr(), n, 0j, C.data_ptr(), n)
        f(); ms = timeit(f)
        opA = A.t() if oa == "N" else A.conj()       # column-major image: math matrix = tensor.t()
        opB = B.t() if ob == "N" else B.conj()
        ref = opA[:32] @ opB
        err = (C.t()[:32] - ref).abs().max().item()
        print(f"zgemm {oa}{ob} n={n}: {ms:.3f} ms {8.0*n**3/ms/1e9:.2f} TFLOP/s (real flops) err={err:.2e}", flush=True)
        if n == 8192 and (oa, ob) != ("N", "N"): break
batch, n = 200000, 32
A0 = torch.rand((batch, n, n), dtype=torch.float64, device="cuda") * 10
B0 = torch.rand((batch, n), dtype=torch.float64, device="cuda") * 10
A = A0.clone(); B = B0.clone()
ipiv = torch.zeros((batch, n), dtype=torch.int32, device="cuda"); info = torch.zeros(batch, dtype=torch.int32, device="cuda")
for keep in (0, 1):
    def run():
        D.dgesv_batched(n, 1, batch, A.data_ptr(), n * n, ipiv.data_ptr() if keep else None, B.data_ptr(), n, info.data_ptr(), keep)
    best = 1e30
    for _ in range(4):
        A.copy_(A0); B.copy_(B0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    byts = batch * (n * n * 8 + n * 8 * 2 + 4) + (batch * (n * n * 8 + n * 4) if keep else 0)
    x = B.unsqueeze(2); Am = A0.transpose(1, 2)
    res = (torch.linalg.norm(Am @ x - B0.unsqueeze(2), dim=(1, 2)) / (n * 2.22e-16 * torch.linalg.norm(Am, dim=(1, 2)) * torch.linalg.norm(x, dim=(1, 2)))).max().item()
    print(f"batched dgesv 200k x 32x32 keep_lu={keep}: {best:.3f} ms  {byts/best/1e6:.0f} GB/s algorithmic  {batch*23893/best/1e9:.2f} TFLOP/s  max resid {res:.3f} info!=0: {(info != 0).sum().item()}", flush=True)
