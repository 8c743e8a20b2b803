"""Standalone timing of the factorisation leaves: dpotrf n=128 (one CTA), dpotrf n=1024 (a panel), dgetrf 1024."""
import sys, torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D
_lib.init(0); D.use_torch_stream()
def bench(f, reps=64):
    for it in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): f()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
info = torch.zeros(1, dtype=torch.int32, device="cuda")
for n in (128, 1024, 4096):
    G = torch.rand((n, n), dtype=torch.float64, device="cuda")
    S0 = G @ G.t() + n * torch.eye(n, dtype=torch.float64, device="cuda")
    A = S0.clone()
    def f():
        A.copy_(S0); D.dpotrf("U", n, A.data_ptr(), n, info.data_ptr())
    def g():
        A.copy_(S0)
    t = bench(f) - bench(g)
    U = torch.triu(A.t())   # A is the column-major image: A.t() is the matrix
    err = (U.t() @ U - S0).abs().max().item() / S0.abs().max().item()
    print(f"dpotrf n={n}: {t:.1f} us  ({n**3/3/t/1e6:.2f} TF)  recon err {err:.2e} info {info.item()}", flush=True)
