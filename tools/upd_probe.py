"""Standalone rates of the trailing-update GEMM shapes (device-resident)."""
import sys, torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D
_lib.init(0); D.use_torch_stream()
def t(f, reps=3):
    best = 1e9
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ld = 16384
A = torch.rand((ld, ld), dtype=torch.float64, device="cuda")
C = torch.rand((ld, ld), dtype=torch.float64, device="cuda")
for n in (14336, 8192, 4096):
    for k in (256, 512, 1024):
        # TN tri-upper (Cholesky 'U'): C(n x n) -= P^T P, P is k x n (ld)
        ms = t(lambda: D.dgemm("T", "N", n, n, k, -1.0, A.data_ptr(), ld, A.data_ptr(), ld, 1.0, C.data_ptr(), ld, 1))
        tn = n * n * k / ms / 1e9
        # NT tri-lower (Cholesky 'L'): C -= P P^T, P is n x k
        ms = t(lambda: D.dgemm("N", "T", n, n, k, -1.0, A.data_ptr(), ld, A.data_ptr(), ld, 1.0, C.data_ptr(), ld, 2))
        nt = n * n * k / ms / 1e9
        # NN full (LU): C(n x n) -= L21 (n x k) U12 (k x n)
        ms = t(lambda: D.dgemm("N", "N", n, n, k, -1.0, A.data_ptr(), ld, A.data_ptr(), ld, 1.0, C.data_ptr(), ld, 0))
        nn = 2 * n * n * k / ms / 1e9
        print(f"n={n} k={k}: TN-upper {tn:.1f} TF  NT-lower {nt:.1f} TF  NN-full {nn:.1f} TF", flush=True)
