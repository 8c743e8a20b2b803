"""Does a power-of-two leading dimension cost the trailing-update GEMMs L2 capacity?  ld = 16384 vs padded."""
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
n = 16384
for ld in (16384, 16384 + 16, 16384 + 136):
    A = torch.rand((n, ld), dtype=torch.float64, device="cuda")
    C = torch.rand((n, ld), dtype=torch.float64, device="cuda")
    for (m_, n_, k) in ((15872, 15360, 512), (14336, 14336, 1024), (16384, 16384, 16384)):
        ms = t(lambda: D.dgemm("N", "N", m_, n_, k, -1.0, A.data_ptr(), ld, A.data_ptr(), ld, 1.0, C.data_ptr(), ld, 0), reps=2 if k > 2048 else 3)
        ms2 = t(lambda: D.dgemm("T", "N", n_, n_, k, -1.0, A.data_ptr(), ld, A.data_ptr(), ld, 1.0, C.data_ptr(), ld, 1), reps=2 if k > 2048 else 3)
        print(f"ld={ld} m={m_} n={n_} k={k}: NN {2*m_*n_*k/ms/1e9:.1f} TF   TN-upper {n_*n_*k/ms2/1e9:.1f} TF", flush=True)
    del A, C
