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
n, k = 14336, 1024
for (oa, ob) in (("T", "N"), ("N", "T"), ("N", "N")):
    for tri in (0, 1, 2):
        ms = t(lambda: D.dgemm(oa, ob, n, n, k, -1.0, A.data_ptr(), ld, A.data_ptr(), ld, 1.0, C.data_ptr(), ld, tri))
        fl = (2 if tri == 0 else 1) * n * n * k
        print(f"{oa}{ob} tri={tri} n={n} k={k}: {ms:.3f} ms {fl / ms / 1e9:.1f} TF", flush=True)
