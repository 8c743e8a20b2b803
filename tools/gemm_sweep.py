import sys, os, torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D
_lib.init(0); D.use_torch_stream()
def t(m, n, k, oa="N", ob="N", reps=3):
    ar, ac = (k, m) if oa != "N" else (m, k)
    br, bc = (n, k) if ob != "N" else (k, n)
    A = torch.rand((ac, ar), dtype=torch.float64, device="cuda"); B = torch.rand((bc, br), dtype=torch.float64, device="cuda")
    C = torch.zeros((n, m), dtype=torch.float64, device="cuda")
    f = lambda: D.dgemm(oa, ob, m, n, k, 1.0, D.colmajor(A)[0], ar, D.colmajor(B)[0], br, 0.0, D.colmajor(C)[0], m)
    f(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    ref = (A.t() if oa == "N" else A)[:32] @ (B.t() if ob == "N" else B)
    err = (C.t()[:32] - ref).abs().max().item()
    print(f"cfg={os.environ.get('LLACUDA_GEMM_CFG','0')} {oa}{ob} {m}x{n}x{k}: {best:.3f} ms {2.0*m*n*k/best/1e9:.2f} TF err={err:.2e}", flush=True)
t(8192, 8192, 8192); t(16384, 16384, 16384, reps=2); t(16384, 16384, 256); t(8192, 8192, 8192, "T", "N"); t(8192, 8192, 8192, "N", "T")
