import sys, torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D
_lib.init(0); D.use_torch_stream()
n = 32768
G = torch.rand((n, n), dtype=torch.float64, device="cuda")
S0 = torch.empty_like(G)
D.dgemm("T", "N", n, n, n, 1.0 / n, G.data_ptr(), n, G.data_ptr(), n, 0.0, S0.data_ptr(), n)
S0.diagonal().add_(float(n)); del G
A = S0.clone(); info = torch.zeros(1, dtype=torch.int32, device="cuda")
best = 1e9
for _ in range(2):
    A.copy_(S0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); D.dpotrf("U", n, A.data_ptr(), n, info.data_ptr()); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"dpotrf 1 GPU n={n}: {best:.2f} ms {n**3/3/best/1e9:.2f} TF info {info.item()}", flush=True)
del S0
A0 = torch.rand((n, n), dtype=torch.float64, device="cuda") * 10
ipiv = torch.zeros(n, dtype=torch.int32, device="cuda")
best = 1e9
for _ in range(2):
    A.copy_(A0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); D.dgetrf(n, n, A.data_ptr(), n, ipiv.data_ptr(), info.data_ptr()); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"dgetrf 1 GPU n={n}: {best:.2f} ms {2*n**3/3/best/1e9:.2f} TF info {info.item()}", flush=True)
