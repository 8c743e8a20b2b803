import sys, os, torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D
_lib.init(0); D.use_torch_stream()
for n in (8192, 16384):
    A0 = torch.rand((n, n), dtype=torch.float64, device="cuda") * 10
    A = A0.clone(); ipiv = torch.zeros(n, dtype=torch.int32, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    best = 1e9
    for _ in range(3):
        A.copy_(A0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); D.dgetrf(n, n, A.data_ptr(), n, ipiv.data_ptr(), info.data_ptr()); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"LU_NB={os.environ.get('LLACUDA_LU_NB')}: dgetrf n={n}: {best:.2f} ms {2*n**3/3/best/1e9:.2f} TF", flush=True)
    del A, A0
