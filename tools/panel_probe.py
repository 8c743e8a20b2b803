"""Standalone timing of the LU leaf (tall m x 32 dgetrf = one cooperative panel launch + nothing else)."""
import sys, os, torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D
_lib.init(0); D.use_torch_stream()
w = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for m in (128, 2048, 6144, 16384):
    A0 = torch.rand((w, m), dtype=torch.float64, device="cuda") * 10   # column-major m x w
    A = A0.clone(); ipiv = torch.zeros(w, dtype=torch.int32, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    reps = 128
    for it in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            D.dgetrf(m, w, A.data_ptr(), m, ipiv.data_ptr(), info.data_ptr())
        e1.record(); torch.cuda.synchronize()
    print(f"panel m={m} w={w}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per call", flush=True)
