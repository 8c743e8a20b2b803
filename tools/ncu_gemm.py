"""Short command for the ncu --set full capture of the dominant kernel (DMMA DGEMM)."""
import sys
import torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D
_lib.init(0); D.use_torch_stream()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
A = torch.rand((n, n), dtype=torch.float64, device="cuda")
B = torch.rand((n, n), dtype=torch.float64, device="cuda")
C = torch.zeros((n, n), dtype=torch.float64, device="cuda")
for _ in range(4):
    D.dgemm("N", "N", n, n, n, 1.0, D.colmajor(A)[0], n, D.colmajor(B)[0], n, 0.0, D.colmajor(C)[0], n)
torch.cuda.synchronize()
print("ok")
