"""Short command for the ncu --set full capture of the tcgen05 SGEMM kernel."""
import sys
import torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D
_lib.init(0); D.use_torch_stream()
n = 4096
A = torch.rand((n, n), dtype=torch.float32, device="cuda"); B = torch.rand((n, n), dtype=torch.float32, device="cuda")
C = torch.zeros((n, n), dtype=torch.float32, device="cuda")
for _ in range(4):
    D.sgemm("N", "N", n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, 0.0, C.data_ptr(), n)
torch.cuda.synchronize(); print("ok")
