"""Small factorisations for cheap ncu launch lists: python tools/ncu_small.py [lu|chol] n"""
import sys, torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D
_lib.init(0); D.use_torch_stream()
which, n = sys.argv[1], int(sys.argv[2])
A = torch.rand((n, n), dtype=torch.float64, device="cuda")
if which == "chol":
    A = A @ A.t() + n * torch.eye(n, dtype=torch.float64, device="cuda")
W = A.clone()
ipiv = torch.zeros(n, dtype=torch.int32, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
for _ in range(2):
    W.copy_(A)
    if which == "lu": D.dgetrf(n, n, D.colmajor(W)[0], n, ipiv.data_ptr(), info.data_ptr())
    else: D.dpotrf("U", n, D.colmajor(W)[0], n, info.data_ptr())
torch.cuda.synchronize(); print("ok", info.item())
