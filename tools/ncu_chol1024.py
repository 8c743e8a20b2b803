import sys, torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D
_lib.init(0); D.use_torch_stream()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
which = sys.argv[2] if len(sys.argv) > 2 else "chol"
info = torch.zeros(1, dtype=torch.int32, device="cuda")
if which == "chol":
    G = torch.rand((n, n), dtype=torch.float64, device="cuda")
    A = G @ G.t() + n * torch.eye(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    D.dpotrf("U", n, A.data_ptr(), n, info.data_ptr())
else:
    A = torch.rand((n, n), dtype=torch.float64, device="cuda") * 10
    ipiv = torch.zeros(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    D.dgetrf(n, n, A.data_ptr(), n, ipiv.data_ptr(), info.data_ptr())
torch.cuda.synchronize()
