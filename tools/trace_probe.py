"""Timeline of the blocked Cholesky / LU drivers (LLACUDA_TRACE=1): one traced call per routine at n."""
import sys, os, torch
sys.path.insert(0, ".")
os.environ["LLACUDA_TRACE"] = "1"
from lla_b200 import _lib, device as D
_lib.init(0); D.use_torch_stream()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
which = sys.argv[1] if len(sys.argv) > 1 else "both"
L = _lib.lib()
def timed(f):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); f(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
if which in ("chol", "both"):
    G = torch.rand((n, n), dtype=torch.float64, device="cuda")
    S0 = G @ G.t() + n * torch.eye(n, dtype=torch.float64, device="cuda"); del G
    A = S0.clone(); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    for it in range(2):
        A.copy_(S0); L.llacuda_trace_enable(it == 1)
        ms = timed(lambda: D.dpotrf("U", n, A.data_ptr(), n, info.data_ptr()))
        print(f"dpotrf n={n}: {ms:.2f} ms {n**3/3/ms/1e9:.2f} TF", file=sys.stderr, flush=True)
    L.llacuda_trace_dump(); del A, S0
if which in ("lu", "both"):
    A0 = torch.rand((n, n), dtype=torch.float64, device="cuda") * 10
    A = A0.clone(); ipiv = torch.zeros(n, dtype=torch.int32, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    for it in range(2):
        A.copy_(A0); L.llacuda_trace_enable(it == 1)
        ms = timed(lambda: D.dgetrf(n, n, A.data_ptr(), n, ipiv.data_ptr(), info.data_ptr()))
        print(f"dgetrf n={n}: {ms:.2f} ms {2*n**3/3/ms/1e9:.2f} TF", file=sys.stderr, flush=True)
    L.llacuda_trace_dump()
