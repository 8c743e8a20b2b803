"""dgesv on device pointers (factorisation + solve, 64 right-hand sides): time and scaled residual."""
import sys, os, torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D
_lib.init(0); D.use_torch_stream()
sizes = [int(v) for v in sys.argv[1:]] or [8192, 16384]
for n in sizes:
    nrhs = 64 if n <= 32768 else 2
    A0 = torch.rand((n, n), dtype=torch.float64, device="cuda") * 10      # column-major image: A0[j, i] = a_ij
    B0 = torch.rand((nrhs, n), dtype=torch.float64, device="cuda") * 10
    A, B = A0.clone(), B0.clone()
    ipiv = torch.zeros(n, dtype=torch.int32, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    best = 1e9
    for _ in range(3 if n <= 32768 else 1):
        A.copy_(A0); B.copy_(B0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); D.dgesv(n, nrhs, A.data_ptr(), n, ipiv.data_ptr(), B.data_ptr(), n, info.data_ptr()); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    # images: X^T (nrhs x n) times A^T (n x n) = B^T
    R = B @ A0 - B0
    res = (R.norm() / (n * 2.2e-16 * A0.norm() * B.norm())).item()
    fl = 2 * n**3 / 3 + 2 * n * n * nrhs
    print(f"rider={'off' if os.environ.get('LLACUDA_LU_NO_RHS_RIDER') else 'on '} dgesv n={n}: {best:.2f} ms {fl/best/1e9:.2f} TF  residual {res:.4f} info {info.item()}", flush=True)
    del A, B, A0, B0
