"""configs[3]: 200 000 independent 32 x 32 double systems on one GPU (device-resident), GB/s against the HBM roof."""
import sys, torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D
_lib.init(0); D.use_torch_stream()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = 32
g = torch.Generator(device="cuda"); g.manual_seed(4)
A0 = torch.rand((batch, n, n), dtype=torch.float64, device="cuda", generator=g) * 10
B0 = torch.rand((batch, n), dtype=torch.float64, device="cuda", generator=g) * 10
A, B = A0.clone(), B0.clone()
info = torch.zeros(batch, dtype=torch.int32, device="cuda")
best = 1e9
for _ in range(reps):
    B.copy_(B0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); D.dgesv_batched(n, 1, batch, A.data_ptr(), n * n, None, B.data_ptr(), n, info.data_ptr(), 0); e1.record()
    torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
# residual: A is column-major per system: A[s] (as stored) = M^T
X = B
R = torch.einsum("sji,sj->si", A0, X) - B0
res = (R.norm(dim=1) / (n * 2.2e-16 * A0.flatten(1).norm(dim=1) * X.norm(dim=1))).max().item()
byts = batch * 8704
print(f"batched dgesv {batch} x 32x32: {best:.3f} ms  {byts/best/1e6:.0f} GB/s algorithmic  {batch*23893/best/1e9:.2f} TFLOP/s  max scaled resid {res:.3f}  info!=0: {(info != 0).sum().item()}", flush=True)
