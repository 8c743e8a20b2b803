"""HBM roofline of the layout kernels: llacuda_dtranspose_dev and llacuda_convert_dev (transposing and straight) at n x n,
CUDA events, L2 flushed between launches; algorithmic bytes = rows cols (sizeof in + sizeof out)."""
import json, os, sys, statistics
import numpy as np
import torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D
_lib.init(0); D.use_torch_stream()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
peak = 6539.5
try:
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(fn, nbytes, label):
    ts = []
    for it in range(6):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if it >= 2: ts.append(e0.elapsed_time(e1))
    ms = statistics.mean(ts)
    print(f"{label:44s} n={n}: {ms:7.3f} ms  {nbytes/ms/1e6:7.1f} GB/s  = {nbytes/ms/1e6/peak:5.1%} of the measured {peak:.0f} GB/s HBM roof", flush=True)
a = torch.rand((n, n), dtype=torch.float64, device="cuda")
b = torch.empty_like(a)
run(lambda: D.dtranspose(n, n, a.data_ptr(), n, b.data_ptr(), n), 2 * n * n * 8, "dtranspose (f64 -> f64, transposed)")
assert torch.equal(b, a.t())
run(lambda: D.convert(np.dtype("float64"), np.dtype("float64"), True, n, n, a.data_ptr(), n, b.data_ptr(), n), 2 * n * n * 8, "convert f64 -> f64, transposed")
run(lambda: D.convert(np.dtype("float64"), np.dtype("float64"), False, n, n, a.data_ptr(), n, b.data_ptr(), n), 2 * n * n * 8, "convert f64 -> f64, straight")
s = torch.rand((n, n), dtype=torch.float32, device="cuda")
run(lambda: D.convert(np.dtype("float32"), np.dtype("float64"), True, n, n, s.data_ptr(), n, b.data_ptr(), n), n * n * 12, "convert f32 -> f64, transposed")
assert torch.equal(b, s.double().t())
i = torch.randint(-1000, 1000, (n, n), dtype=torch.int64, device="cuda")
run(lambda: D.convert(np.dtype("int64"), np.dtype("float64"), True, n, n, i.data_ptr(), n, b.data_ptr(), n), 2 * n * n * 8, "convert i64 -> f64, transposed")
z = torch.empty((n // 2, n), dtype=torch.complex128, device="cuda")
c = torch.rand((n // 2, n), dtype=torch.float64, device="cuda")
run(lambda: D.convert(np.dtype("float64"), np.dtype("complex128"), True, n, n // 2, c.data_ptr(), n, z.data_ptr(), n // 2), n * (n // 2) * 24, "convert f64 -> c128, transposed (n x n/2)")
