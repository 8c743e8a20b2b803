import sys, torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D
_lib.init(0); D.use_torch_stream()
for n in (2048, 4096, 8192):
    for oa, ob in (("N", "N"), ("T", "N"), ("N", "T"), ("T", "T")):
        A = torch.rand((n, n), dtype=torch.float32, device="cuda") * 2 - 1; B = torch.rand((n, n), dtype=torch.float32, device="cuda") * 2 - 1
        C = torch.zeros((n, n), dtype=torch.float32, device="cuda")
        f = lambda: D.sgemm(oa, ob, n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, 0.0, C.data_ptr(), n)
        f(); torch.cuda.synchronize(); best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        opA = (A.t() if oa == "N" else A).double(); opB = (B.t() if ob == "N" else B).double()
        ref = opA[:64] @ opB
        rel = ((C.t()[:64].double() - ref).abs().max() / ref.abs().max()).item()
        print(f"sgemm {oa}{ob} n={n}: {best:.3f} ms {2.0*n**3/best/1e9:.1f} TFLOP/s  max rel err {rel:.2e}", flush=True)
