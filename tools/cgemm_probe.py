"""CGEMM (planar split over the tcgen05 SGEMM) throughput, device-resident, n = 4096 and 8192."""
import sys, ctypes, torch
sys.path.insert(0, ".")
from lla_b200 import _lib, device as D
_lib.init(0); D.use_torch_stream()
L = _lib.lib()
f = L.llacuda_cgemm_dev
f.argtypes = [ctypes.c_char, ctypes.c_char, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_float),
              ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_float), ctypes.c_void_p, ctypes.c_int64]
one, zero = (ctypes.c_float * 2)(1.0, 0.0), (ctypes.c_float * 2)(0.0, 0.0)
for n in (4096, 8192):
    A = torch.complex(torch.rand((n, n), device="cuda"), torch.rand((n, n), device="cuda"))
    B = torch.complex(torch.rand((n, n), device="cuda"), torch.rand((n, n), device="cuda"))
    C = torch.zeros((n, n), dtype=torch.complex64, device="cuda")
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rc = f(b"N", b"N", n, n, n, one, A.data_ptr(), n, B.data_ptr(), n, zero, C.data_ptr(), n); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    assert rc == 0
    # column-major images: C_img = (A_mat B_mat)^T = B_img A_img
    ref = (B[:64].to(torch.complex128) @ A.to(torch.complex128))
    err = ((C[:64].to(torch.complex128) - ref).abs().max() / ref.abs().max()).item()
    print(f"cgemm n={n}: {best:.2f} ms  {8*n**3/best/1e9:.1f} real TFLOP/s  max rel err (64 columns) {err:.2e}", flush=True)
