"""Pinned vs pageable host buffers through the Fortran drop-ins (dgemm_, dgesv_, dpotrf_): wall time and identical results.
LLA passes pageable memory (reference src/pinned-array.lisp:124-141); the pinned staging ring of csrc/hostpipe.cu is what
keeps that path at PCIe speed.  Usage: python tools/pageable_probe.py [n]"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lla_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nrhs = 64
_lib.init(0)
lib = _lib.lib()
vp = ctypes.c_void_p
I = lambda v: ctypes.byref(ctypes.c_int(int(v)))
ch = lambda s: ctypes.byref(ctypes.c_char(s.encode()))
one, zero = ctypes.byref(ctypes.c_double(1.0)), ctypes.byref(ctypes.c_double(0.0))
P = lambda x: x.ctypes.data_as(vp)
for name in ("dgemm_", "dgesv_", "dpotrf_"):
    getattr(lib, name).restype = None


def pinned(rows, cols):
    p = vp()
    _lib.check(lib.llacuda_host_alloc(ctypes.byref(p), ctypes.c_size_t(rows * cols * 8)), "host_alloc")
    return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_double)), shape=(rows, cols))


rng = np.random.default_rng(0)
a0 = rng.uniform(0, 10, (n, n))
b0 = rng.uniform(0, 10, (n, n))
r0 = rng.uniform(0, 10, (nrhs, n))
spd = None
res = {}
for kind in ("pinned", "pageable", "pageable-driver"):
    if kind == "pageable-driver":
        os.environ["LLACUDA_NO_PIPE"] = "1"   # read at load time only: informational, stays as the ring unless reloaded
    alloc = pinned if kind == "pinned" else (lambda r, c: np.empty((r, c)))
    A, B, C, X = alloc(n, n), alloc(n, n), alloc(n, n), alloc(nrhs, n)
    A[...] = a0; B[...] = b0
    ipiv = np.zeros(n, dtype=np.int32)
    info = ctypes.c_int(0)
    out = {}
    for it in range(3):
        t0 = time.perf_counter()
        lib.dgemm_(ch("N"), ch("N"), I(n), I(n), I(n), one, P(B), I(n), P(A), I(n), zero, P(C), I(n))
        t1 = time.perf_counter()
        if it == 0:
            out["c"] = C[:8].copy()
            if spd is None:
                # SPD = A^T A / 100 + n I through the same entry point
                lib.dgemm_(ch("T"), ch("N"), I(n), I(n), I(n), ctypes.byref(ctypes.c_double(0.01)), P(A), I(n), P(A), I(n), zero, P(C), I(n))
                spd = np.array(C)
                spd[np.diag_indices(n)] += n
        np.copyto(C, a0); np.copyto(X, r0)
        t2 = time.perf_counter()
        lib.dgesv_(I(n), I(nrhs), P(C), I(n), P(ipiv), P(X), I(n), ctypes.byref(info))
        t3 = time.perf_counter()
        assert info.value == 0
        if it == 0:
            out["lu"] = C[:8].copy(); out["x"] = X.copy(); out["ipiv"] = ipiv.copy()
        np.copyto(C, spd)
        t4 = time.perf_counter()
        lib.dpotrf_(ch("U"), I(n), P(C), I(n), ctypes.byref(info))
        t5 = time.perf_counter()
        assert info.value == 0
        if it == 0:
            out["chol"] = C[-8:].copy()
        tm = (t1 - t0, t3 - t2, t5 - t4)
    f = (2.0 * n ** 3, 2.0 / 3 * n ** 3 + 2.0 * n * n * nrhs, n ** 3 / 3.0)
    print(f"{kind:16s} n={n}: dgemm_ {tm[0]*1e3:8.1f} ms ({f[0]/tm[0]/1e12:5.2f} TF)  dgesv_ {tm[1]*1e3:8.1f} ms ({f[1]/tm[1]/1e12:5.2f} TF)  "
          f"dpotrf_ {tm[2]*1e3:8.1f} ms ({f[2]/tm[2]/1e12:5.2f} TF)  step {sum(f)/sum(tm)/1e12:5.2f} TFLOP/s", flush=True)
    res[kind] = out
    if kind == "pageable":
        break
ok = all(np.array_equal(res["pinned"][k], res["pageable"][k]) for k in res["pinned"])
print("pinned and pageable results identical:", ok)
sys.exit(0 if ok else 1)
