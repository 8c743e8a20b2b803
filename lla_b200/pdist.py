"""Bindings of the multi-GPU engine in libllacuda.so (csrc/dist*.cu; include/llacuda.h group 4).

Two ways in, both running the same C++ rank functions:

* one process, all GPUs (what LLA's single `foreign-funcall` can reach):
  `init_all()` then `pdgemm / pdgesv / pdgetrf / pdpotrf / pdposv / pdpotrs / pzgemm / pdgesv_batched` on HOST
  (NumPy) arrays with the Fortran argument lists of the single-GPU drop-ins;
* one rank per process under a launcher (bench.py --gpus N, tests): `init_from_torch()` hands the NCCL unique id
  round with torch.distributed, then the `*_local` calls work on block-cyclic operands already in HBM.

Layout helpers (`BlockCyclic`) are pure Python and mirror llacuda_dist_numroc / _padded / _l2g.
PyTorch is plumbing here (device memory, process group); nothing in this file computes.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

_i64, _dbl, _vp, _int = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_int


def _f(name, argtypes, restype=ctypes.c_int):
    f = getattr(_lib.lib(), name)
    f.argtypes = argtypes
    f.restype = restype
    return f


# ------------------------------------------------------------------ engine life cycle
def init_all(ndev: int = 0, P: int = 0, Q: int = 0) -> tuple[int, int, int]:
    """Single-process mode over `ndev` GPUs (0 = all visible). Returns (nranks, P, Q)."""
    _lib.check(_f("llacuda_dist_init_all", [_int, _int, _int])(ndev, P, Q), "llacuda_dist_init_all")
    return info()[:3]


def init_from_torch(P: int = 0, Q: int = 0) -> tuple[int, int, int]:
    """Launcher mode: this process is rank torch.distributed.get_rank() on its current CUDA device."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        _lib.check(_f("llacuda_dist_unique_id", [ctypes.c_char_p])(buf), "llacuda_dist_unique_id")
    box = [bytes(buf.raw)]
    dist.broadcast_object_list(box, src=0)
    _lib.check(_f("llacuda_dist_init_rank", [_int, _int, ctypes.c_char_p, _int, _int])(world, rank, box[0], P, Q),
               "llacuda_dist_init_rank")
    torch.cuda.synchronize()
    return info()[:3]


def info():
    n, p, q, sp, r = (_int(0) for _ in range(5))
    rc = _f("llacuda_dist_info", [ctypes.POINTER(_int)] * 5)(n, p, q, sp, r)
    if rc != 0:
        return 0, 0, 0, 0, -1
    return n.value, p.value, q.value, sp.value, r.value


def set_grid(P: int, Q: int) -> None:
    _lib.check(_f("llacuda_dist_set_grid", [_int, _int])(P, Q), "llacuda_dist_set_grid")


def finalize() -> None:
    _lib.check(_f("llacuda_dist_finalize", [])(), "llacuda_dist_finalize")


def stream_ptr() -> int:
    return _f("llacuda_dist_stream", [], _vp)() or 0


# ------------------------------------------------------------------ layout arithmetic (pure)
def lcm(a, b):
    return a * b // np.gcd(a, b)


def padded(n, nb, P, Q):
    m = nb * lcm(P, Q)
    return (n + m - 1) // m * m


def numroc(n, nb, iproc, nprocs):
    """ScaLAPACK NUMROC with source process 0 (reference: none -- SURVEY.md section 2a; mirrors llacuda_dist_numroc)."""
    nblocks, extra = n // nb, (n // nb) % nprocs
    r = (nblocks // nprocs) * nb
    if iproc < extra:
        r += nb
    elif iproc == extra:
        r += n % nb
    return r


def default_grid(world):
    p = 1
    c = 1
    while c * c <= world:
        if world % c == 0:
            p = c
        c += 1
    return p, world // p


class BlockCyclic:
    """Ownership arithmetic of an Mp x Np matrix in nb x nb blocks on a P x Q grid: block (I, J) lives on rank
    (I mod P) Q + (J mod Q) as local block (I // P, J // Q); local storage column-major (Mp/P) x (Np/Q)."""

    def __init__(self, Mp, Np, nb, P, Q, rank):
        assert Mp % (nb * P) == 0 and Np % (nb * Q) == 0
        self.Mp, self.Np, self.nb, self.P, self.Q = Mp, Np, nb, P, Q
        self.p, self.q = rank // Q, rank % Q
        self.mloc, self.nloc = Mp // P, Np // Q

    def row_blocks(self):
        return range(self.p, self.Mp // self.nb, self.P)

    def col_blocks(self):
        return range(self.q, self.Np // self.nb, self.Q)

    def local_rows(self):
        """global row index of every local row"""
        return np.concatenate([np.arange(I * self.nb, (I + 1) * self.nb) for I in self.row_blocks()])

    def local_cols(self):
        return np.concatenate([np.arange(J * self.nb, (J + 1) * self.nb) for J in self.col_blocks()])

    def scatter(self, img):
        """img: (Np, Mp) array-like = column-major image of the global matrix (img[j, i] = A(i, j)) -> local image (nloc, mloc)."""
        rows, cols = self.local_rows(), self.local_cols()
        if not isinstance(img, np.ndarray):   # a torch tensor (NumPy 2 arrays have a .device too)
            import torch
            rows = torch.as_tensor(rows, device=img.device)
            cols = torch.as_tensor(cols, device=img.device)
            return img[cols][:, rows].contiguous()
        return np.ascontiguousarray(img[np.ix_(cols, rows)])


# ------------------------------------------------------------------ one rank per process, operands in HBM
def pdgemm_local(Mp, Np, Kp, nb, alpha, A, lda, B, ldb, beta, C, ldc):
    f = _f("llacuda_pdgemm_local", [_i64, _i64, _i64, _i64, _dbl, _vp, _i64, _vp, _i64, _dbl, _vp, _i64])
    _lib.check(f(Mp, Np, Kp, nb, alpha, A, lda, B, ldb, beta, C, ldc), "pdgemm_local")


def pzgemm_local(Mp, Np, Kp, nb, alpha, A, lda, B, ldb, beta, C, ldc):
    a2 = (_dbl * 2)(complex(alpha).real, complex(alpha).imag)
    b2 = (_dbl * 2)(complex(beta).real, complex(beta).imag)
    f = _f("llacuda_pzgemm_local", [_i64, _i64, _i64, _i64, ctypes.POINTER(_dbl), _vp, _i64, _vp, _i64, ctypes.POINTER(_dbl), _vp, _i64])
    _lib.check(f(Mp, Np, Kp, nb, a2, A, lda, B, ldb, b2, C, ldc), "pzgemm_local")


def pdpotrf_local(Np, nb, A, lld) -> int:
    inf = _int(0)
    _lib.check(_f("llacuda_pdpotrf_local", [_i64, _i64, _vp, _i64, ctypes.POINTER(_int)])(Np, nb, A, lld, inf), "pdpotrf_local")
    return inf.value


def pdgetrf_local(Np, nb, A, lld, ipiv_dev) -> int:
    inf = _int(0)
    _lib.check(_f("llacuda_pdgetrf_local", [_i64, _i64, _vp, _i64, _vp, ctypes.POINTER(_int)])(Np, nb, A, lld, ipiv_dev, inf),
               "pdgetrf_local")
    return inf.value


def pdgesv_local(Np, nb, A, lld, ipiv_dev, B, ldb, nrhs) -> int:
    """factorisation + solve on device-resident block-cyclic A and replicated B (X on return when the INFO returned is 0)"""
    inf = _int(0)
    f = _f("llacuda_pdgesv_local", [_i64, _i64, _vp, _i64, _vp, _vp, _i64, _i64, ctypes.POINTER(_int)])
    _lib.check(f(Np, nb, A, lld, ipiv_dev, B, ldb, nrhs, inf), "pdgesv_local")
    return inf.value


def pdgetrs_local(Np, nb, LU, lld, ipiv_dev, B, ldb, nrhs):
    f = _f("llacuda_pdgetrs_local", [_i64, _i64, _vp, _i64, _vp, _vp, _i64, _i64])
    _lib.check(f(Np, nb, LU, lld, ipiv_dev, B, ldb, nrhs), "pdgetrs_local")


def pdpotrs_local(Np, nb, U, lld, B, ldb, nrhs):
    f = _f("llacuda_pdpotrs_local", [_i64, _i64, _vp, _i64, _vp, _i64, _i64])
    _lib.check(f(Np, nb, U, lld, B, ldb, nrhs), "pdpotrs_local")


# ------------------------------------------------------------------ one process, host arrays (Fortran argument lists)
def _I(v):
    return ctypes.byref(ctypes.c_int(int(v)))


def _ch(s):
    return ctypes.byref(ctypes.c_char(s.encode()))


def _P(a):
    return a.ctypes.data_as(_vp)


def _raise_if_device_error():
    if _lib.lib().llacuda_last_error_code() != 0:
        msg = _lib.last_error()
        _lib.lib().llacuda_clear_error()
        raise _lib.LlacudaError(msg)


def pdgemm(transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
    """column-major operands in NumPy buffers; C overwritten in place"""
    f = _f("llacuda_pdgemm", None, None)
    f(_ch(transa), _ch(transb), _I(m), _I(n), _I(k), ctypes.byref(_dbl(alpha)), _P(a), _I(lda), _P(b), _I(ldb),
      ctypes.byref(_dbl(beta)), _P(c), _I(ldc))
    _raise_if_device_error()


def pzgemm(transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
    al = np.array([alpha], dtype=np.complex128)
    be = np.array([beta], dtype=np.complex128)
    f = _f("llacuda_pzgemm", None, None)
    f(_ch(transa), _ch(transb), _I(m), _I(n), _I(k), _P(al), _P(a), _I(lda), _P(b), _I(ldb), _P(be), _P(c), _I(ldc))
    _raise_if_device_error()


def pdpotrf(uplo, n, a, lda) -> int:
    inf = _int(0)
    _f("llacuda_pdpotrf", None, None)(_ch(uplo), _I(n), _P(a), _I(lda), ctypes.byref(inf))
    return inf.value


def pdposv(uplo, n, nrhs, a, lda, b, ldb) -> int:
    inf = _int(0)
    _f("llacuda_pdposv", None, None)(_ch(uplo), _I(n), _I(nrhs), _P(a), _I(lda), _P(b), _I(ldb), ctypes.byref(inf))
    return inf.value


def pdpotrs(uplo, n, nrhs, a, lda, b, ldb) -> int:
    inf = _int(0)
    _f("llacuda_pdpotrs", None, None)(_ch(uplo), _I(n), _I(nrhs), _P(a), _I(lda), _P(b), _I(ldb), ctypes.byref(inf))
    return inf.value


def pdgetrf(m, n, a, lda, ipiv) -> int:
    inf = _int(0)
    _f("llacuda_pdgetrf", None, None)(_I(m), _I(n), _P(a), _I(lda), _P(ipiv), ctypes.byref(inf))
    return inf.value


def pdgesv(n, nrhs, a, lda, ipiv, b, ldb) -> int:
    inf = _int(0)
    _f("llacuda_pdgesv", None, None)(_I(n), _I(nrhs), _P(a), _I(lda), _P(ipiv), _P(b), _I(ldb), ctypes.byref(inf))
    return inf.value


def pdgesv_batched(a, b, keep_lu=False):
    """a: (batch, n, n) row-major systems as LLA holds them; b: (batch, n) or (batch, n, nrhs).  Returns (x, ipiv, info[, lu])."""
    a = np.asarray(a, dtype=np.float64)
    batch, n, _ = a.shape
    b = np.asarray(b, dtype=np.float64)
    vec = b.ndim == 2
    bb = b[:, :, None] if vec else b
    nrhs = bb.shape[2]
    acm = np.array(np.transpose(a, (0, 2, 1)), order="C", copy=True)       # column-major image per system (never a view)
    bcm = np.array(np.transpose(bb, (0, 2, 1)), order="C", copy=True)
    ipiv = np.zeros((batch, n), dtype=np.int32)
    info_ = np.zeros(batch, dtype=np.int32)
    f = _f("llacuda_pdgesv_batched", [_int, _int, _i64, _vp, _vp, _vp, _vp, _int])
    _lib.check(f(n, nrhs, batch, _P(acm), _P(bcm), _P(ipiv), _P(info_), 1 if keep_lu else 0), "pdgesv_batched")
    x = np.transpose(bcm, (0, 2, 1))
    x = np.ascontiguousarray(x[:, :, 0] if vec else x)
    if keep_lu:
        return x, ipiv, info_, np.ascontiguousarray(np.transpose(acm, (0, 2, 1)))
    return x, ipiv, info_
