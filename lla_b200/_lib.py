"""ctypes loader for libllacuda.so (built in-tree by __graft_entry__.build / csrc/Makefile).

There is deliberately no fallback: if the shared library is missing, or no sm_100 device is
present, every compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libllacuda.so")
SO64_PATH = os.path.join(_HERE, "libllacuda64.so")   # Fortran symbols with 64-bit integers (LLA compiled with :int64)
CSRC = os.path.join(_HERE, "csrc")

_lib = None


class LlacudaError(RuntimeError):
    """A CUDA-side failure reported through llacuda_last_error() (never a LAPACK INFO code)."""


def build(force: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into lla_b200/libllacuda.so and libllacuda64.so (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", CSRC, "clean", "-s"])
    subprocess.check_call(["make", "-C", CSRC, "-j", str(min(16, os.cpu_count() or 4)), "-s"])
    return SO_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise LlacudaError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(llacuda has no CPU fallback)")
        _lib = ctypes.CDLL(SO_PATH)
        _lib.llacuda_last_error.restype = ctypes.c_char_p
        _lib.llacuda_stream.restype = ctypes.c_void_p
        # BLAS-style entry points: poison the output, record the error and return (the mirror polls
        # llacuda_last_error_code after every call and raises) instead of the XERBLA-style stop
        _lib.llacuda_set_error_mode(1)
    return _lib


def last_error() -> str:
    return lib().llacuda_last_error().decode()


def check(rc: int, what: str = "llacuda") -> None:
    if rc != 0:
        raise LlacudaError(f"{what} failed with code {rc}: {last_error()}")


def init(device: int | None = None) -> None:
    if device is not None:
        check(lib().llacuda_set_device(int(device)), "llacuda_set_device")
    check(lib().llacuda_init(), "llacuda_init")
