"""lla_b200 -- B200-native back end for the dense linear-algebra hot path of tpapp/lla.

`lla_b200.lla` mirrors LLA's user API for the path (mm, gemm!, lu, solve, cholesky, ipiv ...)
on NumPy row-major arrays and drives libllacuda.so through the same Fortran-ABI calls LLA issues.
"""
from . import _lib  # noqa: F401
