"""CPU-only checks of the host-side mirror (lla_b200/lla.py): everything LLA decides BEFORE the foreign call --
type lattice, dimension rules, wrapper semantics, ipiv post-processing -- behaves like the reference and like the
oracle's restatement.  No compute entry point is reached (those need a GPU and fail loudly without one)."""
import numpy as np
import pytest

from lla_b200 import lla as L
from oracle import lla_oracle as O


def test_type_lattice_matches_reference_and_oracle():
    """src/types.lisp:41-107: LOGIOR of the type codes; integers and rationals classify as double."""
    cases = [np.zeros(2, np.float32), np.zeros(2, np.float64), np.zeros(2, np.complex64), np.zeros(2, np.complex128),
             np.zeros(2, np.int64), np.zeros(2, bool)]
    for x in cases:
        for y in cases:
            assert L.common_float_type(x, y) == O.common_float_type(x, y)
    assert L.common_float_type(np.zeros(1, np.int32)) == L.DOUBLE
    assert L.common_float_type(np.zeros(1, np.float32), np.zeros(1, np.float64)) == L.DOUBLE
    assert L.common_float_type(np.zeros(1, np.float64), np.zeros(1, np.complex64)) == L.COMPLEX_DOUBLE


def test_dimension_rules_raise_before_any_foreign_call():
    """src/linear-algebra.lisp:18-26, :44-47, :268, :281: these are Lisp-side asserts, no device needed."""
    assert L.dimensions_as_matrix(np.zeros(3), "row") == (1, 3, True)
    assert L.dimensions_as_matrix(np.zeros(3), "column") == (3, 1, True)
    assert L.dimensions_as_matrix(np.zeros((2, 5)), "row") == (2, 5, False)
    with pytest.raises(ValueError):
        L.mm(np.zeros(3), np.zeros(3))                       # "MM called with two vectors" (tests/linear-algebra.lisp:27-29)
    with pytest.raises(AssertionError):
        L.mm(np.zeros((2, 3)), np.zeros((2, 3)))
    with pytest.raises(L.LlaIncompatibleDimensions):
        L.solve(np.zeros((3, 3)), np.zeros(4))
    with pytest.raises(L.LlaIncompatibleDimensions):
        L.solve(L.LU(np.zeros((3, 3)), np.zeros(3, np.int32)), np.zeros((2, 2)))
    with pytest.raises(L.LlaIncompatibleDimensions):
        L.solve(L.Cholesky(np.zeros((3, 3))), np.zeros(5))
    with pytest.raises(L.LlaIncompatibleDimensions):
        L.cholesky(np.zeros((3, 4)))
    with pytest.raises(L.LlaIncompatibleDimensions):
        L.solve(L.Hermitian(np.zeros((3, 3))), np.zeros(2))
    with pytest.raises(L.LlaIncompatibleDimensions):
        L.trsm(np.zeros((3, 3)), True, np.zeros((4, 1)))
    with pytest.raises(AssertionError):
        L.gemm(1.0, np.zeros((4, 3)), np.zeros((3, 5)), 0.0, np.zeros((4, 5)), lda=2)   # k <= lda, src/blas.lisp:40-53


def test_wrappers_and_ipiv_postprocessing():
    """src/factorizations.lisp:15-65, :140-151 and cl-num-utils' hermitian-matrix (lower triangle defines the matrix)."""
    h = L.Hermitian(np.array([[2.0, 99.0], [-1.0, 3.0]]))
    assert np.array_equal(h.as_array(), [[2.0, -1.0], [-1.0, 3.0]])
    piv = np.array([3, 3, 3], dtype=np.int32)                 # LAPACK swap sequence: 1<->3, 2<->3, 3<->3
    assert np.array_equal(L.ipiv(piv), O.ipiv(piv))
    assert np.array_equal(L.ipiv_inverse(piv), O.ipiv_inverse(piv))
    assert L.permutations(piv) == O.permutations(piv) == 2
    f = L.LU(np.array([[4.0, 3.0], [0.5, 2.0]]), np.array([2, 2], dtype=np.int32))
    assert np.array_equal(L.lu_l(f), [[1.0, 0.0], [0.5, 1.0]]) and np.array_equal(L.lu_u(f), [[4.0, 3.0], [0.0, 2.0]])
    a = np.array([[1.0, 2.0], [8.0, 7.0]])                    # P A = L U with P from ipiv (tests/linear-algebra.lisp:114)
    fo = O.lu(a, O.netlib())
    assert np.allclose(a[O.ipiv(fo.ipiv_internal)], O.lu_l(fo) @ O.lu_u(fo))


def test_efficiency_warnings_follow_the_reference_toggles():
    """reference src/conditions.lisp:49-74: off by default, one warning class per toggle."""
    import warnings
    from lla_b200 import lla as L
    a = np.arange(6).reshape(2, 3)                      # integer array: classifies as double, needs a converting copy
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        L._transposed_to_memory(a, L.DOUBLE)            # toggles off: silent
    L.efficiency_warning_array_conversion = True
    L.efficiency_warning_array_type = True
    try:
        with pytest.warns(L.LlaEfficiencyWarningArrayConversion):
            L._transposed_to_memory(a, L.DOUBLE)
        with pytest.warns(L.LlaEfficiencyWarningArrayType):
            L._as_memory(np.array([[1, 2.5]], dtype=object), L.DOUBLE)
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            L._as_memory(np.ones((2, 2)), L.DOUBLE)     # shared as it is: nothing to warn about
    finally:
        L.efficiency_warning_array_conversion = False
        L.efficiency_warning_array_type = False
    assert issubclass(L.LlaEfficiencyWarningArrayType, L.LlaEfficiencyWarning)


def test_min_dimension_threshold():
    from lla_b200 import lla as L
    assert L.device_worthwhile(4096, 4096) and not L.device_worthwhile(4096, 8)
