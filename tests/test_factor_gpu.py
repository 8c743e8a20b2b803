"""GPU parity: LU / solve / Cholesky through the Fortran drop-ins against the CPU oracle.

Reference tests restated (inputs and expected values only):
  mm-solve-lu    tests/linear-algebra.lisp:89-104
  lu-singular    tests/linear-algebra.lisp:106-107
  lu-random-test tests/linear-algebra.lisp:109-115
  cholesky       tests/linear-algebra.lisp:366-385
Bars (north_star): ipiv bit-exact with the reference LAPACK (oracle: netlib restatement AND host
OpenBLAS); solves with scaled backward residual ||AX-B|| / (n eps ||A|| ||X||) <= 10.
"""
import numpy as np
import pytest

from lla_b200 import lla as L
from oracle import lla_oracle as O

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def num_eq(a, b, tol=1e-5):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b))))


def residual(a, x, b):
    a, x, b = np.asarray(a, dtype=np.float64), np.asarray(x), np.asarray(b)
    n = a.shape[0]
    return np.linalg.norm(a @ x - b) / (n * EPS * np.linalg.norm(a) * np.linalg.norm(x))


def test_mm_solve_lu_golden():
    a = np.array([[1, 2], [3, 4]], dtype=np.int64)
    x = np.array([5.0, 6.0])
    b = L.mm(a, x)
    assert num_eq(b, [17.0, 39.0])
    assert num_eq(L.solve(a, b), x)
    f = L.lu(a)
    assert list(f.ipiv_internal) == [2, 2]
    assert num_eq(L.solve(f, b), x)


def test_lu_singular_does_not_signal():
    f = L.lu(np.full((2, 2), 0.5))
    ref = O.lu(np.full((2, 2), 0.5), O.netlib())
    assert np.array_equal(f.ipiv_internal, ref.ipiv_internal)
    assert np.array_equal(f.lu, ref.lu)
    with pytest.raises(L.LapackFailure) as e:
        L.solve(np.full((2, 2), 0.5), np.ones(2))
    assert e.value.info == 2


def test_lu_random_reconstruction():
    rng = np.random.default_rng(7)
    for n in range(2, 11):
        for _ in range(20):
            a = rng.uniform(0, 10, (n, n))
            f = L.lu(a)
            ref = O.lu(a, O.netlib())
            assert np.array_equal(f.ipiv_internal, ref.ipiv_internal)
            p, pinv = L.ipiv(f.ipiv_internal), L.ipiv_inverse(f.ipiv_internal)
            lu_prod = L.lu_l(f) @ L.lu_u(f)
            assert num_eq(a[p, :], lu_prod)
            assert num_eq(a, lu_prod[pinv, :])


@pytest.mark.parametrize("shape", [(1, 1), (5, 5), (33, 33), (64, 64), (100, 100), (129, 129), (300, 300),
                                   (9, 5), (5, 9), (64, 17), (200, 70), (70, 200), (257, 130), (513, 513)])
def test_lu_vs_oracle(shape):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    a = rng.uniform(0, 10, shape)
    f = L.lu(a)
    r1, r2 = O.lu(a, O.netlib()), O.lu(a, O.openblas())
    assert np.array_equal(r1.ipiv_internal, r2.ipiv_internal)
    assert np.array_equal(f.ipiv_internal, r1.ipiv_internal)
    scale = max(shape) * EPS * np.abs(r1.lu).max() * 50
    assert np.max(np.abs(f.lu - r1.lu)) <= scale


def test_lu_small_is_bit_identical_to_reference_algorithm():
    """n <= 32 is a single DGETF2 leaf: same operations in the same order => identical bits."""
    rng = np.random.default_rng(12)
    for n in (2, 5, 17, 32):
        for _ in range(10):
            a = rng.uniform(-10, 10, (n, n))
            f, r = L.lu(a), O.lu(a, O.netlib())
            assert np.array_equal(f.ipiv_internal, r.ipiv_internal)
            assert np.array_equal(f.lu, r.lu)
        b = rng.integers(-2, 3, (n, n)).astype(np.float64)
        f, r = L.lu(b), O.lu(b, O.netlib())
        assert np.array_equal(f.ipiv_internal, r.ipiv_internal)
        assert np.array_equal(np.nan_to_num(f.lu, nan=1e300), np.nan_to_num(r.lu, nan=1e300))


def test_lu_ties_and_zero_columns():
    """exact ties (first index must win), a zero column (INFO = its index, factorisation goes on)."""
    a = np.ones((40, 40))
    a[np.arange(40), np.arange(40)] = -1.0   # |a| all equal: pivot must always be the diagonal... after updates
    f, r = L.lu(a), O.lu(a, O.netlib())
    assert np.array_equal(f.ipiv_internal, r.ipiv_internal)
    rng = np.random.default_rng(3)
    b = rng.integers(-3, 4, (70, 70)).astype(np.float64)
    f, r = L.lu(b), O.lu(b, O.netlib())
    assert np.array_equal(f.ipiv_internal, r.ipiv_internal)
    c = rng.uniform(0, 10, (50, 50))
    c[:, 7] = 0.0
    f, r = L.lu(c), O.lu(c, O.netlib())
    assert np.array_equal(f.ipiv_internal, r.ipiv_internal)
    assert np.allclose(f.lu, r.lu, atol=1e-10)


def test_lu_nan_pivot_rule():
    """IDAMAX skips NaN unless it is the first candidate."""
    rng = np.random.default_rng(9)
    a = rng.uniform(0, 10, (20, 20))
    a[5, 0] = np.nan
    f, r = L.lu(a), O.lu(a, O.netlib())
    assert f.ipiv_internal[0] == r.ipiv_internal[0]
    a = rng.uniform(0, 10, (20, 20))
    a[0, 0] = np.nan
    f, r = L.lu(a), O.lu(a, O.netlib())
    assert f.ipiv_internal[0] == r.ipiv_internal[0] == 1


@pytest.mark.parametrize("n,nrhs", [(3, 1), (37, 5), (64, 64), (130, 3), (300, 17), (512, 64)])
def test_solve_vs_oracle(n, nrhs):
    rng = np.random.default_rng(n + nrhs)
    a = rng.uniform(0, 10, (n, n))
    b = rng.uniform(0, 10, (n, nrhs)) if nrhs > 1 else rng.uniform(0, 10, n)
    x = L.solve(a, b)
    xr = O.solve(a, b, O.openblas())
    assert x.shape == xr.shape == b.shape
    assert residual(a, x, b) <= 10
    f = L.lu(a)
    x2 = L.solve(f, b)
    assert residual(a, x2, b) <= 10
    cond = np.linalg.cond(a)
    assert np.max(np.abs(x - xr)) <= 100 * cond * EPS * np.abs(xr).max()


def test_solve_dimension_errors():
    with pytest.raises(L.LlaIncompatibleDimensions):
        L.solve(np.eye(3), np.ones(4))
    with pytest.raises(L.LlaIncompatibleDimensions):
        L.solve(np.ones((3, 4)), np.ones(3))


def test_cholesky_golden():
    a = np.array([[2, -1, 0], [-1, 2, -1], [0, -1, 2]], dtype=np.float64)
    l = np.array([[1.414214, 0, 0], [-0.7071068, 1.2247449, 0], [0.0, -0.8164966, 1.1547005]])
    c = L.cholesky(a)
    assert num_eq(c.left, l)
    b = np.array([5.0, 7.0, 13.0])
    ab = L.solve(a, b)
    assert num_eq(L.solve(c, b), ab)
    assert num_eq(a @ ab, b)
    a2 = a.copy()
    a2[0, 2] = 99.0
    a2[0, 1] = -55.0
    assert np.array_equal(L.cholesky(a2).left, c.left)   # only the row-major lower triangle is read
    with pytest.raises(L.LapackFailure) as e:
        L.cholesky(np.array([[1.0, 2.0], [2.0, 1.0]]))
    assert e.value.info == 2


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 64, 100, 257, 600])
def test_cholesky_vs_oracle(n):
    rng = np.random.default_rng(n)
    g = rng.uniform(0, 1, (n, n))
    spd = g @ g.T + n * np.eye(n)
    c = L.cholesky(spd)
    r = O.cholesky(spd, O.openblas())
    assert np.max(np.abs(c.left - r.left)) <= 20 * n * EPS * np.abs(r.left).max()
    assert np.linalg.norm(c.left @ c.left.T - spd) / (n * EPS * np.linalg.norm(spd)) <= 10
    nrhs = 1 + n % 7
    b = rng.uniform(0, 1, (n, nrhs))
    x = L.solve(c, b)
    assert residual(spd, x, b) <= 10
    assert np.allclose(x, O.solve_cholesky(r, b, O.openblas()), rtol=1e-9, atol=1e-12)


def test_cholesky_not_pd_index_and_untouched_triangle():
    rng = np.random.default_rng(2)
    n = 150
    g = rng.uniform(0, 1, (n, n))
    spd = g @ g.T + n * np.eye(n)
    bad = spd.copy()
    bad[100, 100] = -1.0
    with pytest.raises(L.LapackFailure) as e:
        L.cholesky(bad)
    with pytest.raises(O.LapackFailure) as eo:
        O.cholesky(bad, O.netlib())
    assert e.value.info == eo.value.info == 101


@pytest.mark.parametrize("shape", [(1300, 1300), (1500, 1100), (1100, 1500), (2111, 2111)])
def test_lu_blocked_lookahead_path_vs_oracle(shape):
    """min(m, n) > 1024 takes the blocked driver with look-ahead (three streams): ipiv must still be
    bit-exact with both oracles, and P A = L U must hold."""
    rng = np.random.default_rng(shape[0] + 3 * shape[1])
    a = rng.uniform(0, 10, shape)
    f = L.lu(a)
    r2 = O.lu(a, O.openblas())
    assert np.array_equal(f.ipiv_internal, r2.ipiv_internal)
    if shape[0] <= 1500:
        r1 = O.lu(a, O.netlib())
        assert np.array_equal(f.ipiv_internal, r1.ipiv_internal)
    scale = max(shape) * EPS * np.abs(r2.lu).max() * 50
    assert np.max(np.abs(f.lu - r2.lu)) <= scale
    if shape[0] == shape[1]:
        p = L.ipiv(f.ipiv_internal)
        rec = L.lu_l(f) @ L.lu_u(f)
        assert np.linalg.norm(a[p, :] - rec) / (shape[0] * EPS * np.linalg.norm(a)) <= 10


@pytest.mark.parametrize("n", [1100, 1700, 2304])
def test_cholesky_blocked_lookahead_path(n):
    rng = np.random.default_rng(n)
    g = rng.uniform(0, 1, (n, n))
    spd = g @ g.T + n * np.eye(n)
    c = L.cholesky(spd)
    r = O.cholesky(spd, O.openblas())
    assert np.max(np.abs(c.left - r.left)) <= 20 * n * EPS * np.abs(r.left).max()
    assert np.linalg.norm(c.left @ c.left.T - spd) / (n * EPS * np.linalg.norm(spd)) <= 10
    b = rng.uniform(0, 1, (n, 3))
    assert residual(spd, L.solve(c, b), b) <= 10
    bad = spd.copy()
    bad[n - 7, n - 7] = -1.0
    with pytest.raises(L.LapackFailure) as e:
        L.cholesky(bad)
    assert e.value.info == n - 6


def test_solve_large_residual():
    rng = np.random.default_rng(77)
    n, nrhs = 2500, 64
    a, b = rng.uniform(0, 10, (n, n)), rng.uniform(0, 10, (n, nrhs))
    x = L.solve(a, b)
    assert residual(a, x, b) <= 10
    assert np.allclose(x, O.solve(a, b, O.openblas()), rtol=1e-6, atol=1e-8)


# ------------------------------------------------------------------ page-locked host operands: the Fortran drop-ins
# send finished block rows back while the factorisation is still running (RowsFinalHook) and move only the
# named triangle for DPOTRF; the results must be the bytes the plain (pageable) path produces.
def _pinned_like(a):
    import ctypes
    from lla_b200 import _lib
    lib = _lib.lib()
    p = ctypes.c_void_p()
    lib.llacuda_host_alloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
    assert lib.llacuda_host_alloc(ctypes.byref(p), a.nbytes) == 0
    buf = np.ctypeslib.as_array((ctypes.c_char * a.nbytes).from_address(p.value)).view(a.dtype).reshape(a.shape)
    buf[...] = a
    return buf, p


@pytest.mark.parametrize("n", [1536, 2100])
def test_pinned_host_overlapped_downloads_match_pageable(n):
    import ctypes
    from lla_b200 import _lib
    lib = _lib.lib()
    I = lambda v: ctypes.byref(ctypes.c_int(v))
    P = lambda arr: arr.ctypes.data_as(ctypes.c_void_p)
    rng = np.random.default_rng(n)
    # ---- dgetrf_ / dgesv_ (column-major images)
    a = np.asfortranarray(rng.uniform(0, 10, (n, n)))
    b = np.asfortranarray(rng.uniform(0, 10, (n, 3)))
    outs = []
    for pinned in (False, True):
        aa, bb = a.copy(order="F"), b.copy(order="F")
        keep = []
        if pinned:
            aa_, h1 = _pinned_like(aa.T); bb_, h2 = _pinned_like(bb.T); keep = [h1, h2]   # .T of F-order = C-order bytes
            aw, bw = aa_, bb_
        else:
            aw, bw = np.ascontiguousarray(aa.T), np.ascontiguousarray(bb.T)
        ipiv = np.zeros(n, dtype=np.int32); info = ctypes.c_int(0)
        lib.dgesv_(I(n), I(3), P(aw), I(n), P(ipiv), P(bw), I(n), ctypes.byref(info))
        assert info.value == 0
        outs.append((aw.copy(), bw.copy(), ipiv.copy()))
        for h in keep:
            lib.llacuda_host_free(h)
    assert np.array_equal(outs[0][2], outs[1][2])
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert residual(a, outs[1][1].T, b) <= 10
    # ---- dpotrf_ 'U': only the upper triangle of the column-major view is read or written
    spd = a @ a.T + n * np.eye(n)
    img = np.ascontiguousarray(np.triu(spd).T)            # column-major image, lower triangle of the view = sentinel
    img[np.triu_indices(n, 1)] = -777.0                   # image[c, r] with c < r  <=>  matrix (r, c) below the diagonal
    res = []
    for pinned in (False, True):
        if pinned:
            w, h = _pinned_like(img)
        else:
            w, h = img.copy(), None
        info = ctypes.c_int(0)
        lib.dpotrf_(ctypes.c_char_p(b"U"), I(n), P(w), I(n), ctypes.byref(info))
        assert info.value == 0
        res.append(w.copy())
        if h is not None:
            lib.llacuda_host_free(h)
    assert np.array_equal(res[0], res[1])
    u = np.triu(res[1].T)
    assert np.all(res[1][np.triu_indices(n, 1)] == -777.0), "the other triangle must come back untouched"
    assert np.linalg.norm(u.T @ u - spd) / (n * EPS * np.linalg.norm(spd)) <= 10


# ------------------------------------------------------------------ device-resident factorisation objects
@pytest.mark.parametrize("n", [7, 300, 1500])
def test_device_resident_lu_and_cholesky_objects(n):
    """lu_device / cholesky_device keep the factors in HBM (column-major, converted on the device); solves move only
    B and X.  Same kernels on the same numbers as the host-pointer path, so factors and pivots are identical."""
    rng = np.random.default_rng(n + 1)
    a = rng.uniform(0, 10, (n, n))
    b = rng.uniform(0, 10, (n, 3))
    f, fd = L.lu(a), L.lu_device(a)
    assert isinstance(fd, L.LU)
    assert np.array_equal(fd.ipiv_internal, f.ipiv_internal)
    assert np.array_equal(fd.lu, f.lu)
    for rhs in (b, b[:, 0]):
        x, xd = L.solve(f, rhs), L.solve(fd, rhs)
        assert xd.shape == x.shape
        assert np.allclose(xd, x, rtol=0, atol=1e-13 * np.abs(x).max())
        assert residual(a, xd.reshape(n, -1), rhs.reshape(n, -1)) <= 10
    assert np.array_equal(L.lu_u(fd), L.lu_u(f)) and np.array_equal(L.ipiv(fd.ipiv_internal), L.ipiv(f.ipiv_internal))
    spd = a @ a.T + n * np.eye(n)
    c, cd = L.cholesky(spd), L.cholesky_device(spd)
    assert isinstance(cd, L.Cholesky)
    assert np.array_equal(cd.left, c.left)
    y, yd = L.solve(c, b), L.solve(cd, b)
    assert np.allclose(yd, y, rtol=0, atol=1e-13 * np.abs(y).max())
    with pytest.raises(L.LlaIncompatibleDimensions):
        L.solve(fd, np.ones(n + 1))
    bad = spd.copy(); bad[n // 2, n // 2] = -1.0
    with pytest.raises(L.LapackFailure) as e:
        L.cholesky_device(bad)
    assert e.value.info == n // 2 + 1
    # rectangular lu stays rectangular
    r = rng.uniform(0, 10, (n, n + 5))
    fr, frd = L.lu(r), L.lu_device(r)
    assert np.array_equal(frd.ipiv_internal, fr.ipiv_internal) and np.array_equal(frd.lu, fr.lu)


def test_ilp64_library_matches_lp64():
    """libllacuda64.so (64-bit integer cells, LLA :int64): same factors, pivots and solution as libllacuda.so."""
    import ctypes
    from lla_b200 import _lib
    lib32, lib64 = _lib.lib(), ctypes.CDLL(_lib.SO64_PATH)
    n, nrhs = 700, 3
    rng = np.random.default_rng(64)
    a = np.asfortranarray(rng.uniform(0, 10, (n, n)))
    b = np.asfortranarray(rng.uniform(0, 10, (n, nrhs)))
    P = lambda arr: arr.ctypes.data_as(ctypes.c_void_p)
    outs = []
    for lib, ct, dt in ((lib32, ctypes.c_int, np.int32), (lib64, ctypes.c_int64, np.int64)):
        aw, bw = a.T.copy(order="C"), b.T.copy(order="C")                  # column-major images (copies: a.T of an F-ordered array is already C-contiguous)
        ipiv = np.zeros(n, dtype=dt); info = ct(0)
        I = lambda v: ctypes.byref(ct(v))
        lib.dgesv_(I(n), I(nrhs), P(aw), I(n), P(ipiv), P(bw), I(n), ctypes.byref(info))
        assert info.value == 0
        spd = np.ascontiguousarray((a @ a.T + n * np.eye(n)).T)
        lib.dpotrf_(ctypes.c_char_p(b"U"), I(n), P(spd), I(n), ctypes.byref(info))
        assert info.value == 0
        bad = ct(0)
        lib.dpotrf_(ctypes.c_char_p(b"U"), I(-3), P(spd), I(n), ctypes.byref(bad))
        assert bad.value == -2
        outs.append((aw, bw, ipiv.astype(np.int64), spd))
    for x, y in zip(outs[0], outs[1]):
        assert np.array_equal(x, y)


# ------------------------------------------------------------------ "next" rows on the device: dsyrk_, dtrsm_, dposv_
def test_mm_hermitian_trsm_posv_reference_goldens():
    a = np.array([[1.0, 2.0], [3.0, 4.0]])
    assert np.array_equal(L.mm(a, True).as_array(), a @ a.T)          # tests/linear-algebra.lisp:117-124
    assert np.array_equal(L.mm(True, a).as_array(), a.T @ a)
    u, l, b = np.array([[1.0, 2.0], [0.0, 3.0]]), np.array([[1.0, 0.0], [2.0, 3.0]]), np.array([[5.0, 6.0], [7.0, 8.0]])
    assert num_eq(u @ L.solve(L.UpperTriangular(u), b), b)              # :167-181
    assert num_eq(l @ L.solve(L.LowerTriangular(l), b), b)
    spd = np.array([[2.0, -1, 0], [-1, 2, -1], [0, -1, 2]])
    rhs = np.array([5.0, 7.0, 13.0])
    assert num_eq(L.solve(L.Hermitian(spd), rhs), L.solve(spd, rhs))    # :366-385
    with pytest.raises(L.LapackFailure):
        L.solve(L.Hermitian(-spd), rhs)


@pytest.mark.parametrize("m,k", [(33, 7), (300, 129), (1100, 600)])
def test_mm_hermitian_trsm_posv_vs_oracle(m, k):
    rng = np.random.default_rng(m + k)
    r = rng.uniform(-1, 1, (m, k))
    for tl in (False, True):
        got = L.mm_hermitian(r, tl)
        want = O.mm_hermitian(r, tl, O.openblas())
        kk = m if tl else k
        tol = 4 * np.sqrt(kk) * EPS * (np.linalg.norm(r, axis=0 if tl else 1).max() ** 2)
        assert np.max(np.abs(got.as_array() - want)) <= tol
        assert np.all(np.triu(got.elements, 1) == 0.0), "SYRK('U') of the column-major view must not write the other triangle"
    t = np.triu(rng.uniform(-1, 1, (m, m))) + 4 * np.eye(m)
    b = rng.uniform(-1, 1, (m, 5))
    for mat, upper in ((t, True), (t.T.copy(), False)):
        x = L.trsm(mat, upper, b)
        assert residual(mat, x, b) <= 10
        xo = O.trsm(mat, upper, b, O.openblas())
        assert np.max(np.abs(x - xo)) <= 1e-10 * np.abs(xo).max()
    g = rng.uniform(0, 1, (m, m))
    spd = g @ g.T + m * np.eye(m)
    x = L.solve(L.Hermitian(spd), b)
    assert residual(spd, x, b) <= 10


# ------------------------------------------------------------------ row-major-native entry points
@pytest.mark.parametrize("n", [5, 257, 1300])
def test_row_major_native_entry_points_match_the_fortran_path(n):
    """llacuda_dgetrf_rm / dgesv_rm / dgetrs_rm / dpotrs_rm take LLA's row-major arrays as they are: same kernels on
    the same numbers as the Fortran-symbol path behind LLA's transposing copies, so the results are identical."""
    rng = np.random.default_rng(3 * n)
    a = rng.uniform(0, 10, (n, n))
    b = rng.uniform(0, 10, (n, 4))
    f, fn = L.lu(a), L.lu(a, native=True)
    assert np.array_equal(fn.ipiv_internal, f.ipiv_internal) and np.array_equal(fn.lu, f.lu)
    r = rng.uniform(0, 10, (n, n + 3))
    fr, frn = L.lu(r), L.lu(r, native=True)
    assert np.array_equal(frn.ipiv_internal, fr.ipiv_internal) and np.array_equal(frn.lu, fr.lu)
    a_keep = a.copy()
    for rhs in (b, b[:, 0]):
        assert np.array_equal(L.solve(a, rhs, native=True), L.solve(a, rhs))
        assert np.array_equal(L.solve(f, rhs, native=True), L.solve(f, rhs))
    assert np.array_equal(a, a_keep), "A is input-only"
    spd = a @ a.T + n * np.eye(n)
    c = L.cholesky(spd)
    assert np.array_equal(L.solve(c, b, native=True), L.solve(c, b))
    with pytest.raises(L.LapackFailure) as e:
        L.solve(np.full((2, 2), 0.5), np.ones(2), native=True)
    assert e.value.info == 2
