/* oracle/netlib_templ.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Type-generic body of the CPU oracle, included four times by netlib_ref.c
 * (once per LLA internal float type: s, d, c, z -- src/types.lisp:22-26).
 *
 * It restates, as plain scalar C, the *published reference algorithms* of the
 * BLAS/LAPACK routines that LLA's hot path binds to through
 * cffi:foreign-funcall (src/fortran-call.lisp:428-439, name mangling :405-416):
 *
 *   xGEMM                       <- mm        src/linear-algebra.lisp:50-54
 *                                  gemm!     src/blas.lisp:56-63
 *   xGETRF (as xGETF2 + xLASWP) <- lu        src/linear-algebra.lisp:227-234
 *   xGETRS                      <- solve/lu  src/linear-algebra.lisp:269-276
 *   xGESV                       <- solve     src/linear-algebra.lisp:282-289
 *   xPOTRF (as xPOTF2)          <- cholesky  src/linear-algebra.lisp:670-674
 *   xPOTRS                      <- solve/chol src/linear-algebra.lisp:296-301
 *
 * The arithmetic is NOT in /root/reference: LLA loads whatever libblas /
 * liblapack the user has (src/libraries.lisp:10-12, src/configuration.lisp:12-13),
 * unpinned.  The algorithm restated is Reference LAPACK 3.x (netlib):
 * IxAMAX = first index of max |re|+|im|; xGETF2 right-looking with reciprocal
 * scaling when |pivot| >= sfmin; xPOTF2 dot-product form; INFO conventions.
 *
 * Macros expected:  T (element type), R (real type), PFX(name), IS_CPLX (0/1),
 *                   R_EPS_SFMIN (safe minimum of R)
 */

#if IS_CPLX
#define ABS1(x) (R_ABS(creal_(x)) + R_ABS(cimag_(x)))
#define CONJ(x) conj_(x)
#define REAL(x) creal_(x)
#else
#define ABS1(x) R_ABS(x)
#define CONJ(x) (x)
#define REAL(x) (x)
#endif

static int PFX(lsame)(char a, char b) {
  if (a >= 'a' && a <= 'z') a = (char)(a - 32);
  return a == b;
}

/* 'C' on a real type means 'T' (LLA always sends #\C for a transposed
 * operand: src/blas.lisp:57-58). */
/* xGEMM: C <- alpha op(A) op(B) + beta C, column-major, Fortran ABI.
 * beta == 0 must not read C (netlib sets C to zero first). */
void PFX(gemm_)(const char *transa, const char *transb, const int *m_, const int *n_,
                const int *k_, const T *alpha_, const T *a, const int *lda_, const T *b,
                const int *ldb_, const T *beta_, T *c, const int *ldc_) {
  int64_t m = *m_, n = *n_, k = *k_, lda = *lda_, ldb = *ldb_, ldc = *ldc_;
  T alpha = *alpha_, beta = *beta_;
  int ma = PFX(lsame)(*transa, 'N') ? 0 : (PFX(lsame)(*transa, 'T') ? 1 : 2);
  int mb = PFX(lsame)(*transb, 'N') ? 0 : (PFX(lsame)(*transb, 'T') ? 1 : 2);
  if (m <= 0 || n <= 0) return;
  /* netlib order: for j { scale C(:,j) by beta; for l { temp = alpha*B(l,j); for i ... } } */
  for (int64_t j = 0; j < n; j++) {
    if (beta == (T)0) {
      for (int64_t i = 0; i < m; i++) c[i + j * ldc] = 0;
    } else if (beta != (T)1) {
      for (int64_t i = 0; i < m; i++) c[i + j * ldc] *= beta;
    }
    if (alpha == (T)0) continue;
    for (int64_t l = 0; l < k; l++) {
      T blj = (mb == 0) ? b[l + j * ldb] : (mb == 1 ? b[j + l * ldb] : CONJ(b[j + l * ldb]));
      T temp = alpha * blj;
      if (ma == 0) {
        const T *acol = a + l * lda;
        for (int64_t i = 0; i < m; i++) c[i + j * ldc] += temp * acol[i];
      } else if (ma == 1) {
        for (int64_t i = 0; i < m; i++) c[i + j * ldc] += temp * a[l + i * lda];
      } else {
        for (int64_t i = 0; i < m; i++) c[i + j * ldc] += temp * CONJ(a[l + i * lda]);
      }
    }
  }
}

/* IxAMAX, 0-based return: first index attaining max |re|+|im|; comparisons
 * with NaN are false, exactly as the Fortran loop behaves. */
static int64_t PFX(iamax)(int64_t n, const T *x) {
  if (n < 1) return 0;
  int64_t best = 0;
  R dmax = ABS1(x[0]);
  for (int64_t i = 1; i < n; i++) {
    R v = ABS1(x[i]);
    if (v > dmax) { dmax = v; best = i; }
  }
  return best;
}

/* xGETRF, computed with the unblocked xGETF2 recurrence (pivot choice of the
 * blocked and recursive LAPACK variants is the same function of the data). */
void PFX(getrf_)(const int *m_, const int *n_, T *a, const int *lda_, int *ipiv, int *info) {
  int64_t m = *m_, n = *n_, lda = *lda_;
  *info = 0;
  if (m < 0) { *info = -1; return; }
  if (n < 0) { *info = -2; return; }
  if (lda < (m > 1 ? m : 1)) { *info = -4; return; }
  if (m == 0 || n == 0) return;
  int64_t mn = m < n ? m : n;
  for (int64_t j = 0; j < mn; j++) {
    int64_t jp = j + PFX(iamax)(m - j, a + j + j * lda);
    ipiv[j] = (int)(jp + 1);
    if (a[jp + j * lda] != (T)0) {
      if (jp != j)
        for (int64_t c = 0; c < n; c++) {
          T t = a[j + c * lda]; a[j + c * lda] = a[jp + c * lda]; a[jp + c * lda] = t;
        }
      if (j < m - 1) {
        T piv = a[j + j * lda];
        if (ABS1(piv) >= R_EPS_SFMIN || IS_CPLX) {
          /* netlib: DSCAL by ONE/A(j,j) when |A(j,j)| >= sfmin (ZGETF2 tests the modulus) */
          T r = (T)1 / piv;
          for (int64_t i = j + 1; i < m; i++) a[i + j * lda] *= r;
        } else {
          for (int64_t i = j + 1; i < m; i++) a[i + j * lda] /= piv;
        }
      }
    } else if (*info == 0) {
      *info = (int)(j + 1);
    }
    if (j < mn - 1 || n > mn) {
      for (int64_t c = j + 1; c < n; c++) {
        T ujc = a[j + c * lda];
        if (ujc != (T)0)
          for (int64_t i = j + 1; i < m; i++) a[i + c * lda] -= a[i + j * lda] * ujc;
      }
    }
  }
}

/* xGETRS: solve op(A) X = B from the factors; B overwritten by X. */
void PFX(getrs_)(const char *trans, const int *n_, const int *nrhs_, const T *a, const int *lda_,
                 const int *ipiv, T *b, const int *ldb_, int *info) {
  int64_t n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
  int mode = PFX(lsame)(*trans, 'N') ? 0 : (PFX(lsame)(*trans, 'T') ? 1 : (PFX(lsame)(*trans, 'C') ? 2 : -1));
  *info = 0;
  if (mode < 0) { *info = -1; return; }
  if (n < 0) { *info = -2; return; }
  if (nrhs < 0) { *info = -3; return; }
  if (lda < (n > 1 ? n : 1)) { *info = -5; return; }
  if (ldb < (n > 1 ? n : 1)) { *info = -8; return; }
  if (n == 0 || nrhs == 0) return;
  if (mode == 0) {
    for (int64_t i = 0; i < n; i++) {            /* xLASWP forward */
      int64_t p = ipiv[i] - 1;
      if (p != i) for (int64_t c = 0; c < nrhs; c++) { T t = b[i + c * ldb]; b[i + c * ldb] = b[p + c * ldb]; b[p + c * ldb] = t; }
    }
    for (int64_t c = 0; c < nrhs; c++) {
      T *x = b + c * ldb;
      for (int64_t k = 0; k < n; k++)            /* L unit lower */
        if (x[k] != (T)0) for (int64_t i = k + 1; i < n; i++) x[i] -= x[k] * a[i + k * lda];
      for (int64_t k = n - 1; k >= 0; k--) {     /* U upper */
        if (x[k] != (T)0) {
          x[k] /= a[k + k * lda];
          for (int64_t i = 0; i < k; i++) x[i] -= x[k] * a[i + k * lda];
        }
      }
    }
  } else {
    for (int64_t c = 0; c < nrhs; c++) {
      T *x = b + c * ldb;
      for (int64_t j = 0; j < n; j++) {          /* U^T / U^H */
        T t = x[j];
        for (int64_t i = 0; i < j; i++) t -= (mode == 2 ? CONJ(a[i + j * lda]) : a[i + j * lda]) * x[i];
        x[j] = t / (mode == 2 ? CONJ(a[j + j * lda]) : a[j + j * lda]);
      }
      for (int64_t j = n - 1; j >= 0; j--) {     /* L^T / L^H, unit */
        T t = x[j];
        for (int64_t i = j + 1; i < n; i++) t -= (mode == 2 ? CONJ(a[i + j * lda]) : a[i + j * lda]) * x[i];
        x[j] = t;
      }
    }
    for (int64_t i = n - 1; i >= 0; i--) {       /* xLASWP backward */
      int64_t p = ipiv[i] - 1;
      if (p != i) for (int64_t c = 0; c < nrhs; c++) { T t = b[i + c * ldb]; b[i + c * ldb] = b[p + c * ldb]; b[p + c * ldb] = t; }
    }
  }
}

/* xGESV = xGETRF, then xGETRS iff INFO == 0. */
void PFX(gesv_)(const int *n_, const int *nrhs_, T *a, const int *lda_, int *ipiv, T *b,
                const int *ldb_, int *info) {
  int64_t n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
  *info = 0;
  if (n < 0) { *info = -1; return; }
  if (nrhs < 0) { *info = -2; return; }
  if (lda < (n > 1 ? n : 1)) { *info = -4; return; }
  if (ldb < (n > 1 ? n : 1)) { *info = -7; return; }
  PFX(getrf_)(n_, n_, a, lda_, ipiv, info);
  if (*info == 0) PFX(getrs_)("N", n_, nrhs_, a, lda_, ipiv, b, ldb_, info);
}

/* xPOTRF via the xPOTF2 recurrence.  Only the named triangle is touched. */
void PFX(potrf_)(const char *uplo, const int *n_, T *a, const int *lda_, int *info) {
  int64_t n = *n_, lda = *lda_;
  int upper = PFX(lsame)(*uplo, 'U');
  *info = 0;
  if (!upper && !PFX(lsame)(*uplo, 'L')) { *info = -1; return; }
  if (n < 0) { *info = -2; return; }
  if (lda < (n > 1 ? n : 1)) { *info = -4; return; }
  for (int64_t j = 0; j < n; j++) {
    R ajj = REAL(a[j + j * lda]);
    if (upper) { for (int64_t k = 0; k < j; k++) ajj -= REAL(CONJ(a[k + j * lda]) * a[k + j * lda]); }
    else       { for (int64_t k = 0; k < j; k++) ajj -= REAL(a[j + k * lda] * CONJ(a[j + k * lda])); }
    if (ajj <= 0 || ajj != ajj) { a[j + j * lda] = ajj; *info = (int)(j + 1); return; }
    ajj = R_SQRT(ajj);
    a[j + j * lda] = ajj;
    if (upper) {
      for (int64_t c = j + 1; c < n; c++) {
        T t = a[j + c * lda];
        for (int64_t k = 0; k < j; k++) t -= CONJ(a[k + j * lda]) * a[k + c * lda];
        a[j + c * lda] = t / ajj;
      }
    } else {
      for (int64_t i = j + 1; i < n; i++) {
        T t = a[i + j * lda];
        for (int64_t k = 0; k < j; k++) t -= a[i + k * lda] * CONJ(a[j + k * lda]);
        a[i + j * lda] = t / ajj;
      }
    }
  }
}

/* xPOTRS: solve A X = B from the Cholesky factor. */
void PFX(potrs_)(const char *uplo, const int *n_, const int *nrhs_, const T *a, const int *lda_,
                 T *b, const int *ldb_, int *info) {
  int64_t n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
  int upper = PFX(lsame)(*uplo, 'U');
  *info = 0;
  if (!upper && !PFX(lsame)(*uplo, 'L')) { *info = -1; return; }
  if (n < 0) { *info = -2; return; }
  if (nrhs < 0) { *info = -3; return; }
  if (lda < (n > 1 ? n : 1)) { *info = -5; return; }
  if (ldb < (n > 1 ? n : 1)) { *info = -7; return; }
  for (int64_t c = 0; c < nrhs; c++) {
    T *x = b + c * ldb;
    if (upper) {
      for (int64_t j = 0; j < n; j++) {          /* U^H y = b */
        T t = x[j];
        for (int64_t i = 0; i < j; i++) t -= CONJ(a[i + j * lda]) * x[i];
        x[j] = t / CONJ(a[j + j * lda]);
      }
      for (int64_t k = n - 1; k >= 0; k--) {     /* U x = y */
        x[k] /= a[k + k * lda];
        for (int64_t i = 0; i < k; i++) x[i] -= x[k] * a[i + k * lda];
      }
    } else {
      for (int64_t k = 0; k < n; k++) {          /* L y = b */
        x[k] /= a[k + k * lda];
        for (int64_t i = k + 1; i < n; i++) x[i] -= x[k] * a[i + k * lda];
      }
      for (int64_t j = n - 1; j >= 0; j--) {     /* L^H x = y */
        T t = x[j];
        for (int64_t i = j + 1; i < n; i++) t -= CONJ(a[i + j * lda]) * x[i];
        x[j] = t / CONJ(a[j + j * lda]);
      }
    }
  }
}


/* ------------------------------------------------------------------ "next" rows of SURVEY section 8f
 * xSYRK  <- mm-hermitian%  src/linear-algebra.lisp:62-76   (real types; LLA sends herk for complex)
 * xTRSM  <- trsm%          src/linear-algebra.lisp:333-347
 * xPOSV  <- solve (hermitian-matrix)  src/linear-algebra.lisp:303-315  (= xPOTRF + xPOTRS)
 * Reference BLAS loop orders (netlib dsyrk.f / dtrsm.f). */

/* C <- alpha A A^T + beta C (trans = 'N', A n x k) or alpha A^T A + beta C (trans = 'T'/'C', A k x n); only the
 * uplo triangle of C is referenced. */
void PFX(syrk_)(const char *uplo, const char *trans, const int *n_, const int *k_, const T *alpha_, const T *a,
                const int *lda_, const T *beta_, T *c, const int *ldc_) {
  int64_t n = *n_, k = *k_, lda = *lda_, ldc = *ldc_;
  T alpha = *alpha_, beta = *beta_;
  int upper = PFX(lsame)(*uplo, 'U'), notr = PFX(lsame)(*trans, 'N');
  for (int64_t j = 0; j < n; j++) {
    int64_t i0 = upper ? 0 : j, i1 = upper ? j + 1 : n;
    for (int64_t i = i0; i < i1; i++) {
      T s = 0;
      if (notr) { for (int64_t l = 0; l < k; l++) s += a[i + l * lda] * a[j + l * lda]; }
      else { for (int64_t l = 0; l < k; l++) s += a[l + i * lda] * a[l + j * lda]; }
      c[i + j * ldc] = (beta == (T)0) ? alpha * s : alpha * s + beta * c[i + j * ldc];
    }
  }
}

/* xHERK <- mm-hermitian% on complex arrays, src/linear-algebra.lisp:62-76 ("syrk" "herk"): C <- alpha A A^H + beta C
 * (trans = 'N') or alpha A^H A + beta C (trans = 'C'), alpha / beta REAL, only the uplo triangle referenced, the
 * diagonal kept real (netlib zherk.f loop order: per column, the conjugated factor first).  For the real types the
 * symbol is an alias of SYRK. */
void PFX(herk_)(const char *uplo, const char *trans, const int *n_, const int *k_, const R *alpha_, const T *a,
                const int *lda_, const R *beta_, T *c, const int *ldc_) {
  int64_t n = *n_, k = *k_, lda = *lda_, ldc = *ldc_;
  R alpha = *alpha_, beta = *beta_;
  int upper = PFX(lsame)(*uplo, 'U'), notr = PFX(lsame)(*trans, 'N');
  for (int64_t j = 0; j < n; j++) {
    int64_t i0 = upper ? 0 : j, i1 = upper ? j + 1 : n;
    for (int64_t i = i0; i < i1; i++) {
      T s = 0;
      if (notr) { for (int64_t l = 0; l < k; l++) s += a[i + l * lda] * CONJ(a[j + l * lda]); }
      else { for (int64_t l = 0; l < k; l++) s += CONJ(a[l + i * lda]) * a[l + j * lda]; }
      T v = (beta == (R)0) ? alpha * s : alpha * s + beta * c[i + j * ldc];
      if (i == j) v = REAL(v);
      c[i + j * ldc] = v;
    }
  }
}

/* op(A) X = alpha B (side 'L') or X op(A) = alpha B (side 'R'); B (m x n) is overwritten by X. */
void PFX(trsm_)(const char *side, const char *uplo, const char *transa, const char *diag, const int *m_, const int *n_,
                const T *alpha_, const T *a, const int *lda_, T *b, const int *ldb_) {
  int64_t m = *m_, n = *n_, lda = *lda_, ldb = *ldb_;
  T alpha = *alpha_;
  int left = PFX(lsame)(*side, 'L'), upper = PFX(lsame)(*uplo, 'U');
  int tr = PFX(lsame)(*transa, 'N') ? 0 : (PFX(lsame)(*transa, 'T') ? 1 : 2);
  int nounit = PFX(lsame)(*diag, 'N');
  int64_t na = left ? m : n;
  /* E = op(A) as an explicit na x na triangle accessor */
#define OPA(i, j) (tr == 0 ? a[(i) + (j) * lda] : (tr == 1 ? a[(j) + (i) * lda] : CONJ(a[(j) + (i) * lda])))
  int eff_upper = (tr == 0) ? upper : !upper;
  for (int64_t j = 0; j < n; j++)
    for (int64_t i = 0; i < m; i++) b[i + j * ldb] *= alpha;
  if (left) {
    for (int64_t j = 0; j < n; j++) {
      if (eff_upper) {
        for (int64_t i = m - 1; i >= 0; i--) {
          T s = b[i + j * ldb];
          for (int64_t l = i + 1; l < m; l++) s -= OPA(i, l) * b[l + j * ldb];
          b[i + j * ldb] = nounit ? s / OPA(i, i) : s;
        }
      } else {
        for (int64_t i = 0; i < m; i++) {
          T s = b[i + j * ldb];
          for (int64_t l = 0; l < i; l++) s -= OPA(i, l) * b[l + j * ldb];
          b[i + j * ldb] = nounit ? s / OPA(i, i) : s;
        }
      }
    }
  } else {
    /* X E = B  <=>  for each row i of X: E^T x = b */
    for (int64_t i = 0; i < m; i++) {
      if (eff_upper) {   /* E upper: x_j = (b_j - sum_{l<j} x_l E(l,j)) / E(j,j) */
        for (int64_t j = 0; j < n; j++) {
          T s = b[i + j * ldb];
          for (int64_t l = 0; l < j; l++) s -= b[i + l * ldb] * OPA(l, j);
          b[i + j * ldb] = nounit ? s / OPA(j, j) : s;
        }
      } else {
        for (int64_t j = n - 1; j >= 0; j--) {
          T s = b[i + j * ldb];
          for (int64_t l = j + 1; l < n; l++) s -= b[i + l * ldb] * OPA(l, j);
          b[i + j * ldb] = nounit ? s / OPA(j, j) : s;
        }
      }
    }
  }
  (void)na;
#undef OPA
}

void PFX(posv_)(const char *uplo, const int *n_, const int *nrhs_, T *a, const int *lda_, T *b, const int *ldb_, int *info) {
  PFX(potrf_)(uplo, n_, a, lda_, info);
  if (*info == 0) PFX(potrs_)(uplo, n_, nrhs_, a, lda_, b, ldb_, info);
}

/* ------------------------------------------------------------------ the invert family
 * xTRTRI (as xTRTI2) <- invert-triangular%  src/linear-algebra.lisp:381-391
 * xPOTRI (= xTRTRI + xLAUUM) <- invert (cholesky)  :399-406
 * xGETRI (unblocked)  <- invert (lu)  :358-367, LWORK = -1 answers the workspace query of lapack-call-w/query */
void PFX(trtri_)(const char *uplo, const char *diag, const int *n_, T *a, const int *lda_, int *info) {
  int64_t n = *n_, lda = *lda_;
  int upper = PFX(lsame)(*uplo, 'U'), nounit = PFX(lsame)(*diag, 'N');
  *info = 0;
  if (nounit)
    for (int64_t i = 0; i < n; i++)
      if (a[i + i * lda] == (T)0) { *info = (int)(i + 1); return; }
  if (upper) {
    for (int64_t j = 0; j < n; j++) {
      T ajj;
      if (nounit) { a[j + j * lda] = (T)1 / a[j + j * lda]; ajj = -a[j + j * lda]; } else ajj = -(T)1;
      /* x = T(0:j,0:j) * a(0:j, j)  (xTRMV upper, no transpose: ascending i reads only not-yet-overwritten entries) */
      for (int64_t i = 0; i < j; i++) {
        T s = nounit ? a[i + i * lda] * a[i + j * lda] : a[i + j * lda];
        for (int64_t l = i + 1; l < j; l++) s += a[i + l * lda] * a[l + j * lda];
        a[i + j * lda] = s;
      }
      for (int64_t i = 0; i < j; i++) a[i + j * lda] *= ajj;
    }
  } else {
    for (int64_t j = n - 1; j >= 0; j--) {
      T ajj;
      if (nounit) { a[j + j * lda] = (T)1 / a[j + j * lda]; ajj = -a[j + j * lda]; } else ajj = -(T)1;
      for (int64_t i = n - 1; i > j; i--) {
        T s = nounit ? a[i + i * lda] * a[i + j * lda] : a[i + j * lda];
        for (int64_t l = j + 1; l < i; l++) s += a[i + l * lda] * a[l + j * lda];
        a[i + j * lda] = s;
      }
      for (int64_t i = j + 1; i < n; i++) a[i + j * lda] *= ajj;
    }
  }
}

void PFX(potri_)(const char *uplo, const int *n_, T *a, const int *lda_, int *info) {
  int64_t n = *n_, lda = *lda_;
  int upper = PFX(lsame)(*uplo, 'U');
  char nn = 'N';
  PFX(trtri_)(uplo, &nn, n_, a, lda_, info);
  if (*info != 0) return;
  /* xLAUUM: upper: A <- U U^H, lower: A <- L^H L, in place, only the named triangle */
  if (upper) {
    for (int64_t i = 0; i < n; i++)
      for (int64_t j = i; j < n; j++) {
        T s = 0;
        for (int64_t l = j; l < n; l++) s += a[i + l * lda] * CONJ(a[j + l * lda]);
        a[i + j * lda] = s;      /* row i, ascending j: entries (i, l >= j) are still U's when read */
      }
  } else {
    for (int64_t j = 0; j < n; j++)
      for (int64_t i = j; i < n; i++) {
        T s = 0;
        for (int64_t l = i; l < n; l++) s += CONJ(a[l + i * lda]) * a[l + j * lda];
        a[i + j * lda] = s;
      }
  }
}

void PFX(getri_)(const int *n_, T *a, const int *lda_, const int *ipiv, T *work, const int *lwork_, int *info) {
  int64_t n = *n_, lda = *lda_, lwork = *lwork_;
  *info = 0;
  work[0] = (T)(n > 1 ? n : 1);
  if (lwork == -1 || n == 0) return;
  char uu = 'U', nn = 'N';
  PFX(trtri_)(&uu, &nn, n_, a, lda_, info);          /* inv(U); INFO > 0: U singular */
  if (*info > 0) return;
  for (int64_t j = n - 1; j >= 0; j--) {              /* solve inv(A) L = inv(U), column by column from the right */
    for (int64_t i = j + 1; i < n; i++) { work[i] = a[i + j * lda]; a[i + j * lda] = 0; }
    for (int64_t l = j + 1; l < n; l++)
      for (int64_t i = 0; i < n; i++) a[i + j * lda] -= a[i + l * lda] * work[l];
  }
  for (int64_t j = n - 2; j >= 0; j--) {              /* column interchanges undo the row interchanges of the factorisation */
    int64_t jp = ipiv[j] - 1;
    if (jp != j)
      for (int64_t i = 0; i < n; i++) { T tmp = a[i + j * lda]; a[i + j * lda] = a[i + jp * lda]; a[i + jp * lda] = tmp; }
  }
}

#undef ABS1
#undef CONJ
#undef REAL
