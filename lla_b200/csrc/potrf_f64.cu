// lla_b200/csrc/potrf_f64.cu -- DPOTRF / DPOTRS on the device (column-major).
//
// Replaces LLA's `dpotrf_` (cholesky, reference src/linear-algebra.lisp:670-674: UPLO='U' on the
// untransposed row-major buffer, which yields the row-major lower factor) and `dpotrs_`
// (solve on a cholesky object, :296-301).
//
// Recursive formulation (as LAPACK's DPOTRF2): A11 = U11^T U11 ; U12 = U11^-T A12 (TRSM) ;
// A22 -= U12^T U12 (DMMA GEMM restricted to the upper triangle) ; recurse on A22.  Only the named
// triangle is ever written.  A non-positive or NaN pivot stores INFO = its 1-based index in a
// device flag that every later kernel of the call checks, so, as in LAPACK, nothing after the
// failing minor is modified.  UPLO='L' runs the same code on the in-place transpose.
#include "internal.h"
#include "../../include/llacuda.h"

namespace llacuda {

constexpr int CL = 32;  // leaf size

// one CTA of 32x32 threads factors the upper triangle of an n x n (n <= 32) block
__global__ void __launch_bounds__(CL * CL) lla_dpotf2_leaf_kernel(int n, double* __restrict__ A, int64_t lda, int off,
                                                              int* __restrict__ info) {
  if (*info != 0) return;
  __shared__ double s[CL][CL + 1];
  __shared__ int s_fail;
  const int i = threadIdx.x, j = threadIdx.y;  // element (i, j), i = row
  const bool in = i < n && j < n && i <= j;
  s[i][j] = in ? A[i + (int64_t)j * lda] : 0.0;
  if (i == 0 && j == 0) s_fail = 0;
  __syncthreads();
  for (int k = 0; k < n; k++) {
    double d = s[k][k];
    if (!(d > 0.0)) {  // also catches NaN, as DPOTF2's  ajj <= 0 .or. disnan(ajj)
      if (i == 0 && j == 0) s_fail = k + 1;
      break;  // d is read by every thread from the same location: uniform exit
    }
    double r = sqrt(d);
    __syncthreads();
    if (i == k && j == k) s[k][k] = r;
    if (i == k && j > k && j < n) s[k][j] = s[k][j] / r;
    __syncthreads();
    if (i > k && j >= i && j < n) s[i][j] = fma(-s[k][i], s[k][j], s[i][j]);
    __syncthreads();
  }
  __syncthreads();
  if (in) A[i + (int64_t)j * lda] = s[i][j];
  if (i == 0 && j == 0 && s_fail != 0) *info = off + s_fail;
}

static int potrf_rec(int64_t n, double* A, int64_t lda, int64_t off, int* info, cudaStream_t st) {
  if (n <= 0) return 0;
  if (n <= CL) {
    lla_dpotf2_leaf_kernel<<<1, dim3(CL, CL), 0, st>>>((int)n, A, lda, (int)off, info);
    count_launch();
    LLA_CUDA(cudaGetLastError());
    return 0;
  }
  int64_t n1 = ((n / 2 + CL - 1) / CL) * CL;
  int64_t n2 = n - n1;
  double* A12 = A + n1 * lda;
  double* A22 = A12 + n1;
  LLA_TRY(potrf_rec(n1, A, lda, off, info, st));
  LLA_TRY(dtrsm_left_dev(true, 1, false, n1, n2, A, lda, A12, lda, st, info));
  LLA_TRY(dgemm_dev(1, 0, n2, n2, n1, -1.0, A12, lda, A12, lda, 1.0, A22, lda, TRI_UPPER, st, info));
  LLA_TRY(potrf_rec(n2, A22, lda, off + n1, info, st));
  return 0;
}

int dpotrf_dev(bool upper, int64_t n, double* A, int64_t lda, int* info_dev, cudaStream_t st) {
  if (!ctx()) return LLACUDA_ERR_NO_DEVICE;
  LLA_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), st));
  if (n <= 0) return 0;
  if (!upper) LLA_TRY(dtranspose_inplace_dev(n, A, lda, st));
  LLA_TRY(potrf_rec(n, A, lda, 0, info_dev, st));
  if (!upper) LLA_TRY(dtranspose_inplace_dev(n, A, lda, st));
  return 0;
}

int dpotrs_dev(bool upper, int64_t n, int64_t nrhs, const double* A, int64_t lda, double* B, int64_t ldb,
               cudaStream_t st, const int* abort_flag) {
  if (n <= 0 || nrhs <= 0) return 0;
  if (upper) {  // A = U^T U :  U^T y = b ; U x = y
    LLA_TRY(dtrsm_left_dev(true, 1, false, n, nrhs, A, lda, B, ldb, st, abort_flag));
    LLA_TRY(dtrsm_left_dev(true, 0, false, n, nrhs, A, lda, B, ldb, st, abort_flag));
  } else {      // A = L L^T :  L y = b ; L^T x = y
    LLA_TRY(dtrsm_left_dev(false, 0, false, n, nrhs, A, lda, B, ldb, st, abort_flag));
    LLA_TRY(dtrsm_left_dev(false, 1, false, n, nrhs, A, lda, B, ldb, st, abort_flag));
  }
  return 0;
}

}  // namespace llacuda
