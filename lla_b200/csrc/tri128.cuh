// lla_b200/csrc/tri128.cuh -- shared-memory building blocks for 128 x 128 triangular work:
// in-place inversion of a lower-triangular block (used for the diagonal blocks that turn every
// large triangular solve into DMMA GEMMs) and an in-place lower Cholesky factorisation.
// All routines are called by a whole CTA of T128_THREADS threads.
#pragma once
#include <cuda_runtime.h>

namespace llacuda {

constexpr int T128 = 128;
constexpr int T128_SP = 130;       // column pitch (doubles) of the 128 x 128 shared-memory image
constexpr int T128_THREADS = 512;
constexpr size_t T128_SMEM = (size_t)(T128 * T128_SP + 64 * 64) * sizeof(double);

// C (M x N, pitch ldc) = alpha * A (M x K, pitch lda) * B (K x N, pitch ldb), all column-major in
// shared memory; 4 x 4 register tiles handed out round-robin to `nth` threads.
template <int M, int N, int K>
__device__ __forceinline__ void smem_mm(double* C, int ldc, const double* A, int lda, const double* B, int ldb,
                                        double alpha, int tid, int nth) {
  constexpr int TM = M / 4, TN = N / 4;
  for (int t = tid; t < TM * TN; t += nth) {
    const int i0 = (t % TM) * 4, j0 = (t / TM) * 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
#pragma unroll 4
    for (int k = 0; k < K; k++) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; a++) av[a] = A[i0 + a + k * lda];
#pragma unroll
      for (int b = 0; b < 4; b++) bv[b] = B[k + (j0 + b) * ldb];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) C[i0 + a + (j0 + b) * ldc] = alpha * acc[a][b];
  }
}

// S: 128 x 128 lower-triangular matrix (strict upper part zero), pitch T128_SP; tmp: 64 x 64 doubles.
// On return S holds inv(S).  inv([L11 0; L21 L22]) = [X11 0; -X22 L21 X11, X22], applied at the
// 32 -> 64 -> 128 levels; the 32 x 32 diagonal blocks are inverted by forward substitution on the
// identity, one thread per column with the column in registers.
__device__ __forceinline__ void trtri128_lower_smem(double* S, double* tmp, bool unit, int tid) {
  // ---- 32 x 32 diagonal blocks
  double x[32];
  const int b = tid >> 5, c = tid & 31;
  if (tid < 128) {
    const double* D = S + (32 * b) + (32 * b) * T128_SP;
#pragma unroll
    for (int i = 0; i < 32; i++) x[i] = (i == c) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 32; k++) {
      if (!unit) x[k] = x[k] / D[k + k * T128_SP];
#pragma unroll
      for (int i = k + 1; i < 32; i++) x[i] = fma(-D[i + k * T128_SP], x[k], x[i]);
    }
  }
  __syncthreads();
  if (tid < 128) {
    double* D = S + (32 * b) + (32 * b) * T128_SP;
#pragma unroll
    for (int i = 0; i < 32; i++) D[i + c * T128_SP] = x[i];
  }
  __syncthreads();
  // ---- level 64: block pairs (1,0) and (3,2), 128 threads per pair
  if (tid < 256) {
    const int p = tid >> 7, t = tid & 127;
    const double* X21 = S + (64 * p + 32) + (64 * p) * T128_SP;
    const double* X11 = S + (64 * p) + (64 * p) * T128_SP;
    smem_mm<32, 32, 32>(tmp + p * 32 * 32, 32, X21, T128_SP, X11, T128_SP, 1.0, t, 128);
  }
  __syncthreads();
  if (tid < 256) {
    const int p = tid >> 7, t = tid & 127;
    double* X21 = S + (64 * p + 32) + (64 * p) * T128_SP;
    const double* X22 = S + (64 * p + 32) + (64 * p + 32) * T128_SP;
    smem_mm<32, 32, 32>(X21, T128_SP, X22, T128_SP, tmp + p * 32 * 32, 32, -1.0, t, 128);
  }
  __syncthreads();
  // ---- level 128
  smem_mm<64, 64, 64>(tmp, 64, S + 64, T128_SP, S, T128_SP, 1.0, tid, T128_THREADS);
  __syncthreads();
  smem_mm<64, 64, 64>(S + 64, T128_SP, S + 64 + 64 * T128_SP, T128_SP, tmp, 64, -1.0, tid, T128_THREADS);
  __syncthreads();
}

// Loads the r x r triangle of T (lower, or upper read transposed) into S as a LOWER matrix padded
// with the identity to 128 x 128; the strict upper part of S is zeroed.
__device__ __forceinline__ void tri128_load_lower(double* S, const double* __restrict__ T, int64_t ldt, int r,
                                                  bool src_upper, bool unit, int tid) {
  for (int idx = tid; idx < T128 * T128; idx += T128_THREADS) {
    const int i = idx & 127, j = idx >> 7;  // consecutive threads run down a column of T
    double v = 0.0;
    const bool in_tri = src_upper ? (i <= j) : (i >= j);
    if (i < r && j < r && in_tri && !(unit && i == j)) v = T[i + (int64_t)j * ldt];
    if (i == j && (unit || i >= r)) v = 1.0;
    if (src_upper) S[j + i * T128_SP] = v;   // S = T^T
    else S[i + j * T128_SP] = v;
  }
}

// Writes the lower matrix S to out (leading dimension ldo) as lower, or transposed as upper; the
// other triangle of the r x r block is written with zeros when `zero_other` is set.
__device__ __forceinline__ void tri128_store(const double* S, double* __restrict__ out, int64_t ldo, int r,
                                             bool dst_upper, bool zero_other, int tid) {
  for (int idx = tid; idx < T128 * T128; idx += T128_THREADS) {
    const int i = idx & 127, j = idx >> 7;
    if (i >= r || j >= r) continue;
    const bool in_tri = dst_upper ? (i <= j) : (i >= j);
    if (in_tri) out[i + (int64_t)j * ldo] = dst_upper ? S[j + i * T128_SP] : S[i + j * T128_SP];
    else if (zero_other) out[i + (int64_t)j * ldo] = 0.0;
  }
}

// In-place lower Cholesky S = L L^T of the leading r x r block, left-looking in 32-column panels.
// A panel first receives the updates of all earlier columns (4 x 4 register tiles), then is
// factored column by column straight in shared memory: thread (i, h) owns row i and every 4th
// column of the panel.  One barrier per column: column g is scaled on the fly (x * rsqrt(d)) and its
// scaled values are written back one step later, when nobody reads the unscaled ones any more.
// These loops are latency bound (a few warps on one SM), hence the low instruction count per step.
// Returns 0, or k+1 for the first non-positive / NaN pivot (uniform across the CTA).
__device__ __forceinline__ int potrf128_lower_smem(double* S, int r, int tid) {
  const int i = tid & 127, h = tid >> 7;  // T128_THREADS = 512: h in 0..3
  for (int c0 = 0; c0 < r; c0 += 32) {
    if (c0 > 0) {
      // S[c0.., c0..c0+31] -= L[c0.., 0..c0) * L[c0..c0+31, 0..c0)^T
      const int tiles_m = (T128 - c0) / 4;
      for (int t = tid; t < tiles_m * 8; t += T128_THREADS) {
        const int i0 = c0 + (t % tiles_m) * 4, j0 = c0 + (t / tiles_m) * 4;
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
#pragma unroll 4
        for (int k = 0; k < c0; k++) {
          double av[4], bv[4];
#pragma unroll
          for (int a = 0; a < 4; a++) av[a] = S[i0 + a + k * T128_SP];
#pragma unroll
          for (int b = 0; b < 4; b++) bv[b] = S[j0 + b + k * T128_SP];
#pragma unroll
          for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int b = 0; b < 4; b++)
            if (i0 + a >= j0 + b) S[i0 + a + (j0 + b) * T128_SP] -= acc[a][b];
      }
    }
    __syncthreads();
    double lprev = 0.0, dprev = 0.0;  // deferred write-back of the previous column
    const int kend = (r - c0) < 32 ? (r - c0) : 32;
#pragma unroll 1
    for (int k = 0; k < kend; k++) {
      const int g = c0 + k;
      if (k > 0 && h == 0) {
        if (i == g - 1) S[i + (g - 1) * T128_SP] = dprev;
        else if (i > g - 1) S[i + (g - 1) * T128_SP] = lprev;
      }
      const double d = S[g + g * T128_SP];
      if (!(d > 0.0)) return g + 1;
      const double rs = rsqrt(d);
      const double li = S[i + g * T128_SP] * rs;
      lprev = li;
      dprev = sqrt(d);
      if (i > g) {
        const int cend = (i < c0 + 31) ? i : (c0 + 31);
        for (int c = g + 1 + h; c <= cend; c += 4)
          S[i + c * T128_SP] = fma(-li, S[c + g * T128_SP] * rs, S[i + c * T128_SP]);
      }
      __syncthreads();
    }
    if (h == 0) {
      const int g = c0 + kend - 1;
      if (i == g) S[i + g * T128_SP] = dprev;
      else if (i > g) S[i + g * T128_SP] = lprev;
    }
    __syncthreads();
  }
  return 0;
}

}  // namespace llacuda
