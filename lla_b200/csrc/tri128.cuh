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
  // tile t owns rows {ti, ti + TM, ti + 2 TM, ti + 3 TM} and columns 4 tj .. 4 tj + 3: consecutive lanes read
  // consecutive rows of A (conflict-free 64-bit shared loads) and broadcast-read B
  constexpr int TM = M / 4, TN = N / 4;
  for (int t = tid; t < TM * TN; t += nth) {
    const int ti = t % TM, j0 = (t / TM) * 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
#pragma unroll 4
    for (int k = 0; k < K; k++) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; a++) av[a] = A[ti + a * TM + k * lda];
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
      for (int b = 0; b < 4; b++) C[ti + a * TM + (j0 + b) * ldc] = alpha * acc[a][b];
  }
}

// S: 128 x 128 lower-triangular matrix (strict upper part zero), pitch T128_SP; tmp: 64 x 64 doubles.
// On return S holds inv(S).  inv([L11 0; L21 L22]) = [X11 0; -X22 L21 X11, X22], applied at the
// 32 -> 64 -> 128 levels; the 32 x 32 diagonal blocks are inverted by forward substitution on the
// identity, one thread per column with the column in registers.
// `have_rdiag`: tmp[0..127] already holds the reciprocals of the diagonal (the Cholesky leaf leaves them
// there); otherwise they are computed first, so that the substitution chain multiplies instead of dividing.
__device__ __forceinline__ void trtri128_lower_smem(double* S, double* tmp, bool unit, int tid, bool have_rdiag = false) {
  if (!have_rdiag) {
    if (tid < 128) tmp[tid] = unit ? 1.0 : 1.0 / S[tid + tid * T128_SP];
    __syncthreads();
  }
  // ---- 32 x 32 diagonal blocks
  double x[32];
  const int b = tid >> 5, c = tid & 31;
  if (tid < 128) {
    const double* D = S + (32 * b) + (32 * b) * T128_SP;
    const double* rd = tmp + 32 * b;
#pragma unroll
    for (int i = 0; i < 32; i++) x[i] = (i == c) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 32; k++) {
      x[k] = x[k] * rd[k];
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
  constexpr int PER = T128 * T128 / T128_THREADS;  // 32 elements per thread, 8 loads in flight at a time
#pragma unroll 1
  for (int q0 = 0; q0 < PER; q0 += 8) {
    double v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int idx = tid + (q0 + q) * T128_THREADS;
      const int i = idx & 127, j = idx >> 7;  // consecutive threads run down a column of T
      const bool in_tri = src_upper ? (i <= j) : (i >= j);
      v[q] = 0.0;
      if (i < r && j < r && in_tri && !(unit && i == j)) v[q] = T[i + (int64_t)j * ldt];
      if (i == j && (unit || i >= r)) v[q] = 1.0;
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int idx = tid + (q0 + q) * T128_THREADS;
      const int i = idx & 127, j = idx >> 7;
      if (src_upper) S[j + i * T128_SP] = v[q];   // S = T^T
      else S[i + j * T128_SP] = v[q];
    }
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

// One warp factors the 32 x 32 diagonal block at (c0, c0) of S in registers (lane = row) and leaves 1 / l_kk in
// rdiag[c0 ..].  Column k of L travels through shared memory: every lane stores its l_ik, everybody reads the l_ck it
// needs back as 16-byte broadcasts (one LSU instruction per two columns where a 64-bit shuffle costs two per column);
// the next pivot's reciprocal square root is in flight while the remaining columns of the current step update.
// lbuf: 2 x 32 doubles (double-buffered by the parity of k, so one __syncwarp per pivot).  Returns 0 or c0 + k + 1.
__device__ __forceinline__ int potrf32_diag_warp(double* S, int c0, double* rdiag, double* lbuf, int lane) {
  double a[32];
  double* D = S + (c0 + lane) + c0 * T128_SP;  // row c0 + lane
#pragma unroll
  for (int c = 0; c < 32; c++) a[c] = D[c * T128_SP];
  int fail = 0;
  double dmine = 1.0;  // lane k keeps d_k; its square root is taken once, after the loop
  double d = __shfl_sync(0xffffffffu, a[0], 0);
  double rs = rsqrt(d);
#pragma unroll
  for (int k = 0; k < 32; k++) {
    if (!(d > 0.0)) { fail = c0 + k + 1; break; }
    if (lane == k) dmine = d;
    if (lane == 0) rdiag[c0 + k] = rs;
    const double l = a[k] * rs;
    a[k] = l;
    double dn = 1.0, rsn = 1.0;
    if (k + 1 < 32) {
      double* lb = lbuf + (k & 1) * 32;
      lb[lane] = l;
      // next pivot first, without waiting for the broadcast: lane k + 1 needs only its own l for d_{k+1} = a - l * l;
      // its reciprocal square root is in flight under the updates of the other columns
      dn = __shfl_sync(0xffffffffu, fma(-l, l, a[k + 1]), k + 1);
      rsn = rsqrt(dn);
      __syncwarp();
      a[k + 1] = fma(-l, lb[k + 1], a[k + 1]);
      int c = k + 2;
      if (c < 32 && (c & 1)) { a[c] = fma(-l, lb[c], a[c]); c++; }
#pragma unroll
      for (; c < 32; c += 2) {  // meaningful for rows >= c only
        const double2 v = *reinterpret_cast<const double2*>(lb + c);
        a[c] = fma(-l, v.x, a[c]);
        a[c + 1] = fma(-l, v.y, a[c + 1]);
      }
    }
    d = dn; rs = rsn;
  }
  if (fail) return fail;
  const double diag = sqrt(dmine);  // the diagonal itself (lane k's a[k] holds d_k * rs, a rounding away from it)
#pragma unroll
  for (int c = 0; c < 32; c++)
    if (c <= lane) D[c * T128_SP] = (c == lane) ? diag : a[c];
  return 0;
}

// Rank-32 update S[i0.., j0..] -= sum_k S[i0.., c0 + k] S[j0.., c0 + k] of one 4 x 4 tile of the lower triangle
// (entries above the diagonal of a diagonal tile are left alone).  i0, j0 are multiples of 4: 16-byte loads.
__device__ __forceinline__ void potrf128_tile_update(double* S, int c0, int i0, int j0) {
  double acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; u++)
#pragma unroll
    for (int v = 0; v < 4; v++) acc[u][v] = 0.0;
#pragma unroll 8
  for (int k = 0; k < 32; k++) {
    const double2* ap = reinterpret_cast<const double2*>(S + i0 + (c0 + k) * T128_SP);
    const double2* bp = reinterpret_cast<const double2*>(S + j0 + (c0 + k) * T128_SP);
    const double2 a01 = ap[0], a23 = ap[1], b01 = bp[0], b23 = bp[1];
    const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int v = 0; v < 4; v++) acc[u][v] = fma(av[u], bv[v], acc[u][v]);
  }
#pragma unroll
  for (int u = 0; u < 4; u++)
#pragma unroll
    for (int v = 0; v < 4; v++)
      if (i0 + u >= j0 + v) S[i0 + u + (j0 + v) * T128_SP] -= acc[u][v];
}

// In-place lower Cholesky S = L L^T of the leading r x r block (S is identity-padded to 128 x 128),
// right-looking over 32-column block steps with the diagonal factorisations off the critical path:
//   (1) one thread per row below the block solves its row against L11^T by substitution (L11 read as broadcasts);
//   (2) the rank-32 update of the NEXT diagonal block (36 tiles of 4 x 4);
//   (3) warp 0 factors that next diagonal block in registers (potrf32_diag_warp) WHILE the other 15 warps apply the
//       rank-32 update to the rest of the trailing block (tiles enumerated over the trapezoid below it).
// Returns 0, or k+1 for the first non-positive / NaN pivot (uniform across the CTA).  rdiag: 128 doubles.
__device__ __forceinline__ int potrf128_lower_smem(double* S, double* rdiag, int r, int tid) {
  __shared__ int s_fail;
  __shared__ __align__(16) double s_lbuf[64];
  const int lane = tid & 31, warp = tid >> 5;
  if (tid < 128) rdiag[tid] = 1.0;
  if (tid == 0) s_fail = 0;
  __syncthreads();
  if (warp == 0) {
    const int fail = potrf32_diag_warp(S, 0, rdiag, s_lbuf, lane);
    if (fail && lane == 0) s_fail = fail;
  }
  __syncthreads();
  for (int c0 = 0; c0 < r; c0 += 32) {
    if (s_fail) return s_fail;
    const int t0 = c0 + 32;          // first row / column of the trailing block
    const int nbelow = T128 - t0;
    if (nbelow <= 0) break;
    if (tid >= 32 && tid < 32 + nbelow) {
      double x[32];
      double* R = S + (t0 + tid - 32) + c0 * T128_SP;
      const double* L11 = S + c0 + c0 * T128_SP;
#pragma unroll
      for (int c = 0; c < 32; c++) x[c] = R[c * T128_SP];
#pragma unroll
      for (int k = 0; k < 32; k++) {
        x[k] *= rdiag[c0 + k];
        const double* col = L11 + k * T128_SP;   // column k of L11: 16-byte aligned (pitch 130, c0 a multiple of 32)
        int c = k + 1;
        if (c < 32 && (c & 1)) { x[c] = fma(-x[k], col[c], x[c]); c++; }
#pragma unroll
        for (; c < 32; c += 2) {
          const double2 v = *reinterpret_cast<const double2*>(col + c);
          x[c] = fma(-x[k], v.x, x[c]);
          x[c + 1] = fma(-x[k], v.y, x[c + 1]);
        }
      }
#pragma unroll
      for (int c = 0; c < 32; c++) R[c * T128_SP] = x[c];
    }
    __syncthreads();
    // the next diagonal block first: tiles (ti >= tj) of its 8 x 8 tile triangle
    if (tid < 36) {
      int ti = 0, t = tid;
      while (t > ti) { t -= ti + 1; ti++; }
      potrf128_tile_update(S, c0, t0 + 4 * ti, t0 + 4 * t);
    }
    __syncthreads();
    if (warp == 0) {
      if (t0 < r) {   // beyond r the block is identity padding
        const int fail = potrf32_diag_warp(S, t0, rdiag, s_lbuf, lane);
        if (fail && lane == 0) s_fail = fail;
      }
    } else {
      // rows t0 + 32 .. 127 of the trailing block: tile rows ti = 8 .. T - 1, tile columns 0 .. ti
      const int T = nbelow / 4;
      const int count = (T * (T + 1)) / 2 - 36;
      for (int q = tid - 32; q < count; q += T128_THREADS - 32) {
        int ti = 8, t = q;
        while (t > ti) { t -= ti + 1; ti++; }
        potrf128_tile_update(S, c0, t0 + 4 * ti, t0 + 4 * t);
      }
    }
    __syncthreads();
  }
  return s_fail;
}

}  // namespace llacuda
