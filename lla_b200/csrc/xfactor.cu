// lla_b200/csrc/xfactor.cu -- LU / Cholesky / solves / rank-k updates for the single, complex-single and
// complex-double types: {s,c,z}getrf_ / getrs_ / gesv_ / potrf_ / potrs_ / posv_, ssyrk_ / cherk_ / zherk_.
//
// LLA's lu / solve / cholesky / mm-hermitian% dispatch on the common float type of their operands
// (reference src/linear-algebra.lisp:62-76,225-234,264-301,666-674) through the four-way ECASE of
// src/fortran-call.lisp:428-439; the double-precision symbols have their own tuned kernels (lu_f64.cu, potrf_f64.cu),
// this file completes the set with one templated path:
//   * LU: right-looking over 32-column panels.  A panel is factored by ONE cooperative kernel (one thread per row, the
//     256 x 32 slice of a CTA in shared memory): per column a grid-wide arg-max with IxAMAX semantics -- |re| + |im| for
//     complex types (reference LAPACK IZAMAX / DCABS1), first index on ties, NaN never beats a number -- then swap,
//     reciprocal scaling as xGETF2 does, rank-1 update; the 32 interchanges go to the other columns, U12 by a 32 x 32
//     substitution leaf, the trailing matrix by a GEMM (DMMA ZGEMM for complex double, an FP32-FMA tile kernel for the
//     single types: the tcgen05 SGEMM splits into TF32 and is kept for the large products of mm / gemm!).
//   * Cholesky ('U' natively, 'L' through a conjugate transpose): 32 x 32 diagonal blocks by one CTA (xPOTF2 arithmetic:
//     real diagonal, INFO = first non-positive or NaN pivot), block row by the substitution leaf, trailing update by
//     the GEMM restricted to the upper triangle.
//   * GETRS / POTRS: 32-row substitution leaves + GEMM updates, any of N / T / C.
// Measured numbers and what bounds them: DESIGN.md section 4.
#include "internal.h"
#include "staging.h"
#include "../../include/llacuda.h"

#include <cooperative_groups.h>
#include <cfloat>
#include <cstdio>
#include <vector>

namespace llacuda {

namespace cg = cooperative_groups;

// ------------------------------------------------------------------ scalar algebra shared by the three types
typedef cuFloatComplex C32;
typedef cuDoubleComplex C64;

template <class T> struct Real { typedef T type; };
template <> struct Real<C32> { typedef float type; };
template <> struct Real<C64> { typedef double type; };

__host__ __device__ __forceinline__ float xzero(float) { return 0.f; }
__host__ __device__ __forceinline__ C32 xzero(C32) { return make_cuFloatComplex(0.f, 0.f); }
__host__ __device__ __forceinline__ C64 xzero(C64) { return make_cuDoubleComplex(0.0, 0.0); }
__host__ __device__ __forceinline__ float xreal(float a) { return a; }
__host__ __device__ __forceinline__ float xreal(C32 a) { return a.x; }
__host__ __device__ __forceinline__ double xreal(C64 a) { return a.x; }
__host__ __device__ __forceinline__ float xfrom(float, float r) { return r; }
__host__ __device__ __forceinline__ C32 xfrom(C32, float r) { return make_cuFloatComplex(r, 0.f); }
__host__ __device__ __forceinline__ C64 xfrom(C64, double r) { return make_cuDoubleComplex(r, 0.0); }
__host__ __device__ __forceinline__ float xconj(float a) { return a; }
__host__ __device__ __forceinline__ C32 xconj(C32 a) { return make_cuFloatComplex(a.x, -a.y); }
__host__ __device__ __forceinline__ C64 xconj(C64 a) { return make_cuDoubleComplex(a.x, -a.y); }
__host__ __device__ __forceinline__ float xmul(float a, float b) { return a * b; }
__host__ __device__ __forceinline__ C32 xmul(C32 a, C32 b) { return cuCmulf(a, b); }
__host__ __device__ __forceinline__ C64 xmul(C64 a, C64 b) { return cuCmul(a, b); }
__host__ __device__ __forceinline__ float xsub(float a, float b) { return a - b; }
__host__ __device__ __forceinline__ C32 xsub(C32 a, C32 b) { return cuCsubf(a, b); }
__host__ __device__ __forceinline__ C64 xsub(C64 a, C64 b) { return cuCsub(a, b); }
__host__ __device__ __forceinline__ float xadd(float a, float b) { return a + b; }
__host__ __device__ __forceinline__ C32 xadd(C32 a, C32 b) { return cuCaddf(a, b); }
__host__ __device__ __forceinline__ C64 xadd(C64 a, C64 b) { return cuCadd(a, b); }
__host__ __device__ __forceinline__ float xdiv(float a, float b) { return a / b; }
__host__ __device__ __forceinline__ C32 xdiv(C32 a, C32 b) { return cuCdivf(a, b); }
__host__ __device__ __forceinline__ C64 xdiv(C64 a, C64 b) { return cuCdiv(a, b); }
// acc + a * b with the real products fused as the FMA units do it
__device__ __forceinline__ float xfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ C32 xfma(C32 a, C32 b, C32 c) {
  return make_cuFloatComplex(fmaf(-a.y, b.y, fmaf(a.x, b.x, c.x)), fmaf(a.y, b.x, fmaf(a.x, b.y, c.y)));
}
__device__ __forceinline__ C64 xfma(C64 a, C64 b, C64 c) {
  return make_cuDoubleComplex(fma(-a.y, b.y, fma(a.x, b.x, c.x)), fma(a.y, b.x, fma(a.x, b.y, c.y)));
}
// the magnitude IxAMAX compares: |x| for real types, |re| + |im| (xCABS1) for complex ones
__host__ __device__ __forceinline__ float xabs1(float a) { return fabsf(a); }
__host__ __device__ __forceinline__ float xabs1(C32 a) { return fabsf(a.x) + fabsf(a.y); }
__host__ __device__ __forceinline__ double xabs1(C64 a) { return fabs(a.x) + fabs(a.y); }
__host__ __device__ __forceinline__ bool xiszero(float a) { return a == 0.f; }
__host__ __device__ __forceinline__ bool xiszero(C32 a) { return a.x == 0.f && a.y == 0.f; }
__host__ __device__ __forceinline__ bool xiszero(C64 a) { return a.x == 0.0 && a.y == 0.0; }
__device__ __forceinline__ float xsfmin(float) { return FLT_MIN; }
__device__ __forceinline__ double xsfmin(double) { return DBL_MIN; }
// order-preserving integer key of a non-negative magnitude (0 = "no number here": NaN or out of range)
__device__ __forceinline__ unsigned long long xkey(float v) { return v == v ? (unsigned long long)__float_as_uint(v) + 1ull : 0ull; }
__device__ __forceinline__ unsigned long long xkey(double v) { return v == v ? (unsigned long long)__double_as_longlong(v) + 1ull : 0ull; }
template <class T> __device__ __forceinline__ T xop(T v, int op) { return op == 2 ? xconj(v) : v; }

// ------------------------------------------------------------------ GEMM for the updates
// C = alpha op(A) op(B) + beta C, optionally only the upper triangle (tri = TRI_UPPER).  64 x 64 x 16 tiles, 256
// threads, 4 x 4 register tiles, FMA arithmetic in the operand precision.
template <class T>
__global__ void __launch_bounds__(256) lla_xgemm_simt_kernel(int opa, int opb, int64_t M, int64_t N, int64_t K, T alpha,
                                                             const T* __restrict__ A, int64_t lda, const T* __restrict__ B,
                                                             int64_t ldb, T beta, T* __restrict__ C, int64_t ldc, int tri,
                                                             const int* __restrict__ abort_flag) {
  if (abort_flag != nullptr && *abort_flag != 0) return;
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ T As[BK][BM + 1];
  __shared__ T Bs[BK][BN + 1];
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  if (tri == TRI_UPPER && m0 > n0 + BN - 1) return;   // tile strictly below the diagonal
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = xzero(T());
  for (int64_t k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int e = tid; e < BM * BK; e += 256) {
      int mm, kk;
      if (opa == 0) { mm = e % BM; kk = e / BM; } else { kk = e % BK; mm = e / BK; }   // the contiguous index runs fastest
      const int64_t gm = m0 + mm, gk = k0 + kk;
      T v = xzero(T());
      if (gm < M && gk < K) v = (opa == 0) ? A[gm + gk * lda] : xop(A[gk + gm * lda], opa);
      As[kk][mm] = v;
    }
#pragma unroll
    for (int e = tid; e < BN * BK; e += 256) {
      int nn, kk;
      if (opb == 0) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
      const int64_t gn = n0 + nn, gk = k0 + kk;
      T v = xzero(T());
      if (gn < N && gk < K) v = (opb == 0) ? B[gk + gn * ldb] : xop(B[gn + gk * ldb], opb);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      T av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; i++) av[i] = As[kk][tx + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; j++) bv[j] = Bs[kk][ty + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = xfma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool beta0 = xiszero(beta);
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int64_t gm = m0 + tx + 16 * i, gn = n0 + ty + 16 * j;
      if (gm >= M || gn >= N) continue;
      if (tri == TRI_UPPER && gm > gn) continue;
      T* c = C + gm + gn * ldc;
      const T ab = xmul(alpha, acc[i][j]);
      *c = beta0 ? ab : xadd(ab, xmul(beta, *c));   // beta = 0 never reads C
    }
}

template <class T>
static int xgemm_dev(int opa, int opb, int64_t m, int64_t n, int64_t k, T alpha, const T* A, int64_t lda, const T* B,
                     int64_t ldb, T beta, T* C, int64_t ldc, int tri, cudaStream_t st, const int* abort_flag) {
  if (m <= 0 || n <= 0) return 0;
  dim3 grid((unsigned)((m + 63) / 64), (unsigned)((n + 63) / 64));
  lla_xgemm_simt_kernel<T><<<grid, 256, 0, st>>>(opa, opb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri, abort_flag);
  count_launch();
  LLA_CUDA(cudaGetLastError());
  return 0;
}
// complex double goes to the DMMA ZGEMM (it has no abort flag: values after a failed Cholesky minor are unspecified, as in LAPACK)
template <>
int xgemm_dev<C64>(int opa, int opb, int64_t m, int64_t n, int64_t k, C64 alpha, const C64* A, int64_t lda, const C64* B,
                   int64_t ldb, C64 beta, C64* C, int64_t ldc, int tri, cudaStream_t st, const int*) {
  return zgemm_dev(opa, opb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri, st);
}

int sgemm_simt_dev(int opa, int opb, int64_t m, int64_t n, int64_t k, float alpha, const float* A, int64_t lda, const float* B,
                   int64_t ldb, float beta, float* C, int64_t ldc, cudaStream_t st) {
  return xgemm_dev<float>(opa, opb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, TRI_FULL, st, nullptr);
}

// ------------------------------------------------------------------ LU panel (cooperative, one thread per row)
constexpr int XW = 32;      // panel width
constexpr int XR = 256;     // rows per CTA
struct XCand { unsigned long long key; long long row; };
template <class T> constexpr size_t xpanel_smem() { return (size_t)XW * (XR + 1) * sizeof(T); }

template <class T>
__global__ void __launch_bounds__(XR) lla_xgetf2_panel_kernel(T* __restrict__ A, int64_t lda, int64_t m, int w, int64_t off,
                                                              int* __restrict__ ipiv, int* __restrict__ info,
                                                              XCand* __restrict__ cand, T* __restrict__ rowbuf) {
  typedef typename Real<T>::type R;
  extern __shared__ __align__(16) unsigned char xsm_raw[];
  T* S = reinterpret_cast<T*>(xsm_raw);          // column c of my slice at S + c * (XR + 1)
  constexpr int PITCH = XR + 1;
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  const int64_t row0 = (int64_t)blockIdx.x * XR, grow = row0 + tid;   // relative to the panel top
  const bool live = grow < m;
  const int nsteps = (int)(m < w ? m : w);
  for (int c = 0; c < w; c++) S[c * PITCH + tid] = live ? A[grow + (int64_t)c * lda] : xzero(T());
  __shared__ unsigned long long s_key[XR / 32];
  __shared__ long long s_row[XR / 32];
  __shared__ T s_u[XW], s_r[XW];
  __shared__ long long s_p;
  __syncthreads();
  for (int j = 0; j < nsteps; j++) {
    const int par = j & 1;
    // ---- candidate of this CTA: first row >= j with the largest magnitude; NaN only wins as row j (IxAMAX)
    unsigned long long key = 0ull;
    long long r = 0x7fffffffffffffffll;
    if (live && grow >= j) {
      R v = xabs1(S[j * PITCH + tid]);
      if (grow == j && !(v == v)) v = (R)INFINITY;
      key = xkey(v);
      r = grow;
    }
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long k2 = __shfl_down_sync(0xffffffffu, key, o);
      const long long r2 = __shfl_down_sync(0xffffffffu, r, o);
      if (k2 > key || (k2 == key && r2 < r)) { key = k2; r = r2; }
    }
    if (lane == 0) { s_key[warp] = key; s_row[warp] = r; }
    __syncthreads();
    if (tid == 0) {
      for (int q = 1; q < XR / 32; q++)
        if (s_key[q] > key || (s_key[q] == key && s_row[q] < r)) { key = s_key[q]; r = s_row[q]; }
      cand[par * G + blockIdx.x].key = key;
      cand[par * G + blockIdx.x].row = r;
    }
    grid.sync();
    if (warp == 0) {   // every CTA picks the same winner
      unsigned long long bk = 0ull;
      long long br = 0x7fffffffffffffffll;
      for (int q = lane; q < G; q += 32) {
        const XCand cq = cand[par * G + q];
        if (cq.key > bk || (cq.key == bk && cq.row < br)) { bk = cq.key; br = cq.row; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long k2 = __shfl_down_sync(0xffffffffu, bk, o);
        const long long r2 = __shfl_down_sync(0xffffffffu, br, o);
        if (k2 > bk || (k2 == bk && r2 < br)) { bk = k2; br = r2; }
      }
      if (lane == 0) s_p = br;
    }
    __syncthreads();
    const long long p = s_p;   // pivot row (relative to the panel top); always a live row >= j
    T* rb = rowbuf + (size_t)par * 2 * XW;
    if (p >= row0 && p < row0 + XR && tid < w) rb[tid] = S[tid * PITCH + (int)(p - row0)];          // pivot row
    if (blockIdx.x == 0 && tid < w) rb[XW + tid] = S[tid * PITCH + j];                              // the row at position j
    grid.sync();
    if (tid < w) { s_u[tid] = rb[tid]; s_r[tid] = rb[XW + tid]; }
    __syncthreads();
    const T piv = s_u[j];
    const bool nz = !xiszero(piv);
    if (blockIdx.x == 0 && tid == 0) {
      ipiv[off + j] = (int)(off + p + 1);
      if (!nz && *info == 0) *info = (int)(off + j + 1);
    }
    if (nz && p != j) {   // interchange (xGETF2 swaps only when the pivot is non-zero)
      if (p >= row0 && p < row0 + XR && tid < w) S[tid * PITCH + (int)(p - row0)] = s_r[tid];
      if (blockIdx.x == 0 && tid < w) S[tid * PITCH + j] = s_u[tid];
    }
    __syncthreads();
    if (live && grow > j) {
      T l = S[j * PITCH + tid];
      if (nz) {
        if (xabs1(piv) >= xsfmin(R())) l = xmul(l, xdiv(xfrom(T(), (R)1), piv));   // reciprocal scaling, as xGETF2 / xSCAL
        else l = xdiv(l, piv);
        S[j * PITCH + tid] = l;
      }
      // row j now holds the pivot row (a zero pivot means p == j: nothing moved)
      for (int c = j + 1; c < w; c++) S[c * PITCH + tid] = xsub(S[c * PITCH + tid], xmul(l, s_u[c]));
    }
    __syncthreads();
  }
  if (live)
    for (int c = 0; c < w; c++) A[grow + (int64_t)c * lda] = S[c * PITCH + tid];
}

// rows i <-> ipiv[i]-1 for i in [k1, k2) on the columns [0, ncols) of A: one thread per column
template <class T>
__global__ void lla_xlaswp_kernel(T* __restrict__ A, int64_t lda, int64_t ncols, int64_t k1, int64_t k2,
                                  const int* __restrict__ ipiv, bool forward) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  T* col = A + c * lda;
  for (int64_t q = 0; q < k2 - k1; q++) {
    const int64_t i = forward ? k1 + q : k2 - 1 - q;
    const int64_t p = ipiv[i] - 1;
    if (p != i) { const T t = col[i]; col[i] = col[p]; col[p] = t; }
  }
}
template <class T>
static int xlaswp_dev(int64_t ncols, T* A, int64_t lda, int64_t k1, int64_t k2, const int* ipiv, bool forward, cudaStream_t st) {
  if (ncols <= 0 || k2 <= k1) return 0;
  lla_xlaswp_kernel<T><<<(unsigned)((ncols + 127) / 128), 128, 0, st>>>(A, lda, ncols, k1, k2, ipiv, forward);
  count_launch();
  LLA_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ 32-row substitution leaf
// op(Tkk) X = B for one diagonal block (nb <= 32) and ncols right-hand sides, one thread per right-hand side.  The block
// is loaded as the effective matrix E = op(Tkk) (E lower: forward substitution, E upper: backward).
template <class T>
__global__ void __launch_bounds__(128) lla_xtrsm32_kernel(bool upper, int op, bool unit, int nb, int64_t ncols,
                                                          const T* __restrict__ Tk, int64_t ldt, T* __restrict__ B, int64_t ldb,
                                                          const int* __restrict__ abort_flag) {
  if (abort_flag != nullptr && *abort_flag != 0) return;
  __shared__ T E[32][33];
  for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
    const int i = e % 32, j = e / 32;
    T v = xzero(T());
    if (i < nb && j < nb) {
      // E(i, j) = op(T)(i, j)
      const bool in_tri = (op == 0) ? (upper ? i <= j : i >= j) : (upper ? j <= i : j >= i);
      if (in_tri) v = (op == 0) ? Tk[i + (int64_t)j * ldt] : xop(Tk[j + (int64_t)i * ldt], op);
    }
    E[i][j] = v;
  }
  __syncthreads();
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  const bool eff_lower = (upper == (op != 0));
  T x[32];
  T* b = B + c * ldb;
#pragma unroll
  for (int i = 0; i < 32; i++) x[i] = (i < nb) ? b[i] : xzero(T());
  if (eff_lower) {
#pragma unroll
    for (int i = 0; i < 32; i++) {
      if (i < nb) {
        T s = x[i];
#pragma unroll
        for (int j = 0; j < i; j++) s = xsub(s, xmul(E[i][j], x[j]));
        x[i] = unit ? s : xdiv(s, E[i][i]);
      }
    }
  } else {
#pragma unroll
    for (int i = 31; i >= 0; i--) {
      if (i < nb) {
        T s = x[i];
#pragma unroll
        for (int j = i + 1; j < 32; j++)
          if (j < nb) s = xsub(s, xmul(E[i][j], x[j]));
        x[i] = unit ? s : xdiv(s, E[i][i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 32; i++)
    if (i < nb) b[i] = x[i];
}
template <class T>
static int xtrsm32(bool upper, int op, bool unit, int nb, int64_t ncols, const T* Tk, int64_t ldt, T* B, int64_t ldb,
                   cudaStream_t st, const int* abort_flag) {
  if (ncols <= 0 || nb <= 0) return 0;
  lla_xtrsm32_kernel<T><<<(unsigned)((ncols + 127) / 128), 128, 0, st>>>(upper, op, unit, nb, ncols, Tk, ldt, B, ldb, abort_flag);
  count_launch();
  LLA_CUDA(cudaGetLastError());
  return 0;
}

// op(T) X = B, T n x n triangular, B n x nrhs: 32-row leaves coupled by GEMMs
template <class T>
static int xtrsm_blocked(bool upper, int op, bool unit, int64_t n, int64_t nrhs, const T* Tm, int64_t ldt, T* B, int64_t ldb,
                         cudaStream_t st, const int* abort_flag) {
  if (n <= 0 || nrhs <= 0) return 0;
  const bool eff_lower = (upper == (op != 0));
  const T one = xfrom(T(), (typename Real<T>::type)1), mone = xfrom(T(), (typename Real<T>::type)-1);
  const int64_t nblk = (n + 31) / 32;
  for (int64_t s = 0; s < nblk; s++) {
    const int64_t kb0 = eff_lower ? s : nblk - 1 - s;
    const int64_t k0 = kb0 * 32, k1 = (k0 + 32 < n) ? k0 + 32 : n;
    const int nb = (int)(k1 - k0);
    LLA_TRY(xtrsm32<T>(upper, op, unit, nb, nrhs, Tm + k0 + k0 * ldt, ldt, B + k0, ldb, st, abort_flag));
    // fold x_k into the rows still to be solved: below (forward) or above (backward)
    const int64_t r0 = eff_lower ? k1 : 0, r1 = eff_lower ? n : k0;
    if (r1 > r0) {
      // E(r0:r1, k0:k1) = op(T)(r0:r1, k0:k1): T(r0:r1, k0:k1) as it is, or T(k0:k1, r0:r1) (conjugate-)transposed
      const T* blk = (op == 0) ? Tm + r0 + k0 * ldt : Tm + k0 + r0 * ldt;
      LLA_TRY(xgemm_dev<T>(op, 0, r1 - r0, nrhs, nb, mone, blk, ldt, B + k0, ldb, one, B + r0, ldb, TRI_FULL, st, abort_flag));
    }
  }
  return 0;
}

// ------------------------------------------------------------------ LU driver
template <class T>
static int xgetrf_dev(int64_t m, int64_t n, T* A, int64_t lda, int* ipiv, int* info_dev, cudaStream_t st) {
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  LLA_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), st));
  if (m <= 0 || n <= 0) return 0;
  static DeviceOnce once;
  static int per_sm[MAX_DEV] = {};
  if (once.first()) {
    LLA_CUDA(cudaFuncSetAttribute(lla_xgetf2_panel_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xpanel_smem<T>()));
    LLA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[c->device], lla_xgetf2_panel_kernel<T>, XR, xpanel_smem<T>()));
  }
  const int maxG = (int)((m + XR - 1) / XR);
  if ((int64_t)per_sm[c->device] * c->sm_count < maxG) {
    set_error(__FILE__, __LINE__, "xgetrf: panel too tall for one cooperative launch", (int)m);
    return LLACUDA_ERR_BAD_ARG;
  }
  Arena ar;
  LLA_TRY(ar.reserve(sizeof(XCand) * 2 * maxG + sizeof(T) * 4 * XW + 1024, st));
  XCand* cand = (XCand*)ar.take(sizeof(XCand) * 2 * maxG);
  T* rowbuf = (T*)ar.take(sizeof(T) * 4 * XW);
  const T one = xfrom(T(), (typename Real<T>::type)1), mone = xfrom(T(), (typename Real<T>::type)-1);
  const int64_t mn = m < n ? m : n;
  for (int64_t k = 0; k < mn; k += XW) {
    int w = (int)((mn - k) < XW ? (mn - k) : XW);
    T* P = A + k + k * lda;
    int64_t mk = m - k, lda_ = lda, off = k;
    int G = (int)((mk + XR - 1) / XR);
    void* args[] = {&P, &lda_, &mk, &w, &off, &ipiv, &info_dev, &cand, &rowbuf};
    LLA_CUDA(cudaLaunchCooperativeKernel((void*)lla_xgetf2_panel_kernel<T>, dim3(G), dim3(XR), args, xpanel_smem<T>(), st));
    count_launch();
    // interchanges on the columns left and right of the panel
    LLA_TRY(xlaswp_dev<T>(k, A, lda, k, k + w, ipiv, true, st));
    const int64_t nr = n - k - w;
    if (nr > 0) {
      T* A12 = A + k + (k + w) * lda;
      LLA_TRY(xlaswp_dev<T>(nr, A + (k + w) * lda, lda, k, k + w, ipiv, true, st));
      LLA_TRY(xtrsm32<T>(false, 0, true, w, nr, P, lda, A12, lda, st, nullptr));
      if (mk > w)
        LLA_TRY(xgemm_dev<T>(0, 0, mk - w, nr, w, mone, P + w, lda, A12, lda, one, A12 + w, lda, TRI_FULL, st, nullptr));
    }
  }
  return 0;
}

template <class T>
static int xgetrs_dev(int op, int64_t n, int64_t nrhs, const T* A, int64_t lda, const int* ipiv, T* B, int64_t ldb,
                      cudaStream_t st, const int* abort_flag) {
  if (n <= 0 || nrhs <= 0) return 0;
  if (op == 0) {
    LLA_TRY(xlaswp_dev<T>(nrhs, B, ldb, 0, n, ipiv, true, st));
    LLA_TRY(xtrsm_blocked<T>(false, 0, true, n, nrhs, A, lda, B, ldb, st, abort_flag));
    LLA_TRY(xtrsm_blocked<T>(true, 0, false, n, nrhs, A, lda, B, ldb, st, abort_flag));
  } else {
    LLA_TRY(xtrsm_blocked<T>(true, op, false, n, nrhs, A, lda, B, ldb, st, abort_flag));
    LLA_TRY(xtrsm_blocked<T>(false, op, true, n, nrhs, A, lda, B, ldb, st, abort_flag));
    LLA_TRY(xlaswp_dev<T>(nrhs, B, ldb, 0, n, ipiv, false, st));
  }
  return 0;
}

// ------------------------------------------------------------------ Cholesky
// 32 x 32 diagonal block, A = U^H U, one warp: lane = column.  xPOTF2 arithmetic: the diagonal is taken real, a
// non-positive or NaN pivot stores INFO and stops.
template <class T>
__global__ void __launch_bounds__(32) lla_xpotf2_kernel(int nb, T* __restrict__ A, int64_t lda, int off, int* __restrict__ info) {
  typedef typename Real<T>::type R;
  if (*info != 0) return;
  __shared__ T U[32][33];   // U[k][i] = element (k, i), k <= i
  const int i = threadIdx.x;
  for (int k = 0; k < 32; k++) U[k][i] = (i < nb && k <= i && k < nb) ? A[k + (int64_t)i * lda] : xzero(T());
  __syncwarp();
  for (int j = 0; j < nb; j++) {
    // column i >= j: a_ji - sum_{k<j} conj(u_kj) u_ki
    T s = U[j][i];
    for (int k = 0; k < j; k++) s = xsub(s, xmul(xconj(U[k][j]), U[k][i]));
    const R ajj = __shfl_sync(0xffffffffu, xreal(s), j);
    if (!(ajj > (R)0)) {   // also catches NaN
      if (i == 0) *info = off + j + 1;
      return;
    }
    const R d = sqrt(ajj);
    if (i == j) U[j][i] = xfrom(T(), d);
    else if (i > j && i < nb) U[j][i] = xdiv(s, xfrom(T(), d));
    __syncwarp();
  }
  for (int k = 0; k < nb; k++)
    if (i < nb && k <= i) A[k + (int64_t)i * lda] = U[k][i];
}

// square in-place conjugate transpose (UPLO = 'L' runs the 'U' code on it)
template <class T>
__global__ void lla_xconjtrans_inplace_kernel(int64_t n, T* __restrict__ A, int64_t lda) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= n || i > j) return;
  const T a = A[i + j * lda], b = A[j + i * lda];
  A[i + j * lda] = xconj(b);
  A[j + i * lda] = xconj(a);
}

template <class T>
static int xpotrf_dev(bool upper, int64_t n, T* A, int64_t lda, int* info_dev, cudaStream_t st) {
  if (!ctx()) return LLACUDA_ERR_NO_DEVICE;
  LLA_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), st));
  if (n <= 0) return 0;
  dim3 tg((unsigned)((n + 127) / 128), (unsigned)n);
  if (!upper) { lla_xconjtrans_inplace_kernel<T><<<tg, 128, 0, st>>>(n, A, lda); count_launch(); }
  const T one = xfrom(T(), (typename Real<T>::type)1), mone = xfrom(T(), (typename Real<T>::type)-1);
  for (int64_t k = 0; k < n; k += 32) {
    const int nb = (int)((n - k) < 32 ? (n - k) : 32);
    T* Akk = A + k + k * lda;
    lla_xpotf2_kernel<T><<<1, 32, 0, st>>>(nb, Akk, lda, (int)k, info_dev);
    count_launch();
    const int64_t n2 = n - k - nb;
    if (n2 > 0) {
      T* A12 = Akk + (int64_t)nb * lda;
      LLA_TRY(xtrsm32<T>(true, 2, false, nb, n2, Akk, lda, A12, lda, st, info_dev));       // U12 = U11^-H A12
      LLA_TRY(xgemm_dev<T>(2, 0, n2, n2, nb, mone, A12, lda, A12, lda, one, A12 + nb, lda, TRI_UPPER, st, info_dev));
    }
  }
  if (!upper) { lla_xconjtrans_inplace_kernel<T><<<tg, 128, 0, st>>>(n, A, lda); count_launch(); }
  LLA_CUDA(cudaGetLastError());
  return 0;
}

template <class T>
static int xpotrs_dev(bool upper, int64_t n, int64_t nrhs, const T* A, int64_t lda, T* B, int64_t ldb, cudaStream_t st,
                      const int* abort_flag) {
  if (upper) {   // A = U^H U
    LLA_TRY(xtrsm_blocked<T>(true, 2, false, n, nrhs, A, lda, B, ldb, st, abort_flag));
    LLA_TRY(xtrsm_blocked<T>(true, 0, false, n, nrhs, A, lda, B, ldb, st, abort_flag));
  } else {       // A = L L^H
    LLA_TRY(xtrsm_blocked<T>(false, 0, false, n, nrhs, A, lda, B, ldb, st, abort_flag));
    LLA_TRY(xtrsm_blocked<T>(false, 2, false, n, nrhs, A, lda, B, ldb, st, abort_flag));
  }
  return 0;
}

// the imaginary part of a Hermitian diagonal is exactly zero (xHERK sets it)
template <class T>
__global__ void lla_xreal_diag_kernel(int64_t n, T* __restrict__ C, int64_t ldc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) C[i + i * ldc] = xfrom(T(), xreal(C[i + i * ldc]));
}

// ------------------------------------------------------------------ host-pointer bodies
static inline int64_t max1(int64_t x) { return x > 1 ? x : 1; }
static int xop_code(char c) {
  switch (c) {
    case 'N': case 'n': return 0;
    case 'T': case 't': return 1;
    case 'C': case 'c': return 2;
    default: return -1;
  }
}
struct XIntOut {
  fint* dst; int64_t n; std::vector<int> tmp;
  XIntOut(fint* d, int64_t count) : dst(d), n(count) { if (sizeof(fint) != sizeof(int)) tmp.resize((size_t)(count > 0 ? count : 1)); }
  int* host32() { return sizeof(fint) == sizeof(int) ? (int*)dst : tmp.data(); }
  void finish() { if (sizeof(fint) != sizeof(int)) for (int64_t i = 0; i < n; i++) dst[i] = (fint)tmp[(size_t)i]; }
};

template <class T>
static int xgesv_host(int64_t m, int64_t n, int64_t nrhs, T* a, int64_t lda, fint* ipiv, T* b, int64_t ldb, fint* info) {
  // nrhs < 0: GETRF of an m x n matrix; otherwise GESV (m == n)
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  cudaStream_t st = cx->stream;
  Staged sa, sb;
  PoolBuf piv, inf;
  const int64_t mn = m < n ? m : n;
  LLA_TRY(stage_alloc(sa, m, n, sizeof(T)));
  LLA_TRY(piv.alloc(sizeof(int) * (size_t)max1(mn)));
  LLA_TRY(inf.alloc(sizeof(int)));
  LLA_TRY(stage_upload(sa, a, lda, st));
  if (nrhs > 0) {
    LLA_TRY(stage_alloc(sb, n, nrhs, sizeof(T)));
    LLA_TRY(stage_upload(sb, b, ldb, st));
  }
  LLA_TRY(xgetrf_dev<T>(m, n, sa.ptr<T>(), sa.ld, (int*)piv.p, (int*)inf.p, st));
  XIntOut pivots(ipiv, mn), inf_out(info, 1);
  LLA_CUDA(cudaMemcpyAsync(inf_out.host32(), inf.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (nrhs > 0) {   // GESV solves only when the factorisation found no zero pivot
    LLA_CUDA(cudaStreamSynchronize(st));
    if (*inf_out.host32() == 0) {
      LLA_TRY(xgetrs_dev<T>(0, n, nrhs, sa.ptr<T>(), sa.ld, (const int*)piv.p, sb.ptr<T>(), sb.ld, st, nullptr));
      LLA_TRY(stage_download(sb, b, ldb, st));
    }
  }
  LLA_TRY(stage_download(sa, a, lda, st));
  if (mn > 0) LLA_CUDA(cudaMemcpyAsync(pivots.host32(), piv.p, sizeof(int) * (size_t)mn, cudaMemcpyDeviceToHost, st));
  LLA_TRY(stage_sync(st));
  pivots.finish(); inf_out.finish();
  return 0;
}

template <class T>
static int xgetrs_host(int op, int64_t n, int64_t nrhs, const T* a, int64_t lda, const fint* ipiv, T* b, int64_t ldb) {
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  cudaStream_t st = cx->stream;
  Staged sa, sb;
  PoolBuf piv;
  LLA_TRY(stage_alloc(sa, n, n, sizeof(T)));
  LLA_TRY(stage_alloc(sb, n, nrhs, sizeof(T)));
  LLA_TRY(piv.alloc(sizeof(int) * (size_t)n));
  std::vector<int> p32((size_t)n);
  for (int64_t i = 0; i < n; i++) p32[(size_t)i] = (int)ipiv[i];
  LLA_TRY(stage_upload(sa, a, lda, st));
  LLA_TRY(stage_upload(sb, b, ldb, st));
  LLA_CUDA(cudaMemcpyAsync(piv.p, p32.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
  LLA_TRY(xgetrs_dev<T>(op, n, nrhs, sa.ptr<T>(), sa.ld, (const int*)piv.p, sb.ptr<T>(), sb.ld, st, nullptr));
  LLA_TRY(stage_download(sb, b, ldb, st));
  return stage_sync(st);
}

// POTRF (nrhs < 0), POSV (factor + solve), POTRS (solve_only: `a` holds the factor)
template <class T>
static int xposv_host(bool upper, int64_t n, int64_t nrhs, T* a, int64_t lda, T* b, int64_t ldb, fint* info, bool solve_only) {
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  cudaStream_t st = cx->stream;
  Staged sa, sb;
  PoolBuf inf;
  LLA_TRY(stage_alloc(sa, n, n, sizeof(T)));
  LLA_TRY(inf.alloc(sizeof(int)));
  LLA_CUDA(cudaMemsetAsync(inf.p, 0, sizeof(int), st));
  LLA_TRY(stage_upload_tri(sa, a, lda, upper, st));
  if (nrhs > 0) {
    LLA_TRY(stage_alloc(sb, n, nrhs, sizeof(T)));
    LLA_TRY(stage_upload(sb, b, ldb, st));
  }
  XIntOut inf_out(info, 1);
  *inf_out.host32() = 0;
  if (!solve_only) {
    LLA_TRY(xpotrf_dev<T>(upper, n, sa.ptr<T>(), sa.ld, (int*)inf.p, st));
    LLA_CUDA(cudaMemcpyAsync(inf_out.host32(), inf.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    LLA_CUDA(cudaStreamSynchronize(st));
  }
  if (nrhs > 0 && *inf_out.host32() == 0) {   // POSV solves only after a successful factorisation; B stays as it was otherwise
    LLA_TRY(xpotrs_dev<T>(upper, n, nrhs, sa.ptr<T>(), sa.ld, sb.ptr<T>(), sb.ld, st, nullptr));
    LLA_TRY(stage_download(sb, b, ldb, st));
  }
  if (!solve_only) LLA_TRY(stage_download_tri(sa, a, lda, upper, 0, st));
  LLA_TRY(stage_sync(st));
  inf_out.finish();
  return 0;
}

// C <- alpha A A^H + beta C (trans = 'N') or alpha A^H A + beta C, alpha / beta real, only the uplo triangle of C
template <class T>
static int xherk_host(bool upper, int trans, int64_t n, int64_t k, typename Real<T>::type alpha, const T* a, int64_t lda,
                      typename Real<T>::type beta, T* c, int64_t ldc) {
  Context* cx = ctx();
  if (!cx) return LLACUDA_ERR_NO_DEVICE;
  if (n == 0) return 0;
  cudaStream_t st = cx->stream;
  Staged sa, sc;
  const bool need_a = k > 0 && alpha != 0;
  if (need_a) {
    LLA_TRY(stage_alloc(sa, trans ? k : n, trans ? n : k, sizeof(T)));
    LLA_TRY(stage_upload(sa, a, lda, st));
  }
  LLA_TRY(stage_alloc(sc, n, n, sizeof(T)));
  LLA_TRY(stage_upload_tri(sc, c, ldc, upper, st));
  dim3 tg((unsigned)((n + 127) / 128), (unsigned)n);
  // the GEMM below writes the upper triangle; a lower request runs on the conjugate transpose of C
  if (!upper) { lla_xconjtrans_inplace_kernel<T><<<tg, 128, 0, st>>>(n, sc.ptr<T>(), sc.ld); count_launch(); }
  LLA_TRY(xgemm_dev<T>(trans ? 2 : 0, trans ? 0 : 2, n, n, need_a ? k : 0, xfrom(T(), alpha), sa.ptr<T>(), sa.ld, sa.ptr<T>(), sa.ld,
                       xfrom(T(), beta), sc.ptr<T>(), sc.ld, TRI_UPPER, st, nullptr));
  lla_xreal_diag_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, sc.ptr<T>(), sc.ld);
  count_launch();
  if (!upper) { lla_xconjtrans_inplace_kernel<T><<<tg, 128, 0, st>>>(n, sc.ptr<T>(), sc.ld); count_launch(); }
  LLA_CUDA(cudaGetLastError());
  LLA_TRY(stage_download_tri(sc, c, ldc, upper, 0, st));
  return stage_sync(st);
}

}  // namespace llacuda

using namespace llacuda;

// ------------------------------------------------------------------ Fortran symbols (and their llacuda_ aliases)
#define LLA_FAIL_INFO(rc) do { drain_streams(); stage_sync(nullptr); *info = (fint)(rc); } while (0)

#define LLA_X_LAPACK(L, T)                                                                                                        \
  /* lu: reference src/linear-algebra.lisp:227-234 */                                                                            \
  extern "C" void L##getrf_(const fint* m_, const fint* n_, void* a, const fint* lda_, fint* ipiv, fint* info) {                 \
    ApiLock api_lock;                                                                                                             \
    const int64_t m = *m_, n = *n_, lda = *lda_;                                                                                  \
    *info = 0;                                                                                                                    \
    if (m < 0) { *info = -1; return; }                                                                                            \
    if (n < 0) { *info = -2; return; }                                                                                            \
    if (lda < max1(m)) { *info = -4; return; }                                                                                    \
    if (m == 0 || n == 0) return;                                                                                                 \
    int rc = xgesv_host<T>(m, n, -1, (T*)a, lda, ipiv, nullptr, 0, info);                                                         \
    if (rc != 0) LLA_FAIL_INFO(rc);                                                                                               \
  }                                                                                                                               \
  /* solve (lu): reference src/linear-algebra.lisp:269-276 */                                                                    \
  extern "C" void L##getrs_(const char* trans, const fint* n_, const fint* nrhs_, const void* a, const fint* lda_,               \
                            const fint* ipiv, void* b, const fint* ldb_, fint* info) {                                           \
    ApiLock api_lock;                                                                                                             \
    const int op = xop_code(*trans);                                                                                              \
    const int64_t n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;                                                               \
    *info = 0;                                                                                                                    \
    if (op < 0) { *info = -1; return; }                                                                                           \
    if (n < 0) { *info = -2; return; }                                                                                            \
    if (nrhs < 0) { *info = -3; return; }                                                                                         \
    if (lda < max1(n)) { *info = -5; return; }                                                                                    \
    if (ldb < max1(n)) { *info = -8; return; }                                                                                    \
    if (n == 0 || nrhs == 0) return;                                                                                              \
    int rc = xgetrs_host<T>(op, n, nrhs, (const T*)a, lda, ipiv, (T*)b, ldb);                                                     \
    if (rc != 0) LLA_FAIL_INFO(rc);                                                                                               \
  }                                                                                                                               \
  /* solve (array array): reference src/linear-algebra.lisp:282-289 */                                                           \
  extern "C" void L##gesv_(const fint* n_, const fint* nrhs_, void* a, const fint* lda_, fint* ipiv, void* b, const fint* ldb_,  \
                           fint* info) {                                                                                          \
    ApiLock api_lock;                                                                                                             \
    const int64_t n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;                                                               \
    *info = 0;                                                                                                                    \
    if (n < 0) { *info = -1; return; }                                                                                            \
    if (nrhs < 0) { *info = -2; return; }                                                                                         \
    if (lda < max1(n)) { *info = -4; return; }                                                                                    \
    if (ldb < max1(n)) { *info = -7; return; }                                                                                    \
    if (n == 0) return;                                                                                                           \
    int rc = xgesv_host<T>(n, n, nrhs, (T*)a, lda, ipiv, (T*)b, ldb, info);                                                       \
    if (rc != 0) LLA_FAIL_INFO(rc);                                                                                               \
  }                                                                                                                               \
  /* cholesky: reference src/linear-algebra.lisp:670-674 */                                                                      \
  extern "C" void L##potrf_(const char* uplo, const fint* n_, void* a, const fint* lda_, fint* info) {                           \
    ApiLock api_lock;                                                                                                             \
    const int64_t n = *n_, lda = *lda_;                                                                                           \
    const bool upper = (*uplo == 'U' || *uplo == 'u'), lower = (*uplo == 'L' || *uplo == 'l');                                    \
    *info = 0;                                                                                                                    \
    if (!upper && !lower) { *info = -1; return; }                                                                                 \
    if (n < 0) { *info = -2; return; }                                                                                            \
    if (lda < max1(n)) { *info = -4; return; }                                                                                    \
    if (n == 0) return;                                                                                                           \
    int rc = xposv_host<T>(upper, n, -1, (T*)a, lda, nullptr, 0, info, false);                                                    \
    if (rc != 0) LLA_FAIL_INFO(rc);                                                                                               \
  }                                                                                                                               \
  /* solve (cholesky): reference src/linear-algebra.lisp:296-301 */                                                              \
  extern "C" void L##potrs_(const char* uplo, const fint* n_, const fint* nrhs_, const void* a, const fint* lda_, void* b,       \
                            const fint* ldb_, fint* info) {                                                                       \
    ApiLock api_lock;                                                                                                             \
    const int64_t n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;                                                               \
    const bool upper = (*uplo == 'U' || *uplo == 'u'), lower = (*uplo == 'L' || *uplo == 'l');                                    \
    *info = 0;                                                                                                                    \
    if (!upper && !lower) { *info = -1; return; }                                                                                 \
    if (n < 0) { *info = -2; return; }                                                                                            \
    if (nrhs < 0) { *info = -3; return; }                                                                                         \
    if (lda < max1(n)) { *info = -5; return; }                                                                                    \
    if (ldb < max1(n)) { *info = -7; return; }                                                                                    \
    if (n == 0 || nrhs == 0) return;                                                                                              \
    fint dummy = 0;                                                                                                               \
    int rc = xposv_host<T>(upper, n, nrhs, (T*)const_cast<void*>(a), lda, (T*)b, ldb, &dummy, true);                              \
    if (rc != 0) LLA_FAIL_INFO(rc);                                                                                               \
  }                                                                                                                               \
  /* solve (hermitian-matrix): reference src/linear-algebra.lisp:303-315 */                                                      \
  extern "C" void L##posv_(const char* uplo, const fint* n_, const fint* nrhs_, void* a, const fint* lda_, void* b,              \
                           const fint* ldb_, fint* info) {                                                                        \
    ApiLock api_lock;                                                                                                             \
    const int64_t n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;                                                               \
    const bool upper = (*uplo == 'U' || *uplo == 'u'), lower = (*uplo == 'L' || *uplo == 'l');                                    \
    *info = 0;                                                                                                                    \
    if (!upper && !lower) { *info = -1; return; }                                                                                 \
    if (n < 0) { *info = -2; return; }                                                                                            \
    if (nrhs < 0) { *info = -3; return; }                                                                                         \
    if (lda < max1(n)) { *info = -5; return; }                                                                                    \
    if (ldb < max1(n)) { *info = -7; return; }                                                                                    \
    if (n == 0) return;                                                                                                           \
    int rc = xposv_host<T>(upper, n, nrhs, (T*)a, lda, (T*)b, ldb, info, false);                                                  \
    if (rc != 0) LLA_FAIL_INFO(rc);                                                                                               \
  }                                                                                                                               \
  extern "C" void llacuda_##L##getrf(const fint* m, const fint* n, void* a, const fint* lda, fint* ipiv, fint* info) {           \
    L##getrf_(m, n, a, lda, ipiv, info);                                                                                          \
  }                                                                                                                               \
  extern "C" void llacuda_##L##getrs(const char* t, const fint* n, const fint* nrhs, const void* a, const fint* lda,             \
                                     const fint* ipiv, void* b, const fint* ldb, fint* info) {                                   \
    L##getrs_(t, n, nrhs, a, lda, ipiv, b, ldb, info);                                                                            \
  }                                                                                                                               \
  extern "C" void llacuda_##L##gesv(const fint* n, const fint* nrhs, void* a, const fint* lda, fint* ipiv, void* b,              \
                                    const fint* ldb, fint* info) {                                                                \
    L##gesv_(n, nrhs, a, lda, ipiv, b, ldb, info);                                                                                \
  }                                                                                                                               \
  extern "C" void llacuda_##L##potrf(const char* u, const fint* n, void* a, const fint* lda, fint* info) {                       \
    L##potrf_(u, n, a, lda, info);                                                                                                \
  }                                                                                                                               \
  extern "C" void llacuda_##L##potrs(const char* u, const fint* n, const fint* nrhs, const void* a, const fint* lda, void* b,    \
                                     const fint* ldb, fint* info) {                                                               \
    L##potrs_(u, n, nrhs, a, lda, b, ldb, info);                                                                                  \
  }                                                                                                                               \
  extern "C" void llacuda_##L##posv(const char* u, const fint* n, const fint* nrhs, void* a, const fint* lda, void* b,           \
                                    const fint* ldb, fint* info) {                                                                \
    L##posv_(u, n, nrhs, a, lda, b, ldb, info);                                                                                   \
  }

LLA_X_LAPACK(s, float)
LLA_X_LAPACK(c, C32)
LLA_X_LAPACK(z, C64)

/* mm-hermitian%: reference src/linear-algebra.lisp:62-76 -- ("syrk" "herk"): SYRK for real arrays, HERK for complex ones;
 * LLA hands alpha = 1 and beta = 0 as cells of the matrix type, of which HERK reads the (leading) real part */
#define LLA_X_HERK(NAME, T, R)                                                                                                    \
  extern "C" void NAME##_(const char* uplo, const char* trans, const fint* n_, const fint* k_, const R* alpha, const void* a,     \
                          const fint* lda_, const R* beta, void* c, const fint* ldc_) {                                           \
    ApiLock api_lock;                                                                                                             \
    const bool upper = (*uplo == 'U' || *uplo == 'u'), lower = (*uplo == 'L' || *uplo == 'l');                                    \
    const int op = xop_code(*trans);                                                                                              \
    const int64_t n = *n_, k = *k_, lda = *lda_, ldc = *ldc_;                                                                     \
    const int64_t nrowa = op ? k : n;                                                                                             \
    llacuda_clear_error();                                                                                                        \
    if ((!upper && !lower) || op < 0 || n < 0 || k < 0 || lda < max1(nrowa) || ldc < max1(n)) {                                   \
      set_error(__FILE__, __LINE__, #NAME ": illegal argument", -1);                                                              \
      blas_failure(#NAME "_", LLACUDA_ERR_BAD_ARG, nullptr, 0, 0, 0, sizeof(T));                                                  \
      return;                                                                                                                     \
    }                                                                                                                             \
    int rc = xherk_host<T>(upper, op ? 1 : 0, n, k, *alpha, (const T*)a, lda, *beta, (T*)c, ldc);                                 \
    if (rc != 0) blas_failure(#NAME "_", rc, c, ldc, n, n, sizeof(T));                                                            \
  }                                                                                                                               \
  extern "C" void llacuda_##NAME(const char* uplo, const char* trans, const fint* n, const fint* k, const R* alpha,               \
                                 const void* a, const fint* lda, const R* beta, void* c, const fint* ldc) {                       \
    NAME##_(uplo, trans, n, k, alpha, a, lda, beta, c, ldc);                                                                      \
  }

LLA_X_HERK(ssyrk, float, float)
LLA_X_HERK(cherk, C32, float)
LLA_X_HERK(zherk, C64, double)
