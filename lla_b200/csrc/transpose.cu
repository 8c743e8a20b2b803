// lla_b200/csrc/transpose.cu -- layout kernels: the device-side replacement of LLA's element-wise
// transpose-matrix-to-memory / transpose-matrix-from-memory (reference
// src/foreign-memory.lisp:211-255).  HBM-bound: 2 * rows * cols * 8 bytes per call; 32x32 tiles
// through padded shared memory so that both the read and the write are coalesced.
#include "internal.h"
#include "../../include/llacuda.h"

namespace llacuda {

constexpr int TT = 32;

// out[c + r*ldout] = in[r + c*ldin],  in: rows x cols column-major
__global__ void __launch_bounds__(TT * 8) lla_dtranspose_kernel(int64_t rows, int64_t cols, const double* __restrict__ in,
                                                            int64_t ldin, double* __restrict__ out, int64_t ldout) {
  __shared__ double t[TT][TT + 1];
  const int64_t r0 = (int64_t)blockIdx.x * TT, c0 = (int64_t)blockIdx.y * TT;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int k = ty; k < TT; k += 8) {
    int64_t r = r0 + tx, c = c0 + k;
    if (r < rows && c < cols) t[k][tx] = in[r + c * ldin];
  }
  __syncthreads();
#pragma unroll
  for (int k = ty; k < TT; k += 8) {
    int64_t c = c0 + tx, r = r0 + k;
    if (r < rows && c < cols) out[c + r * ldout] = t[tx][k];
  }
}

// A <- A^T for a square n x n matrix; each CTA swaps one tile pair (diagonal tiles transpose in place)
__global__ void __launch_bounds__(TT * 8) lla_dtranspose_inplace_kernel(int64_t n, double* __restrict__ A, int64_t lda) {
  const int64_t bi = blockIdx.x, bj = blockIdx.y;
  if (bi > bj) return;
  __shared__ double t1[TT][TT + 1];
  __shared__ double t2[TT][TT + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t r0 = bi * TT, c0 = bj * TT;
#pragma unroll
  for (int k = ty; k < TT; k += 8) {
    int64_t r = r0 + tx, c = c0 + k;
    if (r < n && c < n) t1[k][tx] = A[r + c * lda];      // tile (bi, bj)
    int64_t r2 = c0 + tx, c2 = r0 + k;
    if (r2 < n && c2 < n) t2[k][tx] = A[r2 + c2 * lda];  // tile (bj, bi)
  }
  __syncthreads();
#pragma unroll
  for (int k = ty; k < TT; k += 8) {
    int64_t r = c0 + tx, c = r0 + k;                      // write tile (bj, bi) <- t1^T
    if (r < n && c < n) A[r + c * lda] = t1[tx][k];
    int64_t r2 = r0 + tx, c2 = c0 + k;                    // write tile (bi, bj) <- t2^T
    if (bi != bj && r2 < n && c2 < n) A[r2 + c2 * lda] = t2[tx][k];
  }
}

int dtranspose_dev(int64_t rows, int64_t cols, const double* in, int64_t ldin, double* out, int64_t ldout,
                   cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return 0;
  dim3 grid((unsigned)((rows + TT - 1) / TT), (unsigned)((cols + TT - 1) / TT));
  lla_dtranspose_kernel<<<grid, dim3(TT, 8), 0, st>>>(rows, cols, in, ldin, out, ldout);
  count_launch();
  LLA_CUDA(cudaGetLastError());
  return 0;
}

int dtranspose_inplace_dev(int64_t n, double* A, int64_t lda, cudaStream_t st) {
  if (n <= 1) return 0;
  unsigned g = (unsigned)((n + TT - 1) / TT);
  lla_dtranspose_inplace_kernel<<<dim3(g, g), dim3(TT, 8), 0, st>>>(n, A, lda);
  count_launch();
  LLA_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ convert (+ transpose) on the device
// LLA's copy-array-to-memory / transpose-matrix-to-memory (reference src/foreign-memory.lisp:168-255) walk the array
// element by element in Lisp, coercing every element to the common float type of the call (single, double, complex
// single, complex double: src/types.lisp:22-26; integer arrays classify as double, :69-80), optionally transposing.
// Here the raw array is uploaded as it is and one HBM-bound kernel does both: out[c + r ldout] = (Tout) in[r + c ldin]
// (transposing) or out[r + c ldout] = (Tout) in[r + c ldin].  A complex source for a real destination is an error, as
// in the reference (tests/pinned-array.lisp:67).  Algorithmic bytes: rows cols (sizeof(Tin) + sizeof(Tout)).
template <class Tout, class Tin> struct Cvt { static __device__ __forceinline__ Tout go(Tin v) { return (Tout)v; } };
template <class Tin> struct Cvt<cuFloatComplex, Tin> { static __device__ __forceinline__ cuFloatComplex go(Tin v) { return make_cuFloatComplex((float)v, 0.f); } };
template <class Tin> struct Cvt<cuDoubleComplex, Tin> { static __device__ __forceinline__ cuDoubleComplex go(Tin v) { return make_cuDoubleComplex((double)v, 0.0); } };
template <> struct Cvt<cuFloatComplex, cuFloatComplex> { static __device__ __forceinline__ cuFloatComplex go(cuFloatComplex v) { return v; } };
template <> struct Cvt<cuFloatComplex, cuDoubleComplex> { static __device__ __forceinline__ cuFloatComplex go(cuDoubleComplex v) { return make_cuFloatComplex((float)v.x, (float)v.y); } };
template <> struct Cvt<cuDoubleComplex, cuFloatComplex> { static __device__ __forceinline__ cuDoubleComplex go(cuFloatComplex v) { return make_cuDoubleComplex((double)v.x, (double)v.y); } };
template <> struct Cvt<cuDoubleComplex, cuDoubleComplex> { static __device__ __forceinline__ cuDoubleComplex go(cuDoubleComplex v) { return v; } };

template <class Tout, class Tin, bool TR>
__global__ void __launch_bounds__(TT * 8) lla_convert_kernel(int64_t rows, int64_t cols, const Tin* __restrict__ in, int64_t ldin,
                                                             Tout* __restrict__ out, int64_t ldout) {
  const int64_t r0 = (int64_t)blockIdx.x * TT, c0 = (int64_t)blockIdx.y * TT;
  const int tx = threadIdx.x, ty = threadIdx.y;
  if constexpr (!TR) {
#pragma unroll
    for (int k = ty; k < TT; k += 8) {
      const int64_t r = r0 + tx, c = c0 + k;
      if (r < rows && c < cols) out[r + c * ldout] = Cvt<Tout, Tin>::go(in[r + c * ldin]);
    }
  } else {
    __shared__ Tout t[TT][TT + 1];
#pragma unroll
    for (int k = ty; k < TT; k += 8) {
      const int64_t r = r0 + tx, c = c0 + k;
      if (r < rows && c < cols) t[k][tx] = Cvt<Tout, Tin>::go(in[r + c * ldin]);
    }
    __syncthreads();
#pragma unroll
    for (int k = ty; k < TT; k += 8) {
      const int64_t c = c0 + tx, r = r0 + k;
      if (r < rows && c < cols) out[c + r * ldout] = t[tx][k];
    }
  }
}

template <class Tout, class Tin>
static int convert_launch(bool tr, int64_t rows, int64_t cols, const void* in, int64_t ldin, void* out, int64_t ldout, cudaStream_t st) {
  dim3 grid((unsigned)((rows + TT - 1) / TT), (unsigned)((cols + TT - 1) / TT));
  if (tr) lla_convert_kernel<Tout, Tin, true><<<grid, dim3(TT, 8), 0, st>>>(rows, cols, (const Tin*)in, ldin, (Tout*)out, ldout);
  else lla_convert_kernel<Tout, Tin, false><<<grid, dim3(TT, 8), 0, st>>>(rows, cols, (const Tin*)in, ldin, (Tout*)out, ldout);
  count_launch();
  LLA_CUDA(cudaGetLastError());
  return 0;
}

template <class Tout>
static int convert_from(int in_type, bool tr, int64_t rows, int64_t cols, const void* in, int64_t ldin, void* out, int64_t ldout,
                        cudaStream_t st) {
  switch (in_type) {
    case 0: return convert_launch<Tout, float>(tr, rows, cols, in, ldin, out, ldout, st);
    case 1: return convert_launch<Tout, double>(tr, rows, cols, in, ldin, out, ldout, st);
    case 4: return convert_launch<Tout, int>(tr, rows, cols, in, ldin, out, ldout, st);
    case 5: return convert_launch<Tout, long long>(tr, rows, cols, in, ldin, out, ldout, st);
    default: return LLACUDA_ERR_BAD_ARG;
  }
}

// type codes: LLA's internal ones (0 single, 1 double, 2 complex single, 3 complex double) plus 4 = int32, 5 = int64 sources
int convert_dev(int in_type, int out_type, bool transpose, int64_t rows, int64_t cols, const void* in, int64_t ldin, void* out,
                int64_t ldout, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return 0;
  const bool in_cplx = in_type == 2 || in_type == 3, out_cplx = out_type == 2 || out_type == 3;
  if (in_cplx && !out_cplx) {
    set_error(__FILE__, __LINE__, "convert: a complex array cannot be coerced to a real float type", in_type);
    return LLACUDA_ERR_BAD_ARG;
  }
  if (in_cplx) {
    if (in_type == 2 && out_type == 2) return convert_launch<cuFloatComplex, cuFloatComplex>(transpose, rows, cols, in, ldin, out, ldout, st);
    if (in_type == 2 && out_type == 3) return convert_launch<cuDoubleComplex, cuFloatComplex>(transpose, rows, cols, in, ldin, out, ldout, st);
    if (in_type == 3 && out_type == 2) return convert_launch<cuFloatComplex, cuDoubleComplex>(transpose, rows, cols, in, ldin, out, ldout, st);
    return convert_launch<cuDoubleComplex, cuDoubleComplex>(transpose, rows, cols, in, ldin, out, ldout, st);
  }
  switch (out_type) {
    case 0: return convert_from<float>(in_type, transpose, rows, cols, in, ldin, out, ldout, st);
    case 1: return convert_from<double>(in_type, transpose, rows, cols, in, ldin, out, ldout, st);
    case 2: return convert_from<cuFloatComplex>(in_type, transpose, rows, cols, in, ldin, out, ldout, st);
    case 3: return convert_from<cuDoubleComplex>(in_type, transpose, rows, cols, in, ldin, out, ldout, st);
    default: return LLACUDA_ERR_BAD_ARG;
  }
}

}  // namespace llacuda

extern "C" int llacuda_convert_dev(int in_type, int out_type, int transpose, int64_t rows, int64_t cols, const void* in,
                                   int64_t ldin, void* out, int64_t ldout) {
  llacuda::ApiLock api_lock;
  llacuda::Context* c = llacuda::ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  if (rows < 0 || cols < 0 || ldin < (rows > 1 ? rows : 1) || ldout < ((transpose ? cols : rows) > 1 ? (transpose ? cols : rows) : 1))
    return LLACUDA_ERR_BAD_ARG;
  return llacuda::convert_dev(in_type, out_type, transpose != 0, rows, cols, in, ldin, out, ldout, c->stream);
}
