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

}  // namespace llacuda
