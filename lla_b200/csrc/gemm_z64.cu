// lla_b200/csrc/gemm_z64.cu -- ZGEMM (complex double, interleaved re/im) on the FP64 tensor pipe.
//
// Replaces the `zgemm_` call of LLA's mm / gemm! for complex-double arrays (reference
// src/linear-algebra.lisp:50-54, src/blas.lisp:56-63; LLA sends #\C for a transposed operand,
// i.e. the CONJUGATE transpose, src/blas.lisp:57-58).  A complex product is four real DMMA
// m8n8k4 products on the same fragments:
//     Cr += Ar Br - Ai Bi ,   Ci += Ar Bi + Ai Br
// so the tensor pipe sees 8 real flop per complex multiply-add, the count SURVEY.md 8(d) uses.
// Tiles are staged as interleaved complex (16-byte elements, one cp.async each); the shared
// pitches (== 2 mod 8 elements for MN-contiguous tiles, BK+4 for K-contiguous ones) make every
// quarter-warp of a 128-bit fragment load hit 32 distinct banks.
#include "internal.h"
#include "../../include/llacuda.h"

namespace llacuda {

__device__ __forceinline__ void zcp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void zcp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void zcp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void zdmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int MN, int BK, bool CONTIG_MN>
struct ZTile {
  static constexpr int PITCH = CONTIG_MN ? (MN + 2) : (BK + 4);   // in complex elements
  static constexpr int ELEMS = CONTIG_MN ? BK * PITCH : MN * PITCH;
};

template <int MN, int BK, bool CONTIG_MN, int NT>
__device__ __forceinline__ void zload_tile(double2* __restrict__ sm, const double2* __restrict__ X, int64_t ld,
                                           int64_t mn0, int64_t k0, int64_t MNtot, int64_t Ktot, int tid) {
  using TS = ZTile<MN, BK, CONTIG_MN>;
  constexpr int TOTAL = MN * BK;
  static_assert(TOTAL % NT == 0, "tile must divide evenly over the CTA");
#pragma unroll
  for (int it = 0; it < TOTAL / NT; it++) {
    const int c = tid + it * NT;
    int mm, kk;
    if (CONTIG_MN) { kk = c / MN; mm = c % MN; }
    else { mm = c / BK; kk = c % BK; }
    const int64_t gm = mn0 + mm, gk = k0 + kk;
    const bool in = gm < MNtot && gk < Ktot;
    const int64_t gms = in ? gm : 0, gks = in ? gk : 0;
    const double2* src = CONTIG_MN ? (X + gms + gks * ld) : (X + gks + gms * ld);
    double2* dst = CONTIG_MN ? (sm + kk * TS::PITCH + mm) : (sm + mm * TS::PITCH + kk);
    zcp_async16(dst, src, in ? 16 : 0);
  }
}

struct ZgemmParams {
  int64_t M, N, K;
  const double2* A; int64_t lda;
  const double2* B; int64_t ldb;
  double2* C; int64_t ldc;
  double2 alpha, beta;
  double sa, sb;   // -1 when the operand is conjugated, else +1
};

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool TA, bool TB>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, 2) lla_zgemm_kernel(ZgemmParams p) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr bool A_MN = !TA, B_MN = TB;
  using TSA = ZTile<BM, BK, A_MN>;
  using TSB = ZTile<BN, BK, B_MN>;
  constexpr int STAGE_ELEMS = TSA::ELEMS + TSB::ELEMS;
  constexpr int MI = WM / 8, NI = WN / 8;
  extern __shared__ __align__(16) double2 zsmem[];

  constexpr int GROUP_M = 16;
  const int tiles_m = (int)((p.M + BM - 1) / BM), tiles_n = (int)((p.N + BN - 1) / BN);
  const int pid = blockIdx.x;
  const int group = pid / (GROUP_M * tiles_n);
  const int first_m = group * GROUP_M;
  const int gsz = (tiles_m - first_m) < GROUP_M ? (tiles_m - first_m) : GROUP_M;
  const int64_t m0 = (int64_t)(first_m + (pid % (GROUP_M * tiles_n)) % gsz) * BM;
  const int64_t n0 = (int64_t)((pid % (GROUP_M * tiles_n)) / gsz) * BN;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm0 = (warp % (BM / WM)) * WM, wn0 = (warp / (BM / WM)) * WN;
  const int lr = lane >> 2, lc = lane & 3;

  double accr[MI][NI][2], acci[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NI; j++) accr[i][j][0] = accr[i][j][1] = acci[i][j][0] = acci[i][j][1] = 0.0;

  const int KT = (int)((p.K + BK - 1) / BK);
  auto issue = [&](int kt) {
    if (kt < KT) {
      double2* sa = zsmem + (kt % STAGES) * STAGE_ELEMS;
      double2* sb = sa + TSA::ELEMS;
      zload_tile<BM, BK, A_MN, NT>(sa, p.A, p.lda, m0, (int64_t)kt * BK, p.M, p.K, tid);
      zload_tile<BN, BK, B_MN, NT>(sb, p.B, p.ldb, n0, (int64_t)kt * BK, p.N, p.K, tid);
    }
    zcp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) issue(s);

  const double sa_ = p.sa, sb_ = p.sb;
  for (int kt = 0; kt < KT; kt++) {
    zcp_async_wait<STAGES - 2>();
    __syncthreads();
    issue(kt + STAGES - 1);
    const double2* sa = zsmem + (kt % STAGES) * STAGE_ELEMS;
    const double2* sb = sa + TSA::ELEMS;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double ar[MI], ai[MI], nai[MI], br[NI], bi[NI];
#pragma unroll
      for (int i = 0; i < MI; i++) {
        const int m = wm0 + i * 8 + lr, k = kk + lc;
        const double2 v = A_MN ? sa[k * TSA::PITCH + m] : sa[m * TSA::PITCH + k];
        ar[i] = v.x; ai[i] = v.y * sa_; nai[i] = -ai[i];
      }
#pragma unroll
      for (int j = 0; j < NI; j++) {
        const int n = wn0 + j * 8 + lr, k = kk + lc;
        const double2 v = B_MN ? sb[k * TSB::PITCH + n] : sb[n * TSB::PITCH + k];
        br[j] = v.x; bi[j] = v.y * sb_;
      }
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) {
          zdmma884(accr[i][j][0], accr[i][j][1], ar[i], br[j]);
          zdmma884(accr[i][j][0], accr[i][j][1], nai[i], bi[j]);
          zdmma884(acci[i][j][0], acci[i][j][1], ar[i], bi[j]);
          zdmma884(acci[i][j][0], acci[i][j][1], ai[i], br[j]);
        }
    }
  }
  zcp_async_wait<0>();

  const double2 al = p.alpha, be = p.beta;
  const bool use_c = (be.x != 0.0 || be.y != 0.0);
#pragma unroll
  for (int j = 0; j < NI; j++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int64_t gn = n0 + wn0 + j * 8 + 2 * lc + e;
      if (gn >= p.N) continue;
      double2* ccol = p.C + gn * p.ldc;
#pragma unroll
      for (int i = 0; i < MI; i++) {
        const int64_t gm = m0 + wm0 + i * 8 + lr;
        if (gm >= p.M) continue;
        const double xr = accr[i][j][e], xi = acci[i][j][e];
        double2 v;
        v.x = al.x * xr - al.y * xi;
        v.y = al.x * xi + al.y * xr;
        if (use_c) {
          const double2 c = ccol[gm];
          v.x += be.x * c.x - be.y * c.y;
          v.y += be.x * c.y + be.y * c.x;
        }
        ccol[gm] = v;
      }
    }
}

__global__ void lla_zscale_c_kernel(int64_t M, int64_t N, double2* C, int64_t ldc, double2 beta) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t j = blockIdx.y;
  if (i >= M) return;
  double2* c = C + i + j * ldc;
  double2 v = *c, r;
  if (beta.x == 0.0 && beta.y == 0.0) { r.x = 0.0; r.y = 0.0; }
  else { r.x = beta.x * v.x - beta.y * v.y; r.y = beta.x * v.y + beta.y * v.x; }
  *c = r;
}

template <bool TA, bool TB>
static int zlaunch(const ZgemmParams& p, cudaStream_t st) {
  constexpr int BM = 64, BN = 64, BK = 8, WM = 32, WN = 32, STAGES = 4;
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  using TSA = ZTile<BM, BK, !TA>;
  using TSB = ZTile<BN, BK, TB>;
  constexpr size_t SMEM = size_t(STAGES) * (TSA::ELEMS + TSB::ELEMS) * sizeof(double2);
  auto kern = lla_zgemm_kernel<BM, BN, BK, WM, WN, STAGES, TA, TB>;
  static DeviceOnce once;
  if (once.first()) {
    LLA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  }
  int64_t tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  kern<<<(unsigned)tiles, NT, SMEM, st>>>(p);
  count_launch();
  LLA_CUDA(cudaGetLastError());
  return 0;
}

int zgemm_dev(int opa, int opb, int64_t m, int64_t n, int64_t k, cuDoubleComplex alpha, const cuDoubleComplex* A,
              int64_t lda, const cuDoubleComplex* B, int64_t ldb, cuDoubleComplex beta, cuDoubleComplex* C,
              int64_t ldc, int tri, cudaStream_t st) {
  (void)tri;
  if (m <= 0 || n <= 0) return 0;
  if (!ctx()) return LLACUDA_ERR_NO_DEVICE;
  const bool alpha0 = (alpha.x == 0.0 && alpha.y == 0.0);
  if (k <= 0 || alpha0) {
    if (beta.x == 1.0 && beta.y == 0.0) return 0;
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)n);
    lla_zscale_c_kernel<<<grid, 256, 0, st>>>(m, n, (double2*)C, ldc, make_double2(beta.x, beta.y));
    count_launch();
    LLA_CUDA(cudaGetLastError());
    return 0;
  }
  ZgemmParams p;
  p.M = m; p.N = n; p.K = k;
  p.A = (const double2*)A; p.lda = lda; p.B = (const double2*)B; p.ldb = ldb; p.C = (double2*)C; p.ldc = ldc;
  p.alpha = make_double2(alpha.x, alpha.y); p.beta = make_double2(beta.x, beta.y);
  p.sa = (opa == 2) ? -1.0 : 1.0;
  p.sb = (opb == 2) ? -1.0 : 1.0;
  const bool ta = opa != 0, tb = opb != 0;
  if (!ta && !tb) return zlaunch<false, false>(p, st);
  if (ta && !tb) return zlaunch<true, false>(p, st);
  if (!ta && tb) return zlaunch<false, true>(p, st);
  return zlaunch<true, true>(p, st);
}

}  // namespace llacuda
