// lla_b200/csrc/gemm_f64.cu -- DGEMM / ZGEMM on the FP64 tensor pipe (DMMA) for sm_100a.
//
// Replaces the `dgemm_` / `zgemm_` calls of LLA's mm (reference src/linear-algebra.lisp:50-54)
// and gemm! (reference src/blas.lisp:56-63), and is the trailing-matrix update of the blocked
// LU / Cholesky factorisations (getrf.cu, potrf.cu).
//
// Design notes
//  * tcgen05 has no f64 kind; the FP64 tensor path on sm_100a is the warp-level
//    mma.sync.m8n8k4.f64 (every larger f64 shape is lowered by ptxas to DMMA.8x8x4 as well).
//  * column-major operands, 64-bit indexing (lda*n exceeds 2^31 at n=32768).
//  * global -> shared through a STAGES-deep cp.async ring (16-byte copies, zero-fill at the
//    edges); shared tiles are padded so that each half-warp of a 64-bit fragment load touches 32
//    distinct banks: MN-contiguous tiles use a row pitch == 4 (mod 16) doubles, K-contiguous
//    tiles a pitch of BK+4.
//  * C = alpha*op(A)*op(B) + beta*C; beta == 0 never reads C (reference BLAS semantics LLA
//    relies on for freshly made result arrays).
//  * `tri` restricts the update to the upper / lower triangle of C (SYRK-shaped updates of
//    POTRF): tiles strictly on the wrong side exit at once, diagonal tiles mask their stores.
#include "internal.h"
#include "../../include/llacuda.h"

#include <cstdlib>

namespace llacuda {

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---------------------------------------------------------------- tile loader
// A "tile" is MN x BK elements of op(X), X column-major with leading dimension ld.
//   CONTIG_MN = true : op(X)[mn][k] = X[mn + k*ld]   -> shared [k][mn], pitch MN+4
//   CONTIG_MN = false: op(X)[mn][k] = X[k + mn*ld]   -> shared [mn][k], pitch BK+4
template <int MN, int BK, bool CONTIG_MN>
struct TileShape {
  static constexpr int PITCH = CONTIG_MN ? (MN + 4) : (BK + 4);
  static constexpr int ELEMS = CONTIG_MN ? BK * PITCH : MN * PITCH;
};

template <int MN, int BK, bool CONTIG_MN, int NT, bool AL16>
__device__ __forceinline__ void load_tile(double* __restrict__ sm, const double* __restrict__ X, int64_t ld,
                                          int64_t mn0, int64_t k0, int64_t MNtot, int64_t Ktot, int tid) {
  using TS = TileShape<MN, BK, CONTIG_MN>;
  constexpr int CH = 2;  // doubles per 16-byte chunk
  if (CONTIG_MN) {
    constexpr int CPR = MN / CH;  // chunks per k-row
    constexpr int TOTAL = CPR * BK;
    static_assert(TOTAL % NT == 0, "tile chunks must divide evenly over the CTA");
#pragma unroll
    for (int it = 0; it < TOTAL / NT; it++) {
      const int c = tid + it * NT;
      int kk = c / CPR, mc = (c % CPR) * CH;
      int64_t gm = mn0 + mc, gk = k0 + kk;
      double* dst = sm + kk * TS::PITCH + mc;
      int64_t rem = MNtot - gm;
      bool kin = gk < Ktot;
      int64_t gks = kin ? gk : 0;
      int64_t gms = rem > 0 ? gm : 0;
      const double* src = X + gms + gks * ld;
      if (AL16) {
        int bytes = !kin ? 0 : (rem >= 2 ? 16 : (rem == 1 ? 8 : 0));
        cp_async16(dst, src, bytes);
      } else {
        cp_async8(dst, src, (kin && rem >= 1) ? 8 : 0);
        cp_async8(dst + 1, (kin && rem >= 2) ? src + 1 : src, (kin && rem >= 2) ? 8 : 0);
      }
    }
  } else {
    constexpr int CPR = BK / CH;  // chunks per mn-row
    constexpr int TOTAL = CPR * MN;
    static_assert(TOTAL % NT == 0, "tile chunks must divide evenly over the CTA");
#pragma unroll
    for (int it = 0; it < TOTAL / NT; it++) {
      const int c = tid + it * NT;
      int mm = c / CPR, kc = (c % CPR) * CH;
      int64_t gm = mn0 + mm, gk = k0 + kc;
      double* dst = sm + mm * TS::PITCH + kc;
      int64_t rem = Ktot - gk;
      bool min_ = gm < MNtot;
      int64_t gms = min_ ? gm : 0;
      int64_t gks = rem > 0 ? gk : 0;
      const double* src = X + gks + gms * ld;
      if (AL16) {
        int bytes = !min_ ? 0 : (rem >= 2 ? 16 : (rem == 1 ? 8 : 0));
        cp_async16(dst, src, bytes);
      } else {
        cp_async8(dst, src, (min_ && rem >= 1) ? 8 : 0);
        cp_async8(dst + 1, (min_ && rem >= 2) ? src + 1 : src, (min_ && rem >= 2) ? 8 : 0);
      }
    }
  }
}

struct DgemmParams {
  int64_t M, N, K;
  const double* A; int64_t lda;
  const double* B; int64_t ldb;
  double* C; int64_t ldc;
  double alpha, beta;
  int tri;
  const int* abort_flag;  // optional device flag (LAPACK INFO): non-zero => leave C untouched
};

// ---------------------------------------------------------------- the kernel
// CTA tile BM x BN, warp tile WM x WN built from m8n8k4 DMMA atoms.
// Fragment ownership (PTX ISA, mma.m8n8k4.f64): lane l holds A[l/4][l%4], B[l%4][l/4],
// C[l/4][2*(l%4) + {0,1}].
template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool TA, bool TB, bool AL16, int MINB>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, MINB) lla_dgemm_kernel(DgemmParams p) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr bool A_MN = !TA;  // op(A)[m][k] = A[m + k*lda] when not transposed
  constexpr bool B_MN = TB;   // op(B)[k][n] = B[n + k*ldb] when transposed
  using TSA = TileShape<BM, BK, A_MN>;
  using TSB = TileShape<BN, BK, B_MN>;
  constexpr int STAGE_ELEMS = TSA::ELEMS + TSB::ELEMS;
  constexpr int MI = WM / 8, NI = WN / 8;

  extern __shared__ __align__(16) double smem[];

  // grouped rasterisation: consecutive CTAs walk GROUP_M row-tiles of one column-tile strip, so the
  // CTAs resident at any time cover a compact block of C and share their A / B panels through L2
  constexpr int GROUP_M = 16;
  const int tiles_m = (int)((p.M + BM - 1) / BM), tiles_n = (int)((p.N + BN - 1) / BN);
  if (p.abort_flag != nullptr && *p.abort_flag != 0) return;
  // persistent form: with gridDim.x == number of tiles this loop runs once; a smaller grid (SM
  // reservation for a concurrent panel stream) makes every CTA walk several tiles
  for (int pid = blockIdx.x; pid < tiles_m * tiles_n; pid += gridDim.x) {
  const int group = pid / (GROUP_M * tiles_n);
  const int first_m = group * GROUP_M;
  const int gsz = (tiles_m - first_m) < GROUP_M ? (tiles_m - first_m) : GROUP_M;
  const int tm = first_m + (pid % (GROUP_M * tiles_n)) % gsz;
  const int tn = (pid % (GROUP_M * tiles_n)) / gsz;
  const int64_t m0 = (int64_t)tm * BM;
  const int64_t n0 = (int64_t)tn * BN;
  if (p.tri == TRI_UPPER && m0 > n0 + BN - 1) continue;  // tile entirely below the diagonal
  if (p.tri == TRI_LOWER && n0 > m0 + BM - 1) continue;  // tile entirely above the diagonal

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm0 = (warp % (BM / WM)) * WM;
  const int wn0 = (warp / (BM / WM)) * WN;
  const int lr = lane >> 2, lc = lane & 3;

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int KT = (int)((p.K + BK - 1) / BK);

  auto issue = [&](int kt) {
    if (kt < KT) {
      double* sa = smem + (kt % STAGES) * STAGE_ELEMS;
      double* sb = sa + TSA::ELEMS;
      load_tile<BM, BK, A_MN, NT, AL16>(sa, p.A, p.lda, m0, (int64_t)kt * BK, p.M, p.K, tid);
      load_tile<BN, BK, B_MN, NT, AL16>(sb, p.B, p.ldb, n0, (int64_t)kt * BK, p.N, p.K, tid);
    }
    cp_async_commit();
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) issue(s);

  for (int kt = 0; kt < KT; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    issue(kt + STAGES - 1);
    const double* sa = smem + (kt % STAGES) * STAGE_ELEMS;
    const double* sb = sa + TSA::ELEMS;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double fa[MI], fb[NI];
#pragma unroll
      for (int i = 0; i < MI; i++) {
        int m = wm0 + i * 8 + lr, k = kk + lc;
        fa[i] = A_MN ? sa[k * TSA::PITCH + m] : sa[m * TSA::PITCH + k];
      }
#pragma unroll
      for (int j = 0; j < NI; j++) {
        int n = wn0 + j * 8 + lr, k = kk + lc;
        fb[j] = B_MN ? sb[k * TSB::PITCH + n] : sb[n * TSB::PITCH + k];
      }
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue: C = alpha*acc + beta*C straight from the fragment registers.
  const double alpha = p.alpha, beta = p.beta;
  // only tiles that straddle the diagonal pay for the per-element triangle test (tile-uniform branch): behind
  // it the loads of C cannot be batched, which cost the SYRK-shaped updates 12 % on B200 when every tile took it
  const bool diag_tile = (p.tri == TRI_UPPER && m0 + BM - 1 > n0) || (p.tri == TRI_LOWER && n0 + BN - 1 > m0);
  const bool interior = m0 + BM <= p.M && n0 + BN <= p.N;
  if (!diag_tile && interior) {
#pragma unroll
    for (int j = 0; j < NI; j++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        double* ccol = p.C + (n0 + wn0 + j * 8 + 2 * lc + e) * p.ldc + m0 + wm0 + lr;
        if (beta != 0.0) {
          double cv[MI];
#pragma unroll
          for (int i = 0; i < MI; i++) cv[i] = ccol[i * 8];
#pragma unroll
          for (int i = 0; i < MI; i++) {
            double v = alpha * acc[i][j][e];
            v += beta * cv[i];  // same expression as the edge path (identical contraction / rounding)
            ccol[i * 8] = v;
          }
        } else {
#pragma unroll
          for (int i = 0; i < MI; i++) ccol[i * 8] = alpha * acc[i][j][e];
        }
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < NI; j++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        int64_t gn = n0 + wn0 + j * 8 + 2 * lc + e;
        if (gn >= p.N) continue;
        double* ccol = p.C + gn * p.ldc;
#pragma unroll
        for (int i = 0; i < MI; i++) {
          int64_t gm = m0 + wm0 + i * 8 + lr;
          if (gm >= p.M) continue;
          if (p.tri == TRI_UPPER && gm > gn) continue;
          if (p.tri == TRI_LOWER && gm < gn) continue;
          double v = alpha * acc[i][j][e];
          if (beta != 0.0) v += beta * ccol[gm];
          ccol[gm] = v;
        }
      }
    }
  }
  __syncthreads();  // the next tile's prologue overwrites the shared-memory ring
  }  // persistent tile loop
}

// C = beta*C when k == 0 or alpha == 0 (netlib quick return path)
__global__ void lla_dscale_c_kernel(int64_t M, int64_t N, double* C, int64_t ldc, double beta, int tri,
                                const int* abort_flag) {
  if (abort_flag != nullptr && *abort_flag != 0) return;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t j = blockIdx.y;
  if (i >= M) return;
  if (tri == TRI_UPPER && i > j) return;
  if (tri == TRI_LOWER && i < j) return;
  double* c = C + i + j * ldc;
  *c = (beta == 0.0) ? 0.0 : beta * *c;
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool TA, bool TB, bool AL16, int MINB>
static int launch_cfg(const DgemmParams& p, cudaStream_t st, int64_t max_ctas = 0) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  using TSA = TileShape<BM, BK, !TA>;
  using TSB = TileShape<BN, BK, TB>;
  constexpr size_t SMEM = size_t(STAGES) * (TSA::ELEMS + TSB::ELEMS) * sizeof(double);
  auto kern = lla_dgemm_kernel<BM, BN, BK, WM, WN, STAGES, TA, TB, AL16, MINB>;
  static DeviceOnce once;
  if (once.first()) {
    LLA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  }
  int64_t tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  if (max_ctas > 0 && tiles > max_ctas) tiles = max_ctas;
  dim3 grid((unsigned)tiles);
  kern<<<grid, NT, SMEM, st>>>(p);
  count_launch();
  LLA_CUDA(cudaGetLastError());
  return 0;
}

template <bool TA, bool TB, bool AL16>
static int launch_shape(const DgemmParams& p, cudaStream_t st, int sm_count, int hint) {
  if (hint == GEMM_INPLACE_B) return launch_cfg<128, 64, 16, 64, 32, 4, TA, TB, AL16, 2>(p, st);
  if (hint >= GEMM_RESERVE_SMS)  // one CTA per SM on all but (hint - GEMM_RESERVE_SMS + 8) SMs, which stay free for the panel stream
    return launch_cfg<128, 128, 32, 64, 32, 3, TA, TB, AL16, 1>(p, st, sm_count - (8 + hint - GEMM_RESERVE_SMS));
  int64_t big_tiles = ((p.M + 127) / 128) * ((p.N + 127) / 128);
  if (p.tri != TRI_FULL) big_tiles = big_tiles / 2 + 1;
  if (big_tiles >= 2 * (int64_t)sm_count) {
    // measured on B200 (profiles/r01_gemm_config_sweep.txt): two 128x64 CTAs per SM hide each
    // other's barrier bubbles (33.6 vs 31.5 TFLOP/s for one 128x128 CTA).
    static int cfg = -1;
    if (cfg < 0) { const char* e = getenv("LLACUDA_GEMM_CFG"); cfg = e ? atoi(e) : 9; }
    if (cfg == 0) return launch_cfg<128, 128, 16, 64, 32, 4, TA, TB, AL16, 1>(p, st);
    if (cfg == 1) return launch_cfg<128, 64, 16, 64, 32, 4, TA, TB, AL16, 2>(p, st);
    if (cfg == 2) return launch_cfg<128, 128, 32, 64, 32, 3, TA, TB, AL16, 1>(p, st);
    if (cfg == 3) return launch_cfg<128, 64, 16, 64, 32, 3, TA, TB, AL16, 2>(p, st);   // 3 stages: two CTAs/SM also when both tiles are K-contiguous
    // op(A) K-contiguous: the 4-stage ring of the 128x64 tile is 123 KB, one CTA per SM only (25.9 TFLOP/s); with 3
    // stages (92 KB) two CTAs fit again: 31.6 TFLOP/s on the SYRK-shaped k = 1024 update against 30.1 for the
    // 128x128x32 single-CTA tile that used to be selected here
    if (TA) return launch_cfg<128, 64, 16, 64, 32, 3, TA, TB, AL16, 2>(p, st);
    return launch_cfg<128, 64, 16, 64, 32, 4, TA, TB, AL16, 2>(p, st);
  }
  if (p.N <= 32 || p.M <= 32) {
    if (p.N <= 32) return launch_cfg<128, 32, 16, 32, 32, 4, TA, TB, AL16, 2>(p, st);
    return launch_cfg<32, 128, 16, 32, 32, 4, TA, TB, AL16, 2>(p, st);
  }
  return launch_cfg<64, 64, 16, 32, 32, 4, TA, TB, AL16, 2>(p, st);
}

static inline bool aligned16(const void* p, int64_t ld) {
  return ((uintptr_t)p & 15) == 0 && (ld & 1) == 0;
}

int dgemm_dev(int opa, int opb, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
              const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int tri, cudaStream_t st,
              const int* abort_flag, int hint) {
  if (m <= 0 || n <= 0) return 0;
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  if (k <= 0 || alpha == 0.0) {
    if (beta == 1.0) return 0;
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)n);
    lla_dscale_c_kernel<<<grid, 256, 0, st>>>(m, n, C, ldc, beta, tri, abort_flag);
    count_launch();
    LLA_CUDA(cudaGetLastError());
    return 0;
  }
  DgemmParams p{m, n, k, A, lda, B, ldb, C, ldc, alpha, beta, tri, abort_flag};
  bool al = aligned16(A, lda) && aligned16(B, ldb);
  bool ta = opa != 0, tb = opb != 0;
  // gridDim.y limit (65535) is far above any n/32 this library sees (n <= 2^21).
#define LLA_DISPATCH(TA_, TB_)                                             \
  (al ? launch_shape<TA_, TB_, true>(p, st, c->sm_count, hint) : launch_shape<TA_, TB_, false>(p, st, c->sm_count, hint))
  if (!ta && !tb) return LLA_DISPATCH(false, false);
  if (ta && !tb) return LLA_DISPATCH(true, false);
  if (!ta && tb) return LLA_DISPATCH(false, true);
  return LLA_DISPATCH(true, true);
#undef LLA_DISPATCH
}

}  // namespace llacuda

using namespace llacuda;

static int opcode(char c) {
  switch (c) {
    case 'N': case 'n': return 0;
    case 'T': case 't': return 1;
    case 'C': case 'c': return 2;
    default: return -1;
  }
}

extern "C" int llacuda_dsyrk_like_dev(char opa, char opb, int64_t m, int64_t n, int64_t k, double alpha,
                                      const double* A, int64_t lda, const double* B, int64_t ldb, double beta,
                                      double* C, int64_t ldc, int tri) {
  ApiLock api_lock;
  int oa = opcode(opa), ob = opcode(opb);
  if (oa < 0 || ob < 0 || m < 0 || n < 0 || k < 0 || tri < 0 || tri > 2) return LLACUDA_ERR_BAD_ARG;
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  return dgemm_dev(oa, ob, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri, c->stream);
}

extern "C" int llacuda_dgemm_dev(char opa, char opb, int64_t m, int64_t n, int64_t k, double alpha,
                                 const double* A, int64_t lda, const double* B, int64_t ldb, double beta,
                                 double* C, int64_t ldc) {
  ApiLock api_lock;
  return llacuda_dsyrk_like_dev(opa, opb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, 0);
}
