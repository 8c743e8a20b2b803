// lla_b200/csrc/gemm_f32.cu -- SGEMM on the 5th-generation tensor cores (tcgen05 / TMEM / TMA).
//
// Replaces the `sgemm_` call of LLA's mm / gemm! for single-float arrays (reference
// src/linear-algebra.lisp:50-54, src/blas.lisp:56-63).  kind::tf32 keeps 10 mantissa bits, so the
// product is computed as an error-compensated 3xTF32 split,
//     A = Ah + Al,  B = Bh + Bl,   C += Ah Bh + Ah Bl + Al Bh      (Al Bl ~ 2^-21 |a||b| is dropped),
// with FP32 accumulation in tensor memory.  Per k-block of 32:
//   warp 0      TMA producer: cp.async.bulk.tensor tiles of op(A), op(B) (128B swizzle) -> smem ring
//   warps 2..5  splitter: every fp32 word x of the landed tiles becomes xh = x & 0xffffe000 (written
//               back in place, so the tensor core's own TF32 conversion is exact whichever way it
//               rounds) and xl = x - xh (second buffer, same offsets => swizzle-agnostic)
//   warp 1      one elected thread issues 3 x 4 tcgen05.mma.kind::tf32 (M=128, N=128, K=8) from
//               shared-memory descriptors into a 128-column TMEM accumulator, tcgen05.commit frees
//               the stage
//   warps 6..9  accumulators: the tensor core sums in FP32 with truncation, which biases long
//               same-sign sums, so every 4 k-blocks the partial sum is read out of TMEM (tcgen05.ld,
//               two ping-pong accumulators) and added to FP32 registers with round-to-nearest;
//               finally C = alpha*acc + beta*C, coalesced stores
// Operand "major-ness" follows the BLAS transposition flags: a column-major, non-transposed A is
// MN-major (four 32-wide TMA boxes per stage), a transposed A is K-major (one box); B the other
// way round.  Edges are zero-filled by TMA.
#include "internal.h"
#include "../../include/llacuda.h"

#include <cuda.h>
#include <cstring>

namespace llacuda {

namespace sg {
constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;              // 16 KiB, same for A and B tiles
constexpr int STAGE_BYTES = 4 * TILE_BYTES;          // A raw, A lo, B raw, B lo
constexpr int NUM_THREADS = 320;                     // 10 warps: TMA, MMA, 4 splitters, 4 accumulators
constexpr int TMEM_COLS = 256;                       // two 128-column accumulators, ping-pong per K chunk
constexpr int KC = 4;                                // k-blocks (of 32) accumulated in TMEM before a flush
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// shared-memory matrix descriptor (PTX ISA "tcgen05 shared memory descriptor"): 128B swizzle,
// version 1, offsets in 16-byte units
// layout type 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B (the only layout the hardware
// accepts for MN-major 32-bit operands: 128-byte rows, swizzle atom of 4 rows with 32-byte granularity)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)(lbo16 & 0x3fff) << 16;
  d |= (uint64_t)(sbo16 & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
}  // namespace sg

struct SgemmParams {
  int M, N, K;
  float* C;
  int64_t ldc;
  float alpha, beta;
  int a_mn_major, b_mn_major;   // 1: operand tile is MN-major in shared memory (4 boxes), 0: K-major (1 box)
};

__global__ void __launch_bounds__(sg::NUM_THREADS, 1)
lla_sgemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, SgemmParams p) {
  using namespace sg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // 128B swizzle atoms need 1024-byte alignment
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  // barrier layout (8 bytes each): full[s], conv[s], empty[s], acc_full[2], acc_empty[2], then the TMEM base slot
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto conv_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto acc_full_bar = [&](int b) { return bars + 8u * (3 * STAGES + b); };
  auto acc_empty_bar = [&](int b) { return bars + 8u * (3 * STAGES + 2 + b); };
  const uint32_t tmem_slot = bars + 8u * (3 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(base_ptr + STAGES * STAGE_BYTES + 8 * (3 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int KT = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(conv_bar(s), 4); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; b++) { mbar_init(acc_full_bar(b), 1); mbar_init(acc_empty_bar(b), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int kt = 0; kt < KT; kt++) {
        const int s = kt % STAGES;
        if (kt >= STAGES) mbar_wait(empty_bar(s), ((kt / STAGES) - 1) & 1);
        const uint32_t sA = base + s * STAGE_BYTES, sB = sA + 2 * TILE_BYTES;
        mbar_expect_tx(full_bar(s), 2 * TILE_BYTES);
        const int k0 = kt * BK;
        if (p.a_mn_major) {
          for (int i = 0; i < 4; i++) tma_load_2d(sA + i * (BK * 128), &tmA, full_bar(s), m0 + 32 * i, k0);
        } else {
          tma_load_2d(sA, &tmA, full_bar(s), k0, m0);
        }
        if (p.b_mn_major) {
          for (int i = 0; i < 4; i++) tma_load_2d(sB + i * (BK * 128), &tmB, full_bar(s), n0 + 32 * i, k0);
        } else {
          tma_load_2d(sB, &tmB, full_bar(s), k0, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D = F32, A = B = TF32, majors, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn_major << 15) |
                           ((uint32_t)p.b_mn_major << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    for (int kt = 0; kt < KT; kt++) {
      const int s = kt % STAGES;
      // The tensor core accumulates in FP32 with truncation, which biases long same-sign sums; so at
      // most KC k-blocks are summed in TMEM, then the accumulator warps add the partial result to
      // their FP32 registers (round to nearest) while the next chunk goes to the other TMEM buffer.
      const int chunk = kt / KC, buf = chunk & 1, use = chunk >> 1;
      if (kt % KC == 0 && use >= 1) mbar_wait(acc_empty_bar(buf), (use - 1) & 1);
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
      mbar_wait(conv_bar(s), (kt / STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      if (lane == 0) {
        const uint32_t sA = base + s * STAGE_BYTES, sAl = sA + TILE_BYTES, sB = sA + 2 * TILE_BYTES, sBl = sA + 3 * TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / 8; kk++) {
          // K-major: 8 tf32 = 32 bytes further along the swizzled 128-byte row; MN-major: next group of 8 k-rows
          const uint32_t offA = p.a_mn_major ? kk * 1024u : kk * 32u;
          const uint32_t offB = p.b_mn_major ? kk * 1024u : kk * 32u;
          // K-major: 8-row groups 1024 B apart (SBO 64).  MN-major: 32-element MN groups one TMA box
          // (BK*128 B) apart (LBO), 4-row K atoms 512 B apart (SBO 32).
          const uint32_t lboA = p.a_mn_major ? (BK * 128) / 16 : 1, lboB = p.b_mn_major ? (BK * 128) / 16 : 1;
          const uint32_t sboA = p.a_mn_major ? 32 : 64, sboB = p.b_mn_major ? 32 : 64;
          const uint32_t ltA = p.a_mn_major ? 1 : 2, ltB = p.b_mn_major ? 1 : 2;
          const uint64_t dA = make_desc(sA + offA, lboA, sboA, ltA), dAl = make_desc(sAl + offA, lboA, sboA, ltA);
          const uint64_t dB = make_desc(sB + offB, lboB, sboB, ltB), dBl = make_desc(sBl + offB, lboB, sboB, ltB);
          umma_tf32(tmem_d, dAl, dB, idesc, ((kt % KC) | kk) != 0);   // small terms first
          umma_tf32(tmem_d, dA, dBl, idesc, 1);
          umma_tf32(tmem_d, dA, dB, idesc, 1);
        }
        umma_commit(empty_bar(s));
        if (kt % KC == KC - 1 || kt == KT - 1) umma_commit(acc_full_bar(buf));
      }
      __syncwarp();
    }
  } else if (warp < 6) {
    // ===================== splitter (warps 2..5) =====================
    const int t = threadIdx.x - 64;   // 0..127
    for (int kt = 0; kt < KT; kt++) {
      const int s = kt % STAGES;
      mbar_wait(full_bar(s), (kt / STAGES) & 1);
      uint8_t* stage = base_ptr + s * STAGE_BYTES;
#pragma unroll
      for (int half = 0; half < 2; half++) {            // A tile, then B tile
        uint4* raw = (uint4*)(stage + half * 2 * TILE_BYTES);
        uint4* lo = (uint4*)(stage + half * 2 * TILE_BYTES + TILE_BYTES);
#pragma unroll
        for (int it = 0; it < TILE_BYTES / 16 / 128; it++) {
          const int idx = t + it * 128;
          uint4 v = raw[idx], h, l;
          h.x = v.x & 0xffffe000u; h.y = v.y & 0xffffe000u; h.z = v.z & 0xffffe000u; h.w = v.w & 0xffffe000u;
          l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
          l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
          l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
          l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
          raw[idx] = h;
          lo[idx] = l;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> visible to UMMA
      __syncwarp();
      if (lane == 0) mbar_arrive(conv_bar(s));
    }
  } else {
    // ===================== accumulators + epilogue (warps 6..9); TMEM lane quarter = warp index mod 4
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int64_t gm = (int64_t)m0 + row;
    float acc[BN];
#pragma unroll
    for (int c = 0; c < BN; c++) acc[c] = 0.0f;
    const int NCH = (KT + KC - 1) / KC;
    for (int chunk = 0; chunk < NCH; chunk++) {
      const int buf = chunk & 1, use = chunk >> 1;
      mbar_wait(acc_full_bar(buf), use & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
        for (int c = 0; c < 32; c++) acc[c0 + c] += __uint_as_float(r[c]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty_bar(buf));
    }
    if (gm < p.M) {
#pragma unroll
      for (int c = 0; c < BN; c++) {
        const int64_t gn = (int64_t)n0 + c;
        if (gn < p.N) {
          float v = p.alpha * acc[c];
          float* dst = p.C + gm + gn * p.ldc;
          if (p.beta != 0.0f) v += p.beta * *dst;
          *dst = v;
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

__global__ void lla_sscale_c_kernel(int64_t M, int64_t N, float* C, int64_t ldc, float beta) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t j = blockIdx.y;
  if (i >= M) return;
  float* c = C + i + j * ldc;
  *c = (beta == 0.0f) ? 0.0f : beta * *c;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// X: column-major rows x cols fp32 matrix (leading dimension ld).  The tensor map's inner dimension is
// the contiguous one (rows); box = {32, box_outer}.
static int make_map(CUtensorMap* map, const float* X, int64_t rows, int64_t cols, int64_t ld, int box_outer,
                    bool mn_major) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled not available", -1); return LLACUDA_ERR_BAD_ARG; }
  cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)X, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled failed", (int)r); return LLACUDA_ERR_BAD_ARG; }
  return 0;
}

int sgemm_dev(int opa, int opb, int64_t m, int64_t n, int64_t k, float alpha, const float* A, int64_t lda, const float* B,
              int64_t ldb, float beta, float* C, int64_t ldc, cudaStream_t st) {
  if (m <= 0 || n <= 0) return 0;
  if (!ctx()) return LLACUDA_ERR_NO_DEVICE;
  if (k <= 0 || alpha == 0.0f) {
    if (beta == 1.0f) return 0;
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)n);
    lla_sscale_c_kernel<<<grid, 256, 0, st>>>(m, n, C, ldc, beta);
    count_launch();
    LLA_CUDA(cudaGetLastError());
    return 0;
  }
  if ((lda & 3) || (ldb & 3) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15))
    // TMA needs 16-byte aligned operands (lda, ldb multiples of 4); LLA's sloppy arrays (padded leading dimensions,
    // reference tests/blas.lisp:13-17) need not be: those go to the FP32-FMA tile kernel instead of being refused
    return sgemm_simt_dev(opa, opb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, st);
  const bool ta = opa != 0, tb = opb != 0;
  CUtensorMap tmA, tmB;
  // A stored (ta ? k x m : m x k), B stored (tb ? n x k : k x n); the contiguous dimension is the inner one
  if (!ta) LLA_TRY(make_map(&tmA, A, m, k, lda, sg::BK, true));       // MN-major: box {32 of M, 32 of K}
  else LLA_TRY(make_map(&tmA, A, k, m, lda, sg::BM, false));           // K-major:  box {32 of K, 128 of M}
  if (!tb) LLA_TRY(make_map(&tmB, B, k, n, ldb, sg::BN, false));       // K-major:  box {32 of K, 128 of N}
  else LLA_TRY(make_map(&tmB, B, n, k, ldb, sg::BK, true));           // MN-major: box {32 of N, 32 of K}
  SgemmParams p;
  p.M = (int)m; p.N = (int)n; p.K = (int)k; p.C = C; p.ldc = ldc; p.alpha = alpha; p.beta = beta;
  p.a_mn_major = ta ? 0 : 1;
  p.b_mn_major = tb ? 1 : 0;
  static DeviceOnce once;
  if (once.first()) {
    LLA_CUDA(cudaFuncSetAttribute(lla_sgemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sg::SMEM_BYTES));
  }
  dim3 grid((unsigned)((m + sg::BM - 1) / sg::BM), (unsigned)((n + sg::BN - 1) / sg::BN));
  lla_sgemm_tcgen05_kernel<<<grid, sg::NUM_THREADS, sg::SMEM_BYTES, st>>>(tmA, tmB, p);
  count_launch();
  LLA_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace llacuda

using namespace llacuda;

static int sopcode(char c) {
  switch (c) {
    case 'N': case 'n': return 0;
    case 'T': case 't': return 1;
    case 'C': case 'c': return 2;
    default: return -1;
  }
}

extern "C" int llacuda_sgemm_dev(char opa, char opb, int64_t m, int64_t n, int64_t k, float alpha, const float* A,
                                 int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc) {
  ApiLock api_lock;
  Context* c = ctx();
  if (!c) return LLACUDA_ERR_NO_DEVICE;
  int oa = sopcode(opa), ob = sopcode(opb);
  if (oa < 0 || ob < 0 || m < 0 || n < 0 || k < 0) return LLACUDA_ERR_BAD_ARG;
  return sgemm_dev(oa, ob, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, c->stream);
}
