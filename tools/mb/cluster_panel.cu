// Prototype for the next round's LU panel engine -- NOT part of libllacuda, never run yet:
// written at the end of round 1 after the GPU budget was spent; it compiles for sm_100a and checks itself against a CPU
// DGETF2 when run.  Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o cluster_panel cluster_panel.cu
//
// Idea under test: the per-column pivot exchange of the cooperative panel kernel costs 2.1-3.7k cycles through L2 for
// 48-128 CTAs (xchg.cu).  One thread-block cluster of 16 CTAs x 1024 threads holds an m x 16 panel (m <= 16384) with one
// row per thread in registers; every CTA pushes its candidate (key, row, the candidate's 16 values, and CTA 0 the old row j)
// into every CTA's shared-memory inbox with st.shared::cluster, one cluster barrier per column, and each CTA picks the same
// winner locally: no second round trip, no global memory on the chain.  Arithmetic as DGETF2 / DGER (reciprocal scaling,
// separately rounded multiply and subtract), first maximum on ties, so ipiv and the factors must equal the CPU restatement
// bit for bit.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;

constexpr int W = 16;            // panel width
constexpr int ROWS = 1024;       // rows (= threads) per CTA
constexpr int NCTA = 16;         // cluster size (non-portable)
constexpr int REC = 2 + 2 * W;   // doubles per candidate record: key, row, candidate row, old row j

struct Inbox { double rec[2][NCTA][REC]; };   // [parity][source CTA][...]

__global__ void __launch_bounds__(ROWS, 1) cluster_panel_kernel(double* __restrict__ A, long long lda, int m, int* __restrict__ ipiv,
                                                                int* __restrict__ info) {
  cg::cluster_group cluster = cg::this_cluster();
  const int cta = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row = cta * ROWS + tid;
  const bool live = row < m;
  __shared__ Inbox inbox;
  __shared__ unsigned long long s_key[ROWS / 32];
  __shared__ int s_row[ROWS / 32];
  __shared__ double s_wrow[ROWS / 32][W];   // each warp's candidate row
  __shared__ double s_rowj[W];              // CTA 0: row j before the interchange

  double x[W];
#pragma unroll
  for (int c = 0; c < W; c++) x[c] = live ? A[row + c * lda] : 0.0;
  const int nsteps = m < W ? m : W;

#pragma unroll
  for (int j = 0; j < W; j++) {
    if (j < nsteps) {
      const int par = j & 1;
      // ---- warp candidate: first row >= j with the largest |x[j]| (NaN only wins as row j: IDAMAX)
      unsigned long long key = 0ull;
      unsigned r = 0x7fffffffu;
      if (live && row >= j) {
        double v = fabs(x[j]);
        const bool isn = (v != v);
        if (row == j && isn) v = INFINITY;
        if (v == v) key = (unsigned long long)__double_as_longlong(v) + 1ull;
        r = (unsigned)row;
      }
      const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
      const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
      const bool c1 = hi == mh;
      const unsigned ml = __reduce_max_sync(0xffffffffu, c1 ? lo : 0u);
      const bool c2 = c1 && lo == ml;
      const unsigned rm = __reduce_min_sync(0xffffffffu, c2 ? r : 0x7fffffffu);
      if (lane == 0) { s_key[warp] = ((unsigned long long)mh << 32) | ml; s_row[warp] = (int)rm; }
      if (rm != 0x7fffffffu && (unsigned)row == rm) {
#pragma unroll
        for (int c = 0; c < W; c++) s_wrow[warp][c] = x[c];
      }
      if (cta == 0 && row == j) {
#pragma unroll
        for (int c = 0; c < W; c++) s_rowj[c] = x[c];
      }
      __syncthreads();
      // ---- CTA candidate (warp 0), pushed into every CTA's inbox
      if (warp == 0) {
        const unsigned long long k32 = s_key[lane];
        const unsigned r32 = (unsigned)s_row[lane];
        const unsigned h2 = (unsigned)(k32 >> 32), l2 = (unsigned)k32;
        const unsigned mh2 = __reduce_max_sync(0xffffffffu, h2);
        const bool d1 = h2 == mh2;
        const unsigned ml2 = __reduce_max_sync(0xffffffffu, d1 ? l2 : 0u);
        const bool d2 = d1 && l2 == ml2;
        const unsigned br = __reduce_min_sync(0xffffffffu, d2 ? r32 : 0x7fffffffu);
        const int bw = (br == 0x7fffffffu) ? 0 : (((int)br - cta * ROWS) >> 5);
        // record word `lane` (and lane + 32 for the two words beyond 32)
        double w0, w1 = 0.0;
        if (lane == 0) w0 = __longlong_as_double((long long)(((unsigned long long)mh2 << 32) | ml2));
        else if (lane == 1) w0 = __longlong_as_double((long long)br);
        else if (lane < 2 + W) w0 = s_wrow[bw][lane - 2];
        else w0 = s_rowj[lane - 2 - W];                       // lanes 18..31 -> rowj[0..13]
        if (lane < 2) w1 = s_rowj[W - 2 + lane];              // words 32, 33 -> rowj[14], rowj[15]
        for (int dst = 0; dst < NCTA; dst++) {
          Inbox* remote = cluster.map_shared_rank(&inbox, dst);
          remote->rec[par][cta][lane] = w0;
          if (lane < 2) remote->rec[par][cta][32 + lane] = w1;
        }
      }
      cluster.sync();   // one cluster barrier per column: every record of this column has landed everywhere
      // ---- every thread picks the winner from its CTA's inbox (16 records, broadcast reads)
      unsigned long long bk = 0ull;
      int ws = 0;
#pragma unroll
      for (int s = 0; s < NCTA; s++) {
        const unsigned long long ks = (unsigned long long)__double_as_longlong(inbox.rec[par][s][0]);
        if (ks > bk) { bk = ks; ws = s; }                    // strict >: the lowest CTA (= lowest row) wins ties
      }
      const double* rec = inbox.rec[par][ws];
      const int p = (int)__double_as_longlong(rec[1]);
      const double piv = rec[2 + j];
      if (cta == 0 && tid == 0) {
        ipiv[j] = p + 1;
        if (piv == 0.0 && *info == 0) *info = j + 1;
      }
      const double* rowj = inbox.rec[par][0] + 2 + W;        // old row j travels with CTA 0's record
      if (piv != 0.0 && p != j) {
        if (row == j) {
#pragma unroll
          for (int c = 0; c < W; c++) x[c] = rec[2 + c];
        } else if (row == p) {
#pragma unroll
          for (int c = 0; c < W; c++) x[c] = rowj[c];
        }
      }
      if (live && row > j) {
        double l = x[j];
        if (piv != 0.0) l = (fabs(piv) >= DBL_MIN) ? x[j] * (1.0 / piv) : x[j] / piv;
        x[j] = l;
#pragma unroll
        for (int c = j + 1; c < W; c++) x[c] = __dsub_rn(x[c], __dmul_rn(l, rec[2 + c]));   // rec = row j after the interchange
      }
      __syncthreads();  // s_key / s_wrow are rewritten by the next column
    }
  }
  if (live) {
#pragma unroll
    for (int c = 0; c < W; c++) A[row + c * lda] = x[c];
  }
}

static void cpu_getf2(int m, int n, std::vector<double>& a, int lda, std::vector<int>& ipiv, int& info) {
  info = 0;
  for (int j = 0; j < (m < n ? m : n); j++) {
    int p = j; double best = -1.0;
    for (int i = j; i < m; i++) { double v = fabs(a[i + (size_t)j * lda]); if (v > best) { best = v; p = i; } }
    ipiv[j] = p + 1;
    const double piv = a[p + (size_t)j * lda];
    if (piv != 0.0) {
      if (p != j) for (int c = 0; c < n; c++) std::swap(a[j + (size_t)c * lda], a[p + (size_t)c * lda]);
      const double rinv = 1.0 / piv;
      for (int i = j + 1; i < m; i++) a[i + (size_t)j * lda] = (fabs(piv) >= DBL_MIN) ? a[i + (size_t)j * lda] * rinv : a[i + (size_t)j * lda] / piv;
    } else if (info == 0) info = j + 1;
    for (int c = j + 1; c < n; c++) {
      const double u = a[j + (size_t)c * lda];
      for (int i = j + 1; i < m; i++) {
        volatile double prod = a[i + (size_t)j * lda] * u;      // separately rounded, as DGER without FMA contraction
        a[i + (size_t)c * lda] = a[i + (size_t)c * lda] - prod;
      }
    }
  }
}

int main() {
  cudaFuncSetAttribute(cluster_panel_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int m : {16, 1000, 4096, 6144, 16384}) {
    const int lda = m;
    std::vector<double> h((size_t)lda * W), ref;
    srand(m);
    for (auto& v : h) v = 10.0 * rand() / RAND_MAX;
    ref = h;
    std::vector<int> piv_ref(W), piv(W);
    int info_ref = 0;
    cpu_getf2(m, W, ref, lda, piv_ref, info_ref);
    double* dA; int *dP, *dI;
    cudaMalloc(&dA, h.size() * 8); cudaMalloc(&dP, W * 4); cudaMalloc(&dI, 4);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(NCTA); cfg.blockDim = dim3(ROWS);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = NCTA; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 20; rep++) {
      cudaMemcpy(dA, h.data(), h.size() * 8, cudaMemcpyHostToDevice); cudaMemset(dI, 0, 4);
      cudaEventRecord(e0);
      cudaError_t e = cudaLaunchKernelEx(&cfg, cluster_panel_kernel, dA, (long long)lda, m, dP, dI);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("m=%d: launch failed: %s\n", m, cudaGetErrorString(e)); break; }
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    std::vector<double> out(h.size());
    cudaMemcpy(out.data(), dA, h.size() * 8, cudaMemcpyDeviceToHost); cudaMemcpy(piv.data(), dP, W * 4, cudaMemcpyDeviceToHost);
    bool piv_ok = piv == piv_ref, val_ok = true;
    for (size_t i = 0; i < out.size(); i++) if (out[i] != ref[i]) { val_ok = false; break; }
    printf("m=%5d w=%d: %.1f us per panel (%.2f us per column)  ipiv %s  factors %s\n", m, W, best * 1e3, best * 1e3 / W,
           piv_ok ? "bit-exact" : "DIFFER", val_ok ? "bit-identical" : "differ");
    cudaFree(dA); cudaFree(dP); cudaFree(dI);
  }
  return 0;
}
