// lla_b200/csrc/dist_algos.cu -- SPMD rank functions of the multi-GPU engine: SUMMA GEMM, block-cyclic Cholesky and
// LU, triangular solves on replicated right-hand sides (BASELINE.json configs[2] and [4]).
//
// The local arithmetic is the single-GPU library (DMMA GEMM, shared-memory Cholesky leaves, cooperative LU panel,
// inverse-based TRSM); this file adds the ownership arithmetic, the panel traffic (NCCL broadcasts on the grid's row /
// column communicators over NVLink) and the look-ahead: panels and collectives run on the rank's high-priority
// `side` stream, trailing updates on `main`, events hand the double-buffered panels over.
#include "dist.h"
#include "staging.h"
#include "../../include/llacuda.h"

#include <cstdlib>
#include <vector>

namespace llacuda {

// ------------------------------------------------------------------ small kernels
// copies n sub-blocks (rows x cols, column-major) between two buffers: block b from src + soff[b] to dst + doff[b]
constexpr int MAXBL = 120;
struct BlockList {
  int n;
  long long soff[MAXBL], doff[MAXBL];
};

template <class T>
__global__ void __launch_bounds__(256) lla_copy_blocks_kernel(const T* __restrict__ src, int64_t lds, T* __restrict__ dst,
                                                              int64_t ldd, int rows, int cols, BlockList L) {
  const int b = blockIdx.y;
  const T* s = src + L.soff[b];
  T* d = dst + L.doff[b];
  const int c0 = blockIdx.x * 8;
  for (int c = c0; c < c0 + 8 && c < cols; c++)
    for (int r = threadIdx.x; r < rows; r += 256) d[r + (int64_t)c * ldd] = s[r + (int64_t)c * lds];
}

template <class T>
static int copy_blocks(const T* src, int64_t lds, T* dst, int64_t ldd, int rows, int cols,
                       const std::vector<long long>& soff, const std::vector<long long>& doff, cudaStream_t st) {
  const size_t n = soff.size();
  for (size_t b0 = 0; b0 < n; b0 += MAXBL) {
    BlockList L;
    L.n = (int)(n - b0 < (size_t)MAXBL ? n - b0 : (size_t)MAXBL);
    for (int i = 0; i < L.n; i++) { L.soff[i] = soff[b0 + i]; L.doff[i] = doff[b0 + i]; }
    dim3 grid((unsigned)((cols + 7) / 8), (unsigned)L.n);
    lla_copy_blocks_kernel<T><<<grid, 256, 0, st>>>(src, lds, dst, ldd, rows, cols, L);
    count_launch();
    LLA_CUDA(cudaGetLastError());
  }
  return 0;
}

// INFO bookkeeping on the device: no host round trip inside a factorisation
__global__ void lla_info_acc_kernel(int* acc, const int* last, int offset) {
  if (*acc == 0 && *last != 0) *acc = *last + offset;
}
__global__ void lla_info_encode_kernel(int* flag, bool decode) {   // 0 <-> INT_MAX so that a MIN all-reduce picks the first failure
  if (!decode) { if (*flag <= 0) *flag = 0x7fffffff; }
  else if (*flag == 0x7fffffff) *flag = 0;
}
void launch_info_encode(int* flag, bool decode, cudaStream_t st) {
  lla_info_encode_kernel<<<1, 1, 0, st>>>(flag, decode);
  count_launch();
}
// ipiv[off + i] = piv[i] + off (panel-relative 1-based pivots -> LAPACK's global ones)
__global__ void lla_ipiv_shift_kernel(int* __restrict__ ipiv, const int* __restrict__ piv, int off, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ipiv[off + i] = piv[i] + off;
}
// X = B - S on an nb x nrhs block (ldx = nb)
__global__ void lla_sub_block_kernel(double* __restrict__ X, const double* __restrict__ B, int64_t ldb,
                                     const double* __restrict__ S, int rows, int cols) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  if (r < rows && c < cols) X[r + (int64_t)c * rows] = B[r + (int64_t)c * ldb] - S[r + (int64_t)c * rows];
}

static int d2d_2d(void* dst, int64_t ldd, const void* src, int64_t lds, int64_t rows, int64_t cols, size_t es, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return 0;
  LLA_CUDA(cudaMemcpy2DAsync(dst, (size_t)ldd * es, src, (size_t)lds * es, (size_t)rows * es, (size_t)cols, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// ------------------------------------------------------------------ SUMMA
template <class T> struct Local;
template <> struct Local<double> {
  static constexpr int words = 1;
  static int gemm(int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda, const double* B, int64_t ldb,
                  double beta, double* C, int64_t ldc, cudaStream_t st) {
    return dgemm_dev(0, 0, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, TRI_FULL, st);
  }
};
template <> struct Local<cuDoubleComplex> {
  static constexpr int words = 2;
  static int gemm(int64_t m, int64_t n, int64_t k, cuDoubleComplex alpha, const cuDoubleComplex* A, int64_t lda,
                  const cuDoubleComplex* B, int64_t ldb, cuDoubleComplex beta, cuDoubleComplex* C, int64_t ldc, cudaStream_t st) {
    return zgemm_dev(0, 0, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, TRI_FULL, st);
  }
};
static inline double one_of(double) { return 1.0; }
static inline cuDoubleComplex one_of(cuDoubleComplex) { return make_cuDoubleComplex(1.0, 0.0); }

// C = alpha A B + beta C.  Step s: the grid column that owns k-block s broadcasts its mloc x nb slice of A along the
// grid rows, the grid row that owns it broadcasts its nb x nloc slice of B along the grid columns; everybody
// multiplies.  The broadcasts of step s+1 are in flight (side stream) while the DMMA GEMM of step s runs (main).
template <class T>
int summa_rank(Rank& R, int64_t Mp, int64_t Np, int64_t Kp, int64_t nb, T alpha, const T* A, int64_t lda, const T* B,
               int64_t ldb, T beta, T* C, int64_t ldc) {
  const NcclApi* nc = nccl();
  const int P = R.P, Q = R.Q, p = R.p, q = R.q;
  const int64_t mloc = Mp / P, nloc = Np / Q, steps = Kp / nb;
  if (Mp % (nb * P) || Np % (nb * Q) || Kp % (nb * lcm(P, Q))) {
    set_error(__FILE__, __LINE__, "summa: padded dimensions must be multiples of nb * grid", (int)nb);
    return LLACUDA_ERR_BAD_ARG;
  }
  PoolBuf abuf[2], bbuf[2];
  for (int i = 0; i < 2; i++) {
    LLA_TRY(abuf[i].alloc((size_t)mloc * nb * sizeof(T)));
    LLA_TRY(bbuf[i].alloc((size_t)nloc * nb * sizeof(T)));
  }
  cudaEvent_t ev_ready[2] = {R.ev[0], R.ev[1]}, ev_free[2] = {R.ev[2], R.ev[3]}, ev_start = R.ev[4];
  LLA_CUDA(cudaEventRecord(ev_start, R.main));
  LLA_CUDA(cudaStreamWaitEvent(R.side, ev_start, 0));
  const size_t w = Local<T>::words;
  auto post = [&](int64_t s) -> int {
    const int slot = (int)(s & 1);
    const int qk = (int)(s % Q), pk = (int)(s % P);
    if (s >= 2) LLA_CUDA(cudaStreamWaitEvent(R.side, ev_free[slot], 0));
    T* ab = (T*)abuf[slot].p;
    T* bb = (T*)bbuf[slot].p;
    const T* asrc = A + (s / Q) * nb * lda;           // my slice when q == qk: mloc x nb, contiguous iff lda == mloc
    if (q == qk && (Q == 1 || lda != mloc)) LLA_TRY(d2d_2d(ab, mloc, asrc, lda, mloc, nb, sizeof(T), R.side));
    if (p == pk) LLA_TRY(d2d_2d(bb, nb, B + (s / P) * nb, ldb, nb, nloc, sizeof(T), R.side));
    if (Q > 1) {
      const void* send = (q == qk && lda == mloc) ? (const void*)asrc : (const void*)ab;
      LLA_NCCL(nc->Broadcast(send, ab, (size_t)mloc * nb * w, ncclDouble, qk, R.rowc, R.side));
    }
    if (P > 1) LLA_NCCL(nc->Broadcast(bb, bb, (size_t)nloc * nb * w, ncclDouble, pk, R.colc, R.side));
    LLA_CUDA(cudaEventRecord(ev_ready[slot], R.side));
    return 0;
  };
  LLA_TRY(post(0));
  for (int64_t s = 0; s < steps; s++) {
    const int slot = (int)(s & 1);
    if (s + 1 < steps) LLA_TRY(post(s + 1));
    LLA_CUDA(cudaStreamWaitEvent(R.main, ev_ready[slot], 0));
    LLA_TRY(Local<T>::gemm(mloc, nloc, nb, alpha, (const T*)abuf[slot].p, mloc, (const T*)bbuf[slot].p, nb,
                           s == 0 ? beta : one_of(alpha), C, ldc, R.main));
    LLA_CUDA(cudaEventRecord(ev_free[slot], R.main));
  }
  LLA_CUDA(cudaEventRecord(ev_start, R.side));
  LLA_CUDA(cudaStreamWaitEvent(R.main, ev_start, 0));
  // the pooled panels go back when this frame ends: everything that reads them must be done
  LLA_CUDA(cudaStreamSynchronize(R.main));
  return 0;
}
template int summa_rank<double>(Rank&, int64_t, int64_t, int64_t, int64_t, double, const double*, int64_t, const double*,
                                int64_t, double, double*, int64_t);
template int summa_rank<cuDoubleComplex>(Rank&, int64_t, int64_t, int64_t, int64_t, cuDoubleComplex, const cuDoubleComplex*,
                                         int64_t, const cuDoubleComplex*, int64_t, cuDoubleComplex, cuDoubleComplex*, int64_t);

// ------------------------------------------------------------------ Cholesky, A = U^T U (LLA's call site: POTRF 'U')
// Block step k on the P x Q grid:
//   1. the owner of the diagonal block factors it (shared-memory leaves + DMMA) and broadcasts U_kk along its grid row;
//   2. that grid row solves U_kk^T X = A(k, J > k) on its own blocks of block row k (inverse-based TRSM = DMMA GEMMs);
//   3. the finished block row travels on the sub-communicators only: each rank of the owning grid row broadcasts its
//      slice down its grid column (these are the U(k, J) every rank multiplies from the right), then the ranks of one
//      grid row exchange among themselves the blocks U(k, I) with I on that grid row (the left factors);
//   4. update A(I, J) -= U(k, I)^T U(k, J), k < I <= J: one DMMA GEMM per local block column over the block rows
//      up to the diagonal.  Blocks below the diagonal are scratch, so no masking is needed.
// Look-ahead: the update of block row k+1 goes first; panel k+1 (kernels and collectives) then runs on `side`
// under the rest of update k.
// Block columns (1 x Q, the default grid): every row of a block column is local, and the critical chain of the
// factorisation is diag(k) -> U(k, k+1) -> diag(k+1).  The diagonal blocks therefore get their updates on a stream of
// their own ("chain"): as soon as a rank has solved its piece of block row k it subtracts U(k, J)^T U(k, J) from its own
// diagonal blocks (the one the next step factors first), so the owner of block column k+1 can factor and broadcast
// diag(k+1) while block row k is still being exchanged and the trailing update of step k runs.  The diagonal broadcasts
// use the world communicator, the row exchange the row communicator: two communicators progress independently.
static int pdpotrf_rank_cols(Rank& R, int64_t Np, int64_t nb, double* A, int64_t lld) {
  const NcclApi* nc = nccl();
  const int Q = R.Q, q = R.q;
  const int64_t nloc = Np / Q, nblk = Np / nb, nbl = nloc / nb;
  PoolBuf wcb[2], wrb[2], xchg, dblk[2], tinv[2];
  for (int i = 0; i < 2; i++) {
    LLA_TRY(wcb[i].alloc((size_t)nloc * nb * 8));
    LLA_TRY(wrb[i].alloc((size_t)Np * nb * 8));
    LLA_TRY(dblk[i].alloc((size_t)nb * nb * 8));
    LLA_TRY(tinv[i].alloc(tinv_bytes(nb)));
  }
  LLA_TRY(xchg.alloc((size_t)Np * nb * 8));
  int* info_acc = R.dflag;
  int* info_last = R.dflag + 1;
  LLA_CUDA(cudaMemsetAsync(R.dflag, 0, 2 * sizeof(int), R.main));
  cudaEvent_t evD[2] = {R.ev[0], R.ev[1]}, evTrsm[2] = {R.ev[2], R.ev[3]}, evRow[2] = {R.ev[4], R.ev[5]},
              evUpd[2] = {R.ev[6], R.ev[7]}, evLook = R.ev[8], evStart = R.ev[9], evSyrk = R.ev[10];
  cudaStream_t sm = R.main, sd = R.side, sc = R.chain;
  LLA_CUDA(cudaEventRecord(evStart, sm));
  LLA_CUDA(cudaStreamWaitEvent(sd, evStart, 0));
  LLA_CUDA(cudaStreamWaitEvent(sc, evStart, 0));
  // diag(J, J) -= U(k, J)^T U(k, J) for my block columns lj in [a, b)
  auto syrk = [&](int64_t k, int64_t a, int64_t b) -> int {
    for (int64_t lj = a; lj < b; lj++) {
      const int64_t J = lj * Q + q;
      const double* Ukj = A + k * nb + lj * nb * lld;
      LLA_TRY(dgemm_dev(1, 0, nb, nb, nb, -1.0, Ukj, lld, Ukj, lld, 1.0, A + J * nb + lj * nb * lld, lld, TRI_UPPER, sc));
    }
    return 0;
  };
  for (int64_t k = 0; k < nblk; k++) {
    const int slot = (int)(k & 1);
    const int qk = (int)(k % Q);
    double* D = (double*)dblk[slot].p;
    double* wc = (double*)wcb[slot].p;
    double* wr = (double*)wrb[slot].p;
    const int64_t lj0 = blocks_le(k, q, Q);           // my first block column with J > k
    const int64_t ncols = nloc - lj0 * nb;
    // ---- chain: factor diag(k), broadcast it, invert its 128-blocks
    if (q == qk) {
      double* diag = A + k * nb + (k / Q) * nb * lld;
      LLA_TRY(dpotrf_dev(true, nb, diag, lld, info_last, sc));
      lla_info_acc_kernel<<<1, 1, 0, sc>>>(info_acc, info_last, (int)(k * nb));
      count_launch();
      LLA_TRY(d2d_2d(D, nb, diag, lld, nb, nb, 8, sc));
    }
    LLA_NCCL(nc->Broadcast(D, D, (size_t)nb * nb, ncclDouble, qk, R.world, sc));
    LLA_TRY(dtrtri_blocks_dev(true, false, nb, D, nb, (double*)tinv[slot].p, sc));
    LLA_CUDA(cudaEventRecord(evD[slot], sc));
    // the rest of the previous row's contributions to my diagonal blocks (the nearest one went ahead of this step's factorisation)
    if (k > 0) {
      const int64_t a = blocks_le(k, q, Q);            // columns J > k; J == k was applied before diag(k) was factored
      LLA_TRY(syrk(k - 1, a, nbl));
    }
    // ---- side: my piece of block row k, then the exchange of the whole row
    LLA_CUDA(cudaStreamWaitEvent(sd, evD[slot], 0));
    if (k > 0) LLA_CUDA(cudaStreamWaitEvent(sd, evLook, 0));
    if (k >= 2) LLA_CUDA(cudaStreamWaitEvent(sd, evUpd[slot], 0));      // update k-2 has finished with this slot's buffers
    if (ncols > 0) {
      double* row = A + k * nb + lj0 * nb * lld;
      LLA_TRY(dtrsm_left_inv_dev(true, 1, nb, ncols, D, nb, (const double*)tinv[slot].p, row, lld, sd));
      LLA_CUDA(cudaEventRecord(evTrsm[slot], sd));
      LLA_TRY(d2d_2d(wc, nb, row, lld, nb, ncols, 8, sd));
    } else {
      LLA_CUDA(cudaEventRecord(evTrsm[slot], sd));
    }
    // chain again: the contribution of row k to the diagonal block the next step factors, if it is mine
    LLA_CUDA(cudaStreamWaitEvent(sc, evTrsm[slot], 0));
    if (k + 1 < nblk && (int)((k + 1) % Q) == q) LLA_TRY(syrk(k, lj0, lj0 + 1));
    LLA_CUDA(cudaEventRecord(evSyrk, sc));
    if (k + 1 >= nblk) break;
    // exchange: every rank contributes U(k, J) for its J > k; everybody needs the whole row as left factors
    {
      double* X = (double*)xchg.p;
      std::vector<int64_t> seg((size_t)Q + 1, 0);
      for (int qq = 0; qq < Q; qq++) seg[(size_t)qq + 1] = seg[(size_t)qq] + (nbl - blocks_le(k, qq, Q));
      if (ncols > 0) LLA_TRY(d2d_2d(X + seg[(size_t)q] * nb * nb, nb, wc, nb, nb, ncols, 8, sd));
      LLA_NCCL(nc->GroupStart());
      for (int qq = 0; qq < Q; qq++) {
        const int64_t cnt = seg[(size_t)qq + 1] - seg[(size_t)qq];
        if (cnt > 0) LLA_NCCL(nc->Broadcast(X + seg[(size_t)qq] * nb * nb, X + seg[(size_t)qq] * nb * nb, (size_t)cnt * nb * nb, ncclDouble, qq, R.rowc, sd));
      }
      LLA_NCCL(nc->GroupEnd());
      std::vector<long long> soff, doff;   // interleave into block order: wr block (I - k - 1) = U(k, I)
      for (int qq = 0; qq < Q; qq++) {
        const int64_t l0 = blocks_le(k, qq, Q);
        for (int64_t lj = l0; lj < nbl; lj++) {
          soff.push_back((long long)((seg[(size_t)qq] + (lj - l0)) * nb * nb));
          doff.push_back((long long)((lj * Q + qq - k - 1) * nb * nb));
        }
      }
      LLA_TRY(copy_blocks<double>(X, nb, wr, nb, (int)nb, (int)nb, soff, doff, sd));
      LLA_CUDA(cudaEventRecord(evRow[slot], sd));
    }
    // ---- main: off-diagonal update; block row k+1 first (the next step's solve needs it)
    LLA_CUDA(cudaStreamWaitEvent(sm, evRow[slot], 0));
    {
      const int64_t l1 = blocks_le(k + 1, q, Q);        // my columns J > k+1
      if (l1 < nbl)   // A(k+1, J) -= U(k, k+1)^T U(k, J)
        LLA_TRY(dgemm_dev(1, 0, nb, (nbl - l1) * nb, nb, -1.0, wr, nb, wc + (l1 - lj0) * nb * nb, nb, 1.0,
                          A + (k + 1) * nb + l1 * nb * lld, lld, TRI_FULL, sm));
      LLA_CUDA(cudaEventRecord(evLook, sm));
      for (int64_t lj = l1; lj < nbl; lj++) {           // rows I = k+2 .. J-1 of column J
        const int64_t J = lj * Q + q;
        const int64_t rows = J - (k + 2);
        if (rows <= 0) continue;
        LLA_TRY(dgemm_dev(1, 0, rows * nb, nb, nb, -1.0, wr + nb * nb, nb, wc + (lj - lj0) * nb * nb, nb, 1.0,
                          A + (k + 2) * nb + lj * nb * lld, lld, TRI_FULL, sm));
      }
      LLA_CUDA(cudaEventRecord(evUpd[slot], sm));
    }
  }
  for (cudaStream_t t : {sd, sc}) {
    LLA_CUDA(cudaEventRecord(evStart, t));
    LLA_CUDA(cudaStreamWaitEvent(sm, evStart, 0));
  }
  LLA_CUDA(cudaStreamSynchronize(sm));   // the pooled buffers are released with this frame
  return 0;
}

int pdpotrf_rank(Rank& R, int64_t Np, int64_t nb, double* A, int64_t lld) {
  const NcclApi* nc = nccl();
  const int P = R.P, Q = R.Q, p = R.p, q = R.q;
  if (Np % (nb * lcm(P, Q)) || nb % IB) {
    set_error(__FILE__, __LINE__, "pdpotrf: the padded order must be a multiple of nb * lcm(P, Q), nb of 128", (int)nb);
    return LLACUDA_ERR_BAD_ARG;
  }
  static const bool no_fast = getenv("LLACUDA_CHOL_NO_DIAG_CHAIN") != nullptr;
  if (P == 1 && Q > 1 && !no_fast) return pdpotrf_rank_cols(R, Np, nb, A, lld);
  const int64_t mloc = Np / P, nloc = Np / Q, nblk = Np / nb, mb = mloc / nb, nbl = nloc / nb;
  PoolBuf wcb[2], wrb[2], xchg, dblk, tinv;
  for (int i = 0; i < 2; i++) {
    LLA_TRY(wcb[i].alloc((size_t)nloc * nb * 8));
    LLA_TRY(wrb[i].alloc((size_t)mloc * nb * 8));
  }
  LLA_TRY(xchg.alloc((size_t)mloc * nb * 8));
  LLA_TRY(dblk.alloc((size_t)nb * nb * 8));
  LLA_TRY(tinv.alloc(tinv_bytes(nb)));
  int* info_acc = R.dflag;
  int* info_last = R.dflag + 1;
  LLA_CUDA(cudaMemsetAsync(R.dflag, 0, 2 * sizeof(int), R.main));
  cudaEvent_t ev_panel[2] = {R.ev[0], R.ev[1]}, ev_row = R.ev[2], ev_start = R.ev[3];
  LLA_CUDA(cudaEventRecord(ev_start, R.main));
  LLA_CUDA(cudaStreamWaitEvent(R.side, ev_start, 0));
  cudaStream_t sd = R.side;

  auto panel = [&](int64_t k) -> int {
    const int slot = (int)(k & 1);
    const int pk = (int)(k % P), qk = (int)(k % Q);
    const int64_t lkr = k / P;
    double* wc = (double*)wcb[slot].p;
    double* wr = (double*)wrb[slot].p;
    double* D = (double*)dblk.p;
    const int64_t lj0 = blocks_le(k, q, Q);           // my first local block column right of the diagonal
    const int64_t ncols = nloc - lj0 * nb;
    if (p == pk) {
      if (q == qk) {
        double* diag = A + lkr * nb + (k / Q) * nb * lld;
        LLA_TRY(dpotrf_dev(true, nb, diag, lld, info_last, sd));
        lla_info_acc_kernel<<<1, 1, 0, sd>>>(info_acc, info_last, (int)(k * nb));
        count_launch();
        LLA_TRY(d2d_2d(D, nb, diag, lld, nb, nb, 8, sd));
      }
      if (Q > 1) LLA_NCCL(nc->Broadcast(D, D, (size_t)nb * nb, ncclDouble, qk, R.rowc, sd));
      if (ncols > 0) {
        double* row = A + lkr * nb + lj0 * nb * lld;
        LLA_TRY(dtrtri_blocks_dev(true, false, nb, D, nb, (double*)tinv.p, sd));
        LLA_TRY(dtrsm_left_inv_dev(true, 1, nb, ncols, D, nb, (const double*)tinv.p, row, lld, sd));
        LLA_TRY(d2d_2d(wc, nb, row, lld, nb, ncols, 8, sd));
      }
    }
    // U(k, J), J on my grid column: down the column communicator
    if (P > 1 && ncols > 0) LLA_NCCL(nc->Broadcast(wc, wc, (size_t)nb * ncols, ncclDouble, pk, R.colc, sd));
    // U(k, I), I on my grid row: each rank of the row contributes the blocks it holds (I mod Q == its q)
    const int64_t li0 = blocks_le(k, p, P);
    std::vector<long long> soff, doff;
    if (Q == 1) {
      for (int64_t li = li0; li < mb; li++) {
        const int64_t I = li * P + p;
        soff.push_back((long long)((I - lj0) * nb * nb));   // Q == 1: wc holds every J > k, block J at (J - lj0)
        doff.push_back((long long)((li - li0) * nb * nb));
      }
      if (!soff.empty()) LLA_TRY(copy_blocks<double>(wc, nb, wr, nb, (int)nb, (int)nb, soff, doff, sd));
    } else {
      double* X = (double*)xchg.p;
      std::vector<int64_t> seg_off((size_t)Q + 1, 0);
      std::vector<std::vector<int64_t>> lis((size_t)Q);
      for (int64_t li = li0; li < mb; li++) lis[(size_t)((li * P + p) % Q)].push_back(li);
      for (int qq = 0; qq < Q; qq++) seg_off[(size_t)qq + 1] = seg_off[(size_t)qq] + (int64_t)lis[(size_t)qq].size();
      // my own contribution: pack from wc
      for (int64_t li : lis[(size_t)q]) {
        const int64_t I = li * P + p;
        soff.push_back((long long)((I / Q - lj0) * nb * nb));
        doff.push_back((long long)((seg_off[(size_t)q] + (int64_t)doff.size()) * nb * nb));
      }
      if (!soff.empty()) LLA_TRY(copy_blocks<double>(wc, nb, X, nb, (int)nb, (int)nb, soff, doff, sd));
      LLA_NCCL(nc->GroupStart());
      for (int qq = 0; qq < Q; qq++) {
        const int64_t cnt = (int64_t)lis[(size_t)qq].size();
        if (cnt > 0) {
          double* seg = X + seg_off[(size_t)qq] * nb * nb;
          LLA_NCCL(nc->Broadcast(seg, seg, (size_t)cnt * nb * nb, ncclDouble, qq, R.rowc, sd));
        }
      }
      LLA_NCCL(nc->GroupEnd());
      soff.clear(); doff.clear();
      for (int qq = 0; qq < Q; qq++)
        for (size_t i = 0; i < lis[(size_t)qq].size(); i++) {
          soff.push_back((long long)((seg_off[(size_t)qq] + (int64_t)i) * nb * nb));
          doff.push_back((long long)((lis[(size_t)qq][i] - li0) * nb * nb));
        }
      if (!soff.empty()) LLA_TRY(copy_blocks<double>(X, nb, wr, nb, (int)nb, (int)nb, soff, doff, sd));
    }
    LLA_CUDA(cudaEventRecord(ev_panel[slot], sd));
    return 0;
  };

  // A(I, J) -= U(k, I)^T U(k, J) on my blocks with local block row in [lo, hi)
  auto update = [&](int64_t k, int64_t lo, int64_t hi) -> int {
    const int slot = (int)(k & 1);
    const double* wc = (const double*)wcb[slot].p;
    const double* wr = (const double*)wrb[slot].p;
    const int64_t lj0 = blocks_le(k, q, Q), li0 = blocks_le(k, p, P);
    for (int64_t lj = lj0; lj < nbl; lj++) {
      const int64_t J = lj * Q + q;
      const int64_t top = blocks_le(J, p, P);          // local block rows with I <= J
      const int64_t r0 = lo > li0 ? lo : li0, r1 = hi < top ? hi : top;
      if (r1 <= r0) continue;
      LLA_TRY(dgemm_dev(1, 0, (r1 - r0) * nb, nb, nb, -1.0, wr + (r0 - li0) * nb * nb, nb, wc + (lj - lj0) * nb * nb, nb, 1.0,
                        A + r0 * nb + lj * nb * lld, lld, TRI_FULL, R.main));
    }
    return 0;
  };

  LLA_TRY(panel(0));
  for (int64_t k = 0; k < nblk; k++) {
    LLA_CUDA(cudaStreamWaitEvent(R.main, ev_panel[k & 1], 0));
    const int64_t li0 = blocks_le(k, p, P);
    const bool own_next = (k + 1 < nblk) && ((k + 1) % P == p);
    const int64_t split = own_next ? li0 + 1 : li0;
    if (own_next) LLA_TRY(update(k, li0, split));       // block row k+1 first: all that panel k+1 needs
    if (k + 1 < nblk) {
      LLA_CUDA(cudaEventRecord(ev_row, R.main));
      LLA_CUDA(cudaStreamWaitEvent(sd, ev_row, 0));
      LLA_TRY(panel(k + 1));
    }
    LLA_TRY(update(k, split, mb));
  }
  LLA_CUDA(cudaEventRecord(ev_start, sd));
  LLA_CUDA(cudaStreamWaitEvent(R.main, ev_start, 0));
  LLA_CUDA(cudaStreamSynchronize(R.main));   // the pooled buffers are released with this frame
  return 0;
}

// ------------------------------------------------------------------ LU with partial pivoting
// Pivoting needs a whole block column in one place: the panel of step k is assembled on the owner of the diagonal
// block (for P > 1 the other ranks of that grid column send their rows: "gather the whole panel onto one GPU",
// SURVEY.md section 8e), factored there by the single-GPU panel code -- the same kernels on the same numbers in
// the same order, which is what keeps ipiv bit-exact for every grid -- and broadcast, pivots included, to all
// ranks.  Every rank then applies the interchanges to its own columns, solves U(k, J) = L_kk^-1 A(k, J) on the grid
// row that owns block row k, receives U(k, J) down its grid column and updates with one DMMA GEMM.
//
// P == 1 (block columns, the default for LU on an NVSwitch box): every row of a column is local, so interchanges,
// U(k, J) and the update need no further traffic.  P > 1: rows are spread over the grid column; the interchanges
// become an exchange of the touched rows inside each grid column (compose the nb swaps into one permutation on the
// device, pack the source rows, all-gather along the column communicator, unpack the destination rows).
namespace {

constexpr int LSW_MAX = 1024;   // nb <= 512: at most 2 nb rows are touched by one panel's interchanges

// One CTA composes the nb interchanges i <-> piv[i] (piv 1-based, relative to row `base`) into a permutation over
// the touched rows: lists dst[e], src[e] (global rows, content of dst after = content of src before), count in *cnt.
__global__ void __launch_bounds__(512) lla_swap_compose_kernel(const int* __restrict__ piv, int nb, int base,
                                                               int* __restrict__ dst, int* __restrict__ src, int* __restrict__ cnt) {
  __shared__ int row[LSW_MAX];   // slot -> global row
  __shared__ int cur[LSW_MAX];   // slot -> row whose original content sits there now
  __shared__ int nslots, hit;
  const int t = threadIdx.x;
  for (int i = t; i < nb; i += blockDim.x) { row[i] = base + i; cur[i] = base + i; }
  if (t == 0) nslots = nb;
  __syncthreads();
  for (int i = 0; i < nb; i++) {
    const int pr = base + piv[i] - 1;
    if (pr == base + i) continue;            // uniform: piv is the same for every thread
    int sp;
    if (pr < base + nb) sp = pr - base;
    else {
      if (t == 0) hit = -1;
      __syncthreads();
      for (int s = nb + t; s < nslots; s += blockDim.x)
        if (row[s] == pr) hit = s;
      __syncthreads();
      if (hit < 0) {
        if (t == 0) { row[nslots] = pr; cur[nslots] = pr; hit = nslots; nslots = nslots + 1; }
        __syncthreads();
      }
      sp = hit;
    }
    if (t == 0) { const int a = cur[i], b = cur[sp]; cur[i] = b; cur[sp] = a; }
    __syncthreads();
  }
  for (int s = t; s < nslots; s += blockDim.x) { dst[s] = row[s]; src[s] = cur[s]; }
  if (t == 0) *cnt = nslots;
}

// local position of global row g on process row (g / nb) mod P
__device__ __forceinline__ int64_t local_row(int g, int nb, int P) { return (int64_t)((g / nb) / P) * nb + g % nb; }

// buf[e, c] = A_loc[local(src[e]), c] for the entries whose source row I own (others left alone)
__global__ void __launch_bounds__(256) lla_swap_pack_kernel(const double* __restrict__ A, int64_t lld, int64_t ncols,
                                                            const int* __restrict__ dst, const int* __restrict__ src,
                                                            const int* __restrict__ cnt, int nb, int P, int p,
                                                            double* __restrict__ buf, int ldbuf) {
  const int n = *cnt;
  const int64_t c = (int64_t)blockIdx.x * 4 + threadIdx.x / 64;
  if (c >= ncols) return;
  for (int e = threadIdx.x % 64; e < n; e += 64) {
    const int s = src[e];
    if (s == dst[e] || (s / nb) % P != p) continue;
    buf[e + c * ldbuf] = A[local_row(s, nb, P) + c * lld];
  }
}
// A_loc[local(dst[e]), c] = gathered[owner(src[e])][e, c] for the entries whose destination row I own
__global__ void __launch_bounds__(256) lla_swap_unpack_kernel(double* __restrict__ A, int64_t lld, int64_t ncols,
                                                              const int* __restrict__ dst, const int* __restrict__ src,
                                                              const int* __restrict__ cnt, int nb, int P, int p,
                                                              const double* __restrict__ gathered, int ldbuf, int64_t part) {
  const int n = *cnt;
  const int64_t c = (int64_t)blockIdx.x * 4 + threadIdx.x / 64;
  if (c >= ncols) return;
  for (int e = threadIdx.x % 64; e < n; e += 64) {
    const int d = dst[e], s = src[e];
    if (s == d || (d / nb) % P != p) continue;
    const int owner = (s / nb) % P;
    A[local_row(d, nb, P) + c * lld] = gathered[owner * part + e + c * ldbuf];
  }
}

}  // namespace

int pdgetrf_rank(Rank& R, int64_t Np, int64_t nb, double* A, int64_t lld, int* ipiv, double* B, int64_t ldb, int64_t nrhs,
                 bool* rode) {
  const NcclApi* nc = nccl();
  const int P = R.P, Q = R.Q, p = R.p, q = R.q;
  if (Np % (nb * lcm(P, Q)) || nb % IB || 2 * nb > LSW_MAX) {
    set_error(__FILE__, __LINE__, "pdgetrf: the padded order must be a multiple of nb * lcm(P, Q), nb of 128, nb <= 512", (int)nb);
    return LLACUDA_ERR_BAD_ARG;
  }
  const int64_t mloc = Np / P, nloc = Np / Q, nblk = Np / nb, mb = mloc / nb, nbl = nloc / nb;
  PoolBuf pan[2], pivb[2], linv[2], lrow[2], urow[2], gath, swp, swl;
  for (int i = 0; i < 2; i++) {
    LLA_TRY(pan[i].alloc((size_t)Np * nb * 8));          // the factored panel in global row order, ld = m_k
    LLA_TRY(pivb[i].alloc((size_t)nb * sizeof(int)));
    LLA_TRY(linv[i].alloc(tinv_bytes(nb)));
    if (P > 1) {
      LLA_TRY(lrow[i].alloc((size_t)mloc * nb * 8));     // my rows of L(:, k) in local row order
      LLA_TRY(urow[i].alloc((size_t)nloc * nb * 8));     // U(k, J) for my columns
    }
  }
  if (P > 1) {
    LLA_TRY(gath.alloc((size_t)mloc * nb * 8));
    LLA_TRY(swp.alloc((size_t)(P + 1) * 2 * nb * nloc * 8));   // pack buffer + the all-gathered copies
    LLA_TRY(swl.alloc((size_t)(4 * LSW_MAX + 8) * sizeof(int)));   // two slots of (dst, src) lists + the two counts
  }
  int* info_acc = R.dflag;
  int* info_last = R.dflag + 1;
  LLA_CUDA(cudaMemsetAsync(R.dflag, 0, 2 * sizeof(int), R.main));
  cudaEvent_t ev_panel[2] = {R.ev[0], R.ev[1]}, ev_col = R.ev[2], ev_start = R.ev[3];
  LLA_CUDA(cudaEventRecord(ev_start, R.main));
  LLA_CUDA(cudaStreamWaitEvent(R.side, ev_start, 0));
  cudaStream_t sd = R.side;
  // right-hand sides riding along (block columns only: every rank sees whole rows of the broadcast panel)
  const bool rider = B != nullptr && nrhs > 0 && P == 1 && getenv("LLACUDA_LU_NO_RHS_RIDER") == nullptr;
  cudaEvent_t ev_rider[2] = {R.ev[4], R.ev[5]};
  cudaStream_t sB = R.chain;
  if (rider) LLA_CUDA(cudaStreamWaitEvent(sB, ev_start, 0));
  if (rode) *rode = rider;

  auto panel = [&](int64_t k) -> int {
    const int slot = (int)(k & 1);
    const int pk = (int)(k % P), qk = (int)(k % Q);
    const int64_t mk = Np - k * nb;
    double* PN = (double*)pan[slot].p;
    int* piv = (int*)pivb[slot].p;
    const int root = pk * Q + qk;
    if (rider && k >= 2) LLA_CUDA(cudaStreamWaitEvent(sd, ev_rider[slot], 0));   // the rider of panel k - 2 read this slot
    if (q == qk) {
      const int64_t lkc = k / Q;
      double* col = A + lkc * nb * lld;                  // my rows of block column k
      if (P == 1) {
        double* mine = col + k * nb;
        LLA_TRY(dgetrf_dev(mk, nb, mine, lld, piv, info_last, sd));
        LLA_TRY(d2d_2d(PN, mk, mine, lld, mk, nb, 8, sd));
      } else {
        // gather the panel on the diagonal owner: my block rows I >= k, packed in local order
        const int64_t li0 = blocks_lt(k, p, P);          // first local block row with I >= k
        const int64_t myrows = (mb - li0) * nb;
        double* G = (double*)gath.p;
        if (p != pk) {
          if (myrows > 0) {
            LLA_TRY(d2d_2d(G, myrows, col + li0 * nb, lld, myrows, nb, 8, sd));
            LLA_NCCL(nc->Send(G, (size_t)myrows * nb, ncclDouble, pk, R.colc, sd));
          }
        } else {
          std::vector<long long> soff, doff;
          for (int64_t li = li0; li < mb; li++) {
            soff.push_back((long long)(li * nb));
            doff.push_back((long long)((li * P + p - k) * nb));
          }
          LLA_TRY(copy_blocks<double>(col, lld, PN, mk, (int)nb, (int)nb, soff, doff, sd));
          for (int pp = 0; pp < P; pp++) {
            if (pp == pk) continue;
            const int64_t l0 = blocks_lt(k, pp, P), rows = (mb - l0) * nb;
            if (rows <= 0) continue;
            LLA_NCCL(nc->Recv(G, (size_t)rows * nb, ncclDouble, pp, R.colc, sd));
            soff.clear(); doff.clear();
            for (int64_t li = l0; li < mb; li++) {
              soff.push_back((long long)((li - l0) * nb));
              doff.push_back((long long)((li * P + pp - k) * nb));
            }
            LLA_TRY(copy_blocks<double>(G, rows, PN, mk, (int)nb, (int)nb, soff, doff, sd));
          }
          LLA_TRY(dgetrf_dev(mk, nb, PN, mk, piv, info_last, sd));
        }
      }
      if (R.rank == root) {
        lla_info_acc_kernel<<<1, 1, 0, sd>>>(info_acc, info_last, (int)(k * nb));
        count_launch();
      }
    }
    if (R.nranks > 1) {
      LLA_NCCL(nc->GroupStart());
      LLA_NCCL(nc->Broadcast(PN, PN, (size_t)mk * nb, ncclDouble, root, R.world, sd));
      LLA_NCCL(nc->Broadcast(piv, piv, (size_t)nb, ncclInt32, root, R.world, sd));
      LLA_NCCL(nc->GroupEnd());
    }
    lla_ipiv_shift_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, sd>>>(ipiv, piv, (int)(k * nb), (int)nb);
    count_launch();
    LLA_TRY(dtrtri_blocks_dev(false, true, nb, PN, mk, (double*)linv[slot].p, sd));
    if (P > 1) {
      // my rows of the factored panel: back into my block column (grid column qk), and as L(I, k) in local row order
      const int64_t li0 = blocks_lt(k, p, P);
      std::vector<long long> soff, doff;
      for (int64_t li = li0; li < mb; li++) {
        soff.push_back((long long)((li * P + p - k) * nb));
        doff.push_back((long long)(li * nb));
      }
      if (q == qk && !soff.empty()) LLA_TRY(copy_blocks<double>(PN, mk, A + (k / Q) * nb * lld, lld, (int)nb, (int)nb, soff, doff, sd));
      const int64_t la0 = blocks_le(k, p, P);            // block rows strictly below the diagonal block
      soff.clear(); doff.clear();
      for (int64_t li = la0; li < mb; li++) {
        soff.push_back((long long)((li * P + p - k) * nb));
        doff.push_back((long long)((li - la0) * nb));
      }
      if (!soff.empty()) LLA_TRY(copy_blocks<double>(PN, mk, (double*)lrow[slot].p, mloc, (int)nb, (int)nb, soff, doff, sd));
      int* lists = (int*)swl.p + slot * 2 * LSW_MAX;
      lla_swap_compose_kernel<<<1, 512, 0, sd>>>(piv, (int)nb, (int)(k * nb), lists, lists + LSW_MAX, (int*)swl.p + 4 * LSW_MAX + slot);
      count_launch();
    }
    LLA_CUDA(cudaGetLastError());
    LLA_CUDA(cudaEventRecord(ev_panel[slot], sd));
    if (rider) {
      // interchanges of this panel, solve with its unit-lower triangle, product with the rows below -- on my copy of B
      double* Bk = B + k * nb;
      LLA_CUDA(cudaStreamWaitEvent(sB, ev_panel[slot], 0));
      LLA_TRY(dlaswp_dev(nrhs, Bk, ldb, 0, nb, piv, true, sB));
      LLA_TRY(dtrsm_left_inv_dev(false, 0, nb, nrhs, PN, mk, (const double*)linv[slot].p, Bk, ldb, sB));
      if (mk > nb) LLA_TRY(dgemm_dev(0, 0, mk - nb, nrhs, nb, -1.0, PN + nb, mk, Bk, ldb, 1.0, Bk + nb, ldb, TRI_FULL, sB));
      LLA_CUDA(cudaEventRecord(ev_rider[slot], sB));
    }
    return 0;
  };

  // interchanges of panel k on my local columns [c0, c1) (element columns)
  auto swap_cols = [&](int64_t k, int64_t c0, int64_t c1, cudaStream_t st) -> int {
    if (c1 <= c0) return 0;
    const int slot = (int)(k & 1);
    if (P == 1) return dlaswp_dev(c1 - c0, A + k * nb + c0 * lld, lld, 0, nb, (const int*)pivb[slot].p, true, st);
    const int64_t ncols = c1 - c0;
    const int ldbuf = (int)(2 * nb);
    double* mine = (double*)swp.p;
    double* all = mine + (size_t)ldbuf * nloc;
    const int* dl = (const int*)swl.p + slot * 2 * LSW_MAX;
    const int* sl = dl + LSW_MAX;
    const int* cn = (const int*)swl.p + 4 * LSW_MAX + slot;
    const int64_t part = (int64_t)ldbuf * ncols;
    lla_swap_pack_kernel<<<(unsigned)((ncols + 3) / 4), 256, 0, st>>>(A + c0 * lld, lld, ncols, dl, sl, cn, (int)nb, P, p, mine, ldbuf);
    count_launch();
    LLA_NCCL(nc->AllGather(mine, all, (size_t)part, ncclDouble, R.colc, st));
    lla_swap_unpack_kernel<<<(unsigned)((ncols + 3) / 4), 256, 0, st>>>(A + c0 * lld, lld, ncols, dl, sl, cn, (int)nb, P, p, all, ldbuf, part);
    count_launch();
    LLA_CUDA(cudaGetLastError());
    return 0;
  };

  // U(k, J) and the trailing update on my local block columns [lj_lo, lj_hi), all right of panel k
  auto update = [&](int64_t k, int64_t lj_lo, int64_t lj_hi, cudaStream_t st) -> int {
    if (lj_hi <= lj_lo) return 0;
    const int slot = (int)(k & 1);
    const int pk = (int)(k % P);
    const int64_t mk = Np - k * nb, ncols = (lj_hi - lj_lo) * nb;
    const double* PN = (const double*)pan[slot].p;
    if (P == 1) {
      double* cols = A + k * nb + lj_lo * nb * lld;
      LLA_TRY(dlaswp_dev(ncols, cols, lld, 0, nb, (const int*)pivb[slot].p, true, st));
      LLA_TRY(dtrsm_left_inv_dev(false, 0, nb, ncols, PN, mk, (const double*)linv[slot].p, cols, lld, st));
      if (mk > nb) LLA_TRY(dgemm_dev(0, 0, mk - nb, ncols, nb, -1.0, PN + nb, mk, cols, lld, 1.0, cols + nb, lld, TRI_FULL, st));
      return 0;
    }
    LLA_TRY(swap_cols(k, lj_lo * nb, lj_hi * nb, st));
    double* U = (double*)urow[slot].p + lj_lo * nb * nb;   // nb x ncols, ld = nb
    if (p == pk) {
      double* row = A + (k / P) * nb + lj_lo * nb * lld;
      LLA_TRY(dtrsm_left_inv_dev(false, 0, nb, ncols, PN, mk, (const double*)linv[slot].p, row, lld, st));
      LLA_TRY(d2d_2d(U, nb, row, lld, nb, ncols, 8, st));
    }
    LLA_NCCL(nc->Broadcast(U, U, (size_t)nb * ncols, ncclDouble, pk, R.colc, st));
    const int64_t la0 = blocks_le(k, p, P);
    const int64_t rows = (mb - la0) * nb;
    if (rows > 0)
      LLA_TRY(dgemm_dev(0, 0, rows, ncols, nb, -1.0, (const double*)lrow[slot].p, mloc, U, nb, 1.0,
                        A + la0 * nb + lj_lo * nb * lld, lld, TRI_FULL, st));
    return 0;
  };

  // With P > 1 every collective of the step (row exchange, U broadcast) is issued on the stream the step's work
  // runs on, in the same order on every rank of the communicator; panels use `side`, so the look-ahead column goes
  // there too and the rest of the update stays on `main` -- two streams, but each communicator call sequence is the
  // same on all ranks because the program is SPMD and both streams are fed by one host thread in program order.
  LLA_TRY(panel(0));
  for (int64_t k = 0; k < nblk; k++) {
    LLA_CUDA(cudaStreamWaitEvent(R.main, ev_panel[k & 1], 0));
    const int64_t lj0 = blocks_le(k, q, Q);               // my first block column right of panel k
    const bool own_next = (k + 1 < nblk) && ((k + 1) % Q == q);
    const int64_t split = own_next ? lj0 + 1 : lj0;
    if (own_next) LLA_TRY(update(k, lj0, split, R.main)); // block column k+1 first
    if (k + 1 < nblk) {
      LLA_CUDA(cudaEventRecord(ev_col, R.main));
      LLA_CUDA(cudaStreamWaitEvent(sd, ev_col, 0));
    }
    if (P == 1) {
      if (k + 1 < nblk) LLA_TRY(panel(k + 1));
      LLA_TRY(update(k, split, nbl, R.main));
    } else {
      // one communicator, one issue order: the rest of update k (its exchanges on the column communicator) is queued
      // before panel k+1's collectives; the kernels still overlap, panel k+1 only waits for its own column
      LLA_TRY(update(k, split, nbl, R.main));
      if (k + 1 < nblk) LLA_TRY(panel(k + 1));
    }
    // interchanges of my block columns left of the panel: nothing depends on them before the end
    const int64_t nleft = blocks_lt(k, q, Q);
    if (nleft > 0) LLA_TRY(swap_cols(k, 0, nleft * nb, R.main));
  }
  LLA_CUDA(cudaEventRecord(ev_start, sd));
  LLA_CUDA(cudaStreamWaitEvent(R.main, ev_start, 0));
  if (rider) {
    LLA_CUDA(cudaEventRecord(ev_start, sB));
    LLA_CUDA(cudaStreamWaitEvent(R.main, ev_start, 0));
  }
  LLA_CUDA(cudaStreamSynchronize(R.main));
  return 0;
}

// ------------------------------------------------------------------ triangular solves on replicated right-hand sides
// op(T) X = B, T block-cyclic triangular (the factor as the factorisations leave it), B: Np x nrhs on every rank.
// Few right-hand sides do not shard: per block k the owners of the known part fold it into a
// partial sum (one DMMA GEMM on their blocks), an all-reduce adds the partial sums, the owner of the diagonal block
// solves and broadcasts x_k.
int pdtrsv_rank(Rank& R, int64_t Np, int64_t nb, const double* T, int64_t lld, bool upper, bool trans, bool unit,
                double* B, int64_t ldb, int64_t nrhs) {
  const NcclApi* nc = nccl();
  const int P = R.P, Q = R.Q, p = R.p, q = R.q;
  const int64_t mloc = Np / P, nloc = Np / Q, nblk = Np / nb, mb = mloc / nb, nbl = nloc / nb;
  const bool forward = (upper == trans);   // effective lower triangular
  cudaStream_t st = R.main;
  PoolBuf sbuf, xbuf, gbuf, tinv;
  const int64_t gl = mloc > nloc ? mloc : nloc;
  LLA_TRY(sbuf.alloc((size_t)nb * nrhs * 8));
  LLA_TRY(xbuf.alloc((size_t)nb * nrhs * 8));
  LLA_TRY(gbuf.alloc((size_t)gl * nrhs * 8));
  LLA_TRY(tinv.alloc(tinv_bytes(nb)));
  double* S = (double*)sbuf.p;
  double* X = (double*)xbuf.p;
  double* G = (double*)gbuf.p;
  for (int64_t step = 0; step < nblk; step++) {
    const int64_t k = forward ? step : nblk - 1 - step;
    const int pk = (int)(k % P), qk = (int)(k % Q);
    LLA_CUDA(cudaMemsetAsync(S, 0, (size_t)nb * nrhs * 8, st));
    // the blocks of op(T)'s block row k that multiply known parts of x: a block row of T (no transpose) spread over
    // the grid row pk, or a block column of T (transpose) spread over the grid column qk
    std::vector<long long> soff, doff;
    if (!trans) {
      if (p == pk) {
        const int64_t lo = forward ? 0 : blocks_le(k, q, Q), hi = forward ? blocks_lt(k, q, Q) : nbl;
        for (int64_t lj = lo; lj < hi; lj++) { soff.push_back((long long)((lj * Q + q) * nb)); doff.push_back((long long)((lj - lo) * nb)); }
        if (hi > lo) {
          const int64_t kk = (hi - lo) * nb;
          LLA_TRY(copy_blocks<double>(B, ldb, G, gl, (int)nb, (int)nrhs, soff, doff, st));
          LLA_TRY(dgemm_dev(0, 0, nb, nrhs, kk, 1.0, T + (k / P) * nb + lo * nb * lld, lld, G, gl, 0.0, S, nb, TRI_FULL, st));
        }
      }
    } else {
      if (q == qk) {
        const int64_t lo = forward ? 0 : blocks_le(k, p, P), hi = forward ? blocks_lt(k, p, P) : mb;
        for (int64_t li = lo; li < hi; li++) { soff.push_back((long long)((li * P + p) * nb)); doff.push_back((long long)((li - lo) * nb)); }
        if (hi > lo) {
          const int64_t kk = (hi - lo) * nb;
          LLA_TRY(copy_blocks<double>(B, ldb, G, gl, (int)nb, (int)nrhs, soff, doff, st));
          LLA_TRY(dgemm_dev(1, 0, nb, nrhs, kk, 1.0, T + lo * nb + (k / Q) * nb * lld, lld, G, gl, 0.0, S, nb, TRI_FULL, st));
        }
      }
    }
    if (R.nranks > 1) LLA_NCCL(nc->AllReduce(S, S, (size_t)nb * nrhs, ncclDouble, ncclSum, R.world, st));
    dim3 grid((unsigned)((nb + 127) / 128), (unsigned)nrhs);
    lla_sub_block_kernel<<<grid, 128, 0, st>>>(X, B + k * nb, ldb, S, (int)nb, (int)nrhs);
    count_launch();
    if (p == pk && q == qk) {
      const double* Tkk = T + (k / P) * nb + (k / Q) * nb * lld;
      LLA_TRY(dtrtri_blocks_dev(upper, unit, nb, Tkk, lld, (double*)tinv.p, st));
      LLA_TRY(dtrsm_left_inv_dev(upper, trans ? 1 : 0, nb, nrhs, Tkk, lld, (const double*)tinv.p, X, nb, st));
    }
    if (R.nranks > 1) LLA_NCCL(nc->Broadcast(X, X, (size_t)nb * nrhs, ncclDouble, pk * Q + qk, R.world, st));
    LLA_TRY(d2d_2d(B + k * nb, ldb, X, nb, nb, nrhs, 8, st));
  }
  LLA_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int pdlaswp_rhs_rank(Rank& R, int64_t n, const int* ipiv, double* B, int64_t ldb, int64_t nrhs) {
  return dlaswp_dev(nrhs, B, ldb, 0, n, ipiv, true, R.main);
}

}  // namespace llacuda
