// lla_b200/csrc/dist.h -- internal interface of the multi-GPU engine (dist.cu, dist_algos.cu).
//
// BASELINE.json north_star: "Large GEMM and blocked LU/Cholesky are partitioned across one 8xB200 box as a 2D
// block-cyclic layout, broadcasting panels with NCCL over NVLink ... The Lisp API stays unchanged".  LLA issues one
// synchronous foreign-funcall from one process (reference src/fortran-call.lisp:428-439), so the engine's primary
// mode is SINGLE PROCESS: ncclCommInitAll over the visible devices, one host thread per device running the same
// SPMD rank function.  The same rank functions run one-rank-per-process under a launcher
// (ncclCommInitRank with a caller-distributed unique id): that is how bench.py --gpus N and the tests drive them.
#pragma once
#include "internal.h"

#include <functional>
#include <vector>
#include <nccl.h>   // types only: the functions are resolved with dlopen (libnccl.so.2), so that single-GPU users never load NCCL

namespace llacuda {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*);
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
  ncclResult_t (*GetVersion)(int*);
};
const NcclApi* nccl();   // nullptr (error recorded) when libnccl.so.2 cannot be loaded

int record_nccl_error(ncclResult_t r, const char* file, int line);
#define LLA_NCCL(call)                                                                  \
  do {                                                                                  \
    ncclResult_t _r = (call);                                                           \
    if (_r != ncclSuccess) return ::llacuda::record_nccl_error(_r, __FILE__, __LINE__); \
  } while (0)

// One rank = one GPU of the P x Q process grid; rank = p * Q + q.
struct Rank {
  int rank = 0, nranks = 1;
  int P = 1, Q = 1, p = 0, q = 0;
  int dev = 0;
  ncclComm_t world = nullptr;   // all ranks
  ncclComm_t rowc = nullptr;    // the Q ranks of my grid row (rank inside = q)
  ncclComm_t colc = nullptr;    // the P ranks of my grid column (rank inside = p)
  cudaStream_t main = nullptr;  // trailing updates
  cudaStream_t side = nullptr;  // panels + every collective (one issue order per communicator)
  cudaStream_t chain = nullptr; // the diagonal chain of the block-column Cholesky (its broadcasts use `world`, the row exchange `rowc`)
  cudaEvent_t ev[24] = {};
  int* dflag = nullptr;         // 8 device ints: [0] INFO accumulator, [1] INFO of the last local factorisation
  int* hflag = nullptr;         // pinned mirror
  char err[320] = "";
  int err_code = 0;
  struct GridComms { int P, Q; ncclComm_t rowc, colc; };
  std::vector<GridComms> grids;   // sub-communicators of every grid shape used so far (switching back is free)
};

struct Engine {
  bool ready = false;
  bool single_process = false;
  int nranks = 0;
  int P = 1, Q = 1;
};
const Engine& engine();
// Runs fn on every local rank (single process: one worker thread per device, in parallel; launcher mode: the one
// local rank, inline) and returns the first non-zero code; messages go to llacuda_last_error().
int run_ranks(const std::function<int(Rank&)>& fn);
int regrid(int P, int Q);   // re-split the row / column communicators for another grid shape (P * Q == nranks)

// ---- block-cyclic arithmetic: block index I lives on process I mod R, as local block I / R
inline int64_t blocks_le(int64_t k, int r, int R) { return k >= r ? (k - r) / R + 1 : 0; }   // #{I <= k : I mod R == r}
inline int64_t blocks_lt(int64_t k, int r, int R) { return k > 0 ? blocks_le(k - 1, r, R) : 0; }
inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
int64_t lcm(int64_t a, int64_t b);

// ---- rank functions (dist_algos.cu).  Local storage of an Mp x Np matrix (Mp, Np multiples of nb * P, nb * Q):
// column-major Mp/P x Np/Q with leading dimension lld, local block (li, lj) = global block (li * P + p, lj * Q + q).
// Everything is queued on R.main / R.side; the functions return with R.main waiting for all of it.
template <class T>
int summa_rank(Rank& R, int64_t Mp, int64_t Np, int64_t Kp, int64_t nb, T alpha, const T* A, int64_t lda,
               const T* B, int64_t ldb, T beta, T* C, int64_t ldc);
// A = U^T U, upper triangle of the block-cyclic square matrix (blocks below the diagonal are scratch);
// INFO (LAPACK's, global index) lands in R.dflag[0] on every rank
int pdpotrf_rank(Rank& R, int64_t Np, int64_t nb, double* A, int64_t lld);
// P A = L U with partial pivoting; ipiv: device int32[Np], LAPACK's global 1-based swap sequence, on every rank
// B (Np x nrhs, ldb, replicated on every rank), optional: on a 1 x Q grid the right-hand sides ride along with the
// factorisation (every rank applies each broadcast panel to its copy on a side stream), so B holds L^-1 P B on return
// and *rode is set; otherwise B is not touched.
int pdgetrf_rank(Rank& R, int64_t Np, int64_t nb, double* A, int64_t lld, int* ipiv, double* B = nullptr, int64_t ldb = 0,
                 int64_t nrhs = 0, bool* rode = nullptr);
// op(T) X = B for a block-cyclic triangular factor and B (Np x nrhs, ldb) replicated on every rank
int pdtrsv_rank(Rank& R, int64_t Np, int64_t nb, const double* T, int64_t lld, bool upper, bool trans, bool unit,
                double* B, int64_t ldb, int64_t nrhs);
// rows i <-> ipiv[i]-1 of the replicated right-hand sides (forward order)
int pdlaswp_rhs_rank(Rank& R, int64_t n, const int* ipiv, double* B, int64_t ldb, int64_t nrhs);
int finish_info(Rank& R, int* info_out);   // min positive INFO over the ranks -> *info_out (host), synchronises

}  // namespace llacuda
