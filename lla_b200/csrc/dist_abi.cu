// lla_b200/csrc/dist_abi.cu -- C entry points of the multi-GPU engine (include/llacuda.h, group 4).
//
//  * `llacuda_pdgemm / pdpotrf / pdgetrf / pdgesv / pdpotrs / pzgemm / pdgesv_batched`: HOST pointers, Fortran-style
//    argument lists identical to the single-GPU drop-ins, callable from ONE process (what LLA's foreign-funcall can
//    reach: reference src/fortran-call.lisp:428-439).  The engine scatters the operands block-cyclically over the
//    visible GPUs (each rank thread uploads its own blocks over its own PCIe link), runs the SPMD rank functions of
//    dist_algos.cu and gathers the results in place.
//  * `llacuda_*_local`: the same rank functions on block-cyclic operands that already live in HBM, one rank per
//    process (bench.py --gpus N, tests).
//  * with LLACUDA_GPUS=N in the environment the plain Fortran symbols (dgemm_, dgesv_, dgetrf_, dpotrf_, zgemm_)
//    route large problems here, so that the Lisp side does not change at all.
#include "dist.h"
#include "staging.h"
#include "../../include/llacuda.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace llacuda {

static inline int64_t max1(int64_t x) { return x > 1 ? x : 1; }

// A(i, i) = 1 for the padding rows / columns i in [n, Np) that this rank owns (keeps the padded matrix non-singular
// and positive definite without touching pivots: a zero never beats a candidate, and row i is its own first maximum)
__global__ void lla_pad_identity_kernel(double* A, int64_t lld, int64_t n, int64_t Np, int nb, int P, int Q, int p, int q) {
  const int64_t i = n + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Np) return;
  const int64_t I = i / nb;
  if (I % P != p || I % Q != q) return;
  A[(I / P) * nb + i % nb + ((I / Q) * nb + i % nb) * lld] = 1.0;
}

// Block-cyclic local part of an m x n host matrix (column-major, ldh), element size es: rows in blocks of nb over a
// grid dimension of size Pr (my index pr), columns over Pc (pc).  `tri` = 1: only blocks I <= J (upper triangle).
struct Scatter {
  int64_t m, n, nb;
  int Pr, pr, Pc, pc;
  size_t es;
  int tri;
};

template <bool UP>
static int scatter_copy(const Scatter& S, void* dev, int64_t lld, void* host, int64_t ldh, cudaStream_t st) {
  const int64_t rb = (S.m + S.nb - 1) / S.nb, cb = (S.n + S.nb - 1) / S.nb;
  for (int64_t J = S.pc; J < cb; J += S.Pc) {
    const int64_t c0 = J * S.nb, wc = (S.n - c0 < S.nb) ? S.n - c0 : S.nb;
    char* dcol = (char*)dev + (size_t)((J / S.Pc) * S.nb) * lld * S.es;
    char* hcol = (char*)host + (size_t)c0 * ldh * S.es;
    if (S.Pr == 1) {   // all rows local: one copy per block column (down to the diagonal block for a triangle)
      int64_t rows = S.m;
      if (S.tri) { rows = (J + 1) * S.nb; if (rows > S.m) rows = S.m; }
      if (UP) LLA_TRY(stage_region_h2d(dcol, (size_t)lld * S.es, hcol, (size_t)ldh * S.es, (size_t)rows * S.es, (size_t)wc, st));
      else LLA_TRY(stage_region_d2h(hcol, (size_t)ldh * S.es, dcol, (size_t)lld * S.es, (size_t)rows * S.es, (size_t)wc, st));
      continue;
    }
    for (int64_t I = S.pr; I < rb; I += S.Pr) {
      if (S.tri && I > J) break;
      const int64_t r0 = I * S.nb, hr = (S.m - r0 < S.nb) ? S.m - r0 : S.nb;
      char* d = dcol + (size_t)((I / S.Pr) * S.nb) * S.es;
      char* h = hcol + (size_t)r0 * S.es;
      if (UP) LLA_TRY(stage_region_h2d(d, (size_t)lld * S.es, h, (size_t)ldh * S.es, (size_t)hr * S.es, (size_t)wc, st));
      else LLA_TRY(stage_region_d2h(h, (size_t)ldh * S.es, d, (size_t)lld * S.es, (size_t)hr * S.es, (size_t)wc, st));
    }
  }
  return 0;
}

static void grid_from_env(const char* name, int nranks, int defP, int* P, int* Q) {
  *P = defP; *Q = nranks / defP;
  if (const char* e = getenv(name)) {
    int a = 0, b = 0;
    if (sscanf(e, "%dx%d", &a, &b) == 2 && a > 0 && b > 0 && a * b == nranks) { *P = a; *Q = b; }
  }
}
static int most_square(int w) { int p = 1; for (int c = 1; c * c <= w; c++) if (w % c == 0) p = c; return p; }

static int64_t env_nb(const char* name, int64_t def) {
  if (const char* e = getenv(name)) { long v = atol(e); if (v >= 128 && v % 128 == 0) return v; }
  return def;
}

// ------------------------------------------------------------------ GEMM
// C = alpha op(A) op(B) + beta C, double.  op = 'T' / 'C' operands are scattered on the transposed grid and
// transposed on the device, so the SUMMA itself is always N, N.
static int pdgemm_host(int oa, int ob, int64_t m, int64_t n, int64_t k, double alpha, const double* a, int64_t lda,
                       const double* b, int64_t ldb, double beta, double* c, int64_t ldc) {
  const Engine& E = engine();
  int P, Q;
  grid_from_env("LLACUDA_GEMM_GRID", E.nranks, most_square(E.nranks), &P, &Q);
  LLA_TRY(regrid(P, Q));
  const int64_t mn = m < n ? (m < k ? m : k) : (n < k ? n : k);
  const int64_t nb = env_nb("LLACUDA_GEMM_NB", mn >= 16384 ? 2048 : (mn >= 8192 ? 1024 : 512));
  const int64_t Mp = round_up(m, nb * P), Np = round_up(n, nb * Q), Kp = round_up(k, nb * lcm(P, Q));
  return run_ranks([&](Rank& R) -> int {
    const int64_t mloc = Mp / P, nloc = Np / Q, kA = Kp / Q, kB = Kp / P;
    PoolBuf A, B, C, T;
    LLA_TRY(A.alloc((size_t)mloc * kA * 8));
    LLA_TRY(B.alloc((size_t)kB * nloc * 8));
    LLA_TRY(C.alloc((size_t)mloc * nloc * 8));
    cudaStream_t st = R.main;
    if (m < Mp || k < Kp) LLA_CUDA(cudaMemsetAsync(A.p, 0, (size_t)mloc * kA * 8, st));
    if (n < Np || k < Kp) LLA_CUDA(cudaMemsetAsync(B.p, 0, (size_t)kB * nloc * 8, st));
    if (oa || ob) LLA_TRY(T.alloc((size_t)(mloc > nloc ? mloc : nloc) * (kA > kB ? kA : kB) * 8));
    if (!oa) {
      LLA_TRY(scatter_copy<true>(Scatter{m, k, nb, P, R.p, Q, R.q, 8, 0}, A.p, mloc, (void*)a, lda, st));
    } else {   // stored k x m: its block (J, I) is op(A)'s block (I, J)
      if (m < Mp || k < Kp) LLA_CUDA(cudaMemsetAsync(T.p, 0, (size_t)kA * mloc * 8, st));
      LLA_TRY(scatter_copy<true>(Scatter{k, m, nb, Q, R.q, P, R.p, 8, 0}, T.p, kA, (void*)a, lda, st));
      LLA_TRY(dtranspose_dev(kA, mloc, (const double*)T.p, kA, (double*)A.p, mloc, st));
    }
    if (!ob) {
      LLA_TRY(scatter_copy<true>(Scatter{k, n, nb, P, R.p, Q, R.q, 8, 0}, B.p, kB, (void*)b, ldb, st));
    } else {
      if (n < Np || k < Kp) LLA_CUDA(cudaMemsetAsync(T.p, 0, (size_t)nloc * kB * 8, st));
      LLA_TRY(scatter_copy<true>(Scatter{n, k, nb, Q, R.q, P, R.p, 8, 0}, T.p, nloc, (void*)b, ldb, st));
      LLA_TRY(dtranspose_dev(nloc, kB, (const double*)T.p, nloc, (double*)B.p, kB, st));
    }
    if (beta != 0.0) {
      if (m < Mp || n < Np) LLA_CUDA(cudaMemsetAsync(C.p, 0, (size_t)mloc * nloc * 8, st));
      LLA_TRY(scatter_copy<true>(Scatter{m, n, nb, P, R.p, Q, R.q, 8, 0}, C.p, mloc, c, ldc, st));
    }
    LLA_TRY(summa_rank<double>(R, Mp, Np, Kp, nb, alpha, (const double*)A.p, mloc, (const double*)B.p, kB, beta,
                               (double*)C.p, mloc));
    LLA_TRY(scatter_copy<false>(Scatter{m, n, nb, P, R.p, Q, R.q, 8, 0}, C.p, mloc, c, ldc, st));
    return stage_sync(st);
  });
}

static int pzgemm_host(int64_t m, int64_t n, int64_t k, cuDoubleComplex alpha, const cuDoubleComplex* a, int64_t lda,
                       const cuDoubleComplex* b, int64_t ldb, cuDoubleComplex beta, cuDoubleComplex* c, int64_t ldc) {
  const Engine& E = engine();
  int P, Q;
  grid_from_env("LLACUDA_GEMM_GRID", E.nranks, most_square(E.nranks), &P, &Q);
  LLA_TRY(regrid(P, Q));
  const int64_t mn = m < n ? (m < k ? m : k) : (n < k ? n : k);
  const int64_t nb = env_nb("LLACUDA_GEMM_NB", mn >= 8192 ? 1024 : 512);
  const int64_t Mp = round_up(m, nb * P), Np = round_up(n, nb * Q), Kp = round_up(k, nb * lcm(P, Q));
  const bool need_c = !(beta.x == 0.0 && beta.y == 0.0);
  return run_ranks([&](Rank& R) -> int {
    const int64_t mloc = Mp / P, nloc = Np / Q, kA = Kp / Q, kB = Kp / P;
    PoolBuf A, B, C;
    LLA_TRY(A.alloc((size_t)mloc * kA * 16));
    LLA_TRY(B.alloc((size_t)kB * nloc * 16));
    LLA_TRY(C.alloc((size_t)mloc * nloc * 16));
    cudaStream_t st = R.main;
    if (m < Mp || k < Kp) LLA_CUDA(cudaMemsetAsync(A.p, 0, (size_t)mloc * kA * 16, st));
    if (n < Np || k < Kp) LLA_CUDA(cudaMemsetAsync(B.p, 0, (size_t)kB * nloc * 16, st));
    LLA_TRY(scatter_copy<true>(Scatter{m, k, nb, P, R.p, Q, R.q, 16, 0}, A.p, mloc, (void*)a, lda, st));
    LLA_TRY(scatter_copy<true>(Scatter{k, n, nb, P, R.p, Q, R.q, 16, 0}, B.p, kB, (void*)b, ldb, st));
    if (need_c) {
      if (m < Mp || n < Np) LLA_CUDA(cudaMemsetAsync(C.p, 0, (size_t)mloc * nloc * 16, st));
      LLA_TRY(scatter_copy<true>(Scatter{m, n, nb, P, R.p, Q, R.q, 16, 0}, C.p, mloc, c, ldc, st));
    }
    LLA_TRY(summa_rank<cuDoubleComplex>(R, Mp, Np, Kp, nb, alpha, (const cuDoubleComplex*)A.p, mloc, (const cuDoubleComplex*)B.p, kB,
                                        beta, (cuDoubleComplex*)C.p, mloc));
    LLA_TRY(scatter_copy<false>(Scatter{m, n, nb, P, R.p, Q, R.q, 16, 0}, C.p, mloc, c, ldc, st));
    return stage_sync(st);
  });
}

// ------------------------------------------------------------------ Cholesky
static int chol_grid(int* P, int* Q) {
  const Engine& E = engine();
  grid_from_env("LLACUDA_CHOL_GRID", E.nranks, 1, P, Q);   // measured default: block columns (1 x N), see DESIGN.md section 7
  return regrid(*P, *Q);
}

static int pdpotrf_host(int64_t n, double* a, int64_t lda, int* info_out, double* b, int64_t ldb, int64_t nrhs) {
  int P, Q;
  LLA_TRY(chol_grid(&P, &Q));
  const int64_t nb = env_nb("LLACUDA_DIST_NB", 512);
  const int64_t Np = round_up(n, nb * lcm(P, Q));
  std::vector<int> infos((size_t)engine().nranks, 0);
  int rc = run_ranks([&](Rank& R) -> int {
    const int64_t mloc = Np / P, nloc = Np / Q;
    PoolBuf A, B;
    LLA_TRY(A.alloc((size_t)mloc * nloc * 8));
    cudaStream_t st = R.main;
    if (n < Np) {
      LLA_CUDA(cudaMemsetAsync(A.p, 0, (size_t)mloc * nloc * 8, st));
      lla_pad_identity_kernel<<<(unsigned)((Np - n + 255) / 256), 256, 0, st>>>((double*)A.p, mloc, n, Np, (int)nb, P, Q, R.p, R.q);
      count_launch();
    }
    LLA_TRY(scatter_copy<true>(Scatter{n, n, nb, P, R.p, Q, R.q, 8, 1}, A.p, mloc, a, lda, st));
    if (nrhs > 0) {
      LLA_TRY(B.alloc((size_t)Np * nrhs * 8));
      if (n < Np) LLA_CUDA(cudaMemsetAsync(B.p, 0, (size_t)Np * nrhs * 8, st));
      LLA_TRY(stage_region_h2d(B.p, (size_t)Np * 8, b, (size_t)ldb * 8, (size_t)n * 8, (size_t)nrhs, st));
    }
    LLA_TRY(pdpotrf_rank(R, Np, nb, (double*)A.p, mloc));
    int info = 0;
    LLA_TRY(finish_info(R, &info));
    infos[(size_t)R.rank] = info;
    if (info == 0 && nrhs > 0) {   // POSV: the solve only runs on a successful factorisation
      LLA_TRY(pdtrsv_rank(R, Np, nb, (const double*)A.p, mloc, true, true, false, (double*)B.p, Np, nrhs));
      LLA_TRY(pdtrsv_rank(R, Np, nb, (const double*)A.p, mloc, true, false, false, (double*)B.p, Np, nrhs));
      if (R.rank == 0) LLA_TRY(stage_region_d2h(b, (size_t)ldb * 8, B.p, (size_t)Np * 8, (size_t)n * 8, (size_t)nrhs, st));
    }
    // as LAPACK: the factor (or whatever was computed up to the failing minor) overwrites the named triangle
    LLA_TRY(scatter_copy<false>(Scatter{n, n, nb, P, R.p, Q, R.q, 8, 1}, A.p, mloc, a, lda, st));
    return stage_sync(st);
  });
  if (rc != 0) return rc;
  *info_out = infos[0];
  return 0;
}

// solve with a factor that lives on the host (LLA's (solve cholesky b)): scatter U, two distributed sweeps
static int pdpotrs_host(int64_t n, int64_t nrhs, const double* a, int64_t lda, double* b, int64_t ldb) {
  int P, Q;
  LLA_TRY(chol_grid(&P, &Q));
  const int64_t nb = env_nb("LLACUDA_DIST_NB", 512);
  const int64_t Np = round_up(n, nb * lcm(P, Q));
  return run_ranks([&](Rank& R) -> int {
    const int64_t mloc = Np / P, nloc = Np / Q;
    PoolBuf A, B;
    LLA_TRY(A.alloc((size_t)mloc * nloc * 8));
    LLA_TRY(B.alloc((size_t)Np * nrhs * 8));
    cudaStream_t st = R.main;
    if (n < Np) {
      LLA_CUDA(cudaMemsetAsync(A.p, 0, (size_t)mloc * nloc * 8, st));
      lla_pad_identity_kernel<<<(unsigned)((Np - n + 255) / 256), 256, 0, st>>>((double*)A.p, mloc, n, Np, (int)nb, P, Q, R.p, R.q);
      count_launch();
      LLA_CUDA(cudaMemsetAsync(B.p, 0, (size_t)Np * nrhs * 8, st));
    }
    LLA_TRY(scatter_copy<true>(Scatter{n, n, nb, P, R.p, Q, R.q, 8, 1}, A.p, mloc, (void*)a, lda, st));
    LLA_TRY(stage_region_h2d(B.p, (size_t)Np * 8, b, (size_t)ldb * 8, (size_t)n * 8, (size_t)nrhs, st));
    LLA_TRY(pdtrsv_rank(R, Np, nb, (const double*)A.p, mloc, true, true, false, (double*)B.p, Np, nrhs));
    LLA_TRY(pdtrsv_rank(R, Np, nb, (const double*)A.p, mloc, true, false, false, (double*)B.p, Np, nrhs));
    if (R.rank == 0) LLA_TRY(stage_region_d2h(b, (size_t)ldb * 8, B.p, (size_t)Np * 8, (size_t)n * 8, (size_t)nrhs, st));
    return stage_sync(st);
  });
}

// ------------------------------------------------------------------ LU / solve
static int lu_grid(int* P, int* Q) {
  const Engine& E = engine();
  grid_from_env("LLACUDA_LU_GRID", E.nranks, 1, P, Q);   // block columns: pivot search, interchanges and U(k, J) stay on one GPU
  return regrid(*P, *Q);
}

// square systems; nrhs < 0: factor only (GETRF), else GESV
static int pdgesv_host(int64_t n, int64_t nrhs, double* a, int64_t lda, fint* ipiv, double* b, int64_t ldb, int* info_out) {
  int P, Q;
  LLA_TRY(lu_grid(&P, &Q));
  const int64_t nb = env_nb("LLACUDA_DIST_NB", 512) > 512 ? 512 : env_nb("LLACUDA_DIST_NB", 512);
  const int64_t Np = round_up(n, nb * lcm(P, Q));
  std::vector<int> infos((size_t)engine().nranks, 0);
  std::vector<int> piv32((size_t)max1(n));
  int rc = run_ranks([&](Rank& R) -> int {
    const int64_t mloc = Np / P, nloc = Np / Q;
    PoolBuf A, B, piv;
    LLA_TRY(A.alloc((size_t)mloc * nloc * 8));
    LLA_TRY(piv.alloc((size_t)Np * sizeof(int)));
    cudaStream_t st = R.main;
    if (n < Np) {
      LLA_CUDA(cudaMemsetAsync(A.p, 0, (size_t)mloc * nloc * 8, st));
      lla_pad_identity_kernel<<<(unsigned)((Np - n + 255) / 256), 256, 0, st>>>((double*)A.p, mloc, n, Np, (int)nb, P, Q, R.p, R.q);
      count_launch();
    }
    LLA_TRY(scatter_copy<true>(Scatter{n, n, nb, P, R.p, Q, R.q, 8, 0}, A.p, mloc, a, lda, st));
    if (nrhs > 0) {
      LLA_TRY(B.alloc((size_t)Np * nrhs * 8));
      if (n < Np) LLA_CUDA(cudaMemsetAsync(B.p, 0, (size_t)Np * nrhs * 8, st));
      LLA_TRY(stage_region_h2d(B.p, (size_t)Np * 8, b, (size_t)ldb * 8, (size_t)n * 8, (size_t)nrhs, st));
    }
    bool rode = false;   // the forward substitution rode along with the factorisation (block columns)
    LLA_TRY(pdgetrf_rank(R, Np, nb, (double*)A.p, mloc, (int*)piv.p, nrhs > 0 ? (double*)B.p : nullptr, Np, nrhs, &rode));
    int info = 0;
    LLA_TRY(finish_info(R, &info));
    infos[(size_t)R.rank] = info;
    if (info == 0 && nrhs > 0) {
      if (!rode) {
        LLA_TRY(pdlaswp_rhs_rank(R, Np, (const int*)piv.p, (double*)B.p, Np, nrhs));
        LLA_TRY(pdtrsv_rank(R, Np, nb, (const double*)A.p, mloc, false, false, true, (double*)B.p, Np, nrhs));
      }
      LLA_TRY(pdtrsv_rank(R, Np, nb, (const double*)A.p, mloc, true, false, false, (double*)B.p, Np, nrhs));
      if (R.rank == 0) LLA_TRY(stage_region_d2h(b, (size_t)ldb * 8, B.p, (size_t)Np * 8, (size_t)n * 8, (size_t)nrhs, st));
    }
    if (R.rank == 0 && n > 0) LLA_CUDA(cudaMemcpyAsync(piv32.data(), piv.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
    LLA_TRY(scatter_copy<false>(Scatter{n, n, nb, P, R.p, Q, R.q, 8, 0}, A.p, mloc, a, lda, st));
    return stage_sync(st);
  });
  if (rc != 0) return rc;
  for (int64_t i = 0; i < n; i++) ipiv[i] = (fint)piv32[(size_t)i];
  *info_out = infos[0];
  return 0;
}

// ------------------------------------------------------------------ batched small solves: one batch slice per GPU
static int pdgesv_batched_host(int n, int nrhs, int64_t batch, double* a, double* b, int* ipiv, int* info, int keep_lu) {
  return run_ranks([&](Rank& R) -> int {
    const int64_t lo = batch * R.rank / R.nranks, hi = batch * (R.rank + 1) / R.nranks, cnt = hi - lo;
    if (cnt <= 0) return 0;
    PoolBuf A, B, Pv, In;
    const size_t sa = (size_t)n * n, sb = (size_t)n * nrhs;
    LLA_TRY(A.alloc(sa * cnt * 8));
    LLA_TRY(B.alloc(sb * cnt * 8));
    LLA_TRY(Pv.alloc((size_t)n * cnt * sizeof(int)));
    LLA_TRY(In.alloc((size_t)cnt * sizeof(int)));
    cudaStream_t st = R.main;
    LLA_TRY(stage_region_h2d(A.p, sa * cnt * 8, a + sa * lo, sa * cnt * 8, sa * cnt * 8, 1, st));
    LLA_TRY(stage_region_h2d(B.p, sb * cnt * 8, b + sb * lo, sb * cnt * 8, sb * cnt * 8, 1, st));
    LLA_TRY(dgesv_batched_dev(n, nrhs, cnt, (double*)A.p, (int64_t)sa, (int*)Pv.p, (double*)B.p, (int64_t)sb, (int*)In.p, keep_lu, st));
    LLA_TRY(stage_region_d2h(b + sb * lo, sb * cnt * 8, B.p, sb * cnt * 8, sb * cnt * 8, 1, st));
    if (keep_lu) LLA_TRY(stage_region_d2h(a + sa * lo, sa * cnt * 8, A.p, sa * cnt * 8, sa * cnt * 8, 1, st));
    if (ipiv) LLA_CUDA(cudaMemcpyAsync(ipiv + (size_t)n * lo, Pv.p, (size_t)n * cnt * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (info) LLA_CUDA(cudaMemcpyAsync(info + lo, In.p, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost, st));
    return stage_sync(st);
  });
}

// ------------------------------------------------------------------ routing of the plain Fortran symbols
static int g_auto = -1;
static int64_t g_min_n = 8192;
bool dist_route(int64_t n) {
  if (g_auto < 0) {
    g_auto = 0;
    if (const char* e = getenv("LLACUDA_GPUS")) {
      const int want = atoi(e);
      if (const char* m = getenv("LLACUDA_DIST_MIN_N")) g_min_n = atol(m);
      if (want > 1) {
        if (llacuda_dist_init_all(want, 0, 0) == 0) g_auto = 1;
        else fprintf(stderr, " ** llacuda: LLACUDA_GPUS=%d ignored: %s\n", want, llacuda_last_error());
      }
    }
  }
  const Engine& E = engine();
  return g_auto == 1 && E.ready && E.single_process && E.nranks > 1 && n >= g_min_n;
}
int dist_dgemm(int oa, int ob, int64_t m, int64_t n, int64_t k, double alpha, const double* a, int64_t lda, const double* b,
               int64_t ldb, double beta, double* c, int64_t ldc) {
  return pdgemm_host(oa, ob, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
int dist_dgesv(int64_t n, int64_t nrhs, double* a, int64_t lda, fint* ipiv, double* b, int64_t ldb, int* info) {
  return pdgesv_host(n, nrhs, a, lda, ipiv, b, ldb, info);
}
int dist_dpotrf(int64_t n, double* a, int64_t lda, int* info) { return pdpotrf_host(n, a, lda, info, nullptr, 0, 0); }

}  // namespace llacuda

using namespace llacuda;

static int opc(char c) {
  switch (c) {
    case 'N': case 'n': return 0;
    case 'T': case 't': return 1;
    case 'C': case 'c': return 2;
    default: return -1;
  }
}
static bool engine_single(const char* who) {
  const Engine& E = engine();
  if (E.ready && E.single_process) return true;
  char msg[160];
  snprintf(msg, sizeof(msg), "%s: call llacuda_dist_init_all first (host-pointer multi-GPU entry points run in one process)", who);
  set_error(__FILE__, __LINE__, msg, -1);
  return false;
}

extern "C" {

/* reference call sites: mm src/linear-algebra.lisp:50-54, gemm! src/blas.lisp:56-63 -- one product over all GPUs */
void llacuda_pdgemm(const char* transa, const char* transb, const fint* m_, const fint* n_, const fint* k_, const double* alpha,
                    const double* a, const fint* lda_, const double* b, const fint* ldb_, const double* beta, double* c,
                    const fint* ldc_) {
  ApiLock api_lock;
  llacuda_clear_error();
  const int oa = opc(*transa), ob = opc(*transb);
  const int64_t m = *m_, n = *n_, k = *k_, lda = *lda_, ldb = *ldb_, ldc = *ldc_;
  const int64_t nrowa = oa ? k : m, nrowb = ob ? n : k;
  if (oa < 0 || ob < 0 || m < 0 || n < 0 || k < 0 || lda < max1(nrowa) || ldb < max1(nrowb) || ldc < max1(m)) {
    set_error(__FILE__, __LINE__, "PDGEMM: illegal argument", -1);
    blas_failure("llacuda_pdgemm", LLACUDA_ERR_BAD_ARG, nullptr, 0, 0, 0, 8);
    return;
  }
  if (m == 0 || n == 0) return;
  int rc = engine_single("llacuda_pdgemm") ? pdgemm_host(oa, ob, m, n, k, *alpha, a, lda, b, ldb, *beta, c, ldc) : LLACUDA_ERR_BAD_ARG;
  if (rc != 0) blas_failure("llacuda_pdgemm", rc, c, ldc, m, n, 8);
}

void llacuda_pzgemm(const char* transa, const char* transb, const fint* m_, const fint* n_, const fint* k_, const void* alpha,
                    const void* a, const fint* lda_, const void* b, const fint* ldb_, const void* beta, void* c, const fint* ldc_) {
  ApiLock api_lock;
  llacuda_clear_error();
  const int oa = opc(*transa), ob = opc(*transb);
  if (oa != 0 || ob != 0) {   // transposed complex operands stay on the single-GPU kernel
    zgemm_(transa, transb, m_, n_, k_, alpha, a, lda_, b, ldb_, beta, c, ldc_);
    return;
  }
  const int64_t m = *m_, n = *n_, k = *k_, lda = *lda_, ldb = *ldb_, ldc = *ldc_;
  if (m < 0 || n < 0 || k < 0 || lda < max1(m) || ldb < max1(k) || ldc < max1(m)) {
    set_error(__FILE__, __LINE__, "PZGEMM: illegal argument", -1);
    blas_failure("llacuda_pzgemm", LLACUDA_ERR_BAD_ARG, nullptr, 0, 0, 0, 16);
    return;
  }
  if (m == 0 || n == 0) return;
  int rc = engine_single("llacuda_pzgemm")
               ? pzgemm_host(m, n, k, *(const cuDoubleComplex*)alpha, (const cuDoubleComplex*)a, lda, (const cuDoubleComplex*)b, ldb,
                             *(const cuDoubleComplex*)beta, (cuDoubleComplex*)c, ldc)
               : LLACUDA_ERR_BAD_ARG;
  if (rc != 0) blas_failure("llacuda_pzgemm", rc, c, ldc, m, n, 16);
}

/* reference call site: cholesky, src/linear-algebra.lisp:670-674 (always UPLO = 'U') */
void llacuda_pdpotrf(const char* uplo, const fint* n_, double* a, const fint* lda_, fint* info) {
  ApiLock api_lock;
  const int64_t n = *n_, lda = *lda_;
  const bool upper = (*uplo == 'U' || *uplo == 'u'), lower = (*uplo == 'L' || *uplo == 'l');
  *info = 0;
  if (!upper && !lower) { *info = -1; return; }
  if (n < 0) { *info = -2; return; }
  if (lda < max1(n)) { *info = -4; return; }
  if (n == 0) return;
  if (lower) { dpotrf_(uplo, n_, a, lda_, info); return; }   // LLA never asks for 'L': single-GPU path
  if (!engine_single("llacuda_pdpotrf")) { *info = LLACUDA_ERR_BAD_ARG; return; }
  int inf = 0;
  int rc = pdpotrf_host(n, a, lda, &inf, nullptr, 0, 0);
  *info = (fint)(rc != 0 ? rc : inf);
}

/* reference call site: solve (hermitian-matrix), src/linear-algebra.lisp:303-315: factor and solve in one call */
void llacuda_pdposv(const char* uplo, const fint* n_, const fint* nrhs_, double* a, const fint* lda_, double* b, const fint* ldb_,
                    fint* info) {
  ApiLock api_lock;
  const int64_t n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
  const bool upper = (*uplo == 'U' || *uplo == 'u'), lower = (*uplo == 'L' || *uplo == 'l');
  *info = 0;
  if (!upper && !lower) { *info = -1; return; }
  if (n < 0) { *info = -2; return; }
  if (nrhs < 0) { *info = -3; return; }
  if (lda < max1(n)) { *info = -5; return; }
  if (ldb < max1(n)) { *info = -7; return; }
  if (n == 0) return;
  if (lower) { dposv_(uplo, n_, nrhs_, a, lda_, b, ldb_, info); return; }
  if (!engine_single("llacuda_pdposv")) { *info = LLACUDA_ERR_BAD_ARG; return; }
  int inf = 0;
  int rc = pdpotrf_host(n, a, lda, &inf, b, ldb, nrhs);
  *info = (fint)(rc != 0 ? rc : inf);
}

/* reference call site: solve (cholesky), src/linear-algebra.lisp:296-301 */
void llacuda_pdpotrs(const char* uplo, const fint* n_, const fint* nrhs_, const double* a, const fint* lda_, double* b,
                     const fint* ldb_, fint* info) {
  ApiLock api_lock;
  const int64_t n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
  const bool upper = (*uplo == 'U' || *uplo == 'u'), lower = (*uplo == 'L' || *uplo == 'l');
  *info = 0;
  if (!upper && !lower) { *info = -1; return; }
  if (n < 0) { *info = -2; return; }
  if (nrhs < 0) { *info = -3; return; }
  if (lda < max1(n)) { *info = -5; return; }
  if (ldb < max1(n)) { *info = -7; return; }
  if (n == 0 || nrhs == 0) return;
  if (lower) { dpotrs_(uplo, n_, nrhs_, a, lda_, b, ldb_, info); return; }
  if (!engine_single("llacuda_pdpotrs")) { *info = LLACUDA_ERR_BAD_ARG; return; }
  int rc = pdpotrs_host(n, nrhs, a, lda, b, ldb);
  if (rc != 0) *info = (fint)rc;
}

/* reference call site: lu, src/linear-algebra.lisp:227-234 (square matrices go over all GPUs, rectangular ones to one) */
void llacuda_pdgetrf(const fint* m_, const fint* n_, double* a, const fint* lda_, fint* ipiv, fint* info) {
  ApiLock api_lock;
  const int64_t m = *m_, n = *n_, lda = *lda_;
  *info = 0;
  if (m < 0) { *info = -1; return; }
  if (n < 0) { *info = -2; return; }
  if (lda < max1(m)) { *info = -4; return; }
  if (m == 0 || n == 0) return;
  if (m != n) { dgetrf_(m_, n_, a, lda_, ipiv, info); return; }
  if (!engine_single("llacuda_pdgetrf")) { *info = LLACUDA_ERR_BAD_ARG; return; }
  int inf = 0;
  int rc = pdgesv_host(n, -1, a, lda, ipiv, nullptr, 0, &inf);
  *info = (fint)(rc != 0 ? rc : inf);
}

/* reference call site: solve (array array), src/linear-algebra.lisp:282-289 */
void llacuda_pdgesv(const fint* n_, const fint* nrhs_, double* a, const fint* lda_, fint* ipiv, double* b, const fint* ldb_,
                    fint* info) {
  ApiLock api_lock;
  const int64_t n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
  *info = 0;
  if (n < 0) { *info = -1; return; }
  if (nrhs < 0) { *info = -2; return; }
  if (lda < max1(n)) { *info = -4; return; }
  if (ldb < max1(n)) { *info = -7; return; }
  if (n == 0) return;
  if (!engine_single("llacuda_pdgesv")) { *info = LLACUDA_ERR_BAD_ARG; return; }
  int inf = 0;
  int rc = pdgesv_host(n, nrhs, a, lda, ipiv, b, ldb, &inf);
  *info = (fint)(rc != 0 ? rc : inf);
}

/* BASELINE configs[3]: `batch` independent n x n systems (n <= 32, column-major, contiguous), one slice per GPU, no
 * collective on the path.  ipiv (batch*n) and info (batch) are int32 host arrays (may be NULL). */
int llacuda_pdgesv_batched(int n, int nrhs, int64_t batch, double* a, double* b, int* ipiv, int* info, int keep_lu) {
  ApiLock api_lock;
  if (n < 0 || n > 32 || nrhs < 0 || nrhs > 8 || batch < 0) return LLACUDA_ERR_BAD_ARG;
  if (n == 0 || batch == 0) return 0;
  if (!engine_single("llacuda_pdgesv_batched")) return LLACUDA_ERR_BAD_ARG;
  return pdgesv_batched_host(n, nrhs, batch, a, b, ipiv, info, keep_lu);
}

/* ---- one rank per process: operands already block-cyclic in HBM (see dist.h for the local layout) ---- */
static bool local_rank(const char* who) {   // true (error recorded) unless this process is one rank of a launcher job
  const Engine& E = engine();
  if (E.ready && !E.single_process) return false;
  char msg[160];
  snprintf(msg, sizeof(msg), "%s: call llacuda_dist_init_rank first (one rank per process)", who);
  set_error(__FILE__, __LINE__, msg, -1);
  return true;
}

int llacuda_pdgemm_local(int64_t Mp, int64_t Np, int64_t Kp, int64_t nb, double alpha, const double* A, int64_t lda,
                         const double* B, int64_t ldb, double beta, double* C, int64_t ldc) {
  ApiLock api_lock;
  if (local_rank("llacuda_pdgemm_local")) return LLACUDA_ERR_BAD_ARG;
  return run_ranks([&](Rank& R) { return summa_rank<double>(R, Mp, Np, Kp, nb, alpha, A, lda, B, ldb, beta, C, ldc); });
}
int llacuda_pzgemm_local(int64_t Mp, int64_t Np, int64_t Kp, int64_t nb, const double* alpha2, const void* A, int64_t lda,
                         const void* B, int64_t ldb, const double* beta2, void* C, int64_t ldc) {
  ApiLock api_lock;
  if (local_rank("llacuda_pzgemm_local")) return LLACUDA_ERR_BAD_ARG;
  return run_ranks([&](Rank& R) {
    return summa_rank<cuDoubleComplex>(R, Mp, Np, Kp, nb, make_cuDoubleComplex(alpha2[0], alpha2[1]), (const cuDoubleComplex*)A, lda,
                                       (const cuDoubleComplex*)B, ldb, make_cuDoubleComplex(beta2[0], beta2[1]), (cuDoubleComplex*)C, ldc);
  });
}
int llacuda_pdpotrf_local(int64_t Np, int64_t nb, double* A, int64_t lld, int* info) {
  ApiLock api_lock;
  if (local_rank("llacuda_pdpotrf_local")) return LLACUDA_ERR_BAD_ARG;
  return run_ranks([&](Rank& R) -> int {
    LLA_TRY(pdpotrf_rank(R, Np, nb, A, lld));
    return finish_info(R, info);
  });
}
int llacuda_pdgetrf_local(int64_t Np, int64_t nb, double* A, int64_t lld, int* ipiv_dev, int* info) {
  ApiLock api_lock;
  if (local_rank("llacuda_pdgetrf_local")) return LLACUDA_ERR_BAD_ARG;
  return run_ranks([&](Rank& R) -> int {
    LLA_TRY(pdgetrf_rank(R, Np, nb, A, lld, ipiv_dev));
    return finish_info(R, info);
  });
}
/* factorisation and solve in one call: B (Np x nrhs, ldb, replicated on every rank) is overwritten by X when *info == 0
   and holds garbage otherwise (the caller keeps the original, as LLA does); on block columns the forward substitution
   rides along with the factorisation */
int llacuda_pdgesv_local(int64_t Np, int64_t nb, double* A, int64_t lld, int* ipiv_dev, double* B, int64_t ldb, int64_t nrhs,
                         int* info) {
  ApiLock api_lock;
  if (local_rank("llacuda_pdgesv_local")) return LLACUDA_ERR_BAD_ARG;
  return run_ranks([&](Rank& R) -> int {
    bool rode = false;
    LLA_TRY(pdgetrf_rank(R, Np, nb, A, lld, ipiv_dev, nrhs > 0 ? B : nullptr, ldb, nrhs, &rode));
    LLA_TRY(finish_info(R, info));
    if (*info != 0 || nrhs <= 0) return 0;
    if (!rode) {
      LLA_TRY(pdlaswp_rhs_rank(R, Np, ipiv_dev, B, ldb, nrhs));
      LLA_TRY(pdtrsv_rank(R, Np, nb, A, lld, false, false, true, B, ldb, nrhs));
    }
    return pdtrsv_rank(R, Np, nb, A, lld, true, false, false, B, ldb, nrhs);
  });
}
/* B: Np x nrhs (ldb) replicated on every rank, overwritten by X */
int llacuda_pdgetrs_local(int64_t Np, int64_t nb, const double* LU, int64_t lld, const int* ipiv_dev, double* B, int64_t ldb,
                          int64_t nrhs) {
  ApiLock api_lock;
  if (local_rank("llacuda_pdgetrs_local")) return LLACUDA_ERR_BAD_ARG;
  return run_ranks([&](Rank& R) -> int {
    LLA_TRY(pdlaswp_rhs_rank(R, Np, ipiv_dev, B, ldb, nrhs));
    LLA_TRY(pdtrsv_rank(R, Np, nb, LU, lld, false, false, true, B, ldb, nrhs));
    return pdtrsv_rank(R, Np, nb, LU, lld, true, false, false, B, ldb, nrhs);
  });
}
int llacuda_pdpotrs_local(int64_t Np, int64_t nb, const double* U, int64_t lld, double* B, int64_t ldb, int64_t nrhs) {
  ApiLock api_lock;
  if (local_rank("llacuda_pdpotrs_local")) return LLACUDA_ERR_BAD_ARG;
  return run_ranks([&](Rank& R) -> int {
    LLA_TRY(pdtrsv_rank(R, Np, nb, U, lld, true, true, false, B, ldb, nrhs));
    return pdtrsv_rank(R, Np, nb, U, lld, true, false, false, B, ldb, nrhs);
  });
}
/* the stream the rank functions run their trailing updates on and join before they return (for event timing) */
void* llacuda_dist_stream(void) {
  void* s = nullptr;
  if (engine().ready && !engine().single_process) run_ranks([&](Rank& R) { s = (void*)R.main; return 0; });
  return s;
}

/* pure ownership arithmetic (also the CPU-testable part of the layout): number of rows / columns of an n-long dimension
 * in blocks of nb that process `iproc` of `nprocs` owns (ScaLAPACK's NUMROC with source process 0) */
int64_t llacuda_dist_numroc(int64_t n, int64_t nb, int iproc, int nprocs) {
  if (n <= 0 || nb <= 0 || nprocs <= 0 || iproc < 0 || iproc >= nprocs) return 0;
  const int64_t nblocks = n / nb, extra = nblocks % nprocs;
  int64_t r = (nblocks / nprocs) * nb;
  if (iproc < extra) r += nb;
  else if (iproc == extra) r += n % nb;
  return r;
}
/* padded order the factorisations run at: the next multiple of nb * lcm(P, Q) */
int64_t llacuda_dist_padded(int64_t n, int64_t nb, int P, int Q) {
  if (n < 0 || nb <= 0 || P <= 0 || Q <= 0) return -1;
  return round_up(n, nb * lcm(P, Q));
}
/* global index of local index l (0-based) on process iproc, blocks of nb over nprocs */
int64_t llacuda_dist_l2g(int64_t l, int64_t nb, int iproc, int nprocs) {
  return ((l / nb) * nprocs + iproc) * nb + l % nb;
}
void llacuda_dist_default_grid(int world, int* P, int* Q) {
  int p = most_square(world > 0 ? world : 1);
  if (P) *P = p;
  if (Q) *Q = (world > 0 ? world : 1) / p;
}

}  // extern "C"
