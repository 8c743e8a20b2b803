/* oracle/netlib_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU oracle for the llacuda hot path: a scalar C restatement of the netlib
 * reference BLAS/LAPACK algorithms behind LLA's FFI call sites, instantiated
 * for LLA's four float types (src/types.lisp:22-26: 0=S 1=D 2=C 3=Z).
 * Symbols: oracle_{s,d,c,z}{gemm,getrf,getrs,gesv,potrf,potrs,syrk,trsm,posv,trtri,potri,getri}_  with the exact
 * Fortran pointer-only signatures LLA passes (src/fortran-call.lisp:418-439).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product (libllacuda.so) never does.
 * Pinned against the reference's own goldens in tests/test_oracle.py
 * (tests/linear-algebra.lisp:9-29,89-115,366-385; tests/blas.lisp:19-46) and
 * cross-checked against host OpenBLAS issuing the same calls.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <float.h>

#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)

/* ---- single ---- */
#define T float
#define R float
#define PFX(name) CAT(oracle_s, name)
#define IS_CPLX 0
#define R_ABS fabsf
#define R_SQRT sqrtf
#define R_EPS_SFMIN FLT_MIN
#include "netlib_templ.h"
#undef T
#undef R
#undef PFX
#undef IS_CPLX
#undef R_ABS
#undef R_SQRT
#undef R_EPS_SFMIN

/* ---- double ---- */
#define T double
#define R double
#define PFX(name) CAT(oracle_d, name)
#define IS_CPLX 0
#define R_ABS fabs
#define R_SQRT sqrt
#define R_EPS_SFMIN DBL_MIN
#include "netlib_templ.h"
#undef T
#undef R
#undef PFX
#undef IS_CPLX
#undef R_ABS
#undef R_SQRT
#undef R_EPS_SFMIN

/* ---- complex single ---- */
#define T float _Complex
#define R float
#define PFX(name) CAT(oracle_c, name)
#define IS_CPLX 1
#define R_ABS fabsf
#define R_SQRT sqrtf
#define R_EPS_SFMIN FLT_MIN
#define creal_ crealf
#define cimag_ cimagf
#define conj_ conjf
#include "netlib_templ.h"
#undef T
#undef R
#undef PFX
#undef IS_CPLX
#undef R_ABS
#undef R_SQRT
#undef R_EPS_SFMIN
#undef creal_
#undef cimag_
#undef conj_

/* ---- complex double ---- */
#define T double _Complex
#define R double
#define PFX(name) CAT(oracle_z, name)
#define IS_CPLX 1
#define R_ABS fabs
#define R_SQRT sqrt
#define R_EPS_SFMIN DBL_MIN
#define creal_ creal
#define cimag_ cimag
#define conj_ conj
#include "netlib_templ.h"
