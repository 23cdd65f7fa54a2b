/* oracle/qb_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's H*v hot path (wztzjhn/quantum_basis).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (quantum_basis_b200/) never does.  Every function cites the reference lines it restates.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle.py) against (1) the reference's
 * own golden values (src/main_test.cc:88; examples/**: E0 of L=16 Heisenberg, 4x4 triangular k-sectors,
 * 4x2 Hubbard, honeycomb GENERAL matrix) and (2) outputs of the unmodified reference compiled here
 * (oracle/_ref/qb_ref: per-product y=H*x vectors, Lanczos a/b coefficients, step counts, CG vectors)
 * committed under tests/golden/.  Exception: the Chebyshev/KPM moments have NO reference implementation
 * (src/kpm.cc holds only energy_scale) -> "parity unpinned" for qbo_kpm_moments_*.
 *
 * Scalars: suffix _d = double, _z = double complex (layout-compatible with std::complex<double>).
 * Indices: int64 (the reference's MKL_INT under -DMKL_ILP64).
 */
#ifndef QB_ORACLE_H
#define QB_ORACLE_H
#include <stdint.h>
#include <complex.h>
#ifdef __cplusplus
extern "C" {
#endif

#define QBO_LANCZOS_PRECISION 2e-12   /* src/miscellaneous.cc:47 */

/* vec_randomize (src/miscellaneous.cc:371-388): minstd_rand0(seed), x_j = g()/2147483647 - 0.5, then /nrm2.
 * seed == 0 -> constant 1/sqrt(n). */
void qbo_vec_randomize_d(int64_t n, double *x, uint32_t seed);
void qbo_vec_randomize_z(int64_t n, double _Complex *x, uint32_t seed);

/* csr_mat<T>::MultMv2 / MultMv (src/sparse.cc:262-297): y (+)= H x.  sym != 0: only the upper triangle is
 * stored; (i,j>i,v) also contributes y_j += conj(v) x_i.  accumulate != 0 -> MultMv2, else MultMv.
 * nthreads <= 1: serial, row order. */
void qbo_spmv_d(int64_t n, const int64_t *ia, const int64_t *ja, const double *val, int sym,
                const double *x, double *y, int accumulate, int nthreads);
void qbo_spmv_z(int64_t n, const int64_t *ia, const int64_t *ja, const double _Complex *val, int sym,
                const double _Complex *x, double _Complex *y, int accumulate, int nthreads);
/* same product accumulated in long double: the arbiter when two fp64 results disagree near 1e-12 */
void qbo_spmv_z_ld(int64_t n, const int64_t *ia, const int64_t *ja, const double _Complex *val, int sym,
                   const double _Complex *x, double _Complex *y);

/* Expansion of an upper-triangle Hermitian CSR to the full matrix, rows sorted by column (the meaning of
 * `sym` in csr_mat::to_dense, src/sparse.cc:299-315).  ia_full has n+1 entries; pass ja_full/val_full = NULL
 * to only count.  Returns the full nnz. */
int64_t qbo_expand_upper_z(int64_t n, const int64_t *ia, const int64_t *ja, const double _Complex *val,
                           int64_t *ia_full, int64_t *ja_full, double _Complex *val_full);

/* hess_eigen (src/lanczos.cc:355-390), order "sr": eigenvalues of the m x m tridiagonal (diag a[0..m),
 * off-diag b[1..m)) ascending in ritz[m]; eigenvectors column-major in s[m*m] (s may be NULL). */
int qbo_hess_eigen(const double *hessenberg, int64_t maxit, int64_t m, double *ritz, double *s);

/* lanczos<T,MAT> (src/lanczos.cc:134-266) started from k=0 with np = maxit-1, purposes "sr_val0",
 * "sr_val1" (phi0 = v+2n) and "dnmcs".  v holds 2 (or 3) vectors of length n, v[0..n) normalised on entry.
 * hessenberg[2*maxit]: b in [0,maxit), a in [maxit,2maxit).  Returns m (steps done). */
int64_t qbo_lanczos_d(int64_t n, const int64_t *ia, const int64_t *ja, const double *val, int sym,
                      double *v, double *hessenberg, int64_t maxit, const char *purpose, int nthreads);
int64_t qbo_lanczos_z(int64_t n, const int64_t *ia, const int64_t *ja, const double _Complex *val, int sym,
                      double _Complex *v, double *hessenberg, int64_t maxit, const char *purpose, int nthreads);

/* eigenvec_CG<T,MAT> (src/lanczos.cc:281-341) from m=0.  Returns steps; *accu = final residual norm. */
int64_t qbo_eigenvec_cg_d(int64_t n, const int64_t *ia, const int64_t *ja, const double *val, int sym,
                          double E0, int64_t maxit, double *accu, double *v, double *r, double *p, double *pp, int nthreads);
int64_t qbo_eigenvec_cg_z(int64_t n, const int64_t *ia, const int64_t *ja, const double _Complex *val, int sym,
                          double _Complex E0, int64_t maxit, double *accu, double _Complex *v, double _Complex *r,
                          double _Complex *p, double _Complex *pp, int nthreads);
/* E0 as a pointer to (re, im) for callers whose FFI cannot pass a complex scalar by value (ctypes) */
int64_t qbo_eigenvec_cg_zp(int64_t n, const int64_t *ia, const int64_t *ja, const double _Complex *val, int sym,
                           const double *E0_reim, int64_t maxit, double *accu, double _Complex *v, double _Complex *r,
                           double _Complex *p, double _Complex *pp, int nthreads);

/* energy_scale<T,MAT> (src/kpm.cc:45-88) with the start vector given in v[0..n) (the reference draws
 * vec_randomize(seed=1)); v needs 2n entries. */
void qbo_energy_scale_z(int64_t n, const int64_t *ia, const int64_t *ja, const double _Complex *val, int sym,
                        double _Complex *v, double *lo, double *hi, double extend, int64_t iters, int nthreads);

/* NEW functionality (no reference counterpart, parity unpinned): Chebyshev moments
 * mu_k = <phi| T_k(Ht) |phi>, Ht = (H - c)/s with c = (hi+lo)/2, s = (hi-lo)/2, k = 0..nmom-1, computed by the
 * plain three-term recurrence (one H*v per moment, no doubling trick). */
void qbo_kpm_moments_z(int64_t n, const int64_t *ia, const int64_t *ja, const double _Complex *val, int sym,
                       const double _Complex *phi, double lo, double hi, int64_t nmom, double *mu, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
