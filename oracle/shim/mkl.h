/* oracle/shim/mkl.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A stand-in for Intel MKL's umbrella header so that the UNMODIFIED reference sources under
 * /root/reference/src compile in an image without MKL.  It provides
 *   - MKL_INT as a 64-bit integer (the reference builds with -DMKL_ILP64, src/Makefile:2),
 *   - CBLAS / LAPACKE names forwarded to the ILP64 OpenBLAS that ships inside numpy
 *     (libscipy_openblas64_: symbols scipy_cblas_*64_ / scipy_LAPACKE_*64_),
 *   - a restatement of the inspector-executor sparse BLAS entry points the reference calls
 *     (mkl_sparse_{d,z}_create_csr, mkl_sparse_{d,z}_mv, mkl_sparse_destroy; call sites
 *     src/sparse.cc:8-40,129,165,185,258,287).  MKL is closed source; the semantics restated here
 *     are the documented ones for descr = {GENERAL | SYMMETRIC | HERMITIAN, FILL_UPPER, NON_UNIT}
 *     and agree with the reference's own csr_mat::to_dense (src/sparse.cc:299-315),
 *   - stubs for FEAST and version queries (out of scope for the H*v path).
 */
#ifndef QB_ORACLE_SHIM_MKL_H
#define QB_ORACLE_SHIM_MKL_H

#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#ifndef MKL_INT
#define MKL_INT long long
#endif
#define MKL_INT64 long long
#ifndef MKL_Complex16
#define MKL_Complex16 std::complex<double>
#endif
#ifndef lapack_int
#define lapack_int MKL_INT
#endif
#ifndef lapack_complex_double
#define lapack_complex_double MKL_Complex16
#endif

/* ---------------------------------------------------------------- CBLAS ---- */
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_LAYOUT;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;

extern "C" {
void   scipy_cblas_daxpy64_(MKL_INT n, double a, const double *x, MKL_INT incx, double *y, MKL_INT incy);
void   scipy_cblas_zaxpy64_(MKL_INT n, const void *a, const void *x, MKL_INT incx, void *y, MKL_INT incy);
void   scipy_cblas_dcopy64_(MKL_INT n, const double *x, MKL_INT incx, double *y, MKL_INT incy);
void   scipy_cblas_zcopy64_(MKL_INT n, const void *x, MKL_INT incx, void *y, MKL_INT incy);
double scipy_cblas_dnrm264_(MKL_INT n, const double *x, MKL_INT incx);
double scipy_cblas_dznrm264_(MKL_INT n, const void *x, MKL_INT incx);
void   scipy_cblas_dscal64_(MKL_INT n, double a, double *x, MKL_INT incx);
void   scipy_cblas_zscal64_(MKL_INT n, const void *a, void *x, MKL_INT incx);
double scipy_cblas_ddot64_(MKL_INT n, const double *x, MKL_INT incx, const double *y, MKL_INT incy);
void   scipy_cblas_zdotc_sub64_(MKL_INT n, const void *x, MKL_INT incx, const void *y, MKL_INT incy, void *res);
void   scipy_cblas_dgemm64_(CBLAS_LAYOUT, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, MKL_INT m, MKL_INT n, MKL_INT k,
                            double alpha, const double *a, MKL_INT lda, const double *b, MKL_INT ldb,
                            double beta, double *c, MKL_INT ldc);
void   scipy_cblas_zgemm64_(CBLAS_LAYOUT, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, MKL_INT m, MKL_INT n, MKL_INT k,
                            const void *alpha, const void *a, MKL_INT lda, const void *b, MKL_INT ldb,
                            const void *beta, void *c, MKL_INT ldc);
}
#define cblas_daxpy     scipy_cblas_daxpy64_
#define cblas_zaxpy     scipy_cblas_zaxpy64_
#define cblas_dcopy     scipy_cblas_dcopy64_
#define cblas_zcopy     scipy_cblas_zcopy64_
#define cblas_dnrm2     scipy_cblas_dnrm264_
#define cblas_dznrm2    scipy_cblas_dznrm264_
#define cblas_dscal     scipy_cblas_dscal64_
#define cblas_zscal     scipy_cblas_zscal64_
#define cblas_ddot      scipy_cblas_ddot64_
#define cblas_zdotc_sub scipy_cblas_zdotc_sub64_
#define cblas_dgemm     scipy_cblas_dgemm64_
#define cblas_zgemm     scipy_cblas_zgemm64_

/* -------------------------------------------------------------- LAPACKE ---- */
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
extern "C" {
MKL_INT scipy_LAPACKE_dstedc64_(int layout, char compz, MKL_INT n, double *d, double *e, double *z, MKL_INT ldz);
MKL_INT scipy_LAPACKE_dsyevd64_(int layout, char jobz, char uplo, MKL_INT n, double *a, MKL_INT lda, double *w);
MKL_INT scipy_LAPACKE_zheevd64_(int layout, char jobz, char uplo, MKL_INT n, void *a, MKL_INT lda, double *w);
MKL_INT scipy_LAPACKE_dgesv64_(int layout, MKL_INT n, MKL_INT nrhs, double *a, MKL_INT lda, MKL_INT *ipiv,
                               double *b, MKL_INT ldb);
}
#define LAPACKE_dstedc scipy_LAPACKE_dstedc64_
#define LAPACKE_dsyevd scipy_LAPACKE_dsyevd64_
#define LAPACKE_zheevd scipy_LAPACKE_zheevd64_
#define LAPACKE_dgesv  scipy_LAPACKE_dgesv64_

/* ------------------------------------------- inspector-executor sparse BLAS */
typedef enum { SPARSE_STATUS_SUCCESS = 0, SPARSE_STATUS_NOT_INITIALIZED = 1, SPARSE_STATUS_ALLOC_FAILED = 2,
               SPARSE_STATUS_INVALID_VALUE = 3, SPARSE_STATUS_EXECUTION_FAILED = 4,
               SPARSE_STATUS_INTERNAL_ERROR = 5, SPARSE_STATUS_NOT_SUPPORTED = 6 } sparse_status_t;
typedef enum { SPARSE_INDEX_BASE_ZERO = 0, SPARSE_INDEX_BASE_ONE = 1 } sparse_index_base_t;
typedef enum { SPARSE_OPERATION_NON_TRANSPOSE = 10, SPARSE_OPERATION_TRANSPOSE = 11,
               SPARSE_OPERATION_CONJUGATE_TRANSPOSE = 12 } sparse_operation_t;
typedef enum { SPARSE_MATRIX_TYPE_GENERAL = 20, SPARSE_MATRIX_TYPE_SYMMETRIC = 21,
               SPARSE_MATRIX_TYPE_HERMITIAN = 22, SPARSE_MATRIX_TYPE_TRIANGULAR = 23,
               SPARSE_MATRIX_TYPE_DIAGONAL = 24 } sparse_matrix_type_t;
typedef enum { SPARSE_FILL_MODE_LOWER = 40, SPARSE_FILL_MODE_UPPER = 41, SPARSE_FILL_MODE_FULL = 42 } sparse_fill_mode_t;
typedef enum { SPARSE_DIAG_NON_UNIT = 50, SPARSE_DIAG_UNIT = 51 } sparse_diag_type_t;
struct matrix_descr { sparse_matrix_type_t type; sparse_fill_mode_t mode; sparse_diag_type_t diag; };

/* The handle aliases the caller's arrays, exactly like MKL's 4-array create_csr (no copy). */
struct qb_shim_sparse_matrix {
    MKL_INT rows, cols;
    const MKL_INT *rows_start, *rows_end, *col_indx;
    const void *values;
    int is_complex;
};
typedef qb_shim_sparse_matrix *sparse_matrix_t;

/* Worker pool size for the restated mat-vec (set by the oracle drivers; 1 = serial). */
extern "C" int  qb_shim_get_spmv_threads(void);
extern "C" void qb_shim_set_spmv_threads(int nthreads);

sparse_status_t mkl_sparse_d_create_csr(sparse_matrix_t *A, sparse_index_base_t indexing, MKL_INT rows, MKL_INT cols,
                                        MKL_INT *rows_start, MKL_INT *rows_end, MKL_INT *col_indx, double *values);
sparse_status_t mkl_sparse_z_create_csr(sparse_matrix_t *A, sparse_index_base_t indexing, MKL_INT rows, MKL_INT cols,
                                        MKL_INT *rows_start, MKL_INT *rows_end, MKL_INT *col_indx,
                                        MKL_Complex16 *values);
sparse_status_t mkl_sparse_destroy(sparse_matrix_t A);
sparse_status_t mkl_sparse_d_mv(sparse_operation_t op, double alpha, const sparse_matrix_t A, struct matrix_descr descr,
                                const double *x, double beta, double *y);
sparse_status_t mkl_sparse_z_mv(sparse_operation_t op, MKL_Complex16 alpha, const sparse_matrix_t A,
                                struct matrix_descr descr, const MKL_Complex16 *x, MKL_Complex16 beta,
                                MKL_Complex16 *y);

/* ------------------------------------------------------- FEAST (stubbed) --- */
void feastinit(MKL_INT *fpm);
void zfeast_hcsrev(const char *uplo, const MKL_INT *n, const MKL_Complex16 *a, const MKL_INT *ia, const MKL_INT *ja,
                   MKL_INT *fpm, double *epsout, MKL_INT *loop, const double *emin, const double *emax, MKL_INT *m0,
                   double *e, MKL_Complex16 *x, MKL_INT *m, double *res, MKL_INT *info);

/* ------------------------------------------------------ version queries ---- */
typedef struct { int MajorVersion, MinorVersion, UpdateVersion; const char *ProductStatus, *Build, *Processor, *Platform; } MKLVersion;
void MKL_Get_Version(MKLVersion *ver);
int  mkl_get_max_threads(void);

#endif
