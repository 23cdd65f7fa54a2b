/* oracle/shim/mkl_shim.cc -- TEST INFRASTRUCTURE ONLY.
 *
 * Definitions behind oracle/shim/mkl.h.  The only arithmetic of interest is the restated
 * mkl_sparse_{d,z}_mv: y := alpha*op(A)*x + beta*y for a zero-based 4-array CSR whose descriptor says
 * GENERAL, or SYMMETRIC/HERMITIAN with FILL_UPPER (what csr_mat<T>::MultMv2 passes, reference
 * src/sparse.cc:262-289).  For the SYMMETRIC/HERMITIAN case only entries with col >= row are read;
 * a stored (i,j,v) with j>i contributes y_i += v*x_j and y_j += conj(v)*x_i, the same meaning the
 * reference gives `sym` in csr_mat::to_dense (src/sparse.cc:299-315).
 *
 * Two execution modes: serial (row order, the arbiter for parity) and a row-partitioned OpenMP mode
 * with per-thread accumulation windows for the transposed half (the CPU baseline "with all the host
 * threads it can use").
 */
#include "mkl.h"
#include <omp.h>
#include <vector>
#include <type_traits>

static int g_spmv_threads = 1;
extern "C" int  qb_shim_get_spmv_threads(void) { return g_spmv_threads; }
extern "C" void qb_shim_set_spmv_threads(int n) { g_spmv_threads = n < 1 ? 1 : n; }

static inline double               cj(double v) { return v; }
static inline std::complex<double> cj(const std::complex<double> &v) { return std::conj(v); }

template <typename T>
static sparse_status_t create_csr(sparse_matrix_t *A, sparse_index_base_t indexing, MKL_INT rows, MKL_INT cols,
                                  MKL_INT *rs, MKL_INT *re, MKL_INT *ci, T *values)
{
    if (A == nullptr || indexing != SPARSE_INDEX_BASE_ZERO || rows < 0 || cols < 0) return SPARSE_STATUS_INVALID_VALUE;
    auto *h = new qb_shim_sparse_matrix;
    h->rows = rows; h->cols = cols; h->rows_start = rs; h->rows_end = re; h->col_indx = ci; h->values = values;
    h->is_complex = std::is_same<T, double>::value ? 0 : 1;
    *A = h;
    return SPARSE_STATUS_SUCCESS;
}

sparse_status_t mkl_sparse_d_create_csr(sparse_matrix_t *A, sparse_index_base_t indexing, MKL_INT rows, MKL_INT cols,
                                        MKL_INT *rs, MKL_INT *re, MKL_INT *ci, double *values)
{ return create_csr<double>(A, indexing, rows, cols, rs, re, ci, values); }

sparse_status_t mkl_sparse_z_create_csr(sparse_matrix_t *A, sparse_index_base_t indexing, MKL_INT rows, MKL_INT cols,
                                        MKL_INT *rs, MKL_INT *re, MKL_INT *ci, MKL_Complex16 *values)
{ return create_csr<MKL_Complex16>(A, indexing, rows, cols, rs, re, ci, values); }

sparse_status_t mkl_sparse_destroy(sparse_matrix_t A)
{
    delete A;          /* deleting nullptr is fine: csr_mat's dtor passes a null handle for empty objects */
    return SPARSE_STATUS_SUCCESS;
}

template <typename T>
static sparse_status_t mv(sparse_operation_t op, T alpha, const sparse_matrix_t A, struct matrix_descr descr,
                          const T *x, T beta, T *y)
{
    if (A == nullptr || x == nullptr || y == nullptr) return SPARSE_STATUS_NOT_INITIALIZED;
    if (op != SPARSE_OPERATION_NON_TRANSPOSE) return SPARSE_STATUS_NOT_SUPPORTED;
    if (descr.type != SPARSE_MATRIX_TYPE_GENERAL &&
        !((descr.type == SPARSE_MATRIX_TYPE_SYMMETRIC || descr.type == SPARSE_MATRIX_TYPE_HERMITIAN) &&
          descr.mode == SPARSE_FILL_MODE_UPPER && descr.diag == SPARSE_DIAG_NON_UNIT))
        return SPARSE_STATUS_NOT_SUPPORTED;
    const MKL_INT n = A->rows;
    const MKL_INT *rs = A->rows_start, *re = A->rows_end, *ci = A->col_indx;
    const T *val = static_cast<const T *>(A->values);
    const bool conj_half = (descr.type == SPARSE_MATRIX_TYPE_HERMITIAN);
    const bool general   = (descr.type == SPARSE_MATRIX_TYPE_GENERAL);

    if (beta != T(1.0)) for (MKL_INT i = 0; i < n; i++) y[i] = (beta == T(0.0)) ? T(0.0) : beta * y[i];

    int nt = g_spmv_threads;
    if (general) {
        #pragma omp parallel for schedule(static) num_threads(nt)
        for (MKL_INT i = 0; i < n; i++) {
            T acc = T(0.0);
            for (MKL_INT p = rs[i]; p < re[i]; p++) acc += val[p] * x[ci[p]];
            y[i] += alpha * acc;
        }
        return SPARSE_STATUS_SUCCESS;
    }
    if (nt == 1) {
        for (MKL_INT i = 0; i < n; i++) {
            T acc = T(0.0);
            const T xi = x[i];
            for (MKL_INT p = rs[i]; p < re[i]; p++) {
                const MKL_INT j = ci[p];
                if (j < i) continue;                       /* FILL_UPPER: the lower part is not referenced */
                acc += val[p] * x[j];
                if (j != i) y[j] += alpha * ((conj_half ? cj(val[p]) : val[p]) * xi);
            }
            y[i] += alpha * acc;
        }
        return SPARSE_STATUS_SUCCESS;
    }
    /* threaded: nnz-balanced row ranges; thread t owns rows [lo_t, hi_t) and scatters the transposed half
       into a private window covering [lo_t, n), reduced afterwards in thread order (deterministic). */
    std::vector<MKL_INT> bound(nt + 1, n);
    bound[0] = 0;
    {
        const MKL_INT nnz = re[n - 1] - rs[0];
        int t = 1;
        for (MKL_INT i = 0; i < n && t < nt; i++)
            while (t < nt && (re[i] - rs[0]) >= (nnz / nt) * t) bound[t++] = i + 1;
    }
    std::vector<std::vector<T>> win(nt);
    #pragma omp parallel num_threads(nt)
    {
        const int t = omp_get_thread_num();
        const MKL_INT lo = bound[t], hi = bound[t + 1];
        std::vector<T> &w = win[t];
        w.assign(static_cast<size_t>(n - lo), T(0.0));
        for (MKL_INT i = lo; i < hi; i++) {
            T acc = T(0.0);
            const T xi = x[i];
            for (MKL_INT p = rs[i]; p < re[i]; p++) {
                const MKL_INT j = ci[p];
                if (j < i) continue;
                acc += val[p] * x[j];
                if (j != i) w[j - lo] += (conj_half ? cj(val[p]) : val[p]) * xi;
            }
            w[i - lo] += acc;
        }
        #pragma omp barrier
        #pragma omp for schedule(static)
        for (MKL_INT i = 0; i < n; i++) {
            T acc = T(0.0);
            for (int s = 0; s < nt; s++) if (i >= bound[s]) acc += win[s][i - bound[s]];
            y[i] += alpha * acc;
        }
    }
    return SPARSE_STATUS_SUCCESS;
}

sparse_status_t mkl_sparse_d_mv(sparse_operation_t op, double alpha, const sparse_matrix_t A, struct matrix_descr descr,
                                const double *x, double beta, double *y)
{ return mv<double>(op, alpha, A, descr, x, beta, y); }

sparse_status_t mkl_sparse_z_mv(sparse_operation_t op, MKL_Complex16 alpha, const sparse_matrix_t A,
                                struct matrix_descr descr, const MKL_Complex16 *x, MKL_Complex16 beta,
                                MKL_Complex16 *y)
{ return mv<MKL_Complex16>(op, alpha, A, descr, x, beta, y); }

void feastinit(MKL_INT *fpm) { for (int i = 0; i < 128; i++) fpm[i] = 0; }
void zfeast_hcsrev(const char *, const MKL_INT *, const MKL_Complex16 *, const MKL_INT *, const MKL_INT *, MKL_INT *,
                   double *, MKL_INT *, const double *, const double *, MKL_INT *, double *, MKL_Complex16 *,
                   MKL_INT *m, double *, MKL_INT *info)
{ *m = 0; *info = 200; /* FEAST is a direct solver outside the H*v path: not provided by the shim */ }

void MKL_Get_Version(MKLVersion *v)
{ v->MajorVersion = 0; v->MinorVersion = 0; v->UpdateVersion = 0; v->ProductStatus = "shim (no MKL)";
  v->Build = "qb oracle shim"; v->Processor = "generic"; v->Platform = "OpenBLAS ILP64 + restated sparse mv"; }
int mkl_get_max_threads(void) { return g_spmv_threads; }
