/* oracle/qb_oracle.c -- TEST INFRASTRUCTURE ONLY (see qb_oracle.h for scope, citations and parity status). */
#include "qb_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

/* Symmetric tridiagonal eigen-decomposition by implicit QL with Wilkinson shifts (the textbook "tql2"
 * algorithm).  The reference calls LAPACKE_dstedc('I') here (src/lanczos.cc:367); LAPACK is a third-party
 * dependency, so the published QL algorithm is restated instead.  d[m] diagonal -> eigenvalues (unsorted),
 * e[m] sub-diagonal in e[0..m-1) (e[m-1] scratch), z = identity on entry (m x m, column-major) or NULL. */
static int tridiag_ql(int64_t m, double *d, double *e, double *z)
{
    for (int64_t l = 0; l < m; l++) {
        int iter = 0;
        int64_t mm;
        do {
            for (mm = l; mm < m - 1; mm++) {
                double dd = fabs(d[mm]) + fabs(d[mm + 1]);
                if (fabs(e[mm]) <= DBL_EPSILON * dd) break;
            }
            if (mm != l) {
                if (iter++ == 200) return 1;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = hypot(g, 1.0);
                g = d[mm] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int64_t i;
                for (i = mm - 1; i >= l; i--) {
                    double f = s * e[i], b = c * e[i];
                    e[i + 1] = (r = hypot(f, g));
                    if (r == 0.0) { d[i + 1] -= p; e[mm] = 0.0; break; }
                    s = f / r; c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    if (z) for (int64_t k = 0; k < m; k++) {
                        f = z[k + (i + 1) * m];
                        z[k + (i + 1) * m] = s * z[k + i * m] + c * f;
                        z[k + i * m]       = c * z[k + i * m] - s * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p; e[l] = g; e[mm] = 0.0;
            }
        } while (mm != l);
    }
    return 0;
}

int qbo_hess_eigen(const double *hess, int64_t maxit, int64_t m, double *ritz, double *s)
{
    double *d = (double *)malloc(sizeof(double) * (size_t)m);
    double *e = (double *)malloc(sizeof(double) * (size_t)m);
    double *z = s ? (double *)calloc((size_t)m * (size_t)m, sizeof(double)) : NULL;
    int64_t *ord = (int64_t *)malloc(sizeof(int64_t) * (size_t)m);
    for (int64_t j = 0; j < m; j++) { d[j] = hess[maxit + j]; e[j] = (j + 1 < m) ? hess[j + 1] : 0.0; ord[j] = j; }
    if (z) for (int64_t j = 0; j < m; j++) z[j + j * m] = 1.0;
    int info = tridiag_ql(m, d, e, z);
    /* ascending ("sr", src/lanczos.cc:375-377); insertion sort on an index keeps it simple */
    for (int64_t i = 1; i < m; i++) { int64_t k = ord[i]; int64_t j = i - 1; while (j >= 0 && d[ord[j]] > d[k]) { ord[j + 1] = ord[j]; j--; } ord[j + 1] = k; }
    for (int64_t j = 0; j < m; j++) {
        ritz[j] = d[ord[j]];
        if (s) memcpy(s + m * j, z + m * ord[j], sizeof(double) * (size_t)m);
    }
    free(d); free(e); free(z); free(ord);
    return info;
}

#define T double
#define SUF _d
#define CONJ(x) (x)
#define REAL(x) (x)
#define ABS2(x) ((x) * (x))
#include "qb_oracle_impl.inc"
#undef T
#undef SUF
#undef CONJ
#undef REAL
#undef ABS2

#define T double _Complex
#define SUF _z
#define CONJ(x) conj(x)
#define REAL(x) creal(x)
#define ABS2(x) (creal(x) * creal(x) + cimag(x) * cimag(x))
#include "qb_oracle_impl.inc"
#undef T
#undef SUF
#undef CONJ
#undef REAL
#undef ABS2

void qbo_spmv_z_ld(int64_t n, const int64_t *ia, const int64_t *ja, const double _Complex *val, int sym,
                   const double _Complex *x, double _Complex *y)
{
    long double _Complex *acc = (long double _Complex *)calloc((size_t)n, sizeof(long double _Complex));
    for (int64_t i = 0; i < n; i++)
        for (int64_t p = ia[i]; p < ia[i + 1]; p++) {
            const int64_t j = ja[p];
            if (sym && j < i) continue;
            acc[i] += (long double _Complex)val[p] * (long double _Complex)x[j];
            if (sym && j != i) acc[j] += (long double _Complex)conj(val[p]) * (long double _Complex)x[i];
        }
    for (int64_t i = 0; i < n; i++) y[i] = (double _Complex)acc[i];
    free(acc);
}

int64_t qbo_expand_upper_z(int64_t n, const int64_t *ia, const int64_t *ja, const double _Complex *val,
                           int64_t *ia_full, int64_t *ja_full, double _Complex *val_full)
{
    /* row i of the full matrix = conj-transposed entries (j<i, stored in row j) followed by the stored
       upper entries (j>=i); both runs are column-sorted when the input rows are. */
    int64_t *cnt = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
    for (int64_t i = 0; i < n; i++)
        for (int64_t p = ia[i]; p < ia[i + 1]; p++) {
            const int64_t j = ja[p];
            if (j < i) continue;
            cnt[i]++;
            if (j != i) cnt[j]++;
        }
    ia_full[0] = 0;
    for (int64_t i = 0; i < n; i++) ia_full[i + 1] = ia_full[i] + cnt[i];
    const int64_t nnz_full = ia_full[n];
    if (ja_full && val_full) {
        int64_t *pos = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
        for (int64_t i = 0; i < n; i++) pos[i] = ia_full[i];
        /* pass 1: transposed half, visiting source rows in increasing order keeps columns ascending */
        for (int64_t i = 0; i < n; i++)
            for (int64_t p = ia[i]; p < ia[i + 1]; p++) {
                const int64_t j = ja[p];
                if (j > i) { ja_full[pos[j]] = i; val_full[pos[j]] = conj(val[p]); pos[j]++; }
            }
        for (int64_t i = 0; i < n; i++)
            for (int64_t p = ia[i]; p < ia[i + 1]; p++) {
                const int64_t j = ja[p];
                if (j >= i) { ja_full[pos[i]] = j; val_full[pos[i]] = val[p]; pos[i]++; }
            }
        free(pos);
    }
    free(cnt);
    return nnz_full;
}

void qbo_energy_scale_z(int64_t n, const int64_t *ia, const int64_t *ja, const double _Complex *val, int sym,
                        double _Complex *v, double *lo, double *hi, double extend, int64_t iters, int nthreads)
{
    /* src/kpm.cc:45-88: iters-1 plain Lanczos steps, no stop rule, then the extreme Ritz values */
    const int64_t mm = iters - 1;
    double *hess = (double *)calloc((size_t)(2 * iters), sizeof(double));
    double *ritz = (double *)malloc(sizeof(double) * (size_t)mm);
    double _Complex *vp[2] = { v, v + n };
    qbo_spmv_z(n, ia, ja, val, sym, vp[0], vp[1], 0, nthreads);
    hess[iters] = creal(dotc_z(n, vp[0], vp[1]));
    axpy_z(n, -hess[iters], vp[0], vp[1]);
    hess[1] = nrm2_z(n, vp[1]);
    scal_z(n, 1.0 / hess[1], vp[1]);
    for (int64_t m = 2; m <= mm; m++) {
        double _Complex *vm = vp[m % 2], *vm1 = vp[(m - 1) % 2];
        for (int64_t l = 0; l < n; l++) vm[l] = -hess[m - 1] * vm[l];
        qbo_spmv_z(n, ia, ja, val, sym, vm1, vm, 1, nthreads);
        hess[iters + m - 1] = creal(dotc_z(n, vm1, vm));
        axpy_z(n, -hess[iters + m - 1], vm1, vm);
        hess[m] = nrm2_z(n, vm);
        scal_z(n, 1.0 / hess[m], vm);
    }
    qbo_hess_eigen(hess, iters, mm, ritz, NULL);
    double l = ritz[0], h = ritz[mm - 1], slack = extend * (h - l);
    *lo = l - slack; *hi = h + slack;
    free(hess); free(ritz);
}

void qbo_kpm_moments_z(int64_t n, const int64_t *ia, const int64_t *ja, const double _Complex *val, int sym,
                       const double _Complex *phi, double lo, double hi, int64_t nmom, double *mu, int nthreads)
{
    const double c = 0.5 * (hi + lo), s = 0.5 * (hi - lo);
    double _Complex *t0 = (double _Complex *)malloc(sizeof(double _Complex) * (size_t)n);
    double _Complex *t1 = (double _Complex *)malloc(sizeof(double _Complex) * (size_t)n);
    double _Complex *w  = (double _Complex *)malloc(sizeof(double _Complex) * (size_t)n);
    memcpy(t0, phi, sizeof(double _Complex) * (size_t)n);
    if (nmom > 0) mu[0] = creal(dotc_z(n, phi, t0));
    /* T1 = Ht phi */
    qbo_spmv_z(n, ia, ja, val, sym, t0, t1, 0, nthreads);
    for (int64_t i = 0; i < n; i++) t1[i] = (t1[i] - c * t0[i]) / s;
    if (nmom > 1) mu[1] = creal(dotc_z(n, phi, t1));
    for (int64_t k = 2; k < nmom; k++) {
        qbo_spmv_z(n, ia, ja, val, sym, t1, w, 0, nthreads);
        for (int64_t i = 0; i < n; i++) w[i] = 2.0 * (w[i] - c * t1[i]) / s - t0[i];
        double _Complex *tmp = t0; t0 = t1; t1 = w; w = tmp;
        mu[k] = creal(dotc_z(n, phi, t1));
    }
    free(t0); free(t1); free(w);
}

/* Same as qbo_eigenvec_cg_z with E0 behind a pointer (re, im): complex scalars by value do not cross every FFI. */
int64_t qbo_eigenvec_cg_zp(int64_t n, const int64_t *ia, const int64_t *ja, const double _Complex *val, int sym,
                           const double *E0_reim, int64_t maxit, double *accu, double _Complex *v, double _Complex *r,
                           double _Complex *p, double _Complex *pp, int nthreads)
{
    return qbo_eigenvec_cg_z(n, ia, ja, val, sym, E0_reim[0] + E0_reim[1] * I, maxit, accu, v, r, p, pp, nthreads);
}
