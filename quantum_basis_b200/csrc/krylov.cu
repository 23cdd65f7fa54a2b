// quantum_basis_b200/csrc/krylov.cu -- device-resident Krylov loops behind the reference's lanczos(), eigenvec_CG()
// and energy_scale() signatures, plus Chebyshev/KPM moments (new functionality).
//
// Reference loops: src/lanczos.cc:134-266 (lanczos), :281-341 (eigenvec_CG), :355-390 (hess_eigen),
// src/kpm.cc:45-88 (energy_scale).  What changes is the data flow, not the mathematics: vectors stay in HBM, each
// step is 2-3 fused kernels (one pass over H, one over each vector), scalars chain through device memory, and the
// host only sees (a_m, b_m) to evaluate the reference's stop rule.  The checkpoint hooks and the per-step log
// files of the reference (ckpt_lanczos_update, log_Lanczos_srval, log_CG.txt) are host I/O outside the path.
#include "internal.hpp"
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace qb {

constexpr double kLanczosPrecision = 2e-12;   // reference src/miscellaneous.cc:47

// ---------------------------------------------------------------------------------------------- hess_eigen
// Implicit QL with Wilkinson shifts for the symmetric tridiagonal (diag d, off-diag e).  The reference calls
// LAPACKE_dstedc('I') (src/lanczos.cc:367).  z: full eigenvector matrix (m x m column-major, identity on entry),
// or null; zlast: only the last row of the eigenvector matrix (e_m^T Q), or null -- the stop rule needs
// s[m-1] of the lowest Ritz vector only (src/lanczos.cc:231), which keeps the per-step cost O(m^2).
static int tridiag_ql(int64_t m, double *d, double *e, double *z, double *zlast)
{
    for (int64_t l = 0; l < m; l++) {
        int iter = 0;
        int64_t mm;
        do {
            for (mm = l; mm < m - 1; mm++) {
                const double dd = fabs(d[mm]) + fabs(d[mm + 1]);
                if (fabs(e[mm]) <= DBL_EPSILON * dd) break;
            }
            if (mm != l) {
                if (iter++ == 300) return 1;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = hypot(g, 1.0);
                g = d[mm] - d[l] + e[l] / (g + copysign(r, g));
                double s = 1.0, c = 1.0, p = 0.0;
                int64_t i;
                bool underflow = false;
                for (i = mm - 1; i >= l; i--) {
                    double f = s * e[i];
                    const double b = c * e[i];
                    r = hypot(f, g);
                    e[i + 1] = r;
                    if (r == 0.0) { d[i + 1] -= p; e[mm] = 0.0; underflow = true; break; }
                    s = f / r; c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    p = s * r;
                    d[i + 1] = g + p;
                    g = c * r - b;
                    if (z) for (int64_t k = 0; k < m; k++) {
                        f = z[k + (i + 1) * m];
                        z[k + (i + 1) * m] = s * z[k + i * m] + c * f;
                        z[k + i * m] = c * z[k + i * m] - s * f;
                    }
                    if (zlast) { f = zlast[i + 1]; zlast[i + 1] = s * zlast[i] + c * f; zlast[i] = c * zlast[i] - s * f; }
                }
                if (underflow) continue;
                d[l] -= p; e[l] = g; e[mm] = 0.0;
            }
        } while (mm != l);
    }
    return 0;
}

// Smallest eigenvalue of the m x m tridiagonal (diag hess[maxit+j], off-diag hess[j+1]) by bisection on the Sturm
// count, to machine precision.  O(m) per evaluation: the per-step part of the stop rule (the relative change of the
// lowest Ritz value, src/lanczos.cc:232-239) needs nothing else; the O(m^2) QL solve for the residual estimate
// |b_m s_{m-1}| (:231) is only run once that change has stayed below the threshold for more than 15 steps.
double tridiag_smallest(const double *hess, int64_t maxit, int64_t m)
{
    const double *a = hess + maxit, *b = hess;             // b[j] couples j-1 and j (j = 1..m-1)
    double lo = a[0], hi = a[0];
    for (int64_t j = 0; j < m; j++) {
        const double r = (j > 0 ? fabs(b[j]) : 0.0) + (j + 1 < m ? fabs(b[j + 1]) : 0.0);
        lo = fmin(lo, a[j] - r); hi = fmax(hi, a[j] + r);
    }
    auto count_below = [&](double x) {                     // number of eigenvalues < x
        int64_t cnt = 0;
        double q = 1.0;
        for (int64_t j = 0; j < m; j++) {
            const double off2 = j > 0 ? b[j] * b[j] : 0.0;
            q = a[j] - x - (j > 0 ? off2 / q : 0.0);
            if (q == 0.0) q = -DBL_MIN;
            if (q < 0.0) cnt++;
        }
        return cnt;
    };
    hi = fmin(hi, a[0]);                                   // the smallest eigenvalue is <= any diagonal entry
    for (int it = 0; it < 200; it++) {
        const double mid = 0.5 * (lo + hi);
        if (mid <= lo || mid >= hi) break;
        if (count_below(mid) >= 1) hi = mid; else lo = mid;
    }
    return 0.5 * (lo + hi);
}

// ritz ascending ("sr"); s (optional) full vectors; s_last0 (optional) = last component of the lowest vector
int hess_eigen_host(const double *hess, int64_t maxit, int64_t m, double *ritz, double *s, double *s_last0)
{
    std::vector<double> d(m), e(m), z, zl;
    std::vector<int64_t> ord(m);
    for (int64_t j = 0; j < m; j++) { d[j] = hess[maxit + j]; e[j] = (j + 1 < m) ? hess[j + 1] : 0.0; ord[j] = j; }
    if (s) { z.assign((size_t)m * m, 0.0); for (int64_t j = 0; j < m; j++) z[j + j * m] = 1.0; }
    if (s_last0) { zl.assign(m, 0.0); zl[m - 1] = 1.0; }
    if (tridiag_ql(m, d.data(), e.data(), s ? z.data() : nullptr, s_last0 ? zl.data() : nullptr)) return 1;
    for (int64_t i = 1; i < m; i++) { const int64_t k = ord[i]; int64_t j = i - 1; while (j >= 0 && d[ord[j]] > d[k]) { ord[j + 1] = ord[j]; j--; } ord[j + 1] = k; }
    for (int64_t j = 0; j < m; j++) {
        ritz[j] = d[ord[j]];
        if (s) memcpy(s + m * j, z.data() + m * ord[j], sizeof(double) * (size_t)m);
    }
    if (s_last0) *s_last0 = zl[ord[0]];
    return 0;
}

// ------------------------------------------------------------------------------------------------- helpers
struct DevBuf {             // RAII-ish device scratch
    void *p = nullptr;
    int alloc(size_t bytes) { cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1); if (e != cudaSuccess) { (void)cudaGetLastError(); p = nullptr; return fail(QBGPU_ERR_ALLOC, "cudaMalloc failed in a Krylov driver"); } return 0; }
    void release() { if (p) cudaFree(p); p = nullptr; }
    ~DevBuf() { release(); }
};

static int check_single(const qbgpu_matrix *A, bool cplx)
{
    if (!A) return fail(QBGPU_ERR_ARG, "null matrix handle");
    if (A->api_complex != cplx) return fail(QBGPU_ERR_ARG, "handle scalar type does not match this entry point");
    if (A->row_lo != 0 || A->row_hi != A->n) return fail(QBGPU_ERR_STATE, "this loop needs an unsharded handle (use quantum_basis_b200.dist for shards)");
    return QBGPU_OK;
}

// ------------------------------------------------------------------------------------------------- lanczos
// state layout: see include/qbgpu.h (lanczos_step_*).
// k > 0 resumes (src/lanczos.cc:144-148, qbasis.h:1044-1061): v holds the normalised v[k-1], v[k] in their slots, hess
// holds a[0..k-1] and b[0..k].  stop_state (optional, in/out) = {cnt_accuE0, accuracy, theta0_prev, theta1_prev}: what the
// reference's checkpoint carries across an interruption (ckpt.cc: lczs_mlns.dat); without it a resumed run restarts the
// stop rule's counters, like the reference's lanczos() called with k > 0 and checkpoints disabled.
static int lanczos_impl(qbgpu_matrix *A, bool cplx, int64_t k, int64_t np, int64_t maxit, int64_t *m_out, void *v,
                        double *hess, const char *purpose, int where, bool stop_on_breakdown = true, double *stop_state = nullptr)
{
    QB_TRY(ensure_init());
    QB_TRY(check_single(A, cplx));
    if (!m_out || !v || !hess || !purpose) return fail(QBGPU_ERR_ARG, "lanczos: null argument");
    const bool is_val = strstr(purpose, "val") != nullptr;
    const bool is_val1 = strstr(purpose, "val1") != nullptr;
    const bool is_dn = strcmp(purpose, "dnmcs") == 0;
    if (!is_val && !is_dn) return fail(QBGPU_ERR_ARG, "lanczos: purpose must be sr_val0, sr_val1 or dnmcs");
    const int64_t mm = k + np;
    if (!(mm < maxit && k >= 0 && np >= 0)) return fail(QBGPU_ERR_ARG, "lanczos: need k >= 0, np >= 0 and k + np < maxit");   // src/lanczos.cc:147
    *m_out = k;
    if (stop_state && is_val && (int)stop_state[0] > 15 && stop_state[1] < kLanczosPrecision) return QBGPU_OK;   // :149 (already converged)
    if (np == 0) return QBGPU_OK;                                                                  // :150
    Context &c = ctx();
    const int64_t n = A->n;
    const size_t vb = cplx ? 16 : 8;
    const int nvec = is_val1 ? 3 : 2;

    DevBuf dv, dstate, dhess;
    char *U;
    if (where == QBGPU_HOST) {
        QB_TRY(dv.alloc(vb * n * nvec));
        U = (char *)dv.p;
        QB_CUDA(cudaMemcpyAsync(U, v, vb * n * (k > 0 ? 2 : 1), cudaMemcpyHostToDevice, c.stream));    // a resumed run needs both live vectors
        if (is_val1) QB_CUDA(cudaMemcpyAsync(U + 2 * vb * n, (char *)v + 2 * vb * n, vb * n, cudaMemcpyHostToDevice, c.stream));
    } else if (where == QBGPU_DEVICE) {
        U = (char *)v;
    } else return fail(QBGPU_ERR_ARG, "where must be QBGPU_HOST or QBGPU_DEVICE");
    char *Ubuf[2] = {U, U + vb * n};
    char *phi = U + 2 * vb * n;
    QB_TRY(dstate.alloc(sizeof(double) * 8));
    QB_TRY(dhess.alloc(sizeof(double) * 2 * maxit));
    double *state = (double *)dstate.p;
    double *b_dev = (double *)dhess.p, *a_dev = b_dev + maxit;
    // state = {scale of ux, scale of uz, b_prev, ...}: both live vectors of a resumed run are normalised and b_prev = b[k]
    const double init[8] = {1.0, k > 0 ? 1.0 : 0.0, k > 0 ? hess[k] : 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    QB_CUDA(cudaMemcpyAsync(state, init, sizeof init, cudaMemcpyHostToDevice, c.stream));
    QB_CUDA(cudaMemsetAsync(dhess.p, 0, sizeof(double) * 2 * maxit, c.stream));
    QB_CUDA(cudaStreamSynchronize(c.stream));              // `init` is on the host stack
    // the caller's array as the reference leaves it: zeros beyond what is known (b[0..k], a[0..k-1] are kept on a resume)
    for (int64_t j = (k > 0 ? k + 1 : 0); j < maxit; j++) hess[j] = 0.0;
    for (int64_t j = maxit + k; j < 2 * maxit; j++) hess[j] = 0.0;

    std::vector<double> ritz(mm + 1);
    int cnt_accuE0 = stop_state ? (int)stop_state[0] : 0;
    double accuracy = stop_state ? stop_state[1] : 0.0;
    double theta0_prev = stop_state ? stop_state[2] : (k > 0 && is_val ? tridiag_smallest(hess, maxit, k) : 0.0);
    bool converged = false;
    int64_t m = k;
    // QBGPU_VERBOSE: per-phase device time of the loop (events around the three passes) and host time of the stop rule
    const bool prof = getenv("QBGPU_VERBOSE") != nullptr;
    EventGuard<4> pg;                                       // (destroyed on every way out of the loop, error returns included)
    double t_a = 0, t_b = 0, t_c = 0, t_host = 0;
    if (prof) for (int i = 0; i < 4; i++) QB_CUDA(pg.create());
    cudaEvent_t *pe = pg.e;
    while (m < mm) {
        m++;
        // one fused Lanczos step (src/lanczos.cc:167-187 for m == 1, :194-214 otherwise)
        const void *ux = Ubuf[(m - 1) % 2];
        void *uz = Ubuf[m % 2];
        if (prof) QB_CUDA(cudaEventRecord(pe[0], c.stream));
        QB_TRY(lanczos_step_a(A, ux, uz, state, true, true));
        if (prof) QB_CUDA(cudaEventRecord(pe[1], c.stream));
        QB_TRY(lanczos_step_b(n, cplx, ux, uz, state));
        if (prof) QB_CUDA(cudaEventRecord(pe[2], c.stream));
        QB_TRY(lanczos_step_c(state, a_dev, b_dev, m));
        if (prof) QB_CUDA(cudaEventRecord(pe[3], c.stream));
        double ab[2];
        QB_CUDA(cudaMemcpyAsync(c.scal_host, a_dev + (m - 1), sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        QB_CUDA(cudaMemcpyAsync(c.scal_host + 1, b_dev + m, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        QB_CUDA(cudaStreamSynchronize(c.stream));
        ab[0] = c.scal_host[0]; ab[1] = c.scal_host[1];
        hess[maxit + m - 1] = ab[0];
        hess[m] = ab[1];
        if (prof) {
            float f;
            cudaEventElapsedTime(&f, pe[0], pe[1]); t_a += f;
            cudaEventElapsedTime(&f, pe[1], pe[2]); t_b += f;
            cudaEventElapsedTime(&f, pe[2], pe[3]); t_c += f;
        }
        const auto th0 = std::chrono::steady_clock::now();
        struct HostTimer { const std::chrono::steady_clock::time_point t0; double &acc; ~HostTimer() { acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); } } host_timer{th0, t_host};
        if (m == 1) continue;                               // the start-up step has no checks in the reference (:167-191)
        if (stop_on_breakdown && fabs(hess[m]) < kLanczosPrecision) break;                         // :216
        if (is_val1) {                                      // re-orthogonalise against phi0, :218-226
            double *dd = c.scal_dev + 52;
            QB_TRY(vec_dotc(n, cplx, phi, uz, dd));
            double h[3];
            QB_TRY(read_scalars(dd, h, 2));
            double sx;
            QB_TRY(read_scalars(state, &sx, 1));
            const double tr = h[0] * sx, ti = cplx ? h[1] * sx : 0.0;
            if (sqrt(tr * tr + ti * ti) > kLanczosPrecision) {
                // v_m = sx*U ;  v_m -= tmp*phi  <=>  U -= (tmp/sx)*phi ;  then renormalise: sx <- 1/||U||
                QB_TRY(vec_axpy(n, cplx, make_double2(-h[0], cplx ? -h[1] : 0.0), phi, uz));
                QB_TRY(vec_nrm2sq(n, cplx, uz, dd));
                double nn;
                QB_TRY(read_scalars(dd, &nn, 1));
                const double nsx = 1.0 / sqrt(nn);
                QB_CUDA(cudaMemcpyAsync(state, &nsx, sizeof(double), cudaMemcpyHostToDevice, c.stream));
                QB_CUDA(cudaStreamSynchronize(c.stream));
            }
        }
        if (is_val) {                                       // stop rule, :228-248
            const double theta0 = tridiag_smallest(hess, maxit, m);
            if (m > 3) {
                const double accu_E0 = fabs((theta0 - theta0_prev) / theta0);
                if (accu_E0 < kLanczosPrecision) cnt_accuE0++; else cnt_accuE0 = 0;
                if (cnt_accuE0 > 15) {                      // only now is the residual estimate |b_m s_{m-1}| needed
                    double s_last = 0.0;
                    if (hess_eigen_host(hess, maxit, m, ritz.data(), nullptr, &s_last)) return fail(QBGPU_ERR_NUMERIC, "hess_eigen: QL did not converge");
                    accuracy = fabs(hess[m] * s_last);
                    if (accuracy < kLanczosPrecision) { converged = true; break; }
                }
            }
            theta0_prev = theta0;
        }
    }
    if (prof) {
        fprintf(stderr, "[qbgpu lanczos] %lld steps, per step: product+epilogue %.3f ms, update pass %.3f ms, scalar kernel %.4f ms, host stop rule %.3f ms (n=%lld, %s vectors)\n",
                (long long)m, t_a / m, t_b / m, t_c / m, t_host / m, (long long)n, cplx ? "complex" : "fp64");
    }
    *m_out = m;
    if (stop_state && is_val && m > k) {
        // what ckpt_lanczos_update would save at this step (src/lanczos.cc:231-248): the residual estimate of step m and,
        // unless the run stopped on the rule (the reference breaks before updating them), this step's Ritz values
        double s_last = 0.0;
        std::vector<double> rz(m + 1);
        if (hess_eigen_host(hess, maxit, m, rz.data(), nullptr, &s_last)) return fail(QBGPU_ERR_NUMERIC, "hess_eigen: QL did not converge");
        stop_state[0] = (double)cnt_accuE0;
        if (m > 3) stop_state[1] = fabs(hess[m] * s_last);
        if (!converged) { stop_state[2] = rz[0]; stop_state[3] = m > 1 ? rz[1] : 0.0; }
        (void)accuracy;
    }
    // hand the two live vectors back normalised, in the reference's slots: v_m at (m%2), v_{m-1} at ((m-1)%2)
    QB_TRY(scale_copy(n, cplx, state + 0, 1.0, Ubuf[m % 2], Ubuf[m % 2]));
    QB_TRY(scale_copy(n, cplx, state + 1, 1.0, Ubuf[(m - 1) % 2], Ubuf[(m - 1) % 2]));
    if (where == QBGPU_HOST) QB_CUDA(cudaMemcpyAsync(v, U, vb * n * 2, cudaMemcpyDeviceToHost, c.stream));
    QB_CUDA(cudaStreamSynchronize(c.stream));
    return QBGPU_OK;
}

// --------------------------------------------------------------------------------------------- eigenvec_CG
// The two branches of the reference's loop body (src/lanczos.cc:295-331) on device vectors; sc: 8 device doubles
// [0]=gamma=|r| [1,2]=delta [3]=|pp|^2 [4]=|r_new|^2 [5]=gamma_next [6]=scratch.  Shared by the whole loop (cg_impl) and by the
// step-level entry points qbgpu_cg_restart_* / qbgpu_cg_step_*.
// :297-314  v <- v / rnorm ; r = (E0 - H) v in one pass, |r|^2 from the epilogue ; p = r ; accu = |r| (left in sc[0] as gamma)
static int cg_restart_piece(qbgpu_matrix *A, bool cplx, double2 E0, double rnorm, double *sc, char *dv, char *dr, char *dp, double *accu)
{
    Context &c = ctx();
    const int64_t n = A->n;
    const size_t vb = cplx ? 16 : 8;
    QB_TRY(vec_scal(n, cplx, make_double2(1.0 / rnorm, 0.0), dv));
    FusedArgs fa;
    fa.x = dv; fa.y = dr; fa.alpha = make_double2(-1.0, 0.0); fa.gamma = E0; fa.dots = sc + 1;
    QB_TRY(launch_spmv(A, fa));
    QB_CUDA(cudaMemcpyAsync(dp, dr, vb * n, cudaMemcpyDeviceToDevice, c.stream));   // p = r
    double h[3];
    QB_TRY(read_scalars(sc + 1, h, 3));
    *accu = sqrt(h[2]);
    QB_CUDA(cudaMemcpyAsync(sc, accu, sizeof(double), cudaMemcpyHostToDevice, c.stream));
    QB_CUDA(cudaStreamSynchronize(c.stream));
    return QBGPU_OK;
}
// :320-330  pp = (H - E0 + eps) p ; delta = <p,pp> ; v += alpha p ; r -= alpha pp ; p = r + beta p ; accu = |r_new|
static int cg_iterate_piece(qbgpu_matrix *A, bool cplx, double2 E0, double *sc, char *dv, char *dr, char *dp, char *dpp, double *accu)
{
    Context &c = ctx();
    const int64_t n = A->n;
    FusedArgs fa;
    fa.x = dp; fa.y = dpp; fa.alpha = make_double2(1.0, 0.0);
    fa.gamma = make_double2(DBL_EPSILON - E0.x, -E0.y); fa.dots = sc + 1;
    QB_TRY(launch_spmv(A, fa));
    QB_TRY(cg_update_vr(n, cplx, sc, dv, dr, dp, dpp));          // :324-327
    QB_TRY(cg_update_p(n, cplx, sc, dr, dp));                    // :327-330
    QB_CUDA(cudaMemcpyAsync(sc, sc + 5, sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
    return read_scalars(sc, accu, 1);
}

static int cg_impl(qbgpu_matrix *A, bool cplx, int64_t maxit, int64_t *m_io, double2 E0, double *accu_out,
                   void *v, void *r, void *p, void *pp, int where)
{
    QB_TRY(ensure_init());
    QB_TRY(check_single(A, cplx));
    if (!m_io || !accu_out || !v || !r || !p || !pp) return fail(QBGPU_ERR_ARG, "eigenvec_CG: null argument");
    if (maxit <= 0 || *m_io < 0 || *m_io >= maxit) return fail(QBGPU_ERR_ARG, "eigenvec_CG: need 0 <= m < maxit");        // src/lanczos.cc:287
    Context &c = ctx();
    const int64_t n = A->n;
    const size_t vb = cplx ? 16 : 8;
    DevBuf buf, dsc;
    char *dv, *dr, *dp, *dpp;
    if (where == QBGPU_HOST) {
        QB_TRY(buf.alloc(vb * n * 4));
        dv = (char *)buf.p; dr = dv + vb * n; dp = dr + vb * n; dpp = dp + vb * n;
        QB_CUDA(cudaMemcpyAsync(dv, v, vb * n, cudaMemcpyHostToDevice, c.stream));
        if (*m_io > 0) {                                    // a resumed run continues from (v, r, p) of step m
            QB_CUDA(cudaMemcpyAsync(dr, r, vb * n, cudaMemcpyHostToDevice, c.stream));
            QB_CUDA(cudaMemcpyAsync(dp, p, vb * n, cudaMemcpyHostToDevice, c.stream));
        }
    } else if (where == QBGPU_DEVICE) {
        dv = (char *)v; dr = (char *)r; dp = (char *)p; dpp = (char *)pp;
    } else return fail(QBGPU_ERR_ARG, "where must be QBGPU_HOST or QBGPU_DEVICE");
    QB_TRY(dsc.alloc(sizeof(double) * 8));
    double *sc = (double *)dsc.p;                           // [0]=gamma [1,2]=delta [3]=|pp|^2 [4]=|r|^2 [5]=gamma_next
    QB_CUDA(cudaMemsetAsync(sc, 0, sizeof(double) * 8, c.stream));

    int64_t m = *m_io;
    double accu = 0.0;                                      // src/lanczos.cc:288-292: 0 at m == 0, |r| on a resume
    if (m > 0) {
        double nn;
        QB_TRY(vec_nrm2sq(n, cplx, dr, sc + 6));
        QB_TRY(read_scalars(sc + 6, &nn, 1));
        accu = sqrt(nn);
        QB_CUDA(cudaMemcpyAsync(sc, &accu, sizeof(double), cudaMemcpyHostToDevice, c.stream));
        QB_CUDA(cudaStreamSynchronize(c.stream));
    }
    while (m < maxit) {
        if (accu < kLanczosPrecision) {                     // :295
            double nn;
            QB_TRY(vec_nrm2sq(n, cplx, dv, sc + 6));
            QB_TRY(read_scalars(sc + 6, &nn, 1));
            const double rnorm = sqrt(nn);
            if (m == 0 || fabs(rnorm - 1.0) > kLanczosPrecision) {       // :297 re-normalise and restart
                QB_TRY(cg_restart_piece(A, cplx, E0, rnorm, sc, dv, dr, dp, &accu));
                m++;
                if (accu < kLanczosPrecision) break;        // :315
            } else {
                break;                                      // :317
            }
        } else {
            QB_TRY(cg_iterate_piece(A, cplx, E0, sc, dv, dr, dp, dpp, &accu));     // :320-330
            m++;
        }
    }
    *m_io = m;
    *accu_out = accu;
    if (where == QBGPU_HOST) {
        QB_CUDA(cudaMemcpyAsync(v, dv, vb * n, cudaMemcpyDeviceToHost, c.stream));
        QB_CUDA(cudaMemcpyAsync(r, dr, vb * n, cudaMemcpyDeviceToHost, c.stream));
        QB_CUDA(cudaMemcpyAsync(p, dp, vb * n, cudaMemcpyDeviceToHost, c.stream));
        QB_CUDA(cudaMemcpyAsync(pp, dpp, vb * n, cudaMemcpyDeviceToHost, c.stream));
    }
    QB_CUDA(cudaStreamSynchronize(c.stream));
    return QBGPU_OK;
}

// -------------------------------------------------------------------------------------------- energy_scale
static int energy_scale_impl(qbgpu_matrix *A, bool cplx, void *v, double *lo, double *hi, double extend, int64_t iters, int where)
{
    QB_TRY(ensure_init());
    QB_TRY(check_single(A, cplx));
    if (!v || !lo || !hi || iters < 3) return fail(QBGPU_ERR_ARG, "energy_scale: bad argument");
    Context &c = ctx();
    const int64_t n = A->n;
    const size_t vb = cplx ? 16 : 8;
    const int64_t mm = iters - 1;                          // src/kpm.cc:53
    // the reference draws vec_randomize(dim, vpt[0]) itself (default seed 1, src/kpm.cc:60)
    DevBuf dv;
    void *U = v;
    if (where == QBGPU_HOST) { QB_TRY(dv.alloc(vb * n * 2)); U = dv.p; }
    QB_TRY(vec_randomize(n, cplx, U, 1));
    std::vector<double> hess(2 * iters, 0.0), ritz(mm);
    int64_t m = 0;
    QB_TRY(lanczos_impl(A, cplx, 0, mm, iters, &m, U, hess.data(), "dnmcs", QBGPU_DEVICE, /*stop_on_breakdown=*/false));
    if (hess_eigen_host(hess.data(), iters, mm, ritz.data(), nullptr, nullptr)) return fail(QBGPU_ERR_NUMERIC, "hess_eigen: QL did not converge");
    const double l = ritz[0], h = ritz[mm - 1], slack = extend * (h - l);                          // :83-87
    *lo = l - slack; *hi = h + slack;
    if (where == QBGPU_HOST) { QB_CUDA(cudaMemcpyAsync(v, U, vb * n * 2, cudaMemcpyDeviceToHost, c.stream)); QB_CUDA(cudaStreamSynchronize(c.stream)); }
    return QBGPU_OK;
}

// --------------------------------------------------------------------------------------------- KPM moments
// mu_k = <phi| T_k(Ht) |phi>, Ht = (H - c)/s.  One fused product per TWO moments:
//   T_{k+1} = 2 Ht T_k - T_{k-1}   (alpha = 2/s, gamma = -2c/s, beta = -1, written over T_{k-1})
//   mu_{2k+1} = 2 <T_k, T_{k+1}> - mu_1 ,  mu_{2k+2} = 2 <T_{k+1}, T_{k+1}> - mu_0
// both inner products come out of the product's epilogue into a device array; the host reads them once at the end.
static int kpm_impl(qbgpu_matrix *A, bool cplx, const void *phi, double lo, double hi, int64_t nmom, double *mu, int where)
{
    QB_TRY(ensure_init());
    QB_TRY(check_single(A, cplx));
    if (!phi || !mu || nmom < 1 || !(hi > lo)) return fail(QBGPU_ERR_ARG, "kpm_moments: bad argument");
    Context &c = ctx();
    const int64_t n = A->n;
    const size_t vb = cplx ? 16 : 8;
    const double cc = 0.5 * (hi + lo), ss = 0.5 * (hi - lo);
    DevBuf t, ddots;
    QB_TRY(t.alloc(vb * n * 2));
    char *T[2] = {(char *)t.p, (char *)t.p + vb * n};
    const int64_t nprod = nmom / 2 + 1;                    // products needed to reach index nmom-1
    QB_TRY(ddots.alloc(sizeof(double) * 4 * (nprod + 1)));
    double *dots = (double *)ddots.p;
    QB_CUDA(cudaMemcpyAsync(T[0], phi, vb * n, where == QBGPU_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, c.stream));
    QB_TRY(vec_nrm2sq(n, cplx, T[0], dots));               // mu_0 = <phi,phi>
    for (int64_t k = 0; k < nprod; k++) {
        FusedArgs fa;
        fa.x = T[k % 2]; fa.y = T[(k + 1) % 2]; fa.dots = dots + 4 * (k + 1);
        if (k == 0) { fa.alpha = make_double2(1.0 / ss, 0.0); fa.gamma = make_double2(-cc / ss, 0.0); }
        else { fa.alpha = make_double2(2.0 / ss, 0.0); fa.gamma = make_double2(-2.0 * cc / ss, 0.0); fa.beta = make_double2(-1.0, 0.0); fa.z = fa.y; }
        QB_TRY(launch_spmv(A, fa));
    }
    std::vector<double> h(4 * (nprod + 1));
    QB_CUDA(cudaMemcpyAsync(h.data(), dots, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, c.stream));
    QB_CUDA(cudaStreamSynchronize(c.stream));
    const double mu0 = h[0];
    const double mu1 = h[4 + 0];                           // <T0,T1>
    for (int64_t j = 0; j < nmom; j++) {
        if (j == 0) mu[j] = mu0;
        else if (j == 1) mu[j] = mu1;
        else if (j % 2 == 0) { const int64_t k = j / 2 - 1; mu[j] = 2.0 * h[4 * (k + 1) + 2] - mu0; }      // 2<T_{k+1},T_{k+1}> - mu0
        else { const int64_t k = (j - 1) / 2; mu[j] = 2.0 * h[4 * (k + 1) + 0] - mu1; }                    // 2<T_k,T_{k+1}> - mu1
    }
    return QBGPU_OK;
}


// ------------------------------------------------------------------------------------------ real-mode dispatch
// Through the reference's public API every vector is complex<double> (only model<complex> is instantiated, SURVEY
// F3), but when the stored values are real (all imaginary parts exactly 0) and the start vector is real -- which
// vec_randomize always is (src/miscellaneous.cc:382: imag = 0) -- the whole Krylov recurrence stays real.  The loop
// is then run on fp64 vectors (half the vector bytes, half the gather sectors) and the results are widened at the
// end.  Row sums go through the identical fma sequence; (a_m, b_m) agree with the complex run to round-off (only the
// grouping of the block-level reductions differs).
static bool real_mode_enabled(const qbgpu_matrix *A, bool cplx) { return cplx && A->val_real && !getenv("QBGPU_NO_REAL_MODE"); }

static int is_all_real(int64_t n, const void *dvec, bool *out)
{
    double *t = ctx().scal_dev + 56;
    QB_TRY(vec_imag_norm2(n, dvec, t));
    double h;
    QB_TRY(read_scalars(t, &h, 1));
    *out = (h == 0.0);
    return QBGPU_OK;
}

// ------------------------------------------------------------------------------- species-order handles (species.cu)
// The loops above are order-agnostic: they see vectors in whatever order launch_spmv multiplies in.  For a handle with an
// internal order the drivers below keep the reference's order at the boundary: the caller's vectors are permuted into
// temporaries on the way in and back on the way out -- the same two passes the real-mode dispatch spends on
// complex -> fp64 -> complex, with which the permutation is fused.  (a_m, b_m), E0, the KPM moments and the spectral
// bounds do not depend on the order; vectors come back in the reference's order.
struct NativeRun {
    qbgpu_matrix R;            // the handle as the loop sees it (fp64 vectors in real mode)
    bool loop_cplx = false;    // scalar type of the loop's vectors
    DevBuf stage, work;        // device copy of host vectors (reference order); the loop's vectors (internal order)
    char *ref = nullptr;       // the caller's vectors on the device, reference order
    size_t vb_api = 8, vb_loop = 8;
};

// in_slots: which of the caller's `nvec` vectors are inputs (bit j = vector j); real mode needs all of them real
static int native_begin(NativeRun &N, qbgpu_matrix *A, bool cplx, void *v, int nvec, unsigned in_slots, int where, bool allow_real)
{
    Context &c = ctx();
    const int64_t n = A->n;
    N.vb_api = cplx ? 16 : 8;
    N.ref = (char *)v;
    if (where == QBGPU_HOST) {
        QB_TRY(N.stage.alloc(N.vb_api * n * nvec));
        N.ref = (char *)N.stage.p;
        for (int j = 0; j < nvec; j++)
            if (in_slots & (1u << j)) QB_CUDA(cudaMemcpyAsync(N.ref + N.vb_api * n * j, (const char *)v + N.vb_api * n * j, N.vb_api * n, cudaMemcpyHostToDevice, c.stream));
    } else if (where != QBGPU_DEVICE) return fail(QBGPU_ERR_ARG, "where must be QBGPU_HOST or QBGPU_DEVICE");
    bool real = allow_real && real_mode_enabled(A, cplx);
    for (int j = 0; real && j < nvec; j++)
        if (in_slots & (1u << j)) { bool r = false; QB_TRY(is_all_real(n, N.ref + N.vb_api * n * j, &r)); real = real && r; }
    N.loop_cplx = cplx && !real;
    N.vb_loop = N.loop_cplx ? 16 : 8;
    N.R = *A;
    N.R.api_complex = N.loop_cplx;
    QB_TRY(N.work.alloc(N.vb_loop * n * nvec));
    QB_CUDA(cudaMemsetAsync(N.work.p, 0, N.vb_loop * n * nvec, c.stream));
    for (int j = 0; j < nvec; j++)
        if (in_slots & (1u << j)) QB_TRY(vec_to_native(A, cplx, N.loop_cplx, N.ref + N.vb_api * n * j, (char *)N.work.p + N.vb_loop * n * j));
    return QBGPU_OK;
}

// out_slots: vectors handed back to the caller (internal order -> reference order, widened to the API's scalar type)
static int native_end(NativeRun &N, qbgpu_matrix *A, bool cplx, void *v, unsigned out_slots, int nvec, int where)
{
    Context &c = ctx();
    const int64_t n = A->n;
    for (int j = 0; j < nvec; j++) {
        if (!(out_slots & (1u << j))) continue;
        QB_TRY(vec_from_native(A, N.loop_cplx, cplx, (char *)N.work.p + N.vb_loop * n * j, N.ref + N.vb_api * n * j, make_double2(1.0, 0.0), make_double2(0.0, 0.0)));
        if (where == QBGPU_HOST) QB_CUDA(cudaMemcpyAsync((char *)v + N.vb_api * n * j, N.ref + N.vb_api * n * j, N.vb_api * n, cudaMemcpyDeviceToHost, c.stream));
    }
    QB_CUDA(cudaStreamSynchronize(c.stream));
    return QBGPU_OK;
}

static int lanczos_native(qbgpu_matrix *A, bool cplx, int64_t k, int64_t np, int64_t maxit, int64_t *m_out, void *v,
                          double *hess, const char *purpose, int where, double *stop_state)
{
    QB_TRY(ensure_init());
    QB_TRY(check_single(A, cplx));
    if (!m_out || !v || !hess || !purpose) return fail(QBGPU_ERR_ARG, "lanczos: null argument");
    if (np <= 0) return lanczos_impl(A, cplx, k, np, maxit, m_out, v, hess, purpose, where, true, stop_state);   // nothing to do / argument errors
    const bool val1 = strstr(purpose, "val1") != nullptr;
    const int nvec = val1 ? 3 : 2;
    NativeRun N;
    QB_TRY(native_begin(N, A, cplx, v, nvec, (val1 ? 0x5u : 0x1u) | (k > 0 ? 0x2u : 0u), where, true));
    QB_TRY(lanczos_impl(&N.R, N.loop_cplx, k, np, maxit, m_out, N.work.p, hess, purpose, QBGPU_DEVICE, true, stop_state));
    return native_end(N, A, cplx, v, 0x3u, nvec, where);    // the two live vectors, like the reference's slots
}

static int cg_native(qbgpu_matrix *A, bool cplx, int64_t maxit, int64_t *m_io, double2 E0, double *accu_out,
                     void *v, void *r, void *p, void *pp, int where)
{
    QB_TRY(ensure_init());
    QB_TRY(check_single(A, cplx));
    if (!m_io || !accu_out || !v || !r || !p || !pp) return fail(QBGPU_ERR_ARG, "eigenvec_CG: null argument");
    Context &c = ctx();
    const int64_t n = A->n;
    const bool resumed = *m_io > 0;
    NativeRun N;
    QB_TRY(native_begin(N, A, cplx, v, 1, 0x1u, where, E0.y == 0.0 && !resumed));   // (a resumed run keeps the API's scalar type)
    DevBuf rest;                                            // r, p, pp of the loop (internal order)
    QB_TRY(rest.alloc(N.vb_loop * n * 3));
    char *w[4] = {(char *)N.work.p, (char *)rest.p, (char *)rest.p + N.vb_loop * n, (char *)rest.p + 2 * N.vb_loop * n};
    if (resumed) {                                          // r and p of step m come in with v
        DevBuf in;
        void *ins[2] = {r, p};
        if (where == QBGPU_HOST) QB_TRY(in.alloc(N.vb_api * n));
        for (int j = 0; j < 2; j++) {
            const void *src = ins[j];
            if (where == QBGPU_HOST) { QB_CUDA(cudaMemcpyAsync(in.p, ins[j], N.vb_api * n, cudaMemcpyHostToDevice, c.stream)); src = in.p; }
            QB_TRY(vec_to_native(A, cplx, N.loop_cplx, src, w[1 + j]));
            QB_CUDA(cudaStreamSynchronize(c.stream));
        }
    }
    QB_TRY(cg_impl(&N.R, N.loop_cplx, maxit, m_io, E0, accu_out, w[0], w[1], w[2], w[3], QBGPU_DEVICE));
    void *outs[4] = {v, r, p, pp};
    DevBuf tmp;                                             // host callers: one device vector in reference order at a time
    if (where == QBGPU_HOST) QB_TRY(tmp.alloc(N.vb_api * n));
    for (int j = 0; j < 4; j++) {
        void *dst = where == QBGPU_HOST ? tmp.p : outs[j];
        QB_TRY(vec_from_native(A, N.loop_cplx, cplx, w[j], dst, make_double2(1.0, 0.0), make_double2(0.0, 0.0)));
        if (where == QBGPU_HOST) { QB_CUDA(cudaMemcpyAsync(outs[j], dst, N.vb_api * n, cudaMemcpyDeviceToHost, c.stream)); QB_CUDA(cudaStreamSynchronize(c.stream)); }
    }
    QB_CUDA(cudaStreamSynchronize(c.stream));
    return QBGPU_OK;
}

static int energy_scale_native(qbgpu_matrix *A, bool cplx, void *v, double *lo, double *hi, double extend, int64_t iters, int where)
{
    QB_TRY(ensure_init());
    QB_TRY(check_single(A, cplx));
    if (!v || !lo || !hi || iters < 3) return fail(QBGPU_ERR_ARG, "energy_scale: bad argument");
    Context &c = ctx();
    const int64_t n = A->n;
    const size_t vb = cplx ? 16 : 8;
    const int64_t mm = iters - 1;                          // src/kpm.cc:53
    DevBuf dref;                                            // the reference draws its start vector itself (seed 1, src/kpm.cc:60)
    char *ref = (char *)v;
    if (where == QBGPU_HOST) { QB_TRY(dref.alloc(vb * n * 2)); ref = (char *)dref.p; }
    else if (where != QBGPU_DEVICE) return fail(QBGPU_ERR_ARG, "where must be QBGPU_HOST or QBGPU_DEVICE");
    QB_TRY(vec_randomize(n, cplx, ref, 1));
    NativeRun N;
    QB_TRY(native_begin(N, A, cplx, ref, 2, 0x1u, QBGPU_DEVICE, true));
    std::vector<double> hess(2 * iters, 0.0), ritz(mm);
    int64_t m = 0;
    QB_TRY(lanczos_impl(&N.R, N.loop_cplx, 0, mm, iters, &m, N.work.p, hess.data(), "dnmcs", QBGPU_DEVICE, /*stop_on_breakdown=*/false));
    if (hess_eigen_host(hess.data(), iters, mm, ritz.data(), nullptr, nullptr)) return fail(QBGPU_ERR_NUMERIC, "hess_eigen: QL did not converge");
    const double l = ritz[0], h = ritz[mm - 1], slack = extend * (h - l);                          // :83-87
    *lo = l - slack; *hi = h + slack;
    QB_TRY(native_end(N, A, cplx, ref, 0x3u, 2, QBGPU_DEVICE));
    if (where == QBGPU_HOST) { QB_CUDA(cudaMemcpyAsync(v, ref, vb * n * 2, cudaMemcpyDeviceToHost, c.stream)); QB_CUDA(cudaStreamSynchronize(c.stream)); }
    return QBGPU_OK;
}

static int kpm_native(qbgpu_matrix *A, bool cplx, const void *phi, double lo, double hi, int64_t nmom, double *mu, int where)
{
    QB_TRY(ensure_init());
    QB_TRY(check_single(A, cplx));
    if (!phi || !mu || nmom < 1 || !(hi > lo)) return fail(QBGPU_ERR_ARG, "kpm_moments: bad argument");
    NativeRun N;
    QB_TRY(native_begin(N, A, cplx, const_cast<void *>(phi), 1, 0x1u, where, true));
    return kpm_impl(&N.R, N.loop_cplx, N.work.p, lo, hi, nmom, mu, QBGPU_DEVICE);
}

static int lanczos_dispatch(qbgpu_matrix *A, bool cplx, int64_t k, int64_t np, int64_t maxit, int64_t *m_out, void *v,
                            double *hess, const char *purpose, int where, double *stop_state = nullptr)
{
    if (A && A->perm) return lanczos_native(A, cplx, k, np, maxit, m_out, v, hess, purpose, where, stop_state);
    if (!A || !real_mode_enabled(A, cplx) || !v || !purpose || np <= 0 || k < 0)
        return lanczos_impl(A, cplx, k, np, maxit, m_out, v, hess, purpose, where, true, stop_state);
    QB_TRY(ensure_init());
    Context &c = ctx();
    const int64_t n = A->n;
    const bool val1 = strstr(purpose, "val1") != nullptr;
    const int nvec = val1 ? 3 : 2;
    DevBuf dc, dr;
    char *Uc = (char *)v;
    if (where == QBGPU_HOST) {
        QB_TRY(dc.alloc(16 * n * nvec));
        Uc = (char *)dc.p;
        QB_CUDA(cudaMemcpyAsync(Uc, v, 16 * n * (k > 0 ? 2 : 1), cudaMemcpyHostToDevice, c.stream));   // a resumed run has two live vectors
        if (val1) QB_CUDA(cudaMemcpyAsync(Uc + 32 * n, (char *)v + 32 * n, 16 * n, cudaMemcpyHostToDevice, c.stream));
    } else if (where != QBGPU_DEVICE) return fail(QBGPU_ERR_ARG, "where must be QBGPU_HOST or QBGPU_DEVICE");
    bool real0 = false, real1 = true, real2 = true;
    QB_TRY(is_all_real(n, Uc, &real0));
    if (k > 0) QB_TRY(is_all_real(n, Uc + 16 * n, &real1));
    if (val1) QB_TRY(is_all_real(n, Uc + 32 * n, &real2));
    int rc;
    if (real0 && real1 && real2) {
        QB_TRY(dr.alloc(8 * n * nvec));
        double *Ur = (double *)dr.p;
        QB_TRY(vec_take_real(n, Uc, Ur));
        if (k > 0) QB_TRY(vec_take_real(n, Uc + 16 * n, Ur + n));
        if (val1) QB_TRY(vec_take_real(n, Uc + 32 * n, Ur + 2 * n));
        qbgpu_matrix R = *A;                               // same arrays, fp64 vectors
        R.api_complex = false;
        rc = lanczos_impl(&R, false, k, np, maxit, m_out, Ur, hess, purpose, QBGPU_DEVICE, true, stop_state);
        if (rc == QBGPU_OK) { QB_TRY(vec_put_real(n, Ur, Uc)); QB_TRY(vec_put_real(n, Ur + n, Uc + 16 * n)); }
    } else {
        rc = lanczos_impl(A, true, k, np, maxit, m_out, Uc, hess, purpose, QBGPU_DEVICE, true, stop_state);
    }
    if (rc == QBGPU_OK && where == QBGPU_HOST) QB_CUDA(cudaMemcpyAsync(v, Uc, 32 * n, cudaMemcpyDeviceToHost, c.stream));
    QB_CUDA(cudaStreamSynchronize(c.stream));
    return rc;
}

static int cg_dispatch(qbgpu_matrix *A, bool cplx, int64_t maxit, int64_t *m_io, double2 E0, double *accu_out,
                       void *v, void *r, void *p, void *pp, int where)
{
    if (A && A->perm) return cg_native(A, cplx, maxit, m_io, E0, accu_out, v, r, p, pp, where);
    if (!A || !real_mode_enabled(A, cplx) || E0.y != 0.0 || !v || !r || !p || !pp) return cg_impl(A, cplx, maxit, m_io, E0, accu_out, v, r, p, pp, where);
    QB_TRY(ensure_init());
    Context &c = ctx();
    const int64_t n = A->n;
    DevBuf dc, dr;
    char *vc = (char *)v;
    if (where == QBGPU_HOST) {
        QB_TRY(dc.alloc(16 * n));
        vc = (char *)dc.p;
        QB_CUDA(cudaMemcpyAsync(vc, v, 16 * n, cudaMemcpyHostToDevice, c.stream));
    } else if (where != QBGPU_DEVICE) return fail(QBGPU_ERR_ARG, "where must be QBGPU_HOST or QBGPU_DEVICE");
    bool real0 = false;
    QB_TRY(is_all_real(n, vc, &real0));
    const bool resumed = m_io && *m_io > 0;
    if (!real0 || resumed) { dc.release(); return cg_impl(A, true, maxit, m_io, E0, accu_out, v, r, p, pp, where); }   // (a resumed run stays complex)
    QB_TRY(dr.alloc(8 * n * 4));
    double *R4 = (double *)dr.p;
    QB_TRY(vec_take_real(n, vc, R4));
    qbgpu_matrix R = *A;
    R.api_complex = false;
    QB_TRY(cg_impl(&R, false, maxit, m_io, E0, accu_out, R4, R4 + n, R4 + 2 * n, R4 + 3 * n, QBGPU_DEVICE));
    void *outs[4] = {v, r, p, pp};
    if (where == QBGPU_DEVICE) {
        for (int j = 0; j < 4; j++) QB_TRY(vec_put_real(n, R4 + j * n, outs[j]));
    } else {
        for (int j = 0; j < 4; j++) {
            QB_TRY(vec_put_real(n, R4 + j * n, vc));
            QB_CUDA(cudaMemcpyAsync(outs[j], vc, 16 * n, cudaMemcpyDeviceToHost, c.stream));
            QB_CUDA(cudaStreamSynchronize(c.stream));
        }
    }
    QB_CUDA(cudaStreamSynchronize(c.stream));
    return QBGPU_OK;
}

static int energy_scale_dispatch(qbgpu_matrix *A, bool cplx, void *v, double *lo, double *hi, double extend, int64_t iters, int where)
{
    if (A && A->perm) return energy_scale_native(A, cplx, v, lo, hi, extend, iters, where);
    if (!A || !real_mode_enabled(A, cplx) || !v) return energy_scale_impl(A, cplx, v, lo, hi, extend, iters, where);
    QB_TRY(ensure_init());
    Context &c = ctx();
    const int64_t n = A->n;
    DevBuf dr, dc;
    QB_TRY(dr.alloc(8 * n * 2));
    qbgpu_matrix R = *A;
    R.api_complex = false;
    QB_TRY(energy_scale_impl(&R, false, dr.p, lo, hi, extend, iters, QBGPU_DEVICE));   // start vector = vec_randomize: real
    char *vc = (char *)v;
    if (where == QBGPU_HOST) { QB_TRY(dc.alloc(32 * n)); vc = (char *)dc.p; }
    QB_TRY(vec_put_real(n, (double *)dr.p, vc));
    QB_TRY(vec_put_real(n, (double *)dr.p + n, vc + 16 * n));
    if (where == QBGPU_HOST) QB_CUDA(cudaMemcpyAsync(v, vc, 32 * n, cudaMemcpyDeviceToHost, c.stream));
    QB_CUDA(cudaStreamSynchronize(c.stream));
    return QBGPU_OK;
}

static int kpm_dispatch(qbgpu_matrix *A, bool cplx, const void *phi, double lo, double hi, int64_t nmom, double *mu, int where)
{
    if (A && A->perm) return kpm_native(A, cplx, phi, lo, hi, nmom, mu, where);
    if (!A || !real_mode_enabled(A, cplx) || !phi) return kpm_impl(A, cplx, phi, lo, hi, nmom, mu, where);
    QB_TRY(ensure_init());
    Context &c = ctx();
    const int64_t n = A->n;
    DevBuf dc, dr;
    const char *pc = (const char *)phi;
    if (where == QBGPU_HOST) {
        QB_TRY(dc.alloc(16 * n));
        QB_CUDA(cudaMemcpyAsync(dc.p, phi, 16 * n, cudaMemcpyHostToDevice, c.stream));
        pc = (const char *)dc.p;
    } else if (where != QBGPU_DEVICE) return fail(QBGPU_ERR_ARG, "where must be QBGPU_HOST or QBGPU_DEVICE");
    bool real0 = false;
    QB_TRY(is_all_real(n, pc, &real0));
    if (!real0) return kpm_impl(A, true, pc, lo, hi, nmom, mu, QBGPU_DEVICE);
    QB_TRY(dr.alloc(8 * n));
    QB_TRY(vec_take_real(n, pc, (double *)dr.p));
    qbgpu_matrix R = *A;
    R.api_complex = false;
    return kpm_impl(&R, false, dr.p, lo, hi, nmom, mu, QBGPU_DEVICE);
}

}  // namespace qb

using namespace qb;

extern "C" {

int qbgpu_hess_eigen(const double *hess, int64_t maxit, int64_t m, double *ritz, double *s)
{
    if (!hess || !ritz || m < 1 || m >= maxit) return fail(QBGPU_ERR_ARG, "hess_eigen: need 0 < m < maxit");   // src/lanczos.cc:358
    if (hess_eigen_host(hess, maxit, m, ritz, s, nullptr)) return fail(QBGPU_ERR_NUMERIC, "hess_eigen: QL did not converge");
    return QBGPU_OK;
}

int qbgpu_hess_smallest(const double *hess, int64_t maxit, int64_t m, double *theta0)
{
    if (!hess || !theta0 || m < 1 || m >= maxit) return fail(QBGPU_ERR_ARG, "hess_smallest: need 0 < m < maxit");
    *theta0 = tridiag_smallest(hess, maxit, m);
    return QBGPU_OK;
}

int qbgpu_lanczos_d(qbgpu_matrix_t A, int64_t k, int64_t np, int64_t maxit, int64_t *m, double *v, double *hess, const char *purpose, int where)
{ return lanczos_dispatch(A, false, k, np, maxit, m, v, hess, purpose, where); }
int qbgpu_lanczos_z(qbgpu_matrix_t A, int64_t k, int64_t np, int64_t maxit, int64_t *m, void *v, double *hess, const char *purpose, int where)
{ return lanczos_dispatch(A, true, k, np, maxit, m, v, hess, purpose, where); }

int qbgpu_lanczos_resume_d(qbgpu_matrix_t A, int64_t k, int64_t np, int64_t maxit, int64_t *m, double *v, double *hess, const char *purpose, int where, double *stop_state)
{ return lanczos_dispatch(A, false, k, np, maxit, m, v, hess, purpose, where, stop_state); }
int qbgpu_lanczos_resume_z(qbgpu_matrix_t A, int64_t k, int64_t np, int64_t maxit, int64_t *m, void *v, double *hess, const char *purpose, int where, double *stop_state)
{ return lanczos_dispatch(A, true, k, np, maxit, m, v, hess, purpose, where, stop_state); }

int qbgpu_eigenvec_cg_d(qbgpu_matrix_t A, int64_t maxit, int64_t *m, double E0, double *accu, double *v, double *r, double *p, double *pp, int where)
{ return cg_dispatch(A, false, maxit, m, make_double2(E0, 0.0), accu, v, r, p, pp, where); }
int qbgpu_eigenvec_cg_z(qbgpu_matrix_t A, int64_t maxit, int64_t *m, const double E0[2], double *accu, void *v, void *r, void *p, void *pp, int where)
{ if (!E0) return fail(QBGPU_ERR_ARG, "null E0"); return cg_dispatch(A, true, maxit, m, make_double2(E0[0], E0[1]), accu, v, r, p, pp, where); }

int qbgpu_energy_scale_d(qbgpu_matrix_t A, double *v, double *lo, double *hi, double extend, int64_t iters, int where)
{ return energy_scale_dispatch(A, false, v, lo, hi, extend, iters, where); }
int qbgpu_energy_scale_z(qbgpu_matrix_t A, void *v, double *lo, double *hi, double extend, int64_t iters, int where)
{ return energy_scale_dispatch(A, true, v, lo, hi, extend, iters, where); }

int qbgpu_kpm_moments_d(qbgpu_matrix_t A, const double *phi, double lo, double hi, int64_t nmom, double *mu, int where)
{ return kpm_dispatch(A, false, phi, lo, hi, nmom, mu, where); }
int qbgpu_kpm_moments_z(qbgpu_matrix_t A, const void *phi, double lo, double hi, int64_t nmom, double *mu, int where)
{ return kpm_dispatch(A, true, phi, lo, hi, nmom, mu, where); }

// ---------------------------------------------------------------- step-level entry points, DEVICE vectors
// The bodies the loops above are made of, one call per step, for a host that keeps the reference's loop and replaces its body
// (eigenvec_CG: src/lanczos.cc:293-332).  Element type and order of the vectors follow the handle (qbgpu_native_order); no
// real-mode dispatch, no permutation.
int qbgpu_cg_restart(qbgpu_matrix_t A, const double E0[2], double *sc_dev, void *v, void *r, void *p, double *vnorm, double *accu)
{
    QB_TRY(ensure_init());
    if (!A || !E0 || !sc_dev || !v || !r || !p || !accu) return fail(QBGPU_ERR_ARG, "cg_restart: null argument");
    QB_TRY(check_single(A, A->api_complex));
    double nn = 0.0;
    QB_TRY(vec_nrm2sq(A->n, A->api_complex, v, sc_dev + 6));
    QB_TRY(read_scalars(sc_dev + 6, &nn, 1));
    const double rnorm = sqrt(nn);
    if (vnorm) *vnorm = rnorm;
    if (!(rnorm > 0.0)) return fail(QBGPU_ERR_NUMERIC, "cg_restart: the vector is zero");
    return cg_restart_piece(A, A->api_complex, make_double2(E0[0], A->api_complex ? E0[1] : 0.0), rnorm, sc_dev, (char *)v, (char *)r, (char *)p, accu);
}

int qbgpu_cg_step(qbgpu_matrix_t A, const double E0[2], double *sc_dev, void *v, void *r, void *p, void *pp, double *accu)
{
    QB_TRY(ensure_init());
    if (!A || !E0 || !sc_dev || !v || !r || !p || !pp || !accu) return fail(QBGPU_ERR_ARG, "cg_step: null argument");
    QB_TRY(check_single(A, A->api_complex));
    return cg_iterate_piece(A, A->api_complex, make_double2(E0[0], A->api_complex ? E0[1] : 0.0), sc_dev, (char *)v, (char *)r, (char *)p, (char *)pp, accu);
}

// One step of the Chebyshev recurrence as ONE fused product (the body of kpm_impl): Ht = (H - c)/s, c = (hi+lo)/2, s = (hi-lo)/2;
// first: t_next = Ht t_cur; otherwise t_next = 2 Ht t_cur - t_prev (t_next may alias t_prev).  dots (may be NULL): the local
// rows' <t_cur, t_next> (2 doubles) and |t_next|^2 -- what the 2k / 2k+1 doubling identities need.  Works on shards too
// (t_cur: the full vector; t_prev / t_next: the local rows), like qbgpu_spmv_fused.
int qbgpu_cheb_step(qbgpu_matrix_t A, double lo, double hi, int first, const void *t_cur_full, const void *t_prev_local, void *t_next_local, double *dots_dev)
{
    QB_TRY(ensure_init());
    if (!A || !t_cur_full || !t_next_local || !(hi > lo)) return fail(QBGPU_ERR_ARG, "cheb_step: bad argument");
    if (!first && !t_prev_local) return fail(QBGPU_ERR_ARG, "cheb_step: t_prev is needed after the first step");
    const double cc = 0.5 * (hi + lo), ss = 0.5 * (hi - lo);
    FusedArgs fa;
    fa.x = t_cur_full; fa.y = t_next_local; fa.dots = dots_dev;
    if (first) { fa.alpha = make_double2(1.0 / ss, 0.0); fa.gamma = make_double2(-cc / ss, 0.0); }
    else { fa.alpha = make_double2(2.0 / ss, 0.0); fa.gamma = make_double2(-2.0 * cc / ss, 0.0); fa.beta = make_double2(-1.0, 0.0); fa.z = t_prev_local; }
    return launch_spmv(A, fa);
}

}  // extern "C"
