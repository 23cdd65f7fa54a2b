// quantum_basis_b200/csrc/trlan.cu -- device-resident thick-restart Lanczos: the lowest `nev` eigenpairs with a
// Krylov basis of `ncv` vectors kept in HBM.
//
// It serves the contract of the reference's iram<T,MAT>() / model<T>::locate_E0_iram (src/lanczos.cc:497-603,
// src/model.cc:1320-1366: nev eigenpairs of smallest algebraic value, basis size ncv, at most `maxit` restarts) without
// ARPACK's reverse communication: there every product crosses PCIe twice (x and y live in ARPACK's host `workd`,
// src/lanczos.cc:476) and ARPACK's own orthogonalisation runs on the CPU.  Here the basis, the products, the
// Gram-Schmidt passes and the restart rotation stay on the device; only the ncv x ncv projected matrix is solved on the
// host.  Thick-restart Lanczos (Wu & Simon, SIAM J. Matrix Anal. Appl. 22 (2000) 602) is the Hermitian equivalent of the
// implicitly restarted method ARPACK implements, so the eigenvalues agree (parity is on eigenvalues).
#include "internal.hpp"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <vector>

namespace qb {

using cd = std::complex<double>;

// ----------------------------------------------------------------- dense Hermitian eigensolver (host, m <= 64)
// Cyclic complex Jacobi: A (m x m, column-major, Hermitian, destroyed) -> w ascending, S columns = eigenvectors.
int herm_eigen_host(int m, cd *A, double *w, cd *S)
{
    for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) S[i + (size_t)j * m] = (i == j) ? 1.0 : 0.0;
    auto at = [&](int i, int j) -> cd & { return A[i + (size_t)j * m]; };
    double norm = 0.0;
    for (int i = 0; i < m * m; i++) norm += std::norm(A[i]);
    norm = std::sqrt(norm);
    if (norm == 0.0) norm = 1.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int p = 0; p < m; p++) for (int q = p + 1; q < m; q++) off += std::norm(at(p, q));
        if (std::sqrt(off) <= 1e-16 * norm) break;
        for (int p = 0; p < m - 1; p++)
            for (int q = p + 1; q < m; q++) {
                const cd c = at(p, q);
                const double ac = std::abs(c);
                if (ac <= 1e-300) continue;
                const double a = at(p, p).real(), b = at(q, q).real();
                const double theta = 0.5 * std::atan2(2.0 * ac, a - b);
                const double cs = std::cos(theta), sn = std::sin(theta);
                const cd ph = c / ac;                         // e^{i phi}
                // J = [[cs, -ph*sn], [conj(ph)*sn, cs]] on the (p,q) plane;  A <- J^H A J ;  S <- S J
                for (int k = 0; k < m; k++) {                 // columns p,q of A and S
                    const cd akp = at(k, p), akq = at(k, q);
                    at(k, p) = akp * cs + akq * std::conj(ph) * sn;
                    at(k, q) = -akp * ph * sn + akq * cs;
                    const cd skp = S[k + (size_t)p * m], skq = S[k + (size_t)q * m];
                    S[k + (size_t)p * m] = skp * cs + skq * std::conj(ph) * sn;
                    S[k + (size_t)q * m] = -skp * ph * sn + skq * cs;
                }
                for (int k = 0; k < m; k++) {                 // rows p,q of A: J^H from the left
                    const cd apk = at(p, k), aqk = at(q, k);
                    at(p, k) = apk * cs + aqk * ph * sn;
                    at(q, k) = -apk * std::conj(ph) * sn + aqk * cs;
                }
                at(p, q) = 0.0; at(q, p) = 0.0;
                at(p, p) = at(p, p).real(); at(q, q) = at(q, q).real();
            }
    }
    std::vector<int> ord(m);
    for (int i = 0; i < m; i++) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](int i, int j) { return at(i, i).real() < at(j, j).real(); });
    std::vector<cd> T((size_t)m * m);
    for (int j = 0; j < m; j++) { w[j] = at(ord[j], ord[j]).real(); for (int i = 0; i < m; i++) T[i + (size_t)j * m] = S[i + (size_t)ord[j] * m]; }
    std::copy(T.begin(), T.end(), S);
    return 0;
}

// -------------------------------------------------------------------------------------------- device kernels
constexpr int kTBlock = 256;
constexpr int kMaxCv = 64;

// h[c] = <V_c, w> for c < nc (conj on V): one pass over w and the nc basis vectors.  nc <= 8 per launch.
template <typename VecT, int NC>
__global__ void __launch_bounds__(kTBlock, 4) multi_dot_kernel(int64_t n, int64_t ld, const VecT *__restrict__ V, const VecT *__restrict__ w,
                                                               double *out /* [2*NC] */, double *partials, unsigned *ticket)
{
    using VT = VecTraits<VecT>;
    double d[2 * NC];
#pragma unroll
    for (int c = 0; c < 2 * NC; c++) d[c] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const VecT wi = w[i];
        VecT v[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) v[c] = V[i + c * ld];
#pragma unroll
        for (int c = 0; c < NC; c++) { const double2 p = VT::conj_mul(v[c], wi); d[2 * c] += p.x; d[2 * c + 1] += p.y; }
    }
    // block reduction with a wide partial row (this kernel has its own partial buffer: 2*NC slots per block)
    __shared__ double sm[2 * NC][kTBlock / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < 2 * NC; s++) {
        double a = d[s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) sm[s][warp] = a;
    }
    __syncthreads();
    if (threadIdx.x < 2 * NC) {
        double a = 0.0;
        for (int k = 0; k < kTBlock / 32; k++) a += sm[threadIdx.x][k];
        partials[(size_t)blockIdx.x * 2 * NC + threadIdx.x] = a;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last) {
        __threadfence();
        if (threadIdx.x < 2 * NC) {
            double a = 0.0;
            for (unsigned b = 0; b < gridDim.x; b++) a += __ldcg(&partials[(size_t)b * 2 * NC + threadIdx.x]);
            out[threadIdx.x] = a;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

// w -= sum_c h[c] V_c  (h: device, complex interleaved; real vectors use the real parts)
template <typename VecT, int NC>
__global__ void __launch_bounds__(kTBlock, 4) multi_axpy_kernel(int64_t n, int64_t ld, const VecT *__restrict__ V, const double *__restrict__ h, VecT *w)
{
    using VT = VecTraits<VecT>;
    double2 hc[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) hc[c] = make_double2(-h[2 * c], -h[2 * c + 1]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        VecT v[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) v[c] = V[i + c * ld];
        VecT acc = w[i];
#pragma unroll
        for (int c = 0; c < NC; c++) acc = VT::add(acc, VT::scale(hc[c], v[c]));
        w[i] = acc;
    }
}

// restart: V[:, 0:k] <- V[:, 0:m] * S[:, 0:k]  row by row, in place (each thread owns one row of the basis)
template <typename VecT>
__global__ void __launch_bounds__(kTBlock, 2) basis_rotate_kernel(int64_t n, int64_t ld, VecT *V, int m, int k, const double2 *__restrict__ S /* m x k col-major */)
{
    using VT = VecTraits<VecT>;
    extern __shared__ double2 sS[];
    for (int t = threadIdx.x; t < m * k; t += blockDim.x) sS[t] = S[t];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        VecT r[kMaxCv];
        for (int c = 0; c < m; c++) r[c] = V[i + c * ld];
        for (int j = 0; j < k; j++) {
            VecT acc = VT::zero();
            for (int c = 0; c < m; c++) acc = VT::add(acc, VT::scale(sS[c + j * m], r[c]));
            V[i + j * ld] = acc;
        }
    }
}

static int tgrid(int64_t n)
{
    int64_t want = (n + kTBlock - 1) / kTBlock, cap = (int64_t)ctx().num_sms * 4;
    if (cap > kMaxPartialBlocks) cap = kMaxPartialBlocks;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

template <typename VecT>
static int multi_dot(int64_t n, int64_t ld, const VecT *V, int nc, const VecT *w, double *out, double *partials)
{
    Context &c = ctx();
    for (int c0 = 0; c0 < nc; c0 += 8) {
        const int m = std::min(8, nc - c0);
        const VecT *Vc = V + (int64_t)c0 * ld;
        double *o = out + 2 * c0;
#define QB_MD(NCV) multi_dot_kernel<VecT, NCV><<<tgrid(n), kTBlock, 0, c.stream>>>(n, ld, Vc, w, o, partials, c.ticket)
        switch (m) { case 1: QB_MD(1); break; case 2: QB_MD(2); break; case 3: QB_MD(3); break; case 4: QB_MD(4); break;
                     case 5: QB_MD(5); break; case 6: QB_MD(6); break; case 7: QB_MD(7); break; default: QB_MD(8); break; }
#undef QB_MD
        QB_LAUNCH_COUNT();
        QB_CUDA(cudaGetLastError());
    }
    return QBGPU_OK;
}

template <typename VecT>
static int multi_axpy(int64_t n, int64_t ld, const VecT *V, int nc, const double *h, VecT *w)
{
    Context &c = ctx();
    for (int c0 = 0; c0 < nc; c0 += 8) {
        const int m = std::min(8, nc - c0);
        const VecT *Vc = V + (int64_t)c0 * ld;
        const double *hc = h + 2 * c0;
#define QB_MA(NCV) multi_axpy_kernel<VecT, NCV><<<tgrid(n), kTBlock, 0, c.stream>>>(n, ld, Vc, hc, w)
        switch (m) { case 1: QB_MA(1); break; case 2: QB_MA(2); break; case 3: QB_MA(3); break; case 4: QB_MA(4); break;
                     case 5: QB_MA(5); break; case 6: QB_MA(6); break; case 7: QB_MA(7); break; default: QB_MA(8); break; }
#undef QB_MA
        QB_LAUNCH_COUNT();
        QB_CUDA(cudaGetLastError());
    }
    return QBGPU_OK;
}

// ------------------------------------------------------------------------------------------------- driver
template <typename VecT>
// largest: the algebraically largest eigenvalues (ARPACK's "LA", the reference's order "lr"): the same iteration on -H --
// every product is taken with alpha = -1 -- and the Ritz values negated on the way out.
static int trlan_typed(qbgpu_matrix *A, int nev, int ncv, int maxit, double tol, int *nconv_out, int *nprod_out, double *evals,
                       void *evecs, int where, uint32_t seed, bool largest = false)
{
    Context &c = ctx();
    constexpr bool cplx = sizeof(VecT) == 16;
    const int64_t n = A->n;
    const size_t vb = sizeof(VecT);
    if (tol <= 0.0) tol = DBL_EPSILON;                     // like ARPACK with tol = 0 (src/lanczos.cc:405,452)
    const double eps23 = std::pow(DBL_EPSILON, 2.0 / 3.0);
    VecT *V = nullptr;
    double *d_h = nullptr, *d_part = nullptr;
    double2 *d_S = nullptr;
    auto cleanup = [&]() { cudaFree(V); cudaFree(d_h); cudaFree(d_part); cudaFree(d_S); };
#define QB_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return cuda_fail(e_, #call, __FILE__, __LINE__); } } while (0)
#define QB_TR(call) do { int rc_ = (call); if (rc_) { cleanup(); return rc_; } } while (0)
    QB_CU(cudaMalloc(&V, vb * (size_t)n * (ncv + 1)));
    QB_CU(cudaMalloc(&d_h, sizeof(double) * 2 * (ncv + 2)));
    QB_CU(cudaMalloc(&d_part, sizeof(double) * 16 * kMaxPartialBlocks));
    QB_CU(cudaMalloc(&d_S, sizeof(double2) * (size_t)ncv * ncv));
    auto col = [&](int j) { return V + (int64_t)j * n; };
    QB_TR(vec_randomize(n, cplx, col(0), seed));           // ARPACK draws its own random start (info = 0, :421,470)

    std::vector<cd> T((size_t)ncv * ncv, 0.0), Tw((size_t)ncv * ncv), S((size_t)ncv * ncv);
    std::vector<double> theta(ncv), hh(2 * (ncv + 2));
    int k = 0, nconv = 0, nprod = 0, restarts = 0;
    double beta = 0.0;
    for (;;) {
        for (int j = k; j < ncv; j++) {                    // extend the basis: columns k..ncv-1, residual in column ncv
            FusedArgs fa;
            fa.x = col(j); fa.y = col(j + 1);
            if (largest) fa.alpha = make_double2(-1.0, 0.0);
            QB_TR(launch_spmv(A, fa));
            nprod++;
            // classical Gram-Schmidt against columns 0..j, repeated when the DGKS test asks for it
            double wnorm_before = 0.0, wnorm_after = 0.0;
            for (int pass = 0; pass < 2; pass++) {
                QB_TR(vec_nrm2sq(n, cplx, col(j + 1), d_h + 2 * (ncv + 1)));
                QB_TR(multi_dot<VecT>(n, n, V, j + 1, col(j + 1), d_h, d_part));
                QB_TR(multi_axpy<VecT>(n, n, V, j + 1, d_h, col(j + 1)));
                QB_CU(cudaMemcpyAsync(hh.data(), d_h, sizeof(double) * 2 * (ncv + 2), cudaMemcpyDeviceToHost, c.stream));
                QB_CU(cudaStreamSynchronize(c.stream));
                for (int i = 0; i <= j; i++) {
                    const cd hi(hh[2 * i], cplx ? hh[2 * i + 1] : 0.0);
                    if (pass == 0) T[i + (size_t)j * ncv] = hi; else T[i + (size_t)j * ncv] += hi;
                }
                wnorm_before = std::sqrt(hh[2 * (ncv + 1)]);
                QB_TR(vec_nrm2sq(n, cplx, col(j + 1), d_h + 2 * (ncv + 1)));
                double t;
                QB_TR(read_scalars(d_h + 2 * (ncv + 1), &t, 1));
                wnorm_after = std::sqrt(t);
                if (wnorm_after > 0.7071 * wnorm_before) break;      // DGKS: no second pass needed
            }
            for (int i = 0; i < j; i++) T[j + (size_t)i * ncv] = std::conj(T[i + (size_t)j * ncv]);
            T[j + (size_t)j * ncv] = T[j + (size_t)j * ncv].real();
            beta = wnorm_after;
            if (beta < 1e-300) { cleanup(); return fail(QBGPU_ERR_NUMERIC, "thick-restart Lanczos: invariant subspace (breakdown)"); }
            QB_TR(vec_scal(n, cplx, make_double2(1.0 / beta, 0.0), col(j + 1)));
            if (j + 1 < ncv) { T[(j + 1) + (size_t)j * ncv] = beta; T[j + (size_t)(j + 1) * ncv] = beta; }
        }
        // Ritz pairs of the projected matrix
        Tw = T;
        herm_eigen_host(ncv, Tw.data(), theta.data(), S.data());
        nconv = 0;
        for (int i = 0; i < nev; i++) {
            const double est = std::abs(beta * S[(ncv - 1) + (size_t)i * ncv]);
            if (est <= tol * std::max(eps23, std::abs(theta[i]))) nconv++; else break;
        }
        if (getenv("QBGPU_VERBOSE")) fprintf(stderr, "[qbgpu trlan] restart %d: %d products, theta0 = %.12f, converged %d/%d\n", restarts, nprod, theta[0], nconv, nev);
        const bool done = (nconv >= nev) || (restarts >= maxit);
        const int keep = done ? nev : std::min(ncv - 1, nev + std::max(1, (ncv - nev) / 2));
        // V[:, 0:keep] <- V[:, 0:ncv] S[:, 0:keep]
        std::vector<double2> Sk((size_t)ncv * keep);
        for (int j = 0; j < keep; j++) for (int i = 0; i < ncv; i++) Sk[i + (size_t)j * ncv] = make_double2(S[i + (size_t)j * ncv].real(), S[i + (size_t)j * ncv].imag());
        QB_CU(cudaMemcpyAsync(d_S, Sk.data(), sizeof(double2) * Sk.size(), cudaMemcpyHostToDevice, c.stream));
        basis_rotate_kernel<VecT><<<tgrid(n), kTBlock, sizeof(double2) * ncv * keep, c.stream>>>(n, n, V, ncv, keep, d_S);
        QB_LAUNCH_COUNT();
        QB_CU(cudaGetLastError());
        QB_CU(cudaStreamSynchronize(c.stream));             // Sk lives on the host stack frame
        if (done) break;
        // the residual vector becomes column `keep`; T = diag(theta) with the coupling row beta * s_last
        QB_CU(cudaMemcpyAsync(col(keep), col(ncv), vb * (size_t)n, cudaMemcpyDeviceToDevice, c.stream));
        std::fill(T.begin(), T.end(), cd(0.0));
        for (int i = 0; i < keep; i++) {
            T[i + (size_t)i * ncv] = theta[i];
            const cd cpl = beta * S[(ncv - 1) + (size_t)i * ncv];
            T[keep + (size_t)i * ncv] = cpl;                 // <v_keep| H |y_i> = beta * s_i[last]
            T[i + (size_t)keep * ncv] = std::conj(cpl);
        }
        k = keep;
        restarts++;
        // the next pass starts at j = keep: its Gram-Schmidt recomputes column `keep` of T (diagonal and couplings)
    }
    for (int i = 0; i < nev; i++) evals[i] = largest ? -theta[i] : theta[i];
    if (evecs && A->perm) {
        // species-order handle: the basis lives in the handle's internal order; hand the Ritz vectors back in the reference's
        // (column ncv, the residual, is free by now and serves as the staging vector of host callers)
        for (int i = 0; i < nev; i++) {
            VecT *dst = where == QBGPU_HOST ? col(ncv) : (VecT *)evecs + (int64_t)i * n;
            QB_TR(vec_from_native(A, cplx, cplx, col(i), dst, make_double2(1.0, 0.0), make_double2(0.0, 0.0)));
            if (where == QBGPU_HOST) QB_CU(cudaMemcpyAsync((char *)evecs + vb * (size_t)n * i, dst, vb * (size_t)n, cudaMemcpyDeviceToHost, c.stream));
        }
    } else if (evecs) {
        if (where == QBGPU_HOST) QB_CU(cudaMemcpyAsync(evecs, V, vb * (size_t)n * nev, cudaMemcpyDeviceToHost, c.stream));
        else QB_CU(cudaMemcpyAsync(evecs, V, vb * (size_t)n * nev, cudaMemcpyDeviceToDevice, c.stream));
    }
    QB_CU(cudaStreamSynchronize(c.stream));
    cleanup();
#undef QB_CU
#undef QB_TR
    *nconv_out = nconv;
    if (nprod_out) *nprod_out = nprod;
    return QBGPU_OK;
}

}  // namespace qb

using namespace qb;

extern "C" {

int qbgpu_herm_eigen(int m, const void *a_colmajor, double *w, void *s_colmajor)
{
    if (m < 1 || m > 1024 || !a_colmajor || !w || !s_colmajor) return fail(QBGPU_ERR_ARG, "herm_eigen: bad argument");
    std::vector<cd> A((const cd *)a_colmajor, (const cd *)a_colmajor + (size_t)m * m);
    return herm_eigen_host(m, A.data(), w, (cd *)s_colmajor);
}

int qbgpu_trlan(qbgpu_matrix_t A, int nev, int ncv, int maxit, double tol, int *nconv, int *nprod, double *eigenvals, void *eigenvecs, int where)
{
    QB_TRY(ensure_init());
    if (!A || !nconv || !eigenvals) return fail(QBGPU_ERR_ARG, "trlan: null argument");
    if (A->row_lo != 0 || A->row_hi != A->n) return fail(QBGPU_ERR_STATE, "trlan needs an unsharded handle");
    if (nev <= 0 || nev >= A->n - 1) return fail(QBGPU_ERR_ARG, "0 < nev < N-1 should be satisfied.");        // src/lanczos.cc:502
    if (ncv <= nev + 1 || ncv > 48 || ncv >= A->n) return fail(QBGPU_ERR_ARG, "trlan: need nev + 1 < ncv <= 48 and ncv < N");   // src/model.cc:1338
    if (maxit <= 0) maxit = nev * 100;                                                                       // src/model.cc:1339
    if (where != QBGPU_HOST && where != QBGPU_DEVICE) return fail(QBGPU_ERR_ARG, "where must be QBGPU_HOST or QBGPU_DEVICE");
    return A->api_complex ? trlan_typed<double2>(A, nev, ncv, maxit, tol, nconv, nprod, eigenvals, eigenvecs, where, 1)
                          : trlan_typed<double>(A, nev, ncv, maxit, tol, nconv, nprod, eigenvals, eigenvecs, where, 1);
}

int qbgpu_trlan_largest(qbgpu_matrix_t A, int nev, int ncv, int maxit, double tol, int *nconv, int *nprod, double *eigenvals, void *eigenvecs, int where)
{
    QB_TRY(ensure_init());
    if (!A || !nconv || !eigenvals) return fail(QBGPU_ERR_ARG, "trlan: null argument");
    if (A->row_lo != 0 || A->row_hi != A->n) return fail(QBGPU_ERR_STATE, "trlan needs an unsharded handle");
    if (nev <= 0 || nev >= A->n - 1) return fail(QBGPU_ERR_ARG, "0 < nev < N-1 should be satisfied.");        // src/lanczos.cc:502
    if (ncv <= nev + 1 || ncv > 48 || ncv >= A->n) return fail(QBGPU_ERR_ARG, "trlan: need nev + 1 < ncv <= 48 and ncv < N");   // src/model.cc:1389
    if (maxit <= 0) maxit = nev * 100;                                                                       // src/model.cc:1390
    if (where != QBGPU_HOST && where != QBGPU_DEVICE) return fail(QBGPU_ERR_ARG, "where must be QBGPU_HOST or QBGPU_DEVICE");
    return A->api_complex ? trlan_typed<double2>(A, nev, ncv, maxit, tol, nconv, nprod, eigenvals, eigenvecs, where, 1, true)
                          : trlan_typed<double>(A, nev, ncv, maxit, tol, nconv, nprod, eigenvals, eigenvecs, where, 1, true);
}

}  // extern "C"
