// quantum_basis_b200/csrc/orbit.cu -- momentum sectors for an ARBITRARY abelian translation group (BASELINE config 4).
//
// Config 4 is the spin-1/2 Heisenberg model on the tilted 31-site triangular cluster (A0 = [5,1], A1 = [-1,6]) in a
// momentum sector.  The reference cannot build it (divide_lattice asserts !q_tilted(), src/lattice.cc:1079,
// and both repr builders divide by L[d], src/model.cc:811), so there is no reference convention to reproduce here --
// unlike sectors.cu, which is bit-identical to the reference on untilted lattices.  This assembler therefore uses the
// textbook convention (Sandvik, AIP Conf. Proc. 1297, 135): the representative of an orbit is its smallest bit pattern,
//     |r_k> = P_k |r> / sqrt(<r|P_k|r>),   P_k = 1/|G| sum_g conj(chi_k(g)) T_g,   <r|P_k|r> = |Stab r| / |G|
// for orbits on whose stabiliser chi_k is trivial (the others do not exist at this momentum and are dropped), and
//     <r'_k| H |r_k> = sum_bonds h_b conj(chi_k(g_b)) sqrt(|Stab r'| / |Stab r|),      T_{g_b} (bond b applied to r) = r'.
// The group is given as site permutations with the characters of the wanted irrep, so any cluster (tilted or not, any
// dimension, point-group-free) works; translated states cost four byte-table lookups.  The upper triangle is assembled
// in HBM and expanded by matrix.cu like every other input.  Checked by spectra: the union over all momenta equals the
// spectrum of the full Sz sector, and on untilted clusters each sector equals the reference-convention sector.
#include "internal.hpp"
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/device/device_reduce.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <algorithm>
#include <cstring>
#include <vector>

namespace qb {

constexpr int kOBlock = 128;
constexpr int kOMaxTrans = 64;
constexpr int kOMaxTerms = 128;

struct OrbitDev {
    int nsites, ntrans, ndown, shift;
    int64_t n;
    const uint32_t *lut;               // [ntrans][4][256]: image of each byte of a state under translation t
    const uint32_t *keys;              // [n] representatives, ascending
    const uint32_t *seg;               // first-level index on key >> shift
    const uint8_t  *stab;              // [n] |Stab r|
    double2 chi[kOMaxTrans];
};
__constant__ unsigned long long c_binom[33][17];   // C(i, k) for the colex unranking (k <= 16 after the up/down symmetry)

__device__ __forceinline__ uint32_t orbit_apply(const OrbitDev &O, int t, uint32_t s)
{
    const uint32_t *L = O.lut + (size_t)t * 1024;
    return L[s & 255u] | L[256 + ((s >> 8) & 255u)] | L[512 + ((s >> 16) & 255u)] | L[768 + (s >> 24)];
}

// colex unranking: the r-th (ascending integer order) state with exactly k bits set among nsites
__device__ __forceinline__ uint32_t orbit_unrank(const OrbitDev &O, unsigned long long r, int k, bool complement)
{
    uint32_t s = 0;
    for (int i = O.nsites - 1; i >= 0 && k > 0; i--) {
        const unsigned long long c = c_binom[i][k];
        if (r >= c) { s |= 1u << i; r -= c; k--; }
    }
    return complement ? (~s & (O.nsites == 32 ? 0xFFFFFFFFu : ((1u << O.nsites) - 1u))) : s;
}

// flag[c] = 1 when candidate c (c-th state with ndown bits, ascending) is the smallest element of its orbit and the
// character is trivial on its stabiliser
__global__ void __launch_bounds__(kOBlock) orbit_flag_kernel(OrbitDev O, int64_t first, int64_t count, int kbits, bool complement, uint8_t *flag, uint32_t *state_out)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < count; c += (int64_t)gridDim.x * blockDim.x) {
        // with complement the enumeration runs over the complemented patterns in DESCENDING order of the state
        const uint32_t s = orbit_unrank(O, (unsigned long long)(first + c), kbits, complement);
        bool ok = true;
        for (int t = 1; t < O.ntrans && ok; t++) {
            const uint32_t u = orbit_apply(O, t, s);
            if (u < s) ok = false;
            else if (u == s && (fabs(O.chi[t].x - 1.0) > 1e-9 || fabs(O.chi[t].y) > 1e-9)) ok = false;
        }
        flag[c] = ok ? 1 : 0;
        state_out[c] = s;
    }
}

__global__ void __launch_bounds__(kOBlock) orbit_stab_kernel(OrbitDev O, uint8_t *stab)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < O.n; r += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t s = O.keys[r];
        int cnt = 0;
        for (int t = 0; t < O.ntrans; t++) cnt += orbit_apply(O, t, s) == s;
        stab[r] = (uint8_t)cnt;
    }
}

__global__ void __launch_bounds__(kOBlock) orbit_seg_kernel(int64_t n, const uint32_t *__restrict__ keys, int shift, uint32_t nseg, uint32_t *seg)
{
    for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h <= nseg; h += gridDim.x * blockDim.x) {
        const uint64_t target = (uint64_t)h << shift;
        int64_t lo = 0, hi = n;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if ((uint64_t)keys[mid] < target) lo = mid + 1; else hi = mid; }
        seg[h] = (uint32_t)lo;
    }
}

__device__ __forceinline__ int64_t orbit_lookup(const OrbitDev &O, uint32_t key)
{
    const uint32_t h = key >> O.shift;
    uint32_t lo = O.seg[h], hi = O.seg[h + 1];
    const uint32_t end = hi;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (O.keys[mid] < key) lo = mid + 1; else hi = mid; }
    return (lo < end && O.keys[lo] == key) ? (int64_t)lo : -1;
}

struct OrbitTerms {
    int nterms;
    double J;
    uint8_t p[kOMaxTerms], q[kOMaxTerms];
};

__global__ void __launch_bounds__(kOBlock) orbit_cap_kernel(OrbitDev O, const OrbitTerms *Tp, int64_t *cap)
{
    __shared__ OrbitTerms T;
    for (int i = threadIdx.x; i < (int)(sizeof(OrbitTerms) / 4); i += blockDim.x) ((int *)&T)[i] = ((const int *)Tp)[i];
    __syncthreads();
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= O.n; r += (int64_t)gridDim.x * blockDim.x) {
        int c = 0;
        if (r < O.n) {
            const uint32_t s = O.keys[r];
            c = 1;
            for (int t = 0; t < T.nterms; t++) c += (((s >> T.p[t]) ^ (s >> T.q[t])) & 1u);
        }
        cap[r] = c;
    }
}

__global__ void __launch_bounds__(kOBlock) orbit_rows_kernel(OrbitDev O, const OrbitTerms *Tp, const int64_t *__restrict__ start, int64_t *row_end,
                                                             int64_t *ocol, double2 *oval)
{
    __shared__ OrbitTerms T;
    for (int i = threadIdx.x; i < (int)(sizeof(OrbitTerms) / 4); i += blockDim.x) ((int *)&T)[i] = ((const int *)Tp)[i];
    __syncthreads();
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < O.n; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t *col = ocol + start[r];
        double2 *val = oval + start[r];
        const uint32_t s = O.keys[r];
        const double stab_r = (double)O.stab[r];
        double dg = 0.0;
        for (int t = 0; t < T.nterms; t++) dg += (((s >> T.p[t]) ^ (s >> T.q[t])) & 1u) ? -0.25 * T.J : 0.25 * T.J;
        int len = 1;
        col[0] = r; val[0] = make_double2(dg, 0.0);
        for (int t = 0; t < T.nterms; t++) {
            if (!(((s >> T.p[t]) ^ (s >> T.q[t])) & 1u)) continue;
            const uint32_t s2 = s ^ ((1u << T.p[t]) | (1u << T.q[t]));
            uint32_t best = s2; int g = 0;
            for (int tt = 1; tt < O.ntrans; tt++) { const uint32_t u = orbit_apply(O, tt, s2); if (u < best) { best = u; g = tt; } }
            const int64_t j = orbit_lookup(O, best);
            if (j < r) continue;                           // upper triangle; -1: the orbit does not exist at this momentum
            // row r, column j:  <r_k|H|j_k> = conj(<j_k|H|r_k>) = h chi(g) sqrt(|Stab j| / |Stab r|)
            const double w = 0.5 * T.J * sqrt((double)O.stab[j] / stab_r);
            const double re = w * O.chi[g].x, im = w * O.chi[g].y;
            int e = 0;
            while (e < len && col[e] != j) e++;
            if (e < len) { val[e].x += re; val[e].y += im; }
            else { col[len] = j; val[len] = make_double2(re, im); len++; }
        }
        // the diagonal of a Hermitian matrix is real: contributions chi(g) + conj(chi(g)) from the two orientations add up
        // to a real number mathematically; drop the rounding residue so that the expansion sees an exactly Hermitian input
        val[0].y = 0.0;
        for (int x1 = 1; x1 < len; x1++) {
            const int64_t c = col[x1]; const double2 v = val[x1];
            int y1 = x1 - 1;
            while (y1 >= 0 && col[y1] > c) { col[y1 + 1] = col[y1]; val[y1 + 1] = val[y1]; y1--; }
            col[y1 + 1] = c; val[y1 + 1] = v;
        }
        row_end[r] = start[r] + len;
    }
}

__global__ void __launch_bounds__(kOBlock) orbit_len_kernel(int64_t n, const int64_t *start, const int64_t *end, int64_t *len)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) len[r] = end[r] - start[r];
}

static int ogrid(int64_t n) { int64_t g = (n + kOBlock - 1) / kOBlock; if (g < 1) g = 1; if (g > 148 * 64) g = 148 * 64; return (int)g; }

}  // namespace qb

using namespace qb;

extern "C" int qbgpu_build_heisenberg_orbit(qbgpu_matrix_t *A, int nsites, int ndown, int ntrans, const int32_t *perms, const double *chi_reim,
                                            int nbonds, const int32_t *bonds, double J, int flags, uint32_t *states_out, int64_t states_capacity)
{
    QB_TRY(ensure_init());
    Context &c = ctx();
    if (!A || !perms || !chi_reim || !bonds) return fail(QBGPU_ERR_ARG, "build_heisenberg_orbit: null argument");
    *A = nullptr;
    if (nsites < 2 || nsites > 32 || ndown < 0 || ndown > nsites) return fail(QBGPU_ERR_ARG, "build_heisenberg_orbit: 2..32 sites, 0 <= ndown <= nsites");
    if (ntrans < 1 || ntrans > kOMaxTrans) return fail(QBGPU_ERR_ARG, "build_heisenberg_orbit: 1..64 group elements");
    if (nbonds < 1 || nbonds > kOMaxTerms) return fail(QBGPU_ERR_ARG, "build_heisenberg_orbit: 1..128 bonds");
    for (int s = 0; s < nsites; s++) if (perms[s] != s) return fail(QBGPU_ERR_ARG, "build_heisenberg_orbit: element 0 must be the identity");
    OrbitDev O;
    memset(&O, 0, sizeof(O));
    O.nsites = nsites; O.ntrans = ntrans; O.ndown = ndown; O.shift = nsites > 16 ? nsites - 16 : 0;
    std::vector<uint32_t> lut((size_t)ntrans * 1024, 0u);
    for (int t = 0; t < ntrans; t++) {
        std::vector<char> seen(nsites, 0);
        for (int s = 0; s < nsites; s++) {
            const int p = perms[(size_t)t * nsites + s];
            if (p < 0 || p >= nsites || seen[p]) return fail(QBGPU_ERR_ARG, "build_heisenberg_orbit: not a permutation");
            seen[p] = 1;
        }
        for (int b = 0; b < 4; b++)
            for (int v = 0; v < 256; v++) {
                uint32_t img = 0;
                for (int bit = 0; bit < 8; bit++) { const int s = 8 * b + bit; if (s < nsites && ((v >> bit) & 1)) img |= 1u << perms[(size_t)t * nsites + s]; }
                lut[(size_t)t * 1024 + b * 256 + v] = img;
            }
        O.chi[t] = make_double2(chi_reim[2 * t], chi_reim[2 * t + 1]);
    }
    if (fabs(O.chi[0].x - 1.0) > 1e-12 || fabs(O.chi[0].y) > 1e-12) return fail(QBGPU_ERR_ARG, "build_heisenberg_orbit: chi(identity) must be 1");
    // enumerate over the smaller of (down spins, up spins); complementing keeps the patterns but reverses their order
    const bool complement = ndown > nsites - ndown;
    const int kbits = complement ? nsites - ndown : ndown;
    unsigned long long binom[33][17];
    for (int i = 0; i <= 32; i++)
        for (int k = 0; k <= 16; k++) {
            unsigned long long v;
            if (k == 0) v = 1; else if (i == 0) v = 0; else v = binom[i - 1][k - 1] + binom[i - 1][k];
            binom[i][k] = v;
        }
    QB_CUDA(cudaMemcpyToSymbol(c_binom, binom, sizeof(binom)));
    const int64_t ncand = (int64_t)binom[nsites][kbits];
    OrbitTerms T;
    memset(&T, 0, sizeof(T));
    T.nterms = nbonds; T.J = J;
    for (int t = 0; t < nbonds; t++) {
        const int p = bonds[2 * t], q = bonds[2 * t + 1];
        if (p < 0 || q < 0 || p >= nsites || q >= nsites || p == q) return fail(QBGPU_ERR_ARG, "build_heisenberg_orbit: bad bond");
        T.p[t] = (uint8_t)p; T.q[t] = (uint8_t)q;
    }

    uint32_t *d_lut = nullptr, *d_keys = nullptr, *d_seg = nullptr, *d_cand = nullptr, *d_sel = nullptr;
    uint8_t *d_flag = nullptr, *d_stab = nullptr;
    int *d_nsel = nullptr;
    OrbitTerms *d_T = nullptr;
    int64_t *d_cap = nullptr, *d_start = nullptr, *d_end = nullptr, *d_col = nullptr;
    double2 *d_val = nullptr;
    void *d_tmp = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_lut); cudaFree(d_keys); cudaFree(d_seg); cudaFree(d_cand); cudaFree(d_sel); cudaFree(d_flag); cudaFree(d_stab); cudaFree(d_nsel);
        cudaFree(d_T); cudaFree(d_cap); cudaFree(d_start); cudaFree(d_end); cudaFree(d_col); cudaFree(d_val); cudaFree(d_tmp);
    };
#define QB_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return cuda_fail(e_, #call, __FILE__, __LINE__); } } while (0)
    QB_CU(cudaMalloc(&d_lut, lut.size() * 4));
    QB_CU(cudaMemcpyAsync(d_lut, lut.data(), lut.size() * 4, cudaMemcpyHostToDevice, c.stream));
    O.lut = d_lut;
    // representatives, in batches of candidates (a batch is flagged, compacted and appended)
    const int64_t batch = 1LL << 27;
    std::vector<uint32_t> dummy;
    int64_t n = 0, capacity = std::max<int64_t>(1024, ncand / std::max(1, ntrans) * 5 / 4 + 4096);
    QB_CU(cudaMalloc(&d_keys, sizeof(uint32_t) * (size_t)capacity));
    QB_CU(cudaMalloc(&d_flag, (size_t)std::min(batch, ncand)));
    QB_CU(cudaMalloc(&d_cand, sizeof(uint32_t) * (size_t)std::min(batch, ncand)));
    QB_CU(cudaMalloc(&d_sel, sizeof(uint32_t) * (size_t)std::min(batch, ncand)));
    QB_CU(cudaMalloc(&d_nsel, sizeof(int)));
    size_t tb = 0;
    QB_CU(cub::DeviceSelect::Flagged(nullptr, tb, d_cand, d_flag, d_sel, d_nsel, (int)std::min(batch, ncand), c.stream));
    QB_CU(cudaMalloc(&d_tmp, tb ? tb : 1));
    for (int64_t first = 0; first < ncand; first += batch) {
        const int64_t cnt = std::min(batch, ncand - first);
        orbit_flag_kernel<<<ogrid(cnt), kOBlock, 0, c.stream>>>(O, first, cnt, kbits, complement, d_flag, d_cand);
        QB_LAUNCH_COUNT();
        QB_CU(cub::DeviceSelect::Flagged(d_tmp, tb, d_cand, d_flag, d_sel, d_nsel, (int)cnt, c.stream));
        int nsel = 0;
        QB_CU(cudaMemcpyAsync(&nsel, d_nsel, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        QB_CU(cudaStreamSynchronize(c.stream));
        if (n + nsel > capacity) {                          // orbits shorter than the group: grow
            const int64_t ncap = std::max(capacity * 2, n + nsel);
            uint32_t *nk = nullptr;
            QB_CU(cudaMalloc(&nk, sizeof(uint32_t) * (size_t)ncap));
            QB_CU(cudaMemcpyAsync(nk, d_keys, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, c.stream));
            QB_CU(cudaStreamSynchronize(c.stream));
            cudaFree(d_keys); d_keys = nk; capacity = ncap;
        }
        QB_CU(cudaMemcpyAsync(d_keys + n, d_sel, sizeof(uint32_t) * (size_t)nsel, cudaMemcpyDeviceToDevice, c.stream));
        n += nsel;
    }
    QB_CU(cudaStreamSynchronize(c.stream));
    cudaFree(d_flag); d_flag = nullptr; cudaFree(d_cand); d_cand = nullptr; cudaFree(d_sel); d_sel = nullptr; cudaFree(d_tmp); d_tmp = nullptr;
    if (n == 0) { cleanup(); return fail(QBGPU_ERR_ARG, "build_heisenberg_orbit: empty sector"); }
    if (complement) {                                       // complemented enumeration came out in descending order
        std::vector<uint32_t> h((size_t)n);
        QB_CU(cudaMemcpy(h.data(), d_keys, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost));
        std::reverse(h.begin(), h.end());
        QB_CU(cudaMemcpy(d_keys, h.data(), sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice));
    }
    O.keys = d_keys; O.n = n;
    const uint32_t nseg = 1u << (nsites - O.shift);
    QB_CU(cudaMalloc(&d_seg, sizeof(uint32_t) * ((size_t)nseg + 1)));
    orbit_seg_kernel<<<ogrid(nseg + 1), kOBlock, 0, c.stream>>>(n, d_keys, O.shift, nseg, d_seg);
    QB_LAUNCH_COUNT();
    O.seg = d_seg;
    QB_CU(cudaMalloc(&d_stab, (size_t)n));
    orbit_stab_kernel<<<ogrid(n), kOBlock, 0, c.stream>>>(O, d_stab);
    QB_LAUNCH_COUNT();
    O.stab = d_stab;
    if (states_out) {
        if (states_capacity < n) { cleanup(); return fail(QBGPU_ERR_ARG, "build_heisenberg_orbit: states_out too small"); }
        QB_CU(cudaMemcpyAsync(states_out, d_keys, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, c.stream));
    }
    // rows: upper triangle with per-row capacity = 1 + antiparallel bonds
    QB_CU(cudaMalloc(&d_T, sizeof(OrbitTerms)));
    QB_CU(cudaMemcpyAsync(d_T, &T, sizeof(OrbitTerms), cudaMemcpyHostToDevice, c.stream));
    QB_CU(cudaMalloc(&d_cap, sizeof(int64_t) * (n + 1)));
    QB_CU(cudaMalloc(&d_start, sizeof(int64_t) * (n + 1)));
    QB_CU(cudaMalloc(&d_end, sizeof(int64_t) * (n + 1)));
    orbit_cap_kernel<<<ogrid(n + 1), kOBlock, 0, c.stream>>>(O, d_T, d_cap);
    QB_LAUNCH_COUNT();
    QB_CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_cap, d_start, n + 1, c.stream));
    QB_CU(cudaMalloc(&d_tmp, tb ? tb : 1));
    QB_CU(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_cap, d_start, n + 1, c.stream));
    int64_t total = 0;
    QB_CU(cudaMemcpyAsync(&total, d_start + n, sizeof(int64_t), cudaMemcpyDeviceToHost, c.stream));
    QB_CU(cudaStreamSynchronize(c.stream));
    QB_CU(cudaMalloc(&d_col, sizeof(int64_t) * (size_t)total));
    QB_CU(cudaMalloc(&d_val, sizeof(double2) * (size_t)total));
    orbit_rows_kernel<<<ogrid(n), kOBlock, 0, c.stream>>>(O, d_T, d_start, d_end, d_col, d_val);
    QB_LAUNCH_COUNT();
    orbit_len_kernel<<<ogrid(n), kOBlock, 0, c.stream>>>(n, d_start, d_end, d_cap);
    QB_LAUNCH_COUNT();
    cudaFree(d_tmp); d_tmp = nullptr;
    int64_t *d_sum = nullptr;
    QB_CU(cudaMalloc(&d_sum, sizeof(int64_t)));
    cudaError_t e1 = cub::DeviceReduce::Sum(nullptr, tb, d_cap, d_sum, n, c.stream);
    if (e1 == cudaSuccess) e1 = cudaMalloc(&d_tmp, tb ? tb : 1);
    if (e1 == cudaSuccess) e1 = cub::DeviceReduce::Sum(d_tmp, tb, d_cap, d_sum, n, c.stream);
    int64_t upper = 0;
    if (e1 == cudaSuccess) e1 = cudaMemcpyAsync(&upper, d_sum, sizeof(int64_t), cudaMemcpyDeviceToHost, c.stream);
    if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(c.stream);
    cudaFree(d_sum);
    QB_CU(e1);
    QB_CU(cudaGetLastError());
#undef QB_CU
    int rc = create_from_device_csr(A, n, d_start, d_end, d_col, d_val, true, upper, 1, flags, true);
    cleanup();
    return rc;
}
