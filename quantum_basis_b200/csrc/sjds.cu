// quantum_basis_b200/csrc/sjds.cu -- the "sliced jagged" H*v kernel and its in-place format conversion.
//
// Why a second layout: in the reference's Lin-table order (src/basis.cc:1144-1190) consecutive rows are basis states
// that differ only in the even-site label, so a given Hamiltonian term sends consecutive rows to (nearly) consecutive
// columns.  The CSR-vector kernel assigns LANES threads to ONE row, so the 32 gathers of a warp instruction hit 32
// unrelated cache lines; assigning one thread per row makes the gathers of a warp fall into a few contiguous
// segments of x (coalesced 128-byte lines instead of 32 separate sectors), which is what bounds the kernel once the
// matrix stream itself is coalesced.  Thread-per-row over plain CSR would un-coalesce the matrix stream, and ELL/SELL
// padding would add bytes, so the entries of every 32-row slice are re-ordered jagged-diagonal-wise:
//     rank the slice's rows by length (descending);  for k = 0,1,2,...: the k-th entry of every row with length > k,
//     in rank order.
// The rows active at step k are always ranks [0, cnt_k), so lane `rank` reads entry off_k + rank: contiguous,
// unpadded, and off_{k+1} = off_k + popc(ballot(k < len)).  A slice occupies exactly the CSR range of its 32 rows,
// so the conversion is an in-place (chunk-buffered) permutation and the algorithmic bytes do not change
// (4 bytes of rowinfo per row replace nothing: rowptr is still read once per slice).
#include "internal.hpp"
#include <algorithm>
#include <vector>

namespace qb {

constexpr int kSBlock = 256;
constexpr uint32_t kLenMask = 0xFFFFFFu;

template <typename ValT, typename VecT, bool DOTS>
__global__ void __launch_bounds__(kSBlock)
spmv_sjds_kernel(int64_t nslices, int64_t nrows, int64_t row_lo, const int64_t *__restrict__ rowptr,
                 const uint32_t *__restrict__ rowinfo, const int32_t *__restrict__ col, const ValT *__restrict__ val,
                 const VecT *__restrict__ x, const VecT *z, VecT *y, double2 alpha, double2 gamma, double2 beta,
                 int scal_mode, const double *__restrict__ sc, double *dots_out, double *partials, unsigned *ticket)
{
    using VT = VecTraits<VecT>;
    constexpr int WPB = kSBlock / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double dot_scale = 1.0;
    if (scal_mode == 1) {
        const double sx = sc[0], sz = sc[1], bprev = sc[2];
        alpha = make_double2(sx, 0.0);
        gamma = make_double2(0.0, 0.0);
        beta = make_double2(-bprev * sz, 0.0);
        dot_scale = sx;
    }
    const bool use_gamma = (gamma.x != 0.0 || gamma.y != 0.0);
    const bool use_beta = (beta.x != 0.0 || beta.y != 0.0);
    double d[3] = {0.0, 0.0, 0.0};

    for (int64_t s = (int64_t)blockIdx.x * WPB + warp; s < nslices; s += (int64_t)gridDim.x * WPB) {
        const uint32_t info = rowinfo[s * 32 + lane];
        const int len = (int)(info & kLenMask);
        const int64_t row = s * 32 + (info >> 24);
        const int maxlen = __shfl_sync(0xffffffffu, len, 0);
        int64_t off = rowptr[s * 32] + lane;               // this lane's entry at step k is off_k + lane
        VecT acc0 = VT::zero(), acc1 = VT::zero();
        int k = 0;
        for (; k + 4 <= maxlen; k += 4) {                  // 4 jagged diagonals per trip: 12 independent loads per lane
            const bool a0 = k < len, a1 = k + 1 < len, a2 = k + 2 < len, a3 = k + 3 < len;
            const int n0 = __popc(__ballot_sync(0xffffffffu, a0)), n1 = __popc(__ballot_sync(0xffffffffu, a1));
            const int n2 = __popc(__ballot_sync(0xffffffffu, a2)), n3 = __popc(__ballot_sync(0xffffffffu, a3));
            const int64_t o0 = off, o1 = o0 + n0, o2 = o1 + n1, o3 = o2 + n2;
            int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
            ValT v0{}, v1{}, v2{}, v3{};
            if (a0) { c0 = ld_stream(col + o0); v0 = ld_stream(val + o0); }
            if (a1) { c1 = ld_stream(col + o1); v1 = ld_stream(val + o1); }
            if (a2) { c2 = ld_stream(col + o2); v2 = ld_stream(val + o2); }
            if (a3) { c3 = ld_stream(col + o3); v3 = ld_stream(val + o3); }
            if (a0) mac(acc0, v0, ld_vec(x + c0));
            if (a1) mac(acc1, v1, ld_vec(x + c1));
            if (a2) mac(acc0, v2, ld_vec(x + c2));
            if (a3) mac(acc1, v3, ld_vec(x + c3));
            off = o3 + n3;
        }
        for (; k < maxlen; k++) {
            const bool a0 = k < len;
            const int n0 = __popc(__ballot_sync(0xffffffffu, a0));
            if (a0) { const int c0 = ld_stream(col + off); const ValT v0 = ld_stream(val + off); mac(acc0, v0, ld_vec(x + c0)); }
            off += n0;
        }
        if (len > 0) {                                      // padding ranks of the last slice have length 0
            const VecT acc = VT::add(acc0, acc1);
            VecT out = VT::scale(alpha, acc);
            VecT xi = VT::zero();
            if (use_gamma || DOTS) xi = x[row_lo + row];
            if (use_gamma) out = VT::add(out, VT::scale(gamma, xi));
            if (use_beta) out = VT::add(out, VT::scale(beta, z[row]));
            y[row] = out;
            if (DOTS) {
                const double2 p = VT::conj_mul(xi, out);
                d[0] += p.x; d[1] += p.y; d[2] += VT::abs2(out);
            }
        }
    }
    if (DOTS) {
        d[0] *= dot_scale; d[1] *= dot_scale;
        block_reduce_finalize<3, kSBlock>(d, partials, ticket, dots_out);
    }
}

template <typename ValT, typename VecT, bool DOTS>
static int launch_sjds_variant(const qbgpu_matrix *A, const FusedArgs &a)
{
    Context &c = ctx();
    auto kern = spmv_sjds_kernel<ValT, VecT, DOTS>;
    static int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        QB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kSBlock, 0));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    const int64_t nrows = A->nrows();
    if (nrows == 0) return QBGPU_OK;
    const int64_t nslices = (nrows + 31) / 32;
    constexpr int WPB = kSBlock / 32;
    int64_t want = (nslices + WPB - 1) / WPB;
    int64_t cap = (int64_t)c.num_sms * blocks_per_sm;
    if (cap > kMaxPartialBlocks) cap = kMaxPartialBlocks;
    const int grid = (int)(want < cap ? want : cap);
    kern<<<grid, kSBlock, 0, c.stream>>>(nslices, nrows, A->row_lo, A->rowptr, A->rowinfo, A->col, (const ValT *)A->val,
                                         (const VecT *)a.x, (const VecT *)a.z, (VecT *)a.y, a.alpha, a.gamma, a.beta,
                                         a.scal_mode, a.sc, a.dots, c.partials, c.ticket);
    QB_LAUNCH_COUNT();
    QB_CUDA(cudaGetLastError());
    return QBGPU_OK;
}

int launch_spmv_sjds(const qbgpu_matrix *A, const FusedArgs &a)
{
    const bool dots = a.dots != nullptr;
    if (!A->api_complex) return dots ? launch_sjds_variant<double, double, true>(A, a) : launch_sjds_variant<double, double, false>(A, a);
    if (A->val_real) return dots ? launch_sjds_variant<double, double2, true>(A, a) : launch_sjds_variant<double, double2, false>(A, a);
    return dots ? launch_sjds_variant<double2, double2, true>(A, a) : launch_sjds_variant<double2, double2, false>(A, a);
}

// ------------------------------------------------------------------------------------------ format conversion
// One warp per slice.  forward: CSR order -> jagged order (and fills rowinfo); backward: the inverse.
// `tmp` receives the re-ordered entries of slices [s0, s1) at offset (position - base0).
template <typename ValT, bool FORWARD>
__global__ void __launch_bounds__(kSBlock)
sjds_permute_kernel(int64_t s0, int64_t s1, int64_t nrows, const int64_t *__restrict__ rowptr, uint32_t *rowinfo,
                    const int32_t *__restrict__ col, const ValT *__restrict__ val, int64_t base0, int32_t *tcol, ValT *tval,
                    int *too_long)
{
    constexpr int WPB = kSBlock / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t s = s0 + (int64_t)blockIdx.x * WPB + warp; s < s1; s += (int64_t)gridDim.x * WPB) {
        const int64_t r = s * 32 + lane;
        const int64_t rs = r < nrows ? rowptr[r] : 0;
        const int64_t len64 = r < nrows ? rowptr[r + 1] - rs : 0;
        if (__any_sync(0xffffffffu, len64 > (int64_t)kLenMask)) { if (lane == 0) atomicExch(too_long, 1); continue; }   // warp-uniform
        const int len = (int)len64;
        int rank = 0, maxlen = 0;
        for (int j = 0; j < 32; j++) {                      // rank by (length desc, lane asc): a stable sort of 32 keys
            const int lj = __shfl_sync(0xffffffffu, len, j);
            rank += (lj > len) || (lj == len && j < lane);
            maxlen = max(maxlen, lj);
        }
        if (FORWARD) rowinfo[s * 32 + rank] = ((uint32_t)lane << 24) | (uint32_t)len;
        int64_t off = rowptr[s * 32] - base0;
        for (int k = 0; k < maxlen; k++) {
            const bool act = k < len;
            const int cnt = __popc(__ballot_sync(0xffffffffu, act));
            if (act) {
                const int64_t jag = off + rank, csr = rs + k;           // jag is relative to base0, csr absolute
                if (FORWARD) { tcol[jag] = col[csr]; tval[jag] = val[csr]; }
                else         { tcol[csr - base0] = col[base0 + jag]; tval[csr - base0] = val[base0 + jag]; }
            }
            off += cnt;
        }
    }
}

template <typename ValT>
static int sjds_convert_typed(qbgpu_matrix *A, bool forward)
{
    Context &c = ctx();
    const int64_t nrows = A->nrows();
    const int64_t nslices = (nrows + 31) / 32;
    if (nslices == 0) return QBGPU_OK;
    if (forward && !A->rowinfo) QB_CUDA(cudaMalloc(&A->rowinfo, sizeof(uint32_t) * nslices * 32));
    // batches of slices re-ordered through a bounded temporary, then copied back over the same range
    const int64_t batch = 1 << 16;                          // 65,536 slices = 2,097,152 rows per batch
    const int64_t nb = (nslices + batch - 1) / batch;
    std::vector<int64_t> bidx(nb + 1), bofs(nb + 1);
    for (int64_t b = 0; b <= nb; b++) bidx[b] = std::min(nrows, b * batch * 32);
    int64_t *d_idx = nullptr, *d_ofs = nullptr;
    int *d_flag = nullptr;
    int32_t *tcol = nullptr;
    ValT *tval = nullptr;
    auto cleanup = [&]() { cudaFree(d_idx); cudaFree(d_ofs); cudaFree(d_flag); cudaFree(tcol); cudaFree(tval); };
#define QB_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return cuda_fail(e_, #call, __FILE__, __LINE__); } } while (0)
    // boundary offsets rowptr[bidx[b]]
    for (int64_t b = 0; b <= nb; b++) QB_CU(cudaMemcpyAsync(&bofs[b], A->rowptr + bidx[b], sizeof(int64_t), cudaMemcpyDeviceToHost, c.stream));
    QB_CU(cudaStreamSynchronize(c.stream));
    int64_t maxspan = 1;
    for (int64_t b = 0; b < nb; b++) maxspan = std::max(maxspan, bofs[b + 1] - bofs[b]);
    QB_CU(cudaMalloc(&tcol, sizeof(int32_t) * maxspan));
    QB_CU(cudaMalloc(&tval, sizeof(ValT) * maxspan));
    QB_CU(cudaMalloc(&d_flag, sizeof(int)));
    QB_CU(cudaMemsetAsync(d_flag, 0, sizeof(int), c.stream));
    for (int64_t b = 0; b < nb; b++) {
        const int64_t s0 = b * batch, s1 = std::min(nslices, (b + 1) * batch);
        const int64_t span = bofs[b + 1] - bofs[b];
        if (span == 0) continue;
        const int grid = (int)std::min<int64_t>((s1 - s0 + 7) / 8, 148 * 16);
        if (forward) sjds_permute_kernel<ValT, true><<<grid, kSBlock, 0, c.stream>>>(s0, s1, nrows, A->rowptr, A->rowinfo, A->col, (const ValT *)A->val, bofs[b], tcol, tval, d_flag);
        else         sjds_permute_kernel<ValT, false><<<grid, kSBlock, 0, c.stream>>>(s0, s1, nrows, A->rowptr, A->rowinfo, A->col, (const ValT *)A->val, bofs[b], tcol, tval, d_flag);
        QB_LAUNCH_COUNT();
        QB_CU(cudaMemcpyAsync(A->col + bofs[b], tcol, sizeof(int32_t) * span, cudaMemcpyDeviceToDevice, c.stream));
        QB_CU(cudaMemcpyAsync((ValT *)A->val + bofs[b], tval, sizeof(ValT) * span, cudaMemcpyDeviceToDevice, c.stream));
    }
    int flag = 0;
    QB_CU(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    QB_CU(cudaStreamSynchronize(c.stream));
    QB_CU(cudaGetLastError());
    cleanup();
#undef QB_CU
    if (flag) return fail(QBGPU_ERR_STATE, "sliced-jagged layout: a row is longer than 2^24-1 entries");
    A->format = forward ? QBGPU_FORMAT_SELL : QBGPU_FORMAT_CSR;
    return QBGPU_OK;
}

int sjds_convert(qbgpu_matrix *A, bool forward)
{
    if (forward == (A->format == QBGPU_FORMAT_SELL)) return QBGPU_OK;
    return A->val_real ? sjds_convert_typed<double>(A, forward) : sjds_convert_typed<double2>(A, forward);
}

}  // namespace qb
