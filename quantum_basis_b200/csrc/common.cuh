// quantum_basis_b200/csrc/common.cuh -- shared device/host helpers for libqbgpu (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace qb {

// ---------------------------------------------------------------------------------------------- errors
void set_error(const std::string &msg);
int  fail(int code, const std::string &msg);
int  cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define QB_CUDA(call)                                                                     \
    do {                                                                                  \
        cudaError_t qb_e_ = (call);                                                       \
        if (qb_e_ != cudaSuccess) return qb::cuda_fail(qb_e_, #call, __FILE__, __LINE__); \
    } while (0)
#define QB_TRY(call)                     \
    do {                                 \
        int qb_rc_ = (call);             \
        if (qb_rc_ != 0) return qb_rc_;  \
    } while (0)

// --------------------------------------------------------------------------------------------- context
constexpr int kMaxPartialBlocks = 148 * 16;   // upper bound on the grid of any reducing kernel
constexpr int kDotSlots = 4;                  // partial sums carried per reducing kernel

struct Context {
    int          device = -1;
    cudaStream_t own_stream = nullptr;   // created at init
    cudaStream_t stream = nullptr;       // stream in use (own_stream or the caller's)
    cudaStream_t copy_stream = nullptr;  // D2H/H2D overlap for host-pointer products
    int          num_sms = 148;
    double      *partials = nullptr;     // [kMaxPartialBlocks * kDotSlots] per-block partial sums
    unsigned    *ticket = nullptr;       // last-block-done counter (self-resetting)
    double      *scal_dev = nullptr;     // small device scalar scratch (64 doubles)
    double      *scal_host = nullptr;    // pinned mirror
    void        *stage_x = nullptr, *stage_y = nullptr;   // device staging for host-pointer products
    size_t       stage_x_bytes = 0, stage_y_bytes = 0;
    long long    launches = 0;
};
Context &ctx();
int ensure_init();

#define QB_LAUNCH_COUNT() (qb::ctx().launches++)

// -------------------------------------------------------------------------------------- complex algebra
// (host + device: species.cu runs its per-row functions on the host too, for the CPU tests of the index logic)
struct cplx { double x, y; };   // layout-compatible with double2 / std::complex<double>

__host__ __device__ inline double2 make_c(double re, double im) { return make_double2(re, im); }
__host__ __device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__host__ __device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

// fma-style accumulate: acc += v * x for the (ValT, VecT) combinations used
__host__ __device__ __forceinline__ void mac(double &acc, double v, double x) { acc = fma(v, x, acc); }
__host__ __device__ __forceinline__ void mac(double2 &acc, double v, double2 x) { acc.x = fma(v, x.x, acc.x); acc.y = fma(v, x.y, acc.y); }
__host__ __device__ __forceinline__ void mac(double2 &acc, double2 v, double2 x)
{
    acc.x = fma(v.x, x.x, acc.x); acc.x = fma(-v.y, x.y, acc.x);
    acc.y = fma(v.x, x.y, acc.y); acc.y = fma(v.y, x.x, acc.y);
}

__host__ __device__ __forceinline__ int popc_hd(uint32_t v)
{
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
// read-only load: the non-coherent path on the device, a plain load on the host
template <typename T> __host__ __device__ __forceinline__ T ld_ro(const T *p)
{
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

template <typename V> struct VecTraits;
template <> struct VecTraits<double> {
    static constexpr int ncomp = 1;
    __host__ __device__ static __forceinline__ double zero() { return 0.0; }
    __host__ __device__ static __forceinline__ double add(double a, double b) { return a + b; }
    // s*a with complex scalar s: real vectors only use the real part
    __host__ __device__ static __forceinline__ double scale(double2 s, double a) { return s.x * a; }
    __host__ __device__ static __forceinline__ double rscale(double s, double a) { return s * a; }
    __host__ __device__ static __forceinline__ double2 conj_mul(double a, double b) { return make_double2(a * b, 0.0); }   // conj(a)*b
    __host__ __device__ static __forceinline__ double abs2(double a) { return a * a; }
    __device__ static __forceinline__ double shfl_xor(double a, int m, int w, unsigned mask) { return __shfl_xor_sync(mask, a, m, w); }
};
template <> struct VecTraits<double2> {
    static constexpr int ncomp = 2;
    __host__ __device__ static __forceinline__ double2 zero() { return make_double2(0.0, 0.0); }
    __host__ __device__ static __forceinline__ double2 add(double2 a, double2 b) { return cadd(a, b); }
    __host__ __device__ static __forceinline__ double2 scale(double2 s, double2 a) { return cmul(s, a); }
    __host__ __device__ static __forceinline__ double2 rscale(double s, double2 a) { return make_double2(s * a.x, s * a.y); }
    __host__ __device__ static __forceinline__ double2 conj_mul(double2 a, double2 b) { return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }
    __host__ __device__ static __forceinline__ double abs2(double2 a) { return a.x * a.x + a.y * a.y; }
    __device__ static __forceinline__ double2 shfl_xor(double2 a, int m, int w, unsigned mask)
    { return make_double2(__shfl_xor_sync(mask, a.x, m, w), __shfl_xor_sync(mask, a.y, m, w)); }
};

// ------------------------------------------------------------------------------------------ cache hints
// matrix streams are read exactly once per product: stream them past L1 and mark them evict-first in L2 so the
// gathered vector keeps the cache; vectors use the default (read-only) path.
__device__ __forceinline__ int    ld_stream(const int *p)    { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ double2 ld_stream(const double2 *p) { return __ldcs(p); }
__device__ __forceinline__ double  ld_vec(const double *p)  { return __ldg(p); }
__device__ __forceinline__ double2 ld_vec(const double2 *p) { return __ldg(p); }

// --------------------------------------------------------------------------- deterministic block reduction
// Each block reduces NS running sums, writes them to partials[block*kDotSlots + s]; the last block to finish
// (ticket) adds the per-block partials in block order and stores the totals to out[0..NS).  Result is
// bit-reproducible for a fixed grid.  Must be called by all threads of the block.
template <int NS, int BLOCK>
__device__ __forceinline__ void block_reduce_finalize(double (&v)[NS], double *partials, unsigned *ticket, double *out)
{
    __shared__ double sm[NS][BLOCK / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        double a = v[s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) sm[s][warp] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) {
            double a = 0.0;
            for (int w = 0; w < BLOCK / 32; w++) a += sm[s][w];
            partials[blockIdx.x * kDotSlots + s] = a;
        }
        __threadfence();
        unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        // fixed-shape tree over the per-block partials: thread t adds blocks t, t+BLOCK, ...; then warp and
        // block trees.  Deterministic for a fixed (grid, BLOCK).
        double acc[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) acc[s] = 0.0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += BLOCK) {
#pragma unroll
            for (int s = 0; s < NS; s++) acc[s] += __ldcg(&partials[b * kDotSlots + s]);
        }
#pragma unroll
        for (int s = 0; s < NS; s++) {
            double a = acc[s];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) sm[s][warp] = a;
        }
        __syncthreads();
        if (threadIdx.x < NS) {
            double a = 0.0;
            for (int w = 0; w < BLOCK / 32; w++) a += sm[threadIdx.x][w];
            out[threadIdx.x] = a;
        }
        if (threadIdx.x == 0) *ticket = 0u;     // self-reset for the next reducing kernel on this stream
    }
}

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace qb
