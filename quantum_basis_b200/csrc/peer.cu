// quantum_basis_b200/csrc/peer.cu -- peer-memory exchange of the Krylov vector over NVLink (one process per GPU).
//
// The reference has no distributed mode; SURVEY section 8(e) asks for the gathered vector to arrive while the local
// block is being multiplied.  Instead of a collective, every rank maps its peers' vector buffers (CUDA IPC) and PULLS
// the slices it needs with copy-engine transfers on one stream per peer; the product of column block p is ordered
// behind the arrival of slice p with an event, so transfers and products overlap block by block and all seven
// incoming NVLink paths of a GPU are busy at once.  Cross-process ordering (a slice is final before it is pulled, and
// is not overwritten while it is being pulled) comes from the collectives the Krylov loop already contains (the
// all-reduce of the Lanczos scalars) or from an explicit barrier in the plain product loop, plus ping-pong buffers.
#include "internal.hpp"
#include <cstring>

namespace qb {

constexpr int kPeerLanes = 16;
struct PeerState {
    cudaStream_t lane[kPeerLanes] = {};
    cudaEvent_t arrived[kPeerLanes] = {};
    cudaEvent_t fence = nullptr;
    int *flags = nullptr;              // [kPeerLanes + 1] arrival flags by ring distance; last entry: a waiter timed out
    int *one = nullptr;                // device constant 1, the source of the flag-setting copies
    bool ready = false;
};
static thread_local PeerState g_peer;

// SM-driven pull: a few CTAs read the peer's slice through the NVLink mapping with 16-byte volatile-cached loads (L1 is
// not trusted across kernels for peer memory; nothing is kept) and store it locally.  Sixteen independent loads per
// thread and trip keep enough requests in flight to cover the NVLink round trip; the lanes are highest-priority streams
// so that these CTAs take SM slots ahead of the queued CTAs of the product that runs beside them.
constexpr int kPullThreads = 256;
constexpr int kPullUnroll = 16;            // 16 x 16 B per thread = 64 KB in flight per CTA (NVLink round trip ~2.5 us)
__global__ void __launch_bounds__(kPullThreads) peer_pull_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src, size_t n16)
{
    constexpr size_t kTile = (size_t)kPullThreads * kPullUnroll;
    for (size_t base = (size_t)blockIdx.x * kTile; base < n16; base += (size_t)gridDim.x * kTile) {
        uint4 r[kPullUnroll];
#pragma unroll
        for (int u = 0; u < kPullUnroll; u++) {
            const size_t i = base + (size_t)u * kPullThreads + threadIdx.x;
            if (i < n16) r[u] = __ldcv(src + i);
        }
#pragma unroll
        for (int u = 0; u < kPullUnroll; u++) {
            const size_t i = base + (size_t)u * kPullThreads + threadIdx.x;
            if (i < n16) __stcs(dst + i, r[u]);
        }
    }
}

static int peer_init()
{
    if (g_peer.ready) return QBGPU_OK;
    int prio_lo = 0, prio_hi = 0;
    QB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    for (int i = 0; i < kPeerLanes; i++) {
        QB_CUDA(cudaStreamCreateWithPriority(&g_peer.lane[i], cudaStreamNonBlocking, prio_hi));
        QB_CUDA(cudaEventCreateWithFlags(&g_peer.arrived[i], cudaEventDisableTiming));
    }
    QB_CUDA(cudaEventCreateWithFlags(&g_peer.fence, cudaEventDisableTiming));
    QB_CUDA(cudaMalloc(&g_peer.flags, sizeof(int) * (kPeerLanes + 1)));
    QB_CUDA(cudaMalloc(&g_peer.one, sizeof(int)));
    QB_CUDA(cudaMemset(g_peer.flags, 0, sizeof(int) * (kPeerLanes + 1)));
    const int h_one = 1;
    QB_CUDA(cudaMemcpy(g_peer.one, &h_one, sizeof(int), cudaMemcpyHostToDevice));
    g_peer.ready = true;
    return QBGPU_OK;
}

// Two-step form of qbgpu_peer_pull_async for callers that want to launch kernels BETWEEN the point the pulls are ordered after
// and the (host-side) enqueueing of the pulls: peer_mark_fence() records that point on the compute stream, peer_pull_fenced()
// enqueues a pull ordered after it -- not after whatever was launched on the compute stream since (csrc/dist.cu: the part of
// the product that needs no remote data is launched first and runs while the host is still enqueueing the transfers).
int peer_mark_fence()
{
    QB_TRY(peer_init());
    QB_CUDA(cudaEventRecord(g_peer.fence, ctx().stream));
    return QBGPU_OK;
}
int peer_pull_fenced(int lane, int slot, void *dst_local, const void *src_peer, size_t bytes)
{
    if (lane < 0 || lane >= kPeerLanes || slot < 0 || slot >= kPeerLanes || !dst_local || !src_peer) return fail(QBGPU_ERR_ARG, "peer_pull: bad argument");
    QB_CUDA(cudaStreamWaitEvent(g_peer.lane[lane], g_peer.fence, 0));
    QB_CUDA(cudaMemcpyAsync(dst_local, src_peer, bytes, cudaMemcpyDefault, g_peer.lane[lane]));
    QB_CUDA(cudaEventRecord(g_peer.arrived[slot], g_peer.lane[lane]));
    return QBGPU_OK;
}

int peer_record_on_lane(int lane, cudaEvent_t ev)
{
    QB_TRY(peer_init());
    if (lane < 0 || lane >= kPeerLanes) return fail(QBGPU_ERR_ARG, "peer_record_on_lane: bad lane");
    QB_CUDA(cudaEventRecord(ev, g_peer.lane[lane]));
    return QBGPU_OK;
}

int peer_ring_flags(const int **flags, int **timeout_flag)
{
    QB_TRY(peer_init());
    *flags = g_peer.flags; *timeout_flag = g_peer.flags + kPeerLanes;
    return QBGPU_OK;
}

}  // namespace qb

using namespace qb;

extern "C" {

int qbgpu_ipc_export(void *dptr, void *handle64)
{
    QB_TRY(ensure_init());
    if (!dptr || !handle64) return fail(QBGPU_ERR_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    QB_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, dptr));
    return QBGPU_OK;
}

int qbgpu_ipc_open(const void *handle64, void **peer_ptr)
{
    QB_TRY(ensure_init());
    if (!handle64 || !peer_ptr) return fail(QBGPU_ERR_ARG, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    QB_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return QBGPU_OK;
}

int qbgpu_ipc_close(void *peer_ptr)
{
    if (peer_ptr) QB_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return QBGPU_OK;
}

/* Enqueue, on copy stream `lane`, a pull of `bytes` from a peer-mapped (or local) source into local memory, ordered
 * after everything issued so far on the compute stream; its arrival is recorded in event `slot` for qbgpu_peer_wait.
 * Pulls on one lane run one after the other (each at the full NVLink rate), so the slices arrive in the order the
 * column blocks are multiplied; several lanes keep more than one transfer in flight. */
int qbgpu_peer_pull_async(int lane, int slot, void *dst_local, const void *src_peer, size_t bytes)
{
    QB_TRY(ensure_init());
    QB_TRY(peer_init());
    if (lane < 0 || lane >= kPeerLanes || slot < 0 || slot >= kPeerLanes || !dst_local || !src_peer) return fail(QBGPU_ERR_ARG, "peer_pull: bad argument");
    Context &c = ctx();
    QB_CUDA(cudaEventRecord(g_peer.fence, c.stream));
    QB_CUDA(cudaStreamWaitEvent(g_peer.lane[lane], g_peer.fence, 0));
    QB_CUDA(cudaMemcpyAsync(dst_local, src_peer, bytes, cudaMemcpyDefault, g_peer.lane[lane]));
    QB_CUDA(cudaEventRecord(g_peer.arrived[slot], g_peer.lane[lane]));
    return QBGPU_OK;
}

/* The same pull done by `ctas` thread blocks instead of a copy engine (bytes and both pointers multiples of 16). */
int qbgpu_peer_pull_sm(int lane, int slot, void *dst_local, const void *src_peer, size_t bytes, int ctas)
{
    QB_TRY(ensure_init());
    QB_TRY(peer_init());
    if (lane < 0 || lane >= kPeerLanes || slot < 0 || slot >= kPeerLanes || !dst_local || !src_peer || ctas < 1) return fail(QBGPU_ERR_ARG, "peer_pull_sm: bad argument");
    if ((bytes & 15) || ((uintptr_t)dst_local & 15) || ((uintptr_t)src_peer & 15)) return fail(QBGPU_ERR_ARG, "peer_pull_sm: 16-byte alignment required");
    Context &c = ctx();
    QB_CUDA(cudaEventRecord(g_peer.fence, c.stream));
    QB_CUDA(cudaStreamWaitEvent(g_peer.lane[lane], g_peer.fence, 0));
    if (bytes) {
        peer_pull_kernel<<<ctas, kPullThreads, 0, g_peer.lane[lane]>>>((uint4 *)dst_local, (const uint4 *)src_peer, bytes / 16);
        QB_LAUNCH_COUNT();
    }
    QB_CUDA(cudaEventRecord(g_peer.arrived[slot], g_peer.lane[lane]));
    return QBGPU_OK;
}

/* Ring-fused product (sjds.cu: spmv_sjds_ring_kernel).  qbgpu_peer_ring_reset clears the arrival flags on the compute
 * stream; qbgpu_peer_pull_flag is qbgpu_peer_pull_async followed, on the same lane, by a 4-byte copy-engine write that
 * sets flags[ring_distance] -- no SM is needed to signal an arrival, so the product kernel may occupy every SM and spin;
 * qbgpu_peer_ring_status reports whether a waiter ever gave up (it does after ~1 s instead of hanging the GPU). */
int qbgpu_peer_ring_reset(void)
{
    QB_TRY(ensure_init());
    QB_TRY(peer_init());
    QB_CUDA(cudaMemsetAsync(g_peer.flags, 0, sizeof(int) * kPeerLanes, ctx().stream));
    return QBGPU_OK;
}

int qbgpu_peer_pull_flag(int lane, int slot, void *dst_local, const void *src_peer, size_t bytes, int ring_distance)
{
    QB_TRY(ensure_init());
    QB_TRY(peer_init());
    if (lane < 0 || lane >= kPeerLanes || slot < 0 || slot >= kPeerLanes || ring_distance < 1 || ring_distance >= kPeerLanes || !dst_local || !src_peer)
        return fail(QBGPU_ERR_ARG, "peer_pull_flag: bad argument");
    Context &c = ctx();
    QB_CUDA(cudaEventRecord(g_peer.fence, c.stream));
    QB_CUDA(cudaStreamWaitEvent(g_peer.lane[lane], g_peer.fence, 0));
    if (bytes) QB_CUDA(cudaMemcpyAsync(dst_local, src_peer, bytes, cudaMemcpyDefault, g_peer.lane[lane]));
    QB_CUDA(cudaMemcpyAsync(g_peer.flags + ring_distance, g_peer.one, sizeof(int), cudaMemcpyDeviceToDevice, g_peer.lane[lane]));
    QB_CUDA(cudaEventRecord(g_peer.arrived[slot], g_peer.lane[lane]));
    return QBGPU_OK;
}

int qbgpu_peer_ring_status(int *timed_out)
{
    QB_TRY(ensure_init());
    QB_TRY(peer_init());
    if (!timed_out) return fail(QBGPU_ERR_ARG, "null argument");
    QB_CUDA(cudaMemcpyAsync(timed_out, g_peer.flags + kPeerLanes, sizeof(int), cudaMemcpyDeviceToHost, ctx().stream));
    QB_CUDA(cudaStreamSynchronize(ctx().stream));
    return QBGPU_OK;
}

int qbgpu_ring_prepare(qbgpu_matrix_t A, int rank, int world, int64_t chunk, qbgpu_matrix_t *ring_view)
{
    QB_TRY(ensure_init());
    QB_TRY(no_species(A, "ring_prepare"));
    return ring_prepare(A, rank, world, chunk, ring_view);
}

/* Order everything issued afterwards on the compute stream behind the pull recorded in `slot`. */
int qbgpu_peer_wait(int slot)
{
    QB_TRY(ensure_init());
    QB_TRY(peer_init());
    if (slot < 0 || slot >= kPeerLanes) return fail(QBGPU_ERR_ARG, "peer_wait: bad slot");
    QB_CUDA(cudaStreamWaitEvent(ctx().stream, g_peer.arrived[slot], 0));
    return QBGPU_OK;
}

}  // extern "C"
