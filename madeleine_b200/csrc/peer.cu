// Latency-bound exchanges over NVLink / NVSwitch peer memory, written for the two small collectives on the path:
//   * the all-gather of the slide embeddings before the contrastive loss (64 KB per rank), and
//   * the all-reduce of the gradients that only become final at the very end of backward (2 MB: the first two layers).
// NCCL needs 60-100 us for either at 8 ranks (launch + multi-step ring / tree protocol); here every rank keeps a
// symmetric buffer that its peers can read directly, a flag exchange over the peers' signal pads is the only
// synchronisation, and each rank pulls what it needs with plain loads in ONE kernel on its compute stream:
//
//   1. the caller has copied its contribution into its own symmetric buffer (stream-ordered before this kernel);
//   2. block 0 writes `epoch` into slot [rank] of every peer's signal pad (st.release.sys);
//   3. every block spins until all `world` slots of its OWN pad hold >= epoch (ld.acquire.sys): all contributions are
//      then visible;
//   4. every rank reads all contributions in rank order (identical summation order everywhere: bitwise identical sums).
// Buffers are double-buffered by the parity of `epoch`, so no exit barrier is needed: a rank can only overwrite the half
// used by exchange k at exchange k + 2, and reaching that point means it passed the entry barrier of exchange k + 1, which
// every peer signals only after its kernel of exchange k has finished.  Waits are bounded (trap instead of hang).
#include "common.cuh"
#include "madeleine_b200.h"

namespace mdl {

constexpr int PEER_MAX = 16;
struct PeerTable {
    const float* buf[PEER_MAX];
    unsigned* sig[PEER_MAX];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void peer_barrier_in(const PeerTable& t, int rank, int world, unsigned epoch, int slot_base) {
    if (blockIdx.x == 0 && threadIdx.x < world) st_release_sys(t.sig[threadIdx.x] + slot_base + rank, epoch);
    if (threadIdx.x < world) {
        const unsigned* mine = t.sig[rank] + slot_base + threadIdx.x;
        const long long t0 = clock64();
        while ((int)(ld_relaxed_sys(mine) - epoch) < 0) {          // relaxed polls, one acquire fence once the flag is there
            if (clock64() - t0 > 8000000000ll) {
                printf("mdl: peer exchange timeout (rank %d waits for rank %d, epoch %u)\n", rank, (int)threadIdx.x, epoch);
                __trap();
            }
        }
        asm volatile("fence.acquire.sys;" ::: "memory");
    }
    __syncthreads();
}

// out[i] = sum_r buf_r[i]   (n % 4 == 0)
__global__ void __launch_bounds__(256)
peer_allreduce_kernel(PeerTable t, int rank, int world, long long n4, float* __restrict__ out, unsigned epoch, int slot_base) {
    peer_barrier_in(t, rank, world, epoch, slot_base);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v[PEER_MAX];
#pragma unroll
        for (int r = 0; r < PEER_MAX; ++r)
            if (r < world) v[r] = reinterpret_cast<const float4*>(t.buf[r])[i];      // all peers' loads in flight together
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < PEER_MAX; ++r)
            if (r < world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
        reinterpret_cast<float4*>(out)[i] = s;
    }
}

// out[r * n + i] = buf_r[i]
__global__ void __launch_bounds__(256)
peer_allgather_kernel(PeerTable t, int rank, int world, long long n4, float* __restrict__ out, unsigned epoch, int slot_base) {
    peer_barrier_in(t, rank, world, epoch, slot_base);
    const long long total = n4 * world;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / n4);
        reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(t.buf[r])[i - (long long)r * n4];
    }
}

static int peer_table(PeerTable& t, void* const* host_bufs, void* const* host_sigs, int rank, int world, long long buf_off_bytes) {
    MDL_REQUIRE(world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world, "peer exchange: bad rank %d / world %d (max %d)", rank, world, PEER_MAX);
    MDL_REQUIRE(buf_off_bytes % 16 == 0, "peer exchange: buffer offset must be 16-byte aligned");
    for (int r = 0; r < world; ++r) {
        MDL_REQUIRE(host_bufs[r] != nullptr && host_sigs[r] != nullptr, "peer exchange: null peer pointer");
        t.buf[r] = reinterpret_cast<const float*>(reinterpret_cast<const char*>(host_bufs[r]) + buf_off_bytes);
        t.sig[r] = reinterpret_cast<unsigned*>(host_sigs[r]);
    }
    return 0;
}

}  // namespace mdl

using namespace mdl;

extern "C" {

int mdl_peer_allreduce_f32(void* const* host_bufs, void* const* host_sigs, int rank, int world, long long buf_off_bytes, long long n,
                           float* out, unsigned epoch, int slot_base, void* stream) {
    MDL_REQUIRE(n > 0 && n % 4 == 0, "peer_allreduce: n must be a positive multiple of 4 (got %lld)", n);
    PeerTable t;
    int rc = peer_table(t, host_bufs, host_sigs, rank, world, buf_off_bytes);
    if (rc) return rc;
    // a thread keeps `world` 16-byte loads in flight; few blocks for small messages (every block polls the signal pad)
    long long blocks = (n / 4 + 1023) / 1024;
    if (blocks < 1) blocks = 1;
    if (blocks > kNumSMs) blocks = kNumSMs;
    peer_allreduce_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(t, rank, world, n / 4, out, epoch, slot_base);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_peer_allgather_f32(void* const* host_bufs, void* const* host_sigs, int rank, int world, long long buf_off_bytes, long long n_per_rank,
                           float* out, unsigned epoch, int slot_base, void* stream) {
    MDL_REQUIRE(n_per_rank > 0 && n_per_rank % 4 == 0, "peer_allgather: n_per_rank must be a positive multiple of 4 (got %lld)", n_per_rank);
    PeerTable t;
    int rc = peer_table(t, host_bufs, host_sigs, rank, world, buf_off_bytes);
    if (rc) return rc;
    long long blocks = (n_per_rank / 4 * world + 1023) / 1024;
    if (blocks < 1) blocks = 1;
    if (blocks > kNumSMs) blocks = kNumSMs;
    peer_allgather_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(t, rank, world, n_per_rank / 4, out, epoch, slot_base);
    MDL_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
