// On-device patch resampling for a feature store that lives in HBM.
//
// The reference's training loader reads every slide's [N, 512] fp32 features from HDF5, keeps `sample` = 2048 of them
// (madeleine/datasets/wsi_dataset.py:42-50: randperm(N)[:sample] when N >= sample, else randint(0, N, (sample,))),
// stacks the stains of a case and ships 4 MB per bag over PCIe every step (Model.py:113).  A pre-training set of a few
// thousand slides is tens of GB — it fits the 180 GB of one B200 — so here the features stay resident and one kernel
// draws the sample AND gathers the rows into the [R, sample, D] batch: no host work and no H2D per step.
//
//   N >= S : slot s takes row  pi(s)  of the bag, pi a keyed pseudo-random PERMUTATION of [0, N) evaluated pointwise
//            (4-round balanced Feistel network on the next even power of two, cycle-walked back into [0, N)): distinct
//            rows without sorting or a sequential shuffle, O(1) work per slot;
//   N <  S : slot s takes row  hash(seed, bag, s) mod N  (with replacement, like randint);
//   N == 0 : the stain is missing -> S zero rows (the reference feeds torch.zeros([2, D]) through the same sampler).
// One warp per (bag, slot): the index is computed redundantly by all lanes, the 2 KB row moves as 16-byte vectors.
#include "common.cuh"
#include "madeleine_b200.h"

namespace mdl {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {   // "lowbias32" integer finaliser: full avalanche
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

// pi(i) for i in [0, n): Feistel permutation of [0, 2^(2h)) restricted to [0, n) by cycle walking.
__device__ __forceinline__ uint32_t feistel_perm(uint32_t i, uint32_t n, uint32_t key) {
    if (n <= 1) return 0;
    int bits = 32 - __clz(n - 1);          // smallest b with 2^b >= n
    const int h = (bits + 1) >> 1;         // half width; domain 2^(2h) < 4n
    const uint32_t mask = (1u << h) - 1u;
    uint32_t x = i;
    do {
        uint32_t l = x >> h, r = x & mask;
#pragma unroll
        for (int round = 0; round < 4; ++round) {
            const uint32_t f = mix32(r * 0x9E3779B1u + key + (uint32_t)round * 0x85EBCA6Bu) & mask;
            const uint32_t nl = r;
            r = l ^ f;
            l = nl;
        }
        x = (l << h) | r;
    } while (x >= n);
    return x;
}

__global__ void __launch_bounds__(256)
sample_gather_kernel(const float* __restrict__ store, const long long* __restrict__ bag_offset, const int* __restrict__ bag_len,
                     int R, int S, int D, unsigned long long seed, float* __restrict__ out, int* __restrict__ idx_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long total = (long long)R * S;
    const int vec = D >> 2;
    for (long long w = (long long)blockIdx.x * 8 + warp; w < total; w += (long long)gridDim.x * 8) {
        const int b = (int)(w / S), s = (int)(w - (long long)b * S);
        const int n = __ldg(bag_len + b);
        float4* dst = reinterpret_cast<float4*>(out + w * D);
        if (n <= 0) {
            for (int v = lane; v < vec; v += 32) dst[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx_out != nullptr && lane == 0) idx_out[w] = -1;
            continue;
        }
        const uint32_t key = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x9E3779B9u * (uint32_t)(b + 1)));
        uint32_t idx;
        if (n >= S) idx = feistel_perm((uint32_t)s, (uint32_t)n, key);
        else idx = mix32(key ^ mix32((uint32_t)s + 0x632BE5ABu)) % (uint32_t)n;
        const float4* src = reinterpret_cast<const float4*>(store + (__ldg(bag_offset + b) + idx) * D);
        for (int v = lane; v < vec; v += 32) dst[v] = __ldg(src + v);
        if (idx_out != nullptr && lane == 0) idx_out[w] = (int)idx;
    }
}

}  // namespace mdl

using namespace mdl;

extern "C" int mdl_sample_gather_f32(const float* store, const long long* bag_offset, const int* bag_len, int n_bags, int n_sample,
                                     int D, unsigned long long seed, float* out, int* idx_out, void* stream) {
    MDL_REQUIRE(D > 0 && D % 4 == 0, "sample_gather: D must be a positive multiple of 4 (got %d)", D);
    MDL_REQUIRE(n_sample > 0, "sample_gather: n_sample must be positive (got %d)", n_sample);
    if (n_bags == 0) return 0;
    long long blocks = ((long long)n_bags * n_sample + 7) / 8;
    const long long cap = (long long)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    sample_gather_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(store, bag_offset, bag_len, n_bags, n_sample, D, seed, out, idx_out);
    MDL_CHECK_LAUNCH();
    return 0;
}
