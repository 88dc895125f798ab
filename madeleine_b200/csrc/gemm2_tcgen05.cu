// CTA-pair (cta_group::2) variant of the K-major fused-pass tcgen05 GEMM.
//
// Measured on B200 (tools/gemm_bounds.py): with one CTA per tile the 128x256 MMA only reaches ~75 % of peak even with
// operand loads and stores disabled, and TMA writes cost another ~30 % — one SM's shared-memory port (128 B/cycle) has
// to serve the tensor core's A+B operand reads (96 B/cycle), the TMA fills and the epilogue staging.  Pairing two SMs
// on a 256x256 tile halves the B traffic of each SM: every CTA stages its own 128 rows of A and HALF of B
// (32 KB per 32-wide k-block instead of 48 KB, so six pipeline stages fit), the leader CTA issues
// tcgen05.mma.cta_group::2 (UMMA M = 256) which reads both halves of B across the pair, and each CTA drains its own
// 128 accumulator rows from its own TMEM.
//
// Protocol (cluster of 2, rank 0 = leader):
//   full[s]       lives in the leader; both CTAs' TMA loads complete_tx on it (peer clears the CTA bit of the barrier
//                 address); the leader's producer arms it with the byte count of BOTH CTAs
//   empty[s]      one per CTA; the leader's tcgen05.commit multicasts the arrival to both
//   tmem_full[a]  one per CTA, multicast commit after the last k-block of a tile
//   tmem_empty[a] lives in the leader, count = 2 x 8 epilogue warps; the peer's warps arrive remotely
#include "gemm_common.cuh"
#include "madeleine_b200.h"

namespace mdl {

constexpr int STAGES2 = 6;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> leader CTA

struct Smem2 {
    static constexpr int A_BYTES = 2 * BLOCK_M * BLOCK_KF * 2;      // hi + lo planes of this CTA's 128 rows
    static constexpr int B_BYTES = 2 * 128 * BLOCK_KF * 2;          // hi + lo planes of this CTA's half of the 256 B rows
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;           // 32 KB
    static constexpr int BAR_OFFSET = STAGES2 * STAGE_BYTES;
    static constexpr int AUX_OFFSET = BAR_OFFSET + 256;
    static constexpr int AUX_BYTES = 8 * 32 * 33 * 4;
    static constexpr int TOTAL = AUX_OFFSET + AUX_BYTES + 1024;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const void* tmap, uint32_t leader_bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {   // arrives on `bar` at the same offset in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}

// kMN = false: K-major operands (A[M,K], B[N,K]), 64-byte swizzle.   kMN = true: MN-major operands (A[t,M], B[t,N],
// contraction over token rows t, split-K over p.ksplit), 128-byte swizzle, [64 cols x 32 rows] TMA boxes.
template <int EPI, bool kMN, bool ONE_PASS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_kf_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmArgs p) {
    using L = Smem2;
    constexpr int BLOCK_N = 256;
    constexpr int BOX = 64 * BLOCK_KF * 2;   // MN-major: one [64 x 32] box = 4 KB
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + L::BAR_OFFSET;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES2 + s); };
    auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * STAGES2 + s); };
    auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES2 + 2 + s); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES2 + 4);
    volatile uint32_t* tmem_ptr_generic = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < STAGES2; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tmem_full_bar(s), 1); mbar_init(tmem_empty_bar(s), 2 * EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) {   // the same warp of both CTAs allocates collectively
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    pdl_wait();      // everything above (barrier init, TMEM allocation, tensor-map prefetch) overlaps the previous kernel's tail
    float* aux = reinterpret_cast<float*>(smem_raw + (smem_base + L::AUX_OFFSET - smem_u32(smem_raw)));
    if constexpr (EPI == EPI_GATED) {
        const int hc = p.n_heads * 512;
        for (int i = threadIdx.x; i < hc; i += GEMM_THREADS) {
            if constexpr (ONE_PASS) { aux[i] = __ldg(p.ba + i); aux[2048 + i] = 0.5f * __ldg(p.bb + i); }      // FAST_GATES
            else { aux[i] = -2.885390081777927f * __ldg(p.ba + i); aux[2048 + i] = -1.4426950408889634f * __ldg(p.bb + i); }   // pre-scaled: see EPI_GATED
            aux[4096 + i] = __ldg(p.wc + i);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // barriers of both CTAs are initialised before any remote arrive / TMA
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_generic;

    // p.nsplit == 1 (bf16 mode, or the backward pass of fp32_fwd): one plane, a stage carries 64 of k; p.k_blocks counts stages
    // (a compile-time switch: a runtime test in the single-thread TMA / MMA issue loops cost the 3-pass GEMMs 6 %)
    constexpr bool one_pass = ONE_PASS;
    constexpr int kstep = one_pass ? 2 : 1;
    const int n_groups = p.num_n_tiles / p.n_inner;
    const int num_units = p.num_m_tiles * n_groups * p.ksplit;     // m tiles of 256 rows
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    auto decode = [&](int unit, int& m_tile, int& n_group, int& kb0, int& kb1) {
        // output tile fastest, k-range slowest: the CTAs (pairs) of one round sweep the SAME token range for different output
        // tiles, so a split-K wgrad fetches every operand row from HBM once and shares it through L2
        const int mn_count = p.num_m_tiles * n_groups;
        const int mn = unit % mn_count;
        const int ks = unit / mn_count;
        n_group = mn % n_groups;
        m_tile = mn / n_groups;
        const int per = (p.k_blocks + p.ksplit - 1) / p.ksplit;
        kb0 = ks * per;
        kb1 = min(p.k_blocks, kb0 + per);
    };

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int unit = pair; unit < num_units; unit += num_pairs) {
                int m_tile, n_group, kb0, kb1;
                decode(unit, m_tile, n_group, kb0, kb1);
                for (int inner = 0; inner < p.n_inner; ++inner) {
                    const int n_tile = n_group * p.n_inner + inner;
                    const int a_k0 = (n_tile / p.grp_n_tiles) * p.a_koff;
                    const int b_c0 = (m_tile / p.grp_m_tiles) * p.b_coff;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
                        const uint32_t sb = sa + L::A_BYTES;
                        const uint32_t leader_full = full_bar(stage) & kPeerMask;
                        if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * L::STAGE_BYTES);
                        const int row_a = m_tile * 256 + rank * 128, row_b = n_tile * BLOCK_N + rank * 128;
                        // the two half-slots of a stage: 3-pass = (hi, lo) planes of one 32-wide k-block;
                        // 1-pass = two consecutive 32-wide k-blocks of the only plane
                        const int kc0 = kb * BLOCK_KF * kstep;
                        const int kc1 = one_pass ? kc0 + BLOCK_KF : kc0;
                        const int pl1 = one_pass ? 0 : 1;
                        if constexpr (!kMN) {
                            tma_load_3d_2sm(sa, &tmap_a, leader_full, a_k0 + kc0, row_a, 0);
                            tma_load_3d_2sm(sa + L::A_BYTES / 2, &tmap_a, leader_full, a_k0 + kc1, row_a, pl1);
                            tma_load_3d_2sm(sb, &tmap_b, leader_full, kc0, row_b, 0);
                            tma_load_3d_2sm(sb + L::B_BYTES / 2, &tmap_b, leader_full, kc1, row_b, pl1);
                        } else {
#pragma unroll
                            for (int pl = 0; pl < 2; ++pl) {
                                const int tok = pl == 0 ? kc0 : kc1, plane = pl == 0 ? 0 : pl1;
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
                                    tma_load_3d_2sm(sa + pl * (L::A_BYTES / 2) + j * BOX, &tmap_a, leader_full, row_a + 64 * j, tok, plane);
                                    tma_load_3d_2sm(sb + pl * (L::B_BYTES / 2) + j * BOX, &tmap_b, leader_full, b_c0 + row_b + 64 * j, tok, plane);
                                }
                            }
                        }
                        if (++stage == STAGES2) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = p.ab_f16 ? idesc_as_f16(make_idesc_bf16(256, BLOCK_N, kMN, kMN)) : make_idesc_bf16(256, BLOCK_N, kMN, kMN);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int unit = pair; unit < num_units; unit += num_pairs) {
                int m_tile, n_group, kb0, kb1;
                decode(unit, m_tile, n_group, kb0, kb1);
                for (int inner = 0; inner < p.n_inner; ++inner) {
                    mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
                    uint32_t accumulate = 0;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
                        const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
                        for (int k = 0; k < BLOCK_KF / UMMA_K; ++k) {
                            uint64_t a_hi, a_lo, b_hi, b_lo;
                            if constexpr (!kMN) {
                                a_hi = make_umma_desc_sw64(sa + k * (UMMA_K * 2), 16, 512);
                                a_lo = make_umma_desc_sw64(sa + L::A_BYTES / 2 + k * (UMMA_K * 2), 16, 512);
                                b_hi = make_umma_desc_sw64(sb + k * (UMMA_K * 2), 16, 512);
                                b_lo = make_umma_desc_sw64(sb + L::B_BYTES / 2 + k * (UMMA_K * 2), 16, 512);
                            } else {
                                // 64-element chunks along M/N are one [64 x 32] box (4 KB) apart; 16 token rows = 2048 B
                                a_hi = make_umma_desc_sw128(sa + k * (UMMA_K * 128), BOX, 1024);
                                a_lo = make_umma_desc_sw128(sa + L::A_BYTES / 2 + k * (UMMA_K * 128), BOX, 1024);
                                b_hi = make_umma_desc_sw128(sb + k * (UMMA_K * 128), BOX, 1024);
                                b_lo = make_umma_desc_sw128(sb + L::B_BYTES / 2 + k * (UMMA_K * 128), BOX, 1024);
                            }
                            umma_bf16_2cta(tmem_d, a_hi, b_hi, idesc, accumulate);
                            if constexpr (one_pass) {
                                umma_bf16_2cta(tmem_d, a_lo, b_lo, idesc, 1u);        // the second k-block of the stage
                            } else {
                                umma_bf16_2cta(tmem_d, a_hi, b_lo, idesc, 1u);
                                umma_bf16_2cta(tmem_d, a_lo, b_hi, idesc, 1u);
                            }
                            accumulate = 1;
                        }
                        umma_commit_2cta(empty_bar(stage));      // frees the slot in BOTH CTAs
                        if (++stage == STAGES2) { stage = 0; phase ^= 1u; }
                    }
                    umma_commit_2cta(tmem_full_bar(acc));        // accumulators ready in BOTH CTAs
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                }
            }
        }
    } else {
        // ===================== epilogue warps (8 per CTA) =====================
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        int acc = 0; uint32_t acc_phase = 0;
        int unit_parity = 0;
        for (int unit = pair; unit < num_units; unit += num_pairs) {
            int m_tile, n_group, kb0, kb1;
            decode(unit, m_tile, n_group, kb0, kb1);
            float gated_partial = 0.f;
            for (int inner = 0; inner < p.n_inner; ++inner) {
                const int n_tile = n_group * p.n_inner + inner;
                mbar_wait(tmem_full_bar(acc), acc_phase);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
                epilogue_tile<BLOCK_N, EPI, ONE_PASS && EPI == EPI_GATED>(p, aux, t_row, m_tile * 256 + rank * 128, quad, half, warp - 2, lane, n_tile, n_group, inner,
                                            kb1 > kb0, gated_partial, unit_parity);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tmem_empty_bar(acc) & kPeerMask);   // the leader's barrier
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
            unit_parity ^= 1;
        }
        // A persistent GEMM is one wave from the start: signalling the dependent kernel at the top would park its blocks next to
        // ours for the whole run (measured: 1 % slower steps).  Signal when this CTA's last tile has been drained instead.
        if (warp == 2 && lane == 0) pdl_trigger();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // nobody leaves (or frees TMEM) while the peer may still touch this CTA's smem / barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

template <int EPI, bool kMN, bool ONE_PASS>
static int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args, cudaStream_t stream) {
    auto kern = gemm2_kf_kernel<EPI, kMN, ONE_PASS>;
    static PerDeviceOnce attr_set;   // per instantiation and per device
    unsigned long long dev_bit;
    if (attr_set.needed(dev_bit)) {
        MDL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem2::TOTAL));
        attr_set.mark(dev_bit);
    }
    const int units = args.num_m_tiles * (args.num_n_tiles / args.n_inner) * args.ksplit;
    if (units == 0) return 0;
    int sms = kNumSMs, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int pairs = sms / 2;
    if (units < pairs) pairs = units;
    launch_k(kern, dim3(2 * pairs), dim3(GEMM_THREADS), Smem2::TOTAL, stream, ta, tb, args);
    MDL_CHECK_LAUNCH();
    return 0;
}

int launch_gemm2_kf(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args, int epi, cudaStream_t stream) {
    if (args.nsplit == 1) {
        if (epi == EPI_GATED) return launch2<EPI_GATED, false, true>(ta, tb, args, stream);
        return launch2<EPI_STORE, false, true>(ta, tb, args, stream);
    }
    if (epi == EPI_GATED) return launch2<EPI_GATED, false, false>(ta, tb, args, stream);
    return launch2<EPI_STORE, false, false>(ta, tb, args, stream);
}

int launch_gemm2_mn(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args, cudaStream_t stream) {
    if (args.nsplit == 1) return launch2<EPI_ATOMIC, true, true>(ta, tb, args, stream);
    return launch2<EPI_ATOMIC, true, false>(ta, tb, args, stream);
}

}  // namespace mdl
