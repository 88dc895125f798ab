// Persistent, warp-specialised tcgen05 GEMM for sm_100a with TMA-fed 128B-swizzled operands and TMEM accumulators.
//
// Operands are "split planes": an fp32 matrix X is held in HBM as two bf16 matrices (hi, lo) with X ~= hi + lo.
// nsplit = 3 evaluates hi*hi + hi*lo + lo*hi in one fp32 TMEM accumulator (fp32-grade products on the bf16
// tensor pipe, |err| ~ 2^-16); nsplit = 1 uses the hi planes only (plain bf16, what the reference's autocast
// scripts compute).  The three passes are just three TMA coordinate schedules over the same kernel.
//
// Roles (320 threads): warp 0 = TMA producer (1 lane), warp 1 = TMEM owner + MMA issuer (1 lane),
// warps 2..9 = epilogue (TMEM lane quadrant = warp_id % 4, two warps per quadrant split the columns).  Two 256-column accumulator stages let the epilogue of
// tile i overlap the MMAs of tile i+1.  Grid = min(#work units, 148): one CTA per SM, static round-robin.
//
// Modes
//   MODE_KF : like MODE_K for the 3-pass split: one pipeline stage holds A_hi, A_lo, B_hi, B_lo of a 32-wide k-block
//             (64-byte swizzle) and feeds all three products, so every operand byte crosses L2->SM once instead of
//             A_hi and B_hi twice (the K=512 GEMMs were L2-operand-bandwidth bound: -35 % traffic, same stage depth).
//   kKMajor : C[M,N] = A[M,K] * B[N,K]^T, both operands contiguous along K (forward and dgrad GEMMs).
//   kMNMajor: C[M,N] = sum_t A[t,M]^T B[t,N], both operands contiguous along their output index (wgrad;
//             t = tokens), split over t across CTAs with fp32 red.add accumulation.
// Epilogues
//   EPI_STORE : C = acc + bias[n] + rowbias[row2bag[m], n]               (fp32 store)
//   EPI_GATED : per head h, logit[m,h] = sum_j tanh(a_j+ba_j) * sigmoid(b_j+bb_j) * wc_j + bc  over 4 n-tiles
//               holding 128 'a' and 128 'b' columns each; optionally stores the (dropout-scaled) gates as fp16.
//   EPI_ATOMIC: C += acc (red.global.add.f32), used by split-K wgrad.
#include "gemm_common.cuh"
#include "madeleine_b200.h"
#include <mutex>
#include <stdlib.h>

namespace mdl {

template <int BLOCK_N, int MODE, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmArgs p) {
    constexpr bool kMNMajor = MODE == MODE_MN;
    constexpr bool kFused = MODE == MODE_KF;
    using L = SmemLayout<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-byte alignment
    const uint32_t bar_base = smem_base + L::BAR_OFFSET;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_ptr_generic =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tmem_full_bar(s), 1); mbar_init(tmem_empty_bar(s), EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_addr, TMEM_COLS);
        tmem_relinquish();
    }
    pdl_wait();      // everything above (barrier init, TMEM allocation, tensor-map prefetch) overlaps the previous kernel's tail
    float* aux = reinterpret_cast<float*>(smem_raw + (smem_base + L::AUX_OFFSET - smem_u32(smem_raw)));
    if constexpr (EPI == EPI_GATED) {
        const int hc = p.n_heads * 512;   // <= 2048
        for (int i = threadIdx.x; i < hc; i += GEMM_THREADS) {
            aux[i] = -2.885390081777927f * __ldg(p.ba + i); aux[2048 + i] = -1.4426950408889634f * __ldg(p.bb + i); aux[4096 + i] = __ldg(p.wc + i);   // pre-scaled: see EPI_GATED
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_generic;

    const int num_units = p.num_m_tiles * (p.num_n_tiles / p.n_inner) * p.ksplit;
    const int n_groups = p.num_n_tiles / p.n_inner;
    const int iters_total = p.k_blocks;  // k-blocks over the whole contraction

    // unit -> (m_tile, n_group, k range). n_group varies fastest so CTAs running concurrently share the A tile in L2.
    auto decode = [&](int unit, int& m_tile, int& n_group, int& kb0, int& kb1) {
        // output tile fastest, k-range slowest: the CTAs (pairs) of one round sweep the SAME token range for different output
        // tiles, so a split-K wgrad fetches every operand row from HBM once and shares it through L2
        const int mn_count = p.num_m_tiles * n_groups;
        const int mn = unit % mn_count;
        const int ks = unit / mn_count;
        n_group = mn % n_groups;
        m_tile = mn / n_groups;
        const int per = (iters_total + p.ksplit - 1) / p.ksplit;
        kb0 = ks * per;
        kb1 = min(iters_total, kb0 + per);
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
                int m_tile, n_group, kb0, kb1;
                decode(unit, m_tile, n_group, kb0, kb1);
                for (int inner = 0; inner < p.n_inner; ++inner) {
                    const int n_tile = n_group * p.n_inner + inner;
                    if constexpr (kFused) {
                        const int a_k0 = (n_tile / p.grp_n_tiles) * p.a_koff;
                        for (int kb = kb0; kb < kb1; ++kb) {
                            mbar_wait(empty_bar(stage), phase ^ 1u);
                            const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
                            const uint32_t sb = sa + L::A_BYTES;
                            if (p.debug_flags & 1) { mbar_arrive(full_bar(stage)); if (++stage == STAGES) { stage = 0; phase ^= 1u; } continue; }
                            mbar_arrive_expect_tx(full_bar(stage), L::STAGE_BYTES);
                            tma_load_3d(sa, &tmap_a, full_bar(stage), a_k0 + kb * BLOCK_KF, m_tile * BLOCK_M, 0);
                            tma_load_3d(sa + L::A_BYTES / 2, &tmap_a, full_bar(stage), a_k0 + kb * BLOCK_KF, m_tile * BLOCK_M, 1);
                            tma_load_3d(sb, &tmap_b, full_bar(stage), kb * BLOCK_KF, n_tile * BLOCK_N, 0);
                            tma_load_3d(sb + L::B_BYTES / 2, &tmap_b, full_bar(stage), kb * BLOCK_KF, n_tile * BLOCK_N, 1);
                            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                        }
                        continue;
                    }
                    for (int kb = kb0; kb < kb1; ++kb) {
                        for (int pass = 0; pass < p.nsplit; ++pass) {
                            const int plane_a = pass == 2 ? 1 : 0;
                            const int plane_b = pass == 1 ? 1 : 0;
                            mbar_wait(empty_bar(stage), phase ^ 1u);
                            const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
                            const uint32_t sb = sa + L::A_BYTES;
                            mbar_arrive_expect_tx(full_bar(stage), L::STAGE_BYTES);
                            if constexpr (!kMNMajor) {
                                const int a_k0 = (n_tile / p.grp_n_tiles) * p.a_koff;
                                tma_load_3d(sa, &tmap_a, full_bar(stage), a_k0 + kb * BLOCK_K, m_tile * BLOCK_M, plane_a);
                                tma_load_3d(sb, &tmap_b, full_bar(stage), kb * BLOCK_K, n_tile * BLOCK_N, plane_b);
                            } else {
                                const int b_c0 = (m_tile / p.grp_m_tiles) * p.b_coff;
#pragma unroll
                                for (int j = 0; j < BLOCK_M / 64; ++j)
                                    tma_load_3d(sa + j * (64 * BLOCK_K * 2), &tmap_a, full_bar(stage),
                                                m_tile * BLOCK_M + 64 * j, kb * BLOCK_K, plane_a);
#pragma unroll
                                for (int j = 0; j < BLOCK_N / 64; ++j)
                                    tma_load_3d(sb + j * (64 * BLOCK_K * 2), &tmap_b, full_bar(stage),
                                                b_c0 + n_tile * BLOCK_N + 64 * j, kb * BLOCK_K, plane_b);
                            }
                            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = p.ab_f16 ? idesc_as_f16(make_idesc_bf16(BLOCK_M, BLOCK_N, kMNMajor, kMNMajor))
                                            : make_idesc_bf16(BLOCK_M, BLOCK_N, kMNMajor, kMNMajor);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
                int m_tile, n_group, kb0, kb1;
                decode(unit, m_tile, n_group, kb0, kb1);
                for (int inner = 0; inner < p.n_inner; ++inner) {
                    mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
                    uint32_t accumulate = 0;
                    if constexpr (kFused) {
                        for (int kb = kb0; kb < kb1; ++kb) {
                            mbar_wait(full_bar(stage), phase);
                            tc_fence_after();
                            const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
                            const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
                            for (int k = 0; k < BLOCK_KF / UMMA_K; ++k) {
                                // rows of 64 B, 8-row atoms 512 B apart; step 16 elements = 32 B along the row
                                const uint64_t a_hi = make_umma_desc_sw64(sa + k * (UMMA_K * 2), 16, 512);
                                const uint64_t a_lo = make_umma_desc_sw64(sa + L::A_BYTES / 2 + k * (UMMA_K * 2), 16, 512);
                                const uint64_t b_hi = make_umma_desc_sw64(sb + k * (UMMA_K * 2), 16, 512);
                                const uint64_t b_lo = make_umma_desc_sw64(sb + L::B_BYTES / 2 + k * (UMMA_K * 2), 16, 512);
                                umma_bf16(tmem_d, a_hi, b_hi, idesc, accumulate);
                                umma_bf16(tmem_d, a_hi, b_lo, idesc, 1u);
                                umma_bf16(tmem_d, a_lo, b_hi, idesc, 1u);
                                accumulate = 1;
                            }
                            umma_commit(empty_bar(stage));
                            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                        }
                        umma_commit(tmem_full_bar(acc));
                        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                        continue;
                    }
                    for (int kb = kb0; kb < kb1; ++kb) {
                        for (int pass = 0; pass < p.nsplit; ++pass) {
                            mbar_wait(full_bar(stage), phase);
                            tc_fence_after();
                            const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
                            const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                uint64_t da, db;
                                if constexpr (!kMNMajor) {
                                    // K-major: rows of 128 B, 8-row atoms 1024 B apart; step 16 elements = 32 B along the row.
                                    da = make_umma_desc_sw128(sa + k * (UMMA_K * 2), 16, 1024);
                                    db = make_umma_desc_sw128(sb + k * (UMMA_K * 2), 16, 1024);
                                } else {
                                    // MN-major: 64-element (128 B) chunks along M/N, LBO = one [64 x BLOCK_K] box,
                                    // 8-row K atoms 1024 B apart; step 16 rows = 2048 B.
                                    da = make_umma_desc_sw128(sa + k * (UMMA_K * 128), 64 * BLOCK_K * 2, 1024);
                                    db = make_umma_desc_sw128(sb + k * (UMMA_K * 128), 64 * BLOCK_K * 2, 1024);
                                }
                                umma_bf16(tmem_d, da, db, idesc, accumulate);
                                accumulate = 1;
                            }
                            umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
                            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                        }
                    }
                    umma_commit(tmem_full_bar(acc));
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                }
            }
        }
    } else {
        // ===================== epilogue warps (8) =====================
        const int quad = warp & 3;             // TMEM lane quadrant this warp may read
        const int half = (warp - 2) >> 2;      // which half of the tile's columns this warp drains
        int acc = 0; uint32_t acc_phase = 0;
        int unit_parity = 0;
        for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
            int m_tile, n_group, kb0, kb1;
            decode(unit, m_tile, n_group, kb0, kb1);
            float gated_partial = 0.f;
            for (int inner = 0; inner < p.n_inner; ++inner) {
                const int n_tile = n_group * p.n_inner + inner;
                mbar_wait(tmem_full_bar(acc), acc_phase);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
                const bool have_k = kb1 > kb0;  // empty split-K slice: accumulator is stale, skip
                epilogue_tile<BLOCK_N, EPI>(p, aux, t_row, m_tile * BLOCK_M, quad, half, warp - 2, lane, n_tile, n_group, inner, have_k,
                                            gated_partial, unit_parity);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
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
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------
// host side: tensor maps and launch
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    });
    return fn;
}

// bf16 planes tensor [planes][rows][cols] (cols contiguous, row stride ld elements, plane stride in elements);
// box = [box_cols=64][box_rows][1], 128B swizzle, OOB -> zeros.
static int make_plane_tmap(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld,
                           long long plane_stride, int planes, int box_rows, int box_cols = 64) {
    PFN_encodeTiled enc = get_encode_fn();
    MDL_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the CUDA driver");
    MDL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15u) == 0, "TMA base pointer must be 16-byte aligned");
    MDL_REQUIRE((ld * 2) % 16 == 0 && (plane_stride * 2) % 16 == 0, "TMA strides must be multiples of 16 bytes");
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane_stride * 2};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1u};
    const CUtensorMapSwizzle swz = box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MDL_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r, rows, cols, ld);
    return 0;
}

template <int BLOCK_N, int MODE, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args, cudaStream_t stream) {
    using L = SmemLayout<BLOCK_N>;
    auto kern = gemm_tcgen05_kernel<BLOCK_N, MODE, EPI>;
    static PerDeviceOnce attr_set;   // per instantiation and per device
    unsigned long long dev_bit;
    if (attr_set.needed(dev_bit)) {
        MDL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set.mark(dev_bit);
    }
    const int units = args.num_m_tiles * (args.num_n_tiles / args.n_inner) * args.ksplit;
    if (units == 0) return 0;
    int sms = kNumSMs;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = units < sms ? units : sms;
    launch_k(kern, dim3(grid), dim3(GEMM_THREADS), L::TOTAL, stream, ta, tb, args);
    MDL_CHECK_LAUNCH();
    return 0;
}

static int g_debug_flags = 0;

// The fused-pass schedule is the default for nsplit = 3; MDL_GEMM_FUSED=0 selects the pass-serial one (debug / A-B).
static bool use_fused() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MDL_GEMM_FUSED");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// CTA-pair kernels are the default for the 3-pass K-major GEMMs with N % 256 == 0; MDL_GEMM_2CTA=0 disables them.
static bool use_2cta() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MDL_GEMM_2CTA");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}
// 1-pass (bf16) problems on the CTA-pair kernel.  Measured on the bench step (tools/step_breakdown.py --precision bf16): the
// gated-attention GEMM gains 8 % (0.483 -> 0.443 ms; its epilogue math leaves the pair's halved operand traffic visible), the
// store and wgrad GEMMs are output-bandwidth-bound at one pass and gain nothing (0.749 -> 0.772, 0.513 -> 0.516 ms).
// MDL_GEMM_2CTA_BF16: 0 = never, 1 (default) = gated GEMM only, 2 = every 1-pass GEMM.
static int cta_pair_1pass_level() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MDL_GEMM_2CTA_BF16");
        v = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
    }
    return v;
}
static bool use_2cta_1pass() { return cta_pair_1pass_level() >= 2; }
static bool use_2cta_1pass_gated() { return cta_pair_1pass_level() >= 1; }

}  // namespace mdl

using namespace mdl;

extern "C" {

int mdl_gemm_nt(const void* a_planes, long long a_rows, long long a_cols, long long lda, long long a_plane_stride,
                const void* b_planes, long long b_rows, long long b_cols, long long ldb, long long b_plane_stride,
                void* out, long long ldc, int M, int N, int K, int nsplit,
                int grp_n_cols, int a_koff,
                const float* bias, const float* rowbias, const int* row2bag, int out_bf16, void* stream) {
    const bool f16 = (nsplit & kPlanesF16) != 0;     // operands are fp16 hi/lo planes, the B planes scaled by MDL_F16_WEIGHT_SCALE
    nsplit &= 0xff;
    MDL_REQUIRE(nsplit == 1 || nsplit == 3, "nsplit must be 1 or 3");
    MDL_REQUIRE(K % BLOCK_K == 0, "K (%d) must be a multiple of %d", K, BLOCK_K);
    MDL_REQUIRE(N % 128 == 0, "N (%d) must be a multiple of 128", N);
    MDL_REQUIRE(M > 0, "M must be positive");
    const int planes = nsplit == 3 ? 2 : 1;
    const int bn = (N % 256 == 0) ? 256 : 128;
    const bool fused = nsplit == 3 && use_fused();
    // the CTA-pair kernel also takes 1-pass problems (a stage then holds two k-blocks of the single plane)
    const bool two_cta = (fused || (nsplit == 1 && use_fused() && use_2cta_1pass())) && bn == 256 && use_2cta() && g_debug_flags == 0;
    const int bk = (fused || two_cta) ? BLOCK_KF : BLOCK_K;
    CUtensorMap ta, tb;
    int rc = make_plane_tmap(&ta, a_planes, a_rows, a_cols, lda, a_plane_stride, planes, BLOCK_M, bk);
    if (rc) return rc;
    rc = make_plane_tmap(&tb, b_planes, b_rows, b_cols, ldb, b_plane_stride, planes, two_cta ? 128 : bn, bk);
    if (rc) return rc;
    GemmArgs g{};
    g.M = M; g.N = N; g.k_blocks = (two_cta && nsplit == 1) ? K / (2 * bk) : K / bk; g.nsplit = nsplit;
    g.num_m_tiles = (M + BLOCK_M - 1) / BLOCK_M; g.num_n_tiles = N / bn; g.n_inner = 1; g.ksplit = 1;
    MDL_REQUIRE(grp_n_cols <= 0 || grp_n_cols % bn == 0, "grp_n_cols (%d) must be a multiple of the N tile (%d)", grp_n_cols, bn);
    MDL_REQUIRE(ldc % 4 == 0, "ldc must be a multiple of 4");
    g.grp_n_tiles = grp_n_cols > 0 ? grp_n_cols / bn : (1 << 30); g.a_koff = a_koff;
    g.grp_m_tiles = 1 << 30; g.b_coff = 0;
    g.out = (float*)out; g.ldc = (int)ldc; g.out_bf16 = out_bf16 ? 1 : 0; g.bias = bias; g.rowbias = rowbias; g.row2bag = row2bag;
    g.debug_flags = g_debug_flags;
    g.ab_f16 = f16 ? 1 : 0; g.acc_scale = f16 ? 1.f / MDL_F16_WEIGHT_SCALE : 1.f;
    if (two_cta) {
        g.num_m_tiles = (M + 255) / 256;
        return launch_gemm2_kf(ta, tb, g, EPI_STORE, (cudaStream_t)stream);
    }
    if (fused) {
        if (bn == 256) return launch_gemm<256, MODE_KF, EPI_STORE>(ta, tb, g, (cudaStream_t)stream);
        return launch_gemm<128, MODE_KF, EPI_STORE>(ta, tb, g, (cudaStream_t)stream);
    }
    if (bn == 256) return launch_gemm<256, MODE_K, EPI_STORE>(ta, tb, g, (cudaStream_t)stream);
    return launch_gemm<128, MODE_K, EPI_STORE>(ta, tb, g, (cudaStream_t)stream);
}

/* benchmark-only knob (tools/gemm_bounds.py): 1 = skip TMA loads, 2 = skip epilogue global writes in mdl_gemm_nt */
int mdl_gemm_debug_flags(int flags) { g_debug_flags = flags; return 0; }

int mdl_gemm_gated(const void* a_planes, long long a_rows, long long a_cols, long long lda, long long a_plane_stride,
                   const void* b_planes, long long b_plane_stride,
                   int M, int n_heads, int nsplit,
                   const float* ba, const float* bb, const float* wc, const float* bc,
                   float* logits, void* gate_a, void* gate_b, float drop_p, unsigned long long seed, void* stream) {
    const bool f16 = (nsplit & kPlanesF16) != 0;
    nsplit &= 0xff;
    MDL_REQUIRE(nsplit == 1 || nsplit == 3, "nsplit must be 1 or 3");
    MDL_REQUIRE(M > 0 && n_heads > 0, "bad sizes");
    const int planes = nsplit == 3 ? 2 : 1;
    const int K = 512, N = n_heads * 1024;
    const bool fused = nsplit == 3 && use_fused();
    const bool two_cta = (fused || (nsplit == 1 && use_fused() && use_2cta_1pass_gated())) && use_2cta();
    const int bk = (fused || two_cta) ? BLOCK_KF : BLOCK_K;
    CUtensorMap ta, tb;
    int rc = make_plane_tmap(&ta, a_planes, a_rows, a_cols, lda, a_plane_stride, planes, BLOCK_M, bk);
    if (rc) return rc;
    rc = make_plane_tmap(&tb, b_planes, N, K, K, b_plane_stride, planes, two_cta ? 128 : 256, bk);
    if (rc) return rc;
    GemmArgs g{};
    g.M = M; g.N = N; g.k_blocks = (two_cta && nsplit == 1) ? K / (2 * bk) : K / bk; g.nsplit = nsplit;
    g.num_m_tiles = (M + BLOCK_M - 1) / BLOCK_M; g.num_n_tiles = N / 256; g.n_inner = 4; g.ksplit = 1;
    g.grp_n_tiles = 4; g.a_koff = 512; g.grp_m_tiles = 1 << 30; g.b_coff = 0;
    g.ba = ba; g.bb = bb; g.wc = wc; g.bc = bc; g.logits = logits;
    g.gate_a = reinterpret_cast<__half*>(gate_a); g.gate_b = reinterpret_cast<__half*>(gate_b);
    g.drop_p = drop_p; g.seed = seed; g.n_heads = n_heads;
    g.ab_f16 = f16 ? 1 : 0; g.acc_scale = f16 ? 1.f / MDL_F16_WEIGHT_SCALE : 1.f;
    if (two_cta) {
        g.num_m_tiles = (M + 255) / 256;
        return launch_gemm2_kf(ta, tb, g, EPI_GATED, (cudaStream_t)stream);
    }
    if (fused) return launch_gemm<256, MODE_KF, EPI_GATED>(ta, tb, g, (cudaStream_t)stream);
    return launch_gemm<256, MODE_K, EPI_GATED>(ta, tb, g, (cudaStream_t)stream);
}

int mdl_gemm_tn_accum(const void* a_planes, long long a_cols, long long lda, long long a_plane_stride,
                      const void* b_planes, long long b_cols, long long ldb, long long b_plane_stride,
                      long long tokens, float* out, long long ldc, int M, int N, int nsplit,
                      int grp_m_rows, int b_coff, int ksplit, void* stream) {
    MDL_REQUIRE(grp_m_rows <= 0 || grp_m_rows % BLOCK_M == 0, "grp_m_rows (%d) must be a multiple of %d", grp_m_rows, BLOCK_M);
    MDL_REQUIRE(nsplit == 1 || nsplit == 3, "nsplit must be 1 or 3");
    MDL_REQUIRE(M % BLOCK_M == 0 && N % 256 == 0, "wgrad output must be a multiple of 128 x 256 (got %d x %d)", M, N);
    MDL_REQUIRE(tokens > 0, "tokens must be positive");
    const int planes = nsplit == 3 ? 2 : 1;
    const bool two_cta = (nsplit == 3 || use_2cta_1pass()) && use_fused() && use_2cta() && M % 256 == 0 &&
                         (grp_m_rows <= 0 || grp_m_rows % 256 == 0);
    const int bk = two_cta ? BLOCK_KF : BLOCK_K;
    CUtensorMap ta, tb;
    int rc = make_plane_tmap(&ta, a_planes, tokens, a_cols, lda, a_plane_stride, planes, bk);
    if (rc) return rc;
    rc = make_plane_tmap(&tb, b_planes, tokens, b_cols, ldb, b_plane_stride, planes, bk);
    if (rc) return rc;
    GemmArgs g{};
    const int k_per_stage = (two_cta && nsplit == 1) ? 2 * bk : bk;
    g.M = M; g.N = N; g.k_blocks = (int)((tokens + k_per_stage - 1) / k_per_stage); g.nsplit = nsplit;
    const int tile_m = two_cta ? 256 : BLOCK_M;
    g.num_m_tiles = M / tile_m; g.num_n_tiles = N / 256; g.n_inner = 1;
    const int mn_tiles = g.num_m_tiles * g.num_n_tiles;
    const int workers = two_cta ? kNumSMs / 2 : kNumSMs;
    if (ksplit <= 0) {
        const char* legacy = getenv("MDL_WGRAD_KSPLIT_LEGACY");      // A/B switch (read per call: tools/ab_probe.py flips it in-process)
        if (legacy != nullptr && legacy[0] == '1') ksplit = (4 * workers + mn_tiles - 1) / mn_tiles;
    }
    if (ksplit <= 0) {
        // Units = output tiles x k-ranges are dealt round-robin to the persistent workers (CTAs or CTA pairs), so the kernel
        // lasts ceil(units / workers) rounds of one k-range each: pick the split with the smallest rounds x (k-blocks per range +
        // a fixed per-unit cost for pipeline refill and the atomic epilogue), i.e. one whose unit count fills its last round
        // (round 2: the former "~4 units per worker" gave 320 units = 4.3 rounds for the attention wgrad and 304 = 4.1 for
        // layer 3 - the fifth round ran nearly empty).
        const int kOverhead = 4;
        long long best_cost = -1;
        const int hi = (16 * workers + mn_tiles - 1) / mn_tiles;
        for (int cand = 1; cand <= hi && cand <= g.k_blocks; ++cand) {
            const long long rounds = ((long long)mn_tiles * cand + workers - 1) / workers;
            const long long per = (g.k_blocks + cand - 1) / cand;
            const long long cost = rounds * (per + kOverhead);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; ksplit = cand; }
        }
    }
    if (ksplit > g.k_blocks) ksplit = g.k_blocks;
    if (ksplit < 1) ksplit = 1;
    g.ksplit = ksplit;
    g.grp_n_tiles = 1 << 30; g.a_koff = 0;
    g.grp_m_tiles = grp_m_rows > 0 ? grp_m_rows / tile_m : (1 << 30); g.b_coff = b_coff;
    g.out = out; g.ldc = (int)ldc;
    if (two_cta) return launch_gemm2_mn(ta, tb, g, (cudaStream_t)stream);
    return launch_gemm<256, MODE_MN, EPI_ATOMIC>(ta, tb, g, (cudaStream_t)stream);
}

}  // extern "C"
