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
#include "common.cuh"
#include "madeleine_b200.h"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>

namespace mdl {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int GEMM_THREADS = 320;   // 2 control warps + 8 epilogue warps
constexpr int EPI_WARPS = 8;
constexpr int TMEM_COLS = 512;

enum { EPI_STORE = 0, EPI_GATED = 1, EPI_ATOMIC = 2 };
enum { MODE_K = 0, MODE_MN = 1, MODE_KF = 2 };
constexpr int BLOCK_KF = 32;  // fused-pass k-block: 32 bf16 = one 64-byte swizzle row

struct GemmArgs {
    int M, N;                 // output extent
    int k_blocks;             // contraction length / BLOCK_K (per pass)
    int nsplit;               // 1 or 3 passes
    int num_m_tiles, num_n_tiles;
    int n_inner;              // consecutive n-tiles processed by one work unit (EPI_GATED: 4)
    int ksplit;               // split-K factor (kMNMajor), else 1
    int grp_n_tiles, a_koff;  // kKMajor: A k-offset = (n_tile / grp_n_tiles) * a_koff
    int grp_m_tiles, b_coff;  // kMNMajor: B column offset = (m_tile / grp_m_tiles) * b_coff
    float* out; int ldc;
    const float* bias;        // [N] or null
    const float* rowbias;     // [R, N] or null (per-bag bias, stain encodings)
    const int* row2bag;       // [M]
    // gated epilogue
    const float* ba; const float* bb; const float* wc; const float* bc;  // [H*512], [H*512], [H*512], [H]
    float* logits;            // [M, H]
    __half* gate_a; __half* gate_b;  // [M, H*512] or null
    float drop_p; unsigned long long seed;
    int n_heads;
};

template <int BLOCK_N>
struct SmemLayout {
    // identical totals in every mode: classic = one plane x 64 k, fused = two planes x 32 k
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int AUX_OFFSET = BAR_OFFSET + 256;
    // aux region: gated epilogue = ba | bb | wc (3 x 2048 floats) + 2 x 128 partials; store/atomic epilogues = one padded
    // 32 x 33 fp32 transpose buffer per epilogue warp
    static constexpr int AUX_BYTES = 8 * 32 * 33 * 4;
    static constexpr int TOTAL = AUX_OFFSET + AUX_BYTES + 1024;         // + alignment slack
};

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
    float* aux = reinterpret_cast<float*>(smem_raw + (smem_base + L::AUX_OFFSET - smem_u32(smem_raw)));
    if constexpr (EPI == EPI_GATED) {
        const int hc = p.n_heads * 512;   // <= 2048
        for (int i = threadIdx.x; i < hc; i += GEMM_THREADS) {
            aux[i] = __ldg(p.ba + i); aux[2048 + i] = __ldg(p.bb + i); aux[4096 + i] = __ldg(p.wc + i);
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
        const int ks = unit % p.ksplit;
        const int mn = unit / p.ksplit;
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
            constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, kMNMajor, kMNMajor);
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
            const int m = m_tile * BLOCK_M + quad * 32 + lane;
            const bool row_ok = m < p.M;
            float gated_partial = 0.f;
            for (int inner = 0; inner < p.n_inner; ++inner) {
                const int n_tile = n_group * p.n_inner + inner;
                mbar_wait(tmem_full_bar(acc), acc_phase);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
                const bool have_k = kb1 > kb0;  // empty split-K slice: accumulator is stale, skip
                if constexpr (EPI == EPI_STORE || EPI == EPI_ATOMIC) {
                    // TMEM -> registers (thread = row) -> +bias -> warp-private padded smem -> row-contiguous global access:
                    // 8 lanes cover one 128-byte row segment, so each vector store / reduction touches 4 full lines instead
                    // of 32 scattered 16-byte pieces.
                    int bag = 0;
                    if (EPI == EPI_STORE && p.rowbias != nullptr && row_ok) bag = p.row2bag[m];
                    float* stg = aux + (warp - 2) * (32 * 33);
                    const int m_warp = m_tile * BLOCK_M + quad * 32;
                    constexpr int CH = BLOCK_N / 64;   // 32-column chunks per half
#pragma unroll 1
                    for (int cc = 0; cc < CH; ++cc) {
                        const int c = half * CH + cc;
                        uint32_t r[32];
                        tmem_ld_32x32(t_row + c * 32, r);
                        tmem_ld_wait();
                        const int n0 = n_tile * BLOCK_N + c * 32;
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            float4 v;
                            v.x = __uint_as_float(r[i]); v.y = __uint_as_float(r[i + 1]);
                            v.z = __uint_as_float(r[i + 2]); v.w = __uint_as_float(r[i + 3]);
                            if (EPI == EPI_STORE && p.bias != nullptr) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + i));
                                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                            }
                            if (EPI == EPI_STORE && p.rowbias != nullptr) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(p.rowbias + (size_t)bag * p.N + n0 + i));
                                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                            }
                            float* d = stg + lane * 33 + i;
                            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
                        }
                        __syncwarp();
                        const int sub = lane >> 3, col4 = (lane & 7) * 4;
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int row = it * 4 + sub;
                            const float* sp = stg + row * 33 + col4;
                            const float x0 = sp[0], x1 = sp[1], x2 = sp[2], x3 = sp[3];
                            if (m_warp + row < p.M) {
                                float* dst = p.out + (size_t)(m_warp + row) * p.ldc + n0 + col4;
                                if constexpr (EPI == EPI_STORE) {
                                    *reinterpret_cast<float4*>(dst) = make_float4(x0, x1, x2, x3);
                                } else {
                                    if (have_k)
                                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x0), "f"(x1), "f"(x2), "f"(x3) : "memory");
                                }
                            }
                        }
                        __syncwarp();
                    }
                } else {  // EPI_GATED
                    const int head = n_group;          // one work unit = (m_tile, head); inner = 128-wide gate group
                    const int j_base = head * 512 + inner * 128;
                    const int HC = p.n_heads * 512;
#pragma unroll 1
                    for (int cc = 0; cc < 2; ++cc) {
                        const int c = half * 2 + cc;
                        uint32_t ra[32], rb[32];
                        tmem_ld_32x32(t_row + c * 32, ra);
                        tmem_ld_32x32(t_row + 128 + c * 32, rb);
                        tmem_ld_wait();
                        const int j0 = j_base + c * 32;
                        uint32_t ha[16], hb[16];
#pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) {
                            const float4 vba = *reinterpret_cast<const float4*>(aux + j0 + 4 * i4);
                            const float4 vbb = *reinterpret_cast<const float4*>(aux + 2048 + j0 + 4 * i4);
                            const float4 vwc = *reinterpret_cast<const float4*>(aux + 4096 + j0 + 4 * i4);
                            const float fba[4] = {vba.x, vba.y, vba.z, vba.w}, fbb[4] = {vbb.x, vbb.y, vbb.z, vbb.w};
                            const float fwc[4] = {vwc.x, vwc.y, vwc.z, vwc.w};
                            float ma[4], mb[4];
                            const uint64_t idx4 = ((uint64_t)m * (uint64_t)HC + (uint64_t)(j0 + 4 * i4)) >> 2;
                            dropout_scale4(p.drop_p, p.seed, 10u, idx4, ma);
                            dropout_scale4(p.drop_p, p.seed, 11u, idx4, mb);
                            float av[4], bv[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                av[i] = tanh_acc(__uint_as_float(ra[4 * i4 + i]) + fba[i]) * ma[i];
                                bv[i] = sigmoid_acc(__uint_as_float(rb[4 * i4 + i]) + fbb[i]) * mb[i];
                                gated_partial = fmaf(av[i] * bv[i], fwc[i], gated_partial);
                            }
                            if (p.gate_a != nullptr) {
                                const __half2 a01 = __floats2half2_rn(av[0], av[1]), a23 = __floats2half2_rn(av[2], av[3]);
                                const __half2 b01 = __floats2half2_rn(bv[0], bv[1]), b23 = __floats2half2_rn(bv[2], bv[3]);
                                ha[2 * i4] = *reinterpret_cast<const uint32_t*>(&a01); ha[2 * i4 + 1] = *reinterpret_cast<const uint32_t*>(&a23);
                                hb[2 * i4] = *reinterpret_cast<const uint32_t*>(&b01); hb[2 * i4 + 1] = *reinterpret_cast<const uint32_t*>(&b23);
                            }
                        }
                        if (p.gate_a != nullptr && row_ok) {
                            uint4* da = reinterpret_cast<uint4*>(p.gate_a + (size_t)m * HC + j0);
                            uint4* db = reinterpret_cast<uint4*>(p.gate_b + (size_t)m * HC + j0);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                da[i] = make_uint4(ha[4 * i], ha[4 * i + 1], ha[4 * i + 2], ha[4 * i + 3]);
                                db[i] = make_uint4(hb[4 * i], hb[4 * i + 1], hb[4 * i + 2], hb[4 * i + 3]);
                            }
                        }
                    }
                    if (inner == p.n_inner - 1) {
                        // combine the two column halves of this row: half 1 hands its partial to half 0 through smem
                        float* part = aux + 6144 + unit_parity * 128;
                        if (half == 1) part[quad * 32 + lane] = gated_partial;
                        asm volatile("bar.sync 1, 256;" ::: "memory");
                        if (half == 0 && row_ok)
                            p.logits[(size_t)m * p.n_heads + head] = gated_partial + part[quad * 32 + lane] + __ldg(p.bc + head);
                    }
                }
                (void)have_k;
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
            unit_parity ^= 1;
        }
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
    static bool attr_set = false;  // per instantiation
    if (!attr_set) {
        MDL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    const int units = args.num_m_tiles * (args.num_n_tiles / args.n_inner) * args.ksplit;
    if (units == 0) return 0;
    int sms = kNumSMs;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = units < sms ? units : sms;
    kern<<<grid, GEMM_THREADS, L::TOTAL, stream>>>(ta, tb, args);
    MDL_CHECK_LAUNCH();
    return 0;
}

// The fused-pass schedule is the default for nsplit = 3; MDL_GEMM_FUSED=0 selects the pass-serial one (debug / A-B).
static bool use_fused() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MDL_GEMM_FUSED");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

}  // namespace mdl

using namespace mdl;

extern "C" {

int mdl_gemm_nt(const void* a_planes, long long a_rows, long long a_cols, long long lda, long long a_plane_stride,
                const void* b_planes, long long b_rows, long long b_cols, long long ldb, long long b_plane_stride,
                float* out, long long ldc, int M, int N, int K, int nsplit,
                int grp_n_cols, int a_koff,
                const float* bias, const float* rowbias, const int* row2bag, void* stream) {
    MDL_REQUIRE(nsplit == 1 || nsplit == 3, "nsplit must be 1 or 3");
    MDL_REQUIRE(K % BLOCK_K == 0, "K (%d) must be a multiple of %d", K, BLOCK_K);
    MDL_REQUIRE(N % 128 == 0, "N (%d) must be a multiple of 128", N);
    MDL_REQUIRE(M > 0, "M must be positive");
    const int planes = nsplit == 3 ? 2 : 1;
    const int bn = (N % 256 == 0) ? 256 : 128;
    const bool fused = nsplit == 3 && use_fused();
    const int bk = fused ? BLOCK_KF : BLOCK_K;
    CUtensorMap ta, tb;
    int rc = make_plane_tmap(&ta, a_planes, a_rows, a_cols, lda, a_plane_stride, planes, BLOCK_M, bk);
    if (rc) return rc;
    rc = make_plane_tmap(&tb, b_planes, b_rows, b_cols, ldb, b_plane_stride, planes, bn, bk);
    if (rc) return rc;
    GemmArgs g{};
    g.M = M; g.N = N; g.k_blocks = K / bk; g.nsplit = nsplit;
    g.num_m_tiles = (M + BLOCK_M - 1) / BLOCK_M; g.num_n_tiles = N / bn; g.n_inner = 1; g.ksplit = 1;
    MDL_REQUIRE(grp_n_cols <= 0 || grp_n_cols % bn == 0, "grp_n_cols (%d) must be a multiple of the N tile (%d)", grp_n_cols, bn);
    MDL_REQUIRE(ldc % 4 == 0, "ldc must be a multiple of 4");
    g.grp_n_tiles = grp_n_cols > 0 ? grp_n_cols / bn : (1 << 30); g.a_koff = a_koff;
    g.grp_m_tiles = 1 << 30; g.b_coff = 0;
    g.out = out; g.ldc = (int)ldc; g.bias = bias; g.rowbias = rowbias; g.row2bag = row2bag;
    if (fused) {
        if (bn == 256) return launch_gemm<256, MODE_KF, EPI_STORE>(ta, tb, g, (cudaStream_t)stream);
        return launch_gemm<128, MODE_KF, EPI_STORE>(ta, tb, g, (cudaStream_t)stream);
    }
    if (bn == 256) return launch_gemm<256, MODE_K, EPI_STORE>(ta, tb, g, (cudaStream_t)stream);
    return launch_gemm<128, MODE_K, EPI_STORE>(ta, tb, g, (cudaStream_t)stream);
}

int mdl_gemm_gated(const void* a_planes, long long a_rows, long long a_cols, long long lda, long long a_plane_stride,
                   const void* b_planes, long long b_plane_stride,
                   int M, int n_heads, int nsplit,
                   const float* ba, const float* bb, const float* wc, const float* bc,
                   float* logits, void* gate_a, void* gate_b, float drop_p, unsigned long long seed, void* stream) {
    MDL_REQUIRE(nsplit == 1 || nsplit == 3, "nsplit must be 1 or 3");
    MDL_REQUIRE(M > 0 && n_heads > 0, "bad sizes");
    const int planes = nsplit == 3 ? 2 : 1;
    const int K = 512, N = n_heads * 1024;
    const bool fused = nsplit == 3 && use_fused();
    const int bk = fused ? BLOCK_KF : BLOCK_K;
    CUtensorMap ta, tb;
    int rc = make_plane_tmap(&ta, a_planes, a_rows, a_cols, lda, a_plane_stride, planes, BLOCK_M, bk);
    if (rc) return rc;
    rc = make_plane_tmap(&tb, b_planes, N, K, K, b_plane_stride, planes, 256, bk);
    if (rc) return rc;
    GemmArgs g{};
    g.M = M; g.N = N; g.k_blocks = K / bk; g.nsplit = nsplit;
    g.num_m_tiles = (M + BLOCK_M - 1) / BLOCK_M; g.num_n_tiles = N / 256; g.n_inner = 4; g.ksplit = 1;
    g.grp_n_tiles = 4; g.a_koff = 512; g.grp_m_tiles = 1 << 30; g.b_coff = 0;
    g.ba = ba; g.bb = bb; g.wc = wc; g.bc = bc; g.logits = logits;
    g.gate_a = reinterpret_cast<__half*>(gate_a); g.gate_b = reinterpret_cast<__half*>(gate_b);
    g.drop_p = drop_p; g.seed = seed; g.n_heads = n_heads;
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
    CUtensorMap ta, tb;
    int rc = make_plane_tmap(&ta, a_planes, tokens, a_cols, lda, a_plane_stride, planes, BLOCK_K);
    if (rc) return rc;
    rc = make_plane_tmap(&tb, b_planes, tokens, b_cols, ldb, b_plane_stride, planes, BLOCK_K);
    if (rc) return rc;
    GemmArgs g{};
    g.M = M; g.N = N; g.k_blocks = (int)((tokens + BLOCK_K - 1) / BLOCK_K); g.nsplit = nsplit;
    g.num_m_tiles = M / BLOCK_M; g.num_n_tiles = N / 256; g.n_inner = 1;
    const int mn_tiles = g.num_m_tiles * g.num_n_tiles;
    if (ksplit <= 0) {
        // aim for ~4 work units per SM, never more splits than k-blocks
        ksplit = (4 * kNumSMs + mn_tiles - 1) / mn_tiles;
    }
    if (ksplit > g.k_blocks) ksplit = g.k_blocks;
    if (ksplit < 1) ksplit = 1;
    g.ksplit = ksplit;
    g.grp_n_tiles = 1 << 30; g.a_koff = 0;
    g.grp_m_tiles = grp_m_rows > 0 ? grp_m_rows / BLOCK_M : (1 << 30); g.b_coff = b_coff;
    g.out = out; g.ldc = (int)ldc;
    return launch_gemm<256, MODE_MN, EPI_ATOMIC>(ta, tb, g, (cudaStream_t)stream);
}

}  // extern "C"
