// Shared pieces of the tcgen05 GEMM kernels (single-CTA gemm_tcgen05.cu and CTA-pair gemm2_tcgen05.cu): argument block,
// shared-memory layout, and the epilogue that drains one accumulator tile from TMEM.
#pragma once
#include "common.cuh"
#include <cuda.h>

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
    int out_bf16;             // EPI_STORE only: `out` is a bf16 matrix (leading dimension ldc in ELEMENTS), values rounded RN
    const float* bias;        // [N] or null
    const float* rowbias;     // [R, N] or null (per-bag bias, stain encodings)
    const int* row2bag;       // [M]
    // gated epilogue
    const float* ba; const float* bb; const float* wc; const float* bc;  // [H*512], [H*512], [H*512], [H]
    float* logits;            // [M, H]
    __half* gate_a; __half* gate_b;  // [M, H*512] or null
    float drop_p; unsigned long long seed;
    int n_heads;
    int debug_flags;          // tools/gemm_bounds.py only: 1 = skip TMA loads, 2 = skip epilogue global writes
    int ab_f16;               // operand planes are fp16 hi/lo (inference format) instead of bf16 hi/lo
    float acc_scale;          // the accumulator is multiplied by this before bias / activations (1, or 1/MDL_F16_WEIGHT_SCALE)
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


// Saved gate activations (fp16 scratch between the gated-attention GEMM and mdl_gate_bwd) are stored TILED: element
// (row m, gate column j) of a [M, HC] matrix lives at  ((((m / 32) * (HC / 16) + j / 16) * 2 + (j % 16) / 8) * 32 + m % 32) * 8 + j % 8,
// i.e. blocks of 32 rows x 8 columns are contiguous.  The buffer holds ceil(M / 32) * 32 rows.
__host__ __device__ __forceinline__ size_t gate_tile_offset(long long m, int j, int HC) {
    return ((((size_t)(m >> 5) * (size_t)(HC >> 4) + (size_t)(j >> 4)) * 2 + (size_t)((j >> 3) & 1)) * 32 + (size_t)(m & 31)) * 8 + (size_t)(j & 7);
}

// Drain one [128 x BLOCK_N] accumulator tile (this CTA's TMEM lanes) for epilogue warp `warp_epi` (0..7):
// quadrant = warp_epi & 3 ... see callers; `row_base` is the global output row of TMEM lane 0 of this CTA.
// FAST_GATES (EPI_GATED, 1-pass modes only): tanh and sigmoid from ONE tanh.approx each (sigmoid(y) = 0.5 tanh(y / 2) + 0.5; 2^-11, below the
// bf16 operands' own 2^-9) instead of ex2 + rcp each: the bf16-mode gated GEMM is bound by its epilogue's MUFU work, not by
// its MMAs.  aux then holds ba and bb / 2 unscaled (see the kernels' prologues).
template <int BLOCK_N, int EPI, bool FAST_GATES = false>
__device__ __forceinline__ void epilogue_tile(const GemmArgs& p, float* aux, uint32_t t_row, int row_base, int quad, int half,
                                              int warp_epi, int lane, int n_tile, int n_group, int inner, bool have_k,
                                              float& gated_partial, int unit_parity) {
    const int m = row_base + quad * 32 + lane;
    const bool row_ok = m < p.M;
    if constexpr (EPI == EPI_STORE || EPI == EPI_ATOMIC) {
        // TMEM -> registers (thread = row) -> +bias -> warp-private padded smem -> row-contiguous global access:
        // 8 lanes cover one 128-byte row segment, so each vector store / reduction touches 4 full lines instead
        // of 32 scattered 16-byte pieces.
        int bag = 0;
        if (EPI == EPI_STORE && p.rowbias != nullptr && row_ok) bag = p.row2bag[m];
        float* stg = aux + warp_epi * (32 * 33);
        const int m_warp = row_base + quad * 32;
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
                if (EPI == EPI_STORE) {
                    // x * 1 and fma(x, 1, b) round exactly like x and x + b: the bf16-plane path is bit-for-bit what it was
                    const float sc = p.acc_scale;
                    if (p.bias != nullptr) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + i));
                        v.x = fmaf(v.x, sc, b.x); v.y = fmaf(v.y, sc, b.y); v.z = fmaf(v.z, sc, b.z); v.w = fmaf(v.w, sc, b.w);
                    } else {
                        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
                    }
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
                if (m_warp + row < p.M && !(p.debug_flags & 2)) {
                    float* dst = p.out + (size_t)(m_warp + row) * p.ldc + n0 + col4;
                    if constexpr (EPI == EPI_STORE) {
                        if (p.out_bf16) {
                            __nv_bfloat16* d16 = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)(m_warp + row) * p.ldc + n0 + col4;
                            const __nv_bfloat162 lo = __floats2bfloat162_rn(x0, x1), hi = __floats2bfloat162_rn(x2, x3);
                            *reinterpret_cast<uint2*>(d16) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
                        } else {
                            *reinterpret_cast<float4*>(dst) = make_float4(x0, x1, x2, x3);
                        }
                    } else {
                        if (have_k)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x0), "f"(x1), "f"(x2), "f"(x3) : "memory");
                    }
                }
            }
            __syncwarp();
        }
    } else {  // EPI_GATED
        // aux holds ba' = -2 log2(e) ba, bb' = -log2(e) bb (pre-scaled by the caller) and wc: tanh(x + ba) = 2 / (1 + 2^(c2 x + ba')) - 1
        // and sigmoid(y + bb) = 1 / (1 + 2^(c1 y + bb')) each start with ONE fma.  The 64 accumulator columns a warp owns per
        // gate branch are drained in four 16-column pieces, the load of piece k + 1 in flight while piece k is evaluated
        // (tcgen05.wait::ld waits for everything issued before it, so: wait, issue the next load, then compute).
        const int head = n_group;          // one work unit = (m_tile, head); inner = 128-wide gate group
        const int j_base = head * 512 + inner * 128 + half * 64;
        const int HC = p.n_heads * 512;
        const DropCfg dcfg = make_drop_cfg(p.drop_p);
        const bool drop8 = dcfg.on && drop_p_is_8bit(p.drop_p);
        const uint32_t thresh24 = ((uint32_t)(p.drop_p * 256.f)) << 24;
        const float C2 = -2.885390081777927f * p.acc_scale, C1 = -1.4426950408889634f * p.acc_scale;
        uint32_t ra[2][16], rb[2][16];
        const uint32_t t_a = t_row + half * 64, t_b = t_row + 128 + half * 64;
        tmem_ld_32x16(t_a, ra[0]);
        tmem_ld_32x16(t_b, rb[0]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            tmem_ld_wait();
            if (k < 3) {
                tmem_ld_32x16(t_a + (k + 1) * 16, ra[(k + 1) & 1]);
                tmem_ld_32x16(t_b + (k + 1) * 16, rb[(k + 1) & 1]);
            }
            const uint32_t (&xa)[16] = ra[k & 1];
            const uint32_t (&xb)[16] = rb[k & 1];
            const int j0 = j_base + k * 16;
            uint32_t ha[8], hb[8];
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
                const float4 vba = *reinterpret_cast<const float4*>(aux + j0 + 4 * i4);
                const float4 vbb = *reinterpret_cast<const float4*>(aux + 2048 + j0 + 4 * i4);
                const float4 vwc = *reinterpret_cast<const float4*>(aux + 4096 + j0 + 4 * i4);
                const float fba[4] = {vba.x, vba.y, vba.z, vba.w}, fbb[4] = {vbb.x, vbb.y, vbb.z, vbb.w};
                const float fwc[4] = {vwc.x, vwc.y, vwc.z, vwc.w};
                float ma[4], mb[4];
                const uint64_t idx4 = ((uint64_t)m * (uint64_t)HC + (uint64_t)(j0 + 4 * i4)) >> 2;
                if (drop8) {
                    dropout_scale4x2_8bit(thresh24, dcfg.keep, p.seed, 12u, idx4, ma, mb);
                } else {
                    dropout_scale4(dcfg, p.seed, 10u, idx4, ma);
                    dropout_scale4(dcfg, p.seed, 11u, idx4, mb);
                }
                float av[4], bv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if constexpr (FAST_GATES) {
                        av[i] = tanh_approx(fmaf(__uint_as_float(xa[4 * i4 + i]), p.acc_scale, fba[i])) * ma[i];
                        bv[i] = fmaf(0.5f, tanh_approx(fmaf(__uint_as_float(xb[4 * i4 + i]), 0.5f * p.acc_scale, fbb[i])), 0.5f) * mb[i];
                    } else {
                        const float ea = ex2_approx(fmaf(__uint_as_float(xa[4 * i4 + i]), C2, fba[i]));
                        const float eb = ex2_approx(fmaf(__uint_as_float(xb[4 * i4 + i]), C1, fbb[i]));
                        av[i] = fmaf(2.f, rcp_approx(1.f + ea), -1.f) * ma[i];
                        bv[i] = rcp_approx(1.f + eb) * mb[i];
                    }
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
                // Tiled scratch layout (gate_tile_offset): the 32 rows of this warp x 8 columns are 512 contiguous bytes, so
                // one 16-byte store per lane is 4 full lines instead of 32 scattered pieces (the LSU handles one line
                // per cycle: row-major stores cost 2 us of every 9.4 us tile).
                uint4* da = reinterpret_cast<uint4*>(p.gate_a + gate_tile_offset(m, j0, HC));
                uint4* db = reinterpret_cast<uint4*>(p.gate_b + gate_tile_offset(m, j0, HC));
                da[0] = make_uint4(ha[0], ha[1], ha[2], ha[3]); da[32] = make_uint4(ha[4], ha[5], ha[6], ha[7]);
                db[0] = make_uint4(hb[0], hb[1], hb[2], hb[3]); db[32] = make_uint4(hb[4], hb[5], hb[6], hb[7]);
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
}

// CTA-pair (cta_group::2) K-major fused-pass GEMM, gemm2_tcgen05.cu.  Tensor maps: A box [32 x 128 rows], B box [32 x 128 rows].
int launch_gemm2_kf(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args, int epi, cudaStream_t stream);
// CTA-pair MN-major (wgrad) variant: tensor maps with [64 cols x 32 rows] boxes, 128-byte swizzle; args.num_m_tiles in 256-row tiles.
int launch_gemm2_mn(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args, cudaStream_t stream);

}  // namespace mdl
