// HBM-bound elementwise / row-normalisation kernels around the tcgen05 GEMMs:
//   split planes (fp32 -> bf16 hi/lo), gather+split weight packing, LayerNorm+GELU(+dropout) forward and backward,
//   gated-attention backward, per-bag column sums, small index utilities.
// All are coalesced, 16-byte vectorised, grid-stride over rows with grids sized as a multiple of the SM count.
#include "common.cuh"
#include "madeleine_b200.h"
#include <stdlib.h>

namespace mdl {

// streaming 16-byte load that does not allocate in L1 (data touched once)
__device__ __forceinline__ uint4 ld_nc_na(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

static inline int grid_for(long long work_items, int per_block, int max_blocks_per_sm = 8) {
    long long b = (work_items + per_block - 1) / per_block;
    long long cap = (long long)kNumSMs * max_blocks_per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ---------------------------------------------------------------------------------------------------
// fp32 [rows, cols] (row stride ld) -> bf16 planes [nplanes][rows][cols]
// ---------------------------------------------------------------------------------------------------
template <bool F16>
__global__ void split_planes_kernel(const float* __restrict__ x, long long rows, int cols, long long ld,
                                    __nv_bfloat16* __restrict__ planes, long long plane_stride, int nplanes) {
    pdl_sync();
    const int vec_per_row = cols >> 2;
    const long long total = rows * vec_per_row;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / vec_per_row;
        const int c = (int)(i - r * vec_per_row) << 2;
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ld + c));
        uint32_t h01, l01, h23, l23;
        split_planes2<F16>(v.x, v.y, h01, l01);
        split_planes2<F16>(v.z, v.w, h23, l23);
        const long long o = r * cols + c;
        *reinterpret_cast<uint2*>(planes + o) = make_uint2(h01, h23);
        if (nplanes > 1)
            *reinterpret_cast<uint2*>(planes + plane_stride + o) = make_uint2(l01, l23);
    }
}

// packed[i] = split(src[idx[i]])  (weight packing: permutations / transposes expressed as an index map)
template <bool F16>
__global__ void gather_split_kernel(const float* __restrict__ src, const int* __restrict__ idx, long long n,
                                    __nv_bfloat16* __restrict__ planes, long long plane_stride, int nplanes) {
    pdl_sync();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float w = __ldg(src + __ldg(idx + i));
        if (F16) {
            // weights: scaled by a power of two so that the lo plane stays in fp16's normal range (undone in the GEMM epilogue)
            uint32_t hi, lo;
            split_f16x2(w * MDL_F16_WEIGHT_SCALE, 0.f, hi, lo);
            reinterpret_cast<unsigned short*>(planes)[i] = (unsigned short)(hi & 0xffffu);
            if (nplanes > 1) reinterpret_cast<unsigned short*>(planes)[plane_stride + i] = (unsigned short)(lo & 0xffffu);
        } else {
            __nv_bfloat16 h, l;
            split_bf16(w, h, l);
            planes[i] = h;
            if (nplanes > 1) planes[plane_stride + i] = l;
        }
    }
}
__global__ void gather_f32_kernel(const float* __restrict__ src, const int* __restrict__ idx, long long n, float* __restrict__ dst) {
    pdl_sync();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = __ldg(src + __ldg(idx + i));
}
// dst[idx[i]] (+)= src[i]; idx is injective so there are no write conflicts.
__global__ void scatter_f32_kernel(const float* __restrict__ src, const int* __restrict__ idx, long long n, float* __restrict__ dst, int accumulate) {
    pdl_sync();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int j = __ldg(idx + i);
        dst[j] = accumulate ? dst[j] + src[i] : src[i];
    }
}

__global__ void row2bag_kernel(const int* __restrict__ cu, int n_bags, int* __restrict__ row2bag, long long rows) {
    pdl_sync();
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < rows; m += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = n_bags;  // find bag with cu[bag] <= m < cu[bag+1]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(cu + mid) <= m) lo = mid; else hi = mid;
        }
        row2bag[m] = lo;
    }
}

// Four consecutive activations as fp32: the GEMM outputs that feed these kernels are fp32 in the fp32-grade modes and bf16
// in the bf16 mode (what torch autocast hands LayerNorm after an nn.Linear).  `off` is an ELEMENT offset.
template <bool BF16>
__device__ __forceinline__ float4 ld_act4(const void* base, size_t off) {
    if (BF16) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + off));
        return make_float4(bf16_lo_of(u.x), bf16_hi_of(u.x), bf16_lo_of(u.y), bf16_hi_of(u.y));
    }
    return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off));
}
template <bool BF16>
__device__ __forceinline__ void prefetch_act(const void* base, size_t off) {
    const char* p = reinterpret_cast<const char*>(base) + off * (BF16 ? 2 : 4);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm + exact GELU (+ dropout) forward.  One warp owns RPW whole rows (C/32 values per lane per row, float4
// columns j*128 + lane*4), so the two row reductions are shuffles only and RPW*C/128 16-byte loads are in flight
// per lane.  No shared memory, no block barriers.
// ---------------------------------------------------------------------------------------------------
template <int C, int RPW, bool ZBF16, bool F16>
__global__ void __launch_bounds__(256, 2)
ln_gelu_fwd_kernel(const void* __restrict__ z, int M, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float eps, float drop_p, unsigned long long seed, unsigned stream_id,
                   __nv_bfloat16* __restrict__ planes, long long plane_stride, int nplanes,
                   float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    pdl_sync();
    constexpr int V = C / 128;            // float4 per lane per row
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps_total = gridDim.x * 8;
    const DropCfg dcfg = make_drop_cfg(drop_p);
    const bool drop = dcfg.on;
    for (int m0 = (blockIdx.x * 8 + warp) * RPW; m0 < M; m0 += warps_total * RPW) {
        {   // next rows of this warp -> L2 (16 warps per SM do not cover the HBM round trip on their own)
            const int mp = m0 + warps_total * RPW;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                if (mp + r < M) {
#pragma unroll
                    for (int j = 0; j < V; ++j) prefetch_act<ZBF16>(z, (size_t)(mp + r) * C + lane * 4 + j * 128);
                }
            }
        }
        float4 v[RPW][V];
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int mr = m0 + r < M ? m0 + r : m0;                 // clamp: a duplicate row is loaded, never stored
#pragma unroll
            for (int j = 0; j < V; ++j) v[r][j] = ld_act4<ZBF16>(z, (size_t)mr * C + lane * 4 + j * 128);
        }
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int m = m0 + r;
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < V; ++j) s += (v[r][j].x + v[r][j].y) + (v[r][j].z + v[r][j].w);
            const float mu = warp_sum(s) * (1.f / C);
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const float a = v[r][j].x - mu, b = v[r][j].y - mu, c = v[r][j].z - mu, d = v[r][j].w - mu;
                q += (a * a + b * b) + (c * c + d * d);
            }
            const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
            if (m >= M) continue;
            if (lane == 0) { mean_out[m] = mu; rstd_out[m] = rstd; }
            const float nmr = -mu * rstd;
            const size_t row_off = (size_t)m * C + lane * 4;
            const float4* gp = reinterpret_cast<const float4*>(gamma) + lane;
            const float4* bp = reinterpret_cast<const float4*>(beta) + lane;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const float4 g = __ldg(gp + j * 32);
                const float4 be = __ldg(bp + j * 32);
                float y0 = gelu_erf(fmaf(fmaf(v[r][j].x, rstd, nmr), g.x, be.x));
                float y1 = gelu_erf(fmaf(fmaf(v[r][j].y, rstd, nmr), g.y, be.y));
                float y2 = gelu_erf(fmaf(fmaf(v[r][j].z, rstd, nmr), g.z, be.z));
                float y3 = gelu_erf(fmaf(fmaf(v[r][j].w, rstd, nmr), g.w, be.w));
                if (drop) {
                    float msk[4];
                    dropout_scale4(dcfg, seed, stream_id, (row_off + j * 128) >> 2, msk);
                    y0 *= msk[0]; y1 *= msk[1]; y2 *= msk[2]; y3 *= msk[3];
                }
                uint32_t h01, l01, h23, l23;
                split_planes2<F16>(y0, y1, h01, l01);
                split_planes2<F16>(y2, y3, h23, l23);
                __nv_bfloat16* o = planes + row_off + j * 128;
                *reinterpret_cast<uint2*>(o) = make_uint2(h01, h23);
                if (nplanes > 1) *reinterpret_cast<uint2*>(o + plane_stride) = make_uint2(l01, l23);
            }
        }
    }
}

// cp.async (LDGSTS): 16 bytes global -> shared without passing through registers; completion tracked per thread in groups
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------------
// LayerNorm + GELU (+dropout) backward.
//   dh = dh_a + dh_b + sum_v p_v[m, head(c)] * dS_v[seg_v(m), c]     (the last term fuses the pooling backward)
//   dy = dh * dropout_scale * gelu'(y),   y = xhat*gamma + beta,  xhat = (z - mean) * rstd
//   dz = rstd * (dy*gamma - mean_c(dy*gamma) - xhat * mean_c(dy*gamma*xhat))
// Writes dz as bf16 planes (operand of the following dgrad / wgrad GEMMs) and accumulates the column sums
// dgamma += dy*xhat, dbeta += dy, dbias += dz.
//
// Layout: a warp owns a 512-column slab of a row — lane l holds 16 FIXED columns (float4 at slab + 128 j + 4 l, j < 4)
// for the whole kernel, so the 48 column-sum accumulators live in registers and every per-row cost (the two shuffle
// reductions, mean / rstd / attention-weight loads, address arithmetic) is spread over 16 elements per thread.  C = 512:
// one warp per row, no shared memory and no barrier in the row loop.  C = 2048: the four warps of a row exchange their two
// partial sums through shared memory behind a 128-thread named barrier (double-buffered, one barrier per row).
// A block sweeps a CONTIGUOUS run of rows (few bag boundaries per thread for the per-bag sums).
// Optional inputs are template flags so the inner loop carries no pointer tests.
// ---------------------------------------------------------------------------------------------------
struct PoolTerm {
    const float* p;        // [M, H] attention probabilities of this view (0 outside the view)
    const float* dS;       // [n_seg, C]
    const int* row2seg;    // [M]
};

// HAS_B: 0 = no second dense gradient, 1 = dh_b is [M, C], 2 = dh_b is COMPACT [n_sel, C] and dh_b_rows[m] gives the compact
// row of token m (or -1): the token-projector gradient exists only for the token window the local loss can read.
// BAGSUM: additionally accumulate per-bag column sums of dz into bag_dz[row2bag[m], c] (the stain-encoding backward needs
// them); a thread flushes its bag accumulator only when the bag id changes.
// ASYNC (fp32 activations, no dense second gradient): the two row streams (z, dh_a) are staged by cp.async into a warp-private
// double buffer one row ahead — every lane copies exactly the 2 x 64 bytes it will read itself, so there is no cross-lane
// synchronisation, only cp.async.wait_group — instead of being requested into registers when they are needed (with an L2
// prefetch two rows ahead).  The loads then overlap the previous row's arithmetic without holding 32 registers per thread.
constexpr int LNB_ASYNC_STAGES = 2;
template <int C, int HAS_B, int NPOOL, bool BAGSUM, bool INBF16, bool ASYNC>
__global__ void __launch_bounds__(256, 2)
ln_gelu_bwd_kernel(const void* __restrict__ z, int M, const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ mean, const float* __restrict__ rstd_in,
                   const void* __restrict__ dh_a, const void* __restrict__ dh_b, const int* __restrict__ dh_b_rows,
                   PoolTerm pt0, PoolTerm pt1, int n_heads,
                   float drop_p, unsigned long long seed, unsigned stream_id,
                   __nv_bfloat16* __restrict__ dz_planes, long long plane_stride, int nplanes,
                   float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
                   const int* __restrict__ row2bag, float* __restrict__ bag_dz) {
    pdl_sync();
    constexpr int WPR = C / 512;          // warps per row
    constexpr int GROUPS = 8 / WPR;       // rows in flight per block
    __shared__ float red[2][GROUPS][2][WPR];
    __shared__ float colacc[3 * C];
    extern __shared__ __align__(16) float lnb_stage[];        // ASYNC: [8 warps][LNB_ASYNC_STAGES][2 streams][512] floats
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = warp / WPR, wq = warp % WPR;
    const int c0 = wq * 512 + lane * 4;                       // this lane's columns: c0 + 128 j + i
    const int cols_per_head = C / n_heads;
    // a warp's 512-column slab lies inside ONE head whenever a head is a whole number of slabs (always for the 2048-wide layer with
    // 4 heads): the attention weight of (row, head) is then one load per row, not a runtime division and a load per 128 columns
    const bool one_head = cols_per_head % 512 == 0;
    const int head_of_warp = one_head ? (wq * 512) / cols_per_head : 0;
    for (int i = tid; i < 3 * C; i += 256) colacc[i] = 0.f;
    float* my_stage = lnb_stage + (size_t)warp * (LNB_ASYNC_STAGES * 2 * 512) + lane * 4;
    float accg[16], accb[16], accz[16], accbag[BAGSUM ? 16 : 1];
#pragma unroll
    for (int i = 0; i < 16; ++i) { accg[i] = 0.f; accb[i] = 0.f; accz[i] = 0.f; }
    if (BAGSUM) {
#pragma unroll
        for (int i = 0; i < 16; ++i) accbag[BAGSUM ? i : 0] = 0.f;
    }
    int cur_bag = -1;
    const DropCfg dcfg = make_drop_cfg(drop_p);
    const int rows_per_block = (M + gridDim.x - 1) / gridDim.x;
    const int row_begin = blockIdx.x * rows_per_block;
    const int row_end = min(M, row_begin + rows_per_block);
    const int iters = (rows_per_block + GROUPS - 1) / GROUPS;
    auto stage_row = [&](int it_) {                            // ASYNC: request row `it_` of this warp into its stage (one group per row)
        const int mm = row_begin + it_ * GROUPS + grp;
        if (it_ < iters && mm < row_end) {
            float* dst = my_stage + (it_ % LNB_ASYNC_STAGES) * (2 * 512);
            const float* sz = reinterpret_cast<const float*>(z) + (size_t)mm * C + c0;
            const float* sd = reinterpret_cast<const float*>(dh_a) + (size_t)mm * C + c0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                cp_async16(dst + 128 * j, sz + 128 * j);
                cp_async16(dst + 512 + 128 * j, sd + 128 * j);
            }
        }
        cp_async_commit();
    };
    if (ASYNC) {
#pragma unroll
        for (int f = 0; f < LNB_ASYNC_STAGES - 1; ++f) stage_row(f);
    }
    for (int it = 0; it < iters; ++it) {
        const int m = row_begin + it * GROUPS + grp;
        const bool ok = m < row_end;
        const int mr = ok ? m : 0;                             // clamp; contributions of padded rows are zeroed below
        const size_t row_off = (size_t)mr * C + c0;
        if (ASYNC) {
            stage_row(it + LNB_ASYNC_STAGES - 1);               // the row after this one is in flight while this one is evaluated
        } else {   // pull the rows this warp will need two iterations from now into L2 (the kernel is latency-bound otherwise:
            // 128 registers per thread leave 16 warps per SM to cover the HBM round trip)
            const int mp = m + 2 * GROUPS;
            if (mp < row_end) {
                const size_t po = (size_t)mp * C + c0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    prefetch_act<INBF16>(z, po + 128 * j);
                    prefetch_act<INBF16>(dh_a, po + 128 * j);
                    if (HAS_B == 1) prefetch_act<INBF16>(dh_b, po + 128 * j);
                }
            }
        }
        float x[16], dx[16];                                    // xhat, then dy*gamma
        {
            float4 zv[4], dv[4];
            if (ASYNC) {
                cp_async_wait<LNB_ASYNC_STAGES - 1>();          // this row's group has landed (the newest one may still be in flight)
                const float* src = my_stage + (it % LNB_ASYNC_STAGES) * (2 * 512);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    zv[j] = ok ? *reinterpret_cast<const float4*>(src + 128 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    dv[j] = ok ? *reinterpret_cast<const float4*>(src + 512 + 128 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    zv[j] = ld_act4<INBF16>(z, row_off + 128 * j);
                    dv[j] = ld_act4<INBF16>(dh_a, row_off + 128 * j);
                }
            }
            if (HAS_B == 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 t = ld_act4<INBF16>(dh_b, row_off + 128 * j);
                    dv[j].x += t.x; dv[j].y += t.y; dv[j].z += t.z; dv[j].w += t.w;
                }
            }
            if (HAS_B == 2) {
                const int sel = __ldg(dh_b_rows + mr);
                if (sel >= 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 t = ld_act4<INBF16>(dh_b, (size_t)sel * C + c0 + 128 * j);
                        dv[j].x += t.x; dv[j].y += t.y; dv[j].z += t.z; dv[j].w += t.w;
                    }
                }
            }
            if (NPOOL >= 1) {
                const float* dS = pt0.dS + (size_t)__ldg(pt0.row2seg + mr) * C + c0;
                const float pw_row = one_head ? __ldg(pt0.p + (size_t)mr * n_heads + head_of_warp) : 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float pw = one_head ? pw_row : __ldg(pt0.p + (size_t)mr * n_heads + (c0 + 128 * j) / cols_per_head);
                    const float4 t = __ldg(reinterpret_cast<const float4*>(dS + 128 * j));
                    dv[j].x = fmaf(pw, t.x, dv[j].x); dv[j].y = fmaf(pw, t.y, dv[j].y);
                    dv[j].z = fmaf(pw, t.z, dv[j].z); dv[j].w = fmaf(pw, t.w, dv[j].w);
                }
            }
            if (NPOOL >= 2) {
                const float* dS = pt1.dS + (size_t)__ldg(pt1.row2seg + mr) * C + c0;
                const float pw_row = one_head ? __ldg(pt1.p + (size_t)mr * n_heads + head_of_warp) : 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float pw = one_head ? pw_row : __ldg(pt1.p + (size_t)mr * n_heads + (c0 + 128 * j) / cols_per_head);
                    const float4 t = __ldg(reinterpret_cast<const float4*>(dS + 128 * j));
                    dv[j].x = fmaf(pw, t.x, dv[j].x); dv[j].y = fmaf(pw, t.y, dv[j].y);
                    dv[j].z = fmaf(pw, t.z, dv[j].z); dv[j].w = fmaf(pw, t.w, dv[j].w);
                }
            }
            const float r_ = ok ? __ldg(rstd_in + mr) : 0.f;
            const float nmr = -__ldg(mean + mr) * r_;
            const float okf = ok ? 1.f : 0.f;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 128 * j));
                const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c0 + 128 * j));
                float msk[4] = {okf, okf, okf, okf};
                if (dcfg.on) {
                    dropout_scale4(dcfg, seed, stream_id, (row_off + 128 * j) >> 2, msk);
#pragma unroll
                    for (int i = 0; i < 4; ++i) msk[i] *= okf;
                }
                const float zz[4] = {zv[j].x, zv[j].y, zv[j].z, zv[j].w};
                const float dd[4] = {dv[j].x, dv[j].y, dv[j].z, dv[j].w};
                const float gg[4] = {g.x, g.y, g.z, g.w};
                const float bb[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = 4 * j + i;
                    const float xh = fmaf(zz[i], r_, nmr);
                    const float t = dd[i] * msk[i] * gelu_erf_grad(fmaf(xh, gg[i], bb[i]));
                    const float dxg = t * gg[i];
                    x[k] = xh; dx[k] = dxg;
                    s1 += dxg;
                    s2 = fmaf(dxg, xh, s2);
                    accg[k] = fmaf(t, xh, accg[k]);
                    accb[k] += t;
                }
            }
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if (WPR > 1) {
                const int buf = it & 1;
                if (lane == 0) { red[buf][grp][0][wq] = s1; red[buf][grp][1][wq] = s2; }
                asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(WPR * 32) : "memory");
                s1 = 0.f; s2 = 0.f;
#pragma unroll
                for (int w = 0; w < WPR; ++w) { s1 += red[buf][grp][0][w]; s2 += red[buf][grp][1][w]; }
            }
            const float m1 = s1 * (1.f / C), m2 = s2 * (1.f / C);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                dx[k] = r_ * (dx[k] - m1 - x[k] * m2);             // dz; r_ = 0 for rows past the end
                accz[k] += dx[k];
            }
        }
        if (ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t h01, l01, h23, l23;
                split_bf16x2(dx[4 * j], dx[4 * j + 1], h01, l01);
                split_bf16x2(dx[4 * j + 2], dx[4 * j + 3], h23, l23);
                __nv_bfloat16* o = dz_planes + row_off + 128 * j;
                *reinterpret_cast<uint2*>(o) = make_uint2(h01, h23);
                if (nplanes > 1) *reinterpret_cast<uint2*>(o + plane_stride) = make_uint2(l01, l23);
            }
            if (BAGSUM) {
                const int bag = __ldg(row2bag + m);
                if (bag != cur_bag) {
                    if (cur_bag >= 0) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                atomicAdd(bag_dz + (size_t)cur_bag * C + c0 + 128 * j + i, accbag[BAGSUM ? 4 * j + i : 0]);
                                accbag[BAGSUM ? 4 * j + i : 0] = 0.f;
                            }
                    }
                    cur_bag = bag;
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) accbag[BAGSUM ? k : 0] += dx[k];
            }
        }
    }
    if (BAGSUM && cur_bag >= 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) atomicAdd(bag_dz + (size_t)cur_bag * C + c0 + 128 * j + i, accbag[BAGSUM ? 4 * j + i : 0]);
    }
    // column sums: the GROUPS warps that own the same columns combine in shared memory, then one atomic per column per block
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = c0 + 128 * j + i, k = 4 * j + i;
            atomicAdd(colacc + c, accg[k]);
            atomicAdd(colacc + C + c, accb[k]);
            atomicAdd(colacc + 2 * C + c, accz[k]);
        }
    __syncthreads();
    for (int c = tid; c < C; c += 256) {
        atomicAdd(dgamma + c, colacc[c]);
        atomicAdd(dbeta + c, colacc[C + c]);
        atomicAdd(dbias + c, colacc[2 * C + c]);
    }
}

// ---------------------------------------------------------------------------------------------------
// Gated-attention backward (elementwise part).  gate_a/gate_b hold the (dropout-scaled) tanh / sigmoid outputs as fp16 in
// the TILED scratch layout the gated GEMM's epilogue writes (32 rows x 8 columns contiguous, see gate_tile_offset).
// Produces d(pre-activation) as bf16 planes in the packed column order of the gated GEMM (per head: 4 groups of
// [128 a-cols | 128 b-cols]) and the column sums d(ba), d(bb), d(wc), d(bc).
//
// A block owns ONE 64-column slice of the gate columns (grid.x % 32) and sweeps 32-row blocks; warp w owns the slice's
// 8-column slab w and lane = row, so a warp's load is one contiguous 512-byte slab of the tiled layout and every thread keeps
// the same 8 columns — and their 24 column-sum accumulators — for the whole kernel (one shuffle reduction over the rows at the
// end).  The results of a row block are transposed through (double-buffered) shared memory so that the row-major dpre planes
// are written in 128-byte row segments.
//
// The dropout masks are NOT regenerated: a dropped gate was stored as an exact zero, and every product below that must
// vanish for a dropped gate does so by itself except the tanh branch's own derivative, which one select on `a == 0`
// handles.  (A KEPT tanh output that underflows fp16 — |tanh| < 3e-8, < 1e-7 of the elements — is thereby treated as
// dropped; the gates themselves carry 2^-11 per element, so this is far inside the precision of what was saved.)
//   dpre_a = dl wc (b mb) ma (1 - a^2),  dpre_b = dl wc (a ma) mb b (1 - b),  ma, mb in {0, 1/(1-p)}
// ---------------------------------------------------------------------------------------------------
constexpr int GB_COLS = 32;                 // gate columns per block (4 warps x 8 columns)
constexpr int GB_THREADS = GB_COLS * 4;
constexpr int GB_PITCH = GB_COLS + 8;       // bf16 elements per staged row (144 bytes: 16-byte accesses of consecutive rows hit distinct banks)
template <int NPL>
__global__ void __launch_bounds__(GB_THREADS, 6)
gate_bwd_kernel(const __half* __restrict__ gate_a, const __half* __restrict__ gate_b, const float* __restrict__ dlogit,
                const float* __restrict__ wc, long long M, int n_heads, float drop_p,
                __nv_bfloat16* __restrict__ dpre, long long plane_stride,
                float* __restrict__ dba, float* __restrict__ dbb, float* __restrict__ dwc, float* __restrict__ dbc) {
    pdl_sync();
    __shared__ __align__(16) __nv_bfloat16 stage[2][NPL][2][32][GB_PITCH];     // double-buffered: one barrier per row block
    const int HC = n_heads * 512;
    const int groups = HC / GB_COLS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cg = blockIdx.x % groups;                      // GB_COLS-column slice of a 128-column gate group
    const int jg = cg * GB_COLS;                             // first gate column of the slice
    const int head = jg >> 9, grp = (jg >> 7) & 3;
    const int j0 = jg + warp * 8;                            // this thread's 8 gate columns (one 32 x 8 slab of the tiled layout per row block)
    const int pcol_a = head * 1024 + grp * 256 + (jg & 127); // packed column of the slice's a-part; b-part is +128
    const bool drop = drop_p > 0.f;
    const float keep_inv = drop ? (1.f - drop_p) : 1.f;      // undoes the 1/(1-p) scaling of a stored gate
    const float keep = drop ? __fdividef(1.f, 1.f - drop_p) : 1.f;
    float kw[8], s_a[8], s_b[8], s_w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { kw[i] = __ldg(wc + j0 + i) * keep; s_a[i] = 0.f; s_b[i] = 0.f; s_w[i] = 0.f; }
    float s_c = 0.f;
    const bool count_c = warp == 0 && (jg & 511) == 0;      // one warp per head sums dlogit
    const long long n_rb = (M + 31) >> 5;
    const int chunks = gridDim.x / groups, chunk = blockIdx.x / groups;
    // slab of row block rb: ((rb * (HC/16) + j0/16) * 2 + (j0%16)/8) * 32 rows * 8 halfs, + lane * 8
    const size_t slab0 = ((size_t)(j0 >> 4) * 2 + (size_t)((j0 >> 3) & 1)) * 256 + (size_t)lane * 8;
    const size_t rb_stride = (size_t)(HC / 16) * 512;
    // software pipeline: the loads of the next two row blocks are in flight while one is evaluated
    constexpr int PF = 2;
    uint4 qa[PF], qb[PF];
    float qdl[PF];
#pragma unroll
    for (int f = 0; f < PF; ++f) {
        const long long rb = chunk + (long long)f * chunks;
        qa[f] = make_uint4(0, 0, 0, 0); qb[f] = make_uint4(0, 0, 0, 0); qdl[f] = 0.f;
        if (rb < n_rb) {
            qa[f] = ld_nc_na(reinterpret_cast<const uint4*>(gate_a + rb * rb_stride + slab0));
            qb[f] = ld_nc_na(reinterpret_cast<const uint4*>(gate_b + rb * rb_stride + slab0));
            const long long mm = rb * 32 + lane;
            qdl[f] = mm < M ? __ldg(dlogit + mm * n_heads + head) : 0.f;
        }
    }
    int buf = 0;
    for (long long rb = chunk; rb < n_rb; rb += chunks) {
        const long long m = rb * 32 + lane;
        const bool ok = m < M;
        const uint4 ua = qa[0], ub = qb[0];
        const float dl = qdl[0];
#pragma unroll
        for (int f = 0; f + 1 < PF; ++f) { qa[f] = qa[f + 1]; qb[f] = qb[f + 1]; qdl[f] = qdl[f + 1]; }
        {
            const long long rn = rb + (long long)PF * chunks;
            if (rn < n_rb) {
                qa[PF - 1] = ld_nc_na(reinterpret_cast<const uint4*>(gate_a + rn * rb_stride + slab0));
                qb[PF - 1] = ld_nc_na(reinterpret_cast<const uint4*>(gate_b + rn * rb_stride + slab0));
                const long long mm = rn * 32 + lane;
                qdl[PF - 1] = mm < M ? __ldg(dlogit + mm * n_heads + head) : 0.f;
            }
        }
        if (count_c) s_c += dl;
        float dpa[8], dpb[8];
        const __half2* ha = reinterpret_cast<const __half2*>(&ua);
        const __half2* hb = reinterpret_cast<const __half2*>(&ub);
#pragma unroll
        for (int i2 = 0; i2 < 4; ++i2) {
            const float2 fa = __half22float2(ha[i2]);
            const float2 fb = __half22float2(hb[i2]);
            const float adv[2] = {fa.x, fa.y}, bdv[2] = {fb.x, fb.y};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int i = 2 * i2 + k;
                const float ad = ok ? adv[k] : 0.f, bd = ok ? bdv[k] : 0.f;   // dropout-scaled gates (0 where dropped); padding rows contribute nothing
                const float a = ad * keep_inv, b = bd * keep_inv;
                const float t = dl * kw[i];                                   // dl wc / (1-p): the surviving mask factor of either branch
                const float da = (t * bd) * fmaf(-a, a, 1.f);
                dpa[i] = (drop && ad == 0.f) ? 0.f : da;
                dpb[i] = (t * ad) * fmaf(-b, b, b);
                s_a[i] += dpa[i]; s_b[i] += dpb[i]; s_w[i] = fmaf(dl, ad * bd, s_w[i]);
            }
        }
        // registers -> shared (this thread: row `lane`, columns warp*8 .. +7 of the slice), hi / lo planes
        {
            uint32_t ah[4], al[4], bh[4], bl[4];
#pragma unroll
            for (int i2 = 0; i2 < 4; ++i2) {
                split_bf16x2(dpa[2 * i2], dpa[2 * i2 + 1], ah[i2], al[i2]);
                split_bf16x2(dpb[2 * i2], dpb[2 * i2 + 1], bh[i2], bl[i2]);
            }
            const int c = warp * 8;
            *reinterpret_cast<uint4*>(&stage[buf][0][0][lane][c]) = make_uint4(ah[0], ah[1], ah[2], ah[3]);
            *reinterpret_cast<uint4*>(&stage[buf][0][1][lane][c]) = make_uint4(bh[0], bh[1], bh[2], bh[3]);
            if (NPL > 1) {
                *reinterpret_cast<uint4*>(&stage[buf][NPL - 1][0][lane][c]) = make_uint4(al[0], al[1], al[2], al[3]);
                *reinterpret_cast<uint4*>(&stage[buf][NPL - 1][1][lane][c]) = make_uint4(bl[0], bl[1], bl[2], bl[3]);
            }
        }
        __syncthreads();     // the only barrier of a row block: the other buffer is rewritten one iteration later, after every thread
                             // has passed this barrier again, i.e. finished reading it
        // shared -> global: GB_COLS / 8 lanes cover one row segment; the block's threads = all 32 rows
        {
            const int row = tid / (GB_COLS / 8), c = (tid % (GB_COLS / 8)) * 8;
            const long long mr = rb * 32 + row;
            if (mr < M) {
#pragma unroll
                for (int pl = 0; pl < NPL; ++pl)
#pragma unroll
                    for (int br = 0; br < 2; ++br) {
                        const uint4 v = *reinterpret_cast<const uint4*>(&stage[buf][pl][br][row][c]);
                        __nv_bfloat16* o = dpre + (size_t)pl * (size_t)plane_stride + (size_t)mr * (size_t)(n_heads * 1024) + pcol_a + br * 128 + c;
                        *reinterpret_cast<uint4*>(o) = v;
                    }
            }
        }
        buf ^= 1;
    }
    // column sums over the rows (lanes) of this warp's 8 columns
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float va = warp_sum(s_a[i]), vb = warp_sum(s_b[i]), vw = warp_sum(s_w[i]);
        if (lane == 0) {
            atomicAdd(dba + j0 + i, va);
            atomicAdd(dbb + j0 + i, vb);
            atomicAdd(dwc + j0 + i, vw);
        }
    }
    if (count_c) {
        s_c = warp_sum(s_c);
        if (lane == 0) atomicAdd(dbc + head, s_c);
    }
}

// out[p][s, :] = planes[p][rows[s], :]  — row gather of a bf16 planes tensor (token window of the local loss: only the
// first few tokens of every bag go through token_projector).  One warp per selected row, 16-byte vectors.
__global__ void __launch_bounds__(256)
gather_rows_planes_kernel(const __nv_bfloat16* __restrict__ planes, long long plane_stride_in, int nplanes, int C,
                          const int* __restrict__ rows, long long n_sel, __nv_bfloat16* __restrict__ out, long long plane_stride_out) {
    pdl_sync();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int vec = C >> 3;                                   // uint4 per row
    for (long long s = (long long)blockIdx.x * 8 + warp; s < n_sel * nplanes; s += (long long)gridDim.x * 8) {
        const int p = (int)(s / n_sel);
        const long long r = s - (long long)p * n_sel;
        const uint4* src = reinterpret_cast<const uint4*>(planes + p * plane_stride_in + (long long)__ldg(rows + r) * C);
        uint4* dst = reinterpret_cast<uint4*>(out + p * plane_stride_out + r * C);
        for (int v = lane; v < vec; v += 32) dst[v] = __ldg(src + v);
    }
}

// G[bag, c] = sum_{t in bag} (hi + lo)[t, c]  — per-bag column sums of a planes tensor (stain-encoding backward).
__global__ void __launch_bounds__(128)
bag_colsum_planes_kernel(const __nv_bfloat16* __restrict__ planes, long long plane_stride, int nplanes, int C,
                         const int* __restrict__ cu, float* __restrict__ out) {
    pdl_sync();
    const int bag = blockIdx.y;
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= C) return;
    const int t0 = cu[bag], t1 = cu[bag + 1];
    float s = 0.f;
    for (int t = t0; t < t1; ++t) {
        float v = __bfloat162float(planes[(long long)t * C + c]);
        if (nplanes > 1) v += __bfloat162float(planes[plane_stride + (long long)t * C + c]);
        s += v;
    }
    out[(long long)bag * C + c] = s;
}

// rowbias[r, n] = sum_s emb[code[r], s] * W1[n, d_in + s]   (stain encodings folded into a per-bag bias)
__global__ void stain_rowbias_kernel(const float* __restrict__ emb, const int* __restrict__ code, const float* __restrict__ w1,
                                     int ldw, int d_in, int se_dim, int n_out, float* __restrict__ rowbias) {
    pdl_sync();
    const int r = blockIdx.x;
    const float* e = emb + (long long)code[r] * se_dim;
    for (int n = threadIdx.x; n < n_out; n += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < se_dim; ++k) s = fmaf(__ldg(e + k), __ldg(w1 + (long long)n * ldw + d_in + k), s);
        rowbias[(long long)r * n_out + n] = s;
    }
}
// Backward of the per-bag bias: G [R, n_out] = per-bag column sums of dz1.
//   dW1[n, d_in + s] += sum_r G[r, n] emb[code r, s];   demb[code r, s] += sum_n G[r, n] W1[n, d_in + s]
__global__ void stain_rowbias_bwd_kernel(const float* __restrict__ G, const float* __restrict__ emb, const int* __restrict__ code,
                                         const float* __restrict__ w1, int ldw, int d_in, int se_dim, int n_out, int R,
                                         float* __restrict__ dw1, float* __restrict__ demb) {
    pdl_sync();
    // grid.x = n_out blocks for dW1 rows, then R blocks for demb rows
    const int bid = blockIdx.x;
    if (bid < n_out) {
        const int n = bid;
        for (int s = threadIdx.x; s < se_dim; s += blockDim.x) {
            float acc = 0.f;
            for (int r = 0; r < R; ++r) acc = fmaf(__ldg(G + (long long)r * n_out + n), __ldg(emb + (long long)code[r] * se_dim + s), acc);
            dw1[(long long)n * ldw + d_in + s] += acc;
        }
    } else {
        const int r = bid - n_out;
        for (int s = threadIdx.x; s < se_dim; s += blockDim.x) {
            float acc = 0.f;
            for (int n = 0; n < n_out; ++n) acc = fmaf(__ldg(G + (long long)r * n_out + n), __ldg(w1 + (long long)n * ldw + d_in + s), acc);
            atomicAdd(demb + (long long)code[r] * se_dim + s, acc);
        }
    }
}

// out[c] (+)= sum_m x[m, c]  for a small fp32 matrix (bias grads of the skinny projections).
__global__ void colsum_f32_kernel(const float* __restrict__ x, long long M, int C, float* __restrict__ out) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (long long m = blockIdx.y; m < M; m += gridDim.y) s += __ldg(x + m * C + c);
    atomicAdd(out + c, s);
}

}  // namespace mdl

using namespace mdl;

extern "C" {

int mdl_split_planes(const float* x, long long rows, int cols, long long ld, void* planes, long long plane_stride, int nplanes, void* stream) {
    MDL_REQUIRE(cols % 4 == 0 && ld % 4 == 0, "split_planes: cols and ld must be multiples of 4");
    if (rows == 0) return 0;
    const long long total = rows * (cols / 4);
    const bool f16 = (nplanes & kPlanesF16) != 0;
    nplanes &= 0xff;
    if (f16) launch_k(split_planes_kernel<true>, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, x, rows, cols, ld, (__nv_bfloat16*)planes, plane_stride, nplanes);
    else launch_k(split_planes_kernel<false>, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, x, rows, cols, ld, (__nv_bfloat16*)planes, plane_stride, nplanes);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_gather_split(const float* src, const int* idx, long long n, void* planes, long long plane_stride, int nplanes, void* stream) {
    if (n == 0) return 0;
    const bool f16 = (nplanes & kPlanesF16) != 0;
    nplanes &= 0xff;
    if (f16) launch_k(gather_split_kernel<true>, dim3(grid_for(n, 256)), dim3(256), 0, (cudaStream_t)stream, src, idx, n, (__nv_bfloat16*)planes, plane_stride, nplanes);
    else launch_k(gather_split_kernel<false>, dim3(grid_for(n, 256)), dim3(256), 0, (cudaStream_t)stream, src, idx, n, (__nv_bfloat16*)planes, plane_stride, nplanes);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_gather_f32(const float* src, const int* idx, long long n, float* dst, void* stream) {
    if (n == 0) return 0;
    launch_k(gather_f32_kernel, dim3(grid_for(n, 256)), dim3(256), 0, (cudaStream_t)stream, src, idx, n, dst);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_scatter_f32(const float* src, const int* idx, long long n, float* dst, int accumulate, void* stream) {
    if (n == 0) return 0;
    launch_k(scatter_f32_kernel, dim3(grid_for(n, 256)), dim3(256), 0, (cudaStream_t)stream, src, idx, n, dst, accumulate);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_row2bag(const int* cu_seqlens, int n_bags, int* row2bag, long long rows, void* stream) {
    if (rows == 0) return 0;
    launch_k(row2bag_kernel, dim3(grid_for(rows, 256)), dim3(256), 0, (cudaStream_t)stream, cu_seqlens, n_bags, row2bag, rows);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_ln_gelu_fwd(const void* z, long long M, int C, const float* gamma, const float* beta, float eps,
                    float drop_p, unsigned long long seed, unsigned stream_id,
                    void* planes, long long plane_stride, int nplanes, float* mean, float* rstd, int z_bf16, void* stream) {
    MDL_REQUIRE(C == 512 || C == 2048, "ln_gelu_fwd: C must be 512 or 2048 (got %d)", C);
    MDL_REQUIRE(M < (1LL << 31), "ln_gelu_fwd: too many rows");
    if (M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool f16 = (nplanes & kPlanesF16) != 0;
    nplanes &= 0xff;
    MDL_REQUIRE(!(f16 && z_bf16), "ln_gelu_fwd: fp16 planes are the fp32-grade inference format (z must be fp32)");
#define MDL_LNF(CC, RPW, grid) \
    do { \
        if (z_bf16) launch_k(ln_gelu_fwd_kernel<CC, RPW, true, false>, dim3(grid), dim3(256), 0, st, z, (int)M, gamma, beta, eps, drop_p, seed, stream_id, (__nv_bfloat16*)planes, plane_stride, nplanes, mean, rstd); \
        else if (f16) launch_k(ln_gelu_fwd_kernel<CC, RPW, false, true>, dim3(grid), dim3(256), 0, st, z, (int)M, gamma, beta, eps, drop_p, seed, stream_id, (__nv_bfloat16*)planes, plane_stride, nplanes, mean, rstd); \
        else launch_k(ln_gelu_fwd_kernel<CC, RPW, false, false>, dim3(grid), dim3(256), 0, st, z, (int)M, gamma, beta, eps, drop_p, seed, stream_id, (__nv_bfloat16*)planes, plane_stride, nplanes, mean, rstd); \
    } while (0)
    if (C == 512) {
        const int grid = grid_for(M, 8 * 2, 6);
        MDL_LNF(512, 2, grid);
    } else {
        const int grid = grid_for(M, 8, 4);
        MDL_LNF(2048, 1, grid);
    }
#undef MDL_LNF
    MDL_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"

// MDL_LN_BWD_ASYNC=0 selects the register-held loads for A/B measurements
static const bool g_ln_bwd_async = [] { const char* e = getenv("MDL_LN_BWD_ASYNC"); return !(e && e[0] == '0'); }();

template <int C, bool INBF16>
static void launch_ln_bwd(int has_b, int npool, int grid, cudaStream_t st, const void* z, int M, const float* gamma, const float* beta,
                          const float* mean, const float* rstd, const void* dh_a, const void* dh_b, const int* dh_b_rows,
                          PoolTerm t0, PoolTerm t1, int n_heads,
                          float drop_p, unsigned long long seed, unsigned stream_id, __nv_bfloat16* dz, long long ps, int npl,
                          float* dgamma, float* dbeta, float* dbias, const int* row2bag, float* bag_dz) {
    // fp32 activations without a dense second gradient take the cp.async-staged variant (64 KB of dynamic shared memory)
    constexpr int kStageBytes = 8 * LNB_ASYNC_STAGES * 2 * 512 * 4;
#define MDL_LN_BWD(HB, NP, BS)                                                                                                          \
    do {                                                                                                                                \
        if (!INBF16 && (HB) != 1 && g_ln_bwd_async) {                                                                                   \
            auto kern = ln_gelu_bwd_kernel<C, HB, NP, BS, INBF16, !INBF16 && (HB) != 1>;                                                \
            static PerDeviceOnce attr_set;                                                                                              \
            unsigned long long dev_bit;                                                                                                 \
            if (attr_set.needed(dev_bit)) {                                                                                             \
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes);                                   \
                attr_set.mark(dev_bit);                                                                                                 \
            }                                                                                                                           \
            launch_k(kern, dim3(grid), dim3(256), kStageBytes, st, z, M, gamma, beta, mean, rstd, dh_a, dh_b, dh_b_rows, t0, t1, n_heads, drop_p, seed,   \
                                                 stream_id, dz, ps, npl, dgamma, dbeta, dbias, row2bag, bag_dz);                        \
        } else {                                                                                                                        \
            launch_k(ln_gelu_bwd_kernel<C, HB, NP, BS, INBF16, false>, dim3(grid), dim3(256), 0, st, z, M, gamma, beta, mean, rstd, dh_a, dh_b, dh_b_rows, \
                                                                                  t0, t1, n_heads, drop_p, seed, stream_id, dz, ps,    \
                                                                                  npl, dgamma, dbeta, dbias, row2bag, bag_dz);          \
        }                                                                                                                               \
    } while (0)
    if constexpr (C == 512) {
        if (bag_dz != nullptr) { MDL_LN_BWD(0, 0, true); return; }   // only the first layer (no second gradient, no pooling term)
    }
    if (has_b == 0 && npool == 0) MDL_LN_BWD(0, 0, false);
    else if (has_b == 0 && npool == 1) MDL_LN_BWD(0, 1, false);
    else if (has_b == 0) MDL_LN_BWD(0, 2, false);
    else if (has_b == 1 && npool == 0) MDL_LN_BWD(1, 0, false);
    else if (has_b == 1 && npool == 1) MDL_LN_BWD(1, 1, false);
    else if (has_b == 1) MDL_LN_BWD(1, 2, false);
    else if (npool == 0) MDL_LN_BWD(2, 0, false);
    else if (npool == 1) MDL_LN_BWD(2, 1, false);
    else MDL_LN_BWD(2, 2, false);
#undef MDL_LN_BWD
}

extern "C" {

int mdl_ln_gelu_bwd(const void* z, long long M, int C, const float* gamma, const float* beta, const float* mean, const float* rstd,
                    const void* dh_a, const void* dh_b, const int* dh_b_rows,
                    const float* pool_p0, const float* pool_dS0, const int* pool_seg0,
                    const float* pool_p1, const float* pool_dS1, const int* pool_seg1, int n_heads,
                    float drop_p, unsigned long long seed, unsigned stream_id,
                    void* dz_planes, long long plane_stride, int nplanes,
                    float* dgamma, float* dbeta, float* dbias, const int* row2bag, float* bag_dz, int in_bf16, void* stream) {
    MDL_REQUIRE(C == 512 || C == 2048, "ln_gelu_bwd: C must be 512 or 2048 (got %d)", C);
    MDL_REQUIRE(n_heads > 0 && C % n_heads == 0 && (C / n_heads) % 4 == 0, "ln_gelu_bwd: bad n_heads");
    MDL_REQUIRE(M < (1LL << 31), "ln_gelu_bwd: too many rows");
    MDL_REQUIRE(dh_a != nullptr || dh_b != nullptr, "ln_gelu_bwd: at least one dense upstream gradient is required");
    MDL_REQUIRE(pool_p0 != nullptr || pool_p1 == nullptr, "ln_gelu_bwd: pool term 1 given without pool term 0");
    MDL_REQUIRE(dh_b_rows == nullptr || (dh_a != nullptr && dh_b != nullptr), "ln_gelu_bwd: a row-indexed dh_b needs a dense dh_a");
    MDL_REQUIRE(bag_dz == nullptr || (row2bag != nullptr && dh_b == nullptr && pool_p0 == nullptr),
                "ln_gelu_bwd: per-bag sums are built for the first layer only (no dh_b, no pooling term) and need row2bag");
    if (M == 0) return 0;
    if (dh_a == nullptr) { dh_a = dh_b; dh_b = nullptr; }
    const int has_b = dh_b == nullptr ? 0 : (dh_b_rows == nullptr ? 1 : 2);
    PoolTerm t0{pool_p0, pool_dS0, pool_seg0}, t1{pool_p1, pool_dS1, pool_seg1};
    const int npool = pool_p0 == nullptr ? 0 : (pool_p1 == nullptr ? 1 : 2);
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 512) {
        const int grid = grid_for(M, 8 * 4, 2);       // two blocks per SM, several rows per warp so the column atomics amortise
        if (in_bf16) launch_ln_bwd<512, true>(has_b, npool, grid, st, z, (int)M, gamma, beta, mean, rstd, dh_a, dh_b, dh_b_rows, t0, t1, n_heads, drop_p, seed, stream_id,
                                              (__nv_bfloat16*)dz_planes, plane_stride, nplanes, dgamma, dbeta, dbias, row2bag, bag_dz);
        else launch_ln_bwd<512, false>(has_b, npool, grid, st, z, (int)M, gamma, beta, mean, rstd, dh_a, dh_b, dh_b_rows, t0, t1, n_heads, drop_p, seed, stream_id,
                                       (__nv_bfloat16*)dz_planes, plane_stride, nplanes, dgamma, dbeta, dbias, row2bag, bag_dz);
    } else {
        const int grid = grid_for(M, 2 * 8, 2);
        MDL_REQUIRE(bag_dz == nullptr, "ln_gelu_bwd: per-bag sums are only built for C == 512");
        if (in_bf16) launch_ln_bwd<2048, true>(has_b, npool, grid, st, z, (int)M, gamma, beta, mean, rstd, dh_a, dh_b, dh_b_rows, t0, t1, n_heads, drop_p, seed, stream_id,
                                               (__nv_bfloat16*)dz_planes, plane_stride, nplanes, dgamma, dbeta, dbias, nullptr, nullptr);
        else launch_ln_bwd<2048, false>(has_b, npool, grid, st, z, (int)M, gamma, beta, mean, rstd, dh_a, dh_b, dh_b_rows, t0, t1, n_heads, drop_p, seed, stream_id,
                                        (__nv_bfloat16*)dz_planes, plane_stride, nplanes, dgamma, dbeta, dbias, nullptr, nullptr);
    }
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_gate_bwd(const void* gate_a, const void* gate_b, const float* dlogit, const float* wc, long long M, int n_heads,
                 float drop_p, unsigned long long seed, void* dpre_planes, long long plane_stride, int nplanes,
                 float* dba, float* dbb, float* dwc, float* dbc, void* stream) {
    MDL_REQUIRE(n_heads == 4, "gate_bwd: only n_heads == 4 is built (got %d)", n_heads);
    if (M == 0) return 0;
    (void)seed;      // the masks are read off the stored gates (a dropped gate is an exact zero), not regenerated
    const int groups = n_heads * 512 / GB_COLS;            // 64-column slices
    const long long n_rb = (M + 31) / 32;
    long long chunks = (long long)kNumSMs * 6 / groups;    // ~6 resident blocks per SM
    if (chunks > n_rb) chunks = n_rb;
    if (chunks < 1) chunks = 1;
    const int grid = (int)(chunks * groups);
    if (nplanes > 1)
        launch_k(gate_bwd_kernel<2>, dim3(grid), dim3(GB_THREADS), 0, (cudaStream_t)stream, (const __half*)gate_a, (const __half*)gate_b, dlogit, wc, M, n_heads, drop_p,
                                                                  (__nv_bfloat16*)dpre_planes, plane_stride, dba, dbb, dwc, dbc);
    else
        launch_k(gate_bwd_kernel<1>, dim3(grid), dim3(GB_THREADS), 0, (cudaStream_t)stream, (const __half*)gate_a, (const __half*)gate_b, dlogit, wc, M, n_heads, drop_p,
                                                                  (__nv_bfloat16*)dpre_planes, plane_stride, dba, dbb, dwc, dbc);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_gather_rows_planes(const void* planes, long long plane_stride_in, int nplanes, int C, const int* rows, long long n_sel,
                           void* out, long long plane_stride_out, void* stream) {
    nplanes &= 0xff;      // a byte copy: the plane format flag does not matter
    MDL_REQUIRE(C % 8 == 0, "gather_rows_planes: C must be a multiple of 8 (got %d)", C);
    if (n_sel == 0) return 0;
    launch_k(gather_rows_planes_kernel, dim3(grid_for(n_sel * nplanes, 8, 8)), dim3(256), 0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)planes, plane_stride_in, nplanes, C, rows, n_sel, (__nv_bfloat16*)out, plane_stride_out);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_bag_colsum_planes(const void* planes, long long plane_stride, int nplanes, int C, const int* cu_seqlens, int n_bags, float* out, void* stream) {
    if (n_bags == 0) return 0;
    dim3 grid((C + 127) / 128, n_bags);
    launch_k(bag_colsum_planes_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, (const __nv_bfloat16*)planes, plane_stride, nplanes, C, cu_seqlens, out);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_stain_rowbias(const float* emb, const int* code, const float* w1, int ldw, int d_in, int se_dim, int n_out, int R, float* rowbias, void* stream) {
    if (R == 0) return 0;
    launch_k(stain_rowbias_kernel, dim3(R), dim3(256), 0, (cudaStream_t)stream, emb, code, w1, ldw, d_in, se_dim, n_out, rowbias);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_stain_rowbias_bwd(const float* G, const float* emb, const int* code, const float* w1, int ldw, int d_in, int se_dim, int n_out, int R,
                          float* dw1, float* demb, void* stream) {
    if (R == 0) return 0;
    launch_k(stain_rowbias_bwd_kernel, dim3(n_out + R), dim3(32), 0, (cudaStream_t)stream, G, emb, code, w1, ldw, d_in, se_dim, n_out, R, dw1, demb);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_colsum_f32(const float* x, long long M, int C, float* out, void* stream) {
    if (M == 0) return 0;
    const long long want = (M + 63) / 64, cap = 4LL * kNumSMs / ((C + 127) / 128) + 1;   // ~4 blocks per SM in total
    dim3 grid((C + 127) / 128, (unsigned)(want > cap ? cap : want));
    launch_k(colsum_f32_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, x, M, C, out);
    MDL_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
