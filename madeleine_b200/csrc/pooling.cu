// Masked (bag-packed) attention pooling over variable-length bags — the HBM-bound kernel of the path.
//
//   forward : S[r, h, e] = sum_{t in bag r} softmax_t(logit[t, h]) * X[t, h, e]
//   backward: dlogit[t, h] = p[t, h] * (sum_e dS[r,h,e] X[t,h,e] - sum_e dS[r,h,e] S[r,h,e])
//             (the dX term p[t,h] * dS[r,h,e] is fused into the LayerNorm/GELU backward of the producing layer)
//
// X is the head-major pre-attention feature matrix [sum_N, H*E] stored as bf16 planes (hi[, lo]); bags are
// delimited by cu_seqlens[R+1]; an optional tok_idx list turns a segment into a gather (n_views = 3 half views).
// Algorithmic bytes per bag (forward): N*H*E*2*nplanes + N*H*4 + H*E*4.
//
// Work decomposition (forward): a tiny statistics kernel (one CTA per bag) turns the logits into attention weights
// p[t, h] once; the streaming kernel runs one CTA per (bag, head, 256-channel half, token chunk of ~192 tokens) so the
// 148 SMs see thousands of equal, prologue-free work items and the tail is short;
// token splits write partial sums to a workspace and the last CTA to finish (atomic ticket) adds them in split
// order, so the result is bit-reproducible run to run.  Each lane streams 16-byte vectors (8 bf16 channels),
// a warp covers 512 contiguous bytes per plane per token, 4 tokens are in flight per warp.
#include "common.cuh"
#include "madeleine_b200.h"

namespace mdl {

constexpr int POOL_THREADS = 256;
constexpr int POOL_WARPS = POOL_THREADS / 32;

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// fp16 planes (inference format): same accumulation, half2 -> float2 conversions
__device__ __forceinline__ void fma8_f16(float (&acc)[8], float w, const uint4& v) {
    const float2 a = f16x2_to_f32(v.x), b = f16x2_to_f32(v.y), c = f16x2_to_f32(v.z), d = f16x2_to_f32(v.w);
    acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]);
    acc[2] = fmaf(w, b.x, acc[2]); acc[3] = fmaf(w, b.y, acc[3]);
    acc[4] = fmaf(w, c.x, acc[4]); acc[5] = fmaf(w, c.y, acc[5]);
    acc[6] = fmaf(w, d.x, acc[6]); acc[7] = fmaf(w, d.y, acc[7]);
}
__device__ __forceinline__ void fma8(float (&acc)[8], float w, const uint4& v) {
    acc[0] = fmaf(w, bf16_lo_of(v.x), acc[0]); acc[1] = fmaf(w, bf16_hi_of(v.x), acc[1]);
    acc[2] = fmaf(w, bf16_lo_of(v.y), acc[2]); acc[3] = fmaf(w, bf16_hi_of(v.y), acc[3]);
    acc[4] = fmaf(w, bf16_lo_of(v.z), acc[4]); acc[5] = fmaf(w, bf16_hi_of(v.z), acc[5]);
    acc[6] = fmaf(w, bf16_lo_of(v.w), acc[6]); acc[7] = fmaf(w, bf16_hi_of(v.w), acc[7]);
}
__device__ __forceinline__ float dot8(const float (&d)[8], const uint4& v) {
    float s = d[0] * bf16_lo_of(v.x);
    s = fmaf(d[1], bf16_hi_of(v.x), s);
    s = fmaf(d[2], bf16_lo_of(v.y), s); s = fmaf(d[3], bf16_hi_of(v.y), s);
    s = fmaf(d[4], bf16_lo_of(v.z), s); s = fmaf(d[5], bf16_hi_of(v.z), s);
    s = fmaf(d[6], bf16_lo_of(v.w), s); s = fmaf(d[7], bf16_hi_of(v.w), s);
    return s;
}

enum { ACT_SOFTMAX = 0, ACT_LEAKY_RELU = 1, ACT_RELU = 2, ACT_SIGMOID = 3 };  // abmil.py:54-63

__device__ __forceinline__ float attn_weight(float l, float mx, float inv, int act) {
    switch (act) {
        case ACT_SOFTMAX: return __expf(l - mx) * inv;
        case ACT_LEAKY_RELU: return l > 0.f ? l : 0.01f * l;
        case ACT_RELU: return fmaxf(l, 0.f);
        default: return sigmoid_acc(l);
    }
}
__device__ __forceinline__ float attn_weight_grad(float l, float w, int act) {  // d act(l) / d l for the pointwise cases
    switch (act) {
        case ACT_LEAKY_RELU: return l > 0.f ? 1.f : 0.01f;
        case ACT_RELU: return l > 0.f ? 1.f : 0.f;
        default: return w * (1.f - w);
    }
}

// attention weights of one (bag, head): p[t, h] = act(logit) (softmax over the bag's tokens, or a pointwise activation).
__global__ void __launch_bounds__(POOL_THREADS)
pool_weights_kernel(const float* __restrict__ logits, const int* __restrict__ cu, const int* __restrict__ tok_idx, int H,
                    float* __restrict__ attn_p, int act, int* __restrict__ tickets, int halves) {
    pdl_sync();
    __shared__ float scratch[33];
    const int r = blockIdx.x;
    if (tickets != nullptr && threadIdx.x < halves) tickets[(r * H + blockIdx.y) * halves + threadIdx.x] = 0;
    const int t0 = cu[r], n = cu[r + 1] - t0;
    {
        const int h = blockIdx.y;
        float mx = 0.f, inv = 1.f;
        if (act == ACT_SOFTMAX) {
            mx = -INFINITY;
            for (int i = threadIdx.x; i < n; i += POOL_THREADS) {
                const long long row = tok_idx ? tok_idx[t0 + i] : (t0 + i);
                mx = fmaxf(mx, __ldg(logits + row * H + h));
            }
            mx = block_max(mx, scratch);
            float sm = 0.f;
            for (int i = threadIdx.x; i < n; i += POOL_THREADS) {
                const long long row = tok_idx ? tok_idx[t0 + i] : (t0 + i);
                sm += __expf(__ldg(logits + row * H + h) - mx);
            }
            sm = block_sum(sm, scratch);
            inv = n > 0 ? 1.f / sm : 0.f;
        }
        for (int i = threadIdx.x; i < n; i += POOL_THREADS) {
            const long long row = tok_idx ? tok_idx[t0 + i] : (t0 + i);
            attn_p[row * H + h] = attn_weight(__ldg(logits + row * H + h), mx, inv, act);
        }
    }
}

template <int NPLANES, bool F16>
__global__ void __launch_bounds__(POOL_THREADS, 4)
pool_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long plane_stride, const float* __restrict__ attn_p,
                const int* __restrict__ cu, const int* __restrict__ tok_idx, int H, int E, int tsplit,
                float* __restrict__ out, float* __restrict__ partial, int* __restrict__ tickets) {
    pdl_sync();
    __shared__ int is_last;
    __shared__ float part[POOL_WARPS][256];
    const int halves = E / 256;
    const int half = blockIdx.x % halves, split = blockIdx.x / halves;
    const int h = blockIdx.y, r = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = cu[r], n = cu[r + 1] - t0;
    const int C = H * E;
    const int per = (n + tsplit - 1) / tsplit;
    const int i_begin = split * per, i_end = min(n, i_begin + per);
    const int col = h * E + half * 256 + lane * 8;

    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;

    constexpr int UNROLL = 4;
    for (int i = i_begin + warp; i < i_end; i += POOL_WARPS * UNROLL) {
        uint4 vh[UNROLL], vl[UNROLL];
        float w[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int ii = i + u * POOL_WARPS;
            const bool ok = ii < i_end;
            const int pos = ok ? ii : i;
            const long long row = tok_idx ? tok_idx[t0 + pos] : (long long)(t0 + pos);
            vh[u] = ldg_stream(x + row * C + col);
            if (NPLANES > 1) vl[u] = ldg_stream(x + plane_stride + row * C + col);
            w[u] = ok ? __ldg(attn_p + row * H + h) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (F16) {
                fma8_f16(acc, w[u], vh[u]);
                if (NPLANES > 1) fma8_f16(acc, w[u], vl[u]);
            } else {
                fma8(acc, w[u], vh[u]);
                if (NPLANES > 1) fma8(acc, w[u], vl[u]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) part[warp][lane * 8 + i] = acc[i];
    __syncthreads();
    {
        const int c = threadIdx.x;  // 256 threads <-> 256 channels
        float s = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < POOL_WARPS; ++w2) s += part[w2][c];
        float* dst = out + (long long)r * C + h * E + half * 256 + c;
        if (tsplit == 1) {
            *dst = s;
        } else {
            const int R = gridDim.z;
            partial[((long long)split * R + r) * C + h * E + half * 256 + c] = s;
            __threadfence();
            __syncthreads();
            int* ticket = tickets + (r * H + h) * halves + half;
            if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1) == tsplit - 1;
            __syncthreads();
            if (is_last) {
                __threadfence();
                float t = 0.f;
                for (int sp = 0; sp < tsplit; ++sp) t += __ldcg(partial + ((long long)sp * R + r) * C + h * E + half * 256 + c);
                *dst = t;
                if (threadIdx.x == 0) *ticket = 0;  // ready for the next launch
            }
        }
    }
}

template <int NPLANES>
__global__ void __launch_bounds__(POOL_THREADS)
pool_bwd_dlogit_kernel(const __nv_bfloat16* __restrict__ x, long long plane_stride, const float* __restrict__ dS,
                       const float* __restrict__ S, const float* __restrict__ attn_p, const int* __restrict__ cu,
                       const int* __restrict__ tok_idx, int H, int E, int tsplit, float* __restrict__ dlogit, int accumulate,
                       const float* __restrict__ logits, int act) {
    pdl_sync();
    const int split = blockIdx.x, h = blockIdx.y, r = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = cu[r], n = cu[r + 1] - t0;
    const int C = H * E;
    const int chunks = E / 256;  // 256-channel chunks per head (E = 512 -> 2)
    // dS slab of this (bag, head) in registers: lane owns channels chunk*256 + lane*8 .. +8
    float d[2][8];
    float cdot = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (k < chunks) {
            const float* ds = dS + (long long)r * C + h * E + k * 256 + lane * 8;
            const float* ss = S + (long long)r * C + h * E + k * 256 + lane * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) { d[k][i] = __ldg(ds + i); cdot = fmaf(d[k][i], __ldg(ss + i), cdot); }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) d[k][i] = 0.f;
        }
    }
    cdot = warp_sum(cdot);

    const int per = (n + tsplit - 1) / tsplit;
    const int i_begin = split * per, i_end = min(n, i_begin + per);
    constexpr int UNROLL = 2;
    for (int i = i_begin + warp; i < i_end; i += POOL_WARPS * UNROLL) {
        uint4 v[UNROLL][2][2];
        long long rows[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int ii = i + u * POOL_WARPS;
            const int pos = ii < i_end ? ii : i;
            rows[u] = tok_idx ? tok_idx[t0 + pos] : (long long)(t0 + pos);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (k < chunks) {
                    const __nv_bfloat16* px = x + rows[u] * C + h * E + k * 256 + lane * 8;
                    v[u][k][0] = ldg_stream(px);
                    if (NPLANES > 1) v[u][k][1] = ldg_stream(px + plane_stride);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            float g = 0.f;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (k < chunks) {
                    g += dot8(d[k], v[u][k][0]);
                    if (NPLANES > 1) g += dot8(d[k], v[u][k][1]);
                }
            }
            g = warp_sum(g);
            if (lane == 0 && (i + u * POOL_WARPS) < i_end) {
                const long long o = rows[u] * H + h;
                const float pw = __ldg(attn_p + o);
                const float val = act == ACT_SOFTMAX ? pw * (g - cdot) : attn_weight_grad(__ldg(logits + o), pw, act) * g;
                dlogit[o] = accumulate ? dlogit[o] + val : val;
            }
        }
    }
}

// Head-major [M, H*E] planes -> reference channel order fp32 [M, E, H] (c_ref = e*H + h); only for
// ABMILEmbedder(return_preattn_feats=True) called on its own.
__global__ void planes_to_ref_order_kernel(const __nv_bfloat16* __restrict__ x, long long plane_stride, int nplanes, int f16,
                                           long long M, int H, int E, float* __restrict__ out) {
    pdl_sync();
    const long long total = M * H * E;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / (H * E);
        const int cr = (int)(i - m * H * E);
        const int e = cr / H, h = cr % H;
        const long long src = m * H * E + h * E + e;
        float v;
        if (f16) {
            const __half* xh = reinterpret_cast<const __half*>(x);
            v = __half2float(xh[src]);
            if (nplanes > 1) v += __half2float(xh[plane_stride + src]);
        } else {
            v = __bfloat162float(x[src]);
            if (nplanes > 1) v += __bfloat162float(x[plane_stride + src]);
        }
        out[i] = v;
    }
}

static int choose_tsplit(int R, int H, int halves, long long total_tokens) {
    (void)H; (void)halves;
    if (R <= 0) return 1;
    const long long avg = total_tokens / R;
    long long s = (avg + 191) / 192;          // ~192 tokens per CTA: thousands of equal work items, short tail
    if (s < 1) s = 1;
    if (s > 64) s = 64;
    return (int)s;
}

}  // namespace mdl

using namespace mdl;

extern "C" {

static int* pool_tickets(void* workspace, int tsplit, int n_bags, int n_heads, int head_dim) {
    if (tsplit <= 1 || workspace == nullptr) return nullptr;
    return reinterpret_cast<int*>(reinterpret_cast<float*>(workspace) + (size_t)tsplit * n_bags * n_heads * head_dim);
}

int mdl_pool_weights(const float* logits, const int* cu_seqlens, const int* tok_idx, int n_bags, int n_heads, int head_dim,
                     float* attn_p, int activation, int tsplit, void* workspace, void* stream) {
    MDL_REQUIRE(activation >= 0 && activation <= 3, "pool_weights: unknown activation %d", activation);
    MDL_REQUIRE(attn_p != nullptr, "pool_weights: attn_p ([tokens, n_heads] fp32) is required");
    MDL_REQUIRE(tsplit >= 1 && tsplit <= 64, "pool_weights: tsplit must be in [1, 64]");
    MDL_REQUIRE(tsplit == 1 || workspace != nullptr, "pool_weights: tsplit > 1 needs the pooling workspace (its tickets are cleared here)");
    if (n_bags == 0) return 0;
    const int halves = head_dim / 256 > 0 ? head_dim / 256 : 1;
    launch_k(pool_weights_kernel, dim3(dim3(n_bags, n_heads)), dim3(POOL_THREADS), 0, (cudaStream_t)stream, 
        logits, cu_seqlens, tok_idx, n_heads, attn_p, activation, pool_tickets(workspace, tsplit, n_bags, n_heads, head_dim), halves);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_pool_fwd(const void* x_planes, long long plane_stride, int nplanes, const float* attn_p, const int* cu_seqlens,
                 const int* tok_idx, int n_bags, long long total_tokens, int n_heads, int head_dim,
                 float* out, int tsplit, void* workspace, void* stream) {
    const bool f16 = (nplanes & kPlanesF16) != 0;       // fp16 hi/lo planes (the fp32-grade inference format)
    nplanes &= 0xff;
    MDL_REQUIRE(head_dim % 256 == 0, "pool_fwd: head_dim must be a multiple of 256 (got %d)", head_dim);
    MDL_REQUIRE(nplanes == 1 || nplanes == 2, "pool_fwd: nplanes must be 1 or 2");
    MDL_REQUIRE(attn_p != nullptr, "pool_fwd: attn_p (from mdl_pool_weights) is required");
    (void)total_tokens;
    if (n_bags == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int halves = head_dim / 256;
    MDL_REQUIRE(tsplit >= 1 && tsplit <= 64, "pool_fwd: tsplit must be in [1, 64] (use mdl_pool_tsplit), got %d", tsplit);
    MDL_REQUIRE(tsplit == 1 || workspace != nullptr, "pool_fwd: tsplit > 1 needs a workspace (mdl_pool_workspace_bytes)");
    float* partial = reinterpret_cast<float*>(workspace);
    int* tickets = pool_tickets(workspace, tsplit, n_bags, n_heads, head_dim);
    dim3 grid(halves * tsplit, n_heads, n_bags);
    if (nplanes == 2 && f16)
        launch_k(pool_fwd_kernel<2, true>, dim3(grid), dim3(POOL_THREADS), 0, st, (const __nv_bfloat16*)x_planes, plane_stride, attn_p, cu_seqlens, tok_idx, n_heads, head_dim, tsplit, out, partial, tickets);
    else if (nplanes == 2)
        launch_k(pool_fwd_kernel<2, false>, dim3(grid), dim3(POOL_THREADS), 0, st, (const __nv_bfloat16*)x_planes, plane_stride, attn_p, cu_seqlens, tok_idx, n_heads, head_dim, tsplit, out, partial, tickets);
    else if (f16)
        launch_k(pool_fwd_kernel<1, true>, dim3(grid), dim3(POOL_THREADS), 0, st, (const __nv_bfloat16*)x_planes, plane_stride, attn_p, cu_seqlens, tok_idx, n_heads, head_dim, tsplit, out, partial, tickets);
    else
        launch_k(pool_fwd_kernel<1, false>, dim3(grid), dim3(POOL_THREADS), 0, st, (const __nv_bfloat16*)x_planes, plane_stride, attn_p, cu_seqlens, tok_idx, n_heads, head_dim, tsplit, out, partial, tickets);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_pool_tsplit(int n_bags, long long total_tokens, int n_heads, int head_dim) {
    return choose_tsplit(n_bags, n_heads, head_dim / 256 > 0 ? head_dim / 256 : 1, total_tokens);
}

long long mdl_pool_workspace_bytes(int n_bags, int n_heads, int head_dim, int tsplit) {
    if (tsplit <= 1) return 0;
    const int halves = head_dim / 256 > 0 ? head_dim / 256 : 1;
    return (long long)sizeof(float) * tsplit * n_bags * n_heads * head_dim + (long long)sizeof(int) * n_bags * n_heads * halves;
}

int mdl_pool_bwd_dlogit(const void* x_planes, long long plane_stride, int nplanes, const float* dS, const float* S, const float* attn_p,
                        const int* cu_seqlens, const int* tok_idx, int n_bags, long long total_tokens, int n_heads, int head_dim,
                        float* dlogit, int accumulate, const float* logits, int activation, int tsplit, void* stream) {
    MDL_REQUIRE(activation >= 0 && activation <= 3, "pool_bwd: unknown activation %d", activation);
    MDL_REQUIRE(head_dim == 256 || head_dim == 512, "pool_bwd: head_dim must be 256 or 512 (got %d)", head_dim);
    MDL_REQUIRE(nplanes == 1 || nplanes == 2, "pool_bwd: nplanes must be 1 or 2");
    if (n_bags == 0) return 0;
    if (tsplit <= 0) tsplit = choose_tsplit(n_bags, n_heads, 1, total_tokens);
    dim3 grid(tsplit, n_heads, n_bags);
    if (nplanes == 2)
        launch_k(pool_bwd_dlogit_kernel<2>, dim3(grid), dim3(POOL_THREADS), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x_planes, plane_stride, dS, S, attn_p, cu_seqlens, tok_idx, n_heads, head_dim, tsplit, dlogit, accumulate, logits, activation);
    else
        launch_k(pool_bwd_dlogit_kernel<1>, dim3(grid), dim3(POOL_THREADS), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x_planes, plane_stride, dS, S, attn_p, cu_seqlens, tok_idx, n_heads, head_dim, tsplit, dlogit, accumulate, logits, activation);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_planes_to_ref_order(const void* x_planes, long long plane_stride, int nplanes, long long M, int n_heads, int head_dim, float* out, void* stream) {
    if (M == 0) return 0;
    const int f16 = (nplanes & kPlanesF16) != 0;
    nplanes &= 0xff;
    const long long total = M * n_heads * head_dim;
    long long blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    launch_k(planes_to_ref_order_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x_planes, plane_stride, nplanes, f16, M, n_heads, head_dim, out);
    MDL_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
