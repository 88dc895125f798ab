// Graph Optimal Transport local loss for problems of 96 < n <= 256 tokens (reference: madeleine/utils/loss.py:162-301;
// GOT(..., subsample=256) caps n at min(#cases with the stain, 256), trainer.py:45 + loss.py:281-284).
//
// Same mathematics and the same hand-written reverse sweep as got.cu (see the header comment there); what changes is
// where the n x n matrices live.  Five 256 x 257 fp32 matrices (1.3 MB) do not fit one SM's shared memory, so a problem's
// matrices stay in its global workspace slab (L2-resident: one CTA streams ~1.5 MB per IPOT iteration) and the kernels
// are restructured so that every global access is a coalesced ROW sweep:
//   * column sums (sigma, dsigma, ...) are accumulated while sweeping rows — lane l keeps partial sums of columns
//     l, l+32, ... in registers — and combined across the 32 warps through a [32][n] shared-memory staging buffer;
//   * the n x n x n products stage 32-row tiles of B (transposed on the fly for A B^T) in shared memory and give each
//     warp two output rows per pass, so B is re-read from L2 n/64 times instead of n times.
// One CTA of 1024 threads per (case, stain) problem; shared memory holds only the O(n) vectors and the staging tile.
#include "got_common.cuh"
#include "madeleine_b200.h"

namespace mdl {

constexpr int BIG_THREADS = 1024;
constexpr int BIG_WARPS = BIG_THREADS / 32;
constexpr int BIG_NU = GOT_BIG_NMAX / 32;     // column chunks per lane

struct BigVecs {
    float *sigma, *signew, *sigprev, *delta, *dsig, *dsig_prev, *ddel, *dc, *dr;
    float *rowsq_s, *rowsq_t, *dcst_r, *dcst_c;
    float *lu, *lw;      // [(MAX_ITERS+1)][n]
    float* stage;        // max(32 x n column-partial staging, 32 x (n|1) B tile)
};

// Sum the per-warp column partials (lane l of every warp holds columns l + 32u) and hand column j's total to fn(j, total).
template <class F>
__device__ __forceinline__ void col_reduce(const float (&acc)[BIG_NU], float* stage, int n, F&& fn) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int u = 0; u < BIG_NU; ++u) {
        const int j = lane + 32 * u;
        if (j < n) stage[warp * n + j] = acc[u];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += BIG_THREADS) {
        float s = 0.f;
#pragma unroll 8
        for (int w = 0; w < BIG_WARPS; ++w) s += stage[w * n + j];
        fn(j, s);
    }
    __syncthreads();
}

// C(i,j) = alpha * sum_k A(i,k) B(k,j);  A(i,k) = TA ? A[k][i] : A[i][k];  B(k,j) = TB ? B[j][k] : B[k][j].
// Output to C (global, leading dimension ld, must not alias A/B) and/or accumulated into the dense matrix Cg (ld = n).
template <bool TA, bool TB>
__device__ void mm_big(const float* A, const float* B, int n, int ld, float alpha, float* C, float* Cg, float* tile) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lds = n | 1;
    for (int rb = 0; rb < n; rb += 2 * BIG_WARPS) {
        const int i0 = rb + 2 * warp, i1 = i0 + 1;
        float acc0[BIG_NU], acc1[BIG_NU];
#pragma unroll
        for (int u = 0; u < BIG_NU; ++u) { acc0[u] = 0.f; acc1[u] = 0.f; }
        for (int k0 = 0; k0 < n; k0 += 32) {
            __syncthreads();                                   // the previous tile has been consumed
            if (!TB) {
                for (int idx = threadIdx.x; idx < 32 * n; idx += BIG_THREADS) {
                    const int kk = idx / n, j = idx - kk * n, k = k0 + kk;
                    tile[kk * lds + j] = k < n ? B[(size_t)k * ld + j] : 0.f;
                }
            } else {
                const int k = k0 + lane;
                for (int j = warp; j < n; j += BIG_WARPS) tile[lane * lds + j] = k < n ? B[(size_t)j * ld + k] : 0.f;
            }
            __syncthreads();
            // this lane's slice of the two A rows for the tile, broadcast by shuffle inside the k loop
            const int k = k0 + lane;
            float a0 = 0.f, a1 = 0.f;
            if (k < n) {
                if (i0 < n) a0 = TA ? A[(size_t)k * ld + i0] : A[(size_t)i0 * ld + k];
                if (i1 < n) a1 = TA ? A[(size_t)k * ld + i1] : A[(size_t)i1 * ld + k];
            }
#pragma unroll 4
            for (int kk = 0; kk < 32; ++kk) {
                const float x0 = __shfl_sync(0xffffffffu, a0, kk), x1 = __shfl_sync(0xffffffffu, a1, kk);
#pragma unroll
                for (int u = 0; u < BIG_NU; ++u) {
                    const int j = lane + 32 * u;
                    if (j < n) {
                        const float b = tile[kk * lds + j];
                        acc0[u] = fmaf(x0, b, acc0[u]);
                        acc1[u] = fmaf(x1, b, acc1[u]);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < BIG_NU; ++u) {
            const int j = lane + 32 * u;
            if (j < n) {
                if (i0 < n) {
                    if (C != nullptr) C[(size_t)i0 * ld + j] = alpha * acc0[u];
                    if (Cg != nullptr) Cg[(size_t)i0 * n + j] += alpha * acc0[u];
                }
                if (i1 < n) {
                    if (C != nullptr) C[(size_t)i1 * ld + j] = alpha * acc1[u];
                    if (Cg != nullptr) Cg[(size_t)i1 * n + j] += alpha * acc1[u];
                }
            }
        }
    }
    __syncthreads();
}

// IPOT forward (loss.py:179-193) on L = -C/beta; T <- plan T_K, lu/lw[t] <- cumulative log scalings, A scratch = exp(L).
__device__ void ipot_forward_big(const float* L, float* T, float* A, int K, int n, int ld, const BigVecs& v) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float inv_n = 1.f / n;
    for (int idx = tid; idx < n * ld; idx += BIG_THREADS) { T[idx] = 1.f; A[idx] = expf(L[idx]); }
    for (int j = tid; j < n; j += BIG_THREADS) { v.sigma[j] = inv_n; v.lu[j] = 0.f; v.lw[j] = 0.f; }
    __syncthreads();
    for (int t = 1; t <= K; ++t) {
        // Q = A * T (stored in T), delta = 1 / (n Q sigma); column partials of Q^T delta in the same sweep
        float cacc[BIG_NU];
#pragma unroll
        for (int u = 0; u < BIG_NU; ++u) cacc[u] = 0.f;
        for (int i = warp; i < n; i += BIG_WARPS) {
            float q[BIG_NU];
            float rs = 0.f;
#pragma unroll
            for (int u = 0; u < BIG_NU; ++u) {
                const int j = lane + 32 * u;
                q[u] = 0.f;
                if (j < n) {
                    q[u] = A[(size_t)i * ld + j] * T[(size_t)i * ld + j];
                    rs = fmaf(q[u], v.sigma[j], rs);
                }
            }
            rs = warp_sum(rs);
            const float d = 1.f / (n * rs);
            if (lane == 0) v.delta[i] = d;
#pragma unroll
            for (int u = 0; u < BIG_NU; ++u) {
                const int j = lane + 32 * u;
                if (j < n) { T[(size_t)i * ld + j] = q[u]; cacc[u] = fmaf(q[u], d, cacc[u]); }
            }
        }
        col_reduce(cacc, v.stage, n, [&](int j, float s) { v.signew[j] = 1.f / (n * s); });
        for (int idx = tid; idx < n * n; idx += BIG_THREADS) {     // T = delta * Q * sigma^T
            const int i = idx / n, j = idx - i * n;
            T[(size_t)i * ld + j] *= v.delta[i] * v.signew[j];
        }
        for (int j = tid; j < n; j += BIG_THREADS) {
            v.lu[t * n + j] = v.lu[(t - 1) * n + j] + logf(v.delta[j]);
            v.lw[t * n + j] = v.lw[(t - 1) * n + j] + logf(v.signew[j]);
            v.sigma[j] = v.signew[j];
        }
        __syncthreads();
    }
}

// Reverse sweep of ipot_forward_big.  dT (in: dLoss/dT_K, destroyed), dL (accumulated), Q scratch.
__device__ void ipot_backward_big(const float* L, float* dT, float* dL, float* Q, int K, int n, int ld, const BigVecs& v) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float inv_n = 1.f / n;
    for (int j = tid; j < n; j += BIG_THREADS) v.dsig[j] = 0.f;
    __syncthreads();
    for (int t = K; t >= 1; --t) {
        for (int j = tid; j < n; j += BIG_THREADS) {
            v.delta[j] = expf(v.lu[t * n + j] - v.lu[(t - 1) * n + j]);
            v.sigma[j] = expf(v.lw[t * n + j] - v.lw[(t - 1) * n + j]);
            v.sigprev[j] = t == 1 ? inv_n : expf(v.lw[(t - 1) * n + j] - v.lw[(t - 2) * n + j]);
        }
        __syncthreads();
        float cacc[BIG_NU];
        // sweep 1: Q_t = exp(t L + lu_{t-1} + lw_{t-1});  ddelta[i] = sum_j dT Q sigma_t[j];  column partials of dT Q delta_t
#pragma unroll
        for (int u = 0; u < BIG_NU; ++u) cacc[u] = 0.f;
        for (int i = warp; i < n; i += BIG_WARPS) {
            const float lui = v.lu[(t - 1) * n + i], di = v.delta[i];
            float acc = 0.f;
#pragma unroll
            for (int u = 0; u < BIG_NU; ++u) {
                const int j = lane + 32 * u;
                if (j < n) {
                    const size_t o = (size_t)i * ld + j;
                    const float q = expf(t * L[o] + lui + v.lw[(t - 1) * n + j]);
                    Q[o] = q;
                    const float x = dT[o] * q;
                    acc = fmaf(x, v.sigma[j], acc);
                    cacc[u] = fmaf(x, di, cacc[u]);
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) v.ddel[i] = acc;
        }
        col_reduce(cacc, v.stage, n, [&](int j, float s) {
            const float ds = v.dsig[j] + s;
            v.dc[j] = -ds * n * v.sigma[j] * v.sigma[j];
        });
        // sweep 2: ddelta[i] += sum_j Q dc[j];  dr[i] = -ddelta[i] n delta_t[i]^2;  column partials of Q dr
#pragma unroll
        for (int u = 0; u < BIG_NU; ++u) cacc[u] = 0.f;
        for (int i = warp; i < n; i += BIG_WARPS) {
            float q[BIG_NU];
            float acc = 0.f;
#pragma unroll
            for (int u = 0; u < BIG_NU; ++u) {
                const int j = lane + 32 * u;
                q[u] = 0.f;
                if (j < n) { q[u] = Q[(size_t)i * ld + j]; acc = fmaf(q[u], v.dc[j], acc); }
            }
            acc = warp_sum(acc);
            const float dd = v.ddel[i] + acc;
            const float dri = -dd * n * v.delta[i] * v.delta[i];
            if (lane == 0) v.dr[i] = dri;
#pragma unroll
            for (int u = 0; u < BIG_NU; ++u) cacc[u] = fmaf(q[u], dri, cacc[u]);
        }
        col_reduce(cacc, v.stage, n, [&](int j, float s) { v.dsig_prev[j] = s; });
        // dQ = dT delta sigma^T + delta dc^T + dr sigma_{t-1}^T;  dL += dQ Q;  dT_{t-1} = dQ A
        for (int idx = tid; idx < n * n; idx += BIG_THREADS) {
            const int i = idx / n, j = idx - i * n;
            const size_t o = (size_t)i * ld + j;
            const float dq = dT[o] * v.delta[i] * v.sigma[j] + v.delta[i] * v.dc[j] + v.dr[i] * v.sigprev[j];
            dL[o] = fmaf(dq, Q[o], dL[o]);
            dT[o] = dq * expf(L[o]);
        }
        __syncthreads();
        for (int j = tid; j < n; j += BIG_THREADS) v.dsig[j] = v.dsig_prev[j];
        __syncthreads();
    }
}

__device__ __forceinline__ void load_mat_big(float* S, const float* g, int n, int ld, float scale, float sub, bool relu) {
    for (int idx = threadIdx.x; idx < n * n; idx += BIG_THREADS) {
        const int i = idx / n, j = idx - i * n;
        float x = g[idx] - sub;
        if (relu) x = fmaxf(x, 0.f);
        S[(size_t)i * ld + j] = x * scale;
    }
}
__device__ __forceinline__ void store_mat_big(const float* S, float* g, int n, int ld) {
    for (int idx = threadIdx.x; idx < n * n; idx += BIG_THREADS) {
        const int i = idx / n, j = idx - i * n;
        g[idx] = S[(size_t)i * ld + j];
    }
}

// ---------------------------------------------------------------------------------------------------
// kernel A: normalised tokens, raw cosine costs of one problem + its extrema
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BIG_THREADS)
got_cost_big_kernel(const float* __restrict__ v, const float* __restrict__ q, GotLayout lay, float* ws) {
    extern __shared__ float sm[];
    const int n = lay.n, D = lay.D, ldd = D + 1;
    float* Ys = sm;                        // [n][D+1]
    __shared__ float red_val[BIG_WARPS];
    __shared__ int red_idx[BIG_WARPS];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* slab = ws + lay.header + lay.per_item * (size_t)b;
    float* vn = slab + lay.vn;
    float* qn = slab + lay.qn;
    float* norms = slab + lay.norms;
    for (int r = warp; r < 2 * n; r += BIG_WARPS) {
        const bool isv = r < n;
        const int i = isv ? r : r - n;
        const float* src = (isv ? v : q) + ((size_t)b * n + i) * D;
        float s = 0.f;
        for (int d = lane; d < D; d += 32) { const float x = src[d]; s = fmaf(x, x, s); }
        s = warp_sum(s);
        const float nrm = sqrtf(s), inv = 1.f / (nrm + 1e-12f);
        float* dst = (isv ? vn : qn) + (size_t)i * D;
        for (int d = lane; d < D; d += 32) dst[d] = src[d] * inv;
        if (lane == 0) norms[r] = nrm;
    }
    __syncthreads();
    for (int which = 0; which < 3; ++which) {
        const float* X = which == 2 ? qn : vn;
        const float* Y = which == 1 ? vn : qn;
        __syncthreads();
        for (int idx = tid; idx < n * D; idx += BIG_THREADS) {
            const int j = idx / D, d = idx - j * D;
            Ys[j * ldd + d] = Y[idx];
        }
        __syncthreads();
        float* out = slab + (which == 0 ? lay.raw0 : (which == 1 ? lay.raws : lay.rawt));
        float mn = INFINITY, mx = -INFINITY;
        int imn = 0x7fffffff, imx = 0x7fffffff;
        for (int idx = tid; idx < n * n; idx += BIG_THREADS) {
            const int i = idx / n, j = idx - i * n;
            const float* xr = X + (size_t)i * D;
            float s = 0.f;
            for (int d = 0; d < D; ++d) s = fmaf(xr[d], Ys[j * ldd + d], s);
            const float c = 1.f - s;
            out[idx] = c;
            if (c < mn) { mn = c; imn = idx; }
            if (c > mx) { mx = c; imx = idx; }
        }
        for (int pass = 0; pass < 2; ++pass) {             // block arg-min / arg-max (ties -> smallest index)
            float val = pass == 0 ? mn : -mx;
            int id = pass == 0 ? imn : imx;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, val, o);
                const int oi = __shfl_xor_sync(0xffffffffu, id, o);
                if (ov < val || (ov == val && oi < id)) { val = ov; id = oi; }
            }
            __syncthreads();
            if (lane == 0) { red_val[warp] = val; red_idx[warp] = id; }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < BIG_WARPS; ++w)
                    if (red_val[w] < val || (red_val[w] == val && red_idx[w] < id)) { val = red_val[w]; id = red_idx[w]; }
                float* e = slab + lay.ext;
                e[which * 2 + pass] = pass == 0 ? val : -val;
                reinterpret_cast<int*>(e)[6 + which * 2 + pass] = id;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// kernel C: forward + reverse sweep of one problem, down to gradients w.r.t. the thresholded costs
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BIG_THREADS, 1)
got_main_big_kernel(GotLayout lay, float* ws, const float* __restrict__ extrema, float* __restrict__ wd_out, float* __restrict__ gwd_out) {
    extern __shared__ float sm[];
    const int n = lay.n, ld = n | 1;
    const size_t msz = (size_t)n * ld;
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* slab = ws + lay.header + lay.per_item * (size_t)b;
    float* S0 = slab + lay.mats; float* S1 = S0 + msz; float* S2 = S1 + msz; float* S3 = S2 + msz; float* S4 = S3 + msz;
    float* vp = sm;
    BigVecs v;
    v.sigma = vp; vp += n; v.signew = vp; vp += n; v.sigprev = vp; vp += n; v.delta = vp; vp += n; v.dsig = vp; vp += n;
    v.dsig_prev = vp; vp += n; v.ddel = vp; vp += n; v.dc = vp; vp += n; v.dr = vp; vp += n;
    v.rowsq_s = vp; vp += n; v.rowsq_t = vp; vp += n; v.dcst_r = vp; vp += n; v.dcst_c = vp; vp += n;
    v.lu = vp; vp += (MAX_ITERS + 1) * n; v.lw = vp; vp += (MAX_ITERS + 1) * n;
    v.stage = vp;
    __shared__ float scratch[33];

    const float thr0 = extrema[0] + THR_BETA * (extrema[1] - extrema[0]);
    const float thrs = extrema[2] + THR_BETA * (extrema[3] - extrema[2]);
    const float thrt = extrema[4] + THR_BETA * (extrema[5] - extrema[4]);
    const float inv_n = 1.f / n;

    // ================= Wasserstein term =================
    load_mat_big(S1, slab + lay.raw0, n, ld, -1.f / WD_BETA, thr0, true);       // S1 = L = -C/beta, S0 = T
    __syncthreads();
    ipot_forward_big(S1, S0, S2, WD_ITERS, n, ld, v);
    {   // wd = <C, T>;  dT_K = C (into S2);  direct dC = T kept in S0;  dL accumulator S3 = 0
        float acc = 0.f;
        for (int idx = tid; idx < n * n; idx += BIG_THREADS) {
            const int i = idx / n, j = idx - i * n;
            const size_t o = (size_t)i * ld + j;
            const float c = -WD_BETA * S1[o];
            acc = fmaf(c, S0[o], acc);
            S2[o] = c;
            S3[o] = 0.f;
        }
        acc = block_sum(acc, scratch);
        if (tid == 0) wd_out[b] = acc;
    }
    __syncthreads();
    ipot_backward_big(S1, S2, S3, S4, WD_ITERS, n, ld, v);
    {   // dC = T_K - dL/beta, masked by C > 0 -> g0 ; dthr0 partial = -sum
        float acc = 0.f;
        float* g0 = slab + lay.g0;
        for (int idx = tid; idx < n * n; idx += BIG_THREADS) {
            const int i = idx / n, j = idx - i * n;
            const size_t o = (size_t)i * ld + j;
            const bool on = S1[o] < 0.f;  // C > 0  <=>  L < 0
            const float g = on ? S0[o] - S3[o] * (1.f / WD_BETA) : 0.f;
            g0[idx] = g;
            acc += g;
        }
        acc = block_sum(acc, scratch);
        if (tid == 0) slab[lay.ext + 12] = -acc;
    }
    __syncthreads();

    // ================= Gromov-Wasserstein term =================
    float* gs = slab + lay.gs;
    float* gt = slab + lay.gt;
    for (int idx = tid; idx < n * n; idx += BIG_THREADS) { gs[idx] = 0.f; gt[idx] = 0.f; }
    load_mat_big(S1, slab + lay.raws, n, ld, 1.f, thrs, true);                  // S1 = Cs, S3 = Ct stay during the forward
    load_mat_big(S3, slab + lay.rawt, n, ld, 1.f, thrt, true);
    __syncthreads();
    for (int i = warp; i < n; i += BIG_WARPS) {
        float a = 0.f, c = 0.f;
        for (int j = lane; j < n; j += 32) {
            const float x = S1[(size_t)i * ld + j], y = S3[(size_t)i * ld + j];
            a = fmaf(x, x, a); c = fmaf(y, y, c);
        }
        a = warp_sum(a); c = warp_sum(c);
        if (lane == 0) { v.rowsq_s[i] = a * inv_n; v.rowsq_t[i] = c * inv_n; }
    }
    for (int idx = tid; idx < n * ld; idx += BIG_THREADS) S0[idx] = inv_n * inv_n;   // gamma_0
    __syncthreads();
    for (int k = 0; k <= GW_OUTER; ++k) {
        // Cg_k = Cst - 2 Cs gamma_k Ct^T : P = Cs gamma (S2), Cg (S4)
        mm_big<false, false>(S1, S0, n, ld, 1.f, S2, nullptr, v.stage);
        mm_big<false, true>(S2, S3, n, ld, -2.f, S4, nullptr, v.stage);
        for (int idx = tid; idx < n * n; idx += BIG_THREADS) {
            const int i = idx / n, j = idx - i * n;
            S4[(size_t)i * ld + j] += v.rowsq_s[i] + v.rowsq_t[j];
        }
        __syncthreads();
        store_mat_big(S4, slab + lay.cg + (size_t)k * lay.nn, n, ld);
        if (k == GW_OUTER) break;
        for (int idx = tid; idx < n * n; idx += BIG_THREADS) {      // gamma_{k+1} = IPOT(Cg_k): L in S4, plan in S0
            const int i = idx / n, j = idx - i * n;
            S4[(size_t)i * ld + j] *= -1.f / GW_BETA;
        }
        __syncthreads();
        ipot_forward_big(S4, S0, S2, GW_INNER, n, ld, v);
        store_mat_big(S0, slab + lay.gamma + (size_t)k * lay.nn, n, ld);
        float* lulw = slab + lay.lulw + (size_t)k * 2 * (GW_INNER + 1) * n;
        for (int idx = tid; idx < (GW_INNER + 1) * n; idx += BIG_THREADS) {
            lulw[idx] = v.lu[idx];
            lulw[(GW_INNER + 1) * n + idx] = v.lw[idx];
        }
        __syncthreads();
    }
    {   // gwd = <Cg_5, gamma_5>   (S4 = Cg_5, S0 = gamma_5)
        float acc = 0.f;
        for (int idx = tid; idx < n * n; idx += BIG_THREADS) {
            const int i = idx / n, j = idx - i * n;
            const size_t o = (size_t)i * ld + j;
            acc = fmaf(S4[o], S0[o], acc);
        }
        acc = block_sum(acc, scratch);
        if (tid == 0) gwd_out[b] = acc;
    }
    for (int j = tid; j < n; j += BIG_THREADS) { v.dcst_r[j] = 0.f; v.dcst_c[j] = 0.f; }
    __syncthreads();

    // ---- reverse sweep.  Invariant at the top of each step k: S0 = dCg_k, S1 = Cs, S3 = Ct. ----
    for (int k = GW_OUTER; k >= 0; --k) {
        {   // row and column sums of dCg -> dCst (one row sweep)
            float cacc[BIG_NU];
#pragma unroll
            for (int u = 0; u < BIG_NU; ++u) cacc[u] = 0.f;
            for (int i = warp; i < n; i += BIG_WARPS) {
                float a = 0.f;
#pragma unroll
                for (int u = 0; u < BIG_NU; ++u) {
                    const int j = lane + 32 * u;
                    if (j < n) { const float x = S0[(size_t)i * ld + j]; a += x; cacc[u] += x; }
                }
                a = warp_sum(a);
                if (lane == 0) v.dcst_r[i] += a;
            }
            col_reduce(cacc, v.stage, n, [&](int j, float s) { v.dcst_c[j] += s; });
        }
        if (k == 0) {                                             // gamma_k -> S2
            for (int idx = tid; idx < n * ld; idx += BIG_THREADS) S2[idx] = inv_n * inv_n;
        } else {
            load_mat_big(S2, slab + lay.gamma + (size_t)(k - 1) * lay.nn, n, ld, 1.f, 0.f, false);
        }
        __syncthreads();
        mm_big<false, false>(S1, S2, n, ld, 1.f, S4, nullptr, v.stage);        // P = Cs gamma_k -> S4
        mm_big<true, false>(S0, S4, n, ld, -2.f, nullptr, gt, v.stage);        // dCt += -2 dCg^T P
        mm_big<false, false>(S0, S3, n, ld, -2.f, S4, nullptr, v.stage);       // dP = -2 dCg Ct -> S4
        mm_big<false, true>(S4, S2, n, ld, 1.f, nullptr, gs, v.stage);         // dCs += dP gamma_k^T
        if (k == 0) break;               // gamma_0 is a constant
        mm_big<true, false>(S1, S4, n, ld, 1.f, S0, nullptr, v.stage);         // dgamma_k = Cs^T dP -> S0
        // through gamma_k = IPOT(Cg_{k-1}):  L -> S2, dL -> S4 (zeroed), Q scratch = S3 (Ct is reloaded afterwards)
        load_mat_big(S2, slab + lay.cg + (size_t)(k - 1) * lay.nn, n, ld, -1.f / GW_BETA, 0.f, false);
        for (int idx = tid; idx < n * ld; idx += BIG_THREADS) S4[idx] = 0.f;
        const float* lulw = slab + lay.lulw + (size_t)(k - 1) * 2 * (GW_INNER + 1) * n;
        for (int idx = tid; idx < (GW_INNER + 1) * n; idx += BIG_THREADS) {
            v.lu[idx] = lulw[idx];
            v.lw[idx] = lulw[(GW_INNER + 1) * n + idx];
        }
        __syncthreads();
        ipot_backward_big(S2, S0, S4, S3, GW_INNER, n, ld, v);
        for (int idx = tid; idx < n * n; idx += BIG_THREADS) {     // dCg_{k-1} = -dL / beta -> S0 ; restore Ct in S3
            const int i = idx / n, j = idx - i * n;
            const size_t o = (size_t)i * ld + j;
            S0[o] = S4[o] * (-1.f / GW_BETA);
        }
        load_mat_big(S3, slab + lay.rawt, n, ld, 1.f, thrt, true);
        __syncthreads();
    }
    __syncthreads();
    {   // Cst terms, relu masks, threshold partial sums
        float accs = 0.f, acct = 0.f;
        for (int idx = tid; idx < n * n; idx += BIG_THREADS) {
            const int i = idx / n, j = idx - i * n;
            const size_t o = (size_t)i * ld + j;
            const float cs = S1[o], ct = S3[o];
            float a = gs[idx] + 2.f * cs * inv_n * v.dcst_r[i];
            float c = gt[idx] + 2.f * ct * inv_n * v.dcst_c[i];
            a = cs > 0.f ? a : 0.f;
            c = ct > 0.f ? c : 0.f;
            gs[idx] = a; gt[idx] = c;
            accs += a; acct += c;
        }
        accs = block_sum(accs, scratch);
        acct = block_sum(acct, scratch);
        if (tid == 0) { slab[lay.ext + 13] = -accs; slab[lay.ext + 14] = -acct; }
    }
}

// ---------------------------------------------------------------------------------------------------
// kernel D: threshold (min/max) gradients + chain rule to the token embeddings
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BIG_THREADS)
got_grad_big_kernel(GotLayout lay, float* ws, const float* __restrict__ extrema, const float* __restrict__ dthr_ext,
                    const float* __restrict__ wd, const float* __restrict__ gwd, float* __restrict__ loss,
                    float* __restrict__ dv, float* __restrict__ dq) {
    const int n = lay.n, D = lay.D;
    __shared__ float dthr[3];
    __shared__ float scratch[33];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* slab = ws + lay.header + lay.per_item * (size_t)b;
    const float* vn = slab + lay.vn;
    const float* qn = slab + lay.qn;
    const float* norms = slab + lay.norms;

    if (b == 0) {
        float s = 0.f;
        for (int i = tid; i < lay.m; i += BIG_THREADS) s += wd[i] + gwd[i];
        s = block_sum(s, scratch);
        if (tid == 0) *loss = s;
    }
    if (tid < 3) {
        float s = 0.f;
        if (dthr_ext != nullptr) {
            s = dthr_ext[tid];                   // sharded run: sums over ALL ranks' problems (all-reduced by the host side)
        } else {
            for (int i = 0; i < lay.m; ++i) s += ws[lay.header + lay.per_item * (size_t)i + lay.ext + 12 + tid];
        }
        dthr[tid] = s;
    }
    __syncthreads();
    // threshold = 0.9 min + 0.1 max: the owning element of each batch extremum receives its share of dthr
    if (tid < 6) {
        const int* h = reinterpret_cast<const int*>(ws);
        const int which = tid >> 1, is_max = tid & 1;
        const float local_best = ws[lay.header + lay.per_item * (size_t)h[tid * 2] + lay.ext + tid];
        if (h[tid * 2] == b && local_best == extrema[tid]) {
            float* g = slab + (which == 0 ? lay.g0 : (which == 1 ? lay.gs : lay.gt));
            atomicAdd(g + h[tid * 2 + 1], (is_max ? THR_BETA : 1.f - THR_BETA) * dthr[which]);
        }
    }
    __syncthreads();
    const float* g0 = slab + lay.g0;
    const float* gs = slab + lay.gs;
    const float* gt = slab + lay.gt;
    for (int r = warp; r < 2 * n; r += BIG_WARPS) {          // warp per token row; lane owns d = lane + 32 u
        const bool isv = r < n;
        const int i = isv ? r : r - n;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const float* G = isv ? gs : gt;
        for (int j = 0; j < n; ++j) {
            // cross term: v: -dC0[i,j] q^_j ; q: -dC0[j,i] v^_j.   intra term: -(G[i,j] + G[j,i]) x^_j
            const float cx = isv ? g0[(size_t)i * n + j] : g0[(size_t)j * n + i];
            const float ci = G[(size_t)i * n + j] + G[(size_t)j * n + i];
            const float* other = (isv ? qn : vn) + (size_t)j * D;
            const float* same = (isv ? vn : qn) + (size_t)j * D;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int d = lane + 32 * u;
                if (d < D) acc[u] -= cx * other[d] + ci * same[d];
            }
        }
        const float* self = (isv ? vn : qn) + (size_t)i * D;
        float dot = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int d = lane + 32 * u; if (d < D) dot = fmaf(acc[u], self[d], dot); }
        dot = warp_sum(dot);
        const float nrm = norms[r];
        const float s = 1.f / (nrm + 1e-12f);
        float* out = (isv ? dv : dq) + ((size_t)b * n + i) * D;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int d = lane + 32 * u;
            if (d < D) out[d] = s * acc[u] - (nrm > 0.f ? dot * self[d] / nrm : 0.f);
        }
    }
}

static size_t big_main_smem_bytes(int n) {
    const size_t stage = (size_t)32 * (n | 1) > (size_t)BIG_WARPS * n ? (size_t)32 * (n | 1) : (size_t)BIG_WARPS * n;
    return sizeof(float) * ((size_t)13 * n + (size_t)2 * (MAX_ITERS + 1) * n + stage);
}

static int big_set_attrs() {
    static PerDeviceOnce attr;
    unsigned long long dev_bit;
    if (attr.needed(dev_bit)) {
        MDL_CHECK_CUDA(cudaFuncSetAttribute(got_cost_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(sizeof(float) * GOT_BIG_NMAX * 129)));
        MDL_CHECK_CUDA(cudaFuncSetAttribute(got_main_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)big_main_smem_bytes(GOT_BIG_NMAX)));
        attr.mark(dev_bit);
    }
    return 0;
}

int got_big_extrema(const float* v, const float* q, int m, int n, int D, void* workspace, float* extrema, cudaStream_t st) {
    GotLayout lay(m, n, D, true);
    if (int rc = big_set_attrs()) return rc;
    got_cost_big_kernel<<<m, BIG_THREADS, sizeof(float) * (size_t)n * (D + 1), st>>>(v, q, lay, (float*)workspace);
    MDL_CHECK_LAUNCH();
    return got_launch_extrema(lay, (float*)workspace, extrema, st);
}

int got_big_main(int m, int n, int D, void* workspace, const float* extrema, float* wd, float* gwd, float* dthr_local, cudaStream_t st) {
    GotLayout lay(m, n, D, true);
    if (int rc = big_set_attrs()) return rc;
    got_main_big_kernel<<<m, BIG_THREADS, big_main_smem_bytes(n), st>>>(lay, (float*)workspace, extrema, wd, gwd);
    MDL_CHECK_LAUNCH();
    if (dthr_local != nullptr) return got_launch_dthr(lay, (const float*)workspace, dthr_local, st);
    return 0;
}

int got_big_finish(const float* v, const float* q, int m, int n, int D, void* workspace, const float* extrema, const float* dthr_global,
                   const float* wd, const float* gwd, float* loss, float* dv, float* dq, cudaStream_t st) {
    (void)v; (void)q;                      // the normalised tokens and norms were kept in the workspace by got_big_extrema
    GotLayout lay(m, n, D, true);
    got_grad_big_kernel<<<m, BIG_THREADS, 0, st>>>(lay, (float*)workspace, extrema, dthr_global, wd, gwd, loss, dv, dq);
    MDL_CHECK_LAUNCH();
    return 0;
}

}  // namespace mdl
