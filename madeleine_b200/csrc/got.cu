// Graph Optimal Transport local loss (reference: madeleine/utils/loss.py:162-301), forward AND backward on the device.
//
//   C0  = 1 - v^ q^T,  Cs0 = 1 - v^ v^T,  Ct0 = 1 - q^ q^T        (x^ = x / (|x| + 1e-12), loss.py:172-175, 220-223)
//   X   = relu(X0 - (min + 0.1 (max - min)))  with min/max over the WHOLE batch tensor (quirk Q5, loss.py:227-233, 289-292)
//   wd  = <C, IPOT(C; beta 0.5, 30 its)>                           (loss.py:179-207, 294)
//   gwd = <Cg_5, stopgrad(gamma_5)>,  Cg_k = Cst - 2 Cs gamma_k Ct^T,  gamma_{k+1} = IPOT(Cg_k; beta 0.1, 20 its),
//         gamma_0 = 1/n^2, Cst[i,j] = mean_k Cs[i,k]^2 + mean_l Ct[j,l]^2    (loss.py:236-275)
//   loss = sum_b (wd_b + gwd_b)                                     (quirk Q4)
//
// The reference differentiates through every IPOT iteration (only the final plan multiplying Cg_5 is detached), through
// the five Cg_k products and through the global min/max of the thresholds; the reverse sweep below does the same by hand.
// IPOT state is not stored: T_t = exp(t L + lu_t[i] + lw_t[j]) with L = -C/beta and lu/lw the running log-products of
// the row/column scalings, so the backward pass recomputes any T_t from two vectors per iteration.
//
// One CTA per (case, stain) problem; n <= 96 tokens, all n x n work in shared memory (5 matrices + vectors),
// GW-level state (Cg_k, gamma_k, lu/lw per outer iteration, gradient accumulators) in a per-problem global workspace.
// Three launches: costs + extrema, main forward/backward, gradient w.r.t. the token embeddings.
#include "got_common.cuh"
#include "madeleine_b200.h"
#include <stdlib.h>

namespace mdl {

// ---------------------------------------------------------------------------------------------------
// shared-memory helpers (n x n matrices with odd leading dimension ld)
// ---------------------------------------------------------------------------------------------------
// C[i][j] = alpha * sum_k A(i,k) B(k,j); A(i,k) = TA ? A[k][i] : A[i][k]; B(k,j) = TB ? B[j][k] : B[k][j].
// warp per row i, lanes over j: A reads are broadcasts, B reads are stride-1 or stride-ld (ld odd) -> conflict free.
// Output goes to shared C (must not alias A/B) and/or is accumulated into a dense global matrix Cg (ld = n).
template <bool TA, bool TB>
__device__ __forceinline__ void mm(const float* A, const float* B, int n, int ld, float alpha, float* C, float* Cg) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < n; i += GOT_WARPS) {
        float acc[GOT_NMAX / 32];
#pragma unroll
        for (int u = 0; u < GOT_NMAX / 32; ++u) acc[u] = 0.f;
        for (int k = 0; k < n; ++k) {
            const float a = TA ? A[k * ld + i] : A[i * ld + k];
#pragma unroll
            for (int u = 0; u < GOT_NMAX / 32; ++u) {
                const int j = lane + 32 * u;
                if (j < n) acc[u] = fmaf(a, TB ? B[j * ld + k] : B[k * ld + j], acc[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < GOT_NMAX / 32; ++u) {
            const int j = lane + 32 * u;
            if (j < n) {
                if (C != nullptr) C[i * ld + j] = alpha * acc[u];
                if (Cg != nullptr) Cg[(size_t)i * n + j] += alpha * acc[u];
            }
        }
    }
}

struct Vecs {
    float *sigma, *signew, *sigprev, *delta, *dsig, *dsig_prev, *ddel, *dc, *dr;
    float *rowsq_s, *rowsq_t, *dcst_r, *dcst_c;
    float *lu, *lw;  // [(MAX_ITERS+1)][n]
};

// IPOT forward (loss.py:179-193) on L = -C/beta.  On exit T holds the plan T_K, lu/lw[t] the cumulative log scalings.
// `A` is scratch for exp(L) (computed once instead of once per iteration).
__device__ void ipot_forward(const float* L, float* T, float* A, int K, int n, int ld, const Vecs& v) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float inv_n = 1.f / n;
    for (int idx = tid; idx < n * ld; idx += GOT_THREADS) { T[idx] = 1.f; A[idx] = expf(L[idx]); }
    for (int j = tid; j < n; j += GOT_THREADS) { v.sigma[j] = inv_n; v.lu[j] = 0.f; v.lw[j] = 0.f; }
    __syncthreads();
    for (int t = 1; t <= K; ++t) {
        for (int i = warp; i < n; i += GOT_WARPS) {          // Q = A * T (in place), delta = 1 / (n Q sigma)
            float rs = 0.f;
            for (int j = lane; j < n; j += 32) {
                const float qv = A[i * ld + j] * T[i * ld + j];
                T[i * ld + j] = qv;
                rs = fmaf(qv, v.sigma[j], rs);
            }
            rs = warp_sum(rs);
            if (lane == 0) v.delta[i] = 1.f / (n * rs);
        }
        __syncthreads();
        for (int j = warp; j < n; j += GOT_WARPS) {          // sigma = 1 / (n Q^T delta)
            float cs = 0.f;
            for (int i = lane; i < n; i += 32) cs = fmaf(T[i * ld + j], v.delta[i], cs);
            cs = warp_sum(cs);
            if (lane == 0) v.signew[j] = 1.f / (n * cs);
        }
        __syncthreads();
        for (int idx = tid; idx < n * n; idx += GOT_THREADS) { // T = delta * Q * sigma^T
            const int i = idx / n, j = idx - i * n;
            T[i * ld + j] *= v.delta[i] * v.signew[j];
        }
        for (int j = tid; j < n; j += GOT_THREADS) {
            v.lu[t * n + j] = v.lu[(t - 1) * n + j] + logf(v.delta[j]);
            v.lw[t * n + j] = v.lw[(t - 1) * n + j] + logf(v.signew[j]);
            v.sigma[j] = v.signew[j];
        }
        __syncthreads();
    }
}

// Reverse sweep of ipot_forward.  dT (in: dLoss/dT_K, destroyed), dL (accumulated: dLoss/dL through A = exp(L)), Q scratch.
__device__ void ipot_backward(const float* L, float* dT, float* dL, float* Q, int K, int n, int ld, const Vecs& v) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float inv_n = 1.f / n;
    for (int j = tid; j < n; j += GOT_THREADS) v.dsig[j] = 0.f;
    __syncthreads();
    for (int t = K; t >= 1; --t) {
        for (int j = tid; j < n; j += GOT_THREADS) {
            v.delta[j] = expf(v.lu[t * n + j] - v.lu[(t - 1) * n + j]);
            v.sigma[j] = expf(v.lw[t * n + j] - v.lw[(t - 1) * n + j]);
            v.sigprev[j] = t == 1 ? inv_n : expf(v.lw[(t - 1) * n + j] - v.lw[(t - 2) * n + j]);
        }
        __syncthreads();
        // Q_t = A * T_{t-1};  ddelta[i] = sum_j dT Q sigma_t[j]
        for (int i = warp; i < n; i += GOT_WARPS) {
            float acc = 0.f;
            const float lui = v.lu[(t - 1) * n + i];
            for (int j = lane; j < n; j += 32) {
                const float l = L[i * ld + j];
                const float q = expf(t * l + lui + v.lw[(t - 1) * n + j]);   // exp(L) * exp((t-1) L + lu + lw)
                Q[i * ld + j] = q;
                acc = fmaf(dT[i * ld + j] * q, v.sigma[j], acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) v.ddel[i] = acc;
        }
        __syncthreads();
        // dsigma_t[j] += sum_i dT Q delta_t[i];  dc[j] = -dsigma_t[j] n sigma_t[j]^2
        for (int j = warp; j < n; j += GOT_WARPS) {
            float acc = 0.f;
            for (int i = lane; i < n; i += 32) acc = fmaf(dT[i * ld + j] * Q[i * ld + j], v.delta[i], acc);
            acc = warp_sum(acc);
            if (lane == 0) {
                const float ds = v.dsig[j] + acc;
                v.dc[j] = -ds * n * v.sigma[j] * v.sigma[j];
            }
        }
        __syncthreads();
        // ddelta[i] += sum_j Q dc[j];  dr[i] = -ddelta[i] n delta_t[i]^2
        for (int i = warp; i < n; i += GOT_WARPS) {
            float acc = 0.f;
            for (int j = lane; j < n; j += 32) acc = fmaf(Q[i * ld + j], v.dc[j], acc);
            acc = warp_sum(acc);
            if (lane == 0) {
                const float dd = v.ddel[i] + acc;
                v.dr[i] = -dd * n * v.delta[i] * v.delta[i];
            }
        }
        __syncthreads();
        // dsigma_{t-1}[j] = sum_i Q dr[i]
        for (int j = warp; j < n; j += GOT_WARPS) {
            float acc = 0.f;
            for (int i = lane; i < n; i += 32) acc = fmaf(Q[i * ld + j], v.dr[i], acc);
            acc = warp_sum(acc);
            if (lane == 0) v.dsig_prev[j] = acc;
        }
        // dQ = dT delta sigma^T + delta dc^T + dr sigma_{t-1}^T;  dL += dQ Q;  dT_{t-1} = dQ A
        for (int idx = tid; idx < n * n; idx += GOT_THREADS) {
            const int i = idx / n, j = idx - i * n;
            const int o = i * ld + j;
            const float dq = dT[o] * v.delta[i] * v.sigma[j] + v.delta[i] * v.dc[j] + v.dr[i] * v.sigprev[j];
            dL[o] = fmaf(dq, Q[o], dL[o]);
            dT[o] = dq * expf(L[o]);
        }
        __syncthreads();
        for (int j = tid; j < n; j += GOT_THREADS) v.dsig[j] = v.dsig_prev[j];
        __syncthreads();
    }
}

__device__ __forceinline__ void load_mat(float* S, const float* g, int n, int ld, float scale, float sub, bool relu) {
    for (int idx = threadIdx.x; idx < n * n; idx += GOT_THREADS) {
        const int i = idx / n, j = idx - i * n;
        float x = g[idx] - sub;
        if (relu) x = fmaxf(x, 0.f);
        S[i * ld + j] = x * scale;
    }
}
__device__ __forceinline__ void store_mat(const float* S, float* g, int n, int ld) {
    for (int idx = threadIdx.x; idx < n * n; idx += GOT_THREADS) {
        const int i = idx / n, j = idx - i * n;
        g[idx] = S[i * ld + j];
    }
}

// ---------------------------------------------------------------------------------------------------
// kernel A: raw cosine costs of one problem + its extrema
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GOT_THREADS)
got_cost_kernel(const float* __restrict__ v, const float* __restrict__ q, GotLayout lay, float* __restrict__ ws) {
    extern __shared__ float sm[];
    const int n = lay.n, D = lay.D, ldd = D + 1;
    float* vh = sm;                    // [n][D+1] normalised
    float* qh = sm + (size_t)n * ldd;
    __shared__ float red_val[GOT_WARPS];
    __shared__ int red_idx[GOT_WARPS];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* slab = ws + lay.header + lay.per_item * (size_t)b;
    for (int r = warp; r < 2 * n; r += GOT_WARPS) {
        const bool isv = r < n;
        const int i = isv ? r : r - n;
        const float* src = (isv ? v : q) + ((size_t)b * n + i) * D;
        float s = 0.f;
        for (int d = lane; d < D; d += 32) { const float x = src[d]; s = fmaf(x, x, s); }
        s = warp_sum(s);
        const float inv = 1.f / (sqrtf(s) + 1e-12f);
        float* dst = (isv ? vh : qh) + (size_t)i * ldd;
        for (int d = lane; d < D; d += 32) dst[d] = src[d] * inv;
    }
    __syncthreads();
    for (int which = 0; which < 3; ++which) {
        const float* X = which == 2 ? qh : vh;
        const float* Y = which == 1 ? vh : qh;
        float* out = slab + (which == 0 ? lay.raw0 : (which == 1 ? lay.raws : lay.rawt));
        float mn = INFINITY, mx = -INFINITY;
        int imn = 0x7fffffff, imx = 0x7fffffff;
        for (int idx = tid; idx < n * n; idx += GOT_THREADS) {
            const int i = idx / n, j = idx - i * n;
            float s = 0.f;
            for (int d = 0; d < D; ++d) s = fmaf(X[i * ldd + d], Y[j * ldd + d], s);
            const float c = 1.f - s;
            out[idx] = c;
            if (c < mn) { mn = c; imn = idx; }
            if (c > mx) { mx = c; imx = idx; }
        }
        // block arg-min / arg-max (ties -> smallest index)
        for (int pass = 0; pass < 2; ++pass) {
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
                for (int w = 1; w < GOT_WARPS; ++w)
                    if (red_val[w] < val || (red_val[w] == val && red_idx[w] < id)) { val = red_val[w]; id = red_idx[w]; }
                float* e = slab + lay.ext;
                e[which * 2 + pass] = pass == 0 ? val : -val;
                reinterpret_cast<int*>(e)[6 + which * 2 + pass] = id;
            }
        }
        __syncthreads();
    }
}

// kernel A2: batch extrema + owning (item, index), first occurrence.
__global__ void got_extrema_kernel(GotLayout lay, float* __restrict__ ws, float* __restrict__ extrema) {
    const int k = threadIdx.x;  // 0..5: which*2 + (0 min | 1 max)
    if (k >= 6) return;
    const bool is_max = k & 1;
    float best = is_max ? -INFINITY : INFINITY;
    int bi = 0, bidx = 0;
    for (int b = 0; b < lay.m; ++b) {
        const float* e = ws + lay.header + lay.per_item * (size_t)b + lay.ext;
        const float x = e[k];
        if (is_max ? x > best : x < best) { best = x; bi = b; bidx = reinterpret_cast<const int*>(e)[6 + k]; }
    }
    extrema[k] = best;
    int* h = reinterpret_cast<int*>(ws);
    h[k * 2] = bi;
    h[k * 2 + 1] = bidx;
}

// ---------------------------------------------------------------------------------------------------
// kernel C: forward + reverse sweep of one problem, down to gradients w.r.t. the thresholded costs
// ---------------------------------------------------------------------------------------------------
// two CTAs per SM when shared memory allows (n <= ~70): the four per-stain problem sets of a step (<= 65 CTAs each, issued on
// side streams) then fit in one wave
__global__ void __launch_bounds__(GOT_THREADS, 2)
got_main_kernel(GotLayout lay, float* __restrict__ ws, const float* __restrict__ extrema, float* __restrict__ wd_out, float* __restrict__ gwd_out) {
    extern __shared__ float sm[];
    const int n = lay.n, ld = n | 1;
    const size_t msz = (size_t)n * ld;
    float* S0 = sm; float* S1 = S0 + msz; float* S2 = S1 + msz; float* S3 = S2 + msz; float* S4 = S3 + msz;
    float* vp = S4 + msz;
    Vecs v;
    v.sigma = vp; vp += n; v.signew = vp; vp += n; v.sigprev = vp; vp += n; v.delta = vp; vp += n; v.dsig = vp; vp += n;
    v.dsig_prev = vp; vp += n; v.ddel = vp; vp += n; v.dc = vp; vp += n; v.dr = vp; vp += n;
    v.rowsq_s = vp; vp += n; v.rowsq_t = vp; vp += n; v.dcst_r = vp; vp += n; v.dcst_c = vp; vp += n;
    v.lu = vp; vp += (MAX_ITERS + 1) * n; v.lw = vp; vp += (MAX_ITERS + 1) * n;
    __shared__ float scratch[33];

    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* slab = ws + lay.header + lay.per_item * (size_t)b;
    const float thr0 = extrema[0] + THR_BETA * (extrema[1] - extrema[0]);
    const float thrs = extrema[2] + THR_BETA * (extrema[3] - extrema[2]);
    const float thrt = extrema[4] + THR_BETA * (extrema[5] - extrema[4]);
    const float inv_n = 1.f / n;

    // ================= Wasserstein term =================
    // S1 = L = -C/beta (C = relu(C0 - thr0));  S0 = T
    load_mat(S1, slab + lay.raw0, n, ld, -1.f / WD_BETA, thr0, true);
    __syncthreads();
    ipot_forward(S1, S0, S2, WD_ITERS, n, ld, v);
    {   // wd = <C, T>;  dT_K = C (into S2);  direct dC = T kept in S0;  dL accumulator S3 = 0
        float acc = 0.f;
        for (int idx = tid; idx < n * n; idx += GOT_THREADS) {
            const int i = idx / n, j = idx - i * n, o = i * ld + j;
            const float c = -WD_BETA * S1[o];
            acc = fmaf(c, S0[o], acc);
            S2[o] = c;
            S3[o] = 0.f;
        }
        acc = block_sum(acc, scratch);
        if (tid == 0) wd_out[b] = acc;
    }
    __syncthreads();
    ipot_backward(S1, S2, S3, S4, WD_ITERS, n, ld, v);
    {   // dC = T_K - dL/beta, masked by C > 0 -> g0 ; dthr0 partial = -sum
        float acc = 0.f;
        float* g0 = slab + lay.g0;
        for (int idx = tid; idx < n * n; idx += GOT_THREADS) {
            const int i = idx / n, j = idx - i * n, o = i * ld + j;
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
    for (int idx = tid; idx < n * n; idx += GOT_THREADS) { gs[idx] = 0.f; gt[idx] = 0.f; }
    // S1 = Cs, S3 = Ct (thresholded) stay resident during the forward
    load_mat(S1, slab + lay.raws, n, ld, 1.f, thrs, true);
    load_mat(S3, slab + lay.rawt, n, ld, 1.f, thrt, true);
    __syncthreads();
    for (int i = warp; i < n; i += GOT_WARPS) {
        float a = 0.f, c = 0.f;
        for (int j = lane; j < n; j += 32) { a = fmaf(S1[i * ld + j], S1[i * ld + j], a); c = fmaf(S3[i * ld + j], S3[i * ld + j], c); }
        a = warp_sum(a); c = warp_sum(c);
        if (lane == 0) { v.rowsq_s[i] = a * inv_n; v.rowsq_t[i] = c * inv_n; }
    }
    for (int idx = tid; idx < n * ld; idx += GOT_THREADS) S0[idx] = inv_n * inv_n;   // gamma_0
    __syncthreads();
    for (int k = 0; k <= GW_OUTER; ++k) {
        // Cg_k = Cst - 2 Cs gamma_k Ct^T : P = Cs gamma (S2), Cg (S4)
        mm<false, false>(S1, S0, n, ld, 1.f, S2, nullptr);
        __syncthreads();
        mm<false, true>(S2, S3, n, ld, -2.f, S4, nullptr);
        __syncthreads();
        for (int idx = tid; idx < n * n; idx += GOT_THREADS) {
            const int i = idx / n, j = idx - i * n;
            S4[i * ld + j] += v.rowsq_s[i] + v.rowsq_t[j];
        }
        __syncthreads();
        store_mat(S4, slab + lay.cg + (size_t)k * lay.nn, n, ld);
        if (k == GW_OUTER) break;
        // gamma_{k+1} = IPOT(Cg_k): L in S4 (scaled in place), plan in S0
        for (int idx = tid; idx < n * n; idx += GOT_THREADS) {
            const int i = idx / n, j = idx - i * n;
            S4[i * ld + j] *= -1.f / GW_BETA;
        }
        __syncthreads();
        ipot_forward(S4, S0, S2, GW_INNER, n, ld, v);
        store_mat(S0, slab + lay.gamma + (size_t)k * lay.nn, n, ld);
        float* lulw = slab + lay.lulw + (size_t)k * 2 * (GW_INNER + 1) * n;
        for (int idx = tid; idx < (GW_INNER + 1) * n; idx += GOT_THREADS) {
            lulw[idx] = v.lu[idx];
            lulw[(GW_INNER + 1) * n + idx] = v.lw[idx];
        }
        __syncthreads();
    }
    {   // gwd = <Cg_5, gamma_5>   (S4 = Cg_5, S0 = gamma_5)
        float acc = 0.f;
        for (int idx = tid; idx < n * n; idx += GOT_THREADS) {
            const int i = idx / n, j = idx - i * n, o = i * ld + j;
            acc = fmaf(S4[o], S0[o], acc);
        }
        acc = block_sum(acc, scratch);
        if (tid == 0) gwd_out[b] = acc;
    }
    for (int j = tid; j < n; j += GOT_THREADS) { v.dcst_r[j] = 0.f; v.dcst_c[j] = 0.f; }
    __syncthreads();

    // ---- reverse sweep.  Invariant at the top of each step k: S0 = dCg_k, S1 = Cs, S3 = Ct. ----
    // dCg_5 = gamma_5 (already in S0; the plan multiplying Cg_5 is detached, so this is the only upstream of Cg_5).
    for (int k = GW_OUTER; k >= 0; --k) {
        // column/row sums of dCg -> dCst
        for (int i = warp; i < n; i += GOT_WARPS) {
            float a = 0.f;
            for (int j = lane; j < n; j += 32) a += S0[i * ld + j];
            a = warp_sum(a);
            if (lane == 0) v.dcst_r[i] += a;
        }
        for (int j = warp; j < n; j += GOT_WARPS) {
            float a = 0.f;
            for (int i = lane; i < n; i += 32) a += S0[i * ld + j];
            a = warp_sum(a);
            if (lane == 0) v.dcst_c[j] += a;
        }
        // gamma_k -> S2
        if (k == 0) {
            for (int idx = tid; idx < n * ld; idx += GOT_THREADS) S2[idx] = inv_n * inv_n;
        } else {
            load_mat(S2, slab + lay.gamma + (size_t)(k - 1) * lay.nn, n, ld, 1.f, 0.f, false);
        }
        __syncthreads();
        // P = Cs gamma_k -> S4 ;  dCt += -2 dCg^T P
        mm<false, false>(S1, S2, n, ld, 1.f, S4, nullptr);
        __syncthreads();
        mm<true, false>(S0, S4, n, ld, -2.f, nullptr, gt);
        __syncthreads();
        // dP = -2 dCg Ct -> S4 ;  dCs += dP gamma_k^T
        mm<false, false>(S0, S3, n, ld, -2.f, S4, nullptr);
        __syncthreads();
        mm<false, true>(S4, S2, n, ld, 1.f, nullptr, gs);
        if (k == 0) break;               // gamma_0 is a constant
        // dgamma_k = Cs^T dP -> S0 (dCg_k no longer needed after the products above)
        __syncthreads();
        mm<true, false>(S1, S4, n, ld, 1.f, S0, nullptr);
        __syncthreads();
        // through gamma_k = IPOT(Cg_{k-1}):  L -> S2, dL -> S4 (zeroed), Q scratch needs a matrix: Ct is reloaded afterwards
        load_mat(S2, slab + lay.cg + (size_t)(k - 1) * lay.nn, n, ld, -1.f / GW_BETA, 0.f, false);
        for (int idx = tid; idx < n * ld; idx += GOT_THREADS) S4[idx] = 0.f;
        const float* lulw = slab + lay.lulw + (size_t)(k - 1) * 2 * (GW_INNER + 1) * n;
        for (int idx = tid; idx < (GW_INNER + 1) * n; idx += GOT_THREADS) {
            v.lu[idx] = lulw[idx];
            v.lw[idx] = lulw[(GW_INNER + 1) * n + idx];
        }
        __syncthreads();
        ipot_backward(S2, S0, S4, S3, GW_INNER, n, ld, v);
        // dCg_{k-1} = -dL / beta -> S0 ; restore Ct in S3
        for (int idx = tid; idx < n * n; idx += GOT_THREADS) {
            const int i = idx / n, j = idx - i * n, o = i * ld + j;
            S0[o] = S4[o] * (-1.f / GW_BETA);
        }
        load_mat(S3, slab + lay.rawt, n, ld, 1.f, thrt, true);
        __syncthreads();
    }
    __syncthreads();
    // Cst terms, relu masks, threshold partial sums
    {
        float accs = 0.f, acct = 0.f;
        for (int idx = tid; idx < n * n; idx += GOT_THREADS) {
            const int i = idx / n, j = idx - i * n, o = i * ld + j;
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
__global__ void __launch_bounds__(GOT_THREADS)
got_grad_kernel(const float* __restrict__ v, const float* __restrict__ q, GotLayout lay, float* __restrict__ ws,
                const float* __restrict__ extrema, const float* __restrict__ dthr_ext,
                const float* __restrict__ wd, const float* __restrict__ gwd, float* __restrict__ loss,
                float* __restrict__ dv, float* __restrict__ dq) {
    extern __shared__ float sm[];
    const int n = lay.n, D = lay.D, ldd = D + 1;
    float* vh = sm;
    float* qh = sm + (size_t)n * ldd;
    float* nv = qh + (size_t)n * ldd;   // norms |v_i|
    float* nq = nv + n;
    __shared__ float dthr[3];
    __shared__ float scratch[33];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* slab = ws + lay.header + lay.per_item * (size_t)b;

    if (b == 0) {
        float s = 0.f;
        for (int i = tid; i < lay.m; i += GOT_THREADS) s += wd[i] + gwd[i];
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
    for (int r = warp; r < 2 * n; r += GOT_WARPS) {
        const bool isv = r < n;
        const int i = isv ? r : r - n;
        const float* src = (isv ? v : q) + ((size_t)b * n + i) * D;
        float s = 0.f;
        for (int d = lane; d < D; d += 32) { const float x = src[d]; s = fmaf(x, x, s); }
        s = warp_sum(s);
        const float nrm = sqrtf(s), inv = 1.f / (nrm + 1e-12f);
        float* dst = (isv ? vh : qh) + (size_t)i * ldd;
        for (int d = lane; d < D; d += 32) dst[d] = src[d] * inv;
        if (lane == 0) (isv ? nv : nq)[i] = nrm;
    }
    __syncthreads();
    // threshold = 0.9 min + 0.1 max: the owning element of each batch extremum receives its share of dthr
    if (tid < 6) {
        const int* h = reinterpret_cast<const int*>(ws);
        const int which = tid >> 1, is_max = tid & 1;
        // the local candidate owns the extremum only if it equals the (possibly all-reduced) batch extremum
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
    // warp per token row; lane owns d = lane + 32 u
    for (int r = warp; r < 2 * n; r += GOT_WARPS) {
        const bool isv = r < n;
        const int i = isv ? r : r - n;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < n; ++j) {
            // cross term: v: -dC0[i,j] q^_j ; q: -dC0[j,i] v^_j.   intra term: -(G[i,j] + G[j,i]) x^_j
            const float cx = isv ? g0[(size_t)i * n + j] : g0[(size_t)j * n + i];
            const float* G = isv ? gs : gt;
            const float ci = G[(size_t)i * n + j] + G[(size_t)j * n + i];
            const float* other = (isv ? qh : vh) + (size_t)j * ldd;
            const float* same = (isv ? vh : qh) + (size_t)j * ldd;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int d = lane + 32 * u;
                if (d < D) acc[u] -= cx * other[d] + ci * same[d];
            }
        }
        const float* self = (isv ? vh : qh) + (size_t)i * ldd;
        float dot = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int d = lane + 32 * u; if (d < D) dot = fmaf(acc[u], self[d], dot); }
        dot = warp_sum(dot);
        const float nrm = (isv ? nv : nq)[i];
        const float s = 1.f / (nrm + 1e-12f);
        float* out = (isv ? dv : dq) + ((size_t)b * n + i) * D;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int d = lane + 32 * u;
            if (d < D) out[d] = s * acc[u] - (nrm > 0.f ? dot * self[d] / nrm : 0.f);
        }
    }
}

// local sums of the three threshold-gradient partials (one launch, 3 threads) for the sharded path
__global__ void got_dthr_kernel(GotLayout lay, const float* __restrict__ ws, float* __restrict__ dthr_out) {
    const int k = threadIdx.x;
    if (k >= 3) return;
    float s = 0.f;
    for (int i = 0; i < lay.m; ++i) s += ws[lay.header + lay.per_item * (size_t)i + lay.ext + 12 + k];
    dthr_out[k] = s;
}

int got_launch_extrema(const GotLayout& lay, float* ws, float* extrema, cudaStream_t st) {
    got_extrema_kernel<<<1, 32, 0, st>>>(lay, ws, extrema);
    MDL_CHECK_LAUNCH();
    return 0;
}
int got_launch_dthr(const GotLayout& lay, const float* ws, float* dthr_out, cudaStream_t st) {
    got_dthr_kernel<<<1, 32, 0, st>>>(lay, ws, dthr_out);
    MDL_CHECK_LAUNCH();
    return 0;
}

static size_t main_smem_bytes(int n) {
    const int ld = n | 1;
    return sizeof(float) * ((size_t)5 * n * ld + (size_t)13 * n + (size_t)2 * (MAX_ITERS + 1) * n);
}

}  // namespace mdl

using namespace mdl;

// Problems above GOT_NMAX tokens go to the global-memory kernels; MDL_GOT_FORCE_BIG=1 (or mdl_got_force_big) sends every
// problem there so the two implementations can be compared on the same inputs.
static int g_force_big = -1;
static bool got_use_big(int n) {
    if (g_force_big < 0) {
        const char* e = getenv("MDL_GOT_FORCE_BIG");
        g_force_big = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return n > GOT_NMAX || g_force_big == 1;
}

extern "C" {

int mdl_got_force_big(int on) { g_force_big = on ? 1 : 0; return 0; }

int mdl_got_max_tokens(void) { return GOT_BIG_NMAX; }

long long mdl_got_workspace_bytes(int m, int n, int D) {
    if (m <= 0 || n <= 0) return 0;
    GotLayout lay(m, n, D, got_use_big(n));
    return (long long)(lay.total_floats() * sizeof(float));
}

int mdl_got_extrema(const float* v, const float* q, int m, int n, int D, void* workspace, float* extrema, void* stream) {
    MDL_REQUIRE(m > 0 && n > 0 && n <= GOT_BIG_NMAX, "GOT: n must be in [1, %d] (got %d)", GOT_BIG_NMAX, n);
    MDL_REQUIRE(D > 0 && D <= 128, "GOT: token dim must be <= 128 (got %d)", D);
    cudaStream_t st = (cudaStream_t)stream;
    if (got_use_big(n)) return got_big_extrema(v, q, m, n, D, workspace, extrema, st);
    GotLayout lay(m, n, D, got_use_big(n));
    const size_t smem = sizeof(float) * (size_t)2 * n * (D + 1);
    static PerDeviceOnce attr;
    unsigned long long dev_bit;
    if (attr.needed(dev_bit)) {
        MDL_CHECK_CUDA(cudaFuncSetAttribute(got_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * 2 * GOT_NMAX * 129)));
        attr.mark(dev_bit);
    }
    got_cost_kernel<<<m, GOT_THREADS, smem, st>>>(v, q, lay, (float*)workspace);
    MDL_CHECK_LAUNCH();
    return got_launch_extrema(lay, (float*)workspace, extrema, st);
}

static int got_set_attrs() {
    static PerDeviceOnce attr;
    unsigned long long dev_bit;
    if (attr.needed(dev_bit)) {
        MDL_CHECK_CUDA(cudaFuncSetAttribute(got_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)main_smem_bytes(GOT_NMAX)));
        MDL_CHECK_CUDA(cudaFuncSetAttribute(got_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * (2 * GOT_NMAX * 129 + 2 * GOT_NMAX))));
        attr.mark(dev_bit);
    }
    return 0;
}

int mdl_got_main(int m, int n, int D, void* workspace, const float* extrema, float* wd, float* gwd, float* dthr_local, void* stream) {
    MDL_REQUIRE(m > 0 && n > 0 && n <= GOT_BIG_NMAX, "GOT: n must be in [1, %d] (got %d)", GOT_BIG_NMAX, n);
    cudaStream_t st = (cudaStream_t)stream;
    if (got_use_big(n)) return got_big_main(m, n, D, workspace, extrema, wd, gwd, dthr_local, st);
    GotLayout lay(m, n, D, got_use_big(n));
    if (int rc = got_set_attrs()) return rc;
    got_main_kernel<<<m, GOT_THREADS, main_smem_bytes(n), st>>>(lay, (float*)workspace, extrema, wd, gwd);
    MDL_CHECK_LAUNCH();
    if (dthr_local != nullptr) {
        got_dthr_kernel<<<1, 32, 0, st>>>(lay, (const float*)workspace, dthr_local);
        MDL_CHECK_LAUNCH();
    }
    return 0;
}

int mdl_got_finish(const float* v, const float* q, int m, int n, int D, void* workspace, const float* extrema, const float* dthr_global,
                   const float* wd, const float* gwd, float* loss, float* dv, float* dq, void* stream) {
    MDL_REQUIRE(m > 0 && n > 0 && n <= GOT_BIG_NMAX, "GOT: n must be in [1, %d] (got %d)", GOT_BIG_NMAX, n);
    MDL_REQUIRE(D > 0 && D <= 128, "GOT: token dim must be <= 128 (got %d)", D);
    cudaStream_t st = (cudaStream_t)stream;
    if (got_use_big(n)) return got_big_finish(v, q, m, n, D, workspace, extrema, dthr_global, wd, gwd, loss, dv, dq, st);
    GotLayout lay(m, n, D, got_use_big(n));
    if (int rc = got_set_attrs()) return rc;
    const size_t smem = sizeof(float) * ((size_t)2 * n * (D + 1) + 2 * n);
    got_grad_kernel<<<m, GOT_THREADS, smem, st>>>(v, q, lay, (float*)workspace, extrema, dthr_global, wd, gwd, loss, dv, dq);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_got_fwd_bwd(const float* v, const float* q, int m, int n, int D, void* workspace, const float* extrema,
                    float* loss, float* wd, float* gwd, float* dv, float* dq, void* stream) {
    MDL_REQUIRE(D > 0 && D <= 128, "GOT: token dim must be <= 128 (got %d)", D);
    int rc = mdl_got_main(m, n, D, workspace, extrema, wd, gwd, nullptr, stream);
    if (rc) return rc;
    return mdl_got_finish(v, q, m, n, D, workspace, extrema, nullptr, wd, gwd, loss, dv, dq, stream);
}

}  // extern "C"
