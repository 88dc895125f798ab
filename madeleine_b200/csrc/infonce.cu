// Fused symmetric InfoNCE with in-batch negatives (reference: madeleine/utils/loss.py:111-127).
//   q^ = q / max(|q|, 1e-12), k^ likewise;  L = q^ k^T;  nll_r[i] = lse_j(L[i,:]/tau) - L[i,i]/tau;
//   nll_c[j] = lse_i(L[:,j]/tau) - L[j,j]/tau  (symmetric term).
// Forward writes L, the two log-sum-exp vectors (as [max | log-sum], 2m floats each) and the per-sample nll vectors; backward turns per-sample upstream
// weights into dL and then into dq, dk through the normalisation.  m <= a few thousand, D arbitrary (multiple of 4);
// the problem is latency-bound, so the kernels are sized for launch count, not bandwidth.
#include "common.cuh"
#include "madeleine_b200.h"

namespace mdl {

// Row i of an operand: base + (rows ? rows[i] : i) * ld.  With an index list the cases that carry a stain are read straight
// out of the encoder's slide-embedding matrix (trainer.py:30-33 selects them with a boolean mask) — no gathered copy.
struct Rows {
    const float* base; const int* rows; long long ld;
    __device__ __forceinline__ const float* row(int i) const { return base + (long long)(rows ? __ldg(rows + i) : i) * ld; }
    __device__ __forceinline__ long long off(int i) const { return (long long)(rows ? __ldg(rows + i) : i) * ld; }
};

__global__ void __launch_bounds__(256)
infonce_norm_kernel(Rows q, Rows k, int m, int D, float* __restrict__ qn, float* __restrict__ kn) {
    pdl_sync();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    if (row >= 2 * m) return;
    const float* x = row < m ? q.row(row) : k.row(row - m);
    float s = 0.f;
    for (int d = lane; d < D; d += 32) { const float v = __ldg(x + d); s = fmaf(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) {
        const float inv = 1.f / fmaxf(sqrtf(s), 1e-12f);
        if (row < m) qn[row] = inv; else kn[row - m] = inv;
    }
}

// L[i, j] = sum_d (q[i,d] * qn[i]) * (k[j,d] * kn[j]); block = one i, warps sweep j.
__global__ void __launch_bounds__(256)
infonce_logits_kernel(Rows q, Rows k, const float* __restrict__ qn, const float* __restrict__ kn,
                      int m, int D, float* __restrict__ L) {
    pdl_sync();
    extern __shared__ float qs[];  // normalised q_i
    const int i = blockIdx.x;
    const float qi = qn[i];
    for (int d = threadIdx.x; d < D; d += blockDim.x) qs[d] = __ldg(q.row(i) + d) * qi;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = warp; j < m; j += 8) {
        const float kj = kn[j];
        const float* kr = k.row(j);
        float s = 0.f;
        for (int d = lane; d < D; d += 32) s = fmaf(qs[d], __ldg(kr + d) * kj, s);
        s = warp_sum(s);
        if (lane == 0) L[(long long)i * m + j] = s;
    }
}

// warp w < m: row w;  m <= w < 2m: column w-m.  lse of L/tau and nll against the diagonal.
__global__ void __launch_bounds__(256)
infonce_lse_kernel(const float* __restrict__ L, int m, float inv_tau, float* __restrict__ lse_r, float* __restrict__ lse_c,
                   float* __restrict__ nll_r, float* __restrict__ nll_c) {
    pdl_sync();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * 8 + warp;
    if (w >= 2 * m) return;
    const bool is_row = w < m;
    const int a = is_row ? w : w - m;
    const long long base = is_row ? (long long)a * m : a, stride = is_row ? 1 : m;
    float mx = -INFINITY;
    for (int t = lane; t < m; t += 32) mx = fmaxf(mx, __ldg(L + base + t * stride) * inv_tau);
    mx = warp_max(mx);
    float s = 0.f;
    for (int t = lane; t < m; t += 32) s += expf(__ldg(L + base + t * stride) * inv_tau - mx);
    s = warp_sum(s);
    if (lane == 0) {
        // keep max and log-sum apart: (l - max) - log(sum) is exact for the dominant element, l - (max + log(sum)) is
        // not once |l| ~ 1/tau = 1000 (fp32 spacing 6e-5), and the gradient is the small difference softmax - 1.
        const float ls = logf(s);
        const float nll = ls - (__ldg(L + (long long)a * m + a) * inv_tau - mx);
        if (is_row) { lse_r[a] = mx; lse_r[m + a] = ls; nll_r[a] = nll; } else { lse_c[a] = mx; lse_c[m + a] = ls; nll_c[a] = nll; }
    }
}

// loss = scale_r * sum nll_r + scale_c * sum nll_c   (single block)
__global__ void __launch_bounds__(256)
infonce_reduce_kernel(const float* __restrict__ nll_r, const float* __restrict__ nll_c, int m, float scale_r, float scale_c, float* __restrict__ loss,
                      float* __restrict__ total) {
    pdl_sync();
    __shared__ float scratch[33];
    float s = 0.f;
    for (int i = threadIdx.x; i < m; i += blockDim.x) s += scale_r * nll_r[i] + (scale_c != 0.f ? scale_c * nll_c[i] : 0.f);
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) {
        *loss = s;
        if (total != nullptr) *total += s;      // running sum over the stains of a step (launches are stream-ordered)
    }
}

// G[i,j] = inv_tau * ( w_r[i] * (softmax_row[i,j] - d_ij) + w_c[j] * (softmax_col[i,j] - d_ij) )
__global__ void __launch_bounds__(256)
infonce_dlogits_kernel(const float* __restrict__ L, const float* __restrict__ lse_r, const float* __restrict__ lse_c,
                       const float* __restrict__ w_r, const float* __restrict__ w_c, int m, float inv_tau, float* __restrict__ G,
                       const float* __restrict__ go, float wr_const, float wc_const) {
    pdl_sync();
    const long long total = (long long)m * m;
    // go != NULL: every sample's weight is (*go) * w{r,c}_const (mean / symmetric scaling folded in by the host)
    const float g0 = go != nullptr ? __ldg(go) : 0.f;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / m), j = (int)(idx % m);
        const float l = __ldg(L + idx) * inv_tau;
        const float dij = i == j ? 1.f : 0.f;
        const float wr = go != nullptr ? g0 * wr_const : __ldg(w_r + i);
        float g = wr * (expf((l - __ldg(lse_r + i)) - __ldg(lse_r + m + i)) - dij);
        if (go != nullptr) {
            if (wc_const != 0.f) g += g0 * wc_const * (expf((l - __ldg(lse_c + j)) - __ldg(lse_c + m + j)) - dij);
        } else if (w_c != nullptr) {
            g += __ldg(w_c + j) * (expf((l - __ldg(lse_c + j)) - __ldg(lse_c + m + j)) - dij);
        }
        G[idx] = g * inv_tau;
    }
}

// rows 0..m-1: dq_i ; rows m..2m-1: dk_j.   dx = (dxh - xh (xh . dxh)) * inv_norm,  dxh_i = sum_j G[i,j] kh_j  (or G^T qh).
// The result goes to dq.row(a) / dk.row(a) (same row addressing as the operands); accumulate != 0 adds to what is there
// (the H&E rows of a step collect the gradients of every stain's term, launch after launch on one stream).
constexpr int INFONCE_MAX_U = 8;   // D <= 128 * 8
__global__ void __launch_bounds__(128)
infonce_grads_kernel(Rows q, Rows k, const float* __restrict__ qn, const float* __restrict__ kn,
                     const float* __restrict__ G, int m, int D, float* __restrict__ dq, float* __restrict__ dk, int accumulate) {
    pdl_sync();
    __shared__ float scratch[33];
    extern __shared__ float gs[];  // coefficients G[i,:] * kn[:]  (or G[:,j] * qn[:])
    const int row = blockIdx.x;
    const bool is_q = row < m;
    const int a = is_q ? row : row - m;
    const Rows& self = is_q ? q : k;
    const Rows& other = is_q ? k : q;
    const float* other_n = is_q ? kn : qn;
    const float self_n = is_q ? qn[a] : kn[a];
    for (int t = threadIdx.x; t < m; t += blockDim.x)
        gs[t] = (is_q ? __ldg(G + (long long)a * m + t) : __ldg(G + (long long)t * m + a)) * __ldg(other_n + t);
    if (threadIdx.x < 8) gs[m + threadIdx.x] = 0.f;           // padding behind the coefficients (see the launch)
    __syncthreads();
    float dot = 0.f;
    float acc[INFONCE_MAX_U];
    const float* xs = self.row(a);
    // each thread owns columns d = threadIdx.x + 128*u
#pragma unroll
    for (int u = 0; u < INFONCE_MAX_U; ++u) {
        acc[u] = 0.f;
        const int d = threadIdx.x + u * 128;
        if (d < D) {
            float s = 0.f;
            for (int t = 0; t < m; ++t) s = fmaf(gs[t], __ldg(other.row(t) + d), s);
            acc[u] = s;
            dot = fmaf(s, __ldg(xs + d) * self_n, dot);
        }
    }
    dot = block_sum(dot, scratch);
    const bool clamped = self_n >= 1e12f;  // |x| < eps: F.normalize divides by eps, a constant
    float* out = (is_q ? dq : dk) + self.off(a);
#pragma unroll
    for (int u = 0; u < INFONCE_MAX_U; ++u) {
        const int d = threadIdx.x + u * 128;
        if (d < D) {
            const float xh = __ldg(xs + d) * self_n;
            const float v = clamped ? acc[u] * self_n : (acc[u] - xh * dot) * self_n;
            out[d] = accumulate ? out[d] + v : v;
        }
    }
}

}  // namespace mdl

using namespace mdl;

extern "C" {

static int infonce_fwd_impl(Rows q, Rows k, int m, int D, float temperature, int symmetric, int reduction,
                            float* qn, float* kn, float* L, float* lse_r, float* lse_c, float* nll_r, float* nll_c, float* loss,
                            float* total, cudaStream_t st) {
    MDL_REQUIRE(m > 0 && D > 0, "infonce: empty input");
    MDL_REQUIRE(temperature > 0.f, "infonce: temperature must be positive");
    MDL_REQUIRE((size_t)D * sizeof(float) <= 48 * 1024, "infonce: D too large (%d)", D);
    const float inv_tau = 1.f / temperature;
    launch_k(infonce_norm_kernel, dim3((2 * m + 7) / 8), dim3(256), 0, st, q, k, m, D, qn, kn);
    MDL_CHECK_LAUNCH();
    launch_k(infonce_logits_kernel, dim3(m), dim3(256), D * sizeof(float), st, q, k, qn, kn, m, D, L);
    MDL_CHECK_LAUNCH();
    launch_k(infonce_lse_kernel, dim3((2 * m + 7) / 8), dim3(256), 0, st, L, m, inv_tau, lse_r, lse_c, nll_r, nll_c);
    MDL_CHECK_LAUNCH();
    if (loss != nullptr && reduction != 0) {  // 1 = mean, 2 = sum
        const float base = reduction == 1 ? 1.f / m : 1.f;
        const float sr = symmetric ? 0.5f * base : base, sc = symmetric ? 0.5f * base : 0.f;
        launch_k(infonce_reduce_kernel, dim3(1), dim3(256), 0, st, nll_r, nll_c, m, sr, sc, loss, total);
        MDL_CHECK_LAUNCH();
    }
    return 0;
}

static int infonce_bwd_impl(Rows q, Rows k, int m, int D, float temperature, const float* qn, const float* kn, const float* L,
                            const float* lse_r, const float* lse_c, const float* w_r, const float* w_c, const float* go, float wr_const,
                            float wc_const, float* G, float* dq, float* dk, int accumulate, cudaStream_t st) {
    MDL_REQUIRE(m > 0 && D > 0, "infonce: empty input");
    MDL_REQUIRE((size_t)m * sizeof(float) <= 40 * 1024, "infonce_bwd: m too large (%d)", m);
    MDL_REQUIRE(D <= 128 * INFONCE_MAX_U, "infonce_bwd: D too large (%d)", D);
    const float inv_tau = 1.f / temperature;
    const long long total = (long long)m * m;
    int blocks = (int)((total + 255) / 256);
    if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
    launch_k(infonce_dlogits_kernel, dim3(blocks), dim3(256), 0, st, L, lse_r, lse_c, w_r, w_c, m, inv_tau, G, go, wr_const, wc_const);
    MDL_CHECK_LAUNCH();
    // + 8 floats: the unrolled dot-product loop may load a few coefficients past m (speculatively, never used)
    launch_k(infonce_grads_kernel, dim3(2 * m), dim3(128), (m + 8) * sizeof(float), st, q, k, qn, kn, G, m, D, dq, dk, accumulate);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_infonce_fwd(const float* q, const float* k, int m, int D, float temperature, int symmetric, int reduction,
                    float* qn, float* kn, float* L, float* lse_r, float* lse_c, float* nll_r, float* nll_c, float* loss, void* stream) {
    return infonce_fwd_impl(Rows{q, nullptr, D}, Rows{k, nullptr, D}, m, D, temperature, symmetric, reduction, qn, kn, L, lse_r, lse_c,
                            nll_r, nll_c, loss, nullptr, (cudaStream_t)stream);
}

int mdl_infonce_bwd(const float* q, const float* k, int m, int D, float temperature,
                    const float* qn, const float* kn, const float* L, const float* lse_r, const float* lse_c,
                    const float* w_r, const float* w_c, float* G, float* dq, float* dk, void* stream) {
    return infonce_bwd_impl(Rows{q, nullptr, D}, Rows{k, nullptr, D}, m, D, temperature, qn, kn, L, lse_r, lse_c, w_r, w_c, nullptr, 0.f, 0.f,
                            G, dq, dk, 0, (cudaStream_t)stream);
}

long long mdl_infonce_rows_workspace_floats(int m) { return (long long)m * m * 2 + 8LL * m + 64; }

int mdl_infonce_rows_fwd(const float* base, long long ld, const int* q_rows, const int* k_rows, int m, int D, float temperature,
                         int symmetric, float* workspace, float* loss, float* total, void* stream) {
    MDL_REQUIRE(base && q_rows && k_rows && workspace && loss, "infonce_rows: null argument");
    float* w = workspace;
    float *L = w, *G = w + (long long)m * m, *qn = G + (long long)m * m, *kn = qn + m, *lse_r = kn + m, *lse_c = lse_r + 2 * m,
          *nll_r = lse_c + 2 * m, *nll_c = nll_r + m;
    (void)G;
    return infonce_fwd_impl(Rows{base, q_rows, ld}, Rows{base, k_rows, ld}, m, D, temperature, symmetric, 1, qn, kn, L, lse_r, lse_c, nll_r,
                            nll_c, loss, total, (cudaStream_t)stream);
}

int mdl_infonce_rows_bwd(const float* base, long long ld, const int* q_rows, const int* k_rows, int m, int D, float temperature,
                         int symmetric, float* workspace, const float* go, float* dbase, void* stream) {
    MDL_REQUIRE(base && q_rows && k_rows && workspace && go && dbase, "infonce_rows: null argument");
    float* w = workspace;
    float *L = w, *G = w + (long long)m * m, *qn = G + (long long)m * m, *kn = qn + m, *lse_r = kn + m, *lse_c = lse_r + 2 * m;
    const float wr = (symmetric ? 0.5f : 1.f) / m, wc = symmetric ? 0.5f / m : 0.f;
    return infonce_bwd_impl(Rows{base, q_rows, ld}, Rows{base, k_rows, ld}, m, D, temperature, qn, kn, L, lse_r, lse_c, nullptr, nullptr, go,
                            wr, wc, G, dbase, dbase, 1, (cudaStream_t)stream);
}

}  // extern "C"
