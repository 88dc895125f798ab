// Fused multi-tensor AdamW (reference: torch.optim.AdamW(lr) at madeleine/utils/setup_components.py:194-196,
// defaults betas (0.9, 0.999), eps 1e-8, weight_decay 0.01): one launch updates every parameter tensor of the model.
// Tensor pointers travel in the kernel parameter block (<= 64 tensors), so there is no pointer table to upload.
#include "common.cuh"
#include "madeleine_b200.h"

namespace mdl {

constexpr int ADAMW_MAX_TENSORS = 64;

struct AdamWTable {
    float* p[ADAMW_MAX_TENSORS];
    const float* g[ADAMW_MAX_TENSORS];
    float* m[ADAMW_MAX_TENSORS];
    float* v[ADAMW_MAX_TENSORS];
    long long off[ADAMW_MAX_TENSORS + 1];  // prefix sums of numel
    int n;
};

__global__ void __launch_bounds__(256)
adamw_kernel(const AdamWTable t, float lr, float beta1, float beta2, float eps, float weight_decay,
             float bias_corr1, float bias_corr2_sqrt, float grad_scale) {
    const long long total = t.off[t.n];
    const float step_size = lr / bias_corr1;
    const float decay = 1.f - lr * weight_decay;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = t.n;               // tensor k with off[k] <= i < off[k+1]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (t.off[mid] <= i) lo = mid; else hi = mid;
        }
        const long long j = i - t.off[lo];
        const float g = t.g[lo][j] * grad_scale;
        float p = t.p[lo][j] * decay;
        const float m = beta1 * t.m[lo][j] + (1.f - beta1) * g;
        const float v = beta2 * t.v[lo][j] + (1.f - beta2) * g * g;
        p -= step_size * m / (sqrtf(v) / bias_corr2_sqrt + eps);
        t.p[lo][j] = p; t.m[lo][j] = m; t.v[lo][j] = v;
    }
}

}  // namespace mdl

using namespace mdl;

extern "C" {

int mdl_adamw_max_tensors(void) { return ADAMW_MAX_TENSORS; }

int mdl_adamw_step(int n_tensors, void* const* host_params, void* const* host_grads, void* const* host_exp_avg,
                   void* const* host_exp_avg_sq, const long long* host_numels, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int step, float grad_scale, void* stream) {
    MDL_REQUIRE(n_tensors > 0 && n_tensors <= ADAMW_MAX_TENSORS, "adamw: n_tensors must be in [1, %d] (got %d)", ADAMW_MAX_TENSORS, n_tensors);
    MDL_REQUIRE(step >= 1, "adamw: step counts from 1");
    AdamWTable t;
    t.n = n_tensors;
    long long o = 0;
    for (int i = 0; i < n_tensors; ++i) {
        t.p[i] = (float*)host_params[i]; t.g[i] = (const float*)host_grads[i];
        t.m[i] = (float*)host_exp_avg[i]; t.v[i] = (float*)host_exp_avg_sq[i];
        t.off[i] = o;
        o += host_numels[i];
    }
    t.off[n_tensors] = o;
    if (o == 0) return 0;
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2 = sqrtf(1.f - powf(beta2, (float)step));
    long long blocks = (o + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    adamw_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(t, lr, beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale);
    MDL_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
