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

__device__ __forceinline__ void adamw_update(float& p, float g, float& m, float& v, float decay, float step_size, float beta1, float beta2,
                                             float eps, float bias_corr2_sqrt, float grad_scale) {
    g *= grad_scale;
    p *= decay;
    m = beta1 * m + (1.f - beta1) * g;
    v = beta2 * v + (1.f - beta2) * g * g;
    p -= step_size * m / (sqrtf(v) / bias_corr2_sqrt + eps);
}

// Tensors are swept one after the other by the whole grid (with the flattened parameter / gradient / moment buffers the
// model is 1-3 long runs): 16-byte vectors wherever the four pointers of a tensor are 16-byte aligned, scalars for the
// unaligned head / tail and for unaligned tensors.  28 bytes move per parameter; the kernel is a pure HBM stream.
__global__ void __launch_bounds__(256)
adamw_kernel(const AdamWTable t, float lr, float beta1, float beta2, float eps, float weight_decay,
             float bias_corr1, float bias_corr2_sqrt, float grad_scale) {
    pdl_sync();
    const float step_size = lr / bias_corr1;
    const float decay = 1.f - lr * weight_decay;
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nthreads = (long long)gridDim.x * blockDim.x;
    for (int k = 0; k < t.n; ++k) {
        float* __restrict__ p = t.p[k];
        const float* __restrict__ g = t.g[k];
        float* __restrict__ m = t.m[k];
        float* __restrict__ v = t.v[k];
        const long long n = t.off[k + 1] - t.off[k];
        const bool aligned = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15u) == 0;
        const long long n4 = aligned ? n / 4 : 0;
        for (long long i = tid; i < n4; i += nthreads) {
            float4 pp = reinterpret_cast<float4*>(p)[i];
            const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
            float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
            adamw_update(pp.x, gg.x, mm.x, vv.x, decay, step_size, beta1, beta2, eps, bias_corr2_sqrt, grad_scale);
            adamw_update(pp.y, gg.y, mm.y, vv.y, decay, step_size, beta1, beta2, eps, bias_corr2_sqrt, grad_scale);
            adamw_update(pp.z, gg.z, mm.z, vv.z, decay, step_size, beta1, beta2, eps, bias_corr2_sqrt, grad_scale);
            adamw_update(pp.w, gg.w, mm.w, vv.w, decay, step_size, beta1, beta2, eps, bias_corr2_sqrt, grad_scale);
            reinterpret_cast<float4*>(p)[i] = pp;
            reinterpret_cast<float4*>(m)[i] = mm;
            reinterpret_cast<float4*>(v)[i] = vv;
        }
        for (long long j = 4 * n4 + tid; j < n; j += nthreads) {
            float pp = p[j], mm = m[j], vv = v[j];
            adamw_update(pp, g[j], mm, vv, decay, step_size, beta1, beta2, eps, bias_corr2_sqrt, grad_scale);
            p[j] = pp; m[j] = mm; v[j] = vv;
        }
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
    long long blocks = (o / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    launch_k(adamw_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, t, lr, beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale);
    MDL_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
