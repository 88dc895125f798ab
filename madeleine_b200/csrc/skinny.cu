// Exact-fp32 "skinny" linear for the slide-level projector: R (bags) is tiny next to the token count, so these
// run as warp-per-output-column dot products on the CUDA cores (no tensor-core tile would be filled).
//   fwd : Y[r, o] = sum_c X[r, c] W[o, c] + b[o]
//   bwd : dX[r, c] = sum_o dY[r, o] W[o, c];  dW[o, c] += sum_r dY[r, o] X[r, c];  db[o] += sum_r dY[r, o]
#include "common.cuh"
#include "madeleine_b200.h"

namespace mdl {

// one warp per output column o; the weight row stays in registers while the warp sweeps the R inputs.
template <int C>
__global__ void __launch_bounds__(256)
skinny_fwd_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ b, int R, int O, float* __restrict__ Y) {
    constexpr int PER = C / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = blockIdx.x * 8 + warp;
    if (o >= O) return;
    float w[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) w[i] = __ldg(W + (long long)o * C + i * 32 + lane);
    const float bias = b ? __ldg(b + o) : 0.f;
    for (int r = blockIdx.y; r < R; r += gridDim.y) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) s = fmaf(w[i], __ldg(X + (long long)r * C + i * 32 + lane), s);
        s = warp_sum(s);
        if (lane == 0) Y[(long long)r * O + o] = s + bias;
    }
}

// dX[r, c]: thread per c, loop over o (W read coalesced along c).
__global__ void __launch_bounds__(256)
skinny_dgrad_kernel(const float* __restrict__ dY, const float* __restrict__ W, int R, int O, int C, float* __restrict__ dX) {
    extern __shared__ float dy[];  // [O]
    const int r = blockIdx.y;
    for (int o = threadIdx.x; o < O; o += blockDim.x) dy[o] = __ldg(dY + (long long)r * O + o);
    __syncthreads();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
#pragma unroll 8
    for (int o = 0; o < O; ++o) s = fmaf(dy[o], __ldg(W + (long long)o * C + c), s);
    dX[(long long)r * C + c] = s;
}

// dW[o, c] += sum_r dY[r, o] X[r, c]; block = (c tile of 256, o); db[o] += sum_r dY[r,o] by the c-tile-0 block.
__global__ void __launch_bounds__(256)
skinny_wgrad_kernel(const float* __restrict__ dY, const float* __restrict__ X, int R, int O, int C, float* __restrict__ dW, float* __restrict__ db) {
    const int o = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    float s = 0.f, sb = 0.f;
    for (int r = 0; r < R; ++r) {
        const float g = __ldg(dY + (long long)r * O + o);
        sb += g;
        if (c < C) s = fmaf(g, __ldg(X + (long long)r * C + c), s);
    }
    if (c < C) dW[(long long)o * C + c] += s;
    if (db != nullptr && blockIdx.x == 0 && threadIdx.x == 0) db[o] += sb;
}

}  // namespace mdl

using namespace mdl;

extern "C" {

int mdl_skinny_linear_fwd(const float* X, const float* W, const float* b, int R, int C, int O, float* Y, void* stream) {
    MDL_REQUIRE(C == 2048 || C == 512, "skinny_linear_fwd: C must be 512 or 2048 (got %d)", C);
    if (R == 0) return 0;
    dim3 grid((O + 7) / 8, R < 16 ? R : 16);
    if (C == 2048) skinny_fwd_kernel<2048><<<grid, 256, 0, (cudaStream_t)stream>>>(X, W, b, R, O, Y);
    else skinny_fwd_kernel<512><<<grid, 256, 0, (cudaStream_t)stream>>>(X, W, b, R, O, Y);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_skinny_linear_bwd(const float* dY, const float* X, const float* W, int R, int C, int O,
                          float* dX, float* dW, float* db, void* stream) {
    if (R == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dX != nullptr) {
        dim3 grid((C + 255) / 256, R);
        skinny_dgrad_kernel<<<grid, 256, O * sizeof(float), st>>>(dY, W, R, O, C, dX);
        MDL_CHECK_LAUNCH();
    }
    if (dW != nullptr) {
        dim3 grid((C + 255) / 256, O);
        skinny_wgrad_kernel<<<grid, 256, 0, st>>>(dY, X, R, O, C, dW, db);
        MDL_CHECK_LAUNCH();
    }
    return 0;
}

}  // extern "C"
