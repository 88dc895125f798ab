// Exact-fp32 "skinny" linear for the slide-level projector: R (bags) is tiny next to the token count, so these
// run as register-tiled dot products on the CUDA cores (no tensor-core tile would be filled).
//   fwd : Y[r, o] = sum_c X[r, c] W[o, c] + b[o]
//   bwd : dX[r, c] = sum_o dY[r, o] W[o, c];  dW[o, c] += sum_r dY[r, o] X[r, c];  db[o] += sum_r dY[r, o]
#include "common.cuh"
#include "madeleine_b200.h"

namespace mdl {

// Forward: one warp per output column o.  The weight row stays in registers (16-byte vectors) while the warp sweeps a tile of
// `rows` (<= 32) input rows, so W crosses L2 -> SM once per tile; every X access is a 512-byte warp-wide vector load; lane i of the warp
// keeps the result of the tile's row i and the tile is written with one store instruction.  Blocks start at different rows of
// the tile (all blocks of a tile read the same X rows: walking them in lockstep would hammer one L2 line at a time).
template <int C>
__global__ void __launch_bounds__(128)
skinny_fwd_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ b, int R, int O, int rows, float* __restrict__ Y) {
    constexpr int PER4 = C / 128;
    pdl_sync();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = blockIdx.x * 4 + warp;
    if (o >= O) return;
    float4 w[PER4];
    const float4* W4 = reinterpret_cast<const float4*>(W + (size_t)o * C);
#pragma unroll
    for (int i = 0; i < PER4; ++i) w[i] = __ldg(W4 + i * 32 + lane);
    const float bias = b ? __ldg(b + o) : 0.f;
    const int r0 = blockIdx.y * rows, n_r = min(R - r0, rows);
    float keep = 0.f;
    int rr = (int)(blockIdx.x % (unsigned)n_r);
    for (int it = 0; it < n_r; ++it) {
        const float4* X4 = reinterpret_cast<const float4*>(X + (size_t)(r0 + rr) * C);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int i = 0; i < PER4; ++i) {
            const float4 x = __ldg(X4 + i * 32 + lane);
            s0 = fmaf(w[i].x, x.x, s0); s1 = fmaf(w[i].y, x.y, s1); s2 = fmaf(w[i].z, x.z, s2); s3 = fmaf(w[i].w, x.w, s3);
        }
        const float s = warp_sum((s0 + s1) + (s2 + s3));
        if (lane == rr) keep = s;
        if (++rr == n_r) rr = 0;
    }
    if (lane < n_r) Y[(size_t)(r0 + lane) * O + o] = keep + bias;
}

// dX[r, c] = sum_o dY[r, o] W[o, c].  Block = 32 columns (lane = column: W rows are read as 128-byte segments) x a tile of 32 rows
// (one accumulator per row in registers); the O contraction is split over the 8 warps, each staging its slice of dY transposed
// in shared memory so that four rows come back per broadcast 16-byte read; the warps' partial tiles are added in warp order.
constexpr int SK_DG_WARPS = 8;
__global__ void __launch_bounds__(32 * SK_DG_WARPS)
skinny_dgrad_kernel(const float* __restrict__ dY, const float* __restrict__ W, int R, int O, int C, float* __restrict__ dX) {
    __shared__ __align__(16) float buf[SK_DG_WARPS][32 * 36];   // per warp: dY slice [32 o][36] (32 rows + pad), later its [32 r][32 c] partial
    pdl_sync();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 32 + lane;
    const int r0 = blockIdx.y * 32;
    const int per = (O + SK_DG_WARPS - 1) / SK_DG_WARPS;
    const int o_lo = warp * per, o_hi = min(O, o_lo + per);
    float* mine = buf[warp];
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    for (int ob = o_lo; ob < o_hi; ob += 32) {
        __syncwarp();
        // stage dY[r0 .. r0+31][ob .. ob+31] as mine[oo][r] (coalesced along o, transposed into shared memory)
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
            const int r = r0 + rr, o = ob + lane;
            mine[lane * 36 + rr] = (r < R && o < o_hi) ? __ldg(dY + (size_t)r * O + o) : 0.f;
        }
        __syncwarp();
        // eight W rows in flight per lane (rows past the slice multiply staged zeros; their loads are clamped to row O - 1)
#pragma unroll 2
        for (int o8 = 0; o8 < 32; o8 += 8) {
            if (ob + o8 >= o_hi) break;
            float wv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) wv[u] = c < C ? __ldg(W + (size_t)min(ob + o8 + u, O - 1) * C + c) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int r4 = 0; r4 < 8; ++r4) {
                    const float4 d = *reinterpret_cast<const float4*>(mine + (o8 + u) * 36 + r4 * 4);
                    acc[4 * r4] = fmaf(d.x, wv[u], acc[4 * r4]); acc[4 * r4 + 1] = fmaf(d.y, wv[u], acc[4 * r4 + 1]);
                    acc[4 * r4 + 2] = fmaf(d.z, wv[u], acc[4 * r4 + 2]); acc[4 * r4 + 3] = fmaf(d.w, wv[u], acc[4 * r4 + 3]);
                }
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int rr = 0; rr < 32; ++rr) mine[rr * 32 + lane] = acc[rr];
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * 32; e += 32 * SK_DG_WARPS) {
        const int rr = e >> 5, cc = e & 31;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < SK_DG_WARPS; ++w) s += buf[w][rr * 32 + cc];
        const int r = r0 + rr, col = blockIdx.x * 32 + cc;
        if (r < R && col < C) dX[(size_t)r * C + col] = s;
    }
}

// dW[o, c] += sum_r dY[r, o] X[r, c]; db[o] += sum_r dY[r, o].  Thread = column c (X rows are read coalesced), block = 256 columns x
// 16 output rows o: 16 accumulators per thread, so X crosses L2 -> SM once per 16 rows of dW; dY[:, o-tile] is staged through
// shared memory 64 rows at a time and read back as broadcast 16-byte vectors.
constexpr int SK_WG_OT = 16;
__global__ void __launch_bounds__(256)
skinny_wgrad_kernel(const float* __restrict__ dY, const float* __restrict__ X, int R, int O, int C, float* __restrict__ dW, float* __restrict__ db) {
    __shared__ __align__(16) float sdy[64][SK_WG_OT];
    pdl_sync();
    const int o0 = blockIdx.y * SK_WG_OT;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const bool c_ok = c < C;
    float acc[SK_WG_OT];
#pragma unroll
    for (int j = 0; j < SK_WG_OT; ++j) acc[j] = 0.f;
    float sb = 0.f;
    for (int rb = 0; rb < R; rb += 64) {
        __syncthreads();
        for (int e = threadIdx.x; e < 64 * SK_WG_OT; e += blockDim.x) {
            const int rr = e / SK_WG_OT, j = e % SK_WG_OT;
            sdy[rr][j] = (rb + rr < R && o0 + j < O) ? __ldg(dY + (size_t)(rb + rr) * O + o0 + j) : 0.f;
        }
        __syncthreads();
        const int n_r = min(64, R - rb);
        if (blockIdx.x == 0 && threadIdx.x < SK_WG_OT)
            for (int rr = 0; rr < n_r; ++rr) sb += sdy[rr][threadIdx.x];
        int rr = (int)((blockIdx.y * 4u) % (unsigned)n_r);      // o-tiles walk the rows from different starting points (same X lines)
#pragma unroll 8
        for (int it = 0; it < n_r; ++it, rr = (rr + 1 == n_r) ? 0 : rr + 1) {
            const float x = c_ok ? __ldg(X + (size_t)(rb + rr) * C + c) : 0.f;
#pragma unroll
            for (int j4 = 0; j4 < SK_WG_OT / 4; ++j4) {
                const float4 d = *reinterpret_cast<const float4*>(&sdy[rr][4 * j4]);
                acc[4 * j4] = fmaf(d.x, x, acc[4 * j4]); acc[4 * j4 + 1] = fmaf(d.y, x, acc[4 * j4 + 1]);
                acc[4 * j4 + 2] = fmaf(d.z, x, acc[4 * j4 + 2]); acc[4 * j4 + 3] = fmaf(d.w, x, acc[4 * j4 + 3]);
            }
        }
    }
    if (c_ok) {
#pragma unroll
        for (int j = 0; j < SK_WG_OT; ++j)
            if (o0 + j < O) dW[(size_t)(o0 + j) * C + c] += acc[j];
    }
    if (db != nullptr && blockIdx.x == 0 && threadIdx.x < SK_WG_OT && o0 + threadIdx.x < O) db[o0 + threadIdx.x] += sb;
}

}  // namespace mdl

using namespace mdl;

extern "C" {

int mdl_skinny_linear_fwd(const float* X, const float* W, const float* b, int R, int C, int O, float* Y, void* stream) {
    MDL_REQUIRE(C == 2048 || C == 512, "skinny_linear_fwd: C must be 512 or 2048 (got %d)", C);
    if (R == 0) return 0;
    // rows per warp: 8 measured best from R = 32 to R = 325 (more rows per warp = fewer W reads but fewer blocks in flight: slower)
    const int rows = 8;
    dim3 grid((O + 3) / 4, (R + rows - 1) / rows);
    if (C == 2048) MDL_CHECK_CUDA(launch_k(skinny_fwd_kernel<2048>, grid, dim3(128), 0, (cudaStream_t)stream, X, W, b, R, O, rows, Y));
    else MDL_CHECK_CUDA(launch_k(skinny_fwd_kernel<512>, grid, dim3(128), 0, (cudaStream_t)stream, X, W, b, R, O, rows, Y));
    return 0;
}

int mdl_skinny_linear_bwd(const float* dY, const float* X, const float* W, int R, int C, int O,
                          float* dX, float* dW, float* db, void* stream) {
    if (R == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dX != nullptr) {
        dim3 grid((C + 31) / 32, (R + 31) / 32);
        MDL_CHECK_CUDA(launch_k(skinny_dgrad_kernel, grid, dim3(32 * SK_DG_WARPS), 0, st, dY, W, R, O, C, dX));
    }
    if (dW != nullptr) {
        dim3 grid((C + 255) / 256, (O + SK_WG_OT - 1) / SK_WG_OT);
        MDL_CHECK_CUDA(launch_k(skinny_wgrad_kernel, grid, dim3(256), 0, st, dY, X, R, O, C, dW, db));
    }
    return 0;
}

}  // extern "C"
