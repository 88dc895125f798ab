// Shared definitions of the Graph-OT kernels (got.cu: problems of n <= 96 tokens entirely in shared memory;
// got_big.cu: 96 < n <= 256, matrices in an L2-resident global workspace).
#pragma once
#include "common.cuh"

namespace mdl {

constexpr int GOT_NMAX = 96;          // largest problem the shared-memory kernels (got.cu) take
constexpr int GOT_BIG_NMAX = 256;     // largest problem at all: GOT(..., subsample=256) caps n at 256 (trainer.py:45)
constexpr int GOT_THREADS = 512;
constexpr int GOT_WARPS = GOT_THREADS / 32;
constexpr int WD_ITERS = 30;
constexpr int GW_OUTER = 5;
constexpr int GW_INNER = 20;
constexpr float WD_BETA = 0.5f;
constexpr float GW_BETA = 0.1f;
constexpr float THR_BETA = 0.1f;
constexpr int MAX_ITERS = WD_ITERS;  // lu/lw capacity (>= GW_INNER)

struct GotLayout {
    int m, n, D;
    size_t nn;
    // float offsets inside one problem's slab
    size_t raw0, raws, rawt;          // raw costs
    size_t g0, gs, gt;                // gradient accumulators (w.r.t. thresholded costs, then raw costs)
    size_t cg;                        // 6 x nn
    size_t gamma;                     // 5 x nn (gamma_1..gamma_5)
    size_t lulw;                      // GW_OUTER x 2 x (GW_INNER+1) x n
    size_t ext;                       // 6 floats extrema + 6 ints argidx + 3 floats dthr + pad
    // big path only (got_big.cu, n > GOT_NMAX or forced): the five working matrices (leading dimension n|1), the normalised tokens and norms
    size_t mats, vn, qn, norms;
    size_t per_item;
    size_t header;                    // global header floats (6 extrema + 12 ints)
    __host__ __device__ GotLayout(int m_, int n_, int D_, bool big) : m(m_), n(n_), D(D_) {
        nn = (size_t)n * n;
        size_t o = 0;
        raw0 = o; o += nn; raws = o; o += nn; rawt = o; o += nn;
        g0 = o; o += nn; gs = o; o += nn; gt = o; o += nn;
        cg = o; o += 6 * nn;
        gamma = o; o += 5 * nn;
        lulw = o; o += (size_t)GW_OUTER * 2 * (GW_INNER + 1) * n;
        ext = o; o += 16;
        mats = vn = qn = norms = 0;
        if (big) {
            o = (o + 31) / 32 * 32;
            mats = o; o += (size_t)5 * n * (n | 1);
            vn = o; o += (size_t)n * D;
            qn = o; o += (size_t)n * D;
            norms = o; o += (size_t)2 * n;
        }
        per_item = (o + 31) / 32 * 32;
        header = 32;
    }
    __host__ __device__ size_t total_floats() const { return header + per_item * (size_t)m; }
};


// got_big.cu
int got_big_extrema(const float* v, const float* q, int m, int n, int D, void* workspace, float* extrema, cudaStream_t st);
int got_big_main(int m, int n, int D, void* workspace, const float* extrema, float* wd, float* gwd, float* dthr_local, cudaStream_t st);
int got_big_finish(const float* v, const float* q, int m, int n, int D, void* workspace, const float* extrema, const float* dthr_global,
                   const float* wd, const float* gwd, float* loss, float* dv, float* dq, cudaStream_t st);
// small single-thread-block kernels shared by both paths (defined in got.cu)
int got_launch_extrema(const GotLayout& lay, float* ws, float* extrema, cudaStream_t st);
int got_launch_dthr(const GotLayout& lay, const float* ws, float* dthr_out, cudaStream_t st);

}  // namespace mdl
