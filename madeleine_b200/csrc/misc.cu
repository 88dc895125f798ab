// Error plumbing, version queries and the CUDA-core cross-check GEMMs used by the tests.
#include "common.cuh"
#include "madeleine_b200.h"
#include <stdarg.h>
#include <stdlib.h>

namespace mdl {

static thread_local char g_err[1024] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_last_error() { return g_err; }

static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
    int v = g_pdl.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("MADELEINE_B200_PDL");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
        g_pdl.store(v, std::memory_order_relaxed);
    }
    return v != 0;
}
void set_pdl(int on) { g_pdl.store(on ? 1 : 0, std::memory_order_relaxed); }

__device__ __forceinline__ float plane_val(const __nv_bfloat16* p, long long off, long long plane_stride, int plane) {
    return __bfloat162float(p[off + (long long)plane * plane_stride]);
}

// out[m,n] = sum_k A[m,k] B[n,k] with the same pass structure as the tensor-core kernel (hi*hi + hi*lo + lo*hi).
__global__ void gemm_nt_simt_kernel(const __nv_bfloat16* __restrict__ a, long long lda, long long aps,
                                    const __nv_bfloat16* __restrict__ b, long long ldb, long long bps,
                                    float* __restrict__ out, long long ldc, int M, int N, int K, int nsplit) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (n >= N || m >= M) return;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
        const float ah = plane_val(a, m * lda + k, aps, 0), bh = plane_val(b, n * ldb + k, bps, 0);
        acc = fmaf(ah, bh, acc);
        if (nsplit == 3) {
            const float al = plane_val(a, m * lda + k, aps, 1), bl = plane_val(b, n * ldb + k, bps, 1);
            acc = fmaf(ah, bl, acc);
            acc = fmaf(al, bh, acc);
        }
    }
    out[m * ldc + n] = acc;
}

// out[i,j] += sum_t A[t,i] B[t,j]
__global__ void gemm_tn_simt_kernel(const __nv_bfloat16* __restrict__ a, long long lda, long long aps,
                                    const __nv_bfloat16* __restrict__ b, long long ldb, long long bps, long long tokens,
                                    float* __restrict__ out, long long ldc, int M, int N, int nsplit) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= N || i >= M) return;
    float acc = 0.f;
    for (long long t = 0; t < tokens; ++t) {
        const float ah = plane_val(a, t * lda + i, aps, 0), bh = plane_val(b, t * ldb + j, bps, 0);
        acc = fmaf(ah, bh, acc);
        if (nsplit == 3) {
            const float al = plane_val(a, t * lda + i, aps, 1), bl = plane_val(b, t * ldb + j, bps, 1);
            acc = fmaf(ah, bl, acc);
            acc = fmaf(al, bh, acc);
        }
    }
    out[i * ldc + j] += acc;
}

}  // namespace mdl

using namespace mdl;

extern "C" {

const char* mdl_last_error(void) { return get_last_error(); }
int mdl_version(void) { return 100; }
int mdl_built_arch(void) { return 100; }
int mdl_set_pdl(int on) {
    const int before = pdl_enabled() ? 1 : 0;
    set_pdl(on);
    return before;
}

int mdl_gemm_nt_simt(const void* a_planes, long long lda, long long a_plane_stride, const void* b_planes, long long ldb,
                     long long b_plane_stride, float* out, long long ldc, int M, int N, int K, int nsplit, void* stream) {
    dim3 grid((N + 127) / 128, M);
    gemm_nt_simt_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a_planes, lda, a_plane_stride,
                                                               (const __nv_bfloat16*)b_planes, ldb, b_plane_stride, out, ldc, M, N, K, nsplit);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_gemm_tn_simt(const void* a_planes, long long lda, long long a_plane_stride, const void* b_planes, long long ldb,
                     long long b_plane_stride, long long tokens, float* out, long long ldc, int M, int N, int nsplit, void* stream) {
    dim3 grid((N + 127) / 128, M);
    gemm_tn_simt_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a_planes, lda, a_plane_stride,
                                                               (const __nv_bfloat16*)b_planes, ldb, b_plane_stride, tokens, out, ldc, M, N, nsplit);
    MDL_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
