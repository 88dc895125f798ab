// Shared device helpers for the madeleine_b200 sm_100a kernels: PTX wrappers (mbarrier, TMA, tcgen05/TMEM),
// warp/block reductions, the fp32 -> (bf16 hi, bf16 lo) operand split and a counter-based dropout RNG.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

namespace mdl {

constexpr int kNumSMs = 148;

// ---------------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
const char* get_last_error();

#define MDL_CHECK_CUDA(expr)                                                                     \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            mdl::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

#define MDL_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) {                                           \
            mdl::set_last_error(__VA_ARGS__);                    \
            return 2;                                            \
        }                                                        \
    } while (0)

#define MDL_CHECK_LAUNCH() MDL_CHECK_CUDA(cudaGetLastError())

// ---------------------------------------------------------------------------------------------------
// programmatic dependent launch
// ---------------------------------------------------------------------------------------------------
// Every kernel of the hot path starts with pdl_sync() and is launched through launch_k().  With the launch attribute set
// (MADELEINE_B200_PDL, default on) a kernel's blocks may become resident while the previous kernel of the stream is still
// draining: block scheduling, the launch latency and everything in front of pdl_sync() (barrier initialisation, TMEM allocation,
// tensor-map prefetch in the GEMMs) overlap the predecessor's tail; griddepcontrol.wait then blocks until the predecessor
// grid has completed and its writes are visible, so no kernel touches memory earlier than it would in plain stream order.
// launch_dependents follows the wait immediately: at most two kernels of a stream are in flight, and because every kernel
// of the chain waits before it signals, "my predecessor completed" implies "everything before it completed".  Kernels launched
// without the attribute (torch's, NCCL's, the <<< >>> launches that remain) keep full stream semantics on both sides.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_wait(); pdl_trigger(); }

bool pdl_enabled();        // misc.cu: MADELEINE_B200_PDL != "0", read once; mdl_set_pdl() overrides

template <typename... P, typename... A>
inline cudaError_t launch_k(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);
}

// Function attributes (opt-in dynamic shared memory) belong to a device, not to the process: a host that drives several
// GPUs from one process (nn.DataParallel replicas, one thread per GPU) must set them on each.  One bit per device ordinal;
// setting an attribute twice from racing threads is harmless.
struct PerDeviceOnce {
    std::atomic<unsigned long long> done{0};
    bool needed(unsigned long long& bit) const {
        int dev = 0;
        cudaGetDevice(&dev);
        bit = 1ull << (dev & 63);
        return (done.load(std::memory_order_acquire) & bit) == 0;
    }
    void mark(unsigned long long bit) { done.fetch_or(bit, std::memory_order_release); }
};

// ---------------------------------------------------------------------------------------------------
// small math
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32). `scratch` holds >= 33 floats. Result broadcast to all.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect scratch reuse
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < nw ? scratch[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < nw ? scratch[lane] : -INFINITY;
        t = warp_max(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// fp32 -> bf16 hi + bf16 lo with x ~= hi + lo (|err| <= 2^-17 |x|): the operand format of the 3-pass tcgen05 GEMM.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ float join_bf16(__nv_bfloat16 hi, __nv_bfloat16 lo) {
    return __bfloat162float(hi) + __bfloat162float(lo);
}
__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
// two fp32 -> packed (hi, lo) bf16x2 words: one cvt.rn.bf16x2 per plane instead of scalar converts + packing.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<uint32_t*>(&h);
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
    __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
    lo = *reinterpret_cast<uint32_t*>(&l);
}
// fp16 operand planes (inference): x ~= hi + lo with 11 + 11 mantissa bits (|err| <= 2^-22 |x| for |x| within fp16's normal range;
// the absolute floor is fp16's subnormal spacing 6e-8).  Conversions saturate (+-65504) instead of producing infinities.
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));        // low half <- a, high half <- b
    const __half2 h = *reinterpret_cast<const __half2*>(&hi);
    const float2 hf = __half22float2(h);
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - hf.y), "f"(a - hf.x));
}
// one packed pair of planes values -> two floats
__device__ __forceinline__ float2 f16x2_to_f32(uint32_t packed) { return __half22float2(*reinterpret_cast<const __half2*>(&packed)); }
// PLANE FORMAT: 0 = bf16 hi/lo, 1 = fp16 hi/lo.  `nplanes` / `nsplit` arguments of the C ABI carry it as a flag bit.
constexpr int kPlanesF16 = 0x100;
template <bool F16>
__device__ __forceinline__ void split_planes2(float a, float b, uint32_t& hi, uint32_t& lo) {
    if (F16) split_f16x2(a, b, hi, lo); else split_bf16x2(a, b, hi, lo);
}
template <bool F16>
__device__ __forceinline__ float2 planes2_to_f32(uint32_t packed) {
    if (F16) return f16x2_to_f32(packed);
    return make_float2(__uint_as_float(packed << 16), __uint_as_float(packed & 0xffff0000u));
}
// weights are multiplied by this power of two before the fp16 split (|w| ~ 0.04 would put the lo plane into fp16's
// subnormal range); the GEMM epilogues multiply the accumulator by its inverse
#define MDL_F16_WEIGHT_SCALE 64.0f

__device__ __forceinline__ float bf16_lo_of(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi_of(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }

// Raw SFU instructions (one MUFU each).  __fdividef / exp2f / __expf wrap the same instructions in denormal-range guards
// (FSETP + two predicated FMUL + FSEL per call) that the arguments here never need: denominators are >= 1, and an
// exponential that underflows should flush to zero.
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// one MUFU, |relative error| <= 2^-11: only where the operands are bf16 anyway (the 1-pass modes)
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// GELU(x) = x Phi(x) and its derivative Phi(x) + x phi(x) from ONE exponential: erfc(|x|/sqrt2) by Abramowitz-Stegun
// 7.1.26 (|abs err| <= 1.5e-7, far inside the 1e-4 parity budget) uses exp(-x^2/2), which is also the Gaussian pdf.
//   u = |x| sqrt(log2 e) / sqrt2  (so that exp2(-u^2) = exp(-x^2/2)),  t = 1 / (1 + p |x| / sqrt2),  e = exp2(-u^2)
//   erfc(|x|/sqrt2) = poly(t) e
//   gelu(x) = max(x, 0) - (|x|/2) erfc(|x|/sqrt2) = max(x, 0) - u (poly_g(t) e),   poly_g = poly / (2 sqrt(log2 e)/sqrt2)
//   Phi(x)  = 1/2 + copysign(1/2 - erfc/2, x)
// Branch-free: 13 instructions (2 MUFU) for the value, 15 for the derivative.
#define MDL_GELU_K 0.84932180028801907f               /* sqrt(log2 e) / sqrt2 */
#define MDL_GELU_P 0.27273617448364717f               /* 0.3275911 / sqrt(log2 e): p |x| / sqrt2 = MDL_GELU_P * u */
__device__ __forceinline__ float gelu_erf(float x) {
    const float u = fabsf(x) * MDL_GELU_K;
    const float t = rcp_approx(fmaf(MDL_GELU_P, u, 1.f));
    const float e = ex2_approx(-u * u);
    // A-S coefficients times 1 / (2 K)
    const float poly = t * (0.1500194578f + t * (-0.1674846542f + t * (0.8367933924f + t * (-0.8554778804f + t * 0.624854695f))));
    return fmaf(-u, poly * e, fmaxf(x, 0.f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float u = fabsf(x) * MDL_GELU_K;
    const float t = rcp_approx(fmaf(MDL_GELU_P, u, 1.f));
    const float e = ex2_approx(-u * u);
    // A-S coefficients times 1/2: hw = erfc(|x|/sqrt2) / 2
    const float hw = (t * (0.127414796f + t * (-0.142248368f + t * (0.7107068705f + t * (-0.7265760135f + t * 0.5307027145f))))) * e;
    const float phi_cdf = 0.5f + copysignf(0.5f - hw, x);
    return fmaf(x * e, 0.39894228040143268f, phi_cdf);
}
__device__ __forceinline__ void gelu_and_grad(float x, float& g, float& dg) { g = gelu_erf(x); dg = gelu_erf_grad(x); }
// sigmoid / tanh from one ex2 + one rcp each; abs error ~1e-7, saturate correctly at +-inf.
__device__ __forceinline__ float sigmoid_acc(float x) { return rcp_approx(1.f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_acc(float x) { return fmaf(2.f, sigmoid_acc(2.f * x), -1.f); }

// Stateless counter-based RNG for dropout masks: the same (seed, stream, index) gives the same bit in fwd and bwd.
// One 64-bit hash serves FOUR consecutive elements (16 bits each), so kernels that touch float4 / 4-column groups
// pay one hash per vector.  `idx4` is the element index divided by 4.
// Philox-2x32-style mixing (5 rounds of 32x32->64 multiply + xor): 64 well-mixed bits from ~15 integer instructions.
__device__ __forceinline__ uint64_t hash_u64(uint64_t seed, uint32_t stream, uint64_t idx4) {
    uint32_t c0 = (uint32_t)idx4, c1 = (uint32_t)(idx4 >> 32) ^ (stream * 0x9E3779B9u);
    uint32_t k = (uint32_t)seed ^ (uint32_t)(seed >> 32) * 0x85EBCA6Bu;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const uint64_t p = (uint64_t)c0 * 0xD2511F53u;
        c0 = (uint32_t)(p >> 32) ^ k ^ c1;
        c1 = (uint32_t)p;
        k += 0x9E3779B9u;
    }
    return ((uint64_t)c0 << 32) | c1;
}
// multiplicative masks for elements 4*idx4 .. 4*idx4+3: 0 (dropped) or 1/(1-p) (kept); p == 0 -> all 1.
// Element i is kept when the i-th 16-bit field of the hash is >= p * 65536.  The fields are compared in place: a field in
// the top half of a 32-bit word compares as (word >= thresh << 16), one in the bottom half after a 16-bit left shift.
struct DropCfg {
    uint32_t thresh_hi;      // (uint32_t)(p * 65536) << 16
    float keep;              // 1 / (1 - p)
    bool on;
};
__device__ __forceinline__ DropCfg make_drop_cfg(float p) {
    DropCfg c;
    c.on = p > 0.f;
    c.thresh_hi = ((uint32_t)(p * 65536.f)) << 16;
    c.keep = c.on ? __fdividef(1.f, 1.f - p) : 1.f;
    return c;
}
__device__ __forceinline__ void dropout_scale4(const DropCfg& c, uint64_t seed, uint32_t stream, uint64_t idx4, float (&m)[4]) {
    if (!c.on) { m[0] = m[1] = m[2] = m[3] = 1.f; return; }
    const uint64_t h = hash_u64(seed, stream, idx4);
    const uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
    m[0] = (lo << 16) >= c.thresh_hi ? c.keep : 0.f;
    m[1] = lo >= c.thresh_hi ? c.keep : 0.f;
    m[2] = (hi << 16) >= c.thresh_hi ? c.keep : 0.f;
    m[3] = hi >= c.thresh_hi ? c.keep : 0.f;
}
// Two mask vectors (elements 4*idx4 .. +3 of stream `stream` and of stream `stream + 1`) from ONE hash, 8-bit fields: exact
// whenever p * 256 is an integer (the gates' nn.Dropout(0.25), abmil.py:33-35).  Used where only one kernel ever needs
// the masks (the gated-attention epilogue; its backward reads them off the stored gates).
__device__ __forceinline__ bool drop_p_is_8bit(float p) { const float t = p * 256.f; return t == floorf(t); }
__device__ __forceinline__ void dropout_scale4x2_8bit(uint32_t thresh24, float keep, uint64_t seed, uint32_t stream, uint64_t idx4,
                                                      float (&ma)[4], float (&mb)[4]) {
    const uint64_t h = hash_u64(seed, stream, idx4);
    const uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
    ma[0] = (lo << 24) >= thresh24 ? keep : 0.f;
    ma[1] = (lo << 16) >= thresh24 ? keep : 0.f;
    ma[2] = (lo << 8) >= thresh24 ? keep : 0.f;
    ma[3] = lo >= thresh24 ? keep : 0.f;
    mb[0] = (hi << 24) >= thresh24 ? keep : 0.f;
    mb[1] = (hi << 16) >= thresh24 ? keep : 0.f;
    mb[2] = (hi << 8) >= thresh24 ? keep : 0.f;
    mb[3] = hi >= thresh24 ? keep : 0.f;
}
__device__ __forceinline__ void dropout_scale4(float p, uint64_t seed, uint32_t stream, uint64_t idx4, float (&m)[4]) {
    dropout_scale4(make_drop_cfg(p), seed, stream, idx4, m);
}

// ---------------------------------------------------------------------------------------------------
// PTX: shared-memory address, mbarrier, TMA, tcgen05
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {  // ~2 s at 2 GHz
            printf("mdl: mbarrier wait timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
}

__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 3D tiled TMA load global -> shared, completion on an mbarrier.
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate, issued by ONE thread for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread i gets lane i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// TMEM -> registers: 32 lanes x 16 consecutive fp32 columns (half of the above; lets an epilogue software-pipeline its loads).
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   bits [0,14)  start address >> 4      bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4 bits [46,48) version = 1        bits [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Same, 64-byte swizzle (layout type 4): rows of 64 B, 8-row atoms 512 B apart.
__device__ __forceinline__ uint64_t make_umma_desc_sw64(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D.
// kind::f16 operand formats live in bits [7,10) (A) and [10,13) (B): 1 = BF16, 0 = F16
__host__ __device__ constexpr uint32_t idesc_as_f16(uint32_t idesc_bf16) { return idesc_bf16 & ~((7u << 7) | (7u << 10)); }
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

}  // namespace mdl
