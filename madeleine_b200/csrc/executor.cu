// Native step executor: the whole encoder forward / backward launch sequence behind ONE C call each.
//
// The reference's training step (madeleine/utils/trainer.py:108-131) spends its host time in Python — ~45 dispatches per
// encoder pass.  Here the sequence  split -> 3 x (GEMM, LayerNorm+GELU) -> gated-attention GEMM -> pooling -> projector /
// token_projector  (Model.py:346-451, 138-146) and its reverse are issued natively on the caller's stream over a
// caller-provided arena: no allocation, no synchronisation, no per-launch host round trip.  The arena layout is computed by
// running the same code with a null base pointer ("dry run"), so the size query and the real pass cannot disagree.
#include "common.cuh"
#include "madeleine_b200.h"
#include <mutex>
#include <vector>

namespace mdl {

// ------------------------------------------------------------------------------------------------------------------
// small kernels that only the executor needs
// ------------------------------------------------------------------------------------------------------------------
__global__ void permute_f32_kernel(const float* __restrict__ src, const int* __restrict__ pos, const int* __restrict__ dst,
                                   long long n, float* __restrict__ out) {
    pdl_sync();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[dst[i]] = __ldg(src + pos[i]);
}

// y += x over n elements (fp32 or bf16 storage): the rare "gradient through the pre-attention features AND token_projector" case
template <typename T>
__global__ void add_inplace_kernel(T* __restrict__ y, const T* __restrict__ x, long long n) {
    pdl_sync();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = (T)((float)y[i] + (float)x[i]);
}

// dst[r, c] += src[c, r]   (dst [rows, cols] row-major, src [cols, rows]); first-layer wgrad of widths that are 128- but not 256-aligned
__global__ void add_transposed_kernel(float* __restrict__ dst, const float* __restrict__ src, int rows, int cols) {
    pdl_sync();
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < cols && r < rows) ? src[(long long)c * rows + r] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) dst[(long long)r * cols + c] += tile[threadIdx.x][i];
    }
}

// G[i, j] = sum_n E[n, i] * E[n, j] accumulated in fp64 (E fp32 [n, D]); one 32 x 32 tile of G per block, upper triangle
// only (mirrored on the way out).  The singular values the rank metric needs (smooth_rank_measure, utils.py:180-201) are
// the square roots of this matrix's eigenvalues, so the [n, 512] embedding matrix never leaves the device.
__global__ void __launch_bounds__(1024)
gram_f64_kernel(const float* __restrict__ E, long long n, int D, double* __restrict__ G) {
    pdl_sync();
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (tj < ti) return;
    __shared__ float a[32][33], b[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    double acc = 0.0;
    for (long long r0 = 0; r0 < n; r0 += 32) {
        const long long r = r0 + ty;
        const int ci = ti * 32 + tx, cj = tj * 32 + tx;
        a[ty][tx] = (r < n && ci < D) ? __ldg(E + r * D + ci) : 0.f;
        b[ty][tx] = (r < n && cj < D) ? __ldg(E + r * D + cj) : 0.f;
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) acc = fma((double)a[k][ty], (double)b[k][tx], acc);
        __syncthreads();
    }
    const int i = ti * 32 + ty, j = tj * 32 + tx;
    if (i < D && j < D) {
        G[(long long)i * D + j] = acc;
        G[(long long)j * D + i] = acc;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// process-wide launch profiler (bench.py's live roofline); forward runs on the caller's thread, backward on autograd's
// ------------------------------------------------------------------------------------------------------------------
struct ProfRec { int tag; cudaEvent_t e0, e1; };
struct Profiler {
    unsigned mask = 0;        // bit per mdl_prof_tag
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
};
static Profiler g_prof;
static std::mutex g_prof_mu;
static std::atomic<long long> g_launches{0};

struct Scope {
    cudaStream_t st; int idx = -1;
    Scope(int tag, cudaStream_t s) : st(s) {
        if (g_prof.mask & (1u << tag)) {
            std::lock_guard<std::mutex> lk(g_prof_mu);
            ProfRec r{tag, g_prof.get(), g_prof.get()};
            cudaEventRecord(r.e0, st);
            g_prof.recs.push_back(r);
            idx = (int)g_prof.recs.size() - 1;
        }
    }
    ~Scope() {
        if (idx >= 0) {
            std::lock_guard<std::mutex> lk(g_prof_mu);
            if (idx < (int)g_prof.recs.size()) cudaEventRecord(g_prof.recs[idx].e1, st);
        }
    }
};

#define RUN(tag, call)                               \
    do {                                             \
        if (!dry) {                                  \
            ++g_launches;                            \
            Scope _s(tag, st);                       \
            int _rc = (call);                        \
            if (_rc != 0) return _rc;                \
        }                                            \
    } while (0)

// ------------------------------------------------------------------------------------------------------------------
// arena
// ------------------------------------------------------------------------------------------------------------------
struct Arena {
    char* base; size_t off = 0, peak = 0;
    explicit Arena(void* b) : base((char*)b) {}
    void* raw(size_t bytes) {
        off = (off + 255) & ~(size_t)255;
        void* p = base ? base + off : nullptr;
        off += bytes;
        if (off > peak) peak = off;
        return p;
    }
    template <typename T> T* get(size_t n) { return (T*)raw(n * sizeof(T)); }
    size_t mark() const { return off; }
    void release(size_t m) { off = m; }
};

constexpr int HID = 512, GATE = 512, TOK = 128;
constexpr float LN_EPS = 1e-5f;

struct Desc {
    const long long* ip; const double* fp; void* const* pp;
    long long M; int R, d_in, d_in_total, se, H, C;
    int nsplit_f, npl_f, nsplit_b, npl_b, act_bf16, act, keep, want_tok, want_proj, want_ref;
    long long n_view_tok; int R2; long long n_sel; unsigned long long seed; int phase;
    float p_pre, p_gate;
    cudaStream_t st;
    size_t act_size;           // bytes per element of z / dgrad outputs
    explicit Desc(const long long* ip_, const double* fp_, void* const* pp_) : ip(ip_), fp(fp_), pp(pp_) {
        M = ip[MDL_ENC_I_M]; R = (int)ip[MDL_ENC_I_R]; d_in = (int)ip[MDL_ENC_I_D_IN]; d_in_total = (int)ip[MDL_ENC_I_D_IN_TOTAL];
        se = (int)ip[MDL_ENC_I_SE_DIM]; H = (int)ip[MDL_ENC_I_N_HEADS]; C = HID * H;
        nsplit_f = (int)ip[MDL_ENC_I_NSPLIT_FWD]; npl_f = (int)ip[MDL_ENC_I_NPL_FWD];
        nsplit_b = (int)ip[MDL_ENC_I_NSPLIT_BWD]; npl_b = (int)ip[MDL_ENC_I_NPL_BWD];
        act_bf16 = (int)ip[MDL_ENC_I_ACT_BF16]; act = (int)ip[MDL_ENC_I_ACTIVATION]; keep = (int)ip[MDL_ENC_I_KEEP];
        want_tok = (int)ip[MDL_ENC_I_WANT_TOKENS]; want_proj = (int)ip[MDL_ENC_I_WANT_PROJECTOR]; want_ref = (int)ip[MDL_ENC_I_WANT_REF];
        n_view_tok = ip[MDL_ENC_I_N_VIEW_TOK]; R2 = (int)ip[MDL_ENC_I_R2]; n_sel = ip[MDL_ENC_I_N_SEL];
        seed = (unsigned long long)ip[MDL_ENC_I_SEED]; phase = (int)ip[MDL_ENC_I_PHASE];
        p_pre = fp ? (float)fp[MDL_ENC_F_P_PRE] : 0.f; p_gate = fp ? (float)fp[MDL_ENC_F_P_GATE] : 0.f;
        st = pp ? (cudaStream_t)pp[MDL_ENC_P_STREAM] : nullptr;
        act_size = act_bf16 ? 2 : 4;
    }
    template <typename T> T* ptr(int which) const { return pp ? (T*)pp[which] : nullptr; }
    const __nv_bfloat16* wbf(int which) const {
        const __nv_bfloat16* b = ptr<const __nv_bfloat16>(MDL_ENC_P_WBF);
        return b ? b + ip[which] : nullptr;
    }
    const float* wf(int which) const {
        const float* b = ptr<const float>(MDL_ENC_P_WF32);
        return b ? b + ip[which] : nullptr;
    }
    bool has_views() const { return n_view_tok > 0; }
    bool has_window() const { return n_sel > 0; }
};

// everything the forward pass leaves in the arena (pointers are null in a dry run)
struct FwdState {
    int* row2bag; float* rowbias; __nv_bfloat16* xp;
    void *z1, *z2, *z3; float *mean1, *rstd1, *mean2, *rstd2, *mean3, *rstd3;
    __nv_bfloat16 *h1, *h2, *h3; __half *gate_a, *gate_b; float *attn_p, *attn_p2; __nv_bfloat16* h3_sel;
    void *pool_ws, *pool_ws2; int tsplit, tsplit2;
};

static void layout_fwd(Arena& a, const Desc& d, FwdState& s) {
    const long long M = d.M; const int C = d.C, npl = d.npl_f;
    s.row2bag = a.get<int>(M);
    s.rowbias = d.se > 0 ? a.get<float>((size_t)d.R * HID) : nullptr;
    s.xp = a.get<__nv_bfloat16>((size_t)npl * M * d.d_in);
    s.z1 = a.raw((size_t)M * HID * d.act_size); s.mean1 = a.get<float>(M); s.rstd1 = a.get<float>(M);
    s.h1 = a.get<__nv_bfloat16>((size_t)npl * M * HID);
    s.z2 = a.raw((size_t)M * HID * d.act_size); s.mean2 = a.get<float>(M); s.rstd2 = a.get<float>(M);
    s.h2 = a.get<__nv_bfloat16>((size_t)npl * M * HID);
    s.z3 = a.raw((size_t)M * C * d.act_size); s.mean3 = a.get<float>(M); s.rstd3 = a.get<float>(M);
    s.h3 = a.get<__nv_bfloat16>((size_t)npl * M * C);
    const size_t gate_rows = ((size_t)M + 31) / 32 * 32;     // tiled scratch layout: whole 32-row blocks (gemm_common.cuh::gate_tile_offset)
    s.gate_a = d.keep ? a.get<__half>(gate_rows * d.H * GATE) : nullptr;
    s.gate_b = d.keep ? a.get<__half>(gate_rows * d.H * GATE) : nullptr;
    s.attn_p = a.get<float>((size_t)M * d.H);
    s.tsplit = mdl_pool_tsplit(d.R, M, d.H, HID);
    s.pool_ws = s.tsplit > 1 ? a.raw((size_t)mdl_pool_workspace_bytes(d.R, d.H, HID, s.tsplit)) : nullptr;
    s.attn_p2 = nullptr; s.pool_ws2 = nullptr; s.tsplit2 = 1;
    if (d.has_views()) {
        s.attn_p2 = a.get<float>((size_t)M * d.H);
        s.tsplit2 = mdl_pool_tsplit(d.R2, d.n_view_tok, d.H, HID);
        s.pool_ws2 = s.tsplit2 > 1 ? a.raw((size_t)mdl_pool_workspace_bytes(d.R2, d.H, HID, s.tsplit2)) : nullptr;
    }
    s.h3_sel = (d.want_tok && d.has_window()) ? a.get<__nv_bfloat16>((size_t)npl * d.n_sel * C) : nullptr;
}

static int run_fwd(const Desc& d, bool dry, size_t* bytes_out) {
    Arena a(dry ? nullptr : d.ptr<void>(MDL_ENC_P_ARENA));
    FwdState s;
    layout_fwd(a, d, s);
    if (bytes_out) *bytes_out = a.peak;
    if (dry) return 0;
    cudaStream_t st = d.st;
    void* stv = (void*)st;
    const long long M = d.M; const int C = d.C, H = d.H, R = d.R;
    // MDL_ENC_I_PLANES_F16: the operand planes (activations and the packed weights handed in) are fp16 hi/lo — the
    // inference format of the fp32-grade mode; the flag rides on the nplanes / nsplit arguments of the entry points
    const int f16 = d.ip[MDL_ENC_I_PLANES_F16] ? kPlanesF16 : 0;
    const int npl = d.npl_f | f16, nsplit = d.nsplit_f | f16;
    const long long bfn = d.ip[MDL_ENC_I_BF_NUMEL];
    const float* x = d.ptr<const float>(MDL_ENC_P_X);
    const int* cu = d.ptr<const int>(MDL_ENC_P_CU);
    const float* master = d.ptr<const float>(MDL_ENC_P_MASTER);
    float* logits = d.ptr<float>(MDL_ENC_P_LOGITS);
    float* slide_hm = d.ptr<float>(MDL_ENC_P_SLIDE_HM);
    MDL_REQUIRE(x && cu && logits && slide_hm && d.ptr<void>(MDL_ENC_P_WBF) && d.ptr<void>(MDL_ENC_P_WF32), "mdl_encoder_fwd: null input");
    MDL_REQUIRE(!(f16 && (d.keep || d.act_bf16 || d.nsplit_f != 3)), "mdl_encoder_fwd: fp16 planes are the fp32-grade INFERENCE format (keep = 0, 3-pass)");

    RUN(MDL_PROF_OTHER, mdl_row2bag(cu, R, s.row2bag, M, stv));
    if (d.se > 0) {
        MDL_REQUIRE(master && d.ptr<void>(MDL_ENC_P_CODES), "mdl_encoder_fwd: stain encodings need MASTER and CODES");
        RUN(MDL_PROF_OTHER, mdl_stain_rowbias(master + d.ip[MDL_ENC_I_MASTER_EMB], d.ptr<const int>(MDL_ENC_P_CODES),
                                              master + d.ip[MDL_ENC_I_MASTER_PRE0W], d.d_in_total, d.d_in, d.se, HID, R, s.rowbias, stv));
    }
    RUN(MDL_PROF_OTHER, mdl_split_planes(x, M, d.d_in, d.d_in, s.xp, M * d.d_in, npl, stv));
    // layer 1..3: Linear (+ per-bag stain bias) -> LayerNorm -> GELU -> Dropout   (Model.py:350-362)
    RUN(MDL_PROF_GEMM_NT, mdl_gemm_nt(s.xp, M, d.d_in, d.d_in, M * d.d_in, d.wbf(MDL_ENC_I_BF_W1), HID, d.d_in, d.d_in, bfn, s.z1, HID,
                                      (int)M, HID, d.d_in, nsplit, 0, 0, d.wf(MDL_ENC_I_F32_B1), s.rowbias, s.row2bag, d.act_bf16, stv));
    RUN(MDL_PROF_LN_FWD, mdl_ln_gelu_fwd(s.z1, M, HID, d.wf(MDL_ENC_I_F32_G1), d.wf(MDL_ENC_I_F32_BE1), LN_EPS, d.p_pre, d.seed, 1, s.h1,
                                         M * HID, npl, s.mean1, s.rstd1, d.act_bf16, stv));
    RUN(MDL_PROF_GEMM_NT, mdl_gemm_nt(s.h1, M, HID, HID, M * HID, d.wbf(MDL_ENC_I_BF_W2), HID, HID, HID, bfn, s.z2, HID, (int)M, HID, HID,
                                      nsplit, 0, 0, d.wf(MDL_ENC_I_F32_B2), nullptr, nullptr, d.act_bf16, stv));
    RUN(MDL_PROF_LN_FWD, mdl_ln_gelu_fwd(s.z2, M, HID, d.wf(MDL_ENC_I_F32_G2), d.wf(MDL_ENC_I_F32_BE2), LN_EPS, d.p_pre, d.seed, 2, s.h2,
                                         M * HID, npl, s.mean2, s.rstd2, d.act_bf16, stv));
    RUN(MDL_PROF_GEMM_NT, mdl_gemm_nt(s.h2, M, HID, HID, M * HID, d.wbf(MDL_ENC_I_BF_W3), C, HID, HID, bfn, s.z3, C, (int)M, C, HID,
                                      nsplit, 0, 0, d.wf(MDL_ENC_I_F32_B3), nullptr, nullptr, d.act_bf16, stv));
    RUN(MDL_PROF_LN_FWD, mdl_ln_gelu_fwd(s.z3, M, C, d.wf(MDL_ENC_I_F32_G3), d.wf(MDL_ENC_I_F32_BE3), LN_EPS, d.p_pre, d.seed, 3, s.h3,
                                         M * C, npl, s.mean3, s.rstd3, d.act_bf16, stv));
    // gated attention, all heads (abmil.py:49-52)
    RUN(MDL_PROF_GEMM_GATED, mdl_gemm_gated(s.h3, M, C, C, M * C, d.wbf(MDL_ENC_I_BF_WAB), bfn, (int)M, H, nsplit, d.wf(MDL_ENC_I_F32_BA),
                                            d.wf(MDL_ENC_I_F32_BB), d.wf(MDL_ENC_I_F32_WC), d.wf(MDL_ENC_I_F32_BC), logits, s.gate_a, s.gate_b,
                                            d.p_gate, d.seed, stv));
    // attention pooling (abmil.py:54-63, Model.py:413-417): whole view with the configured activation
    RUN(MDL_PROF_POOL_WEIGHTS, mdl_pool_weights(logits, cu, nullptr, R, H, HID, s.attn_p, d.act, s.tsplit, s.pool_ws, stv));
    RUN(MDL_PROF_POOL_FWD, mdl_pool_fwd(s.h3, M * C, npl, s.attn_p, cu, nullptr, R, M, H, HID, slide_hm, s.tsplit, s.pool_ws, stv));
    int n_slide = R;
    if (d.has_views()) {
        // the two half views are always re-normalised with a softmax over the raw logits (Model.py:436)
        const int* tok_idx = d.ptr<const int>(MDL_ENC_P_VIEW_TOK_IDX);
        const int* cu2 = d.ptr<const int>(MDL_ENC_P_VIEW_CU);
        MDL_REQUIRE(tok_idx && cu2, "mdl_encoder_fwd: n_views=3 needs the view index lists");
        if (d.keep) MDL_CHECK_CUDA(cudaMemsetAsync(s.attn_p2, 0, (size_t)M * H * sizeof(float), st));
        RUN(MDL_PROF_POOL_WEIGHTS, mdl_pool_weights(logits, cu2, tok_idx, d.R2, H, HID, s.attn_p2, 0, s.tsplit2, s.pool_ws2, stv));
        RUN(MDL_PROF_POOL_FWD, mdl_pool_fwd(s.h3, M * C, npl, s.attn_p2, cu2, tok_idx, d.R2, d.n_view_tok, H, HID, slide_hm + (size_t)R * C,
                                            s.tsplit2, s.pool_ws2, stv));
        n_slide += d.R2;
    }
    if (d.want_proj) {
        float* slide = d.ptr<float>(MDL_ENC_P_SLIDE);
        MDL_REQUIRE(slide, "mdl_encoder_fwd: SLIDE output missing");
        RUN(MDL_PROF_SKINNY, mdl_skinny_linear_fwd(slide_hm, d.wf(MDL_ENC_I_F32_WP), d.wf(MDL_ENC_I_F32_BP), n_slide, C, HID, slide, stv));
    }
    if (d.want_tok) {
        float* tokens = d.ptr<float>(MDL_ENC_P_TOKENS);
        MDL_REQUIRE(tokens, "mdl_encoder_fwd: TOKENS output missing");
        if (d.has_window()) {
            RUN(MDL_PROF_OTHER, mdl_gather_rows_planes(s.h3, M * C, npl, C, d.ptr<const int>(MDL_ENC_P_TOKEN_ROWS), d.n_sel, s.h3_sel,
                                                       d.n_sel * C, stv));
            RUN(MDL_PROF_GEMM_NT, mdl_gemm_nt(s.h3_sel, d.n_sel, C, C, d.n_sel * C, d.wbf(MDL_ENC_I_BF_TP), TOK, C, C, bfn, tokens, TOK,
                                              (int)d.n_sel, TOK, C, nsplit, 0, 0, d.wf(MDL_ENC_I_F32_BTP), nullptr, nullptr, 0, stv));
        } else {
            RUN(MDL_PROF_GEMM_NT, mdl_gemm_nt(s.h3, M, C, C, M * C, d.wbf(MDL_ENC_I_BF_TP), TOK, C, C, bfn, tokens, TOK, (int)M, TOK, C,
                                              nsplit, 0, 0, d.wf(MDL_ENC_I_F32_BTP), nullptr, nullptr, 0, stv));
        }
    }
    if (d.want_ref) {
        float* ref = d.ptr<float>(MDL_ENC_P_REF);
        MDL_REQUIRE(ref, "mdl_encoder_fwd: REF output missing");
        RUN(MDL_PROF_OTHER, mdl_planes_to_ref_order(s.h3, M * C, npl, M, H, HID, ref, stv));
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------------
static int run_bwd(const Desc& d, bool dry, size_t* bytes_out) {
    // forward state: same layout function over the forward arena
    Arena fa(dry ? nullptr : d.ptr<void>(MDL_ENC_P_ARENA));
    FwdState s;
    layout_fwd(fa, d, s);
    Arena a(dry ? nullptr : d.ptr<void>(MDL_ENC_P_BWD_ARENA));
    cudaStream_t st = d.st;
    void* stv = (void*)st;
    const long long M = d.M; const int C = d.C, H = d.H, R = d.R, npl = d.npl_b, nsplit = d.nsplit_b, fnpl = d.npl_f;
    const long long bfn = d.ip[MDL_ENC_I_BF_NUMEL];
    const int n_slide = R + (d.has_views() ? d.R2 : 0);
    const int phase = d.phase;
    const bool first = phase == 0 || phase == 1, second = phase == 0 || phase == 2;

    // persistent across the two phases: packed gradient buffer, per-bag dz1 sums, and the layer-2 input gradient
    float* gp = a.get<float>((size_t)d.ip[MDL_ENC_I_GR_NUMEL]);
    void* dh2 = a.raw((size_t)M * HID * d.act_size);
    auto g = [&](int which) -> float* { return gp ? gp + d.ip[which] : nullptr; };
    const size_t base_mark = a.mark();

    float* gmaster = d.ptr<float>(MDL_ENC_P_GMASTER);
    const float* master = d.ptr<const float>(MDL_ENC_P_MASTER);
    const int* cu = d.ptr<const int>(MDL_ENC_P_CU);
    if (!dry) MDL_REQUIRE(gmaster && cu && d.ptr<void>(MDL_ENC_P_LOGITS) && d.ptr<void>(MDL_ENC_P_SLIDE_HM), "mdl_encoder_bwd: null input");

    if (first) {
        if (!dry) {
            MDL_CHECK_CUDA(cudaMemsetAsync(gp, 0, (size_t)d.ip[MDL_ENC_I_GR_NUMEL] * sizeof(float), st));
            MDL_CHECK_CUDA(cudaMemsetAsync(gmaster, 0, (size_t)d.ip[MDL_ENC_I_MASTER_NUMEL] * sizeof(float), st));
        }
        const float* logits = d.ptr<const float>(MDL_ENC_P_LOGITS);
        const float* slide_hm = d.ptr<const float>(MDL_ENC_P_SLIDE_HM);
        // ---- projector / pooled gradient (head-major) ----
        float* dS_all = a.get<float>((size_t)n_slide * C);
        const float* d_slide = d.ptr<const float>(MDL_ENC_P_D_SLIDE);
        if (!dry) {
            if (!d_slide) {
                MDL_CHECK_CUDA(cudaMemsetAsync(dS_all, 0, (size_t)n_slide * C * sizeof(float), st));
            } else if (d.want_proj) {
                RUN(MDL_PROF_SKINNY, mdl_skinny_linear_bwd(d_slide, slide_hm, d.wf(MDL_ENC_I_F32_WP), n_slide, C, HID, dS_all,
                                                           g(MDL_ENC_I_GR_WP), g(MDL_ENC_I_GR_BP), stv));
            } else {
                MDL_CHECK_CUDA(cudaMemcpyAsync(dS_all, d_slide, (size_t)n_slide * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
            }
        }
        // ---- pooling backward: dlogit ----
        float* dlogit = a.get<float>((size_t)M * H);
        const float* d_logits = d.ptr<const float>(MDL_ENC_P_D_LOGITS);
        int accumulate = 0;
        if (d_logits) {
            if (!dry) MDL_CHECK_CUDA(cudaMemcpyAsync(dlogit, d_logits, (size_t)M * H * sizeof(float), cudaMemcpyDeviceToDevice, st));
            accumulate = 1;
        }
        RUN(MDL_PROF_POOL_BWD, mdl_pool_bwd_dlogit(s.h3, M * C, fnpl, dS_all, slide_hm, s.attn_p, cu, nullptr, R, M, H, HID, dlogit,
                                                   accumulate, logits, d.act, 0, stv));
        const float *p1 = nullptr, *dS1 = nullptr; const int* seg1 = nullptr;
        if (d.has_views()) {
            const int* tok_idx = d.ptr<const int>(MDL_ENC_P_VIEW_TOK_IDX);
            const int* cu2 = d.ptr<const int>(MDL_ENC_P_VIEW_CU);
            float* dS2 = dS_all ? dS_all + (size_t)R * C : nullptr;
            RUN(MDL_PROF_POOL_BWD, mdl_pool_bwd_dlogit(s.h3, M * C, fnpl, dS2, slide_hm + (size_t)R * C, s.attn_p2, cu2, tok_idx, d.R2,
                                                       d.n_view_tok, H, HID, dlogit, 1, logits, 0, 0, stv));
            p1 = s.attn_p2; dS1 = dS2; seg1 = d.ptr<const int>(MDL_ENC_P_VIEW_ROW2SEG);
        }
        // ---- gated attention backward ----
        const size_t m_attn = a.mark();
        void* dh3_attn = a.raw((size_t)M * C * d.act_size);
        {
            const size_t m0 = a.mark();
            __nv_bfloat16* dpre = a.get<__nv_bfloat16>((size_t)npl * M * H * 1024);
            RUN(MDL_PROF_GATE_BWD, mdl_gate_bwd(s.gate_a, s.gate_b, dlogit, d.wf(MDL_ENC_I_F32_WC), M, H, d.p_gate, d.seed, dpre,
                                                M * H * 1024, npl, g(MDL_ENC_I_GR_BA), g(MDL_ENC_I_GR_BB), g(MDL_ENC_I_GR_WC),
                                                g(MDL_ENC_I_GR_BC), stv));
            RUN(MDL_PROF_GEMM_NT, mdl_gemm_nt(dpre, M, (long long)H * 1024, (long long)H * 1024, M * H * 1024, d.wbf(MDL_ENC_I_BF_WABT),
                                              (long long)H * HID, 1024, 1024, bfn, dh3_attn, C, (int)M, C, 1024, nsplit, HID, 1024,
                                              nullptr, nullptr, nullptr, d.act_bf16, stv));
            RUN(MDL_PROF_GEMM_TN, mdl_gemm_tn_accum(dpre, (long long)H * 1024, (long long)H * 1024, M * H * 1024, s.h3, C, C, M * C, M,
                                                    g(MDL_ENC_I_GR_WAB), HID, H * 1024, HID, nsplit, 1024, HID, 0, stv));
            a.release(m0);
        }
        // ---- token projector backward ----
        void* dh3_tok = nullptr;
        const int* dh3_tok_rows = nullptr;
        const float* d_tokens = d.ptr<const float>(MDL_ENC_P_D_TOKENS);
        const void* d_ref = d.ptr<const void>(MDL_ENC_P_D_REF_HM);
        if (d_tokens || (dry && d.want_tok)) {
            const bool win = d.has_window();
            const long long n_tok = win ? d.n_sel : M;
            const __nv_bfloat16* tok_src = win ? s.h3_sel : s.h3;
            dh3_tok = a.raw((size_t)n_tok * C * d.act_size);
            const size_t m0 = a.mark();
            __nv_bfloat16* dtp = a.get<__nv_bfloat16>((size_t)npl * n_tok * TOK);
            RUN(MDL_PROF_OTHER, mdl_split_planes(d_tokens, n_tok, TOK, TOK, dtp, n_tok * TOK, npl, stv));
            RUN(MDL_PROF_GEMM_NT, mdl_gemm_nt(dtp, n_tok, TOK, TOK, n_tok * TOK, d.wbf(MDL_ENC_I_BF_TPT), C, TOK, TOK, bfn, dh3_tok, C,
                                              (int)n_tok, C, TOK, nsplit, 0, 0, nullptr, nullptr, nullptr, d.act_bf16, stv));
            RUN(MDL_PROF_GEMM_TN, mdl_gemm_tn_accum(dtp, TOK, TOK, n_tok * TOK, tok_src, C, C, n_tok * C, n_tok, g(MDL_ENC_I_GR_TP), C,
                                                    TOK, C, nsplit, 0, 0, 0, stv));
            RUN(MDL_PROF_OTHER, mdl_colsum_f32(d_tokens, n_tok, TOK, g(MDL_ENC_I_GR_BTP), stv));
            a.release(m0);
            if (win) dh3_tok_rows = d.ptr<const int>(MDL_ENC_P_TOKEN_SEL_OF_ROW);
        }
        if (d_ref) {
            if (!dry) MDL_REQUIRE(dh3_tok_rows == nullptr, "a token window cannot be combined with gradients through the pre-attention features");
            if (dh3_tok == nullptr) {
                dh3_tok = const_cast<void*>(d_ref);
            } else if (!dry) {
                const long long n = M * C;
                const int blocks = (int)((n + 255) / 256);
                if (d.act_bf16) launch_k(add_inplace_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, (__nv_bfloat16*)dh3_tok, (const __nv_bfloat16*)d_ref, n);
                else launch_k(add_inplace_kernel<float>, dim3(blocks), dim3(256), 0, st, (float*)dh3_tok, (const float*)d_ref, n);
                MDL_CHECK_LAUNCH();
            }
        }
        // ---- layer 3 ----
        __nv_bfloat16* dz3 = a.get<__nv_bfloat16>((size_t)npl * M * C);
        RUN(MDL_PROF_LN_BWD, mdl_ln_gelu_bwd(s.z3, M, C, d.wf(MDL_ENC_I_F32_G3), d.wf(MDL_ENC_I_F32_BE3), s.mean3, s.rstd3, dh3_attn, dh3_tok,
                                             dh3_tok_rows, s.attn_p, dS_all, s.row2bag, p1, dS1, seg1, H, d.p_pre, d.seed, 3, dz3, M * C, npl,
                                             g(MDL_ENC_I_GR_G3), g(MDL_ENC_I_GR_BE3), g(MDL_ENC_I_GR_B3), nullptr, nullptr, d.act_bf16, stv));
        RUN(MDL_PROF_GEMM_NT, mdl_gemm_nt(dz3, M, C, C, M * C, d.wbf(MDL_ENC_I_BF_W3T), HID, C, C, bfn, dh2, HID, (int)M, HID, C, nsplit, 0, 0,
                                          nullptr, nullptr, nullptr, d.act_bf16, stv));
        RUN(MDL_PROF_GEMM_TN, mdl_gemm_tn_accum(dz3, C, C, M * C, s.h2, HID, HID, M * HID, M, g(MDL_ENC_I_GR_W3), HID, C, HID, nsplit, 0, 0, 0, stv));
        (void)m_attn;
        if (phase == 1 && d.ip[MDL_ENC_I_GR_N_EARLY] > 0) {
            RUN(MDL_PROF_OTHER, mdl_permute_f32(gp, d.ptr<const int>(MDL_ENC_P_GR_POS_EARLY), d.ptr<const int>(MDL_ENC_P_GR_DST_EARLY),
                                                d.ip[MDL_ENC_I_GR_N_EARLY], gmaster, stv));
        }
    }
    a.release(base_mark);
    if (second) {
        __nv_bfloat16* dz2 = a.get<__nv_bfloat16>((size_t)npl * M * HID);
        RUN(MDL_PROF_LN_BWD, mdl_ln_gelu_bwd(s.z2, M, HID, d.wf(MDL_ENC_I_F32_G2), d.wf(MDL_ENC_I_F32_BE2), s.mean2, s.rstd2, dh2, nullptr, nullptr,
                                             nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1, d.p_pre, d.seed, 2, dz2, M * HID, npl,
                                             g(MDL_ENC_I_GR_G2), g(MDL_ENC_I_GR_BE2), g(MDL_ENC_I_GR_B2), nullptr, nullptr, d.act_bf16, stv));
        void* dh1 = a.raw((size_t)M * HID * d.act_size);
        RUN(MDL_PROF_GEMM_NT, mdl_gemm_nt(dz2, M, HID, HID, M * HID, d.wbf(MDL_ENC_I_BF_W2T), HID, HID, HID, bfn, dh1, HID, (int)M, HID, HID,
                                          nsplit, 0, 0, nullptr, nullptr, nullptr, d.act_bf16, stv));
        RUN(MDL_PROF_GEMM_TN, mdl_gemm_tn_accum(dz2, HID, HID, M * HID, s.h1, HID, HID, M * HID, M, g(MDL_ENC_I_GR_W2), HID, HID, HID, nsplit,
                                                0, 0, 0, stv));
        float* G = d.se > 0 ? a.get<float>((size_t)R * HID) : nullptr;      // per-bag column sums of dz1 (stain-encoding backward)
        if (G && !dry) MDL_CHECK_CUDA(cudaMemsetAsync(G, 0, (size_t)R * HID * sizeof(float), st));
        __nv_bfloat16* dz1 = a.get<__nv_bfloat16>((size_t)npl * M * HID);
        RUN(MDL_PROF_LN_BWD, mdl_ln_gelu_bwd(s.z1, M, HID, d.wf(MDL_ENC_I_F32_G1), d.wf(MDL_ENC_I_F32_BE1), s.mean1, s.rstd1, dh1, nullptr, nullptr,
                                             nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1, d.p_pre, d.seed, 1, dz1, M * HID, npl,
                                             g(MDL_ENC_I_GR_G1), g(MDL_ENC_I_GR_BE1), g(MDL_ENC_I_GR_B1), G ? s.row2bag : nullptr, G,
                                             d.act_bf16, stv));
        if (d.d_in % 256 == 0) {
            RUN(MDL_PROF_GEMM_TN, mdl_gemm_tn_accum(dz1, HID, HID, M * HID, s.xp, d.d_in, d.d_in, M * d.d_in, M, g(MDL_ENC_I_GR_W1), d.d_in, HID,
                                                    d.d_in, nsplit, 0, 0, 0, stv));
        } else {
            // the wgrad tile is 128 x 256: for feature widths that are a multiple of 128 only compute dW1^T and add its transpose
            float* w1t = a.get<float>((size_t)d.d_in * HID);
            if (!dry) MDL_CHECK_CUDA(cudaMemsetAsync(w1t, 0, (size_t)d.d_in * HID * sizeof(float), st));
            RUN(MDL_PROF_GEMM_TN, mdl_gemm_tn_accum(s.xp, d.d_in, d.d_in, M * d.d_in, dz1, HID, HID, M * HID, M, w1t, HID, d.d_in, HID, nsplit,
                                                    0, 0, 0, stv));
            if (!dry) {
                dim3 grid((d.d_in + 31) / 32, (HID + 31) / 32), block(32, 8);
                launch_k(add_transposed_kernel, dim3(grid), dim3(block), 0, st, g(MDL_ENC_I_GR_W1), w1t, HID, d.d_in);
                MDL_CHECK_LAUNCH();
            }
        }
        if (phase == 2) {
            if (d.ip[MDL_ENC_I_GR_N_LATE] > 0)
                RUN(MDL_PROF_OTHER, mdl_permute_f32(gp, d.ptr<const int>(MDL_ENC_P_GR_POS_LATE), d.ptr<const int>(MDL_ENC_P_GR_DST_LATE),
                                                    d.ip[MDL_ENC_I_GR_N_LATE], gmaster, stv));
        } else {
            RUN(MDL_PROF_OTHER, mdl_permute_f32(gp, d.ptr<const int>(MDL_ENC_P_GR_POS), d.ptr<const int>(MDL_ENC_P_GR_DST), d.ip[MDL_ENC_I_GR_N],
                                                gmaster, stv));
        }
        if (d.se > 0) {
            if (!dry) MDL_REQUIRE(master && d.ptr<void>(MDL_ENC_P_CODES), "mdl_encoder_bwd: stain encodings need MASTER and CODES");
            RUN(MDL_PROF_OTHER, mdl_stain_rowbias_bwd(G, master + d.ip[MDL_ENC_I_MASTER_EMB], d.ptr<const int>(MDL_ENC_P_CODES),
                                                      master + d.ip[MDL_ENC_I_MASTER_PRE0W], d.d_in_total, d.d_in, d.se, HID, R,
                                                      gmaster + d.ip[MDL_ENC_I_MASTER_PRE0W], gmaster + d.ip[MDL_ENC_I_MASTER_EMB], stv));
        }
    }
    if (bytes_out) *bytes_out = a.peak;
    return 0;
}

}  // namespace mdl

using namespace mdl;

extern "C" {

int mdl_encoder_abi(void) { return MDL_ENC_I_COUNT * 10000 + MDL_ENC_F_COUNT * 1000 + MDL_ENC_P_COUNT; }

long long mdl_encoder_fwd_arena_bytes(const long long* ip) {
    Desc d(ip, nullptr, nullptr);
    size_t bytes = 0;
    run_fwd(d, true, &bytes);
    return (long long)bytes + 256;
}

long long mdl_encoder_bwd_arena_bytes(const long long* ip) {
    // the backward arena must serve every phase: size it for the whole pass
    std::vector<long long> tmp(ip, ip + MDL_ENC_I_COUNT);
    tmp[MDL_ENC_I_PHASE] = 0;
    Desc d(tmp.data(), nullptr, nullptr);
    size_t bytes = 0;
    run_bwd(d, true, &bytes);
    return (long long)bytes + 256;
}

int mdl_encoder_fwd(const long long* ip, const double* fp, void* const* pp) {
    MDL_REQUIRE(ip && fp && pp, "mdl_encoder_fwd: null argument array");
    Desc d(ip, fp, pp);
    MDL_REQUIRE(d.H == 4 && d.M > 0 && d.R > 0, "mdl_encoder_fwd: unsupported shape (M=%lld, R=%d, heads=%d)", d.M, d.R, d.H);
    MDL_REQUIRE(d.ptr<void>(MDL_ENC_P_ARENA), "mdl_encoder_fwd: arena missing");
    return run_fwd(d, false, nullptr);
}

int mdl_encoder_bwd(const long long* ip, const double* fp, void* const* pp) {
    MDL_REQUIRE(ip && fp && pp, "mdl_encoder_bwd: null argument array");
    Desc d(ip, fp, pp);
    MDL_REQUIRE(d.keep, "mdl_encoder_bwd: the forward pass ran without keeping its state");
    MDL_REQUIRE(d.ptr<void>(MDL_ENC_P_ARENA) && d.ptr<void>(MDL_ENC_P_BWD_ARENA), "mdl_encoder_bwd: arena missing");
    return run_bwd(d, false, nullptr);
}

int mdl_permute_f32(const float* src, const int* pos, const int* dst, long long n, float* out, void* stream) {
    if (n <= 0) return 0;
    launch_k(permute_f32_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, src, pos, dst, n, out);
    MDL_CHECK_LAUNCH();
    return 0;
}

int mdl_gram_f64(const float* E, long long n, int D, double* G, void* stream) {
    MDL_REQUIRE(E && G && n > 0 && D > 0, "gram: empty input");
    dim3 grid((D + 31) / 32, (D + 31) / 32);
    launch_k(gram_f64_kernel, dim3(grid), dim3(1024), 0, (cudaStream_t)stream, E, n, D, G);
    MDL_CHECK_LAUNCH();
    return 0;
}

long long mdl_executor_launches(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}

int mdl_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.mask = on == 1 ? 0xffffffffu : (unsigned)on;
    return 0;
}

int mdl_profile_read(int* tags, float* ms, int max) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    const int n = (int)g_prof.recs.size();
    for (int i = 0; i < n; ++i) {
        ProfRec& r = g_prof.recs[i];
        float t = 0.f;
        cudaEventSynchronize(r.e1);
        cudaEventElapsedTime(&t, r.e0, r.e1);
        if (i < max) { tags[i] = r.tag; ms[i] = t; }
        g_prof.pool.push_back(r.e0);
        g_prof.pool.push_back(r.e1);
    }
    g_prof.recs.clear();
    return n;
}

}  // extern "C"
