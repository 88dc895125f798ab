/*
 * madeleine_b200 — C ABI of the B200-native (sm_100a) MADELEINE hot path.
 *
 * The reference (mahmoodlab/MADELEINE) is pure Python/PyTorch and has no FFI of its own; the entry points below
 * are what a binding for its hot path calls instead of the torch ops cited on each function (file:line under
 * /root/reference).  Host code (madeleine_b200/ops.py, or the ctypes stub in INTEGRATION.md) loads
 * libmadeleine_b200.so and passes raw device pointers.
 *
 * Conventions
 *   - every function returns 0 on success; non-zero means failure and mdl_last_error() describes it (thread-local);
 *   - all pointers are DEVICE pointers unless named host_*; the library never allocates or frees device memory
 *     and keeps no global mutable state besides cached function attributes;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, nothing synchronises;
 *   - "planes": an fp32 matrix held as bf16 hi plane followed (plane_stride elements later) by a bf16 lo plane,
 *     x ~= hi + lo.  nplanes/nsplit = 2/3 gives fp32-grade products on the bf16 tensor pipe, 1/1 plain bf16;
 *   - MDL_PLANES_F16 OR-ed into an `nplanes` or `nsplit` argument selects fp16 hi/lo planes instead (x ~= hi + lo with
 *     2 x 11 mantissa bits, 2^-22 per element instead of 2^-17): the fp32-grade INFERENCE format.  Supported by
 *     mdl_split_planes, mdl_gather_split (weights: values are multiplied by 64 before the split so that the lo plane
 *     stays in fp16's normal range), mdl_ln_gelu_fwd, mdl_gemm_nt and mdl_gemm_gated (both operands fp16, B planes from
 *     mdl_gather_split; the accumulator is multiplied by 1/64), mdl_pool_fwd and mdl_planes_to_ref_order.  Conversions
 *     saturate at +-65504.  Backward entry points take bf16 planes only;
 *   - token rows are bag-packed: bag r owns rows [cu_seqlens[r], cu_seqlens[r+1]).
 */
#ifndef MADELEINE_B200_H
#define MADELEINE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MDL_PLANES_F16 0x100

const char* mdl_last_error(void);
/* Library/ABI version and the compute capability the kernels were built for (100 = sm_100a). */
int mdl_version(void);
int mdl_built_arch(void);

/* ---- operand preparation ---------------------------------------------------------------------------------- */
/* fp32 [rows, cols] (row stride ld) -> bf16 planes.  Replaces the implicit autocast casts of torch.amp.autocast
 * (madeleine/utils/trainer.py:108). */
int mdl_split_planes(const float* x, long long rows, int cols, long long ld, void* planes, long long plane_stride,
                     int nplanes, void* stream);
/* planes[i] = split(src[idx[i]]): weight packing (row permutation to head-major = einops rearrange
 * 'b t (e c) -> b t e c' at Model.py:396, transposes for dgrad, gate interleave for abmil.py:49-52). */
int mdl_gather_split(const float* src, const int* idx, long long n, void* planes, long long plane_stride, int nplanes,
                     void* stream);
int mdl_gather_f32(const float* src, const int* idx, long long n, float* dst, void* stream);
int mdl_scatter_f32(const float* src, const int* idx, long long n, float* dst, int accumulate, void* stream);
int mdl_row2bag(const int* cu_seqlens, int n_bags, int* row2bag, long long rows, void* stream);

/* ---- tcgen05 GEMMs ------------------------------------------------------------------------------------------ */
/* out[M,N] = A[M,K] * B[N,K]^T + bias[n] + rowbias[row2bag[m], n]   (nn.Linear forward, Model.py:351,355,359,
 * 80-83, and every dgrad with a pre-transposed B).  A's k offset is (n / grp_n_cols) * a_koff when grp_n_cols > 0
 * (per-head operand slabs).  bias / rowbias / row2bag may be NULL.  out is fp32 [M, ldc]; with out_bf16 != 0 it is a
 * bf16 matrix (ldc in elements) holding the round-to-nearest of the fp32 result — what torch autocast returns from
 * nn.Linear in the reference's --precision bfloat16 runs (trainer.py:108). */
int mdl_gemm_nt(const void* a_planes, long long a_rows, long long a_cols, long long lda, long long a_plane_stride,
                const void* b_planes, long long b_rows, long long b_cols, long long ldb, long long b_plane_stride,
                void* out, long long ldc, int M, int N, int K, int nsplit, int grp_n_cols, int a_koff,
                const float* bias, const float* rowbias, const int* row2bag, int out_bf16, void* stream);
/* Gated-attention scores of all heads in one launch (BatchedABMIL.forward, abmil.py:49-52; head loop Model.py:406-409):
 * logits[m,h] = sum_j tanh(x_h Wa_h^T + ba)_j * sigmoid(x_h Wb_h^T + bb)_j * wc_hj + bc_h, x_h = A[:, h*512:(h+1)*512].
 * b_planes = packed [n_heads*4][128 Wa rows | 128 Wb rows][512].  gate_a/gate_b (optional, may be NULL): fp16 scratch for
 * mdl_gate_bwd, ceil(M/32)*32 * n_heads*512 elements each, holding the dropout-scaled tanh / sigmoid outputs in a TILED
 * layout — element (row m, gate column j) at ((((m/32) * (n_heads*32) + j/16) * 2 + (j%16)/8) * 32 + m%32) * 8 + j%8, i.e.
 * blocks of 32 rows x 8 columns are contiguous (full-line stores from the epilogue's row-per-lane registers). */
int mdl_gemm_gated(const void* a_planes, long long a_rows, long long a_cols, long long lda, long long a_plane_stride,
                   const void* b_planes, long long b_plane_stride, int M, int n_heads, int nsplit,
                   const float* ba, const float* bb, const float* wc, const float* bc,
                   float* logits, void* gate_a, void* gate_b, float drop_p, unsigned long long seed, void* stream);
/* out[M,N] += sum_t A[t, m]^T B[t, n]  (weight gradients; split over tokens, fp32 red.add).  B's column offset is
 * (m / grp_m_rows) * b_coff when grp_m_rows > 0.  ksplit <= 0 picks a split automatically. */
int mdl_gemm_tn_accum(const void* a_planes, long long a_cols, long long lda, long long a_plane_stride,
                      const void* b_planes, long long b_cols, long long ldb, long long b_plane_stride,
                      long long tokens, float* out, long long ldc, int M, int N, int nsplit,
                      int grp_m_rows, int b_coff, int ksplit, void* stream);
/* Benchmark-only knob for tools/gemm_bounds.py (results are WRONG while set): 1 = skip TMA operand loads,
 * 2 = skip epilogue global writes in mdl_gemm_nt.  0 restores normal operation. */
int mdl_gemm_debug_flags(int flags);
/* CUDA-core cross-checks of the two GEMM shapes above (tests only; never on the product path). */
int mdl_gemm_nt_simt(const void* a_planes, long long lda, long long a_plane_stride, const void* b_planes, long long ldb,
                     long long b_plane_stride, float* out, long long ldc, int M, int N, int K, int nsplit, void* stream);
int mdl_gemm_tn_simt(const void* a_planes, long long lda, long long a_plane_stride, const void* b_planes, long long ldb,
                     long long b_plane_stride, long long tokens, float* out, long long ldc, int M, int N, int nsplit,
                     void* stream);

/* ---- LayerNorm + GELU (+dropout) ------------------------------------------------------------------------------ */
/* nn.LayerNorm -> nn.GELU -> nn.Dropout of ABMILEmbedder.pre_attn (Model.py:352-354, 356-358, 360-362).
 * z is fp32 [M, C], or bf16 when z_bf16 != 0 (bf16 mode: the Linear outputs are stored as autocast returns them). */
int mdl_ln_gelu_fwd(const void* z, long long M, int C, const float* gamma, const float* beta, float eps,
                    float drop_p, unsigned long long seed, unsigned stream_id,
                    void* planes, long long plane_stride, int nplanes, float* mean, float* rstd, int z_bf16, void* stream);
/* Backward of the above; dh = dh_a + dh_b + sum_v pool_p_v[m, head] * pool_dS_v[pool_seg_v[m], c] (any may be NULL).
 * dh_b_rows == NULL: dh_b is dense [M, C]; otherwise dh_b is compact [n_sel, C] and dh_b_rows[m] is token m's compact row
 * or -1 (the token_projector gradient of the token window, Model.py:138-146 + loss.py:281-284).
 * Accumulates dgamma, dbeta and the preceding Linear's bias grad dbias (all [C], caller zero-fills).
 * in_bf16 != 0: z, dh_a and dh_b are bf16 instead of fp32 (bf16 mode).
 * bag_dz != NULL (C == 512, no dh_b / pooling term): also accumulates per-bag column sums of dz into bag_dz[row2bag[m], c]
 * ([n_bags, C], caller zero-fills) — the stain-encoding backward of Model.py:126-133. */
int mdl_ln_gelu_bwd(const void* z, long long M, int C, const float* gamma, const float* beta, const float* mean,
                    const float* rstd, const void* dh_a, const void* dh_b, const int* dh_b_rows,
                    const float* pool_p0, const float* pool_dS0, const int* pool_seg0,
                    const float* pool_p1, const float* pool_dS1, const int* pool_seg1, int n_heads,
                    float drop_p, unsigned long long seed, unsigned stream_id,
                    void* dz_planes, long long plane_stride, int nplanes,
                    float* dgamma, float* dbeta, float* dbias, const int* row2bag, float* bag_dz, int in_bf16, void* stream);
/* Backward of the gate nonlinearities of mdl_gemm_gated; gate_a / gate_b in the tiled scratch layout described there;
 * dpre planes [M, n_heads*1024] row-major in packed gate order.  `seed` is unused: the dropout masks are read off the saved
 * gates (a dropped gate is an exact zero). */
int mdl_gate_bwd(const void* gate_a, const void* gate_b, const float* dlogit, const float* wc, long long M, int n_heads,
                 float drop_p, unsigned long long seed, void* dpre_planes, long long plane_stride, int nplanes,
                 float* dba, float* dbb, float* dwc, float* dbc, void* stream);

/* ---- attention pooling (the HBM-bound kernel) ---------------------------------------------------------------- */
/* softmax over tokens + weighted sum (abmil.py:55, Model.py:416-417) in two launches:
 *   mdl_pool_weights: attn_p[t, h] = act(logits) per (bag, head); activation 0 softmax over the bag's tokens, 1 leaky_relu,
 *                     2 relu, 3 sigmoid (abmil.py:54-63).  Also clears the split tickets in `workspace`.
 *   mdl_pool_fwd:     out[r, h, :] = sum_t attn_p[t, h] * X[t, h, :]   (the HBM-bound streaming kernel), head-major out.
 * tok_idx (optional) gathers rows: segment r pools rows tok_idx[cu[r] .. cu[r+1]) (n_views=3, Model.py:427-437); rows
 * outside every segment are left untouched in attn_p.  Each bag is split over `tsplit` CTAs along tokens (mdl_pool_tsplit
 * suggests ~192 tokens per CTA); partial sums go through `workspace` (mdl_pool_workspace_bytes; may be NULL when
 * tsplit == 1) and are combined in a fixed order, so results are bit-reproducible. */
int mdl_pool_tsplit(int n_bags, long long total_tokens, int n_heads, int head_dim);
long long mdl_pool_workspace_bytes(int n_bags, int n_heads, int head_dim, int tsplit);
int mdl_pool_weights(const float* logits, const int* cu_seqlens, const int* tok_idx, int n_bags, int n_heads, int head_dim,
                     float* attn_p, int activation, int tsplit, void* workspace, void* stream);
int mdl_pool_fwd(const void* x_planes, long long plane_stride, int nplanes, const float* attn_p, const int* cu_seqlens,
                 const int* tok_idx, int n_bags, long long total_tokens, int n_heads, int head_dim,
                 float* out, int tsplit, void* workspace, void* stream);
int mdl_pool_bwd_dlogit(const void* x_planes, long long plane_stride, int nplanes, const float* dS, const float* S,
                        const float* attn_p, const int* cu_seqlens, const int* tok_idx, int n_bags,
                        long long total_tokens, int n_heads, int head_dim, float* dlogit, int accumulate,
                        const float* logits, int activation, int tsplit, void* stream);
/* head-major planes -> fp32 [M, head_dim, n_heads] (reference channel order), ABMILEmbedder(return_preattn_feats). */
int mdl_planes_to_ref_order(const void* x_planes, long long plane_stride, int nplanes, long long M, int n_heads,
                            int head_dim, float* out, void* stream);

/* ---- slide-level projector, stain encodings -------------------------------------------------------------------- */
/* MADELEINE.projector (Model.py:88-91,145) on [n_bags, 2048]: exact fp32. */
int mdl_skinny_linear_fwd(const float* X, const float* W, const float* b, int R, int C, int O, float* Y, void* stream);
int mdl_skinny_linear_bwd(const float* dY, const float* X, const float* W, int R, int C, int O,
                          float* dX, float* dW, float* db, void* stream);
/* Stain encodings (Model.py:125-132) folded into a per-bag bias: rowbias[r,:] = W1[:, d_in:] emb[code[r]]. */
int mdl_stain_rowbias(const float* emb, const int* code, const float* w1, int ldw, int d_in, int se_dim, int n_out,
                      int R, float* rowbias, void* stream);
/* out[p][s, :] = planes[p][rows[s], :]: the token rows that token_projector (Model.py:80-83,138-146) has to see when only
 * a window of every bag's tokens can reach the local loss (GOT's permutation is over the batch size, loss.py:281-284). */
int mdl_gather_rows_planes(const void* planes, long long plane_stride_in, int nplanes, int C, const int* rows,
                           long long n_sel, void* out, long long plane_stride_out, void* stream);
int mdl_bag_colsum_planes(const void* planes, long long plane_stride, int nplanes, int C, const int* cu_seqlens,
                          int n_bags, float* out, void* stream);
int mdl_stain_rowbias_bwd(const float* G, const float* emb, const int* code, const float* w1, int ldw, int d_in,
                          int se_dim, int n_out, int R, float* dw1, float* demb, void* stream);
int mdl_colsum_f32(const float* x, long long M, int C, float* out, void* stream);

/* ---- losses ------------------------------------------------------------------------------------------------------ */
/* InfoNCE with in-batch negatives (madeleine/utils/loss.py:111-127). reduction: 0 none, 1 mean, 2 sum.
 * lse_r / lse_c hold 2m floats each: [row max | log-sum]. */
int mdl_infonce_fwd(const float* q, const float* k, int m, int D, float temperature, int symmetric, int reduction,
                    float* qn, float* kn, float* L, float* lse_r, float* lse_c, float* nll_r, float* nll_c,
                    float* loss, void* stream);
int mdl_infonce_bwd(const float* q, const float* k, int m, int D, float temperature,
                    const float* qn, const float* kn, const float* L, const float* lse_r, const float* lse_c,
                    const float* w_r, const float* w_c, float* G, float* dq, float* dk, void* stream);
/* The same loss on rows picked out of one matrix: q_i = base + q_rows[i] * ld, k_i = base + k_rows[i] * ld (the boolean-mask
 * row selection of calculate_losses, madeleine/utils/trainer.py:28-33, fused into the kernels — no gathered copies, no
 * index_put in backward).  reduction = mean.  `workspace`: mdl_infonce_rows_workspace_floats(m) floats, kept by the caller
 * from forward to backward.  loss: this term; total (optional): running sum over the terms of a step (+=, stream-ordered).
 * Backward ADDS (*go) * d loss / d row into dbase (same layout as base; the caller zero-fills it once per step). */
long long mdl_infonce_rows_workspace_floats(int m);
int mdl_infonce_rows_fwd(const float* base, long long ld, const int* q_rows, const int* k_rows, int m, int D, float temperature,
                         int symmetric, float* workspace, float* loss, float* total, void* stream);
int mdl_infonce_rows_bwd(const float* base, long long ld, const int* q_rows, const int* k_rows, int m, int D, float temperature,
                         int symmetric, float* workspace, const float* go, float* dbase, void* stream);
/* Graph-OT local loss, forward and backward in one call (GOT, madeleine/utils/loss.py:278-301 with
 * cost_matrix_batch_torch :162-176, IPOT :179-207, cos_batch_torch :210-233, GW :236-275).
 * v, q [m, n, D] (already token-subsampled); loss = sum_b (wd_b + gwd_b); dv, dq = d loss / d v, q.
 * extrema (host-visible optional hook for sharded runs): 6 floats {min,max} x {cross, intra_v, intra_q}; when
 * use_external_extrema != 0 the thresholds use these instead of the local batch extrema (quirk Q5 under sharding). */
long long mdl_got_workspace_bytes(int m, int n, int D);
int mdl_got_max_tokens(void);
/* Test hook: send problems of any size to the global-memory kernels used for 96 < n <= 256 (got_big.cu), so that both
 * implementations can be compared on the same inputs.  Also settable with MDL_GOT_FORCE_BIG=1. */
int mdl_got_force_big(int on);
int mdl_got_extrema(const float* v, const float* q, int m, int n, int D, void* workspace, float* extrema, void* stream);
int mdl_got_fwd_bwd(const float* v, const float* q, int m, int n, int D, void* workspace, const float* extrema,
                    float* loss, float* wd, float* gwd, float* dv, float* dq, void* stream);
/* The same in two steps for case-sharded runs (quirk Q5 couples all problems of a stain through the batch-wide min/max):
 *   mdl_got_extrema -> all-reduce MIN/MAX of `extrema` across ranks -> mdl_got_main (also writes the 3 local
 *   threshold-gradient sums to dthr_local) -> all-reduce SUM of dthr -> mdl_got_finish(dthr_global).
 * A rank adds the threshold gradient to its arg-min/max element only if its local extremum equals the global one. */
int mdl_got_main(int m, int n, int D, void* workspace, const float* extrema, float* wd, float* gwd, float* dthr_local,
                 void* stream);
int mdl_got_finish(const float* v, const float* q, int m, int n, int D, void* workspace, const float* extrema,
                   const float* dthr_global, const float* wd, const float* gwd, float* loss, float* dv, float* dq,
                   void* stream);

/* ---- native step executor (train-step caller) ------------------------------------------------------ */
/* ONE call runs the whole encoder forward (ABMILEmbedder.forward + projector + token_projector, Model.py:110-159,
 * 346-451) and ONE call its backward, over a caller-provided arena: every launch of the sequence that
 * madeleine_b200/ops.py used to issue through ~45 separate host calls is issued here, natively, on `stream`.
 * Arguments travel in three flat arrays indexed by the enums below (so the ABI survives new fields):
 *   ip  const long long[MDL_ENC_I_COUNT]   sizes, flags and element offsets
 *   fp  const double[MDL_ENC_F_COUNT]      dropout probabilities
 *   pp  void* const[MDL_ENC_P_COUNT]       device pointers (and the stream)
 * The arena (mdl_encoder_fwd_arena_bytes) holds every activation the backward pass needs; the caller keeps it alive
 * until mdl_encoder_bwd ran.  mdl_encoder_bwd needs a scratch arena of mdl_encoder_bwd_arena_bytes as well.
 * MDL_ENC_I_PHASE: 0 = whole backward; 1 = up to and including the moment the gradients of everything except the first
 * two pre-attention layers and the stain embedding are final in `gmaster` (the caller may start reducing that range
 * across ranks); 2 = the rest.  Weight/gradient segment offsets are those of ops.PackSpec. */
enum mdl_enc_i {
    MDL_ENC_I_M = 0, MDL_ENC_I_R, MDL_ENC_I_D_IN, MDL_ENC_I_D_IN_TOTAL, MDL_ENC_I_SE_DIM, MDL_ENC_I_N_HEADS,
    MDL_ENC_I_NSPLIT_FWD, MDL_ENC_I_NPL_FWD, MDL_ENC_I_NSPLIT_BWD, MDL_ENC_I_NPL_BWD, MDL_ENC_I_ACT_BF16,
    MDL_ENC_I_ACTIVATION, MDL_ENC_I_KEEP, MDL_ENC_I_WANT_TOKENS, MDL_ENC_I_WANT_PROJECTOR, MDL_ENC_I_WANT_REF,
    MDL_ENC_I_N_VIEW_TOK, MDL_ENC_I_R2, MDL_ENC_I_N_SEL, MDL_ENC_I_SEED, MDL_ENC_I_PHASE,
    MDL_ENC_I_BF_NUMEL, MDL_ENC_I_BF_W1, MDL_ENC_I_BF_W2, MDL_ENC_I_BF_W2T, MDL_ENC_I_BF_W3, MDL_ENC_I_BF_W3T,
    MDL_ENC_I_BF_WAB, MDL_ENC_I_BF_WABT, MDL_ENC_I_BF_TP, MDL_ENC_I_BF_TPT,
    /* fp32 vectors, in this order: b1 g1 be1 b2 g2 be2 b3 g3 be3 ba bb wc bc btp wp bp */
    MDL_ENC_I_F32_B1, MDL_ENC_I_F32_G1, MDL_ENC_I_F32_BE1, MDL_ENC_I_F32_B2, MDL_ENC_I_F32_G2, MDL_ENC_I_F32_BE2,
    MDL_ENC_I_F32_B3, MDL_ENC_I_F32_G3, MDL_ENC_I_F32_BE3, MDL_ENC_I_F32_BA, MDL_ENC_I_F32_BB, MDL_ENC_I_F32_WC,
    MDL_ENC_I_F32_BC, MDL_ENC_I_F32_BTP, MDL_ENC_I_F32_WP, MDL_ENC_I_F32_BP,
    /* packed gradient buffer: matrices then the same 16 vectors */
    MDL_ENC_I_GR_NUMEL, MDL_ENC_I_GR_W1, MDL_ENC_I_GR_W2, MDL_ENC_I_GR_W3, MDL_ENC_I_GR_WAB, MDL_ENC_I_GR_TP,
    MDL_ENC_I_GR_B1, MDL_ENC_I_GR_G1, MDL_ENC_I_GR_BE1, MDL_ENC_I_GR_B2, MDL_ENC_I_GR_G2, MDL_ENC_I_GR_BE2,
    MDL_ENC_I_GR_B3, MDL_ENC_I_GR_G3, MDL_ENC_I_GR_BE3, MDL_ENC_I_GR_BA, MDL_ENC_I_GR_BB, MDL_ENC_I_GR_WC,
    MDL_ENC_I_GR_BC, MDL_ENC_I_GR_BTP, MDL_ENC_I_GR_WP, MDL_ENC_I_GR_BP,
    MDL_ENC_I_MASTER_NUMEL, MDL_ENC_I_MASTER_PRE0W, MDL_ENC_I_MASTER_EMB,
    MDL_ENC_I_GR_N, MDL_ENC_I_GR_N_EARLY, MDL_ENC_I_GR_N_LATE, MDL_ENC_I_PLANES_F16,
    MDL_ENC_I_COUNT
};
enum mdl_enc_f { MDL_ENC_F_P_PRE = 0, MDL_ENC_F_P_GATE, MDL_ENC_F_COUNT };
enum mdl_enc_p {
    MDL_ENC_P_STREAM = 0, MDL_ENC_P_X, MDL_ENC_P_CU, MDL_ENC_P_CODES, MDL_ENC_P_MASTER, MDL_ENC_P_WBF, MDL_ENC_P_WF32,
    MDL_ENC_P_ARENA, MDL_ENC_P_SLIDE_HM, MDL_ENC_P_SLIDE, MDL_ENC_P_LOGITS, MDL_ENC_P_TOKENS, MDL_ENC_P_REF,
    MDL_ENC_P_VIEW_TOK_IDX, MDL_ENC_P_VIEW_CU, MDL_ENC_P_VIEW_ROW2SEG, MDL_ENC_P_TOKEN_ROWS, MDL_ENC_P_TOKEN_SEL_OF_ROW,
    MDL_ENC_P_BWD_ARENA, MDL_ENC_P_D_SLIDE, MDL_ENC_P_D_LOGITS, MDL_ENC_P_D_TOKENS, MDL_ENC_P_D_REF_HM, MDL_ENC_P_GMASTER,
    MDL_ENC_P_GR_POS, MDL_ENC_P_GR_DST, MDL_ENC_P_GR_POS_EARLY, MDL_ENC_P_GR_DST_EARLY, MDL_ENC_P_GR_POS_LATE,
    MDL_ENC_P_GR_DST_LATE,
    MDL_ENC_P_COUNT
};
/* MDL_ENC_I_COUNT * 10000 + MDL_ENC_F_COUNT * 1000 + MDL_ENC_P_COUNT of the built library (binding sanity check). */
int mdl_encoder_abi(void);
long long mdl_encoder_fwd_arena_bytes(const long long* ip);
long long mdl_encoder_bwd_arena_bytes(const long long* ip);
/* Inputs: X fp32 [M, d_in] bag-packed, CU int32 [R+1], CODES int32 [R] (se_dim > 0), MASTER flat fp32 parameters,
 * WBF / WF32 packed operand planes / vectors (mdl_gather_split / mdl_gather_f32 of MASTER).  Outputs: SLIDE_HM fp32
 * [R + R2, n_heads*512] head-major pooled vectors (always), SLIDE [R + R2, 512] (want_projector), LOGITS [M, n_heads],
 * TOKENS [M or n_sel, 128] (want_tokens), REF [M, 512, n_heads] (want_ref). */
int mdl_encoder_fwd(const long long* ip, const double* fp, void* const* pp);
/* D_SLIDE: [R + R2, 512] (want_projector) or head-major [R + R2, n_heads*512]; D_LOGITS [M, n_heads]; D_TOKENS
 * [M or n_sel, 128]; D_REF_HM [M, n_heads*512] head-major in the activation dtype; any may be NULL.
 * GMASTER [master_numel] receives the parameter gradients in MASTER's layout (overwritten). */
int mdl_encoder_bwd(const long long* ip, const double* fp, void* const* pp);
/* out[dst[i]] = src[pos[i]] — packed gradient buffer -> parameter layout in one pass. */
int mdl_permute_f32(const float* src, const int* pos, const int* dst, long long n, float* out, void* stream);
/* G [D, D] fp64 = E^T E for E fp32 [n, D]: the Gram matrix whose eigenvalues are the squared singular values that
 * smooth_rank_measure needs (madeleine/utils/utils.py:180-201: torch.svd of the [n, 512] H&E embedding matrix on the CPU). */
int mdl_gram_f64(const float* E, long long n, int D, double* G, void* stream);
/* Number of C-ABI kernel entry points mdl_encoder_fwd / mdl_encoder_bwd called (all threads) since the last reset. */
long long mdl_executor_launches(int reset);

/* Per-launch device timing of the executor's kernels (process-wide switch; for bench.py's roofline only).  While enabled every
 * launch issued by mdl_encoder_fwd / mdl_encoder_bwd whose tag is selected is bracketed by CUDA events on its stream
 * (`on`: 0 = off, 1 = every launch, otherwise a bit mask, bit t = tag t; bit 0 set alone is spelled 1 = all).  mdl_profile_read
 * synchronises on the recorded events, writes up to `max` (tag, milliseconds) pairs in launch order, clears the
 * record and returns the number of pairs that were available. */
enum mdl_prof_tag {
    MDL_PROF_OTHER = 0, MDL_PROF_GEMM_NT, MDL_PROF_GEMM_GATED, MDL_PROF_GEMM_TN, MDL_PROF_LN_FWD, MDL_PROF_LN_BWD,
    MDL_PROF_GATE_BWD, MDL_PROF_POOL_WEIGHTS, MDL_PROF_POOL_FWD, MDL_PROF_POOL_BWD, MDL_PROF_SKINNY, MDL_PROF_COUNT
};
int mdl_profile_enable(int on);
int mdl_profile_read(int* tags, float* ms, int max);

/* Programmatic dependent launch of the hot-path kernels (csrc/common.cuh::launch_k): on by default, MADELEINE_B200_PDL=0 or
 * mdl_set_pdl(0) restores plain stream-ordered launches (same results; used for A/B timing).  Returns the previous setting. */
int mdl_set_pdl(int on);

/* ---- small exchanges over NVLink peer memory (case-sharded runs) -------------------------------------- */
/* One-kernel all-reduce (sum) / all-gather of small fp32 messages over symmetric peer buffers, replacing NCCL where the
 * message is latency-bound: the slide-embedding all-gather before the contrastive loss and the all-reduce of the 2 MB of
 * gradients that only become final at the end of backward.  host_bufs / host_sigs: HOST arrays of `world` DEVICE pointers
 * to every rank's symmetric buffer and uint32 signal pad (peer-mapped; this rank's own at [rank]).  The caller has put its
 * contribution at buf[rank] + buf_off_bytes (stream-ordered before the call) and alternates between two buffer halves from
 * call to call (epoch parity); `epoch` increases by one per call on a channel and is the same on every rank; slots
 * [slot_base, slot_base + world) of the signal pad belong to the channel.  out: [n] (all-reduce, may alias nothing in the
 * symmetric buffers) or [world * n_per_rank] rank-major (all-gather).  Sums are taken in rank order on every rank. */
int mdl_peer_allreduce_f32(void* const* host_bufs, void* const* host_sigs, int rank, int world, long long buf_off_bytes, long long n,
                           float* out, unsigned epoch, int slot_base, void* stream);
int mdl_peer_allgather_f32(void* const* host_bufs, void* const* host_sigs, int rank, int world, long long buf_off_bytes,
                           long long n_per_rank, float* out, unsigned epoch, int slot_base, void* stream);

/* ---- optimiser (train-step caller) -------------------------------------------------------------- */
/* Fused multi-tensor AdamW, one launch for all parameters (torch.optim.AdamW semantics; reference:
 * madeleine/utils/setup_components.py:194-196).  host_* are HOST arrays of n_tensors DEVICE pointers / element counts
 * (n_tensors <= mdl_adamw_max_tensors()); `step` counts from 1; gradients are multiplied by grad_scale first. */
int mdl_adamw_max_tensors(void);
int mdl_adamw_step(int n_tensors, void* const* host_params, void* const* host_grads, void* const* host_exp_avg,
                   void* const* host_exp_avg_sq, const long long* host_numels, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int step, float grad_scale, void* stream);

/* ---- on-device patch resampling from an HBM-resident feature store --------------------------------------------- */
/* SlideDataset.sample_n + collate (madeleine/datasets/wsi_dataset.py:42-50, 85-99) + the H2D copy of Model.py:113 in one
 * kernel: store [sum N, D] fp32 holds every bag; bag b = rows [bag_offset[b], bag_offset[b] + bag_len[b]).  For each of the
 * n_bags bags writes n_sample rows to out [n_bags, n_sample, D]: a random subset without replacement when
 * bag_len >= n_sample (keyed Feistel permutation, evaluated pointwise), draws with replacement when shorter, zero rows when
 * bag_len == 0 (missing stain).  idx_out (optional, [n_bags, n_sample]) receives the chosen row of each slot (-1: zero row). */
int mdl_sample_gather_f32(const float* store, const long long* bag_offset, const int* bag_len, int n_bags, int n_sample,
                          int D, unsigned long long seed, float* out, int* idx_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MADELEINE_B200_H */
