/* tan_b200.h -- C ABI of the B200-native TAN hot path (libtan_b200.so).
 *
 * The reference (TengdaHan/TemporalAlignNet) is pure Python on PyTorch and has no FFI of its own;
 * every arithmetic step of its hot path is a torch library call.  Each entry point below replaces
 * one group of those call sites (cited per function, paths relative to the reference root).  The
 * Python classes in temporalalignnet_b200/ (same names / arguments / state-dict keys as the
 * reference's TemporalAligner, TemporalEncoder, get_loss) bind these symbols with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *   - bf16 tensors are passed as `const void*` / `void*` (raw __nv_bfloat16 storage), row-major,
 *     innermost dimension contiguous; `ld*` arguments are row pitches in ELEMENTS;
 *   - `stream` is a cudaStream_t passed as void*; no call allocates, synchronises or uses the
 *     default stream implicitly; every call is re-entrant per stream;
 *   - return value: TAN_OK (0) or a negative TAN_ERR_* code; `tan_last_error_string()` gives a
 *     thread-local human-readable message.  Shape violations are reported, never "fixed up".
 *   - the library only contains sm_100a code: on any other device every compute entry point
 *     returns TAN_ERR_ARCH.  There is no CPU path.
 */
#ifndef TAN_B200_H_
#define TAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* exported symbols (the library is built with -fvisibility=hidden) */
#if defined(__GNUC__)
#define TAN_API __attribute__((visibility("default")))
#else
#define TAN_API
#endif

#define TAN_OK 0
#define TAN_ERR_SHAPE (-1)     /* unsupported / inconsistent dimensions or alignment */
#define TAN_ERR_ARCH (-2)      /* device is not sm_100 */
#define TAN_ERR_WORKSPACE (-3) /* workspace pointer null or too small */
#define TAN_ERR_CUDA (-4)      /* a CUDA runtime / driver call failed (see error string) */
#define TAN_ERR_ARG (-5)       /* null pointer or invalid flag */

/* Epilogue activation for tan_linear_bf16. */
#define TAN_ACT_NONE 0
#define TAN_ACT_QUICKGELU 1 /* x * sigmoid(1.702 x), model/tfm_model.py:11-13 */
#define TAN_ACT_RELU 2      /* max(x, 0), model/word2vec_model.py:86 */
#define TAN_ACT_QUICKGELU_GRAD 3 /* internal to tan_linear_gelu_bwd_bf16: multiply by d QuickGELU / dx of a second operand */

/* ---- library ------------------------------------------------------------------------------- */

/* ABI version of this header (bumped on any signature change). */
TAN_API int tan_abi_version(void);
/* Thread-local message for the last non-zero return code on this thread. */
TAN_API const char* tan_last_error_string(void);
/* TAN_OK if the current device is sm_100 (B200), TAN_ERR_ARCH otherwise, TAN_ERR_CUDA if no device. */
TAN_API int tan_device_check(void);

/* Development aid: point the GEMM kernels' per-CTA event trace at a device buffer of
 * 64 * gridDim int64 slots (NULL disables it, the default).  Slot 0 = globaltimer at CTA start, slots 1-3
 * = clock64 at start / after prologue / at exit, slots 4+6i.. = producer first/last issue, MMA first/last
 * issue, epilogue start/end of the CTA's i-th tile. */
TAN_API int tan_debug_set_trace(void* device_buffer);

/* ---- dtype conversion ------------------------------------------------------------------------ */

/* out[i] = bf16(in[i]), i < n.  Replaces the implicit fp32->half cast torch autocast inserts in
 * front of every F.linear under train/main.py:81.  n % 8 == 0, 16-byte aligned pointers. */
TAN_API int tan_cast_f32_to_bf16(const float* in, void* out, size_t n, void* stream);

/* ---- GEMM (+ fused epilogue) ----------------------------------------------------------------- */

/* out = act(A @ W^T + bias) [+ residual]
 *   A   [M, K] bf16 (lda), W [N, K] bf16 (nn.Linear weight layout, ldw), bias [N] fp32 or NULL,
 *   residual [M, N] fp32 (ldr) or NULL (may alias out_f32: each element is read before written),
 *   out_f32 [M, N] fp32 (ldo_f32) and/or out_bf16 [M, N] bf16 (ldo_bf16); at least one non-NULL.
 * tcgen05.mma (bf16 in, fp32 accumulate in TMEM), operands staged by TMA with 128-byte swizzle.
 * Requirements: K % 64 == 0, N % 128 == 0, lda/ldw % 8 == 0, 16-byte aligned bases, ld* % 4 == 0.
 * Replaces F.linear at: model/tan_model.py:155,:187,:233 (pre-projections), torch
 * nn/functional.py `_in_projection_packed` + out_proj reached from model/tfm_model.py:32 (QKV and
 * output projections, residual add of :36), model/tfm_model.py:23-27,:37 (c_fc + QuickGELU,
 * c_proj + residual). */
TAN_API int tan_linear_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                    const float* residual, int64_t ldr, float* out_f32, int64_t ldo_f32,
                    void* out_bf16, int64_t ldo_bf16, int M, int N, int K, int act, void* stream);

/* Training forward of c_fc: out_act = act(A W^T + bias) and out_pre = A W^T + bias (both bf16 [M, N]) from ONE GEMM --
 * QuickGELU's backward needs the pre-activation, the next GEMM the activation (model/tfm_model.py:23-25,:37).  Same
 * operand requirements as tan_linear_bf16. */
TAN_API int tan_linear_dual_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                 void* out_act, int64_t ldo_act, void* out_pre, int64_t ldo_pre, int M, int N, int K,
                                 int act, void* stream);

/* Backward through c_proj and QuickGELU in one GEMM: out = (A W^T) o gelu'(u), bf16 [M, N]; u [M, N] bf16 (ldu) are the
 * pre-activations stored by tan_linear_dual_bf16 (autograd of model/tfm_model.py:11-13,:37). */
TAN_API int tan_linear_gelu_bwd_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const void* u, int64_t ldu,
                                     void* out, int64_t ldo, int M, int N, int K, void* stream);

/* Fused projection + residual + LayerNorm for an output width of 512 (= the model width):
 *     x   <- x + A @ W^T + bias                 x [M, 512] fp32 (ldx), updated in place
 *     out <- LayerNorm(x) * gamma + beta         out [M, 512] bf16 (ldo); eps = 1e-5, biased variance
 *   A [M, K] bf16 (lda), W [512, K] bf16 (ldw), bias [512] fp32 or NULL, gamma / beta [512] fp32.
 * Requirements: N == 512, K % 64 == 0, lda/ldw/ldo % 8 == 0, ldx % 4 == 0, 16-byte aligned bases.
 * Replaces the attention out-projection + residual add (torch MHA out_proj reached from model/tfm_model.py:32,
 * `x = x + attn` at :36) together with `self.ln_2(x)` at :37: the fp32 residual stream is read and written
 * once instead of twice. */
TAN_API int tan_linear_res_ln_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                           float* x, int64_t ldx, const float* gamma, const float* beta, void* out_bf16,
                           int64_t ldo, int M, int N, int K, void* stream);

/* The same with the stage-feature emission of tan_layernorm (below): additionally (or instead of `out`, which may
 * be NULL) writes y / ||y||_2 as bf16, y = LayerNorm(x) * gamma + beta, for token r = b * L + l to
 *     nrmA_bf16[b * strideA + l]            (l <  l_split: video tokens)
 *     nrmB_bf16[b * strideB + l - l_split]  (l >= l_split: text tokens of the joint sequence; may be NULL if l_split == L).
 * Requirements: as above, M % L == 0, L % 32 == 0, l_split % 32 == 0 (a warp stores 32 consecutive tokens as one
 * TMA box).  Replaces c_proj + residual (model/tfm_model.py:37) together with the NEXT block's ln_1 (:35) -- whose
 * output is also the previous stage's feature (:50-53) -- or with ln_video_post_enc / ln_joint_post_enc
 * (model/tan_model.py:174,:206), and the L2 normalisation of model/tan_model.py:116-117,:136-137. */
TAN_API int tan_linear_res_ln_stage_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                 float* x, int64_t ldx, const float* gamma, const float* beta, void* out_bf16,
                                 int64_t ldo, int M, int N, int K, int L, int l_split, void* nrmA_bf16,
                                 int64_t strideA, void* nrmB_bf16, int64_t strideB, void* stream);

/* ---- LayerNorm (+ positional add, scatter, L2-normalised stage features) ---------------------- */

/* Row-wise LayerNorm over the last dimension with optional fused extras.  For input row r
 * (0 <= r < rows), with b = r / L_in and l = r % L_in:
 *     y = LN(in[r]) * gamma + beta                (gamma == NULL: y = in[r], no normalisation)
 *     y += add[(l % add_rows)]                     (add != NULL; fp32 [add_rows, d])
 *     dst = b * L_out + l_off + l
 *     out_f32[dst] = y ; out_bf16[dst] = bf16(y)   (each optional)
 *   stage emission (each pointer optional), rows split at l_split into an "A" part (l < l_split,
 *   video tokens) and a "B" part (text tokens of the joint sequence):
 *     rowA = b * strideA + l ,  rowB = b * strideB + (l - l_split)       (strides in rows)
 *     rawA_f32[rowA] / rawB_f32[rowB]   = y      (with raw_strideA / raw_strideB != 0: rows b * raw_stride + l ...)
 *     nrmA_bf16[rowA] / nrmB_bf16[rowB] = bf16(y / ||y||_2)  ; nrmA_f32 / nrmB_f32 = y / ||y||_2
 * in is fp32 (in_is_bf16 == 0) or bf16.  d % 128 == 0, d <= 1024.  eps = 1e-5 (torch default).
 * Replaces: LayerNorm at model/tfm_model.py:35,:37 and model/tan_model.py:155,:167,:174,:206,:233;
 * the positional add (:167,:199); torch.cat of video|text tokens (:201); torch.stack/permute of
 * stage features (:176,:208); `x / x.norm(dim=-1)` (:116-117,:136-137). */
typedef struct tan_ln_args {
  const void* in;
  int in_is_bf16;
  int rows;
  int d;
  const float* gamma;
  const float* beta;
  const float* add;
  int add_rows;
  int L_in;
  int L_out;
  int l_off;
  float* out_f32;
  void* out_bf16;
  int l_split;
  int64_t strideA;
  int64_t strideB;
  float* rawA_f32;
  float* rawB_f32;
  void* nrmA_bf16;
  void* nrmB_bf16;
  float* nrmA_f32;
  float* nrmB_f32;
  int64_t raw_strideA;   /* row strides of rawA_f32 / rawB_f32 when they differ from strideA / strideB (0: the same): */
  int64_t raw_strideB;   /* the training tape keeps raw features stage-major while the normalised ones keep the sink's layout */
} tan_ln_args;
TAN_API int tan_layernorm(const tan_ln_args* args, void* stream);

/* ---- multi-head attention core ------------------------------------------------------------------ */

/* out[b, i, h*64:(h+1)*64] = softmax_j(q_i . k_j / 8 + mask_j) @ v   for every clip b and head h.
 *   q [B*Lq, *] bf16 (ldq), k/v [B*Lk, *] bf16 (ldk / ldv): head h occupies columns [64h, 64h+64)
 *   relative to each base pointer (so a packed [M, 3d] in-projection output is passed as
 *   q = qkv, k = qkv + d, v = qkv + 2d, ldq = ldk = ldv = 3d);
 *   key_padding_mask [B, Lk] uint8 (1 = ignore key) or NULL; out [B*Lq, H*64] bf16 (ldo).
 * head_dim is fixed at 64 (the reference hard-codes width 512 / 8 heads, model/tan_model.py:43-46;
 * config 4 uses width 768 / 12 heads).  A row whose keys are all masked yields NaN, as
 * torch.softmax over all -inf does in the reference.
 * Replaces F.scaled_dot_product_attention + the [L,B,C]<->[B*H,L,hd] transposes reached from
 * nn.MultiheadAttention at model/tfm_model.py:32 (self-attention, Lq == Lk) and :80
 * (cross-attention of the unused decoder, Lq != Lk).
 * lse (optional, may be NULL): [B, H, pad64(Lq)] fp32, receives the log2-domain log-sum-exp of every query row's
 * scaled scores, log2(sum_j exp(q_i . k_j / 8)) (+inf in the padding rows Lq .. pad64(Lq)): what
 * tan_attention_bwd_bf16 needs from the forward pass (pad64(n) = n rounded up to a multiple of 64). */
TAN_API int tan_attention_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                               const uint8_t* key_padding_mask, void* out, int64_t ldo, int B, int H, int Lq, int Lk,
                               float* lse, void* stream);

/* ---- cosine-similarity matrix + MIL-NCE statistics ----------------------------------------------- */

/* Geometry shared by the similarity / NCE entry points.
 *   rows   r = (b * S + s) * T + t   for local clip b < B_loc, stage s < S, frame t < T
 *   cols   c = b' * N + n            for GLOBAL clip b' and sentence n;  C = B_glob * N
 *   z[r, c] = <vfeat[r], tfeat_s[c]> / 0.07                       (train/loss.py:65-67)
 *   positive(r, c) <=> b_off + b == b' and bit n of posbits[b][t] (below)
 * Sums use the fixed shift 1/0.07 (|cos| <= 1):  e = exp(z - 1/0.07). */
typedef struct tan_sim_geom {
  int B_loc;   /* local clips (rows) */
  int S;       /* stages */
  int T;       /* frames per clip */
  int C;       /* global text columns: B_glob * N, or col_off[B_glob] with ragged columns */
  int N;       /* sentences per clip (padded): bits per target row */
  int d;       /* feature width */
  int b_off;   /* global index of local clip 0 */
  /* Ragged ("compact") columns, optional (NULL = every clip owns N columns): device array [B_glob + 1] of prefix
   * offsets, clip b' owns columns [col_off[b'], col_off[b' + 1]) and its sentence n is column col_off[b'] + n.
   * The reference drops padded sentences before the loss (train/loss.py:235); with this layout they are never
   * computed at all (25 % of the columns at BASELINE's n_b ~ U[N/2, N]).  Not available with materialised logits
   * (logits_out / tan_nce_from_logits keep the [B, S, T, B, N] layout). */
  const int32_t* col_off;
} tan_sim_geom;

/* Targets.  posbits [B_loc, T, W] uint32, W = ceil(N / 32): bit (n % 32) of word n / 32 of posbits[b][t] is
 * set iff sentence n of LOCAL clip b is a positive of frame t.  Replaces the reference's float target tensor
 * [B, T, B, N] (train/loss.py:80-85: block-diagonal, 2.1 GB at BASELINE config 3); bits of padded sentences
 * and of n >= N must be 0.  The same format carries the interval targets of get_mask_from_time and the
 * self-labelled targets of train/loss.py:88-229.
 * tan_pos_from_time builds it from sentence times: bit = valid[b][n] && start[b][n] <= t < end[b][n]
 * (train/loss.py:26-41; start/end [B, N] fp32, valid [B, N] uint8 or NULL = all valid). */
TAN_API int tan_pos_from_time(const float* start, const float* end, const uint8_t* valid, int B, int T, int N,
                      uint32_t* posbits, void* stream);

/* Workspace (bytes) for the partial sums of tan_sim_nce_fwd / tan_nce_from_logits. */
TAN_API size_t tan_sim_nce_workspace_bytes(const tan_sim_geom* g);

/* Fused similarity GEMM + NCE statistics.
 *   vfeat [B_loc*S*T, d] bf16 L2-normalised video stage features (layout [B_loc, S, T, d]);
 *   tfeat bf16 L2-normalised text features: [C, d] shared by all stages (tfeat_stage_stride == 0,
 *         dual encoder, model/tan_model.py:118-119) or [S, C, d] (tfeat_stage_stride = C*d,
 *         joint encoder, :138-139);
 *   posbits as above; col_valid [C] uint8 (1 = real sentence, i.e. ~text_padding_mask, GLOBAL columns);
 *   row_kill [B_loc, T] uint8 or NULL: 1 = the own-clip entries of this frame count as exp(-inf) (the
 *         reference's in-place -6e4 fill of padded frames under --learn_agreement, train/loss.py:96-97);
 *   logits_out: NULL (fused mode, the matrix never leaves the SM) or bf16 [B_loc*S*T, C] (ld = C),
 *         written once = the reference's `logits_*` tensor [B,S,T,B,N] (cosines, NOT divided by 0.07);
 *   row_sums [2, B_loc*S*T] fp32: sum_c e (valid columns) and sum_{c positive} e;
 *   col_sums [2, S, C] fp32: sum_r e over the LOCAL rows (all rows, video padding is not applied
 *         to rows, train/loss.py:241-254) and over positive rows.
 * Replaces torch.einsum at model/tan_model.py:118-119,:138-139 and the logits passes of
 * train/loss.py:65-67,:241-254,:261-271. */
TAN_API int tan_sim_nce_fwd(const void* vfeat, const void* tfeat, int64_t tfeat_stage_stride,
                    const tan_sim_geom* g, const uint32_t* posbits, const uint8_t* col_valid,
                    const uint8_t* row_kill, void* logits_out, float* row_sums, float* col_sums,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Same statistics from MATERIALISED logits (API-preserving mode: a caller hands get_loss a plain
 * `logits_*` tensor).  logits [B_loc*S*T, C] bf16 (logits_is_f32 == 0) or fp32, unscaled cosines,
 * 16-byte aligned.  HBM-bound: one coalesced pass, register column accumulators, butterfly row sums. */
TAN_API int tan_nce_from_logits(const void* logits, int logits_is_f32, const tan_sim_geom* g,
                        const uint32_t* posbits, const uint8_t* col_valid, const uint8_t* row_kill,
                        float* row_sums, float* col_sums, void* workspace, size_t workspace_bytes,
                        void* stream);

/* Loss reduction (train/loss.py:248-256): from row_sums [2, R] (R = B_loc*S*T) and (globally reduced)
 * col_sums [2, S*C] accumulate
 *   out[0] += sum over rows with a positive of log(all) - log(pos),  out[1] += their count,
 *   out[2] += sum over valid columns with a positive of log(all) - log(pos), out[3] += their count.
 * Either side may be NULL (rows and columns may be reduced by separate calls / ranks).
 * row_sel [B_loc, T] / col_sel [C] uint8 (or NULL) further restrict the rows / columns that count: the
 * thresholded loss of train/loss.py:277-304 keeps the frames with a positive among the kept sentences and
 * the kept sentences only.  `out` [4] fp64 must be zeroed by the caller (fp64 so that the atomic accumulation
 * order is invisible at fp32 resolution). */
TAN_API int tan_nce_reduce(const float* row_sums, int64_t R, int S, int T, const uint8_t* row_sel,
                   const float* col_sums, int64_t SC, int C, const uint8_t* col_sel, double* out,
                   void* stream);

/* ---- own-clip similarity blocks, agreement self-labelling ----------------------------------------- */

/* out[b, j, t, n] = <vfeat[b, s_first + j, t], tfeat_(s)[b * N + n]>  for j < s_count: the own-clip (diagonal)
 * blocks of the [B,S,T,B,N] cosine matrix, fp32 [B, s_count, T, N].  vfeat [B*S*T, d] bf16; tfeat [B*N, d]
 * (tfeat_stage_stride == 0) or [S, B*N, d] (stride B*N*d) bf16; B counts the LOCAL clips and tfeat holds their
 * sentences.  d % 64 == 0.  Replaces torch.diagonal(logits, dim1=0, dim2=3) at train/loss.py:92-95,:150-153,
 * :280-283 (last stage) and the eval einsum 'bstc,b(s)kc->bstk' at model/tan_model.py:261-262,:280-281. */
TAN_API int tan_own_clip_sim(const void* vfeat, const void* tfeat, int64_t tfeat_stage_stride, int B, int S, int T,
                     int N, int d, int s_first, int s_count, float* out, void* stream);

/* Self-labelling scan of one model (train/loss.py:96-145): own [B, T, N] fp32 own-clip cosines of the last
 * stage; with z = own / 0.07 and -6e4 on padded frames / sentences: p = softmax over sentences, q = softmax
 * over time of p / 0.07; for every sentence the window of its original duration (popcount of its posbits
 * column, >= 1; padded sentences: none) with the best mean q, never frames 0 and T-1, first best wins.
 *   win [B, N, 2] int32 kept frame range [lo, hi) (lo == hi: none); mean_logit [B, N] = mean of z over it;
 *   max_logit [B, N] = max_t z (train/loss.py:280; padded frames filled only if fill_max != 0, which is the
 *   reference's in-place side effect in `init` mode).
 * video_padding_mask [B, T] uint8 or NULL, text_padding_mask [B, N] uint8 (1 = padded).
 * Replaces the [B,N,T,T] circulant box filter (2.1 GB at BASELINE config 5). */
TAN_API size_t tan_agree_scan_workspace_bytes(int B, int T, int N);
TAN_API int tan_agree_scan(const float* own, const uint32_t* posbits, const uint8_t* video_padding_mask,
                   const uint8_t* text_padding_mask, int B, int T, int N, int fill_max, int* win,
                   float* mean_logit, float* max_logit, void* workspace, size_t workspace_bytes, void* stream);

/* New target bits from the two models' windows (train/loss.py:196-226).  kind: 0 'i' (intersection where
 * replace[b,n]), 1 'u' (union where replace), 2 'keep' (union where replace, else the old timestamps),
 * 3 'keep-joint' (joint window where replace, else old); then at most one sentence per frame (the first)
 * and sentences left with no frame get their old timestamps back.  replace [B, N] uint8. */
TAN_API int tan_agree_targets(const uint32_t* old_posbits, const int* win_joint, const int* win_dual,
                      const uint8_t* replace, int B, int T, int N, int kind, uint32_t* new_posbits,
                      void* stream);

/* ---- backward pass (training step) ----------------------------------------------------------------
 * train/main.py:112 calls loss.backward(); in the reference every gradient is produced by torch autograd over the
 * call sites cited above.  Here the GEMM-shaped gradients (dgrad = dY @ W, wgrad = dY^T @ X, the similarity
 * recomputation and its two gradient products) are tan_linear_bf16 calls on transposed operands; the entry
 * points below are the rest.  Round-1 status: first correct path. */

/* out[c, r] = in[r, c] (bf16) for r < R; columns R <= r < R_pad of out are written as zero (contraction padding to
 * the GEMM's K % 64).  in [R, C] (ldi), out [C, R_pad] (ldo); C, R_pad, ldi, ldo even. */
TAN_API int tan_transpose_bf16(const void* in, int64_t ldi, void* out, int64_t ldo, int R, int C, int R_pad,
                               void* stream);

/* out[n] (+)= sum_m in[m, n]: bias gradients (autograd of the `+ bias` in F.linear).  in [M, N] bf16 or fp32 (ld),
 * out [N] fp32; deterministic two-stage reduction through `workspace`. */
TAN_API size_t tan_colsum_workspace_bytes(int M, int N);
TAN_API int tan_colsum(const void* in, int in_is_bf16, int64_t ld, int M, int N, float* out, int accumulate,
                       void* workspace, size_t workspace_bytes, void* stream);

/* QuickGELU on a stored pre-activation and its derivative (model/tfm_model.py:11-13): h = u * sigmoid(1.702 u);
 * du = dh * s (1 + 1.702 u (1 - s)), s = sigmoid(1.702 u).  bf16 [n], n % 8 == 0, 16-byte aligned. */
TAN_API int tan_quickgelu_fwd(const void* u, void* h, size_t n, void* stream);
TAN_API int tan_quickgelu_bwd(const void* dh, const void* u, void* du, size_t n, void* stream);

/* Backward of tan_layernorm's y = LayerNorm(x) * gamma + beta with the same row map (row r of x <-> row
 * (r / L_in) * L_out + l_off + r % L_in of dy):  dx[r] (+)= the LayerNorm input gradient (accumulate_dx != 0 adds
 * to dx: the residual stream's gradient), dgamma / dbeta [d] are ACCUMULATED (NULL: skipped).  x, dx [rows, d]
 * fp32, dy fp32.  Replaces autograd of nn.LayerNorm at model/tfm_model.py:31,:37, model/tan_model.py:155,:161-167,
 * :174,:187,:206,:233.  * dx_bf16 (optional, [rows, d] bf16): a bf16 copy of the UPDATED dx; dx_colsum (optional, [d] fp32, accumulated;
 * needs dgamma / dbeta): its column sums -- the operand and the bias gradient of the linear layer whose backward comes
 * next, produced here instead of by a cast pass and a column-sum pass over the residual-stream gradient. */
TAN_API size_t tan_layernorm_bwd_workspace_bytes(int rows, int d);
TAN_API int tan_layernorm_bwd(const float* dy, const float* x, const float* gamma, float* dx, int accumulate_dx,
                              int rows, int d, int L_in, int L_out, int l_off, float* dgamma, float* dbeta,
                              void* dx_bf16, float* dx_colsum, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of y = x / ||x|| (model/tan_model.py:116-117,:136-137): dst = (g - y <y, g>) / ||x||.  Row r of the raw
 * features x lies at (r / L_in) * src_stride + r % L_in, of the incoming gradient g at (r / L_in) * g_stride +
 * r % L_in; dst is token-major (row (r / L_in) * L_out + l_off + r % L_in); written, or added when
 * accumulate != 0.  fp32. */
TAN_API int tan_l2norm_bwd(const float* x, const float* g, float* dst, int accumulate, int rows, int d, int L_in,
                           int64_t src_stride, int64_t g_stride, int L_out, int l_off, void* stream);

/* out[l, :] (+)= sum_b in[b * L_out + l_off + l, :]: gradient of a positional table broadcast over the clips
 * (model/tan_model.py:161-167).  fp32. */
TAN_API int tan_batch_sum(const float* in, float* out, int B, int L, int d, int L_out, int l_off, int accumulate,
                          void* stream);

/* Gradient of the MIL-NCE loss with respect to a block of cosines (train/loss.py:231-275 differentiated).
 * z [Rc, ldz] fp32: cosines of rows r0 .. r0+Rc of ONE stage (row = b * T + t, local clips) against all C global
 * columns (g->C; g->S is ignored).  With e = exp((z - 1) / 0.07) on valid columns,
 *   G[r, c] = e (ra[r] + cb[c] - positive(r, c) (rap[r] + cbp[c])) / 0.07
 * ra / rap [B_loc * T]: weight / sum_all and weight / sum_pos of the row (0 when the row does not count),
 * cb / cbp [C] the same for the stage's columns.  Writes G [Rc, ldg] bf16 (columns up to ldg zero-filled) and
 * GT [C, ldgt] bf16 = G^T with columns Rc .. Rc_pad zero-filled: the operands of dA = G @ tfeat and
 * dB = G^T @ vfeat. */
TAN_API int tan_sim_grad_tiles(const float* z, int64_t ldz, int Rc, int Rc_pad, int r0, const tan_sim_geom* g,
                               const uint32_t* posbits, const uint8_t* col_valid, const uint8_t* row_kill,
                               const float* ra, const float* rap, const float* cb, const float* cbp, void* G,
                               int64_t ldg, void* GT, int64_t ldgt, void* stream);

/* The similarity recomputation fused with tan_sim_grad_tiles: G [Rc, ldg] bf16 (columns g->C .. C_pad are zero) =
 * d loss / d cos of rows r0 .. r0+Rc of one stage, computed in the epilogue of the tcgen05 pair GEMM
 * <vfeat[r], tfeat[c]> (vfeat [Rc, d] bf16 (ldv), tfeat [C_pad, d] bf16 (ldt), rows beyond g->C zero) -- the
 * fp32 cosines never reach HBM.  Same coefficient vectors / targets as tan_sim_grad_tiles; g->N <= 64,
 * C_pad % 128 == 0, d % 64 == 0.  dB = G^T @ vfeat then runs on tan_gemm_tn_bf16 straight from G (no transpose). */
TAN_API int tan_sim_grad_gemm(const void* vfeat, int64_t ldv, const void* tfeat, int64_t ldt, int Rc, int r0,
                              const tan_sim_geom* g, int C_pad, const uint32_t* posbits, const uint8_t* col_valid,
                              const uint8_t* row_kill, const float* ra, const float* rap, const float* cb,
                              const float* cbp, void* G, int64_t ldg, void* stream);

/* out[P, Q] (+)= A[R, P]^T @ B[R, Q]: bf16 row-major operands consumed as they lie in HBM (MN-major UMMA operands,
 * no transposes), the contraction running over their R rows; fp32 accumulation and output (row pitch ldo; added to
 * the existing contents when accumulate != 0).  Weight gradients dW = dY^T X (autograd of F.linear,
 * model/tfm_model.py:21,35-37, model/tan_model.py:155,:232) and the text-side similarity gradient dB = G^T V
 * (autograd of the einsum at model/tan_model.py:119,:139).  The contraction is split over the CTA pairs inside the
 * launch; partial tiles go through `workspace` (tan_gemm_tn_workspace_bytes; may be 0) and are summed in a fixed
 * order (deterministic).  P % 8 == 0, Q % 32 == 0, lda / ldb multiples of 8, ldo of 4. */
TAN_API size_t tan_gemm_tn_workspace_bytes(int R, int P, int Q);
TAN_API int tan_gemm_tn_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, int R, int P, int Q, float* out,
                             int64_t ldo, int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of tan_attention_bf16 (same operand conventions; o = the forward output, d_out its gradient) on tcgen05
 * tensor cores: writes dq [B*Lq, *], dk / dv [B*Lk, *] (bf16).  lse [B, H, pad64(Lq)] fp32 is the row statistic the
 * FORWARD call stored (16-byte aligned); delta [B, H, pad64(Lq)] fp32 is workspace (receives <d_out_i, o_i>).
 * Deterministic (two kernels, no atomics).  Replaces autograd of F.scaled_dot_product_attention reached from
 * model/tfm_model.py:32. */
TAN_API int tan_attention_bwd_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                   const void* o, int64_t ldo, const void* d_out, int64_t lddo,
                                   const uint8_t* key_padding_mask, void* dq, int64_t lddq, void* dk, int64_t lddk,
                                   void* dv, int64_t lddv, float* lse, float* delta, int B, int H, int Lq, int Lk,
                                   void* stream);

/* ---- text embedder (model/word2vec_model.py:76-102) ----------------------------- */

/* out[r, :] = table[ids[r], :] for r < n: rows of the (frozen) bf16 embedding table [V, ld] (nn.Embedding lookup at
 * model/word2vec_model.py:84-85; ld = 320 = the 300 word2vec dimensions zero-padded to a multiple of 64 so that the
 * result is the K-major A operand of fc1).  ids outside [0, V) read row 0 (the padding / unknown word). */
TAN_API int tan_embed_gather_bf16(const int64_t* ids, const void* table, int64_t ld, int V, int64_t n, void* out,
                                  void* stream);

/* pooled[s, f] = max over the 32 words w of sentence s of  keep(s, w) ? relu(x[s*32 + w, :] . w1[f, :] + b1[f]) : -6e4
 * (model/word2vec_model.py:86,:92-96): the fc1 GEMM on tcgen05 with bias + ReLU + masked max-pool fused into its
 * epilogue (an epilogue warp's 32 TMEM lanes are one sentence).  x [S*32, K] bf16 (ldx), w1 [F, K] bf16 (ldw),
 * b1 [F] fp32 or NULL, keep [S*32] uint8 attention mask (1 = keep) or NULL; a sentence whose words are all ignored
 * keeps all of them (:93).  pooled [S, F] bf16; argmax [S, F] uint8 (optional): the first word attaining the maximum,
 * for tan_text_pool_bwd.  F % 128 == 0, K % 64 == 0. */
TAN_API int tan_text_pool_fc1(const void* x, int64_t ldx, const void* w1, int64_t ldw, const float* b1,
                              const uint8_t* keep, int S, int F, int K, void* pooled, uint8_t* argmax, void* stream);

/* Backward of the pooling + ReLU: dH[s*32 + w, f] = (w == argmax[s, f] && pooled[s, f] > 0) ? dpool[s, f] : 0
 * (bf16 [S*32, F]), the operand of dW1 = dH^T x (tan_gemm_tn_bf16) and db1 (tan_colsum).  Autograd of
 * torch.max / F.relu at model/word2vec_model.py:86,:95. */
TAN_API int tan_text_pool_bwd(const void* dpool, const void* pooled, const uint8_t* argmax, int S, int F, void* dH,
                              void* stream);

/* ---- optimizer step (co-training step) ------------------------------------------ */

/* One parameter tensor of the multi-tensor optimizer step (all pointers device, fp32, numel elements). */
typedef struct {
  float* param;
  const float* grad;       /* NULL: no gradient this step, the tensor is skipped (as torch does) */
  float* exp_avg;
  float* exp_avg_sq;
  float* ema;              /* NULL or the EMA target copy: ema = m * ema + (1 - m) * param_new */
  int64_t numel;
  float decay;             /* 1 - lr * weight_decay of the tensor's parameter group */
  float neg_step_size;     /* -lr / (1 - beta1^step) */
} tan_optim_tensor;

/* Per-parameter gradient clipping (utils/train_utils.py:3-13: g *= min(1, clip / (||g|| + 1e-6)); clip_grad <= 0
 * disables), AdamW (torch.optim.AdamW as constructed at train/main.py:397, decoupled weight decay) and the EMA
 * update of the target network (model/tan_model.py:340-344) over ALL tensors of `table` [n_tensors] in two launches,
 * no host synchronisation.  Chunking (built by the caller): chunk j covers elements [chunk_start[j],
 * chunk_start[j] + chunk_elems) of tensor chunk_tensor[j]; the chunks of tensor t are tensor_first_chunk[t] ..
 * tensor_first_chunk[t + 1].  partial [n_chunks] fp32 workspace; norms_out [n_tensors] (optional) receives the
 * gradient norms (what clip_gradients returns); inv_scale (optional device scalar) multiplies the gradients first
 * (GradScaler.unscale_, train/main.py:114).  An unclipped, unscaled step is bit-identical to
 * torch.optim.AdamW(foreach=True) in fp32 (same operations, one rounding each; the hyper-parameters are doubles, as
 * Python hands them to torch, and are rounded to fp32 where the tensor operation consumes them). */
TAN_API int tan_optim_adamw_step(const tan_optim_tensor* table, int n_tensors, const int* chunk_tensor,
                                 const int64_t* chunk_start, const int* tensor_first_chunk, int n_chunks,
                                 int chunk_elems, double clip_grad, double beta1, double beta2, double eps, int step,
                                 double ema_m, const float* inv_scale, float* partial, float* norms_out,
                                 void* stream);

/* ema = m * ema + (1 - m) * param for every tensor of `table` with a non-NULL ema (model/tan_model.py:340-344,
 * TwinTemporalAligner._momentum_update), one launch. */
TAN_API int tan_ema_update(const tan_optim_tensor* table, const int* chunk_tensor, const int64_t* chunk_start,
                           int n_chunks, int chunk_elems, double m, void* stream);

/* Sliding-window alignment inference (eval/eval_zeroshot_align.py 'overlap-seq').  Stitching of :198-205: windows
 * [W, 4] int32 = (t0, t1, n0, n1) per window (device, 16-byte aligned); blk_joint / blk_dual [W, T, N] fp32 = the last
 * stage's own-clip cosine blocks of the batched windows (tan_own_clip_sim).  For every (sentence n, frame t):
 *     sim_x[n, t] = (sum over the windows covering (n, t), in window order, of blk_x[w, t - t0, n - n0] / 0.07)
 *                   / max(cover[n, t], 1e-5),      cover[n, t] = number of such windows
 * -- bit-identical to the reference's python loop of slice additions followed by the division.  Outputs [n_text, vlen]
 * fp32.  A long video comes in several batches of windows: accumulate != 0 continues the running (un-normalised) sums
 * and counts a previous call left in the outputs, finalize != 0 divides (the last batch). */
TAN_API int tan_align_stitch(const float* blk_joint, const float* blk_dual, const int* windows, int W, int T, int N,
                             float* sim_joint, float* sim_dual, float* cover, int n_text, int vlen, int accumulate,
                             int finalize, void* stream);

/* The decision of eval/eval_zeroshot_align.py:222-238: out[n] = argmax_t softmax_t(s[n, :]), s = sim where sim != 0,
 * -6e4 for uncovered entries; first index on ties.  sim [n_text, vlen] fp32, out [n_text] int64. */
TAN_API int tan_align_argmax(const float* sim, int n_text, int vlen, int64_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TAN_B200_H_ */
