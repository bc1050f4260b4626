/*
 * lavender_b200 — C ABI of the B200-native (sm_100a) kernels behind LAVENDER's data-parallel
 * forward/backward hot path.
 *
 * The reference (microsoft/LAVENDER) is pure Python/PyTorch: it has NO FFI / plugin / operator interface
 * (SURVEY.md §8b).  Every entry point below therefore replaces an implicit ATen/cuBLAS/cuDNN call site of
 * the reference; the call site is cited per function as `file:line` relative to the reference tree.
 * The Python host modules under lavender_b200/ bind these symbols with ctypes.
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs; no torch / C++ types cross the boundary.
 *   - the CALLER owns every buffer (inputs, outputs, workspaces); the library never allocates or frees
 *     device memory and keeps no pointer after a call returns.
 *   - all work is enqueued on the `stream` argument (a cudaStream_t passed as void*); no call synchronises.
 *   - return value: 0 on success, negative LAV_E_* otherwise; lav_last_error() gives the message of the
 *     last failure on the calling thread.  Nothing throws across the boundary.
 *   - "f16" tensors are IEEE binary16 (the reference's GPU path is fp16 autocast: agent.py:219,
 *     utils/deepspeed.py:21-24); accumulation, LayerNorm/softmax statistics, residual stream, losses and
 *     parameter gradients are fp32.
 *   - row-major everywhere; `ld*` are leading dimensions in ELEMENTS.
 */
#ifndef LAVENDER_B200_H_
#define LAVENDER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAV_OK 0
#define LAV_E_INVALID (-1)  /* bad argument / unsupported shape */
#define LAV_E_CUDA (-2)     /* CUDA runtime / driver error      */
#define LAV_E_NO_DEVICE (-3)

/* ---- library ------------------------------------------------------------------------------------ */
int lav_abi_version(void);
const char* lav_last_error(void);
/* number of kernels this library has launched in this process (bench.py's `gpu_launches`) */
int64_t lav_launch_count(void);
/* queries device 0..n: fills sm_count; returns LAV_E_NO_DEVICE when no sm_100 device is present */
int lav_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/* Profiling aid: a caller-owned device buffer of n_u64 uint64 that the instrumented attention kernels fill with clock64()
 * stamps (8 per CTA: tools/trace_attn.py); NULL switches it off (the default).  Not used on the hot path. */
int lav_debug_set_trace(void* buf, int64_t n_u64);

/* ---- dropout ---------------------------------------------------------------------------------------
 * HF BERT's nn.Dropout sites (hidden_dropout_prob / attention_probs_dropout_prob = 0.1, active in train() mode as
 * Agent_Pretrain_MLM.step runs it, main_pretrain_mlm.py:147) are folded into the kernels that produce the dropped
 * tensor.  A mask bit is a pure function of (rng[0] = seed, rng[1] = step, site, row, column[, head]) — Philox-4x32-7,
 * 16 random bits per element — so backward kernels regenerate it and nothing mask-shaped is stored.  `rng` is a
 * DEVICE pointer (two uint64), read when the kernel runs: a CUDA-graph replay sees the step counter the host (or a
 * captured increment) wrote before it.  p == 0 or a NULL struct disables the site.  Kept elements are scaled by
 * 1 / (1 - round(p * 65536) / 65536). */
typedef struct LavDropout {
  const uint64_t* rng;     /* device: [seed, step]                                                        */
  uint32_t site;           /* distinct per dropout call site (and per call of the same module in a step)  */
  float p;                 /* drop probability                                                            */
} LavDropout;

/* out[r, c] = x[r, c] * keep(r, c) / (1 - p)  on fp32 [rows, C] (C % 8 == 0; in place allowed).  BertEmbeddings.dropout
 * forward, and its backward on the incoming gradient. */
int lav_dropout_f32(const float* x, int64_t ldx, float* out, int64_t ldo, int rows, int C, const LavDropout* drop,
                    void* stream);
/* Test helper: keep[r, c] (uint8 0/1) of the elementwise sites (head < 0: index (r, c)) or of the attention
 * probabilities of head `head` (index (query row r, key column c, head)). */
int lav_dropout_mask(uint8_t* keep, int rows, int C, int head, const LavDropout* drop, void* stream);

/* ---- GEMM (tcgen05 + TMA) -------------------------------------------------------------------------
 * D[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) )
 * Replaces every nn.Linear / Conv3d-as-GEMM call on the path: video_swin.py:74,77 (Mlp), :147,168
 * (qkv/proj), :285 (PatchMerging.reduction), :398 (PatchEmbed3D.proj), model.py:49 (EncVideo.fc), HF
 * BertSelfAttention/BertSelfOutput/BertIntermediate/BertOutput linears (model.py:242) and
 * BertLMPredictionHead (main_pretrain_mlm.py:69,115) — forward, dgrad and wgrad.
 *
 * Operand storage ("major"): LAV_MAJOR_K  : X is [rows(M or N), K] with K contiguous  (ldx >= K)
 *                            LAV_MAJOR_MN : X is [K, rows(M or N)] with rows contiguous (ldx >= rows)
 *   forward  Y = X W^T      : A = X  (K-major),  B = W  (K-major)
 *   dgrad    dX = dY W      : A = dY (K-major),  B = W  (MN-major: [N_out(k), K_in(n)])
 *   wgrad    dW = dY^T X    : A = dY (MN-major), B = X  (MN-major), contraction over tokens, split_k > 1
 * All leading dimensions must be multiples of 8 elements (16 bytes) and base pointers 16-byte aligned. */
#define LAV_MAJOR_K 0
#define LAV_MAJOR_MN 1

#define LAV_ACT_NONE 0
#define LAV_ACT_GELU 1       /* out = gelu_erf(v); if aux != NULL also aux = gelu_erf'(v) (f16), saved for backward */
#define LAV_ACT_GELU_BWD 2   /* out = v * aux, aux = the gelu_erf'(pre-activation) the forward epilogue saved   */

#define LAV_OUT_F16 0
#define LAV_OUT_F32 1

#define LAV_STORE 0          /* out  = value                                                              */
#define LAV_ACCUMULATE 1     /* out += value (fp32 out; atomic when split_k > 1)                          */

typedef struct LavGemmEpilogue {
  void* out;               /* [M, ldo] f16 or f32 (row r is written to row row_map[r] when row_map set)   */
  int64_t ldo;
  int32_t out_dtype;       /* LAV_OUT_F16 / LAV_OUT_F32                                                   */
  int32_t act;             /* LAV_ACT_*                                                                   */
  const float* bias;       /* [N] or NULL; added before the activation                                    */
  void* aux;               /* f16 [M, ldaux]; see LAV_ACT_*                                               */
  int64_t ldaux;
  const float* residual;   /* fp32 [*, ldres] or NULL: out = residual + row_scale * value (same row map)  */
  int64_t ldres;
  const int32_t* row_map;  /* [M] destination row per GEMM row, or NULL (identity)                        */
  const float* row_scale;  /* [ceil(M / rows_per_scale)] or NULL: DropPath keep/keep_prob per sample      */
  int32_t rows_per_scale;
  float alpha;
  int32_t accumulate;      /* LAV_STORE / LAV_ACCUMULATE                                                  */
  int32_t reserved;
  float* bias_grad;        /* wgrad only (A MN-major, LAV_ACCUMULATE): bias_grad[m] += alpha * sum_k A(m,k), i.e. the
                              column sum of dY = the nn.Linear bias gradient, from one extra N=16 MMA per k-step    */
  LavDropout drop;         /* dropout of (value + bias) before row_scale / residual (BertSelfOutput / BertOutput
                              .dropout, element index = (GEMM row, column)); drop.p == 0 disables           */
} LavGemmEpilogue;

int lav_gemm_f16(const void* A, int64_t lda, int a_major, const void* B, int64_t ldb, int b_major, int M, int N,
                 int K, const LavGemmEpilogue* epi, int split_k, void* stream);


/* ---- LayerNorm (one warp per row, fp32 statistics) -------------------------------------------------
 * Forward of nn.LayerNorm at video_swin.py:209 (norm1), :246 (norm2), :284 (PatchMerging.norm), :402, :477,
 * model.py:85 and the HF BERT LayerNorms.  Output row r (width G*C) is the concatenation of G source rows
 * of width C taken at row_map[r*G+g] (identity when row_map == NULL).  G=1 with a row map folds
 * torch.roll + window_partition (video_swin.py:218-227) into the load; G=4 folds the PatchMerging
 * gather/concat (video_swin.py:278-282).  Writes fp16 and/or fp32 outputs and the per-row mean / rstd. */
int lav_layernorm_fwd(const float* x, int64_t ldx, const int32_t* row_map, int G, int C, const float* gamma,
                      const float* beta, float eps, void* y16, int64_t ldy16, float* y32, int64_t ldy32,
                      float* mean, float* rstd, int rows, void* stream);

/* Backward of the above (autograd of nn.LayerNorm in the reference).  dy is fp16 (dy_is_f32=0) or fp32.
 *   dx32[src row] = LN'(dy) (+ add32[src row])      — scatter through the same row map
 *   dx16[r]       = the same value as fp16, row r     — operand of the following dgrad / wgrad GEMMs
 *   dgamma / dbeta (fp32, accumulated with atomics; both or neither). */
int lav_layernorm_bwd(const void* dy, int64_t lddy, int dy_is_f32, const float* x, int64_t ldx,
                      const int32_t* row_map, int G, int C, const float* gamma, const float* mean,
                      const float* rstd, const float* add32, int64_t ldadd, float* dx32, int64_t lddx32,
                      void* dx16, int64_t lddx16, float* dgamma, float* dbeta, float* param_ws, int64_t ws_floats,
                      int rows, const LavDropout* drop16, void* stream);
/* Same, with the gradient cast that follows the LayerNorm backward in the Swin block folded in (the scale_cast of
 * video_swin.py's backward: DropPath scale, :46-54, and the roll + window_partition gather, :218-227, of the gradient):
 *   dx16[dst row] = fp16( dx16_row_scale[src row / dx16_rows_per_scale] * value ),
 *   dst row = dx16_row_map[src row] if given, else the src row if dx16_at_src, else r.  G must be 1. */
int lav_layernorm_bwd_ex(const void* dy, int64_t lddy, int dy_is_f32, const float* x, int64_t ldx,
                         const int32_t* row_map, int G, int C, const float* gamma, const float* mean,
                         const float* rstd, const float* add32, int64_t ldadd, float* dx32, int64_t lddx32,
                         void* dx16, int64_t lddx16, const int32_t* dx16_row_map, int dx16_at_src,
                         const float* dx16_row_scale, int dx16_rows_per_scale, float* dgamma, float* dbeta,
                         float* param_ws, int64_t ws_floats, int rows, const LavDropout* drop16, void* stream);
/* param_ws (may be NULL): caller-owned scratch of ws_floats fp32 (>= 8 * SM count * 2 * C is always enough) for the
 * per-block partial sums of dgamma / dbeta; without it the blocks add to dgamma / dbeta with atomics. */
/* drop16 (may be NULL): dx16 is additionally multiplied by the dropout mask of site drop16 at (r, column) — the
 * gradient entering a dense layer whose output was dropped in the forward epilogue — while dx32 stays unmasked
 * (the residual path). */

/* out16[r, 0:C] = fp16( alpha * row_scale[r / rows_per_scale] * x[row_map[r], 0:C] )
 * (gradient gather for window-major GEMM operands; DropPath scale of video_swin.py:46-54 in backward). */
int lav_scale_cast_f16(const float* x, int64_t ldx, const int32_t* row_map, const float* row_scale,
                       int rows_per_scale, float alpha, void* out16, int64_t ldo, int rows, int C, void* stream);

/* High-precision mode (LAV_PRECISION=high; the stated parity mode): out16[r] = [hi | lo | hi] (mode 0,
 * the A operand) or [hi | hi | lo] (mode 1, the B operand) of x[r, 0:C] with hi = fp16(x), lo = fp16(x - hi), so that
 * one lav_gemm_f16 call over K' = 3C computes the split product Ah*Bh + Al*Bh + Ah*Bl (~2^-22 relative operand error)
 * on the same tcgen05 kernel.  Replaces nothing in the reference: its fp32 CPU path multiplies fp32 operands directly
 * (e.g. video_swin.py:147 nn.Linear); this restores that operand precision on the fp16 tensor-core path. */
int lav_split3_f16(const float* x, int64_t ldx, void* out16, int64_t ldo, int rows, int C, int mode, void* stream);

/* flat fp32 -> fp16 copy (per-step fp16 shadow of the fp32 master parameters; autocast's weight cast,
 * agent.py:219) */
int lav_cast_f32_to_f16(const float* src, void* dst, int64_t n, void* stream);

/* out16 = dy16 * dgelu16, flat fp16, dgelu16 = gelu_erf'(pre-activation) as saved by the LAV_ACT_GELU epilogue (backward of
 * BertPredictionHeadTransform's GELU, main_pretrain_mlm.py:46-48; the other GELUs are fused into GEMM epilogues) */
int lav_gelu_bwd_f16(const void* dy16, const void* dgelu16, void* out16, int64_t n, void* stream);

/* out[c] += alpha * sum_r x16[r, c]  (bias gradients of every nn.Linear on the path) */
int lav_colsum_f16(const void* x16, int64_t ld, int rows, int N, float* out, float alpha, void* stream);

/* ---- fused attention (tcgen05 + TMA) ---------------------------------------------------------------
 * O = softmax(scale * Q K^T + bias) V for `nprob` independent problems of L tokens each whose rows are
 * contiguous in the fused QKV activation [rows_total, ld] (Q/K/V of head h at columns *_off + h*head_dim).
 *   - WindowAttention3D.forward video_swin.py:147-167: head_dim 32, L = 245 (<= 256); `bias16` is the dense
 *     [ncls][nheads][256][256] fp16 tensor of lav_relpos_bias_expand built with inv_scale = 1 / scale (relative
 *     position bias :153-155 + shift mask :157-160 + masked padded keys); problem p uses class
 *     prob_class[p % class_period] (ncls <= 8).
 *   - HF BertSelfAttention (model.py:242): head_dim 64, any L (blocked online-softmax kernel, 128-key blocks);
 *     key_bias is the additive [nprob][NPk] fp32 row (0 for kept keys, -inf for masked / padded keys; NPk >= L
 *     rounded up to 128) of get_extended_attention_mask (model.py:239).
 *   - windows longer than 256 tokens (8 x 12 x 12 -> 720 at 384^2) use the same blocked kernel with the dense bias.
 * Writes O (fp16, [rows_total, ldo], head h at column h*head_dim) and lse[h][row] = log-sum-exp (fp32). */
int lav_attn_fwd_f16(const void* qkv, int64_t ld, int64_t rows_total, int q_off, int k_off, int v_off, int head_dim,
                     int nheads, int nprob, int L, float scale, const void* bias16, int NPb,
                     const int32_t* prob_class, int class_period, const float* key_bias, int NPk, int causal_from,
                     void* out16, int64_t ldo, float* lse, const LavDropout* drop, void* stream);
/* Same as lav_attn_fwd_f16 with an optional fp32 copy of O (out32, [rows_total, ldo32]; may be NULL; out16 may then be
 * NULL too) - the high-precision mode feeds the un-rounded O to the output projection (video_swin.py:168, HF
 * BertSelfOutput.dense). */
int lav_attn_fwd_ex(const void* qkv, int64_t ld, int64_t rows_total, int q_off, int k_off, int v_off, int head_dim,
                    int nheads, int nprob, int L, float scale, const void* bias16, int NPb,
                    const int32_t* prob_class, int class_period, const float* key_bias, int NPk, int causal_from,
                    void* out16, int64_t ldo, float* out32, int64_t ldo32, float* lse, const LavDropout* drop,
                    void* stream);
/* causal_from (-1: off): the seq2seq mask of LAVENDER_Base.get_attn_mask (model.py:208-218) without materialising
 * [B, L, L]: keys j >= causal_from (the text part) are visible to queries i >= j only — causal among the text tokens,
 * invisible to the video / prefix queries; keys j < causal_from follow key_bias. */
/* drop (may be NULL): dropout of the attention probabilities (HF BertSelfAttention.dropout), element index
 * (global query row, key column, head); lse is that of the un-dropped softmax. */

/* Backward of the above.  delta_ws: caller-owned fp32 [nheads][rows_total] workspace (rowsum(dO * O), filled by a
 * small pre-pass).  dq_acc: fp32 [rows_total, nheads*head_dim] scratch, zeroed by the pre-pass (dQ is reduced
 * over key chunks with atomics); dK and dV are written as fp16 into dqkv16 at k_off / v_off.  When ds16 is
 * given ([nprob][nheads][NPs][NPs] fp16) the gradient wrt the pre-softmax logits is stored for
 * lav_relpos_bias_grad. */
int lav_attn_bwd_f16(const void* qkv, int64_t ld, int64_t rows_total, int q_off, int k_off, int v_off, int head_dim,
                     int nheads, int nprob, int L, float scale, const void* bias16, int NPb,
                     const int32_t* prob_class, int class_period, const float* key_bias, int NPk, int causal_from,
                     const void* out16, int64_t ldo, const void* dout16, int64_t lddo, const float* lse,
                     float* delta_ws, float* dq_acc, int64_t lddq, void* dqkv16, int64_t lddqkv, void* ds16, int NPs,
                     const LavDropout* drop, void* stream);

/* dense16[cls][h][i][j] = inv_scale * (table[rel_index[i*L+j]][h] + (labels[cls][i] != labels[cls][j] ? -100 : 0)),
 * and -30000 for the padded key columns j >= L (video_swin.py:153-160 and compute_mask :290-305); labels may be NULL
 * (unshifted block, ncls = 1).  inv_scale = 1 / (softmax scale passed to lav_attn_*_f16): the attention kernels add
 * the tile to the raw Q K^T accumulator with a tensor-core identity product and scale the sum. */
int lav_relpos_bias_expand(const float* table, int nheads, const int32_t* rel_index, int L, const uint8_t* labels,
                           int ncls, void* dense16, int NP, float inv_scale, void* stream);

/* dtable[rel_index[i*L+j]][h] += sum_p ds16[p][h][i][j]  (gradient of relative_position_bias_table) */
int lav_relpos_bias_grad(const void* ds16, int nprob, int nheads, int NP, int L, const int32_t* rel_index,
                         float* dtable, void* stream);

/* ---- token embeddings (HBM-bound row kernels) ------------------------------------------------------
 * HF BertEmbeddings forward as reached through EncTxt.forward (model.py:125-142, txt_backbone_embed_only):
 *   sum32[r] = word[ids[r]] + type[type_ids ? type_ids[r] : 0] + pos[pos_ids ? pos_ids[r] : r % Lt]
 *   y32[r]   = LayerNorm(sum32[r]) (eps 1e-12), mean / rstd saved for lav_layernorm_bwd. */
int lav_bert_embed_ln_fwd(const int64_t* ids, const int64_t* pos_ids, const int64_t* type_ids, int rows, int Lt, int C,
                          int vocab, int max_pos, int n_types, const float* word, const float* pos, const float* type,
                          const float* gamma, const float* beta, float eps, float* sum32, float* y32, float* mean,
                          float* rstd, void* stream);
/* d{word,pos,type}[index[r]] += dsum32[r]  (autograd of the three nn.Embedding lookups; fp32 atomics).  Rows whose id
 * equals padding_idx (HF: word_embeddings has padding_idx = pad_token_id = 0; pass -1 for none) add nothing to dword. */
int lav_bert_embed_bwd(const float* dsum32, const int64_t* ids, const int64_t* pos_ids, const int64_t* type_ids,
                       int rows, int Lt, int C, int vocab, int max_pos, int n_types, float* dword, float* dpos,
                       float* dtype, int padding_idx, void* stream);

/* EncVideo.forward token assembly, model.py:69-85: row (b,t,s), s in [0, 1+hw):
 *   sum32 = (s == 0 ? emb_cls : feat[(b*T+t)*hw + s-1]) + emb_pos[s] + (odr_swap[b*T+t] ? emb_odr : emb_len[t])
 *   y32   = LayerNorm(sum32) (eps 1e-5).  odr_swap (uint8 [B*T], may be NULL) marks the frames whose emb_len is
 *   replaced by emb_odr (model.py:72-81). */
int lav_vid_embed_ln_fwd(const float* feat, int64_t ldf, const float* emb_cls, const float* emb_pos,
                         const float* emb_len, const float* emb_odr, const uint8_t* odr_swap, int B, int T, int hw,
                         int C, const float* gamma, const float* beta, float eps, float* sum32, float* y32,
                         float* mean, float* rstd, void* stream);
/* Backward of the assembly from dsum32 [B*T*(1+hw), C]: dfeat (fp16 and/or fp32, rows s >= 1) and
 * demb_cls / demb_pos / demb_len / demb_odr (+=, any may be NULL). */
int lav_vid_embed_bwd(const float* dsum32, int B, int T, int hw, int C, const uint8_t* odr_swap, void* dfeat16,
                      int64_t lddf16, float* dfeat32, int64_t lddf32, float* demb_cls, float* demb_pos,
                      float* demb_len, float* demb_odr, void* stream);

/* ---- cross entropy over the vocabulary -------------------------------------------------------------
 * nn.CrossEntropyLoss(ignore_index) of agent.py:73 on the MLM / VTM logits (main_pretrain_mlm.py:158-163).
 * Forward: row_lse[r] = logsumexp(logits[r, :V]); loss_sum += sum over labelled rows of (lse - logit[label]);
 * count += number of labelled rows (both fp32 device scalars, zeroed by the caller; loss = loss_sum / count). */
int lav_xent_fwd(const float* logits, int64_t ld, const int64_t* labels, int rows, int V, int64_t ignore_index,
                 float* row_lse, float* row_loss, float* loss_sum, float* count, void* stream);
/* Backward: d[r,c] = (*gout / *count) * (softmax(logits[r])[c] - [c == label[r]]) on labelled rows, 0 elsewhere;
 * written as fp32 (d32) and/or fp16 (d16, padding columns up to a multiple of 8 zeroed). */
int lav_xent_bwd(const float* logits, int64_t ld, const int64_t* labels, int rows, int V, int64_t ignore_index,
                 const float* row_lse, const float* gout, const float* count, float* d32, int64_t ldd32, void* d16,
                 int64_t ldd16, void* stream);

/* ---- GPU input pipeline ------------------------------------------------------------
 * Per-frame transform of the reference's loader (dataset.py:118-175: Resize(size_img) -> RandomCrop / CenterCrop ->
 * ToTensor -> Normalize on PIL images) on decoded uint8 RGB frames already in device memory:
 *   src  [T][Hs][Ws][3] uint8 (frames `frame_stride` bytes apart), resized to Hr x Wr with Pillow's two-pass antialiased
 *   bilinear resample (bit-exact: 22-bit fixed-point weights, uint8 rounding after each pass), cropped to the S x S window
 *   at (top, left), scaled by 1/255 and normalised:  out[t][c][y][x] = (pixel / 255 - mean[c]) / std[c]   (fp32 [T][3][S][S]).
 * mean / std are HOST pointers to 3 floats.  JPEG entropy decoding stays on the host (cv2 / PIL as in dataset.py:177-186). */
int lav_frames_resize_crop_norm_u8(const uint8_t* src, int T, int Hs, int Ws, int64_t frame_stride, int Hr, int Wr, int S,
                                   int top, int left, const float* mean, const float* stdv, float* out, void* stream);

/* ---- fused optimizer step on the flat arena ----------------------------------------
 * Device-side restatement of Agent_Base.backward_step (agent.py:240-250): GradScaler.unscale_ with the inf check,
 * clip_grad_norm_(max_grad_norm), AdamW (agent.py:137-139: betas (0.9, 0.98), decoupled weight decay, per-group lr /
 * weight decay from the name rules of agent.py:98-119) and GradScaler.update, without a host round trip.
 * `state` is a caller-owned fp32[16] device array: [0] loss scale, [1] growth tracker, [2] optimizer step count,
 * [3] sum of squares and [4] non-finite count of the (scaled) gradient — both accumulated by lav_grad_stats and
 * cleared by lav_adamw_step —, [5] unscaled gradient norm (out), [6] found_inf (out), [7..9] internal. */
int lav_grad_stats(const float* grad, int64_t n, float* state, float* ws, int64_t ws_floats, void* stream);
/* ws (may be NULL): caller-owned fp32 scratch of ws_floats >= 8 * SM count + 1 elements, ZEROED ONCE by the caller: per-block
 * partial sums + a ticket counter.  With it the sum of squares is reduced in a fixed order (bit-identical on every
 * data-parallel rank, so replicas stay bit-identical after the clip); without it the blocks use atomicAdd. */
/* group_of_block[i] = parameter group (0..ngroups-1) of elements [8i, 8i+8), 255 = no gradient (skipped);
 * group_lr / group_wd are fp32 device arrays indexed by group.  The update is skipped when found_inf. */
int lav_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                   const uint8_t* group_of_block, const float* group_lr, const float* group_wd, float beta1, float beta2,
                   float eps, float max_grad_norm, float* state, float growth_factor, float backoff_factor,
                   int growth_interval, void* param16, float* group_step0, float* group_bc, int ngroups,
                   void* stream);
/* group_step0 (fp32 [ngroups], may be NULL = all 0): number of non-skipped optimizer steps taken before the group's
 * parameters got their first gradient (a NEGATIVE entry is replaced by the kernel with the current count on the group's
 * first step) — torch.optim.AdamW keeps a per-parameter step that only advances when the parameter has a gradient,
 * so late starters (emb_odr, task heads) are bias-corrected with step - step0; group_bc: fp32 [2 * ngroups] scratch. */
/* param16 (may be NULL): f16 [n] shadow of `param`, rewritten for every updated element (the tensor-core operand copy
 * of the weights; saves the per-step lav_cast_f32_to_f16 pass over the arena). */

#ifdef __cplusplus
}
#endif
#endif /* LAVENDER_B200_H_ */
