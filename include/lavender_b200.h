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
#define LAV_ACT_GELU 1       /* out = gelu_erf(v); if aux != NULL also aux = v (f16 pre-activation)      */
#define LAV_ACT_GELU_BWD 2   /* out = v * gelu_erf'(aux)                                                  */

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
} LavGemmEpilogue;

int lav_gemm_f16(const void* A, int64_t lda, int a_major, const void* B, int64_t ldb, int b_major, int M, int N,
                 int K, const LavGemmEpilogue* epi, int split_k, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAVENDER_B200_H_ */
