// Token-embedding kernels of the text / video front ends (HBM-bound, one warp per row, fp32):
//   * BertEmbeddings forward: word + position + token-type gather, LayerNorm (HF BertEmbeddings through
//     EncTxt.forward, model.py:125-142) and its backward scatter-add into the three tables;
//   * EncVideo's token assembly (model.py:69-85): [emb_cls ; fc(swin)] + emb_pos + emb_len, LayerNorm, and the
//     reductions that give the gradients of emb_cls / emb_pos / emb_len.
#include "runtime.h"
#include "sm100.cuh"

namespace lav {

constexpr int kEmbThreads = 256;

static inline int emb_grid(int64_t rows) {
  int64_t ctas = (rows + 7) / 8;
  return (int)std::max<int64_t>(1, std::min<int64_t>(ctas, (int64_t)sm_count() * 16));
}

// LayerNorm of one row held as `n4` float4 per lane (C <= 32*4*MAXV); two-pass statistics like ln_fwd_kernel.
template <int MAXV>
__device__ __forceinline__ void ln_row(float4 (&v)[MAXV], int n4, int lane, int C, const float* gamma, const float* beta,
                                       float eps, float* y32, __half* y16, float* mean_out, float* rstd_out) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k)
    if (lane + 32 * k < n4) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  const float mean = warp_sum(s) / C;
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k)
    if (lane + 32 * k < n4) {
      float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
      ss += (a * a + b * b) + (c * c + d * d);
    }
  const float rstd = 1.0f / sqrtf(warp_sum(ss) / C + eps);
  if (lane == 0) {
    if (mean_out) *mean_out = mean;
    if (rstd_out) *rstd_out = rstd;
  }
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int i = lane + 32 * k;
    if (i < n4) {
      float4 ga = *reinterpret_cast<const float4*>(gamma + 4 * i);
      float4 be = *reinterpret_cast<const float4*>(beta + 4 * i);
      float4 o;
      o.x = (v[k].x - mean) * rstd * ga.x + be.x;
      o.y = (v[k].y - mean) * rstd * ga.y + be.y;
      o.z = (v[k].z - mean) * rstd * ga.z + be.z;
      o.w = (v[k].w - mean) * rstd * ga.w + be.w;
      if (y32) *reinterpret_cast<float4*>(y32 + 4 * i) = o;
      if (y16) *reinterpret_cast<uint2*>(y16 + 4 * i) = make_uint2(pack_half2(o.x, o.y), pack_half2(o.z, o.w));
    }
  }
}

constexpr int kMaxV = 8;  // C <= 1024

struct BertEmbParams {
  const int64_t* ids; const int64_t* pos_ids; const int64_t* type_ids;
  int rows, Lt, C, vocab, max_pos, n_types;
  const float* word; const float* pos; const float* type;
  const float* gamma; const float* beta; float eps;
  float* sum32; float* y32; float* mean; float* rstd;
};

__global__ void __launch_bounds__(kEmbThreads) bert_embed_ln_fwd_kernel(const BertEmbParams p) {
  const int lane = threadIdx.x & 31;
  const int n4 = p.C >> 2;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < p.rows; r += gridDim.x * 8) {
    int64_t w = p.ids[r];
    int64_t ps = p.pos_ids ? p.pos_ids[r] : (r % p.Lt);
    int64_t tt = p.type_ids ? p.type_ids[r] : 0;
    w = min(max(w, (int64_t)0), (int64_t)p.vocab - 1);  // ids are validated on the host; clamp keeps loads in-bounds
    ps = min(max(ps, (int64_t)0), (int64_t)p.max_pos - 1);
    tt = min(max(tt, (int64_t)0), (int64_t)p.n_types - 1);
    const float4* a = reinterpret_cast<const float4*>(p.word + w * p.C);
    const float4* b = reinterpret_cast<const float4*>(p.pos + ps * p.C);
    const float4* c = reinterpret_cast<const float4*>(p.type + tt * p.C);
    float4 v[kMaxV];
#pragma unroll
    for (int k = 0; k < kMaxV; ++k) {
      const int i = lane + 32 * k;
      if (i < n4) {
        float4 x = a[i], y = b[i], z = c[i];
        // same association as HF: (word + token_type) + position
        v[k] = make_float4((x.x + z.x) + y.x, (x.y + z.y) + y.y, (x.z + z.z) + y.z, (x.w + z.w) + y.w);
        if (p.sum32) *reinterpret_cast<float4*>(p.sum32 + (int64_t)r * p.C + 4 * i) = v[k];
      }
    }
    ln_row<kMaxV>(v, n4, lane, p.C, p.gamma, p.beta, p.eps, p.y32 + (int64_t)r * p.C, nullptr,
                  p.mean ? p.mean + r : nullptr, p.rstd ? p.rstd + r : nullptr);
  }
}

// d{word,pos,type}[index[r]] += d[r]   (fp32 atomics; a few hundred rows per step)
__global__ void __launch_bounds__(kEmbThreads)
bert_embed_bwd_kernel(const float* d, const int64_t* ids, const int64_t* pos_ids, const int64_t* type_ids, int rows,
                      int Lt, int C, int vocab, int max_pos, int n_types, float* dword, float* dpos, float* dtype,
                      int padding_idx) {
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8) {
    int64_t w = min(max(ids[r], (int64_t)0), (int64_t)vocab - 1);
    int64_t ps = pos_ids ? pos_ids[r] : (r % Lt);
    ps = min(max(ps, (int64_t)0), (int64_t)max_pos - 1);
    int64_t tt = type_ids ? type_ids[r] : 0;
    tt = min(max(tt, (int64_t)0), (int64_t)n_types - 1);
    for (int c = lane; c < C; c += 32) {
      const float g = d[(int64_t)r * C + c];
      if (dword && w != padding_idx) atomicAdd(dword + w * C + c, g);   // nn.Embedding(padding_idx): the pad row gets no gradient
      if (dpos) atomicAdd(dpos + ps * C + c, g);
      if (dtype) atomicAdd(dtype + tt * C + c, g);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// EncVideo token assembly.  Output row (b, t, s), s in [0, 1+hw):
//   v = (s == 0 ? emb_cls : feat[(b*T+t)*hw + s-1]) + emb_pos[s] + (odr_swap[b*T+t] ? emb_odr : emb_len[t])
//   sum32[row] = v ; y32[row] = LN(v)
// ---------------------------------------------------------------------------------------------------------
struct VidEmbParams {
  const float* feat; int64_t ldf;   // [B*T*hw, C]
  const float* emb_cls; const float* emb_pos; const float* emb_len; const float* emb_odr;
  const uint8_t* odr_swap;          // [B*T] or null
  int B, T, hw, C;
  const float* gamma; const float* beta; float eps;
  float* sum32; float* y32; float* mean; float* rstd;
};

__global__ void __launch_bounds__(kEmbThreads) vid_embed_ln_fwd_kernel(const VidEmbParams p) {
  const int lane = threadIdx.x & 31;
  const int n4 = p.C >> 2;
  const int S = 1 + p.hw;
  const int rows = p.B * p.T * S;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8) {
    const int bt = r / S, s = r - bt * S, t = bt % p.T;
    const float4* a = (s == 0) ? reinterpret_cast<const float4*>(p.emb_cls)
                               : reinterpret_cast<const float4*>(p.feat + ((int64_t)bt * p.hw + (s - 1)) * p.ldf);
    const float4* b = reinterpret_cast<const float4*>(p.emb_pos + (int64_t)s * p.C);
    const float4* c = (p.odr_swap && p.odr_swap[bt]) ? reinterpret_cast<const float4*>(p.emb_odr)
                                                      : reinterpret_cast<const float4*>(p.emb_len + (int64_t)t * p.C);
    float4 v[kMaxV];
#pragma unroll
    for (int k = 0; k < kMaxV; ++k) {
      const int i = lane + 32 * k;
      if (i < n4) {
        float4 x = a[i], y = b[i], z = c[i];
        v[k] = make_float4((x.x + y.x) + z.x, (x.y + y.y) + z.y, (x.z + y.z) + z.z, (x.w + y.w) + z.w);
        if (p.sum32) *reinterpret_cast<float4*>(p.sum32 + (int64_t)r * p.C + 4 * i) = v[k];
      }
    }
    ln_row<kMaxV>(v, n4, lane, p.C, p.gamma, p.beta, p.eps, p.y32 + (int64_t)r * p.C, nullptr,
                  p.mean ? p.mean + r : nullptr, p.rstd ? p.rstd + r : nullptr);
  }
}

// Backward of the assembly: d = gradient wrt sum32 rows [B*T*S, C].
//   dfeat16[(bt*hw + s-1)] = fp16(d[bt, s])  (s >= 1)       -> operand of EncVideo.fc's dgrad / wgrad
//   demb_pos[s] += sum_bt d ; demb_cls += sum_bt d[bt, 0] ; demb_len[t] / demb_odr += sum_{b, s} d
// One CTA per (s, 32-column slab): loops over bt, so emb_pos / emb_cls need no atomics; emb_len / emb_odr use atomics.
__global__ void __launch_bounds__(256)
vid_embed_bwd_kernel(const float* d, int B, int T, int hw, int C, const uint8_t* odr_swap, __half* dfeat16, int64_t lddf,
                     float* dfeat32, int64_t lddf32, float* demb_cls, float* demb_pos, float* demb_len, float* demb_odr) {
  const int S = 1 + hw;
  const int s = blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 columns x 8 bt lanes
  const int col = blockIdx.x * 32 + tx;
  __shared__ float sm[8][33];
  float acc = 0.f;
  if (col < C) {
    for (int bt = ty; bt < B * T; bt += 8) {
      const float g = d[((int64_t)bt * S + s) * C + col];
      acc += g;
      if (s > 0) {
        if (dfeat16) dfeat16[((int64_t)bt * hw + (s - 1)) * lddf + col] = __float2half_rn(g);
        if (dfeat32) dfeat32[((int64_t)bt * hw + (s - 1)) * lddf32 + col] = g;
      }
      const bool swap = odr_swap && odr_swap[bt];
      if (swap) {
        if (demb_odr) atomicAdd(demb_odr + col, g);
      } else if (demb_len) {
        atomicAdd(demb_len + (int64_t)(bt % T) * C + col, g);
      }
    }
  }
  sm[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && col < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) acc += sm[k][tx];
    if (demb_pos) atomicAdd(demb_pos + (int64_t)s * C + col, acc);  // atomics only for "+=" across calls
    if (s == 0 && demb_cls) atomicAdd(demb_cls + col, acc);
  }
}

}  // namespace lav

using namespace lav;

extern "C" int lav_bert_embed_ln_fwd(const int64_t* ids, const int64_t* pos_ids, const int64_t* type_ids, int rows,
                                     int Lt, int C, int vocab, int max_pos, int n_types, const float* word,
                                     const float* pos, const float* type, const float* gamma, const float* beta,
                                     float eps, float* sum32, float* y32, float* mean, float* rstd, void* stream) {
  LAV_REQUIRE(ids && word && pos && type && gamma && beta && y32, "lav_bert_embed_ln_fwd: null pointer");
  LAV_REQUIRE(C > 0 && (C % 4) == 0 && C <= 128 * kMaxV && Lt > 0, "lav_bert_embed_ln_fwd: need C%%4==0, C<=1024");
  if (rows <= 0) return LAV_OK;
  BertEmbParams p{ids, pos_ids, type_ids, rows, Lt, C, vocab, max_pos, n_types, word, pos, type,
                  gamma, beta, eps, sum32, y32, mean, rstd};
  bert_embed_ln_fwd_kernel<<<emb_grid(rows), kEmbThreads, 0, (cudaStream_t)stream>>>(p);
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}

extern "C" int lav_bert_embed_bwd(const float* dsum32, const int64_t* ids, const int64_t* pos_ids,
                                  const int64_t* type_ids, int rows, int Lt, int C, int vocab, int max_pos, int n_types,
                                  float* dword, float* dpos, float* dtype, int padding_idx, void* stream) {
  LAV_REQUIRE(dsum32 && ids && Lt > 0, "lav_bert_embed_bwd: null pointer");
  if (rows <= 0) return LAV_OK;
  bert_embed_bwd_kernel<<<emb_grid(rows), kEmbThreads, 0, (cudaStream_t)stream>>>(dsum32, ids, pos_ids, type_ids, rows, Lt, C,
                                                                                 vocab, max_pos, n_types, dword, dpos, dtype, padding_idx);
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}

extern "C" int lav_vid_embed_ln_fwd(const float* feat, int64_t ldf, const float* emb_cls, const float* emb_pos,
                                    const float* emb_len, const float* emb_odr, const uint8_t* odr_swap, int B, int T,
                                    int hw, int C, const float* gamma, const float* beta, float eps, float* sum32,
                                    float* y32, float* mean, float* rstd, void* stream) {
  LAV_REQUIRE(feat && emb_cls && emb_pos && emb_len && gamma && beta && y32, "lav_vid_embed_ln_fwd: null pointer");
  LAV_REQUIRE(!odr_swap || emb_odr, "lav_vid_embed_ln_fwd: odr_swap needs emb_odr");
  LAV_REQUIRE(C > 0 && (C % 4) == 0 && C <= 128 * kMaxV && (ldf % 4) == 0, "lav_vid_embed_ln_fwd: need C%%4==0, C<=1024");
  const int64_t rows = (int64_t)B * T * (1 + hw);
  if (rows <= 0) return LAV_OK;
  VidEmbParams p{feat, ldf, emb_cls, emb_pos, emb_len, emb_odr, odr_swap, B, T, hw, C, gamma, beta, eps,
                 sum32, y32, mean, rstd};
  vid_embed_ln_fwd_kernel<<<emb_grid(rows), kEmbThreads, 0, (cudaStream_t)stream>>>(p);
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}

extern "C" int lav_vid_embed_bwd(const float* dsum32, int B, int T, int hw, int C, const uint8_t* odr_swap,
                                 void* dfeat16, int64_t lddf16, float* dfeat32, int64_t lddf32, float* demb_cls,
                                 float* demb_pos, float* demb_len, float* demb_odr, void* stream) {
  LAV_REQUIRE(dsum32, "lav_vid_embed_bwd: null pointer");
  if (B <= 0 || T <= 0) return LAV_OK;
  dim3 grid((C + 31) / 32, 1 + hw);
  vid_embed_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dsum32, B, T, hw, C, odr_swap, (__half*)dfeat16, lddf16,
                                                              dfeat32, lddf32, demb_cls, demb_pos, demb_len, demb_odr);
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}
