// Cross-entropy over the vocabulary (nn.CrossEntropyLoss(ignore_index=-1) at agent.py:73 applied to the MLM / VTM
// logits, main_pretrain_mlm.py:158-163).  HBM-bound: one CTA per logit row, online log-sum-exp in one pass.
#include "runtime.h"
#include "sm100.cuh"

namespace lav {

constexpr int kXentThreads = 256;

__device__ __forceinline__ void lse_combine(float& m, float& s, float m2, float s2) {
  const float mm = fmaxf(m, m2);
  if (mm == -INFINITY) { m = mm, s = 0.f; return; }
  s = s * expf(m - mm) + s2 * expf(m2 - mm);
  m = mm;
}

// row_lse[r] = logsumexp(logits[r, :V]); row_loss[r] = lse - logits[r, label] (0 for ignored rows);
// loss_sum += sum of row losses, count += number of labelled rows.
__global__ void __launch_bounds__(kXentThreads)
xent_fwd_kernel(const float* logits, int64_t ld, const int64_t* labels, int V, int64_t ignore_index, float* row_lse,
                float* row_loss, float* loss_sum, float* count) {
  const int r = blockIdx.x;
  const float* x = logits + (int64_t)r * ld;
  float m = -INFINITY, s = 0.f;
  const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  if (vec) {
    const int n4 = V >> 2;
    for (int i = threadIdx.x; i < n4; i += kXentThreads) {
      float4 v = reinterpret_cast<const float4*>(x)[i];
      const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
      if (mx > m) { s *= expf(m - mx); m = mx; }
      s += (expf(v.x - m) + expf(v.y - m)) + (expf(v.z - m) + expf(v.w - m));
    }
    for (int c = (n4 << 2) + threadIdx.x; c < V; c += kXentThreads) {
      const float v = x[c];
      if (v > m) { s *= expf(m - v); m = v; }
      s += expf(v - m);
    }
  } else {
    for (int c = threadIdx.x; c < V; c += kXentThreads) {
      const float v = x[c];
      if (v > m) { s *= expf(m - v); m = v; }
      s += expf(v - m);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    lse_combine(m, s, m2, s2);
  }
  __shared__ float sm_m[kXentThreads / 32], sm_s[kXentThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sm_m[warp] = m, sm_s[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kXentThreads / 32; ++w) lse_combine(m, s, sm_m[w], sm_s[w]);
    const float lse = m + logf(s);
    row_lse[r] = lse;
    const int64_t lab = labels[r];
    float l = 0.f;
    if (lab != ignore_index && lab >= 0 && lab < V) {
      l = lse - x[lab];
      atomicAdd(loss_sum, l);
      atomicAdd(count, 1.0f);
    }
    if (row_loss) row_loss[r] = l;
  }
}

// d[r, c] = g * (exp(logits[r,c] - lse[r]) - [c == label[r]]) for labelled rows, 0 otherwise; g = *gout / *count.
// Writes fp32 (autograd contract of the logits tensor) and/or fp16 (operand of the decoder dgrad / wgrad GEMMs).
__global__ void __launch_bounds__(kXentThreads)
xent_bwd_kernel(const float* logits, int64_t ld, const int64_t* labels, int V, int64_t ignore_index, const float* row_lse,
                const float* gout, const float* count, float* d32, int64_t ldd32, __half* d16, int64_t ldd16, int Vpad16) {
  const int r = blockIdx.x;
  const int64_t lab = labels[r];
  const bool valid = lab != ignore_index && lab >= 0 && lab < V;
  const float g = valid ? (*gout) / (*count) : 0.f;
  const float lse = row_lse[r];
  const float* x = logits + (int64_t)r * ld;
  for (int c = threadIdx.x; c < V; c += kXentThreads) {
    float v = 0.f;
    if (valid) v = g * (expf(x[c] - lse) - (c == lab ? 1.f : 0.f));
    if (d32) d32[(int64_t)r * ldd32 + c] = v;
    if (d16) d16[(int64_t)r * ldd16 + c] = __float2half_rn(v);
  }
  if (d16)
    for (int c = V + threadIdx.x; c < Vpad16; c += kXentThreads) d16[(int64_t)r * ldd16 + c] = __float2half_rn(0.f);
}

}  // namespace lav

using namespace lav;

extern "C" int lav_xent_fwd(const float* logits, int64_t ld, const int64_t* labels, int rows, int V,
                            int64_t ignore_index, float* row_lse, float* row_loss, float* loss_sum, float* count,
                            void* stream) {
  LAV_REQUIRE(logits && labels && row_lse && loss_sum && count, "lav_xent_fwd: null pointer");
  LAV_REQUIRE(V > 0 && ld >= V, "lav_xent_fwd: bad shape");
  if (rows <= 0) return LAV_OK;
  xent_fwd_kernel<<<rows, kXentThreads, 0, (cudaStream_t)stream>>>(logits, ld, labels, V, ignore_index, row_lse, row_loss,
                                                                  loss_sum, count);
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}

extern "C" int lav_xent_bwd(const float* logits, int64_t ld, const int64_t* labels, int rows, int V,
                            int64_t ignore_index, const float* row_lse, const float* gout, const float* count,
                            float* d32, int64_t ldd32, void* d16, int64_t ldd16, void* stream) {
  LAV_REQUIRE(logits && labels && row_lse && gout && count && (d32 || d16), "lav_xent_bwd: null pointer");
  LAV_REQUIRE(V > 0 && ld >= V && (!d32 || ldd32 >= V) && (!d16 || ldd16 >= V), "lav_xent_bwd: bad shape");
  if (rows <= 0) return LAV_OK;
  const int vpad = d16 ? (int)std::min<int64_t>(ldd16, (V + 7) / 8 * 8) : V;
  xent_bwd_kernel<<<rows, kXentThreads, 0, (cudaStream_t)stream>>>(logits, ld, labels, V, ignore_index, row_lse, gout, count,
                                                                  d32, ldd32, (__half*)d16, ldd16, vpad);
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}
