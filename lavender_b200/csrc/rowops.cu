// HBM-bound row kernels of the hot path: LayerNorm forward/backward (with the window-partition / cyclic-shift /
// patch-merging gather folded into the row map), fp32->fp16 casts, column sums (bias gradients).
// One warp per row, 128-bit accesses, fp32 statistics.
#include "rng.cuh"
#include "runtime.h"
#include "sm100.cuh"

namespace lav {

constexpr int kRowThreads = 256;  // 8 warps = 8 rows per CTA

static inline int row_grid(int64_t rows) {
  int64_t ctas = (rows + (kRowThreads / 32) - 1) / (kRowThreads / 32);
  int64_t cap = (int64_t)sm_count() * 16;
  return (int)std::max<int64_t>(1, std::min<int64_t>(ctas, cap));
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm forward.  Output row r (width G*C) is the concatenation of G source rows of width C:
//   src row = map ? map[r*G+g] : r*G+g      (G=1: plain / window gather;  G=4: PatchMerging gather)
// ---------------------------------------------------------------------------------------------------------
struct LnFwdParams {
  const float* x; int64_t ldx;
  const int32_t* map; int G; int C;
  const float* gamma; const float* beta; float eps;
  __half* y16; int64_t ldy16; float* y32; int64_t ldy32;
  float* mean; float* rstd;
  int rows;
};

__global__ void __launch_bounds__(kRowThreads) ln_fwd_kernel(const LnFwdParams p) {
  griddep_launch();  // the next kernel may start its prologue under this kernel's tail
  griddep_wait();    // (this one may have started under its predecessor's: wait before touching global data)
  const int lane = threadIdx.x & 31;
  const int wpb = kRowThreads / 32;
  const int W = p.G * p.C;  // normalised width
  const int c4 = p.C >> 2;  // float4 per source row
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < p.rows; r += gridDim.x * wpb) {
    const float4* src[4];
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < p.G) {
        const int64_t sr = p.map ? p.map[(int64_t)r * p.G + g] : (int64_t)r * p.G + g;
        src[g] = reinterpret_cast<const float4*>(p.x + sr * p.ldx);
      }
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < p.G)
        for (int i = lane; i < c4; i += 32) {
          float4 v = src[g][i];
          s += (v.x + v.y) + (v.z + v.w);
        }
    const float mean = warp_sum(s) / W;
    float ss = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < p.G)
        for (int i = lane; i < c4; i += 32) {
          float4 v = src[g][i];
          float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
          ss += (a * a + b * b) + (c * c + d * d);
        }
    const float rstd = 1.0f / sqrtf(warp_sum(ss) / W + p.eps);
    if (lane == 0) {
      if (p.mean) p.mean[r] = mean;
      if (p.rstd) p.rstd[r] = rstd;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < p.G)
        for (int i = lane; i < c4; i += 32) {
          const int col = g * p.C + i * 4;
          float4 v = src[g][i];
          float4 ga = *reinterpret_cast<const float4*>(p.gamma + col);
          float4 be = *reinterpret_cast<const float4*>(p.beta + col);
          float4 o;
          o.x = (v.x - mean) * rstd * ga.x + be.x;
          o.y = (v.y - mean) * rstd * ga.y + be.y;
          o.z = (v.z - mean) * rstd * ga.z + be.z;
          o.w = (v.w - mean) * rstd * ga.w + be.w;
          if (p.y32) *reinterpret_cast<float4*>(p.y32 + (int64_t)r * p.ldy32 + col) = o;
          if (p.y16) {
            uint2 u = make_uint2(pack_half2(o.x, o.y), pack_half2(o.z, o.w));
            *reinterpret_cast<uint2*>(p.y16 + (int64_t)r * p.ldy16 + col) = u;
          }
        }
  }
}

// Register-resident variant (G == 1, C <= 128 * NV): the row is loaded ONCE, all loads issued before the first use, and
// stays in registers for the statistics and the normalisation (the generic kernel re-reads it through L1 twice).
template <int NV>
__global__ void __launch_bounds__(kRowThreads) ln_fwd_reg_kernel(const LnFwdParams p) {
  griddep_launch();
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const int wpb = kRowThreads / 32;
  const int c4 = p.C >> 2;
  const float invC = 1.0f / (float)p.C;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < p.rows; r += gridDim.x * wpb) {
    const int64_t sr = p.map ? p.map[r] : r;
    const float4* src = reinterpret_cast<const float4*>(p.x + sr * p.ldx);
    float4 v[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = lane + 32 * k;
      v[k] = i < c4 ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    const float mean = warp_sum(s) * invC;
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      if (lane + 32 * k < c4) {
        const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
        ss += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(ss) * invC + p.eps);
    if (lane == 0) {
      if (p.mean) p.mean[r] = mean;
      if (p.rstd) p.rstd[r] = rstd;
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = lane + 32 * k;
      if (i < c4) {
        const float4 ga = __ldg(reinterpret_cast<const float4*>(p.gamma) + i);
        const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta) + i);
        float4 o;
        o.x = (v[k].x - mean) * rstd * ga.x + be.x;
        o.y = (v[k].y - mean) * rstd * ga.y + be.y;
        o.z = (v[k].z - mean) * rstd * ga.z + be.z;
        o.w = (v[k].w - mean) * rstd * ga.w + be.w;
        if (p.y32) *reinterpret_cast<float4*>(p.y32 + (int64_t)r * p.ldy32 + i * 4) = o;
        if (p.y16)
          *reinterpret_cast<uint2*>(p.y16 + (int64_t)r * p.ldy16 + i * 4) = make_uint2(pack_half2(o.x, o.y), pack_half2(o.z, o.w));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm backward (input gradient).  For output row r (same gather as forward):
//   xhat = (x - mean) * rstd ; a = dy*gamma ; dx = rstd * (a - mean_c(a) - xhat * mean_c(a*xhat))
//   dx32[src row] = dx (+ add32[src row])   ; dx16[r] = (fp16) of the same value (identity rows, G == 1)
// dy is fp16 or fp32.  Parameter gradients are produced by ln_bwd_params_kernel.
// ---------------------------------------------------------------------------------------------------------
struct LnBwdParams {
  const void* dy; int64_t lddy; int dy_f32;
  const float* x; int64_t ldx;
  const int32_t* map; int G; int C;
  const float* gamma; const float* mean; const float* rstd;
  const float* add32; int64_t ldadd;
  float* dx32; int64_t lddx32;
  __half* dx16; int64_t lddx16;
  float* dgamma; float* dbeta;
  int rows;
  DropParams drop;  // mask applied to dx16 only
  float* param_ws;  // fused parameter gradients: per-block partial sums [grid][2][C] (null: atomics straight to dgamma/dbeta)
  // where dx16 goes (fused gradient casts, lav_layernorm_bwd_ex): row dx16_map[src row], else the src row, else row r;
  // multiplied by dx16_scale[src row / dx16_rps] (DropPath keep factor of the consumer)
  const int32_t* dx16_map; int dx16_at_src; const float* dx16_scale; int dx16_rps;
};

__device__ __forceinline__ int64_t dx16_row(const LnBwdParams& p, int r, int64_t srow) {
  return p.dx16_map ? (int64_t)p.dx16_map[srow] : (p.dx16_at_src ? srow : (int64_t)r);
}
__device__ __forceinline__ float dx16_factor(const LnBwdParams& p, int64_t srow) {
  return p.dx16_scale ? p.dx16_scale[srow / p.dx16_rps] : 1.0f;
}

__device__ __forceinline__ float4 load_dy4(const LnBwdParams& p, int r, int col) {
  if (p.dy_f32) return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.dy) + (int64_t)r * p.lddy + col);
  uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p.dy) + (int64_t)r * p.lddy + col);
  float2 a = __half22float2(*reinterpret_cast<__half2*>(&u.x));
  float2 b = __half22float2(*reinterpret_cast<__half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// PARAMS: also accumulate dgamma / dbeta (G == 1, C <= 1024): every lane keeps the partial sums of its columns over
// the rows its warp walks, the 8 warps of a block meet in shared memory and the block issues one atomic per column -
// the separate parameter-gradient pass (a second read of x and dy) disappears.
template <bool PARAMS, int NV = 8>  // NV: float4 column groups per lane kept in registers (C <= 128 * NV)
__global__ void __launch_bounds__(kRowThreads) ln_bwd_kernel(const LnBwdParams p) {
  griddep_launch();  // the next kernel may start its prologue under this kernel's tail
  griddep_wait();    // (this one may have started under its predecessor's: wait before touching global data)
  const int lane = threadIdx.x & 31;
  const int wpb = kRowThreads / 32;
  const int W = p.G * p.C;
  const int c4 = p.C >> 2;
  DropKey dkey{};
  if (p.drop.on) dkey = drop_key(p.drop);
  float4 ag[PARAMS ? NV : 1], ab[PARAMS ? NV : 1];
  if (PARAMS) {
#pragma unroll
    for (int k = 0; k < NV; ++k) ag[k] = make_float4(0.f, 0.f, 0.f, 0.f), ab[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < p.rows; r += gridDim.x * wpb) {
    int64_t srow[4];
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < p.G) srow[g] = p.map ? p.map[(int64_t)r * p.G + g] : (int64_t)r * p.G + g;
    const float mean = p.mean[r], rstd = p.rstd[r];
    const int64_t drow16 = p.dx16 ? dx16_row(p, r, srow[0]) : 0;
    const float sc16 = p.dx16 ? dx16_factor(p, srow[0]) : 1.0f;
    float s1 = 0.f, s2 = 0.f;
    if (PARAMS) {  // G == 1: the row lives in registers; every global load is issued before the first use
      float4 v[NV], d[NV], ad[NV];
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int i = lane + 32 * k;
        const bool in = i < c4;
        v[k] = in ? *reinterpret_cast<const float4*>(p.x + srow[0] * p.ldx + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        d[k] = in ? load_dy4(p, r, i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        ad[k] = (in && p.add32) ? *reinterpret_cast<const float4*>(p.add32 + srow[0] * p.ldadd + i * 4)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int i = lane + 32 * k;
        if (i < c4) {
          const float4 ga = __ldg(reinterpret_cast<const float4*>(p.gamma) + i);
          const float x0 = (v[k].x - mean) * rstd, x1 = (v[k].y - mean) * rstd, x2 = (v[k].z - mean) * rstd,
                      x3 = (v[k].w - mean) * rstd;
          ag[k].x += d[k].x * x0, ag[k].y += d[k].y * x1, ag[k].z += d[k].z * x2, ag[k].w += d[k].w * x3;
          ab[k].x += d[k].x, ab[k].y += d[k].y, ab[k].z += d[k].z, ab[k].w += d[k].w;
          d[k].x *= ga.x, d[k].y *= ga.y, d[k].z *= ga.z, d[k].w *= ga.w;  // a = dy * gamma
          v[k] = make_float4(x0, x1, x2, x3);                              // xhat
          s1 += (d[k].x + d[k].y) + (d[k].z + d[k].w);
          s2 += (d[k].x * x0 + d[k].y * x1) + (d[k].z * x2 + d[k].w * x3);
        }
      }
      s1 = warp_sum(s1) / W;
      s2 = warp_sum(s2) / W;  // mean_c(a * xhat)
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int i = lane + 32 * k;
        if (i < c4) {
          const int col = i * 4;
          float4 o;
          o.x = rstd * (d[k].x - s1 - v[k].x * s2) + ad[k].x;
          o.y = rstd * (d[k].y - s1 - v[k].y * s2) + ad[k].y;
          o.z = rstd * (d[k].z - s1 - v[k].z * s2) + ad[k].z;
          o.w = rstd * (d[k].w - s1 - v[k].w * s2) + ad[k].w;
          if (p.dx32) *reinterpret_cast<float4*>(p.dx32 + srow[0] * p.lddx32 + col) = o;
          if (p.dx16) {
            if (p.drop.on) {  // gradient entering a dense layer whose output was dropped at (r, col) in the forward
              const uint32_t m = drop_keep8(dkey, p.drop.thresh, (uint32_t)r, (uint32_t)(col >> 3), 0u) >> (col & 7);
              o.x = (m & 1u) ? o.x * p.drop.inv_keep : 0.f, o.y = (m & 2u) ? o.y * p.drop.inv_keep : 0.f;
              o.z = (m & 4u) ? o.z * p.drop.inv_keep : 0.f, o.w = (m & 8u) ? o.w * p.drop.inv_keep : 0.f;
            }
            *reinterpret_cast<uint2*>(p.dx16 + drow16 * p.lddx16 + col) =
                make_uint2(pack_half2(o.x * sc16, o.y * sc16), pack_half2(o.z * sc16, o.w * sc16));
          }
        }
      }
      continue;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < p.G)
        for (int i = lane; i < c4; i += 32) {
          const int col = g * p.C + i * 4;
          float4 v = *reinterpret_cast<const float4*>(p.x + srow[g] * p.ldx + i * 4);
          float4 d = load_dy4(p, r, col);
          float4 ga = *reinterpret_cast<const float4*>(p.gamma + col);
          float a0 = d.x * ga.x, a1 = d.y * ga.y, a2 = d.z * ga.z, a3 = d.w * ga.w;
          s1 += (a0 + a1) + (a2 + a3);
          s2 += (a0 * (v.x - mean) + a1 * (v.y - mean)) + (a2 * (v.z - mean) + a3 * (v.w - mean));
        }
    s1 = warp_sum(s1) / W;
    s2 = warp_sum(s2) * rstd / W;  // mean_c(a * xhat)
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < p.G)
        for (int i = lane; i < c4; i += 32) {
          const int col = g * p.C + i * 4;
          float4 v = *reinterpret_cast<const float4*>(p.x + srow[g] * p.ldx + i * 4);
          float4 d = load_dy4(p, r, col);
          float4 ga = *reinterpret_cast<const float4*>(p.gamma + col);
          float4 o;
          o.x = rstd * (d.x * ga.x - s1 - (v.x - mean) * rstd * s2);
          o.y = rstd * (d.y * ga.y - s1 - (v.y - mean) * rstd * s2);
          o.z = rstd * (d.z * ga.z - s1 - (v.z - mean) * rstd * s2);
          o.w = rstd * (d.w * ga.w - s1 - (v.w - mean) * rstd * s2);
          if (p.add32) {
            float4 a = *reinterpret_cast<const float4*>(p.add32 + srow[g] * p.ldadd + i * 4);
            o.x += a.x, o.y += a.y, o.z += a.z, o.w += a.w;
          }
          if (p.dx32) *reinterpret_cast<float4*>(p.dx32 + srow[g] * p.lddx32 + i * 4) = o;
          if (p.dx16) {
            if (p.drop.on) {  // gradient entering a dense layer whose output was dropped at (r, col) in the forward
              const uint32_t m = drop_keep8(dkey, p.drop.thresh, (uint32_t)r, (uint32_t)(col >> 3), 0u) >> (col & 7);
              o.x = (m & 1u) ? o.x * p.drop.inv_keep : 0.f, o.y = (m & 2u) ? o.y * p.drop.inv_keep : 0.f;
              o.z = (m & 4u) ? o.z * p.drop.inv_keep : 0.f, o.w = (m & 8u) ? o.w * p.drop.inv_keep : 0.f;
            }
            *reinterpret_cast<uint2*>(p.dx16 + drow16 * p.lddx16 + col) =
                make_uint2(pack_half2(o.x * sc16, o.y * sc16), pack_half2(o.z * sc16, o.w * sc16));
          }
        }
  }
  if (PARAMS) {
    // block reduction of the per-lane column sums: [8 warps][C] floats, dgamma first, then dbeta through the same buffer
    extern __shared__ float red[];
    const int wy = threadIdx.x >> 5;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      __syncthreads();
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int i = lane + 32 * k;
        if (i < c4) *reinterpret_cast<float4*>(red + wy * p.C + i * 4) = pass == 0 ? ag[k] : ab[k];
      }
      __syncthreads();
      // hundreds of blocks adding to the same C addresses serialise in L2 (~17 us at C = 512): with a workspace the
      // block stores its partial row instead and ln_param_reduce_kernel sums the rows
      float* dst = p.param_ws ? p.param_ws + ((size_t)blockIdx.x * 2 + pass) * p.C : (pass == 0 ? p.dgamma : p.dbeta);
      for (int c = threadIdx.x; c < p.C; c += kRowThreads) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w * p.C + c];
        if (p.param_ws) dst[c] = t;
        else atomicAdd(dst + c, t);
      }
    }
  }
}

// Wide rows (C = 512 ... 1024: Swin stages 2-3, BERT): the register-resident kernel above runs ONE block of 8 warps per
// SM there (the row + the per-lane dgamma / dbeta sums take 130-200 registers), and a warp alternates between "loads in
// flight" and "math + stores", so only ~30 KB per SM are in flight on average: 1.7 TB/s.  Here the in-flight bytes live
// in shared memory instead: every warp owns a ring of `depth` row slots (x row | dy row | add row) filled by 1-D bulk
// copies (cp.async.bulk -> mbarrier), issued `depth` rows ahead by lane 0, so 120-190 KB per SM stay in flight while the
// warps do math on rows that have already landed.  Same arithmetic, same outputs, same parameter-gradient path.
template <int NV>
__global__ void __launch_bounds__(kRowThreads, 1) ln_bwd_staged_kernel(const LnBwdParams p, int depth) {
  extern __shared__ __align__(128) uint8_t sm_raw[];
  griddep_launch();
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  constexpr int wpb = kRowThreads / 32;
  const int c4 = p.C >> 2;
  const uint32_t xb = p.C * 4u, dyb = p.dy_f32 ? p.C * 4u : p.C * 2u, adb = p.add32 ? p.C * 4u : 0u;
  const uint32_t stage_bytes = xb + dyb + adb;
  float* red = reinterpret_cast<float*>(sm_raw);                                   // [8 warps][C]
  uint64_t* mybar = reinterpret_cast<uint64_t*>(sm_raw + (size_t)wpb * p.C * 4) + wy * 4;  // 4 barriers per warp
  uint8_t* my = sm_raw + (size_t)wpb * p.C * 4 + 256 + (size_t)wy * depth * stage_bytes;
  if (lane == 0) {
    for (int s = 0; s < depth; ++s) mbar_init(mybar + s, 1);
    fence_barrier_init();
  }
  __syncwarp();
  griddep_wait();
  DropKey dkey{};
  if (p.drop.on) dkey = drop_key(p.drop);
  const int r0 = blockIdx.x * wpb + wy, stride = gridDim.x * wpb;
  auto issue = [&](int r, int s) {  // lane 0: the three rows of output row r into slot s
    const int64_t srow = p.map ? p.map[r] : (int64_t)r;
    uint8_t* dst = my + (size_t)s * stage_bytes;
    mbar_arrive_expect_tx(mybar + s, stage_bytes);
    bulk_load_1d(dst, p.x + srow * p.ldx, xb, mybar + s);
    bulk_load_1d(dst + xb, reinterpret_cast<const uint8_t*>(p.dy) + (int64_t)r * p.lddy * (p.dy_f32 ? 4 : 2), dyb, mybar + s);
    if (adb) bulk_load_1d(dst + xb + dyb, p.add32 + srow * p.ldadd, adb, mybar + s);
  };
  if (lane == 0)
    for (int s = 0; s < depth; ++s)
      if (r0 + s * stride < p.rows) issue(r0 + s * stride, s);
  float4 ag[NV], ab[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) ag[k] = make_float4(0.f, 0.f, 0.f, 0.f), ab[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  float mean_n = 0.f, rstd_n = 0.f;
  int64_t srow_n = 0;
  if (r0 < p.rows) mean_n = p.mean[r0], rstd_n = p.rstd[r0], srow_n = p.map ? p.map[r0] : (int64_t)r0;
  int s = 0;
  uint32_t par = 0;
  for (int r = r0; r < p.rows; r += stride) {
    const float mean = mean_n, rstd = rstd_n;
    const int64_t srow = srow_n;
    if (r + stride < p.rows) {  // the next row's scalars travel under this row's math
      const int rn = r + stride;
      mean_n = p.mean[rn], rstd_n = p.rstd[rn], srow_n = p.map ? p.map[rn] : (int64_t)rn;
    }
    const int64_t drow16 = p.dx16 ? dx16_row(p, r, srow) : 0;
    const float sc16 = p.dx16 ? dx16_factor(p, srow) : 1.0f;
    mbar_wait(mybar + s, par, 40);
    const uint8_t* st = my + (size_t)s * stage_bytes;
    float4 v[NV], d[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = lane + 32 * k;
      if (i < c4) {
        v[k] = *reinterpret_cast<const float4*>(st + i * 16);
        if (p.dy_f32) d[k] = *reinterpret_cast<const float4*>(st + xb + i * 16);
        else {
          const uint2 u = *reinterpret_cast<const uint2*>(st + xb + i * 8);
          const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
          const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
          d[k] = make_float4(a.x, a.y, b.x, b.y);
        }
        const float4 ga = __ldg(reinterpret_cast<const float4*>(p.gamma) + i);
        const float x0 = (v[k].x - mean) * rstd, x1 = (v[k].y - mean) * rstd, x2 = (v[k].z - mean) * rstd,
                    x3 = (v[k].w - mean) * rstd;
        ag[k].x += d[k].x * x0, ag[k].y += d[k].y * x1, ag[k].z += d[k].z * x2, ag[k].w += d[k].w * x3;
        ab[k].x += d[k].x, ab[k].y += d[k].y, ab[k].z += d[k].z, ab[k].w += d[k].w;
        d[k].x *= ga.x, d[k].y *= ga.y, d[k].z *= ga.z, d[k].w *= ga.w;  // a = dy * gamma
        v[k] = make_float4(x0, x1, x2, x3);                              // xhat
        s1 += (d[k].x + d[k].y) + (d[k].z + d[k].w);
        s2 += (d[k].x * x0 + d[k].y * x1) + (d[k].z * x2 + d[k].w * x3);
      }
    }
    s1 = warp_sum(s1) / p.C;
    s2 = warp_sum(s2) / p.C;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = lane + 32 * k;
      if (i < c4) {
        const int col = i * 4;
        float4 o;
        o.x = rstd * (d[k].x - s1 - v[k].x * s2), o.y = rstd * (d[k].y - s1 - v[k].y * s2);
        o.z = rstd * (d[k].z - s1 - v[k].z * s2), o.w = rstd * (d[k].w - s1 - v[k].w * s2);
        if (adb) {
          const float4 a = *reinterpret_cast<const float4*>(st + xb + dyb + i * 16);
          o.x += a.x, o.y += a.y, o.z += a.z, o.w += a.w;
        }
        if (p.dx32) *reinterpret_cast<float4*>(p.dx32 + srow * p.lddx32 + col) = o;
        if (p.dx16) {
          if (p.drop.on) {
            const uint32_t m = drop_keep8(dkey, p.drop.thresh, (uint32_t)r, (uint32_t)(col >> 3), 0u) >> (col & 7);
            o.x = (m & 1u) ? o.x * p.drop.inv_keep : 0.f, o.y = (m & 2u) ? o.y * p.drop.inv_keep : 0.f;
            o.z = (m & 4u) ? o.z * p.drop.inv_keep : 0.f, o.w = (m & 8u) ? o.w * p.drop.inv_keep : 0.f;
          }
          *reinterpret_cast<uint2*>(p.dx16 + drow16 * p.lddx16 + col) =
              make_uint2(pack_half2(o.x * sc16, o.y * sc16), pack_half2(o.z * sc16, o.w * sc16));
        }
      }
    }
    // every lane has consumed its part of the slot (the values above were used): refill it `depth` rows ahead
    __syncwarp();
    if (lane == 0 && r + depth * stride < p.rows) {
      fence_proxy_async_smem();
      issue(r + depth * stride, s);
    }
    if (++s == depth) s = 0, par ^= 1;
  }
  // block reduction of the per-lane column sums, as in ln_bwd_kernel<true>
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = lane + 32 * k;
      if (i < c4) *reinterpret_cast<float4*>(red + wy * p.C + i * 4) = pass == 0 ? ag[k] : ab[k];
    }
    __syncthreads();
    float* dst = p.param_ws ? p.param_ws + ((size_t)blockIdx.x * 2 + pass) * p.C : (pass == 0 ? p.dgamma : p.dbeta);
    for (int c = threadIdx.x; c < p.C; c += kRowThreads) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w * p.C + c];
      if (p.param_ws) dst[c] = t;
      else atomicAdd(dst + c, t);
    }
  }
}

// dgamma[c] += sum_b ws[b][0][c] ; dbeta[c] += sum_b ws[b][1][c]   (second stage of the fused parameter gradients)
__global__ void __launch_bounds__(1024) ln_param_reduce_kernel(const float* ws, int nblocks, int C, float* dgamma, float* dbeta) {
  griddep_launch();
  griddep_wait();
  // block = 32 columns x 32 row slabs: coalesced 128-byte reads, 32-way parallel over the partial rows
  __shared__ float sm[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;  // column of the [2C]-wide partial rows
  float t = 0.f;
  if (j < 2 * C)
    for (int b = ty; b < nblocks; b += 32) t += ws[(size_t)b * 2 * C + j];
  sm[ty][tx] = t;
  __syncthreads();
  if (ty == 0 && j < 2 * C) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) a += sm[k][tx];
    if (j < C) dgamma[j] += a;
    else dbeta[j - C] += a;
  }
}

// dgamma[c] += sum_r dy[r,c]*xhat[r,c] ; dbeta[c] += sum_r dy[r,c].
// Block = 8 warps; a warp covers 128 consecutive columns (one float4 / 4 halfs per lane) of one row at a time and
// walks the rows of its slab with stride 8; partial sums meet in shared memory, one atomic per column per block.
__global__ void __launch_bounds__(256) ln_bwd_params_kernel(const LnBwdParams p, int rows_per_block) {
  griddep_launch();  // the next kernel may start its prologue under this kernel's tail
  griddep_wait();    // (this one may have started under its predecessor's: wait before touching global data)
  __shared__ float sg[8][128], sb[8][128];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int W = p.G * p.C;
  const int col = blockIdx.x * 128 + lane * 4;  // C % 4 == 0, so a quad never straddles two gathered source rows
  const int g = col / p.C, cc = col - g * p.C;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(p.rows, r0 + rows_per_block);
  float ag[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
  if (col < W) {
    for (int r = r0 + wy; r < r1; r += 8) {
      const int64_t sr = p.map ? p.map[(int64_t)r * p.G + g] : (int64_t)r * p.G + g;
      const float4 xv = *reinterpret_cast<const float4*>(p.x + sr * p.ldx + cc);
      const float4 d = load_dy4(p, r, col);
      const float mean = p.mean[r], rstd = p.rstd[r];
      ag[0] += d.x * (xv.x - mean) * rstd, ag[1] += d.y * (xv.y - mean) * rstd;
      ag[2] += d.z * (xv.z - mean) * rstd, ag[3] += d.w * (xv.w - mean) * rstd;
      ab[0] += d.x, ab[1] += d.y, ab[2] += d.z, ab[3] += d.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) sg[wy][lane * 4 + j] = ag[j], sb[wy][lane * 4 + j] = ab[j];
  __syncthreads();
  if (threadIdx.x < 128) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c < W) {
      float tg = 0.f, tb = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) tg += sg[k][threadIdx.x], tb += sb[k][threadIdx.x];
      atomicAdd(p.dgamma + c, tg);
      atomicAdd(p.dbeta + c, tb);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// out16[r, 0:C] = (fp16)( x[map ? map[r] : r, 0:C] * (scale ? scale[r / rows_per_scale] : 1) * alpha )
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowThreads)
scale_cast_kernel(const float* x, int64_t ldx, const int32_t* map, const float* scale, int rows_per_scale, float alpha,
                  __half* out, int64_t ldo, int rows, int C) {
  griddep_launch();  // the next kernel may start its prologue under this kernel's tail
  griddep_wait();    // (this one may have started under its predecessor's: wait before touching global data)
  const int lane = threadIdx.x & 31, wpb = kRowThreads / 32;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const int64_t sr = map ? map[r] : r;
    const float s = alpha * (scale ? scale[r / rows_per_scale] : 1.f);
    if ((ldx & 3) == 0) {
      const float4* src = reinterpret_cast<const float4*>(x + sr * ldx);
      for (int i = lane; i < (C >> 2); i += 32) {
        float4 v = src[i];
        *reinterpret_cast<uint2*>(out + (int64_t)r * ldo + i * 4) =
            make_uint2(pack_half2(v.x * s, v.y * s), pack_half2(v.z * s, v.w * s));
      }
      for (int c = (C & ~3) + lane; c < C; c += 32) out[(int64_t)r * ldo + c] = __float2half_rn(x[sr * ldx + c] * s);
    } else {  // unaligned rows (e.g. fp32 logit gradients with ld = 30522)
      for (int c = lane; c < C; c += 32) out[(int64_t)r * ldo + c] = __float2half_rn(x[sr * ldx + c] * s);
    }
  }
}

// Same, for FEW WIDE rows (logit gradients: 32-160 rows x 30522): the warp-per-row kernel above would run 4-20 blocks,
// each lane walking ~1000 dependent iterations (measured 120-135 us for 15 MB).  Block = (row, 2048-column slab), thread =
// 8 consecutive columns; fp32 rows of odd length are only 4-byte aligned, so loads are scalar but all 8 are in flight at
// once and the fp16 store is one 16-byte vector when the output row allows it.
__global__ void __launch_bounds__(256)
scale_cast_wide_kernel(const float* x, int64_t ldx, const int32_t* map, const float* scale, int rows_per_scale, float alpha,
                       __half* out, int64_t ldo, int C) {
  griddep_launch();
  griddep_wait();
  const int r = blockIdx.y;
  const int c0 = (blockIdx.x * 256 + threadIdx.x) * 8;
  if (c0 >= C) return;
  const int64_t sr = map ? map[r] : r;
  const float s = alpha * (scale ? scale[r / rows_per_scale] : 1.f);
  const float* src = x + sr * ldx + c0;
  __half* dst = out + (int64_t)r * ldo + c0;
  float v[8];
  if (c0 + 8 <= C) {
    if (((reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
      const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
      v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    } else if (((reinterpret_cast<uintptr_t>(src)) & 7) == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = *reinterpret_cast<const float2*>(src + 2 * j);
        v[2 * j] = a.x, v[2 * j + 1] = a.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = src[j];
    }
    if (((reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
      *reinterpret_cast<uint4*>(dst) = make_uint4(pack_half2(v[0] * s, v[1] * s), pack_half2(v[2] * s, v[3] * s),
                                                  pack_half2(v[4] * s, v[5] * s), pack_half2(v[6] * s, v[7] * s));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = __float2half_rn(v[j] * s);
    }
  } else {
    for (int j = 0; c0 + j < C; ++j) dst[j] = __float2half_rn(src[j] * s);
  }
}

// Split-fp16 operand of the high-precision GEMM mode (LAV_PRECISION=high): x = hi + lo, hi = fp16(x), lo = fp16(x - hi)
// (|x - hi - lo| <= 2^-22 |x|).  out16[r] = [hi | lo | hi] (mode 0, A operand) or [hi | hi | lo] (mode 1, B operand),
// each C wide, so that ONE fp16 GEMM over K' = 3C accumulates Ah*Bh + Al*Bh + Ah*Bl in fp32.
__global__ void __launch_bounds__(kRowThreads)
split3_kernel(const float* x, int64_t ldx, __half* out, int64_t ldo, int rows, int C, int mode) {
  griddep_launch();
  griddep_wait();
  const int lane = threadIdx.x & 31, wpb = kRowThreads / 32;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const float* src = x + (int64_t)r * ldx;
    __half* o = out + (int64_t)r * ldo;
    for (int c = lane; c < C; c += 32) {
      const float v = src[c];
      const __half hi = __float2half_rn(v);
      const __half lo = __float2half_rn(v - __half2float(hi));
      o[c] = hi;
      o[C + c] = mode == 0 ? lo : hi;
      o[2 * C + c] = mode == 0 ? hi : lo;
    }
  }
}

// flat fp32 -> fp16 (parameter shadow copy)
__global__ void __launch_bounds__(256) cast_flat_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t n) {
  griddep_launch();  // the next kernel may start its prologue under this kernel's tail
  griddep_wait();    // (this one may have started under its predecessor's: wait before touching global data)
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_half2(v.x, v.y), pack_half2(v.z, v.w));
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = __float2half_rn(src[i]);
}

// out[c] += alpha * sum_r x16[r, c]   (bias gradients).  Block = 8 warps; a warp covers 256 consecutive columns of
// one row (one 16-byte load per lane) and walks its slab of rows with stride 8.  Needs ld % 8 == 0.
__global__ void __launch_bounds__(256)
colsum_f16_kernel(const __half* x, int64_t ld, int rows, int N, float* out, float alpha, int rows_per_block) {
  griddep_launch();  // the next kernel may start its prologue under this kernel's tail
  griddep_wait();    // (this one may have started under its predecessor's: wait before touching global data)
  __shared__ float sm[8][256];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col + 8 <= ld && col < N) {
    for (int r = r0 + wy; r < r1; r += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + (int64_t)r * ld + col);
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __half22float2(h[q]);
        a[2 * q] += f.x, a[2 * q + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[wy][lane * 8 + j] = a[j];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x];
    atomicAdd(out + c, t * alpha);
  }
}

// out16 = dy16 * dgelu16   (flat, n % 2 == 0; dgelu16 = gelu'(pre-activation) as saved by the forward GEMM epilogue)
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const __half2* dy, const __half2* pre, __half2* out, int64_t n2) {
  griddep_launch();  // the next kernel may start its prologue under this kernel's tail
  griddep_wait();    // (this one may have started under its predecessor's: wait before touching global data)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    float2 d = __half22float2(dy[i]), x = __half22float2(pre[i]);
    out[i] = __floats2half2_rn(d.x * x.x, d.y * x.y);
  }
}

// out = x * keep / (1 - p), fp32 [rows, C], one thread per 8 consecutive columns (one Philox call)
__global__ void __launch_bounds__(256)
dropout_f32_kernel(const float* x, int64_t ldx, float* out, int64_t ldo, int rows, int C8, const DropParams d) {
  griddep_launch();  // the next kernel may start its prologue under this kernel's tail
  griddep_wait();    // (this one may have started under its predecessor's: wait before touching global data)
  const DropKey key = drop_key(d);
  const int64_t n = (int64_t)rows * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / C8), b = (int)(i - (int64_t)r * C8);
    const uint32_t m = drop_keep8(key, d.thresh, (uint32_t)r, (uint32_t)b, 0u);
    const float4* src = reinterpret_cast<const float4*>(x + (int64_t)r * ldx + b * 8);
    float4* dst = reinterpret_cast<float4*>(out + (int64_t)r * ldo + b * 8);
    float4 v0 = src[0], v1 = src[1];
    v0.x = (m & 1u) ? v0.x * d.inv_keep : 0.f, v0.y = (m & 2u) ? v0.y * d.inv_keep : 0.f;
    v0.z = (m & 4u) ? v0.z * d.inv_keep : 0.f, v0.w = (m & 8u) ? v0.w * d.inv_keep : 0.f;
    v1.x = (m & 16u) ? v1.x * d.inv_keep : 0.f, v1.y = (m & 32u) ? v1.y * d.inv_keep : 0.f;
    v1.z = (m & 64u) ? v1.z * d.inv_keep : 0.f, v1.w = (m & 128u) ? v1.w * d.inv_keep : 0.f;
    dst[0] = v0, dst[1] = v1;
  }
}

__global__ void __launch_bounds__(256)
dropout_mask_kernel(uint8_t* keep, int rows, int C, int head, const DropParams d) {
  griddep_launch();  // the next kernel may start its prologue under this kernel's tail
  griddep_wait();    // (this one may have started under its predecessor's: wait before touching global data)
  const DropKey key = drop_key(d);
  const int C8 = (C + 7) / 8;
  const int64_t n = (int64_t)rows * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / C8), b = (int)(i - (int64_t)r * C8);
    const uint32_t m = drop_keep8(key, d.thresh, (uint32_t)r, (uint32_t)b, head < 0 ? 0u : (uint32_t)head);
    for (int q = 0; q < 8; ++q)
      if (b * 8 + q < C) keep[(int64_t)r * C + b * 8 + q] = (m >> q) & 1u;
  }
}

static int slab_rows(int rows, int col_blocks) {
  // enough row slabs to fill the machine a few times, at least 32 rows each
  int want = std::max(1, (6 * sm_count()) / std::max(1, col_blocks));
  int rpb = std::max(32, (rows + want - 1) / want);
  return (rpb + 7) / 8 * 8;
}

}  // namespace lav

using namespace lav;

extern "C" int lav_layernorm_fwd(const float* x, int64_t ldx, const int32_t* row_map, int G, int C, const float* gamma,
                                 const float* beta, float eps, void* y16, int64_t ldy16, float* y32, int64_t ldy32,
                                 float* mean, float* rstd, int rows, void* stream) {
  LAV_REQUIRE(x && gamma && beta && (y16 || y32), "lav_layernorm_fwd: null pointer");
  LAV_REQUIRE(G >= 1 && G <= 4 && C > 0 && (C % 4) == 0 && (ldx % 4) == 0, "lav_layernorm_fwd: need C%%4==0, 1<=G<=4");
  LAV_REQUIRE((!y16 || ldy16 % 4 == 0) && (!y32 || ldy32 % 4 == 0), "lav_layernorm_fwd: output ld must be %%4");
  if (rows <= 0) return LAV_OK;
  LnFwdParams p{x, ldx, row_map, G, C, gamma, beta, eps, (__half*)y16, ldy16, y32, ldy32, mean, rstd, rows};
  if (G == 1 && C <= 1024) {
    const int grid = row_grid(rows);
    cudaStream_t st = (cudaStream_t)stream;
    if (C <= 128) LAV_CHECK_CUDA(launch_pdl(ln_fwd_reg_kernel<1>, dim3(grid), dim3(kRowThreads), 0, st, p));
    else if (C <= 256) LAV_CHECK_CUDA(launch_pdl(ln_fwd_reg_kernel<2>, dim3(grid), dim3(kRowThreads), 0, st, p));
    else if (C <= 512) LAV_CHECK_CUDA(launch_pdl(ln_fwd_reg_kernel<4>, dim3(grid), dim3(kRowThreads), 0, st, p));
    else LAV_CHECK_CUDA(launch_pdl(ln_fwd_reg_kernel<8>, dim3(grid), dim3(kRowThreads), 0, st, p));
  } else {
    LAV_CHECK_CUDA(launch_pdl(ln_fwd_kernel, dim3(row_grid(rows)), dim3(kRowThreads), 0, (cudaStream_t)stream, p));
  }
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}

static bool ln_staged_enabled() {  // LAV_LN_STAGED=0: the register-resident kernel for every width (A/B testing)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LAV_LN_STAGED");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

extern "C" int lav_layernorm_bwd(const void* dy, int64_t lddy, int dy_is_f32, const float* x, int64_t ldx,
                                 const int32_t* row_map, int G, int C, const float* gamma, const float* mean,
                                 const float* rstd, const float* add32, int64_t ldadd, float* dx32, int64_t lddx32,
                                 void* dx16, int64_t lddx16, float* dgamma, float* dbeta, float* param_ws, int64_t ws_floats,
                                 int rows, const LavDropout* drop16, void* stream) {
  return lav_layernorm_bwd_ex(dy, lddy, dy_is_f32, x, ldx, row_map, G, C, gamma, mean, rstd, add32, ldadd, dx32, lddx32, dx16,
                              lddx16, nullptr, 0, nullptr, 1, dgamma, dbeta, param_ws, ws_floats, rows, drop16, stream);
}

extern "C" int lav_layernorm_bwd_ex(const void* dy, int64_t lddy, int dy_is_f32, const float* x, int64_t ldx,
                                    const int32_t* row_map, int G, int C, const float* gamma, const float* mean,
                                    const float* rstd, const float* add32, int64_t ldadd, float* dx32, int64_t lddx32,
                                    void* dx16, int64_t lddx16, const int32_t* dx16_row_map, int dx16_at_src,
                                    const float* dx16_row_scale, int dx16_rows_per_scale, float* dgamma, float* dbeta,
                                    float* param_ws, int64_t ws_floats, int rows, const LavDropout* drop16, void* stream) {
  LAV_REQUIRE(!(dx16_row_map || dx16_at_src || dx16_row_scale) || (dx16 && G == 1 && dx16_rows_per_scale >= 1),
              "lav_layernorm_bwd_ex: the dx16 placement / scale options need dx16, G == 1 and rows_per_scale >= 1");
  LAV_REQUIRE(dy && x && gamma && mean && rstd && (dx32 || dx16), "lav_layernorm_bwd: null pointer");
  LAV_REQUIRE(G >= 1 && G <= 4 && C > 0 && (C % 4) == 0 && (ldx % 4) == 0 && (lddy % 4) == 0,
              "lav_layernorm_bwd: need C%%4==0, 1<=G<=4, ld%%4==0");
  LAV_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "lav_layernorm_bwd: dgamma/dbeta must come together");
  if (rows <= 0) return LAV_OK;
  LnBwdParams p{dy, lddy, dy_is_f32, x, ldx, row_map, G, C, gamma, mean, rstd, add32, ldadd,
                dx32, lddx32, (__half*)dx16, lddx16, dgamma, dbeta, rows, make_drop(drop16), nullptr,
                dx16_row_map, dx16_at_src, dx16_row_scale, dx16_rows_per_scale > 0 ? dx16_rows_per_scale : 1};
  LAV_REQUIRE(!p.drop.on || (dx16 && G == 1), "lav_layernorm_bwd: drop16 needs dx16 and G == 1");
  cudaStream_t s = (cudaStream_t)stream;
  if (dgamma && G == 1 && C >= 512 && C <= 1024 && (C % 8) == 0 && ln_staged_enabled() &&
      ((uintptr_t)x % 16) == 0 && ((uintptr_t)dy % 16) == 0 && (dy_is_f32 || (lddy % 8) == 0) &&
      (!add32 || (((uintptr_t)add32 % 16) == 0 && (ldadd % 4) == 0))) {
    // wide rows: shared-memory staged kernel, one persistent block per SM (see ln_bwd_staged_kernel)
    const size_t stage = (size_t)C * 4 + (dy_is_f32 ? (size_t)C * 4 : (size_t)C * 2) + (add32 ? (size_t)C * 4 : 0);
    const size_t fixed = (size_t)8 * C * 4 + 256;
    const int depth = (int)std::min<size_t>(4, (227 * 1024 - fixed) / (8 * stage));
    if (depth >= 2) {
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((rows + 7) / 8, (int64_t)sm_count()));
      if (param_ws && ws_floats >= (int64_t)grid * 2 * C) p.param_ws = param_ws;
      const size_t smem = fixed + (size_t)8 * depth * stage;
      static bool attr_set = false;
      if (!attr_set) {
        LAV_CHECK_CUDA(cudaFuncSetAttribute(ln_bwd_staged_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        LAV_CHECK_CUDA(cudaFuncSetAttribute(ln_bwd_staged_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        LAV_CHECK_CUDA(cudaFuncSetAttribute(ln_bwd_staged_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
      }
      if (C <= 512) LAV_CHECK_CUDA(launch_pdl(ln_bwd_staged_kernel<4>, dim3(grid), dim3(kRowThreads), smem, s, p, depth));
      else if (C <= 768) LAV_CHECK_CUDA(launch_pdl(ln_bwd_staged_kernel<6>, dim3(grid), dim3(kRowThreads), smem, s, p, depth));
      else LAV_CHECK_CUDA(launch_pdl(ln_bwd_staged_kernel<8>, dim3(grid), dim3(kRowThreads), smem, s, p, depth));
      LAV_CHECK_CUDA(cudaGetLastError());
      count_launch();
      if (p.param_ws) {
        LAV_CHECK_CUDA(launch_pdl(ln_param_reduce_kernel, dim3((2 * C + 31) / 32), dim3(1024), 0, s, p.param_ws, grid, C, dgamma, dbeta));
        LAV_CHECK_CUDA(cudaGetLastError());
        count_launch();
      }
      return LAV_OK;
    }
  }
  if (dgamma && G == 1 && C <= 1024) {
    // fused: the per-block column reduction (2 * C atomics per block) must amortise over the rows a block walks, and
    // narrow rows need many resident warps to cover the load latency: blocks per SM grow as the rows get narrower
    const size_t red_bytes = (size_t)8 * C * sizeof(float);
    const int nv = C <= 128 ? 1 : C <= 256 ? 2 : C <= 512 ? 4 : 8;
    const int per_sm = nv == 1 ? 8 : nv == 2 ? 5 : nv == 4 ? 2 : 1;  // = resident blocks per SM at each variant's register count
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((rows + 7) / 8, (int64_t)sm_count() * per_sm));
    if (param_ws && ws_floats >= (int64_t)grid * 2 * C) p.param_ws = param_ws;
    static bool attr_set = false;
    if (!attr_set) {
      LAV_CHECK_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 1024 * 4));
      attr_set = true;
    }
    if (nv == 1) LAV_CHECK_CUDA(launch_pdl(ln_bwd_kernel<true, 1>, dim3(grid), dim3(kRowThreads), red_bytes, s, p));
    else if (nv == 2) LAV_CHECK_CUDA(launch_pdl(ln_bwd_kernel<true, 2>, dim3(grid), dim3(kRowThreads), red_bytes, s, p));
    else if (nv == 4) LAV_CHECK_CUDA(launch_pdl(ln_bwd_kernel<true, 4>, dim3(grid), dim3(kRowThreads), red_bytes, s, p));
    else LAV_CHECK_CUDA(launch_pdl(ln_bwd_kernel<true, 8>, dim3(grid), dim3(kRowThreads), red_bytes, s, p));
    LAV_CHECK_CUDA(cudaGetLastError());
    count_launch();
    if (p.param_ws) {
      LAV_CHECK_CUDA(launch_pdl(ln_param_reduce_kernel, dim3((2 * C + 31) / 32), dim3(1024), 0, s, p.param_ws, grid, C, dgamma, dbeta));
      LAV_CHECK_CUDA(cudaGetLastError());
      count_launch();
    }
    return LAV_OK;
  }
  if (dgamma) {  // must read x before an in-place dx32 overwrite of the same rows
    const int cb = (G * C + 127) / 128;
    const int rpb = slab_rows(rows, cb);
    dim3 grid(cb, (rows + rpb - 1) / rpb);
    LAV_CHECK_CUDA(launch_pdl(ln_bwd_params_kernel, dim3(grid), dim3(256), 0, s, p, rpb));
    LAV_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  LAV_CHECK_CUDA(launch_pdl(ln_bwd_kernel<false>, dim3(row_grid(rows)), dim3(kRowThreads), 0, s, p));
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}

extern "C" int lav_scale_cast_f16(const float* x, int64_t ldx, const int32_t* row_map, const float* row_scale,
                                  int rows_per_scale, float alpha, void* out16, int64_t ldo, int rows, int C,
                                  void* stream) {
  LAV_REQUIRE(x && out16, "lav_scale_cast_f16: null pointer");
  LAV_REQUIRE((ldo % 4) == 0, "lav_scale_cast_f16: output ld must be %%4");
  if (rows <= 0) return LAV_OK;
  if (rows <= 2048 && C >= 8192) {  // few wide rows: (row, column slab) blocks
    LAV_CHECK_CUDA(launch_pdl(scale_cast_wide_kernel, dim3((C + 2047) / 2048, rows), dim3(256), 0, (cudaStream_t)stream, x, ldx,
                              row_map, row_scale, rows_per_scale > 0 ? rows_per_scale : 1, alpha, (__half*)out16, ldo, C));
    LAV_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return LAV_OK;
  }
  LAV_CHECK_CUDA(launch_pdl(scale_cast_kernel, dim3(row_grid(rows)), dim3(kRowThreads), 0, (cudaStream_t)stream, 
      x, ldx, row_map, row_scale, rows_per_scale > 0 ? rows_per_scale : 1, alpha, (__half*)out16, ldo, rows, C));
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}

extern "C" int lav_split3_f16(const float* x, int64_t ldx, void* out16, int64_t ldo, int rows, int C, int mode,
                              void* stream) {
  LAV_REQUIRE(x && out16 && C > 0 && ldo >= 3 * (int64_t)C && (mode == 0 || mode == 1), "lav_split3_f16: bad arguments");
  if (rows <= 0) return LAV_OK;
  LAV_CHECK_CUDA(launch_pdl(split3_kernel, dim3(row_grid(rows)), dim3(kRowThreads), 0, (cudaStream_t)stream, x, ldx,
                            (__half*)out16, ldo, rows, C, mode));
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}

extern "C" int lav_cast_f32_to_f16(const float* src, void* dst, int64_t n, void* stream) {
  LAV_REQUIRE(src && dst, "lav_cast_f32_to_f16: null pointer");
  LAV_REQUIRE(((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 8) == 0, "lav_cast_f32_to_f16: unaligned");
  if (n <= 0) return LAV_OK;
  int grid = (int)std::min<int64_t>((n / 4 + 255) / 256 + 1, (int64_t)sm_count() * 8);
  LAV_CHECK_CUDA(launch_pdl(cast_flat_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, src, (__half*)dst, n));
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}

extern "C" int lav_colsum_f16(const void* x16, int64_t ld, int rows, int N, float* out, float alpha, void* stream) {
  LAV_REQUIRE(x16 && out, "lav_colsum_f16: null pointer");
  LAV_REQUIRE((ld % 8) == 0 && ((uintptr_t)x16 % 16) == 0, "lav_colsum_f16: need ld %% 8 == 0 and a 16-byte aligned base");
  if (rows <= 0 || N <= 0) return LAV_OK;
  const int cb = (N + 255) / 256;
  const int rpb = slab_rows(rows, cb);
  dim3 grid(cb, (rows + rpb - 1) / rpb);
  LAV_CHECK_CUDA(launch_pdl(colsum_f16_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const __half*)x16, ld, rows, N, out, alpha, rpb));
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}

extern "C" int lav_gelu_bwd_f16(const void* dy16, const void* pre16, void* out16, int64_t n, void* stream) {
  LAV_REQUIRE(dy16 && pre16 && out16 && (n % 2) == 0, "lav_gelu_bwd_f16: bad arguments");
  if (n <= 0) return LAV_OK;
  int grid = (int)std::min<int64_t>((n / 2 + 255) / 256, (int64_t)sm_count() * 8);
  LAV_CHECK_CUDA(launch_pdl(gelu_bwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const __half2*)dy16, (const __half2*)pre16, (__half2*)out16, n / 2));
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}

extern "C" int lav_dropout_f32(const float* x, int64_t ldx, float* out, int64_t ldo, int rows, int C,
                               const LavDropout* drop, void* stream) {
  LAV_REQUIRE(x && out && drop && drop->rng, "lav_dropout_f32: null pointer");
  LAV_REQUIRE((C % 8) == 0 && (ldx % 4) == 0 && (ldo % 4) == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0,
              "lav_dropout_f32: need C %% 8 == 0 and 16-byte aligned rows");
  if (rows <= 0 || C <= 0) return LAV_OK;
  DropParams d = make_drop(drop);
  if (!d.on) d.on = 1, d.rng = drop->rng, d.site = drop->site, d.thresh = 0, d.inv_keep = 1.f;  // p == 0: copy
  const int64_t n = (int64_t)rows * (C / 8);
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8);
  LAV_CHECK_CUDA(launch_pdl(dropout_f32_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x, ldx, out, ldo, rows, C / 8, d));
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}

extern "C" int lav_dropout_mask(uint8_t* keep, int rows, int C, int head, const LavDropout* drop, void* stream) {
  LAV_REQUIRE(keep && drop && drop->rng, "lav_dropout_mask: null pointer");
  if (rows <= 0 || C <= 0) return LAV_OK;
  DropParams d = make_drop(drop);
  if (!d.on) d.on = 1, d.rng = drop->rng, d.site = drop->site, d.thresh = 0, d.inv_keep = 1.f;
  const int64_t n = (int64_t)rows * ((C + 7) / 8);
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8);
  LAV_CHECK_CUDA(launch_pdl(dropout_mask_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, keep, rows, C, head, d));
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}
