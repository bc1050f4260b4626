// Fused optimizer step on the flat parameter / gradient arena (HBM-bound, ~16 B per parameter):
//   GradScaler.unscale_ + inf check, clip_grad_norm_, AdamW with per-group lr / weight decay, GradScaler.update
//   (Agent_Base.backward_step, agent.py:240-250, and build_optimizer, agent.py:96-140) without any host round trip.
#include "runtime.h"
#include "sm100.cuh"

namespace lav {

// state[] layout (fp32, device): 0 loss_scale | 1 growth_tracker | 2 step | 3 sumsq | 4 nonfinite | 5 grad_norm (out)
//                                6 found_inf (out) | 7 gmul = clip_coef / loss_scale | 8 bias_corr1 | 9 sqrt(bias_corr2)
constexpr int ST_SCALE = 0, ST_TRACK = 1, ST_STEP = 2, ST_SUMSQ = 3, ST_NONFIN = 4, ST_NORM = 5, ST_FOUND = 6, ST_GMUL = 7,
              ST_BC1 = 8, ST_BC2S = 9;

// Sum of squares + non-finite check of the gradient arena.  DETERMINISTIC: every block writes its partial sum to
// ws[blockIdx.x] and the last block to finish (ticket counter in ws[gridDim.x]) adds the partials in index order, so every
// data-parallel rank computes bit-identical norms / clip factors from the identical all-reduced gradients and the replicas'
// weights stay bit-identical (an atomicAdd per block would make the summation order - and the last bits - vary per rank).
__global__ void __launch_bounds__(256) grad_stats_kernel(const float* __restrict__ g, int64_t n, float* state, float* ws) {
  float ss = 0.f;
  int bad = 0;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    ss += g[i] * g[i];
    bad |= !isfinite(g[i]);
  }
  ss = warp_sum(ss);
  bad = __any_sync(0xffffffffu, bad);
  __shared__ float s_ss[8];
  __shared__ int s_bad[8];
  __shared__ bool s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_ss[warp] = ss, s_bad[warp] = bad;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    int b = 0;
    for (int w = 0; w < 8; ++w) t += s_ss[w], b |= s_bad[w];
    if (b || !isfinite(t)) atomicAdd(state + ST_NONFIN, 1.0f);   // (a flag: order-independent)
    if (ws == nullptr) {
      atomicAdd(state + ST_SUMSQ, t);
      s_last = false;
    } else {
      ws[blockIdx.x] = t;
      __threadfence();
      const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(ws + gridDim.x), 1u);
      s_last = ticket == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (s_last && warp == 0) {   // fixed-order tree over the partials: lane l adds ws[l], ws[l + 32], ...; then a shuffle tree
    __threadfence();
    float t = 0.f;
    for (unsigned i = lane; i < gridDim.x; i += 32) t += __ldcg(ws + i);
    t = warp_sum(t);
    if (lane == 0) {
      state[ST_SUMSQ] += t;
      *reinterpret_cast<unsigned*>(ws + gridDim.x) = 0u;   // ready for the next step
    }
  }
}

// group_step0[g]: number of (non-skipped) optimizer steps taken before group g's parameters received their first
// gradient (0 for parameters active from the start); a negative entry is filled in by this kernel on the group's first step.  torch.optim.AdamW keeps a step PER PARAMETER that only advances when the parameter has a
// gradient, so the bias corrections of late starters (emb_odr on the first `odr` batch, task heads) use
// step - step0.  group_bc[2g] = 1 - b1^t, group_bc[2g+1] = sqrt(1 - b2^t).
__global__ void adamw_prepare_kernel(float* state, float max_norm, float beta1, float beta2, float growth, float backoff,
                                     int growth_interval, float* group_step0, float* group_bc, int ngroups) {
  const float scale = state[ST_SCALE];
  const float inv = 1.0f / scale;
  const float sumsq = state[ST_SUMSQ];
  const bool found = state[ST_NONFIN] > 0.f || !isfinite(sumsq);
  const float norm = sqrtf(sumsq) * inv;                 // norm of the UNSCALED gradient (what clip_grad_norm_ sees)
  float coef = 1.0f;
  if (max_norm > 0.f) coef = fminf(1.0f, max_norm / (norm + 1e-6f));
  state[ST_NORM] = norm;
  state[ST_FOUND] = found ? 1.f : 0.f;
  state[ST_GMUL] = inv * coef;
  if (!found) {
    const float step = state[ST_STEP] + 1.0f;
    state[ST_STEP] = step;
    state[ST_BC1] = 1.0f - powf(beta1, step);
    state[ST_BC2S] = sqrtf(1.0f - powf(beta2, step));
    for (int g = 0; g < ngroups; ++g) {
      if (group_step0 && group_step0[g] < 0.f) group_step0[g] = step - 1.0f;   // the group's first step: recorded here
      const float t = fmaxf(1.0f, step - (group_step0 ? group_step0[g] : 0.f));
      group_bc[2 * g] = 1.0f - powf(beta1, t);
      group_bc[2 * g + 1] = sqrtf(1.0f - powf(beta2, t));
    }
  }
  // GradScaler.update (torch/amp/grad_scaler.py: _amp_update_scale_)
  if (found) {
    state[ST_SCALE] = scale * backoff;
    state[ST_TRACK] = 0.f;
  } else {
    const float tr = state[ST_TRACK] + 1.0f;
    if (growth_interval > 0 && tr >= (float)growth_interval) {
      state[ST_SCALE] = scale * growth;
      state[ST_TRACK] = 0.f;
    } else {
      state[ST_TRACK] = tr;
    }
  }
  state[ST_SUMSQ] = 0.f;
  state[ST_NONFIN] = 0.f;
}

// torch.optim.AdamW (decoupled weight decay, no amsgrad):
//   p *= 1 - lr*wd ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(256)
adamw_update_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                    int64_t nblocks8, const uint8_t* __restrict__ group_of_block, const float* __restrict__ group_lr,
                    const float* __restrict__ group_wd, float beta1, float beta2, float eps, const float* __restrict__ state,
                    const float* __restrict__ group_bc, __half* __restrict__ p16) {
  if (state[ST_FOUND] != 0.f) return;  // GradScaler.step: skip the update when a non-finite gradient was found
  const float gmul = state[ST_GMUL];
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nblocks8; b += (int64_t)gridDim.x * blockDim.x) {
    const int grp = group_of_block[b];
    if (grp == 255) continue;  // parameter without a gradient (optimizer skips it, as torch does for p.grad is None)
    const float lr = group_lr[grp], wd = group_wd[grp];
    const float bc1 = group_bc[2 * grp], bc2s = group_bc[2 * grp + 1];
    const float decay = 1.0f - lr * wd, step_size = lr / bc1;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t i = b * 2 + h;
      float4 pv = reinterpret_cast<float4*>(p)[i];
      const float4 gv = reinterpret_cast<const float4*>(g)[i];
      float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      float* pp = &pv.x;
      const float* gp = &gv.x;
      float* mp = &mv.x;
      float* vp = &vv.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gg = gp[j] * gmul;
        mp[j] = beta1 * mp[j] + (1.0f - beta1) * gg;
        vp[j] = beta2 * vp[j] + (1.0f - beta2) * gg * gg;
        pp[j] = pp[j] * decay - step_size * mp[j] / (sqrtf(vp[j]) / bc2s + eps);
      }
      reinterpret_cast<float4*>(p)[i] = pv;
      reinterpret_cast<float4*>(m)[i] = mv;
      reinterpret_cast<float4*>(v)[i] = vv;
      if (p16)  // fp16 shadow of the updated weights (tensor-core operand of the next step): no separate cast pass
        reinterpret_cast<uint2*>(p16)[i] = make_uint2(pack_half2(pv.x, pv.y), pack_half2(pv.z, pv.w));
    }
  }
}

}  // namespace lav

using namespace lav;

extern "C" int lav_grad_stats(const float* grad, int64_t n, float* state, float* ws, int64_t ws_floats, void* stream) {
  LAV_REQUIRE(grad && state && ((uintptr_t)grad % 16) == 0, "lav_grad_stats: bad arguments");
  if (n <= 0) return LAV_OK;
  int grid = (int)std::min<int64_t>((n / 4 + 255) / 256 + 1, (int64_t)sm_count() * 8);
  if (ws) {
    LAV_REQUIRE(ws_floats >= 2, "lav_grad_stats: workspace too small");
    grid = (int)std::min<int64_t>(grid, ws_floats - 1);
  }
  grad_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(grad, n, state, ws);
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}

extern "C" int lav_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                              const uint8_t* group_of_block, const float* group_lr, const float* group_wd, float beta1,
                              float beta2, float eps, float max_grad_norm, float* state, float growth_factor,
                              float backoff_factor, int growth_interval, void* param16, float* group_step0,
                              float* group_bc, int ngroups, void* stream) {
  LAV_REQUIRE(param && grad && exp_avg && exp_avg_sq && group_of_block && group_lr && group_wd && state && group_bc,
              "lav_adamw_step: null pointer");
  LAV_REQUIRE(ngroups > 0 && ngroups < 255, "lav_adamw_step: 1..254 parameter groups");
  LAV_REQUIRE((n % 8) == 0 && ((uintptr_t)param % 16) == 0 && ((uintptr_t)grad % 16) == 0 &&
                  ((uintptr_t)exp_avg % 16) == 0 && ((uintptr_t)exp_avg_sq % 16) == 0,
              "lav_adamw_step: buffers must be 16-byte aligned and n a multiple of 8");
  cudaStream_t s = (cudaStream_t)stream;
  adamw_prepare_kernel<<<1, 1, 0, s>>>(state, max_grad_norm, beta1, beta2, growth_factor, backoff_factor, growth_interval,
                                       group_step0, group_bc, ngroups);
  LAV_CHECK_CUDA(cudaGetLastError());
  if (n > 0) {
    const int64_t nb = n / 8;
    const int grid = (int)std::min<int64_t>((nb + 255) / 256, (int64_t)sm_count() * 8);
    adamw_update_kernel<<<grid, 256, 0, s>>>(param, grad, exp_avg, exp_avg_sq, nb, group_of_block, group_lr, group_wd, beta1,
                                            beta2, eps, state, group_bc, (__half*)param16);
    LAV_CHECK_CUDA(cudaGetLastError());
  }
  count_launch(2);
  return LAV_OK;
}
