// Blocked ("flash") attention forward for sm_100a: any sequence / window length, 2 CTAs per SM.
//   O = softmax(scale * Q K^T + bias) V   per (problem, head, 128-row query tile), streaming 128-key blocks:
//     TMA   : Q once, K_c / V_c double-buffered (+ the 128 x 128 tile of the dense relative-position bias)
//     MMA   : S_c = Q K_c^T (+ I * Bias_c, see attention_fwd.cu) -> TMEM[0,128) ;  O (+)= P_c V_c -> TMEM[128,128+HD)
//     warps : online softmax, thread = query row: running max m and sum l in registers, P_c (fp16) to shared memory,
//             O rescaled in TMEM by exp(m_old - m_new) (tcgen05.ld / tcgen05.st) when a block raises the maximum
//   Shared memory per CTA: 112 KB (HD 64) / 80 KB (HD 32 + bias), TMEM 256 columns -> two CTAs per SM overlap each
//   other's load / MMA / softmax phases.  Used for HF BertSelfAttention (HD 64, L = 283 / 284 in pre-training, up to
//   758 with 384^2 frames) and for WindowAttention3D windows longer than 256 tokens (8 x 12 x 12 -> N = 720);
//   windows of <= 256 tokens keep the one-shot kernel of attention_fwd.cu.
#include "rng.cuh"
#include "runtime.h"
#include "sm100.cuh"

namespace lav {

constexpr int kFlashThreads = 160;  // warps 0-3: softmax (one TMEM lane quarter each); warp 4: TMA + MMA + TMEM alloc
constexpr int kFlashIdentBytes = 30 * 256;

struct FlashParams {
  int L, nheads, nprob;
  int q_off, k_off, v_off;
  float scale;
  int NPb;                                  // dense bias [ncls][nheads][NPb][NPb] (values / scale), via tmBias
  int has_bias;
  const int32_t* prob_class; int period;
  const float* key_bias; int NPk;           // [nprob][NPk] additive (0 / -inf) or null
  int causal_from;                          // >= 0: keys j >= causal_from are visible to queries i >= j only (seq2seq)
  __half* out; int64_t ldo;
  float* out32; int64_t ldo32;              // optional fp32 copy of O (high-precision mode)
  float* lse; int64_t rows_total;
  DropParams drop;
};

template <int HD, bool BMMA>
struct FlashCfg {
  static constexpr int ROWB = HD * 2;
  static constexpr int TILE = 128 * ROWB;                      // one Q / K / V block
  static constexpr int OFF_Q = 0, OFF_KV = TILE;               // stage s: K at OFF_KV + s*2*TILE, V right behind it
  static constexpr int OFF_P = OFF_KV + 4 * TILE;              // P block (fp16 [128][128]); first the bias tile (BMMA)
  static constexpr int OFF_ID = OFF_P + 32768;
  static constexpr int OFF_BAR = OFF_ID + (BMMA ? kFlashIdentBytes : 0);
  static constexpr int SMEM_BYTES = OFF_BAR + 128;
  static constexpr uint32_t SWZ = (HD == 64) ? SWZ_128B : SWZ_64B;
  static constexpr uint32_t SBO = 8 * ROWB;
  static constexpr int COL_S = 0, COL_O = 128;
};

template <int HD, bool BMMA>
__global__ void __launch_bounds__(kFlashThreads)
attn_fwd_flash_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmBias,
                      const FlashParams p) {
  griddep_launch();  // dependents (GEMMs) may start their prologue under this kernel's tail
  using Cfg = FlashCfg<HD, BMMA>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  // barriers: 0,1 kv_full[stage] | 2 bias_full | 3 s_ready | 4 p_ready | 5 o_ready
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x, h = blockIdx.y, prob = blockIdx.z;
  const int row0 = prob * p.L;
  const int nblk = (p.L + 127) >> 7;

  if (BMMA && warp < 4) {  // identity strip (attention_fwd.cu): zeros with a 16 x 16 identity block at groups 14-15
    uint8_t* id = smem + Cfg::OFF_ID;
    for (int i = threadIdx.x; i < kFlashIdentBytes / 16; i += 128) reinterpret_cast<uint4*>(id)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x < 16) {
      const int r = threadIdx.x;
      const int off = r < 8 ? 14 * 256 + r * 16 + r * 2 : 15 * 256 + 128 + (r - 8) * 16 + (r - 8) * 2;
      *reinterpret_cast<__half*>(id + off) = __float2half_rn(1.0f);
    }
    fence_proxy_async_smem();
  }
  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV);
      if (BMMA) tma_prefetch_desc(&tmBias);
      mbar_init(bars + 0, 1);
      mbar_init(bars + 1, 1);
      mbar_init(bars + 2, 1);
      mbar_init(bars + 3, 1);
      mbar_init(bars + 4, 128);
      mbar_init(bars + 5, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }
  tc_fence_before();
  griddep_wait();  // the prologue above touched no global data; from here on it does (PDL, see gemm.cu)
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      const int bcls = (BMMA && p.prob_class) ? p.prob_class[prob % p.period] : 0;
      const int brow = (bcls * p.nheads + h) * p.NPb + t * 128;
      auto load_kv = [&](int c) {
        const int s = c & 1;
        uint8_t* sk = smem + Cfg::OFF_KV + s * 2 * Cfg::TILE;
        mbar_arrive_expect_tx(bars + s, (c == 0 ? 3 : 2) * Cfg::TILE);
        if (c == 0) tma_load_2d(smem + Cfg::OFF_Q, &tmQKV, bars + s, p.q_off + h * HD, row0 + t * 128);
        tma_load_2d(sk, &tmQKV, bars + s, p.k_off + h * HD, row0 + c * 128);
        tma_load_2d(sk + Cfg::TILE, &tmQKV, bars + s, p.v_off + h * HD, row0 + c * 128);
      };
      auto load_bias = [&](int c) {
        mbar_arrive_expect_tx(bars + 2, 32768);
        tma_load_2d(smem + Cfg::OFF_P, &tmBias, bars + 2, c * 128, brow);
        tma_load_2d(smem + Cfg::OFF_P + 16384, &tmBias, bars + 2, c * 128 + 64, brow);
      };
      auto issue_s = [&](int c) {
        const int s = c & 1;
        const int nc = min(128, (p.L - c * 128 + 31) & ~31);
        mbar_wait(bars + s, (c >> 1) & 1, 30);
        if (BMMA) mbar_wait(bars + 2, c & 1, 31);
        tc_fence_after();
        const uint32_t idesc_s = make_idesc_f16(128, nc, 0, 0);
        const uint32_t sq = smem_u32(smem + Cfg::OFF_Q), sk = smem_u32(smem + Cfg::OFF_KV + s * 2 * Cfg::TILE);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_f16_ss(tmem + Cfg::COL_S, make_smem_desc(sq + k * 32, 0, Cfg::SBO, Cfg::SWZ),
                      make_smem_desc(sk + k * 32, 0, Cfg::SBO, Cfg::SWZ), idesc_s, k > 0);
        if (BMMA) {
          constexpr uint32_t idesc_b = make_idesc_f16(128, 128, 0, 1);
          const uint32_t sid = smem_u32(smem + Cfg::OFF_ID), sbias = smem_u32(smem + Cfg::OFF_P);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_f16_ss(tmem + Cfg::COL_S, make_smem_desc(sid + (14 - 2 * kk) * 256, 128, 256, SWZ_NONE),
                        make_smem_desc(sbias + kk * 2048, 16384, 1024, SWZ_128B), idesc_b, 1u);
        }
        umma_commit(bars + 3);
      };
      load_kv(0);
      if (BMMA) load_bias(0);
      if (nblk > 1) load_kv(1);
      issue_s(0);
      constexpr uint32_t idesc_o = make_idesc_f16(128, HD, 0, 1);
      const uint32_t sp = smem_u32(smem + Cfg::OFF_P);
      for (int c = 0; c < nblk; ++c) {
        const int s = c & 1;
        const int nc = min(128, (p.L - c * 128 + 31) & ~31);
        const uint32_t sv = smem_u32(smem + Cfg::OFF_KV + s * 2 * Cfg::TILE + Cfg::TILE);
        mbar_wait(bars + 4, c & 1, 32);  // P_c in shared memory, O rescaled
        tc_fence_after();
#pragma unroll 4
        for (int k = 0; k < nc / 16; ++k)
          umma_f16_ss(tmem + Cfg::COL_O, make_smem_desc(sp + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024, SWZ_128B),
                      make_smem_desc(sv + k * 16 * Cfg::ROWB, 0, Cfg::SBO, Cfg::SWZ), idesc_o, (c > 0 || k > 0) ? 1u : 0u);
        umma_commit(bars + 5);
        if (!BMMA && c + 1 < nblk) issue_s(c + 1);  // S of the next block does not wait for this PV
        if (c + 1 < nblk) {
          mbar_wait(bars + 5, c & 1, 33);           // PV_c done: stage s and the P buffer are free again
          if (c + 2 < nblk) load_kv(c + 2);
          if (BMMA) {
            load_bias(c + 1);
            issue_s(c + 1);
          }
        }
      }
    }
  } else {
    const int i = warp * 32 + lane;
    const int qi = t * 128 + i;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const float sc2 = p.scale * 1.4426950408889634f;  // scores in log2 units
    DropKey dkey{};
    if (p.drop.on) dkey = drop_key(p.drop);
    uint8_t* prow = smem + Cfg::OFF_P + i * 128;
    float m = -INFINITY, l = 0.f;  // running max (log2 units) and sum

    for (int c = 0; c < nblk; ++c) {
      const int nc = min(128, (p.L - c * 128 + 31) & ~31);
      const int nvalid = p.L - c * 128;  // columns >= nvalid are padding
      const float* kb = p.key_bias ? p.key_bias + (size_t)prob * p.NPk + c * 128 : nullptr;
      auto scores = [&](const uint32_t(&sraw)[32], int j0, float(&v)[32]) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(sraw[j]) * sc2;
        if (kb) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(kb + j0) + j);
            v[4 * j] += f.x * 1.4426950408889634f, v[4 * j + 1] += f.y * 1.4426950408889634f;
            v[4 * j + 2] += f.z * 1.4426950408889634f, v[4 * j + 3] += f.w * 1.4426950408889634f;
          }
        }
        if (j0 + 32 > nvalid) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j0 + j >= nvalid) v[j] = -INFINITY;
        }
        // seq2seq mask of LAVENDER_Base.get_attn_mask (model.py:208-218): the text keys (j >= causal_from) are seen
        // causally by the text queries and not at all by the video / prefix queries, i.e. exactly when j <= i
        if (p.causal_from >= 0 && c * 128 + j0 + 31 >= p.causal_from) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = c * 128 + j0 + j;
            if (col >= p.causal_from && col > qi) v[j] = -INFINITY;
          }
        }
      };
      mbar_wait(bars + 3, c & 1, 34);
      tc_fence_after();
      float mb = m;
#pragma unroll 1
      for (int j0 = 0; j0 < nc; j0 += 32) {
        uint32_t sraw[32];
        float v[32];
        tmem_ld_32x32(trow + Cfg::COL_S + j0, sraw);
        tmem_ld_wait();
        scores(sraw, j0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) mb = fmaxf(mb, v[j]);
      }
      const float m_use = mb == -INFINITY ? 0.f : mb;
      const float alpha = exp2f(m - m_use);  // m = -inf on the first block -> 0
      l *= alpha;
      if (c > 0) {  // PV_{c-1} must have landed before O is rescaled and before P is overwritten
        mbar_wait(bars + 5, (c - 1) & 1, 35);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
          for (int c0 = 0; c0 < HD; c0 += 32) {
            uint32_t o[32];
            tmem_ld_32x32(trow + Cfg::COL_O + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
            tmem_st_32x32(trow + Cfg::COL_O + c0, o);
          }
          tmem_st_wait();
        }
      }
      m = mb;
#pragma unroll 1
      for (int j0 = 0; j0 < nc; j0 += 32) {
        uint32_t sraw[32];
        float v[32];
        tmem_ld_32x32(trow + Cfg::COL_S + j0, sraw);
        tmem_ld_wait();
        scores(sraw, j0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = exp2f(v[j] - m_use);
          l += v[j];
        }
        if (p.drop.on) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t mk = drop_keep8(dkey, p.drop.thresh, (uint32_t)(row0 + qi),
                                           (uint32_t)(((c * 128 + j0) >> 3) + j), (uint32_t)h);
#pragma unroll
            for (int q = 0; q < 8; ++q) v[8 * j + q] = ((mk >> q) & 1u) ? v[8 * j + q] * p.drop.inv_keep : 0.f;
          }
        }
        uint8_t* atom = prow + (j0 >> 6) * 16384;
        const int chunk0 = (j0 & 63) >> 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          u.x = pack_half2(v[8 * j], v[8 * j + 1]), u.y = pack_half2(v[8 * j + 2], v[8 * j + 3]);
          u.z = pack_half2(v[8 * j + 4], v[8 * j + 5]), u.w = pack_half2(v[8 * j + 6], v[8 * j + 7]);
          *reinterpret_cast<uint4*>(atom + (((chunk0 + j) ^ (i & 7)) << 4)) = u;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bars + 4);
    }

    mbar_wait(bars + 5, (nblk - 1) & 1, 36);
    tc_fence_after();
    const float inv = 1.f / l;
    const bool valid = qi < p.L;
    if (valid && p.lse) p.lse[(size_t)h * p.rows_total + row0 + qi] = (m == -INFINITY ? 0.f : m) * 0.6931471805599453f + __logf(l);
#pragma unroll
    for (int c0 = 0; c0 < HD; c0 += 32) {
      uint32_t o[32];
      tmem_ld_32x32(trow + Cfg::COL_O + c0, o);
      tmem_ld_wait();
      if (valid && p.out) {
        __half* dst = p.out + (size_t)(row0 + qi) * p.ldo + h * HD + c0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          u.x = pack_half2(__uint_as_float(o[8 * j]) * inv, __uint_as_float(o[8 * j + 1]) * inv);
          u.y = pack_half2(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv);
          u.z = pack_half2(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv);
          u.w = pack_half2(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv);
          reinterpret_cast<uint4*>(dst)[j] = u;
        }
      }
      if (valid && p.out32) {
        float* dst = p.out32 + (size_t)(row0 + qi) * p.ldo32 + h * HD + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          reinterpret_cast<float4*>(dst)[j] =
              make_float4(__uint_as_float(o[4 * j]) * inv, __uint_as_float(o[4 * j + 1]) * inv,
                          __uint_as_float(o[4 * j + 2]) * inv, __uint_as_float(o[4 * j + 3]) * inv);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<256>(tmem);
}

template <int HD, bool BMMA>
static int launch_flash(const void* qkv, int64_t ld, int64_t rows_total, const void* bias16, const FlashParams& p,
                        cudaStream_t s) {
  using Cfg = FlashCfg<HD, BMMA>;
  CUtensorMap tm, tmb;
  int rc = encode_tmap_2d_f16(&tm, qkv, rows_total, ld, ld, 128, HD,
                              HD == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  tmb = tm;
  if (BMMA) {
    rc = encode_tmap_2d_f16(&tmb, bias16, (uint64_t)8 * p.nheads * p.NPb, p.NPb, p.NPb, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  auto kern = attn_fwd_flash_kernel<HD, BMMA>;
  static bool attr_set = false;
  if (!attr_set) {
    LAV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((p.L + 127) / 128, p.nheads, p.nprob);
  LAV_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(kFlashThreads), Cfg::SMEM_BYTES, s, tm, tmb, p));
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}

// called by lav_attn_fwd_f16 (attention_fwd.cu) for the shapes the one-shot kernel does not take
int attn_fwd_flash(const void* qkv, int64_t ld, int64_t rows_total, int q_off, int k_off, int v_off, int head_dim,
                   int nheads, int nprob, int L, float scale, const void* bias16, int NPb, const int32_t* prob_class,
                   int class_period, const float* key_bias, int NPk, int causal_from, void* out16, int64_t ldo,
                   float* out32, int64_t ldo32, float* lse, const LavDropout* drop, cudaStream_t s) {
  FlashParams p;
  p.causal_from = causal_from;
  p.L = L, p.nheads = nheads, p.nprob = nprob, p.q_off = q_off, p.k_off = k_off, p.v_off = v_off, p.scale = scale;
  p.NPb = NPb, p.has_bias = bias16 != nullptr, p.prob_class = prob_class, p.period = class_period > 0 ? class_period : 1;
  p.key_bias = key_bias, p.NPk = NPk, p.out = (__half*)out16, p.ldo = ldo, p.lse = lse, p.rows_total = rows_total;
  p.out32 = out32, p.ldo32 = ldo32;
  p.drop = make_drop(drop);
  const int nblk = (L + 127) / 128;
  LAV_REQUIRE(!bias16 || NPb >= nblk * 128, "lav_attn_fwd_f16: dense bias smaller than the padded length");
  LAV_REQUIRE(!key_bias || NPk >= nblk * 128, "lav_attn_fwd_f16: key_bias rows shorter than the padded length");
  if (head_dim == 64) {
    LAV_REQUIRE(!bias16, "lav_attn_fwd_f16: a dense bias is supported for head_dim 32 only");
    return launch_flash<64, false>(qkv, ld, rows_total, bias16, p, s);
  }
  if (head_dim == 32) {
    if (bias16) return launch_flash<32, true>(qkv, ld, rows_total, bias16, p, s);
    return launch_flash<32, false>(qkv, ld, rows_total, bias16, p, s);
  }
  return set_error(LAV_E_INVALID, "lav_attn_fwd_f16: unsupported head_dim %d (32 or 64)", head_dim);
}

}  // namespace lav
