// Fused attention forward for sm_100a:  O = softmax(scale * Q Kᵀ + bias) V   per (problem, head, 128-row q tile).
//   * problem = one 3-D shifted window (WindowAttention3D.forward, video_swin.py:145-170; L = 245 tokens, hd = 32)
//               or one BERT sequence (HF BertSelfAttention via model.py:242; L <= 384, hd = 64).
//   * Q/K/V are column slices of the fused QKV activation [rows, ld] (rows of a problem are contiguous), staged
//     by TMA; S = Q Kᵀ is a tcgen05 MMA into TMEM (128 x NKC*128 fp32), softmax runs in registers straight from
//     tcgen05.ld (one thread per query row), P goes to shared memory as fp16 in the UMMA K-major 128B-swizzle
//     layout, O = P V is a second tcgen05 MMA whose accumulator aliases the dead S columns.
//   * relative-position bias + shift mask (+ a large negative on the padded key columns) come pre-expanded as a dense
//     fp16 [class][head][NP][NP] tensor (lav_relpos_bias_expand, values already divided by `scale`).  The window
//     kernel TMA-stages the 128 x 256 bias tile into the shared memory that later holds P; each softmax thread adds
//     its own row of it to the S values it reads from TMEM (round 1 added it on the tensor core through an identity
//     MMA: 8 more tcgen05.mma per tile on an issue-bound warp, removed in round 2).  BERT passes an additive per-key
//     fp32 row instead.
// Nothing of size [B_, nh, N, N] ever reaches HBM.
#include <cstdlib>

#include "rng.cuh"
#include "runtime.h"
#include "sm100.cuh"

namespace lav {

// warps 0-7: softmax — warp w owns the TMEM lane quarter (w & 3) and the key-column half (w >> 2) of the score tile, i.e.
// two threads per query row (the one-thread-per-row chain over 256 columns was the critical path of a CTA);
// warp 8: TMA + MMA + TMEM alloc
constexpr int kAttnThreads = 288;

struct AttnFwdParams {
  int L, nheads, nprob, HD;
  int q_off, k_off, v_off;
  float scale;
  const __half* bias16; int NPb;           // dense bias [ncls][nheads][NPb][NPb] or null
  const int32_t* prob_class; int period;   // class of problem p = prob_class[p % period] (null: class 0)
  const float* key_bias;                   // [nprob][NKC*128] additive (0 / -inf) or null
  __half* out; int64_t ldo;
  float* out32; int64_t ldo32;             // optional fp32 copy of O (high-precision mode)
  float* lse; int64_t rows_total;
  DropParams drop;                         // attention-probability dropout (BERT, train mode)
  unsigned long long* trace;               // profiling: 8 clock64() stamps per CTA, or null
};

template <int HD, int NKC, bool BMMA>
struct AttnFwdCfg {
  static constexpr int ROWB = HD * 2;                    // bytes per Q/K/V smem row == TMA swizzle span
  static constexpr int Q_BYTES = 128 * ROWB;
  static constexpr int KV_BYTES = NKC * 128 * ROWB;      // per K or V
  static constexpr int P_BYTES = 128 * NKC * 128 * 2;    // bias tile (fp16 [128][NKC*128]) first, P afterwards
  static constexpr int OFF_K = Q_BYTES, OFF_V = OFF_K + KV_BYTES, OFF_P = OFF_V + KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 64;        // the dynamic window itself is 1024-byte aligned (checked)
  static constexpr int TMEM_COLS = (NKC * 128 <= 256) ? 256 : 512;
  static constexpr uint32_t SWZ = (HD == 64) ? SWZ_128B : SWZ_64B;
  static constexpr uint32_t SBO = 8 * ROWB;              // 8-row core-matrix group stride for Q/K/V tiles
};

template <int HD, int NKC, bool BMMA>
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmBias,
                const AttnFwdParams p) {
  griddep_launch();  // dependents (GEMMs) may start their prologue under this kernel's tail
  using Cfg = AttnFwdCfg<HD, NKC, BMMA>;
  const int cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  auto stamp = [&](int k) {
    if (p.trace) p.trace[(size_t)cta_lin * 8 + k] = clock64();
  };
  if (threadIdx.x == 0) stamp(0);
  constexpr int NP = NKC * 128;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // swizzled TMA / UMMA tiles need the 1024-byte alignment
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);  // 0 load, 1 S ready, 2 P ready, 3 O ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x, h = blockIdx.y, prob = blockIdx.z;
  const int row0 = prob * p.L;  // first token row of this problem
  const int ncol = min(NP, (p.L + 31) & ~31);  // key columns that can hold a valid key (BERT: 283 -> 288 of 384)

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV);
      if (BMMA) tma_prefetch_desc(&tmBias);
      mbar_init(bars + 0, 1);
      mbar_init(bars + 1, 1);
      mbar_init(bars + 2, 256);
      mbar_init(bars + 3, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  griddep_wait();  // the prologue above touched no global data; from here on it does (PDL, see gemm.cu)
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) stamp(1);

  if (warp == 8) {
    if (lane == 0) {
      // ---- loads
      mbar_arrive_expect_tx(bars + 0, Cfg::Q_BYTES + 2 * Cfg::KV_BYTES + (BMMA ? Cfg::P_BYTES : 0));
      if (BMMA) {  // bias tile rows t*128.., all NP key columns: NP/64 boxes of [128 rows x 64 columns]
        const int cls = p.prob_class ? p.prob_class[prob % p.period] : 0;
        const int brow = (cls * p.nheads + h) * p.NPb + t * 128;
#pragma unroll
        for (int jb = 0; jb < NP / 64; ++jb)
          tma_load_2d(smem + Cfg::OFF_P + jb * 16384, &tmBias, bars + 0, jb * 64, brow);
      }
      tma_load_2d(smem, &tmQKV, bars + 0, p.q_off + h * HD, row0 + t * 128);
#pragma unroll
      for (int c = 0; c < NKC; ++c) {
        tma_load_2d(smem + Cfg::OFF_K + c * 128 * Cfg::ROWB, &tmQKV, bars + 0, p.k_off + h * HD, row0 + c * 128);
        tma_load_2d(smem + Cfg::OFF_V + c * 128 * Cfg::ROWB, &tmQKV, bars + 0, p.v_off + h * HD, row0 + c * 128);
      }
      mbar_wait(bars + 0, 0, 10);
      stamp(2);
      tc_fence_after();
      // ---- S_c = Q K_cᵀ  (M=128, N=128, K=HD; both operands K-major)
      const uint32_t sq = smem_u32(smem), sk = smem_u32(smem + Cfg::OFF_K);
#pragma unroll
      for (int c = 0; c < NKC; ++c) {
        const int nc = min(128, ncol - c * 128);  // the last chunk may be narrow (MMA N is any multiple of 16)
        if (nc <= 0) break;
        const uint32_t idesc_s = make_idesc_f16(128, nc, 0, 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_f16_ss(tmem + c * 128, make_smem_desc(sq + k * 32, 0, Cfg::SBO, Cfg::SWZ),
                      make_smem_desc(sk + c * 128 * Cfg::ROWB + k * 32, 0, Cfg::SBO, Cfg::SWZ), idesc_s, k > 0);
      }
      umma_commit(bars + 1);
      // ---- O = P V  (M=128, N=HD, K=NP; A = P K-major 128B swizzle, B = V MN-major)
      mbar_wait(bars + 2, 0, 11);
      tc_fence_after();
      constexpr uint32_t idesc_o = make_idesc_f16(128, HD, 0, 1);
      const uint32_t sp = smem_u32(smem + Cfg::OFF_P), sv = smem_u32(smem + Cfg::OFF_V);
#pragma unroll 4
      for (int k = 0; k < ncol / 16; ++k)
        umma_f16_ss(tmem, make_smem_desc(sp + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024, SWZ_128B),
                    make_smem_desc(sv + k * 16 * Cfg::ROWB, 0, Cfg::SBO, Cfg::SWZ), idesc_o, k > 0);
      umma_commit(bars + 3);
    }
  } else {
    // ---- softmax: two threads per query row, each over one half of the key columns
    const int hsel = warp >> 2;
    const int i = (warp & 3) * 32 + lane;  // row within the q tile == TMEM lane
    const int qi = t * 128 + i;            // token index within the problem
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    constexpr int HW = HD / 2;             // O columns written by each thread of the pair
    constexpr int HC = NP / 2;             // key columns per thread
    const int jlo = hsel * HC, jhi = min(ncol, jlo + HC);
    const float sc = p.scale;
    const __half* brow = nullptr;
    if (!BMMA && p.bias16) {
      const int cls = p.prob_class ? p.prob_class[prob % p.period] : 0;
      brow = p.bias16 + (((size_t)cls * p.nheads + h) * p.NPb + min(qi, p.NPb - 1)) * p.NPb;
    }
    const float* kb = p.key_bias ? p.key_bias + (size_t)prob * NP : nullptr;
    DropKey dkey{};
    if (p.drop.on) dkey = drop_key(p.drop);
    // row-pair exchange (maximum, then partial sum) in the first 2 KB of the Q tile: Q is dead once S has been computed
    float* xch = reinterpret_cast<float*>(smem);
    const int pair_bar = 2 + (warp & 3);

    auto biased = [&](const uint32_t(&s)[32], int j0, float(&v)[32]) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(s[j]) * sc;
      if (BMMA) {
        // relative-position bias + shift mask: this row's 32 entries of the TMA-staged tile (same [128 rows x 64 keys]
        // 128B-swizzled atoms as P, which later overwrites them in place); the tile holds bias / scale
        const uint8_t* brow_s = smem + Cfg::OFF_P + (j0 >> 6) * 16384 + i * 128;
        const int chunk0 = (j0 & 63) >> 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 u = *reinterpret_cast<const uint4*>(brow_s + (((chunk0 + j) ^ (i & 7)) << 4));
          const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = __half22float2(hh[q]);
            v[8 * j + 2 * q] = fmaf(f.x, sc, v[8 * j + 2 * q]), v[8 * j + 2 * q + 1] = fmaf(f.y, sc, v[8 * j + 2 * q + 1]);
          }
        }
      } else if (brow) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u = *reinterpret_cast<const uint4*>(brow + j0 + 8 * j);
          const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float2 f = __half22float2(hh[q]);
            v[8 * j + 2 * q] += f.x, v[8 * j + 2 * q + 1] += f.y;
          }
        }
      }
      if (kb) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 f = __ldg(reinterpret_cast<const float4*>(kb + j0) + j);
          v[4 * j] += f.x, v[4 * j + 1] += f.y, v[4 * j + 2] += f.z, v[4 * j + 3] += f.w;
        }
      }
    };

    if (BMMA) mbar_wait(bars + 0, 0, 14);   // the TMA-staged bias tile (the MMA thread waited for it too)
    mbar_wait(bars + 1, 0, 12);
    if (threadIdx.x == 0) stamp(3);
    tc_fence_after();
    float m = -INFINITY;
#pragma unroll 1
    for (int j0 = jlo; j0 < jhi; j0 += 32) {
      uint32_t s[32];
      float v[32];
      tmem_ld_32x32(trow + j0, s);
      tmem_ld_wait();
      biased(s, j0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) m = fmaxf(m, v[j]);
    }
    if (threadIdx.x == 0) stamp(4);
    xch[hsel * 128 + i] = m;
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    m = fmaxf(m, xch[(hsel ^ 1) * 128 + i]);
    if (m == -INFINITY) m = 0.f;
    const float mlog = m * 1.4426950408889634f;
    float l = 0.f;
    uint8_t* prow = smem + Cfg::OFF_P + i * 128;
#pragma unroll 1
    for (int j0 = jlo; j0 < jhi; j0 += 32) {
      uint32_t s[32];
      float v[32];
      tmem_ld_32x32(trow + j0, s);
      tmem_ld_wait();
      biased(s, j0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = exp2f(v[j] * 1.4426950408889634f - mlog);
        l += v[j];
      }
      if (p.drop.on) {  // P -> keep * P / (1 - p); the normaliser l stays that of the un-dropped softmax
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t mk = drop_keep8(dkey, p.drop.thresh, (uint32_t)(row0 + qi), (uint32_t)((j0 >> 3) + j), (uint32_t)h);
#pragma unroll
          for (int q = 0; q < 8; ++q) v[8 * j + q] = ((mk >> q) & 1u) ? v[8 * j + q] * p.drop.inv_keep : 0.f;
        }
      }
      // P (fp16) -> smem, K-major 128B-swizzle atoms of [128 rows x 64 keys]
      uint8_t* atom = prow + (j0 >> 6) * 16384;
      const int chunk0 = (j0 & 63) >> 3;  // first 16-byte chunk inside the 128-byte row
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack_half2(v[8 * j], v[8 * j + 1]), u.y = pack_half2(v[8 * j + 2], v[8 * j + 3]);
        u.z = pack_half2(v[8 * j + 4], v[8 * j + 5]), u.w = pack_half2(v[8 * j + 6], v[8 * j + 7]);
        *reinterpret_cast<uint4*>(atom + (((chunk0 + j) ^ (i & 7)) << 4)) = u;
      }
    }
    xch[256 + hsel * 128 + i] = l;      // (second exchange array: the partner may still be reading the maximum)
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(bars + 2);
    if (threadIdx.x == 0) stamp(5);
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    l += xch[256 + (hsel ^ 1) * 128 + i];

    mbar_wait(bars + 3, 0, 13);
    if (threadIdx.x == 0) stamp(6);
    tc_fence_after();
    const float inv = 1.f / l;
    const bool valid = qi < p.L;
    if (valid && hsel == 0 && p.lse) p.lse[(size_t)h * p.rows_total + row0 + qi] = m + __logf(l);
    {
      uint32_t o[HW];
      tmem_ld_cols<HW>(trow + hsel * HW, o);
      tmem_ld_wait();
      if (valid && p.out) {
        __half* dst = p.out + (size_t)(row0 + qi) * p.ldo + h * HD + hsel * HW;
#pragma unroll
        for (int j = 0; j < HW / 8; ++j) {
          uint4 u;
          u.x = pack_half2(__uint_as_float(o[8 * j]) * inv, __uint_as_float(o[8 * j + 1]) * inv);
          u.y = pack_half2(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv);
          u.z = pack_half2(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv);
          u.w = pack_half2(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv);
          reinterpret_cast<uint4*>(dst)[j] = u;
        }
      }
      if (valid && p.out32) {
        float* dst = p.out32 + (size_t)(row0 + qi) * p.ldo32 + h * HD + hsel * HW;
#pragma unroll
        for (int j = 0; j < HW / 4; ++j)
          reinterpret_cast<float4*>(dst)[j] =
              make_float4(__uint_as_float(o[4 * j]) * inv, __uint_as_float(o[4 * j + 1]) * inv,
                          __uint_as_float(o[4 * j + 2]) * inv, __uint_as_float(o[4 * j + 3]) * inv);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) stamp(7);
  if (warp == 8) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

template <int HD, int NKC, bool BMMA>
static int launch_attn_fwd(const void* qkv, int64_t ld, int64_t rows_total, const AttnFwdParams& p, int ncls,
                           cudaStream_t s) {
  using Cfg = AttnFwdCfg<HD, NKC, BMMA>;
  CUtensorMap tm, tmb;
  int rc = encode_tmap_2d_f16(&tm, qkv, rows_total, ld, ld, 128, HD,
                              HD == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  tmb = tm;
  if (BMMA) {
    rc = encode_tmap_2d_f16(&tmb, p.bias16, (uint64_t)ncls * p.nheads * p.NPb, p.NPb, p.NPb, 128, 64,
                            CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  auto kern = attn_fwd_kernel<HD, NKC, BMMA>;
  static bool attr_set = false;
  if (!attr_set) {
    LAV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((p.L + 127) / 128, p.nheads, p.nprob);
  LAV_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(kAttnThreads), Cfg::SMEM_BYTES, s, tm, tmb, p));
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}

// dense[cls][h][i][j] = (table[rel_index(i,j)][h] + (label[cls][i] != label[cls][j] ? -100 : 0)) * inv_scale, and
// kMaskedKey for the padded key columns j >= L (video_swin.py:153-160, compute_mask :290-305).  rel_index is passed in
// (int32 [L][L], the [:N,:N] slice).  inv_scale = 1 / softmax scale: the attention kernels add the tile to the RAW
// Q K^T accumulator and apply `scale` to the sum.  kMaskedKey is finite on purpose: the flash kernel (attention_flash.cu)
// still adds the tile through an identity MMA, which multiplies every bias element by 0 or 1, and 0 * -inf would be NaN;
// exp() of it is exactly 0 all the same.
constexpr float kMaskedKey = -30000.0f;
__global__ void relpos_bias_expand_kernel(const float* table, int nheads, const int32_t* rel_index, int L,
                                          const uint8_t* labels, int ncls, __half* dense, int NP, float inv_scale) {
  griddep_launch();  // dependents (GEMMs) may start their prologue under this kernel's tail
  griddep_wait();    // launched with the PDL attribute: wait before touching global data
  const int i = blockIdx.x, h = blockIdx.y, cls = blockIdx.z;
  __half* drow = dense + (((size_t)cls * nheads + h) * NP + i) * NP;
  for (int j = threadIdx.x; j < NP; j += blockDim.x) {
    float v;
    if (j >= L) v = kMaskedKey;
    else if (i >= L) v = 0.f;
    else {
      v = table[(size_t)rel_index[i * L + j] * nheads + h];
      if (labels && labels[cls * NP + i] != labels[cls * NP + j]) v += -100.0f;
      v *= inv_scale;
    }
    drow[j] = __float2half_rn(v);
  }
}

int attn_fwd_flash(const void* qkv, int64_t ld, int64_t rows_total, int q_off, int k_off, int v_off, int head_dim,
                   int nheads, int nprob, int L, float scale, const void* bias16, int NPb, const int32_t* prob_class,
                   int class_period, const float* key_bias, int NPk, int causal_from, void* out16, int64_t ldo,
                   float* out32, int64_t ldo32, float* lse, const LavDropout* drop, cudaStream_t s);  // attention_flash.cu

}  // namespace lav

using namespace lav;

extern "C" int lav_attn_fwd_ex(const void* qkv, int64_t ld, int64_t rows_total, int q_off, int k_off, int v_off,
                               int head_dim, int nheads, int nprob, int L, float scale, const void* bias16, int NPb,
                               const int32_t* prob_class, int class_period, const float* key_bias, int NPk,
                               int causal_from, void* out16, int64_t ldo, float* out32, int64_t ldo32, float* lse,
                               const LavDropout* drop, void* stream) {
  LAV_REQUIRE(qkv && (out16 || out32), "lav_attn_fwd_f16: null pointer");
  LAV_REQUIRE(!out32 || ((ldo32 % 4) == 0 && ((uintptr_t)out32 % 16) == 0), "lav_attn_fwd_ex: out32 must be 16-byte aligned");
  LAV_REQUIRE(nprob > 0 && nheads > 0 && L > 0, "lav_attn_fwd_f16: empty problem");
  LAV_REQUIRE((ldo % 8) == 0 && (q_off % 8) == 0 && (k_off % 8) == 0 && (v_off % 8) == 0,
              "lav_attn_fwd_f16: offsets / ld must be multiples of 8");
  AttnFwdParams p;
  p.L = L, p.nheads = nheads, p.nprob = nprob, p.HD = head_dim;
  p.q_off = q_off, p.k_off = k_off, p.v_off = v_off, p.scale = scale;
  p.bias16 = (const __half*)bias16, p.NPb = NPb, p.prob_class = prob_class, p.period = class_period > 0 ? class_period : 1;
  p.key_bias = key_bias, p.out = (__half*)out16, p.ldo = ldo, p.lse = lse, p.rows_total = rows_total;
  p.out32 = out32, p.ldo32 = ldo32;
  p.drop = make_drop(drop);
  {
    const int64_t ctas = (int64_t)((L + 127) / 128) * nheads * nprob;
    p.trace = (trace_buffer() && trace_capacity() >= ctas * 8) ? trace_buffer() : nullptr;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int ncls = 8;  // row extent of the bias tensor map: an upper bound on the classes a dense tensor holds (2^3
                       // shifted axes); only the classes named by prob_class are ever addressed
  // one-shot kernels (whole key range in TMEM): windows of <= 256 tokens; LAV_ATTN_ONESHOT=1 also routes BERT
  // sequences of <= 384 tokens to the one-shot hd-64 kernel (one CTA per SM: A/B reference for the blocked kernel)
  static int oneshot64 = -1;
  if (oneshot64 < 0) {
    const char* e = getenv("LAV_ATTN_ONESHOT");
    oneshot64 = (e && e[0] == '1') ? 1 : 0;
  }
  if (causal_from < 0 && head_dim == 32 && L <= 256 && (!key_bias || NPk == 256)) {
    LAV_REQUIRE(!bias16 || NPb == 256, "lav_attn_fwd_f16: dense bias must be [*, *, 256, 256] for L <= 256");
    if (bias16) return launch_attn_fwd<32, 2, true>(qkv, ld, rows_total, p, ncls, s);
    return launch_attn_fwd<32, 2, false>(qkv, ld, rows_total, p, ncls, s);
  }
  if (causal_from < 0 && oneshot64 && head_dim == 64 && L <= 384 && !bias16 && (!key_bias || NPk == 384))
    return launch_attn_fwd<64, 3, false>(qkv, ld, rows_total, p, ncls, s);
  return attn_fwd_flash(qkv, ld, rows_total, q_off, k_off, v_off, head_dim, nheads, nprob, L, scale, bias16, NPb,
                        prob_class, class_period, key_bias, NPk, causal_from, out16, ldo, out32, ldo32, lse, drop, s);
}

extern "C" int lav_attn_fwd_f16(const void* qkv, int64_t ld, int64_t rows_total, int q_off, int k_off, int v_off,
                                int head_dim, int nheads, int nprob, int L, float scale, const void* bias16, int NPb,
                                const int32_t* prob_class, int class_period, const float* key_bias, int NPk,
                                int causal_from, void* out16, int64_t ldo, float* lse, const LavDropout* drop,
                                void* stream) {
  LAV_REQUIRE(out16, "lav_attn_fwd_f16: null pointer");
  return lav_attn_fwd_ex(qkv, ld, rows_total, q_off, k_off, v_off, head_dim, nheads, nprob, L, scale, bias16, NPb,
                         prob_class, class_period, key_bias, NPk, causal_from, out16, ldo, nullptr, 0, lse, drop, stream);
}

extern "C" int lav_relpos_bias_expand(const float* table, int nheads, const int32_t* rel_index, int L,
                                      const uint8_t* labels, int ncls, void* dense16, int NP, float inv_scale,
                                      void* stream) {
  LAV_REQUIRE(table && rel_index && dense16 && ncls >= 1 && L <= NP, "lav_relpos_bias_expand: bad arguments");
  dim3 grid(NP, nheads, ncls);
  LAV_CHECK_CUDA(launch_pdl(relpos_bias_expand_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, table, nheads, rel_index, L, labels, ncls,
                                                                   (__half*)dense16, NP, inv_scale));
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}
