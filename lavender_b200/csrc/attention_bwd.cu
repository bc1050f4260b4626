// Fused attention backward for sm_100a (recompute from the saved log-sum-exp; nothing [N,N]-shaped is kept from
// the forward).  Persistent CTAs (one per SM) walk work items (128-key chunk c, head, problem); an item keeps K_c / V_c
// resident (double-buffered against the next item's) and loops over the 128-row query tiles t:
//     S  = Q_t K_cᵀ , dP = dO_t V_cᵀ                 (tcgen05 -> TMEM)
//     P  = exp(scale*S + bias - lse) , dS = P ∘ (dP - delta)          (registers, thread = query row)
//     dV_c += Pᵀ dO_t , dK_c += dSᵀ Q_t , dQ_t = dS K_c               (tcgen05; P/dS staged in smem as fp16)
// dQ_t is reduced over the key chunks with fp32 atomics into `dq_acc`; dK/dV are written once as fp16.
// For the Swin relative-position bias the per-window dS is also written out (fp16) and reduced over windows by
// relpos_bias_grad_kernel (autograd's index_put of video_swin.py:153 in the reference).
#include "rng.cuh"
#include "runtime.h"
#include "sm100.cuh"

namespace lav {

constexpr int kAttnBwdThreads = 288;  // warps 0-7: softmax (lane quarter w & 3, key-column half w >> 2); warp 8: TMA + MMA + TMEM

struct AttnBwdParams {
  int L, nheads, nprob;
  int q_off, k_off, v_off;
  float scale;
  const __half* bias16; int NPb;
  const int32_t* prob_class; int period;
  const float* key_bias; int NPk;           // [nprob][NPk]
  int causal_from;                          // >= 0: keys j >= causal_from visible to queries i >= j only (seq2seq mask)
  const __half* out; int64_t ldo;           // forward output O
  const __half* dout; int64_t lddo;         // dO
  const float* lse; int64_t rows_total;
  const float* delta;                       // [nheads][rows_total]: rowsum(dO * O), from attn_delta_kernel
  float* dq_acc; int64_t lddq;              // fp32 [rows_total, nheads*HD], pre-zeroed
  __half* dqkv; int64_t lddqkv;             // dK / dV written at k_off / v_off
  __half* ds_out; int NPs;                  // optional [nprob][nheads][NPs][NPs]
  DropParams drop;                          // attention-probability dropout of the forward (regenerated here)
};

// BMMA (window attention with the dense relative-position bias): the 128 x 128 bias tile of each query tile is TMA-staged
// (double-buffered like Q / dO) and each softmax thread adds its own row of it to the S values it reads from TMEM.
template <int HD, bool BMMA>
struct AttnBwdCfg {
  static constexpr int ROWB = HD * 2;
  static constexpr int TILE = 128 * ROWB;
  // K / V double-buffered by work item (persistent CTAs prefetch the next item's chunk), Q / dO by (item, query tile) step
  static constexpr int OFF_K = 0, OFF_V = 2 * TILE, OFF_Q = 4 * TILE, OFF_DO = 6 * TILE;
  static constexpr int OFF_P = 8 * TILE, OFF_DS = OFF_P + 32768;
  static constexpr int OFF_BIAS = OFF_DS + 32768;                                     // 2 x [128][128] fp16
  static constexpr int OFF_BAR = OFF_BIAS + (BMMA ? 65536 : 0);
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
  static constexpr uint32_t SWZ = (HD == 64) ? SWZ_128B : SWZ_64B;
  static constexpr uint32_t SBO = 8 * ROWB;
  static constexpr int COL_S = 0, COL_DP = 128, COL_DQ = 256, COL_DK = 320, COL_DV = 384;
};

template <int HD, bool BMMA>
__global__ void __launch_bounds__(kAttnBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmBias, const AttnBwdParams p) {
  griddep_launch();  // dependents (GEMMs) may start their prologue under this kernel's tail
  using Cfg = AttnBwdCfg<HD, BMMA>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // barriers: 0,1 K/V buffer | 2,3 Q/dO(/bias) buffer | 4 S,dP ready | 5 P,dS ready | 6 dV,dK,dQ MMAs done | 7 dK,dV read out
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // PERSISTENT: one CTA per SM walks work items (key chunk c, head h, problem) = blockIdx.x, + gridDim.x, ...; an item is
  // nqt steps (query tiles).  Round 1 launched one CTA per item: with nqt = 2-3 the per-CTA prologue (512-column TMEM
  // allocation, barriers), the first K/V/Q/dO/bias loads and the dK/dV read-out were ~half of a CTA's life
  // and nothing overlapped them at 1 CTA/SM.  Now the next item's operands are prefetched and its S / dP products issued
  // while the softmax warps finish the current item.
  const int nkc = (p.L + 127) / 128;
  const int nqt = (p.L + 127) / 128;
  const int nitems = nkc * p.nheads * p.nprob;
  auto item_c = [&](int it) { return it % nkc; };
  auto item_h = [&](int it) { return (it / nkc) % p.nheads; };
  auto item_prob = [&](int it) { return it / (nkc * p.nheads); };

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV);
      tma_prefetch_desc(&tmDO);
      if (BMMA) tma_prefetch_desc(&tmBias);
      mbar_init(bars + 0, 1);
      mbar_init(bars + 1, 1);
      mbar_init(bars + 2, 1);
      mbar_init(bars + 3, 1);
      mbar_init(bars + 4, 1);
      mbar_init(bars + 5, 256);
      mbar_init(bars + 6, 1);
      mbar_init(bars + 7, 256);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  griddep_wait();  // the prologue above touched no global data; from here on it does (PDL, see gemm.cu)
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      constexpr uint32_t id_t = make_idesc_f16(128, HD, 1, 1);    // dV / dK: MN-major A (P/dS transposed), MN-major B
      constexpr uint32_t id_q = make_idesc_f16(128, HD, 0, 1);    // dQ     : K-major A (dS), MN-major B (K)
      const uint32_t sp = smem_u32(smem + Cfg::OFF_P), sds = smem_u32(smem + Cfg::OFF_DS);
      const int nmine = blockIdx.x < nitems ? (nitems - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;   // items of this CTA
      const int nsteps = nmine * nqt;
      auto item_of = [&](int k) { return (int)blockIdx.x + k * (int)gridDim.x; };
      auto jmax_of = [&](int it) { return min(128, (p.L - item_c(it) * 128 + 31) & ~31); };
      auto load_kv = [&](int k) {   // item ordinal k -> K/V buffer k & 1
        const int it = item_of(k), kb = k & 1;
        const int row0 = item_prob(it) * p.L;
        mbar_arrive_expect_tx(bars + kb, 2 * Cfg::TILE);
        tma_load_2d(smem + Cfg::OFF_K + kb * Cfg::TILE, &tmQKV, bars + kb, p.k_off + item_h(it) * HD, row0 + item_c(it) * 128);
        tma_load_2d(smem + Cfg::OFF_V + kb * Cfg::TILE, &tmQKV, bars + kb, p.v_off + item_h(it) * HD, row0 + item_c(it) * 128);
      };
      auto load_q = [&](int s) {    // step s = (item ordinal s / nqt, query tile s % nqt) -> Q/dO/bias buffer s & 1
        const int it = item_of(s / nqt), t = s % nqt, b = s & 1;
        const int h = item_h(it), row0 = item_prob(it) * p.L;
        mbar_arrive_expect_tx(bars + 2 + b, 2 * Cfg::TILE + (BMMA ? 32768 : 0));
        tma_load_2d(smem + Cfg::OFF_Q + b * Cfg::TILE, &tmQKV, bars + 2 + b, p.q_off + h * HD, row0 + t * 128);
        tma_load_2d(smem + Cfg::OFF_DO + b * Cfg::TILE, &tmDO, bars + 2 + b, h * HD, row0 + t * 128);
        if (BMMA) {  // bias rows t*128.., key columns c*128..: two boxes of [128 rows x 64 columns]
          const int bcls = p.prob_class ? p.prob_class[item_prob(it) % p.period] : 0;
          const int brow = (bcls * p.nheads + h) * p.NPb + t * 128;
          tma_load_2d(smem + Cfg::OFF_BIAS + b * 32768, &tmBias, bars + 2 + b, item_c(it) * 128, brow);
          tma_load_2d(smem + Cfg::OFF_BIAS + b * 32768 + 16384, &tmBias, bars + 2 + b, item_c(it) * 128 + 64, brow);
        }
      };
      // S = Q K^T (+ I * Bias), dP = dO V^T of step s.  Issued one step AHEAD of the softmax warps (also across items).
      auto issue_s_dp = [&](int s) {
        const int k = s / nqt, b = s & 1, kb = k & 1;
        const uint32_t sq = smem_u32(smem + Cfg::OFF_Q + b * Cfg::TILE);
        const uint32_t sdo = smem_u32(smem + Cfg::OFF_DO + b * Cfg::TILE);
        const uint32_t sk = smem_u32(smem + Cfg::OFF_K + kb * Cfg::TILE), sv = smem_u32(smem + Cfg::OFF_V + kb * Cfg::TILE);
        if (s % nqt == 0) mbar_wait(bars + kb, (k >> 1) & 1, 20);   // first step of an item: its K / V chunk has landed
        mbar_wait(bars + 2 + b, (s >> 1) & 1, 21);
        tc_fence_after();
        const uint32_t id_s = make_idesc_f16(128, jmax_of(item_of(k)), 0, 0);  // K-major x K-major, narrow last chunk
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk)
          umma_f16_ss(tmem + Cfg::COL_S, make_smem_desc(sq + kk * 32, 0, Cfg::SBO, Cfg::SWZ),
                      make_smem_desc(sk + kk * 32, 0, Cfg::SBO, Cfg::SWZ), id_s, kk > 0);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk)
          umma_f16_ss(tmem + Cfg::COL_DP, make_smem_desc(sdo + kk * 32, 0, Cfg::SBO, Cfg::SWZ),
                      make_smem_desc(sv + kk * 32, 0, Cfg::SBO, Cfg::SWZ), id_s, kk > 0);
        umma_commit(bars + 4);
      };
      if (nsteps > 0) {
        load_kv(0);
        load_q(0);
        if (nsteps > 1) load_q(1);
        if (nmine > 1) load_kv(1);
        issue_s_dp(0);
      }
      for (int s = 0; s < nsteps; ++s) {
        const int k = s / nqt, t = s % nqt, b = s & 1, kb = k & 1;
        const int jmax = jmax_of(item_of(k));
        const uint32_t sq = smem_u32(smem + Cfg::OFF_Q + b * Cfg::TILE);
        const uint32_t sdo = smem_u32(smem + Cfg::OFF_DO + b * Cfg::TILE);
        const uint32_t sk = smem_u32(smem + Cfg::OFF_K + kb * Cfg::TILE);
        mbar_wait(bars + 5, s & 1, 22);       // P, dS of this step are in shared memory
        if (t == 0 && k > 0) mbar_wait(bars + 7, (k - 1) & 1, 26);   // the previous item's dK / dV have been read out of TMEM
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // contraction over the 128 query rows, 16 per MMA
          const uint64_t b_do = make_smem_desc(sdo + kk * 16 * Cfg::ROWB, 0, Cfg::SBO, Cfg::SWZ);
          const uint64_t b_q = make_smem_desc(sq + kk * 16 * Cfg::ROWB, 0, Cfg::SBO, Cfg::SWZ);
          umma_f16_ss(tmem + Cfg::COL_DV, make_smem_desc(sp + kk * 2048, 16384, 1024, SWZ_128B), b_do, id_t,
                      (t > 0 || kk > 0));
          umma_f16_ss(tmem + Cfg::COL_DK, make_smem_desc(sds + kk * 2048, 16384, 1024, SWZ_128B), b_q, id_t,
                      (t > 0 || kk > 0));
        }
#pragma unroll 4
        for (int kk = 0; kk < jmax / 16; ++kk)  // contraction over the (valid) keys of this chunk
          umma_f16_ss(tmem + Cfg::COL_DQ, make_smem_desc(sds + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024, SWZ_128B),
                      make_smem_desc(sk + kk * 16 * Cfg::ROWB, 0, Cfg::SBO, Cfg::SWZ), id_q, kk > 0);
        umma_commit(bars + 6);
        if (s + 1 < nsteps) issue_s_dp(s + 1);
        if (s + 2 < nsteps || (t == nqt - 1 && k + 2 < nmine)) {
          mbar_wait(bars + 6, s & 1, 23);     // the MMAs above have consumed this step's Q / dO buffer (and, on an item's
          if (s + 2 < nsteps) load_q(s + 2);  // last step, its K / V buffer): refill them for step s + 2 / item k + 2
          if (t == nqt - 1 && k + 2 < nmine) load_kv(k + 2);
        }
      }
    }
  } else {
    // 8 softmax warps: warp w owns accumulator rows (w & 3) * 32 .. +31 (its TMEM lane quarter) and the key columns
    // [hsel * 64, hsel * 64 + 64) of the chunk (hsel = w >> 2): two threads per query row halve the latency-bound
    // exp / dS phase; no cross-thread reduction is needed in backward (lse and delta are known per row).
    const int hsel = warp >> 2;
    const int i = (warp & 3) * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    constexpr int HW = HD / 2;  // output columns of dQ / dK / dV handled by each half
    const float sc_log2 = p.scale * 1.4426950408889634f;
    uint8_t* prow = smem + Cfg::OFF_P + i * 128;
    uint8_t* dsrow = smem + Cfg::OFF_DS + i * 128;
    DropKey dkey{};
    if (p.drop.on) dkey = drop_key(p.drop);

    int s = 0;   // step counter of this CTA (same sequence as the MMA warp's)
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int c = item_c(item), h = item_h(item), prob = item_prob(item);
      const int row0 = prob * p.L;
      const int jmax = min(128, (p.L - c * 128 + 31) & ~31);  // key columns of this chunk that can hold a valid key
      const int cls = (p.bias16 && p.prob_class) ? p.prob_class[prob % p.period] : 0;
      const float* kb = p.key_bias ? p.key_bias + (size_t)prob * p.NPk + c * 128 : nullptr;
      for (int t = 0; t < nqt; ++t, ++s) {
        const int qi = t * 128 + i;
        const bool valid = qi < p.L;
        float lse_l2 = 0.f, delta = 0.f;
        if (valid) {  // two scalars per row, requested before the wait on the tensor core
          lse_l2 = p.lse[(size_t)h * p.rows_total + row0 + qi] * 1.4426950408889634f;
          delta = p.delta[(size_t)h * p.rows_total + row0 + qi];
        }
        const __half* brow = (!BMMA && p.bias16) ? p.bias16 + (((size_t)cls * p.nheads + h) * p.NPb + min(qi, p.NPb - 1)) * p.NPb + c * 128
                                      : nullptr;
        __half* dsg = (p.ds_out && valid) ? p.ds_out + (((size_t)prob * p.nheads + h) * p.NPs + qi) * p.NPs + c * 128
                                          : nullptr;
        if (BMMA) mbar_wait(bars + 2 + (s & 1), (s >> 1) & 1, 27);   // this step's bias tile (TMA) is in shared memory
        mbar_wait(bars + 4, s & 1, 24);
        tc_fence_after();
#pragma unroll 1
        for (int j0 = hsel * 64; j0 < min(jmax, hsel * 64 + 64); j0 += 32) {
          uint32_t sr[32], dp[32];
          tmem_ld_32x32(trow + Cfg::COL_S + j0, sr);
          tmem_ld_32x32(trow + Cfg::COL_DP + j0, dp);
          tmem_ld_wait();
          float pv[32], dsv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) pv[j] = __uint_as_float(sr[j]) * sc_log2 - lse_l2;
          if (BMMA) {
            // relative-position bias + shift mask from the TMA-staged [128 x 128] tile of this step (same 128B-swizzled
            // [128 rows x 64 keys] atoms as P / dS); it holds bias / scale, so one FMA brings it to log2 units
            const uint8_t* bt = smem + Cfg::OFF_BIAS + (s & 1) * 32768 + (j0 >> 6) * 16384 + i * 128;
            const int chunk0 = (j0 & 63) >> 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 u = *reinterpret_cast<const uint4*>(bt + (((chunk0 + j) ^ (i & 7)) << 4));
              const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 f = __half22float2(hh[q]);
                pv[8 * j + 2 * q] = fmaf(f.x, sc_log2, pv[8 * j + 2 * q]);
                pv[8 * j + 2 * q + 1] = fmaf(f.y, sc_log2, pv[8 * j + 2 * q + 1]);
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
                // the dense bias holds bias / scale (lav_relpos_bias_expand): back to log2 units with scale * log2(e)
                pv[8 * j + 2 * q] += f.x * sc_log2, pv[8 * j + 2 * q + 1] += f.y * sc_log2;
              }
            }
          }
          if (kb) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 f = __ldg(reinterpret_cast<const float4*>(kb + j0) + j);
              pv[4 * j] += f.x * 1.4426950408889634f, pv[4 * j + 1] += f.y * 1.4426950408889634f;
              pv[4 * j + 2] += f.z * 1.4426950408889634f, pv[4 * j + 3] += f.w * 1.4426950408889634f;
            }
          }
          if (p.causal_from >= 0 && c * 128 + j0 + 31 >= p.causal_from) {  // seq2seq mask, as in attention_flash.cu
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = c * 128 + j0 + j;
              if (col >= p.causal_from && col > qi) pv[j] = -INFINITY;
            }
          }
          if (!p.drop.on) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float pe = valid ? exp2f(pv[j]) : 0.f;
              pv[j] = pe;
              dsv[j] = pe * (__uint_as_float(dp[j]) - delta);
            }
          } else {  // O = (keep * P / (1-p)) V: dV uses the dropped P, dP = keep * dP_drop / (1-p), dS = P (dP - delta)
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              const uint32_t m = drop_keep8(dkey, p.drop.thresh, (uint32_t)(row0 + qi), (uint32_t)(((c * 128 + j0) >> 3) + j8),
                                            (uint32_t)h);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const int j = 8 * j8 + q;
                const float pe = valid ? exp2f(pv[j]) : 0.f;
                const float kc = ((m >> q) & 1u) ? p.drop.inv_keep : 0.f;
                pv[j] = pe * kc;
                dsv[j] = pe * (kc * __uint_as_float(dp[j]) - delta);
              }
            }
          }
          uint8_t* pa = prow + (j0 >> 6) * 16384;
          uint8_t* da = dsrow + (j0 >> 6) * 16384;
          const int chunk0 = (j0 & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u, w;
            u.x = pack_half2(pv[8 * j], pv[8 * j + 1]), u.y = pack_half2(pv[8 * j + 2], pv[8 * j + 3]);
            u.z = pack_half2(pv[8 * j + 4], pv[8 * j + 5]), u.w = pack_half2(pv[8 * j + 6], pv[8 * j + 7]);
            w.x = pack_half2(dsv[8 * j], dsv[8 * j + 1]), w.y = pack_half2(dsv[8 * j + 2], dsv[8 * j + 3]);
            w.z = pack_half2(dsv[8 * j + 4], dsv[8 * j + 5]), w.w = pack_half2(dsv[8 * j + 6], dsv[8 * j + 7]);
            const int off = ((chunk0 + j) ^ (i & 7)) << 4;
            *reinterpret_cast<uint4*>(pa + off) = u;
            *reinterpret_cast<uint4*>(da + off) = w;
            if (dsg) reinterpret_cast<uint4*>(dsg + j0)[j] = w;
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bars + 5);

        mbar_wait(bars + 6, s & 1, 25);
        tc_fence_after();
        {
          uint32_t o[HW];
          tmem_ld_cols<HW>(trow + Cfg::COL_DQ + hsel * HW, o);
          tmem_ld_wait();
          if (valid) {
            float* dst = p.dq_acc + (size_t)(row0 + qi) * p.lddq + h * HD + hsel * HW;
#pragma unroll
            for (int j = 0; j < HW / 4; ++j)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j),
                           "f"(__uint_as_float(o[4 * j]) * p.scale), "f"(__uint_as_float(o[4 * j + 1]) * p.scale),
                           "f"(__uint_as_float(o[4 * j + 2]) * p.scale), "f"(__uint_as_float(o[4 * j + 3]) * p.scale)
                           : "memory");
          }
        }
        tc_fence_before();
      }
      // ---- dK_c, dV_c of this item: thread = key row of the chunk (the last step's MMAs have completed: bars + 6 above)
      const int kj = c * 128 + i;
      const bool kvalid = kj < p.L;
      {
        uint32_t dk[HW], dv[HW];
        tmem_ld_cols<HW>(trow + Cfg::COL_DK + hsel * HW, dk);
        tmem_ld_cols<HW>(trow + Cfg::COL_DV + hsel * HW, dv);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(bars + 7);   // the accumulators may be overwritten by the next item's first dV / dK MMAs
        if (kvalid) {
          __half* gk = p.dqkv + (size_t)(row0 + kj) * p.lddqkv + p.k_off + h * HD + hsel * HW;
          __half* gv = p.dqkv + (size_t)(row0 + kj) * p.lddqkv + p.v_off + h * HD + hsel * HW;
#pragma unroll
          for (int j = 0; j < HW / 8; ++j) {
            uint4 u, w;
            u.x = pack_half2(__uint_as_float(dk[8 * j]) * p.scale, __uint_as_float(dk[8 * j + 1]) * p.scale);
            u.y = pack_half2(__uint_as_float(dk[8 * j + 2]) * p.scale, __uint_as_float(dk[8 * j + 3]) * p.scale);
            u.z = pack_half2(__uint_as_float(dk[8 * j + 4]) * p.scale, __uint_as_float(dk[8 * j + 5]) * p.scale);
            u.w = pack_half2(__uint_as_float(dk[8 * j + 6]) * p.scale, __uint_as_float(dk[8 * j + 7]) * p.scale);
            w.x = pack_half2(__uint_as_float(dv[8 * j]), __uint_as_float(dv[8 * j + 1]));
            w.y = pack_half2(__uint_as_float(dv[8 * j + 2]), __uint_as_float(dv[8 * j + 3]));
            w.z = pack_half2(__uint_as_float(dv[8 * j + 4]), __uint_as_float(dv[8 * j + 5]));
            w.w = pack_half2(__uint_as_float(dv[8 * j + 6]), __uint_as_float(dv[8 * j + 7]));
            reinterpret_cast<uint4*>(gk)[j] = u;
            reinterpret_cast<uint4*>(gv)[j] = w;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem);
}

// delta[h][row] = sum_d dO[row, h*HD + d] * O[row, h*HD + d]   (softmax backward's row term; one warp per row,
// 16-byte loads, HD/8 lanes per head).  Read once here instead of once per key chunk by a single thread per row.
template <int HD>
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __half* o, int64_t ldo, const __half* dout, int64_t lddo, float* delta, int64_t rows, int nheads,
                  float* dq_acc, int64_t lddq) {
  griddep_launch();  // dependents (GEMMs) may start their prologue under this kernel's tail
  griddep_wait();    // launched with the PDL attribute: wait before touching global data
  constexpr int LPH = HD / 8;  // lanes per head
  const int lane = threadIdx.x & 31;
  const int C = nheads * HD;
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
    for (int cb = 0; cb < C; cb += 256) {  // warp-uniform trip count (the shuffles below need every lane)
      const int c0 = cb + lane * 8;
      const bool in = c0 < C;
      const uint4 a = in ? *reinterpret_cast<const uint4*>(o + r * ldo + c0) : make_uint4(0u, 0u, 0u, 0u);
      const uint4 b = in ? *reinterpret_cast<const uint4*>(dout + r * lddo + c0) : make_uint4(0u, 0u, 0u, 0u);
      const __half2* ha = reinterpret_cast<const __half2*>(&a);
      const __half2* hb = reinterpret_cast<const __half2*>(&b);
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 fa = __half22float2(ha[q]), fb = __half22float2(hb[q]);
        acc += fa.x * fb.x + fa.y * fb.y;
      }
#pragma unroll
      for (int o2 = 1; o2 < LPH; o2 <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2);
      if (in && (lane & (LPH - 1)) == 0) delta[(int64_t)(c0 / HD) * rows + r] = acc;
      if (in) {  // the dQ accumulator of this row segment starts from zero (the main kernel reduces into it)
        float4* z = reinterpret_cast<float4*>(dq_acc + r * lddq + c0);
        z[0] = make_float4(0.f, 0.f, 0.f, 0.f), z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

// dtable[rel_index[i][j]][h] += sum_p ds[p][h][i][j]   (i, j < L).  One warp per (h, i) row and slab of problems.
__global__ void __launch_bounds__(256)
relpos_bias_grad_kernel(const __half* ds, int nprob, int nheads, int NP, int L, const int32_t* rel_index,
                        float* dtable, int probs_per_block) {
  griddep_launch();  // dependents (GEMMs) may start their prologue under this kernel's tail
  griddep_wait();    // launched with the PDL attribute: wait before touching global data
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);  // (h, i)
  if (row >= nheads * L) return;
  const int h = row / L, i = row - h * L;
  const int p0 = blockIdx.y * probs_per_block, p1 = min(nprob, p0 + probs_per_block);
  for (int jb = lane * 8; jb < NP; jb += 256) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const size_t pstride = (size_t)nheads * NP * NP;
    const __half* src = ds + (((size_t)p0 * nheads + h) * NP + i) * NP + jb;
    int pr = p0;
    // 8 problems per trip, all eight 16-byte loads issued before the first add: the rows of one (h, i) are a whole
    // [nheads, NP, NP] slab apart, so the memory-level parallelism has to come from here (r1: one load in flight per lane)
    for (; pr + 8 <= p1; pr += 8, src += 8 * pstride) {
      uint4 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] = __ldcs(reinterpret_cast<const uint4*>(src + k * pstride));
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const __half2* hh = reinterpret_cast<const __half2*>(&u[k]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __half22float2(hh[q]);
          acc[2 * q] += f.x, acc[2 * q + 1] += f.y;
        }
      }
    }
    for (; pr < p1; ++pr, src += pstride) {
      const uint4 u = __ldcs(reinterpret_cast<const uint4*>(src));
      const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __half22float2(hh[q]);
        acc[2 * q] += f.x, acc[2 * q + 1] += f.y;
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (jb + q < L) atomicAdd(dtable + (size_t)rel_index[i * L + jb + q] * nheads + h, acc[q]);
  }
}

template <int HD, bool BMMA>
static int launch_attn_bwd(const void* qkv, int64_t ld, const AttnBwdParams& p, int nkc, cudaStream_t s) {
  using Cfg = AttnBwdCfg<HD, BMMA>;
  CUtensorMap tq, tdo, tb;
  const CUtensorMapSwizzle sw = HD == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  int rc = encode_tmap_2d_f16(&tq, qkv, p.rows_total, ld, ld, 128, HD, sw);
  if (rc) return rc;
  rc = encode_tmap_2d_f16(&tdo, p.dout, p.rows_total, p.nheads * HD, p.lddo, 128, HD, sw);
  if (rc) return rc;
  tb = tq;
  if (BMMA) {
    rc = encode_tmap_2d_f16(&tb, p.bias16, (uint64_t)8 * p.nheads * p.NPb, p.NPb, p.NPb, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  auto kern = attn_bwd_kernel<HD, BMMA>;
  static bool attr_set = false;
  if (!attr_set) {
    LAV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int nitems = nkc * p.nheads * p.nprob;
  dim3 grid(std::min(nitems, sm_count()));   // persistent: one CTA per SM walks the (key chunk, head, problem) items
  LAV_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(kAttnBwdThreads), Cfg::SMEM_BYTES, s, tq, tdo, tb, p));
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}

}  // namespace lav

using namespace lav;

extern "C" int lav_attn_bwd_f16(const void* qkv, int64_t ld, int64_t rows_total, int q_off, int k_off, int v_off,
                                int head_dim, int nheads, int nprob, int L, float scale, const void* bias16, int NPb,
                                const int32_t* prob_class, int class_period, const float* key_bias, int NPk,
                                int causal_from, const void* out16, int64_t ldo, const void* dout16, int64_t lddo, const float* lse,
                                float* delta_ws, float* dq_acc, int64_t lddq, void* dqkv16, int64_t lddqkv, void* ds16, int NPs,
                                const LavDropout* drop, void* stream) {
  LAV_REQUIRE(qkv && out16 && dout16 && lse && delta_ws && dq_acc && dqkv16, "lav_attn_bwd_f16: null pointer");
  LAV_REQUIRE(nprob > 0 && nheads > 0 && L > 0, "lav_attn_bwd_f16: empty problem");
  LAV_REQUIRE((ldo % 8) == 0 && (lddo % 8) == 0 && (lddqkv % 8) == 0 && (q_off % 8) == 0 && (k_off % 8) == 0 &&
                  (v_off % 8) == 0, "lav_attn_bwd_f16: offsets / ld must be multiples of 8");
  LAV_REQUIRE(head_dim == 32 || head_dim == 64, "lav_attn_bwd_f16: head_dim must be 32 or 64");
  const int nkc = (L + 127) / 128;
  LAV_REQUIRE(!bias16 || NPb >= nkc * 128, "lav_attn_bwd_f16: dense bias too small");
  LAV_REQUIRE(!key_bias || NPk >= nkc * 128, "lav_attn_bwd_f16: key_bias too small");
  LAV_REQUIRE(!ds16 || NPs >= nkc * 128, "lav_attn_bwd_f16: ds workspace too small");
  AttnBwdParams p;
  p.L = L, p.nheads = nheads, p.nprob = nprob, p.q_off = q_off, p.k_off = k_off, p.v_off = v_off, p.scale = scale;
  p.bias16 = (const __half*)bias16, p.NPb = NPb, p.prob_class = prob_class, p.period = class_period > 0 ? class_period : 1;
  p.key_bias = key_bias, p.NPk = NPk, p.out = (const __half*)out16, p.ldo = ldo, p.dout = (const __half*)dout16;
  p.lddo = lddo, p.lse = lse, p.rows_total = rows_total, p.dq_acc = dq_acc, p.lddq = lddq;
  p.dqkv = (__half*)dqkv16, p.lddqkv = lddqkv, p.ds_out = (__half*)ds16, p.NPs = NPs;
  p.drop = make_drop(drop);
  p.causal_from = causal_from;
  cudaStream_t s = (cudaStream_t)stream;
  p.delta = delta_ws;
  LAV_REQUIRE((lddq % 4) == 0 && ((uintptr_t)dq_acc % 16) == 0, "lav_attn_bwd_f16: dq_acc rows must be 16-byte aligned");
  {
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((rows_total + 7) / 8, (int64_t)sm_count() * 8));
    if (head_dim == 32)
      LAV_CHECK_CUDA(launch_pdl(attn_delta_kernel<32>, dim3(grid), dim3(256), 0, s, (const __half*)out16, ldo, (const __half*)dout16, lddo, delta_ws, rows_total, nheads,
                                                 dq_acc, lddq));
    else
      LAV_CHECK_CUDA(launch_pdl(attn_delta_kernel<64>, dim3(grid), dim3(256), 0, s, (const __half*)out16, ldo, (const __half*)dout16, lddo, delta_ws, rows_total, nheads,
                                                 dq_acc, lddq));
    LAV_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  if (head_dim == 32 && bias16) return launch_attn_bwd<32, true>(qkv, ld, p, nkc, s);
  return head_dim == 32 ? launch_attn_bwd<32, false>(qkv, ld, p, nkc, s) : launch_attn_bwd<64, false>(qkv, ld, p, nkc, s);
}

extern "C" int lav_relpos_bias_grad(const void* ds16, int nprob, int nheads, int NP, int L, const int32_t* rel_index,
                                    float* dtable, void* stream) {
  LAV_REQUIRE(ds16 && rel_index && dtable && L <= NP && (NP % 8) == 0, "lav_relpos_bias_grad: bad arguments");
  const int rows = nheads * L;
  const int xb = (rows + 7) / 8;
  int slabs = std::max(1, std::min(nprob, (4 * sm_count() + xb - 1) / xb));
  const int ppb = (nprob + slabs - 1) / slabs;
  dim3 grid(xb, (nprob + ppb - 1) / ppb);
  LAV_CHECK_CUDA(launch_pdl(relpos_bias_grad_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const __half*)ds16, nprob, nheads, NP, L, rel_index,
                                                                 dtable, ppb));
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}
