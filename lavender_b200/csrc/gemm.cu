// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//   D[M,N] = epilogue(alpha * A·Bᵀ),  fp16 operands staged by TMA (128B swizzle), fp32 accumulators in TMEM.
//   warp 0: TMA producer | warp 1: MMA issuer | warp 2: TMEM allocator | warps 4-11: two epilogue warpgroups,
//   one per TMEM accumulator stage, so the epilogue of tile i overlaps the MMAs of tile i+1.
// Both operands may be K-major or MN-major (UMMA descriptor bit), which gives forward (K,K), dgrad (K,MN)
// and wgrad (MN,MN; split-K with fp32 atomics) from one kernel without transposed copies.
#include <algorithm>
#include <cstdlib>

#include "rng.cuh"
#include "runtime.h"
#include "sm100.cuh"

namespace lav {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 fp16 = one 128-byte swizzle atom
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 128 + kEpiWarps * 32;
// Per-warp epilogue staging: ONE 32 x 32 fp32 chunk (4 KB), 16-byte slots XOR-swizzled by the row so that both the row
// writes (thread = row) and the transposed reads (8 lanes per row) are conflict-free without padding; the TMA-store path
// uses it as two 2 KB fp16 buffers / one 4 KB fp32 buffer.  Kept small on purpose: with 64 KB of staging the 128 x 256
// tile had only 3 operand stages (3 x 48 KB in flight against ~0.8 us of TMA latency + 0.27 us of MMA per k-block made
// the main loop feed-bound at 0.37 us per k-block, profiles/r2_gemm_ablation.md); 32 KB buys the 4th stage.
constexpr int kEpiWarpBytes = 4096;
__device__ __forceinline__ int epi_slot(int row, int slot) { return row * 32 + ((slot ^ (row & 7)) << 2); }  // float index

// Persistent work streams: one CTA per SM, or one CTA pair per TPC.  The pair path (LAV_GEMM_PAIR=1) is correct and
// wins on large square problems (8192^3: 1342 vs 1150 TFLOP/s) but not on the hot path's shapes, whose cost is the
// epilogue and the per-launch prologue rather than the operand feed (profiles/r1d_gemm_sweep.md), so it is opt-in.
static int work_streams(int ncta) { return ncta == 2 ? std::max(1, sm_count() / 2) : sm_count(); }
static bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LAV_GEMM_PAIR");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// NCTA = 2: a CTA pair (cluster of two on one TPC) owns a 256 x BN tile and issues cta_group::2 MMAs; each CTA stages
// its own 128 rows of A and its own BN/2 rows of B, so the L2 -> smem traffic and the smem footprint per FLOP halve
// for B and the ring is deep enough (5 x 32 KB at BN = 256) to cover the TMA latency.
template <int BN, int NCTA>
struct GemmCfg {
  static constexpr int BNL = BN / NCTA;  // rows of B staged by one CTA
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BNL * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = kEpiWarps * kEpiWarpBytes;  // per-warp staging: transposes / TMA-store buffers
  static constexpr int BAR_BYTES = 1024;  // mbarriers + TMEM slot in [0, 256), the all-ones operand tile in [256, 768)
  static constexpr int STAGES_MAX = (232448 - 1024 - BAR_BYTES - EPI_BYTES) / STAGE_BYTES;
  static constexpr int STAGES_CAP = 8;
  static constexpr int STAGES = STAGES_MAX > STAGES_CAP ? STAGES_CAP : STAGES_MAX;  // 1-CTA: 4 / 4 / 6 / 8 for BN = 256 / 192 / 128 / 64
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment
  static constexpr int BIAS_COLS = (BN <= 192 && NCTA == 1) ? 32 : 0;         // 2 x 16 columns: fused bias gradient
  static constexpr int TMEM_COLS = 2 * BN + BIAS_COLS;                        // 2 accumulator stages
  static constexpr int TMEM_ALLOC = TMEM_COLS <= 128 ? 128 : TMEM_COLS <= 256 ? 256 : 512;  // power of 2
};

struct GemmParams {
  int M, N, K;
  int m_blocks, n_blocks, k_blocks, splits, kb_per_split;  // m_blocks counts tiles of BM * NCTA rows
  DropParams drop;  // resolved from epi.drop
  float* bias_grad; // wgrad only: bias_grad[m] += alpha * sum_k A(m, k), by one extra N = 16 MMA against an all-ones tile
  int tma_store;    // 1: plain fp16 / fp32 store through TMA (no row map / residual / accumulation)
  int debug;  // LAV_GEMM_DEBUG (profiling only): bit 0 = epilogue drains without math / stores, bit 1 = no MMA issue
  LavGemmEpilogue epi;
};

struct TileCoord {
  int m_blk, n_blk, kb0, kb1;
};
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int item) {
  TileCoord t;
  t.n_blk = item % p.n_blocks;
  int r = item / p.n_blocks;
  t.m_blk = r % p.m_blocks;
  int split = r / p.m_blocks;
  t.kb0 = split * p.kb_per_split;
  t.kb1 = min(p.k_blocks, t.kb0 + p.kb_per_split);
  return t;
}

// ---------------------------------------------------------------------------------------------------------
// Epilogue.  tcgen05.ld gives each thread one accumulator ROW (32 consecutive columns of a 32x32 chunk).  The
// per-element math (alpha, bias, GELU, DropPath row scale) runs in that layout; every global access goes through
// a per-warp shared-memory transpose so that 8 lanes cover the 32 columns of one row and a warp instruction
// touches 4 rows x 128 contiguous bytes (fp32) / 64 bytes (fp16) instead of 32 different 128-byte lines.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ void stage_rows(float* stg, int lane, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stg + epi_slot(lane, j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

enum { ST_F32 = 0, ST_F32_RMW = 1, ST_F32_RED = 2, ST_F16 = 3 };

// Every lane of the warp has finished reading its rows of the accumulator stage (tcgen05.wait::ld + fence::before_thread_sync
// executed by the caller): one lane tells the MMA warp — in a CTA pair that is a remote arrive on the leader's barrier, so
// 8 per CTA and tile instead of 256.
template <int NCTA>
__device__ __forceinline__ void release_tmem_stage(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) {
    if (NCTA == 2) mbar_arrive_cluster(bar, 0);
    else mbar_arrive(bar);
  }
}

// staged chunk -> out (+ prefetched residual, at the pre-resolved output rows).  (rr, cg) = lane's row-in-group /
// 4-column group; orow[it] < 0 marks rows beyond M.
template <int MODE>
__device__ __forceinline__ void store_phase(const GemmParams& p, const float* stg, const int (&orow)[8],
                                            const uint4 (&res)[8], bool use_res, int col, int rr, int cg) {
  const LavGemmEpilogue& e = p.epi;
  const int nv = p.N - col;
  if (nv <= 0) return;
  const bool vec_o = nv >= 4 && (e.ldo & 3) == 0;
  if (MODE == ST_F32_RMW && vec_o) {
    // read-modify-write of the gradient rows: all 8 loads first, then the adds and stores.  Written as load / add / store per
    // row the stores fence the next load (possible aliasing), i.e. 8 serial HBM round trips per chunk: the decoder weight
    // gradient (30522 x 768, K = 160: 956 tiles of pure epilogue) took 129 us for 188 MB of traffic.
    float4 x[8];
#pragma unroll
    for (int it = 0; it < 8; ++it)
      if (orow[it] >= 0) x[it] = *reinterpret_cast<const float4*>(reinterpret_cast<float*>(e.out) + (size_t)orow[it] * e.ldo + col);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      if (orow[it] < 0) continue;
      const float4 a = *reinterpret_cast<const float4*>(stg + epi_slot(it * 4 + rr, cg));
      float4 y = make_float4(x[it].x + a.x, x[it].y + a.y, x[it].z + a.z, x[it].w + a.w);
      if (use_res) {
        y.x += __uint_as_float(res[it].x), y.y += __uint_as_float(res[it].y);
        y.z += __uint_as_float(res[it].z), y.w += __uint_as_float(res[it].w);
      }
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + (size_t)orow[it] * e.ldo + col) = y;
    }
    return;
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + rr;
    if (orow[it] < 0) continue;
    float4 a = *reinterpret_cast<const float4*>(stg + epi_slot(r, cg));
    if (use_res) {
      a.x += __uint_as_float(res[it].x), a.y += __uint_as_float(res[it].y);
      a.z += __uint_as_float(res[it].z), a.w += __uint_as_float(res[it].w);
    }
    const float v[4] = {a.x, a.y, a.z, a.w};
    if (MODE == ST_F16) {
      __half* o = reinterpret_cast<__half*>(e.out) + (size_t)orow[it] * e.ldo + col;
      if (vec_o) {
        *reinterpret_cast<uint2*>(o) = make_uint2(pack_half2(a.x, a.y), pack_half2(a.z, a.w));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < nv) o[j] = __float2half_rn(v[j]);
      }
    } else {
      float* o = reinterpret_cast<float*>(e.out) + (size_t)orow[it] * e.ldo + col;
      if (MODE == ST_F32_RED) {
        if (vec_o) red_add_v4(o, a);
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < nv) atomicAdd(o + j, v[j]);
        }
      } else if (MODE == ST_F32_RMW) {
        if (vec_o) {
          const float4 x = *reinterpret_cast<float4*>(o);
          *reinterpret_cast<float4*>(o) = make_float4(x.x + a.x, x.y + a.y, x.z + a.z, x.w + a.w);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < nv) o[j] += v[j];
        }
      } else {
        if (vec_o) *reinterpret_cast<float4*>(o) = a;
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < nv) o[j] = v[j];
        }
      }
    }
  }
}

// staged chunk -> aux (fp16 [row][col], identity rows): gelu'(pre-activation) for the backward
__device__ __forceinline__ void store_aux_phase(const GemmParams& p, const float* stg, int row_base, int col, int rr, int cg) {
  const LavGemmEpilogue& e = p.epi;
  const int nv = p.N - col;
  if (nv <= 0) return;
  const bool vec = nv >= 4 && (e.ldaux & 3) == 0;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + rr;
    const int row = row_base + r;
    if (row >= p.M) continue;
    const float4 a = *reinterpret_cast<const float4*>(stg + epi_slot(r, cg));
    __half* o = reinterpret_cast<__half*>(e.aux) + (size_t)row * e.ldaux + col;
    if (vec) *reinterpret_cast<uint2*>(o) = make_uint2(pack_half2(a.x, a.y), pack_half2(a.z, a.w));
    else {
      const float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nv) o[j] = __float2half_rn(v[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// TMA-store epilogue for the plain outputs (fp16 activations of qkv / fc1 / dgrad GEMMs, fp32 logits): the thread
// that owns accumulator row r packs 32 columns into its row of a swizzled shared-memory tile and ONE lane hands the
// 32 x 32 tile to the TMA unit.  No transposed read-back, no per-thread global addresses or bounds predicates (the
// tensor map clips rows >= M and columns >= N); up to 3 stores per warp stay in flight.
// ---------------------------------------------------------------------------------------------------------
template <bool F32>
__device__ __forceinline__ void stage_and_store(const CUtensorMap* tm, uint8_t* buf, uint32_t& nb, int lane,
                                                const float (&v)[32], int col0, int row0, bool rows_valid, int dbg = 0) {
  constexpr int ROWB = F32 ? 128 : 64;
  constexpr int BUFB = 32 * ROWB;
  constexpr int NB = kEpiWarpBytes / BUFB;
  uint8_t* b = buf + (nb % NB) * BUFB;
  if (dbg & 16) return;                          // (profiling) no staging, no store
  if (lane == 0) tma_store_wait_read<NB - 1>();  // the store that last used this buffer has drained it
  __syncwarp();
  if (F32) {
    uint8_t* row = b + lane * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(row + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
    uint8_t* row = b + lane * 64;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<uint4*>(row + ((j ^ ((lane >> 1) & 3)) << 4)) =
          make_uint4(pack_half2(v[8 * j], v[8 * j + 1]), pack_half2(v[8 * j + 2], v[8 * j + 3]),
                     pack_half2(v[8 * j + 4], v[8 * j + 5]), pack_half2(v[8 * j + 6], v[8 * j + 7]));
  }
  if (!(dbg & 8)) fence_proxy_async_smem();      // (profiling bit 8: no proxy fence)
  __syncwarp();
  if (lane == 0) {
    if (rows_valid && !(dbg & 4)) tma_store_2d(tm, b, col0, row0);   // (profiling bit 4: no TMA store)
    tma_store_commit();
  }
  ++nb;
}

// ---------------------------------------------------------------------------------------------------------
// The two epilogues (TMA-store path for plain outputs, register path for scatter / residual / accumulation), both
// software-pipelined.  profiles/r1d_gemm_sweep.md: for K <= 1024 the kernels are bound by the accumulator drain, and in
// round 1 a warp spent ~2000 clk per 32 x 32 chunk in a serial chain  tcgen05.ld -> wait -> bias (8 broadcast LDG at L2
// latency: the L1 is carved down to nothing by the 227 KB of shared memory) -> math -> stores.  Here the tcgen05.ld (and
// the GELU' input) of the warp's NEXT chunk is issued before the math of the current one, so its latency hides under
// math + stores; the TMEM stage is released one chunk earlier; the bias arrives as ONE coalesced load per lane per
// chunk, issued before the accumulator is even ready, and is broadcast with shuffles.  All 384 threads run at 168
// registers (a setmaxnreg re-partition towards the epilogue warpgroups made ptxas spill kilobytes: not used).
// ---------------------------------------------------------------------------------------------------------
template <int BN, int NCTA, int ACT, bool F32>
__device__ __forceinline__ void epilogue_warps_tma2(const GemmParams& p, const CUtensorMap* tmOut, const CUtensorMap* tmAux,
                                                    float* epi_stage, uint64_t* tmem_full, uint64_t* tmem_empty,
                                                    uint32_t tmem_base, int warp, int lane, int total, int stream_id,
                                                    int nstreams, int rank) {
  const int wg = (warp - 4) >> 2;
  const int q = warp & 3;
  uint8_t* buf = reinterpret_cast<uint8_t*>(epi_stage) + (warp - 4) * kEpiWarpBytes;
  const LavGemmEpilogue& e = p.epi;
  const bool aux_out = ACT == LAV_ACT_GELU && e.aux != nullptr;
  constexpr int KMAX = (BN / 32 + 1) / 2;  // chunks one warpgroup drains per tile
  uint32_t nb = 0;
  int iter = 0;
  for (int item = stream_id; item < total; item += nstreams, ++iter) {
    const int as = iter & 1;
    const TileCoord t = decode_tile(p, item);
    const int row_base = (t.m_blk * NCTA + rank) * BM + q * 32;
    const bool rows_valid = row_base < p.M;
    const int row = row_base + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
    const int n0 = t.n_blk * BN;
    const int nchunks = min(BN / 32, (p.N - n0 + 31) / 32);
    // bias of this warp's chunks: lane j holds column (chunk, j); requested while the MMAs of the tile still run
    float bl[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int col = n0 + (wg + 2 * k) * 32 + lane;
      bl[k] = (e.bias && wg + 2 * k < nchunks && col < p.N) ? __ldg(e.bias + col) : 0.f;
    }
    // (r2: a coalesced load of the GELU' input + transpose through the staging buffer was measured SLOWER than this
    // thread = row load - 51 vs 48 us on 7840 x 2048 x 512 - because the transpose has to wait for the staging buffer's
    // previous TMA store; see profiles/r2_gemm_ablation.md)
    auto load_aux = [&](uint4(&ax)[4], int c) {
      const int col0 = n0 + c * 32;
      const __half* xa = reinterpret_cast<const __half*>(e.aux) + (size_t)min(row, p.M - 1) * e.ldaux + col0;
      if (col0 + 32 <= p.N && (e.ldaux & 7) == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ax[j] = __ldg(reinterpret_cast<const uint4*>(xa) + j);
      } else {
        __half* h = reinterpret_cast<__half*>(ax);
#pragma unroll
        for (int j = 0; j < 32; ++j) h[j] = col0 + j < p.N ? xa[j] : __float2half_rn(0.f);
      }
    };
    mbar_wait(tmem_full + as, (iter >> 1) & 1, 4);
    tc_fence_after();
    if ((p.debug & 1) || wg >= nchunks) {  // (a warpgroup without a chunk of a narrow tile just releases the stage)
      tc_fence_before();
      release_tmem_stage<NCTA>(tmem_empty + as, lane);
      continue;
    }
    uint4 axA[4], axB[4];
    if (ACT == LAV_ACT_GELU_BWD) load_aux(axA, wg);
    // math + stores of one chunk whose accumulator values (cur) and GELU' input (axc) are in registers
    auto body = [&](uint32_t(&cur)[32], uint4(&axc)[4], int c, float bias_lane) {
      const int col0 = n0 + c * 32;
      float v[32];
      if (e.alpha != 1.0f) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(cur[j]) * e.alpha;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(cur[j]);
      }
      if (e.bias && !(p.debug & 32)) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bias_lane, j);
      }
      if (ACT == LAV_ACT_GELU) {
        // aux receives gelu'(pre-activation), not the pre-activation: the cdf / pdf are in registers here anyway (one FMA
        // more per output), the backward epilogue becomes a multiply, and the saved tensor has the same size
        if (aux_out) {
          float g[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) gelu_erf_both(v[j], v[j], g[j]);
          stage_and_store<false>(tmAux, buf, nb, lane, g, col0, row_base, rows_valid, p.debug);
        } else if (!(p.debug & 128)) {  // (profiling bit 128: no activation math)
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        }
      } else if (ACT == LAV_ACT_GELU_BWD) {
        const __half2* h2 = reinterpret_cast<const __half2*>(axc);   // gelu'(pre-activation), saved by the forward epilogue
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 x = __half22float2(h2[j]);
          v[2 * j] *= x.x, v[2 * j + 1] *= x.y;
        }
      }
      stage_and_store<F32>(tmOut, buf, nb, lane, v, col0, row_base, rows_valid, p.debug);
    };
    if constexpr (ACT == LAV_ACT_NONE) {
      // plain epilogue: the next chunk's accumulator is in flight under this chunk's conversion + stores
      uint32_t accA[32], accB[32];
      tmem_ld_32x32(taddr + wg * 32, accA);
      auto step = [&](uint32_t(&cur)[32], uint32_t(&nxt)[32], int c, float bias_lane) {
        tmem_ld_wait();
        if (c + 2 < nchunks) {
          tmem_ld_32x32(taddr + (c + 2) * 32, nxt);
        } else {                // the whole accumulator has been read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          release_tmem_stage<NCTA>(tmem_empty + as, lane);
        }
        body(cur, axA, c, bias_lane);
      };
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int c = wg + 2 * k;
        if (c >= nchunks) break;
        if (k & 1) step(accB, accA, c, bl[k]);
        else step(accA, accB, c, bl[k]);
      }
    } else {
      // GELU / GELU': a chunk is ~20 instructions per output, the tcgen05.ld round trip is noise next to it, and the
      // second accumulator buffer's 32 registers are what ptxas needs to keep several GELU evaluations in flight
      // (with both buffers live it evaluated the 32 outputs one after the other through the same three registers)
      uint32_t acc[32];
      auto step = [&](uint4(&axc)[4], uint4(&axn)[4], int c, float bias_lane) {
        tmem_ld_32x32(taddr + c * 32, acc);
        if (ACT == LAV_ACT_GELU_BWD && c + 2 < nchunks) load_aux(axn, c + 2);   // GELU' input of the next chunk (HBM latency)
        tmem_ld_wait();
        if (c + 2 >= nchunks) {
          tc_fence_before();
          release_tmem_stage<NCTA>(tmem_empty + as, lane);
        }
        body(acc, axc, c, bias_lane);
      };
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int c = wg + 2 * k;
        if (c >= nchunks) break;
        if (k & 1) step(axB, axA, c, bl[k]);
        else step(axA, axB, c, bl[k]);
      }
    }
  }
  if (lane == 0) tma_store_wait_all();  // smem must outlive the reads; the writes complete before the grid does
}

// register-path epilogue (row map / residual / DropPath / dropout / accumulation): one accumulator chunk in registers,
// the chunk's global input (residual rows / GELU pre-activation) prefetched one chunk ahead
// Body of the 8 epilogue warps (two warpgroups, one per TMEM accumulator stage), specialised at compile time on the
// activation, on whether a global input is prefetched (PRE: residual rows) and on the store mode: the epilogue is
// instruction-issue bound for the short-K GEMMs of the Swin stages, so unused paths must not cost instructions.
template <int BN, int NCTA, int ACT, int PRE, int STORE, bool DROP = false>
__device__ __forceinline__ void epilogue_warps(const GemmParams& p, float* epi_stage, uint64_t* tmem_full,
                                               uint64_t* tmem_empty, uint32_t tmem_base, int warp, int lane, int total,
                                               int stream_id, int nstreams, int rank) {
  const int wg = (warp - 4) >> 2;
  const int q = warp & 3;  // TMEM lane quarter this warp may access
  float* stg = epi_stage + (warp - 4) * (kEpiWarpBytes / 4);
  const int rr = lane >> 3, cg = lane & 7;  // this lane's row-within-group / 4-column group in the store phase
  const LavGemmEpilogue& e = p.epi;
  constexpr bool use_res = PRE == 1;
  constexpr bool use_auxin = ACT == LAV_ACT_GELU_BWD;
  constexpr bool use_pre = use_res || use_auxin;
  DropKey dkey{};
  if (DROP) dkey = drop_key(p.drop);
  int iter = 0;
  for (int item = stream_id; item < total; item += nstreams, ++iter) {
    const int as = iter & 1;  // accumulator stage; BOTH warpgroups drain every tile (even / odd 32-column chunks)
    const TileCoord t = decode_tile(p, item);
    const int row_base = (t.m_blk * NCTA + rank) * BM + q * 32;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
    const int nchunks = min(BN / 32, (p.N - t.n_blk * BN + 31) / 32);
    // output rows of this lane's 8 store-phase rows (row map resolved once per tile)
    int orow[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = row_base + it * 4 + rr;
      orow[it] = row < p.M ? (e.row_map ? __ldg(e.row_map + row) : row) : -1;
    }
    // Prefetch of the chunk's global INPUT (residual rows, or the GELU pre-activation for GELU_BWD) into registers,
    // one chunk ahead: chunk 0 is requested before the accumulator is ready, so the latency hides under the MMAs.
    auto prefetch = [&](uint4(&dst)[8], int c) {
      const int col = t.n_blk * BN + c * 32 + 4 * cg;
      const int nv = p.N - col;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        dst[it] = make_uint4(0u, 0u, 0u, 0u);
        if (orow[it] < 0 || nv <= 0) continue;
        if (use_res) {
          const float* rs = e.residual + (size_t)orow[it] * e.ldres + col;
          if (nv >= 4 && (e.ldres & 3) == 0) dst[it] = *reinterpret_cast<const uint4*>(rs);
          else {
            dst[it].x = __float_as_uint(rs[0]);
            if (nv > 1) dst[it].y = __float_as_uint(rs[1]);
            if (nv > 2) dst[it].z = __float_as_uint(rs[2]);
            if (nv > 3) dst[it].w = __float_as_uint(rs[3]);
          }
        } else if (use_auxin) {
          const __half* x = reinterpret_cast<const __half*>(e.aux) + (size_t)(row_base + it * 4 + rr) * e.ldaux + col;
          if (nv >= 4 && (e.ldaux & 3) == 0) {
            const uint2 u = *reinterpret_cast<const uint2*>(x);
            dst[it].x = u.x, dst[it].y = u.y;
          } else {
            const __half z = __float2half_rn(0.f);
            __half2 h0 = __halves2half2(x[0], nv > 1 ? x[1] : z), h1 = __halves2half2(nv > 2 ? x[2] : z, nv > 3 ? x[3] : z);
            dst[it].x = *reinterpret_cast<uint32_t*>(&h0), dst[it].y = *reinterpret_cast<uint32_t*>(&h1);
          }
        }
      }
    };
    uint4 cur[8], nxt[8];
    if (use_pre && wg < nchunks) prefetch(cur, wg);
    mbar_wait(tmem_full + as, (iter >> 1) & 1, 4);
    tc_fence_after();
    if (GemmCfg<BN, NCTA>::BIAS_COLS > 0 && p.bias_grad != nullptr && t.n_blk == 0 && wg == 0) {
      // column 0 of the 128 x 16 side accumulator = row sums of A over this tile's k range
      uint32_t bsum;
      tmem_ld_32x1(tmem_base + ((uint32_t)(q * 32) << 16) + 2 * BN + as * 16, bsum);
      tmem_ld_wait();
      const int brow = row_base + lane;
      if (brow < p.M) atomicAdd(p.bias_grad + brow, __uint_as_float(bsum) * e.alpha);
    }
    if ((p.debug & 1) || wg >= nchunks) {  // (a warpgroup without a chunk of a narrow tile just releases the stage)
      tc_fence_before();
      release_tmem_stage<NCTA>(tmem_empty + as, lane);
      continue;
    }
#pragma unroll 1
    for (int c = wg; c < nchunks; c += 2) {
      const int col0 = t.n_blk * BN + c * 32;
      const int col = col0 + 4 * cg;  // this lane's columns in the transposed (store) phases
      uint32_t acc[32];
      tmem_ld_32x32(taddr + c * 32, acc);
      if (use_pre && c + 2 < nchunks) prefetch(nxt, c + 2);
      tmem_ld_wait();
      if (c + 2 >= nchunks) {  // accumulator fully read: hand the TMEM stage back to the MMA warp early
        tc_fence_before();
        release_tmem_stage<NCTA>(tmem_empty + as, lane);  // (pair: the leader's MMA warp waits for both CTAs)
      }
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) * e.alpha;
      if (e.bias) {
        if (col0 + 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + col0) + j);
            v[4 * j] += b.x, v[4 * j + 1] += b.y, v[4 * j + 2] += b.z, v[4 * j + 3] += b.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) v[j] += __ldg(e.bias + col0 + j);
        }
      }
      if (ACT == LAV_ACT_GELU) {
        if (e.aux) {   // aux = gelu'(pre-activation) (see the TMA-store epilogue)
          float g[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) gelu_erf_both(v[j], v[j], g[j]);
          stage_rows(stg, lane, g);
          __syncwarp();
          store_aux_phase(p, stg, row_base, col, rr, cg);
          __syncwarp();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        }
      } else if (use_auxin) {
        // gelu'(pre-activation): (store-layout registers) -> smem -> (row-layout registers)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&cur[it].x));
          const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&cur[it].y));
          *reinterpret_cast<float4*>(stg + epi_slot(it * 4 + rr, cg)) = make_float4(f0.x, f0.y, f1.x, f1.y);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 x = *reinterpret_cast<const float4*>(stg + epi_slot(lane, j));
          v[4 * j] *= x.x, v[4 * j + 1] *= x.y;
          v[4 * j + 2] *= x.z, v[4 * j + 3] *= x.w;
        }
        __syncwarp();
      }
      if (DROP) {  // BertSelfOutput / BertOutput .dropout on (dense + bias), element index (GEMM row, column)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t m = drop_keep8(dkey, p.drop.thresh, (uint32_t)(row_base + lane), (uint32_t)((col0 >> 3) + j), 0u);
#pragma unroll
          for (int q = 0; q < 8; ++q) v[8 * j + q] = ((m >> q) & 1u) ? v[8 * j + q] * p.drop.inv_keep : 0.f;
        }
      }
      if (e.row_scale) {
        const int row = min(row_base + lane, p.M - 1);
        const float sc = __ldg(e.row_scale + row / e.rows_per_scale);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= sc;
      }
      stage_rows(stg, lane, v);
      __syncwarp();
      store_phase<STORE>(p, stg, orow, cur, use_res, col, rr, cg);
      __syncwarp();
      if (use_pre) {
#pragma unroll
        for (int it = 0; it < 8; ++it) cur[it] = nxt[it];
      }
    }
  }
}


template <int BN, int AMAJ, int BMAJ, int NCTA>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAux,
                const GemmParams p) {
  using Cfg = GemmCfg<BN, NCTA>;
  constexpr int BNL = Cfg::BNL;
  extern __shared__ uint8_t smem_raw[];
  // (the dynamic smem window starts at the same offset in both CTAs of a pair, so the aligned offsets match too)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_stage = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* ones_tile = reinterpret_cast<uint8_t*>(bars) + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = NCTA == 2 ? (int)cluster_ctarank() : 0;          // 0 = leader (issues the MMAs)
  const int stream_id = NCTA == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;  // persistent work stream of this CTA (pair)
  const int nstreams = NCTA == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tmem_full + s, 1);
      mbar_init(tmem_empty + s, kEpiWarps * NCTA);  // one elected arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 3 && p.bias_grad != nullptr) {  // 16 x 16 tile of fp16 ones (0x3C00), visible to the async proxy
    reinterpret_cast<uint4*>(ones_tile)[lane] = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_proxy_async_smem();
  }
  if (warp == 2) {
    if (NCTA == 2) tmem_alloc_pair<Cfg::TMEM_ALLOC>(tmem_slot);
    else tmem_alloc<Cfg::TMEM_ALLOC>(tmem_slot);
  }
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all();  // the peer's barriers must be initialised before any remote arrive / TMA signal
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) touches no global data and
  // may overlap the tail of the previous kernel; from here on operands / outputs of earlier kernels are accessed.
  griddep_launch();
  griddep_wait();

  const int total = p.m_blocks * p.n_blocks * p.splits;

  if (warp < 4) {
  if (warp == 0) {
    // ------------------------------------------------ TMA producer (one lane; in a pair, both CTAs run one)
    if (lane == 0 && !(p.debug & 64)) {   // (profiling bit 64: no operand loads - the MMA warp runs on stale smem)
      int stage = 0;
      uint32_t phase = 0;
      for (int item = stream_id; item < total; item += nstreams) {
        const TileCoord t = decode_tile(p, item);
        const int m0 = (t.m_blk * NCTA + rank) * BM;
        const int n0 = t.n_blk * BN + rank * BNL;
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(empty + stage, phase ^ 1, 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          // the bytes of BOTH CTAs complete on the leader's barrier (the leader's MMA warp is the only consumer)
          if (rank == 0) mbar_arrive_expect_tx(full + stage, Cfg::STAGE_BYTES * NCTA);
          auto load = [&](void* dst, const CUtensorMap* m, int c0, int c1) {
            if (NCTA == 2) tma_load_2d_pair(dst, m, full + stage, c0, c1);
            else tma_load_2d(dst, m, full + stage, c0, c1);
          };
          if (AMAJ == LAV_MAJOR_K) {
            load(sa, &tmA, kb * BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) load(sa + j * (BK * 128), &tmA, m0 + j * 64, kb * BK);
          }
          if (BMAJ == LAV_MAJOR_K) {
            load(sb, &tmB, kb * BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BNL / 64; ++j) load(sb + j * (BK * 128), &tmB, n0 + j * 64, kb * BK);
          }
          if (++stage == Cfg::STAGES) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (one lane of the leader CTA)
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM * NCTA, BN, AMAJ, BMAJ);
      // K-major: 8-row atoms are 1024 B apart (SBO); MN-major: 64-element atoms are BK*128 B apart (LBO),
      // 8-k-row groups 1024 B apart (SBO).
      constexpr uint32_t a_lbo = (AMAJ == LAV_MAJOR_K) ? 0 : BK * 128;
      constexpr uint32_t b_lbo = (BMAJ == LAV_MAJOR_K) ? 0 : BK * 128;
      constexpr uint32_t a_kstep = (AMAJ == LAV_MAJOR_K) ? 32 : 16 * 128;  // bytes per UMMA_K = 16
      constexpr uint32_t b_kstep = (BMAJ == LAV_MAJOR_K) ? 32 : 16 * 128;
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int item = stream_id; item < total; item += nstreams, ++iter) {
        const TileCoord t = decode_tile(p, item);
        const int as = iter & 1;
        mbar_wait(tmem_empty + as, ((iter >> 1) & 1) ^ 1, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          if (!(p.debug & 64)) mbar_wait(full + stage, phase, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if (p.debug & 2) break;
            const uint64_t ad = make_smem_desc(sa + k * a_kstep, a_lbo, 1024, SWZ_128B);
            const uint64_t bd = make_smem_desc(sb + k * b_kstep, b_lbo, 1024, SWZ_128B);
            if (NCTA == 2) umma_f16_ss_pair(d_tmem, ad, bd, idesc, (kb > t.kb0 || k > 0) ? 1u : 0u);
            else umma_f16_ss(d_tmem, ad, bd, idesc, (kb > t.kb0 || k > 0) ? 1u : 0u);
          }
          if (Cfg::BIAS_COLS > 0 && p.bias_grad != nullptr && t.n_blk == 0) {
            // bias gradient: side accumulator [128 x 16] += A_tile * ones (K-major, no swizzle; every element is 1)
            constexpr uint32_t idesc_ones = make_idesc_f16(BM, 16, AMAJ, 0);
            const uint64_t od = make_smem_desc(smem_u32(ones_tile), 128, 256, SWZ_NONE);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_f16_ss(tmem_base + 2 * BN + as * 16, make_smem_desc(sa + k * a_kstep, a_lbo, 1024, SWZ_128B), od,
                          idesc_ones, (kb > t.kb0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
          if (NCTA == 2) umma_commit_pair(empty + stage);
          else umma_commit(empty + stage);
          if (kb == t.kb1 - 1) {
            if (NCTA == 2) umma_commit_pair(tmem_full + as);
            else umma_commit(tmem_full + as);
          }
          if (++stage == Cfg::STAGES) stage = 0, phase ^= 1;
        }
      }
    }
  }
  } else {
    // ------------------------------------------------ epilogue: both warpgroups drain every accumulator stage
    const LavGemmEpilogue& e = p.epi;
    const int st = e.out_dtype == LAV_OUT_F16 ? ST_F16
                   : e.accumulate != LAV_ACCUMULATE ? ST_F32 : (p.splits > 1 ? ST_F32_RED : ST_F32_RMW);
    const int pre = e.residual != nullptr ? 1 : 0;
#define LAV_EPI(A, P, S)                                                                                          \
  epilogue_warps<BN, NCTA, A, P, S>(p, epi_stage, tmem_full, tmem_empty, tmem_base, warp, lane, total, stream_id,    \
                                    nstreams, rank)
#define LAV_EPI_TMA(A, F)                                                                                          \
  epilogue_warps_tma2<BN, NCTA, A, F>(p, &tmOut, &tmAux, epi_stage, tmem_full, tmem_empty, tmem_base, warp, lane,  \
                                      total, stream_id, nstreams, rank)
    // Only the (operand layout, epilogue) combinations the hot path uses are instantiated (epilogue_supported() on the
    // host rejects the others): forward (K,K): bias / GELU / residual / scatter / dropout; dgrad (K,MN): plain, GELU',
    // residual; wgrad (MN,MN): fp32 store / accumulate.
    if constexpr (AMAJ == LAV_MAJOR_MN) {
      if (st == ST_F32_RED) LAV_EPI(LAV_ACT_NONE, 0, ST_F32_RED);
      else if (st == ST_F32_RMW) LAV_EPI(LAV_ACT_NONE, 0, ST_F32_RMW);
      else LAV_EPI(LAV_ACT_NONE, 0, ST_F32);
    } else if constexpr (BMAJ == LAV_MAJOR_MN) {
      if (e.accumulate == LAV_ACCUMULATE) {  // split-K dgrad of a skinny problem (MLM decoder: M <= 160, K = vocab): fp32 atomics
        LAV_EPI(LAV_ACT_NONE, 0, ST_F32_RED);
      } else if (p.tma_store) {
        if (e.act == LAV_ACT_GELU_BWD) LAV_EPI_TMA(LAV_ACT_GELU_BWD, false);
        else if (st == ST_F16) LAV_EPI_TMA(LAV_ACT_NONE, false);
        else LAV_EPI_TMA(LAV_ACT_NONE, true);
      } else if (e.act == LAV_ACT_GELU_BWD) {
        if (st == ST_F16) LAV_EPI(LAV_ACT_GELU_BWD, 0, ST_F16);
        else LAV_EPI(LAV_ACT_GELU_BWD, 0, ST_F32);
      } else if (!pre) {
        if (st == ST_F16) LAV_EPI(LAV_ACT_NONE, 0, ST_F16);
        else LAV_EPI(LAV_ACT_NONE, 0, ST_F32);
      } else {
        LAV_EPI(LAV_ACT_NONE, 1, ST_F32);
      }
    } else {
      if (p.tma_store) {
        if (e.act == LAV_ACT_GELU) LAV_EPI_TMA(LAV_ACT_GELU, false);
        else if (st == ST_F16) LAV_EPI_TMA(LAV_ACT_NONE, false);
        else LAV_EPI_TMA(LAV_ACT_NONE, true);
      } else if (e.act == LAV_ACT_GELU) {
        if (st == ST_F16) LAV_EPI(LAV_ACT_GELU, 0, ST_F16);
        else LAV_EPI(LAV_ACT_GELU, 0, ST_F32);
      } else if (!pre) {
        if (st == ST_F16) LAV_EPI(LAV_ACT_NONE, 0, ST_F16);
        else LAV_EPI(LAV_ACT_NONE, 0, ST_F32);
      } else if (p.drop.on) {  // host restricts dropout to (no activation, residual, fp32 store)
        epilogue_warps<BN, NCTA, LAV_ACT_NONE, 1, ST_F32, true>(p, epi_stage, tmem_full, tmem_empty, tmem_base, warp, lane,
                                                                total, stream_id, nstreams, rank);
      } else {
        LAV_EPI(LAV_ACT_NONE, 1, ST_F32);
      }
    }
#undef LAV_EPI
#undef LAV_EPI_TMA
  }

  tc_fence_before();
  if (NCTA == 2) cluster_sync_all();  // neither CTA may exit (or free TMEM) while the peer can still signal it
  else __syncthreads();
  if (warp == 2) {
    if (NCTA == 2) tmem_dealloc_pair<Cfg::TMEM_ALLOC>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_ALLOC>(tmem_base);
  }
}

template <int BN, int AMAJ, int BMAJ, int NCTA>
static int launch_gemm(const void* A, int64_t lda, const void* B, int64_t ldb, GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, NCTA>;
  CUtensorMap tmA, tmB;
  int rc;
  if (AMAJ == LAV_MAJOR_K)
    rc = encode_tmap_2d_f16(&tmA, A, p.M, p.K, lda, BM, BK, CU_TENSOR_MAP_SWIZZLE_128B);
  else
    rc = encode_tmap_2d_f16(&tmA, A, p.K, p.M, lda, BK, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  if (BMAJ == LAV_MAJOR_K)
    rc = encode_tmap_2d_f16(&tmB, B, p.N, p.K, ldb, Cfg::BNL, BK, CU_TENSOR_MAP_SWIZZLE_128B);
  else
    rc = encode_tmap_2d_f16(&tmB, B, p.K, p.N, ldb, BK, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  p.n_blocks = (p.N + BN - 1) / BN;
  CUtensorMap tmOut = tmA, tmAux = tmA;  // placeholders unless the TMA-store epilogue is selected
  if (p.tma_store) {
    const LavGemmEpilogue& e = p.epi;
    const bool f32 = e.out_dtype == LAV_OUT_F32;
    rc = encode_tmap_2d(&tmOut, e.out, f32 ? 4 : 2, p.M, p.N, e.ldo, 32, 32,
                        f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    if (e.act == LAV_ACT_GELU && e.aux) {
      rc = encode_tmap_2d(&tmAux, e.aux, 2, p.M, p.N, e.ldaux, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
    }
  }
  auto kern = gemm_f16_kernel<BN, AMAJ, BMAJ, NCTA>;
  static bool attr_set = false;
  if (!attr_set) {
    LAV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int total = p.m_blocks * p.n_blocks * p.splits;
  const int streams = std::min(total, work_streams(NCTA));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(streams * NCTA);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (NCTA == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = NCTA, attr[na].val.clusterDim.y = 1, attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {  // start this kernel's prologue under the previous kernel's tail (griddepcontrol.wait in the kernel)
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  LAV_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, tmAux, p));
  count_launch();
  return LAV_OK;
}

template <int BN, int NCTA>
static int dispatch_major(const void* A, int64_t lda, int a_major, const void* B, int64_t ldb, int b_major,
                          GemmParams& p, cudaStream_t s) {
  if (a_major == LAV_MAJOR_K && b_major == LAV_MAJOR_K) return launch_gemm<BN, 0, 0, NCTA>(A, lda, B, ldb, p, s);
  if (a_major == LAV_MAJOR_K && b_major == LAV_MAJOR_MN) return launch_gemm<BN, 0, 1, NCTA>(A, lda, B, ldb, p, s);
  if (a_major == LAV_MAJOR_MN && b_major == LAV_MAJOR_MN) return launch_gemm<BN, 1, 1, NCTA>(A, lda, B, ldb, p, s);
  return set_error(LAV_E_INVALID, "lav_gemm_f16: (A MN-major, B K-major) is not instantiated (no caller on the hot path)");
}

}  // namespace lav

extern "C" int lav_gemm_f16(const void* A, int64_t lda, int a_major, const void* B, int64_t ldb, int b_major, int M,
                            int N, int K, const LavGemmEpilogue* epi, int split_k, void* stream) {
  using namespace lav;
  LAV_REQUIRE(A && B && epi && epi->out, "lav_gemm_f16: null pointer");
  LAV_REQUIRE(M > 0 && N > 0 && K > 0, "lav_gemm_f16: empty problem %dx%dx%d", M, N, K);
  LAV_REQUIRE((a_major | b_major) >= 0 && a_major <= 1 && b_major <= 1, "lav_gemm_f16: bad major");
  LAV_REQUIRE(!(epi->accumulate == LAV_ACCUMULATE && epi->out_dtype != LAV_OUT_F32),
              "lav_gemm_f16: accumulation needs an fp32 output");
  LAV_REQUIRE(!(epi->act != LAV_ACT_NONE && split_k > 1), "lav_gemm_f16: activation with split-K");
  LAV_REQUIRE(!(epi->act == LAV_ACT_GELU_BWD && !epi->aux), "lav_gemm_f16: GELU_BWD needs aux");
  LAV_REQUIRE(!(epi->act != LAV_ACT_NONE && epi->residual), "lav_gemm_f16: activations cannot be combined with a residual");
  LAV_REQUIRE(!(epi->act != LAV_ACT_NONE && epi->accumulate == LAV_ACCUMULATE), "lav_gemm_f16: activations cannot accumulate");
  {
    const bool f16o = epi->out_dtype == LAV_OUT_F16, acc = epi->accumulate == LAV_ACCUMULATE;
    bool ok;
    if (a_major == LAV_MAJOR_MN)        // wgrad: plain fp32 store / accumulate
      ok = epi->act == LAV_ACT_NONE && !f16o && !epi->residual && !epi->row_map && !epi->row_scale;
    else if (b_major == LAV_MAJOR_MN)   // dgrad: plain / GELU' / fp32 residual / plain fp32 accumulate (split-K)
      ok = epi->act != LAV_ACT_GELU && !(epi->residual && f16o) && !(epi->residual && epi->act != LAV_ACT_NONE) &&
           (!acc || (epi->act == LAV_ACT_NONE && !f16o && !epi->residual && !epi->row_map && !epi->row_scale && !epi->bias));
    else                                // forward
      ok = epi->act != LAV_ACT_GELU_BWD && !acc && !(epi->residual && f16o);
    LAV_REQUIRE(ok, "lav_gemm_f16: this (operand layout, epilogue) combination is not instantiated (a_major %d b_major %d "
                "act %d out_dtype %d accumulate %d residual %d)", a_major, b_major, epi->act, epi->out_dtype,
                epi->accumulate, epi->residual != nullptr);
  }
  GemmParams p;
  p.M = M, p.N = N, p.K = K;
  p.k_blocks = (K + BK - 1) / BK;
  p.epi = *epi;
  p.drop = make_drop(&epi->drop);
  p.bias_grad = epi->bias_grad;
  LAV_REQUIRE(!epi->bias_grad || (a_major == LAV_MAJOR_MN && epi->accumulate == LAV_ACCUMULATE && epi->act == LAV_ACT_NONE &&
                                  !epi->row_map && !epi->residual),
              "lav_gemm_f16: bias_grad is a wgrad option (A MN-major, accumulating plain epilogue)");
  LAV_REQUIRE(!p.drop.on || (epi->act == LAV_ACT_NONE && epi->residual && epi->out_dtype == LAV_OUT_F32 &&
                             epi->accumulate != LAV_ACCUMULATE),
              "lav_gemm_f16: epilogue dropout needs (no activation, residual, fp32 store)");
  {
    // TMA-store epilogue: plain stores only (no scatter / residual / DropPath / accumulation / dropout), GELU outputs
    // fp16, 16-byte aligned rows.  LAV_GEMM_TMA_STORE=0 keeps the register path (A/B testing).
    static int tma_ok = -1;
    if (tma_ok < 0) {
      const char* ev = getenv("LAV_GEMM_TMA_STORE");
      tma_ok = (ev && ev[0] == '0') ? 0 : 1;
    }
    const int eb = epi->out_dtype == LAV_OUT_F32 ? 4 : 2;
    bool ok = tma_ok && !epi->row_map && !epi->residual && !epi->row_scale && epi->accumulate != LAV_ACCUMULATE &&
              !p.drop.on && ((uintptr_t)epi->out % 16) == 0 && ((epi->ldo * eb) % 16) == 0;
    if (epi->act != LAV_ACT_NONE) ok = ok && epi->out_dtype == LAV_OUT_F16;
    if (epi->act == LAV_ACT_GELU && epi->aux) ok = ok && ((uintptr_t)epi->aux % 16) == 0 && (epi->ldaux % 8) == 0;
    p.tma_store = ok ? 1 : 0;
  }
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("LAV_GEMM_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    p.debug = dbg;
  }
  // CTA pairs (256-row tiles, cta_group::2) whenever the problem has more than one 128-row block.
  const int ncta = (pair_enabled() && M > BM && !epi->bias_grad) ? 2 : 1;
  p.m_blocks = (M + BM * ncta - 1) / (BM * ncta);
  // Tile width BN and split-K factor chosen by a small cost model (SM clocks):
  //   per k-block  max(MMA issue 2*BN, operand feed (BM + BN/ncta)*BK*2 bytes at ~110 B/clk/SM), times the number of
  //   waves over the work streams, plus one exposed epilogue (atomics cost more).
  const int nstr = work_streams(ncta);
  const bool may_split = split_k != 1 && epi->accumulate == LAV_ACCUMULATE && epi->act == LAV_ACT_NONE;
  int bn = 128, splits = 1;
  double best = 1e30;
  const int cands[4] = {256, 192, 128, 64};
  static int forced_bn = -1;  // LAV_GEMM_BN=64|128|192|256: tile-width override for tools/bench_gemm.py (tuning the model)
  if (forced_bn < 0) {
    const char* e = getenv("LAV_GEMM_BN");
    forced_bn = e ? atoi(e) : 0;
  }
  for (int ci = 0; ci < 4; ++ci) {
    const int c = cands[ci];
    if (forced_bn > 0 && c != forced_bn && !(epi->bias_grad && forced_bn > 192 && c == 192)) continue;
    if (ncta == 2 && c != 256 && c != 128) continue;        // pair tiles: B halves must be whole 64-row atoms
    if (epi->bias_grad && c > 192) continue;                // the side accumulator needs 32 spare TMEM columns
    if (c > 64 && c >= 2 * ((N + 63) / 64 * 64) && !(ncta == 2 && c == 128)) continue;  // far wider than the problem
    const int tiles = p.m_blocks * ((N + c - 1) / c);
    const double kb_clk = std::max(2.0 * c, (BM + c / ncta) * BK * 2 / 110.0);
    const int smax = may_split ? (split_k > 1 ? split_k : std::max(1, std::min(p.k_blocks / 4, 64))) : 1;
    for (int sp = (split_k > 1 ? split_k : 1); sp <= smax; ++sp) {
      const int kbs = (p.k_blocks + sp - 1) / sp;
      const int items = tiles * ((p.k_blocks + kbs - 1) / kbs);
      const int waves = (items + nstr - 1) / nstr;
      const double cost = waves * (kbs * kb_clk + 300.0) + c * (sp > 1 ? 10.0 : 6.0);
      if (cost < best) best = cost, bn = c, splits = sp;
    }
  }
  splits = std::max(1, std::min(splits, p.k_blocks));
  LAV_REQUIRE(splits == 1 || epi->accumulate == LAV_ACCUMULATE, "lav_gemm_f16: split-K needs LAV_ACCUMULATE");
  p.kb_per_split = (p.k_blocks + splits - 1) / splits;
  p.splits = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;  // no empty split
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (ncta == 2) {
    if (bn == 256) return dispatch_major<256, 2>(A, lda, a_major, B, ldb, b_major, p, s);
    return dispatch_major<128, 2>(A, lda, a_major, B, ldb, b_major, p, s);
  }
  switch (bn) {
    case 64: return dispatch_major<64, 1>(A, lda, a_major, B, ldb, b_major, p, s);
    case 192: return dispatch_major<192, 1>(A, lda, a_major, B, ldb, b_major, p, s);
    case 256: return dispatch_major<256, 1>(A, lda, a_major, B, ldb, b_major, p, s);
    default: return dispatch_major<128, 1>(A, lda, a_major, B, ldb, b_major, p, s);
  }
}
