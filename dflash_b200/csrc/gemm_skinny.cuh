// Weight-streaming "skinny" GEMM for the DFlash draft step (sm_100a, tcgen05 + TMEM + TMA).
//
//   Y[m, n] = sum_k X[m, k] * W[n, k]        X: [rows<=MB, K] bf16, W: [N, K] bf16 (nn.Linear layout)
//
// The draft step has 16..256 activation rows against 10^7..10^9 weight elements, so the kernel is an
// HBM stream of W. It runs swap-AB: a 128-row slab of W is the UMMA "A" operand (M=128), the MB
// activation rows are the UMMA "B" operand (N=MB), and the fp32 accumulator D[128 x MB] lives in
// TMEM. One elected thread issues tcgen05.mma; one elected thread issues TMA; four warps drain TMEM.
//
// Work split ("stream-K"): the (n-tile, k-block) grid is flattened n-major and cut into gridDim.y
// equal contiguous unit ranges, so every CTA streams the same number of bytes regardless of N/128.
// A CTA that ends inside a tile writes an fp32 partial into slot (cta - first_cta_of_tile); the
// consumer kernel sums the slots in slot order (deterministic, no atomics).
//
// Programmatic dependent launch: weights never depend on the previous kernel, so the producer
// issues the first kStages W tiles BEFORE griddepcontrol.wait and only the activation tiles after
// it. The HBM pipe therefore stays full across kernel boundaries.
//
// Replaces the reference's nn.Linear call sites on the hot path: model/dflash.py:70-76,101,177,
// Qwen3MLP (transformers) via :143, and target.lm_head at :238-245.
#pragma once
#include "ptx.cuh"

namespace dfl {

constexpr int kTileN = 128;   // weight rows per tile (UMMA M)
constexpr int kTileK = 64;    // bf16 elements per k-block (= one 128-byte swizzle row)
constexpr int kUmmaK = 16;

enum GemmMode : int {
  kModePartials = 0,  // write fp32 partial sums to ws[slot][m][n]
  kModeArgmax = 1,    // whole tiles per CTA; per-CTA (max, argmax) of the bf16-rounded logits
  kModeArgmaxDump = 2,  // kModeArgmax that also stores the bf16 logits (parity tests; keeps the store addressing
                        // out of the hot instantiation's registers)
  kModeTopK = 3,      // whole tiles per CTA; per-CTA top-4 of the bf16-rounded logits per activation row
                      // (multi-candidate drafting: benchmark_candidate_solutions.py:181-249)
  kModeSample = 4,    // whole tiles per CTA; per-CTA argmax of  logit / T + Gumbel noise  per activation row = a draw
                      // from softmax(logits / T) (Gumbel-max, the construction the posterior sampler uses too)
};
constexpr int kTopK = 4;

struct GemmArgs {
  int n_tiles;    // ceil(N / 128)
  int k_blocks;   // K / 64
  int N;          // weight rows in range (output columns)
  int w_row0;     // first weight row of the range inside the TMA tensor
  int x_row0;     // first activation row inside the activation TMA tensor
  int m_valid;    // activation rows that are written out (<= MB)
  // kModePartials
  float* ws;          // [slots][ws_rows][ws_ld]
  int ws_rows;
  long long ws_ld;
  // kModeArgmax
  float* cand_val;            // [ranges][cand_ld] (kModeTopK: [ranges][cand_ld][4], best first)
  int* cand_idx;              // [ranges][cand_ld]
  __nv_bfloat16* logits;      // optional [m_valid][logits_ld] (bf16-rounded), may be null
  long long logits_ld;
  // Pre-wait L2 prefetch: while this (PDL-launched) kernel waits for its predecessor, the four idle
  // epilogue warps prefetch the CTA's next `pf_units` weight units (beyond the kStages tiles already
  // in flight to smem) into L2 with plain prefetch.global.L2, one 128-byte weight-row segment per thread.
  const void* w_ptr;   // weight matrix base (row-major bf16, pitch w_ld elements)
  long long w_ld;
  int w_rows;          // rows of the weight matrix (prefetch bound)
  int pf_units;        // 0 = off
  int late_w;          // experiment: issue the first weight tiles only after griddepcontrol.wait
  // optional per-CTA phase timestamps (globaltimer ns), [ranges * groups][8]: 0 kernel entry, 1 prologue done,
  // 2 producer past griddepcontrol.wait, 3 first stage landed (MMA warp), 4 last MMA issued, 5 last accumulator
  // complete (epilogue), 6 epilogue stores issued (scripts/gemm_trace.py)
  unsigned long long* trace;
  // Column groups (wide batches): the activation rows are cut into `groups` slabs of MB rows; the grid is
  // (groups, ranges) with the group index fastest, so the `groups` CTAs that stream one weight range are
  // launched side by side and share it through L2 (HBM sees every weight byte once per step).
  int groups;          // >= 1
  int cand_ld;         // kModeArgmax: row pitch of cand_val/cand_idx (= groups * MB)
  // kModeSample
  float inv_temp;                        // 1 / temperature
  unsigned long long seed;               // Philox key
  unsigned long long step_base;          // Philox counter words 2-3 = step_base + *rng_step
  const unsigned long long* rng_step;    // optional device counter (bumped once per cycle by the accept kernel)
};

// The CTA that owns flat unit x when T units are cut into G ranges [floor(g*T/G), floor((g+1)*T/G)).
__host__ __device__ inline int cta_of_unit(long long x, long long T, long long G) {
  return static_cast<int>(((x + 1) * G - 1) / T);
}
__host__ __device__ inline long long unit_begin(long long g, long long T, long long G) {
  return g * T / G;
}
// Number of partial slots tile t has (>= 1), and the first CTA that touches it.
__host__ __device__ inline int tile_first_cta(int t, int k_blocks, long long T, long long G) {
  return cta_of_unit(static_cast<long long>(t) * k_blocks, T, G);
}
__host__ __device__ inline int tile_num_slots(int t, int k_blocks, long long T, long long G) {
  return cta_of_unit(static_cast<long long>(t + 1) * k_blocks - 1, T, G) -
         tile_first_cta(t, k_blocks, T, G) + 1;
}

// smem budget of the TMA pipeline per mode (measured on B200, profiles/r1_summary.md):
//  * partials GEMMs (33-200 MB each, 21 per step, chained by PDL): ~110 KB, so that the successor's CTA
//    can be co-resident and pre-load its first stages while the predecessor drains (755 us/step vs 768 us
//    with 215 KB);
//  * lm_head argmax GEMM (1.24 GB in one launch): everything, 11 stages -> 0.967 of the measured copy
//    bandwidth instead of 0.93.
#ifndef DFLASH_GEMM_SMEM_KB_PARTIALS
#define DFLASH_GEMM_SMEM_KB_PARTIALS 110
#endif
#ifndef DFLASH_GEMM_SMEM_KB_ARGMAX
#define DFLASH_GEMM_SMEM_KB_ARGMAX 215
#endif
#ifndef DFLASH_GEMM_SMEM_KB_WIDE   // partials GEMMs with >= 64 activation rows per group (+ 16 KB store staging)
#define DFLASH_GEMM_SMEM_KB_WIDE 196
#endif

template <int MB, int MODE = 0>
struct GemmCfg {
  static constexpr int kWBytes = kTileN * kTileK * 2;   // 16 KB
  static constexpr int kXBytes = MB * kTileK * 2;
  static constexpr int kStageBytes = kWBytes + kXBytes;
  // wide activation tiles (batched engines) need the whole SM to keep >= 4 stages in flight
  static constexpr int kBudget =
      (MODE != 0 ? DFLASH_GEMM_SMEM_KB_ARGMAX : (MB >= 64 ? DFLASH_GEMM_SMEM_KB_WIDE : DFLASH_GEMM_SMEM_KB_PARTIALS)) * 1024;
  static constexpr int kStages = kBudget / kStageBytes < 3 ? 3 : kBudget / kStageBytes;
  static constexpr int kTmemCols = (2 * MB < 32) ? 32 : 2 * MB;
  // epilogue warps: warp w drains TMEM lane quarter w % 4. The 256-wide argmax epilogue keeps one packed running
  // best per activation row in registers, so it splits the columns over two warp sets (128 registers each).
  static constexpr int kEpiWarps = (MODE != 0 && MB >= 128) ? 8 : 4;
  static constexpr int kThreads = (kEpiWarps + 2) * 32;  // + TMA warp + MMA warp
  static constexpr int kColsPerThread = MB / (kEpiWarps / 4);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Order-preserving 16-bit key of a bf16 value (larger value <=> larger key; -0 == +0; key 0 is below every value).
__device__ __forceinline__ uint32_t bf16_order_key(float v) {
  uint32_t u = __float_as_uint(bf16_round(v)) >> 16;
  if (u == 0x8000u) u = 0;
  return (u & 0x8000u) ? (~u & 0xFFFFu) : (u | 0x8000u);
}
__device__ __forceinline__ float bf16_from_order_key(uint32_t k) {
  const uint32_t u = (k & 0x8000u) ? (k & 0x7FFFu) : (~k & 0xFFFFu);
  return __uint_as_float(u << 16);
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <int MB, int MODE>
__global__ void __launch_bounds__(GemmCfg<MB, MODE>::kThreads, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                   const GemmArgs a) {
  using Cfg = GemmCfg<MB, MODE>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sW = smem;
  uint8_t* sX = smem + S * Cfg::kWBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* empty = full + S;
  uint64_t* tfull = empty + S;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  unsigned long long* tr = a.trace ? a.trace + (static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x) * 8 : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = global_ns();
  DFL_TRACE(0);
#ifdef DFLASH_GEMM_TOP_TRIGGER
  pdl_trigger();  // (experiment) release the dependent before this kernel's own prologue
#endif
  const long long T = static_cast<long long>(a.n_tiles) * a.k_blocks;
  const long long G = gridDim.y;           // weight ranges
  const int cta = blockIdx.y;              // this CTA's weight range
  const int m0 = blockIdx.x * MB;          // first activation row of this CTA's column group
  const int mv = a.m_valid - m0;           // valid rows in the group (may exceed MB)
  long long u0, u1;
  constexpr bool kArgmax = MODE != kModePartials;
  if (kArgmax) {  // whole tiles only
    u0 = (cta * static_cast<long long>(a.n_tiles) / G) * a.k_blocks;
    u1 = ((cta + 1) * static_cast<long long>(a.n_tiles) / G) * a.k_blocks;
  } else {
    u0 = unit_begin(cta, T, G);
    u1 = unit_begin(cta + 1, T, G);
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], Cfg::kEpiWarps * 32);
    }
    mbar_fence_init();
  }
  constexpr int kTmaWarp = Cfg::kEpiWarps, kMmaWarp = Cfg::kEpiWarps + 1;
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmX);
  }
  if (warp == kMmaWarp) {
    tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (tr && threadIdx.x == 0) tr[1] = global_ns();
  // Let the next kernel in the stream start its own prologue / weight prefetch right away.
#if !defined(DFLASH_GEMM_LATE_TRIGGER) && !defined(DFLASH_GEMM_TOP_TRIGGER)
  pdl_trigger();
#endif

  if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // one group: every weight byte is read once -> evict first; several groups re-read it from L2
      const uint64_t polW = gridDim.x == 1 ? l2_policy_evict_first() : l2_policy_evict_normal();
      const uint64_t polX = l2_policy_evict_last();
      const long long n_units = u1 - u0;
      const int npre = n_units < S ? static_cast<int>(n_units) : S;
      // weight tiles first: they do not depend on the predecessor kernel
      if (a.late_w == 1) pdl_wait();
#ifdef DFLASH_PREWAIT_STAGES   // (experiment) only this many weight tiles before griddepcontrol.wait, the rest after
      const int n_early = npre < DFLASH_PREWAIT_STAGES ? npre : DFLASH_PREWAIT_STAGES;
#else
      const int n_early = npre;
#endif
      auto issue_w = [&](int i) {
        const long long u = u0 + i;
        const int tile = static_cast<int>(u / a.k_blocks);
        const int kb = static_cast<int>(u % a.k_blocks);
        mbar_expect_tx(&full[i], Cfg::kStageBytes);
        tma_load_2d(sW + i * Cfg::kWBytes, &tmW, &full[i], kb * kTileK, a.w_row0 + tile * kTileN,
                    polW);
      };
      for (int i = 0; i < n_early; ++i) issue_w(i);
      pdl_wait();
      for (int i = n_early; i < npre; ++i) issue_w(i);
#ifdef DFLASH_GEMM_LATE_TRIGGER
      pdl_trigger();  // (experiment) the dependent small kernel becomes resident only once this GEMM really starts
#endif
      if (tr) tr[2] = global_ns();
      DFL_TRACE_ANY(1);
      for (int i = 0; i < npre; ++i) {
        const int kb = static_cast<int>((u0 + i) % a.k_blocks);
        tma_load_2d(sX + i * Cfg::kXBytes, &tmX, &full[i], kb * kTileK, a.x_row0 + m0, polX);
      }
      int stage = npre % S;
      uint32_t phase = (npre == S) ? 1u : 0u;
      for (long long u = u0 + npre; u < u1; ++u) {
        const int tile = static_cast<int>(u / a.k_blocks);
        const int kb = static_cast<int>(u % a.k_blocks);
        mbar_wait(&empty[stage], phase ^ 1u);
        if (a.late_w == 2) {  // timing experiment only (wrong results): no activation tile after the first S units
          mbar_expect_tx(&full[stage], Cfg::kWBytes);
          tma_load_2d(sW + stage * Cfg::kWBytes, &tmW, &full[stage], kb * kTileK, a.w_row0 + tile * kTileN, polW);
          if (++stage == S) { stage = 0; phase ^= 1u; }
          continue;
        }
        mbar_expect_tx(&full[stage], Cfg::kStageBytes);
        tma_load_2d(sW + stage * Cfg::kWBytes, &tmW, &full[stage], kb * kTileK,
                    a.w_row0 + tile * kTileN, polW);
        tma_load_2d(sX + stage * Cfg::kXBytes, &tmX, &full[stage], kb * kTileK, a.x_row0 + m0, polX);
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileN, MB);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long u = u0;
      while (u < u1) {
        const long long tile = u / a.k_blocks;
        const long long seg_end = (tile + 1) * a.k_blocks < u1 ? (tile + 1) * a.k_blocks : u1;
        mbar_wait(&tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * MB);
        uint32_t accumulate = 0;
        for (; u < seg_end; ++u) {
          mbar_wait(&full[stage], phase);
          if (tr && u == u0) tr[3] = global_ns();
          if (u == u0) DFL_TRACE_ANY(3);
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(smem_u32(sW + stage * Cfg::kWBytes));
          const uint64_t db = umma_desc_sw128(smem_u32(sX + stage * Cfg::kXBytes));
#pragma unroll
          for (int k = 0; k < kTileK / kUmmaK; ++k) {
            // +32 B per K=16 step inside the 128-byte swizzle row (address field is in 16 B units)
            umma_bf16_ss(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k),
                         idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(&empty[stage]);  // smem slot is free once these MMAs retire
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull[acc]);  // accumulator complete
        if (tr && u >= u1) tr[4] = global_ns();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0..kEpiWarps-1)
    if (a.pf_units > 0) {
      // pre-wait L2 prefetch of units [u0 + S, u0 + S + pf_units): thread t takes weight row t of each
      // unit's 128 x 128 B box, so one pass of the 128 threads covers one 16 KB unit
      const char* wb = static_cast<const char*>(a.w_ptr);
      long long u = u0 + S;
      const long long ue = u + a.pf_units < u1 ? u + a.pf_units : u1;
      for (; u < ue; ++u) {
        const int tile = static_cast<int>(u / a.k_blocks);
        const int kb = static_cast<int>(u % a.k_blocks);
        const int wrow = a.w_row0 + tile * kTileN + static_cast<int>(threadIdx.x);
        if (threadIdx.x < kTileN && wrow < a.w_rows)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], 128;\n" ::"l"(
              wb + (static_cast<long long>(wrow) * a.w_ld + kb * kTileK) * 2));
      }
    }
    pdl_wait();
    int acc = 0;
    uint32_t acc_phase = 0;
    const int quarter = warp & 3;                       // TMEM lanes [32 * quarter, +32)
    constexpr int kCols = Cfg::kColsPerThread;          // activation rows this thread drains per accumulator
    const int col0 = (warp >> 2) * kCols;               // second warp set: the upper half of the columns
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const int row_in_tile = quarter * 32 + lane;
    const int epi_tid = static_cast<int>(threadIdx.x);  // epilogue warps are warps [0, kEpiWarps)

    // kModeArgmax: one running best per activation row in a register, packed as
    //   (order key of the bf16-rounded logit) << 16 | (0xFFFF - tile)
    // so that keeping the best is ONE integer max per logit, and a tie keeps the lower tile = the lower vocab
    // index (this thread's weight row inside the tile is fixed). Decoded and reduced over threads at the end.
    constexpr int kKeep = MODE == kModeTopK ? kTopK : 1;  // running bests per activation row (sorted, descending)
    constexpr bool kSample = MODE == kModeSample;
    uint32_t best[kArgmax ? kCols * kKeep : 1];
    float bestf[kSample ? kCols : 1];  // kModeSample: fp32 perturbed key per row (best[] then holds its tile)
    if (kArgmax) {
#pragma unroll
      for (int j = 0; j < kCols * kKeep; ++j) best[j] = 0u;
    }
    if (kSample) {
#pragma unroll
      for (int j = 0; j < kCols; ++j) bestf[j] = -INFINITY;
    }
    unsigned long long rstep = 0;
    if (kSample) rstep = a.step_base + (a.rng_step != nullptr ? *a.rng_step : 0ull);

    long long u = u0;
    while (u < u1) {
      const int tile = static_cast<int>(u / a.k_blocks);
      const long long seg_end =
          static_cast<long long>(tile + 1) * a.k_blocks < u1 ? static_cast<long long>(tile + 1) * a.k_blocks : u1;
      const int n = tile * kTileN + row_in_tile;
      mbar_wait(&tfull[acc], acc_phase);
      if (tr && seg_end >= u1 && threadIdx.x == 0) tr[5] = global_ns();
      tc_fence_after();
      float* dst = nullptr;
      if (MODE == kModePartials) {
        const int slot = cta - tile_first_cta(tile, a.k_blocks, T, G);
        dst = a.ws + (static_cast<long long>(slot) * a.ws_rows + m0) * a.ws_ld + n;
      }
      const uint32_t tile_tag = 0xFFFFu - static_cast<uint32_t>(tile);
      // kChunk columns per TMEM round trip: the loads of a chunk are all issued before the one wait
      constexpr bool kStageOut = (MODE == kModePartials) && (MB >= 64);
      constexpr int kChunk = (MODE == kModePartials) ? (kStageOut ? 32 : kCols)
                                                      : ((kCols >= 32 && kCols < 128) ? 32 : 16);  // (register budget)
      // Wide partial tiles go out through a 16 KB transpose buffer: a thread owns one weight row (column n of the
      // output) and would store its 128-256 values 4 bytes at a time, one 128-byte warp store per activation row
      // -- measured 4.4 / 8.5 us per tile at 128 / 256 rows (scripts/gemm_trace.py), exposed at the end of every
      // GEMM. Transposed, a warp writes one 512-byte output row per store.
      __shared__ __align__(16) float s_out[kStageOut ? 32 * kTileN : 4];
      const bool vec_ok = kStageOut && ((a.ws_ld & 3) == 0);
#pragma unroll
      for (int c = 0; c < kCols / kChunk; ++c) {
        float v[kChunk];
#pragma unroll
        for (int q = 0; q < kChunk / 16; ++q)
          tmem_ld16(tmem_base + lane_addr + static_cast<uint32_t>(acc * MB + col0 + c * kChunk + q * 16), v + q * 16);
        tmem_ld_wait();
        if (c == kCols / kChunk - 1) {
          // all of this thread's share of the accumulator is in registers: hand the TMEM stage back
          tc_fence_before();
          mbar_arrive(&tempty[acc]);
        }
        if (kStageOut) {
#pragma unroll
          for (int j = 0; j < kChunk; ++j) s_out[j * kTileN + row_in_tile] = v[j];
          asm volatile("bar.sync 1, 128;\n" ::: "memory");
          const int nn = tile * kTileN + lane * 4;
          float* row0 = a.ws + (static_cast<long long>(cta - tile_first_cta(tile, a.k_blocks, T, G)) * a.ws_rows + m0) *
                                   a.ws_ld + nn;
          for (int j = quarter; j < kChunk; j += 4) {
            const int m = col0 + c * kChunk + j;
            if (m >= mv) continue;
            const float4 x = *reinterpret_cast<const float4*>(&s_out[j * kTileN + lane * 4]);
            float* o = row0 + static_cast<long long>(m) * a.ws_ld;
            if (vec_ok && nn + 3 < a.N) {
              *reinterpret_cast<float4*>(o) = x;
            } else {
              if (nn < a.N) o[0] = x.x;
              if (nn + 1 < a.N) o[1] = x.y;
              if (nn + 2 < a.N) o[2] = x.z;
              if (nn + 3 < a.N) o[3] = x.w;
            }
          }
          asm volatile("bar.sync 1, 128;\n" ::: "memory");
          continue;
        }
        if (MODE == kModePartials) {
          if (n < a.N) {
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
              const int m = col0 + c * kChunk + j;
              if (m < mv) dst[static_cast<long long>(m) * a.ws_ld] = v[j];
            }
          }
        } else if (kSample) {
          if (n < a.N) {
#pragma unroll
            for (int j4 = 0; j4 < kChunk; j4 += 4) {
              // one Philox call per (vocab row n, group of four activation rows): four uniforms in (0, 1]
              uint32_t rnd[4];
              philox4x32(static_cast<uint32_t>(n), static_cast<uint32_t>((m0 + col0 + c * kChunk + j4) >> 2),
                         static_cast<uint32_t>(rstep), static_cast<uint32_t>(rstep >> 32) ^ 0x40000000u,
                         static_cast<uint32_t>(a.seed), static_cast<uint32_t>(a.seed >> 32), rnd);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int j = j4 + q;
                const float e = -__logf(u32_to_unit(rnd[q]));                       // Exp(1)
                const float key = bf16_round(v[j]) * a.inv_temp - logf(fmaxf(e, 1e-30f));  // = logit/T + Gumbel
                if (key > bestf[kSample ? c * kChunk + j : 0]) {
                  bestf[kSample ? c * kChunk + j : 0] = key;
                  best[kSample ? c * kChunk + j : 0] = static_cast<uint32_t>(tile);
                }
              }
            }
          }
        } else if (n < a.N) {
#pragma unroll
          for (int j = 0; j < kChunk; ++j) {
            uint32_t key = (bf16_order_key(v[j]) << 16) | tile_tag;
            if (MODE == kModeTopK) {
              // insertion into the sorted 4-entry list: a compare-exchange per entry
#pragma unroll
              for (int q = 0; q < kKeep; ++q) {
                uint32_t& b = best[kArgmax ? (c * kChunk + j) * kKeep + q : 0];
                const uint32_t hi = key > b ? key : b;
                key = key > b ? b : key;
                b = hi;
              }
            } else {
              uint32_t& b = best[kArgmax ? c * kChunk + j : 0];
              b = key > b ? key : b;
            }
            if (MODE == kModeArgmaxDump) {
              const int m = col0 + c * kChunk + j;
              if (m < mv) a.logits[static_cast<long long>(m0 + m) * a.logits_ld + n] = __float2bfloat16_rn(v[j]);
            }
          }
        }
      }
      u = seg_end;
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (tr && threadIdx.x == 0) tr[6] = global_ns();
    DFL_TRACE(2);

    if (kSample) {
      // reduce (perturbed key, tile) over the 128 weight rows through the idle pipeline smem
      constexpr int kPitchS = MB + 1;
      float* skey = reinterpret_cast<float*>(smem);
      uint32_t* stile = reinterpret_cast<uint32_t*>(smem) + kTileN * kPitchS;
#pragma unroll
      for (int j = 0; j < (kSample ? kCols : 1); ++j) {
        skey[row_in_tile * kPitchS + col0 + j] = bestf[j];
        stile[row_in_tile * kPitchS + col0 + j] = best[j];
      }
      asm volatile("bar.sync 1, %0;\n" ::"n"(Cfg::kEpiWarps * 32) : "memory");
      for (int j = epi_tid; j < MB; j += Cfg::kEpiWarps * 32) {
        float bk = -INFINITY;
        int bn = 0x7fffffff;
        for (int r = 0; r < kTileN; ++r) {
          const float k = skey[r * kPitchS + j];
          if (k > bk) { bk = k; bn = static_cast<int>(stile[r * kPitchS + j]) * kTileN + r; }
        }
        a.cand_val[static_cast<long long>(cta) * a.cand_ld + m0 + j] = bk;
        a.cand_idx[static_cast<long long>(cta) * a.cand_ld + m0 + j] = bn;
      }
    } else if (kArgmax) {
      // Reduce over the 128 weight rows of the tile shape through shared memory (the pipeline stages are idle by
      // now): keys[row][col], padded pitch so that both the row-wise writes and the column-wise reads are
      // conflict-free. Rows are visited in ascending order with a strict compare, so among equal keys (same value,
      // same tile) the lowest row, i.e. the lowest vocab index, wins.
      constexpr int kPitch = MB * kKeep + 1;
      static_assert(kTileN * kPitch * 4 <= S * Cfg::kStageBytes, "argmax reduction scratch exceeds the pipeline smem");
      uint32_t* keys = reinterpret_cast<uint32_t*>(smem);
#pragma unroll
      for (int j = 0; j < (kArgmax ? kCols * kKeep : 1); ++j) keys[row_in_tile * kPitch + col0 * kKeep + j] = best[j];
      asm volatile("bar.sync 1, %0;\n" ::"n"(Cfg::kEpiWarps * 32) : "memory");
      for (int j = epi_tid; j < MB; j += Cfg::kEpiWarps * 32) {
        uint32_t bk[kKeep];
        int brow[kKeep];
#pragma unroll
        for (int q = 0; q < kKeep; ++q) { bk[q] = 0u; brow[q] = 0; }
        for (int r = 0; r < kTileN; ++r) {
#pragma unroll
          for (int e = 0; e < kKeep; ++e) {
            uint32_t k = keys[r * kPitch + j * kKeep + e];
            int kr = r;
#pragma unroll
            for (int q = 0; q < kKeep; ++q) {  // strict compare: among equal keys the lower row stays in front
              if (k > bk[q]) {
                const uint32_t tk = bk[q]; bk[q] = k; k = tk;
                const int tr = brow[q]; brow[q] = kr; kr = tr;
              }
            }
          }
        }
        const long long o = (static_cast<long long>(cta) * a.cand_ld + m0 + j) * kKeep;
#pragma unroll
        for (int q = 0; q < kKeep; ++q) {
          a.cand_val[o + q] = bk[q] ? bf16_from_order_key(bk[q] >> 16) : -INFINITY;
          a.cand_idx[o + q] = bk[q] ? static_cast<int>(0xFFFFu - (bk[q] & 0xFFFFu)) * kTileN + brow[q] : 0x7fffffff;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

}  // namespace dfl
