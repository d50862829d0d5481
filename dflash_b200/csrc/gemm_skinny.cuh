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
#include "epilogues.cuh"
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
  // Fused SwiGLU epilogue (epilogues.cuh): stream-K as kModePartials, but a split tile is FINISHED inside the GEMM --
  // the CTA that owns the tile's first k-blocks (always the last segment of its range) adds the partial accumulators
  // the other CTAs of the tile left in `part`, in slot order, and runs the epilogue on the finished tile. Tiles that
  // one CTA owns entirely never leave TMEM/registers. No fp32 partial plane, no consumer kernel.
  kModeSwiglu = 5,    // tile = 64 gate rows + 64 up rows of the same columns -> bf16 silu(gate) * up
  // The target-context injection in ONE kernel (model/dflash.py:177 + :237): stream-K partial planes as kModePartials,
  // then -- after a device-wide arrival count over the CTAs of the launch, which are all co-resident -- the CTAs
  // themselves run the row pass over the context rows (sum of the slots -> bf16 -> hidden_norm -> a_in). Before their
  // first tile the epilogue warps gather the block rows' embeddings and apply layer 0's input_layernorm.
  kModeCtxNorm = 6,
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
  // Fused modes: exchange of partial accumulators between the CTAs of a split tile
  float* part;            // [groups * ranges][MB/4][128][4] fp32: the partial of CTA (group, range)'s FIRST segment
  unsigned int* flags;    // [groups][n_tiles] arrivals of a tile's non-finishing CTAs (reset by the finisher)
  SwigluEpi sw;           // kModeSwiglu
  CtxNormEpi cn;          // kModeCtxNorm
  // kModeArgmax / kModeSample inside the engine: the LAST CTA to finish also reduces the per-CTA candidates to the
  // drafted tokens (block_ids[:, 1:bs], model/dflash.py:247) -- no separate reduce launch. Off when tok_counter is null.
  unsigned int* tok_counter;
  long long* tok_block_ids;     // [R][tok_bs]
  long long* tok_draft_tokens;  // [R*tok_SL]
  int tok_SL, tok_bs, tok_rows;
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
//  * chained projection GEMMs (33-200 MB each, 21 per step, chained by PDL): ~100 KB + the 8 KB epilogue tile, so that
//    the successor's CTA can be co-resident and pre-load its first stages while the predecessor drains (755 us/step
//    vs 768 us with 215 KB);
//  * lm_head argmax GEMM (1.24 GB in one launch): everything, 11 stages -> 0.967 of the measured copy
//    bandwidth instead of 0.93.
#ifndef DFLASH_GEMM_SMEM_KB_PARTIALS
#define DFLASH_GEMM_SMEM_KB_PARTIALS 110
#endif
#ifndef DFLASH_GEMM_SMEM_KB_FUSED
#define DFLASH_GEMM_SMEM_KB_FUSED 100
#endif
#ifndef DFLASH_GEMM_SMEM_KB_ARGMAX
#define DFLASH_GEMM_SMEM_KB_ARGMAX 215
#endif
#ifndef DFLASH_GEMM_SMEM_KB_WIDE   // projection GEMMs with >= 64 activation rows per group (+ 16 KB epilogue tile)
#define DFLASH_GEMM_SMEM_KB_WIDE 196
#endif

constexpr bool mode_is_fused(int mode) { return mode == kModeSwiglu; }
constexpr bool mode_is_whole_tile(int mode) { return mode >= kModeArgmax && mode <= kModeSample; }

template <int MB, int MODE = 0>
struct GemmCfg {
  static constexpr bool kFused = mode_is_fused(MODE);
  static constexpr bool kWhole = mode_is_whole_tile(MODE);
  static constexpr int kWBytes = kTileN * kTileK * 2;   // 16 KB
  static constexpr int kXBytes = MB * kTileK * 2;
  static constexpr int kStageBytes = kWBytes + kXBytes;
  // wide activation tiles (batched engines) need the whole SM to keep >= 4 stages in flight
  static constexpr int kBudget =
      (kWhole ? DFLASH_GEMM_SMEM_KB_ARGMAX
              : (MB >= 64 ? DFLASH_GEMM_SMEM_KB_WIDE : (kFused ? DFLASH_GEMM_SMEM_KB_FUSED : DFLASH_GEMM_SMEM_KB_PARTIALS))) * 1024;
  static constexpr int kStages = kBudget / kStageBytes < 3 ? 3 : kBudget / kStageBytes;
  static constexpr int kTmemCols = (2 * MB < 32) ? 32 : 2 * MB;
  // epilogue warps: warp w drains TMEM lane quarter w % 4. The 256-wide argmax epilogue keeps one packed running
  // best per activation row in registers, so it splits the columns over two warp sets (128 registers each).
  static constexpr int kEpiWarps = (kWhole && MB >= 128) ? 8 : 4;
  static constexpr int kThreads = (kEpiWarps + 2) * 32;  // + TMA warp + MMA warp
  static constexpr int kColsPerThread = MB / (kEpiWarps / 4);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  // fused modes: activation rows per shared-memory epilogue tile (tile[m][128] fp32)
  static constexpr int kEpiRows = MB < 64 ? 16 : 32;
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

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// Activation tensor maps of the context-injection kernel when it reads the selected target hidden states IN PLACE
// (CtxNormEpi::direct): one 3-D map [R][block_size][hidden] per selected layer, box [MB / SL][SL][64] -- the box rows
// come out in the a_in row order (request r, slot j), slots past block_size are zero-filled by the TMA unit. The
// concatenation of extract_context_feature (model/utils.py:16-25) is then just "k-block kb belongs to tensor
// kb / (hidden / 64)": no feature matrix is ever materialised.
struct alignas(64) XMaps {
  CUtensorMap m[8];
};

template <int MB, int MODE>
__device__ __forceinline__ void gemm_skinny_body(const CUtensorMap& tmW, const CUtensorMap& tmX, const CUtensorMap* xsel,
                                                 const GemmArgs& a) {
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
  const long long T = static_cast<long long>(a.n_tiles) * a.k_blocks;
  const long long G = gridDim.y;           // weight ranges
  const int cta = blockIdx.y;              // this CTA's weight range
  const int m0 = blockIdx.x * MB;          // first activation row of this CTA's column group
  const int mv = a.m_valid - m0;           // valid rows in the group (may exceed MB)
  long long u0, u1;
  constexpr bool kArgmax = Cfg::kWhole;
  constexpr bool kFused = Cfg::kFused;
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
    if (MODE == kModeCtxNorm && a.cn.direct)
      for (int s = 0; s * a.cn.kb_per_sel < a.k_blocks; ++s) tma_prefetch_desc(xsel + s);
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
  // Let the next kernel in the stream start its own prologue / weight prefetch right away. (kModeCtxNorm: the next
  // kernel is a GEMM that would sit on the SMs, pipeline filled, for this kernel's whole duration -- measured 1 us
  // slower than releasing it when the last MMA has been issued.)
  if (MODE != kModeCtxNorm) pdl_trigger();

  if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // one group: every weight byte is read once -> evict first; several groups re-read it from L2
      const uint64_t polW = gridDim.x == 1 ? l2_policy_evict_first() : l2_policy_evict_normal();
      const uint64_t polX = l2_policy_evict_last();
      const long long n_units = u1 - u0;
      const int npre = n_units < S ? static_cast<int>(n_units) : S;
      // the weight tile of unit (tile, kb) into stage s. kModeSwiglu: 64 gate rows + 64 up rows of the same columns
      // (two 64-row boxes of the [gate; up] stack; the 128-byte swizzle is a function of the shared-memory address,
      // so the two halves form the same 128-row operand tile one box would)
      auto load_w = [&](int s, int tile, int kb) {
        if (MODE == kModeSwiglu) {
          tma_load_2d(sW + s * Cfg::kWBytes, &tmW, &full[s], kb * kTileK, a.w_row0 + tile * (kTileN / 2), polW);
          tma_load_2d(sW + s * Cfg::kWBytes + Cfg::kWBytes / 2, &tmW, &full[s], kb * kTileK,
                      a.w_row0 + a.sw.I + tile * (kTileN / 2), polW);
        } else {
          tma_load_2d(sW + s * Cfg::kWBytes, &tmW, &full[s], kb * kTileK, a.w_row0 + tile * kTileN, polW);
        }
      };
      // weight tiles first: they do not depend on the predecessor kernel
      for (int i = 0; i < npre; ++i) {
        const long long u = u0 + i;
        mbar_expect_tx(&full[i], Cfg::kStageBytes);
        load_w(i, static_cast<int>(u / a.k_blocks), static_cast<int>(u % a.k_blocks));
      }
      // the activation tile of k-block kb into stage s
      const bool direct = MODE == kModeCtxNorm && a.cn.direct != 0;
      auto load_x = [&](int s, int kb) {
        if (MODE == kModeCtxNorm && direct) {
          const int sel = kb / a.cn.kb_per_sel;
          tma_load_3d(sX + s * Cfg::kXBytes, xsel + sel, &full[s], (kb - sel * a.cn.kb_per_sel) * kTileK, 0,
                      m0 / a.cn.SL, polX);
        } else {
          tma_load_2d(sX + s * Cfg::kXBytes, &tmX, &full[s], kb * kTileK, a.x_row0 + m0, polX);
        }
      };
      // (direct context injection: the hidden states were complete before the verify kernel in front of this one
      // released it -- that kernel waits before it triggers -- so the main loop does not wait for the verify kernel)
      if (!direct) pdl_wait();
      if (tr) tr[2] = global_ns();
      DFL_TRACE_ANY(1);
      for (int i = 0; i < npre; ++i) load_x(i, static_cast<int>((u0 + i) % a.k_blocks));
      int stage = npre % S;
      uint32_t phase = (npre == S) ? 1u : 0u;
      for (long long u = u0 + npre; u < u1; ++u) {
        const int tile = static_cast<int>(u / a.k_blocks);
        const int kb = static_cast<int>(u % a.k_blocks);
        mbar_wait(&empty[stage], phase ^ 1u);
        mbar_expect_tx(&full[stage], Cfg::kStageBytes);
        load_w(stage, tile, kb);
        load_x(stage, kb);
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
        // (direct mode releases its dependents only after its own wait for the verify kernel, in the tail: kernels
        // further down the chain read request state before THEIR waits, on the strength of "the verify kernel is complete")
        if (MODE == kModeCtxNorm && !a.cn.direct && u >= u1) pdl_trigger();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0..kEpiWarps-1)
    if (!(MODE == kModeCtxNorm && a.cn.direct != 0)) pdl_wait();
    int acc = 0;
    uint32_t acc_phase = 0;
    const int quarter = warp & 3;                       // TMEM lanes [32 * quarter, +32)
    constexpr int kCols = Cfg::kColsPerThread;          // activation rows this thread drains per accumulator
    const int col0 = (warp >> 2) * kCols;               // second warp set: the upper half of the columns
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const int row_in_tile = quarter * 32 + lane;
    const int epi_tid = static_cast<int>(threadIdx.x);  // epilogue warps are warps [0, kEpiWarps)

    if constexpr (kFused) {
      // ---------------------------------------------------------------- fused row epilogues
      constexpr int kCh = Cfg::kEpiRows;
      constexpr int kRpw = kCh / 4;  // activation rows of a chunk per warp (row j = quarter + 4 * i)
      static_assert(MB % kCh == 0 && kCh % 16 == 0, "epilogue tile");
      __shared__ __align__(16) float s_t[kCh * kTileN];  // tile[m][n] of the current chunk of activation rows
      const long long part_stride = static_cast<long long>(kTileN) * MB;  // floats per CTA partial
      float* my_part = a.part + (static_cast<long long>(blockIdx.x) * G + cta) * part_stride;
      long long u = u0;
      while (u < u1) {
        const int tile = static_cast<int>(u / a.k_blocks);
        const long long seg_end =
            static_cast<long long>(tile + 1) * a.k_blocks < u1 ? static_cast<long long>(tile + 1) * a.k_blocks : u1;
        const int first = tile_first_cta(tile, a.k_blocks, T, G);
        const int nslots = tile_num_slots(tile, a.k_blocks, T, G);
        const int slot = cta - first;
        const bool writer = slot > 0;                  // the tile began in an earlier CTA: leave a partial for it
        const bool finisher = slot == 0 && nslots > 1; // the tile continues in later CTAs: they left partials for us
        unsigned int* flag = a.flags + static_cast<long long>(blockIdx.x) * a.n_tiles + tile;
        constexpr int kSlotBatch = 64 / kCh;  // other CTAs' partials in flight together (register budget)
        float4 p[kSlotBatch][kCh / 4];
        // partial layout: [m / 4][n][4] -> a warp's float4 accesses are 512 contiguous bytes
        auto part_ofs = [&](int c) { return (static_cast<long long>(c * (kCh / 4)) * kTileN + row_in_tile) * 4; };
        auto load_partials = [&](int c, int s0) {  // slots [s0, s0 + kSlotBatch) of chunk c
#pragma unroll
          for (int b = 0; b < kSlotBatch; ++b)
            if (s0 + b < nslots) {
              const float* op = my_part + static_cast<long long>(s0 + b) * part_stride + part_ofs(c);
#pragma unroll
              for (int i = 0; i < kCh / 4; ++i)
                p[b][i] = __ldcg(reinterpret_cast<const float4*>(op + static_cast<long long>(i) * kTileN * 4));
            }
        };
        if (finisher) {
          // The other CTAs of this tile computed their share at the START of their ranges (gate/up: a tile spans two
          // CTAs), so their partials are normally long there: wait for them and request them BEFORE the wait for
          // this CTA's own accumulator -- at the end of the kernel the last tile's epilogue is exposed.
          if (epi_tid == 0) {
            while (ld_acquire_u32(flag) < static_cast<unsigned int>(nslots - 1)) { }
            *flag = 0u;  // next use is a later launch of this plan
          }
          asm volatile("bar.sync 1, 128;\n" ::: "memory");
          load_partials(0, 1);
        }
        mbar_wait(&tfull[acc], acc_phase);
        if (tr && seg_end >= u1 && threadIdx.x == 0) tr[5] = global_ns();
        if (seg_end >= u1) DFL_TRACE(4);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < MB / kCh; ++c) {
          float v[kCh];
#pragma unroll
          for (int q = 0; q < kCh / 16; ++q)
            tmem_ld16(tmem_base + lane_addr + static_cast<uint32_t>(acc * MB + c * kCh + q * 16), v + q * 16);
          const long long pofs = part_ofs(c);
          tmem_ld_wait();
          if (c == MB / kCh - 1) {
            tc_fence_before();
            mbar_arrive(&tempty[acc]);  // the accumulator stage goes back to the MMA warp
          }
          if (writer) {
#pragma unroll
            for (int i = 0; i < kCh / 4; ++i)
              *reinterpret_cast<float4*>(my_part + pofs + static_cast<long long>(i) * kTileN * 4) =
                  make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            continue;
          }
          if (finisher) {
            for (int s0 = 1; s0 < nslots; s0 += kSlotBatch) {  // slot order: the summation order is fixed
              if (s0 > 1) load_partials(c, s0);
#pragma unroll
              for (int b = 0; b < kSlotBatch; ++b)
                if (s0 + b < nslots) {
#pragma unroll
                  for (int i = 0; i < kCh / 4; ++i) {
                    v[4 * i] += p[b][i].x; v[4 * i + 1] += p[b][i].y; v[4 * i + 2] += p[b][i].z; v[4 * i + 3] += p[b][i].w;
                  }
                }
            }
          }
#pragma unroll
          for (int j = 0; j < kCh; ++j) s_t[j * kTileN + row_in_tile] = v[j];
          asm volatile("bar.sync 1, 128;\n" ::: "memory");
#pragma unroll
          for (int i = 0; i < kRpw; ++i) {
            const int j = quarter + 4 * i;
            const int m = c * kCh + j;
            // row of the activation matrix == row of the output
            if (m < mv) swiglu_epi_apply(a.sw, &s_t[j * kTileN], tile, a.x_row0 + m0 + m, lane);
          }
          if (finisher && c + 1 < MB / kCh) load_partials(c + 1, 1);
          asm volatile("bar.sync 1, 128;\n" ::: "memory");
        }
        if (writer) {
          // publish: every thread's stores are ordered before the barrier, the release covers them (cumulativity)
          __threadfence();
          asm volatile("bar.sync 1, 128;\n" ::: "memory");
          if (epi_tid == 0) red_release_add_u32(flag, 1u);
        }
        u = seg_end;
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
      if (tr && threadIdx.x == 0) tr[6] = global_ns();
      DFL_TRACE(2);
    } else {
    // ------------------------------------------------------------------ partial planes / whole-tile reductions
    // kModeArgmax: one running best per activation row in a register, packed as
    //   (order key of the bf16-rounded logit) << 16 | (0xFFFF - tile)
    // so that keeping the best is ONE integer max per logit, and a tie keeps the lower tile = the lower vocab
    // index (this thread's weight row inside the tile is fixed). Decoded and reduced over threads at the end.
    constexpr bool kPlanes = MODE == kModePartials || MODE == kModeCtxNorm;  // fp32 partial planes in `ws`
    if constexpr (MODE == kModeCtxNorm) {
      // block rows: embedding gather + first input_layernorm, while the first accumulator is still being built
      // (direct mode: the block's first token is written by the verify kernel this launch overlaps -- done at the end)
      __shared__ float s_red[4];
      const int n_cta = static_cast<int>(gridDim.x * gridDim.y);
      if (!a.cn.direct)
        for (int row = static_cast<int>(blockIdx.y * gridDim.x + blockIdx.x); row < a.cn.n_blk_rows; row += n_cta)
          ctxnorm_embed_row<128, true>(a.cn, row, epi_tid, s_red);
    }
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
      if (kPlanes) {
        const int slot = cta - tile_first_cta(tile, a.k_blocks, T, G);
        dst = a.ws + (static_cast<long long>(slot) * a.ws_rows + m0) * a.ws_ld + n;
      }
      const uint32_t tile_tag = 0xFFFFu - static_cast<uint32_t>(tile);
      // kChunk columns per TMEM round trip: the loads of a chunk are all issued before the one wait
      constexpr bool kStageOut = kPlanes && (MB >= 64);
      constexpr int kChunk = kPlanes ? (kStageOut ? 32 : kCols)
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
        if (kPlanes) {
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
    if constexpr (MODE == kModeArgmax || MODE == kModeArgmaxDump || MODE == kModeSample) {
      // Drafted tokens: the last CTA of the grid to get here reduces everybody's candidates (one warp per block row;
      // ties -> lowest vocab index = torch.argmax on the bf16 logits, model/utils.py:27-29) and writes slots 1..bs-1
      // of block_ids (slot 0 is the committed token, model/dflash.py:247).
      if (a.tok_counter != nullptr) {
        __shared__ unsigned int s_last;
        __threadfence();
        asm volatile("bar.sync 1, %0;\n" ::"n"(Cfg::kEpiWarps * 32) : "memory");
        if (epi_tid == 0) {
          const unsigned int total = gridDim.x * gridDim.y;
          const unsigned int prev = atomicAdd(a.tok_counter, 1u);
          s_last = (prev == total - 1u) ? 1u : 0u;
          if (prev == total - 1u) *a.tok_counter = 0u;
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(Cfg::kEpiWarps * 32) : "memory");
        if (s_last) {
          __threadfence();
          const int n_cta = static_cast<int>(gridDim.y);
          for (int row = warp; row < a.tok_rows; row += Cfg::kEpiWarps) {
            float bv = -INFINITY;
            int bi = 0x7fffffff;
            // (all of a lane's candidates requested before the first compare: this runs after the last weight byte, on
            // the critical path in front of the target's forward -- one L2 round trip per row, not one per 32 CTAs)
            constexpr int kMaxPerLane = 8;  // up to 256 CTAs
            float cv[kMaxPerLane];
            int ci[kMaxPerLane];
#pragma unroll
            for (int k = 0; k < kMaxPerLane; ++k) {
              const int g = lane + 32 * k;
              cv[k] = -INFINITY;
              ci[k] = 0x7fffffff;
              if (g < n_cta) {
                cv[k] = __ldcg(a.cand_val + static_cast<long long>(g) * a.cand_ld + row);
                ci[k] = __ldcg(a.cand_idx + static_cast<long long>(g) * a.cand_ld + row);
              }
            }
#pragma unroll
            for (int k = 0; k < kMaxPerLane; ++k)
              if (cv[k] > bv || (cv[k] == bv && ci[k] < bi)) { bv = cv[k]; bi = ci[k]; }
            for (int g = lane + 32 * kMaxPerLane; g < n_cta; g += 32) {
              const float v = __ldcg(a.cand_val + static_cast<long long>(g) * a.cand_ld + row);
              const int i = __ldcg(a.cand_idx + static_cast<long long>(g) * a.cand_ld + row);
              if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
              const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
              if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) {
              a.tok_draft_tokens[row] = bi;
              const int r = row / a.tok_SL, i = row % a.tok_SL;
              if (i >= 1 && i < a.tok_bs) a.tok_block_ids[static_cast<long long>(r) * a.tok_bs + i] = bi;
            }
          }
        }
      }
    }
    }
  }

  if constexpr (MODE == kModeCtxNorm) {
    // ---------------------------------------------------------------- in-kernel row pass over the context rows
    // Every thread of the CTA (the TMA / MMA warps are done too). This CTA's partial planes are stored; count it in,
    // and if it owns live context rows wait until every CTA of the launch is in (they are all co-resident: at most
    // one CTA per SM's worth of them exists), then finish those rows. The LAST CTA to leave resets both counters.
    __shared__ int s_ns[64];       // partial slots per column tile (hidden <= 8192)
    __shared__ float s_red6[Cfg::kThreads / 32];
    __shared__ int s_live;
    __shared__ __align__(8) uint64_t s_rowbar;
    const int n_cta = static_cast<int>(gridDim.x * gridDim.y);
    const int me = static_cast<int>(blockIdx.y * gridDim.x + blockIdx.x);
    const int tid = static_cast<int>(threadIdx.x);
    // the idle lanes of the TMA / MMA warps get here at once: they fill the slot table while the main loop runs
    if (warp >= Cfg::kEpiWarps && lane != 0) {
      const int helper = (warp - Cfg::kEpiWarps) * 31 + lane - 1;
      for (int t = helper; t < a.n_tiles; t += 62) s_ns[t] = tile_num_slots(t, a.k_blocks, T, G);
      if (helper == 0) {
        mbar_init(&s_rowbar, 1);
        mbar_fence_init();
      }
    }
    // direct mode: everything from here on reads what the verify kernel wrote (accepted lengths, the next block).
    // The idle lanes must not sit in griddepcontrol.wait while lane 0 of their warp is still issuing loads / MMAs (a
    // blocked divergent path starves the other one -- measured: not one stage landed before the verify kernel had
    // finished), so the warp reconverges first.
    __syncwarp();
    if (a.cn.direct) {
      pdl_wait();
      pdl_trigger();
    }
    if (tid == 0) {
      int live = 0;
      for (int row = me; row < a.m_valid; row += n_cta)
        live |= (row % a.cn.SL < __ldg(a.cn.ctx_len + row / a.cn.SL)) ? 1 : 0;
      s_live = live;
    }
    __threadfence();
    __syncthreads();
    unsigned int left = 0;
    if (tid == 0) {
      red_release_add_u32(&a.cn.sync[0], 1u);
      if (s_live)
        while (ld_acquire_u32(&a.cn.sync[0]) < static_cast<unsigned int>(n_cta)) { }
      // past the wait (or never waiting): count out now, so that the round trip overlaps the row pass
      left = atomicAdd(&a.cn.sync[1], 1u);
    }
    __syncthreads();
    DFL_TRACE(5);
    if (s_live) {
      const long long slot_stride = static_cast<long long>(a.ws_rows) * a.ws_ld;
      float* stage = reinterpret_cast<float*>(smem);
      const bool bulk = static_cast<long long>(a.cn.max_slots) * a.cn.ld * 4 <= static_cast<long long>(S) * Cfg::kStageBytes;
      uint32_t parity = 0;
      for (int row = me; row < a.m_valid; row += n_cta)
        if (row % a.cn.SL < __ldg(a.cn.ctx_len + row / a.cn.SL)) {
          if (bulk) {
            ctxnorm_row_pass_bulk<Cfg::kThreads>(a.cn, a.ws, slot_stride, a.ws_ld, row, tid, s_ns, stage, &s_rowbar,
                                                 parity, s_red6);
            parity ^= 1u;
          } else {
            ctxnorm_row_pass<Cfg::kThreads>(a.cn, a.ws, slot_stride, a.ws_ld, row, tid, s_ns, stage, s_red6);
          }
        }
    }
    // the LAST CTA past the wait resets both counters for the next launch (nobody can still be polling them)
    if (tid == 0 && left == static_cast<unsigned int>(n_cta) - 1u) {
      a.cn.sync[0] = 0u;
      a.cn.sync[1] = 0u;
    }
    if (a.cn.direct) {
      // block rows (embedding gather + first input_layernorm), dealt from the LAST CTA downwards: the context rows
      // above were dealt from the first CTA upwards, so at small batches no CTA gets both
      __syncthreads();  // (s_red6 of the row pass has been consumed)
      for (int row = n_cta - 1 - me; row < a.cn.n_blk_rows; row += n_cta)
        ctxnorm_embed_row<Cfg::kThreads, false>(a.cn, row, tid, s_red6);
    }
    DFL_TRACE(2);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

template <int MB, int MODE>
__global__ void __launch_bounds__(GemmCfg<MB, MODE>::kThreads, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                   const __grid_constant__ GemmArgs a) {
  gemm_skinny_body<MB, MODE>(tmW, tmX, nullptr, a);
}

// The context-injection kernel (kModeCtxNorm) with the per-layer activation maps of its direct mode
// (register cap: this kernel shares its SMs with the verify kernel it overlaps -- per SM sub-partition, four warps of
// that kernel at 64 registers plus two of this one at 112 fit the 16 K registers)
template <int MB>
__global__ void __maxnreg__(112)
gemm_ctx_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ XMaps xm, const __grid_constant__ GemmArgs a) {
  gemm_skinny_body<MB, kModeCtxNorm>(tmW, tmX, xm.m, a);
}

}  // namespace dfl
