// Persistent draft-step kernel: the whole draft step (embed -> ctx inject -> L layers -> lm_head + argmax) in
// ONE launch of one CTA per SM.
//
// Why: with one kernel per operator the step is 55 launches and ~36 of them are tiny; the timing ablation in
// DESIGN.md §8 shows ~156 us of a 743 us step is the serial latency of kernel boundaries and ~70 us is GEMM
// ramp-up/drain -- HBM idles although the weights of the next GEMM are known. Here the phases of the step are
// separated by grid barriers (one 64-bit counter per phase, monotonic across steps), and the roles of the GEMM
// pipeline live for the whole step:
//   warp 4  W producer : streams the weight tiles of ALL GEMMs of the step, in order, into an 8-stage smem ring;
//                        it never waits for a grid barrier (weights do not depend on activations), so HBM keeps
//                        streaming through every small phase and barrier (ring = 148 x 128 KB = 19 MB ahead)
//   warp 6  X producer : loads the activation tile of each unit once the phase that produces it has completed
//                        on every CTA (grid barrier), into the same stage (same mbarrier transaction count)
//   warp 5  MMA issuer : tcgen05.mma, accumulators double-buffered in TMEM across tiles AND across GEMMs
//   warps 0-3 workers  : GEMM epilogues (TMEM -> fp32 partials / fused argmax) and, between GEMMs, their share
//                        of the small phases (the same __device__ bodies the stand-alone kernels run)
// Cross-CTA data inside the kernel goes through L2 (ld.global.cg / TMA), with fence.proxy.async on both sides
// of a barrier where generic-proxy stores feed TMA loads.
#pragma once
#include "attention.cuh"
#include "fused_ops.cuh"

namespace dfl {

constexpr int kMegaThreads = 256;
constexpr int kMegaWorkers = 128;
constexpr int kMegaStages = 8;
constexpr int kMegaWBytes = kTileN * kTileK * 2;  // 16 KB
constexpr int kMegaXBytes = 4096;                 // up to 32 activation rows per k-block
constexpr int kMegaStageBytes = kMegaWBytes + kMegaXBytes;
constexpr int kMegaScratch = 2 * kAttnTileBytes;  // 32 KB: attention K/V tile, or the finalize row buffer
constexpr int kMegaTmemCols = 64;                 // 2 accumulators x 32 columns
constexpr int kMegaSmemBytes = kMegaStages * kMegaStageBytes + kMegaScratch + 1024 /*align*/ + 1024 /*barriers etc*/;
constexpr long long kMegaSpinLimit = 1ll << 27;   // ~seconds; then trap instead of hanging the GPU

enum MegaPhaseKind : int { kPhGemm = 0, kPhRows, kPhQkvPost, kPhAttn, kPhCombine, kPhSwiglu, kPhTokens };

struct alignas(128) MegaGemm {
  CUtensorMap tmW;
  CUtensorMap tmX;
  GemmArgs args;
  int mb, mode, grid, pad;
};

struct alignas(16) MegaPhase {
  int kind;
  int gemm;     // index into the GEMM table (kPhGemm)
  int n_items;  // rows / warp items / float4 items of the small phase
  int dep;      // index of the phase whose grid barrier must complete first (-1: inputs ready at kernel start)
  union {
    RowsArgs rows;
    QkvPostArgs qkv;
    AttnArgs attn;
    SwigluArgs sw;
    DraftTokArgs tok;
  } u;
};

// Register copy of the per-GEMM fields a pipeline role needs (the table itself is in global memory).
struct MegaGemmRegs {
  const CUtensorMap* tmW;
  const CUtensorMap* tmX;
  int n_tiles, k_blocks, w_row0, x_row0, mb, mode, grid;
};
__device__ __forceinline__ MegaGemmRegs mega_load_gemm(const MegaGemm* g) {
  MegaGemmRegs r;
  r.tmW = &g->tmW; r.tmX = &g->tmX;
  r.n_tiles = g->args.n_tiles; r.k_blocks = g->args.k_blocks; r.w_row0 = g->args.w_row0; r.x_row0 = g->args.x_row0;
  r.mb = g->mb; r.mode = g->mode; r.grid = g->grid;
  return r;
}

struct MegaArgs {
  const MegaGemm* gemms;
  const MegaPhase* phases;
  int n_phases;
  unsigned long long* bars;   // [n_phases] monotonic arrival counters
  unsigned long long* epoch;  // steps completed so far
  int* err;
  unsigned long long* trace;  // optional [n_phases + 1] globaltimer stamps of CTA 0 (phase ends), debug
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
  return t;
}

__device__ __forceinline__ void mega_fail(int* err, int code) {
  *err = code;
  __threadfence_system();
  __trap();
}
__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, uint32_t parity, int* err, int code) {
  uint32_t done = 0;
  long long spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
    if (++spins > kMegaSpinLimit) mega_fail(err, code);
  }
}
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

// one thread: publish this CTA's completion of a phase
__device__ __forceinline__ void grid_arrive(unsigned long long* ctr) {
  __threadfence();
  atomicAdd(ctr, 1ull);
}
// one thread: wait until every CTA has completed the phase (in this step)
__device__ __forceinline__ void grid_wait(const unsigned long long* ctr, unsigned long long target, int* err, int code) {
  long long spins = 0;
  while (true) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(ctr) : "memory");  // no L1 invalidate per poll
    if (v >= target) {
      asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
      return;
    }
    if (++spins > kMegaSpinLimit) mega_fail(err, code);
    __nanosleep(32);
  }
}

__device__ __forceinline__ void mega_unit_range(const MegaGemmRegs& g, int cta, long long& u0, long long& u1) {
  if (cta >= g.grid) { u0 = u1 = 0; return; }
  if (g.mode == kModeArgmax) {
    u0 = (cta * static_cast<long long>(g.n_tiles) / g.grid) * g.k_blocks;
    u1 = ((cta + 1) * static_cast<long long>(g.n_tiles) / g.grid) * g.k_blocks;
  } else {
    const long long T = static_cast<long long>(g.n_tiles) * g.k_blocks;
    u0 = unit_begin(cta, T, g.grid);
    u1 = unit_begin(cta + 1, T, g.grid);
  }
}

// MBA = activation rows of the fused-argmax GEMM (register budget of its epilogue)
template <int MBA>
__global__ void __launch_bounds__(kMegaThreads, 1) draft_step_mega_kernel(const MegaArgs m) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* ring = smem;
  uint8_t* scratch = smem + kMegaStages * kMegaStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(scratch + kMegaScratch);
  uint64_t* empty = full + kMegaStages;
  uint64_t* tfull = empty + kMegaStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* red = reinterpret_cast<float*>(tmem_slot + 4);      // 8 floats
  int* ns_tab = reinterpret_cast<int*>(red + 8);              // kRowsMaxTiles ints

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const unsigned long long target = (*m.epoch + 1ull) * static_cast<unsigned long long>(G);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMegaStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], kMegaWorkers); }
    mbar_fence_init();
  }
  if (warp == 5) {
    tmem_alloc<kMegaTmemCols>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ================================================================== W producer (never waits on the grid)
    if (lane == 0) {
      const uint64_t polW = l2_policy_evict_first();
      int stage = 0;
      uint32_t phase = 0;
      for (int p = 0; p < m.n_phases; ++p) {
        if (m.phases[p].kind != kPhGemm) continue;
        const MegaGemmRegs g = mega_load_gemm(&m.gemms[m.phases[p].gemm]);
        long long u0, u1;
        mega_unit_range(g, cta, u0, u1);
        const uint32_t tx = kMegaWBytes + g.mb * kTileK * 2;
        for (long long u = u0; u < u1; ++u) {
          const int tile = static_cast<int>(u / g.k_blocks), kb = static_cast<int>(u % g.k_blocks);
          mbar_wait_guard(&empty[stage], phase ^ 1u, m.err, 100 + p);
          mbar_expect_tx(&full[stage], tx);
          tma_load_2d(ring + stage * kMegaStageBytes, g.tmW, &full[stage], kb * kTileK, g.w_row0 + tile * kTileN, polW);
          if (++stage == kMegaStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 6) {
    // ================================================================== X producer (follows the grid barriers)
    if (lane == 0) {
      const uint64_t polX = l2_policy_evict_last();
      int stage = 0;
      uint32_t phase = 0;
      for (int p = 0; p < m.n_phases; ++p) {
        if (m.phases[p].kind != kPhGemm) continue;
        const MegaGemmRegs g = mega_load_gemm(&m.gemms[m.phases[p].gemm]);
        long long u0, u1;
        mega_unit_range(g, cta, u0, u1);
        const int dep = m.phases[p].dep;
        if (u1 > u0 && dep >= 0) {
          grid_wait(&m.bars[dep], target, m.err, 200 + p);  // the phase that wrote these activations is done
          proxy_fence_async();
        }
        for (long long u = u0; u < u1; ++u) {
          const int kb = static_cast<int>(u % g.k_blocks);
          mbar_wait_guard(&empty[stage], phase ^ 1u, m.err, 300 + p);
          tma_load_2d(ring + stage * kMegaStageBytes + kMegaWBytes, g.tmX, &full[stage], kb * kTileK, g.x_row0, polX);
          if (++stage == kMegaStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 5) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int p = 0; p < m.n_phases; ++p) {
        if (m.phases[p].kind != kPhGemm) continue;
        const MegaGemmRegs g = mega_load_gemm(&m.gemms[m.phases[p].gemm]);
        long long u0, u1;
        mega_unit_range(g, cta, u0, u1);
        const uint32_t idesc = umma_idesc_bf16(kTileN, g.mb);
        long long u = u0;
        while (u < u1) {
          const long long tile = u / g.k_blocks;
          const long long seg_end = (tile + 1) * g.k_blocks < u1 ? (tile + 1) * g.k_blocks : u1;
          mbar_wait_guard(&tempty[acc], acc_phase ^ 1u, m.err, 400 + p);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 32);
          uint32_t accumulate = 0;
          for (; u < seg_end; ++u) {
            mbar_wait_guard(&full[stage], phase, m.err, 500 + p);
            tc_fence_after();
            const uint64_t da = umma_desc_sw128(smem_u32(ring + stage * kMegaStageBytes));
            const uint64_t db = umma_desc_sw128(smem_u32(ring + stage * kMegaStageBytes + kMegaWBytes));
#pragma unroll
            for (int k = 0; k < kTileK / kUmmaK; ++k) {
              umma_bf16_ss(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                           accumulate);
              accumulate = 1;
            }
            umma_commit(&empty[stage]);
            if (++stage == kMegaStages) { stage = 0; phase ^= 1u; }
          }
          umma_commit(&tfull[acc]);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      }
    }
  } else if (warp < 4) {
    // ================================================================== workers
    const int wtid = threadIdx.x;  // 0..127
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    if (m.trace != nullptr && cta == 0 && wtid == 0) m.trace[m.n_phases] = global_timer_ns();
    for (int p = 0; p < m.n_phases; ++p) {
      const MegaPhase& ph = m.phases[p];
      const int ph_kind = ph.kind, ph_items = ph.n_items, ph_dep = ph.dep;
      if (ph_kind == kPhGemm) {
        // ---------------------------------------------------------------- GEMM epilogue
        const MegaGemmRegs g = mega_load_gemm(&m.gemms[ph.gemm]);
        const GemmArgs a = m.gemms[ph.gemm].args;  // by value: registers, not global reloads
        long long u0, u1;
        mega_unit_range(g, cta, u0, u1);
        const long long T = static_cast<long long>(a.n_tiles) * a.k_blocks;
        float best_v[MBA];
        int best_i[MBA];
#pragma unroll
        for (int j = 0; j < MBA; ++j) { best_v[j] = -INFINITY; best_i[j] = 0x7fffffff; }
        long long u = u0;
        while (u < u1) {
          const int tile = static_cast<int>(u / a.k_blocks);
          const long long seg_end =
              static_cast<long long>(tile + 1) * a.k_blocks < u1 ? static_cast<long long>(tile + 1) * a.k_blocks : u1;
          const int n = tile * kTileN + warp * 32 + lane;
          mbar_wait_guard(&tfull[acc], acc_phase, m.err, 600 + p);
          tc_fence_after();
          float* dst = nullptr;
          if (g.mode == kModePartials) {
            const int slot = cta - tile_first_cta(tile, a.k_blocks, T, g.grid);
            dst = a.ws + (static_cast<long long>(slot) * a.ws_rows) * a.ws_ld + n;
          }
          const int nchunks = g.mb / 16;
          for (int c = 0; c < nchunks; ++c) {
            float v[16];
            tmem_ld16(tmem_base + lane_addr + static_cast<uint32_t>(acc * 32 + c * 16), v);
            tmem_ld_wait();
            if (c == nchunks - 1) {
              tc_fence_before();
              mbar_arrive(&tempty[acc]);
            }
            if (n < a.N) {
              if (g.mode == kModePartials) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int mm = c * 16 + j;
                  if (mm < a.m_valid) dst[static_cast<long long>(mm) * a.ws_ld] = v[j];
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int mm = c * 16 + j;
                  const float r = bf16_round(v[j]);
                  if (mm < MBA && r > best_v[mm < MBA ? mm : 0]) { best_v[mm < MBA ? mm : 0] = r; best_i[mm < MBA ? mm : 0] = n; }
                  if (a.logits != nullptr && mm < a.m_valid)
                    a.logits[static_cast<long long>(mm) * a.logits_ld + n] = __float2bfloat16_rn(v[j]);
                }
              }
            }
          }
          u = seg_end;
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
        if (g.mode == kModeArgmax) {
          float* red_v = reinterpret_cast<float*>(scratch);
          int* red_i = reinterpret_cast<int*>(scratch + 4 * MBA * sizeof(float));
#pragma unroll
          for (int j = 0; j < MBA; ++j) {
            float bv = best_v[j];
            int bi = best_i[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
              const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
              if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { red_v[warp * MBA + j] = bv; red_i[warp * MBA + j] = bi; }
          }
          group_sync(1, kMegaWorkers);
          if (wtid < MBA && cta < g.grid) {
            float bv = red_v[wtid];
            int bi = red_i[wtid];
            for (int w = 1; w < 4; ++w) {
              const float ov = red_v[w * MBA + wtid];
              const int oi = red_i[w * MBA + wtid];
              if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            a.cand_val[static_cast<long long>(cta) * g.mb + wtid] = bv;
            a.cand_idx[static_cast<long long>(cta) * g.mb + wtid] = bi;
          }
        }
      } else {
        // ---------------------------------------------------------------- small phase
        if (ph_dep >= 0) {
          if (wtid == 0) grid_wait(&m.bars[ph_dep], target, m.err, 700 + p);
          group_sync(1, kMegaWorkers);
        }
        if (m.trace != nullptr && cta == 0 && wtid == 0) m.trace[m.n_phases + 1 + p] = global_timer_ns();
        switch (ph_kind) {
          case kPhRows: {
            const RowsArgs ra = ph.u.rows;
            for (int row = cta; row < ph_items; row += G)
              finalize_row_body<kMegaWorkers>(ra, row, wtid, reinterpret_cast<float*>(scratch), red, ns_tab, 1);
            break;
          }
          case kPhQkvPost: {
            const QkvPostArgs qa = ph.u.qkv;
            for (int it = cta * 4 + warp; it < ph_items; it += G * 4) qkv_post_item(qa, it, lane);
            break;
          }
          case kPhAttn: {
            const AttnArgs aa = ph.u.attn;
            const int group = aa.Hq / aa.Hkv;
            const int tiles_per_req = aa.SL / 16;
            for (int it = cta; it < ph_items; it += G) {
              const int split = it % aa.nsplit;
              const int h = (it / aa.nsplit) % aa.Hkv;
              const int z = it / (aa.nsplit * aa.Hkv);
              if (warp < group)
                attn_split_body<1>(aa, z / tiles_per_req, z % tiles_per_req, h, split, wtid, 32 * group,
                                   smem_u32(scratch), 2);
            }
            break;
          }
          case kPhCombine: {
            const AttnArgs aa = ph.u.attn;
            for (int it = cta * 4 + warp; it < ph_items; it += G * 4) attn_combine_item(aa, it, lane);
            break;
          }
          case kPhSwiglu: {
            const SwigluArgs sa = ph.u.sw;
            const int stride = G * kMegaWorkers;
            for (int it = cta * kMegaWorkers + wtid; it < ph_items; it += 3 * stride)
              swiglu_items3(sa, it, stride, ph_items);
            break;
          }
          case kPhTokens: {
            const DraftTokArgs ta = ph.u.tok;
            for (int row = cta * 4 + warp; row < ph_items; row += G * 4) draft_tokens_row(ta, row, lane);
            break;
          }
          default: break;
        }
      }
      // publish: all workers' stores are ordered before the elected thread's release (bar.sync + fence)
      group_sync(1, kMegaWorkers);
      if (m.trace != nullptr && cta == 0 && wtid == 0) m.trace[2 * m.n_phases + 1 + p] = global_timer_ns();
      if (wtid == 0) grid_arrive(&m.bars[p]);
      if (m.trace != nullptr && cta == 0 && wtid == 0) m.trace[p] = global_timer_ns();
    }
    // last CTA-0 worker closes the step: everyone has read the old epoch long ago
    if (cta == 0 && wtid == 0) {
      grid_wait(&m.bars[m.n_phases - 1], target, m.err, 900);
      *m.epoch = *m.epoch + 1ull;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tmem_dealloc<kMegaTmemCols>(tmem_base);
  }
}

}  // namespace dfl
