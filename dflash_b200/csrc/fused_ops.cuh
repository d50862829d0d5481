// Small fused kernels around the skinny GEMMs: they consume the fp32 split-K partials of fc / qkv / o / down, apply
// the reference's bf16 rounding points (Linear output (+ bias) -> bf16, RMSNorm in fp32 -> bf16 -> * weight,
// residual add in bf16, RoPE in bf16) and produce the next GEMM's activations. (SwiGLU runs in the gate/up GEMM's
// own epilogue, epilogues.cuh; the drafted tokens are reduced by the lm_head GEMM's last CTA.)
//
// All of them are latency-bound (a few hundred KB each): 128-bit accesses, rolled loops (small code:
// the first version unrolled 32x with 64-bit divisions inline and spent ~45 us in instruction fetch),
// slot counts per 128-column tile computed once with 32-bit arithmetic.
#pragma once
#include "gemm_host.cuh"

namespace dfl {

// How a consumer finds the partial slots of output column n (mirrors the GEMM's stream-K split).
struct SlotMap {
  int k_blocks;
  uint32_t T;   // n_tiles * k_blocks   (T * G < 2^31 is checked on the host)
  uint32_t G;   // CTAs of the producing GEMM
  int ws_rows;
  long long ws_ld;
};

inline SlotMap slot_map_of(const GemmPlan& p) {
  SlotMap s;
  s.k_blocks = p.args.k_blocks;
  s.T = static_cast<uint32_t>(p.args.n_tiles) * static_cast<uint32_t>(p.args.k_blocks);
  s.G = static_cast<uint32_t>(p.grid);
  s.ws_rows = p.args.ws_rows;
  s.ws_ld = p.args.ws_ld;
  return s;
}

// 32-bit forms of cta_of_unit / tile_num_slots (gemm_skinny.cuh)
__device__ __forceinline__ uint32_t cta_of_unit32(uint32_t x, uint32_t T, uint32_t G) {
  return ((x + 1u) * G - 1u) / T;
}
__device__ __forceinline__ int tile_slots32(int t, const SlotMap& sm) {
  const uint32_t kb = static_cast<uint32_t>(sm.k_blocks);
  return static_cast<int>(cta_of_unit32((t + 1) * kb - 1u, sm.T, sm.G) - cta_of_unit32(t * kb, sm.T, sm.G)) + 1;
}

// fp32 sum over the slots of (row m, column n), in slot order.
__device__ __forceinline__ float sum_slots_n(const float* __restrict__ ws, const SlotMap& sm, int m, int n, int ns) {
  const float* p = ws + static_cast<long long>(m) * sm.ws_ld + n;
  const long long slot_stride = static_cast<long long>(sm.ws_rows) * sm.ws_ld;
  float acc = __ldcg(p);
  for (int s = 1; s < ns; ++s) acc += __ldcg(p + s * slot_stride);
  return acc;
}
__device__ __forceinline__ float4 sum_slots_4(const float* __restrict__ ws, const SlotMap& sm, int m, int n, int ns) {
  const float* p = ws + static_cast<long long>(m) * sm.ws_ld + n;
  const long long slot_stride = static_cast<long long>(sm.ws_rows) * sm.ws_ld;
  // issue every slot's load before the first add (these are L2 round trips; the adds stay in slot order)
  float4 v[8];
#pragma unroll
  for (int s = 0; s < 8; ++s)
    if (s < ns) v[s] = __ldcg(reinterpret_cast<const float4*>(p + s * slot_stride));
  float4 acc = v[0];
#pragma unroll
  for (int s = 1; s < 8; ++s)
    if (s < ns) { acc.x += v[s].x; acc.y += v[s].y; acc.z += v[s].z; acc.w += v[s].w; }
  for (int s = 8; s < ns; ++s) {
    const float4 w = __ldcg(reinterpret_cast<const float4*>(p + s * slot_stride));
    acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------
// test helper: out[m][n] = sum of partial slots (fp32)
__global__ void sum_slots_kernel(const float* __restrict__ ws, SlotMap sm, int rows, int N, float* out,
                                 long long out_ld) {
  pdl_wait();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n < N && m < rows)
    out[static_cast<long long>(m) * out_ld + n] = sum_slots_n(ws, sm, m, n, tile_slots32(n / kTileN, sm));
}

inline cudaError_t launch_sum_slots(const GemmPlan& p, float* out, long long out_ld, cudaStream_t st) {
  dim3 grid((p.args.N + 255) / 256, p.args.m_valid);
  sum_slots_kernel<<<grid, 256, 0, st>>>(p.args.ws, slot_map_of(p), p.args.m_valid, p.args.N, out, out_ld);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// lm_head second stage: per activation row, max over the per-CTA candidates (ties -> lowest index,
// matching torch.argmax on the bf16 logits: model/utils.py:27-29).
__device__ __forceinline__ void argmax_candidates(const float* __restrict__ cand_val, const int* __restrict__ cand_idx,
                                                  int n_cta, int mb, int row, float& bv, int& bi) {
  bv = -INFINITY;
  bi = 0x7fffffff;
  for (int g = threadIdx.x; g < n_cta; g += 32) {
    const float v = __ldcg(cand_val + static_cast<long long>(g) * mb + row);
    const int i = __ldcg(cand_idx + static_cast<long long>(g) * mb + row);
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
}

__global__ void __launch_bounds__(32) reduce_candidates_kernel(const float* __restrict__ cand_val,
                                                               const int* __restrict__ cand_idx, int n_cta, int mb,
                                                               int rows, long long* __restrict__ tokens) {
  pdl_wait();
  const int row = blockIdx.x;
  if (row >= rows) return;
  float bv;
  int bi;
  argmax_candidates(cand_val, cand_idx, n_cta, mb, row, bv, bi);
  if (threadIdx.x == 0) tokens[row] = bi;
}

inline cudaError_t launch_reduce_candidates(const float* cand_val, const int* cand_idx, int n_cta, int mb,
                                            int rows, long long* tokens, cudaStream_t st) {
  reduce_candidates_kernel<<<rows, 32, 0, st>>>(cand_val, cand_idx, n_cta, mb, rows, tokens);
  return cudaGetLastError();
}

// =============================================================================================
// Row layout shared by all step kernels. R requests, SL row slots per request (bs <= SL).
//   activation matrix a_in [2*R*SL, H]: rows [0, R*SL) = context rows (request r, slot j valid iff
//   j < ctx_len[r]); rows [R*SL, 2*R*SL) = block rows (request r, slot i valid iff i < bs).
// Absolute positions: context row j -> start[r] - ctx_len[r] + j ; block row i -> start[r] + i
// (model/dflash.py:241: position_ids[:, cache_len : start + block_size]).
// =============================================================================================

enum RowValid : int { kRowsAll = 0, kRowsCtx = 1 };

struct RowsArgs {
  // source A: split-K partials of the producing GEMM (row m of ws = row m here)
  const float* ws;
  SlotMap sm;
  const __nv_bfloat16* bias;  // optional Linear bias [H] added to the fp32 sum (config.attention_bias: o_proj)
  // source B (if embed != null): embedding gather, token = ids[(row / SL) * ids_ld + row % SL]
  const __nv_bfloat16* embed;
  const long long* ids;
  int ids_ld;
  long long pad_token;  // used for slots >= bs
  int bs;
  int H;
  int SL;
  int valid_mode;
  const int* ctx_len;
  __nv_bfloat16* resid;   // optional residual stream [rows, H] (in/out), or written (embed mode)
  const __nv_bfloat16* norm_w;  // optional RMSNorm weight [H]
  __nv_bfloat16* out;     // [rows, H]
  float eps;
};

#ifndef DFLASH_ROWS_THREADS
#define DFLASH_ROWS_THREADS 512
#endif
constexpr int kRowsThreads = DFLASH_ROWS_THREADS;
constexpr int kRowsMaxTiles = 64;  // H <= 8192

// One CTA per row:  v = bf16(sum of partials)           (nn.Linear output dtype)
//                   v = bf16(resid + v); resid = v       (residual add, model/dflash.py:140,144)
//                   out = w * bf16(v * rsqrt(mean(v^2) + eps))   (Qwen3RMSNorm, fp32 inside)
// rowbuf: 6*H bytes of shared memory (the row as H floats + H bf16 norm weights); red: NT/32 floats; ns_tab:
// kRowsMaxTiles ints. NT threads (tid in [0, NT)) work on one row and meet at named barrier bar_id.
// NS_READY: the caller has already filled ns_tab (and synchronised) -- the stand-alone kernel does that before
// griddepcontrol.wait, since the slot counts do not depend on the producing GEMM's data.
template <int NT, bool NS_READY = false>
__device__ __forceinline__ void finalize_row_body(const RowsArgs& a, int row, int tid, float* rowbuf, float* red,
                                                  int* ns_tab, int bar_id) {
  if (a.valid_mode == kRowsCtx) {
    const int r = row / a.SL, j = row % a.SL;
    if (j >= a.ctx_len[r]) return;
  }
  const long long roff = static_cast<long long>(row) * a.H;
  __nv_bfloat16* wbuf = reinterpret_cast<__nv_bfloat16*>(rowbuf + a.H);  // H bf16 norm weights behind the H floats
  const bool from_embed = a.embed != nullptr;
  long long tok = 0;
  if (from_embed) {
    const int r = row / a.SL, i = row % a.SL;
    if (a.ids == nullptr) tok = row;  // `embed` already is a [rows, H] embedding matrix (forward(noise_embedding=))
    else tok = (i < a.bs) ? a.ids[static_cast<long long>(r) * a.ids_ld + i] : a.pad_token;
  } else if (!NS_READY) {
    const int nt = (a.H + kTileN - 1) / kTileN;
    for (int t = tid; t < nt; t += NT) ns_tab[t] = tile_slots32(t, a.sm);
    group_sync(bar_id, NT);
  }
  float ss = 0.f;
  constexpr int U = 4;  // iterations whose loads are issued together (latency-bound: ~1 us per L2 round trip)
  for (int n0 = tid * 4; n0 < a.H; n0 += NT * 4 * U) {
    float4 x[U];
    uint2 rs[U], wv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int n = n0 + u * NT * 4;
      if (n < a.H) {
        if (a.norm_w != nullptr) wv[u] = *reinterpret_cast<const uint2*>(a.norm_w + n);  // for pass 2, fetched now
        if (from_embed) {
          rs[u] = *reinterpret_cast<const uint2*>(a.embed + tok * a.H + n);
        } else {
          x[u] = sum_slots_4(a.ws, a.sm, row, n, ns_tab[n / kTileN]);
          if (a.bias != nullptr) {
            const float4 b = unpack4_bf16(*reinterpret_cast<const uint2*>(a.bias + n));
            x[u].x += b.x; x[u].y += b.y; x[u].z += b.z; x[u].w += b.w;
          }
          if (a.resid != nullptr) rs[u] = __ldcg(reinterpret_cast<const uint2*>(a.resid + roff + n));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int n = n0 + u * NT * 4;
      if (n >= a.H) continue;
      float4 v;
      if (from_embed) {
        v = unpack4_bf16(rs[u]);
        if (a.resid != nullptr) *reinterpret_cast<uint2*>(a.resid + roff + n) = rs[u];
      } else {
        v = x[u];
        v.x = bf16_round(v.x); v.y = bf16_round(v.y); v.z = bf16_round(v.z); v.w = bf16_round(v.w);
        if (a.resid != nullptr) {
          const float4 rsd = unpack4_bf16(rs[u]);
          v.x = bf16_round(rsd.x + v.x); v.y = bf16_round(rsd.y + v.y);
          v.z = bf16_round(rsd.z + v.z); v.w = bf16_round(rsd.w + v.w);
          *reinterpret_cast<uint2*>(a.resid + roff + n) = pack4_bf16(v.x, v.y, v.z, v.w);
        }
      }
      if (a.norm_w == nullptr) {
        *reinterpret_cast<uint2*>(a.out + roff + n) = pack4_bf16(v.x, v.y, v.z, v.w);
      } else {
        *reinterpret_cast<float4*>(rowbuf + n) = v;
        *reinterpret_cast<uint2*>(wbuf + n) = wv[u];
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
    }
  }
  if (a.norm_w == nullptr) return;
  DFL_TRACE(4);
  ss = warp_sum(ss);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  group_sync(bar_id, NT);
  DFL_TRACE(5);
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) tot += red[w];
  const float rstd = 1.0f / sqrtf(tot / static_cast<float>(a.H) + a.eps);
  for (int n = tid * 4; n < a.H; n += NT * 4) {
    const float4 x = *reinterpret_cast<const float4*>(rowbuf + n);
    const float4 w = unpack4_bf16(*reinterpret_cast<const uint2*>(wbuf + n));
    *reinterpret_cast<uint2*>(a.out + roff + n) =
        pack4_bf16(w.x * bf16_round(x.x * rstd), w.y * bf16_round(x.y * rstd), w.z * bf16_round(x.z * rstd),
                   w.w * bf16_round(x.w * rstd));
  }
}

__global__ void __launch_bounds__(kRowsThreads) finalize_rows_kernel(const RowsArgs a) {
  extern __shared__ __align__(16) float rowbuf[];
  __shared__ float red[kRowsThreads / 32];
  __shared__ int ns_tab[kRowsMaxTiles];
  // everything that does not depend on the producing GEMM happens while this kernel waits for it
  if (a.embed == nullptr) {
    const int nt = (a.H + kTileN - 1) / kTileN;
    for (int t = threadIdx.x; t < nt; t += kRowsThreads) ns_tab[t] = tile_slots32(t, a.sm);
  }
  __syncthreads();
  // wait first, trigger second: the GEMM behind this kernel then becomes resident exactly when this
  // kernel starts its real work
  DFL_WAIT_THEN_TRIGGER();
  finalize_row_body<kRowsThreads, true>(a, blockIdx.x, threadIdx.x, rowbuf, red, ns_tab, 0);
  DFL_TRACE(2);
}

// The same row pass with kRowCtas CTAs per row (one thread-block cluster): a lone CTA has to pull ~110 KB of partials
// through one SM's load path (3-4 us in the step timeline, scripts/step_trace.py); four SMs share it here. Each CTA
// owns a quarter of the columns, keeps its values in registers, and the row's sum of squares is exchanged through
// distributed shared memory (every CTA pushes its partial sum into all peers, one cluster barrier, fixed summation
// order). Partials + optional residual + RMSNorm only (what the per-layer launches need).
#ifndef DFLASH_ROW_CTAS
#define DFLASH_ROW_CTAS 4
#endif
constexpr int kRowCtas = DFLASH_ROW_CTAS;
constexpr int kRowClThreads = 256;
constexpr int kRowClGroups = 2;  // float4 groups per thread: H / kRowCtas <= 256 * 4 * 2 (H <= 8192)

__global__ void __launch_bounds__(kRowClThreads) finalize_rows_cluster_kernel(const RowsArgs a) {
  __shared__ int ns_tab[kRowsMaxTiles];
  __shared__ float red[kRowClThreads / 32];
  __shared__ float peer_ss[kRowCtas];
  const int rank = static_cast<int>(cluster_cta_rank());
  const int row = blockIdx.y, tid = threadIdx.x;
  const int cols = a.H / kRowCtas, c0 = rank * cols;
  {
    const int nt = (a.H + kTileN - 1) / kTileN;
    for (int t = tid; t < nt; t += kRowClThreads) ns_tab[t] = tile_slots32(t, a.sm);
  }
  __syncthreads();
  DFL_WAIT_THEN_TRIGGER();
  if (a.valid_mode == kRowsCtx) {  // the whole cluster takes the same branch (no barrier is left half-entered)
    const int r = row / a.SL, j = row % a.SL;
    if (j >= a.ctx_len[r]) return;
  }
  const long long roff = static_cast<long long>(row) * a.H;
  float4 v[kRowClGroups];
  uint2 rs[kRowClGroups], wv[kRowClGroups];
#pragma unroll
  for (int g = 0; g < kRowClGroups; ++g) {
    const int n = c0 + tid * 4 + g * kRowClThreads * 4;
    if (n < c0 + cols) {
      wv[g] = *reinterpret_cast<const uint2*>(a.norm_w + n);
      v[g] = sum_slots_4(a.ws, a.sm, row, n, ns_tab[n / kTileN]);
      if (a.bias != nullptr) {
        const float4 b = unpack4_bf16(*reinterpret_cast<const uint2*>(a.bias + n));
        v[g].x += b.x; v[g].y += b.y; v[g].z += b.z; v[g].w += b.w;
      }
      if (a.resid != nullptr) rs[g] = __ldcg(reinterpret_cast<const uint2*>(a.resid + roff + n));
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int g = 0; g < kRowClGroups; ++g) {
    const int n = c0 + tid * 4 + g * kRowClThreads * 4;
    if (n >= c0 + cols) continue;
    float4& x = v[g];
    x.x = bf16_round(x.x); x.y = bf16_round(x.y); x.z = bf16_round(x.z); x.w = bf16_round(x.w);
    if (a.resid != nullptr) {
      const float4 rsd = unpack4_bf16(rs[g]);
      x.x = bf16_round(rsd.x + x.x); x.y = bf16_round(rsd.y + x.y);
      x.z = bf16_round(rsd.z + x.z); x.w = bf16_round(rsd.w + x.w);
      *reinterpret_cast<uint2*>(a.resid + roff + n) = pack4_bf16(x.x, x.y, x.z, x.w);
    }
    ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
  }
  DFL_TRACE(4);
  ss = warp_sum(ss);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  if (tid < kRowCtas) {  // thread p pushes this CTA's partial sum into CTA p's peer_ss[rank]
    float part = 0.f;
#pragma unroll
    for (int w = 0; w < kRowClThreads / 32; ++w) part += red[w];
    dsmem_st_f32(dsmem_map(smem_u32(&peer_ss[rank]), static_cast<uint32_t>(tid)), part);
  }
  cluster_sync_all();
  DFL_TRACE(5);
  float tot = 0.f;
#pragma unroll
  for (int p = 0; p < kRowCtas; ++p) tot += peer_ss[p];
  const float rstd = 1.0f / sqrtf(tot / static_cast<float>(a.H) + a.eps);
#pragma unroll
  for (int g = 0; g < kRowClGroups; ++g) {
    const int n = c0 + tid * 4 + g * kRowClThreads * 4;
    if (n >= c0 + cols) continue;
    const float4 w = unpack4_bf16(wv[g]);
    *reinterpret_cast<uint2*>(a.out + roff + n) =
        pack4_bf16(w.x * bf16_round(v[g].x * rstd), w.y * bf16_round(v[g].y * rstd), w.z * bf16_round(v[g].z * rstd),
                   w.w * bf16_round(v[g].w * rstd));
  }
  DFL_TRACE(2);
}

// The same row pass on ONE CTA of 1024 threads per row (hidden <= 4096: four columns per thread, everything in
// registers, all of a thread's slot loads in flight together, the row's sum of squares through shared memory): no
// cluster barrier and no distributed-shared-memory exchange, which were 1.6-3 us of the cluster form's 4.9 us.
#ifndef DFLASH_ROW_BLOCK
#define DFLASH_ROW_BLOCK 1
#endif
// rows up to which the one-CTA form is used (tuning switch). Measured better at every batch width -- 1 / 2 / 8 / 16 /
// 32 / 64 streams: -10 / -7 / -14 / -11 / -6 / -100 us per step (at wide batches the cluster form's 4096 CTAs each
// live for a cluster barrier) -- so the cluster form only serves hidden sizes above 4096.
#ifndef DFLASH_ROW_BLOCK_MAX_ROWS
#define DFLASH_ROW_BLOCK_MAX_ROWS (1 << 30)
#endif
constexpr int kRowBlkThreads = 1024;

__global__ void __launch_bounds__(kRowBlkThreads, 1) finalize_rows_block_kernel(const RowsArgs a) {
  __shared__ int ns_tab[kRowsMaxTiles];
  __shared__ float red[kRowBlkThreads / 32];
  const int row = blockIdx.x, tid = threadIdx.x;
  {
    const int nt = (a.H + kTileN - 1) / kTileN;
    for (int t = tid; t < nt; t += kRowBlkThreads) ns_tab[t] = tile_slots32(t, a.sm);
  }
  __syncthreads();
  DFL_WAIT_THEN_TRIGGER();
  if (a.valid_mode == kRowsCtx) {
    const int r = row / a.SL, j = row % a.SL;
    if (j >= a.ctx_len[r]) return;
  }
  const long long roff = static_cast<long long>(row) * a.H;
  const int n = tid * 4;
  const bool live = n < a.H;
  float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
  uint2 rs = make_uint2(0u, 0u), wv = make_uint2(0u, 0u);
  if (live) {
    wv = *reinterpret_cast<const uint2*>(a.norm_w + n);
    if (a.resid != nullptr) rs = __ldcg(reinterpret_cast<const uint2*>(a.resid + roff + n));
    x = sum_slots_4(a.ws, a.sm, row, n, ns_tab[n / kTileN]);
    if (a.bias != nullptr) {
      const float4 b = unpack4_bf16(*reinterpret_cast<const uint2*>(a.bias + n));
      x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
    }
  }
  float ss = 0.f;
  if (live) {
    x.x = bf16_round(x.x); x.y = bf16_round(x.y); x.z = bf16_round(x.z); x.w = bf16_round(x.w);
    if (a.resid != nullptr) {
      const float4 rsd = unpack4_bf16(rs);
      x.x = bf16_round(rsd.x + x.x); x.y = bf16_round(rsd.y + x.y);
      x.z = bf16_round(rsd.z + x.z); x.w = bf16_round(rsd.w + x.w);
      *reinterpret_cast<uint2*>(a.resid + roff + n) = pack4_bf16(x.x, x.y, x.z, x.w);
    }
    ss = x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
  }
  DFL_TRACE(4);
  ss = warp_sum(ss);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  DFL_TRACE(5);
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kRowBlkThreads / 32; ++w) tot += red[w];
  const float rstd = 1.0f / sqrtf(tot / static_cast<float>(a.H) + a.eps);
  if (live) {
    const float4 w = unpack4_bf16(wv);
    *reinterpret_cast<uint2*>(a.out + roff + n) =
        pack4_bf16(w.x * bf16_round(x.x * rstd), w.y * bf16_round(x.y * rstd), w.z * bf16_round(x.z * rstd),
                   w.w * bf16_round(x.w * rstd));
  }
  DFL_TRACE(2);
}

// ---------------------------------------------------------------------------------------------
// QKV post-processing: per (row, head) warp. q/k: per-head RMSNorm over D=128 then RoPE
// (half-split rotate, cos/sin rounded to bf16, products and sum rounded to bf16:
// model/dflash.py:22-28,70-82); v: bf16 round. K/V go straight into the static draft cache at
// their absolute position, q into the query buffer.
struct QkvPostArgs {
  const float* ws;
  SlotMap sm;
  int R, SL, bs;
  int Hq, Hkv;      // D == 128
  int q_cols;       // Hq*128, or 0 when the GEMM covered only the K/V weight rows (prefill)
  int row0, rows;   // rows of the activation matrix covered by this launch
  const int* start;
  const int* ctx_len;
  const int* blk_len;
  const __nv_bfloat16* q_norm_w;
  const __nv_bfloat16* k_norm_w;
  const __nv_bfloat16* bias;  // optional [q_cols + 2*Hkv*128] = [q_proj.bias; k_proj.bias; v_proj.bias] (attention_bias)
  const float* inv_freq;  // [64]
  float rope_scale;
  float eps;
  __nv_bfloat16* q_out;    // [R*SL][Hq][128]
  __nv_bfloat16* k_cache;  // [R][Hkv][S_max][128] (this layer)
  __nv_bfloat16* v_cache;
  int S_max;
  // prompt pass (pf_rows > 0): every row is a context row of request pf_req at position pf_pos0 + row
  int pf_rows, pf_req, pf_pos0;
};

// one (activation row, head column block hh) item per warp. Split in two so that the stand-alone kernel can run the
// part that only reads request state, weights and the rope table BEFORE griddepcontrol.wait (request state is written
// by the previous step's accept kernel, never by the GEMM in front of this kernel).
struct QkvItem {
  int kind;        // 0 q, 1 k, 2 v, -1 nothing to do
  int ws_row, hh;
  __nv_bfloat16* dst;
  float4 wv;       // norm weights of this lane's 4 elements
  float cs[4], sn[4];
};

__device__ __forceinline__ QkvItem qkv_post_prepare(const QkvPostArgs& a, int row, int hh, int lane) {
  QkvItem it;
  it.kind = -1;
  const int heads_q = a.q_cols / 128;
  const int RS = a.R * a.SL;
  const bool is_block = a.pf_rows == 0 && row >= RS;
  const int rl = is_block ? row - RS : row;
  const int r = a.pf_rows > 0 ? a.pf_req : rl / a.SL, slot = rl % a.SL;
  int pos;
  if (a.pf_rows > 0) {
    if (row >= a.pf_rows) return it;
    pos = a.pf_pos0 + row;
  } else if (is_block) {
    if (slot >= a.blk_len[r]) return it;
    pos = a.start[r] + slot;
  } else {
    const int c = a.ctx_len[r];
    if (slot >= c) return it;
    pos = a.start[r] - c + slot;
  }
  const int kind = hh < heads_q ? 0 : (hh < heads_q + a.Hkv ? 1 : 2);  // q, k, v
  if (kind == 0 && !is_block) return it;  // context rows carry no queries
  if (pos < 0 || pos >= a.S_max) return it;
  const int head = kind == 0 ? hh : (kind == 1 ? hh - heads_q : hh - heads_q - a.Hkv);
  it.ws_row = row - a.row0;
  it.hh = hh;
  if (kind == 0) {
    it.dst = a.q_out + (static_cast<long long>(rl) * a.Hq + head) * 128;
  } else {
    __nv_bfloat16* base = kind == 1 ? a.k_cache : a.v_cache;
    it.dst = base + ((static_cast<long long>(r) * a.Hkv + head) * a.S_max + pos) * 128;
  }
  it.kind = kind;
  if (kind == 2) return it;
  const __nv_bfloat16* w = kind == 0 ? a.q_norm_w : a.k_norm_w;
  it.wv = unpack4_bf16(*reinterpret_cast<const uint2*>(w + lane * 4));
  // RoPE: element d pairs with d +- 64 -> held by lane ^ 16; frequency index = d mod 64
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int f = (lane & 15) * 4 + t;
    const float ang = static_cast<float>(pos) * a.inv_freq[f];
    float sn, cs;
    sincosf(ang, &sn, &cs);
    it.cs[t] = bf16_round(cs * a.rope_scale);
    it.sn[t] = bf16_round(sn * a.rope_scale);
  }
  return it;
}

// the item's four fp32 sums per lane (the slots are L2 round trips: a warp that owns several items requests all of them
// before it finishes the first)
__device__ __forceinline__ float4 qkv_post_load(const QkvPostArgs& a, const QkvItem& it, int lane) {
  if (it.kind < 0) return make_float4(0.f, 0.f, 0.f, 0.f);
  // head hh covers output columns [hh*128, hh*128+128) = exactly stream-K tile hh
  return sum_slots_4(a.ws, a.sm, it.ws_row, it.hh * 128 + lane * 4, tile_slots32(it.hh, a.sm));
}

__device__ __forceinline__ void qkv_post_finish(const QkvPostArgs& a, const QkvItem& it, float4 xv, int lane) {
  if (it.kind < 0) return;
  if (a.bias != nullptr) {
    const float4 b = unpack4_bf16(*reinterpret_cast<const uint2*>(a.bias + it.hh * 128 + lane * 4));
    xv.x += b.x; xv.y += b.y; xv.z += b.z; xv.w += b.w;
  }
  float x[4] = {bf16_round(xv.x), bf16_round(xv.y), bf16_round(xv.z), bf16_round(xv.w)};  // d = 4*lane + t
  if (it.kind == 2) {
    *reinterpret_cast<uint2*>(it.dst + lane * 4) = pack4_bf16(x[0], x[1], x[2], x[3]);
    return;
  }
  float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
  ss = warp_sum(ss);
  const float rstd = 1.0f / sqrtf(ss * (1.0f / 128.0f) + a.eps);
  x[0] = bf16_round(it.wv.x * bf16_round(x[0] * rstd));
  x[1] = bf16_round(it.wv.y * bf16_round(x[1] * rstd));
  x[2] = bf16_round(it.wv.z * bf16_round(x[2] * rstd));
  x[3] = bf16_round(it.wv.w * bf16_round(x[3] * rstd));
  float o[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float other = __shfl_xor_sync(0xffffffffu, x[t], 16);
    // first half (lane < 16): x*cos + (-x_hi)*sin ; second half: x*cos + x_lo*sin
    const float rot = (lane < 16) ? -other : other;
    o[t] = bf16_round(bf16_round(x[t] * it.cs[t]) + bf16_round(rot * it.sn[t]));
  }
  *reinterpret_cast<uint2*>(it.dst + lane * 4) = pack4_bf16(o[0], o[1], o[2], o[3]);
}

__device__ __forceinline__ void qkv_post_apply(const QkvPostArgs& a, const QkvItem& it, int lane) {
  qkv_post_finish(a, it, qkv_post_load(a, it, lane), lane);
}

__device__ __forceinline__ void qkv_post_rowhead(const QkvPostArgs& a, int row, int hh, int lane) {
  const QkvItem it = qkv_post_prepare(a, row, hh, lane);
  qkv_post_apply(a, it, lane);
  DFL_TRACE(2);
}

// item in [0, rows * (q_cols/128 + 2*Hkv))
__device__ __forceinline__ void qkv_post_item(const QkvPostArgs& a, int item, int lane) {
  const int heads_per_row = a.q_cols / 128 + 2 * a.Hkv;
  qkv_post_rowhead(a, a.row0 + item / heads_per_row, item % heads_per_row, lane);
}

__global__ void __launch_bounds__(32 * kItemWarps) qkv_post_kernel(const QkvPostArgs a) {
  const int item = blockIdx.x * kItemWarps + (threadIdx.x >> 5);
  const int heads_per_row = a.q_cols / 128 + 2 * a.Hkv;
  const int lane = threadIdx.x & 31;
  QkvItem it;
  it.kind = -1;
  // positions, rope table and norm weights while the QKV GEMM in front of this kernel is still running. The request
  // state read here was written by the previous cycle's verify kernel, which has completed: every kernel of the draft
  // step sits behind the context-injection kernel, and that kernel releases its dependents only after its own
  // griddepcontrol.wait for the verify kernel has returned (gathered form: at its last MMA, whose operands were loaded
  // after the wait; direct form: explicitly after the wait in its tail).
  if (item < a.rows * heads_per_row) it = qkv_post_prepare(a, a.row0 + item / heads_per_row, item % heads_per_row, lane);
  DFL_WAIT_THEN_TRIGGER();
  qkv_post_apply(a, it, lane);
  DFL_TRACE(2);
}

}  // namespace dfl
