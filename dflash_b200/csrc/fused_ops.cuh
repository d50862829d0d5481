// Small fused kernels around the skinny GEMMs: they consume the fp32 split-K partials, apply the
// reference's bf16 rounding points (Linear output -> bf16, RMSNorm in fp32 -> bf16 -> * weight,
// residual add in bf16, SiLU*up in bf16, RoPE in bf16) and produce the next GEMM's activations.
#pragma once
#include "gemm_host.cuh"

namespace dfl {

// How a consumer finds the partial slots of output column n (mirrors the GEMM's stream-K split).
struct SlotMap {
  int k_blocks;
  long long T;
  long long G;
  int ws_rows;
  long long ws_ld;
};

inline SlotMap slot_map_of(const GemmPlan& p) {
  SlotMap s;
  s.k_blocks = p.args.k_blocks;
  s.T = static_cast<long long>(p.args.n_tiles) * p.args.k_blocks;
  s.G = p.grid;
  s.ws_rows = p.args.ws_rows;
  s.ws_ld = p.args.ws_ld;
  return s;
}

// fp32 sum over the slots of (row m, column n), in slot order.
__device__ __forceinline__ float sum_slots(const float* __restrict__ ws, const SlotMap& sm, int m, int n) {
  const int t = n / kTileN;
  const int ns = tile_num_slots(t, sm.k_blocks, sm.T, sm.G);
  const float* p = ws + static_cast<long long>(m) * sm.ws_ld + n;
  const long long slot_stride = static_cast<long long>(sm.ws_rows) * sm.ws_ld;
  float acc = p[0];
  for (int s = 1; s < ns; ++s) acc += p[s * slot_stride];
  return acc;
}

// ---------------------------------------------------------------------------------------------
// test helper: out[m][n] = sum of partial slots (fp32)
__global__ void sum_slots_kernel(const float* __restrict__ ws, SlotMap sm, int rows, int N, float* out,
                                 long long out_ld) {
  pdl_wait();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n < N && m < rows) out[static_cast<long long>(m) * out_ld + n] = sum_slots(ws, sm, m, n);
}

inline cudaError_t launch_sum_slots(const GemmPlan& p, float* out, long long out_ld, cudaStream_t st) {
  dim3 grid((p.args.N + 255) / 256, p.args.m_valid);
  sum_slots_kernel<<<grid, 256, 0, st>>>(p.args.ws, slot_map_of(p), p.args.m_valid, p.args.N, out, out_ld);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// lm_head second stage: per activation row, max over the per-CTA candidates (ties -> lowest index,
// matching torch.argmax on the bf16 logits: model/utils.py:27-29).
// tokens[row] (int64) <- argmax. Rows with row_mask[row]==0 are skipped when a mask is given.
__global__ void reduce_candidates_kernel(const float* __restrict__ cand_val, const int* __restrict__ cand_idx,
                                         int n_cta, int mb, int rows, long long* __restrict__ tokens) {
  pdl_wait();
  const int row = blockIdx.x;
  if (row >= rows) return;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int g = threadIdx.x; g < n_cta; g += 32) {
    const float v = cand_val[static_cast<long long>(g) * mb + row];
    const int i = cand_idx[static_cast<long long>(g) * mb + row];
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (threadIdx.x == 0) tokens[row] = bi;
}

inline cudaError_t launch_reduce_candidates(const float* cand_val, const int* cand_idx, int n_cta, int mb,
                                            int rows, long long* tokens, cudaStream_t st) {
  reduce_candidates_kernel<<<rows, 32, 0, st>>>(cand_val, cand_idx, n_cta, mb, rows, tokens);
  return cudaGetLastError();
}


// Engine form of the lm_head second stage: block row (r, i) -> drafted token for slot i, written
// into block_ids[r][i] for 1 <= i < bs (slot 0 is the committed token: dflash.py:247).
struct DraftTokArgs {
  const float* cand_val;
  const int* cand_idx;
  int n_cta, mb;
  int R, SL, bs;
  long long* block_ids;     // [R][bs]
  long long* draft_tokens;  // [R*SL] every row's argmax (slot 0 included), for inspection
};

__global__ void __launch_bounds__(32) draft_tokens_kernel(const DraftTokArgs a) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int g = threadIdx.x; g < a.n_cta; g += 32) {
    const float v = a.cand_val[static_cast<long long>(g) * a.mb + row];
    const int i = a.cand_idx[static_cast<long long>(g) * a.mb + row];
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (threadIdx.x == 0) {
    a.draft_tokens[row] = bi;
    const int r = row / a.SL, i = row % a.SL;
    if (i >= 1 && i < a.bs) a.block_ids[static_cast<long long>(r) * a.bs + i] = bi;
  }
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

constexpr int kRowsThreads = 256;
constexpr int kRowsMaxPerThread = 32;  // H <= 8192

// One CTA per row:  v = bf16(sum of partials)           (nn.Linear output dtype)
//                   v = bf16(resid + v); resid = v       (residual add, model/dflash.py:140,144)
//                   out = w * bf16(v * rsqrt(mean(v^2) + eps))   (Qwen3RMSNorm, fp32 inside)
__global__ void __launch_bounds__(kRowsThreads) finalize_rows_kernel(const RowsArgs a) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x;
  if (a.valid_mode == kRowsCtx) {
    const int r = row / a.SL, j = row % a.SL;
    if (j >= a.ctx_len[r]) return;
  }
  __shared__ float red[kRowsThreads / 32];
  float v[kRowsMaxPerThread];
  float ss = 0.f;
  const long long roff = static_cast<long long>(row) * a.H;
  long long tok = 0;
  if (a.embed != nullptr) {
    const int r = row / a.SL, i = row % a.SL;
    if (a.ids == nullptr) tok = row;  // `embed` is a [rows, H] embedding matrix already (forward(noise_embedding=...))
    else tok = (i < a.bs) ? a.ids[static_cast<long long>(r) * a.ids_ld + i] : a.pad_token;
  }
#pragma unroll
  for (int k = 0; k < kRowsMaxPerThread; ++k) {
    const int n = k * kRowsThreads + threadIdx.x;
    if (n < a.H) {
      float x;
      if (a.embed != nullptr) {
        x = __bfloat162float(a.embed[tok * a.H + n]);
        if (a.resid != nullptr) a.resid[roff + n] = __float2bfloat16_rn(x);
      } else {
        x = bf16_round(sum_slots(a.ws, a.sm, row, n));
        if (a.resid != nullptr) {
          x = bf16_round(__bfloat162float(a.resid[roff + n]) + x);
          a.resid[roff + n] = __float2bfloat16_rn(x);
        }
      }
      v[k] = x;
      ss += x * x;
    }
  }
  if (a.norm_w == nullptr) {
#pragma unroll
    for (int k = 0; k < kRowsMaxPerThread; ++k) {
      const int n = k * kRowsThreads + threadIdx.x;
      if (n < a.H) a.out[roff + n] = __float2bfloat16_rn(v[k]);
    }
    return;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kRowsThreads / 32; ++w) tot += red[w];
  const float rstd = 1.0f / sqrtf(tot / static_cast<float>(a.H) + a.eps);
#pragma unroll
  for (int k = 0; k < kRowsMaxPerThread; ++k) {
    const int n = k * kRowsThreads + threadIdx.x;
    if (n < a.H) {
      const float y = bf16_round(v[k] * rstd);
      a.out[roff + n] = __float2bfloat16_rn(__bfloat162float(a.norm_w[n]) * y);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// SwiGLU: out[m, n] = bf16( bf16(silu(gate[m,n])) * up[m,n] ), gate = cols [0,I), up = cols [I,2I) of
// the fused gate/up GEMM (Qwen3MLP: down_proj(act_fn(gate_proj(x)) * up_proj(x))).
struct SwigluArgs {
  const float* ws;
  SlotMap sm;
  int rows;
  int I;
  __nv_bfloat16* out;  // [rows, I]
};

__global__ void __launch_bounds__(256) swiglu_kernel(const SwigluArgs a) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.x * 256 + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= a.I) return;
  const float g = bf16_round(sum_slots(a.ws, a.sm, m, n));
  const float u = bf16_round(sum_slots(a.ws, a.sm, m, a.I + n));
  const float s = bf16_round(g / (1.0f + expf(-g)));
  a.out[static_cast<long long>(m) * a.I + n] = __float2bfloat16_rn(s * u);
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
  const float* inv_freq;  // [64]
  float rope_scale;
  float eps;
  __nv_bfloat16* q_out;    // [R*SL][Hq][128]
  __nv_bfloat16* k_cache;  // [R][Hkv][S_max][128] (this layer)
  __nv_bfloat16* v_cache;
  int S_max;
};

__global__ void __launch_bounds__(256) qkv_post_kernel(const QkvPostArgs a) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int heads_q = a.q_cols / 128;
  const int heads_per_row = heads_q + 2 * a.Hkv;
  const int item = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (item >= a.rows * heads_per_row) return;
  const int row = a.row0 + item / heads_per_row;
  const int hh = item % heads_per_row;
  const int RS = a.R * a.SL;
  const bool is_block = row >= RS;
  const int rl = is_block ? row - RS : row;
  const int r = rl / a.SL, slot = rl % a.SL;
  int pos;
  if (is_block) {
    if (slot >= a.blk_len[r]) return;
    pos = a.start[r] + slot;
  } else {
    const int c = a.ctx_len[r];
    if (slot >= c) return;
    pos = a.start[r] - c + slot;
  }
  const int kind = hh < heads_q ? 0 : (hh < heads_q + a.Hkv ? 1 : 2);  // q, k, v
  if (kind == 0 && !is_block) return;  // context rows carry no queries
  if (pos < 0 || pos >= a.S_max) return;
  const int head = kind == 0 ? hh : (kind == 1 ? hh - heads_q : hh - heads_q - a.Hkv);
  const int col0 = hh * 128;
  const int ws_row = row - a.row0;

  float x[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) x[t] = bf16_round(sum_slots(a.ws, a.sm, ws_row, col0 + lane + 32 * t));

  __nv_bfloat16* dst;
  if (kind == 0) {
    dst = a.q_out + (static_cast<long long>(rl) * a.Hq + head) * 128;
  } else {
    __nv_bfloat16* base = kind == 1 ? a.k_cache : a.v_cache;
    dst = base + ((static_cast<long long>(r) * a.Hkv + head) * a.S_max + pos) * 128;
  }
  if (kind == 2) {
#pragma unroll
    for (int t = 0; t < 4; ++t) dst[lane + 32 * t] = __float2bfloat16_rn(x[t]);
    return;
  }
  const __nv_bfloat16* w = kind == 0 ? a.q_norm_w : a.k_norm_w;
  float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
  ss = warp_sum(ss);
  const float rstd = 1.0f / sqrtf(ss * (1.0f / 128.0f) + a.eps);
#pragma unroll
  for (int t = 0; t < 4; ++t)
    x[t] = bf16_round(__bfloat162float(w[lane + 32 * t]) * bf16_round(x[t] * rstd));
  // RoPE: element d pairs with d +- 64; this lane holds d = lane, lane+32 (first half) and
  // lane+64, lane+96 (second half); frequency index = d mod 64.
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const float ang = static_cast<float>(pos) * a.inv_freq[lane + 32 * t];
    float sn, cs;
    sincosf(ang, &sn, &cs);
    cs = bf16_round(cs * a.rope_scale);
    sn = bf16_round(sn * a.rope_scale);
    const float lo = x[t], hi = x[t + 2];
    const float olo = bf16_round(bf16_round(lo * cs) + bf16_round(-hi * sn));
    const float ohi = bf16_round(bf16_round(hi * cs) + bf16_round(lo * sn));
    dst[lane + 32 * t] = __float2bfloat16_rn(olo);
    dst[lane + 32 * t + 64] = __float2bfloat16_rn(ohi);
  }
}

}  // namespace dfl
