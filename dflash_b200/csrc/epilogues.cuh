// Row-wise epilogues that run INSIDE the streaming GEMM (gemm_skinny.cuh), on a finished 128-column output tile
// staged in shared memory as tile[m][n] (m = activation row, n = column inside the tile): one warp per activation
// row, a lane owns 4 consecutive columns. They apply the reference's bf16 rounding points
// (Linear output -> bf16, residual add in bf16, SiLU*up in bf16, per-head RMSNorm + RoPE in bf16) and write the
// next kernel's bf16 operands directly, so no fp32 partial sum of a finished tile ever reaches HBM.
//
//   RowsEpi    fc / o_proj / down_proj   model/dflash.py:177,101,140,144 ; Qwen3MLP.down_proj
//   SwigluEpi  gate_proj + up_proj       Qwen3MLP: act_fn(gate_proj(x)) * up_proj(x)
//   QkvPostArgs q/k/v_proj               model/dflash.py:22-28,70-85 (q_norm/k_norm, RoPE, cache append)
#pragma once
#include "ptx.cuh"

namespace dfl {

__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&lo);
  r.y = *reinterpret_cast<uint32_t*>(&hi);
  return r;
}
__device__ __forceinline__ float4 unpack4_bf16(uint2 r) {
  const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
  const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// ---------------------------------------------------------------------------------------------
// Linear output (+ bias) -> bf16, optional residual add (bf16), and the tile's share of the row's sum of squares
// (the RMSNorm that follows needs the whole row: the light norm kernel adds the per-tile sums in tile order).
struct RowsEpi {
  const __nv_bfloat16* bias;  // [N] or null (config.attention_bias: o_proj)
  __nv_bfloat16* resid;       // [rows][ld] residual stream (in/out), or null
  __nv_bfloat16* out;         // [rows][ld]; == resid for o_proj / down_proj
  long long ld;
  float* tile_ss;             // [n_tiles][ss_ld]
  int ss_ld;
};

// rs: the row's residual values of these 4 columns, requested by the caller ahead of time (ignored without a residual)
__device__ __forceinline__ void rows_epi_apply(const RowsEpi& e, float4 x, uint2 rs, int tile, int row, int lane) {
  const int n = tile * 128 + lane * 4;
  if (e.bias != nullptr) {
    const float4 b = unpack4_bf16(*reinterpret_cast<const uint2*>(e.bias + n));
    x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
  }
  x.x = bf16_round(x.x); x.y = bf16_round(x.y); x.z = bf16_round(x.z); x.w = bf16_round(x.w);
  const long long off = static_cast<long long>(row) * e.ld + n;
  if (e.resid != nullptr) {
    const float4 r = unpack4_bf16(rs);
    x.x = bf16_round(r.x + x.x); x.y = bf16_round(r.y + x.y);
    x.z = bf16_round(r.z + x.z); x.w = bf16_round(r.w + x.w);
  }
  *reinterpret_cast<uint2*>(e.out + off) = pack4_bf16(x.x, x.y, x.z, x.w);
  const float ss = warp_sum(x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w);
  if (lane == 0) e.tile_ss[static_cast<long long>(tile) * e.ss_ld + row] = ss;
}

// ---------------------------------------------------------------------------------------------
// SwiGLU: a tile is 64 gate rows + 64 up rows of the SAME 64 intermediate columns (two 64-row TMA boxes out of the
// unchanged [gate; up] weight stack), so gate and up of a column meet in one CTA.
struct SwigluEpi {
  __nv_bfloat16* out;  // [rows][ld] = hmid
  long long ld;
  int I;               // intermediate size (up rows start at weight row I)
};

__device__ __forceinline__ float silu_mul_bf16(float g, float u) {
  g = bf16_round(g);
  u = bf16_round(u);
  const float s = bf16_round(g / (1.0f + expf(-g)));
  return s * u;
}

// tile_row: the 128 staged values of one activation row (gate 0..63, up 64..127)
__device__ __forceinline__ void swiglu_epi_apply(const SwigluEpi& e, const float* tile_row, int tile, int row, int lane) {
  const float2 g = *reinterpret_cast<const float2*>(tile_row + 2 * lane);
  const float2 u = *reinterpret_cast<const float2*>(tile_row + 64 + 2 * lane);
  const __nv_bfloat162 o = __floats2bfloat162_rn(silu_mul_bf16(g.x, u.x), silu_mul_bf16(g.y, u.y));
  *reinterpret_cast<__nv_bfloat162*>(e.out + static_cast<long long>(row) * e.ld + tile * 64 + 2 * lane) = o;
}

// ---------------------------------------------------------------------------------------------
// QKV post-processing: a tile is one head (D = 128). q/k: per-head RMSNorm over D then RoPE (half-split rotate,
// cos/sin rounded to bf16, products and sum rounded to bf16: model/dflash.py:22-28,70-82); v: bf16 round. K/V go
// straight into the static draft cache at their absolute position, q into the query buffer.
//
// Row layout shared by all step kernels. R requests, SL row slots per request (bs <= SL).
//   activation matrix a_in [2*R*SL, H]: rows [0, R*SL) = context rows (request r, slot j live iff
//   j < ctx_len[r]); rows [R*SL, 2*R*SL) = block rows (request r, slot i live iff i < blk_len[r]).
struct QkvPostArgs {
  int R, SL;
  int Hq, Hkv;      // D == 128
  int q_cols;       // Hq*128, or 0 when the GEMM covered only the K/V weight rows (prompt pass)
  // Per activation row, filled once per step by rows_pre_kernel (per pass by rope_table_kernel in the prompt pass):
  // the row's absolute position (-1: dead row) and its rotary table, cos | sin of pos * inv_freq, scaled and rounded
  // to bf16 as Qwen3RotaryEmbedding does -- shared by all layers, so no epilogue evaluates a sine.
  const int* row_pos;          // [rows]
  const float* rope;           // [rows][128]: cos[0..63] | sin[0..63]
  const __nv_bfloat16* q_norm_w;
  const __nv_bfloat16* k_norm_w;
  const __nv_bfloat16* bias;   // [q_cols + 2*Hkv*128] or null (config.attention_bias)
  float eps;
  __nv_bfloat16* q_out;    // [R*SL][Hq][128]
  __nv_bfloat16* k_cache;  // [R][Hkv][S_max][128] (this layer)
  __nv_bfloat16* v_cache;
  int S_max;
  // prompt pass (pf_rows > 0): every row is a context row of request pf_req
  int pf_rows, pf_req;
};

// What a lane needs for one (row, head) item besides the GEMM's data; requested early (it does not depend on it).
struct QkvRowPre {
  int pos;          // < 0: nothing to do
  float4 cs, sn;    // rope table entries of this lane's 4 elements (frequency index (lane & 15) * 4 + t)
};

__device__ __forceinline__ QkvRowPre qkv_row_prefetch(const QkvPostArgs& a, int row, int lane) {
  QkvRowPre p;
  p.pos = __ldcg(a.row_pos + row);
  const float* t = a.rope + static_cast<long long>(row) * 128 + (lane & 15) * 4;
  p.cs = __ldcg(reinterpret_cast<const float4*>(t));
  p.sn = __ldcg(reinterpret_cast<const float4*>(t + 64));
  return p;
}

// kind of head column block hh: 0 q, 1 k, 2 v
__device__ __forceinline__ int qkv_kind(const QkvPostArgs& a, int hh) {
  const int heads_q = a.q_cols / 128;
  return hh < heads_q ? 0 : (hh < heads_q + a.Hkv ? 1 : 2);
}

// xv: this lane's 4 fp32 sums (elements d = 4*lane .. 4*lane+3 of head column block hh) of activation row `row`;
// wv: the q/k norm weights of those elements.
__device__ __forceinline__ void qkv_post_apply(const QkvPostArgs& a, const QkvRowPre& pre, float4 wv, float4 xv, int row,
                                               int hh, int lane) {
  const int kind = qkv_kind(a, hh);
  const int RS = a.R * a.SL;
  const bool is_block = a.pf_rows == 0 && row >= RS;
  if (pre.pos < 0 || (kind == 0 && !is_block)) return;  // dead row; context rows carry no queries
  const int rl = is_block ? row - RS : row;
  const int heads_q = a.q_cols / 128;
  __nv_bfloat16* dst;
  if (kind == 0) {
    dst = a.q_out + (static_cast<long long>(rl) * a.Hq + hh) * 128;
  } else {
    const int r = a.pf_rows > 0 ? a.pf_req : rl / a.SL;
    const int head = kind == 1 ? hh - heads_q : hh - heads_q - a.Hkv;
    dst = (kind == 1 ? a.k_cache : a.v_cache) + ((static_cast<long long>(r) * a.Hkv + head) * a.S_max + pre.pos) * 128;
  }
  if (a.bias != nullptr) {
    const float4 b = unpack4_bf16(*reinterpret_cast<const uint2*>(a.bias + hh * 128 + lane * 4));
    xv.x += b.x; xv.y += b.y; xv.z += b.z; xv.w += b.w;
  }
  float x[4] = {bf16_round(xv.x), bf16_round(xv.y), bf16_round(xv.z), bf16_round(xv.w)};
  if (kind == 2) {
    *reinterpret_cast<uint2*>(dst + lane * 4) = pack4_bf16(x[0], x[1], x[2], x[3]);
    return;
  }
  float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
  ss = warp_sum(ss);
  const float rstd = 1.0f / sqrtf(ss * (1.0f / 128.0f) + a.eps);
  x[0] = bf16_round(wv.x * bf16_round(x[0] * rstd));
  x[1] = bf16_round(wv.y * bf16_round(x[1] * rstd));
  x[2] = bf16_round(wv.z * bf16_round(x[2] * rstd));
  x[3] = bf16_round(wv.w * bf16_round(x[3] * rstd));
  const float cs[4] = {pre.cs.x, pre.cs.y, pre.cs.z, pre.cs.w};
  const float sn[4] = {pre.sn.x, pre.sn.y, pre.sn.z, pre.sn.w};
  float o[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    // RoPE: element d pairs with d +- 64 -> held by lane ^ 16
    const float other = __shfl_xor_sync(0xffffffffu, x[t], 16);
    // first half (lane < 16): x*cos + (-x_hi)*sin ; second half: x*cos + x_lo*sin
    const float rot = (lane < 16) ? -other : other;
    o[t] = bf16_round(bf16_round(x[t] * cs[t]) + bf16_round(rot * sn[t]));
  }
  *reinterpret_cast<uint2*>(dst + lane * 4) = pack4_bf16(o[0], o[1], o[2], o[3]);
}

// Row positions + rotary table of the step's activation rows (layout above). One thread per (row, frequency).
//   context row (r, j): pos = start[r] - ctx_len[r] + j, live iff j < ctx_len[r]
//   block row   (r, i): pos = start[r] + i,              live iff i < blk_len[r]     (model/dflash.py:241)
struct RopeTableArgs {
  int R, SL, S_max;
  const int* start;
  const int* ctx_len;
  const int* blk_len;
  const float* inv_freq;  // [64]
  float rope_scale;
  int* row_pos;           // [2*R*SL]
  float* rope;            // [2*R*SL][128]
  // prompt pass (pf_rows > 0): row j is position pf_pos0 + j
  int pf_rows, pf_pos0;
};

__device__ __forceinline__ void rope_table_row(const RopeTableArgs& a, int row, int tid) {
  if (tid >= 64) return;
  int pos;
  if (a.pf_rows > 0) {
    pos = row < a.pf_rows ? a.pf_pos0 + row : -1;
  } else {
    const int RS = a.R * a.SL;
    const bool is_block = row >= RS;
    const int rl = is_block ? row - RS : row;
    const int r = rl / a.SL, slot = rl % a.SL;
    if (is_block) {
      pos = slot < a.blk_len[r] ? a.start[r] + slot : -1;
    } else {
      const int c = a.ctx_len[r];
      pos = slot < c ? a.start[r] - c + slot : -1;
    }
  }
  if (pos >= a.S_max) pos = -1;
  if (tid == 0) a.row_pos[row] = pos;
  if (pos < 0) return;
  float sn, cs;
  sincosf(static_cast<float>(pos) * a.inv_freq[tid], &sn, &cs);
  a.rope[static_cast<long long>(row) * 128 + tid] = bf16_round(cs * a.rope_scale);
  a.rope[static_cast<long long>(row) * 128 + 64 + tid] = bf16_round(sn * a.rope_scale);
}

__global__ void __launch_bounds__(64) rope_table_kernel(const RopeTableArgs a) {
  pdl_wait();
  rope_table_row(a, blockIdx.x, threadIdx.x);
}

}  // namespace dfl
