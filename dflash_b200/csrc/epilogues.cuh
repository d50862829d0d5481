// The row-wise epilogues that run INSIDE the streaming GEMM (gemm_skinny.cuh) on a finished output tile staged in
// shared memory as tile[m][n] (m = activation row, n = column inside the tile), one warp per activation row:
//   kModeSwiglu   SwiGLU with the reference's bf16 rounding points, written as the down projection's bf16 operand, so
//                 the widest fp32 plane of the step (gate/up: 2 x intermediate columns) never reaches HBM;
//   kModeCtxNorm  the target-context injection (model/dflash.py:177, hidden_norm(fc(target_hidden))) as ONE kernel: the
//                 block rows' embedding gather + first input_layernorm (model/dflash.py:237, layer 0) on the epilogue
//                 warps while they are idle, and after the main loop a device-wide arrival count over the CTAs'
//                 split-K partials followed by the row pass (sum -> bf16 -> hidden_norm) on the CTAs themselves.
// (The same treatment of qkv / o / down was built and measured slower than consumer kernels at every batch width --
// DESIGN.md section 7 -- because only a tile's finishing CTA does the row work and the last tile's epilogue is exposed
// at the end of every GEMM.)
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
// Context injection: hidden_norm(fc(target_hidden)) -> a_in context rows; embed_tokens(block) -> x, ln1 -> a_in block rows
struct CtxNormEpi {
  __nv_bfloat16* out;            // [rows][ld] a_in context rows
  long long ld;                  // hidden size
  const __nv_bfloat16* norm_w;   // hidden_norm.weight [H]
  unsigned int* sync;            // [2]: CTAs whose partials are stored; CTAs that are past the wait (self-resetting)
  int max_slots;                 // most partial slots any column tile has
  // direct mode: the activation tiles are read IN PLACE from the selected target hidden states (XMaps, one 3-D map
  // per selected layer) and the launch overlaps the verify kernel in front of it: the main loop does not wait for that
  // kernel, everything that reads its results (ctx_len, the next block's first token) happens after the main loop
  int direct;
  int kb_per_sel;                // hidden / 64: k-blocks per selected layer
  const int* ctx_len;            // [R]: context row (r, j) is live iff j < ctx_len[r]
  int SL;
  float eps;
  // block rows (n_blk_rows = R * SL; row i of request r is token ids[r * ids_ld + i], pad_token past bs)
  const __nv_bfloat16* embed;    // embedding matrix [V][H], or (ids == null) already-embedded rows [n_blk_rows][H]
  const long long* ids;
  int ids_ld, bs, n_blk_rows;
  long long pad_token;
  __nv_bfloat16* resid;          // [n_blk_rows][H] residual stream (written)
  const __nv_bfloat16* ln_w;     // layers[0].input_layernorm.weight
  __nv_bfloat16* blk_out;        // [n_blk_rows][H]
};

// Context row `row` on all NT threads of the CTA, after every CTA's split-K partials are in `ws`:
//   v = bf16(sum of the partial slots, in slot order)      (nn.Linear output dtype)
//   out = w * bf16(v * rsqrt(mean(v^2) + eps))              (Qwen3RMSNorm, fp32 inside)
// ns_tab[t] = partial slots of column tile t; rowbuf: H floats of shared memory (the idle TMA pipeline); red: NT / 32
// floats of shared memory.
constexpr int kCtxMaxGroups = 11;  // float4 groups per thread: hidden <= 8192 over 192 threads
template <int NT>
__device__ __forceinline__ void ctxnorm_row_pass(const CtxNormEpi& e, const float* __restrict__ ws, long long slot_stride,
                                                 long long ws_ld, int row, int tid, const int* ns_tab, float* rowbuf,
                                                 float* red) {
  const int H = static_cast<int>(e.ld);
  // (fallback of ctxnorm_row_pass_bulk for rows that do not fit the pipeline's shared memory: kept small in registers,
  // the kernel shares its SMs with the verify kernel it overlaps)
  constexpr int kBatch = 1;  // column groups whose slot loads are in flight together (these are L2 round trips)
  constexpr int kSlots = 4;  // slots loaded up front; more are added after
  float ss = 0.f;
  for (int n0 = tid * 4; n0 < H; n0 += NT * 4 * kBatch) {
    float4 p[kBatch][kSlots];
#pragma unroll
    for (int b = 0; b < kBatch; ++b) {
      const int n = n0 + b * NT * 4;
      if (n < H) {
        const int ns = ns_tab[n >> 7];
        const float* src = ws + static_cast<long long>(row) * ws_ld + n;
#pragma unroll
        for (int s = 0; s < kSlots; ++s)
          if (s < ns) p[b][s] = __ldcg(reinterpret_cast<const float4*>(src + s * slot_stride));
      }
    }
#pragma unroll
    for (int b = 0; b < kBatch; ++b) {
      const int n = n0 + b * NT * 4;
      if (n < H) {
        const int ns = ns_tab[n >> 7];
        const float* src = ws + static_cast<long long>(row) * ws_ld + n;
        float4 acc = p[b][0];
#pragma unroll
        for (int s = 1; s < kSlots; ++s)
          if (s < ns) { acc.x += p[b][s].x; acc.y += p[b][s].y; acc.z += p[b][s].z; acc.w += p[b][s].w; }
        for (int s = kSlots; s < ns; ++s) {
          const float4 q = __ldcg(reinterpret_cast<const float4*>(src + s * slot_stride));
          acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
        }
        acc.x = bf16_round(acc.x); acc.y = bf16_round(acc.y); acc.z = bf16_round(acc.z); acc.w = bf16_round(acc.w);
        *reinterpret_cast<float4*>(rowbuf + n) = acc;
        ss += acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
      }
    }
  }
  ss = warp_sum(ss);
  __syncthreads();  // red[] of the previous row has been consumed
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) tot += red[w];
  const float rstd = 1.0f / sqrtf(tot / static_cast<float>(H) + e.eps);
  for (int n = tid * 4; n < H; n += NT * 4) {  // (each thread re-reads exactly what it wrote)
    const float4 x = *reinterpret_cast<const float4*>(rowbuf + n);
    const float4 w = unpack4_bf16(*reinterpret_cast<const uint2*>(e.norm_w + n));
    *reinterpret_cast<uint2*>(e.out + static_cast<long long>(row) * e.ld + n) =
        pack4_bf16(w.x * bf16_round(x.x * rstd), w.y * bf16_round(x.y * rstd), w.z * bf16_round(x.z * rstd),
                   w.w * bf16_round(x.w * rstd));
  }
}

// The same row pass with the row's partial slots pulled into shared memory by the TMA engine (max_slots bulk copies of
// H floats each into stage[slot][H], one mbarrier): ONE memory round trip however few threads the CTA has, instead of
// max_slots * H / (4 * NT) dependent-latency loads per thread. Needs max_slots * H * 4 bytes of (idle pipeline) smem.
template <int NT>
__device__ __forceinline__ void ctxnorm_row_pass_bulk(const CtxNormEpi& e, const float* __restrict__ ws, long long slot_stride,
                                                      long long ws_ld, int row, int tid, const int* ns_tab, float* stage,
                                                      uint64_t* bar, uint32_t parity, float* red) {
  const int H = static_cast<int>(e.ld);
  __syncthreads();  // the previous row's readers are done with stage[] and red[]
  if (tid == 0) {
    fence_proxy_async();  // (acquired partials of other CTAs; this CTA's earlier generic reads of stage[])
    mbar_expect_tx(bar, static_cast<uint32_t>(e.max_slots) * static_cast<uint32_t>(H) * 4u);
    for (int s = 0; s < e.max_slots; ++s)
      bulk_load_1d(stage + static_cast<long long>(s) * H, ws + s * slot_stride + static_cast<long long>(row) * ws_ld,
                   static_cast<uint32_t>(H) * 4u, bar);
  }
  constexpr int kWG = 6;  // norm weights requested before the wait (hidden <= 4608 over 192 threads: all of them)
  uint2 wv[kWG];
#pragma unroll
  for (int g = 0; g < kWG; ++g) {
    const int n = (tid + g * NT) * 4;
    if (n < H) wv[g] = *reinterpret_cast<const uint2*>(e.norm_w + n);
  }
  mbar_wait(bar, parity);
  float ss = 0.f;
  for (int n = tid * 4; n < H; n += NT * 4) {
    const int ns = ns_tab[n >> 7];
    float4 acc = *reinterpret_cast<const float4*>(stage + n);
    for (int s = 1; s < ns; ++s) {  // slot order: the summation order is fixed
      const float4 q = *reinterpret_cast<const float4*>(stage + static_cast<long long>(s) * H + n);
      acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
    }
    acc.x = bf16_round(acc.x); acc.y = bf16_round(acc.y); acc.z = bf16_round(acc.z); acc.w = bf16_round(acc.w);
    *reinterpret_cast<float4*>(stage + n) = acc;  // (slot 0's place; this thread is its only reader)
    ss += acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
  }
  ss = warp_sum(ss);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) tot += red[w];
  const float rstd = 1.0f / sqrtf(tot / static_cast<float>(H) + e.eps);
#pragma unroll
  for (int g = 0; g < kCtxMaxGroups; ++g) {
    const int n = (tid + g * NT) * 4;
    if (n < H) {
      const float4 x = *reinterpret_cast<const float4*>(stage + n);
      const float4 w = unpack4_bf16(g < kWG ? wv[g < kWG ? g : 0] : *reinterpret_cast<const uint2*>(e.norm_w + n));
      *reinterpret_cast<uint2*>(e.out + static_cast<long long>(row) * e.ld + n) =
          pack4_bf16(w.x * bf16_round(x.x * rstd), w.y * bf16_round(x.y * rstd), w.z * bf16_round(x.z * rstd),
                     w.w * bf16_round(x.w * rstd));
    }
  }
}

// Block row `row` on NT threads: embedding gather -> residual stream; input_layernorm -> blk_out.
// red: NT / 32 floats of shared memory. NAMED: the NT = 128 epilogue threads meet at named barrier 1; otherwise the
// whole CTA (NT threads) at barrier 0.
template <int NT, bool NAMED>
__device__ __forceinline__ void ctxnorm_embed_row(const CtxNormEpi& e, int row, int tid, float* red) {
  constexpr int kEmbedMaxIt = (8192 + NT * 4 - 1) / (NT * 4);  // hidden <= 8192
  auto sync = [&]() {
    if (NAMED) asm volatile("bar.sync 1, 128;\n" ::: "memory");
    else __syncthreads();
  };
  const int H = static_cast<int>(e.ld);
  long long tok = row;  // ids == null: `embed` already holds this row
  if (e.ids != nullptr) {
    const int r = row / e.SL, i = row % e.SL;
    tok = (i < e.bs) ? e.ids[static_cast<long long>(r) * e.ids_ld + i] : e.pad_token;
  }
  const __nv_bfloat16* src = e.embed + tok * H;
  const long long roff = static_cast<long long>(row) * H;
  uint2 v[kEmbedMaxIt];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < kEmbedMaxIt; ++k) {
    const int n = tid * 4 + k * NT * 4;
    if (n < H) v[k] = *reinterpret_cast<const uint2*>(src + n);
  }
#pragma unroll
  for (int k = 0; k < kEmbedMaxIt; ++k) {
    const int n = tid * 4 + k * NT * 4;
    if (n < H) {
      *reinterpret_cast<uint2*>(e.resid + roff + n) = v[k];
      const float4 x = unpack4_bf16(v[k]);
      ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
  }
  ss = warp_sum(ss);
  sync();  // red[] of the previous row has been consumed
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  sync();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) tot += red[w];
  const float rstd = 1.0f / sqrtf(tot / static_cast<float>(H) + e.eps);
#pragma unroll
  for (int k = 0; k < kEmbedMaxIt; ++k) {
    const int n = tid * 4 + k * NT * 4;
    if (n < H) {
      const float4 x = unpack4_bf16(v[k]);
      const float4 w = unpack4_bf16(*reinterpret_cast<const uint2*>(e.ln_w + n));
      *reinterpret_cast<uint2*>(e.blk_out + roff + n) =
          pack4_bf16(w.x * bf16_round(x.x * rstd), w.y * bf16_round(x.y * rstd), w.z * bf16_round(x.z * rstd),
                     w.w * bf16_round(x.w * rstd));
    }
  }
}

}  // namespace dfl
