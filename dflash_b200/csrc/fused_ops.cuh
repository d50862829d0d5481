// The small kernels left between the streaming GEMMs. Everything that works on a finished output tile now runs in
// the GEMM's own epilogue (epilogues.cuh); what remains needs a whole activation row: the RMSNorm scale
// (Qwen3RMSNorm: fp32 inside, -> bf16, * weight) over the bf16 rows and per-tile sums of squares the GEMM left, and
// the block embedding. Latency-bound (<= 0.3 MB each): 128-bit accesses, one L2 round trip, no block-wide barrier
// in the norm pass.
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
// RMSNorm pass over rows whose bf16 values and per-tile sums of squares were written by a kModeRows GEMM:
//   out = w * bf16(x * rsqrt(mean(x^2) + eps))            (Qwen3RMSNorm, model/dflash.py:115,141,177,189)
// One CTA per row, 16 bytes per thread per iteration; every warp adds the (<= 64) per-tile sums itself in a fixed
// order, so the pass is one L2 round trip with no block barrier.
struct NormArgs {
  const __nv_bfloat16* x;   // [rows][H]
  const float* tile_ss;     // [H / 128][ss_ld]
  int ss_ld;
  int H;
  const __nv_bfloat16* w;   // [H]
  __nv_bfloat16* out;       // [rows][H]
  float eps;
  const int* ctx_len;       // optional: row (r, j) = (row / SL, row % SL) is live iff j < ctx_len[r]
  int SL;
};

constexpr int kNormThreads = 512;

__device__ __forceinline__ void norm_row_body(const NormArgs& a, int row, int tid) {
  const int lane = tid & 31;
  const int n0 = tid * 8;
  // the norm weights do not depend on the producing GEMM: requested before griddepcontrol.wait
  uint4 wv = make_uint4(0, 0, 0, 0);
  if (n0 < a.H) wv = *reinterpret_cast<const uint4*>(a.w + n0);
  DFL_WAIT_THEN_TRIGGER();
  if (a.ctx_len != nullptr && (row % a.SL) >= a.ctx_len[row / a.SL]) return;
  const long long roff = static_cast<long long>(row) * a.H;
  uint4 xv = make_uint4(0, 0, 0, 0);
  if (n0 < a.H) xv = __ldcg(reinterpret_cast<const uint4*>(a.x + roff + n0));
  const int nt = a.H / 128;
  float s = 0.f;
  for (int t = lane; t < nt; t += 32) s += __ldcg(a.tile_ss + static_cast<long long>(t) * a.ss_ld + row);
  const float tot = warp_sum(s);
  const float rstd = 1.0f / sqrtf(tot / static_cast<float>(a.H) + a.eps);
  for (int n = n0; n < a.H; n += kNormThreads * 8) {
    if (n != n0) {
      wv = *reinterpret_cast<const uint4*>(a.w + n);
      xv = __ldcg(reinterpret_cast<const uint4*>(a.x + roff + n));
    }
    const float4 x0 = unpack4_bf16(make_uint2(xv.x, xv.y)), x1 = unpack4_bf16(make_uint2(xv.z, xv.w));
    const float4 w0 = unpack4_bf16(make_uint2(wv.x, wv.y)), w1 = unpack4_bf16(make_uint2(wv.z, wv.w));
    const uint2 o0 = pack4_bf16(w0.x * bf16_round(x0.x * rstd), w0.y * bf16_round(x0.y * rstd),
                                w0.z * bf16_round(x0.z * rstd), w0.w * bf16_round(x0.w * rstd));
    const uint2 o1 = pack4_bf16(w1.x * bf16_round(x1.x * rstd), w1.y * bf16_round(x1.y * rstd),
                                w1.z * bf16_round(x1.z * rstd), w1.w * bf16_round(x1.w * rstd));
    *reinterpret_cast<uint4*>(a.out + roff + n) = make_uint4(o0.x, o0.y, o1.x, o1.y);
  }
}

__global__ void __launch_bounds__(kNormThreads) norm_rows_kernel(const NormArgs a) {
  norm_row_body(a, blockIdx.x, threadIdx.x);
  DFL_TRACE(2);
}

// =============================================================================================
// Block rows: embedding gather -> residual stream, then the first input_layernorm
//   x = embed_tokens(block_ids)  (model/dflash.py:237), out = ln1_0(x)  (:115)
struct EmbedArgs {
  const __nv_bfloat16* embed;  // [vocab][H], or (ids == null) a [rows][H] matrix of already-embedded rows
  const long long* ids;        // [R][ids_ld]
  int ids_ld;
  long long pad_token;         // used for slots >= bs
  int bs, H, SL;
  __nv_bfloat16* resid;        // [rows][H] residual stream (written)
  const __nv_bfloat16* norm_w; // [H]
  __nv_bfloat16* out;          // [rows][H]
  float eps;
};

// One CTA (kNormThreads threads) per row; the row's sum of squares needs one block reduction here.
__device__ __forceinline__ void embed_row_body(const EmbedArgs& a, int row, int tid, float* red) {
  const int n0 = tid * 8;
  uint4 wv = make_uint4(0, 0, 0, 0);
  if (n0 < a.H) wv = *reinterpret_cast<const uint4*>(a.norm_w + n0);
  DFL_WAIT_THEN_TRIGGER();
  const int r = row / a.SL, i = row % a.SL;
  long long tok;
  if (a.ids == nullptr) tok = row;  // forward(noise_embedding=...): the rows are given
  else tok = (i < a.bs) ? a.ids[static_cast<long long>(r) * a.ids_ld + i] : a.pad_token;
  const long long roff = static_cast<long long>(row) * a.H;
  // H <= 8192: at most two 16-byte groups per thread, kept in registers across the reduction
  uint4 xv[2];
  float ss = 0.f;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int n = n0 + g * kNormThreads * 8;
    xv[g] = make_uint4(0, 0, 0, 0);
    if (n < a.H) {
      xv[g] = *reinterpret_cast<const uint4*>(a.embed + tok * a.H + n);
      *reinterpret_cast<uint4*>(a.resid + roff + n) = xv[g];
      const float4 x0 = unpack4_bf16(make_uint2(xv[g].x, xv[g].y)), x1 = unpack4_bf16(make_uint2(xv[g].z, xv[g].w));
      ss += x0.x * x0.x + x0.y * x0.y + x0.z * x0.z + x0.w * x0.w + x1.x * x1.x + x1.y * x1.y + x1.z * x1.z + x1.w * x1.w;
    }
  }
  ss = warp_sum(ss);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kNormThreads / 32; ++w) tot += red[w];
  const float rstd = 1.0f / sqrtf(tot / static_cast<float>(a.H) + a.eps);
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int n = n0 + g * kNormThreads * 8;
    if (n >= a.H) continue;
    if (g == 1) wv = *reinterpret_cast<const uint4*>(a.norm_w + n);
    const float4 x0 = unpack4_bf16(make_uint2(xv[g].x, xv[g].y)), x1 = unpack4_bf16(make_uint2(xv[g].z, xv[g].w));
    const float4 w0 = unpack4_bf16(make_uint2(wv.x, wv.y)), w1 = unpack4_bf16(make_uint2(wv.z, wv.w));
    const uint2 o0 = pack4_bf16(w0.x * bf16_round(x0.x * rstd), w0.y * bf16_round(x0.y * rstd),
                                w0.z * bf16_round(x0.z * rstd), w0.w * bf16_round(x0.w * rstd));
    const uint2 o1 = pack4_bf16(w1.x * bf16_round(x1.x * rstd), w1.y * bf16_round(x1.y * rstd),
                                w1.z * bf16_round(x1.z * rstd), w1.w * bf16_round(x1.w * rstd));
    *reinterpret_cast<uint4*>(a.out + roff + n) = make_uint4(o0.x, o0.y, o1.x, o1.y);
  }
}

// The step's first small kernel: CTAs [0, rows0) finish the context injection (hidden_norm over the fc GEMM's bf16
// rows, model/dflash.py:177), the rest embed the block and apply layer 0's input_layernorm. CTA b also fills row b of
// the step's position / rotary table (Qwen3RotaryEmbedding, model/dflash.py:178), which every layer's QKV epilogue
// reads: request state is only read after griddepcontrol.wait (the verify kernel of the previous cycle wrote it).
__global__ void __launch_bounds__(kNormThreads) rows_pre_kernel(const NormArgs c, const EmbedArgs e, const RopeTableArgs t,
                                                               const int rows0) {
  __shared__ float red[kNormThreads / 32];
  if (static_cast<int>(blockIdx.x) < rows0) norm_row_body(c, blockIdx.x, threadIdx.x);
  else embed_row_body(e, blockIdx.x - rows0, threadIdx.x, red);
  rope_table_row(t, blockIdx.x, threadIdx.x);
  DFL_TRACE(2);
}

}  // namespace dfl
