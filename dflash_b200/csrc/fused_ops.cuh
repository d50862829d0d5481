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

}  // namespace dfl
