// The row-wise epilogue that runs INSIDE the streaming GEMM (gemm_skinny.cuh, kModeSwiglu) on a finished output tile
// staged in shared memory as tile[m][n] (m = activation row, n = column inside the tile), one warp per activation
// row: SwiGLU with the reference's bf16 rounding points, written as the down projection's bf16 operand, so the
// widest fp32 plane of the step (gate/up: 2 x intermediate columns) never reaches HBM.
// (The same treatment of fc / qkv / o / down was built and measured slower than consumer kernels at every batch
// width -- DESIGN.md section 7 -- because only a tile's finishing CTA does the row work.)
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

}  // namespace dfl
