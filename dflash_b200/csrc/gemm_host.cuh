// Host side of the skinny GEMM: TMA tensor maps (driver entry point fetched through the runtime,
// so the library does not link libcuda) and the PDL launch.
#pragma once
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gemm_skinny.cuh"

namespace dfl {

void set_error(const char* fmt, ...);  // api.cu

inline PFN_cuTensorMapEncodeTiled get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || p == nullptr) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return nullptr;
  }
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  return fn;
}

// Row-major bf16 matrix [rows, cols] (cols contiguous, row pitch ld elements) tiled as
// [box_rows x 64] boxes with the 128-byte swizzle the UMMA descriptors in ptx.cuh expect.
inline int make_tmap_bf16(CUtensorMap* out, const void* base, long long rows, long long cols,
                          long long ld, int box_rows) {
  PFN_cuTensorMapEncodeTiled fn = get_encode_fn();
  if (!fn) return -2;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0) {
    set_error("tensor map: base/pitch must be 16-byte aligned");
    return -1;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kTileK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: %d (rows=%lld cols=%lld ld=%lld box=%d)",
              static_cast<int>(r), rows, cols, ld, box_rows);
    return -3;
  }
  return 0;
}

// One selected target hidden state [n_req * rows_per_req, hidden] bf16 seen as [n_req][rows_per_req][hidden], tiled as
// [box_req x SL x 64] boxes (128-byte swizzle): box row (r, j) is activation row r * SL + j of the a_in layout; slots
// j >= rows_per_req and requests past n_req are out of bounds and therefore zero-filled.
inline int make_tmap_hidden3d(CUtensorMap* out, const void* base, int n_req, int rows_per_req, long long hidden,
                              int SL, int box_req) {
  PFN_cuTensorMapEncodeTiled fn = get_encode_fn();
  if (!fn) return -2;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (hidden * 2) % 16 != 0) {
    set_error("tensor map: hidden-state base/pitch must be 16-byte aligned");
    return -1;
  }
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(hidden), static_cast<cuuint64_t>(rows_per_req),
                        static_cast<cuuint64_t>(n_req)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(hidden) * 2, static_cast<cuuint64_t>(hidden) * 2 * rows_per_req};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(kTileK), static_cast<cuuint32_t>(SL), static_cast<cuuint32_t>(box_req)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (hidden state, 3-D) failed: %d (req=%d rows=%d hidden=%lld SL=%d box_req=%d)",
              static_cast<int>(r), n_req, rows_per_req, hidden, SL, box_req);
    return -3;
  }
  return 0;
}

struct GemmPlan {
  CUtensorMap tmW;
  CUtensorMap tmX;
  XMaps xm;   // kModeCtxNorm direct mode only
  GemmArgs args;
  int mb;     // UMMA N (padded activation rows): 16, 32, 64, 128, 256
  int mode;   // GemmMode
  int grid;       // weight ranges (gridDim.y); the launch has groups * grid CTAs
  int groups;     // column groups of mb activation rows (gridDim.x)
  int max_slots;  // partial slots the consumer must allocate
};

// Largest number of partial slots any tile can get when T units are cut into G ranges.
inline int max_slots_for(int n_tiles, int k_blocks, int grid) {
  const long long T = static_cast<long long>(n_tiles) * k_blocks;
  int mx = 1;
  for (int t = 0; t < n_tiles; ++t) {
    int s = tile_num_slots(t, k_blocks, T, grid);
    if (s > mx) mx = s;
  }
  return mx;
}

template <int MB, int MODE>
inline cudaError_t launch_gemm_t(const GemmPlan& p, cudaStream_t stream, bool pdl) {
  using Cfg = GemmCfg<MB, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e;
    if constexpr (MODE == kModeCtxNorm)
      e = cudaFuncSetAttribute(gemm_ctx_kernel<MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    else
      e = cudaFuncSetAttribute(gemm_skinny_kernel<MB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(p.groups, p.grid);
  cfg.blockDim = dim3(Cfg::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  if constexpr (MODE == kModeCtxNorm) return cudaLaunchKernelEx(&cfg, gemm_ctx_kernel<MB>, p.tmW, p.tmX, p.xm, p.args);
  else return cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<MB, MODE>, p.tmW, p.tmX, p.args);
}

template <int MODE>
inline cudaError_t launch_gemm_any_mb(const GemmPlan& p, cudaStream_t stream, bool pdl) {
  switch (p.mb) {
    case 16: return launch_gemm_t<16, MODE>(p, stream, pdl);
    case 32: return launch_gemm_t<32, MODE>(p, stream, pdl);
    case 64: return launch_gemm_t<64, MODE>(p, stream, pdl);
    case 128: return launch_gemm_t<128, MODE>(p, stream, pdl);
    case 256: return launch_gemm_t<256, MODE>(p, stream, pdl);
    default: return cudaErrorInvalidValue;
  }
}

inline cudaError_t launch_gemm(const GemmPlan& p, cudaStream_t stream, bool pdl) {
  if (p.mode == kModeSwiglu) return launch_gemm_any_mb<kModeSwiglu>(p, stream, pdl);
  if (p.mode == kModeCtxNorm) return launch_gemm_any_mb<kModeCtxNorm>(p, stream, pdl);
  if (p.mode == kModeSample) {
    switch (p.mb) {
      case 16: return launch_gemm_t<16, kModeSample>(p, stream, pdl);
      case 32: return launch_gemm_t<32, kModeSample>(p, stream, pdl);
      default: return cudaErrorInvalidValue;
    }
  }
  if (p.mode == kModeTopK) {
    switch (p.mb) {
      case 16: return launch_gemm_t<16, kModeTopK>(p, stream, pdl);
      case 32: return launch_gemm_t<32, kModeTopK>(p, stream, pdl);
      default: return cudaErrorInvalidValue;
    }
  }
  if (p.mode == kModeArgmaxDump) {
    switch (p.mb) {
      case 16: return launch_gemm_t<16, kModeArgmaxDump>(p, stream, pdl);
      case 32: return launch_gemm_t<32, kModeArgmaxDump>(p, stream, pdl);
      case 64: return launch_gemm_t<64, kModeArgmaxDump>(p, stream, pdl);
      case 128: return launch_gemm_t<128, kModeArgmaxDump>(p, stream, pdl);
      case 256: return launch_gemm_t<256, kModeArgmaxDump>(p, stream, pdl);
      default: return cudaErrorInvalidValue;
    }
  }
  if (p.mode == kModeArgmax) {
    switch (p.mb) {
      case 16: return launch_gemm_t<16, kModeArgmax>(p, stream, pdl);
      case 32: return launch_gemm_t<32, kModeArgmax>(p, stream, pdl);
      case 64: return launch_gemm_t<64, kModeArgmax>(p, stream, pdl);
      case 128: return launch_gemm_t<128, kModeArgmax>(p, stream, pdl);
      case 256: return launch_gemm_t<256, kModeArgmax>(p, stream, pdl);
      default: return cudaErrorInvalidValue;
    }
  }
  switch (p.mb) {
    case 16: return launch_gemm_t<16, kModePartials>(p, stream, pdl);
    case 32: return launch_gemm_t<32, kModePartials>(p, stream, pdl);
    case 64: return launch_gemm_t<64, kModePartials>(p, stream, pdl);
    case 128: return launch_gemm_t<128, kModePartials>(p, stream, pdl);
    case 256: return launch_gemm_t<256, kModePartials>(p, stream, pdl);
    default: return cudaErrorInvalidValue;
  }
}

// Weight ranges a GEMM with `groups` column groups is cut into: the groups of one range run side by side,
// so ranges * groups ~ one wave of CTAs.
inline int ranges_for(int grid, int groups) { return grid / groups > 0 ? grid / groups : 1; }

// CTAs for a whole-tile GEMM: the smallest grid whose busiest CTA has as many tiles as with `grid` CTAs.
inline int balanced_tile_grid(int n_tiles, int grid) {
  int g = grid < n_tiles ? grid : n_tiles;
  if (g < 1) g = 1;
  const int per = (n_tiles + g - 1) / g;
  return (n_tiles + per - 1) / per;
}

// Fill a plan. W: [w_rows_total, K] bf16 (pitch K); the GEMM covers weight rows [w_row0, w_row0+N).
// kModeSwiglu: W is the [gate; up] stack, N = 2 * intermediate; tile t = gate rows [64t, 64t+64) + up rows
// [I + 64t, I + 64t + 64), so there are I / 64 tiles and the weight map has 64-row boxes.
// X: [x_rows_total, K] bf16 (pitch K); activation rows [x_row0, x_row0+groups*mb) feed the MMA in `groups`
// slabs of mb rows (groups = ceil(m_valid / mb)).
inline int make_gemm_plan(GemmPlan* p, const void* W, long long w_rows_total, int w_row0, int N, int K,
                          const void* X, int x_rows_total, int x_row0, int mb, int m_valid, int mode,
                          int grid) {
  memset(p, 0, sizeof(*p));
  if (K % kTileK != 0) { set_error("gemm: K=%d must be a multiple of %d", K, kTileK); return -1; }
  if (!(mb == 16 || mb == 32 || mb == 64 || mb == 128 || mb == 256)) {
    set_error("gemm: mb=%d unsupported", mb);
    return -1;
  }
  const int groups = m_valid > mb ? (m_valid + mb - 1) / mb : 1;
  if (x_row0 + groups * mb > x_rows_total) { set_error("gemm: activation buffer too small"); return -1; }
  grid = ranges_for(grid, groups);
  if (groups > 1 && (mode == kModePartials || mode == kModeCtxNorm)) {
    // Wide batches are tensor-bound and their consumers pay for every fp32 partial plane (64 streams: 150 MB per QKV
    // launch): when a range count that divides the tile count exists within 20 % of the full one, cut on tile
    // boundaries -- every tile then has ONE slot (4 x 32 instead of 4 x 37 ranges for the 32 tiles of o / down).
    const int nt = (N + kTileN - 1) / kTileN;
    for (int g = grid; g * 5 >= grid * 4 && g >= 1; --g)
      if (nt % g == 0) { grid = g; break; }
  }
  if (mode_is_fused(mode) && N % kTileN != 0) {
    set_error("gemm: fused epilogues need N (%d) to be a multiple of %d", N, kTileN);
    return -1;
  }
  int rc = make_tmap_bf16(&p->tmW, W, w_rows_total, K, K, mode == kModeSwiglu ? kTileN / 2 : kTileN);
  if (rc) return rc;
  rc = make_tmap_bf16(&p->tmX, X, x_rows_total, K, K, mb);
  if (rc) return rc;
  p->mb = mb;
  p->mode = mode;
  p->groups = groups;
  p->args.groups = groups;
  p->args.cand_ld = groups * mb;
  p->args.n_tiles = (N + kTileN - 1) / kTileN;
  p->args.k_blocks = K / kTileK;
  p->args.N = N;
  p->args.w_row0 = w_row0;
  p->args.x_row0 = x_row0;
  p->args.m_valid = m_valid;
  const long long T = static_cast<long long>(p->args.n_tiles) * p->args.k_blocks;
  if (mode == kModeSwiglu) p->args.sw.I = N / 2;
  if ((mode == kModePartials || mode == kModeCtxNorm) && (T + 1) * grid >= (1ll << 31)) {
    set_error("gemm: %lld work units x %d CTAs overflows the consumers' 32-bit slot arithmetic", T, grid);
    return -1;
  }
  if (mode_is_whole_tile(mode)) {
    // Whole 128-row tiles per CTA: 1187 vocab tiles over 148 CTAs would leave three CTAs with 9 tiles and the rest
    // with 8, and the kernel would last 9 tiles' time at 8/9 of the bandwidth (in-graph timeline: block 0 done 31 us
    // before the kernel). Take the smallest grid with the same maximum -- ceil(1187 / 9) = 132 CTAs of 9 tiles; the
    // stream stays HBM-bound with 132 SMs pulling (measured 694 -> 685 us per step).
    p->grid = balanced_tile_grid(p->args.n_tiles, grid);
    p->max_slots = 1;
  } else {
    p->grid = static_cast<long long>(grid) < T ? grid : static_cast<int>(T);
    p->max_slots = max_slots_for(p->args.n_tiles, p->args.k_blocks, p->grid);
  }
  return 0;
}

}  // namespace dfl
