// C-ABI of libdflash_b200.so (declared in include/dflash_b200.h).
#include "../../include/dflash_b200.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "engine.cuh"

namespace dfl {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return DFLASH_ERR_CUDA;
}
}  // namespace dfl

using namespace dfl;

extern "C" {

int dflash_abi_version(void) { return DFLASH_ABI_VERSION; }

const char* dflash_last_error(void) { return g_err; }

int dflash_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
  if (prop.major != 10) {
    set_error("dflash_b200 needs an sm_100a device (found sm_%d%d)", prop.major, prop.minor);
    return DFLASH_ERR_ARCH;
  }
  return prop.multiProcessorCount;
}

int dflash_gemm_max_slots(int N, int K, int grid) {
  if (N <= 0 || K <= 0 || K % kTileK != 0 || grid <= 0) return DFLASH_ERR_ARG;
  const int n_tiles = (N + kTileN - 1) / kTileN;
  const int kb = K / kTileK;
  const long long T = static_cast<long long>(n_tiles) * kb;
  const int g = static_cast<long long>(grid) < T ? grid : static_cast<int>(T);
  return max_slots_for(n_tiles, kb, g);
}

int dflash_gemm_skinny(const void* W, int w_rows_total, int w_row0, int N, int K, const void* X,
                       int x_rows_total, int x_row0, int mb, int m_valid, float* ws, int ws_rows,
                       long long ws_ld, float* out, long long out_ld, int grid, int use_pdl,
                       void* stream) {
  if (!W || !X || !ws || !out) { set_error("gemm_skinny: null pointer"); return DFLASH_ERR_ARG; }
  GemmPlan p;
  int rc = make_gemm_plan(&p, W, w_rows_total, w_row0, N, K, X, x_rows_total, x_row0, mb, m_valid,
                          kModePartials, grid);
  if (rc) return DFLASH_ERR_ARG;
  p.args.ws = ws;
  p.args.ws_rows = ws_rows;
  p.args.ws_ld = ws_ld;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = launch_gemm(p, st, use_pdl != 0);
  if (e != cudaSuccess) return cuda_fail(e, "gemm_skinny launch");
  e = launch_sum_slots(p, out, out_ld, st);
  if (e != cudaSuccess) return cuda_fail(e, "sum_slots launch");
  return DFLASH_OK;
}

int dflash_gemm_argmax(const void* W, int w_rows_total, int N, int K, const void* X, int x_rows_total,
                       int x_row0, int mb, int m_valid, float* cand_val, int* cand_idx, void* logits,
                       long long logits_ld, long long* tokens_out, int grid, int use_pdl,
                       void* stream) {
  if (!W || !X || !cand_val || !cand_idx || !tokens_out) {
    set_error("gemm_argmax: null pointer");
    return DFLASH_ERR_ARG;
  }
  GemmPlan p;
  int rc = make_gemm_plan(&p, W, w_rows_total, 0, N, K, X, x_rows_total, x_row0, mb, m_valid,
                          kModeArgmax, grid);
  if (rc) return DFLASH_ERR_ARG;
  p.args.cand_val = cand_val;
  p.args.cand_idx = cand_idx;
  p.args.logits = static_cast<__nv_bfloat16*>(logits);
  p.args.logits_ld = logits_ld;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = launch_gemm(p, st, use_pdl != 0);
  if (e != cudaSuccess) return cuda_fail(e, "gemm_argmax launch");
  e = launch_reduce_candidates(cand_val, cand_idx, p.grid, mb, m_valid, tokens_out, st);
  if (e != cudaSuccess) return cuda_fail(e, "reduce_candidates launch");
  return DFLASH_OK;
}

}  // extern "C"
