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

struct dflash_engine {
  Engine* impl;
};

static int device_sm_count(int* sms) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
  if (major != 10) {
    set_error("dflash_b200 needs an sm_100a device (compute capability major %d found)", major);
    return DFLASH_ERR_ARCH;
  }
  e = cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
  return DFLASH_OK;
}

size_t dflash_workspace_bytes(const dflash_config_t* cfg) {
  if (!cfg) { set_error("null config"); return 0; }
  if (check_config(*cfg)) return 0;
  int sms = 148;
  if (cfg->gemm_grid <= 0) {
    if (device_sm_count(&sms)) return 0;
  }
  Region reg[DFLASH_BUF_COUNT];
  return layout_workspace(*cfg, reg, sms);
}

int dflash_engine_create(const dflash_config_t* cfg, const dflash_weights_t* weights, void* workspace,
                         size_t workspace_bytes, dflash_engine_t** out) {
  if (!cfg || !weights || !out || !weights->layers_host) {
    set_error("engine_create: null argument");
    return DFLASH_ERR_ARG;
  }
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc) return rc;
  Engine* impl = nullptr;
  rc = engine_create(*cfg, *weights, workspace, workspace_bytes, sms, &impl);
  if (rc) return rc;
  *out = new dflash_engine{impl};
  return DFLASH_OK;
}

void dflash_engine_destroy(dflash_engine_t* e) {
  if (!e) return;
  delete e->impl;
  delete e;
}

int dflash_engine_buffer(const dflash_engine_t* e, int id, void** ptr_out, size_t* bytes_out) {
  if (!e || id < 0 || id >= DFLASH_BUF_COUNT) { set_error("engine_buffer: bad argument"); return DFLASH_ERR_ARG; }
  if (ptr_out) *ptr_out = e->impl->base + e->impl->reg[id].off;
  if (bytes_out) *bytes_out = e->impl->reg[id].bytes;
  return DFLASH_OK;
}

int dflash_prefill_context(dflash_engine_t* e, int r, const void* const* hidden, int P, void* stream) {
  if (!e || !hidden) { set_error("prefill_context: null argument"); return DFLASH_ERR_ARG; }
  return enqueue_prefill(e->impl, r, hidden, P, 0, static_cast<cudaStream_t>(stream));
}

int dflash_prefill_context_at(dflash_engine_t* e, int r, const void* const* hidden, int n_rows, int pos0,
                              void* stream) {
  if (!e || !hidden) { set_error("prefill_context_at: null argument"); return DFLASH_ERR_ARG; }
  return enqueue_prefill(e->impl, r, hidden, n_rows, pos0, static_cast<cudaStream_t>(stream));
}

int dflash_draft_step(dflash_engine_t* e, const void* noise_embedding, int run_lm_head, void* stream) {
  if (!e) { set_error("draft_step: null engine"); return DFLASH_ERR_ARG; }
  return enqueue_draft_step(e->impl, noise_embedding, run_lm_head != 0, static_cast<cudaStream_t>(stream));
}

int dflash_verify_step(dflash_engine_t* e, const void* target_logits, long long logits_ld,
                       const long long* posterior_in, const void* const* hidden, float temperature,
                       const float* noise, unsigned long long seed, const long long* stop_ids, int n_stop,
                       const int* forced_k, int forced_ld, int clamp_tail, void* stream) {
  if (!e || !hidden || (!target_logits && !posterior_in)) {
    set_error("verify_step: null argument");
    return DFLASH_ERR_ARG;
  }
  VerifyInputs v;
  memset(&v, 0, sizeof(v));
  v.target_logits = target_logits;
  v.logits_ld = logits_ld;
  v.posterior_in = posterior_in;
  for (int s = 0; s < e->impl->nsel; ++s) {
    if (!hidden[s]) { set_error("verify_step: hidden[%d] is null", s); return DFLASH_ERR_ARG; }
    v.hidden[s] = hidden[s];
  }
  v.temperature = temperature;
  v.noise = noise;
  v.seed = seed;
  v.stop_ids = stop_ids;
  v.n_stop = stop_ids ? n_stop : 0;
  v.forced_k = forced_k;
  v.forced_ld = forced_ld;
  v.clamp_tail = clamp_tail;
  return enqueue_verify_step(e->impl, v, static_cast<cudaStream_t>(stream));
}

int dflash_verify_inject_step(dflash_engine_t* e, const void* target_logits, long long logits_ld,
                              const long long* posterior_in, const void* const* hidden, float temperature,
                              const float* noise, unsigned long long seed, const long long* stop_ids, int n_stop,
                              const int* forced_k, int forced_ld, int clamp_tail, void* stream) {
  if (!e || !hidden || (!target_logits && !posterior_in)) {
    set_error("verify_inject_step: null argument");
    return DFLASH_ERR_ARG;
  }
  VerifyInputs v;
  memset(&v, 0, sizeof(v));
  v.target_logits = target_logits;
  v.logits_ld = logits_ld;
  v.posterior_in = posterior_in;
  for (int s = 0; s < e->impl->nsel; ++s) {
    if (!hidden[s]) { set_error("verify_inject_step: hidden[%d] is null", s); return DFLASH_ERR_ARG; }
    v.hidden[s] = hidden[s];
  }
  v.temperature = temperature;
  v.noise = noise;
  v.seed = seed;
  v.stop_ids = stop_ids;
  v.n_stop = stop_ids ? n_stop : 0;
  v.forced_k = forced_k;
  v.forced_ld = forced_ld;
  v.clamp_tail = clamp_tail;
  return enqueue_verify_step(e->impl, v, static_cast<cudaStream_t>(stream), true);
}

int dflash_draft_step_injected(dflash_engine_t* e, int run_lm_head, float temperature, unsigned long long seed,
                               void* stream) {
  if (!e) { set_error("draft_step_injected: null engine"); return DFLASH_ERR_ARG; }
  if (temperature >= 1e-5f && !e->impl->has_sample) {
    set_error("draft_step_injected: the sampling epilogue needs max_requests * row slots <= 32");
    return DFLASH_ERR_ARG;
  }
  return enqueue_draft_step(e->impl, nullptr, run_lm_head != 0, static_cast<cudaStream_t>(stream), 1, 0, temperature,
                            seed, false);
}

int dflash_embed_block(dflash_engine_t* e, void* stream) {
  if (!e) { set_error("embed_block: null engine"); return DFLASH_ERR_ARG; }
  return enqueue_embed_block(e->impl, static_cast<cudaStream_t>(stream));
}

int dflash_engine_launches(const dflash_engine_t* e, int which) {
  if (!e) { set_error("engine_launches: null engine"); return DFLASH_ERR_ARG; }
  const Engine* en = e->impl;
  const int per_layer = en->nsplit_attn > 1 ? 9 : 8;  // (a single KV split needs no merge kernel)
  switch (which) {
    case 0: return 1 + per_layer * en->L + 1;  // dflash_draft_step
    case 1: return per_layer * en->L + 1;      // dflash_draft_step_injected
    case 2: return 1;                          // dflash_verify_step
    case 3: return 2;                          // dflash_verify_inject_step
    default: set_error("engine_launches: which must be 0..3"); return DFLASH_ERR_ARG;
  }
}

int dflash_draft_step_sampled(dflash_engine_t* e, float temperature, unsigned long long seed, void* stream) {
  if (!e) { set_error("draft_step_sampled: null engine"); return DFLASH_ERR_ARG; }
  if (temperature >= 1e-5f && !e->impl->has_sample) {
    set_error("draft_step_sampled: the sampling epilogue needs max_requests * row slots <= 32");
    return DFLASH_ERR_ARG;
  }
  return enqueue_draft_step(e->impl, nullptr, true, static_cast<cudaStream_t>(stream), 1, 0, temperature, seed);
}

int dflash_gemm_sample(const void* W, int w_rows_total, int N, int K, const void* X, int x_rows_total, int x_row0,
                       int mb, int m_valid, float temperature, unsigned long long seed, unsigned long long step,
                       float* cand_val, int* cand_idx, long long* tokens_out, int grid, int use_pdl, void* stream) {
  if (!W || !X || !cand_val || !cand_idx || !tokens_out || temperature < 1e-5f || m_valid > mb || mb > 32) {
    set_error("gemm_sample: bad argument (needs temperature > 0, m_valid <= mb <= 32)");
    return DFLASH_ERR_ARG;
  }
  GemmPlan p;
  int rc = make_gemm_plan(&p, W, w_rows_total, 0, N, K, X, x_rows_total, x_row0, mb, m_valid, kModeSample, grid);
  if (rc) return DFLASH_ERR_ARG;
  p.args.cand_val = cand_val;
  p.args.cand_idx = cand_idx;
  p.args.inv_temp = 1.0f / temperature;
  p.args.seed = seed;
  p.args.step_base = step;
  p.args.rng_step = nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = launch_gemm(p, st, use_pdl != 0);
  if (e != cudaSuccess) return cuda_fail(e, "gemm_sample launch");
  e = launch_reduce_candidates(cand_val, cand_idx, p.grid, p.args.cand_ld, m_valid, tokens_out, st);
  if (e != cudaSuccess) return cuda_fail(e, "reduce_candidates launch");
  return DFLASH_OK;
}

int dflash_draft_step_candidates(dflash_engine_t* e, int n_candidates, int fixed_prefix_len, void* stream) {
  if (!e) { set_error("draft_step_candidates: null engine"); return DFLASH_ERR_ARG; }
  if (n_candidates < 2 || n_candidates > e->impl->max_cand || fixed_prefix_len < 0) {
    set_error("draft_step_candidates: n_candidates %d outside [2, %d] (dflash_config_t.max_candidates)", n_candidates,
              e->impl->max_cand);
    return DFLASH_ERR_ARG;
  }
  return enqueue_draft_step(e->impl, nullptr, true, static_cast<cudaStream_t>(stream), n_candidates, fixed_prefix_len);
}

int dflash_verify_step_candidates(dflash_engine_t* e, int n_candidates, const void* target_logits, long long logits_ld,
                                  const void* const* hidden, float temperature, const float* noise,
                                  unsigned long long seed, const long long* stop_ids, int n_stop, int clamp_tail,
                                  void* stream) {
  if (!e || !hidden || !target_logits) { set_error("verify_step_candidates: null argument"); return DFLASH_ERR_ARG; }
  if (n_candidates < 2 || n_candidates > e->impl->max_cand) {
    set_error("verify_step_candidates: n_candidates %d outside [2, %d]", n_candidates, e->impl->max_cand);
    return DFLASH_ERR_ARG;
  }
  VerifyInputs v;
  memset(&v, 0, sizeof(v));
  v.target_logits = target_logits;
  v.logits_ld = logits_ld;
  for (int s = 0; s < e->impl->nsel; ++s) {
    if (!hidden[s]) { set_error("verify_step_candidates: hidden[%d] is null", s); return DFLASH_ERR_ARG; }
    v.hidden[s] = hidden[s];
  }
  v.temperature = temperature;
  v.noise = noise;
  v.seed = seed;
  v.stop_ids = stop_ids;
  v.n_stop = stop_ids ? n_stop : 0;
  v.clamp_tail = clamp_tail;
  v.n_candidates = n_candidates;
  return enqueue_verify_step(e->impl, v, static_cast<cudaStream_t>(stream));
}

int dflash_sample(const void* logits, long long logits_ld, int rows, int vocab, float temperature,
                  const float* noise, unsigned long long seed, float* scratch_val, int* scratch_idx,
                  int nsplit, long long* tokens_out, void* stream) {
  if (!logits || !scratch_val || !scratch_idx || !tokens_out || rows < 1 || vocab < 1 || nsplit < 1) {
    set_error("sample: bad argument");
    return DFLASH_ERR_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PosteriorArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.logits = static_cast<const __nv_bfloat16*>(logits);
  pa.ld = logits_ld;
  pa.rows = rows;
  pa.V = vocab;
  pa.nsplit = nsplit;
  pa.inv_temp = temperature < 1e-5f ? 0.f : 1.0f / temperature;
  pa.noise = noise;
  pa.seed = seed;
  pa.rng_step = nullptr;
  pa.cand_val = scratch_val;
  pa.cand_idx = scratch_idx;
  posterior_kernel<<<dim3(nsplit, rows), 256, 0, st>>>(pa);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "posterior launch");
  // second stage: candidates are laid out [rows][nsplit] -> reuse the lm_head reducer with mb = 1 stride trick
  sample_reduce_kernel<<<rows, 32, 0, st>>>(scratch_val, scratch_idx, nsplit, tokens_out);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "sample reduce launch");
  return DFLASH_OK;
}

int dflash_gemm_argmax_grid(int N, int grid) {
  if (N <= 0 || grid <= 0) return DFLASH_ERR_ARG;
  return balanced_tile_grid((N + kTileN - 1) / kTileN, grid);
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
  if (ws_rows < p.groups * p.mb) { set_error("gemm_skinny: ws_rows %d < %d", ws_rows, p.groups * p.mb); return DFLASH_ERR_ARG; }
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

int dflash_gemm_swiglu(const void* Wgu, int I, int K, const void* X, int x_rows_total, int mb, int m_valid, void* out,
                       long long ld, float* part, unsigned int* flags, int grid, int use_pdl, void* stream) {
  if (!Wgu || !X || !out || !part || !flags || I % 64 != 0) { set_error("gemm_swiglu: bad argument"); return DFLASH_ERR_ARG; }
  GemmPlan p;
  int rc = make_gemm_plan(&p, Wgu, 2ll * I, 0, 2 * I, K, X, x_rows_total, 0, mb, m_valid, kModeSwiglu, grid);
  if (rc) return DFLASH_ERR_ARG;
  p.args.part = part;
  p.args.flags = flags;
  p.args.sw.out = static_cast<__nv_bfloat16*>(out);
  p.args.sw.ld = ld;
  cudaError_t e = launch_gemm(p, static_cast<cudaStream_t>(stream), use_pdl != 0);
  if (e != cudaSuccess) return cuda_fail(e, "gemm_swiglu launch");
  return DFLASH_OK;
}

// Debug: dflash_gemm_skinny with per-CTA phase timestamps (not part of the reference-facing surface; scripts/gemm_trace.py)
int dflash_gemm_trace(const void* W, int N, int K, const void* X, int x_rows_total, int mb, int m_valid, float* ws,
                      int ws_rows, unsigned long long* trace, int grid, void* stream) {
  GemmPlan p;
  int rc = make_gemm_plan(&p, W, N, 0, N, K, X, x_rows_total, 0, mb, m_valid, kModePartials, grid);
  if (rc) return DFLASH_ERR_ARG;
  p.args.ws = ws;
  p.args.ws_rows = ws_rows;
  p.args.ws_ld = N;
  p.args.trace = trace;
  cudaError_t e = launch_gemm(p, static_cast<cudaStream_t>(stream), false);
  if (e != cudaSuccess) return cuda_fail(e, "gemm_trace launch");
  return p.grid;
}

#ifdef DFLASH_STEP_TRACE
// Debug build only: device buffer for the in-step timeline (ptx.cuh). cap = number of (tag, time) records.
int dflash_step_trace_set(unsigned long long* buf, unsigned int cap) {
  unsigned int zero = 0;
  cudaError_t e = cudaMemcpyToSymbol(g_trace_buf, &buf, sizeof(buf));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(zero));
  return e == cudaSuccess ? DFLASH_OK : cuda_fail(e, "step_trace_set");
}
#endif

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
                          logits ? kModeArgmaxDump : kModeArgmax, grid);
  if (rc) return DFLASH_ERR_ARG;
  p.args.cand_val = cand_val;
  p.args.cand_idx = cand_idx;
  p.args.logits = static_cast<__nv_bfloat16*>(logits);
  p.args.logits_ld = logits_ld;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = launch_gemm(p, st, use_pdl != 0);
  if (e != cudaSuccess) return cuda_fail(e, "gemm_argmax launch");
  e = launch_reduce_candidates(cand_val, cand_idx, p.grid, p.args.cand_ld, m_valid, tokens_out, st);
  if (e != cudaSuccess) return cuda_fail(e, "reduce_candidates launch");
  return DFLASH_OK;
}

}  // extern "C"
