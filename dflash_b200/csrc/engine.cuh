// Draft/verify engine: workspace layout, GEMM plans (TMA maps built once), per-step kernel schedule.
// The engine object is host memory only; every device byte belongs to the caller's workspace.
#pragma once
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "../../include/dflash_b200.h"
#include "attention.cuh"
#include "fused_ops.cuh"
#include "verify.cuh"

namespace dfl {
int cuda_fail(cudaError_t e, const char* what);  // api.cu

struct Region {
  size_t off = 0, bytes = 0;
};

struct Engine {
  dflash_config_t cfg;
  dflash_weights_t w;
  std::vector<dflash_layer_weights_t> layers;
  uint8_t* base = nullptr;
  size_t total = 0;
  Region reg[DFLASH_BUF_COUNT];
  int R, SL, RS, H, I, L, Hq, Hkv, V, nsel, bs, grid, nsplit_attn, nsplit_post, sm_count;
  int row_block_max_rows = DFLASH_ROW_BLOCK_MAX_ROWS;
  bool pdl;
  // plans
  GemmPlan fc;                  // ctx_feat -> partials
  // the same kernel reading the selected target hidden states in place (3-D maps, rebuilt when the pointers change),
  // launched right behind the verify kernel it overlaps (enqueue_verify_step with inject)
  GemmPlan fc_direct;
  const void* direct_hidden[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  std::vector<GemmPlan> qkv;    // a_in (ctx + block rows)
  // prompt pass (c = P rows at once): fc and the K/V rows of wqkv over the dedicated prompt buffers,
  // one plan per UMMA width so that short prompts do not pay for 256 columns
  GemmPlan fc_pf[5];            // mb = 16 << i
  std::vector<GemmPlan> kv_pf;  // [layer * 5 + i]
  std::vector<GemmPlan> o, gu, d;
  GemmPlan lm;
  GemmPlan lm_topk;             // same GEMM with the top-4 epilogue (multi-candidate drafting)
  GemmPlan lm_sample;           // same GEMM with the Gumbel-max epilogue (draft tokens sampled at temperature > 0)
  bool has_sample = false;
  int max_cand = 1;
  size_t flags_used = 0;        // bump allocator over DFLASH_BUF_FLAGS (one arrival counter per (group, tile) per plan)

  template <class T>
  T* buf(int id) const { return reinterpret_cast<T*>(base + reg[id].off); }
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kPrefillRows = 256;  // prompt rows projected per pass (one full-width UMMA)

// UMMA N (activation rows per MMA) for a GEMM over `rows` activation rows; more rows run as column groups.
inline int round_mb(int rows) {
  for (int mb : {16, 32, 64, 128, 256})
    if (rows <= mb) return mb;
  return 256;
}
inline int groups_of(int rows) { const int mb = round_mb(rows); return (rows + mb - 1) / mb; }
// rows an activation buffer is allocated with, so that a TMA box of mb rows never leaves it
inline int rows_padded(int rows) { return groups_of(rows) * round_mb(rows); }

// KV splits of the draft attention: enough (request, kv head, split) CTAs for about two waves, at most 16.
// One stream needs all 16 to fill the machine; 64 streams already give 512 CTAs and would only pay for the
// fp32 partials (16 splits x 1024 rows x 32 heads x 128 x 4 B = 268 MB per layer).
inline int default_attn_splits(const dflash_config_t& c, int sm_count) {
  if (c.attn_splits > 0) return c.attn_splits;
  const int SL = c.block_size <= 16 ? 16 : 32;
  const int ctas = c.max_requests * c.n_kv_heads * (SL / 16);
  // from about one CTA per SM up, ONE split: the attention kernel then writes the normalised output itself and the
  // merge kernel (one more full-grid dependency per layer) is not launched (16 streams, 128 CTAs: 1379 us per step
  // against 1405 / 1428 with 2 / 3 splits; 8 streams: 1022 against 1021 with 5; 4 and 2 streams want their 10 / 16)
  if (ctas * 5 >= sm_count * 4) return 1;
  int n = (2 * sm_count + ctas - 1) / ctas;
  return n < 1 ? 1 : (n > 16 ? 16 : n);
}

// Vocab splits of the posterior sampler, 1..32.
inline int default_post_splits(const dflash_config_t& c, int sm_count) {
  if (c.post_splits > 0) return c.post_splits;
  const int rows = c.max_requests * c.block_size;
  // about TWO CTAs per SM in all (the verify kernel walks the rows when there are more of them): the context-injection
  // kernel that overlaps it needs its own CTA on every SM, and the register file of an SM sub-partition (16 K
  // registers) holds four warps of this kernel (4 x 32 x 64) next to two of that one (2 x 32 x 112), not more
  // (measured: -10..13 us per step at one stream). Wide batches (more block rows than SMs) gain nothing from the overlap
  // -- the injection GEMM is tensor-bound there and from 32 streams up the board is power-capped -- and want the
  // memory parallelism of about four waves of CTAs instead (R = 16 / 64: 1459 / 4242 us per step against 1484 / 4312).
  if (rows > sm_count) {
    const int n = (4 * sm_count + rows - 1) / rows;
    return n < 2 ? 2 : n;
  }
  int n = (2 * sm_count) / rows;
  return n < 1 ? 1 : (n > 32 ? 32 : n);
}

// Fills reg[] (offsets/sizes) for cfg; returns total bytes or 0 on a bad config.
inline size_t layout_workspace(const dflash_config_t& c, Region* reg, int sm_count) {
  const int SL = c.block_size <= 16 ? 16 : 32;
  const int R = c.max_requests;
  const int RS = R * SL;
  const int RSp = rows_padded(RS), RS2p = rows_padded(2 * RS);
  const int H = c.hidden, I = c.intermediate, Hq = c.n_q_heads, Hkv = c.n_kv_heads;
  const int grid = c.gemm_grid > 0 ? c.gemm_grid : sm_count;
  const int nsa = default_attn_splits(c, sm_count);
  const int nsp = default_post_splits(c, sm_count);
  const int qkv_cols = (Hq + 2 * Hkv) * 128;
  // widest fp32 partial plane over the GEMMs whose consumers are separate kernels (fc, qkv, o, down + prompt pass)
  long long ws_elems = 0;
  struct G { int rows, N, K; } gs[] = {{RS, H, c.n_sel * H}, {2 * RS, qkv_cols, H}, {RS, H, Hq * 128}, {RS, H, I},
                                       {kPrefillRows, H, c.n_sel * H}, {kPrefillRows, 2 * Hkv * 128, H}};
  for (auto& g : gs) {
    const int nt = (g.N + kTileN - 1) / kTileN, kb = g.K / kTileK;
    const long long T = static_cast<long long>(nt) * kb;
    const int ranges = ranges_for(grid, groups_of(g.rows));
    const int gg = T < ranges ? static_cast<int>(T) : ranges;
    const long long e = static_cast<long long>(max_slots_for(nt, kb, gg)) * rows_padded(g.rows) * g.N;
    if (e > ws_elems) ws_elems = e;
  }
  // gate/up GEMM (SwiGLU epilogue): one [128 x mb] fp32 partial tile per CTA (groups * ranges <= grid CTAs) and one
  // arrival counter per (column group, tile) per layer
  const long long part_elems = static_cast<long long>(grid) * kTileN * round_mb(RS);
  // (+ the context-injection kernel's two CTA counters)
  const long long n_flags = static_cast<long long>(c.n_layers) * (I / (kTileN / 2)) * groups_of(RS) + 2;
  size_t sz[DFLASH_BUF_COUNT] = {0};
  sz[DFLASH_BUF_X] = static_cast<size_t>(RSp) * H * 2;
  sz[DFLASH_BUF_A_IN] = static_cast<size_t>(RS2p) * H * 2;
  sz[DFLASH_BUF_CTX_FEAT] = static_cast<size_t>(RSp) * c.n_sel * H * 2;
  sz[DFLASH_BUF_Q] = static_cast<size_t>(RSp) * Hq * 128 * 2;
  sz[DFLASH_BUF_ATTN_OUT] = static_cast<size_t>(RSp) * Hq * 128 * 2;
  sz[DFLASH_BUF_A2] = static_cast<size_t>(RSp) * H * 2;
  sz[DFLASH_BUF_HMID] = static_cast<size_t>(RSp) * I * 2;
  sz[DFLASH_BUF_HN] = static_cast<size_t>(RSp) * H * 2;
  sz[DFLASH_BUF_KV] = static_cast<size_t>(c.n_layers) * 2 * R * Hkv * c.max_seq * 128 * 2;
  sz[DFLASH_BUF_WS] = static_cast<size_t>(ws_elems) * 4;
  sz[DFLASH_BUF_PART] = static_cast<size_t>(part_elems) * 4;
  sz[DFLASH_BUF_FLAGS] = static_cast<size_t>(n_flags) * 4;
  sz[DFLASH_BUF_COUNTERS] = static_cast<size_t>(R + 2) * 4;
  sz[DFLASH_BUF_ATTN_PO] = static_cast<size_t>(nsa) * RS * Hq * 128 * 4;
  sz[DFLASH_BUF_ATTN_ML] = static_cast<size_t>(nsa) * RS * Hq * 2 * 4;
  const int ncand = c.max_candidates > 1 ? 4 : 1;
  sz[DFLASH_BUF_CAND_VAL] = static_cast<size_t>(grid) * RSp * 4 * ncand;
  sz[DFLASH_BUF_CAND_IDX] = static_cast<size_t>(grid) * RSp * 4 * ncand;
  sz[DFLASH_BUF_POST_VAL] = static_cast<size_t>(R) * c.block_size * nsp * 4 * ncand;
  sz[DFLASH_BUF_POST_IDX] = static_cast<size_t>(R) * c.block_size * nsp * 4 * ncand;
  sz[DFLASH_BUF_TOPK_IDX] = static_cast<size_t>(RS) * 4 * 4;
  sz[DFLASH_BUF_TOPK_VAL] = static_cast<size_t>(RS) * 4 * 4;
  sz[DFLASH_BUF_CAND_IDS] = static_cast<size_t>(R) * 4 * c.block_size * 8;
  sz[DFLASH_BUF_CAND_SCORES] = static_cast<size_t>(R) * 4 * 4;
  sz[DFLASH_BUF_CHOSEN] = static_cast<size_t>(R) * 4;
  sz[DFLASH_BUF_DRAFT_TOKENS] = static_cast<size_t>(RS) * 8;
  sz[DFLASH_BUF_BLOCK_IDS] = static_cast<size_t>(R) * c.block_size * 8;
  sz[DFLASH_BUF_POSTERIOR] = static_cast<size_t>(R) * c.block_size * 8;
  sz[DFLASH_BUF_OUTPUT_IDS] = static_cast<size_t>(R) * c.out_len * 8;
  sz[DFLASH_BUF_START] = sz[DFLASH_BUF_CTX_LEN] = sz[DFLASH_BUF_DONE] = sz[DFLASH_BUF_N_CYCLES] =
      sz[DFLASH_BUF_BLK_LEN] = sz[DFLASH_BUF_MAX_LEN] = static_cast<size_t>(R) * 4;
  sz[DFLASH_BUF_ACC_HIST] = static_cast<size_t>(R) * c.hist_len * 4;
  sz[DFLASH_BUF_RNG_STEP] = 8;
  sz[DFLASH_BUF_DRAFT_LOGITS] = c.keep_draft_logits ? static_cast<size_t>(RSp) * c.vocab * 2 : 0;
  sz[DFLASH_BUF_PF_FEAT] = static_cast<size_t>(kPrefillRows) * c.n_sel * H * 2;
  sz[DFLASH_BUF_PF_A] = static_cast<size_t>(kPrefillRows) * H * 2;
  size_t off = 0;
  for (int i = 0; i < DFLASH_BUF_COUNT; ++i) {
    reg[i].off = off;
    reg[i].bytes = sz[i];
    off += align_up(sz[i], 1024);
  }
  return off;
}

inline int check_config(const dflash_config_t& c) {
  if (c.head_dim != 128) { set_error("head_dim %d unsupported (128 only)", c.head_dim); return DFLASH_ERR_ARG; }
  if (c.hidden % 128 || c.intermediate % 64 || c.hidden > 8192) {
    set_error("hidden must be a multiple of 128 (<= 8192) and intermediate a multiple of 64");
    return DFLASH_ERR_ARG;
  }
  if (c.block_size < 2 || c.block_size > 32) { set_error("block_size must be in [2,32]"); return DFLASH_ERR_ARG; }
  if (c.n_q_heads % c.n_kv_heads || c.n_q_heads / c.n_kv_heads > 8) {
    set_error("GQA group must divide and be <= 8");
    return DFLASH_ERR_ARG;
  }
  const int SL = c.block_size <= 16 ? 16 : 32;
  if (c.max_requests < 1 || c.max_requests > 64) {
    set_error("max_requests %d unsupported: 1..64 request streams per engine", c.max_requests);
    return DFLASH_ERR_ARG;
  }
  if (c.max_candidates < 0 || c.max_candidates > 4 || (c.max_candidates > 1 && c.max_requests * SL > 32)) {
    set_error("max_candidates must be 0..4 and needs max_requests * row slots <= 32");
    return DFLASH_ERR_ARG;
  }
  if (c.n_sel < 1 || c.n_sel > 8) { set_error("n_sel must be in [1,8]"); return DFLASH_ERR_ARG; }
  if (c.attn_splits > 16) { set_error("attn_splits must be <= 16"); return DFLASH_ERR_ARG; }
  if (c.max_seq < 2 * c.block_size || c.out_len < 1 || c.hist_len < 1) {
    set_error("max_seq/out_len/hist_len too small");
    return DFLASH_ERR_ARG;
  }
  return DFLASH_OK;
}

#define DFL_CUDA(expr, what)                              \
  do {                                                    \
    cudaError_t _e = (expr);                              \
    if (_e != cudaSuccess) return cuda_fail(_e, what);    \
  } while (0)

template <class Kern, class... Args>
inline cudaError_t launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                              const Args&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

inline int engine_create(const dflash_config_t& c, const dflash_weights_t& w, void* workspace,
                         size_t workspace_bytes, int sm_count, Engine** out) {
  int rc = check_config(c);
  if (rc) return rc;
  Engine* e = new Engine();
  e->cfg = c;
  e->w = w;
  e->layers.assign(w.layers_host, w.layers_host + c.n_layers);
  e->w.layers_host = nullptr;
  e->R = c.max_requests;
  e->SL = c.block_size <= 16 ? 16 : 32;
  e->RS = e->R * e->SL;
  e->H = c.hidden; e->I = c.intermediate; e->L = c.n_layers; e->Hq = c.n_q_heads; e->Hkv = c.n_kv_heads;
  e->V = c.vocab; e->nsel = c.n_sel; e->bs = c.block_size;
  e->grid = c.gemm_grid > 0 ? c.gemm_grid : sm_count;
  e->sm_count = sm_count;
  e->nsplit_attn = default_attn_splits(c, sm_count);
  e->nsplit_post = default_post_splits(c, sm_count);
  e->pdl = c.use_pdl != 0;
  e->total = layout_workspace(c, e->reg, sm_count);
  if (workspace == nullptr || workspace_bytes < e->total) {
    set_error("workspace too small: need %zu bytes, got %zu", e->total, workspace_bytes);
    delete e;
    return DFLASH_ERR_ARG;
  }
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) {
    set_error("workspace must be 1024-byte aligned");
    delete e;
    return DFLASH_ERR_ARG;
  }
  e->base = static_cast<uint8_t*>(workspace);

  const int RS = e->RS, H = e->H, I = e->I;
  const int RSp = rows_padded(RS), RS2p = rows_padded(2 * RS);
  const int qkv_cols = (e->Hq + 2 * e->Hkv) * 128;
  const int mb_blk = round_mb(RS), mb_all = round_mb(2 * RS);
  float* ws = e->buf<float>(DFLASH_BUF_WS);
  // GEMMs whose consumer is a separate kernel: fp32 partial planes in `ws`
  auto finish = [&](GemmPlan& p) {
    p.args.ws = ws;
    p.args.ws_rows = p.groups * p.mb;
    p.args.ws_ld = p.args.N;
  };
  // gate/up GEMM with the SwiGLU epilogue: the shared partial-exchange region and its own arrival counters
  auto finish_fused = [&](GemmPlan& p) -> int {
    p.args.part = e->buf<float>(DFLASH_BUF_PART);
    const size_t need = static_cast<size_t>(p.groups) * p.grid * kTileN * p.mb * 4;
    const size_t nflags = static_cast<size_t>(p.groups) * p.args.n_tiles;
    if (need > e->reg[DFLASH_BUF_PART].bytes || (e->flags_used + nflags) * 4 > e->reg[DFLASH_BUF_FLAGS].bytes) {
      set_error("internal: fused GEMM scratch regions too small");
      return -1;
    }
    p.args.flags = e->buf<unsigned int>(DFLASH_BUF_FLAGS) + e->flags_used;
    e->flags_used += nflags;
    return 0;
  };
#define DFL_PLAN(call)        \
  do {                        \
    int _rc = (call);         \
    if (_rc) { delete e; return DFLASH_ERR_ARG; } \
  } while (0)
  // activation TMA tensors are declared with the buffers' padded row counts (a multiple of the UMMA width), so a box
  // never leaves the allocation; rows past the live ones are zero and their outputs are never used.
  // context injection (dflash.py:177) + block embedding + layer 0's input_layernorm: ONE kernel (kModeCtxNorm)
  // (its CTAs wait for each other: never more of them than SMs)
  DFL_PLAN(make_gemm_plan(&e->fc, w.fc, H, 0, H, e->nsel * H, e->buf<void>(DFLASH_BUF_CTX_FEAT), RSp, 0, mb_blk, RS,
                          kModeCtxNorm, e->grid < sm_count ? e->grid : sm_count));
  finish(e->fc);
  {
    CtxNormEpi& cn = e->fc.args.cn;
    cn.out = e->buf<__nv_bfloat16>(DFLASH_BUF_A_IN);
    cn.ld = H;
    cn.norm_w = static_cast<const __nv_bfloat16*>(w.hidden_norm);
    cn.max_slots = e->fc.max_slots;
    if ((e->flags_used + 2) * 4 > e->reg[DFLASH_BUF_FLAGS].bytes) {
      set_error("internal: flag region too small");
      delete e;
      return DFLASH_ERR_ARG;
    }
    cn.sync = e->buf<unsigned int>(DFLASH_BUF_FLAGS) + e->flags_used;
    e->flags_used += 2;
    cn.ctx_len = e->buf<int>(DFLASH_BUF_CTX_LEN);
    cn.SL = e->SL;
    cn.eps = c.rms_eps;
    cn.ids_ld = e->bs;
    cn.bs = e->bs;
    cn.n_blk_rows = RS;
    cn.pad_token = c.mask_token_id;
    cn.resid = e->buf<__nv_bfloat16>(DFLASH_BUF_X);
    cn.ln_w = static_cast<const __nv_bfloat16*>(e->layers[0].ln1);
    cn.blk_out = e->buf<__nv_bfloat16>(DFLASH_BUF_A_IN) + static_cast<size_t>(RS) * H;
    cn.direct = 0;
    cn.kb_per_sel = H / kTileK;
    cn.embed = static_cast<const __nv_bfloat16*>(w.embed);
    cn.ids = e->buf<long long>(DFLASH_BUF_BLOCK_IDS);
    e->fc_direct = e->fc;
    e->fc_direct.args.cn.direct = 1;
  }
  e->qkv.resize(e->L); e->kv_pf.resize(e->L * 5); e->o.resize(e->L); e->gu.resize(e->L); e->d.resize(e->L);
  for (int i = 0; i < 5; ++i) {
    DFL_PLAN(make_gemm_plan(&e->fc_pf[i], w.fc, H, 0, H, e->nsel * H, e->buf<void>(DFLASH_BUF_PF_FEAT), kPrefillRows, 0,
                            16 << i, 16 << i, kModePartials, e->grid));
    finish(e->fc_pf[i]);
  }
  for (int l = 0; l < e->L; ++l) {
    const dflash_layer_weights_t& lw = e->layers[l];
    DFL_PLAN(make_gemm_plan(&e->qkv[l], lw.wqkv, qkv_cols, 0, qkv_cols, H, e->buf<void>(DFLASH_BUF_A_IN), RS2p, 0,
                            mb_all, 2 * RS, kModePartials, e->grid));
    finish(e->qkv[l]);
    for (int i = 0; i < 5; ++i) {
      GemmPlan& kp = e->kv_pf[l * 5 + i];
      DFL_PLAN(make_gemm_plan(&kp, lw.wqkv, qkv_cols, e->Hq * 128, 2 * e->Hkv * 128, H, e->buf<void>(DFLASH_BUF_PF_A),
                              kPrefillRows, 0, 16 << i, 16 << i, kModePartials, e->grid));
      finish(kp);
    }
    DFL_PLAN(make_gemm_plan(&e->o[l], lw.wo, H, 0, H, e->Hq * 128, e->buf<void>(DFLASH_BUF_ATTN_OUT), RSp, 0, mb_blk, RS,
                            kModePartials, e->grid));
    finish(e->o[l]);
    DFL_PLAN(make_gemm_plan(&e->gu[l], lw.wgu, 2 * I, 0, 2 * I, H, e->buf<void>(DFLASH_BUF_A2), RSp, 0, mb_blk, RS,
                            kModeSwiglu, e->grid));
    DFL_PLAN(finish_fused(e->gu[l]));
    e->gu[l].args.sw.out = e->buf<__nv_bfloat16>(DFLASH_BUF_HMID);
    e->gu[l].args.sw.ld = I;
    DFL_PLAN(make_gemm_plan(&e->d[l], lw.wd, H, 0, H, I, e->buf<void>(DFLASH_BUF_HMID), RSp, 0, mb_blk, RS,
                            kModePartials, e->grid));
    finish(e->d[l]);
  }
  unsigned int* counters = e->buf<unsigned int>(DFLASH_BUF_COUNTERS);
  auto tok_fields = [&](GemmPlan& p) {
    p.args.cand_val = e->buf<float>(DFLASH_BUF_CAND_VAL);
    p.args.cand_idx = e->buf<int>(DFLASH_BUF_CAND_IDX);
    p.args.tok_counter = counters + e->R + 1;
    p.args.tok_block_ids = e->buf<long long>(DFLASH_BUF_BLOCK_IDS);
    p.args.tok_draft_tokens = e->buf<long long>(DFLASH_BUF_DRAFT_TOKENS);
    p.args.tok_SL = e->SL;
    p.args.tok_bs = e->bs;
    p.args.tok_rows = RS;
  };
  DFL_PLAN(make_gemm_plan(&e->lm, w.lm_head, e->V, 0, e->V, H, e->buf<void>(DFLASH_BUF_HN), RSp, 0, mb_blk, RS,
                          c.keep_draft_logits ? kModeArgmaxDump : kModeArgmax, e->grid));
  tok_fields(e->lm);
  e->lm.args.logits = c.keep_draft_logits ? e->buf<__nv_bfloat16>(DFLASH_BUF_DRAFT_LOGITS) : nullptr;
  e->lm.args.logits_ld = e->V;
  if (mb_blk <= 32) {
    DFL_PLAN(make_gemm_plan(&e->lm_sample, w.lm_head, e->V, 0, e->V, H, e->buf<void>(DFLASH_BUF_HN), RSp, 0, mb_blk, RS,
                            kModeSample, e->grid));
    tok_fields(e->lm_sample);
    e->lm_sample.args.rng_step = e->buf<unsigned long long>(DFLASH_BUF_RNG_STEP);
    e->has_sample = true;
  }
  e->max_cand = c.max_candidates > 1 ? c.max_candidates : 1;
  if (e->max_cand > 1) {
    DFL_PLAN(make_gemm_plan(&e->lm_topk, w.lm_head, e->V, 0, e->V, H, e->buf<void>(DFLASH_BUF_HN), RSp, 0, mb_blk, RS,
                            kModeTopK, e->grid));
    e->lm_topk.args.cand_val = e->buf<float>(DFLASH_BUF_CAND_VAL);
    e->lm_topk.args.cand_idx = e->buf<int>(DFLASH_BUF_CAND_IDX);
  }
#undef DFL_PLAN
  cudaError_t ce = cudaFuncSetAttribute(attn_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
  if (ce != cudaSuccess) { delete e; return cuda_fail(ce, "attn smem attribute"); }
  ce = cudaFuncSetAttribute(finalize_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 6);
  if (ce != cudaSuccess) { delete e; return cuda_fail(ce, "finalize smem attribute"); }
  // One shared-memory carveout for every kernel of the step: consecutive kernels with different L1/smem splits
  // cannot share an SM, which would serialise exactly the PDL overlaps the schedule relies on.
  {
    const int mx = cudaSharedmemCarveoutMaxShared;
    cudaFuncSetAttribute(finalize_rows_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(finalize_rows_cluster_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(finalize_rows_block_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(qkv_post_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(attn_split_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(attn_combine_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(verify_fused_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(verify_fused_wide_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(posterior_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(accept_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(ctx_gather_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaGetLastError();
  }
  *out = e;
  return DFLASH_OK;
}

template <class Kern, class Args>
inline cudaError_t launch_cluster_pdl(Kern kern, dim3 grid, dim3 block, dim3 cluster, size_t smem, cudaStream_t st,
                                      bool pdl, const Args& args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster.x;
  at[0].val.clusterDim.y = cluster.y;
  at[0].val.clusterDim.z = cluster.z;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, args);
}

// ------------------------------------------------------------------------------------------------
inline RowsArgs rows_args_base(const Engine* e) {
  RowsArgs a;
  memset(&a, 0, sizeof(a));
  a.H = e->H;
  a.SL = e->SL;
  a.bs = e->bs;
  a.eps = e->cfg.rms_eps;
  a.ctx_len = e->buf<int>(DFLASH_BUF_CTX_LEN);
  a.valid_mode = kRowsAll;
  return a;
}

// Per-layer row pass (partials + residual + RMSNorm): four CTAs per row when the width allows it
inline cudaError_t launch_finalize(Engine* e, const RowsArgs& a, int rows, cudaStream_t st) {
  const bool ok = a.ws != nullptr && a.embed == nullptr && a.norm_w != nullptr &&
                  e->H % (4 * kRowCtas) == 0 && e->H / kRowCtas <= kRowClThreads * 4 * kRowClGroups;
  // one 1024-thread CTA per row -- no cluster exchange -- whenever the row fits (hidden <= 4096)
  if (DFLASH_ROW_BLOCK && ok && e->H <= 4 * kRowBlkThreads && rows <= e->row_block_max_rows)
    return launch_pdl(finalize_rows_block_kernel, dim3(rows), dim3(kRowBlkThreads), 0, st, e->pdl, a);
  if (ok)
    return launch_cluster_pdl(finalize_rows_cluster_kernel, dim3(kRowCtas, rows), dim3(kRowClThreads),
                              dim3(kRowCtas, 1, 1), 0, st, e->pdl, a);
  return launch_pdl(finalize_rows_kernel, dim3(rows), dim3(kRowsThreads), static_cast<size_t>(e->H) * 6, st, e->pdl, a);
}

inline QkvPostArgs qkv_post_args(Engine* e, int l, const GemmPlan& p, bool kv_only) {
  QkvPostArgs a;
  memset(&a, 0, sizeof(a));
  a.ws = p.args.ws;
  a.sm = slot_map_of(p);
  a.R = e->R; a.SL = e->SL; a.bs = e->bs; a.Hq = e->Hq; a.Hkv = e->Hkv;
  a.q_cols = kv_only ? 0 : e->Hq * 128;
  a.row0 = 0;
  a.rows = kv_only ? e->RS : 2 * e->RS;
  a.start = e->buf<int>(DFLASH_BUF_START);
  a.ctx_len = e->buf<int>(DFLASH_BUF_CTX_LEN);
  a.blk_len = e->buf<int>(DFLASH_BUF_BLK_LEN);
  a.q_norm_w = static_cast<const __nv_bfloat16*>(e->layers[l].q_norm);
  a.k_norm_w = static_cast<const __nv_bfloat16*>(e->layers[l].k_norm);
  {
    const __nv_bfloat16* b = static_cast<const __nv_bfloat16*>(e->layers[l].bqkv);
    a.bias = b == nullptr ? nullptr : (kv_only ? b + e->Hq * 128 : b);
  }
  a.inv_freq = e->w.inv_freq;
  a.rope_scale = e->cfg.rope_scale;
  a.eps = e->cfg.rms_eps;
  a.q_out = e->buf<__nv_bfloat16>(DFLASH_BUF_Q);
  const size_t per = static_cast<size_t>(e->R) * e->Hkv * e->cfg.max_seq * 128;
  a.k_cache = e->buf<__nv_bfloat16>(DFLASH_BUF_KV) + (static_cast<size_t>(l) * 2 + 0) * per;
  a.v_cache = e->buf<__nv_bfloat16>(DFLASH_BUF_KV) + (static_cast<size_t>(l) * 2 + 1) * per;
  a.S_max = e->cfg.max_seq;
  return a;
}

// One draft step: ctx injection -> block embedding -> L layers -> final norm -> lm_head + argmax.
// Writes the drafted tokens into block_ids[:, 1:bs]  (dflash.py:235-247). 1 + 9 L + 1 launches:
//   fc GEMM [epilogue: hidden_norm of the ctx rows | embed + ln1 of the block]
//   per layer: qkv GEMM - qkv_post [q/k norm, RoPE, cache write] - attention split - merge - o GEMM -
//              row kernel [+ residual, ln2] - gate/up GEMM [SwiGLU epilogue] - down GEMM - row kernel [+ residual, norm]
//   lm_head GEMM [argmax; its last CTA reduces the candidates to the drafted tokens]
// n_candidates > 1: top-4 lm_head epilogue + candidate blocks (fixed_prefix_rank) instead of the plain argmax tail
inline int enqueue_draft_step(Engine* e, const void* noise_embedding, bool run_lm_head, cudaStream_t st,
                              int n_candidates = 1, int fixed_prefix_len = 0, float draft_temperature = 0.f,
                              unsigned long long draft_seed = 0, bool inject = true) {
  const int RS = e->RS;
  __nv_bfloat16* x = e->buf<__nv_bfloat16>(DFLASH_BUF_X);
  __nv_bfloat16* a_in = e->buf<__nv_bfloat16>(DFLASH_BUF_A_IN);
  // Context injection as ONE kernel: the fc GEMM over the features gathered by the previous verify step finishes its
  // tiles in its own epilogue (bf16 + hidden_norm -> a_in ctx rows), and its epilogue warps also do the block rows
  // (embed_tokens(block_ids) -> residual stream, input_layernorm of layer 0 -> a_in block rows) before their first tile.
  // inject == false: the previous verify step (or the prompt pass) already did it (enqueue_verify_step with inject).
  if (inject) {
    GemmPlan fc = e->fc;
    if (noise_embedding != nullptr) {
      fc.args.cn.embed = static_cast<const __nv_bfloat16*>(noise_embedding);  // [R*SL, H] rows, already embedded
      fc.args.cn.ids = nullptr;
    } else {
      fc.args.cn.embed = static_cast<const __nv_bfloat16*>(e->w.embed);
      fc.args.cn.ids = e->buf<long long>(DFLASH_BUF_BLOCK_IDS);
    }
    DFL_CUDA(launch_gemm(fc, st, e->pdl), "context injection (fc gemm + hidden_norm + embed + ln1)");
  }
  AttnArgs aa;
  memset(&aa, 0, sizeof(aa));
  aa.R = e->R; aa.SL = e->SL; aa.bs = e->bs; aa.Hq = e->Hq; aa.Hkv = e->Hkv; aa.S_max = e->cfg.max_seq;
  aa.nsplit = e->nsplit_attn;
  aa.start = e->buf<int>(DFLASH_BUF_START);
  aa.blk_len = e->buf<int>(DFLASH_BUF_BLK_LEN);
  aa.ctx_len = e->buf<int>(DFLASH_BUF_CTX_LEN);
  aa.q = e->buf<__nv_bfloat16>(DFLASH_BUF_Q);
  aa.part_o = e->buf<float>(DFLASH_BUF_ATTN_PO);
  aa.part_ml = e->buf<float>(DFLASH_BUF_ATTN_ML);
  aa.scale_log2 = 1.4426950408889634f / sqrtf(128.0f);
  aa.out = e->buf<__nv_bfloat16>(DFLASH_BUF_ATTN_OUT);
  const int group = e->Hq / e->Hkv;
  for (int l = 0; l < e->L; ++l) {
    DFL_CUDA(launch_gemm(e->qkv[l], st, e->pdl), "qkv gemm");
    QkvPostArgs qa = qkv_post_args(e, l, e->qkv[l], false);
    const int items = qa.rows * (e->Hq + 2 * e->Hkv);
    aa.k_cache = qa.k_cache;
    aa.v_cache = qa.v_cache;
    // (four items per warp with their loads in flight together, for the ten waves of one-item CTAs at 64 streams: 170
    // registers, 1563 / 4384 us per step at 16 / 64 streams against 1440 / 4112 -- occupancy is what hides the latency)
    DFL_CUDA(launch_pdl(qkv_post_kernel, dim3((items + kItemWarps - 1) / kItemWarps), dim3(32 * kItemWarps), 0, st,
                        e->pdl, qa), "qkv post");
    DFL_CUDA(launch_pdl(attn_split_kernel, dim3(e->nsplit_attn, e->Hkv, e->R * (e->SL / 16)), dim3(32 * group),
                        kAttnSmem, st, e->pdl, aa), "attention");
    if (e->nsplit_attn > 1)  // (a single split writes the normalised output itself)
      DFL_CUDA(launch_pdl(attn_combine_kernel, dim3((RS * e->Hq + kItemWarps - 1) / kItemWarps), dim3(32 * kItemWarps), 0,
                          st, e->pdl, aa), "attention combine");
    DFL_CUDA(launch_gemm(e->o[l], st, e->pdl), "o gemm");
    {
      RowsArgs a = rows_args_base(e);
      a.ws = e->o[l].args.ws;
      a.sm = slot_map_of(e->o[l]);
      a.bias = static_cast<const __nv_bfloat16*>(e->layers[l].bo);
      a.resid = x;
      a.norm_w = static_cast<const __nv_bfloat16*>(e->layers[l].ln2);
      a.out = e->buf<__nv_bfloat16>(DFLASH_BUF_A2);
      DFL_CUDA(launch_finalize(e, a, RS, st), "o finalize");
    }
    DFL_CUDA(launch_gemm(e->gu[l], st, e->pdl), "gate/up gemm + swiglu");
    DFL_CUDA(launch_gemm(e->d[l], st, e->pdl), "down gemm");
    {
      RowsArgs a = rows_args_base(e);
      a.ws = e->d[l].args.ws;
      a.sm = slot_map_of(e->d[l]);
      a.resid = x;
      if (l + 1 < e->L) {
        a.norm_w = static_cast<const __nv_bfloat16*>(e->layers[l + 1].ln1);
        a.out = a_in + static_cast<size_t>(RS) * e->H;
      } else {
        a.norm_w = static_cast<const __nv_bfloat16*>(e->w.final_norm);
        a.out = e->buf<__nv_bfloat16>(DFLASH_BUF_HN);
      }
      DFL_CUDA(launch_finalize(e, a, RS, st), "down finalize");
    }
  }
  if (!run_lm_head) return DFLASH_OK;
  if (n_candidates > 1) {
    DFL_CUDA(launch_gemm(e->lm_topk, st, e->pdl), "lm_head top-k gemm");
    CandArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.cand_val = e->lm_topk.args.cand_val;
    ca.cand_idx = e->lm_topk.args.cand_idx;
    ca.n_cta = e->lm_topk.grid;
    ca.cand_ld = e->lm_topk.args.cand_ld;
    ca.R = e->R; ca.SL = e->SL; ca.bs = e->bs; ca.prefix_len = fixed_prefix_len;
    ca.blk_len = e->buf<int>(DFLASH_BUF_BLK_LEN);
    ca.block_ids = e->buf<long long>(DFLASH_BUF_BLOCK_IDS);
    ca.draft_tokens = e->buf<long long>(DFLASH_BUF_DRAFT_TOKENS);
    ca.topk_idx = e->buf<int>(DFLASH_BUF_TOPK_IDX);
    ca.topk_val = e->buf<float>(DFLASH_BUF_TOPK_VAL);
    ca.cand_ids = e->buf<long long>(DFLASH_BUF_CAND_IDS);
    ca.cand_scores = e->buf<float>(DFLASH_BUF_CAND_SCORES);
    DFL_CUDA(launch_pdl(candidates_kernel, dim3(e->R), dim3(256), 0, st, e->pdl, ca), "candidate blocks");
    return DFLASH_OK;
  }
  if (draft_temperature >= 1e-5f) {
    GemmPlan p = e->lm_sample;
    p.args.inv_temp = 1.0f / draft_temperature;
    p.args.seed = draft_seed;
    p.args.step_base = 0;
    DFL_CUDA(launch_gemm(p, st, e->pdl), "lm_head sampling gemm");
  } else {
    DFL_CUDA(launch_gemm(e->lm, st, e->pdl), "lm_head gemm");
  }
  return DFLASH_OK;
}

struct VerifyInputs {
  const void* target_logits;  // [R*bs][V] bf16 (row pitch logits_ld), or null with posterior_in
  long long logits_ld;
  const long long* posterior_in;
  const void* hidden[8];      // n_sel x [R*bs][H] bf16
  float temperature;
  const float* noise;
  unsigned long long seed;
  const long long* stop_ids;
  int n_stop;
  const int* forced_k;
  int forced_ld;
  int clamp_tail;
  int n_candidates;  // > 1: rows are [R][n_candidates][bs]
};

// Posterior sampling -> acceptance/commit/state -> next-cycle context gather  (dflash.py:257-268).
// The plain path is ONE kernel (verify_fused_kernel); multi-candidate verify and the given-posterior harness hook
// keep the three-kernel form (the winner's rows are only known after the acceptance).
// Block rows only: embed_tokens(block_ids) -> residual stream, input_layernorm of layer 0 -> a_in block rows
// (model/dflash.py:237 + the first norm). What the context-injection kernel does for the block rows, as its own launch:
// after a request reset (or any other rewrite of block_ids), when the next draft step runs without injection.
inline int enqueue_embed_block(Engine* e, cudaStream_t st) {
  RowsArgs a = rows_args_base(e);
  a.embed = static_cast<const __nv_bfloat16*>(e->w.embed);
  a.ids = e->buf<long long>(DFLASH_BUF_BLOCK_IDS);
  a.ids_ld = e->bs;
  a.pad_token = e->cfg.mask_token_id;
  a.resid = e->buf<__nv_bfloat16>(DFLASH_BUF_X);
  a.norm_w = static_cast<const __nv_bfloat16*>(e->layers[0].ln1);
  a.out = e->buf<__nv_bfloat16>(DFLASH_BUF_A_IN) + static_cast<size_t>(e->RS) * e->H;
  DFL_CUDA(launch_pdl(finalize_rows_kernel, dim3(e->RS), dim3(kRowsThreads), static_cast<size_t>(e->H) * 6, st, e->pdl, a),
           "block embedding");
  return DFLASH_OK;
}

// The next cycle's context injection reading the selected hidden states in place (dflash.py:263 + :177 + :237 as one
// kernel), launched behind the verify kernel: its fc main loop overlaps that kernel, its row pass waits for it.
inline int enqueue_inject_direct(Engine* e, const void* const* hidden, cudaStream_t st) {
  bool same = true;
  for (int s = 0; s < e->nsel; ++s) same = same && hidden[s] == e->direct_hidden[s];
  if (!same) {
    GemmPlan& p = e->fc_direct;
    for (int s = 0; s < e->nsel; ++s) {
      if (make_tmap_hidden3d(&p.xm.m[s], hidden[s], e->R, e->bs, e->H, e->SL, p.mb / e->SL)) return DFLASH_ERR_ARG;
      e->direct_hidden[s] = hidden[s];
    }
  }
  DFL_CUDA(launch_gemm(e->fc_direct, st, e->pdl), "context injection (direct: concat + fc + hidden_norm + embed + ln1)");
  return DFLASH_OK;
}

inline int enqueue_verify_step(Engine* e, const VerifyInputs& v, cudaStream_t st, bool inject = false) {
  const int K = v.n_candidates > 1 ? v.n_candidates : 1;
  const int rows = e->R * K * e->bs;
  PosteriorArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.logits = static_cast<const __nv_bfloat16*>(v.target_logits);
  pa.ld = v.logits_ld;
  pa.rows = rows;
  pa.V = e->V;
  pa.nsplit = e->nsplit_post;
  pa.inv_temp = v.temperature < 1e-5f ? 0.f : 1.0f / v.temperature;
  pa.noise = v.noise;
  pa.seed = v.seed;
  pa.rng_step = e->buf<unsigned long long>(DFLASH_BUF_RNG_STEP);
  pa.cand_val = e->buf<float>(DFLASH_BUF_POST_VAL);
  pa.cand_idx = e->buf<int>(DFLASH_BUF_POST_IDX);
  AcceptArgs aa;
  memset(&aa, 0, sizeof(aa));
  aa.R = e->R; aa.bs = e->bs; aa.nsplit = e->nsplit_post;
  aa.cand_val = e->buf<float>(DFLASH_BUF_POST_VAL);
  aa.cand_idx = e->buf<int>(DFLASH_BUF_POST_IDX);
  aa.posterior_in = v.posterior_in;
  aa.posterior = e->buf<long long>(DFLASH_BUF_POSTERIOR);
  aa.block_ids = e->buf<long long>(DFLASH_BUF_BLOCK_IDS);
  aa.ids_ld = e->bs;
  aa.output_ids = e->buf<long long>(DFLASH_BUF_OUTPUT_IDS);
  aa.out_ld = e->cfg.out_len;
  aa.start = e->buf<int>(DFLASH_BUF_START);
  aa.ctx_len = e->buf<int>(DFLASH_BUF_CTX_LEN);
  aa.done = e->buf<int>(DFLASH_BUF_DONE);
  aa.n_cycles = e->buf<int>(DFLASH_BUF_N_CYCLES);
  aa.blk_len = e->buf<int>(DFLASH_BUF_BLK_LEN);
  aa.acc_hist = e->buf<int>(DFLASH_BUF_ACC_HIST);
  aa.hist_ld = e->cfg.hist_len;
  aa.max_len = e->buf<int>(DFLASH_BUF_MAX_LEN);
  aa.stop_ids = v.stop_ids;
  aa.n_stop = v.n_stop;
  aa.mask_token = e->cfg.mask_token_id;
  aa.forced_k = v.forced_k;
  aa.forced_ld = v.forced_ld > 0 ? v.forced_ld : 1;
  aa.clamp_tail = v.clamp_tail;
  aa.rng_step = e->buf<unsigned long long>(DFLASH_BUF_RNG_STEP);
  aa.K = K;
  aa.cand_ids = e->buf<long long>(DFLASH_BUF_CAND_IDS);
  aa.cand_scores = e->buf<float>(DFLASH_BUF_CAND_SCORES);
  aa.chosen = e->buf<int>(DFLASH_BUF_CHOSEN);
  GatherArgs ga;
  memset(&ga, 0, sizeof(ga));
  for (int s = 0; s < e->nsel; ++s) ga.src[s] = static_cast<const __nv_bfloat16*>(v.hidden[s]);
  ga.n_sel = e->nsel; ga.H = e->H; ga.SL = e->SL;
  ga.r0 = 0; ga.nreq = e->R;
  ga.src_rows = e->bs;
  ga.src_row0 = 0;
  ga.ctx_len = e->buf<int>(DFLASH_BUF_CTX_LEN);
  ga.ctx_feat = e->buf<__nv_bfloat16>(DFLASH_BUF_CTX_FEAT);
  ga.K = K;
  ga.chosen = K > 1 ? e->buf<int>(DFLASH_BUF_CHOSEN) : nullptr;
  if (K == 1 && v.posterior_in == nullptr) {
    VerifyFusedArgs fa;
    fa.post = pa;
    fa.acc = aa;
    fa.gather = ga;
    fa.counters = e->buf<unsigned int>(DFLASH_BUF_COUNTERS);
    if (inject) fa.gather.n_sel = 0;  // the injection kernel reads the hidden states in place: nothing to gather
    int gy = rows;  // one CTA per (split, row); a narrower grid walks the rows
    if (rows <= e->sm_count && e->nsplit_post * rows > 2 * e->sm_count) gy = (2 * e->sm_count) / e->nsplit_post;
    gy = gy < 1 ? 1 : (gy > rows ? rows : gy);
    if (rows > e->sm_count)
      DFL_CUDA(launch_pdl(verify_fused_wide_kernel, dim3(e->nsplit_post, gy), dim3(256), 0, st, e->pdl, fa), "verify");
    else
      DFL_CUDA(launch_pdl(verify_fused_kernel, dim3(e->nsplit_post, gy), dim3(256), 0, st, e->pdl, fa), "verify");
    if (inject) return enqueue_inject_direct(e, v.hidden, st);
    return DFLASH_OK;
  }
  if (v.posterior_in == nullptr)
    DFL_CUDA(launch_pdl(posterior_kernel, dim3(e->nsplit_post, rows), dim3(256), 0, st, e->pdl, pa), "posterior");
  DFL_CUDA(launch_pdl(accept_kernel, dim3(e->R), dim3(32), 0, st, e->pdl, aa), "accept");
  DFL_CUDA(launch_pdl(ctx_gather_kernel, dim3(e->RS, e->nsel), dim3(256), 0, st, e->pdl, ga), "ctx gather");
  if (inject) {  // (the gathered form of the injection kernel: these paths only know the rows after the acceptance)
    GemmPlan fc = e->fc;
    DFL_CUDA(launch_gemm(fc, st, e->pdl), "context injection (fc gemm + hidden_norm + embed + ln1)");
  }
  return DFLASH_OK;
}

// Prompt context of request r (cycle 0 of dflash.py:229,238-246 with c = P): `n_rows` rows of the selected target
// hidden states are projected in passes of up to 256 rows -- fc + hidden_norm, then per layer the K/V rows of wqkv,
// k_norm, RoPE -- and land at cache positions [pos0, pos0 + n_rows). Real M = P GEMMs: every weight byte is read
// once per 256 prompt rows. Uses its own feature / activation buffers, so requests that are mid-generation in the
// same engine keep their pending context rows. Afterwards start[r] = pos0 + n_rows and ctx_len[r] = 0.
inline int enqueue_prefill(Engine* e, int r, const void* const* hidden, int n_rows, int pos0, cudaStream_t st) {
  if (r < 0 || r >= e->R || n_rows < 1 || pos0 < 0 || pos0 + n_rows + 2 * e->bs > e->cfg.max_seq) {
    set_error("prefill: bad request %d, rows %d or position %d (max_seq %d)", r, n_rows, pos0, e->cfg.max_seq);
    return DFLASH_ERR_ARG;
  }
  for (int c0 = 0; c0 < n_rows; c0 += kPrefillRows) {
    const int n = n_rows - c0 < kPrefillRows ? n_rows - c0 : kPrefillRows;
    int pi = 0;
    while ((16 << pi) < n) ++pi;
    GatherArgs ga;
    memset(&ga, 0, sizeof(ga));
    for (int s = 0; s < e->nsel; ++s) ga.src[s] = static_cast<const __nv_bfloat16*>(hidden[s]);
    ga.n_sel = e->nsel; ga.H = e->H; ga.SL = e->SL;
    ga.src_rows = n_rows;
    ga.src_row0 = c0;
    ga.pf_rows = n;
    ga.ctx_feat = e->buf<__nv_bfloat16>(DFLASH_BUF_PF_FEAT);
    DFL_CUDA(launch_pdl(ctx_gather_kernel, dim3(n, e->nsel), dim3(256), 0, st, false, ga), "prefill gather");
    GemmPlan fc = e->fc_pf[pi];
    fc.args.m_valid = n;
    DFL_CUDA(launch_gemm(fc, st, e->pdl), "prefill fc gemm");
    RowsArgs a = rows_args_base(e);
    a.ws = fc.args.ws;
    a.sm = slot_map_of(fc);
    a.norm_w = static_cast<const __nv_bfloat16*>(e->w.hidden_norm);
    a.out = e->buf<__nv_bfloat16>(DFLASH_BUF_PF_A);
    DFL_CUDA(launch_pdl(finalize_rows_kernel, dim3(n), dim3(kRowsThreads), static_cast<size_t>(e->H) * 6, st, e->pdl, a),
             "prefill fc finalize");
    for (int l = 0; l < e->L; ++l) {
      GemmPlan kp = e->kv_pf[l * 5 + pi];
      kp.args.m_valid = n;
      DFL_CUDA(launch_gemm(kp, st, e->pdl), "prefill kv gemm");
      QkvPostArgs qa = qkv_post_args(e, l, kp, true);
      qa.rows = n;
      qa.pf_rows = n;
      qa.pf_req = r;
      qa.pf_pos0 = pos0 + c0;
      const int items = n * (2 * e->Hkv);
      DFL_CUDA(launch_pdl(qkv_post_kernel, dim3((items + kItemWarps - 1) / kItemWarps), dim3(32 * kItemWarps), 0, st,
                          e->pdl, qa), "prefill kv post");
    }
  }
  SetStateArgs sa;
  sa.r = r; sa.start = pos0 + n_rows; sa.ctx_len = 0;
  sa.start_p = e->buf<int>(DFLASH_BUF_START);
  sa.ctx_len_p = e->buf<int>(DFLASH_BUF_CTX_LEN);
  DFL_CUDA(launch_pdl(set_state_kernel, dim3(1), dim3(1), 0, st, false, sa), "set state");
  return DFLASH_OK;
}

}  // namespace dfl
