/* dflash_b200 — C ABI of the B200-native DFlash draft-and-verify hot path.
 *
 * Plain pointers and sizes only (no torch types). Every pointer is a DEVICE pointer unless the
 * name ends in _host. `stream` is a cudaStream_t passed as void*. No entry point allocates device
 * memory or synchronises; all return DFLASH_OK (0) or a negative error (dflash_last_error() holds
 * the message for the calling thread). The library contains sm_100a code only; there is no CPU
 * path behind it.
 *
 * The reference (AtharvRN/dflash) is pure Python and has no FFI of its own. Each entry point names
 * the reference call site(s) (file:line under the reference tree) it replaces; INTEGRATION.md shows
 * the ctypes binding a reference maintainer would add to model/dflash.py.
 */
#ifndef DFLASH_B200_H_
#define DFLASH_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFLASH_ABI_VERSION 3

#define DFLASH_OK 0
#define DFLASH_ERR_ARG (-1)   /* bad argument / unsupported shape */
#define DFLASH_ERR_CUDA (-2)  /* CUDA runtime / driver call failed */
#define DFLASH_ERR_ARCH (-3)  /* not an sm_100 device */

int dflash_abi_version(void);
const char* dflash_last_error(void);

/* Returns the SM count of the current device (>0) or DFLASH_ERR_ARCH if it is not sm_100. */
int dflash_device_check(void);

/* =============================================================================================
 * Engine: the draft+verify step for `max_requests` independent request streams on one GPU.
 * ===========================================================================================*/

/* Mirrors the Qwen3Config keys DFlashDraftModel consumes (model/dflash.py:33-56,157-163). */
typedef struct dflash_config {
  int hidden;           /* hidden_size */
  int intermediate;     /* intermediate_size */
  int n_layers;         /* num_hidden_layers of the draft */
  int n_q_heads;        /* num_attention_heads */
  int n_kv_heads;       /* num_key_value_heads */
  int head_dim;         /* must be 128 */
  int vocab;            /* rows of the target's lm_head / embed_tokens */
  int n_sel;            /* len(target_layer_ids) */
  int block_size;       /* slots per block (2..32); slot 0 is the last committed token */
  int max_requests;     /* request streams resident in this engine, 1..64. Up to 256 activation rows ride on one
                           UMMA; wider batches run as column groups of 256 rows inside the same launch (the groups
                           of one weight range share it through L2) */
  int max_seq;          /* positions per request in the static draft KV cache */
  int out_len;          /* row length of output_ids (>= prompt + max_new_tokens + block_size) */
  int hist_len;         /* acceptance-length history entries per request */
  float rms_eps;
  float rope_scale;     /* rotary attention_scaling (1.0 for default rope) */
  long long mask_token_id;
  int attn_splits;      /* KV splits of the draft attention (0 = about two waves of CTAs, 1..16) */
  int post_splits;      /* vocab splits of the posterior sampler (0 = about two CTAs per SM in all, 1..32) */
  int gemm_grid;        /* CTAs per streaming GEMM (0 = SM count) */
  int use_pdl;          /* programmatic dependent launch between the step's kernels */
  int keep_draft_logits;/* also store the draft's bf16 logits (parity tests) */
  int max_candidates;   /* candidate blocks per request for multi-candidate verify: 0/1 = off, up to 4 (needs
                           max_requests * row slots <= 32: the lm_head epilogue then keeps a top-4 per row) */
} dflash_config_t;

/* Packed bf16 weights of one draft layer. wqkv = [q_proj; k_proj; v_proj] rows, wgu = [gate; up].
 * bqkv / bo: the projections' biases when config.attention_bias is set (model/dflash.py:41-50), else NULL. */
typedef struct dflash_layer_weights {
  const void* wqkv;   /* [(Hq+2Hkv)*128, hidden] */
  const void* wo;     /* [hidden, Hq*128] */
  const void* wgu;    /* [2*intermediate, hidden] */
  const void* wd;     /* [hidden, intermediate] */
  const void* ln1;    /* input_layernorm.weight [hidden] */
  const void* ln2;    /* post_attention_layernorm.weight [hidden] */
  const void* q_norm; /* [128] */
  const void* k_norm; /* [128] */
  const void* bqkv;   /* [(Hq+2Hkv)*128] = [q_proj.bias; k_proj.bias; v_proj.bias], or NULL */
  const void* bo;     /* [hidden] o_proj.bias, or NULL */
} dflash_layer_weights_t;

typedef struct dflash_weights {
  const void* embed;        /* target.model.embed_tokens.weight [vocab, hidden] */
  const void* lm_head;      /* target.lm_head.weight [vocab, hidden] */
  const void* fc;           /* fc.weight [hidden, n_sel*hidden] */
  const void* hidden_norm;  /* [hidden] */
  const void* final_norm;   /* norm.weight [hidden] */
  const float* inv_freq;    /* rotary_emb.inv_freq [64] fp32 */
  const dflash_layer_weights_t* layers_host; /* HOST array of n_layers entries */
} dflash_weights_t;

/* Named regions of the caller-owned workspace (dflash_engine_buffer). */
enum dflash_buffer_id {
  DFLASH_BUF_X = 0,        /* bf16 [R*SL, hidden] residual stream of the block rows */
  DFLASH_BUF_A_IN,         /* bf16 [2*R*SL, hidden] context rows then normed block rows */
  DFLASH_BUF_CTX_FEAT,     /* bf16 [R*SL, n_sel*hidden] pending target-context features (filled by the gathering forms
                              only: dflash_verify_step, the candidate verify, forward(target_hidden=...);
                              dflash_verify_inject_step reads the hidden states in place) */
  DFLASH_BUF_Q,            /* bf16 [R*SL, Hq, 128] */
  DFLASH_BUF_ATTN_OUT,     /* bf16 [R*SL, Hq*128] */
  DFLASH_BUF_A2,           /* bf16 [R*SL, hidden] */
  DFLASH_BUF_HMID,         /* bf16 [R*SL, intermediate] */
  DFLASH_BUF_HN,           /* bf16 [R*SL, hidden] final-normed hidden = DFlashDraftModel.forward output */
  DFLASH_BUF_KV,           /* bf16 [n_layers, 2, R, Hkv, max_seq, 128] static draft KV cache */
  DFLASH_BUF_WS,           /* fp32 split-K partial planes of fc / qkv / o / down (consumed by the row kernels) */
  DFLASH_BUF_PART,         /* fp32 partial accumulators exchanged between the CTAs of a split gate/up tile */
  DFLASH_BUF_FLAGS,        /* uint32 arrival counters per (layer, column group, gate/up tile); zero between launches */
  DFLASH_BUF_COUNTERS,     /* uint32 [R + 2] arrival counters of the verify kernel and of the token reduce */
  DFLASH_BUF_ATTN_PO,
  DFLASH_BUF_ATTN_ML,
  DFLASH_BUF_CAND_VAL,
  DFLASH_BUF_CAND_IDX,
  DFLASH_BUF_POST_VAL,
  DFLASH_BUF_POST_IDX,
  DFLASH_BUF_DRAFT_TOKENS, /* int64 [R*SL] argmax of every block row */
  DFLASH_BUF_BLOCK_IDS,    /* int64 [R, block_size] block tokens (feeds the target's verify forward) */
  DFLASH_BUF_POSTERIOR,    /* int64 [R, block_size] */
  DFLASH_BUF_OUTPUT_IDS,   /* int64 [R, out_len] */
  DFLASH_BUF_START,        /* int32 [R] committed length == both cache lengths */
  DFLASH_BUF_CTX_LEN,      /* int32 [R] pending context rows (= previous tau) */
  DFLASH_BUF_DONE,         /* int32 [R] */
  DFLASH_BUF_N_CYCLES,     /* int32 [R] */
  DFLASH_BUF_BLK_LEN,      /* int32 [R] effective block length of the next cycle */
  DFLASH_BUF_MAX_LEN,      /* int32 [R] prompt length + max_new_tokens */
  DFLASH_BUF_ACC_HIST,     /* int32 [R, hist_len] tau per cycle */
  DFLASH_BUF_RNG_STEP,     /* uint64 [1] */
  DFLASH_BUF_DRAFT_LOGITS, /* bf16 [R*SL, vocab] when keep_draft_logits */
  DFLASH_BUF_PF_FEAT,      /* bf16 [256, n_sel*hidden] prompt pass: gathered target features */
  DFLASH_BUF_PF_A,         /* bf16 [256, hidden] prompt pass: hidden_norm(fc(features)) */
  DFLASH_BUF_TOPK_IDX,     /* int32 [R*SL, 4] top-4 vocab indices of every block row's draft logits */
  DFLASH_BUF_TOPK_VAL,     /* fp32 [R*SL, 4] their bf16-rounded logits */
  DFLASH_BUF_CAND_IDS,     /* int64 [R, 4, block_size] candidate blocks (feeds the target's batched verify forward) */
  DFLASH_BUF_CAND_SCORES,  /* fp32 [R, 4] draft score per candidate */
  DFLASH_BUF_CHOSEN,       /* int32 [R] candidate committed by the last verify step */
  DFLASH_BUF_COUNT
};

typedef struct dflash_engine dflash_engine_t;

/* Bytes of device workspace an engine of this config needs (0 + error on a bad config). */
size_t dflash_workspace_bytes(const dflash_config_t* cfg_host);

/* Builds the engine (host object: TMA descriptors, kernel schedule) over caller-owned memory.
 * `workspace` must be 1024-byte aligned, zero-initialised device memory. */
int dflash_engine_create(const dflash_config_t* cfg_host, const dflash_weights_t* weights_host,
                         void* workspace, size_t workspace_bytes, dflash_engine_t** out_host);
void dflash_engine_destroy(dflash_engine_t* e);

/* Device pointer + size of a named workspace region. */
int dflash_engine_buffer(const dflash_engine_t* e, int buffer_id, void** ptr_out_host, size_t* bytes_out_host);

/* Prompt context for request r: P rows of each selected target hidden state
 * (hidden_host[s] -> device [P, hidden] bf16) are projected (fc + hidden_norm) and their per-layer
 * K/V written to cache positions [0, P), in passes of up to 256 rows (M = P GEMMs: the weights are
 * streamed once per 256 prompt rows). Sets start[r] = P, ctx_len[r] = 0. Other requests of the engine
 * are not disturbed (the pass has its own feature / activation buffers).
 * Replaces cycle 0 of model/dflash.py:229,238-246 (c = P) and the fc/hidden_norm/k_proj/v_proj/
 * k_norm/RoPE call sites at :73-82,177. */
int dflash_prefill_context(dflash_engine_t* e, int r, const void* const* hidden_host, int P, void* stream);

/* Same, for n_rows context rows that continue an existing cache: positions [pos0, pos0 + n_rows)
 * (a forward(target_hidden=...) call whose context is longer than one block: benchmark.py:122-129).
 * Sets start[r] = pos0 + n_rows, ctx_len[r] = 0. */
int dflash_prefill_context_at(dflash_engine_t* e, int r, const void* const* hidden_host, int n_rows, int pos0,
                              void* stream);

/* One draft step for all requests: embed(block_ids) -> ctx injection of the pending context rows
 * -> n_layers x (QKV, attention over [ctx | block], O, SwiGLU MLP) -> norm -> lm_head + argmax.
 * Writes the drafted tokens to block_ids[:, 1:]; the final-normed hidden (what
 * DFlashDraftModel.forward returns) stays in DFLASH_BUF_HN.
 * noise_embedding: NULL to gather target.embed_tokens(block_ids) in-kernel, or device bf16
 *   [R*SL, hidden] rows supplied by a forward(noise_embedding=...) caller.
 * run_lm_head: 0 stops after the final norm (plain forward()).
 * Replaces model/dflash.py:235-247 (embed_tokens, DFlashDraftModel.forward, target.lm_head,
 * past_key_values_draft.crop, sample). */
int dflash_draft_step(dflash_engine_t* e, const void* noise_embedding, int run_lm_head, void* stream);

/* Verify step after the target's forward over block_ids:
 *   posterior = sample(target_logits, temperature); acceptance = longest prefix with
 *   block[1:] == posterior[:-1]; commit accepted tokens + bonus token to output_ids; advance
 *   start (== crop of both caches); stop check; block_ids <- [bonus, mask...]; gather the first
 *   tau rows of the selected hidden states as the next context features.
 * target_logits: [R*block_size, vocab] bf16, row pitch logits_ld elements (or NULL when
 *   posterior_in [R, block_size] int64 holds already-sampled tokens).
 * hidden_host[s]: device [R*block_size, hidden] bf16 for each selected target layer.
 * noise: optional fp32 [R*block_size, vocab] Exp(1) draws (torch.multinomial's race), else Philox.
 * forced_k: optional int32 [R, forced_ld] harness hook: posterior[:k] = block[1:k+1] (k indexed by cycle).
 * Replaces model/dflash.py:257-268 and model/utils.py:16-34. */
int dflash_verify_step(dflash_engine_t* e, const void* target_logits, long long logits_ld,
                       const long long* posterior_in, const void* const* hidden_host, float temperature,
                       const float* noise, unsigned long long seed, const long long* stop_ids, int n_stop,
                       const int* forced_k, int forced_ld, int clamp_tail, void* stream);

/* The verify step AND the next cycle's target-context injection (north-star subsystems (d) + (a)) as two kernels
 * that overlap: dflash_verify_step's kernel, and behind it the context-injection kernel reading the first tau rows of
 * the selected hidden states IN PLACE through one tensor map per selected layer -- the concatenation of
 * extract_context_feature (model/utils.py:16-25), the fc projection, hidden_norm (model/dflash.py:177), the next
 * block's embedding and layer 0's input_layernorm (:237) in ONE kernel whose fc main loop runs while the verify kernel
 * is still sampling and accepting; only its row pass waits for the accepted lengths. Same arguments as
 * dflash_verify_step; hidden_host[s] must stay valid until the work enqueued here has run (they are re-read, not
 * copied). The next draft step is then dflash_draft_step_injected.
 * Replaces model/dflash.py:257-268 plus, for the next cycle, :263 (extract_context_feature), :177 and :237. */
int dflash_verify_inject_step(dflash_engine_t* e, const void* target_logits, long long logits_ld,
                              const long long* posterior_in, const void* const* hidden_host, float temperature,
                              const float* noise, unsigned long long seed, const long long* stop_ids, int n_stop,
                              const int* forced_k, int forced_ld, int clamp_tail, void* stream);

/* dflash_draft_step without its first kernel: the context rows and the block rows were already injected by
 * dflash_verify_inject_step (a caller that rewrites block_ids in between -- a request reset -- runs
 * dflash_embed_block first). temperature >= 1e-5 samples the drafted tokens as dflash_draft_step_sampled does.
 * Replaces model/dflash.py:238-247. */
int dflash_draft_step_injected(dflash_engine_t* e, int run_lm_head, float temperature, unsigned long long seed,
                               void* stream);

/* Block rows only: embed_tokens(block_ids) -> residual stream, layer 0's input_layernorm -> activation rows
 * (model/dflash.py:237). For callers that rewrite block_ids before a dflash_draft_step_injected. */
int dflash_embed_block(dflash_engine_t* e, void* stream);

/* Kernel launches one call enqueues: which = 0 dflash_draft_step, 1 dflash_draft_step_injected, 2 dflash_verify_step,
 * 3 dflash_verify_inject_step (plain verify path). */
int dflash_engine_launches(const dflash_engine_t* e, int which);

/* dflash_draft_step with the drafted tokens SAMPLED from softmax(draft_logits / temperature) instead of argmax-ed
 * (the reference's policy loop does that: benchmark_dynamic_schedule.py:342; spec_generate itself always drafts
 * greedily). Gumbel-max inside the lm_head epilogue -- key = bf16(logit) / T + Gumbel(Philox(seed, cycle, row, vocab))
 * -- so the logits still never reach HBM. temperature < 1e-5 is the plain greedy step. */
int dflash_draft_step_sampled(dflash_engine_t* e, float temperature, unsigned long long seed, void* stream);

/* Multi-candidate drafting ("fixed_prefix_rank", benchmark_candidate_solutions.py:181-249): the draft step with a
 * top-4 lm_head epilogue; builds n_candidates (2..4) candidate blocks per request in DFLASH_BUF_CAND_IDS: candidate 0
 * is the greedy block, candidate k keeps the first fixed_prefix_len positions and takes the rank-(k+1) token at every
 * later position; DFLASH_BUF_CAND_SCORES holds their draft scores. block_ids[:, 1:] gets the greedy tokens as usual. */
int dflash_draft_step_candidates(dflash_engine_t* e, int n_candidates, int fixed_prefix_len, void* stream);

/* Verify after ONE target forward over all candidates (batch n_candidates per request):
 * target_logits [R*n_candidates*block_size, vocab], hidden_host[s] [R*n_candidates*block_size, hidden]; rows of
 * candidate k of request r start at (r*n_candidates + k)*block_size. Commits the candidate with the longest accepted
 * prefix (ties: higher draft score, then lower index -- the reference's composite, :597-604), writes its index to
 * DFLASH_BUF_CHOSEN (the caller keeps that branch of the target's KV cache, :610-614) and gathers the next context
 * features from its rows. Greedy or sampled posterior as dflash_verify_step. */
int dflash_verify_step_candidates(dflash_engine_t* e, int n_candidates, const void* target_logits, long long logits_ld,
                                  const void* const* hidden_host, float temperature, const float* noise,
                                  unsigned long long seed, const long long* stop_ids, int n_stop, int clamp_tail,
                                  void* stream);

/* tokens_out[row] = sample(logits[row, :], temperature) for standalone use (prefill's first token). */
int dflash_sample(const void* logits, long long logits_ld, int rows, int vocab, float temperature,
                  const float* noise, unsigned long long seed, float* scratch_val, int* scratch_idx,
                  int nsplit, long long* tokens_out, void* stream);

/* =============================================================================================
 * Raw operators (unit-test granularity).
 * ===========================================================================================*/

/* out[m, n] = sum_k X[x_row0+m, k] * W[w_row0+n, k]   (fp32 out; bf16 in; nn.Linear layout)
 * Replaces torch.nn.Linear on the draft path (model/dflash.py:70-76,101,177; Qwen3MLP via :143).
 * mb = activation rows per tensor-core instruction (16/32/64/128/256); m_valid rows are written. m_valid > mb
 * runs ceil(m_valid/mb) column groups in one launch over grid/groups weight ranges (X must hold groups*mb rows).
 * ws: fp32 scratch [dflash_gemm_max_slots(N,K,grid/groups)][ws_rows >= groups*mb][ws_ld] for the split-K partials. */
int dflash_gemm_max_slots(int N, int K, int grid);
/* CTAs dflash_gemm_argmax actually launches for N weight rows when offered `grid`: whole 128-row tiles per CTA, the
 * smallest grid with the same busiest CTA (151936 rows, 148 -> 132 CTAs of 9 tiles). */
int dflash_gemm_argmax_grid(int N, int grid);
int dflash_gemm_skinny(const void* W, int w_rows_total, int w_row0, int N, int K, const void* X,
                       int x_rows_total, int x_row0, int mb, int m_valid, float* ws, int ws_rows,
                       long long ws_ld, float* out, long long out_ld, int grid, int use_pdl,
                       void* stream);

/* The gate/up GEMM with its fused SwiGLU epilogue, as a raw operator:
 *   out[m, i] = bf16(bf16(silu(gate[m, i])) * up[m, i]),  Wgu = [gate_proj; up_proj] stacked (rows [0, I), [I, 2I)),
 *   I a multiple of 64. A tile = 64 gate rows + 64 up rows of the same columns; a tile that several CTAs share
 *   (stream-K) is finished INSIDE the launch: `part` (fp32, grid * 128 * mb elements) carries the partial
 *   accumulators, `flags` (uint32, ceil(m_valid/mb) * I/64, zero on entry and on exit) the arrivals.
 * Replaces Qwen3MLP's act_fn(gate_proj(x)) * up_proj(x) (transformers, via model/dflash.py:143). */
int dflash_gemm_swiglu(const void* Wgu, int I, int K, const void* X, int x_rows_total, int mb, int m_valid, void* out,
                       long long ld, float* part, unsigned int* flags, int grid, int use_pdl, void* stream);

/* tokens_out[m] = argmax_n bf16(sum_k X[x_row0+m,k] * W[n,k]), ties -> lowest n; optionally also
 * writes the bf16 logits. Replaces target.lm_head(...) + sample(draft_logits)
 * (model/dflash.py:238-247, model/utils.py:27-29) without materialising the logits.
 * cand_val/cand_idx: scratch [grid][groups*mb]. */
int dflash_gemm_argmax(const void* W, int w_rows_total, int N, int K, const void* X, int x_rows_total,
                       int x_row0, int mb, int m_valid, float* cand_val, int* cand_idx, void* logits,
                       long long logits_ld, long long* tokens_out, int grid, int use_pdl, void* stream);

/* tokens_out[m] ~ softmax(bf16(X[m] . W^T) / temperature): the Gumbel-max epilogue as a raw operator (mb <= 32);
 * `step` selects the noise stream (the engine uses its cycle counter). */
int dflash_gemm_sample(const void* W, int w_rows_total, int N, int K, const void* X, int x_rows_total, int x_row0,
                       int mb, int m_valid, float temperature, unsigned long long seed, unsigned long long step,
                       float* cand_val, int* cand_idx, long long* tokens_out, int grid, int use_pdl, void* stream);

/* Debug only (scripts/gemm_trace.py): dflash_gemm_skinny without the slot sum, recording per-CTA phase timestamps
 * (globaltimer ns) into trace[ranges * groups][8]: kernel entry, prologue done, producer past griddepcontrol.wait,
 * first stage landed, last MMA issued, last accumulator complete, epilogue stores issued. Returns the number of weight
 * ranges (> 0) or a negative error. */
int dflash_gemm_trace(const void* W, int N, int K, const void* X, int x_rows_total, int mb, int m_valid, float* ws,
                      int ws_rows, unsigned long long* trace, int grid, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFLASH_B200_H_ */
