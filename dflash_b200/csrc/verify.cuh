// Verify step: posterior sampling over the target logits, longest-matching-prefix acceptance,
// token commit (+ bonus token), cache-length rollback, stop check, next-block setup and the gather
// of the next cycle's context features  (model/dflash.py:257-268, model/utils.py:16-34).
//
// All state lives on the device (start / ctx_len / done per request) so the whole draft+verify
// step replays from a CUDA graph with no host round trip. With static caches the reference's two
// crops (dflash.py:246,262) are length writes: rows past `start` are dead and get overwritten.
#pragma once
#include "ptx.cuh"

namespace dfl {

struct PosteriorArgs {
  const __nv_bfloat16* logits;  // [rows][ld]
  long long ld;
  int rows, V, nsplit;
  float inv_temp;           // 0 -> greedy argmax (temperature < 1e-5, utils.py:28)
  const float* noise;       // optional Exp(1) draws [rows][V] (torch.multinomial's q); null -> Philox
  unsigned long long seed;
  const unsigned long long* rng_step;  // device counter, bumped by accept_kernel every cycle
  float* cand_val;          // [rows][nsplit]
  int* cand_idx;
};

// key_i = logit_i (greedy) or logit_i / T - log(e_i), e_i ~ Exp(1): argmax_i key_i is a draw from
// softmax(logits / T) -- the same exponential race torch.multinomial runs (argmax(p / q)).
__device__ __forceinline__ void posterior_body(const PosteriorArgs& a, int row, int split) {
  int seg = (a.V + a.nsplit - 1) / a.nsplit;
  seg = (seg + 7) & ~7;
  const int c0 = split * seg;
  const int c1 = min(a.V, c0 + seg);
  const __nv_bfloat16* lp = a.logits + static_cast<long long>(row) * a.ld;
  const bool greedy = a.inv_temp == 0.f;
  const unsigned long long step = (a.rng_step != nullptr && !greedy && a.noise == nullptr) ? *a.rng_step : 0ull;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(lp) & 15) == 0);
  auto consider = [&](float logit, int col, uint32_t rnd) {
    float key = logit;
    if (!greedy) {
      float e;
      if (a.noise != nullptr) e = a.noise[static_cast<long long>(row) * a.V + col];
      else e = -__logf(u32_to_unit(rnd));
      key = logit * a.inv_temp - logf(fmaxf(e, 1e-30f));
    }
    if (key > bv) { bv = key; bi = col; }
  };
  if (vec_ok) {
    // kU 128-bit loads in flight per thread: at 64 streams this pass reads 311 MB of logits (one load per iteration
    // left it at 2.3 TB/s). Columns are still visited in ascending order per thread (ties keep the lowest index).
    constexpr int kU = 4;
    for (int cb = c0 + threadIdx.x * 8; cb < c1; cb += 256 * 8 * kU) {
      uint4 raw[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int c = cb + u * 256 * 8;
        if (c + 8 <= c1) raw[u] = *reinterpret_cast<const uint4*>(lp + c);
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int c = cb + u * 256 * 8;
        if (c >= c1) break;
        uint32_t rnd[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (!greedy && a.noise == nullptr) {
          philox4x32(static_cast<uint32_t>(c), static_cast<uint32_t>(row), static_cast<uint32_t>(step),
                     static_cast<uint32_t>(step >> 32), static_cast<uint32_t>(a.seed),
                     static_cast<uint32_t>(a.seed >> 32), rnd);
          philox4x32(static_cast<uint32_t>(c + 4), static_cast<uint32_t>(row), static_cast<uint32_t>(step),
                     static_cast<uint32_t>(step >> 32), static_cast<uint32_t>(a.seed),
                     static_cast<uint32_t>(a.seed >> 32), rnd + 4);
        }
        if (c + 8 <= c1) {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[u]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(h[j]);
            consider(f.x, c + 2 * j, rnd[2 * j]);
            consider(f.y, c + 2 * j + 1, rnd[2 * j + 1]);
          }
        } else {  // (unrolled with constant indices: a runtime index would put rnd[] in local memory for every path)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c + j < c1) consider(__bfloat162float(lp[c + j]), c + j, rnd[j]);
        }
      }
    }
  } else {
    for (int c = c0 + threadIdx.x; c < c1; c += 256) {
      uint32_t rnd[4] = {0, 0, 0, 0};
      if (!greedy && a.noise == nullptr)
        philox4x32(static_cast<uint32_t>(c), static_cast<uint32_t>(row), static_cast<uint32_t>(step),
                   static_cast<uint32_t>(step >> 32) | 0x80000000u, static_cast<uint32_t>(a.seed),
                   static_cast<uint32_t>(a.seed >> 32), rnd);
      consider(__bfloat162float(lp[c]), c, rnd[0]);
    }
  }
  // block argmax, ties -> lowest column (torch.argmax on CPU; CUDA ties are unspecified)
  __shared__ float sv[8];
  __shared__ int si[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
    a.cand_val[row * a.nsplit + split] = bv;
    a.cand_idx[row * a.nsplit + split] = bi;
  }
}

__global__ void __launch_bounds__(256) posterior_kernel(const PosteriorArgs a) {
  DFL_VERIFY_SYNC();
  posterior_body(a, blockIdx.y, blockIdx.x);
}

// tokens[row] = best candidate over the splits of posterior_kernel (standalone sampler)
__global__ void __launch_bounds__(32) sample_reduce_kernel(const float* __restrict__ cand_val,
                                                           const int* __restrict__ cand_idx, int nsplit,
                                                           long long* __restrict__ tokens) {
  const int row = blockIdx.x;
  if (threadIdx.x != 0) return;
  float bv = cand_val[row * nsplit];
  int bi = cand_idx[row * nsplit];
  for (int s = 1; s < nsplit; ++s) {
    const float v = cand_val[row * nsplit + s];
    const int ix = cand_idx[row * nsplit + s];
    if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
  }
  tokens[row] = bi;
}

struct AcceptArgs {
  int R, bs, nsplit;
  const float* cand_val;  // [R*bs][nsplit] from posterior_kernel; ignored if posterior_in != null
  const int* cand_idx;
  const long long* posterior_in;  // optional [R][bs]: already-sampled posterior tokens
  long long* posterior;           // [R][bs] out
  long long* block_ids;           // [R][ids_ld]; in: block tokens (slot 0 committed, 1.. drafted)
  int ids_ld;
  long long* output_ids;          // [R][out_ld], pre-filled with mask_token
  long long out_ld;
  int* start;
  int* ctx_len;
  int* done;
  int* n_cycles;
  int* blk_len;                   // [R] effective block length of the NEXT cycle (tail clamp)
  int* acc_hist;                  // [R][hist_ld] tau per cycle (acceptance_length + 1)
  int hist_ld;
  const int* max_len;             // [R] prompt length + max_new_tokens
  const long long* stop_ids;
  int n_stop;
  long long mask_token;
  const int* forced_k;            // optional [R][forced_ld] test/bench hook: posterior[:k] = block[1:k+1]
  int forced_ld;
  int clamp_tail;                 // benchmark.py:104-105 effective block size at the tail
  unsigned long long* rng_step;
  // multi-candidate verify (benchmark_candidate_solutions.py:590-625): K candidate blocks per request were verified
  // in one target call; posterior rows are (r*K + k)*bs + i. K <= 1: the plain path.
  int K;
  const long long* cand_ids;      // [R][4][bs] candidate blocks (k = 0 is the greedy block)
  const float* cand_scores;       // [R][4] draft score of each candidate
  int* chosen;                    // [R] out: index of the committed candidate
};

// One warp per request.
__device__ __forceinline__ void accept_request(const AcceptArgs& a, int r, int lane) {
  // every scalar of the request's state is requested up front: one L2 round trip instead of one per use
  const int done = a.done[r];
  const int st = a.start[r];
  const int cyc = a.n_cycles[r];
  const int eff = a.blk_len[r];  // == bs unless the tail clamp shortened this block
  const int max_len = a.max_len[r];
  const int fk = a.forced_k != nullptr ? a.forced_k[r * a.forced_ld + (cyc % a.forced_ld)] : 0;
  if (done) return;
  const int bs = a.bs;
  const int K = a.K > 1 ? a.K : 1;
  long long* blk = a.block_ids + static_cast<long long>(r) * a.ids_ld;
  __shared__ long long s_post[4][64];
  __shared__ long long s_btok[4][64];
  for (int k = 0; k < K; ++k) {
    for (int i = lane; i < bs; i += 32) {
      long long p;
      if (a.posterior_in != nullptr) {
        p = a.posterior_in[(r * K + k) * bs + i];
      } else {
        const int row = (r * K + k) * bs + i;
        const float* cv = a.cand_val + static_cast<long long>(row) * a.nsplit;
        const int* ci = a.cand_idx + static_cast<long long>(row) * a.nsplit;
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int s0 = 0; s0 < a.nsplit; s0 += 8) {  // eight candidates in flight per round trip
          float v[8];
          int ix[8];
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (s0 + q < a.nsplit) { v[q] = cv[s0 + q]; ix[q] = ci[s0 + q]; }
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (s0 + q < a.nsplit && (v[q] > bv || (v[q] == bv && ix[q] < bi))) { bv = v[q]; bi = ix[q]; }
        }
        p = bi;
      }
      s_post[k][i] = p;
      s_btok[k][i] = (a.K > 1) ? a.cand_ids[(static_cast<long long>(r) * 4 + k) * bs + i] : blk[i];
    }
  }
  __syncwarp();
  if (a.forced_k != nullptr) {
    for (int i = lane; i < bs - 1; i += 32)
      if (i < fk) s_post[0][i] = s_btok[0][i + 1];
    __syncwarp();
  }
  // candidate choice: maximise tau, then the draft score, then the lower index -- with the reference's own fp32
  // composite  tau * 1e6 + score - idx * 1e-3  (benchmark_candidate_solutions.py:597-604), first maximum wins
  int kc = 0;
  if (K > 1) {
    float best = -INFINITY;
    for (int k = 0; k < K; ++k) {
      int acc = 0;
      while (acc < eff - 1 && s_btok[k][acc + 1] == s_post[k][acc]) ++acc;
      const float comp = __fsub_rn(__fadd_rn(__fmul_rn(static_cast<float>(acc + 1), 1e6f), a.cand_scores[r * 4 + k]),
                                   __fmul_rn(static_cast<float>(k), 1e-3f));
      if (comp > best) { best = comp; kc = k; }
    }
    if (lane == 0 && a.chosen != nullptr) a.chosen[r] = kc;
  }
  const long long* post = s_post[kc];
  const long long* btok = s_btok[kc];
  for (int i = lane; i < bs; i += 32) a.posterior[r * bs + i] = post[i];
  if (lane != 0) return;
  int acc = 0;  // (blk[1:] == post[:-1]).cumprod().sum()
  while (acc < eff - 1 && btok[acc + 1] == post[acc]) ++acc;
  long long* out = a.output_ids + static_cast<long long>(r) * a.out_ld;
  bool stop = false;
  for (int i = 0; i <= acc; ++i) {
    out[st + i] = btok[i];
    for (int s = 0; s < a.n_stop; ++s) stop |= (btok[i] == a.stop_ids[s]);
  }
  const long long bonus = post[acc];
  out[st + acc + 1] = bonus;
  for (int s = 0; s < a.n_stop; ++s) stop |= (bonus == a.stop_ids[s]);
  const int nst = st + acc + 1;
  a.start[r] = nst;
  a.ctx_len[r] = acc + 1;
  if (cyc < a.hist_ld) a.acc_hist[r * a.hist_ld + cyc] = acc + 1;
  a.n_cycles[r] = cyc + 1;
  if (stop || nst >= max_len) a.done[r] = 1;
  if (a.clamp_tail) {
    const int remaining = max_len - nst;
    a.blk_len[r] = remaining < bs ? (remaining < 1 ? 1 : remaining) : bs;
  }
  blk[0] = bonus;
  for (int i = 1; i < bs; ++i) blk[i] = a.mask_token;
}

__global__ void __launch_bounds__(32) accept_kernel(const AcceptArgs a) {
  DFL_VERIFY_SYNC();
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.rng_step != nullptr) *a.rng_step += 1ull;
  accept_request(a, blockIdx.x, threadIdx.x);
}

// Next-cycle context features: the first ctx_len[r] rows of each selected target hidden state,
// concatenated on the feature dim (extract_context_feature + [:, :tau] slice, dflash.py:263).
struct GatherArgs {
  const __nv_bfloat16* src[8];  // per selected layer: [nreq * src_rows][H]
  int n_sel, H, SL;
  int r0, nreq;      // requests [r0, r0 + nreq) are covered by src
  int src_rows;      // rows per request in src
  int src_row0;      // first source row (prefill chunk offset)
  const int* ctx_len;
  __nv_bfloat16* ctx_feat;  // [R*SL][n_sel*H]
  int pf_rows;       // > 0: prompt pass -- source rows [src_row0, src_row0 + pf_rows) -> ctx_feat rows [0, pf_rows)
  const int* chosen; // multi-candidate verify: request rr's rows are those of candidate chosen[r] ([nreq][K][src_rows])
  int K;
};

__global__ void __launch_bounds__(256) ctx_gather_kernel(const GatherArgs a) {
  DFL_VERIFY_SYNC();
  int rr = blockIdx.x / a.SL, j = blockIdx.x % a.SL;
  long long drow;
  if (a.pf_rows > 0) {
    rr = 0;
    j = blockIdx.x;
    if (j >= a.pf_rows) return;
    drow = j;
  } else {
    const int r = a.r0 + rr;
    if (j >= a.ctx_len[r]) return;
    drow = static_cast<long long>(r) * a.SL + j;
    if (a.chosen != nullptr) rr = rr * a.K + a.chosen[r];
  }
  const int sel = blockIdx.y;
  const uint4* s = reinterpret_cast<const uint4*>(
      a.src[sel] + (static_cast<long long>(rr) * a.src_rows + a.src_row0 + j) * a.H);
  uint4* d = reinterpret_cast<uint4*>(a.ctx_feat + drow * a.n_sel * a.H + static_cast<long long>(sel) * a.H);
  for (int i = threadIdx.x; i < a.H / 8; i += 256) d[i] = s[i];
}

// ---------------------------------------------------------------------------------------------
// The whole verify step as ONE kernel (model/dflash.py:257-268): grid = (vocab splits, R * bs block rows).
//   every CTA   copies its share of the next cycle's context features -- ALL bs rows of the selected hidden states
//               go to ctx_feat; the rows past the accepted length are dead (ctx_len masks them), so the copy does not
//               depend on the acceptance -- and reduces its vocab slice of the posterior (argmax / exponential race);
//   last CTA of a request (arrival counter)  runs acceptance, commit, bonus token, both cache-length rollbacks,
//               stop check and the next block for that request;
//   last CTA of the grid  bumps the Philox step.
struct VerifyFusedArgs {
  PosteriorArgs post;
  AcceptArgs acc;
  GatherArgs gather;
  unsigned int* counters;  // [R + 1] arrivals per request, then of the whole grid (self-resetting)
};

// MINB = 4: at most 64 registers, which is what lets two CTAs per SM sit next to the context-injection kernel's CTA
// (small batches: the overlap pays, and the spills this costs in the vocab loop are hidden under that kernel's fc
// stream). MINB = 3: 85 registers (no spills in the loop), for wide batches, where the kernel reads 78-311 MB of logits and is on the critical path.
template <int MINB>
__device__ __forceinline__ void verify_fused_body(const VerifyFusedArgs& v);

__global__ void __launch_bounds__(256, 4) verify_fused_kernel(const VerifyFusedArgs v) { verify_fused_body<4>(v); }
__global__ void __launch_bounds__(256, 3) verify_fused_wide_kernel(const VerifyFusedArgs v) { verify_fused_body<3>(v); }

template <int MINB>
__device__ __forceinline__ void verify_fused_body(const VerifyFusedArgs& v) {
  // wait first, release second: the context-injection kernel behind this one starts its fc main loop without waiting
  // for this kernel, on the strength of "everything in front of the verify kernel (the target's forward, the draft
  // step) is complete once the verify kernel has released its dependent"
  DFL_WAIT_THEN_TRIGGER();
  const int split = blockIdx.x, nsplit = gridDim.x;
  const int bs = v.acc.bs;
  // gridDim.y may be smaller than the number of block rows (wide batches): the grid is kept to about two CTAs per SM
  // so that the context-injection kernel behind this one finds room on every SM while this kernel is still running
  for (int row = blockIdx.y; row < v.post.rows; row += gridDim.y) {
  const int r = row / bs, i = row % bs;
  {
    const GatherArgs& g = v.gather;
    const int per_sel = g.H / 8;                 // uint4 per (row, selected layer)
    const int total = g.n_sel * per_sel;
    const int share = (total + nsplit - 1) / nsplit;
    const int e0 = split * share, e1 = min(total, e0 + share);
    uint4* d = reinterpret_cast<uint4*>(g.ctx_feat + (static_cast<long long>(r) * g.SL + i) * g.n_sel * g.H);
    for (int e = e0 + threadIdx.x; e < e1; e += 256) {
      const int sel = e / per_sel, c = e % per_sel;
      d[e] = reinterpret_cast<const uint4*>(g.src[sel] + static_cast<long long>(row) * g.H)[c];
    }
  }
  posterior_body(v.post, row, split);
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int n_req = static_cast<unsigned int>(nsplit) * bs;
    const unsigned int prev = atomicAdd(&v.counters[r], 1u);
    s_last = (prev == n_req - 1u) ? 1u : 0u;
    if (prev == n_req - 1u) v.counters[r] = 0u;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (threadIdx.x < 32) accept_request(v.acc, r, threadIdx.x);
  }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int n_all = gridDim.x * gridDim.y;
    const unsigned int prev = atomicAdd(&v.counters[v.acc.R], 1u);
    if (prev == n_all - 1u) {
      v.counters[v.acc.R] = 0u;
      if (v.acc.rng_step != nullptr) *v.acc.rng_step += 1ull;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Multi-candidate drafting, "fixed_prefix_rank" mode (benchmark_candidate_solutions.py:181-249): candidate 0 is the
// greedy block; candidate k > 0 keeps the first `prefix_len` block positions and takes the rank-(k+1) token of the
// draft logits at every later position. Input: the per-CTA top-4 lists of the lm_head GEMM (kModeTopK).
struct CandArgs {
  const float* cand_val;   // [n_cta][cand_ld][4], best first
  const int* cand_idx;
  int n_cta, cand_ld;
  int R, SL, bs, prefix_len;
  const int* blk_len;        // [R] effective block length of this cycle
  long long* block_ids;      // [R][bs]: slots 1.. <- rank-1 tokens (as the plain draft step)
  long long* draft_tokens;   // [R*SL]
  int* topk_idx;             // [R*SL][4]
  float* topk_val;           // [R*SL][4] bf16-rounded logits
  long long* cand_ids;       // [R][4][bs]
  float* cand_scores;        // [R][4]: sum of the rank-k logits over the varied positions (bf16-rounded sum)
};

__global__ void __launch_bounds__(256) candidates_kernel(const CandArgs a) {
  DFL_VERIFY_SYNC();
  const int r = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ int s_idx[32][4];
  __shared__ float s_val[32][4];
  for (int i = warp; i < a.SL; i += 8) {
    const int row = r * a.SL + i;
    // lane-local sorted top-4 over this lane's share of the n_cta * 4 per-CTA candidates
    float lv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int li[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
    for (int g = lane; g < a.n_cta * 4; g += 32) {
      const long long o = (static_cast<long long>(g >> 2) * a.cand_ld + row) * 4 + (g & 3);
      float v = __ldcg(a.cand_val + o);
      int ix = __ldcg(a.cand_idx + o);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (v > lv[q] || (v == lv[q] && ix < li[q])) {
          const float tv = lv[q]; lv[q] = v; v = tv;
          const int ti = li[q]; li[q] = ix; ix = ti;
        }
      }
    }
    // four rounds of warp argmax over the lane heads (value desc, vocab index asc); the winner pops its head
    for (int q = 0; q < 4; ++q) {
      float bv = lv[0];
      int bi = li[0];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lv[0] == bv && li[0] == bi) {  // vocab indices are unique: exactly one lane
        lv[0] = lv[1]; li[0] = li[1]; lv[1] = lv[2]; li[1] = li[2]; lv[2] = lv[3]; li[2] = li[3];
        lv[3] = -INFINITY; li[3] = 0x7fffffff;
      }
      if (lane == 0) {
        s_idx[i][q] = bi; s_val[i][q] = bv;
        a.topk_idx[row * 4 + q] = bi;
        a.topk_val[row * 4 + q] = bv;
      }
    }
    if (lane == 0) a.draft_tokens[row] = s_idx[i][0];
  }
  __syncthreads();
  const int eff = a.blk_len[r];
  int suffix_start = a.prefix_len < eff ? a.prefix_len : eff;
  if (suffix_start < 1) suffix_start = 1;
  long long* blk = a.block_ids + static_cast<long long>(r) * a.bs;
  const long long tok0 = blk[0];
  __syncthreads();
  for (int t = threadIdx.x; t < a.bs; t += 256) {
    for (int k = 0; k < 4; ++k) {
      long long tok;
      if (t == 0) tok = tok0;
      else tok = (t < suffix_start || k == 0) ? s_idx[t][0] : s_idx[t][k];
      a.cand_ids[(static_cast<long long>(r) * 4 + k) * a.bs + t] = tok;
    }
    if (t >= 1) blk[t] = s_idx[t][0];
  }
  if (threadIdx.x < 4) {
    float sum = 0.f;
    for (int t = suffix_start; t < eff; ++t) sum += s_val[t][threadIdx.x];
    a.cand_scores[r * 4 + threadIdx.x] = bf16_round(sum);
  }
}

struct SetStateArgs {
  int r;
  int start, ctx_len;
  int* start_p;
  int* ctx_len_p;
};
__global__ void set_state_kernel(const SetStateArgs a) {
  pdl_wait();
  a.start_p[a.r] = a.start;
  a.ctx_len_p[a.r] = a.ctx_len;
}

}  // namespace dfl
