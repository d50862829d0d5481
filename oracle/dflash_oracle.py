"""CPU oracle for the DFlash draft-and-verify hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, operation by operation, what the reference (AtharvRN/dflash, pure Python on
torch + transformers) computes on the path `DFlashDraftModel.spec_generate` drives. It exists so
that the CUDA path can be checked against something that runs anywhere; it is never imported by
the product package (`dflash_b200/`). Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
CPU-baseline / `--impl reference` legs may import it.

Parity pinning: the reference ships no tests or golden vectors. The oracle is
pinned against outputs of the reference itself, generated in the build container by
`tests/golden/make_golden.py` (imports /root/reference/model unchanged) and committed under
`tests/golden/*.pt`; `tests/test_oracle_golden.py` replays them.

Third-party arithmetic the reference delegates to (restated here from the installed sources,
transformers 5.5.0 / torch 2.11.0, both unpinned by the reference's requirements.txt):
  Qwen3RMSNorm        transformers/models/qwen3/modeling_qwen3.py:50-64
  rotate_half         transformers/models/qwen3/modeling_qwen3.py:81-83 (half split, not interleaved)
  Qwen3MLP            transformers/models/qwen3/modeling_qwen3.py (down(silu(gate(x)) * up(x)))
  Qwen3RotaryEmbedding.forward   inv_freq (x) position in fp32, cat(freqs, freqs), cos/sin * scaling,
                                 cast to the activation dtype
  eager_attention_forward        softmax(q k^T * scaling) in fp32, cast back, @ v ; GQA by repeat_kv
  DynamicCache.update / crop     concat on the sequence dim / slice
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# model/utils.py:4-34
# --------------------------------------------------------------------------------------------
def build_target_layer_ids(num_target_layers: int, num_draft_layers: int) -> List[int]:
    """model/utils.py:4-14."""
    if num_draft_layers == 1:
        return [num_target_layers // 2]
    start, end = 1, num_target_layers - 3
    span = end - start
    return [int(round(start + (i * span) / (num_draft_layers - 1))) for i in range(num_draft_layers)]


def extract_context_feature(hidden_states: Sequence[torch.Tensor], layer_ids: Sequence[int]) -> torch.Tensor:
    """model/utils.py:16-25: hidden_states[id + 1] (entry 0 is the embedding output), concatenated on -1."""
    return torch.cat([hidden_states[i + 1] for i in layer_ids], dim=-1)


def sample(logits: torch.Tensor, temperature: float = 0.0) -> torch.Tensor:
    """model/utils.py:27-34."""
    if temperature < 1e-5:
        return torch.argmax(logits, dim=-1)
    bsz, seq_len, vocab = logits.shape
    probs = torch.softmax(logits.view(-1, vocab) / temperature, dim=-1)
    return torch.multinomial(probs, num_samples=1).view(bsz, seq_len)


# --------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------
def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """Qwen3RMSNorm: fp32 variance, cast back to the input dtype BEFORE the weight multiply."""
    dt = x.dtype
    h = x.to(torch.float32)
    var = h.pow(2).mean(-1, keepdim=True)
    h = h * torch.rsqrt(var + eps)
    return weight * h.to(dt)


def rope_cos_sin(inv_freq: torch.Tensor, position_ids: torch.Tensor, scaling: float, dtype: torch.dtype):
    """Qwen3RotaryEmbedding.forward: [B, S] positions -> cos, sin [B, S, D] in `dtype`."""
    freqs = position_ids[:, :, None].to(torch.float32) * inv_freq[None, None, :].to(torch.float32)
    emb = torch.cat((freqs, freqs), dim=-1)
    return (emb.cos() * scaling).to(dtype), (emb.sin() * scaling).to(dtype)


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary_pos_emb(q, k, cos, sin):
    """model/dflash.py:22-28: q uses the LAST q_len positions, k all of them. [B, h, S, D] layout."""
    cos = cos.unsqueeze(1)
    sin = sin.unsqueeze(1)
    q_len = q.size(-2)
    q_embed = (q * cos[..., -q_len:, :]) + (rotate_half(q) * sin[..., -q_len:, :])
    k_embed = (k * cos) + (rotate_half(k) * sin)
    return q_embed, k_embed


@dataclass
class DraftConfig:
    """The Qwen3Config keys DFlashDraftModel reads (model/dflash.py:33-56,157-163)."""
    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    head_dim: int
    rms_norm_eps: float
    block_size: int
    mask_token_id: int
    target_layer_ids: List[int]
    rope_theta: float = 1_000_000.0
    attention_scaling: float = 1.0
    inv_freq: Optional[torch.Tensor] = None  # overrides rope_theta (llama3 scaling etc.)

    def get_inv_freq(self) -> torch.Tensor:
        if self.inv_freq is not None:
            return self.inv_freq.to(torch.float32)
        d = self.head_dim
        return 1.0 / (self.rope_theta ** (torch.arange(0, d, 2, dtype=torch.int64).to(torch.float32) / d))

    @staticmethod
    def from_hf(model) -> "DraftConfig":
        c = model.config
        return DraftConfig(
            hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
            num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads,
            num_key_value_heads=c.num_key_value_heads,
            head_dim=getattr(c, "head_dim", c.hidden_size // c.num_attention_heads),
            rms_norm_eps=c.rms_norm_eps, block_size=model.block_size, mask_token_id=model.mask_token_id,
            target_layer_ids=list(model.target_layer_ids),
            attention_scaling=float(getattr(model.rotary_emb, "attention_scaling", 1.0)),
            inv_freq=model.rotary_emb.inv_freq.detach().clone().to(torch.float32))


@dataclass
class DraftCache:
    """Per-layer context K/V, [B, Hkv, S, D]: what DynamicCache holds after crop."""
    keys: List[Optional[torch.Tensor]] = field(default_factory=list)
    values: List[Optional[torch.Tensor]] = field(default_factory=list)

    def get_seq_length(self) -> int:
        return 0 if not self.keys or self.keys[0] is None else int(self.keys[0].shape[-2])

    def update(self, k, v, layer):
        while len(self.keys) <= layer:
            self.keys.append(None)
            self.values.append(None)
        if self.keys[layer] is None:
            self.keys[layer], self.values[layer] = k, v
        else:
            self.keys[layer] = torch.cat([self.keys[layer], k], dim=-2)
            self.values[layer] = torch.cat([self.values[layer], v], dim=-2)
        return self.keys[layer], self.values[layer]

    def crop(self, n: int):
        for i in range(len(self.keys)):
            if self.keys[i] is not None:
                self.keys[i] = self.keys[i][..., :n, :]
                self.values[i] = self.values[i][..., :n, :]


ATTN_IMPL = "eager"  # "sdpa" = F.scaled_dot_product_attention, what the reference dispatches to by default


def _lin(x, w, b=None):
    return F.linear(x, w, b)


def _proj(sd, pfx, name, x):
    """nn.Linear of the attention block; its bias exists iff config.attention_bias (model/dflash.py:41-50)."""
    return _lin(x, sd[pfx + f"self_attn.{name}.weight"], sd.get(pfx + f"self_attn.{name}.bias"))


def draft_attention(sd, pfx, cfg: DraftConfig, hidden, target_hidden, cos, sin, cache: Optional[DraftCache], layer):
    """Qwen3DFlashAttention.forward, model/dflash.py:58-102 (eager attention, no mask, non-causal)."""
    bsz, q_len = hidden.shape[:-1]
    ctx_len = target_hidden.shape[1]
    D, Hq, Hkv = cfg.head_dim, cfg.num_attention_heads, cfg.num_key_value_heads
    q = _proj(sd, pfx, "q_proj", hidden).view(bsz, q_len, -1, D)
    q = rms_norm(q, sd[pfx + "self_attn.q_norm.weight"], cfg.rms_norm_eps).transpose(1, 2)
    k_ctx = _proj(sd, pfx, "k_proj", target_hidden)
    k_noise = _proj(sd, pfx, "k_proj", hidden)
    v_ctx = _proj(sd, pfx, "v_proj", target_hidden)
    v_noise = _proj(sd, pfx, "v_proj", hidden)
    k = torch.cat([k_ctx, k_noise], dim=1).view(bsz, ctx_len + q_len, -1, D)
    v = torch.cat([v_ctx, v_noise], dim=1).view(bsz, ctx_len + q_len, -1, D)
    k = rms_norm(k, sd[pfx + "self_attn.k_norm.weight"], cfg.rms_norm_eps).transpose(1, 2)
    v = v.transpose(1, 2)
    q, k = apply_rotary_pos_emb(q, k, cos, sin)
    if cache is not None:
        k, v = cache.update(k, v, layer)
    if ATTN_IMPL == "sdpa":  # transformers integrations/sdpa_attention.py: the reference's default dispatch
        out = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False,
                                             scale=D ** -0.5, enable_gqa=True)
        out = out.transpose(1, 2).reshape(bsz, q_len, -1)
        return _proj(sd, pfx, "o_proj", out)
    rep = Hq // Hkv
    kk = k.repeat_interleave(rep, dim=1)
    vv = v.repeat_interleave(rep, dim=1)
    scores = torch.matmul(q, kk.transpose(2, 3)) * (D ** -0.5)
    probs = torch.softmax(scores, dim=-1, dtype=torch.float32).to(q.dtype)
    out = torch.matmul(probs, vv).transpose(1, 2).reshape(bsz, q_len, -1)
    return _proj(sd, pfx, "o_proj", out)


def draft_forward(sd, cfg: DraftConfig, target_hidden, noise_embedding, position_ids,
                  cache: Optional[DraftCache] = None) -> torch.Tensor:
    """DFlashDraftModel.forward, model/dflash.py:166-190. Returns the final-normed hidden [B, q_len, H]."""
    hidden = noise_embedding
    target_hidden = rms_norm(_lin(target_hidden, sd["fc.weight"]), sd["hidden_norm.weight"], cfg.rms_norm_eps)
    cos, sin = rope_cos_sin(cfg.get_inv_freq(), position_ids, cfg.attention_scaling, hidden.dtype)
    for l in range(cfg.num_hidden_layers):
        pfx = f"layers.{l}."
        residual = hidden
        h = rms_norm(hidden, sd[pfx + "input_layernorm.weight"], cfg.rms_norm_eps)
        h = draft_attention(sd, pfx, cfg, h, target_hidden, cos, sin, cache, l)
        hidden = residual + h
        residual = hidden
        h = rms_norm(hidden, sd[pfx + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
        g = _lin(h, sd[pfx + "mlp.gate_proj.weight"])
        u = _lin(h, sd[pfx + "mlp.up_proj.weight"])
        h = _lin(F.silu(g) * u, sd[pfx + "mlp.down_proj.weight"])
        hidden = residual + h
    return rms_norm(hidden, sd["norm.weight"], cfg.rms_norm_eps)


# --------------------------------------------------------------------------------------------
# verify step: integer logic of model/dflash.py:258-268
# --------------------------------------------------------------------------------------------
def acceptance_length(block_ids: Sequence[int], posterior: Sequence[int]) -> int:
    """(block[1:] == posterior[:-1]).cumprod().sum()  — model/dflash.py:258."""
    n = 0
    for i in range(len(block_ids) - 1):
        if int(block_ids[i + 1]) != int(posterior[i]):
            break
        n += 1
    return n


def fixed_prefix_rank_candidates(base_block: torch.Tensor, draft_logits: torch.Tensor, fixed_prefix_len: int,
                                 rank_top_k: int, max_candidates: int):
    """benchmark_candidate_solutions.py:181-249 (`build_fixed_prefix_rank_candidates`), restated.
    base_block [1, eff] (slot 0 committed, 1.. greedy draft tokens), draft_logits [1, eff-1, V] (row j -> block
    position j+1). Returns (candidates [n, eff] int64, draft_scores list[float]): candidate 0 is the greedy block;
    candidate r keeps positions < max(1, min(fixed_prefix_len, eff)) and takes the rank-(r+1) token at every later
    position; the score of candidate r is the sum of its rank-(r+1) logits over those positions, summed in the logits'
    own dtype as torch does (`.sum(dim=1)`). With no suffix position, or fewer than 2 candidates: the base block."""
    eff = int(base_block.shape[1])
    suffix_start = max(1, min(int(fixed_prefix_len), eff))
    if suffix_start >= eff:
        return base_block.clone(), [0.0]
    n = min(int(max_candidates), int(rank_top_k), int(draft_logits.shape[-1]))
    if n <= 1:
        return base_block.clone(), [0.0]
    vals, idx = torch.topk(draft_logits[:, suffix_start - 1:, :], k=n, dim=-1)  # [1, suffix, n]
    cands = base_block.expand(n, -1).clone()
    cands[:, suffix_start:] = idx[0].transpose(0, 1)
    scores = vals[0].transpose(0, 1).sum(dim=1)
    return cands, [float(x) for x in scores.tolist()]


def choose_candidate(candidates: torch.Tensor, posterior_all: torch.Tensor, draft_scores: Sequence[float]):
    """benchmark_candidate_solutions.py:590-607: acceptance length per candidate, then the reference's fp32 composite
    `tau * 1e6 + draft_score - idx * 1e-3` and the first maximum. Returns (chosen index, acceptance lengths)."""
    acc = (candidates[:, 1:] == posterior_all[:, :-1]).cumprod(dim=1).sum(dim=1)
    tau = acc + 1
    sc = torch.tensor([float(x) for x in draft_scores], dtype=torch.float32)
    ids = torch.arange(candidates.shape[0], dtype=torch.float32)
    comp = tau.float().cpu() * 1e6 + sc - ids * 1e-3
    return int(torch.argmax(comp).item()), [int(x) for x in acc.tolist()]


def verify_commit(output_ids: List[int], start: int, block_ids: Sequence[int], posterior: Sequence[int]):
    """model/dflash.py:258-261. Mutates output_ids; returns (new_start, tau)."""
    a = acceptance_length(block_ids, posterior)
    output_ids[start:start + a + 1] = [int(t) for t in block_ids[:a + 1]]
    output_ids[start + a + 1] = int(posterior[a])
    return start + a + 1, a + 1


def finalize_output(output_ids: Sequence[int], num_input: int, max_length: int, mask_token_id: int,
                    stop_token_ids: Optional[Sequence[int]]) -> List[int]:
    """model/dflash.py:269-276: trim to max_length, drop mask ids, cut after the first stop token."""
    out = [int(t) for t in output_ids[:max_length] if int(t) != mask_token_id]
    if stop_token_ids is not None:
        stops = set(int(s) for s in stop_token_ids)
        for i in range(num_input, len(out)):
            if out[i] in stops:
                return out[: i + 1]
    return out


# --------------------------------------------------------------------------------------------
# the decode loop, model/dflash.py:192-277  (benchmark.py:43-251 with clamp_tail=True)
# --------------------------------------------------------------------------------------------
@torch.inference_mode()
def spec_generate(sd, cfg: DraftConfig, target, input_ids: torch.Tensor, max_new_tokens: int,
                  stop_token_ids: Optional[Sequence[int]], temperature: float, clamp_tail: bool = False,
                  forced_k: Optional[Sequence[int]] = None, trace: Optional[list] = None,
                  draft_fn: Optional[Callable] = None):
    """Restatement of DFlashDraftModel.spec_generate. `target` is the HF target model (not rewritten
    by this project, so the oracle calls it as is). `forced_k[cycle % len]` is the SURVEY §4 harness
    hook: posterior[:, :k] = block[:, 1:k+1] before the acceptance test (None = honest).
    Returns (output_ids[1, n], acceptance_lengths)."""
    from transformers import DynamicCache

    P = input_ids.shape[1]
    max_length = P + max_new_tokens
    bs = cfg.block_size
    dev = input_ids.device
    output_ids = torch.full((1, max_length + bs), cfg.mask_token_id, dtype=torch.long, device=dev)
    position_ids = torch.arange(output_ids.shape[1], device=dev).unsqueeze(0)
    cache_t = DynamicCache()
    cache_d = DraftCache()
    out = target(input_ids, position_ids=position_ids[:, :P], past_key_values=cache_t, use_cache=True,
                 logits_to_keep=1, output_hidden_states=True)
    output_ids[:, :P] = input_ids
    output_ids[:, P:P + 1] = sample(out.logits, temperature)
    target_hidden = extract_context_feature(out.hidden_states, cfg.target_layer_ids)
    if trace is not None:
        trace.append(dict(prefill=True, first_token=int(output_ids[0, P]),
                          hidden_sel=[out.hidden_states[i + 1][0].clone() for i in cfg.target_layer_ids]))
    acc_lengths = []
    start = P
    fwd = draft_fn or (lambda th, ne, pos, cache: draft_forward(sd, cfg, th, ne, pos, cache))
    while start < max_length:
        eff = min(bs, max_length - start) if clamp_tail else bs
        block = output_ids[:, start:start + eff].clone()
        block_pos = position_ids[:, start:start + eff]
        if eff > 1:
            noise = target.model.embed_tokens(block)
            hid = fwd(target_hidden, noise, position_ids[:, cache_d.get_seq_length(): start + eff], cache_d)
            draft_logits = target.lm_head(hid[:, -eff + 1:, :])
            cache_d.crop(start)
            block[:, 1:] = sample(draft_logits)
        else:
            draft_logits = None
        out = target(block, position_ids=block_pos, past_key_values=cache_t, use_cache=True,
                     output_hidden_states=True)
        posterior = sample(out.logits, temperature)
        if forced_k is not None:
            k = int(forced_k[len(acc_lengths) % len(forced_k)])
            k = min(k, eff - 1)
            posterior[:, :k] = block[:, 1:k + 1]
        a = acceptance_length(block[0].tolist(), posterior[0].tolist())
        output_ids[:, start:start + a + 1] = block[:, :a + 1]
        output_ids[:, start + a + 1] = posterior[:, a]
        if trace is not None:
            trace.append(dict(start=start, eff=eff, block=block[0].tolist(), posterior=posterior[0].tolist(),
                              tau=a + 1, ctx_feat=target_hidden[0].clone(),
                              draft_hidden=None if eff <= 1 else hid[0].clone(),
                              draft_logits=None if draft_logits is None else draft_logits[0].clone(),
                              target_logits=out.logits[0].clone(),
                              hidden_sel=[out.hidden_states[i + 1][0].clone() for i in cfg.target_layer_ids]))
        start += a + 1
        cache_t.crop(start)
        target_hidden = extract_context_feature(out.hidden_states, cfg.target_layer_ids)[:, :a + 1, :]
        acc_lengths.append(a + 1)
        if stop_token_ids is not None and any(int(s) in output_ids[0, P:].tolist() for s in stop_token_ids):
            break
    final = finalize_output(output_ids[0].tolist(), P, max_length, cfg.mask_token_id, stop_token_ids)
    return torch.tensor([final], dtype=torch.long, device=dev), acc_lengths


# --------------------------------------------------------------------------------------------
# the timed "step" of bench.py's CPU legs: everything of one cycle except target(...)
# --------------------------------------------------------------------------------------------
def draft_verify_step_cpu(sd, cfg: DraftConfig, embed_w, lm_head_w, block_ids, target_hidden, position_ids,
                          cache: DraftCache, start: int, target_logits, hidden_states_sel, temperature: float):
    """One pass of the hot path on the CPU (SURVEY §8d step boundary): embed -> draft forward ->
    lm_head -> argmax -> posterior sample -> acceptance -> next ctx features. Returns (block, tau, next_th)."""
    noise = F.embedding(block_ids, embed_w)
    hid = draft_forward(sd, cfg, target_hidden, noise, position_ids, cache)
    logits = _lin(hid[:, 1:, :], lm_head_w)
    cache.crop(start)
    block = block_ids.clone()
    block[:, 1:] = sample(logits)
    posterior = sample(target_logits, temperature)
    a = acceptance_length(block[0].tolist(), posterior[0].tolist())
    next_th = torch.cat(hidden_states_sel, dim=-1)[:, :a + 1, :]
    return block, a + 1, next_th
