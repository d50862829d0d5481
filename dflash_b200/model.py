"""Drop-in `DFlashDraftModel` for the reference's public API (model/dflash.py:147-277).

Same class contract as the reference: a `Qwen3PreTrainedModel` with `config_class = Qwen3Config`, the same
sub-module names (so checkpoints / state dicts load unchanged), the same public attributes
(`block_size`, `mask_token_id`, `target_layer_ids`) and the same `forward` / `spec_generate` signatures.
The modules below only hold parameters; the arithmetic of the draft-and-verify hot path runs in
hand-written sm_100a CUDA behind the C ABI (`include/dflash_b200.h`). The target model stays the caller's
HF module and is called exactly as the reference calls it.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn
from transformers import DynamicCache
from transformers.cache_utils import Cache
from transformers.models.qwen3.modeling_qwen3 import (Qwen3Config, Qwen3MLP, Qwen3PreTrainedModel, Qwen3RMSNorm,
                                                      Qwen3RotaryEmbedding)

from . import _lib
from .engine import DraftEngine
from .utils import ContextTap, build_target_layer_ids, sample


class Qwen3DFlashAttention(nn.Module):
    """Parameter container mirroring model/dflash.py:30-56 (q/k/v/o projections, per-head q/k RMSNorm)."""

    def __init__(self, config: Qwen3Config, layer_idx: int):
        super().__init__()
        self.config = config
        self.layer_idx = layer_idx
        self.head_dim = getattr(config, "head_dim", config.hidden_size // config.num_attention_heads)
        self.num_key_value_groups = config.num_attention_heads // config.num_key_value_heads
        self.scaling = self.head_dim ** -0.5
        self.is_causal = False
        bias = config.attention_bias
        self.q_proj = nn.Linear(config.hidden_size, config.num_attention_heads * self.head_dim, bias=bias)
        self.k_proj = nn.Linear(config.hidden_size, config.num_key_value_heads * self.head_dim, bias=bias)
        self.v_proj = nn.Linear(config.hidden_size, config.num_key_value_heads * self.head_dim, bias=bias)
        self.o_proj = nn.Linear(config.num_attention_heads * self.head_dim, config.hidden_size, bias=bias)
        self.q_norm = Qwen3RMSNorm(self.head_dim, eps=config.rms_norm_eps)
        self.k_norm = Qwen3RMSNorm(self.head_dim, eps=config.rms_norm_eps)


class Qwen3DFlashDecoderLayer(nn.Module):
    """Parameter container mirroring model/dflash.py:104-111."""

    def __init__(self, config: Qwen3Config, layer_idx: int):
        super().__init__()
        self.hidden_size = config.hidden_size
        self.self_attn = Qwen3DFlashAttention(config=config, layer_idx=layer_idx)
        self.mlp = Qwen3MLP(config)
        self.input_layernorm = Qwen3RMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.post_attention_layernorm = Qwen3RMSNorm(config.hidden_size, eps=config.rms_norm_eps)


class DFlashStaticCache(Cache):
    """A `transformers.Cache` that is only a handle on the engine's static draft KV cache: the three methods the
    decode loop uses (`get_seq_length`, `crop`, and the implicit append done by `forward`: model/dflash.py:215,241,246)
    work on a length; the K/V rows live in the engine workspace, so cropping is a length write."""

    def __init__(self):
        super().__init__(layers=[])
        self.length = 0

    def get_seq_length(self, layer_idx: int = 0) -> int:
        return self.length

    def crop(self, max_length: int):
        n = int(max_length)
        self.length = max(0, self.length + n) if n < 0 else min(self.length, n)

    def reset(self):
        self.length = 0

    def update(self, key_states, value_states, layer_idx, *args, **kwargs):
        raise RuntimeError("DFlashStaticCache holds no tensors: the draft K/V rows are written by the CUDA engine")


def _grow_foreign_cache(cache, n_rows: int, n_layers: int, device):
    """A caller-owned HF cache (e.g. the `DynamicCache()` of benchmark.py:60,122-129) only has to report the right
    length to its owner (`get_seq_length()` / `crop()`); the K/V rows themselves stay in the engine. Append `n_rows`
    empty positions per layer: [1, 1, n_rows, 1] placeholders, a few bytes each."""
    z = torch.zeros(1, 1, int(n_rows), 1, dtype=torch.bfloat16, device=device)
    for layer_idx in range(n_layers):
        cache.update(z, z, layer_idx)


class DFlashDraftModel(Qwen3PreTrainedModel):
    config_class = Qwen3Config
    _no_split_modules = ["Qwen3DFlashDecoderLayer"]

    def __init__(self, config) -> None:
        super().__init__(config)
        self.config = config
        self.layers = nn.ModuleList(
            [Qwen3DFlashDecoderLayer(config, layer_idx) for layer_idx in range(config.num_hidden_layers)])
        self.target_layer_ids = self.config.dflash_config.get(
            "target_layer_ids", build_target_layer_ids(config.num_target_layers, config.num_hidden_layers))
        self.norm = Qwen3RMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.rotary_emb = Qwen3RotaryEmbedding(config)
        self.fc = nn.Linear(len(self.target_layer_ids) * config.hidden_size, config.hidden_size, bias=False)
        self.hidden_norm = Qwen3RMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.block_size = config.block_size
        self.mask_token_id = self.config.dflash_config.get("mask_token_id", None)
        self._engine: Optional[DraftEngine] = None
        self._engine_key = None
        self.last_acceptance_lengths: List[int] = []
        self.post_init()

    # ------------------------------------------------------------------------------------------
    def _get_engine(self, embed_w: torch.Tensor, lm_head_w: torch.Tensor, max_seq: int, out_len: int,
                    keep_draft_logits: bool = False, max_candidates: int = 0) -> DraftEngine:
        key = (embed_w.data_ptr(), lm_head_w.data_ptr(), self.block_size, keep_draft_logits,
               self.fc.weight.data_ptr(), max_candidates)
        e = self._engine
        if e is not None and self._engine_key == key and e.max_seq >= max_seq and e.out_len >= out_len:
            return e
        if e is not None:
            e.close()
        # round capacities up so that successive calls with similar lengths reuse the engine
        cap_seq = max(1024, 1 << (int(max_seq) - 1).bit_length())
        cap_out = max(1024, 1 << (int(out_len) - 1).bit_length())
        self._engine = DraftEngine(self, embed_w, lm_head_w, max_seq=cap_seq, out_len=cap_out, max_requests=1,
                                   block_size=self.block_size, keep_draft_logits=keep_draft_logits,
                                   max_candidates=max_candidates, device=embed_w.device)
        self._engine_key = key
        return self._engine

    def release_engine(self):
        """Drop the cached engines (single-stream and batched), their CUDA graphs and the graphed targets."""
        if self._engine is not None:
            self._engine.close()
            self._engine = None
        cached = getattr(self, "_batch_engine_cache", None)
        if cached is not None:
            cached[1].close()
            self._batch_engine_cache = None
        self._graphed_target = None

    # ------------------------------------------------------------------------------------------
    def forward(self, position_ids: torch.LongTensor, attention_mask: Optional[torch.Tensor] = None,
                noise_embedding: Optional[torch.Tensor] = None, target_hidden: Optional[torch.Tensor] = None,
                past_key_values=None, use_cache: bool = False, **kwargs) -> torch.Tensor:
        """model/dflash.py:166-190. Returns the final-normed hidden states [1, q_len, H] (no lm_head).
        `past_key_values`: None (a fresh context starting at position 0), a `DFlashStaticCache`, or any HF `Cache` the
        caller owns (`DynamicCache()` as in benchmark.py:60,122-129): its length is adopted and kept in step, the K/V
        rows themselves live in the engine (one sequence at a time: a cache that was not grown by this model must be
        empty). Side effect, as in the reference: the cache grows by ctx_len + q_len rows (the caller crops)."""
        if attention_mask is not None:
            raise NotImplementedError("the DFlash draft attends without a mask (model/dflash.py:244)")
        if noise_embedding.shape[0] != 1:
            raise RuntimeError("DFlashDraftModel.forward: batch size 1 only (as spec_generate, model/dflash.py:206-211)")
        foreign = past_key_values is not None and not isinstance(past_key_values, DFlashStaticCache)
        if foreign and past_key_values.get_seq_length() > 0 and getattr(self, "_foreign_cache_id", None) != id(past_key_values):
            raise RuntimeError("DFlashDraftModel.forward: this cache holds K/V rows written by another model; the CUDA "
                               "engine keeps the draft K/V of one sequence (start from an empty cache)")
        dev = noise_embedding.device
        q_len, c = noise_embedding.shape[1], target_hidden.shape[1]
        cache_len = 0 if past_key_values is None else past_key_values.get_seq_length()
        pos0 = int(position_ids[0, 0])
        if pos0 != cache_len or position_ids.shape[1] != c + q_len:
            raise RuntimeError(f"position_ids must cover [cache_len, cache_len + ctx + q_len): got first={pos0}, "
                               f"n={position_ids.shape[1]}, cache_len={cache_len}, ctx={c}, q_len={q_len}")
        start = cache_len + c
        bf = torch.bfloat16
        dummy = getattr(self, "_fwd_dummy", None)
        if dummy is None or dummy.device != dev:
            dummy = torch.zeros(128, self.config.hidden_size, dtype=bf, device=dev)
            self._fwd_dummy = dummy
        e = self._engine
        if e is None or e.max_seq < start + 2 * q_len + 32:
            e = self._get_engine(dummy, dummy, max_seq=start + 2 * q_len + 32, out_len=1024)
        if q_len > e.SL:
            raise RuntimeError(f"q_len {q_len} exceeds the engine's block rows {e.SL}")
        H, nsel = self.config.hidden_size, len(self.target_layer_ids)
        th = target_hidden[0].to(bf)
        full = (c // e.SL) * e.SL if c > e.SL else 0
        if full:  # whole SL-row chunks go through the context-only pass
            parts = [th[:full, s * H:(s + 1) * H].contiguous() for s in range(nsel)]
            e.prefill_context(0, parts, pos0=cache_len)  # rows land at [cache_len, cache_len + full)
        rem = c - full
        e.buf["ctx_feat"].view(e.R * e.SL, nsel * H)[:rem] = th[full:]
        e.buf["start"][0] = start
        e.buf["ctx_len"][0] = rem
        e.buf["blk_len"][0] = q_len
        noise = torch.zeros(e.SL, H, dtype=bf, device=dev)
        noise[:q_len] = noise_embedding[0].to(bf)
        e.draft_step(noise_embedding=noise, lm_head=False)
        if foreign:
            _grow_foreign_cache(past_key_values, c + q_len, self.config.num_hidden_layers, dev)
            self._foreign_cache_id = id(past_key_values)
        elif past_key_values is not None:
            past_key_values.length = start + q_len
        return e.hn[:q_len].clone().unsqueeze(0).to(noise_embedding.dtype)

    # ------------------------------------------------------------------------------------------
    def spec_generate_batch(self, target: nn.Module, prompts, max_new_tokens: int, stop_token_ids: Optional[List[int]],
                            temperature: float, **kwargs):
        """Many prompts through one engine (up to 64 request streams share the draft's weight stream); see
        `dflash_b200/batched.py`. Returns a list of `LongTensor[1, P_i + n_i]` in prompt order."""
        from .batched import spec_generate_batch
        return spec_generate_batch(self, target, prompts, max_new_tokens, stop_token_ids, temperature, **kwargs)

    @torch.inference_mode()
    def spec_generate(self, target: nn.Module, input_ids: torch.LongTensor, max_new_tokens: int,
                      stop_token_ids: Optional[List[int]], temperature: float, *, clamp_tail: bool = False,
                      forced_k: Optional[List[int]] = None, seed: Optional[int] = None,
                      graph_target: Optional[bool] = None, sync_every: int = 1):
        """model/dflash.py:192-277, same signature and return value (`LongTensor[1, P + n_out]`).
        Keyword-only extras: `clamp_tail` = benchmark.py:104-105 behaviour, `forced_k` = harness hook that
        forces the first k draft tokens of each cycle to be accepted, `seed` for T > 0,
        `graph_target` (default: `self.graph_target`, "auto") runs the target's verify forward from a CUDA graph
        over a static KV cache (`dflash_b200/target_graph.py`) — the target module itself is
        untouched, and the whole cycle then needs no host round trip, so the host may poll the device state only
        every `sync_every` cycles. "auto" falls back to the reference's eager call when the target cannot be
        captured; False forces the eager call."""
        self.eval()
        if input_ids.shape[0] != 1:
            raise RuntimeError("spec_generate: batch size 1 only (the expanded size of the tensor must match: "
                               "model/dflash.py:206-211,227)")
        dev = target.device
        P = input_ids.shape[1]
        max_length = P + max_new_tokens
        bs = self.block_size
        embed_w = target.model.embed_tokens.weight
        lm_head_w = target.lm_head.weight
        e = self._get_engine(embed_w, lm_head_w, max_seq=max_length + 2 * bs + 1, out_len=max_length + bs + 1)
        position_ids = torch.arange(max_length + bs, device=dev).unsqueeze(0)
        if seed is None:
            seed = int(torch.randint(0, 2**62, (1,)).item()) if temperature >= 1e-5 else 0
        # Target verify forward: replayed from a CUDA graph over a static KV cache whenever the target can be captured
        #  -- same module, same arithmetic, no host round trip in the cycle. `graph_target=False`
        # (or `model.graph_target = False`) keeps the reference's eager call with a DynamicCache; the default "auto"
        # falls back to it when capture fails (e.g. attn_implementation="eager", data-dependent expert routing).
        if graph_target is None:
            graph_target = getattr(self, "graph_target", "auto")
        auto = graph_target == "auto"
        if auto and getattr(self, "_graph_target_failed", None) == id(target):
            graph_target = False
        gt = None
        logits0 = hidden0 = None
        if graph_target:
            from .target_graph import GraphedVerifyTarget
            key = (id(target), bs, e.buf["start"].data_ptr())
            try:
                gt = getattr(self, "_graphed_target", None)
                # (+ one more block: a blind cycle behind the end of the request -- sync_every > 1 -- still runs the target
                # over the frozen block at positions [start, start + bs), start < max_length + bs)
                if gt is None or getattr(self, "_graphed_target_key", None) != key or gt.max_cache_len < max_length + 2 * bs:
                    cap = max(1024, 1 << (max_length + 2 * bs - 1).bit_length())
                    gt = GraphedVerifyTarget(target, bs, cap, self.target_layer_ids, e.buf["start"], e.block_ids)
                    self._graphed_target, self._graphed_target_key = gt, key
                logits0, hidden0 = gt.prefill(input_ids)
                if gt.graph is None:
                    e.buf["start"][0] = P  # capture replays the forward on the live state: any in-range start will do
                    gt.capture()
            except Exception as ex:  # noqa: BLE001 -- capture failures surface as many exception types
                if not auto:
                    raise
                import warnings
                warnings.warn(f"dflash_b200: the target's verify forward could not be captured in a CUDA graph "
                              f"({type(ex).__name__}: {ex}); calling it eagerly as the reference does")
                self._graph_target_failed = id(target)
                self._graphed_target = None
                gt = None
                torch.cuda.synchronize(dev)
        if gt is None:
            # the selected residual streams come from forward hooks (ContextTap), not output_hidden_states=True:
            # same tensors, without the target keeping all L + 1 of them
            cache_t = DynamicCache()
            tap = ContextTap(target, self.target_layer_ids)
            with tap:
                out = target(input_ids, position_ids=position_ids[:, :P], past_key_values=cache_t, use_cache=True,
                             logits_to_keep=1)
            logits0, hidden0 = out.logits, list(tap.states)
        first = sample(logits0, temperature, seed=seed ^ 0x5DEECE66D)
        e.reset_request(0, input_ids[0], first.view(-1)[0], max_new_tokens)
        if clamp_tail:
            e.buf["blk_len"][0] = min(bs, max_new_tokens)
        e.prefill_context(0, [h[0] for h in hidden0])
        stop_t = None
        if stop_token_ids is not None and len(stop_token_ids) > 0:
            stop_t = torch.tensor(list(stop_token_ids), dtype=torch.int64, device=dev)
        forced_t = None
        if forced_k is not None:
            forced_t = torch.tensor([list(forced_k)], dtype=torch.int32, device=dev)

        start = P
        state = torch.empty(2, dtype=torch.int32, device="cpu").pin_memory()
        cyc = 0
        while gt is not None and start < max_length:
            # fully device-driven cycle: draft graph -> target graph -> verify kernels, state stays on the GPU
            e.draft_step_graphed()
            logits, hidden = gt.verify_forward()
            if logits.dtype != torch.bfloat16:
                logits = logits.to(torch.bfloat16)
            e.verify_step(logits, hidden, temperature=temperature, seed=seed, stop_ids=stop_t, forced_k=forced_t,
                          clamp_tail=clamp_tail, inject=True)
            cyc += 1
            # (polling one cycle late, so that the GPU never waits for the host, was measured: the blind cycle at the
            # end of the request costs more than the per-cycle sync it saves -- 11.49 vs 11.21 ms per cycle)
            if cyc % max(1, sync_every) == 0:  # blind cycles past the end are frozen on the device (`done`)
                state[0:1].copy_(e.buf["start"][0:1], non_blocking=True)
                state[1:2].copy_(e.buf["done"][0:1], non_blocking=True)
                torch.cuda.current_stream(dev).synchronize()
                start = int(state[0])
                cyc = 0
                if int(state[1]):
                    break  # stop token committed, or max_length reached (the device froze the request)
        while gt is None and start < max_length:
            eff = min(bs, max_length - start) if clamp_tail else bs
            e.draft_step_graphed() if eff > 1 else None
            block = e.block_ids[:, :eff]
            with tap:
                out = target(block, position_ids=position_ids[:, start:start + eff], past_key_values=cache_t,
                             use_cache=True)
            logits = out.logits[0]
            if logits.dtype != torch.bfloat16:
                logits = logits.to(torch.bfloat16)
            hidden = [h[0].contiguous() for h in tap.states]
            if eff < bs:  # tail-clamped block: pad the rows the kernels index to the block stride
                logits = torch.nn.functional.pad(logits, (0, 0, 0, bs - eff))
                hidden = [torch.nn.functional.pad(h, (0, 0, 0, bs - eff)) for h in hidden]
            e.verify_step(logits, hidden, temperature=temperature, seed=seed, stop_ids=stop_t, forced_k=forced_t,
                          clamp_tail=clamp_tail, inject=True)
            # the one host sync of the cycle: the HF target needs `start` to slice positions / crop its cache
            state[0:1].copy_(e.buf["start"][0:1], non_blocking=True)
            state[1:2].copy_(e.buf["done"][0:1], non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            start = int(state[0])
            cache_t.crop(start)
            if int(state[1]) and start < max_length:
                break  # stop token committed
        n_cyc = int(e.buf["n_cycles"][0])
        self.last_acceptance_lengths = e.acc_hist[0, :n_cyc].tolist()
        output_ids = e.output_ids[0:1, :max_length].clone()
        output_ids = output_ids[:, output_ids[0] != self.mask_token_id]
        if stop_t is not None:
            idx = torch.isin(output_ids[0][P:], stop_t).nonzero(as_tuple=True)[0]
            if idx.numel() > 0:
                output_ids = output_ids[:, : P + idx[0] + 1]
        return output_ids
