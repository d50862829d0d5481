"""Host side of the draft/verify engine: packs the draft's weights once, owns the device workspace
(one torch allocation carved by the C library), and enqueues the step kernels through the C ABI.

PyTorch is plumbing here (device memory, streams, CUDA-graph capture); all arithmetic of the hot path
runs in libdflash_b200.so. There is no fallback: a missing library or a non-sm_100 device raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_float, c_int, c_longlong, c_size_t, c_ulonglong, c_void_p)
from typing import List, Optional, Sequence

import torch

from . import _lib


class CConfig(Structure):
    _fields_ = [
        ("hidden", c_int), ("intermediate", c_int), ("n_layers", c_int), ("n_q_heads", c_int),
        ("n_kv_heads", c_int), ("head_dim", c_int), ("vocab", c_int), ("n_sel", c_int), ("block_size", c_int),
        ("max_requests", c_int), ("max_seq", c_int), ("out_len", c_int), ("hist_len", c_int),
        ("rms_eps", c_float), ("rope_scale", c_float), ("mask_token_id", c_longlong), ("attn_splits", c_int),
        ("post_splits", c_int), ("gemm_grid", c_int), ("use_pdl", c_int), ("keep_draft_logits", c_int),
        ("max_candidates", c_int),
    ]


class CLayerWeights(Structure):
    _fields_ = [(n, c_void_p) for n in ("wqkv", "wo", "wgu", "wd", "ln1", "ln2", "q_norm", "k_norm", "bqkv", "bo")]


class CWeights(Structure):
    _fields_ = [("embed", c_void_p), ("lm_head", c_void_p), ("fc", c_void_p), ("hidden_norm", c_void_p),
                ("final_norm", c_void_p), ("inv_freq", c_void_p), ("layers_host", POINTER(CLayerWeights))]


# enum dflash_buffer_id (include/dflash_b200.h) -> (name, torch dtype)
BUFFERS = [
    ("x", torch.bfloat16), ("a_in", torch.bfloat16), ("ctx_feat", torch.bfloat16), ("q", torch.bfloat16),
    ("attn_out", torch.bfloat16), ("a2", torch.bfloat16), ("hmid", torch.bfloat16), ("hn", torch.bfloat16),
    ("kv", torch.bfloat16), ("ws", torch.float32), ("part", torch.float32), ("flags", torch.int32),
    ("counters", torch.int32), ("attn_po", torch.float32), ("attn_ml", torch.float32),
    ("cand_val", torch.float32), ("cand_idx", torch.int32), ("post_val", torch.float32), ("post_idx", torch.int32),
    ("draft_tokens", torch.int64), ("block_ids", torch.int64), ("posterior", torch.int64),
    ("output_ids", torch.int64), ("start", torch.int32), ("ctx_len", torch.int32), ("done", torch.int32),
    ("n_cycles", torch.int32), ("blk_len", torch.int32), ("max_len", torch.int32), ("acc_hist", torch.int32),
    ("rng_step", torch.int64), ("draft_logits", torch.bfloat16), ("pf_feat", torch.bfloat16), ("pf_a", torch.bfloat16),
    ("topk_idx", torch.int32), ("topk_val", torch.float32), ("cand_ids", torch.int64),
    ("cand_scores", torch.float32), ("chosen", torch.int32),
]


def _declare(lib):
    if getattr(lib, "_dflash_engine_declared", False):
        return
    lib.dflash_workspace_bytes.restype = c_size_t
    lib.dflash_workspace_bytes.argtypes = [POINTER(CConfig)]
    lib.dflash_engine_create.restype = c_int
    lib.dflash_engine_create.argtypes = [POINTER(CConfig), POINTER(CWeights), c_void_p, c_size_t, POINTER(c_void_p)]
    lib.dflash_engine_destroy.restype = None
    lib.dflash_engine_destroy.argtypes = [c_void_p]
    lib.dflash_engine_buffer.restype = c_int
    lib.dflash_engine_buffer.argtypes = [c_void_p, c_int, POINTER(c_void_p), POINTER(c_size_t)]
    lib.dflash_prefill_context.restype = c_int
    lib.dflash_prefill_context.argtypes = [c_void_p, c_int, POINTER(c_void_p), c_int, c_void_p]
    lib.dflash_prefill_context_at.restype = c_int
    lib.dflash_prefill_context_at.argtypes = [c_void_p, c_int, POINTER(c_void_p), c_int, c_int, c_void_p]
    lib.dflash_draft_step.restype = c_int
    lib.dflash_draft_step.argtypes = [c_void_p, c_void_p, c_int, c_void_p]
    lib.dflash_verify_step.restype = c_int
    lib.dflash_verify_step.argtypes = [c_void_p, c_void_p, c_longlong, c_void_p, POINTER(c_void_p), c_float,
                                       c_void_p, c_ulonglong, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]
    lib.dflash_verify_inject_step.restype = c_int
    lib.dflash_verify_inject_step.argtypes = lib.dflash_verify_step.argtypes
    lib.dflash_draft_step_injected.restype = c_int
    lib.dflash_draft_step_injected.argtypes = [c_void_p, c_int, c_float, c_ulonglong, c_void_p]
    lib.dflash_embed_block.restype = c_int
    lib.dflash_embed_block.argtypes = [c_void_p, c_void_p]
    lib.dflash_engine_launches.restype = c_int
    lib.dflash_engine_launches.argtypes = [c_void_p, c_int]
    lib.dflash_draft_step_sampled.restype = c_int
    lib.dflash_draft_step_sampled.argtypes = [c_void_p, c_float, c_ulonglong, c_void_p]
    lib.dflash_gemm_sample.restype = c_int
    lib.dflash_gemm_sample.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_float,
                                       c_ulonglong, c_ulonglong, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]
    lib.dflash_draft_step_candidates.restype = c_int
    lib.dflash_draft_step_candidates.argtypes = [c_void_p, c_int, c_int, c_void_p]
    lib.dflash_verify_step_candidates.restype = c_int
    lib.dflash_verify_step_candidates.argtypes = [c_void_p, c_int, c_void_p, c_longlong, POINTER(c_void_p), c_float,
                                                  c_void_p, c_ulonglong, c_void_p, c_int, c_int, c_void_p]
    lib.dflash_sample.restype = c_int
    lib.dflash_sample.argtypes = [c_void_p, c_longlong, c_int, c_int, c_float, c_void_p, c_ulonglong, c_void_p,
                                  c_void_p, c_int, c_void_p, c_void_p]
    lib._dflash_engine_declared = True


def _p(t: Optional[torch.Tensor]):
    return None if t is None else c_void_p(t.data_ptr())


def _stream(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


class PackedDraftWeights:
    """bf16 device copies of the draft's matrices in the fused layouts the kernels stream
    (q/k/v stacked, gate/up stacked). When `alias` is set the module's own parameters are re-pointed at
    views of the packed buffers, so the state-dict keys survive and nothing is stored twice."""

    def __init__(self, draft, device, alias: bool = True):
        bf = torch.bfloat16
        self.layers = []
        for layer in draft.layers:
            at, mlp = layer.self_attn, layer.mlp
            wqkv = torch.cat([at.q_proj.weight, at.k_proj.weight, at.v_proj.weight], 0).to(device=device, dtype=bf).contiguous()
            wgu = torch.cat([mlp.gate_proj.weight, mlp.up_proj.weight], 0).to(device=device, dtype=bf).contiguous()
            ent = dict(
                wqkv=wqkv, wgu=wgu,
                wo=at.o_proj.weight.detach().to(device=device, dtype=bf).contiguous(),
                wd=mlp.down_proj.weight.detach().to(device=device, dtype=bf).contiguous(),
                ln1=layer.input_layernorm.weight.detach().to(device=device, dtype=bf).contiguous(),
                ln2=layer.post_attention_layernorm.weight.detach().to(device=device, dtype=bf).contiguous(),
                q_norm=at.q_norm.weight.detach().to(device=device, dtype=bf).contiguous(),
                k_norm=at.k_norm.weight.detach().to(device=device, dtype=bf).contiguous(),
                bqkv=None, bo=None,
            )
            if getattr(at.q_proj, "bias", None) is not None:  # config.attention_bias (model/dflash.py:41-50)
                ent["bqkv"] = torch.cat([at.q_proj.bias, at.k_proj.bias, at.v_proj.bias], 0).detach().to(
                    device=device, dtype=bf).contiguous()
                ent["bo"] = at.o_proj.bias.detach().to(device=device, dtype=bf).contiguous()
            if alias and at.q_proj.weight.dtype == bf and at.q_proj.weight.device == wqkv.device:
                nq, nk = at.q_proj.weight.shape[0], at.k_proj.weight.shape[0]
                at.q_proj.weight.data = wqkv[:nq]
                at.k_proj.weight.data = wqkv[nq:nq + nk]
                at.v_proj.weight.data = wqkv[nq + nk:]
                ni = mlp.gate_proj.weight.shape[0]
                mlp.gate_proj.weight.data = wgu[:ni]
                mlp.up_proj.weight.data = wgu[ni:]
            self.layers.append(ent)
        self.fc = draft.fc.weight.detach().to(device=device, dtype=bf).contiguous()
        self.hidden_norm = draft.hidden_norm.weight.detach().to(device=device, dtype=bf).contiguous()
        self.final_norm = draft.norm.weight.detach().to(device=device, dtype=bf).contiguous()
        self.inv_freq = draft.rotary_emb.inv_freq.detach().to(device=device, dtype=torch.float32).contiguous()
        self.rope_scale = float(getattr(draft.rotary_emb, "attention_scaling", 1.0))
        self.device = torch.device(device)
        self._fc_src = draft.fc.weight.data_ptr()

    def still_matches(self, draft, device) -> bool:
        """True while the module's parameters are the views this packing created (or, for un-aliased packings, the
        same storage it was copied from): several engines of one draft then share ONE packed copy."""
        if torch.device(device) != self.device or draft.fc.weight.data_ptr() != self._fc_src:
            return False
        if len(draft.layers) != len(self.layers):
            return False
        at = draft.layers[0].self_attn
        return at.q_proj.weight.data_ptr() == self.layers[0]["wqkv"].data_ptr()

    @classmethod
    def for_draft(cls, draft, device):
        cached = getattr(draft, "_dflash_packed", None)
        if cached is not None and cached.still_matches(draft, device):
            return cached
        packed = cls(draft, device)
        at = draft.layers[0].self_attn
        if at.q_proj.weight.data_ptr() == packed.layers[0]["wqkv"].data_ptr():  # aliased: safe to share
            draft._dflash_packed = packed
        return packed


class DraftEngine:
    """One GPU's draft+verify engine for `max_requests` request streams (1..64)."""

    def __init__(self, draft, embed_weight: torch.Tensor, lm_head_weight: torch.Tensor, *, max_seq: int,
                 out_len: int, max_requests: int = 1, block_size: Optional[int] = None, use_pdl: bool = True,
                 keep_draft_logits: bool = False, gemm_grid: int = 0, attn_splits: int = 0,
                 hist_len: Optional[int] = None, max_candidates: int = 0, device=None):
        self.lib = _lib.load()
        _declare(self.lib)
        if not torch.cuda.is_available():
            raise _lib.DFlashNativeError("dflash_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device(device if device is not None else embed_weight.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dflash_device_check(), "dflash_device_check")
        cfg = draft.config
        self.block_size = int(block_size or draft.block_size)
        self.R = int(max_requests)
        self.n_sel = len(draft.target_layer_ids)
        self.hidden = cfg.hidden_size
        self.vocab = int(lm_head_weight.shape[0])
        self.mask_token_id = int(draft.mask_token_id)
        bf = torch.bfloat16
        if embed_weight.dtype != bf or lm_head_weight.dtype != bf:
            raise _lib.DFlashNativeError("the CUDA path computes in bf16: load the target with dtype=torch.bfloat16")
        self.embed = embed_weight.detach().contiguous()
        self.lm_head = lm_head_weight.detach().contiguous()
        self.weights = PackedDraftWeights.for_draft(draft, self.device)
        head_dim = getattr(cfg, "head_dim", cfg.hidden_size // cfg.num_attention_heads)
        if any(t != "full_attention" for t in (getattr(cfg, "layer_types", None) or [])):
            raise _lib.DFlashNativeError("sliding-window draft layers are not supported by the CUDA path")
        # one acceptance-length entry per cycle; a cycle commits at least one token, so out_len bounds the cycles
        hist_len = int(out_len) if hist_len is None else int(hist_len)
        self.hist_len = hist_len
        self.ccfg = CConfig(
            hidden=cfg.hidden_size, intermediate=cfg.intermediate_size, n_layers=cfg.num_hidden_layers,
            n_q_heads=cfg.num_attention_heads, n_kv_heads=cfg.num_key_value_heads, head_dim=head_dim,
            vocab=self.vocab, n_sel=self.n_sel, block_size=self.block_size, max_requests=self.R,
            max_seq=int(max_seq), out_len=int(out_len), hist_len=int(hist_len), rms_eps=float(cfg.rms_norm_eps),
            rope_scale=self.weights.rope_scale, mask_token_id=self.mask_token_id,
            attn_splits=int(os.environ.get("DFLASH_ATTN_SPLITS", attn_splits)),
            post_splits=int(os.environ.get("DFLASH_POST_SPLITS", 0)), gemm_grid=int(os.environ.get("DFLASH_GEMM_GRID", gemm_grid)), use_pdl=int(use_pdl),
            keep_draft_logits=int(keep_draft_logits), max_candidates=int(max_candidates))
        self.max_candidates = int(max_candidates)
        self.max_seq, self.out_len = int(max_seq), int(out_len)
        with torch.cuda.device(self.device):
            nbytes = self.lib.dflash_workspace_bytes(byref(self.ccfg))
            if nbytes == 0:
                _lib.check(-1, "dflash_workspace_bytes")
            self.workspace = torch.zeros(nbytes + 1024, dtype=torch.uint8, device=self.device)
            off = (-self.workspace.data_ptr()) % 1024
            self._ws_view = self.workspace[off:off + nbytes]
            arr = (CLayerWeights * len(self.weights.layers))()
            for i, ent in enumerate(self.weights.layers):
                for k in ("wqkv", "wo", "wgu", "wd", "ln1", "ln2", "q_norm", "k_norm", "bqkv", "bo"):
                    setattr(arr[i], k, None if ent[k] is None else ent[k].data_ptr())
            self._layer_arr = arr
            cw = CWeights(embed=self.embed.data_ptr(), lm_head=self.lm_head.data_ptr(), fc=self.weights.fc.data_ptr(),
                          hidden_norm=self.weights.hidden_norm.data_ptr(), final_norm=self.weights.final_norm.data_ptr(),
                          inv_freq=self.weights.inv_freq.data_ptr(), layers_host=arr)
            h = c_void_p()
            _lib.check(self.lib.dflash_engine_create(byref(self.ccfg), byref(cw), c_void_p(self._ws_view.data_ptr()),
                                                     nbytes, byref(h)), "dflash_engine_create")
        self.handle = h
        self.buf = {}
        base = self._ws_view.data_ptr()
        for i, (name, dt) in enumerate(BUFFERS):
            p, n = c_void_p(), c_size_t()
            _lib.check(self.lib.dflash_engine_buffer(self.handle, i, byref(p), byref(n)), "dflash_engine_buffer")
            o = (p.value or base) - base
            self.buf[name] = self._ws_view[o:o + n.value].view(dt)
        R, bs = self.R, self.block_size
        self.block_ids = self.buf["block_ids"].view(R, bs)
        self.posterior = self.buf["posterior"].view(R, bs)
        self.output_ids = self.buf["output_ids"].view(R, self.out_len)
        self.acc_hist = self.buf["acc_hist"].view(R, hist_len)
        self.cand_ids = self.buf["cand_ids"].view(R, 4, bs)        # candidate blocks (multi-candidate drafting)
        self.cand_scores = self.buf["cand_scores"].view(R, 4)
        self.SL = 16 if bs <= 16 else 32
        self.hn = self.buf["hn"].view(-1, self.hidden)[:R * self.SL]  # (the buffer is padded to the UMMA width)
        # launches of the schedule actually enqueued (engine.cuh), as the library counts them. Per layer {qkv GEMM,
        # qkv_post, attention, merge (only with several KV splits), o GEMM, row kernel, gate/up GEMM (SwiGLU epilogue),
        # down GEMM, row kernel}, lm_head GEMM (argmax + drafted tokens); the context-injection kernel (concat + fc +
        # hidden_norm + block embedding + first layernorm) is the first kernel of draft_step() or, with
        # verify_step(inject=True), the second kernel of the verify step (where it overlaps the verify kernel)
        n = [self.lib.dflash_engine_launches(self.handle, w) for w in range(4)]
        self.kernels_per_draft_step, self.kernels_per_draft_step_injected = n[0], n[1]
        self.kernels_per_verify_step, self.kernels_per_verify_inject_step = n[2], n[3]
        self._graph = None
        self._graph_injected = None
        # where the next draft step finds its activation rows: "no" = nothing injected (draft_step runs the injection
        # kernel itself, from the gathered features); "fresh" = injected by verify_step(inject=True);
        # "stale" = a draft step has already consumed them (the block rows are re-embedded before another one)
        self._injected = "no"

    # ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.dflash_engine_destroy(self.handle)
            self.handle = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def reset_request(self, r: int, prompt: torch.Tensor, first_token: int, max_new_tokens: int):
        """Host-side (not hot path) initialisation of request r, mirroring dflash.py:206-229."""
        P = int(prompt.numel())
        if P + max_new_tokens + self.block_size > self.out_len:
            raise ValueError("out_len too small for prompt + max_new_tokens + block_size")
        if P + max_new_tokens + 2 * self.block_size > self.max_seq:
            raise ValueError("max_seq too small for prompt + max_new_tokens + 2*block_size")
        out = self.output_ids[r]
        out.fill_(self.mask_token_id)
        out[:P] = prompt.to(out.device).view(-1)
        out[P] = first_token
        blk = self.block_ids[r]
        blk.fill_(self.mask_token_id)
        blk[0] = first_token
        self.buf["start"][r] = P
        self.buf["ctx_len"][r] = 0
        self.buf["done"][r] = 0
        self.buf["n_cycles"][r] = 0
        self.buf["blk_len"][r] = self.block_size
        self.buf["max_len"][r] = P + max_new_tokens
        if P + max_new_tokens > self.hist_len:
            raise ValueError("hist_len too small: one acceptance-length entry per cycle is kept")
        if self.R == 1:  # a fresh request restarts the Philox step: same seed -> same draws (engines are cached)
            self.buf["rng_step"].zero_()
        if self._injected == "fresh":
            self._injected = "stale"  # the block rows of this request are embedded again before the next draft step

    def _call(self, name: str, *args):
        """Every library call runs with the engine's device current (the C side launches on the current device)."""
        with torch.cuda.device(self.device):
            _lib.check(getattr(self.lib, name)(*args), name)

    def prefill_context(self, r: int, hidden: Sequence[torch.Tensor], pos0: int = 0):
        """hidden[s]: [P, H] bf16 rows of the selected target layers; they become the context at cache positions
        [pos0, pos0 + P) of request r (pos0 = 0: the prompt)."""
        assert len(hidden) == self.n_sel
        hs = [h.contiguous() for h in hidden]
        P = hs[0].shape[0]
        arr = (c_void_p * self.n_sel)(*[h.data_ptr() for h in hs])
        self._call("dflash_prefill_context_at", self.handle, r, arr, P, int(pos0), _stream(self.device))
        self._keep = hs

    def embed_block(self):
        """embed_tokens(block_ids) -> residual stream + layer 0's input_layernorm (model/dflash.py:237) for every stream."""
        self._call("dflash_embed_block", self.handle, _stream(self.device))

    def draft_step(self, noise_embedding: Optional[torch.Tensor] = None, lm_head: bool = True):
        if noise_embedding is None and self._injected != "no":
            return self.draft_step_injected(lm_head=lm_head)
        self._call("dflash_draft_step", self.handle, _p(noise_embedding), int(lm_head), _stream(self.device))
        self._injected = "no"

    def draft_step_injected(self, lm_head: bool = True, temperature: float = 0.0, seed: int = 0):
        """The draft step without its injection kernel (the previous verify_step(inject=True) ran it)."""
        if self._injected == "stale":
            self.embed_block()
        self._call("dflash_draft_step_injected", self.handle, int(lm_head), float(temperature), int(seed) & (2**64 - 1),
                   _stream(self.device))
        self._injected = "stale"

    def verify_step(self, target_logits: Optional[torch.Tensor], hidden: Sequence[torch.Tensor], *,
                    temperature: float = 0.0, posterior_in: Optional[torch.Tensor] = None,
                    noise: Optional[torch.Tensor] = None, seed: int = 0, stop_ids: Optional[torch.Tensor] = None,
                    forced_k: Optional[torch.Tensor] = None, clamp_tail: bool = False, inject: bool = False):
        """target_logits: [R*bs, V] bf16 (last dim contiguous); hidden[s]: [R*bs, H] bf16 contiguous.
        inject: also run the NEXT cycle's context injection, overlapped with the verify kernel and reading `hidden` in
        place (dflash_verify_inject_step); the next draft step then skips its injection kernel."""
        arr = (c_void_p * self.n_sel)(*[h.data_ptr() for h in hidden])
        ld = 0 if target_logits is None else target_logits.stride(-2)
        n_stop = 0 if stop_ids is None else int(stop_ids.numel())
        fld = 0 if forced_k is None else int(forced_k.shape[-1])
        self._call("dflash_verify_inject_step" if inject else "dflash_verify_step", self.handle, _p(target_logits), ld,
                   _p(posterior_in), arr, float(temperature), _p(noise), int(seed) & (2**64 - 1), _p(stop_ids),
                   n_stop, _p(forced_k), fld, int(clamp_tail), _stream(self.device))
        self._injected = "fresh" if inject else "no"
        self._keep_hidden = list(hidden) if inject else None  # re-read in place by the injection kernel

    def draft_step_sampled(self, temperature: float, seed: int = 0):
        """Draft step whose tokens are drawn from softmax(draft_logits / temperature) in the lm_head epilogue
        (benchmark_dynamic_schedule.py:342); temperature 0 = the greedy step."""
        if self._injected != "no":
            return self.draft_step_injected(temperature=temperature, seed=seed)
        self._call("dflash_draft_step_sampled", self.handle, float(temperature), int(seed) & (2**64 - 1),
                                                      _stream(self.device))

    def draft_step_candidates(self, n_candidates: int, fixed_prefix_len: int):
        """Draft step with the top-4 lm_head epilogue; fills cand_ids[:, :n_candidates] / cand_scores
        (fixed_prefix_rank candidates, benchmark_candidate_solutions.py:181-249)."""
        self._call("dflash_draft_step_candidates", self.handle, int(n_candidates), int(fixed_prefix_len), _stream(self.device))
        self._injected = "no"

    def verify_step_candidates(self, n_candidates: int, target_logits: torch.Tensor, hidden: Sequence[torch.Tensor], *,
                               temperature: float = 0.0, noise: Optional[torch.Tensor] = None, seed: int = 0,
                               stop_ids: Optional[torch.Tensor] = None, clamp_tail: bool = False):
        """target_logits [R*K*bs, V] bf16, hidden[s] [R*K*bs, H] bf16 from ONE target forward over all candidates."""
        arr = (c_void_p * self.n_sel)(*[h.data_ptr() for h in hidden])
        n_stop = 0 if stop_ids is None else int(stop_ids.numel())
        self._call("dflash_verify_step_candidates", self.handle, int(n_candidates), _p(target_logits),
                                                          target_logits.stride(-2), arr, float(temperature), _p(noise),
                                                          int(seed) & (2**64 - 1), _p(stop_ids), n_stop, int(clamp_tail),
                                                          _stream(self.device))
        self._injected = "no"

    def sample(self, logits: torch.Tensor, temperature: float, seed: int = 0, noise: Optional[torch.Tensor] = None):
        """sample() of model/utils.py:27-34 on [rows, V] bf16 logits -> int64 [rows]."""
        rows, V = logits.shape
        nsplit = 8
        sv = torch.empty(rows * nsplit, dtype=torch.float32, device=logits.device)
        si = torch.empty(rows * nsplit, dtype=torch.int32, device=logits.device)
        out = torch.empty(rows, dtype=torch.int64, device=logits.device)
        self._call("dflash_sample", _p(logits), logits.stride(0), rows, V, float(temperature), _p(noise),
                                          int(seed) & (2**64 - 1), _p(sv), _p(si), nsplit, _p(out), _stream(self.device))
        return out

    # ------------------------------------------------------------------------------------------
    def capture_draft_graph(self, injected: bool = False):
        """Capture one draft step (static pointers, device-resident state) into a CUDA graph: the full step, or
        (injected) the step without its injection kernel."""
        torch.cuda.synchronize(self.device)
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        # the warm-up step runs on live request state: keep the block (slots 1.. are mask tokens until the first real
        # draft step fills them, model/dflash.py:233-235) and the draft tokens it would overwrite
        keep_blk, keep_tok = self.buf["block_ids"].clone(), self.buf["draft_tokens"].clone()
        state = self._injected
        with torch.cuda.stream(s):
            # warm up (module load, func attributes) outside capture
            if injected:
                self._call("dflash_draft_step_injected", self.handle, 1, 0.0, 0, _stream(self.device))
            else:
                self._call("dflash_draft_step", self.handle, None, 1, _stream(self.device))
            self.buf["block_ids"].copy_(keep_blk)
            self.buf["draft_tokens"].copy_(keep_tok)
            if injected:  # the warm-up consumed the block rows: put them back
                self.embed_block()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            if injected:
                self._call("dflash_draft_step_injected", self.handle, 1, 0.0, 0, _stream(self.device))
            else:
                self._call("dflash_draft_step", self.handle, None, 1, _stream(self.device))
        self._injected = state
        if injected:
            self._graph_injected = g
        else:
            self._graph = g
        return g

    def draft_step_graphed(self):
        """One draft step replayed from a CUDA graph: without the injection kernel when the previous
        verify_step(inject=True) (or the prompt pass) has already prepared the activation rows."""
        if self._injected != "no":
            if self._graph_injected is None:
                self.capture_draft_graph(injected=True)
            if self._injected == "stale":
                self.embed_block()
            self._graph_injected.replay()
            self._injected = "stale"
            return
        if self._graph is None:
            self.capture_draft_graph()
        self._graph.replay()
