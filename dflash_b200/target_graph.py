"""the target's verify forward under a CUDA graph with a static KV cache.

The target stays the caller's unmodified HF module (`model/dflash.py:249-255` calls it with a `DynamicCache`,
whose `torch.cat` per layer per cycle and ~1.5 k eager launches are ~86 % of the reference's cycle time,
`results.md:363`). Here the same module is called with transformers' `StaticCache`; the reference's
`past_key_values_target.crop(start)` (`dflash.py:262`) becomes a device-side write of the cache's
`cumulative_length` from the engine's `start` buffer, and positions come from the same buffer, so the
whole verify forward replays from a CUDA graph with no host round trip. Module math is untouched.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
from transformers import DynamicCache, StaticCache
from transformers.cache_utils import Cache

from .utils import ContextTap


class GraphedVerifyTarget:
    def __init__(self, target, block_size: int, max_cache_len: int, layer_ids: Sequence[int],
                 start_buf: torch.Tensor, block_ids: torch.Tensor):
        """start_buf: device int32 [1] (engine `start`), block_ids: device int64 [1, block_size] (engine buffer)."""
        self.target = target
        self.bs = int(block_size)
        self.layer_ids = list(layer_ids)
        self.start_buf = start_buf
        self.block_ids = block_ids
        self.device = block_ids.device
        self.max_cache_len = int(max_cache_len)
        self.cache = StaticCache(config=target.config, max_cache_len=self.max_cache_len)
        self.pos = torch.zeros(1, self.bs, dtype=torch.long, device=self.device)
        self._arange = torch.arange(self.bs, dtype=torch.long, device=self.device).unsqueeze(0)
        self.graph = None
        self.logits = None
        self.hidden: List[torch.Tensor] = []

    # -- prompt: eager, once (positions [0, P)) ------------------------------------------------------
    def prefill(self, input_ids: torch.Tensor):
        P = input_ids.shape[1]
        if P + self.bs > self.max_cache_len:
            raise ValueError("static target cache too small for the prompt")
        for layer in self.cache.layers:
            if getattr(layer, "is_initialized", False):
                layer.cumulative_length.zero_()
        pos = torch.arange(P, device=self.device).unsqueeze(0)
        with ContextTap(self.target, self.layer_ids) as tap:
            out = self.target(input_ids, position_ids=pos, past_key_values=self.cache, use_cache=True, logits_to_keep=1)
        return out.logits, list(tap.states)

    # -- one verify forward over the engine's current block -------------------------------------------
    def _forward(self):
        start = self.start_buf[0:1].to(torch.long)
        self.pos.copy_(self._arange + start)
        for layer in self.cache.layers:  # == past_key_values_target.crop(start): a length write
            layer.cumulative_length.copy_(start.view(layer.cumulative_length.shape).to(layer.cumulative_length.dtype))
        with ContextTap(self.target, self.layer_ids) as tap:  # hooks fire during capture: their outputs are static
            out = self.target(self.block_ids, position_ids=self.pos, past_key_values=self.cache, use_cache=True)
        return out.logits, list(tap.states)

    def capture(self):
        torch.cuda.synchronize(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up outside capture (allocator, lazy inits, autotune)
                self._forward()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            logits, hidden = self._forward()
        self.graph = g
        self.logits = logits[0]                      # [bs, V]
        self.hidden = [h[0] for h in hidden]         # n_sel x [bs, H]
        return self

    def verify_forward(self):
        """Runs the captured forward on the current stream; returns (logits [bs, V], hidden n_sel x [bs, H]) —
        static tensors that the next replay overwrites."""
        if self.graph is None:
            self.capture()
        self.graph.replay()
        return self.logits, self.hidden


class GraphedCandidateTarget(GraphedVerifyTarget):
    """Multi-candidate verify on the graphed target: the K candidate blocks of a cycle are the K
    rows of ONE batch-K forward over a batch-K static cache. The reference clones and repeats its whole `DynamicCache`
    per cycle and selects the winning branch afterwards (`benchmark_candidate_solutions.py:574-577,610-614`); here all
    K rows share the committed prefix by construction and only the last block differs, so keeping the winner is a
    copy of <= block_size cache rows per layer -- done at the START of the next replay, from the engine's device-side
    `chosen` / `start`, so the cycle still has no host round trip.

    start_buf: device int32 [1]; cand_ids: device int64 [K, block_size] (engine buffer); chosen_buf: device int32 [1]."""

    def __init__(self, target, block_size: int, n_candidates: int, max_cache_len: int, layer_ids: Sequence[int],
                 start_buf: torch.Tensor, cand_ids: torch.Tensor, chosen_buf: torch.Tensor):
        super().__init__(target, block_size, max_cache_len, layer_ids, start_buf, cand_ids)
        self.K = int(n_candidates)
        self.chosen_buf = chosen_buf
        self.pos = torch.zeros(self.K, self.bs, dtype=torch.long, device=self.device)

    def prefill(self, input_ids: torch.Tensor):
        """The prompt goes through all K cache rows (identical rows: K x the prompt's compute, once)."""
        logits, hidden = super().prefill(input_ids.expand(self.K, -1))
        self.chosen_buf.zero_()
        return logits[0:1], [h[0:1] for h in hidden]

    def _forward(self):
        start = self.start_buf[0:1].to(torch.long)
        chosen = self.chosen_buf[0:1].to(torch.long)
        # keep the previous cycle's winner: rows [start - bs, start) of batch row `chosen` -> every batch row. (The
        # accepted tokens sit at [previous start, start) inside that window; older rows are equal already.)
        idx = (self._arange[0] + start - self.bs).clamp_min(0)
        for layer in self.cache.layers:
            for t in (layer.keys, layer.values):
                win = t.index_select(0, chosen).index_select(2, idx)        # [1, H, bs, D]
                t.index_copy_(2, idx, win.expand(t.shape[0], -1, -1, -1))
            layer.cumulative_length.copy_(start.view(layer.cumulative_length.shape).to(layer.cumulative_length.dtype))
        self.pos.copy_((self._arange + start).expand(self.K, -1))
        with ContextTap(self.target, self.layer_ids) as tap:
            out = self.target(self.block_ids, position_ids=self.pos, past_key_values=self.cache, use_cache=True)
        return out.logits, list(tap.states)

    def verify_forward(self):
        if self.graph is None:
            self._capture_all()
        self.graph.replay()
        return self.logits, self.hidden

    def _capture_all(self):
        torch.cuda.synchronize(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._forward()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            logits, hidden = self._forward()
        self.graph = g
        self.logits = logits.reshape(self.K * self.bs, -1)
        self.hidden = [h.reshape(self.K * self.bs, -1) for h in hidden]


class _RaggedStaticLayer:
    """One layer of `RaggedStaticCache`: K/V `[R, Hkv, max_len, D]`, every batch row at its own length. `update` scatters
    row r's new K/V to positions `owner.write_pos[r, :]` and returns the first `owner.kv_len` positions of the buffers
    (the additive mask built from `start[r]` hides whatever lies beyond a row's own length)."""

    is_compileable = True
    is_sliding = False

    def __init__(self, owner):
        self.owner = owner
        self.keys = self.values = None
        self.is_initialized = False

    def lazy_initialization(self, key_states, value_states):
        o = self.owner
        R, Hkv = key_states.shape[:2]
        self.keys = torch.zeros(R, Hkv, o.capacity, key_states.shape[-1], dtype=key_states.dtype, device=key_states.device)
        self.values = torch.zeros(R, Hkv, o.capacity, value_states.shape[-1], dtype=value_states.dtype,
                                  device=value_states.device)
        self.is_initialized = True

    def update(self, key_states, value_states, *args, **kwargs):
        if not self.is_initialized:
            self.lazy_initialization(key_states, value_states)
        o = self.owner
        idx = o.write_pos[:, None, :, None]
        self.keys.scatter_(2, idx.expand_as(key_states), key_states)
        self.values.scatter_(2, idx.expand_as(value_states), value_states)
        return self.keys[:, :, :o.kv_len], self.values[:, :, :o.kv_len]

    def get_seq_length(self):
        return 0

    def get_mask_sizes(self, query_length):
        return self.owner.kv_len, 0

    def get_max_cache_shape(self):
        return self.owner.capacity


class RaggedStaticCache(Cache):
    """A `transformers.Cache` for R request streams that advance independently (ragged acceptance lengths): static
    `[R, Hkv, max_len, D]` buffers per layer, new rows scattered to per-stream positions. The reference's
    `past_key_values_target.crop(start)` (`model/dflash.py:262`) needs no data motion at all here: a stream's length IS
    the engine's `start[r]`, and the mask / write positions are derived from it on the device."""

    def __init__(self, n_layers: int, max_cache_len: int):
        self.capacity = int(max_cache_len)
        self.kv_len = self.capacity
        self.write_pos = None  # [R, q] long, set by the owner before every forward
        super().__init__(layers=[_RaggedStaticLayer(self) for _ in range(n_layers)])

    def get_seq_length(self, layer_idx: int = 0):
        return 0


RAGGED_ATTN = "dflash_ragged_sdpa"


def _ragged_sdpa_attention(module, query, key, value, attention_mask, dropout: float = 0.0, scaling=None, **kwargs):
    """Attention backend of the batched verify forward, registered with transformers' `AttentionInterface` (the
    library's plug-in point; the target module is untouched). Same arithmetic as the stock "sdpa" backend --
    `F.scaled_dot_product_attention` with the boolean mask -- but grouped-query attention is expressed by folding the
    q heads of a KV group into the query axis ([B, Hkv, g * q, D] against the UNrepeated [B, Hkv, S, D] cache) instead of
    `repeat_kv`, which transformers falls back to whenever a mask is present and which copies the whole static cache
    g times per layer per cycle (measured: 19.8 of 47 ms of a 64-stream forward of a Qwen3-8B target)."""
    B, Hq, q_len, D = query.shape
    Hkv = key.shape[1]
    g = Hq // Hkv
    mask = attention_mask
    if g > 1:
        query = query.reshape(B, Hkv, g * q_len, D)
        if mask is not None:
            mask = mask.expand(B, 1, q_len, mask.shape[-1]).repeat(1, 1, g, 1)  # row (h, j) of a group -> query j
    out = torch.nn.functional.scaled_dot_product_attention(query, key, value, attn_mask=mask, dropout_p=0.0,
                                                           scale=scaling, is_causal=False)
    out = out.reshape(B, Hq, q_len, out.shape[-1]).transpose(1, 2).contiguous()
    return out, None


def _register_ragged_attention() -> bool:
    try:
        from transformers import AttentionInterface
        if RAGGED_ATTN not in AttentionInterface._global_mapping:
            AttentionInterface.register(RAGGED_ATTN, _ragged_sdpa_attention)
        return True
    except Exception:  # noqa: BLE001 -- older transformers: keep the stock backend
        return False


class _attn_backend:
    """with _attn_backend(target, name): the target's config selects attention backend `name` for the duration."""

    def __init__(self, target, name):
        self.cfg, self.name, self.prev = target.config, name, None

    def __enter__(self):
        if self.name is not None:
            self.prev = self.cfg._attn_implementation
            self.cfg._attn_implementation = self.name

    def __exit__(self, *exc):
        if self.name is not None:
            self.cfg._attn_implementation = self.prev
        return False


class BatchedVerifyTarget:
    """ONE verify forward of the caller's unmodified HF target for all R request streams of an engine per cycle
    (`model/dflash.py:249-255`, batched): input = the engine's `block_ids [R, bs]`, `position_ids[r, j] = start[r] + j`,
    a 4-D boolean mask `[R, 1, bs, kv_len]` (key p visible to query j of stream r iff p <= start[r] + j: causal inside
    the block, the stream's own committed prefix before it, nothing stale), K/V in a `RaggedStaticCache`. Everything
    is derived from the engine's device-side `start`, so the forward replays from a CUDA graph with no host round trip;
    graphs are kept per KV bucket (`kv_len` a multiple of `bucket`) so that short contexts do not pay for
    `max_cache_len` keys. Falls back to calling the same forward eagerly if the target cannot be captured.

    start_buf: device int32 [R]; block_ids: device int64 [R, bs] (engine buffers)."""

    def __init__(self, target, block_size: int, n_streams: int, max_cache_len: int, layer_ids: Sequence[int],
                 start_buf: torch.Tensor, block_ids: torch.Tensor, bucket: int = 512, use_graph: bool = True):
        if any(t != "full_attention" for t in (getattr(target.config, "layer_types", None) or [])):
            raise NotImplementedError("BatchedVerifyTarget: sliding-window target layers are not supported")
        self.target, self.bs, self.R = target, int(block_size), int(n_streams)
        self.layer_ids = list(layer_ids)
        self.start_buf, self.block_ids = start_buf, block_ids
        self.device = block_ids.device
        self.bucket = int(bucket)
        self.max_cache_len = -(-int(max_cache_len) // self.bucket) * self.bucket
        self.cache = RaggedStaticCache(target.config.num_hidden_layers, self.max_cache_len)
        self.pos = torch.zeros(self.R, self.bs, dtype=torch.long, device=self.device)
        self.cache.write_pos = self.pos
        self._arange = torch.arange(self.bs, dtype=torch.long, device=self.device).unsqueeze(0)
        self._keys = torch.arange(self.max_cache_len, dtype=torch.long, device=self.device)
        self.use_graph = bool(use_graph)
        # GQA without repeat_kv (see _ragged_sdpa_attention) when the target runs the stock sdpa backend
        self.attn_backend = RAGGED_ATTN if (getattr(target.config, "_attn_implementation", None) == "sdpa" and
                                            _register_ragged_attention()) else None
        self.graphs = {}      # kv_len -> (graph, logits [R*bs, V], hidden n_sel x [R*bs, H])
        self.n_forwards = 0   # verify forwards issued (one per cycle whatever R is)

    # -- prompt of one stream: eager, batch 1, exactly the reference's prefill call (dflash.py:218-225); its K/V rows
    #    are then copied into row r of the static buffers
    def prefill(self, r: int, input_ids: torch.Tensor):
        P = input_ids.shape[1]
        if P + 2 * self.bs > self.max_cache_len:
            raise ValueError("static target cache too small for the prompt")
        cache = DynamicCache()
        pos = torch.arange(P, device=self.device).unsqueeze(0)
        with ContextTap(self.target, self.layer_ids) as tap:
            out = self.target(input_ids, position_ids=pos, past_key_values=cache, use_cache=True, logits_to_keep=1)
        for l, layer in enumerate(cache.layers):
            dst = self.cache.layers[l]
            if not dst.is_initialized:
                shape_k, shape_v = list(layer.keys.shape), list(layer.values.shape)
                shape_k[0] = shape_v[0] = self.R
                dst.lazy_initialization(layer.keys.new_empty(shape_k[:2] + [0, shape_k[3]]).expand(self.R, -1, -1, -1),
                                        layer.values.new_empty(shape_v[:2] + [0, shape_v[3]]).expand(self.R, -1, -1, -1))
            dst.keys[r, :, :P].copy_(layer.keys[0])
            dst.values[r, :, :P].copy_(layer.values[0])
        return out.logits, list(tap.states)

    def _forward(self, kv_len: int):
        start = self.start_buf[: self.R].to(torch.long).unsqueeze(1)
        # (a finished / empty stream keeps computing on whatever its start is; keep its writes inside the buffers)
        self.pos.copy_((self._arange + start).clamp_(max=kv_len - 1))
        mask = (self._keys[:kv_len].view(1, 1, 1, kv_len) <= self.pos.view(self.R, 1, self.bs, 1))
        self.cache.kv_len = kv_len
        # (hooks fire during capture: their outputs are static)
        with ContextTap(self.target, self.layer_ids) as tap, _attn_backend(self.target, self.attn_backend):
            out = self.target(self.block_ids, position_ids=self.pos, attention_mask=mask, past_key_values=self.cache,
                              use_cache=True)
        V = out.logits.shape[-1]
        return out.logits.reshape(self.R * self.bs, V), [h.reshape(self.R * self.bs, h.shape[-1]) for h in tap.states]

    def _capture(self, kv_len: int):
        torch.cuda.synchronize(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up outside capture (allocator, lazy inits, autotune)
                self._forward(kv_len)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            logits, hidden = self._forward(kv_len)
        self.graphs[kv_len] = (g, logits, hidden)

    def bucket_for(self, max_end: int) -> int:
        """Smallest KV bucket that holds positions [0, max_end)."""
        return min(self.max_cache_len, -(-max(int(max_end), 1) // self.bucket) * self.bucket)

    def verify_forward(self, max_end: int):
        """One forward for all R streams. `max_end`: a host-side upper bound of max_r(start[r]) + block_size (it picks
        the KV bucket; the exact lengths are read from the device). Returns (logits [R*bs, V], hidden n_sel x
        [R*bs, H]); with graphs these are static tensors that the next replay of the same bucket overwrites."""
        kv_len = self.bucket_for(max_end)
        self.n_forwards += 1
        if self.use_graph:
            if kv_len not in self.graphs:
                try:
                    self._capture(kv_len)
                except Exception as ex:  # noqa: BLE001 -- capture failures surface as many exception types
                    import warnings
                    warnings.warn(f"dflash_b200: the batched target forward could not be captured in a CUDA graph "
                                  f"({type(ex).__name__}: {ex}); calling it eagerly")
                    self.use_graph = False
                    torch.cuda.synchronize(self.device)
            if self.use_graph:
                g, logits, hidden = self.graphs[kv_len]
                g.replay()
                return logits, hidden
        return self._forward(kv_len)
