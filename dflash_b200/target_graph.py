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
from transformers import StaticCache

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
