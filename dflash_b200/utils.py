"""Hot-path helpers with the reference's names and signatures (model/utils.py:4-34).

`build_target_layer_ids` and `extract_context_feature` are host-side index/view logic. `sample`
dispatches CUDA bf16 logits to the fused sampler in libdflash_b200.so.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


def build_target_layer_ids(num_target_layers: int, num_draft_layers: int) -> List[int]:
    """Evenly spaced target layer ids in [1, L-3] (model/utils.py:4-14)."""
    if num_draft_layers == 1:
        return [num_target_layers // 2]
    start = 1
    end = num_target_layers - 3
    span = end - start
    return [int(round(start + (i * span) / (num_draft_layers - 1))) for i in range(num_draft_layers)]


def select_context_states(hidden_states: Sequence[torch.Tensor], layer_ids: Optional[Sequence[int]]) -> List[torch.Tensor]:
    """The tensors extract_context_feature would concatenate: hidden_states[id + 1] per selected layer
    (entry 0 is the embedding output). The CUDA path gathers from these directly and never builds the cat."""
    return [hidden_states[i + 1] for i in layer_ids]


class ContextTap:
    """Forward hooks on the selected target decoder layers. `hidden_states[id + 1]` of a HF
    causal LM is the output of decoder layer `id` (entry 0 is the embedding output; only the LAST entry is
    post-norm, and `build_target_layer_ids` never selects it), so hooking those layers yields the same tensors
    `output_hidden_states=True` would, without the model keeping all L + 1 residual streams alive. The target
    module's arithmetic is untouched.

        with ContextTap(target, ids) as tap:
            out = target(input_ids, ...)          # no output_hidden_states
        hs = tap.states                           # == [out.hidden_states[i + 1] for i in ids]
    """

    def __init__(self, target, layer_ids: Sequence[int]):
        layers = target.model.layers
        n = len(layers)
        for i in layer_ids:
            if not 0 <= i < n - 1:
                raise ValueError(f"target layer id {i} is not a raw residual-stream output of a {n}-layer target")
        self.layer_ids = list(layer_ids)
        self._layers = [layers[i] for i in self.layer_ids]
        self.states: List[Optional[torch.Tensor]] = [None] * len(self.layer_ids)
        self._handles = []

    def _hook(self, slot):
        def fn(_module, _args, output):
            self.states[slot] = output[0] if isinstance(output, (tuple, list)) else output
        return fn

    def __enter__(self):
        self._handles = [m.register_forward_hook(self._hook(s)) for s, m in enumerate(self._layers)]
        return self

    def __exit__(self, *exc):
        for h in self._handles:
            h.remove()
        self._handles = []
        return False


def extract_context_feature(hidden_states: Sequence[torch.Tensor], layer_ids: Optional[Sequence[int]]) -> torch.Tensor:
    """Same result as the reference (model/utils.py:16-25); kept for callers that want the tensor."""
    return torch.cat(select_context_states(hidden_states, layer_ids), dim=-1)


def sample(logits: torch.Tensor, temperature: float = 0.0, *, seed: Optional[int] = None,
           noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B, S, V] logits -> [B, S] int64 tokens (model/utils.py:27-34): argmax below T=1e-5, otherwise a draw
    from softmax(logits / T) by the exponential race torch.multinomial uses, fused into one pass over the
    vocab in CUDA. `noise` (fp32 [B*S, V] Exp(1) draws) reproduces a given torch.multinomial stream."""
    from . import _lib
    from .engine import _declare, _p, _stream
    if not logits.is_cuda:
        raise _lib.DFlashNativeError("dflash_b200.sample needs CUDA logits; there is no CPU path")
    lib = _lib.load()
    _declare(lib)
    bsz, seq_len, vocab = logits.shape
    flat = logits.reshape(-1, vocab)
    if flat.dtype != torch.bfloat16:
        flat = flat.to(torch.bfloat16)
    if flat.stride(-1) != 1:
        flat = flat.contiguous()
    rows, nsplit = flat.shape[0], 8
    sv = torch.empty(rows * nsplit, dtype=torch.float32, device=flat.device)
    si = torch.empty(rows * nsplit, dtype=torch.int32, device=flat.device)
    out = torch.empty(rows, dtype=torch.int64, device=flat.device)
    if seed is None:
        seed = int(torch.randint(0, 2**62, (1,)).item()) if temperature >= 1e-5 else 0
    with torch.cuda.device(flat.device):
        _lib.check(lib.dflash_sample(_p(flat), flat.stride(0), rows, vocab, float(temperature), _p(noise),
                                     int(seed), _p(sv), _p(si), nsplit, _p(out), _stream()), "dflash_sample")
    return out.view(bsz, seq_len)
