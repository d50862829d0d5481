"""Tiny seeded target/draft pairs shared by the golden generator, the oracle tests and the GPU parity tests.

Weights are drawn per parameter NAME from a seeded generator (not from the module init order), so the
reference class, the product class and the bare oracle state dict all get identical values.
"""
from __future__ import annotations

import hashlib

import torch

TINY = dict(vocab=1000, hidden=256, intermediate=512, target_layers=6, draft_layers=2, heads=2, kv_heads=1,
            head_dim=128, max_pos=4096, eps=1e-6, rope_theta=1_000_000.0)


def target_config(dims=TINY):
    from transformers import Qwen3Config
    return Qwen3Config(vocab_size=dims["vocab"], hidden_size=dims["hidden"], intermediate_size=dims["intermediate"],
                       num_hidden_layers=dims["target_layers"], num_attention_heads=dims["heads"],
                       num_key_value_heads=dims["kv_heads"], head_dim=dims["head_dim"],
                       max_position_embeddings=dims["max_pos"], rms_norm_eps=dims["eps"], tie_word_embeddings=False,
                       rope_parameters={"rope_type": "default", "rope_theta": dims["rope_theta"]})


def draft_config(block_size=16, dims=TINY, target_layer_ids=None, attention_bias=False):
    from transformers import Qwen3Config
    cfg = Qwen3Config(attention_bias=attention_bias, vocab_size=dims["vocab"], hidden_size=dims["hidden"], intermediate_size=dims["intermediate"],
                      num_hidden_layers=dims["draft_layers"], num_attention_heads=dims["heads"],
                      num_key_value_heads=dims["kv_heads"], head_dim=dims["head_dim"],
                      max_position_embeddings=dims["max_pos"], rms_norm_eps=dims["eps"],
                      rope_parameters={"rope_type": "default", "rope_theta": dims["rope_theta"]})
    cfg.num_target_layers = dims["target_layers"]
    cfg.block_size = block_size
    cfg.dflash_config = {"mask_token_id": dims["vocab"] - 1}
    if target_layer_ids is not None:
        cfg.dflash_config["target_layer_ids"] = list(target_layer_ids)
    return cfg


def seeded_fill_(module_or_sd, seed: int):
    """Overwrite every parameter, visiting names in sorted order: matrices ~ N(0, 0.05^2) scaled for
    width, norm vectors ~ 1 + 0.1 N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    items = module_or_sd.items() if isinstance(module_or_sd, dict) else dict(module_or_sd.named_parameters()).items()
    with torch.no_grad():
        for name, p in sorted(items, key=lambda kv: kv[0]):
            if name.endswith(".bias"):
                v = 0.1 * torch.randn(p.shape, generator=g)
            elif p.dim() == 1:
                v = 1.0 + 0.1 * torch.randn(p.shape, generator=g)
            else:
                v = torch.randn(p.shape, generator=g) * (1.0 / (p.shape[-1] ** 0.5))
            p.copy_(v.to(p.dtype))


def build_pair(draft_cls, seed=1234, block_size=16, dims=TINY, dtype=torch.float32, device="cpu", attention_bias=False):
    """(target HF Qwen3ForCausalLM, draft draft_cls) with seeded weights, eval mode."""
    from transformers import Qwen3ForCausalLM
    target = Qwen3ForCausalLM(target_config(dims))
    draft = draft_cls(draft_config(block_size, dims, attention_bias=attention_bias))
    seeded_fill_(target, seed)
    seeded_fill_(draft, seed + 1)  # (with attention_bias the q/k/v/o biases are ~N(0, 0.1^2))
    target = target.to(dtype=dtype, device=device).eval()
    draft = draft.to(dtype=dtype, device=device).eval()
    return target, draft


def rig_lm_head(target, live=(3, 17, 101), seed=99):
    """Zero every lm_head row except `live`: both the target posterior and the draft's tokens (which go
    through the same head) then come from a 3-token alphabet and agree often, so acceptance lengths
    cover 1..block_size without trained weights."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        w = target.lm_head.weight
        new = torch.zeros_like(w, device="cpu", dtype=torch.float32)
        for t in live:
            new[t] = torch.randn(w.shape[1], generator=g) * (1.0 / (w.shape[1] ** 0.5))
        w.copy_(new.to(dtype=w.dtype, device=w.device))


def fingerprint(*modules) -> str:
    h = hashlib.sha256()
    for m in modules:
        for name, p in sorted(dict(m.named_parameters()).items()):
            h.update(name.encode())
            h.update(p.detach().to(torch.float32).cpu().contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


def draft_state_dict(draft) -> dict:
    return {k: v.detach() for k, v in draft.state_dict().items()}
