"""Golden vectors for `config.attention_bias = True` (model/dflash.py:41-50: q/k/v/o projections with bias) from the
UNMODIFIED reference. Run in the build container only:  python tests/golden/make_bias_golden.py
Writes tests/golden/reference_bias.pt; tests/test_oracle_golden.py replays it against the oracle."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.tiny_models import TINY, build_pair, fingerprint, rig_lm_head  # noqa: E402
from tests.golden.make_golden import LIVE  # noqa: E402

sys.path.insert(0, "/root/reference")


def main():
    from model import DFlashDraftModel  # the reference, unchanged
    from transformers import DynamicCache
    torch.set_num_threads(4)
    bs = 16
    target, draft = build_pair(DFlashDraftModel, seed=1234, block_size=bs, attention_bias=True)
    assert draft.layers[0].self_attn.q_proj.bias is not None and draft.layers[0].self_attn.o_proj.bias is not None
    rig_lm_head(target, live=LIVE, seed=99)
    fp = fingerprint(target, draft)
    prompt = torch.randint(0, TINY["vocab"] - 1, (1, 12), generator=torch.Generator().manual_seed(7))
    ids = draft.spec_generate(target, prompt, max_new_tokens=48, stop_token_ids=None, temperature=0.0)
    ar = target.generate(prompt, max_new_tokens=48, do_sample=False)
    g = torch.Generator().manual_seed(11)
    H, nsel = TINY["hidden"], len(draft.target_layer_ids)
    th_old = torch.randn(1, 5, nsel * H, generator=g)
    th_new = torch.randn(1, 3, nsel * H, generator=g)
    noise = torch.randn(1, bs, H, generator=g)
    cache = DynamicCache()
    h0 = draft(target_hidden=th_old, noise_embedding=noise, position_ids=torch.arange(0, 5 + bs).unsqueeze(0),
               past_key_values=cache, use_cache=True, is_causal=False)
    cache.crop(5)
    h1 = draft(target_hidden=th_new, noise_embedding=noise, position_ids=torch.arange(5, 8 + bs).unsqueeze(0),
               past_key_values=cache, use_cache=True, is_causal=False)
    out = dict(prompt=prompt, output=ids, autoregressive=ar, fingerprint=fp, th_old=th_old, th_new=th_new, noise=noise,
               h0=h0.detach(), h1=h1.detach())
    path = os.path.join(HERE, "reference_bias.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes; n_out", ids.shape[1] - prompt.shape[1])


if __name__ == "__main__":
    main()
