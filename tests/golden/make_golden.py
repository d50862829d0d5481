"""Generate golden vectors from the UNMODIFIED reference (imports /root/reference/model).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/*.pt (small). tests/test_oracle_golden.py replays them against oracle/.

Cases (all tiny, fp32, CPU, seeded; weights are re-created from the seed by the tests and
fingerprinted here so a torch init drift is detected instead of silently compared):
  honest     random-init target + draft, greedy: tau == 1 every cycle
  rigged     target lm_head has only 12 live rows (a small alphabet with repeats and a usable stop token)
  forced     forced-acceptance harness mode: tau = k + 1 for a seeded schedule k in [0, bs-1]
  sampled    rigged weights at temperature 1.0 (torch.multinomial under torch.manual_seed)
  stop       rigged weights with a stop token
  forward    one DFlashDraftModel.forward call with a pre-filled DynamicCache
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.tiny_models import TINY, build_pair, fingerprint, rig_lm_head  # noqa: E402

sys.path.insert(0, "/root/reference")
LIVE = (3, 17, 101, 250, 251, 400, 512, 640, 777, 801, 900, 998)
FORCED = (3, 0, 7, 15, 1, 5, 2, 11)


def main():
    from model import DFlashDraftModel  # the reference, unchanged
    from transformers import DynamicCache

    torch.set_num_threads(4)
    out = {}
    for name, bs in (("bs16", 16), ("bs8", 8)):
        target, draft = build_pair(DFlashDraftModel, seed=1234, block_size=bs)
        fp = fingerprint(target, draft)
        prompt = torch.randint(0, TINY["vocab"] - 1, (1, 12), generator=torch.Generator().manual_seed(7))
        # honest
        ids = draft.spec_generate(target, prompt, max_new_tokens=24, stop_token_ids=None, temperature=0.0)
        ar = target.generate(prompt, max_new_tokens=24, do_sample=False)
        out[f"{name}/honest"] = dict(prompt=prompt, output=ids, autoregressive=ar, fingerprint=fp)
        # rigged
        rig_lm_head(target, live=LIVE, seed=99)
        fp_r = fingerprint(target, draft)
        ids = draft.spec_generate(target, prompt, max_new_tokens=48, stop_token_ids=None, temperature=0.0)
        ar = target.generate(prompt, max_new_tokens=48, do_sample=False)
        out[f"{name}/rigged"] = dict(prompt=prompt, output=ids, autoregressive=ar, fingerprint=fp_r)
        # rigged + stop token: the token the rigged stream emits at generated position 9
        stop = [int(ids[0, prompt.shape[1] + 9])]
        ids = draft.spec_generate(target, prompt, max_new_tokens=48, stop_token_ids=stop, temperature=0.0)
        out[f"{name}/stop"] = dict(prompt=prompt, output=ids, stop=stop, fingerprint=fp_r)
        # forced acceptance (SURVEY §4 harness mode): the reference's loop is untouched, only the `sample`
        # symbol it calls is wrapped so that posterior[:, :k] = drafted tokens[:, :k] for a seeded k schedule.
        import model.dflash as ref_mod
        orig_sample = ref_mod.sample
        st = dict(draft=None, cycle=0)

        def forced_sample(logits, temperature=None):
            if temperature is None:  # draft call: sample(draft_logits)
                st["draft"] = orig_sample(logits)
                return st["draft"]
            post = orig_sample(logits, temperature)
            if st["draft"] is not None:  # not the prefill call
                k = min(FORCED[st["cycle"] % len(FORCED)], post.shape[1] - 1)
                post[:, :k] = st["draft"][:, :k]
                st["cycle"] += 1
            return post

        ref_mod.sample = forced_sample
        try:
            ids = draft.spec_generate(target, prompt, max_new_tokens=64, stop_token_ids=None, temperature=0.0)
        finally:
            ref_mod.sample = orig_sample
        out[f"{name}/forced"] = dict(prompt=prompt, output=ids, forced=list(FORCED), fingerprint=fp_r)
        # sampled
        torch.manual_seed(4321)
        ids = draft.spec_generate(target, prompt, max_new_tokens=32, stop_token_ids=None, temperature=1.0)
        out[f"{name}/sampled"] = dict(prompt=prompt, output=ids, seed=4321, fingerprint=fp_r)
        # one forward call with an existing cache: ctx 5 rows cached, 3 new ctx rows, bs block rows
        g = torch.Generator().manual_seed(11)
        H, nsel = TINY["hidden"], len(draft.target_layer_ids)
        th_old = torch.randn(1, 5, nsel * H, generator=g)
        th_new = torch.randn(1, 3, nsel * H, generator=g)
        noise = torch.randn(1, bs, H, generator=g)
        cache = DynamicCache()
        pos = torch.arange(0, 5 + bs).unsqueeze(0)
        h0 = draft(target_hidden=th_old, noise_embedding=noise, position_ids=pos, past_key_values=cache,
                   use_cache=True, is_causal=False)
        cache.crop(5)
        pos = torch.arange(5, 8 + bs).unsqueeze(0)
        h1 = draft(target_hidden=th_new, noise_embedding=noise, position_ids=pos, past_key_values=cache,
                   use_cache=True, is_causal=False)
        out[f"{name}/forward"] = dict(th_old=th_old, th_new=th_new, noise=noise, h0=h0.detach(), h1=h1.detach(),
                                      fingerprint=fp_r)
    path = os.path.join(HERE, "reference_tiny.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k, v in out.items():
        if "output" in v:
            print(k, "n_out", v["output"].shape[1] - v["prompt"].shape[1])


if __name__ == "__main__":
    main()
