"""The oracle (oracle/dflash_oracle.py) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py, run in the build container). CPU only."""
import pytest
import torch

from oracle import dflash_oracle as O
from tests.tiny_models import build_pair, draft_state_dict, fingerprint, rig_lm_head
from tests.golden.make_golden import LIVE


def _pair(bs, rigged):
    from dflash_b200 import DFlashDraftModel  # product class: same parameter names as the reference
    target, draft = build_pair(DFlashDraftModel, seed=1234, block_size=bs)
    if rigged:
        rig_lm_head(target, live=LIVE, seed=99)
    return target, draft


@pytest.mark.parametrize("bs", [16, 8])
@pytest.mark.parametrize("case", ["honest", "rigged", "stop", "sampled", "forced"])
def test_spec_generate_matches_reference(golden, bs, case):
    g = golden[f"bs{bs}/{case}"]
    target, draft = _pair(bs, rigged=(case != "honest"))
    assert fingerprint(target, draft) == g["fingerprint"], "seeded weights drifted from the golden run"
    cfg = O.DraftConfig.from_hf(draft)
    sd = draft_state_dict(draft)
    n_new = {"honest": 24, "rigged": 48, "stop": 48, "sampled": 32, "forced": 64}[case]
    temp = 1.0 if case == "sampled" else 0.0
    if case == "sampled":
        torch.manual_seed(g["seed"])
    trace = []
    out, taus = O.spec_generate(sd, cfg, target, g["prompt"], n_new, g.get("stop"), temp, trace=trace,
                                forced_k=g.get("forced"))
    assert out.tolist() == g["output"].tolist()
    assert sum(taus) >= out.shape[1] - g["prompt"].shape[1] - 1
    if case == "honest":
        assert all(t == 1 for t in taus)  # random weights never agree
    if case == "forced":
        ks = [min(k, bs - 1) for k in g["forced"]]
        assert all(t >= ks[i % len(ks)] + 1 for i, t in enumerate(taus[:-1]))
        assert max(taus) == bs
    if case == "rigged":
        assert out.tolist() == g["autoregressive"].tolist()  # lossless greedy speculative decoding


@pytest.mark.parametrize("bs", [16, 8])
def test_draft_forward_matches_reference(golden, bs):
    g = golden[f"bs{bs}/forward"]
    target, draft = _pair(bs, rigged=True)
    cfg = O.DraftConfig.from_hf(draft)
    sd = draft_state_dict(draft)
    cache = O.DraftCache()
    pos = torch.arange(0, 5 + bs).unsqueeze(0)
    h0 = O.draft_forward(sd, cfg, g["th_old"], g["noise"], pos, cache)
    assert cache.get_seq_length() == 5 + bs
    cache.crop(5)
    pos = torch.arange(5, 8 + bs).unsqueeze(0)
    h1 = O.draft_forward(sd, cfg, g["th_new"], g["noise"], pos, cache)
    torch.testing.assert_close(h0, g["h0"], rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(h1, g["h1"], rtol=2e-4, atol=2e-5)


def test_attention_bias_matches_reference():
    """config.attention_bias = True (model/dflash.py:41-50): biased q/k/v/o projections, forward with a live cache and a
    whole greedy spec_generate, against vectors from the unmodified reference (tests/golden/make_bias_golden.py)."""
    import os
    from dflash_b200 import DFlashDraftModel
    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_bias.pt"))
    bs = 16
    target, draft = build_pair(DFlashDraftModel, seed=1234, block_size=bs, attention_bias=True)
    rig_lm_head(target, live=LIVE, seed=99)
    assert fingerprint(target, draft) == g["fingerprint"], "seeded weights drifted from the golden run"
    cfg, sd = O.DraftConfig.from_hf(draft), draft_state_dict(draft)
    assert "layers.0.self_attn.q_proj.bias" in sd and "layers.1.self_attn.o_proj.bias" in sd
    cache = O.DraftCache()
    h0 = O.draft_forward(sd, cfg, g["th_old"], g["noise"], torch.arange(0, 5 + bs).unsqueeze(0), cache)
    cache.crop(5)
    h1 = O.draft_forward(sd, cfg, g["th_new"], g["noise"], torch.arange(5, 8 + bs).unsqueeze(0), cache)
    torch.testing.assert_close(h0, g["h0"], rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(h1, g["h1"], rtol=2e-4, atol=2e-5)
    out, taus = O.spec_generate(sd, cfg, target, g["prompt"], 48, None, 0.0)
    assert out.tolist() == g["output"].tolist() == g["autoregressive"].tolist()


def test_layer_ids_and_helpers():
    assert O.build_target_layer_ids(36, 5) == [1, 9, 17, 25, 33]
    assert O.build_target_layer_ids(36, 1) == [18]
    assert O.build_target_layer_ids(6, 2) == [1, 3]
    from dflash_b200.utils import build_target_layer_ids
    for lt, ld in [(36, 5), (32, 5), (48, 8), (6, 2), (28, 1)]:
        assert build_target_layer_ids(lt, ld) == O.build_target_layer_ids(lt, ld)


def test_acceptance_exhaustive():
    bs = 16
    blk = list(range(100, 100 + bs))
    for a in range(bs):
        post = [blk[i + 1] if i < a else -1 for i in range(bs - 1)] + [7]
        if a < bs - 1:
            post[a] = 999
        out = [0] * 64
        new_start, tau = O.verify_commit(out, 5, blk, post)
        assert tau == a + 1 and new_start == 5 + a + 1
        assert out[5:5 + a + 1] == blk[:a + 1] and out[5 + a + 1] == post[a]
        ref = (torch.tensor(blk[1:]) == torch.tensor(post[:-1])).cumprod(0).sum().item()
        assert ref == a


def test_candidate_builder_and_choice_match_reference():
    """the oracle's fixed_prefix_rank candidates / candidate choice against vectors generated from the
    reference's own function and expressions (tests/golden/make_cand_golden.py)."""
    import os
    import torch
    from oracle import dflash_oracle as O
    cases = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "candidates.pt"))
    assert len(cases) >= 10
    for c in cases:
        cands, scores = O.fixed_prefix_rank_candidates(c["base"], c["logits"], c["prefix"], c["topk"], c["maxc"])
        assert torch.equal(cands, c["cands"])
        assert scores == c["scores"]
        chosen, acc = O.choose_candidate(cands, c["posterior"], scores)
        assert acc == c["acc"] and chosen == c["chosen"]
