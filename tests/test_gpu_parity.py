"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle.

Tolerances (BASELINE.json north_star): acceptance decisions / commit indices bit-exact given identical
logits; greedy draft tokens identical except at near-ties (top-2 margin < 1e-3 relative to the logit scale
in bf16 -- we use 2 bf16 ulps of the top logit); draft logits within 2e-2 relative in bf16.
"""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

REL_TOL = 2e-2


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


# ------------------------------------------------------------------------------------------------
# raw GEMMs
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,K,mb", [(256, 128, 16), (4096, 4096, 16), (6144, 4096, 32), (1000, 512, 16),
                                    (4096, 12288, 16), (512, 20480, 16), (2048, 2048, 64), (1024, 1024, 256)])
@pytest.mark.parametrize("grid", [148, 37])
def test_gemm_skinny(N, K, mb, grid):
    dev = _cuda()
    from dflash_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(N + K + mb)
    W = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    X = torch.randn(mb, K, device=dev).to(torch.bfloat16)
    slots = lib.dflash_gemm_max_slots(N, K, grid)
    assert slots >= 1
    ws = torch.zeros(slots, mb, N, dtype=torch.float32, device=dev)
    out = torch.zeros(mb, N, dtype=torch.float32, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.dflash_gemm_skinny(_ptr(W), N, 0, N, K, _ptr(X), mb, 0, mb, mb, _ptr(ws), mb, N, _ptr(out), N,
                                      grid, 0, st))
    torch.cuda.synchronize()
    ref = X.float() @ W.float().t()
    assert (out - ref).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("N,K,mb,rows", [(4096, 4096, 256, 1024), (6144, 1024, 256, 2048), (1000, 512, 128, 384),
                                         (2048, 12288, 256, 512), (512, 2048, 16, 48)])
def test_gemm_skinny_column_groups(N, K, mb, rows):
    """rows > mb: the launch runs ceil(rows/mb) column groups over grid/groups weight ranges (batched engines)."""
    dev = _cuda()
    from dflash_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(N + K + rows)
    grid = 148
    groups = (rows + mb - 1) // mb
    W = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    X = torch.randn(groups * mb, K, device=dev).to(torch.bfloat16)
    slots = lib.dflash_gemm_max_slots(N, K, max(1, grid // groups))
    ws = torch.zeros(slots, groups * mb, N, dtype=torch.float32, device=dev)
    out = torch.zeros(rows, N, dtype=torch.float32, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.dflash_gemm_skinny(_ptr(W), N, 0, N, K, _ptr(X), groups * mb, 0, mb, rows, _ptr(ws), groups * mb, N,
                                      _ptr(out), N, grid, 0, st))
    torch.cuda.synchronize()
    ref = X[:rows].float() @ W.float().t()
    assert (out - ref).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())


def _fused_scratch(dev, N_tiles, mb, rows, grid=148):
    groups = max(1, (rows + mb - 1) // mb)
    part = torch.zeros(grid * 128 * mb, dtype=torch.float32, device=dev)
    flags = torch.zeros(groups * N_tiles, dtype=torch.int32, device=dev)
    return groups, part, flags


@pytest.mark.parametrize("I,K,mb,rows,grid", [(12288, 4096, 16, 16, 148), (9728, 2560, 16, 16, 148), (6144, 2048, 64, 64, 148),
                                              (1024, 512, 16, 9, 37), (12288, 4096, 256, 1024, 148), (14336, 4096, 128, 128, 148)])
def test_gemm_swiglu_fused_epilogue(I, K, mb, rows, grid):
    """kModeSwiglu: 64 gate + 64 up rows per tile out of the [gate; up] stack -> bf16(bf16(silu(g)) * u) (Qwen3MLP)."""
    dev = _cuda()
    from dflash_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(I + K + rows)
    groups, part, flags = _fused_scratch(dev, I // 64, mb, rows, grid)
    W = (torch.randn(2 * I, K, device=dev) * 0.05).to(torch.bfloat16)
    X = torch.randn(groups * mb, K, device=dev).to(torch.bfloat16)
    out = torch.zeros(rows, I, dtype=torch.bfloat16, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for rep in range(2):
        out.zero_()
        _lib.check(lib.dflash_gemm_swiglu(_ptr(W), I, K, _ptr(X), groups * mb, mb, rows, _ptr(out), I, _ptr(part),
                                          _ptr(flags), grid, 0, st))
        torch.cuda.synchronize()
        assert int(flags.abs().sum()) == 0
        gu = (X[:rows].float() @ W.float().t()).to(torch.bfloat16)
        g, u = gu[:, :I], gu[:, I:]
        exp = torch.nn.functional.silu(g) * u
        assert _rel_err(out, exp) < 4e-3
        diff = (out.float() - exp.float()).abs()
        assert (diff <= exp.float().abs() * 2 ** -5 + 2e-3).all(), diff.max().item()


@pytest.mark.parametrize("mb,rows", [(128, 512), (256, 256), (256, 1024)])
def test_gemm_argmax_column_groups(mb, rows):
    dev = _cuda()
    from dflash_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(rows)
    V, K = 9000 + 40, 512
    W = (torch.randn(V, K, device=dev) * 0.03).to(torch.bfloat16)
    X = torch.randn(rows, K, device=dev).to(torch.bfloat16)
    grid = 148
    cv = torch.empty(grid, rows, dtype=torch.float32, device=dev)
    ci = torch.empty(grid, rows, dtype=torch.int32, device=dev)
    logits = torch.empty(rows, V, dtype=torch.bfloat16, device=dev)
    toks = torch.empty(rows, dtype=torch.int64, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.dflash_gemm_argmax(_ptr(W), V, V, K, _ptr(X), rows, 0, mb, rows, _ptr(cv), _ptr(ci), _ptr(logits), V,
                                      _ptr(toks), grid, 0, st))
    torch.cuda.synchronize()
    ref = X.float() @ W.float().t()
    assert _rel_err(logits, ref) < 1e-2
    assert toks.cpu().tolist() == logits.float().cpu().argmax(-1).tolist()


def test_gemm_sample_epilogue_distribution_and_greedy_limit():
    """Gumbel-max in the lm_head epilogue: tokens ~ softmax(bf16 logits / T) (chi-square over 6400 draws), a different
    draw per step / seed, and the argmax in the T -> 0 limit."""
    dev = _cuda()
    from dflash_b200 import _lib
    from dflash_b200.engine import _declare
    lib = _lib.load()
    _declare(lib)
    torch.manual_seed(3)
    V, K, mb = 300, 128, 16  # three tiles, the last one partial
    W = (torch.randn(V, K, device=dev) * 0.25).to(torch.bfloat16)
    x = torch.randn(1, K, device=dev).to(torch.bfloat16)
    X = x.repeat(mb, 1).contiguous()  # sixteen identical rows: sixteen draws from one distribution per launch
    logits = (x.float() @ W.float().t()).to(torch.bfloat16).float()[0]
    T = 0.8
    p = torch.softmax(logits / T, -1).cpu()
    grid = 148
    cv = torch.empty(grid, mb, dtype=torch.float32, device=dev)
    ci = torch.empty(grid, mb, dtype=torch.int32, device=dev)
    toks = torch.empty(mb, dtype=torch.int64, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    counts = torch.zeros(V)
    draws = []
    for step in range(400):
        _lib.check(lib.dflash_gemm_sample(_ptr(W), V, V, K, _ptr(X), mb, 0, mb, mb, T, 1234, step, _ptr(cv), _ptr(ci),
                                          _ptr(toks), grid, 0, st))
        t = toks.cpu()
        draws.append(t.tolist())
        counts += torch.bincount(t, minlength=V).float()
    assert int(counts.sum()) == 6400 and counts[V:].sum() == 0
    assert len({tuple(d) for d in draws}) > 350          # fresh noise per step
    assert len(set(draws[0])) > 1                          # and per activation row
    # chi-square with the tail merged so that every expected count is >= 8
    order = torch.argsort(p, descending=True)
    exp, obs, k = [], [], 0
    while k < V and p[order[k]] * 6400 >= 8:
        exp.append(float(p[order[k]] * 6400)); obs.append(float(counts[order[k]])); k += 1
    exp.append(float(p[order[k:]].sum() * 6400)); obs.append(float(counts[order[k:]].sum()))
    chi2 = sum((o - e) ** 2 / e for o, e in zip(obs, exp))
    dof = len(exp) - 1
    assert dof >= 10 and chi2 < dof + 5 * (2 * dof) ** 0.5, (chi2, dof)
    # another seed: a different stream; T -> 0: the argmax
    _lib.check(lib.dflash_gemm_sample(_ptr(W), V, V, K, _ptr(X), mb, 0, mb, mb, T, 99, 0, _ptr(cv), _ptr(ci), _ptr(toks),
                                      grid, 0, st))
    assert toks.cpu().tolist() != draws[0]
    _lib.check(lib.dflash_gemm_sample(_ptr(W), V, V, K, _ptr(X), mb, 0, mb, mb, 1e-3, 1, 0, _ptr(cv), _ptr(ci),
                                      _ptr(toks), grid, 0, st))
    top2 = torch.topk(logits, 2).values
    if (top2[0] - top2[1]).item() > 0.05:
        assert toks.cpu().tolist() == [int(logits.argmax())] * mb


def test_gemm_sample_epilogue_two_streams_width():
    """The 32-row instantiation (two request streams): 20 valid rows, each row its own distribution; T -> 0 gives
    every row's argmax, T = 1 gives tokens that vary with the step and stay inside the vocabulary."""
    dev = _cuda()
    from dflash_b200 import _lib
    from dflash_b200.engine import _declare
    lib = _lib.load()
    _declare(lib)
    torch.manual_seed(5)
    V, K, mb, rows = 1000, 256, 32, 20
    W = (torch.randn(V, K, device=dev) * 0.3).to(torch.bfloat16)
    X = torch.randn(mb, K, device=dev).to(torch.bfloat16)
    logits = (X[:rows].float() @ W.float().t()).to(torch.bfloat16).float()
    cv = torch.empty(148, mb, dtype=torch.float32, device=dev)
    ci = torch.empty(148, mb, dtype=torch.int32, device=dev)
    toks = torch.full((mb,), -1, dtype=torch.int64, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.dflash_gemm_sample(_ptr(W), V, V, K, _ptr(X), mb, 0, mb, rows, 1e-3, 1, 0, _ptr(cv), _ptr(ci),
                                      _ptr(toks), 148, 0, st))
    got = toks[:rows].cpu()
    ref = logits.argmax(-1).cpu()
    top2 = torch.topk(logits, 2, dim=-1).values.cpu()
    clear = (top2[:, 0] - top2[:, 1]) > 0.05
    assert torch.equal(got[clear], ref[clear]) and int(clear.sum()) >= rows // 2
    seen = set()
    for step in range(8):
        _lib.check(lib.dflash_gemm_sample(_ptr(W), V, V, K, _ptr(X), mb, 0, mb, rows, 1.0, 9, step, _ptr(cv), _ptr(ci),
                                          _ptr(toks), 148, 0, st))
        t = toks[:rows].cpu().tolist()
        assert all(0 <= x < V for x in t)
        seen.add(tuple(t))
    assert len(seen) == 8


def test_gemm_argmax_matches_own_logits():
    dev = _cuda()
    from dflash_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(0)
    V, K = 151936 // 8 + 40, 1024  # not a multiple of 128
    W = (torch.randn(V, K, device=dev) * 0.03).to(torch.bfloat16)
    X = torch.randn(16, K, device=dev).to(torch.bfloat16)
    grid = 148
    cv = torch.empty(grid, 16, dtype=torch.float32, device=dev)
    ci = torch.empty(grid, 16, dtype=torch.int32, device=dev)
    logits = torch.empty(16, V, dtype=torch.bfloat16, device=dev)
    toks = torch.empty(16, dtype=torch.int64, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.dflash_gemm_argmax(_ptr(W), V, V, K, _ptr(X), 16, 0, 16, 16, _ptr(cv), _ptr(ci), _ptr(logits), V,
                                      _ptr(toks), grid, 0, st))
    torch.cuda.synchronize()
    ref = X.float() @ W.float().t()
    assert _rel_err(logits, ref) < 1e-2
    # lowest-index argmax of the bf16 logits (torch CPU argmax semantics)
    assert toks.cpu().tolist() == logits.float().cpu().argmax(-1).tolist()


# ------------------------------------------------------------------------------------------------
# sampler: greedy ties, Gumbel race with supplied noise, Philox distribution
# ------------------------------------------------------------------------------------------------
def test_sample_greedy_and_noise_parity():
    dev = _cuda()
    from dflash_b200.utils import sample
    torch.manual_seed(1)
    V = 151936
    logits = torch.randn(2, 16, V, device=dev).to(torch.bfloat16)
    logits[0, 3, 77] = 9.0
    logits[0, 3, 5000] = 9.0  # exact tie -> lowest index
    got = sample(logits, 0.0)
    assert got.cpu().tolist() == logits.float().cpu().argmax(-1).tolist()
    assert got[0, 3].item() == 77
    # temperature: same exponential race as torch.multinomial (argmax(p / q), q ~ Exp(1))
    q = torch.empty(32, V, device=dev, dtype=torch.float32).exponential_(1.0)
    got = sample(logits, 0.7, noise=q)
    probs = torch.softmax(logits.float().view(-1, V) / 0.7, dim=-1)
    ref = (probs / q).argmax(-1)
    assert (got.view(-1) == ref).float().mean().item() >= 31 / 32  # fp32 rounding of p/q vs log-domain race


def test_sample_philox_distribution():
    dev = _cuda()
    from dflash_b200.utils import sample
    V = 8
    base = torch.tensor([2.0, 1.0, 0.0, -1.0, 0.5, -3.0, 1.5, 0.0])
    rows = 4096
    logits = base.to(torch.bfloat16).to(dev).repeat(1, rows, 1)
    p = torch.softmax(base.to(torch.bfloat16).float() / 1.3, -1)
    counts = torch.zeros(V)
    for seed in range(4):
        t = sample(logits, 1.3, seed=1000 + seed)
        counts += torch.bincount(t.view(-1).cpu(), minlength=V).float()
    n = counts.sum()
    chi2 = (((counts - n * p) ** 2) / (n * p)).sum().item()
    assert chi2 < 30.0, (chi2, counts.tolist(), (n * p).tolist())  # 7 dof: p(chi2 > 30) ~ 1e-4


# ------------------------------------------------------------------------------------------------
# engine vs oracle on the tiny pair: teacher-forced replay of an oracle trace
# ------------------------------------------------------------------------------------------------
def _tiny(bs, rigged=False, device="cuda:0"):
    from dflash_b200 import DFlashDraftModel
    from tests.tiny_models import build_pair, rig_lm_head
    from tests.golden.make_golden import LIVE
    target, draft = build_pair(DFlashDraftModel, seed=1234, block_size=bs, dtype=torch.bfloat16, device=device)
    if rigged:
        rig_lm_head(target, live=LIVE, seed=99)
    return target, draft


def _near_tie(logits_row, tok_a, tok_b):
    """True if the two candidate tokens' oracle logits are within 2 bf16 ulps of the top logit."""
    la, lb = logits_row[tok_a].float().item(), logits_row[tok_b].float().item()
    top = logits_row.float().abs().max().item()
    return abs(la - lb) <= 2 * top * 2.0 ** -8 + 1e-6


@pytest.mark.parametrize("bs", [16, 8])
@pytest.mark.parametrize("forced", [None, (3, 0, 7, 15, 1, 5, 2, 11)])
@pytest.mark.parametrize("inject", [False, True])
def test_engine_replays_oracle_trace(bs, forced, inject):
    dev = _cuda()
    from oracle import dflash_oracle as O
    from dflash_b200.engine import DraftEngine
    from tests.tiny_models import TINY, draft_state_dict
    target, draft = _tiny(bs)
    cfg = O.DraftConfig.from_hf(draft)
    sd = draft_state_dict(draft)
    prompt = torch.randint(0, TINY["vocab"] - 1, (1, 37), generator=torch.Generator().manual_seed(7)).to(dev)
    n_new = 80
    trace = []
    out_ref, taus = O.spec_generate(sd, cfg, target, prompt, n_new, None, 0.0, forced_k=forced, trace=trace)
    P = prompt.shape[1]
    eng = DraftEngine(draft, target.model.embed_tokens.weight, target.lm_head.weight, max_seq=P + n_new + 3 * bs,
                      out_len=P + n_new + 2 * bs, max_requests=1, block_size=bs, keep_draft_logits=True, use_pdl=True)
    pre = trace[0]
    eng.reset_request(0, prompt[0], pre["first_token"], n_new)
    eng.prefill_context(0, pre["hidden_sel"])
    forced_t = None if forced is None else torch.tensor([list(forced)], dtype=torch.int32, device=dev)
    worst = 0.0
    flips = 0
    flip_margins = []
    for cyc, tr in enumerate(trace[1:]):
        torch.cuda.synchronize()
        assert int(eng.buf["start"][0]) == tr["start"]
        assert eng.block_ids[0, 0].item() == tr["block"][0]
        eng.draft_step()
        torch.cuda.synchronize()
        # final-normed draft hidden and draft logits within 2e-2 relative (bf16)
        hn = eng.hn[:bs]
        e_h = _rel_err(hn, tr["draft_hidden"])
        dl = eng.buf["draft_logits"].view(eng.R * eng.SL, eng.vocab)[1:bs]
        e_l = _rel_err(dl, tr["draft_logits"])
        worst = max(worst, e_h, e_l)
        assert e_h < REL_TOL and e_l < REL_TOL, (cyc, e_h, e_l)
        # the fused argmax is exact on the engine's own bf16 logits (lowest index on ties) ...
        got = eng.block_ids[0].cpu().tolist()
        assert got[1:] == dl.float().cpu().argmax(-1).tolist()
        # ... so a token can differ from the oracle's only where the oracle's own margin between the two
        # candidates is inside the logit tolerance (a near-tie at bf16 resolution)
        for i in range(1, bs):
            if got[i] != tr["block"][i]:
                ref_row = tr["draft_logits"][i - 1].float()
                margin = (ref_row[tr["block"][i]] - ref_row[got[i]]).item()
                err = (dl[i - 1].float() - ref_row).abs().max().item()
                assert 0 <= margin <= 2 * err + 1e-6, (cyc, i, got[i], tr["block"][i], margin, err)
                assert margin <= REL_TOL * ref_row.abs().max().item(), (cyc, i, margin)
                flips += 1
                flip_margins.append(round(margin / ref_row.abs().max().item(), 5))
        # teacher forcing: the target saw the oracle's block, so verify with exactly that block
        eng.block_ids[0].copy_(torch.tensor(tr["block"], device=dev))
        eng.verify_step(tr["target_logits"].contiguous(), [h.contiguous() for h in tr["hidden_sel"]],
                        temperature=0.0, forced_k=forced_t, inject=inject)
        torch.cuda.synchronize()
        # acceptance / commit / rollback indices are integer work: bit-exact
        assert eng.posterior[0].cpu().tolist() == tr["posterior"]
        assert int(eng.buf["ctx_len"][0]) == tr["tau"]
        assert int(eng.buf["start"][0]) == tr["start"] + tr["tau"]
        assert int(eng.acc_hist[0, cyc]) == tr["tau"]
        nxt = tr["ctx_feat"]  # ctx features consumed by THIS cycle came from the previous verify
        if cyc > 0:
            pass
        if not inject:
            feat = eng.buf["ctx_feat"].view(eng.R * eng.SL, -1)[:tr["tau"]]
            exp = torch.cat(tr["hidden_sel"], dim=-1)[:tr["tau"]]
            assert torch.equal(feat, exp)
    n_cyc = len(trace) - 1
    assert eng.acc_hist[0, :n_cyc].cpu().tolist() == taus
    final = eng.output_ids[0, :P + n_new].cpu().tolist()
    exp_final = out_ref[0].cpu().tolist()
    assert [t for t in final if t != cfg.mask_token_id][:len(exp_final)] == exp_final
    if forced is not None:
        assert max(taus) == bs
    print(f"bs={bs} forced={forced is not None}: cycles={n_cyc} worst rel err={worst:.4f} near-tie flips={flips}/{n_cyc * (bs - 1)} relative margins={flip_margins}")
    eng.close()


def test_forward_dropin_matches_oracle():
    dev = _cuda()
    from oracle import dflash_oracle as O
    from dflash_b200 import DFlashStaticCache
    from tests.tiny_models import TINY, draft_state_dict
    bs = 16
    target, draft = _tiny(bs)
    cfg = O.DraftConfig.from_hf(draft)
    sd = draft_state_dict(draft)
    g = torch.Generator().manual_seed(11)
    H, nsel = TINY["hidden"], len(draft.target_layer_ids)
    th_old = torch.randn(1, 5, nsel * H, generator=g).to(dev, torch.bfloat16)
    th_new = torch.randn(1, 3, nsel * H, generator=g).to(dev, torch.bfloat16)
    noise = torch.randn(1, bs, H, generator=g).to(dev, torch.bfloat16)
    oc = O.DraftCache()
    r0 = O.draft_forward(sd, cfg, th_old, noise, torch.arange(0, 5 + bs, device=dev).unsqueeze(0), oc)
    oc.crop(5)
    r1 = O.draft_forward(sd, cfg, th_new, noise, torch.arange(5, 8 + bs, device=dev).unsqueeze(0), oc)
    cache = DFlashStaticCache()
    h0 = draft(target_hidden=th_old, noise_embedding=noise, position_ids=torch.arange(0, 5 + bs, device=dev).unsqueeze(0),
               past_key_values=cache, use_cache=True, is_causal=False)
    assert cache.get_seq_length() == 5 + bs
    cache.crop(5)
    h1 = draft(target_hidden=th_new, noise_embedding=noise,
               position_ids=torch.arange(5, 8 + bs, device=dev).unsqueeze(0), past_key_values=cache, use_cache=True,
               is_causal=False)
    assert h0.shape == r0.shape and h1.shape == r1.shape
    assert _rel_err(h0, r0) < REL_TOL and _rel_err(h1, r1) < REL_TOL, (_rel_err(h0, r0), _rel_err(h1, r1))
    draft.release_engine()


@pytest.mark.parametrize("c0,c1", [(300, 41), (16, 270)])
def test_forward_long_context_continues_cache(c0, c1):
    """forward() with more context rows than one block (the prompt pass, M = c GEMMs in 256-row passes), first from
    position 0 and then continuing a cropped cache at position c0 (benchmark.py:122-129 shapes)."""
    dev = _cuda()
    from oracle import dflash_oracle as O
    from dflash_b200 import DFlashStaticCache
    from tests.tiny_models import TINY, draft_state_dict
    bs = 16
    target, draft = _tiny(bs)
    cfg = O.DraftConfig.from_hf(draft)
    sd = draft_state_dict(draft)
    g = torch.Generator().manual_seed(c0 + c1)
    H, nsel = TINY["hidden"], len(draft.target_layer_ids)
    th0 = torch.randn(1, c0, nsel * H, generator=g).to(dev, torch.bfloat16)
    th1 = torch.randn(1, c1, nsel * H, generator=g).to(dev, torch.bfloat16)
    noise = torch.randn(1, bs, H, generator=g).to(dev, torch.bfloat16)
    oc = O.DraftCache()
    r0 = O.draft_forward(sd, cfg, th0, noise, torch.arange(0, c0 + bs, device=dev).unsqueeze(0), oc)
    oc.crop(c0)
    r1 = O.draft_forward(sd, cfg, th1, noise, torch.arange(c0, c0 + c1 + bs, device=dev).unsqueeze(0), oc)
    cache = DFlashStaticCache()
    h0 = draft(target_hidden=th0, noise_embedding=noise, position_ids=torch.arange(0, c0 + bs, device=dev).unsqueeze(0),
               past_key_values=cache, use_cache=True, is_causal=False)
    cache.crop(c0)
    h1 = draft(target_hidden=th1, noise_embedding=noise,
               position_ids=torch.arange(c0, c0 + c1 + bs, device=dev).unsqueeze(0), past_key_values=cache,
               use_cache=True, is_causal=False)
    assert cache.get_seq_length() == c0 + c1 + bs
    assert _rel_err(h0, r0) < REL_TOL and _rel_err(h1, r1) < REL_TOL, (_rel_err(h0, r0), _rel_err(h1, r1))
    draft.release_engine()


@pytest.mark.parametrize("temperature", [0.0, 1.0])
def test_spec_generate_dropin_is_lossless(temperature):
    """Product spec_generate end to end with the HF target. Greedy: every generated token must be the target's
    own greedy choice given the committed prefix (checked in one teacher-forced pass, near-ties excused)."""
    dev = _cuda()
    from tests.tiny_models import TINY
    bs = 16
    target, draft = _tiny(bs)
    prompt = torch.randint(0, TINY["vocab"] - 1, (1, 21), generator=torch.Generator().manual_seed(3)).to(dev)
    out = draft.spec_generate(target, prompt, max_new_tokens=40, stop_token_ids=None, temperature=temperature, seed=5)
    assert out.shape == (1, 21 + 40) and out.dtype == torch.int64
    assert torch.equal(out[:, :21], prompt)
    assert sum(draft.last_acceptance_lengths) >= 40
    if temperature == 0.0:
        with torch.inference_mode():
            logits = target(out).logits[0].float()
        pred = logits.argmax(-1)
        for i in range(21 - 1, out.shape[1] - 1):
            tok = out[0, i + 1].item()
            if pred[i].item() != tok:
                assert _near_tie(logits[i], pred[i].item(), tok), (i, pred[i].item(), tok)
    # stop token: generation ends right after its first occurrence
    stop = [int(out[0, 21 + 7])]
    out2 = draft.spec_generate(target, prompt, max_new_tokens=40, stop_token_ids=stop, temperature=0.0)
    gen = out2[0, 21:].tolist()
    if temperature == 0.0:
        assert gen[-1] == stop[0] and stop[0] not in gen[:-1] and len(gen) <= 8
    # tail clamp (benchmark.py:104-105) gives the same greedy tokens
    out3 = draft.spec_generate(target, prompt, max_new_tokens=40, stop_token_ids=None, temperature=0.0, clamp_tail=True)
    assert out3.shape == (1, 61)
    with pytest.raises(RuntimeError):
        draft.spec_generate(target, torch.cat([prompt, prompt]), 8, None, 0.0)
    draft.release_engine()


# ------------------------------------------------------------------------------------------------
# two request streams in one engine (ragged acceptance), block sizes 16 and 32, custom rope table
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bs,R,rope", [(16, 2, "default"), (32, 1, "default"), (8, 2, "scaled"), (16, 4, "default"),
                                        (16, 8, "default"), (32, 4, "default"), (16, 16, "default"),
                                        (8, 32, "scaled"), (16, 64, "default"), (32, 16, "default"),
                                        (8, 64, "default"), (32, 64, "default")])
@pytest.mark.parametrize("inject", [False, True])
def test_engine_batched_ragged_vs_oracle(bs, R, rope, inject):
    """inject: the next cycle's context injection runs behind the verify kernel and reads the hidden states in place
    (dflash_verify_inject_step + dflash_draft_step_injected) instead of at the head of the draft step from the features
    the verify kernel gathers."""
    dev = _cuda()
    from oracle import dflash_oracle as O
    from dflash_b200.engine import DraftEngine
    from tests.tiny_models import TINY, draft_state_dict
    target, draft = _tiny(bs)
    if rope == "scaled":  # llama3-style: a non-default inv_freq table and attention_scaling reach the kernels as data
        inv = draft.rotary_emb.inv_freq.clone()
        inv[32:] = inv[32:] / 8.0
        draft.rotary_emb.inv_freq.copy_(inv)
        draft.rotary_emb.attention_scaling = 0.75
    cfg = O.DraftConfig.from_hf(draft)
    sd = draft_state_dict(draft)
    H, V, nsel = TINY["hidden"], TINY["vocab"], len(draft.target_layer_ids)
    g = torch.Generator(device=dev).manual_seed(5)
    P = [[23, 40, 17, 31, 52, 9, 44, 28][r % 8] + r // 8 for r in range(R)]
    n_new = max(64, 5 * bs + 8)  # five cycles never reach max_length (a finished request is frozen on the device)
    eng = DraftEngine(draft, target.model.embed_tokens.weight, target.lm_head.weight, max_seq=max(P) + n_new + 3 * bs,
                      out_len=max(P) + n_new + 2 * bs, max_requests=R, block_size=bs, keep_draft_logits=True)
    caches = [O.DraftCache() for _ in range(R)]
    starts = list(P)
    pend = []  # pending ctx features per request [c, nsel*H]
    first = [3 + r for r in range(R)]
    for r in range(R):
        hs = [(torch.randn(P[r], H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(nsel)]
        prompt = torch.randint(0, V - 1, (P[r],), device=dev, generator=g)
        eng.reset_request(r, prompt, first[r], n_new)
        eng.prefill_context(r, hs)
        pend.append(torch.cat(hs, dim=-1))
    blocks = [torch.tensor([first[r]] + [cfg.mask_token_id] * (bs - 1), device=dev) for r in range(R)]
    base = [[3, 0, bs - 1, 1, 5], [0, bs - 1, 2, 7, 1]]
    sched = [[(k + 3 * (r // 2)) % bs for k in base[r % 2]] for r in range(R)]
    forced = torch.tensor([[min(k, bs - 1) for k in sched[r]] for r in range(R)], dtype=torch.int32, device=dev)
    for cyc in range(5):
        eng.draft_step()
        torch.cuda.synchronize()
        tl = torch.randn(R * bs, V, device=dev, generator=g).to(torch.bfloat16)
        hsel = [(torch.randn(R * bs, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(nsel)]
        ref_blocks = []
        for r in range(R):
            c = pend[r].shape[0]
            pos = torch.arange(caches[r].get_seq_length(), starts[r] + bs, device=dev).unsqueeze(0)
            assert caches[r].get_seq_length() + c == starts[r]
            noise = target.model.embed_tokens(blocks[r].unsqueeze(0))
            hid = O.draft_forward(sd, cfg, pend[r].unsqueeze(0), noise, pos, caches[r])
            caches[r].crop(starts[r])
            got = eng.hn[r * eng.SL: r * eng.SL + bs]
            assert _rel_err(got, hid[0]) < REL_TOL, (cyc, r, _rel_err(got, hid[0]))
            dl = eng.buf["draft_logits"].view(R * eng.SL, V)[r * eng.SL + 1: r * eng.SL + bs]
            assert eng.block_ids[r, 1:].cpu().tolist() == dl.float().cpu().argmax(-1).tolist()
            ref_blocks.append(eng.block_ids[r].clone())  # continue from the engine's own drafted tokens
        eng.verify_step(tl, hsel, temperature=0.0, forced_k=forced, inject=inject)
        torch.cuda.synchronize()
        for r in range(R):
            post = tl[r * bs:(r + 1) * bs].float().cpu().argmax(-1)
            k = int(forced[r, cyc])
            blk = ref_blocks[r].cpu()
            post[:k] = blk[1:k + 1]
            a = O.acceptance_length(blk.tolist(), post.tolist())
            assert eng.posterior[r].cpu().tolist() == post.tolist()
            assert int(eng.acc_hist[r, cyc]) == a + 1
            out = eng.output_ids[r].cpu()
            assert out[starts[r]:starts[r] + a + 1].tolist() == blk[:a + 1].tolist()
            assert int(out[starts[r] + a + 1]) == int(post[a])
            starts[r] += a + 1
            assert int(eng.buf["start"][r]) == starts[r] and int(eng.buf["ctx_len"][r]) == a + 1
            pend[r] = torch.cat([h[r * bs: r * bs + a + 1] for h in hsel], dim=-1)
            if not inject:  # (the injecting form never materialises the concatenated features)
                feat = eng.buf["ctx_feat"].view(R * eng.SL, -1)[r * eng.SL: r * eng.SL + a + 1]
                assert torch.equal(feat, pend[r])
            blocks[r] = torch.tensor([int(post[a])] + [cfg.mask_token_id] * (bs - 1), device=dev)
            assert eng.block_ids[r].cpu().tolist() == blocks[r].cpu().tolist()
    eng.close()


# ------------------------------------------------------------------------------------------------
# BASELINE config 3 shape: 16 request streams, temperature 1.0 (posterior sampled, draft greedy: SURVEY F2)
# ------------------------------------------------------------------------------------------------
def test_engine_verify_temperature_batch16():
    dev = _cuda()
    from oracle import dflash_oracle as O
    from dflash_b200.engine import DraftEngine
    from tests.tiny_models import TINY
    bs, R, T = 16, 16, 1.0
    target, draft = _tiny(bs, rigged=True)
    H, V, nsel = TINY["hidden"], TINY["vocab"], len(draft.target_layer_ids)
    g = torch.Generator(device=dev).manual_seed(21)
    eng = DraftEngine(draft, target.model.embed_tokens.weight, target.lm_head.weight, max_seq=256, out_len=256,
                      max_requests=R, block_size=bs)
    P = [11 + 3 * r for r in range(R)]
    for r in range(R):
        hs = [(torch.randn(P[r], H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(nsel)]
        eng.reset_request(r, torch.randint(0, V - 1, (P[r],), device=dev, generator=g), 3, 100)
        eng.prefill_context(r, hs)
    starts = list(P)
    for cyc in range(3):
        eng.draft_step()
        blocks = eng.block_ids.clone().cpu()
        # a posterior that agrees with the draft often: peaked at the drafted token for the first slots
        tl = torch.randn(R * bs, V, device=dev, generator=g)
        for r in range(R):
            for i in range((r + cyc) % bs):
                tl[r * bs + i, int(blocks[r, i + 1])] += 14.0
        tl = tl.to(torch.bfloat16)
        hsel = [(torch.randn(R * bs, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(nsel)]
        q = torch.empty(R * bs, V, device=dev, dtype=torch.float32).exponential_(1.0, generator=g)
        eng.verify_step(tl, hsel, temperature=T, noise=q)
        torch.cuda.synchronize()
        ref = (torch.softmax(tl.float() / T, dim=-1) / q).argmax(-1).view(R, bs).cpu()  # torch.multinomial's race
        got = eng.posterior.cpu()
        assert (got == ref).float().mean().item() >= 0.99  # fp32 rounding of p/q vs the log-domain race
        for r in range(R):
            a = O.acceptance_length(blocks[r].tolist(), got[r].tolist())  # decisions bit-exact given the posterior
            assert int(eng.acc_hist[r, cyc]) == a + 1
            out = eng.output_ids[r].cpu()
            assert out[starts[r]:starts[r] + a + 1].tolist() == blocks[r, :a + 1].tolist()
            assert int(out[starts[r] + a + 1]) == int(got[r, a])
            starts[r] += a + 1
            assert int(eng.buf["start"][r]) == starts[r] and int(eng.buf["ctx_len"][r]) == a + 1
        assert max(int(eng.acc_hist[r, cyc]) for r in range(R)) > 4
    eng.close()


# ------------------------------------------------------------------------------------------------
# batched public API: many prompts through one engine, slots refilled as requests finish
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_prompts,R,temperature,rigged", [(5, 4, 0.0, False), (16, 16, 0.0, True), (6, 2, 1.0, True),
                                                            (9, 8, 0.0, True)])
def test_spec_generate_batch_lossless_with_refill(n_prompts, R, temperature, rigged):
    dev = _cuda()
    from tests.tiny_models import TINY
    bs = 16
    target, draft = _tiny(bs, rigged=rigged)
    g = torch.Generator().manual_seed(n_prompts)
    lens = [9 + (7 * i) % 30 for i in range(n_prompts)]
    prompts = [torch.randint(0, TINY["vocab"] - 1, (1, n), generator=g).to(dev) for n in lens]
    n_new = 24
    outs = draft.spec_generate_batch(target, prompts, n_new, None, temperature, max_requests=R, seed=7)
    assert len(outs) == n_prompts
    for i, (p, o) in enumerate(zip(prompts, outs)):
        assert o.shape == (1, lens[i] + n_new) and o.dtype == torch.int64
        assert torch.equal(o[:, :lens[i]], p)
        assert sum(draft.last_batch_acceptance_lengths[i]) >= n_new
        if temperature == 0.0:  # lossless: every token is the target's own greedy choice given the prefix
            with torch.inference_mode():
                logits = target(o).logits[0].float()
            pred = logits.argmax(-1)
            for t in range(lens[i] - 1, o.shape[1] - 1):
                tok = o[0, t + 1].item()
                if pred[t].item() != tok:
                    assert _near_tie(logits[t], pred[t].item(), tok), (i, t, pred[t].item(), tok)
    if temperature == 0.0:
        # same prompts one at a time through spec_generate: identical acceptance bookkeeping per request
        o1 = draft.spec_generate(target, prompts[1], n_new, None, 0.0)
        assert o1.shape == outs[1].shape
        # stop token, ragged: every generation ends right after its first stop token
        stop = [int(outs[0][0, lens[0] + 5])]
        outs2 = draft.spec_generate_batch(target, prompts, n_new, stop, 0.0, max_requests=R)
        gen0 = outs2[0][0, lens[0]:].tolist()
        assert gen0[-1] == stop[0] and stop[0] not in gen0[:-1] and len(gen0) <= 6
        for i, o in enumerate(outs2):
            gen = o[0, lens[i]:].tolist()
            assert stop[0] not in gen[:-1]
            ref_gen = outs[i][0, lens[i]:].tolist()
            assert gen == ref_gen[:len(gen)]  # a prefix of the unconstrained generation
        # forced-acceptance harness hook per prompt: first cycle k = 3 -> tau >= 4
        draft.spec_generate_batch(target, prompts, n_new, None, 0.0, max_requests=R, forced_k=[[3, 0, 7, 1]] * n_prompts)
        for i in range(n_prompts):
            assert draft.last_batch_acceptance_lengths[i][0] >= 4
    draft.release_engine()


@pytest.mark.parametrize("sync_every", [1, 3])
def test_spec_generate_batch_with_graphed_targets(sync_every):
    """Batched loop with ONE target verify forward per cycle for all streams (ragged static cache, per-row positions and a
    4-D mask from the device-side start[r]), replayed from a CUDA graph, host polling every `sync_every` cycles, slots
    refilled: still lossless per prompt, and exactly one target forward per cycle whatever the stream count."""
    dev = _cuda()
    from tests.tiny_models import TINY
    target, draft = _tiny(16, rigged=True)
    g = torch.Generator().manual_seed(17)
    lens = [12, 31, 9, 22, 17, 40]
    prompts = [torch.randint(0, TINY["vocab"] - 1, (1, n), generator=g).to(dev) for n in lens]
    n_new = 20
    outs = draft.spec_generate_batch(target, prompts, n_new, None, 0.0, max_requests=4, graph_target=True,
                                     sync_every=sync_every)
    assert draft.last_batch_target_forwards == draft.last_batch_cycles
    assert draft.last_batch_cycles < sum(len(t) for t in draft.last_batch_acceptance_lengths)  # streams share cycles
    eager = draft.spec_generate_batch(target, prompts, n_new, None, 0.0, max_requests=4, graph_target=False)
    assert draft.last_batch_target_forwards is None  # per-request eager target calls, exactly the reference's
    for res in (outs, eager):
        for i, (p, o) in enumerate(zip(prompts, res)):
            assert o.shape == (1, lens[i] + n_new) and torch.equal(o[:, :lens[i]], p)
            with torch.inference_mode():
                logits = target(o).logits[0].float()
            pred = logits.argmax(-1)
            for t in range(lens[i] - 1, o.shape[1] - 1):
                tok = o[0, t + 1].item()
                if pred[t].item() != tok:
                    assert _near_tie(logits[t], pred[t].item(), tok), (i, t, pred[t].item(), tok)
    draft.release_engine()


def test_dflash_generate_twin_of_benchmark_loop():
    """benchmark.py:44-272 twin: same record, tail clamp, block_size == 1 baseline, --collect-profile spans."""
    dev = _cuda()
    from dflash_b200 import dflash_generate
    from tests.tiny_models import TINY
    bs = 16
    target, draft = _tiny(bs, rigged=True)
    prompt = torch.randint(0, TINY["vocab"] - 1, (1, 27), generator=torch.Generator().manual_seed(4)).to(dev)
    res = dflash_generate(draft, target, prompt, draft.mask_token_id, 40, bs, None, 0.0, collect_profile=True)
    ref = draft.spec_generate(target, prompt, 40, None, 0.0, clamp_tail=True)
    assert torch.equal(res.output_ids, ref)
    assert res.num_input_tokens == 27 and res.num_output_tokens == 40
    assert res.acceptance_lengths == draft.last_acceptance_lengths and sum(res.acceptance_lengths) >= 40
    ps = res.profile_summary
    assert set(ps) == {"target_prefill_s", "target_decode_s", "draft_decode_s", "cycle_decode_s_sum", "decode_wall_s",
                       "profiled_cycles", "draft_share_decode", "target_share_decode"}
    assert ps["profiled_cycles"] == len(res.cycle_trace) == len(res.acceptance_lengths)
    assert abs(ps["draft_share_decode"] + ps["target_share_decode"] - 1.0) < 1e-6
    for row in res.cycle_trace:
        assert row["cycle_s"] >= row["target_s"] > 0 and row["tau"] <= row["effective_block_size"]
    assert res.cycle_trace[-1]["effective_block_size"] <= bs
    # block_size 1 = the autoregressive baseline: the target's own greedy decode
    base = dflash_generate(draft, target, prompt, draft.mask_token_id, 12, 1, None, 0.0)
    assert base.acceptance_lengths == [1] * 12 and base.output_ids.shape == (1, 39)
    with torch.inference_mode():
        logits = target(base.output_ids).logits[0].float()
    pred = logits.argmax(-1)
    for t in range(26, 38):
        tok = base.output_ids[0, t + 1].item()
        assert pred[t].item() == tok or _near_tie(logits[t], pred[t].item(), tok)
    # the speculative output is the same continuation (lossless), modulo near-ties
    n = min(base.output_ids.shape[1], res.output_ids.shape[1])
    diff = (base.output_ids[0, :n] != res.output_ids[0, :n]).nonzero()
    if diff.numel():
        t = int(diff[0]) - 1
        assert _near_tie(logits[t], base.output_ids[0, t + 1].item(), res.output_ids[0, t + 1].item())
    with pytest.raises(NotImplementedError):
        dflash_generate(draft, target, prompt, draft.mask_token_id, 8, bs, None, 0.0, draft_steps=2)
    draft.release_engine()


def test_draft_step_sampled_in_engine():
    """dflash_draft_step_sampled: temperature 0 is the greedy step; at temperature 1 the drafted tokens are valid vocab
    ids that change from cycle to cycle (the noise is keyed by the device-side cycle counter)."""
    dev = _cuda()
    from dflash_b200.engine import DraftEngine
    from tests.tiny_models import TINY
    bs = 16
    target, draft = _tiny(bs)
    H, V, nsel = TINY["hidden"], TINY["vocab"], len(draft.target_layer_ids)
    g = torch.Generator(device=dev).manual_seed(2)
    eng = DraftEngine(draft, target.model.embed_tokens.weight, target.lm_head.weight, max_seq=256, out_len=256,
                      max_requests=1, block_size=bs)
    hs = [(torch.randn(20, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(nsel)]
    eng.reset_request(0, torch.randint(0, V - 1, (20,), device=dev, generator=g), 4, 100)
    eng.prefill_context(0, hs)
    blk0 = eng.block_ids.clone()
    eng.draft_step()
    greedy = eng.block_ids.clone()
    eng.block_ids.copy_(blk0)
    eng.draft_step_sampled(0.0)
    assert torch.equal(eng.block_ids, greedy)
    seen = set()
    for cyc in range(6):
        eng.block_ids.copy_(blk0)
        eng.draft_step_sampled(1.0, seed=5)
        toks = eng.block_ids[0].cpu().tolist()
        assert toks[0] == 4 and all(0 <= t < V for t in toks[1:])
        seen.add(tuple(toks))
        eng.buf["rng_step"] += 1  # what the accept kernel does once per cycle
    assert len(seen) == 6 and tuple(greedy[0].cpu().tolist()) not in seen
    eng.close()


def test_dflash_generate_with_block_size_scheduler():
    """the block size is chosen per cycle by a host policy; every choice is a device-side blk_len.
    Output stays the target's greedy continuation for any schedule (verification is lossless)."""
    dev = _cuda()
    from dflash_b200 import EwmaBlockScheduler, dflash_generate
    from tests.tiny_models import TINY
    target, draft = _tiny(16, rigged=True)
    prompt = torch.randint(0, TINY["vocab"] - 1, (1, 19), generator=torch.Generator().manual_seed(8)).to(dev)
    sched = EwmaBlockScheduler([4, 8, 16, 32], warmup_cycles=8, probe_interval=5, required_streak=1, cooldown_cycles=0)
    res = dflash_generate(draft, target, prompt, draft.mask_token_id, 96, 32, None, 0.0, collect_profile=True,
                          scheduler=sched)
    assert res.num_output_tokens == 96
    ebs = [row["effective_block_size"] for row in res.cycle_trace]
    assert ebs[:4] == [4, 8, 16, 32] and len(set(ebs)) >= 3
    for row in res.cycle_trace:
        assert 1 <= row["tau"] <= row["effective_block_size"]
    # every cycle that ran a candidate size fed the estimates (clamped tail blocks are ignored)
    assert sum(sched.n_obs.values()) == sum(1 for b in ebs if b in (4, 8, 16, 32))
    with torch.inference_mode():
        logits = target(res.output_ids).logits[0].float()
    pred = logits.argmax(-1)
    for t in range(18, res.output_ids.shape[1] - 1):
        tok = res.output_ids[0, t + 1].item()
        if pred[t].item() != tok:
            assert _near_tie(logits[t], pred[t].item(), tok), (t, pred[t].item(), tok)
    with pytest.raises(ValueError):
        dflash_generate(draft, target, prompt, draft.mask_token_id, 8, 16, None, 0.0, scheduler=sched)
    # sampled drafts (the policy loop's `sample(draft_logits, temperature)`): verification keeps the output the target's
    res2 = dflash_generate(draft, target, prompt, draft.mask_token_id, 40, 16, None, 0.0, draft_temperature=1.0, seed=3)
    assert res2.num_output_tokens == 40
    n = min(res2.output_ids.shape[1], res.output_ids.shape[1])
    diff = (res2.output_ids[0, :n] != res.output_ids[0, :n]).nonzero()
    if diff.numel():
        t = int(diff[0]) - 1
        assert _near_tie(logits[t], res2.output_ids[0, t + 1].item(), res.output_ids[0, t + 1].item())
    draft.release_engine()


# ------------------------------------------------------------------------------------------------
# multi-candidate drafting / verify ("fixed_prefix_rank", benchmark_candidate_solutions.py)
# ------------------------------------------------------------------------------------------------
def _cand_engine(bs, K=4):
    from dflash_b200.engine import DraftEngine
    from tests.tiny_models import TINY
    dev = _cuda()
    target, draft = _tiny(bs, rigged=False)
    H, V, nsel = TINY["hidden"], TINY["vocab"], len(draft.target_layer_ids)
    g = torch.Generator(device=dev).manual_seed(31)
    mk = lambda keep: DraftEngine(draft, target.model.embed_tokens.weight, target.lm_head.weight, max_seq=256,  # noqa: E731
                                  out_len=256, max_requests=1, block_size=bs, keep_draft_logits=keep, max_candidates=K)
    return dev, target, draft, H, V, nsel, g, mk


@pytest.mark.parametrize("bs,prefix", [(16, 2), (16, 5), (8, 0), (32, 3)])
def test_topk_epilogue_and_candidate_blocks_vs_oracle(bs, prefix):
    from oracle import dflash_oracle as O
    dev, target, draft, H, V, nsel, g, mk = _cand_engine(bs)
    eng = mk(True)
    hs = [(torch.randn(21, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(nsel)]
    eng.reset_request(0, torch.randint(0, V - 1, (21,), device=dev, generator=g), 5, 100)
    eng.prefill_context(0, hs)
    blk0 = eng.block_ids.clone()          # [committed token, mask, mask, ...]: the draft step's input
    eng.draft_step()                      # argmax epilogue + bf16 logits dump
    torch.cuda.synchronize()
    logits = eng.buf["draft_logits"].view(eng.SL, V)[:bs].clone()
    greedy = eng.block_ids.clone()
    eng.block_ids.copy_(blk0)
    eng.draft_step_candidates(4, prefix)  # same input state: top-4 epilogue + candidate blocks
    torch.cuda.synchronize()
    assert torch.equal(eng.block_ids, greedy)
    tv = eng.buf["topk_val"].view(eng.SL, 4)[:bs]
    ti = eng.buf["topk_idx"].view(eng.SL, 4)[:bs].long()
    ref_v, ref_i = torch.topk(logits.float(), 5, dim=-1)
    assert torch.equal(tv, ref_v[:, :4])                                   # the top-4 VALUES of the engine's own logits
    assert torch.equal(torch.gather(logits.float(), 1, ti), tv)            # each index carries its value
    assert all(len(set(row)) == 4 for row in ti.cpu().tolist())            # four distinct tokens per row
    clean = [len(set(ref_v[i].tolist())) == 5 for i in range(bs)]          # rows without a tie among the top 5
    for i in range(bs):
        if clean[i]:
            assert ti[i].cpu().tolist() == ref_i[i, :4].cpu().tolist()
    # candidate blocks and scores against the reference's builder (restated in the oracle, pinned by goldens)
    base = greedy.cpu()
    cands, scores = O.fixed_prefix_rank_candidates(base, logits[1:].unsqueeze(0).cpu(), prefix, 4, 4)
    got = eng.cand_ids[0].cpu()
    assert got.shape == (4, bs)
    suffix_start = max(1, min(prefix, bs))
    for k in range(4):
        for t in range(bs):
            if t < suffix_start or clean[t]:  # (torch.topk's order among exactly tied logits is unspecified)
                assert int(got[k, t]) == int(cands[k, t]), (k, t)
            if k == 0:
                assert int(got[0, t]) == int(base[0, t])  # candidate 0 is the greedy block (ties -> lowest index)
    sc = eng.cand_scores[0].cpu().tolist()
    for k in range(4):
        assert abs(sc[k] - scores[k]) <= 2.0 ** -7 * max(1.0, abs(scores[k])), (k, sc, scores)
    eng.close()


@pytest.mark.parametrize("R", [1, 2])
def test_verify_candidates_choice_commit_and_gather(R):
    """K candidates per request verified by one (synthetic) target forward: chosen index, commit, posterior, state and
    the context rows gathered from the chosen candidate, against the oracle's restatement of the reference's choice
    rule. R = 2: two request streams, rows ordered [request][candidate][slot]."""
    from oracle import dflash_oracle as O
    from dflash_b200.engine import DraftEngine
    bs, K = 16, 4
    dev, target, draft, H, V, nsel, g, mk = _cand_engine(bs)
    eng = DraftEngine(draft, target.model.embed_tokens.weight, target.lm_head.weight, max_seq=256, out_len=256,
                      max_requests=R, block_size=bs, max_candidates=K)
    starts = [17 + 6 * r for r in range(R)]
    for r in range(R):
        hs = [(torch.randn(starts[r], H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(nsel)]
        eng.reset_request(r, torch.randint(0, V - 1, (starts[r],), device=dev, generator=g), 5 + r, 200)
        eng.prefill_context(r, hs)
    plans = [[3, 9, 1, 9], [0, 0, 0, 0], [2, 5, 14, 7], [15, 15, 3, 3], [6, 2, 2, 11]]
    for cyc in range(len(plans)):
        eng.draft_step_candidates(K, 2)
        torch.cuda.synchronize()
        cands = eng.cand_ids.clone().cpu()          # [R, 4, bs]
        scores = eng.cand_scores.cpu().tolist()
        force = [plans[(cyc + r) % len(plans)] for r in range(R)]
        # target logits whose argmax agrees with candidate k of request r on exactly force[r][k] positions
        tl = torch.randn(R * K * bs, V, device=dev, generator=g)
        for r in range(R):
            for k in range(K):
                for i in range(bs):
                    nxt = int(cands[r, k, i + 1]) if i + 1 < bs else 0
                    tok = nxt if i < force[r][k] else (nxt + 1 + i) % V
                    tl[(r * K + k) * bs + i, tok] += 20.0
        tl = tl.to(torch.bfloat16)
        hsel = [(torch.randn(R * K * bs, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(nsel)]
        eng.verify_step_candidates(K, tl, hsel)
        torch.cuda.synchronize()
        post_all = tl.float().argmax(-1).view(R, K, bs).cpu()
        for r in range(R):
            chosen, acc = O.choose_candidate(cands[r], post_all[r], scores[r])
            assert acc == [min(f, bs - 1) for f in force[r]]
            assert int(eng.buf["chosen"][r]) == chosen, (cyc, r, acc, scores[r])
            a = acc[chosen]
            assert eng.posterior[r].cpu().tolist() == post_all[r, chosen].tolist()
            out = eng.output_ids[r].cpu()
            assert out[starts[r]:starts[r] + a + 1].tolist() == cands[r, chosen, :a + 1].tolist()
            assert int(out[starts[r] + a + 1]) == int(post_all[r, chosen, a])
            starts[r] += a + 1
            assert int(eng.buf["start"][r]) == starts[r] and int(eng.buf["ctx_len"][r]) == a + 1
            feat = eng.buf["ctx_feat"].view(R * eng.SL, -1)[r * eng.SL: r * eng.SL + a + 1]
            row0 = (r * K + chosen) * bs
            assert torch.equal(feat, torch.cat([h[row0: row0 + a + 1] for h in hsel], dim=-1))
            assert eng.block_ids[r].cpu().tolist() == [int(post_all[r, chosen, a])] + [draft.mask_token_id] * (bs - 1)
    eng.close()


@pytest.mark.parametrize("graph_target", [False, True])
def test_dflash_generate_candidates_is_lossless(graph_target):
    """The whole multi-candidate loop with the HF target (one batched verify forward per cycle, the chosen branch of the
    target cache kept): greedy output is still the target's own greedy continuation. graph_target: the batch-K forward
    replays from a CUDA graph over a batch-K static cache; the winning branch is kept by a device-side row copy."""
    dev = _cuda()
    from dflash_b200 import dflash_generate, dflash_generate_candidates
    from tests.tiny_models import TINY
    bs = 16
    target, draft = _tiny(bs, rigged=True)
    prompt = torch.randint(0, TINY["vocab"] - 1, (1, 23), generator=torch.Generator().manual_seed(6)).to(dev)
    res = dflash_generate_candidates(draft, target, prompt, draft.mask_token_id, 48, bs, None, fixed_prefix_len=2,
                                     rank_top_k=4, max_candidates=4, graph_target=graph_target,
                                     sync_every=2 if graph_target else 1)
    assert res.output_ids.shape == (1, 23 + 48) and sum(res.acceptance_lengths) >= 48
    assert res.candidate_summary["avg_candidates_per_cycle"] == 4.0
    with torch.inference_mode():
        logits = target(res.output_ids).logits[0].float()
    pred = logits.argmax(-1)
    for t in range(22, res.output_ids.shape[1] - 1):
        tok = res.output_ids[0, t + 1].item()
        if pred[t].item() != tok:
            assert _near_tie(logits[t], pred[t].item(), tok), (t, pred[t].item(), tok)
    plain = dflash_generate(draft, target, prompt, draft.mask_token_id, 48, bs, None, 0.0)
    # more candidates never commit fewer tokens per cycle on the same prefix: compare cycle counts loosely
    assert len(res.acceptance_lengths) <= len(plain.acceptance_lengths) + 2
    with pytest.raises(ValueError):
        dflash_generate_candidates(draft, target, prompt, draft.mask_token_id, 8, bs, None, temperature=1.0)
    draft.release_engine()


def test_verify_with_given_posterior_and_stop_and_clamp():
    """Integer path only: posterior tokens supplied by the caller (e.g. sampled elsewhere at T>0), stop ids,
    tail clamp. Exhaustive over the acceptance length."""
    dev = _cuda()
    from oracle import dflash_oracle as O
    from dflash_b200.engine import DraftEngine
    bs = 16
    target, draft = _tiny(bs)
    H, nsel = 256, 2
    eng = DraftEngine(draft, target.model.embed_tokens.weight, target.lm_head.weight, max_seq=512, out_len=512,
                      max_requests=1, block_size=bs)
    hsel = [torch.zeros(bs, H, dtype=torch.bfloat16, device=dev) for _ in range(nsel)]
    stop = torch.tensor([777], dtype=torch.int64, device=dev)
    for a in range(bs):
        for stop_at in (None, a):
            eng.reset_request(0, torch.arange(10, device=dev), 5, 300)
            blk = torch.arange(100, 100 + bs, device=dev)
            if stop_at is not None:
                blk[stop_at] = 777
            eng.block_ids[0].copy_(blk)
            post = torch.cat([blk[1:], torch.tensor([42], device=dev)]).clone()
            if a < bs - 1:
                post[a] = 999
            eng.verify_step(None, hsel, posterior_in=post.view(1, bs).contiguous(), stop_ids=stop)
            torch.cuda.synchronize()
            out = [0] * 400
            ns, tau = O.verify_commit(out, 10, blk.tolist(), post.tolist())
            assert int(eng.buf["start"][0]) == ns and int(eng.buf["ctx_len"][0]) == tau == a + 1
            assert eng.output_ids[0, 10:ns + 1].cpu().tolist() == out[10:ns + 1]
            assert int(eng.buf["done"][0]) == (1 if stop_at is not None else 0)
    # tail clamp: only 5 tokens left -> effective block 5 -> at most 4 accepted + bonus
    eng.reset_request(0, torch.arange(10, device=dev), 5, 5)
    eng.buf["blk_len"][0] = 5
    blk = torch.arange(100, 100 + bs, device=dev)
    eng.block_ids[0].copy_(blk)
    post = torch.cat([blk[1:], torch.tensor([42], device=dev)])
    eng.verify_step(None, hsel, posterior_in=post.view(1, bs).contiguous(), clamp_tail=True)
    torch.cuda.synchronize()
    assert int(eng.buf["ctx_len"][0]) == 5 and int(eng.buf["start"][0]) == 15 and int(eng.buf["done"][0]) == 1
    eng.close()


# ------------------------------------------------------------------------------------------------
# hidden sizes above 4096 (Qwen3-14B / 32B width): the row pass takes its four-CTA cluster form there, the context
# injection's row pass and the embedding rows their widest loops
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hidden,R,inject", [(5120, 1, True), (5120, 2, False), (8192, 1, True)])
def test_wide_hidden_draft_vs_oracle(hidden, R, inject):
    dev = _cuda()
    from oracle import dflash_oracle as O
    from dflash_b200 import DFlashDraftModel
    from dflash_b200.engine import DraftEngine
    from tests.tiny_models import TINY, build_pair, draft_state_dict
    bs = 16
    dims = dict(TINY, hidden=hidden, intermediate=1024, heads=4, kv_heads=2, vocab=1200)
    target, draft = build_pair(DFlashDraftModel, seed=77, block_size=bs, dims=dims, dtype=torch.bfloat16, device=dev)
    cfg = O.DraftConfig.from_hf(draft)
    sd = draft_state_dict(draft)
    H, V, nsel = dims["hidden"], dims["vocab"], len(draft.target_layer_ids)
    g = torch.Generator(device=dev).manual_seed(3)
    P = [19, 33][:R]
    n_new = 4 * bs
    eng = DraftEngine(draft, target.model.embed_tokens.weight, target.lm_head.weight, max_seq=max(P) + n_new + 3 * bs,
                      out_len=max(P) + n_new + 2 * bs, max_requests=R, block_size=bs, keep_draft_logits=True)
    caches = [O.DraftCache() for _ in range(R)]
    starts, pend, first = list(P), [], [5 + r for r in range(R)]
    for r in range(R):
        hs = [(torch.randn(P[r], H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(nsel)]
        eng.reset_request(r, torch.randint(0, V - 1, (P[r],), device=dev, generator=g), first[r], n_new)
        eng.prefill_context(r, hs)
        pend.append(torch.cat(hs, dim=-1))
    blocks = [torch.tensor([first[r]] + [cfg.mask_token_id] * (bs - 1), device=dev) for r in range(R)]
    forced = torch.tensor([[4, 0, bs - 1][r:] + [2] * r for r in range(R)], dtype=torch.int32, device=dev)
    for cyc in range(3):
        eng.draft_step()
        torch.cuda.synchronize()
        tl = torch.randn(R * bs, V, device=dev, generator=g).to(torch.bfloat16)
        hsel = [(torch.randn(R * bs, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(nsel)]
        for r in range(R):
            pos = torch.arange(caches[r].get_seq_length(), starts[r] + bs, device=dev).unsqueeze(0)
            noise = target.model.embed_tokens(blocks[r].unsqueeze(0))
            hid = O.draft_forward(sd, cfg, pend[r].unsqueeze(0), noise, pos, caches[r])
            caches[r].crop(starts[r])
            got = eng.hn[r * eng.SL: r * eng.SL + bs]
            assert _rel_err(got, hid[0]) < REL_TOL, (cyc, r, _rel_err(got, hid[0]))
        blk = eng.block_ids.clone()
        eng.verify_step(tl, hsel, temperature=0.0, forced_k=forced, inject=inject)
        torch.cuda.synchronize()
        for r in range(R):
            post = tl[r * bs:(r + 1) * bs].float().cpu().argmax(-1)
            k = int(forced[r, cyc])
            post[:k] = blk[r].cpu()[1:k + 1]
            a = O.acceptance_length(blk[r].cpu().tolist(), post.tolist())
            starts[r] += a + 1
            assert int(eng.buf["start"][r]) == starts[r] and int(eng.buf["ctx_len"][r]) == a + 1
            pend[r] = torch.cat([h[r * bs: r * bs + a + 1] for h in hsel], dim=-1)
            blocks[r] = torch.tensor([int(post[a])] + [cfg.mask_token_id] * (bs - 1), device=dev)
    eng.close()


# ------------------------------------------------------------------------------------------------
# the two forms of the context injection give the same draft: at the head of the draft step from gathered features, or
# behind the verify kernel reading the hidden states in place (dflash_verify_inject_step / dflash_draft_step_injected)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bs,R", [(16, 1), (8, 3), (32, 2)])
def test_injection_forms_agree(bs, R):
    dev = _cuda()
    from dflash_b200.engine import DraftEngine
    from tests.tiny_models import TINY
    target, draft = _tiny(bs)
    H, V, nsel = TINY["hidden"], TINY["vocab"], len(draft.target_layer_ids)
    g = torch.Generator(device=dev).manual_seed(11)
    P = 29
    hs0 = [(torch.randn(P, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(nsel)]
    prompt = torch.randint(0, V - 1, (P,), device=dev, generator=g)
    engs = [DraftEngine(draft, target.model.embed_tokens.weight, target.lm_head.weight, max_seq=P + 8 * bs,
                        out_len=P + 7 * bs, max_requests=R, block_size=bs, keep_draft_logits=True) for _ in range(2)]
    for e in engs:
        for r in range(R):
            e.reset_request(r, prompt, 7 + r, 5 * bs)
            e.prefill_context(r, hs0)
    forced = torch.tensor([[2, bs - 1, 0, 5][r % 4:] + [1] * (r % 4) for r in range(R)], dtype=torch.int32, device=dev)
    for cyc in range(4):
        for e in engs:
            e.draft_step()
        torch.cuda.synchronize()
        # same inputs -> same draft, whichever kernel injected the context (the block rows' layernorm sums its squares
        # over a different thread layout in the two forms: allow the last bf16 bit of the hidden state)
        a = engs[0].hn.view(R, engs[0].SL, H)[:, :bs].float()
        b = engs[1].hn.view(R, engs[1].SL, H)[:, :bs].float()
        assert _rel_err(b, a) < 4e-3, (cyc, _rel_err(b, a))
        same = (engs[0].block_ids == engs[1].block_ids).float().mean().item()
        assert same >= 0.9, (cyc, same)  # (argmax flips only at near-ties of the tiny random head)
        engs[1].block_ids.copy_(engs[0].block_ids)  # keep the integer state of the two engines comparable
        tl = torch.randn(R * bs, V, device=dev, generator=g).to(torch.bfloat16)
        hsel = [(torch.randn(R * bs, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(nsel)]
        if cyc == 2:  # the three-kernel verify (given posterior) with injection: the gathered form of the same kernel
            post = tl.float().argmax(-1).view(R, bs).contiguous()
            engs[0].verify_step(None, hsel, posterior_in=post, inject=False)
            engs[1].verify_step(None, hsel, posterior_in=post, inject=True)
        else:
            engs[0].verify_step(tl, hsel, temperature=0.0, forced_k=forced, inject=False)
            engs[1].verify_step(tl, hsel, temperature=0.0, forced_k=forced, inject=True)
        torch.cuda.synchronize()
        for name in ("start", "ctx_len", "done", "n_cycles"):
            assert torch.equal(engs[0].buf[name], engs[1].buf[name]), (cyc, name)
        assert torch.equal(engs[0].output_ids, engs[1].output_ids)
        if cyc == 1 and R > 1:  # a slot is refilled between an injecting verify step and the next draft step
            for e in engs:
                e.reset_request(R - 1, prompt, 3, 5 * bs)
                e.prefill_context(R - 1, hs0)
    for e in engs:
        e.close()


# ------------------------------------------------------------------------------------------------
# target verify forward from a CUDA graph over a static cache
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sync_every", [1, 4])
def test_graphed_target_matches_eager_target(sync_every):
    dev = _cuda()
    from tests.tiny_models import TINY
    bs = 16
    target, draft = _tiny(bs)
    prompt = torch.randint(0, TINY["vocab"] - 1, (1, 21), generator=torch.Generator().manual_seed(3)).to(dev)
    forced = [3, 0, 7, 15, 1, 5, 2, 11]
    # honest greedy: the output depends only on the target's argmax -> equal to the eager-target run except at
    # near-ties (same module math, but a different attention kernel: static length + mask)
    ref = draft.spec_generate(target, prompt, 40, None, 0.0, graph_target=False)
    out = draft.spec_generate(target, prompt, 40, None, 0.0, graph_target=True, sync_every=sync_every)
    assert out.shape == ref.shape
    if not torch.equal(out, ref):
        with torch.inference_mode():
            logits = target(ref).logits[0].float()
        i = int((out[0] != ref[0]).nonzero()[0])
        assert _near_tie(logits[i - 1], int(out[0, i]), int(ref[0, i])), (i, out[0, i].item(), ref[0, i].item())
    # forced acceptance: the cycle structure (tau per cycle) is integer work and must follow the schedule
    out = draft.spec_generate(target, prompt, 48, None, 0.0, forced_k=forced, graph_target=True, sync_every=sync_every)
    assert out.shape == (1, 21 + 48)
    taus = draft.last_acceptance_lengths
    assert all(t >= forced[i % len(forced)] + 1 for i, t in enumerate(taus[:-1])) and max(taus) == bs
    # honest greedy run must be lossless w.r.t. one teacher-forced pass of the target
    out = draft.spec_generate(target, prompt, 40, None, 0.0, graph_target=True, sync_every=sync_every)
    with torch.inference_mode():
        logits = target(out).logits[0].float()
    pred = logits.argmax(-1)
    for i in range(20, out.shape[1] - 1):
        if pred[i].item() != out[0, i + 1].item():
            assert _near_tie(logits[i], pred[i].item(), out[0, i + 1].item()), i
    # stop token
    stop = [int(out[0, 21 + 6])]
    out2 = draft.spec_generate(target, prompt, 40, stop, 0.0, graph_target=True, sync_every=sync_every)
    gen = out2[0, 21:].tolist()
    assert gen[-1] == stop[0] and stop[0] not in gen[:-1]
    draft.release_engine()


# ------------------------------------------------------------------------------------------------
# full BASELINE size (Qwen3-8B + DFlash-b16 dims): one prompt context + two cycles against the oracle
# ------------------------------------------------------------------------------------------------
def test_full_size_qwen3_8b_step_vs_oracle():
    dev = _cuda()
    import bench
    from oracle import dflash_oracle as O
    dims = bench.Q8
    H, V, L, bs = dims["hidden"], dims["vocab"], dims["draft_layers"], dims["block_size"]
    draft, eng, embed, lm_head = bench.build_engine(dims, dev, seed=3)
    cfg = O.DraftConfig.from_hf(draft)
    sd = {k: v.detach() for k, v in draft.state_dict().items()}
    g = torch.Generator(device=dev).manual_seed(11)
    P = 77
    hs = [(torch.randn(P, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(L)]
    prompt = torch.randint(0, V - 1, (P,), device=dev, generator=g)
    eng.reset_request(0, prompt, 12345, 256)
    eng.prefill_context(0, hs)
    cache = O.DraftCache()
    cache32 = O.DraftCache()
    sd32 = {k: v.float() for k, v in sd.items()}
    pend = torch.cat(hs, dim=-1).unsqueeze(0)
    start = P
    block = torch.tensor([[12345] + [cfg.mask_token_id] * (bs - 1)], device=dev)
    for cyc in range(2):
        eng.draft_step()
        torch.cuda.synchronize()
        with torch.inference_mode():
            pos = torch.arange(cache.get_seq_length(), start + bs, device=dev).unsqueeze(0)
            noise = torch.nn.functional.embedding(block, embed)
            # the reference's default dispatch is sdpa (fp32 scores inside); "eager" additionally rounds the scores
            # to bf16. Both in bf16, plus an fp32 run of the same weights as the yardstick for bf16 noise.
            O.ATTN_IMPL = "sdpa"
            c_sdpa = O.DraftCache([None if k is None else k.clone() for k in cache.keys],
                                  [None if v is None else v.clone() for v in cache.values])
            hid_sdpa = O.draft_forward(sd, cfg, pend, noise, pos, c_sdpa)
            O.ATTN_IMPL = "eager"
            c32 = O.DraftCache([None if k is None else k.float() for k in cache32.keys],
                               [None if v is None else v.float() for v in cache32.values])
            hid32 = O.draft_forward(sd32, cfg, pend.float(), noise.float(), pos, cache32)
            del c32
            cache32.crop(start)
            hid = O.draft_forward(sd, cfg, pend, noise, pos, cache)
            cache.crop(start)
            ref_logits = torch.nn.functional.linear(hid_sdpa[0, 1:], lm_head)
        e_sdpa, e_eager = _rel_err(eng.hn[:bs], hid_sdpa[0]), _rel_err(eng.hn[:bs], hid[0])
        e_32 = _rel_err(eng.hn[:bs], hid32[0])
        n_sdpa, n_eager = _rel_err(hid_sdpa[0], hid32[0]), _rel_err(hid[0], hid32[0])
        print(f"cycle {cyc}: engine vs bf16-sdpa {e_sdpa:.4f}, vs bf16-eager {e_eager:.4f}, vs fp32 {e_32:.4f}; "
              f"bf16-sdpa vs fp32 {n_sdpa:.4f}, bf16-eager vs fp32 {n_eager:.4f}")
        # within 2e-2 of the reference's sdpa path, and no further from the fp32 result than the reference's own bf16 runs
        assert e_sdpa < REL_TOL, (cyc, e_sdpa)
        assert e_32 < 1.25 * max(n_sdpa, n_eager), (cyc, e_32, n_sdpa, n_eager)
        got = eng.block_ids[0, 1:].cpu().tolist()
        ref_tok = ref_logits.float().argmax(-1).cpu().tolist()
        for i, (a, b) in enumerate(zip(got, ref_tok)):
            if a != b:  # only where the oracle's own margin is inside the bf16 logit tolerance
                row = ref_logits[i].float()
                assert (row[b] - row[a]).item() <= REL_TOL * row.abs().max().item(), (cyc, i, a, b)
        # verify with a synthetic target: accept 5 tokens, then the bonus
        tl = torch.randn(bs, V, device=dev, generator=g).to(torch.bfloat16)
        hsel = [(torch.randn(bs, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(L)]
        forced = torch.tensor([[5]], dtype=torch.int32, device=dev)
        blk = eng.block_ids[0].clone()
        eng.verify_step(tl, hsel, forced_k=forced)
        torch.cuda.synchronize()
        post = tl.float().argmax(-1)
        post[:5] = blk[1:6]
        assert eng.posterior[0].tolist() == post.tolist()
        assert int(eng.buf["ctx_len"][0]) == 6 and int(eng.buf["start"][0]) == start + 6
        start += 6
        pend = torch.cat([h[:6] for h in hsel], dim=-1).unsqueeze(0)
        block = torch.tensor([[int(post[5])] + [cfg.mask_token_id] * (bs - 1)], device=dev)
        assert eng.block_ids[0].tolist() == block[0].tolist()
    eng.close()


# Element-wise bound for full-size hidden states / logits: no single element may be further from the reference's bf16
# sdpa result than this fraction of its row's largest magnitude. (A relative L2 error of 2e-2 spread over a row as
# independent rounding noise puts the LARGEST of 151 936 element errors at ~5 sigma = 2e-2 * 5 / 4.5 of the row maximum,
# so 2e-2 in norm corresponds to ~2.3e-2 element-wise; the bound is 3e-2.)
ELEM_TOL = 3e-2


def _elem_err(a, b):
    """max over rows of max_i |a_i - b_i| / max_i |b_i|"""
    a, b = a.float(), b.float()
    return ((a - b).abs().amax(-1) / b.abs().amax(-1).clamp_min(1e-12)).max().item()


def test_full_size_qwen3_8b_long_context_elementwise():
    """BASELINE configs[1] at the END of its 2048-token generation: Qwen3-8B + DFlash-b16 dims, a 2048-row prompt pass
    (eight 256-row GEMM passes), then four full-acceptance cycles to S = 2112 -- the 16-way split-KV attention at its
    full 32/8-head dims with ~33 tiles per split. Hidden states AND bf16 draft logits against the reference's bf16 sdpa
    path: relative L2 < 2e-2, every element within ELEM_TOL of its row maximum, mutual top-1-in-top-4 agreement per
    row, fused argmax exact on the engine's own logits; and no further from an fp32 run than the reference's own bf16
    run is."""
    dev = _cuda()
    import bench
    from oracle import dflash_oracle as O
    dims = bench.Q8
    H, V, L, bs = dims["hidden"], dims["vocab"], dims["draft_layers"], dims["block_size"]
    draft, eng, embed, lm_head = bench.build_engine(dims, dev, seed=9, keep_draft_logits=True)
    cfg = O.DraftConfig.from_hf(draft)
    sd = {k: v.detach() for k, v in draft.state_dict().items()}
    sd32 = {k: v.float() for k, v in sd.items()}
    g = torch.Generator(device=dev).manual_seed(21)
    P = 2048
    hs = [(torch.randn(P, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(L)]
    prompt = torch.randint(0, V - 1, (P,), device=dev, generator=g)
    eng.reset_request(0, prompt, 4242, 128)
    eng.prefill_context(0, hs)
    cache, cache32 = O.DraftCache(), O.DraftCache()
    pend = torch.cat(hs, dim=-1).unsqueeze(0)
    del hs
    start = P
    block = torch.tensor([[4242] + [cfg.mask_token_id] * (bs - 1)], device=dev)
    forced = torch.tensor([[bs - 1]], dtype=torch.int32, device=dev)
    O.ATTN_IMPL = "sdpa"
    stats = []
    for cyc in range(4):
        eng.draft_step()
        torch.cuda.synchronize()
        with torch.inference_mode():
            pos = torch.arange(cache.get_seq_length(), start + bs, device=dev).unsqueeze(0)
            noise = torch.nn.functional.embedding(block, embed)
            hid = O.draft_forward(sd, cfg, pend, noise, pos, cache)
            cache.crop(start)
            hid32 = O.draft_forward(sd32, cfg, pend.float(), noise.float(), pos, cache32)
            cache32.crop(start)
            ref_logits = torch.nn.functional.linear(hid[0, 1:], lm_head)                 # bf16, as target.lm_head
            ref_logits32 = torch.nn.functional.linear(hid32[0, 1:], lm_head.float())
        assert cache.get_seq_length() == start
        got_h = eng.hn[:bs]
        got_l = eng.buf["draft_logits"].view(eng.R * eng.SL, V)[1:bs]
        e_h, e_l = _rel_err(got_h, hid[0]), _rel_err(got_l, ref_logits)
        m_h, m_l = _elem_err(got_h, hid[0]), _elem_err(got_l, ref_logits)
        e_h32, n_h32 = _rel_err(got_h, hid32[0]), _rel_err(hid[0], hid32[0])
        e_l32, n_l32 = _rel_err(got_l, ref_logits32), _rel_err(ref_logits, ref_logits32)
        m_l32, mn_l32 = _elem_err(got_l, ref_logits32), _elem_err(ref_logits, ref_logits32)
        top_e, top_o = got_l.float().topk(4, dim=-1).indices, ref_logits.float().topk(4, dim=-1).indices
        in_e = (top_e == top_o[:, :1]).any(-1)   # the oracle's argmax is among the engine's four best
        in_o = (top_o == top_e[:, :1]).any(-1)   # and the other way round
        same1 = (top_e[:, 0] == top_o[:, 0]).float().mean().item()
        stats.append((start, e_h, m_h, e_l, m_l, e_h32, n_h32, e_l32, n_l32, m_l32, mn_l32, same1))
        print(f"S={start + bs}: hidden rel {e_h:.4f} elem {m_h:.4f} | logits rel {e_l:.4f} elem {m_l:.4f} | vs fp32: hidden "
              f"{e_h32:.4f} (oracle bf16 {n_h32:.4f}), logits {e_l32:.4f} (oracle bf16 {n_l32:.4f}), logits elem {m_l32:.4f} "
              f"(oracle bf16 {mn_l32:.4f}) | same argmax {same1:.2f}")
        assert e_h < REL_TOL and e_l < REL_TOL, (cyc, e_h, e_l)
        assert m_h < ELEM_TOL and m_l < ELEM_TOL, (cyc, m_h, m_l)
        assert e_h32 <= 1.25 * n_h32 and e_l32 <= 1.25 * n_l32 and m_l32 <= 1.5 * mn_l32, (cyc, e_h32, n_h32, e_l32, n_l32)
        assert bool(in_e.all()) and bool(in_o.all()), (cyc, in_e.tolist(), in_o.tolist())
        assert eng.block_ids[0, 1:].tolist() == got_l.float().argmax(-1).tolist()  # fused argmax = argmax of own logits
        # verify with a synthetic target that accepts the whole block (forced), then the bonus token
        tl = torch.randn(bs, V, device=dev, generator=g).to(torch.bfloat16)
        hsel = [(torch.randn(bs, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(L)]
        blk = eng.block_ids[0].clone()
        eng.verify_step(tl, hsel, forced_k=forced)
        torch.cuda.synchronize()
        post = tl.float().argmax(-1)
        post[:bs - 1] = blk[1:]
        assert eng.posterior[0].tolist() == post.tolist()
        assert int(eng.buf["ctx_len"][0]) == bs and int(eng.buf["start"][0]) == start + bs
        start += bs
        pend = torch.cat(hsel, dim=-1).unsqueeze(0)
        block = torch.tensor([[int(post[bs - 1])] + [cfg.mask_token_id] * (bs - 1)], device=dev)
    assert start == P + 4 * bs
    eng.close()


# ------------------------------------------------------------------------------------------------
# BASELINE configs[0], [2] and [3] at their full draft dimensions (configs[0]: Qwen3-4B shape, hidden 2560, batch 1):
#   LLaMA-3.1-8B shape (I = 14336, V = 128256, llama3 rope table) with the posterior sampled at temperature 1.0;
#   Qwen3-Coder-30B-A3B shape (H = 2048, 8 draft layers / 8 selected target layers, GQA 8:1), greedy.
# ------------------------------------------------------------------------------------------------
# Relative L2 tolerance of the draft hidden vs the reference's bf16 sdpa path, per configuration. 2e-2 is the north
# star's bound; it holds for every 5-layer draft. The 8-layer Qwen3-Coder-30B-A3B draft is the documented exception: two
# bf16 runs of 8 layers that round at different points (e.g. the fp32 split-K sum here vs cuBLAS' own order) are each
# ~0.017 from the exact (fp32) result and 0.022-0.025 from each other (measured over 32 streams), so the bound there is
# 3e-2 -- and in EVERY
# configuration the CUDA path must also be no further from the fp32 result than 1.25x the reference's own bf16 run is.
FULL_SIZE_REL_TOL = {"llama31": 2e-2, "qwen3_4b": 2e-2, "coder30b": 3e-2, "q8_bs8": 2e-2, "q8_bs32": 2e-2}


@pytest.mark.parametrize("name,R,temperature", [("llama31", 16, 1.0), ("coder30b", 4, 0.0), ("qwen3_4b", 1, 0.0),
                                                ("coder30b", 32, 0.0), ("q8_bs8", 64, 0.0), ("q8_bs32", 64, 0.0)])
def test_full_size_batched_configs_vs_oracle(name, R, temperature):
    """BASELINE.json configs[0], [2], [3], [4] at their full draft dimensions and stated batch: configs[2] = 16 streams at
    temperature 1.0 (LLaMA-3.1-8B shape), configs[3] = the global batch of 32 (Qwen3-Coder-30B-A3B draft shape) as one
    engine (the 1-GPU end of its 2/4/8-GPU shards, which are the 16/8/4-stream cases of the same code), configs[4] =
    block sizes 8 and 32 at batch 64 (Qwen3-8B shape; block size 16 at batch 1 is the test above)."""
    dev = _cuda()
    import bench
    from oracle import dflash_oracle as O
    dims = {"llama31": bench.LLAMA31_8B, "coder30b": bench.QWEN3_CODER_30B_A3B, "qwen3_4b": bench.QWEN3_4B,
            "q8_bs8": dict(bench.Q8, block_size=8), "q8_bs32": dict(bench.Q8, block_size=32)}[name]
    H, V, L, bs = dims["hidden"], dims["vocab"], dims["draft_layers"], dims["block_size"]
    tol = FULL_SIZE_REL_TOL[name]
    draft, eng, embed, lm_head = bench.build_engine(dims, dev, seed=5, R=R, max_new=192)
    assert len(draft.target_layer_ids) == L
    cfg = O.DraftConfig.from_hf(draft)
    sd = {k: v.detach() for k, v in draft.state_dict().items()}
    sd32 = {k: v.float() for k, v in sd.items()}  # fp32 run of the same weights: the yardstick for bf16 noise
    g = torch.Generator(device=dev).manual_seed(13)
    P = [40 + 9 * (r % 16) + r // 16 for r in range(R)]
    first = [100 + r for r in range(R)]
    caches, caches32, pend, blocks, starts = [], [], [], [], list(P)
    for r in range(R):
        hs = [(torch.randn(P[r], H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(L)]
        eng.reset_request(r, torch.randint(0, V - 1, (P[r],), device=dev, generator=g), first[r], 100)
        eng.prefill_context(r, hs)
        caches.append(O.DraftCache())
        caches32.append(O.DraftCache())
        pend.append(torch.cat(hs, dim=-1).unsqueeze(0))
        blocks.append(torch.tensor([[first[r]] + [cfg.mask_token_id] * (bs - 1)], device=dev))
    O.ATTN_IMPL = "sdpa"
    worst, worst32 = 0.0, 0.0
    for cyc in range(2):
        eng.draft_step()
        torch.cuda.synchronize()
        for r in range(R):
            with torch.inference_mode():
                pos = torch.arange(caches[r].get_seq_length(), starts[r] + bs, device=dev).unsqueeze(0)
                noise = torch.nn.functional.embedding(blocks[r], embed)
                hid = O.draft_forward(sd, cfg, pend[r], noise, pos, caches[r])
                caches[r].crop(starts[r])
                hid32 = O.draft_forward(sd32, cfg, pend[r].float(), noise.float(), pos, caches32[r])
                caches32[r].crop(starts[r])
                ref_logits = torch.nn.functional.linear(hid[0, 1:], lm_head).float()
            got_h = eng.hn[r * eng.SL: r * eng.SL + bs]
            err, e32, n32 = _rel_err(got_h, hid[0]), _rel_err(got_h, hid32[0]), _rel_err(hid[0], hid32[0])
            worst, worst32 = max(worst, err), max(worst32, e32 / n32)
            # within the configuration's stated bound of the reference's bf16 path AND no further from the fp32 result
            # than the reference's own bf16 run is (x1.25)
            assert err < tol, (cyc, r, err, tol)
            assert e32 <= 1.25 * n32, (cyc, r, e32, n32)
            got = eng.block_ids[r, 1:].cpu().tolist()
            for i, (a, b) in enumerate(zip(got, ref_logits.argmax(-1).cpu().tolist())):
                if a != b:  # only where the oracle's own margin is inside the bf16 logit tolerance
                    assert (ref_logits[i, b] - ref_logits[i, a]).item() <= REL_TOL * ref_logits[i].abs().max().item()
        # synthetic target outputs; posterior at the config's temperature with supplied Exp(1) noise
        blk = eng.block_ids.clone().cpu()
        tl = torch.randn(R * bs, V, device=dev, generator=g)
        for r in range(R):
            for i in range((3 * r + cyc + 5) % bs):
                tl[r * bs + i, int(blk[r, i + 1])] += 16.0  # the target agrees with the first drafted tokens
        tl = tl.to(torch.bfloat16)
        hsel = [(torch.randn(R * bs, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(L)]
        q = torch.empty(R * bs, V, device=dev, dtype=torch.float32).exponential_(1.0, generator=g)
        eng.verify_step(tl, hsel, temperature=temperature, noise=q if temperature > 0 else None)
        torch.cuda.synchronize()
        if temperature > 0:
            ref_post = (torch.softmax(tl.float() / temperature, dim=-1) / q).argmax(-1).view(R, bs).cpu()
            assert (eng.posterior.cpu() == ref_post).float().mean().item() >= 0.99
        else:
            assert torch.equal(eng.posterior.cpu(), tl.float().argmax(-1).view(R, bs).cpu())
        post = eng.posterior.cpu()
        for r in range(R):
            a = O.acceptance_length(blk[r].tolist(), post[r].tolist())  # bit-exact given the posterior
            assert int(eng.acc_hist[r, cyc]) == a + 1 and int(eng.buf["start"][r]) == starts[r] + a + 1
            starts[r] += a + 1
            pend[r] = torch.cat([h[r * bs: r * bs + a + 1] for h in hsel], dim=-1).unsqueeze(0)
            assert torch.equal(eng.buf["ctx_feat"].view(R * eng.SL, -1)[r * eng.SL: r * eng.SL + a + 1], pend[r][0])
            blocks[r] = torch.tensor([[int(post[r, a])] + [cfg.mask_token_id] * (bs - 1)], device=dev)
        assert max(int(eng.acc_hist[r, cyc]) for r in range(R)) >= 4
    print(f"{name} R={R}: worst relative error of the draft hidden vs the bf16 sdpa oracle {worst:.4f} (bound {tol}); "
          f"worst (engine vs fp32) / (bf16 oracle vs fp32) {worst32:.3f}")
    eng.close()


@pytest.mark.parametrize("P", [44, 256, 300, 1030])
def test_prompt_pass_then_first_block_vs_oracle(P):
    """Cycle 0 with c = P context rows (dflash.py:229,238-246): the prompt pass (256 rows per GEMM pass, UMMA width
    picked per pass) followed by the first block, against one oracle forward over all P context rows."""
    dev = _cuda()
    from oracle import dflash_oracle as O
    from dflash_b200.engine import DraftEngine
    from tests.tiny_models import TINY, draft_state_dict
    bs = 16
    target, draft = _tiny(bs)
    cfg = O.DraftConfig.from_hf(draft)
    sd = draft_state_dict(draft)
    H, V, nsel = TINY["hidden"], TINY["vocab"], len(draft.target_layer_ids)
    g = torch.Generator(device=dev).manual_seed(P)
    eng = DraftEngine(draft, target.model.embed_tokens.weight, target.lm_head.weight, max_seq=P + 64, out_len=P + 64,
                      max_requests=1, block_size=bs, keep_draft_logits=True)
    hs = [(torch.randn(P, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(nsel)]
    eng.reset_request(0, torch.randint(0, V - 1, (P,), device=dev, generator=g), 9, 32)
    eng.prefill_context(0, hs)
    eng.draft_step()
    torch.cuda.synchronize()
    block = torch.tensor([[9] + [cfg.mask_token_id] * (bs - 1)], device=dev)
    noise = target.model.embed_tokens(block)
    hid = O.draft_forward(sd, cfg, torch.cat(hs, dim=-1).unsqueeze(0), noise,
                          torch.arange(0, P + bs, device=dev).unsqueeze(0), O.DraftCache())
    err = _rel_err(eng.hn[:bs], hid[0])
    assert err < REL_TOL, err
    eng.close()


@pytest.mark.parametrize("P,n_new", [(1, 1), (1, 2), (5, 17), (300, 33), (1030, 20)])
def test_spec_generate_edge_lengths_match_oracle(P, n_new):
    """Shortest prompt, a single new token, a generation that ends inside the first / second block, a prompt longer
    than one 256-row prompt pass and one that spans five of them: same tokens as the oracle's spec_generate (which is
    pinned to the reference), near-ties excused by the lossless check against the target."""
    dev = _cuda()
    from oracle import dflash_oracle as O
    from tests.tiny_models import TINY, draft_state_dict
    target, draft = _tiny(16, rigged=True)
    prompt = torch.randint(0, TINY["vocab"] - 1, (1, P), generator=torch.Generator().manual_seed(P)).to(dev)
    out = draft.spec_generate(target, prompt, max_new_tokens=n_new, stop_token_ids=None, temperature=0.0)
    assert out.shape == (1, P + n_new) and torch.equal(out[:, :P], prompt)
    ref, taus = O.spec_generate(draft_state_dict(draft), O.DraftConfig.from_hf(draft), target, prompt, n_new, None, 0.0)
    assert ref.shape == out.shape
    if not torch.equal(out, ref):  # a near-tie somewhere: every token must still be the target's greedy choice
        with torch.inference_mode():
            logits = target(out).logits[0].float()
        pred = logits.argmax(-1)
        for t in range(P - 1, out.shape[1] - 1):
            tok = out[0, t + 1].item()
            if pred[t].item() != tok:
                assert _near_tie(logits[t], pred[t].item(), tok), (t, pred[t].item(), tok)
    # (acceptance lengths may differ from the oracle's where the DRAFT's own logits are near-tied: with equal
    #  committed tokens that only moves tokens between cycles)
    assert sum(draft.last_acceptance_lengths) >= n_new
    draft.release_engine()


@pytest.mark.parametrize("bs", [8, 32])
def test_spec_generate_block_size_override_is_lossless(bs):
    """benchmark.py --block-size: the number of mask slots is overridden at inference (benchmark.py:104-108,419);
    greedy output stays the target's own greedy continuation for any block size."""
    dev = _cuda()
    from tests.tiny_models import TINY
    target, draft = _tiny(16)
    draft.block_size = bs
    prompt = torch.randint(0, TINY["vocab"] - 1, (1, 33), generator=torch.Generator().manual_seed(9)).to(dev)
    out = draft.spec_generate(target, prompt, max_new_tokens=30, stop_token_ids=None, temperature=0.0)
    assert out.shape == (1, 63)
    with torch.inference_mode():
        logits = target(out).logits[0].float()
    pred = logits.argmax(-1)
    for i in range(32, out.shape[1] - 1):
        if pred[i].item() != out[0, i + 1].item():
            assert _near_tie(logits[i], pred[i].item(), out[0, i + 1].item()), i
    out_f = draft.spec_generate(target, prompt, 30, None, 0.0, forced_k=[bs - 1, 0, 3])
    assert max(draft.last_acceptance_lengths) == bs and out_f.shape == (1, 63)
    draft.release_engine()


# ------------------------------------------------------------------------------------------------
# round-2 boundary items
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("graph_target", [False, None])
def test_fresh_engine_first_cycle_matches_oracle(graph_target):
    """The first cycle on a FRESH engine (graph capture warm-up included) drafts from [token, mask, mask, ...] exactly
    as the reference does (model/dflash.py:233-235): acceptance lengths per cycle equal the oracle's. graph_target=None
    is the default path (graphed target when capturable)."""
    dev = _cuda()
    from oracle import dflash_oracle as O
    from tests.tiny_models import TINY, draft_state_dict
    target, draft = _tiny(16, rigged=True)
    for seed in (11, 12):
        prompt = torch.randint(0, TINY["vocab"] - 1, (1, 23), generator=torch.Generator().manual_seed(seed)).to(dev)
        draft.release_engine()  # every prompt starts on a new engine (and a new draft graph)
        out = draft.spec_generate(target, prompt, 48, None, 0.0, graph_target=graph_target)
        ref, taus = O.spec_generate(draft_state_dict(draft), O.DraftConfig.from_hf(draft), target, prompt, 48, None, 0.0)
        # The first cycle is the one a stale block would change; later cycles may shift at a bf16 near-tie of the draft
        # (other drafted token -> other acceptance length, same committed tokens) or of the target.
        got = draft.last_acceptance_lengths
        assert got[0] == taus[0], (got, taus)
        lead = next((i for i, (a, b) in enumerate(zip(got, taus)) if a != b), min(len(got), len(taus)))
        assert lead >= 4, (got, taus)
        again = draft.spec_generate(target, prompt, 48, None, 0.0, graph_target=graph_target)
        assert torch.equal(again, out)
    draft.release_engine()


def test_same_seed_same_samples_on_cached_engine():
    """temperature > 0 with a given seed is reproducible on the cached engine (the Philox step restarts per request)."""
    dev = _cuda()
    from tests.tiny_models import TINY
    target, draft = _tiny(16, rigged=True)
    prompt = torch.randint(0, TINY["vocab"] - 1, (1, 19), generator=torch.Generator().manual_seed(5)).to(dev)
    a = draft.spec_generate(target, prompt, 40, None, 1.0, seed=1234)
    draft.spec_generate(target, prompt, 24, None, 1.0, seed=99)  # other work on the same engine in between
    b = draft.spec_generate(target, prompt, 40, None, 1.0, seed=1234)
    c = draft.spec_generate(target, prompt, 40, None, 1.0, seed=1235)
    assert torch.equal(a, b)
    assert not torch.equal(a, c)
    draft.release_engine()


def test_forward_with_callers_dynamic_cache_like_benchmark_loop():
    """benchmark.py:60,122-129 owns a DynamicCache and calls draft.forward(...) / cache.crop(start) itself: the product
    forward adopts the cache's length (the K/V rows stay in the engine) and returns what the oracle's forward returns."""
    dev = _cuda()
    from transformers import DynamicCache
    from oracle import dflash_oracle as O
    from tests.tiny_models import TINY, draft_state_dict
    bs = 16
    target, draft = _tiny(bs)
    sd, cfg = draft_state_dict(draft), O.DraftConfig.from_hf(draft)
    H, nsel = TINY["hidden"], len(draft.target_layer_ids)
    g = torch.Generator().manual_seed(3)
    cache_p, cache_o = DynamicCache(), O.DraftCache()
    start = 0
    for c in (21, 5, 16, 1):  # prompt-sized context first, then ragged accepted lengths
        th = (torch.randn(1, c, nsel * H, generator=g) * 0.5).to(dev).to(torch.bfloat16)
        noise = (torch.randn(1, bs, H, generator=g) * 0.1).to(dev).to(torch.bfloat16)
        start += c
        pos = torch.arange(cache_p.get_seq_length(), start + bs, device=dev).unsqueeze(0)
        assert cache_p.get_seq_length() == cache_o.get_seq_length()
        out = draft(target_hidden=th, noise_embedding=noise, position_ids=pos, past_key_values=cache_p, use_cache=True,
                    is_causal=False)
        ref = O.draft_forward(sd, cfg, th, noise, pos, cache_o)
        assert cache_p.get_seq_length() == start + bs
        assert _rel_err(out, ref) < REL_TOL
        cache_p.crop(start)
        cache_o.crop(start)
    other = DynamicCache()
    other.update(torch.zeros(1, 1, 3, 128, device=dev), torch.zeros(1, 1, 3, 128, device=dev), 0)
    with pytest.raises(RuntimeError):
        draft(target_hidden=th, noise_embedding=noise, position_ids=torch.arange(3, 3 + 1 + bs, device=dev).unsqueeze(0),
              past_key_values=other, use_cache=True)
    draft.release_engine()


def test_attention_bias_draft_matches_oracle():
    """config.attention_bias = True (model/dflash.py:41-50): the q/k/v biases are added in the QKV post-processing (before
    q/k-norm, as nn.Linear does), o_proj's in the row pass; prompt pass, block forward with a live cache and a whole
    greedy spec_generate against the oracle (pinned to the reference by tests/golden/reference_bias.pt)."""
    dev = _cuda()
    from oracle import dflash_oracle as O
    from dflash_b200 import DFlashDraftModel, DFlashStaticCache
    from tests.golden.make_golden import LIVE
    from tests.tiny_models import TINY, build_pair, draft_state_dict, rig_lm_head
    bs = 16
    target, draft = build_pair(DFlashDraftModel, seed=1234, block_size=bs, dtype=torch.bfloat16, device=dev,
                               attention_bias=True)
    rig_lm_head(target, live=LIVE, seed=99)
    sd, cfg = draft_state_dict(draft), O.DraftConfig.from_hf(draft)
    assert "layers.0.self_attn.q_proj.bias" in sd
    H, nsel = TINY["hidden"], len(draft.target_layer_ids)
    g = torch.Generator().manual_seed(11)
    cache_p, cache_o = DFlashStaticCache(), O.DraftCache()
    start = 0
    for c in (300, 3, 16):  # a two-pass prompt context, then block-sized contexts
        th = (torch.randn(1, c, nsel * H, generator=g) * 0.5).to(dev).to(torch.bfloat16)
        noise = (torch.randn(1, bs, H, generator=g) * 0.1).to(dev).to(torch.bfloat16)
        start += c
        pos = torch.arange(cache_p.get_seq_length(), start + bs, device=dev).unsqueeze(0)
        out = draft(target_hidden=th, noise_embedding=noise, position_ids=pos, past_key_values=cache_p, use_cache=True,
                    is_causal=False)
        ref = O.draft_forward(sd, cfg, th, noise, pos, cache_o)
        assert _rel_err(out, ref) < REL_TOL, (c, _rel_err(out, ref))
        # the biases matter: the same forward without them is far outside the tolerance
        cache_p.crop(start)
        cache_o.crop(start)
    sd_nb = {k: v for k, v in sd.items() if not k.endswith(".bias")}
    ref_nb = O.draft_forward(sd_nb, cfg, th, noise, torch.arange(0, c + bs, device=dev).unsqueeze(0), O.DraftCache())
    ref_b = O.draft_forward(sd, cfg, th, noise, torch.arange(0, c + bs, device=dev).unsqueeze(0), O.DraftCache())
    assert _rel_err(ref_nb, ref_b) > 5 * REL_TOL
    draft.release_engine()
    prompt = torch.randint(0, TINY["vocab"] - 1, (1, 12), generator=torch.Generator().manual_seed(7)).to(dev)
    out = draft.spec_generate(target, prompt, 48, None, 0.0)
    ref, taus = O.spec_generate(sd, cfg, target, prompt, 48, None, 0.0)
    if torch.equal(out, ref):
        assert draft.last_acceptance_lengths == taus
    else:  # a target near-tie shifts the cycles; the first cycle cannot move
        assert draft.last_acceptance_lengths[0] == taus[0]
    with torch.inference_mode():
        logits = target(out).logits[0].float()
    pred = logits.argmax(-1)
    for i in range(11, out.shape[1] - 1):
        tok = out[0, i + 1].item()
        if pred[i].item() != tok:
            assert _near_tie(logits[i], pred[i].item(), tok), (i, pred[i].item(), tok)
    draft.release_engine()
