"""Batched speculative decoding: many prompts share one draft engine (BASELINE configs 3-5).

The reference decodes one prompt at a time; its "batched" harness (`benchmark_batched.py`) only keeps several
batch-1 decodes in flight. Here up to 64 request streams live in ONE engine: every cycle the draft step, the
posterior sampling, the acceptance / commit and the context gather run once for all of them (one weight stream,
ragged acceptance lengths kept as device state). The target stays the caller's unmodified HF module; by default its
verify forward (`model/dflash.py:249-255`) also runs ONCE per cycle for all streams (`BatchedVerifyTarget`: per-row
positions, a 4-D mask and a ragged static KV cache derived from the engine's device-side `start[r]`, replayed from a
CUDA graph), `graph_target=False` calls it per request exactly as the reference does. A finished request's slot is
refilled with the next prompt (its prompt pass does not disturb the streams that are mid-generation).

Results per prompt obey the same contract as `spec_generate`: `LongTensor[1, P + n_out]`, ends at the first stop
token, mask ids removed; greedy outputs are the target's own greedy continuation.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
from transformers import DynamicCache

from .engine import DraftEngine
from .utils import ContextTap, sample


@torch.inference_mode()
def spec_generate_batch(draft, target, prompts: Sequence[torch.Tensor], max_new_tokens: int,
                        stop_token_ids: Optional[List[int]], temperature: float, *, max_requests: Optional[int] = None,
                        clamp_tail: bool = False, forced_k: Optional[Sequence[Sequence[int]]] = None,
                        seed: Optional[int] = None, noise_fn: Optional[Callable[[int], torch.Tensor]] = None,
                        graph_target="auto", sync_every: int = 1) -> List[torch.Tensor]:
    """prompts: LongTensor[1, P_i] each (ragged). Returns one LongTensor[1, P_i + n_i] per prompt, in order.

    max_requests: request streams resident in the engine (<= 64; default: enough for all prompts).
    forced_k[i]: harness hook, per-prompt forced-acceptance schedule. noise_fn(cycle) -> fp32
    [R * block_size, V] Exp(1) draws for the posterior race at temperature > 0 (tests); otherwise Philox(seed).
    graph_target ("auto" | True | False): unless False, the (unmodified) target's verify forward runs once per cycle for
    ALL streams over a ragged static KV cache whose per-stream lengths are the engine's device-side `start[r]`
    (`target_graph.BatchedVerifyTarget`), replayed from a CUDA graph when the target can be captured; the host then
    only polls `start` / `done` every `sync_every` cycles (finished streams are frozen on the device meanwhile).
    "auto" falls back to the per-request eager calls for targets the batched forward does not cover (sliding-window
    layers); False always uses them. `draft.last_batch_target_forwards` = verify forwards issued.
    Side effect: `draft.last_batch_acceptance_lengths[i]` = tau per cycle of prompt i."""
    draft.eval()
    n = len(prompts)
    if n == 0:
        return []
    dev = target.device
    bs = draft.block_size
    R = min(n, max_requests or 64)
    if R > 64:
        raise ValueError("at most 64 request streams per engine")
    Pmax = max(int(p.shape[1]) for p in prompts)
    for p in prompts:
        if p.dim() != 2 or p.shape[0] != 1:
            raise RuntimeError("spec_generate_batch: every prompt is a LongTensor[1, P]")
    max_len_all = Pmax + max_new_tokens
    eng = _batch_engine(draft, target, R, bs, max_len_all, dev)
    return _run(draft, target, eng, list(prompts), max_new_tokens, stop_token_ids, temperature, clamp_tail,
                forced_k, seed, noise_fn, graph_target, max(1, int(sync_every)), max_len_all + 2 * bs)


def _batch_engine(draft, target, R: int, bs: int, max_len_all: int, dev) -> DraftEngine:
    """The engine of the previous call is reused when it fits (same target weights, stream count, block size, and at
    least this capacity): a serving loop calls spec_generate_batch many times. `draft.release_engine()` drops it."""
    key = (target.model.embed_tokens.weight.data_ptr(), target.lm_head.weight.data_ptr(), R, bs, str(dev))
    cached = getattr(draft, "_batch_engine_cache", None)
    if cached is not None:
        k, eng, _ = cached
        if k == key and eng.max_seq >= max_len_all + 2 * bs + 1 and eng.handle.value:
            return eng
        eng.close()
        draft._batch_engine_cache = None
    cap = max(1024, 1 << (max_len_all + 2 * bs).bit_length())
    eng = DraftEngine(draft, target.model.embed_tokens.weight, target.lm_head.weight, max_seq=cap + 1, out_len=cap + 1,
                      max_requests=R, block_size=bs, device=dev)
    draft._batch_engine_cache = (key, eng, {})
    return eng


def _run(draft, target, eng: DraftEngine, prompts, max_new_tokens, stop_token_ids, temperature, clamp_tail, forced_k,
         seed, noise_fn, graph_target=False, sync_every=1, cache_len=0):
    dev, bs, R, n = eng.device, eng.block_size, eng.R, len(prompts)
    H, V, nsel = eng.hidden, eng.vocab, eng.n_sel
    layer_ids = draft.target_layer_ids
    if seed is None:
        seed = int(torch.randint(0, 2**62, (1,)).item()) if temperature >= 1e-5 else 0
    stop_t = None
    if stop_token_ids is not None and len(stop_token_ids) > 0:
        stop_t = torch.tensor(list(stop_token_ids), dtype=torch.int64, device=dev)
    fk_len = max((len(f) for f in forced_k), default=0) if forced_k is not None else 0
    forced_t = torch.zeros(R, max(1, fk_len), dtype=torch.int32, device=dev) if forced_k is not None else None

    # the verify step's inputs for all R streams: each request's target outputs land in its rows
    tl = torch.zeros(R * bs, V, dtype=torch.bfloat16, device=dev)
    hs = [torch.zeros(R * bs, H, dtype=torch.bfloat16, device=dev) for _ in range(nsel)]
    eng.buf["done"].fill_(1)  # empty slots stay frozen on the device
    slot_req = [-1] * R       # prompt index living in slot r
    slot_cache = [None] * R   # its target KV cache
    slot_P = [0] * R
    slot_start = [0] * R
    results: List[Optional[torch.Tensor]] = [None] * n
    taus: List[List[int]] = [[] for _ in range(n)]
    tap = ContextTap(target, layer_ids)
    state = torch.empty(2, R, dtype=torch.int32).pin_memory()
    state_dev = torch.empty(2, R, dtype=torch.int32, device=dev)
    next_req = 0
    bt = None                 # one batched verify forward per cycle for all streams
    if graph_target:
        from .target_graph import BatchedVerifyTarget
        try:
            bt = _batched_target(draft, target, eng, int(cache_len), layer_ids)
        except NotImplementedError:
            if graph_target != "auto":
                raise
    graph_target = bt is not None
    since_sync = 0
    fwd0 = bt.n_forwards if bt is not None else 0

    def admit(r: int, i: int):
        ids = prompts[i].to(dev)
        P = ids.shape[1]
        cache = None
        if graph_target:
            logits0, hidden0 = bt.prefill(r, ids)
        else:
            cache = DynamicCache()
            with tap:
                out = target(ids, position_ids=torch.arange(P, device=dev).unsqueeze(0), past_key_values=cache,
                             use_cache=True, logits_to_keep=1)
            logits0, hidden0 = out.logits, list(tap.states)
        first = sample(logits0, temperature, seed=(seed ^ 0x5DEECE66D) + i)
        eng.reset_request(r, ids[0], first.view(-1)[0], max_new_tokens)
        if clamp_tail:
            eng.buf["blk_len"][r] = min(bs, max_new_tokens)
        eng.prefill_context(r, [h[0] for h in hidden0])
        if forced_t is not None:
            f = list(forced_k[i]) or [0]
            forced_t[r] = torch.tensor([f[c % len(f)] for c in range(forced_t.shape[1])], dtype=torch.int32)
        slot_req[r], slot_cache[r], slot_P[r], slot_start[r] = i, cache, P, P

    def harvest(r: int):
        i, P = slot_req[r], slot_P[r]
        max_length = P + max_new_tokens
        n_cyc = int(eng.buf["n_cycles"][r])
        taus[i] = eng.acc_hist[r, :n_cyc].tolist()
        out = eng.output_ids[r:r + 1, :max_length].clone()
        out = out[:, out[0] != draft.mask_token_id]
        if stop_t is not None:
            idx = torch.isin(out[0][P:], stop_t).nonzero(as_tuple=True)[0]
            if idx.numel() > 0:
                out = out[:, : P + idx[0] + 1]
        results[i] = out
        slot_req[r], slot_cache[r] = -1, None

    cycle = 0
    while True:
        for r in range(R):  # refill free slots
            if slot_req[r] < 0 and next_req < n:
                admit(r, next_req)
                next_req += 1
        live = [r for r in range(R) if slot_req[r] >= 0]
        if not live:
            break
        eng.draft_step_graphed()
        tl_c, hs_c = tl, hs
        if graph_target:  # positions, mask and cache lengths come from start[r] on the device; the host only bounds them
            since_sync += 1
            tl_c, hs_c = bt.verify_forward(max(slot_start[r] for r in live) + bs * since_sync)
        for r in () if graph_target else live:  # the caller's target, per request, exactly as the reference calls it
            start = slot_start[r]
            eff = min(bs, slot_P[r] + max_new_tokens - start) if clamp_tail else bs
            with tap:
                out = target(eng.block_ids[r:r + 1, :eff],
                             position_ids=torch.arange(start, start + eff, device=dev).unsqueeze(0),
                             past_key_values=slot_cache[r], use_cache=True)
            tl[r * bs: r * bs + eff] = out.logits[0]
            for s in range(nsel):
                hs[s][r * bs: r * bs + eff] = tap.states[s][0]
        noise = noise_fn(cycle) if (noise_fn is not None and temperature >= 1e-5) else None
        eng.verify_step(tl_c, hs_c, temperature=temperature, noise=noise, seed=seed, stop_ids=stop_t, forced_k=forced_t,
                        clamp_tail=clamp_tail, inject=True)
        cycle += 1
        if graph_target and cycle % sync_every != 0:
            continue  # nothing on the host depends on this cycle's outcome
        since_sync = 0
        # the one host sync of the cycle: the HF target caches need every stream's new length
        state_dev[0].copy_(eng.buf["start"])
        state_dev[1].copy_(eng.buf["done"])
        state.copy_(state_dev, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        for r in live:
            slot_start[r] = int(state[0, r])
            if not graph_target:
                slot_cache[r].crop(slot_start[r])
            if int(state[1, r]):
                harvest(r)
    draft.last_batch_acceptance_lengths = taus
    draft.last_batch_target_forwards = bt.n_forwards - fwd0 if bt is not None else None
    draft.last_batch_cycles = cycle
    return results


def _batched_target(draft, target, eng: DraftEngine, cache_len: int, layer_ids):
    """The batched verify target lives with the cached engine (its graphs hold the engine's buffers)."""
    from .target_graph import BatchedVerifyTarget
    cached = getattr(draft, "_batch_engine_cache", None)
    store = cached[2] if cached is not None and cached[1] is eng else {}
    bt = store.get("bt")
    if bt is None or bt.target is not target or bt.max_cache_len < cache_len or bt.R != eng.R:
        bt = BatchedVerifyTarget(target, eng.block_size, eng.R, max(cache_len, eng.max_seq), layer_ids, eng.buf["start"],
                                 eng.block_ids)
        store["bt"] = bt
    return bt
