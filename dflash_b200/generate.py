"""`dflash_generate`: the reference's instrumented decode loop (benchmark.py:44-272) over the CUDA engine.

Same arguments and the same result record (`output_ids`, `num_input_tokens`, `num_output_tokens`,
`time_to_first_token`, `time_per_output_token`, `acceptance_lengths`, `cycle_trace`, `profile_summary`), the same
tail clamp (`effective_block_size = min(block_size, remaining)`, benchmark.py:104-105), the same `block_size == 1`
autoregressive baseline, and with `collect_profile=True` the same three CUDA-event spans per cycle (draft / target /
cycle) and the same summary keys, so numbers from this loop line up with the reference's `--collect-profile` output.
`draft_steps > 1` (the reference's refinement experiment, which recomputes the block without a cache) is not part
of the hot path and raises.
"""
from __future__ import annotations

import time
from types import SimpleNamespace
from typing import List, Optional

import torch
from transformers import DynamicCache

from .utils import ContextTap, sample


def cuda_time() -> float:
    torch.cuda.synchronize()
    return time.perf_counter()


@torch.inference_mode()
def dflash_generate(model, target, input_ids: torch.Tensor, mask_token_id: int, max_new_tokens: int, block_size: int,
                    stop_token_ids: Optional[List[int]], temperature: float = 0.0, collect_profile: bool = False,
                    draft_steps: int = 1, seed: Optional[int] = None, scheduler=None,
                    draft_temperature: float = 0.0) -> SimpleNamespace:
    """`scheduler` (optional, `dflash_b200.schedule.EwmaBlockScheduler`): per-cycle block-size policy, the
    reference's `dflash_generate_policy` (benchmark_dynamic_schedule.py:260-434). `block_size` must then be the
    largest candidate: the engine is built for it and a smaller block is a shorter device-side `blk_len`.
    `draft_temperature` > 0 samples the DRAFTED tokens from softmax(draft_logits / T) in the lm_head epilogue, as that
    policy loop does (benchmark_dynamic_schedule.py:342); `spec_generate` / `benchmark.py` always draft greedily."""
    if draft_steps != 1:
        raise NotImplementedError("draft_steps > 1 (cache-less block refinement, benchmark.py:114-142) is a research "
                                  "variant outside the draft-and-verify hot path")
    if input_ids.shape[0] != 1:
        raise RuntimeError("dflash_generate: batch size 1 (benchmark.py:58-66)")
    dev = target.device
    P = input_ids.shape[1]
    max_length = P + max_new_tokens
    bs = int(block_size)
    if scheduler is not None and (bs != scheduler.max_block_size or bs < 2):
        raise ValueError("with a scheduler, block_size must equal its largest candidate")
    if seed is None:
        seed = int(torch.randint(0, 2**62, (1,)).item()) if temperature >= 1e-5 else 0
    position_ids = torch.arange(max_length + bs, device=dev).unsqueeze(0)
    cache_t = DynamicCache()
    stop_t = None
    if stop_token_ids is not None and len(stop_token_ids) > 0:
        stop_t = torch.tensor(list(stop_token_ids), dtype=torch.int64, device=dev)

    eng = None
    tap = ContextTap(target, model.target_layer_ids) if bs > 1 else None
    old_bs = model.block_size
    prefill_start = cuda_time()
    if bs > 1:
        if int(mask_token_id) != int(model.mask_token_id):
            raise ValueError("mask_token_id differs from the draft's config")
        model.block_size = bs
        eng = model._get_engine(target.model.embed_tokens.weight, target.lm_head.weight,
                                max_seq=max_length + 2 * bs + 1, out_len=max_length + bs + 1)
        with tap:
            out = target(input_ids, position_ids=position_ids[:, :P], past_key_values=cache_t, use_cache=True,
                         logits_to_keep=1)
        first = sample(out.logits, temperature, seed=seed ^ 0x5DEECE66D)
        eng.reset_request(0, input_ids[0], first.view(-1)[0], max_new_tokens)
        eng.buf["blk_len"][0] = min(bs, max_new_tokens)
        eng.prefill_context(0, [h[0] for h in tap.states])
    else:
        out = target(input_ids, position_ids=position_ids[:, :P], past_key_values=cache_t, use_cache=True,
                     logits_to_keep=1)
        ar_ids = torch.full((1, max_length + 1), int(mask_token_id), dtype=torch.long, device=dev)
        ar_ids[:, :P] = input_ids
        ar_ids[:, P] = sample(out.logits, temperature, seed=seed ^ 0x5DEECE66D).view(-1)[0]
    time_to_first_token = cuda_time() - prefill_start

    decode_start = cuda_time()
    start = P
    acceptance_lengths: List[int] = []
    cycle_trace = []
    draft_prefill = True
    state = torch.empty(2, dtype=torch.int32).pin_memory()
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    try:
        while start < max_length:
            cyc_ev = (ev(), ev()) if collect_profile else None
            if collect_profile:
                cyc_ev[0].record()
            t_cycle = time.perf_counter()
            want = bs if scheduler is None else int(scheduler.select(len(acceptance_lengths)))
            eff = min(want, max_length - start)
            if scheduler is not None:
                eng.buf["blk_len"][0] = eff  # the device-side block length of this cycle
            draft_ev = None
            if eff > 1:
                if collect_profile:
                    draft_ev = (ev(), ev())
                    draft_ev[0].record()
                if draft_temperature >= 1e-5:
                    eng.draft_step_sampled(draft_temperature, seed ^ 0x2545F491)
                else:
                    eng.draft_step_graphed()  # embed -> ctx injection -> layers -> lm_head + argmax (benchmark.py:110-140)
                if collect_profile:
                    draft_ev[1].record()
                if draft_prefill:
                    draft_prefill = False
                    decode_start = cuda_time()
            tgt_ev = (ev(), ev()) if collect_profile else None
            if collect_profile:
                tgt_ev[0].record()
            if bs > 1:
                with tap:
                    out = target(eng.block_ids[:, :eff], position_ids=position_ids[:, start:start + eff],
                                 past_key_values=cache_t, use_cache=True)
            else:
                out = target(ar_ids[:, start:start + 1], position_ids=position_ids[:, start:start + 1],
                             past_key_values=cache_t, use_cache=True)
            if collect_profile:
                tgt_ev[1].record()
            if bs > 1:
                logits = out.logits[0]
                if logits.dtype != torch.bfloat16:
                    logits = logits.to(torch.bfloat16)
                hidden = [h[0].contiguous() for h in tap.states]
                if eff < bs:
                    logits = torch.nn.functional.pad(logits, (0, 0, 0, bs - eff))
                    hidden = [torch.nn.functional.pad(h, (0, 0, 0, bs - eff)) for h in hidden]
                eng.verify_step(logits, hidden, temperature=temperature, seed=seed, stop_ids=stop_t, clamp_tail=True, inject=True)
                state[0:1].copy_(eng.buf["start"][0:1], non_blocking=True)
                state[1:2].copy_(eng.buf["done"][0:1], non_blocking=True)
                torch.cuda.current_stream(dev).synchronize()
                tau = int(state[0]) - start
                done = bool(int(state[1]))
            else:
                tok = sample(out.logits, temperature, seed=seed + start).view(-1)[0]
                ar_ids[:, start + 1] = tok
                tau = 1
                done = stop_t is not None and bool(torch.isin(tok, stop_t).item())
            if scheduler is not None:  # the host has just synchronised: wall clock == cycle time
                scheduler.update(tau=tau, cycle_s=time.perf_counter() - t_cycle, effective_bs=eff,
                                 cycle_idx=len(acceptance_lengths))
            acceptance_lengths.append(tau)
            if collect_profile:
                cyc_ev[1].record()
                cycle_trace.append({"cycle_idx": len(acceptance_lengths) - 1, "generated_tokens_before": start - P,
                                    "effective_block_size": int(eff), "tau": int(tau),
                                    "acceptance_ratio": float(tau / max(1, eff)),
                                    "_events": {"draft": draft_ev, "target": tgt_ev, "cycle": cyc_ev}})
            start += tau
            cache_t.crop(start)
            if done and start < max_length:
                break
        if bs > 1:
            output_ids = eng.output_ids[0:1, :max_length].clone()
        else:
            output_ids = ar_ids[:, :max_length].clone()
    finally:
        model.block_size = old_bs
    output_ids = output_ids[:, output_ids[0] != int(mask_token_id)]
    if stop_t is not None:
        idx = torch.isin(output_ids[0][P:], stop_t).nonzero(as_tuple=True)[0]
        if idx.numel() > 0:
            output_ids = output_ids[:, : P + idx[0] + 1]
    num_output_tokens = output_ids.shape[1] - P
    total_decode_time = cuda_time() - decode_start
    time_per_output_token = total_decode_time / max(1, num_output_tokens)

    profile_summary = None
    if collect_profile:
        torch.cuda.synchronize()
        tot = {"draft": 0.0, "target": 0.0, "cycle": 0.0}
        for row in cycle_trace:
            evs = row.pop("_events")
            for k in ("draft", "target", "cycle"):
                s = 0.0 if evs[k] is None else evs[k][0].elapsed_time(evs[k][1]) / 1000.0
                row[f"{k}_s"] = float(s)
                tot[k] += s
        denom = max(1e-12, tot["draft"] + tot["target"])
        profile_summary = {"target_prefill_s": float(time_to_first_token), "target_decode_s": float(tot["target"]),
                           "draft_decode_s": float(tot["draft"]), "cycle_decode_s_sum": float(tot["cycle"]),
                           "decode_wall_s": float(total_decode_time), "profiled_cycles": len(cycle_trace),
                           "draft_share_decode": float(tot["draft"] / denom),
                           "target_share_decode": float(tot["target"] / denom)}
    return SimpleNamespace(output_ids=output_ids, num_input_tokens=P, num_output_tokens=num_output_tokens,
                           time_to_first_token=time_to_first_token, time_per_output_token=time_per_output_token,
                           acceptance_lengths=acceptance_lengths, cycle_trace=cycle_trace,
                           profile_summary=profile_summary)


@torch.inference_mode()
def dflash_generate_candidates(model, target, input_ids: torch.Tensor, mask_token_id: int, max_new_tokens: int,
                               block_size: int, stop_token_ids: Optional[List[int]], *, fixed_prefix_len: int = 2,
                               rank_top_k: int = 4, max_candidates: int = 4, temperature: float = 0.0,
                               candidate_mode: str = "fixed_prefix_rank", graph_target: bool = False,
                               sync_every: int = 1) -> SimpleNamespace:
    """Multi-candidate speculative decoding, the reference's `dflash_generate_candidate_solutions`
    (benchmark_candidate_solutions.py:417-741) in its `fixed_prefix_rank` mode -- the variant the reference's
    results single out (`results.md:461-521`). Per cycle: the draft step keeps a top-4 per block row in the lm_head
    epilogue and builds up to 4 candidate blocks on the device; ONE target forward verifies them all (batch = number
    of candidates, the target's KV cache repeated per candidate exactly as the reference does, :574-577); one kernel
    picks the candidate with the longest accepted prefix (ties: draft score, then index), commits it and gathers
    the next context features from its rows; the caller keeps that branch of the target cache (:610-614).
    Greedy only, like the reference (:441-442). graph_target=True: the batch-K verify forward replays from a CUDA graph
    over a batch-K static cache (`GraphedCandidateTarget`): no per-cycle cache clone, the winning branch is kept by a
    <= block_size-row copy on the device, and the host polls only every `sync_every` cycles."""
    if candidate_mode != "fixed_prefix_rank":
        raise NotImplementedError("only candidate_mode='fixed_prefix_rank' is on this path")
    if temperature >= 1e-5:
        raise ValueError("multi-candidate verification supports only temperature=0.0 (benchmark_candidate_solutions.py:441)")
    if input_ids.shape[0] != 1:
        raise RuntimeError("dflash_generate_candidates: batch size 1")
    K = min(int(max_candidates), int(rank_top_k))
    if not 2 <= K <= 4:
        raise ValueError("2..4 candidates per cycle")
    dev = target.device
    P = input_ids.shape[1]
    max_length = P + max_new_tokens
    bs = int(block_size)
    if int(mask_token_id) != int(model.mask_token_id):
        raise ValueError("mask_token_id differs from the draft's config")
    position_ids = torch.arange(max_length + bs, device=dev).unsqueeze(0)
    cache_t = DynamicCache()
    stop_t = None
    if stop_token_ids is not None and len(stop_token_ids) > 0:
        stop_t = torch.tensor(list(stop_token_ids), dtype=torch.int64, device=dev)
    tap = ContextTap(target, model.target_layer_ids)
    old_bs = model.block_size
    prefill_start = cuda_time()
    model.block_size = bs
    gt = None
    try:
        eng = model._get_engine(target.model.embed_tokens.weight, target.lm_head.weight, max_seq=max_length + 2 * bs + 1,
                                out_len=max_length + bs + 1, max_candidates=4)
        if graph_target:
            from .target_graph import GraphedCandidateTarget
            cap = max(1024, 1 << (max_length + bs - 1).bit_length())
            gt = GraphedCandidateTarget(target, bs, K, cap, model.target_layer_ids, eng.buf["start"][0:1],
                                        eng.cand_ids[0, :K], eng.buf["chosen"][0:1])
            logits0, hidden0 = gt.prefill(input_ids)
        else:
            with tap:
                out = target(input_ids, position_ids=position_ids[:, :P], past_key_values=cache_t, use_cache=True,
                             logits_to_keep=1)
            logits0, hidden0 = out.logits, list(tap.states)
        first = sample(logits0, 0.0)
        eng.reset_request(0, input_ids[0], first.view(-1)[0], max_new_tokens)
        eng.buf["blk_len"][0] = min(bs, max_new_tokens)
        eng.prefill_context(0, [h[0] for h in hidden0])
        time_to_first_token = cuda_time() - prefill_start
        decode_start = cuda_time()
        start = P
        acceptance_lengths: List[int] = []
        cycle_trace = []
        state = torch.empty(3, dtype=torch.int32).pin_memory()
        state_dev = torch.empty(3, dtype=torch.int32, device=dev)
        n_cand_sum = 0
        cyc = 0
        while gt is not None and start < max_length:
            # device-driven cycle: draft (top-4 epilogue + candidate blocks) -> batch-K target graph -> verify kernels
            eng.draft_step_candidates(K, fixed_prefix_len)
            logits, hidden = gt.verify_forward()
            if logits.dtype != torch.bfloat16:
                logits = logits.to(torch.bfloat16)
            eng.verify_step_candidates(K, logits, hidden, stop_ids=stop_t, clamp_tail=True)
            cyc += 1
            if cyc % max(1, sync_every):
                continue  # blind cycles past the end are frozen on the device (`done`)
            state_dev[0:1].copy_(eng.buf["start"][0:1])
            state_dev[1:2].copy_(eng.buf["done"][0:1])
            state.copy_(state_dev, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            start = int(state[0])
            if int(state[1]):
                break
        if gt is not None:
            n_cyc = int(eng.buf["n_cycles"][0])
            acceptance_lengths = eng.acc_hist[0, :n_cyc].tolist()
            n_cand_sum = K * n_cyc
            start = max_length  # skip the eager loop below
        while start < max_length:
            eff = min(bs, max_length - start)
            if eff > 1:
                eng.draft_step_candidates(K, fixed_prefix_len)
            else:  # a one-token tail block has nothing to draft: every candidate is the committed token
                eng.cand_ids[0, :, 0] = eng.block_ids[0, 0]
                eng.cand_scores.zero_()
            cands = eng.cand_ids[0, :K, :eff]                       # [K, eff], the engine's buffer
            cache_t.batch_repeat_interleave(K)                      # == clone + repeat of the reference (:574-577)
            with tap:
                out = target(cands, position_ids=position_ids[:, start:start + eff].expand(K, -1),
                             past_key_values=cache_t, use_cache=True)
            logits = out.logits
            if logits.dtype != torch.bfloat16:
                logits = logits.to(torch.bfloat16)
            hidden = [h for h in tap.states]
            if eff < bs:
                logits = torch.nn.functional.pad(logits, (0, 0, 0, bs - eff))
                hidden = [torch.nn.functional.pad(h, (0, 0, 0, bs - eff)) for h in hidden]
            eng.verify_step_candidates(K, logits.reshape(K * bs, -1), [h.reshape(K * bs, -1).contiguous() for h in hidden],
                                       stop_ids=stop_t, clamp_tail=True)
            state_dev[0:1].copy_(eng.buf["start"][0:1])
            state_dev[1:2].copy_(eng.buf["done"][0:1])
            state_dev[2:3].copy_(eng.buf["chosen"][0:1])
            state.copy_(state_dev, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            tau = int(state[0]) - start
            chosen = int(state[2])
            acceptance_lengths.append(tau)
            n_cand_sum += K
            cycle_trace.append({"cycle_idx": len(acceptance_lengths) - 1, "generated_tokens_before": start - P,
                                "effective_block_size": int(eff), "tau": int(tau), "num_candidates": K,
                                "chosen_candidate_idx": chosen})
            start += tau
            cache_t.batch_select_indices(torch.tensor([chosen], dtype=torch.long, device=dev))   # keep that branch
            cache_t.crop(start)
            if int(state[1]) and start < max_length:
                break
        output_ids = eng.output_ids[0:1, :max_length].clone()
    finally:
        model.block_size = old_bs
    output_ids = output_ids[:, output_ids[0] != int(mask_token_id)]
    if stop_t is not None:
        idx = torch.isin(output_ids[0][P:], stop_t).nonzero(as_tuple=True)[0]
        if idx.numel() > 0:
            output_ids = output_ids[:, : P + idx[0] + 1]
    num_output_tokens = output_ids.shape[1] - P
    total_decode_time = cuda_time() - decode_start
    return SimpleNamespace(output_ids=output_ids, num_input_tokens=P, num_output_tokens=num_output_tokens,
                           time_to_first_token=time_to_first_token,
                           time_per_output_token=total_decode_time / max(1, num_output_tokens),
                           acceptance_lengths=acceptance_lengths, cycle_trace=cycle_trace,
                           candidate_summary={"candidate_mode": candidate_mode, "fixed_prefix_len": int(fixed_prefix_len),
                                              "avg_candidates_per_cycle": n_cand_sum / max(1, len(acceptance_lengths)),
                                              "candidate_verify_calls": len(acceptance_lengths),
                                              "candidate_count_sum": n_cand_sum},
                           profile_summary=None)
