#!/usr/bin/env python
"""bench.py — the DFlash draft+verify hot path at Qwen3-8B + DFlash-b16 dims (BASELINE.json configs[1]).

A "step" is one pass of the hot path for one request stream: block embedding -> context
injection -> 5 draft layers -> lm_head + argmax, then (given the target's logits / hidden states for the
block, which are synthetic here and NOT timed as target work) posterior sampling -> acceptance -> commit ->
cache-length rollback -> next context gather. The HF target forward is outside the step by definition.

  python bench.py --gpus 1 --steps K --warmup W          # this repo's CUDA path
  python bench.py --impl reference ...                   # the reference algorithm on the host CPU (oracle port)

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

Q8 = dict(hidden=4096, intermediate=12288, draft_layers=5, heads=32, kv_heads=8, head_dim=128, vocab=151936,
          target_layers=36, eps=1e-6, rope_theta=1_000_000.0, mask_token_id=151669, block_size=16)
# BASELINE.json configs[2] / configs[3] draft shapes (parity-test cases, not bench lines)
LLAMA31_8B = dict(hidden=4096, intermediate=14336, draft_layers=5, heads=32, kv_heads=8, head_dim=128, vocab=128256,
                  target_layers=32, eps=1e-5, rope_theta=500000.0, mask_token_id=128255, block_size=16,
                  rope_parameters={"rope_type": "llama3", "rope_theta": 500000.0, "factor": 8.0, "low_freq_factor": 1.0,
                                   "high_freq_factor": 4.0, "original_max_position_embeddings": 8192})
# BASELINE.json configs[0]: Qwen3-4B target shape (hidden 2560 != heads * head_dim = 4096, tied lm_head / embeddings)
QWEN3_4B = dict(hidden=2560, intermediate=9728, draft_layers=5, heads=32, kv_heads=8, head_dim=128, vocab=151936,
                target_layers=36, eps=1e-6, rope_theta=1_000_000.0, mask_token_id=151669, block_size=16)
QWEN3_CODER_30B_A3B = dict(hidden=2048, intermediate=6144, draft_layers=8, heads=32, kv_heads=4, head_dim=128,
                           vocab=151936, target_layers=48, eps=1e-6, rope_theta=10_000_000.0, mask_token_id=151669,
                           block_size=16)
if os.environ.get("DFLASH_BENCH_BS"):  # block-size sweep (BASELINE configs[4] shape): 8 / 16 / 32 slots per block
    Q8 = dict(Q8, block_size=int(os.environ["DFLASH_BENCH_BS"]))
PROMPT_LEN = 128
MAX_NEW = 2048
SHARDED_BATCH = 64  # BASELINE.json configs[4]: global batch sharded over the GPUs of the job
TAU_SCHEDULE_LEN = 64
MEAN_TAU_TARGET = 7.3  # published Qwen3-8B-DFlash-b16 math-average acceptance length


def forced_schedule(seed=0, n=TAU_SCHEDULE_LEN, bs=None):
    """Seeded per-cycle forced-acceptance counts k (tau = k + 1) with mean tau ~= 7.3. The draws are
    arranged so that every window of the schedule has about the same mean (alternating from both ends of the sorted
    draws), which keeps tokens/s comparable between short and long runs."""
    import random
    bs = Q8["block_size"] if bs is None else bs
    rng = random.Random(seed)
    ks = []
    for _ in range(n):
        # geometric-like mixture clipped to [0, bs-1]
        k = min(bs - 1, int(rng.expovariate(1.0 / 6.2)))  # seed 0, n 64 -> mean tau 7.28
        ks.append(k)
    mean = sum(ks) / len(ks)
    rest, out, tot = sorted(ks), [], 0
    while rest:
        want = mean * (len(out) + 1) - tot
        k = min(rest, key=lambda v: (abs(v - want), v))
        rest.remove(k)
        out.append(k)
        tot += k
    return out


def workload_config(world, R=1):
    """The `config` object BOTH arms print (the driver compares them): BASELINE.json configs[1]."""
    ks = forced_schedule(seed=0)
    mean_tau = sum(k + 1 for k in ks) / len(ks)
    return dict(workload="Qwen3-8B + DFlash-b16 draft+verify step (target forward excluded, SURVEY 8d), "
                         f"batch {R} per GPU, bs {Q8['block_size']}, prompt {PROMPT_LEN}, up to {MAX_NEW} new tokens, "
                         f"forced-tau schedule mean {mean_tau:.2f} (BASELINE.json configs[1])",
                l2="inputs larger than L2: 3.34 GB of weights streamed per step vs 126 MB L2",
                parallelism=f"dp{world} (independent request streams per GPU; one packed all-gather of the results "
                            "per generation)",
                mean_tau=mean_tau)


def algorithmic_bytes(dims, S, c, R=1):
    """SURVEY §8(d) per-step bytes (bf16): draft weights + lm_head + KV read + ctx features + embeddings + new KV."""
    H, I, L, V = dims["hidden"], dims["intermediate"], dims["draft_layers"], dims["vocab"]
    Hq, Hkv, D = dims["heads"], dims["kv_heads"], dims["head_dim"]
    nsel = L
    per_layer = (Hq * D * H) + 2 * (Hkv * D * H) + (H * Hq * D) + 3 * H * I + 2 * H + 2 * D
    params = L * per_layer + nsel * H * H + 2 * H
    weights = 2 * params
    lm = 2 * V * H
    kv = R * L * 2 * Hkv * D * 2 * (S + c + dims["block_size"])
    ctx = R * nsel * H * 2 * c
    emb = R * dims["block_size"] * H * 2
    kv_new = R * L * 2 * Hkv * D * 2 * c
    return dict(total=weights + lm + kv + ctx + emb + kv_new, draft_weights=weights, lm_head=lm, kv=kv)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------
# CPU port of the step (oracle) -- cpu_baseline and --impl reference
# ------------------------------------------------------------------------------------------------
def cpu_step_bench(dims, steps, warmup, seed=0):
    """Times oracle.draft_verify_step_cpu (fp32, all host threads) at the same dims / cache lengths.
    Returns (tokens_per_s, ms_per_step, cores, sample description)."""
    import torch
    from oracle import dflash_oracle as O
    torch.manual_seed(seed)
    cores = torch.get_num_threads()
    H, I, L, V = dims["hidden"], dims["intermediate"], dims["draft_layers"], dims["vocab"]
    Hq, Hkv, D, bs = dims["heads"], dims["kv_heads"], dims["head_dim"], dims["block_size"]
    nsel = L

    def rnd(*shape):
        return torch.empty(*shape, dtype=torch.float32).normal_(0, 0.02)

    sd = {"fc.weight": rnd(H, nsel * H), "hidden_norm.weight": torch.ones(H), "norm.weight": torch.ones(H)}
    for l in range(L):
        p = f"layers.{l}."
        sd[p + "self_attn.q_proj.weight"] = rnd(Hq * D, H)
        sd[p + "self_attn.k_proj.weight"] = rnd(Hkv * D, H)
        sd[p + "self_attn.v_proj.weight"] = rnd(Hkv * D, H)
        sd[p + "self_attn.o_proj.weight"] = rnd(H, Hq * D)
        sd[p + "self_attn.q_norm.weight"] = torch.ones(D)
        sd[p + "self_attn.k_norm.weight"] = torch.ones(D)
        sd[p + "mlp.gate_proj.weight"] = rnd(I, H)
        sd[p + "mlp.up_proj.weight"] = rnd(I, H)
        sd[p + "mlp.down_proj.weight"] = rnd(H, I)
        sd[p + "input_layernorm.weight"] = torch.ones(H)
        sd[p + "post_attention_layernorm.weight"] = torch.ones(H)
    embed = rnd(V, H)
    lm_head = rnd(V, H)
    cfg = O.DraftConfig(hidden_size=H, intermediate_size=I, num_hidden_layers=L, num_attention_heads=Hq,
                        num_key_value_heads=Hkv, head_dim=D, rms_norm_eps=dims["eps"], block_size=bs,
                        mask_token_id=dims["mask_token_id"], target_layer_ids=O.build_target_layer_ids(
                            dims["target_layers"], L), rope_theta=dims["rope_theta"])
    cache = O.DraftCache()
    # context for the prompt (cycle 0 of the reference: c = P rows through fc / k_proj / v_proj)
    start = PROMPT_LEN
    block = torch.full((1, bs), dims["mask_token_id"], dtype=torch.long)
    block[0, 0] = 1
    th = torch.randn(1, PROMPT_LEN, nsel * H) * 0.5
    ks = forced_schedule(seed)
    tlogits = torch.randn(1, bs, V)
    hsel = [torch.randn(1, bs, H) * 0.5 for _ in range(nsel)]
    tokens = 0
    t0 = None
    with torch.inference_mode():
        for it in range(warmup + steps):
            if it == warmup:
                t0 = time.perf_counter()
                tokens = 0
            pos = torch.arange(cache.get_seq_length(), start + bs).unsqueeze(0)
            blk, tau, nth = O.draft_verify_step_cpu(sd, cfg, embed, lm_head, block, th, pos, cache, start, tlogits,
                                                    hsel, 0.0)
            k = ks[it % len(ks)]
            tau = k + 1  # forced-tau harness mode, same schedule as the CUDA arm
            th = torch.cat(hsel, dim=-1)[:, :tau, :]
            start += tau
            tokens += tau
            block = torch.full((1, bs), dims["mask_token_id"], dtype=torch.long)
            block[0, 0] = int(blk[0, tau - 1]) % V
    dt = time.perf_counter() - t0
    return tokens / dt, dt / steps * 1e3, cores, (f"{steps} draft+verify steps (after {warmup} warm-up) of the same "
                                                   f"Qwen3-8B/DFlash-b16 workload, fp32, {cores} host threads")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    # torchrun exports OMP_NUM_THREADS=1 to its workers: take every core this process may run on
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        torch.set_num_threads(os.cpu_count() or 1)
    steps = min(args.steps, 60)   # bounded sample: ~10 s of CPU work at ~0.17 s per step
    warmup = args.warmup
    tps, ms, cores, sample = cpu_step_bench(Q8, steps, warmup)
    line = dict(metric="draft_verify_tokens_per_s", value=tps, unit="tokens/s", n_gpus=args.gpus, steps=steps,
                warmup=warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=workload_config(args.gpus, args.requests), step_us=dict(median=ms * 1e3),
                cpu_baseline=dict(value=tps, unit="tokens/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=tps, unit="tokens/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                torch_threads=torch.get_num_threads())
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
def build_engine(dims, device, seed, R=1, max_new=None, keep_draft_logits=False):
    import torch
    from transformers import Qwen3Config
    from dflash_b200 import DFlashDraftModel
    from dflash_b200.engine import DraftEngine
    cfg = Qwen3Config(vocab_size=dims["vocab"], hidden_size=dims["hidden"], intermediate_size=dims["intermediate"],
                      num_hidden_layers=dims["draft_layers"], num_attention_heads=dims["heads"],
                      num_key_value_heads=dims["kv_heads"], head_dim=dims["head_dim"], max_position_embeddings=40960,
                      rms_norm_eps=dims["eps"],
                      rope_parameters=dims.get("rope_parameters",
                                               {"rope_type": "default", "rope_theta": dims["rope_theta"]}))
    cfg.num_target_layers = dims["target_layers"]
    cfg.block_size = dims["block_size"]
    cfg.dflash_config = {"mask_token_id": dims["mask_token_id"]}
    torch.manual_seed(seed)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(device):
            draft = DFlashDraftModel(cfg)  # HF default init (normal 0.02, norms 1) directly on the GPU
    finally:
        torch.set_default_dtype(old)
    draft = draft.eval()
    g = torch.Generator(device=device).manual_seed(seed + 1)
    embed = torch.empty(dims["vocab"], dims["hidden"], dtype=torch.bfloat16, device=device).normal_(0, 0.02, generator=g)
    lm_head = torch.empty(dims["vocab"], dims["hidden"], dtype=torch.bfloat16, device=device).normal_(0, 0.02, generator=g)
    span = PROMPT_LEN + (MAX_NEW if max_new is None else max_new) + 64
    eng = DraftEngine(draft, embed, lm_head, max_seq=span, out_len=span,
                      max_requests=R, block_size=dims["block_size"], keep_draft_logits=keep_draft_logits,
                      use_pdl=os.environ.get("DFLASH_PDL", "1") != "0", device=device)
    return draft, eng, embed, lm_head


def gpu_reference_bench(dims, draft, embed, lm_head, device, ks, steps=100, warmup=10):
    """The reference's torch op sequence for the SAME step on the SAME GPU (north-star comparison: ">= 5x lower
    draft+verify step latency than the reference torch/sdpa path"). The reference source cannot travel to the GPU box,
    so this runs the oracle's op-for-op restatement (oracle/dflash_oracle.py, pinned to the reference by
    tests/test_oracle_golden.py) in bf16 on the engine's own weights, with the reference's concat cache and per-cycle
    host sync (model/dflash.py:235-268), once with sdpa (transformers' default dispatch) and once with eager attention.
    Spans as benchmark.py:99-160 (CUDA events around the step). A reported baseline: the oracle is the thing timed
    here, never the thing shipped."""
    import torch
    from oracle import dflash_oracle as O
    H, V, L, bs = dims["hidden"], dims["vocab"], dims["draft_layers"], dims["block_size"]
    nsel = L
    bf = torch.bfloat16
    sd = {k: v for k, v in draft.state_dict().items()}
    cfg = O.DraftConfig.from_hf(draft)
    cfg.inv_freq = cfg.get_inv_freq().to(device)
    g = torch.Generator(device=device).manual_seed(100)
    tlogits = torch.randn(1, bs, V, device=device, generator=g).to(bf)
    hsel = [(torch.randn(1, bs, H, device=device, generator=g) * 0.5).to(bf) for _ in range(nsel)]
    out = {}
    import contextlib
    from torch.nn.attention import SDPBackend, sdpa_kernel
    for impl in ("sdpa", "sdpa_no_cudnn", "eager"):
        O.ATTN_IMPL = "sdpa" if impl.startswith("sdpa") else impl
        # sdpa_no_cudnn: the same dispatch with torch's flash / memory-efficient / math backends only -- the cuDNN
        # backend torch prefers on this GPU re-plans for every new KV length, which is most of the `sdpa` row
        backends = (sdpa_kernel([SDPBackend.FLASH_ATTENTION, SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH])
                    if impl == "sdpa_no_cudnn" else contextlib.nullcontext())
        cache = O.DraftCache()
        start = PROMPT_LEN
        block = torch.full((1, bs), dims["mask_token_id"], dtype=torch.long, device=device)
        block[0, 0] = 1
        th = (torch.randn(1, PROMPT_LEN, nsel * H, device=device, generator=g) * 0.5).to(bf)
        times = []
        try:
            with torch.inference_mode(), backends:
                for it in range(warmup + steps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record()
                    pos = torch.arange(cache.get_seq_length(), start + bs, device=device).unsqueeze(0)
                    O.draft_verify_step_cpu(sd, cfg, embed, lm_head, block, th, pos, cache, start, tlogits, hsel, 0.0)
                    e1.record()
                    torch.cuda.synchronize()
                    if it >= warmup:
                        times.append(e0.elapsed_time(e1) * 1e3)
                    tau = ks[it % len(ks)] + 1  # the same forced-acceptance schedule as the CUDA arm
                    th = torch.cat(hsel, dim=-1)[:, :tau, :]
                    start += tau
        except Exception as ex:  # noqa: BLE001 -- a backend set torch cannot serve on this GPU
            out[impl] = dict(error=f"{type(ex).__name__}: {ex}"[:200])
            continue
        times.sort()
        out[impl] = dict(step_us_median=times[len(times) // 2], step_us_p10=times[len(times) // 10],
                         step_us_p90=times[len(times) * 9 // 10], steps=steps, final_cache_len=cache.get_seq_length())
    O.ATTN_IMPL = "sdpa"
    out["note"] = ("oracle port of model/dflash.py:235-268 (torch bf16 ops, concat KV cache, host sync per cycle) on this "
                   "GPU, same weights / dims / forced-tau schedule as the CUDA arm; sdpa = transformers' default dispatch "
                   "(its per-step KV length change makes the cuDNN fused-attention backend re-plan every call), "
                   "sdpa_no_cudnn = the same with torch's flash / memory-efficient / math backends only, eager = "
                   "softmax(QK^T)V in torch ops")
    return out


def full_cycle_bench(dims, draft, eng, embed, lm_head, device, ks, new_tokens=512):
    """Whole spec-decode cycles through the public API (`draft.spec_generate`) with a random-init HF target of
    Qwen3-8B shape: (a) the target called eagerly with a DynamicCache exactly as the reference does, (b) the same
    module replayed from a CUDA graph over a static cache. Forced-tau schedule as above.
    Context for the headline numbers, not part of them."""
    import torch
    from transformers import Qwen3Config, Qwen3ForCausalLM
    cfg = Qwen3Config(vocab_size=dims["vocab"], hidden_size=dims["hidden"], intermediate_size=dims["intermediate"],
                      num_hidden_layers=dims["target_layers"], num_attention_heads=dims["heads"],
                      num_key_value_heads=dims["kv_heads"], head_dim=dims["head_dim"], max_position_embeddings=40960,
                      rms_norm_eps=dims["eps"], tie_word_embeddings=False,
                      rope_parameters={"rope_type": "default", "rope_theta": dims["rope_theta"]})
    cfg._attn_implementation = "sdpa"
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(device):
            target = Qwen3ForCausalLM(cfg)
    finally:
        torch.set_default_dtype(old)
    target = target.eval()
    # share the benchmark's embedding / head tensors so the draft engine streams the same weights
    target.model.embed_tokens.weight.data = embed
    target.lm_head.weight.data = lm_head
    g = torch.Generator(device=device).manual_seed(7)
    prompt = torch.randint(0, dims["vocab"] - 1, (1, PROMPT_LEN), device=device, generator=g)
    out = {}
    # (attn_implementation="eager" cannot be captured: transformers builds a CPU tensor in its mask path)
    rows = (("default", None), ("eager_target", False), ("graphed_target", True))
    for name, graph in rows:
        draft.spec_generate(target, prompt, 64, None, 0.0, forced_k=ks, graph_target=graph)  # warm-up / capture
        if graph is not False:  # (and one generation of the timed length: the first one after a capture runs ~1 ms
            draft.spec_generate(target, prompt, new_tokens, None, 0.0, forced_k=ks, graph_target=graph)  # per cycle slower)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ids = draft.spec_generate(target, prompt, new_tokens, None, 0.0, forced_k=ks, graph_target=graph)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        taus = draft.last_acceptance_lengths
        n = ids.shape[1] - prompt.shape[1]
        out[name] = dict(tokens_per_s=n / dt, new_tokens=n, cycles=len(taus), mean_tau=sum(taus) / max(1, len(taus)),
                         ms_per_cycle=dt / max(1, len(taus)) * 1e3, wall_s=dt)
    # the reference's own loop (oracle port of model/dflash.py:192-277: torch ops for the draft, the same HF target
    # called eagerly with a DynamicCache), same prompt / schedule / GPU
    from oracle import dflash_oracle as O
    sd = {k: v for k, v in draft.state_dict().items()}
    ocfg = O.DraftConfig.from_hf(draft)
    ocfg.inv_freq = ocfg.get_inv_freq().to(device)
    O.spec_generate(sd, ocfg, target, prompt, 64, None, 0.0, forced_k=ks)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ids, taus = O.spec_generate(sd, ocfg, target, prompt, new_tokens, None, 0.0, forced_k=ks)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = ids.shape[1] - prompt.shape[1]
    out["reference_loop"] = dict(tokens_per_s=n / dt, new_tokens=n, cycles=len(taus),
                                 mean_tau=sum(taus) / max(1, len(taus)), ms_per_cycle=dt / max(1, len(taus)) * 1e3,
                                 wall_s=dt)
    out["speedup_vs_reference_loop"] = {k: out[k]["tokens_per_s"] / out["reference_loop"]["tokens_per_s"]
                                        for k in ("default", "eager_target", "graphed_target")}
    # batched serving (BASELINE configs[2]-[4] shape): R request streams in one engine, ONE target verify forward per
    # cycle for all of them (ragged static cache, CUDA graph); the reference's shape of the same work is one eager
    # target call per request per cycle (graph_target=False), timed at R = 16 on a shorter generation
    batched = {}
    for R, n_new, graph in ((16, 256, "auto"), (64, 256, "auto"), (16, 48, False)):
        gp = torch.Generator(device=device).manual_seed(40 + R)
        prompts = [torch.randint(0, dims["vocab"] - 1, (1, PROMPT_LEN), device=device, generator=gp) for _ in range(R)]
        fk = [ks[r % len(ks):] + ks[:r % len(ks)] for r in range(R)]  # every stream its own rotation of the schedule
        draft.spec_generate_batch(target, prompts, 2 * dims["block_size"], None, 0.0, forced_k=fk, graph_target=graph)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        outs = draft.spec_generate_batch(target, prompts, n_new, None, 0.0, forced_k=fk, graph_target=graph,
                                         sync_every=4 if graph else 1)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n = sum(o.shape[1] - PROMPT_LEN for o in outs)
        name = f"R{R}_" + ("batched_target" if graph else "per_request_eager_target")
        batched[name] = dict(tokens_per_s=n / dt, new_tokens=n, streams=R, cycles=draft.last_batch_cycles,
                             target_forwards=draft.last_batch_target_forwards if graph else draft.last_batch_cycles * R,
                             ms_per_cycle=dt / max(1, draft.last_batch_cycles) * 1e3, wall_s=dt)
        draft.release_engine()
        torch.cuda.empty_cache()
    batched["speedup_R16_batched_vs_per_request"] = (batched["R16_batched_target"]["tokens_per_s"] /
                                                     batched["R16_per_request_eager_target"]["tokens_per_s"])
    out["batched"] = batched
    out["note"] = ("prefill + decode wall clock of spec_generate, batch 1, random-init Qwen3-8B target (bf16, sdpa), "
                   "forced-tau schedule; the target forward is the caller's HF module in every row; `default` = what "
                   "spec_generate does without extra arguments (graphed target when capturable); reference_loop = the "
                   "oracle port of the reference's loop on the same GPU")
    draft.release_engine()
    del target
    torch.cuda.empty_cache()
    return out


TAU_HIST_COLS = 512  # acceptance lengths per request carried by the result gather


class StepRunner:
    """R request streams of the BASELINE workload in one engine: synthetic target outputs resident in HBM, the whole
    draft+verify step (device-resident state, PDL edges) captured in ONE CUDA graph."""

    def __init__(self, eng, dims, device, R, seed, ks, inject=True):
        import torch
        self.eng, self.R, self.ks, self.device, self.inject = eng, R, ks, device, inject
        bs, H, V, nsel = dims["block_size"], dims["hidden"], dims["vocab"], dims["draft_layers"]
        g = torch.Generator(device=device).manual_seed(seed)
        self.tlogits = torch.randn(R * bs, V, device=device, generator=g).to(torch.bfloat16)
        self.hsel = [(torch.randn(R * bs, H, device=device, generator=g) * 0.5).to(torch.bfloat16) for _ in range(nsel)]
        self.prompt_hidden = [(torch.randn(PROMPT_LEN, H, device=device, generator=g) * 0.5).to(torch.bfloat16)
                              for _ in range(nsel)]
        self.prompt = torch.randint(0, V - 1, (PROMPT_LEN,), device=device, generator=g)
        self.forced = torch.tensor([ks] * R, dtype=torch.int32, device=device)  # every stream: the same schedule
        self.steps_per_gen, cum = 0, 0  # cycles one 2048-token generation lasts under the schedule
        while cum + ks[self.steps_per_gen % len(ks)] + 1 <= MAX_NEW - bs:
            cum += ks[self.steps_per_gen % len(ks)] + 1
            self.steps_per_gen += 1
        self.since_reset, self.tokens = 0, 0
        self.reset()
        self.enqueue_step()
        torch.cuda.synchronize()
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream())
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            self.enqueue_step()
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph, stream=side):
            self.enqueue_step()
        torch.cuda.synchronize()
        self.side = side
        self.reset()

    def enqueue_step(self):
        """One step = draft (layers + lm_head) -> verify + the NEXT cycle's context injection (overlapped with the verify
        kernel, reading the hidden states in place). --no-inject: the injection kernel at the head of the draft step,
        fed from the features the verify kernel gathers."""
        if self.inject:
            self.eng._injected = "fresh"  # the graph is a fixed kernel list: reset() re-embeds the block rows itself
            self.eng.draft_step_injected()
            self.eng.verify_step(self.tlogits, self.hsel, temperature=0.0, forced_k=self.forced, inject=True)
        else:
            self.eng._injected = "no"
            self.eng.draft_step()
            self.eng.verify_step(self.tlogits, self.hsel, temperature=0.0, forced_k=self.forced)

    def reset(self):
        for r in range(self.R):
            self.eng.reset_request(r, self.prompt, 1, MAX_NEW)
            self.eng.prefill_context(r, self.prompt_hidden)
        if self.inject:
            self.eng.embed_block()
        self.since_reset = 0

    def run(self, n, events=None):
        """n graph replays on the current stream. When the 2048-token generation is used up the requests are
        re-initialised (a few memsets + the prompt context pass: inside the timed region, ~0.3% of it)."""
        ks = self.ks
        for i in range(n):
            if self.since_reset >= self.steps_per_gen:
                self.reset()
            if events is not None:
                events[i][0].record()
            self.graph.replay()
            if events is not None:
                events[i][1].record()
            self.tokens += self.R * (ks[self.since_reset % len(ks)] + 1)
            self.since_reset += 1

    def packed_result(self):
        """The generation's result rows [R, 1 + MAX_NEW + TAU_HIST_COLS] int32 (what a DP rank contributes)."""
        from dflash_b200 import dist as ddist
        eng = self.eng
        n_out = (eng.buf["start"][: self.R] - PROMPT_LEN).to(self.tlogits.device)
        toks = eng.output_ids[: self.R, PROMPT_LEN:PROMPT_LEN + MAX_NEW]
        taus = eng.acc_hist[: self.R, :TAU_HIST_COLS]
        return ddist.pack_streams(n_out, toks, taus)

    def check(self):
        import torch
        eng, ks = self.eng, self.ks
        n_dev = int(eng.buf["n_cycles"][0])
        assert n_dev == self.since_reset, (n_dev, self.since_reset)
        assert eng.acc_hist[0, :n_dev].tolist() == [ks[i % len(ks)] + 1 for i in range(n_dev)], "forced-tau drifted"
        assert int(eng.buf["done"][: self.R].sum()) == 0
        assert torch.equal(eng.buf["start"][: self.R], eng.buf["start"][:1].expand(self.R))


def timed_steps(runner, steps, warmup, device, local, sample_clocks=True):
    """W warm-up steps, then exactly `steps` steps + ONE packed all-gather of the results (the data path's only
    collective, once per generation) between two events, barrier + synchronize on both sides. Returns a dict."""
    import torch
    from dflash_b200 import dist as ddist
    world = ddist.world_size()
    runner.reset()
    runner.run(warmup)
    packed = runner.packed_result()
    gathered = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=device)
    ddist.all_gather_packed(packed, gathered)  # communicator set-up happens here, outside the timed region
    torch.cuda.synchronize()
    ddist.barrier()
    clocks = ClockSampler(local) if sample_clocks else None
    if clocks:
        clocks.start()
    events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t_begin, t_end, t_g0 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize()
    ddist.barrier()
    runner.tokens = 0
    t_begin.record()
    runner.run(steps, events)
    t_g0.record()
    ddist.all_gather_packed(runner.packed_result(), gathered)
    t_end.record()
    torch.cuda.synchronize()
    ddist.barrier()
    clock_info = clocks.stop() if clocks else None
    runner.check()
    total_ms = t_begin.elapsed_time(t_end)
    step_us = sorted(e0.elapsed_time(e1) * 1e3 for e0, e1 in events)
    total_ms_max = ddist.max_over_ranks(total_ms, device)
    tokens_all = ddist.sum_over_ranks(runner.tokens, device)
    n_expect = int(runner.eng.buf["start"][0]) - PROMPT_LEN
    assert gathered.shape[0] == world and bool((gathered[:, :, 0] == n_expect).all()), "gathered lengths differ"
    return dict(value=tokens_all / (total_ms_max / 1e3), total_ms=total_ms_max, step_us=step_us,
                # the collective itself = what the LAST rank to arrive sees (min over ranks); the max also contains the
                # wait for the slowest rank's steps, which `value` already counts (max over ranks of the whole region)
                gather_us=-ddist.max_over_ranks(-t_g0.elapsed_time(t_end) * 1e3, device),
                gather_wait_us_max=ddist.max_over_ranks(t_g0.elapsed_time(t_end) * 1e3, device), clocks=clock_info,
                gather_bytes_per_rank=packed.numel() * 4)


def run_cuda_arm(args):
    import torch
    from dflash_b200 import dist as ddist
    from dflash_b200.engine import DraftEngine
    rank, world, local = ddist.init()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dims = Q8
    bs, H, V, L = dims["block_size"], dims["hidden"], dims["vocab"], dims["draft_layers"]
    nsel = L
    R = args.requests
    draft, eng, embed, lm_head = build_engine(dims, device, seed=rank, R=R)
    ks = forced_schedule(seed=0)
    mean_tau = sum(k + 1 for k in ks) / len(ks)
    runner = StepRunner(eng, dims, device, R, 100 + rank, ks, inject=not args.no_inject)
    tlogits, hsel, forced, side = runner.tlogits, runner.hsel, runner.forced, runner.side
    steps_per_gen = runner.steps_per_gen
    launches_per_step = (eng.kernels_per_draft_step_injected + eng.kernels_per_verify_inject_step if runner.inject
                         else eng.kernels_per_draft_step + eng.kernels_per_verify_step)

    res = timed_steps(runner, args.steps, args.warmup, device, local)
    value, total_ms_max, step_us, clock_info = res["value"], res["total_ms"], res["step_us"], res["clocks"]
    med = step_us[len(step_us) // 2]
    p10, p90 = step_us[len(step_us) // 10], step_us[(len(step_us) * 9) // 10]

    def do_reset():
        runner.reset()

    # ---- e2e: host buffers in, result out, through the C-ABI call sequence, copies inside the timed region
    # the step's host inputs live in ONE pinned staging buffer (logits rows, then the n_sel hidden-state blocks) so
    # that a step is one cudaMemcpyAsync; the device side is one buffer with views at the same offsets
    n_tl, n_h = tlogits.numel(), hsel[0].numel()
    stage_host = torch.empty(n_tl + nsel * n_h, dtype=torch.bfloat16).pin_memory()
    stage_host[:n_tl].copy_(tlogits.view(-1).cpu())
    for i, h in enumerate(hsel):
        stage_host[n_tl + i * n_h: n_tl + (i + 1) * n_h].copy_(h.view(-1).cpu())
    stage_dev = torch.empty_like(stage_host, device=device)
    tl_dev = stage_dev[:n_tl].view_as(tlogits)
    hs_dev = [stage_dev[n_tl + i * n_h: n_tl + (i + 1) * n_h].view_as(hsel[0]) for i in range(nsel)]
    res_host = torch.empty(2 + bs, dtype=torch.int64).pin_memory()
    res_dev = torch.empty(2 + bs, dtype=torch.int64, device=device)
    h2d = stage_host.numel() * 2
    d2h = res_host.numel() * 8
    # ONE graph per step: the H2D copy of the step's host inputs (pinned memory) forks onto a copy stream and joins in
    # front of the verify half -- the target's logits and hidden states are only consumed there, so the copy overlaps
    # the draft half -- and the D2H copy of the result is the graph's last node. Per step the host launches one graph
    # and waits for it.
    copy_stream = torch.cuda.Stream(device=device)

    def verify_enqueue():
        eng.verify_step(tl_dev, hs_dev, temperature=0.0, forced_k=forced, inject=runner.inject)
        res_dev[0] = eng.buf["start"][0]
        res_dev[1] = eng.buf["ctx_len"][0]
        res_dev[2:] = eng.posterior[0]

    def whole_step():
        cur = torch.cuda.current_stream()
        copy_stream.wait_stream(cur)
        with torch.cuda.stream(copy_stream):
            stage_dev.copy_(stage_host, non_blocking=True)
        eng._injected = "fresh" if runner.inject else "no"
        eng.draft_step()  # (injected state: the step without its injection kernel)
        cur.wait_stream(copy_stream)
        verify_enqueue()
        res_host.copy_(res_dev, non_blocking=True)

    with torch.cuda.stream(side):
        whole_step()
    torch.cuda.synchronize()
    step_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(step_graph, stream=side):
        whole_step()
    torch.cuda.synchronize()

    def e2e_step():
        step_graph.replay()
        torch.cuda.current_stream().synchronize()  # the caller reads the accepted length every cycle
        return int(res_host[1])

    do_reset()
    for _ in range(max(3, args.warmup // 4)):
        e2e_step()
    torch.cuda.synchronize()
    ddist.barrier()
    e2e_steps = min(args.steps, steps_per_gen - 8)
    t0 = time.perf_counter()
    e2e_tokens = 0
    for _ in range(e2e_steps):
        e2e_tokens += R * e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_s_max = ddist.max_over_ranks(e2e_s, device)
    e2e_tokens_all = ddist.sum_over_ranks(e2e_tokens, device)
    e2e_value = e2e_tokens_all / e2e_s_max

    # ---- roofline of the dominant kernel (lm_head GEMM + fused argmax: 1.245 GB of the 3.34 GB step), timed
    #      alone, back to back, with CUDA events on the launching stream. Weights (1.2 GB) >> L2 (126 MB).
    import ctypes
    from dflash_b200 import _lib
    lib = _lib.load()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    sms = eng.lib.dflash_device_check()
    cv = torch.empty(sms * 16, dtype=torch.float32, device=device)
    ci = torch.empty(sms * 16, dtype=torch.int32, device=device)
    toks = torch.empty(16, dtype=torch.int64, device=device)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def lm_once():
        lib.dflash_gemm_argmax(ctypes.c_void_p(lm_head.data_ptr()), V, V, H, ctypes.c_void_p(eng.hn.data_ptr()), 16, 0,
                               16, 16, ctypes.c_void_p(cv.data_ptr()), ctypes.c_void_p(ci.data_ptr()), None, 0,
                               ctypes.c_void_p(toks.data_ptr()), sms, 0, st)

    for _ in range(3):
        lm_once()
    torch.cuda.synchronize()
    n_lm = 20
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for _ in range(n_lm):
        lm_once()
    eb.record()
    torch.cuda.synchronize()
    lm_us = ea.elapsed_time(eb) * 1e3 / n_lm  # includes the 16-warp candidate reduce that follows each GEMM
    lm_bytes = 2 * V * H + 16 * H * 2
    achieved = lm_bytes / (lm_us * 1e-6) / 1e9

    # DRAM traffic of that kernel: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this
    # command, kept under profiles/ by scripts/ncu_table.py (not re-measured per run; null when no capture is committed)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_lm_head_r2.json")))
        traffic, traffic_src = float(tj["dram_bytes_per_launch"]), "profiles/ncu_lm_head_r2.json (" + tj.get("source", "") + ")"
    except Exception:
        pass

    # whole-step roofline at the mean cache length of the timed region
    S_mid = PROMPT_LEN + int(mean_tau * min(args.steps, steps_per_gen) / 2)
    ab = algorithmic_bytes(dims, S_mid, round(mean_tau), R)
    step_gbs = ab["total"] / (med * 1e-6) / 1e9

    # ---- BASELINE configs[4] shape: a global batch of 64 requests sharded over the GPUs of the job (64 / N request
    #      streams per engine), same step, the packed all-gather of all 64 results inside the timed region
    sharded = None
    if not args.no_sharded and R == 1 and SHARDED_BATCH % world == 0:
        Rs = SHARDED_BATCH // world
        span = PROMPT_LEN + MAX_NEW + 64
        eng_s = DraftEngine(draft, embed, lm_head, max_seq=span, out_len=span, max_requests=Rs, block_size=bs,
                            device=device)
        runner_s = StepRunner(eng_s, dims, device, Rs, 200 + rank, ks, inject=not args.no_inject)
        rs = timed_steps(runner_s, max(10, min(args.steps, 40)), 5, device, local, sample_clocks=False)
        su = rs["step_us"]
        sharded = dict(workload=f"global batch {SHARDED_BATCH} requests, {Rs} request streams per GPU (BASELINE.json "
                                "configs[4] shape at bs 16), forced-tau schedule, one packed all-gather per generation",
                       scaling="strong", global_batch=SHARDED_BATCH, streams_per_gpu=Rs, tokens_per_s=rs["value"],
                       step_us_median=su[len(su) // 2], steps=len(su), gather_us=rs["gather_us"],
                       gather_wait_us_max=rs["gather_wait_us_max"],
                       gather_bytes_per_rank=rs["gather_bytes_per_rank"])
        eng_s.close()
        del runner_s, eng_s
        torch.cuda.empty_cache()

    gpu_ref = None
    if rank == 0 and world == 1 and R == 1 and not args.no_gpu_reference:
        try:
            gpu_ref = gpu_reference_bench(dims, draft, embed, lm_head, device, ks)
        except Exception as ex:  # the headline numbers do not depend on this section
            gpu_ref = dict(error=f"{type(ex).__name__}: {ex}")

    full_cycle = None
    if rank == 0 and world == 1 and R == 1 and not args.no_full_cycle:
        try:
            full_cycle = full_cycle_bench(dims, draft, eng, embed, lm_head, device, ks)
        except Exception as ex:  # the headline numbers above do not depend on this section
            full_cycle = dict(error=f"{type(ex).__name__}: {ex}")

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                tps, ms, cores, sample = cpu_step_bench(dims, steps=48, warmup=2)  # ~10 s of CPU work
                cpu = dict(value=tps, unit="tokens/s", cores=cores, kind="port", sample=sample, ms_per_step=ms)
            except Exception as ex:  # e.g. not enough host RAM
                cpu = dict(value=None, unit="tokens/s", cores=os.cpu_count(), kind="port", sample=f"failed: {ex}")
        line = dict(
            metric="draft_verify_tokens_per_s", value=value, unit="tokens/s", n_gpus=world, steps=args.steps,
            warmup=args.warmup, ms_per_step=total_ms_max / args.steps, higher_is_better=True, scaling="weak",
            vs_baseline=None, dtype="bf16", data="synthetic",
            config=workload_config(world, R),
            step_us=dict(median=med, p10=p10, p90=p90),
            hbm_gbs_step=step_gbs,
            step_roofline=dict(bound="hbm", achieved=step_gbs, peak=peak_gbs, unit="GB/s", frac=step_gbs / peak_gbs,
                               algorithmic_bytes=ab["total"], note=f"all {launches_per_step} kernels of the step, median step time"),
            roofline=dict(bound="hbm", achieved=achieved, peak=peak_gbs, unit="GB/s", frac=achieved / peak_gbs,
                          traffic=traffic, traffic_source=traffic_src,
                          kernel="gemm_skinny_kernel<16,argmax> (lm_head 151936x4096 + fused argmax)",
                          launch_us=lm_us, algorithmic_bytes=lm_bytes, peak_source=peak_src),
            cpu_baseline=cpu,
            e2e=dict(value=e2e_value, unit="tokens/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                     steps=e2e_steps, ms_per_step=e2e_s_max / e2e_steps * 1e3),
            gpu_launches=launches_per_step * args.steps,
            launches_per_step=launches_per_step,
            clocks=clock_info,
            gather_us=res["gather_us"],
            gather=dict(collective="all_gather_into_tensor (NCCL), one packed int32 buffer per generation, inside the "
                                   "timed region", bytes_per_rank=res["gather_bytes_per_rank"], us=res["gather_us"],
                        wait_us_max=res["gather_wait_us_max"],
                        note="us = the collective as the last-arriving rank sees it; wait_us_max also holds the wait for "
                             "the slowest rank's steps"),
            sharded_batch=sharded,
            gpu_reference=gpu_ref,
            step_speedup_vs_torch=(None if not gpu_ref or "error" in gpu_ref else
                                   {k: gpu_ref[k]["step_us_median"] / med for k in ("sdpa", "sdpa_no_cudnn", "eager")
                                    if "step_us_median" in gpu_ref.get(k, {})}),
            full_cycle=full_cycle,
        )
        emit(line)
    ddist.barrier()
    eng.close()
    if world > 1:
        import torch.distributed as tdist
        tdist.destroy_process_group()
    return 0


_RESULT_FD = None


def _claim_stdout():
    """stdout carries exactly ONE line, the result JSON: everything else this process or its libraries write to
    fd 1 (NCCL prints its version banner there on some hosts) is sent to stderr."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--requests", type=int, default=1, choices=[1, 2, 4, 8, 16, 32, 64],
                    help="request streams per GPU sharing one weight stream (headline metric is quoted at 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inject", action="store_true",
                    help="A/B switch: context injection at the head of the draft step (fed from gathered features) "
                         "instead of behind the verify kernel it overlaps")
    ap.add_argument("--no-sharded", action="store_true", help="skip the global-batch-64 sharded section")
    ap.add_argument("--no-gpu-reference", action="store_true",
                    help="skip the torch op sequence of the same step on the GPU (north-star comparison)")
    ap.add_argument("--no-full-cycle", action="store_true",
                    help="skip the whole-cycle section (spec_generate with a random-init Qwen3-8B target)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    _claim_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_cuda_arm(args)


if __name__ == "__main__":
    sys.exit(main())
