"""Times the reference's torch/sdpa eager op sequence for ONE draft+verify step on the GPU (bf16, Qwen3-8B +
DFlash-b16 dims) -- the number the north-star ">= 5x lower step latency than the reference torch/sdpa path"
refers to. The reference source cannot travel to the GPU box, so this runs the oracle's op-for-op restatement
(oracle/dflash_oracle.py, validated against the reference in tests/test_oracle_golden.py) with the same
DynamicCache-style concat cache, per-cycle .item() sync and sdpa dispatch as model/dflash.py:235-268.

This is a measurement script (writes profiles/ref_eager_gpu_r1.json); it is not part of the product or of bench.py.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import PROMPT_LEN, Q8, forced_schedule  # noqa: E402
from oracle import dflash_oracle as O  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    dims = Q8
    H, I, L, V = dims["hidden"], dims["intermediate"], dims["draft_layers"], dims["vocab"]
    Hq, Hkv, D, bs = dims["heads"], dims["kv_heads"], dims["head_dim"], dims["block_size"]
    nsel = L
    bf = torch.bfloat16
    torch.manual_seed(0)

    def rnd(*shape):
        return torch.empty(*shape, dtype=bf, device=dev).normal_(0, 0.02)

    sd = {"fc.weight": rnd(H, nsel * H), "hidden_norm.weight": torch.ones(H, dtype=bf, device=dev),
          "norm.weight": torch.ones(H, dtype=bf, device=dev)}
    for l in range(L):
        p = f"layers.{l}."
        sd[p + "self_attn.q_proj.weight"] = rnd(Hq * D, H)
        sd[p + "self_attn.k_proj.weight"] = rnd(Hkv * D, H)
        sd[p + "self_attn.v_proj.weight"] = rnd(Hkv * D, H)
        sd[p + "self_attn.o_proj.weight"] = rnd(H, Hq * D)
        sd[p + "self_attn.q_norm.weight"] = torch.ones(D, dtype=bf, device=dev)
        sd[p + "self_attn.k_norm.weight"] = torch.ones(D, dtype=bf, device=dev)
        sd[p + "mlp.gate_proj.weight"] = rnd(I, H)
        sd[p + "mlp.up_proj.weight"] = rnd(I, H)
        sd[p + "mlp.down_proj.weight"] = rnd(H, I)
        sd[p + "input_layernorm.weight"] = torch.ones(H, dtype=bf, device=dev)
        sd[p + "post_attention_layernorm.weight"] = torch.ones(H, dtype=bf, device=dev)
    embed, lm_head = rnd(V, H), rnd(V, H)
    cfg = O.DraftConfig(hidden_size=H, intermediate_size=I, num_hidden_layers=L, num_attention_heads=Hq,
                        num_key_value_heads=Hkv, head_dim=D, rms_norm_eps=dims["eps"], block_size=bs,
                        mask_token_id=dims["mask_token_id"],
                        target_layer_ids=O.build_target_layer_ids(dims["target_layers"], L), rope_theta=dims["rope_theta"])
    cfg.inv_freq = cfg.get_inv_freq().to(dev)
    ks = forced_schedule(0)
    tlogits = torch.randn(1, bs, V, device=dev).to(bf)
    hsel = [(torch.randn(1, bs, H, device=dev) * 0.5).to(bf) for _ in range(nsel)]
    results = {}
    for impl in ("sdpa", "eager"):
        O.ATTN_IMPL = impl
        cache = O.DraftCache()
        start = PROMPT_LEN
        block = torch.full((1, bs), dims["mask_token_id"], dtype=torch.long, device=dev)
        block[0, 0] = 1
        th = (torch.randn(1, PROMPT_LEN, nsel * H, device=dev) * 0.5).to(bf)
        times = []
        with torch.inference_mode():
            for it in range(20 + 200):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record()
                pos = torch.arange(cache.get_seq_length(), start + bs, device=dev).unsqueeze(0)
                blk, tau, _ = O.draft_verify_step_cpu(sd, cfg, embed, lm_head, block, th, pos, cache, start, tlogits,
                                                      hsel, 0.0)  # includes the .tolist() host sync of the acceptance
                e1.record()
                torch.cuda.synchronize()
                wall = (time.perf_counter() - t0) * 1e6
                if it >= 20:
                    times.append((e0.elapsed_time(e1) * 1e3, wall))
                tau = ks[it % len(ks)] + 1
                th = torch.cat(hsel, dim=-1)[:, :tau, :]
                start += tau
                block = torch.full((1, bs), dims["mask_token_id"], dtype=torch.long, device=dev)
                block[0, 0] = 1
        ev = sorted(t[0] for t in times)
        wl = sorted(t[1] for t in times)
        results[impl] = dict(step_us_median=ev[len(ev) // 2], step_us_p10=ev[len(ev) // 10], step_us_p90=ev[len(ev) * 9 // 10],
                             wall_us_median=wl[len(wl) // 2], final_cache_len=cache.get_seq_length())
        print(impl, results[impl], flush=True)
    out = dict(what="torch eager op sequence of the reference's draft+verify step (oracle port) on B200, bf16, "
                    "Qwen3-8B + DFlash-b16 dims, batch 1, 200 steps after 20 warm-up",
               gpu=torch.cuda.get_device_name(0), torch=torch.__version__, results=results)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ref_eager_gpu_r1.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
