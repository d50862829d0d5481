"""Where one batched verify forward of the HF target goes (BatchedVerifyTarget, R streams): torch profiler kernel table.
   python scripts/batched_target_profile.py [R] [kv_len]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from transformers import Qwen3Config, Qwen3ForCausalLM
from dflash_b200.target_graph import BatchedVerifyTarget

R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kv = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dims, dev, bs = bench.Q8, torch.device("cuda:0"), 16
cfg = Qwen3Config(vocab_size=dims["vocab"], hidden_size=dims["hidden"], intermediate_size=dims["intermediate"],
                  num_hidden_layers=dims["target_layers"], num_attention_heads=dims["heads"],
                  num_key_value_heads=dims["kv_heads"], head_dim=dims["head_dim"], max_position_embeddings=40960,
                  rms_norm_eps=dims["eps"], tie_word_embeddings=False,
                  rope_parameters={"rope_type": "default", "rope_theta": dims["rope_theta"]})
cfg._attn_implementation = "sdpa"
torch.set_default_dtype(torch.bfloat16)
with torch.device(dev):
    target = Qwen3ForCausalLM(cfg).eval()
torch.set_default_dtype(torch.float32)
start = torch.full((R,), 140, dtype=torch.int32, device=dev)
blocks = torch.randint(0, 1000, (R, bs), device=dev)
bt = BatchedVerifyTarget(target, bs, R, 2304, bench.build_target_layer_ids(36, 5) if hasattr(bench, "build_target_layer_ids") else [1, 9, 17, 25, 33], start, blocks)
with torch.inference_mode():
    for r in range(min(R, 2)):
        bt.prefill(r, torch.randint(0, 1000, (1, 128), device=dev))
    for use_graph in (False, True):
        bt.use_graph = use_graph
        for _ in range(3):
            bt.verify_forward(kv)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            bt.verify_forward(kv)
        torch.cuda.synchronize()
        print(f"R={R} kv_len={kv} graph={use_graph}: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per forward", flush=True)
    bt.use_graph = False
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        bt.verify_forward(kv)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=90))
