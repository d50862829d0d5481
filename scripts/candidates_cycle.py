"""Whole cycle at Qwen3-8B dims with a random-init HF target replayed from CUDA graphs: plain verify (batch 1) vs
multi-candidate verify (batch 4 over a batch-4 static cache).  python scripts/candidates_cycle.py   (GPU box)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dflash_b200 import dflash_generate_candidates  # noqa: E402
from transformers import Qwen3Config, Qwen3ForCausalLM  # noqa: E402

dev = torch.device("cuda:0")
dims = bench.Q8
draft, eng0, embed, lm_head = bench.build_engine(dims, dev, seed=0)
eng0.close()
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
target.model.embed_tokens.weight.data = embed
target.lm_head.weight.data = lm_head
prompt = torch.randint(0, dims["vocab"] - 1, (1, 128), device=dev, generator=torch.Generator(device=dev).manual_seed(7))
n_new = 96

draft.spec_generate(target, prompt, 16, None, 0.0, graph_target=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
draft.spec_generate(target, prompt, n_new, None, 0.0, graph_target=True)
torch.cuda.synchronize()
t1 = time.perf_counter()
c1 = len(draft.last_acceptance_lengths)
print(f"plain verify, graphed target: {(t1 - t0) / c1 * 1e3:.2f} ms/cycle over {c1} cycles", flush=True)
draft.release_engine()
draft._graphed_target = None
torch.cuda.empty_cache()

dflash_generate_candidates(draft, target, prompt, draft.mask_token_id, 16, 16, None, graph_target=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
res = dflash_generate_candidates(draft, target, prompt, draft.mask_token_id, n_new, 16, None, graph_target=True)
torch.cuda.synchronize()
t1 = time.perf_counter()
c2 = len(res.acceptance_lengths)
print(f"4 candidates per cycle, graphed batch-4 target: {(t1 - t0) / c2 * 1e3:.2f} ms/cycle over {c2} cycles "
      f"(includes the batch-4 prompt pass of the target)", flush=True)
