"""Draft step with the plain argmax epilogue vs the top-4 epilogue + candidate blocks (Qwen3-8B dims, 1 stream).
python scripts/candidates_timing.py   (GPU box)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dflash_b200.engine import DraftEngine  # noqa: E402

dev = torch.device("cuda:0")
draft, eng0, embed, lm_head = bench.build_engine(bench.Q8, dev, seed=0)
eng0.close()
eng = DraftEngine(draft, embed, lm_head, max_seq=2304, out_len=2304, max_requests=1, block_size=16, max_candidates=4)
g = torch.Generator(device=dev).manual_seed(1)
H, V = 4096, bench.Q8["vocab"]
hs = [(torch.randn(128, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(5)]
eng.reset_request(0, torch.randint(0, V - 1, (128,), device=dev, generator=g), 1, 2048)
eng.prefill_context(0, hs)


def timed(fn, n=60):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n


mask = eng.block_ids.clone()


def plain():
    eng.block_ids.copy_(mask)
    eng.draft_step()


def cands():
    eng.block_ids.copy_(mask)
    eng.draft_step_candidates(4, 2)


def sampled():
    eng.block_ids.copy_(mask)
    eng.draft_step_sampled(1.0, 7)


print(f"draft step, Gumbel-max sampling epilogue (T = 1): {timed(sampled):.1f} us")
print(f"draft step, argmax epilogue: {timed(plain):.1f} us; top-4 epilogue + 4 candidate blocks: {timed(cands):.1f} us")
