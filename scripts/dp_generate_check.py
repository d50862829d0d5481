"""Multi-GPU check of the request-level DP driver with real NCCL:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/dp_generate_check.py
Every rank decodes its round-robin share of the prompts with `spec_generate_batch` (tiny seeded pair, greedy), the
results are all-gathered, and rank 0 compares them with decoding ALL prompts locally."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from dflash_b200 import DFlashDraftModel, dist as ddist
    from tests.tiny_models import TINY, build_pair, rig_lm_head
    rank, world, local = ddist.init()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    target, draft = build_pair(DFlashDraftModel, seed=1234, block_size=16, dtype=torch.bfloat16, device=dev)
    rig_lm_head(target, live=(3, 17, 101, 250, 251, 400, 512, 640, 777, 801, 900, 998), seed=99)
    g = torch.Generator().manual_seed(5)
    n, n_new = 11, 32
    prompts = [torch.randint(0, TINY["vocab"] - 1, (1, 8 + 3 * i), generator=g).to(dev) for i in range(n)]

    def local_generate(mine):
        outs = draft.spec_generate_batch(target, mine, n_new, None, 0.0, max_requests=4)
        return outs, draft.last_batch_acceptance_lengths

    n_out, tokens, taus = ddist.generate_data_parallel(local_generate, prompts, n_new, dev, max_cycles=n_new)
    torch.cuda.synchronize()
    if rank == 0:
        ref, ref_taus = local_generate(prompts)
        ok = True
        for i in range(n):
            gen = ref[i][0, prompts[i].shape[1]:]
            ok &= int(n_out[i]) == gen.numel() and torch.equal(tokens[i, :gen.numel()], gen)
            ok &= taus[i, :len(ref_taus[i])].tolist() == list(ref_taus[i])
        print(f"dp_generate_check world={world}: {'OK' if ok else 'MISMATCH'} "
              f"({n} prompts, {int(n_out.sum())} tokens, mean tau {float(taus[taus > 0].float().mean()):.2f})", flush=True)
        if not ok:
            sys.exit(1)
    ddist.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
