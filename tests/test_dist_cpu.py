"""Request-level data parallelism on CPU: world_size-2 gloo run of the shard + all-gather logic
(dflash_b200/dist.py), the N>1 path of bench.py minus the GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from dflash_b200 import dist as ddist
    r, w, _ = ddist.init(backend="gloo")
    assert (r, w) == (rank, world)
    mine = ddist.shard_indices(n_items, rank, world)
    max_new, max_cyc = 6, 4
    n_out = torch.tensor([10 + g for g in mine], dtype=torch.int32)
    toks = torch.stack([torch.arange(max_new, dtype=torch.int64) + 100 * g for g in mine]) if mine else \
        torch.zeros(0, max_new, dtype=torch.int64)
    taus = torch.stack([torch.full((max_cyc,), g + 1, dtype=torch.int32) for g in mine]) if mine else \
        torch.zeros(0, max_cyc, dtype=torch.int32)
    g_n, g_t, g_a = ddist.gather_streams(n_out, toks, taus, n_items)
    t = ddist.max_over_ranks(float(rank + 1), "cpu")
    s = ddist.sum_over_ranks(float(rank + 1), "cpu")
    ddist.barrier()
    if rank == 0:
        q.put((g_n.tolist(), g_t.tolist(), g_a.tolist(), t, s))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [4, 5, 1])
def test_gather_streams_gloo_world2(n_items):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    g_n, g_t, g_a, t, s = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert g_n == [10 + g for g in range(n_items)]
    assert g_t == [[100 * g + i for i in range(6)] for g in range(n_items)]
    assert g_a == [[g + 1] * 4 for g in range(n_items)]
    assert t == 2.0 and s == 3.0


def test_shard_indices_round_robin():
    from dflash_b200.dist import shard_indices
    assert shard_indices(10, 1, 4) == [1, 5, 9]  # benchmark.py:445: range(rank, N, world)
    assert sorted(sum((shard_indices(7, r, 3) for r in range(3)), [])) == list(range(7))
    assert shard_indices(2, 3, 4) == []


def _dp_worker(rank, world, port, n_items, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from dflash_b200 import dist as ddist
    ddist.init(backend="gloo")
    prompts = [torch.arange(3 + i, dtype=torch.int64).unsqueeze(0) for i in range(n_items)]

    def local_generate(mine):  # stand-in for draft.spec_generate_batch: prompt i generates i % 3 + 1 tokens 7, 8, ...
        outs, taus = [], []
        for p in mine:
            P = p.shape[1]
            n = (P - 3) % 3 + 1
            outs.append(torch.cat([p, 7 + torch.arange(n, dtype=torch.int64).unsqueeze(0)], dim=1))
            taus.append([n])
        return outs, taus

    g_n, g_t, g_a = ddist.generate_data_parallel(local_generate, prompts, 4, "cpu", max_cycles=2)
    if rank == 0:
        q.put((g_n.tolist(), g_t.tolist(), g_a.tolist()))
    ddist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [5, 2])
def test_generate_data_parallel_gloo_world2(n_items):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    g_n, g_t, g_a = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp_n = [i % 3 + 1 for i in range(n_items)]
    assert g_n == exp_n
    assert g_t == [[7 + j if j < n else 0 for j in range(4)] for n in exp_n]
    assert g_a == [[n, 0] for n in exp_n]
