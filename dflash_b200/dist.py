"""Request-level data parallelism (the only parallelism the reference has: benchmark.py:445,
distributed.py:18-83). One process per GPU, each with a full target + draft replica; requests are
dealt round-robin; nothing crosses GPUs during decoding. At the end of a run the per-request accepted
lengths and token streams are all-gathered as fixed-shape integer tensors over NCCL (NVLink/NVSwitch)
instead of the reference's pickled `gather_object`.
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def init(backend: str | None = None) -> Tuple[int, int, int]:
    """env:// init like distributed.py:18-22 (silently single-process when RANK is unset).
    Returns (rank, world_size, local_rank)."""
    if "RANK" not in os.environ:
        return 0, 1, 0
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, init_method="env://")
    return rank, world, local


def is_main() -> bool:
    return (not dist.is_initialized()) or dist.get_rank() == 0


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin request assignment: range(rank, N, world) (benchmark.py:445)."""
    return list(range(rank, n_items, world))


def pack_streams(n_out: torch.Tensor, tokens: torch.Tensor, taus: torch.Tensor) -> torch.Tensor:
    """One int32 row per request: [n_out | tokens (max_new) | taus (max_cyc)] -- the whole result of a request in a
    single fixed-shape buffer, so the end-of-generate exchange is ONE collective (token ids fit int32)."""
    return torch.cat([n_out.view(-1, 1).to(torch.int32), tokens.to(torch.int32), taus.to(torch.int32)], dim=1).contiguous()


def unpack_streams(packed: torch.Tensor, max_new: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    return (packed[:, 0].contiguous(), packed[:, 1:1 + max_new].to(torch.int64).contiguous(),
            packed[:, 1 + max_new:].contiguous())


def all_gather_packed(packed: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """The data path's only collective: all-gather of the packed per-request rows [B_local, W] -> [world, B_local, W]
    (NCCL over NVLink on GPUs; replaces the reference's pickled `dist.gather_object`, distributed.py:66-83)."""
    world = world_size()
    if out is None:
        out = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=packed.device)
    if world == 1:
        out[0].copy_(packed)
    else:
        dist.all_gather_into_tensor(out.view(-1, packed.shape[-1]), packed)
    return out


def gather_streams(n_out: torch.Tensor, tokens: torch.Tensor, taus: torch.Tensor,
                   n_items: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """All-gather the per-request results of every rank and put them back in global request order.

    n_out  int32 [B_local]            generated tokens per local request
    tokens int64 [B_local, max_new]   generated token streams (padded)
    taus   int32 [B_local, max_cyc]   acceptance lengths per cycle (0-padded)
    Local request j of rank r is global request r + j * world (see shard_indices). Ranks that hold fewer
    requests than ceil(N / world) pad with rows of zeros. ONE collective (`all_gather_packed`). Returns tensors of
    leading size n_items.
    """
    world = world_size()
    if world == 1:
        return n_out[:n_items], tokens[:n_items], taus[:n_items]
    per = (n_items + world - 1) // world
    max_new = tokens.shape[1]
    packed = pack_streams(n_out, tokens, taus)
    if packed.shape[0] != per:
        pad = torch.zeros((per, packed.shape[1]), dtype=packed.dtype, device=packed.device)
        pad[: packed.shape[0]] = packed
        packed = pad
    buf = all_gather_packed(packed)  # [world, per, W]; row (r, j) is global request r + j * world
    g = buf.transpose(0, 1).reshape(world * per, -1)[:n_items]
    return unpack_streams(g, max_new)


def generate_data_parallel(local_generate, prompts: Sequence[torch.Tensor], max_new_tokens: int, device,
                           max_cycles: int = 0):
    """Request-level DP over the ranks of the job (benchmark.py:445,539-551 without the pickling).

    Every rank passes the SAME global prompt list. `local_generate(my_prompts)` decodes this rank's share
    (round-robin: `shard_indices`) and returns `(outputs, taus)`: `outputs[j]` = LongTensor[1, P_j + n_j] and
    `taus[j]` = acceptance lengths per cycle -- `DFlashDraftModel.spec_generate_batch` partially applied is the
    intended callee. One all-gather of fixed-shape integer tensors then gives every rank all results:
    returns `(n_out int32 [N], tokens int64 [N, max_new_tokens], taus int32 [N, max_cycles])` in prompt order."""
    n_items = len(prompts)
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = world_size()
    mine = shard_indices(n_items, rank, world)
    outs, taus = local_generate([prompts[i] for i in mine]) if mine else ([], [])
    max_cycles = int(max_cycles) if max_cycles > 0 else max_new_tokens
    n_out = torch.zeros(len(mine), dtype=torch.int32, device=device)
    tokens = torch.zeros(len(mine), max_new_tokens, dtype=torch.int64, device=device)
    tau_t = torch.zeros(len(mine), max_cycles, dtype=torch.int32, device=device)
    for j, i in enumerate(mine):
        P = int(prompts[i].shape[1])
        gen = outs[j][0, P:]
        n_out[j] = gen.numel()
        tokens[j, : gen.numel()] = gen.to(device)
        t = list(taus[j])[:max_cycles]
        if t:
            tau_t[j, : len(t)] = torch.tensor(t, dtype=torch.int32)
    return gather_streams(n_out, tokens, tau_t, n_items)


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(x: float, device) -> float:
    if not dist.is_initialized():
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float, device) -> float:
    if not dist.is_initialized():
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
