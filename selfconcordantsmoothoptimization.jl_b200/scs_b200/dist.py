"""Row sharding across GPUs: one process per GPU, rows of A split into contiguous blocks.

The data path has two exchange steps per iteration — the fp64 sum all-reduce of [g ‖ loss] and of the Gram — and
both happen inside the library over NCCL.  The host only has to agree on a 128-byte NCCL unique id once; that
is done here over whatever torch.distributed backend the launcher initialised (gloo in the CPU tests).
"""
from __future__ import annotations

import os


def shard_rows(n_total: int, world: int, rank: int):
    """Contiguous row block [k*n/P, (k+1)*n/P) of rank k.  Returns (row0, n_local)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    r0 = n_total * rank // world
    r1 = n_total * (rank + 1) // world
    return r0, r1 - r0


def broadcast_unique_id(make_id, rank: int, world: int) -> bytes:
    """Rank 0 calls make_id() -> 128 bytes; everyone gets it through torch.distributed (already initialised)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return make_id()
    if rank == 0:
        raw = make_id()
        if len(raw) != 128:
            raise ValueError("unique id must be 128 bytes")
        t = torch.tensor(list(raw), dtype=torch.uint8)
    else:
        t = torch.zeros(128, dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


def context_from_env():
    """Context for this torchrun rank (RANK / LOCAL_RANK / WORLD_SIZE); single process -> world 1."""
    from .api import Context
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return Context(local)
    uid = broadcast_unique_id(Context.unique_id, rank, world)
    return Context(local, rank, world, uid)
