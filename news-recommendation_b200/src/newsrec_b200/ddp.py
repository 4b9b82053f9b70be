"""Data-parallel plumbing: impressions shard by batch across ranks (one process per GPU); the only
collective on the path is ONE NCCL all-reduce of a flat fp32 gradient buffer per step.

`FlatGradients` makes every parameter's `.grad` a view into one contiguous buffer, so the all-reduce
needs no packing copies and autograd accumulates straight into the communication buffer.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).
    Returns (rank, world_size, local_rank).  Single-process runs need no process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of `n_items` impressions for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatGradients:
    def __init__(self, params, world: int = 1):
        self.params = [p for p in params if p.requires_grad]
        seen, uniq = set(), []
        for p in self.params:  # tied parameters (NAML's shared embeddings) appear once
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        self.params = uniq
        self.world = world
        pad4 = lambda n: (n + 3) // 4 * 4  # every view starts on a 16-byte boundary (the kernels reduce with 16-byte vectors)
        total = sum(pad4(p.numel()) for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += pad4(n)

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        """sum over ranks, then 1/world: rank-local losses are batch means, so this equals the gradient of
        the mean loss over the global batch (reference semantics of a single process on the whole batch)."""
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / self.world)
