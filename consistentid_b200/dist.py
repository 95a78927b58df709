"""Multi-GPU plumbing for batched generation: one process per GPU, images sharded across ranks, exactly ONE
NCCL broadcast of the flat weight arena at init and no per-step collective (the reference's inference path is
single-GPU, infer.py:10; its only collectives are training-time DDP, train_bash.sh:1-8)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).
    Returns (rank, local_rank, world_size).  No-op (0, 0, 1) outside torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 0, 1
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    if not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def broadcast_arena(arena: torch.Tensor, src=0):
    """The single init-time collective: rank ``src``'s packed weights overwrite everybody's arena."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(arena, src=src)
    return arena


def shard_batch(total_images: int, rank: int, world: int):
    """Contiguous shard [start, stop) of the global batch for this rank (images are independent: no data-path collective)."""
    per = (total_images + world - 1) // world
    start = min(rank * per, total_images)
    return start, min(start + per, total_images)


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
