"""Multi-GPU plumbing: one process per GPU, drones sharded along the batch axis, ONE sum-allreduce of the flat
gradient per train step (the loss is a sum over drones, so the summed gradient equals the single-device gradient of
the concatenated batch; no mean scaling).  NCCL on GPUs, gloo in the CPU tests."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style initialisation (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 or dist.is_initialized():
        return int(os.environ.get("RANK", "0")), world
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend)
    return dist.get_rank(), dist.get_world_size()


def shard_bounds(n_total, rank, world):
    """contiguous split of the drone axis: rank r owns [lo, hi)"""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(tensor, rank, world):
    if tensor is None:
        return None
    lo, hi = shard_bounds(tensor.shape[0], rank, world)
    return tensor[lo:hi]


def allreduce_sum_(flat_grad, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def sgd_momentum_step_(flat, grad, buf, lr, momentum=0.9):
    """identical update on every rank after the allreduce (optim.SGD semantics, train_base.py:139-143)"""
    buf.mul_(momentum).add_(grad)
    flat.add_(buf, alpha=-lr)
