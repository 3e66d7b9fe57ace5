"""Data-parallel plumbing: the role tf.distribute.MirroredStrategy plays in the reference
(main.py:22, vangan.py:86,459-490), as one process per GPU over torch.distributed.

The train step shards by batch only (InstanceNorm is per sample), so the one exchange step is the
SUM all-reduce of each network's flat gradient buffer — issued right after that network's backward
sweep so it overlaps the next network's backward — plus one 10-float all-reduce of the result dict.
"""
import contextlib
import os

import torch
import torch.distributed as dist


class Strategy:
    """Minimal MirroredStrategy look-alike: `.num_replicas_in_sync`, `.scope()`, `.run()`, `.reduce()`."""

    def __init__(self, group=None):
        self.group = group
        self.enabled = dist.is_available() and dist.is_initialized()
        self.num_replicas_in_sync = dist.get_world_size(group) if self.enabled else 1
        self.rank = dist.get_rank(group) if self.enabled else 0

    @contextlib.contextmanager
    def scope(self):
        yield self

    def run(self, fn, args=(), kwargs=None):
        return fn(*args, **(kwargs or {}))

    def all_reduce_async(self, tensor):
        """SUM all-reduce; returns a handle with .wait() (None when single replica)."""
        if not self.enabled or self.num_replicas_in_sync == 1:
            return None
        return dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def reduce(self, op, value, axis=None):
        """strategy.reduce(ReduceOp.SUM, v, axis=None) on a tensor."""
        if self.enabled and self.num_replicas_in_sync > 1:
            dist.all_reduce(value, op=dist.ReduceOp.SUM, group=self.group)
        return value


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*).
    Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local
