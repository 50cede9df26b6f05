"""Data-parallel plumbing (one process per GPU; torch.distributed over NCCL/NVLink, gloo on CPU for tests).

The path shards by samples only: every rank runs the same step on its own batches and the
single exchange is the gradient all-reduce.  What this module replaces / restates from the reference:
  * DDP(find_unused_parameters=True) bucketed all-reduce      pretrain_src/utils/misc.py:57-71
      -> one flat fp32 gradient arena, all-reduced in a few large async buckets (no graph traversal, no
         per-parameter hooks; parameters a task does not touch simply contribute zeros)
  * the per-step 1-int task-id broadcast from rank 0           pretrain_src/data/loader.py:56-59
      -> a seeded task schedule every rank derives locally (zero collectives)
  * DistributedSampler sharding                                pretrain_src/data/loader.py:148-150
"""
import os

import torch
import torch.distributed as dist


def init_distributed(device=None):
    """env:// rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT), like utils/distributed.py:73."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    return rank, world, local


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class FlatAllReduce:
    """Averaging all-reduce of a flat gradient buffer in `n_buckets` contiguous async pieces, issued from the
    END of the buffer backwards (the order in which backward finishes the gradients)."""

    def __init__(self, flat, bucket_bytes=64 << 20):
        self.flat = flat
        n = flat.numel()
        per = max(1, bucket_bytes // flat.element_size())
        self.bounds = []
        hi = n
        while hi > 0:
            lo = max(0, hi - per)
            self.bounds.append((lo, hi))
            hi = lo
        self.world = world_size()

    def start(self):
        """Issue the bucket all-reduces (asynchronously, on the backend's own stream)."""
        self._handles, self._avg = [], True
        if self.world == 1:
            return
        # NCCL averages inside the collective (no extra pass over the arena); gloo (CPU tests) has no AVG
        self._avg = dist.get_backend() == "nccl"
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self._handles = [dist.all_reduce(self.flat[lo:hi], op=op, async_op=True) for lo, hi in self.bounds]

    def finish(self):
        """Order the current stream after the exchange started by `start()`."""
        for h in self._handles:
            h.wait()
        if self._handles and not self._avg:
            self.flat.mul_(1.0 / self.world)
        self._handles = []

    def __call__(self):
        self.start()
        self.finish()


def broadcast_flat(flat, src=0):
    """DDP's parameter broadcast at wrap time (utils/misc.py:63-66) on the flat parameter arena."""
    if world_size() > 1:
        dist.broadcast(flat, src)


def task_schedule(seed, n_steps, tasks, ratios):
    """The MetaLoader's multinomial task sampling (data/loader.py:50-75) as a pre-agreed seeded schedule:
    every rank computes the same list, so no per-step broadcast is needed."""
    g = torch.Generator().manual_seed(int(seed))
    p = torch.tensor([float(r) for r in ratios])
    idx = torch.multinomial(p / p.sum(), n_steps, replacement=True, generator=g)
    return [tasks[i] for i in idx.tolist()]


def shard_indices(n, rank, world, seed=0, shuffle=True, drop_last=False):
    """DistributedSampler semantics (pads by wrapping so every rank gets the same count)."""
    if shuffle:
        g = torch.Generator().manual_seed(int(seed))
        order = torch.randperm(n, generator=g).tolist()
    else:
        order = list(range(n))
    if drop_last:
        total = n // world * world
        order = order[:total]
    else:
        total = (n + world - 1) // world * world
        order = order + order[: total - len(order)]
    return order[rank:total:world]
