"""Data-parallel plumbing: one process per GPU, torch.distributed only to bootstrap (rank discovery,
broadcast of the NCCL unique id, barriers); gradients move through the engine's own NCCL communicator
(csrc/comm.cuh).  The reference has no multi-GPU path; semantics are the usual DP ones:
each rank steps on its own shard of the global batch with *local* batch-norm statistics, gradients are
averaged, every rank applies the identical optimiser update.  `enable_sync_bn` switches the statistics to the
global batch: world ranks x local batch n then compute what one device computes at batch world*n."""
from __future__ import annotations

import os
from typing import Tuple


def env_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Patches [lo, hi) of a global batch owned by `rank` (equal shards, like tf.data batch(drop_remainder))."""
    if global_batch % world:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


def batch_seed(step: int, rank: int) -> int:
    """Dropout seed per (step, rank): ranks must draw different masks, steps must not repeat."""
    return (step << 16) ^ (rank * 0x9E3779B1 & 0xFFFFFFFF)


def init_engine_comm(engine, backend: str = "nccl"):
    """Create the engine's NCCL communicator; the id travels over the torch.distributed process group."""
    import torch.distributed as dist
    rank, world, _ = env_world()
    if world == 1:
        return rank, world
    if not dist.is_initialized():
        dist.init_process_group(backend)
    uid = [engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    engine.comm_init(rank, world, uid[0])
    return rank, world


def broadcast_params(engine, src: int = 0):
    """Make every rank start from rank `src`'s variables (weights are drawn per process otherwise)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    for name in engine.variables():
        t = torch.from_numpy(np.ascontiguousarray(engine.get_param(name)))
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=src)
        engine.set_param(name, t.cpu().numpy())


def enable_sync_bn(engine, on: bool = True):
    """Synchronised batch norm.  With the NCCL process group the engine exchanges the statistic rows itself
    (vnb_comm_sync_bn, after init_engine_comm); with gloo (CPU tests) they travel through torch.distributed."""
    import torch
    import torch.distributed as dist
    _, world, _ = env_world()
    if world == 1 or not on:
        if not on and dist.is_initialized() and dist.get_backend() == "nccl":
            engine.comm_sync_bn(False)
        else:
            engine.set_stats_allreduce(None, 1)
        return
    if dist.get_backend() == "nccl":
        engine.comm_sync_bn(True)
    else:
        engine.set_stats_allreduce(lambda row: dist.all_reduce(torch.from_numpy(row)), world)
