"""Data-parallel host logic on CPU: two gloo ranks, each with its own (emulated) engine, shard a global
batch of whole patches, exchange gradients, and apply the identical update -- the protocol the NCCL
communicator (csrc/comm.cuh) implements on the GPUs, with the all-reduce done by torch.distributed here."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import EMUL_LIB


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ref_vnet as R
    from tests.helpers import analytically_zero, engine_for, perturbed_params
    from vnet_tensorflow_b200 import _ffi, distributed as D
    from vnet_tensorflow_b200.synthetic import synth_batch
    lib = _ffi.Library(EMUL_LIB)
    spec = R.VNetSpec(num_classes=2, in_channels=1, num_channels=4, num_levels=1, num_convolutions=(2,), bottom_convolutions=1)
    P, G = 8, 4
    lo, hi = D.shard_range(G, rank, world)
    assert (lo, hi) == (rank * 2, rank * 2 + 2)
    img, lab = synth_batch(0, G, P, 1, 2)           # every rank builds the same global batch, keeps its shard
    eng = engine_for(spec, P, hi - lo, "weighted_sorensen", (0.1, 1.0), lib, learning_rate=1e-2)
    if rank == 0:
        eng.set_params(perturbed_params(spec))
    D.broadcast_params(eng, src=0)                   # identical start on every rank
    params0 = eng.get_params()
    loss = eng.forward_backward(img[lo:hi], lab[lo:hi], update_moving_stats=True)
    _, _, g_ref, _ = R.loss_and_grads(params0, img[lo:hi], lab[lo:hi], spec, "weighted_sorensen", (0.1, 1.0))
    grads = eng.get_grads()
    for k, v in grads.items():                       # local gradient == oracle gradient of the local shard
        if not analytically_zero(k, spec):
            assert np.abs(v - g_ref[k].numpy()).max() <= 5e-4 * max(np.abs(g_ref[k].numpy()).max(), 1e-6), k
    for k, v in grads.items():                       # all-reduce (mean) -- what Comm::ring_allreduce + 1/world do
        t = torch.from_numpy(v.copy())
        dist.all_reduce(t)
        eng.set_param(k, (t / world).numpy(), _ffi.SLOT_GRAD)
    eng.apply_gradients()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), loss=loss, **{k.replace("/", "|"): eng.get_param(k) for k, (_, tr) in eng.variables().items() if tr})
    assert eng.global_step == 1
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_data_parallel_step(emul_lib, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert float(a["loss"]) != float(b["loss"])      # different shards
    for k in a.files:
        if k != "loss":
            assert np.array_equal(a[k], b[k]), k     # identical replicas after the step


def test_shard_and_seed_helpers():
    from vnet_tensorflow_b200 import distributed as D
    assert [D.shard_range(16, r, 8) for r in (0, 7)] == [(0, 2), (14, 16)]
    with pytest.raises(ValueError):
        D.shard_range(10, 0, 4)
    seeds = {D.batch_seed(s, r) for s in range(4) for r in range(8)}
    assert len(seeds) == 32
