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


def _sync_bn_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ref_vnet as R
    from tests.helpers import engine_for, perturbed_params
    from vnet_tensorflow_b200 import _ffi, distributed as D
    from vnet_tensorflow_b200.synthetic import synth_batch
    lib = _ffi.Library(EMUL_LIB)
    # both block kinds (x + BN(x) decoder chain, residual encoder block), the tiled single-channel input layer, down / up convs
    spec = R.VNetSpec(num_classes=2, in_channels=1, num_channels=4, num_levels=2, num_convolutions=(1, 2), bottom_convolutions=1)
    P, G = 8, 2
    lo, hi = D.shard_range(G, rank, world)
    img, lab = synth_batch(3, G, P, 1, 2)
    eng = engine_for(spec, P, hi - lo, "weighted_sorensen", (0.1, 1.0), lib, learning_rate=1e-2)
    eng.set_params(perturbed_params(spec, 7))
    D.enable_sync_bn(eng)
    loss = eng.forward_backward(img[lo:hi], lab[lo:hi], update_moving_stats=True)
    out = {"loss": np.float64(loss)}
    for k, v in eng.get_grads().items():             # the gradient exchange: mean over the ranks
        t = torch.from_numpy(v.copy())
        dist.all_reduce(t)
        out["grad|" + k] = (t / world).numpy()
    for k, (_, trainable) in eng.variables().items():
        if not trainable:
            out["state|" + k] = eng.get_param(k)     # moving statistics: global-batch values on every rank
    logits = eng.forward(img[lo:hi], want_softmax=False, want_argmax=False)[0]   # inference runs on batch statistics too
    out["logits"] = logits
    np.savez(os.path.join(out_dir, "sync%d.npz" % rank), **{k.replace("/", "|"): v for k, v in out.items()})
    D.enable_sync_bn(eng, False)                     # back to local statistics: the loss changes
    assert eng.forward_backward(img[lo:hi], lab[lo:hi]) != loss
    dist.barrier()
    dist.destroy_process_group()


def test_sync_bn_two_ranks_equal_one_device_at_the_global_batch(emul_lib, tmp_path):
    """with the batch statistics (and the two backward sums) of every batch norm summed over the ranks,
    two ranks with one patch each reproduce one device with a batch of two - loss, every averaged gradient, the moving
    statistics and the logits - where local statistics do not."""
    from oracle import ref_vnet as R
    from tests.helpers import analytically_zero, engine_for, perturbed_params
    from vnet_tensorflow_b200.synthetic import synth_batch
    port = _free_port()
    mp.spawn(_sync_bn_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ranks = [np.load(tmp_path / ("sync%d.npz" % r)) for r in (0, 1)]
    spec = R.VNetSpec(num_classes=2, in_channels=1, num_channels=4, num_levels=2, num_convolutions=(1, 2), bottom_convolutions=1)
    img, lab = synth_batch(3, 2, 8, 1, 2)
    one = engine_for(spec, 8, 2, "weighted_sorensen", (0.1, 1.0), emul_lib, learning_rate=1e-2)
    one.set_params(perturbed_params(spec, 7))
    loss = one.forward_backward(img, lab, update_moving_stats=True)
    assert abs((float(ranks[0]["loss"]) + float(ranks[1]["loss"])) / 2 - loss) < 2e-6
    grads = one.get_grads()
    scale = max(float(np.abs(v).max()) for v in grads.values())
    for k, v in grads.items():
        key = ("grad|" + k).replace("/", "|")
        assert np.array_equal(ranks[0][key], ranks[1][key]), k
        if analytically_zero(k, spec):
            assert np.abs(ranks[0][key]).max() <= 1e-6 * scale
        else:
            assert np.abs(ranks[0][key] - v).max() <= 2e-5 * max(np.abs(v).max(), 1e-3 * scale), k
    n_state = 0
    for k, (_, trainable) in one.variables().items():
        if not trainable:
            key = ("state|" + k).replace("/", "|")
            assert np.allclose(ranks[0][key], one.get_param(k), rtol=1e-6, atol=1e-7), k
            assert np.array_equal(ranks[0][key], ranks[1][key]), k
            n_state += 1
    assert n_state > 10
    logits = one.forward(img, want_softmax=False, want_argmax=False)[0]
    both = np.concatenate([ranks[0]["logits"], ranks[1]["logits"]], 0)
    assert np.abs(both - logits).max() <= 1e-5 * np.abs(logits).max()
    one.close()


def test_shard_and_seed_helpers():
    from vnet_tensorflow_b200 import distributed as D
    assert [D.shard_range(16, r, 8) for r in (0, 7)] == [(0, 2), (14, 16)]
    with pytest.raises(ValueError):
        D.shard_range(10, 0, 4)
    seeds = {D.batch_seed(s, r) for s in range(4) for r in range(8)}
    assert len(seeds) == 32
