"""Multi-GPU correctness check (run under torchrun on N GPUs): after one data-parallel step every rank holds
bit-identical parameters, and the ring all-reduce result equals the NCCL all-reduce / mean of local grads."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vnet_tensorflow_b200 import _ffi
from vnet_tensorflow_b200.engine import VNetEngine
from vnet_tensorflow_b200.init import initialize
from vnet_tensorflow_b200.synthetic import synth_batch


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P, B = 32, 2
    eng = VNetEngine(num_classes=2, in_channels=1, patch_shape=(P, P, P), max_batch=B, precision="bf16x3",
                     loss="weighted_sorensen", loss_weights=(0.1, 1.0), device=local)
    initialize(eng, 42)
    uid = [eng.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.comm_init(rank, world, uid[0])
    img, lab = synth_batch(0, B, P, 1, 2, rank=rank)
    # reference: local gradients from a second, communicator-less engine, averaged with torch.distributed
    ref = VNetEngine(num_classes=2, in_channels=1, patch_shape=(P, P, P), max_batch=B, precision="bf16x3",
                     loss="weighted_sorensen", loss_weights=(0.1, 1.0), device=local)
    initialize(ref, 42)
    ref.forward_backward(img, lab, update_moving_stats=True)
    want = {}
    for k, g in ref.get_grads().items():
        t = torch.from_numpy(g).cuda()
        dist.all_reduce(t)
        want[k] = (t / world).cpu().numpy()
    loss = eng.forward_backward(img, lab, update_moving_stats=True)   # launches bucketed ring all-reduces on the side stream
    eng.apply_gradients()                                             # waits for them, Adam with 1/world
    got = eng.get_grads()                                             # summed (not yet averaged) gradients
    worst = 0.0
    for k in want:
        err = np.abs(got[k] / world - want[k]).max() / max(np.abs(want[k]).max(), 1e-12)
        worst = max(worst, err)
    # all ranks must hold identical parameters after the step
    # (trainable variables only: BN moving statistics are per-rank by design)
    names = [k for k, (_, tr) in eng.variables().items() if tr]
    digest = torch.tensor([float(np.float64(sum(float(np.abs(eng.get_param(k).astype(np.float64)).sum()) for k in names)))],
                          dtype=torch.float64, device="cuda")
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    same = all(float(g) == float(gathered[0]) for g in gathered)
    if rank == 0:
        print("DP_CHECK world=%d loss=%.5f ring-vs-nccl grad rel err=%.3e replicas identical=%s" % (world, loss, worst, same))
    dist.barrier()
    dist.destroy_process_group()
    assert worst < 1e-5 and same


if __name__ == "__main__":
    main()
