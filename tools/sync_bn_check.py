"""Synchronised batch norm on real GPUs (run under torchrun on N GPUs): N ranks with batch B each and
`vnb_comm_sync_bn` against one engine on rank 0's GPU that sees the whole global batch of N*B patches.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/sync_bn_check.py [--precision fp32|bf16x3|bf16] [--patch 32]

Prints one SYNC_BN_CHECK line (rank 0) with the worst relative deviation of the loss, the averaged gradients, the
moving statistics and the logits, and the per-step cost of the statistic exchanges.  The CPU counterpart of this
check is tests/test_dp_gloo.py::test_sync_bn_two_ranks_equal_one_device_at_the_global_batch.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vnet_tensorflow_b200.engine import VNetEngine
from vnet_tensorflow_b200.init import initialize
from vnet_tensorflow_b200.synthetic import synth_batch


def _engine(P, B, precision, device):
    eng = VNetEngine(num_classes=2, in_channels=1, patch_shape=(P, P, P), max_batch=B, precision=precision,
                     loss="weighted_sorensen", loss_weights=(0.1, 1.0), device=device)
    initialize(eng, 42)
    return eng


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16x3", "bf16"])
    ap.add_argument("--patch", type=int, default=32)
    ap.add_argument("--batch", type=int, default=1)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P, B = args.patch, args.batch
    img, lab = synth_batch(0, B * world, P, 1, 2)          # the same global batch on every rank
    lo, hi = rank * B, (rank + 1) * B
    eng = _engine(P, B, args.precision, local)
    uid = [eng.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.comm_init(rank, world, uid[0])
    eng.comm_sync_bn(True)
    loss = eng.forward_backward(img[lo:hi], lab[lo:hi], update_moving_stats=True)
    # logits with the step's own weights (apply_gradients below runs the optimiser), statistics of the global batch
    logits = eng.forward(img[lo:hi], want_softmax=False, want_argmax=False)[0]
    eng.apply_gradients()                                   # exchanges the gradient buckets; gradients now hold the sums
    grads = {k: v / world for k, v in eng.get_grads().items()}
    losses = torch.tensor([loss], dtype=torch.float64, device="cuda")
    dist.all_reduce(losses)
    gathered = [torch.zeros(logits.shape, dtype=torch.float32, device="cuda") for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(logits).cuda())
    # cost of the exchanges: the same step with and without them
    times = {}
    for mode in (True, False):
        eng.comm_sync_bn(mode)
        for _ in range(2):
            eng.forward_backward(img[lo:hi], lab[lo:hi])
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            eng.forward_backward(img[lo:hi], lab[lo:hi])
        eng.sync()
        times[mode] = (time.perf_counter() - t0) / 5 * 1e3
    ok = True
    if rank == 0:
        one = _engine(P, B * world, args.precision, local)
        ref_loss = one.forward_backward(img, lab, update_moving_stats=True)
        ref_grads = one.get_grads()
        scale = max(float(np.abs(v).max()) for v in ref_grads.values())
        g_err = max(float(np.abs(grads[k] - v).max()) / max(float(np.abs(v).max()), 1e-3 * scale) for k, v in ref_grads.items())
        s_err = 0.0
        for k, (_, trainable) in one.variables().items():
            if not trainable:
                a, b = eng.get_param(k), one.get_param(k)
                s_err = max(s_err, float(np.abs(a - b).max()) / max(float(np.abs(b).max()), 1e-6))
        ref_logits = one.forward(img, want_softmax=False, want_argmax=False)[0]
        both = torch.cat(gathered, 0).cpu().numpy()
        l_err = float(np.abs(both - ref_logits).max()) / float(np.abs(ref_logits).max())
        loss_err = abs(float(losses) / world - ref_loss)
        tol = {"fp32": 1e-4, "bf16x3": 2e-3, "bf16": 0.2}[args.precision]
        ok = loss_err < 1e-5 * (1 if args.precision != "bf16" else 1e3) and g_err < 50 * tol and s_err < tol and l_err < tol
        print("SYNC_BN_CHECK world=%d precision=%s patch=%d loss err=%.2e grad err=%.2e moving-stat err=%.2e logit err=%.2e "
              "step ms sync=%.2f local=%.2f ok=%s" % (world, args.precision, P, loss_err, g_err, s_err, l_err,
                                                      times[True], times[False], ok))
    dist.barrier()
    dist.destroy_process_group()
    assert ok


if __name__ == "__main__":
    main()
