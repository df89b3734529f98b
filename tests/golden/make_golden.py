"""Generates tests/golden/*.npz from the CPU oracle (oracle/ref_vnet.py) with fixed seeds.

The reference has no golden vectors and TensorFlow cannot run here (parity unpinned),
so these fixtures pin the *oracle* (regression guard) and give the GPU tests vectors that do not need
the oracle at run time.  Parameters are never stored: they are regenerated from PCG64(42).
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_vnet as R  # noqa: E402
from vnet_tensorflow_b200.synthetic import synth_batch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (spec kwargs, P, N, loss, weights)
    "tiny_m1_k2": (dict(num_classes=2, in_channels=1, num_channels=16, num_levels=2, num_convolutions=(1, 2),
                        bottom_convolutions=2), 16, 2, "weighted_sorensen", (0.1, 1.0)),
    "tiny_m2_k3": (dict(num_classes=3, in_channels=2, num_channels=16, num_levels=2, num_convolutions=(3, 1),
                        bottom_convolutions=1), 16, 1, "mixed_weighted_jaccard", (0.01, 0.1, 1.0)),
    "default_m1_k2_p32": (dict(num_classes=2, in_channels=1), 32, 1, "weighted_sorensen", (0.1, 1.0)),
}


def perturbed_params(spec, seed=42):
    """Xavier weights from PCG64(seed) plus non-trivial gamma/beta/alpha so every BN term is exercised."""
    p = R.init_params(spec, seed)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    for k in p:
        if k.endswith(("gamma", "alpha")):
            p[k] = (p[k] * rng.uniform(0.5, 1.5, p[k].shape)).astype(np.float32)
        if k.endswith(("beta", "biases")):
            p[k] = rng.normal(0, 0.3, p[k].shape).astype(np.float32)
    return p


def main():
    torch.manual_seed(0)
    for name, (kw, P, N, loss, weights) in CASES.items():
        spec = R.VNetSpec(**kw)
        params = perturbed_params(spec)
        img, lab = synth_batch(0, N, P, spec.in_channels, spec.num_classes)
        l, logits, grads, upd = R.loss_and_grads(params, img, lab, spec, loss, weights)
        out = {"loss": np.float32(l), "logits": logits.numpy(), "argmax": R.predict(logits).numpy().astype(np.int8),
               "dice_terms": R.dice_terms(logits, torch.from_numpy(lab), "jaccard" if "jaccard" in loss else "sorensen").numpy()}
        for k, g in grads.items():
            g = g.numpy()
            out["gnorm/" + k] = np.float32(np.sqrt((g.astype(np.float64) ** 2).sum()))
            out["gsample/" + k] = g.reshape(-1)[:: max(1, g.size // 64)][:64].copy()
        for k, u in upd.items():
            out["moving/" + k] = u.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "loss", float(l), "logits", logits.shape)


if __name__ == "__main__":
    main()
