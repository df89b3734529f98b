"""Shared helpers of the test-suite (golden fixtures, engine construction, comparison metrics)."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)
CASES = make_golden.CASES
perturbed_params = make_golden.perturbed_params


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def engine_for(spec, P, N, loss, weights, library, precision="fp32", **kw):
    from vnet_tensorflow_b200.engine import VNetEngine
    return VNetEngine(num_classes=spec.num_classes, in_channels=spec.in_channels, patch_shape=(P, P, P), max_batch=N,
                      num_channels=spec.num_channels, num_levels=spec.num_levels,
                      num_convolutions=spec.num_convolutions, bottom_convolutions=spec.bottom_convolutions,
                      precision=precision, loss=loss, loss_weights=weights, library=library,
                      flavour=getattr(spec, "flavour", "networks"), **kw)


def rel_err(a, b):
    """max |a-b| / max |b|  (the 'relative fp32' tolerance of BASELINE.json is on the tensor scale)."""
    b = np.asarray(b, np.float64)
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def analytically_zero(name, spec):
    """Variables whose gradient is exactly 0 in the reference graph, where TF/torch autodiff only
    produces rounding noise: conv biases, and every gamma/beta of a batch norm whose
    output is discarded or re-normalised (the dead BN, and betas of all but the last BN of a chain)."""
    if name.endswith("/biases"):
        return True
    if getattr(spec, "flavour", "networks") == "legacy":
        # VNet.py: whenever a conv has two batch norms (BN -> [+ residual] -> BN) the first BN's beta only shifts
        # the second BN's input by a constant and is normalised away; only decoder conv_1 with n > 1 has one BN
        if not name.endswith("/batch_normalization/beta") or "/conv_" not in name:
            return False
        parts = name.split("/")
        i = int(parts[-3].split("_")[1]) - 1
        if parts[1] != "decoder":
            return True
        n = spec.num_convolutions[int(parts[2].split("_")[1]) - 1]
        return not (i == 0 and n > 1)
    if "/decoder/" not in name or "batch_normalization" not in name:
        return False
    level = int(name.split("/decoder/level_")[1].split("/")[0]) - 1
    n = spec.num_convolutions[level]
    conv = name.split("/")[3]
    if not conv.startswith("conv_"):
        return False
    i = int(conv.split("_")[1]) - 1
    bn = name.split("/")[4]
    k = 0 if bn == "batch_normalization" else int(bn.rsplit("_", 1)[1])
    is_beta = name.endswith("/beta")
    if n == 1:                       # CH_T: only beta of BN2 is live
        return is_beta and k < 2
    if i == 0:
        return False                 # plain BN
    if i == n - 1:                   # CH_Q: beta of BN0 is re-normalised away
        return is_beta and k == 0
    return k == 0                    # CH_D: BN0 entirely dead


def perturbed_attention_params(spec, nch, seed=42, weight_scale=1.0):
    """V-Net + attention / output module variables with every term exercised: perturbed
    gamma / beta / alpha / biases, and non-trivial *moving* statistics in the modules (their batch norms run in
    inference mode, train.py:538-540).  `weight_scale` shrinks the sigma=0.1 truncated-normal module weights so
    that a 64-channel module stays well conditioned (the reference init grows activations ~4x per conv)."""
    from oracle import ref_vnet as R
    p = R.init_attention_params(spec, seed, seed + 1, nch)
    base = perturbed_params(spec, seed)
    for k in base:
        p[k] = base[k]
    rng = np.random.Generator(np.random.PCG64(seed + 2))
    for name, shape, kind in R.attention_param_specs(spec, nch)[len(R.param_specs(spec)):]:
        if kind == "mod_w":
            p[name] = (p[name] * weight_scale).astype(np.float32)
        elif kind in ("mod_b", "beta", "moving_mean"):
            p[name] = rng.normal(0, 0.2, shape).astype(np.float32)
        elif kind in ("gamma", "moving_variance"):
            p[name] = rng.uniform(0.6, 1.4, shape).astype(np.float32)
    return p


def argmax_parity(argmax, logits, ref_logits, ref_argmax=None):
    """Label-volume parity with the oracle (north_star: bit-exact argmax).  Two fp32 evaluations of the same graph that
    differ only in summation order (the oracle on oneDNN vs the CUDA engine; the oracle itself in fp32 vs fp64; TensorFlow
    with another thread count) cannot agree on argmax at a voxel whose two top logits are closer than their own rounding
    error, so the bar is: every mismatch lies inside the rounding band |top1 - top2| <= 2 * max|logits - ref| of the
    oracle's logits, and none outside it.  Returns (mismatches, mismatches outside the band, band width)."""
    ref_logits = np.asarray(ref_logits, np.float64)
    if ref_argmax is None:
        ref_argmax = np.argmax(ref_logits, -1)
    flips = np.asarray(argmax) != np.asarray(ref_argmax)
    band = 2.0 * float(np.abs(np.asarray(logits, np.float64) - ref_logits).max())
    top = np.sort(ref_logits, -1)
    margin = top[..., -1] - top[..., -2]
    return int(flips.sum()), int((flips & (margin > band)).sum()), band


def assert_argmax_parity(argmax, logits, ref_logits, ref_argmax=None, max_frac=1e-4):
    n, outside, band = argmax_parity(argmax, logits, ref_logits, ref_argmax)
    assert outside == 0, "%d argmax mismatches outside the rounding band (%.3g)" % (outside, band)
    assert n <= max(1, int(max_frac * np.asarray(argmax).size)), "%d argmax mismatches" % n
    return n
