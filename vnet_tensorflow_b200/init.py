"""Variable initialisers of the reference graph (layers2.py:4-30, 60-61, 98; tf.layers BN defaults).

The reference draws Xavier weights from NumPy's unseeded global RNG at graph-construction time; here
the draw is made reproducible with PCG64(seed), consumed in variable-creation order.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np


def xavier_uniform(shape, rng: np.random.Generator) -> np.ndarray:
    """layers2.py:16-21: U(-lim, lim), lim = sqrt(6 / (prod(spatial) * (Cin + Cout)))."""
    s = len(shape) - 2
    num_activations = np.prod(shape[:s]) * np.sum(shape[s:])
    lim = np.sqrt(6.0 / num_activations)
    return rng.uniform(-lim, lim, size=tuple(shape)).astype(np.float32)


def truncated_normal(shape, stddev: float, rng: np.random.Generator) -> np.ndarray:
    """attention.py:25-27 tf.truncated_normal: N(0, stddev), values beyond two sigma are redrawn."""
    w = rng.normal(0.0, stddev, size=tuple(shape))
    bad = np.abs(w) > 2.0 * stddev
    while bad.any():
        w[bad] = rng.normal(0.0, stddev, size=int(bad.sum()))
        bad = np.abs(w) > 2.0 * stddev
    return w.astype(np.float32)


def initial_values(variables, seed: int = 42) -> "OrderedDict[str, np.ndarray]":
    """variables: name -> (shape, trainable) as returned by VNetEngine.variables()."""
    rng = np.random.Generator(np.random.PCG64(seed))
    mod_rng = np.random.Generator(np.random.PCG64(seed + 1))  # attention / output modules draw from their own stream
    out = OrderedDict()
    for name, (shape, _) in variables.items():
        if name.endswith("/weights"):
            out[name] = xavier_uniform(shape, rng)
        elif "/Variable" in name and len(shape) == 5:
            out[name] = truncated_normal(shape, 0.1, mod_rng)
        elif name.endswith(("/gamma", "/moving_variance")):
            out[name] = np.ones(shape, np.float32)
        elif name.endswith("/alpha"):
            out[name] = np.full(shape, 0.1, np.float32)
        else:  # biases, beta, moving_mean
            out[name] = np.zeros(shape, np.float32)
    return out


def initialize(engine, seed: int = 42):
    """tf.initializers.global_variables() of model.py:673."""
    vals = initial_values(engine.variables(), seed)
    engine.set_params(vals)
    return vals
