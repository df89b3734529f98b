"""Drop-in for the reference's `networks.VNet` constructor / GetNetwork interface (networks.py:209-305).

In the reference `VNet(...).GetNetwork(x)` appends the V-Net ops to the TF1 graph and returns the logits
tensor.  Here it returns the logits *array* computed by the CUDA engine for the NumPy batch `x`
([N,X,Y,Z,M] float32), with the same constructor arguments.  Only the configuration the reference's live
path uses is accelerated: 3-D inputs and `activation_fn="prelu"` (model.py:428-438).
"""
from __future__ import annotations

import numpy as np

from .engine import VNetEngine
from .init import initialize


class VNet(object):
    def __init__(self, num_classes, dropout_rate=0.01, num_channels=16, num_levels=4, num_convolutions=(1, 2, 3, 3),
                 bottom_convolutions=3, is_training=True, activation_fn="prelu", precision="bf16x3", device=0):
        assert num_levels == len(num_convolutions)  # networks.py:228
        if activation_fn != "prelu":
            raise NotImplementedError("only activation_fn='prelu' (the setting model.py uses for VNet) is accelerated")
        self.num_classes = num_classes
        self.dropout_rate = dropout_rate
        self.num_channels = num_channels
        self.num_levels = num_levels
        self.num_convolutions = tuple(num_convolutions)
        self.bottom_convolutions = bottom_convolutions
        self.is_training = is_training
        self.train_phase = True  # the reference always feeds train_phase=True (model.py:747,788,917)
        self.precision, self.device = precision, device
        self.engine = None

    def _ensure(self, x, **kw):
        if self.engine is None:
            self.engine = VNetEngine(num_classes=self.num_classes, in_channels=int(x.shape[-1]), patch_shape=x.shape[1:4],
                                     max_batch=int(x.shape[0]), num_channels=self.num_channels, num_levels=self.num_levels,
                                     num_convolutions=self.num_convolutions, bottom_convolutions=self.bottom_convolutions,
                                     precision=self.precision, device=self.device, **kw)
            initialize(self.engine)
        return self.engine

    def GetNetwork(self, x):
        x = np.ascontiguousarray(x, np.float32)
        if x.ndim != 5:
            raise ValueError("only the 3-D branch of networks.VNet is accelerated: x must be [N,X,Y,Z,M]")
        logits, _, _ = self._ensure(x).forward(x, want_softmax=False, want_argmax=False)
        return logits


class LegacyVNet(VNet):
    """Drop-in for the legacy `VNet.VNet(num_classes, keep_prob, ...).network_fn(x)` (VNet.py:75-155) that only
    `train.py:271-279` uses: two batch norms per convolution with the residual added between them, and a
    true residual to the up-convolution output in the decoder."""

    def __init__(self, num_classes, keep_prob=1.0, num_channels=16, num_levels=4, num_convolutions=(1, 2, 3, 3),
                 bottom_convolutions=3, is_training=True, activation_fn="prelu", precision="bf16x3", device=0):
        super().__init__(num_classes, 1.0 - keep_prob, num_channels, num_levels, num_convolutions, bottom_convolutions,
                         is_training, activation_fn, precision, device)
        self.keep_prob = keep_prob

    def network_fn(self, x):
        x = np.ascontiguousarray(x, np.float32)
        logits, _, _ = self._ensure(x, flavour="legacy").forward(x, want_softmax=False, want_argmax=False)
        return logits
