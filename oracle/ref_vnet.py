"""CPU oracle: TF1-semantics restatement of the reference V-Net hot path (TEST INFRASTRUCTURE ONLY).

This module is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product path
(``vnet_tensorflow_b200``) never routes through it and fails loudly without its CUDA library.

PINNING.  The reference ships no tests and no golden vectors, and TensorFlow cannot run in this environment
(SURVEY.md §4, §8c), so the numerics of TensorFlow's own kernels stay UNPINNED: every op below follows the
documented TensorFlow-1.15 semantics (SAME padding, training-mode batch norm with biased variance,
conv3d_transpose = input-gradient of conv3d, Adam epsilon-hat form, non-staircase exponential decay).
What IS pinned is the graph the reference builds: ``tests/golden/make_reference_fixtures.py`` imports the
reference's own ``networks.py`` / ``layers2.py`` / ``VNet.py`` / ``Layers.py`` / ``attention.py`` /
``OutputModule.py`` unmodified (and compiles ``dice_coe`` out of ``model.py`` / ``train.py``), executes them over an
eager stand-in for the ~30 TF symbols they touch (``tests/tf1_shim.py``) and stores logits, losses, every gradient,
the UPDATE_OPS moving statistics and the variable names in creation order; ``tests/test_reference_pin.py`` holds this
restatement to those vectors (1e-9 on tensors in fp64).  Further self-checks (fp64 finite differences, structural
invariances, analytic Dice cases) live in ``tests/test_oracle.py``.

Reference files restated (paths relative to /root/reference):
  layers2.py:4-30    xavier / constant initialisers        -> xavier_uniform, init_params
  layers2.py:59-63   convolution  (tf.nn.convolution + b)  -> conv_same
  layers2.py:65-74   deconvolution (conv3d_transpose + b)  -> deconv_k2s2
  layers2.py:78-94   down_convolution / up_convolution     -> used by VNet.forward
  layers2.py:97-99   prelu                                 -> prelu
  networks.py:209-365 VNet (GetNetwork, convolution_block, convolution_block_2) -> VNet.forward
  VNet.py:26-155, Layers.py:88-131  legacy graph flavour used by train.py -> forward(spec.flavour='legacy')
  model.py:26-85     dice_coe                              -> dice_coe
  model.py:87-92     weighted softmax xent                 -> weighted_xent
  model.py:447,477,495-560,568  softmax / one_hot / loss zoo / argmax -> loss_from_logits, predict
  model.py:641-666   exponential_decay, Adam/SGD, UPDATE_OPS -> learning_rate, adam_update, train_step

Layout everywhere: NDHWC ([N, X, Y, Z, C], channels fastest) as in model.py:306-312.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_MOMENTUM = 0.99  # networks.py:259 (momentum=0.99)
BN_EPS = 1e-3  # networks.py:259 (epsilon=0.001)
PRELU_INIT = 0.1  # layers2.py:98


@dataclass
class VNetSpec:
    """Constructor arguments of networks.VNet (networks.py:209-244)."""

    num_classes: int = 2
    in_channels: int = 1
    num_channels: int = 16
    num_levels: int = 4
    num_convolutions: Tuple[int, ...] = (1, 2, 3, 3)
    bottom_convolutions: int = 3
    flavour: str = "networks"  # "networks" = networks.VNet (live path); "legacy" = VNet.py / Layers.py (train.py)

    def __post_init__(self):
        self.num_convolutions = tuple(self.num_convolutions)
        assert self.num_levels == len(self.num_convolutions)  # networks.py:228


# --------------------------------------------------------------------------------------------------
# variable inventory, in TF variable-creation order
# --------------------------------------------------------------------------------------------------
def _bn_names(scope: str, idx: int) -> List[Tuple[str, str]]:
    base = scope + "/batch_normalization" + ("" if idx == 0 else "_%d" % idx)
    return [(base + "/gamma", "gamma"), (base + "/beta", "beta"),
            (base + "/moving_mean", "moving_mean"), (base + "/moving_variance", "moving_variance")]


def param_specs(spec: VNetSpec) -> List[Tuple[str, Tuple[int, ...], str]]:
    """[(tf_variable_name, shape, kind)] in creation order; kind in
    {weights, biases, gamma, beta, moving_mean, moving_variance, alpha}."""
    out: List[Tuple[str, Tuple[int, ...], str]] = []
    C0 = spec.num_channels

    def conv(scope, shape, bias_c):
        out.append((scope + "/weights", tuple(shape), "weights"))
        out.append((scope + "/biases", (bias_c,), "biases"))

    def bn(scope, idx, c):
        for name, kind in _bn_names(scope, idx):
            out.append((name, (c,), kind))

    def alpha(scope, c):
        out.append((scope + "/alpha", (c,), "alpha"))

    s = "vnet/input_layer"
    if spec.in_channels == 1:  # networks.py:254-259
        bn(s, 0, C0)
    else:  # networks.py:260-266
        conv(s, (5, 5, 5, spec.in_channels, C0), C0)
        bn(s, 0, C0)
        alpha(s, C0)
    legacy = spec.flavour == "legacy"
    for l in range(spec.num_levels):  # networks.py:270-280
        c = C0 * 2 ** l
        for i in range(spec.num_convolutions[l]):
            s = "vnet/encoder/level_%d/conv_%d" % (l + 1, i + 1)
            conv(s, (5, 5, 5, c, c), c)
            bn(s, 0, c)
            if legacy:  # VNet.py:31-35: two batch norms per convolution
                bn(s, 1, c)
            alpha(s, c)
        s = "vnet/encoder/level_%d/down_convolution" % (l + 1)
        conv(s, (2, 2, 2, c, 2 * c), 2 * c)
        bn(s, 0, 2 * c)
        alpha(s, 2 * c)
    c = C0 * 2 ** spec.num_levels
    for i in range(spec.bottom_convolutions):  # networks.py:282-283
        s = "vnet/bottom_level/conv_%d" % (i + 1)
        conv(s, (5, 5, 5, c, c), c)
        bn(s, 0, c)
        if legacy:
            bn(s, 1, c)
        alpha(s, c)
    for l in reversed(range(spec.num_levels)):  # networks.py:285-296
        c = C0 * 2 ** l
        s = "vnet/decoder/level_%d/up_convolution" % (l + 1)
        # layers2.py:92: filter = kernel + [num_channels // factor, num_channels] = [2,2,2,c,2c]; bias c
        conv(s, (2, 2, 2, c, 2 * c), c)
        bn(s, 0, c)
        alpha(s, c)
        n = spec.num_convolutions[l]
        s = "vnet/decoder/level_%d/conv_1" % (l + 1)
        conv(s, (5, 5, 5, 2 * c, c), c)
        if legacy:  # VNet.py:45-72
            bn(s, 0, c)
            if n == 1:
                bn(s, 1, c)
            alpha(s, c)
            for i in range(1, n):
                s = "vnet/decoder/level_%d/conv_%d" % (l + 1, i + 1)
                conv(s, (5, 5, 5, c, c), c)
                bn(s, 0, c)
                bn(s, 1, c)
                alpha(s, c)
        elif n == 1:  # networks.py:328-340: three BNs
            bn(s, 0, c)
            bn(s, 1, c)
            bn(s, 2, c)
            alpha(s, c)
        else:  # networks.py:342-349
            bn(s, 0, c)
            alpha(s, c)
            for i in range(1, n):  # networks.py:351-363: two BNs each
                s = "vnet/decoder/level_%d/conv_%d" % (l + 1, i + 1)
                conv(s, (5, 5, 5, c, c), c)
                bn(s, 0, c)
                bn(s, 1, c)
                alpha(s, c)
    s = "vnet/output_layer"  # networks.py:298-303
    conv(s, (1, 1, 1, C0, spec.num_classes), spec.num_classes)
    bn(s, 0, spec.num_classes)
    return out


def xavier_uniform(shape: Sequence[int], rng: np.random.Generator) -> np.ndarray:
    """layers2.py:16-21: lim = sqrt(6 / (prod(spatial) * (Cin + Cout))), U(-lim, lim), float32."""
    s = len(shape) - 2
    num_activations = np.prod(shape[:s]) * np.sum(shape[s:])
    lim = np.sqrt(6.0 / num_activations)
    return rng.uniform(-lim, lim, size=tuple(shape)).astype(np.float32)


def init_params(spec: VNetSpec, seed: int = 42) -> "OrderedDict[str, np.ndarray]":
    """Initial values: weights Xavier-uniform (layers2.py:60), biases 0 (layers2.py:61), gamma 1,
    beta 0, moving_mean 0, moving_variance 1 (tf.layers defaults), alpha 0.1 (layers2.py:98).
    The reference draws from NumPy's unseeded global RNG; here PCG64(seed) in creation order."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, shape, kind in param_specs(spec):
        if kind == "weights":
            p[name] = xavier_uniform(shape, rng)
        elif kind in ("biases", "beta", "moving_mean"):
            p[name] = np.zeros(shape, np.float32)
        elif kind in ("gamma", "moving_variance"):
            p[name] = np.ones(shape, np.float32)
        elif kind == "alpha":
            p[name] = np.full(shape, PRELU_INIT, np.float32)
        else:
            raise ValueError(kind)
    return p


def trainable_names(spec: VNetSpec) -> List[str]:
    return [n for n, _, k in param_specs(spec) if k in ("weights", "biases", "gamma", "beta", "alpha")]


# --------------------------------------------------------------------------------------------------
# ops (NDHWC in / out)
# --------------------------------------------------------------------------------------------------
def _same_pads(size: int, k: int, stride: int) -> Tuple[int, int]:
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return total // 2, total - total // 2


def conv_same(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """layers2.py:59-63: tf.nn.convolution(x, w, 'SAME', strides) + b; w is [kd,kh,kw,Cin,Cout]."""
    k = w.shape[0]
    xin = x.permute(0, 4, 1, 2, 3)
    pads = []
    for dim in (3, 2, 1):  # F.pad order: last spatial dim first
        lo, hi = _same_pads(x.shape[dim], k, stride)
        pads += [lo, hi]
    if any(pads):
        xin = F.pad(xin, pads)
    y = F.conv3d(xin, w.permute(4, 3, 0, 1, 2), stride=stride)
    return y.permute(0, 2, 3, 4, 1) + b


def deconv_k2s2(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, out_spatial) -> torch.Tensor:
    """layers2.py:65-74 + 88-94: tf.nn.conv3d_transpose(x, w, output_shape, [1,2,2,2,1], 'SAME') + b
    with w = [2,2,2,Cout,Cin]: out[n,2i+a,2j+b,2k+d,co] = sum_ci x[n,i,j,k,ci] * w[a,b,d,co,ci]."""
    y = F.conv_transpose3d(x.permute(0, 4, 1, 2, 3), w.permute(4, 3, 0, 1, 2), stride=2)
    y = y.permute(0, 2, 3, 4, 1)
    assert tuple(y.shape[1:4]) == tuple(out_spatial), "odd skip sizes (SURVEY H7) not restated"
    return y + b


def prelu(x: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    """layers2.py:97-99: max(0,x) + alpha * min(0,x)."""
    zero = torch.zeros((), dtype=x.dtype)
    return torch.maximum(zero, x) + alpha * torch.minimum(zero, x)


class _Ctx:
    """Carries parameters and collects BN moving-statistic updates / intermediate tensors."""

    def __init__(self, params: Dict[str, torch.Tensor], collect: Optional[dict]):
        self.p = params
        self.updates: Dict[str, torch.Tensor] = {}
        self.collect = collect

    def bn(self, x: torch.Tensor, scope: str, idx: int = 0) -> torch.Tensor:
        """tf.layers.batch_normalization(momentum=.99, epsilon=1e-3, training=True) on a 5-D tensor
        (non-fused path): biased batch variance over (N,X,Y,Z); moving stats updated with the same."""
        base = scope + "/batch_normalization" + ("" if idx == 0 else "_%d" % idx)
        if self.collect is not None and idx == 0:
            self.collect[scope + ":bn_in"] = x  # what the first batch norm of the scope normalises (conv [+ block input])
        mean = x.mean(dim=(0, 1, 2, 3))
        var = ((x - mean) ** 2).mean(dim=(0, 1, 2, 3))
        y = (x - mean) * torch.rsqrt(var + BN_EPS) * self.p[base + "/gamma"] + self.p[base + "/beta"]
        with torch.no_grad():
            mm, mv = self.p[base + "/moving_mean"], self.p[base + "/moving_variance"]
            self.updates[base + "/moving_mean"] = mm * BN_MOMENTUM + mean.detach() * (1 - BN_MOMENTUM)
            self.updates[base + "/moving_variance"] = mv * BN_MOMENTUM + var.detach() * (1 - BN_MOMENTUM)
        return y

    def conv(self, x, scope, stride=1):
        return conv_same(x, self.p[scope + "/weights"], self.p[scope + "/biases"], stride)

    def act(self, x, scope):
        return prelu(x, self.p[scope + "/alpha"])

    def tap(self, name, t):
        if self.collect is not None:
            self.collect[name] = t


def _dropout(x: torch.Tensor, rate: float, masks, key: str) -> torch.Tensor:
    """tf.nn.dropout(x, rate): keep where u >= rate, scale 1/(1-rate); rate 0 is the identity.
    `masks` (optional dict name -> {0,1} tensor) injects an explicit keep-mask for parity tests."""
    if rate == 0.0:
        return x
    if masks is None or key not in masks:
        raise ValueError("oracle dropout needs an explicit keep-mask for %s (TF RNG not reproducible)" % key)
    return x * masks[key].to(x.dtype) / (1.0 - rate)


def forward(params: Dict[str, torch.Tensor], images: torch.Tensor, spec: VNetSpec,
            dropout_rate: float = 0.0, masks=None, collect: Optional[dict] = None):
    """networks.VNet.GetNetwork (networks.py:246-305) with activation_fn = prelu (model.py:437).
    Returns (logits [N,X,Y,Z,K], bn_moving_updates)."""
    cx = _Ctx(params, collect)
    x = images
    C0 = spec.num_channels
    s = "vnet/input_layer"
    if spec.in_channels == 1:
        x = x.repeat(1, 1, 1, 1, C0)  # tf.tile, networks.py:258
        x = cx.bn(x, s)
    else:
        x = cx.conv(x, s)
        x = cx.bn(x, s)
        x = cx.act(x, s)
    cx.tap(s, x)

    def legacy_block(x, n, scope):  # VNet.py:26-40
        layer_input = x
        for i in range(n):
            sc = "%s/conv_%d" % (scope, i + 1)
            x = cx.conv(x, sc)
            x = cx.bn(x, sc, 0)
            if i == n - 1:
                x = x + layer_input
            x = cx.bn(x, sc, 1)
            x = cx.act(x, sc)
            x = _dropout(x, dropout_rate, masks, sc)
            cx.tap(sc, x)
        return x

    def legacy_block_2(x, f, n, scope):  # VNet.py:43-73 (true residual to the up-convolution output)
        layer_input = x
        x = torch.cat((x, f), dim=-1)
        sc = scope + "/conv_1"
        x = cx.conv(x, sc)
        x = cx.bn(x, sc, 0)
        if n == 1:
            x = x + layer_input
            x = cx.bn(x, sc, 1)
        x = cx.act(x, sc)
        x = _dropout(x, dropout_rate, masks, sc)
        cx.tap(sc, x)
        for i in range(1, n):
            sc = "%s/conv_%d" % (scope, i + 1)
            x = cx.conv(x, sc)
            x = cx.bn(x, sc, 0)
            if i == n - 1:
                x = x + layer_input
            x = cx.bn(x, sc, 1)
            x = cx.act(x, sc)
            x = _dropout(x, dropout_rate, masks, sc)
            cx.tap(sc, x)
        return x

    def convolution_block(x, n, scope):  # networks.py:307-322
        if spec.flavour == "legacy":
            return legacy_block(x, n, scope)
        layer_input = x
        for i in range(n):
            sc = "%s/conv_%d" % (scope, i + 1)
            x = cx.conv(x, sc)
            if i == n - 1:
                x = x + layer_input
            x = cx.bn(x, sc)
            x = cx.act(x, sc)
            x = _dropout(x, dropout_rate, masks, sc)
            cx.tap(sc, x)
        return x

    def convolution_block_2(x, f, n, scope):  # networks.py:324-365
        if spec.flavour == "legacy":
            return legacy_block_2(x, f, n, scope)
        x = torch.cat((x, f), dim=-1)
        sc = scope + "/conv_1"
        if n == 1:
            x = cx.conv(x, sc)
            x = cx.bn(x, sc, 0)
            layer_input = cx.bn(x, sc, 1)
            x = x + layer_input
            x = cx.bn(x, sc, 2)
            x = cx.act(x, sc)
            x = _dropout(x, dropout_rate, masks, sc)
            cx.tap(sc, x)
            return x
        x = cx.conv(x, sc)
        x = cx.bn(x, sc, 0)
        x = cx.act(x, sc)
        x = _dropout(x, dropout_rate, masks, sc)
        cx.tap(sc, x)
        for i in range(1, n):
            sc = "%s/conv_%d" % (scope, i + 1)
            x = cx.conv(x, sc)
            layer_input = cx.bn(x, sc, 0)  # computed even when unused (moving stats still update)
            if i == n - 1:
                x = x + layer_input
            x = cx.bn(x, sc, 1)
            x = cx.act(x, sc)
            x = _dropout(x, dropout_rate, masks, sc)
            cx.tap(sc, x)
        return x

    features = []
    for l in range(spec.num_levels):
        scope = "vnet/encoder/level_%d" % (l + 1)
        x = convolution_block(x, spec.num_convolutions[l], scope)
        features.append(x)
        sc = scope + "/down_convolution"
        x = cx.conv(x, sc, stride=2)  # layers2.py:78-84
        x = cx.bn(x, sc)
        x = cx.act(x, sc)
        cx.tap(sc, x)
    x = convolution_block(x, spec.bottom_convolutions, "vnet/bottom_level")
    for l in reversed(range(spec.num_levels)):
        scope = "vnet/decoder/level_%d" % (l + 1)
        f = features[l]
        sc = scope + "/up_convolution"
        x = deconv_k2s2(x, params[sc + "/weights"], params[sc + "/biases"], f.shape[1:4])
        x = cx.bn(x, sc)
        x = cx.act(x, sc)
        cx.tap(sc, x)
        x = convolution_block_2(x, f, spec.num_convolutions[l], scope)
    s = "vnet/output_layer"
    logits = cx.conv(x, s)
    logits = cx.bn(logits, s)
    cx.tap(s, logits)
    return logits, cx.updates


# --------------------------------------------------------------------------------------------------
# loss / metrics (model.py)
# --------------------------------------------------------------------------------------------------
def dice_coe(output, target, loss_type="jaccard", axis=(1, 2, 3), weights=(), smooth=1e-5):
    """model.py:26-85, verbatim semantics (note: weighted form adds `smooth` once per class)."""
    inse = torch.sum(output * target, dim=axis)
    if loss_type == "jaccard":
        l = torch.sum(output * output, dim=axis)
        r = torch.sum(target * target, dim=axis)
    elif loss_type == "sorensen":
        l = torch.sum(output, dim=axis)
        r = torch.sum(target, dim=axis)
    else:
        raise Exception("Unknown loss_type")
    if len(weights) != 0:
        assert len(weights) == target.shape[-1]
        w = torch.tensor(list(weights), dtype=torch.float32).to(output.dtype)
        dice = torch.sum(2.0 * w * inse + smooth, dim=-1) / torch.sum(w * (l + r) + smooth, dim=-1)
        return torch.mean(dice)
    dice = (2.0 * inse + smooth) / (l + r + smooth)
    return torch.mean(dice)


def softmax_xent(labels_onehot, logits):
    """tf.nn.softmax_cross_entropy_with_logits: -sum_c t_c * log_softmax(logits)_c per voxel."""
    return -(labels_onehot * F.log_softmax(logits, dim=-1)).sum(dim=-1)


def weighted_xent(labels_onehot, logits, weights):
    """model.py:87-92."""
    cw = torch.tensor([list(weights)], dtype=torch.float32).to(logits.dtype)
    wv = torch.sum(cw * labels_onehot, dim=-1)
    return torch.mean(softmax_xent(labels_onehot, logits) * wv)


LOSS_NAMES = ("xent", "weighted_xent", "sorensen", "weighted_sorensen", "jaccard", "weighted_jaccard",
              "mixed_sorensen", "mixed_weighted_sorensen", "mixed_jaccard", "mixed_weighted_jaccard")


def loss_from_logits(logits, labels, loss_name="weighted_sorensen", weights=(), alpha=1.0):
    """model.py:447 (softmax), :477 (one_hot), :495-560 (loss zoo). labels: int [N,X,Y,Z]."""
    K = logits.shape[-1]
    softmax = torch.softmax(logits, dim=-1)
    lab = labels.long()
    valid = ((lab >= 0) & (lab < K)).unsqueeze(-1)
    onehot = F.one_hot(lab.clamp(0, K - 1), K).to(logits.dtype) * valid.to(logits.dtype)
    if loss_name == "xent":
        return torch.mean(softmax_xent(onehot, logits))
    if loss_name == "weighted_xent":
        return weighted_xent(onehot, logits, weights)
    kind = "sorensen" if "sorensen" in loss_name else "jaccard"
    weighted = "weighted" in loss_name
    d = dice_coe(softmax, onehot, loss_type=kind, weights=tuple(weights) if weighted else ())
    loss = 1.0 - d
    if loss_name.startswith("mixed"):
        x = weighted_xent(onehot, logits, weights) if weighted else torch.mean(softmax_xent(onehot, logits))
        loss = loss + alpha * x
    elif loss_name not in LOSS_NAMES:
        raise SystemExit("Invalid loss function")
    return loss


def predict(logits):
    """model.py:568 tf.argmax(logits, -1): int64, lowest index wins ties."""
    K = logits.shape[-1]
    best = logits[..., 0]
    idx = torch.zeros(logits.shape[:-1], dtype=torch.int64)
    for c in range(1, K):
        better = logits[..., c] > best
        idx = torch.where(better, torch.full_like(idx, c), idx)
        best = torch.where(better, logits[..., c], best)
    return idx


def dice_terms(logits, labels, kind="sorensen"):
    """Per-(n,c) reductions I, L, R of dice_coe (model.py:60-66) as an [N,K,3] tensor."""
    K = logits.shape[-1]
    softmax = torch.softmax(logits, dim=-1)
    lab = labels.long()
    valid = ((lab >= 0) & (lab < K)).unsqueeze(-1)
    t = F.one_hot(lab.clamp(0, K - 1), K).to(logits.dtype) * valid.to(logits.dtype)
    inse = (softmax * t).sum(dim=(1, 2, 3))
    if kind == "jaccard":
        l, r = (softmax * softmax).sum(dim=(1, 2, 3)), (t * t).sum(dim=(1, 2, 3))
    else:
        l, r = softmax.sum(dim=(1, 2, 3)), t.sum(dim=(1, 2, 3))
    return torch.stack([inse, l, r], dim=-1)


# --------------------------------------------------------------------------------------------------
# optimiser (model.py:641-666)
# --------------------------------------------------------------------------------------------------
def learning_rate(lr0: float, global_step: int, decay_steps: float, decay_factor: float) -> float:
    """tf.train.exponential_decay(staircase=False): lr0 * factor ** (step / steps)."""
    return lr0 * decay_factor ** (global_step / decay_steps)


ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-8  # tf.train.AdamOptimizer defaults (model.py:652)


def adam_update(p, g, m, v, t: int, lr: float):
    """TF Adam (epsilon-hat form): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA; p -= lr_t*m/(sqrt(v)+eps)."""
    lr_t = lr * math.sqrt(1.0 - ADAM_B2 ** t) / (1.0 - ADAM_B1 ** t)
    m = m + (g - m) * (1.0 - ADAM_B1)
    v = v + (g * g - v) * (1.0 - ADAM_B2)
    p = p - lr_t * m / (torch.sqrt(v) + ADAM_EPS)
    return p, m, v


def to_torch(params: Dict[str, np.ndarray], dtype=torch.float32, requires_grad=False):
    out = OrderedDict()
    for k, a in params.items():
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
        if requires_grad and not k.endswith(("moving_mean", "moving_variance")):
            t.requires_grad_(True)
        out[k] = t
    return out


def loss_and_grads(params_np, images, labels, spec: VNetSpec, loss_name="weighted_sorensen",
                   weights=(), alpha=1.0, dropout_rate=0.0, masks=None, dtype=torch.float32,
                   collect: Optional[dict] = None):
    """One forward + backward of the reference training graph. Returns (loss, logits, grads, bn_updates)."""
    p = to_torch(params_np, dtype, requires_grad=True)
    x = torch.from_numpy(np.ascontiguousarray(images)).to(dtype)
    y = torch.from_numpy(np.ascontiguousarray(labels))
    logits, updates = forward(p, x, spec, dropout_rate, masks, collect)
    loss = loss_from_logits(logits, y, loss_name, weights, alpha)
    names = [k for k, t in p.items() if t.requires_grad]
    grads = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
    g = OrderedDict()
    for k, t in zip(names, grads):
        g[k] = torch.zeros_like(p[k]) if t is None else t
    return loss.detach(), logits.detach(), g, updates


@dataclass
class TrainState:
    params: "OrderedDict[str, np.ndarray]"
    m: Dict[str, np.ndarray] = field(default_factory=dict)
    v: Dict[str, np.ndarray] = field(default_factory=dict)
    global_step: int = 0


def train_step(state: TrainState, images, labels, spec: VNetSpec, loss_name="weighted_sorensen",
               weights=(), alpha=1.0, lr0=1e-2, decay_steps=100, decay_factor=0.99,
               optimizer="Adam", dtype=torch.float32, momentum=0.9):
    """sess.run(train_op) of model.py:743-748 with dropout 0: fwd + bwd + optimiser + BN UPDATE_OPS."""
    loss, logits, grads, updates = loss_and_grads(state.params, images, labels, spec, loss_name,
                                                  weights, alpha, 0.0, None, dtype)
    lr = learning_rate(lr0, state.global_step, decay_steps, decay_factor)
    t = state.global_step + 1
    for k, g in grads.items():
        p = torch.from_numpy(state.params[k]).to(dtype)
        if optimizer == "Adam":
            m = torch.from_numpy(state.m.get(k, np.zeros_like(state.params[k]))).to(dtype)
            v = torch.from_numpy(state.v.get(k, np.zeros_like(state.params[k]))).to(dtype)
            p, m, v = adam_update(p, g, m, v, t, lr)
            state.m[k] = m.to(torch.float32).numpy()
            state.v[k] = v.to(torch.float32).numpy()
        elif optimizer == "SGD":
            p = p - lr * g
        elif optimizer in ("Momentum", "NesterovMomentum"):  # tf.train.MomentumOptimizer (model.py:653-656)
            a = torch.from_numpy(state.m.get(k, np.zeros_like(state.params[k]))).to(dtype)
            a = momentum * a + g
            p = p - (lr * (g + momentum * a) if optimizer == "NesterovMomentum" else lr * a)
            state.m[k] = a.to(torch.float32).numpy()
        else:
            raise SystemExit("Invalid optimizer")
        state.params[k] = p.to(torch.float32).numpy()
    for k, u in updates.items():
        state.params[k] = u.to(torch.float32).numpy()
    state.global_step = t
    return float(loss), logits.to(torch.float32).numpy(), {k: g.to(torch.float32).numpy() for k, g in grads.items()}


# --------------------------------------------------------------------------------------------------
# sliding-window inference geometry (model.py:866-937) -- pure index arithmetic
# --------------------------------------------------------------------------------------------------
def window_starts(dim: int, patch: int, stride: int) -> List[int]:
    """model.py:866-868,879-882: ceil((dim-patch)/stride)+1 windows, last one clamped to dim-patch."""
    num = int(math.ceil((dim - patch) / float(stride))) + 1
    starts = []
    for i in range(num):
        s = i * stride
        if s + patch > dim:
            s = dim - patch
        starts.append(s)
    return starts


# --------------------------------------------------------------------------------------------------
# Legacy attention path: train.py:269-312, attention.py, OutputModule.py
# --------------------------------------------------------------------------------------------------
MODULE_CHANNELS = 64  # attention.py:41 / OutputModule.py:41 num_channels default


def module_param_specs(var_scope: str, bn_scope: str, in_ch: int, num_classes: int, nch: int = MODULE_CHANNELS):
    """Variables of AttentionModule / OutputModule in creation order.  `tf.Variable`s are unnamed, so they are
    `<name_scope>/<variable_scope>/Variable[_k]` (attention.py:25-31); the tf.layers batch norms live under the
    variable scope only.  kind in {mod_w, mod_b, gamma, beta, moving_mean, moving_variance}."""
    out, vi, bi = [], 0, 0

    def var(scope, shape, kind):
        nonlocal vi
        out.append(("%s/Variable%s" % (scope, "" if vi == 0 else "_%d" % vi), tuple(shape), kind))
        vi += 1

    def bn(scope, c):
        nonlocal bi
        base = "%s/batch_normalization%s" % (scope, "" if bi == 0 else "_%d" % bi)
        for k in ("gamma", "beta", "moving_mean", "moving_variance"):
            out.append((base + "/" + k, (c,), k))
        bi += 1

    enc_v, enc_b = var_scope + "/encoder", bn_scope + "/encoder"
    cin = in_ch
    for _ in range(3):  # attention.py:105-109: three residual blocks
        var(enc_v, (3, 3, 3, cin, nch), "mod_w"); var(enc_v, (nch,), "mod_b"); bn(enc_b, nch)   # conv1 + BN
        var(enc_v, (3, 3, 3, nch, nch), "mod_w"); var(enc_v, (nch,), "mod_b"); bn(enc_b, nch)   # conv2 + BN
        var(enc_v, (1, 1, 1, cin, nch), "mod_w"); var(enc_v, (nch,), "mod_b")                   # shortcut 1x1x1
        bn(enc_b, nch)                                                                            # BN after the add
        cin = nch
    vi, bi = 0, 0
    var(var_scope + "/output", (1, 1, 1, nch, num_classes), "mod_w"); var(var_scope + "/output", (num_classes,), "mod_b")
    bn(bn_scope + "/output", num_classes)
    return out


def attention_param_specs(spec: VNetSpec, nch: int = MODULE_CHANNELS):
    """V-Net (legacy flavour) + AttentionModule + OutputModule variables (train.py:271-310)."""
    K = spec.num_classes
    return (param_specs(spec)
            + module_param_specs("attention/AttentionModule", "AttentionModule", K, K, nch)
            + module_param_specs("output/output", "output", K, K, nch))


def init_attention_params(spec: VNetSpec, seed: int = 42, module_seed: int = 43, nch: int = MODULE_CHANNELS):
    p = init_params(spec, seed)
    rng = np.random.Generator(np.random.PCG64(module_seed))
    for name, shape, kind in attention_param_specs(spec, nch)[len(param_specs(spec)):]:
        if kind == "mod_w":  # tf.truncated_normal(stddev=0.1): redraw beyond two sigma (attention.py:25-27)
            w = rng.normal(0.0, 0.1, size=shape)
            bad = np.abs(w) > 0.2
            while bad.any():
                w[bad] = rng.normal(0.0, 0.1, size=int(bad.sum()))
                bad = np.abs(w) > 0.2
            p[name] = w.astype(np.float32)
        elif kind in ("mod_b", "beta", "moving_mean"):
            p[name] = np.zeros(shape, np.float32)
        else:  # gamma, moving_variance
            p[name] = np.ones(shape, np.float32)
    return p


def _bn_inference(x, p, base):
    """tf.layers.batch_normalization(training=False): the modules are fed train_phase=False even while
    training (train.py:538-540), so they normalise with their never-updated moving statistics."""
    return (x - p[base + "/moving_mean"]) * torch.rsqrt(p[base + "/moving_variance"] + BN_EPS) * p[base + "/gamma"] + p[base + "/beta"]


def _conv3_valid_padded(x, w, b):
    """tf.pad 1 voxel then tf.nn.conv3d(..., 'VALID') + b (attention.py:84-90,63-70) == SAME 3^3 convolution."""
    y = F.conv3d(x.permute(0, 4, 1, 2, 3), w.permute(4, 3, 0, 1, 2), padding=1)
    return y.permute(0, 2, 3, 4, 1) + b


def module_forward(p, x, var_scope: str, bn_scope: str, collect=None):
    """AttentionModule.GetNetwork / OutputModule.GetNetwork (attention.py:83-114, OutputModule.py:83-114)."""
    vi, bi = [0], [0]

    def var(scope):
        n = "%s/Variable%s" % (scope, "" if vi[0] == 0 else "_%d" % vi[0])
        vi[0] += 1
        return p[n]

    def bn(scope, t):
        base = "%s/batch_normalization%s" % (scope, "" if bi[0] == 0 else "_%d" % bi[0])
        bi[0] += 1
        return _bn_inference(t, p, base)

    ev, eb = var_scope + "/encoder", bn_scope + "/encoder"
    for blk in range(3):
        c1 = torch.relu(bn(eb, _conv3_valid_padded(x, var(ev), var(ev))))          # ConvActivate3d_block (keep_prob 1)
        c2 = bn(eb, _conv3_valid_padded(c1, var(ev), var(ev)))                     # Conv3d_block
        up = F.conv3d(x.permute(0, 4, 1, 2, 3), var(ev).permute(4, 3, 0, 1, 2)).permute(0, 2, 3, 4, 1) + var(ev)
        x = torch.relu(bn(eb, c2 + up))
        if collect is not None:
            collect["%s/block_%d" % (bn_scope, blk + 1)] = x
    vi[0], bi[0] = 0, 0
    w, b = var(var_scope + "/output"), var(var_scope + "/output")
    y = F.conv3d(x.permute(0, 4, 1, 2, 3), w.permute(4, 3, 0, 1, 2)).permute(0, 2, 3, 4, 1) + b
    return bn(bn_scope + "/output", y)


def attention_forward(params, images, spec: VNetSpec, collect=None):
    """train.py:269-312: V-Net (VNet.py flavour, batch statistics) -> attention module -> (1 + softmax) gating ->
    output module.  Returns dict of the named tensors and the V-Net's moving-statistic updates."""
    logits_vnet, updates = forward(params, images, spec, collect=collect)
    logits_att = module_forward(params, logits_vnet, "attention/AttentionModule", "AttentionModule", collect)
    softmax_att = torch.softmax(logits_att, dim=-1)
    logits_masked = (1.0 + softmax_att) * logits_vnet                                # train.py:302
    logits_out = module_forward(params, logits_masked, "output/output", "output", collect)
    return {"logits_vnet": logits_vnet, "logits_attention": logits_att, "softmax_attention": softmax_att,
            "logits_masked": logits_masked, "logits_output": logits_out}, updates


def attention_total_loss(out, labels, distmap, loss="jaccard", att_loss="l2", weights=(), alpha=1.0):
    """train.py:351-418: segmentation loss on logits_output + attention loss against the distance map.
    `loss`: train.py's own names 'sorensen_fg' (= --loss_function sorensen, train.py:373-377: foreground channel
    against the label volume, mean over the batch), 'jaccard' and 'xent' (train.py:353-357,378-382; these two
    coincide with model.py's forms), or any other Loss.Name of model.py:495-560 (used for K > 2, where
    train.py's binary-only forms do not apply)."""
    lo = out["logits_output"]
    lab = labels.long()
    if loss == "sorensen_fg":
        o = torch.softmax(lo, dim=-1)[..., 1:2]
        t = (lab == 1).to(lo.dtype).unsqueeze(-1)  # tf.cast(labels) for binary labels
        inse = (o * t).sum(dim=(1, 2, 3))
        dice = (2.0 * inse + 1e-5) / (o.sum(dim=(1, 2, 3)) + t.sum(dim=(1, 2, 3)) + 1e-5)
        seg = 1.0 - dice.mean()
    else:
        seg = loss_from_logits(lo, lab, loss, weights, alpha)
    sa = out["softmax_attention"]
    d = distmap.to(lo.dtype)
    if att_loss == "l2":     # train.py:387-393
        att = torch.mean(torch.square(sa[..., 1] - d) * 100.0)
    elif att_loss == "abs":  # train.py:394-399
        att = torch.mean(torch.abs(sa - torch.stack([1.0 - d, d], dim=-1)))
    elif att_loss in (None, "none"):
        att = torch.zeros((), dtype=lo.dtype)
    else:
        raise SystemExit("Invalid loss function")
    return att + seg, seg, att


def attention_loss_and_grads(params_np, images, labels, distmap, spec, loss="jaccard", att_loss="l2", dtype=torch.float32,
                             weights=(), alpha=1.0):
    p = to_torch(params_np, dtype, requires_grad=True)
    for k in p:  # module moving statistics are constants
        if k.endswith(("moving_mean", "moving_variance")):
            p[k].requires_grad_(False)
    x = torch.from_numpy(np.ascontiguousarray(images)).to(dtype)
    out, updates = attention_forward(p, x, spec)
    total, seg, att = attention_total_loss(out, torch.from_numpy(np.ascontiguousarray(labels)),
                                           torch.from_numpy(np.ascontiguousarray(distmap)), loss, att_loss, weights, alpha)
    names = [k for k, t in p.items() if t.requires_grad]
    grads = torch.autograd.grad(total, [p[k] for k in names], allow_unused=True)
    g = OrderedDict((k, torch.zeros_like(p[k]) if t is None else t) for k, t in zip(names, grads))
    return total.detach(), seg.detach(), att.detach(), {k: v.detach() for k, v in out.items()}, g, updates
