"""Eager stand-in for the handful of TensorFlow-1 symbols the reference's network code touches, so that the
reference's OWN source (`/root/reference/networks.py`, `layers2.py`, and `dice_coe` cut out of `model.py`) can be
imported and executed in a container without TensorFlow.  TEST INFRASTRUCTURE ONLY.

What this pins and what it does not: running `networks.VNet(...).GetNetwork(x)` over this shim executes the
reference's graph-building code verbatim -- which ops, in which order, in which `variable_scope`, the `x + BN(x)`
decoder quirk, the dead batch norms, the variable names and their creation order -- and torch autograd through the
same code gives the gradients TF autodiff would form.  The arithmetic of each op is this file's restatement of the
documented TF-1.15 semantics (SAME padding, NDHWC cross-correlation, conv3d_transpose as the input-gradient of a
convolution, biased batch variance with epsilon inside the square root), NOT TensorFlow's kernels: numerics of
individual ops stay unpinned, the wiring does not.

Usage: `tf = install(params)`; import the reference modules; run; read `tf.STATE.created` / `.bn_updates`.
"""
from __future__ import annotations

import contextlib
import math
import sys
import types
from collections import OrderedDict

import numpy as np
import torch


class Shape(list):
    """`Tensor.get_shape()`: indexable / sliceable / len()-able list of ints."""


class T:
    """Eager tensor: a torch tensor with the few methods the reference calls."""

    def __init__(self, v):
        self.v = v

    def get_shape(self):
        return Shape(int(s) for s in self.v.shape)

    @property
    def dtype(self):
        return self.v.dtype

    def _u(self, o):
        return o.v if isinstance(o, T) else o

    def __add__(self, o): return T(self.v + self._u(o))
    def __radd__(self, o): return T(self._u(o) + self.v)
    def __sub__(self, o): return T(self.v - self._u(o))
    def __rsub__(self, o): return T(self._u(o) - self.v)
    def __mul__(self, o): return T(self.v * self._u(o))
    def __rmul__(self, o): return T(self._u(o) * self.v)
    def __truediv__(self, o): return T(self.v / self._u(o))
    def __rtruediv__(self, o): return T(self._u(o) / self.v)
    def __neg__(self): return T(-self.v)


class Placeholder:
    """tf.placeholder: only `train_phase` reaches the ops.  model.py feeds True in training, testing AND evaluation
    (model.py:747,788,917 -- SURVEY R2), hence the default; train.py feeds False to the attention / output modules
    (train.py:538-540): the fixture generator sets `.value = False` on those."""

    def __init__(self, dtype, name=None):
        self.dtype, self.name, self.value = dtype, name, True

    def __bool__(self):
        return bool(self.value)


class State:
    def __init__(self, params, dtype):
        self.params = params                   # name -> numpy array to inject
        self.dtype = dtype
        self.scope = []                        # variable_scope stack
        self.names = []                        # name_scope stack (variable_scope pushes here too, as in TF)
        self.created = OrderedDict()           # name -> torch leaf, in creation order
        self.trainable = OrderedDict()         # name -> bool
        self.uid = {}                          # (scope path, default name) -> next index
        self.bn_updates = OrderedDict()        # moving_mean / moving_variance name -> updated value (UPDATE_OPS)

    def path(self):
        return "/".join(self.scope)


def _unwrap(x):
    return x.v if isinstance(x, T) else x


def _same_pad(size, k, s):
    """TF 'SAME': out = ceil(in / stride); pad_total = max((out - 1) * stride + k - in, 0); before = total // 2."""
    out = int(math.ceil(size / float(s)))
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def install(params, dtype=torch.float64):
    st = State(params, dtype)
    tf = types.ModuleType("tensorflow")
    tf.STATE = st
    tf.bool, tf.float32, tf.float64, tf.int32 = "bool", torch.float32, torch.float64, torch.int32

    @contextlib.contextmanager
    def variable_scope(name, default_name=None, reuse=None):
        if name is None:   # tf.variable_scope(None, default_name=...): uniquified inside the enclosing scope
            key = (st.path(), default_name)
            n = st.uid.get(key, 0)
            st.uid[key] = n + 1
            name = default_name if n == 0 else "%s_%d" % (default_name, n)
        st.scope.append(name)
        st.names.append(name)
        try:
            yield
        finally:
            st.scope.pop()
            st.names.pop()

    @contextlib.contextmanager
    def name_scope(name):
        st.names.append(name)
        try:
            yield
        finally:
            st.names.pop()

    def Variable(initial_value, name=None, trainable=True):
        """tf.Variable without a name: `<name scope>/Variable[_k]`, uniquified inside the current name scope."""
        ns = "/".join(st.names)
        key = (ns, "tf.Variable")
        n = st.uid.get(key, 0)
        st.uid[key] = n + 1
        full = (ns + "/" if ns else "") + ("Variable" if n == 0 else "Variable_%d" % n)
        if full not in st.params:
            raise KeyError("the reference creates variable %r, which the injected parameter set does not have" % full)
        val = np.asarray(st.params[full])
        want = initial_value[1] if isinstance(initial_value, tuple) and initial_value[0] == "init" else None
        if want is not None and tuple(want) != tuple(val.shape):
            raise ValueError("variable %s: the reference asks for shape %r, injected %r" % (full, tuple(want), val.shape))
        leaf = torch.tensor(val, dtype=st.dtype, requires_grad=True)
        st.created[full] = leaf
        st.trainable[full] = True
        return T(leaf)

    def _shape_tuple(shape):
        return (int(shape),) if isinstance(shape, (int, np.integer)) else tuple(int(v) for v in shape)

    def pad(x, paddings, mode="CONSTANT"):
        assert mode == "CONSTANT"
        pw = np.asarray(_unwrap(paddings)).astype(int)
        flat = []
        for a in reversed(range(pw.shape[0])):
            flat += [int(pw[a][0]), int(pw[a][1])]
        return T(torch.nn.functional.pad(_unwrap(x), flat))

    tf.name_scope, tf.Variable, tf.pad = name_scope, Variable, pad
    tf.truncated_normal = lambda shape, mean=0.0, stddev=1.0: ("init", _shape_tuple(shape))
    tf.zeros = lambda shape: ("init", _shape_tuple(shape))
    def constant(value, dtype=None, shape=None):
        a = np.asarray(value)
        return T(torch.as_tensor(a) if a.dtype.kind in "iu" else torch.as_tensor(a.astype(np.float64)).to(st.dtype))

    tf.constant = constant
    tf.multiply = lambda a, b: T(_unwrap(a) * _unwrap(b))

    def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True):
        full = (st.path() + "/" if st.scope else "") + name
        if full in st.created:
            raise ValueError("Variable %s already exists, disallowed (reuse=False)" % full)
        if full not in st.params:
            raise KeyError("the reference creates variable %r, which the injected parameter set does not have" % full)
        if isinstance(initializer, np.ndarray):
            shape = initializer.shape
        elif isinstance(shape, int):
            shape = (shape,)
        val = np.asarray(st.params[full])
        if shape is not None and tuple(int(s) for s in shape) != tuple(val.shape):
            raise ValueError("variable %s: the reference asks for shape %r, injected %r" % (full, tuple(shape), val.shape))
        leaf = torch.tensor(val, dtype=st.dtype, requires_grad=bool(trainable))
        st.created[full] = leaf
        st.trainable[full] = bool(trainable)
        return T(leaf)

    def constant_initializer(value):
        return ("constant", value)

    def placeholder(dtype, shape=None, name=None):
        return Placeholder(dtype, name)

    def shape_fn(x):
        return [int(s) for s in _unwrap(x).shape]

    def tile(x, multiples):
        return T(_unwrap(x).repeat(*[int(m) for m in multiples]))

    def concat(values, axis):
        return T(torch.cat([_unwrap(v) for v in values], dim=axis))

    def maximum(a, b):
        a, b = _unwrap(a), _unwrap(b)
        t = a if torch.is_tensor(a) else b
        return T(torch.maximum(torch.as_tensor(a, dtype=t.dtype), torch.as_tensor(b, dtype=t.dtype)))

    def minimum(a, b):
        a, b = _unwrap(a), _unwrap(b)
        t = a if torch.is_tensor(a) else b
        return T(torch.minimum(torch.as_tensor(a, dtype=t.dtype), torch.as_tensor(b, dtype=t.dtype)))

    def reduce_sum(x, axis=None, name=None):
        v = _unwrap(x)
        return T(v.sum() if axis is None else v.sum(dim=axis))

    def reduce_mean(x, axis=None, name=None):
        v = _unwrap(x)
        return T(v.mean() if axis is None else v.mean(dim=axis))

    def cast(x, dtype):
        # values are rounded to the requested type (dice_coe casts Loss.Weights to float32), then carried in the
        # working precision of the run
        v = _unwrap(x)
        v = torch.as_tensor(v, dtype=torch.float64) if not torch.is_tensor(v) else v
        if dtype in (torch.float32, torch.int32):
            v = v.to(dtype)
        return T(v.to(st.dtype))

    def one_hot(idx, depth):
        v = _unwrap(idx).long()
        oh = torch.zeros(tuple(v.shape) + (depth,), dtype=st.dtype)
        valid = (v >= 0) & (v < depth)
        oh.scatter_(-1, v.clamp(0, depth - 1).unsqueeze(-1), valid.to(st.dtype).unsqueeze(-1))
        return T(oh)

    tf.variable_scope, tf.get_variable, tf.constant_initializer, tf.placeholder = variable_scope, get_variable, constant_initializer, placeholder
    tf.shape, tf.tile, tf.concat, tf.maximum, tf.minimum = shape_fn, tile, concat, maximum, minimum
    tf.reduce_sum, tf.reduce_mean, tf.cast, tf.one_hot = reduce_sum, reduce_mean, cast, one_hot
    tf.add = lambda a, b, name=None: T(_unwrap(a) + _unwrap(b))

    # ---- tf.nn ------------------------------------------------------------------------------------------------
    nn = types.ModuleType("tensorflow.nn")

    def convolution(x, w, padding="SAME", strides=None, dilation_rate=None):
        """tf.nn.convolution, N-D channels-last cross-correlation, filter [k..., Cin, Cout]."""
        xv, wv = _unwrap(x), _unwrap(w)
        rank = xv.dim() - 2
        assert rank == 3 and padding == "SAME" and dilation_rate is None
        strides = list(strides) if strides else [1] * rank
        xc = xv.permute(0, 4, 1, 2, 3)
        pads = []
        for a in reversed(range(rank)):   # F.pad wants the last axis first
            b, e = _same_pad(xv.shape[1 + a], wv.shape[a], strides[a])
            pads += [b, e]
        xc = torch.nn.functional.pad(xc, pads)
        y = torch.nn.functional.conv3d(xc, wv.permute(4, 3, 0, 1, 2), stride=strides)
        return T(y.permute(0, 2, 3, 4, 1))

    def conv3d_transpose(x, w, output_shape, strides, padding="SAME"):
        """tf.nn.conv3d_transpose: gradient of conv3d w.r.t. its input; filter [kd, kh, kw, Cout, Cin]."""
        xv, wv = _unwrap(x), _unwrap(w)
        s = [int(v) for v in strides[1:4]]
        assert padding == "SAME" and all(int(wv.shape[a]) == s[a] for a in range(3)), "only the k == stride case of the reference"
        y = torch.nn.functional.conv_transpose3d(xv.permute(0, 4, 1, 2, 3), wv.permute(4, 3, 0, 1, 2), stride=s)
        y = y.permute(0, 2, 3, 4, 1)
        want = [int(v) for v in output_shape]
        assert list(y.shape[1:4]) == want[1:4], "output_shape %r vs %r" % (want, list(y.shape))
        return T(y)

    def dropout(x, keep_prob=None, rate=None):
        r = rate if rate is not None else 1.0 - keep_prob
        assert float(r) == 0.0, "fixtures are generated with dropout 0 (tf.nn.dropout(rate=0) is the identity)"
        return x

    def conv3d(x, w, strides, padding):
        """tf.nn.conv3d, NDHWC, filter [kd, kh, kw, Cin, Cout], unit strides."""
        assert list(strides) == [1, 1, 1, 1, 1]
        xv, wv = _unwrap(x), _unwrap(w)
        if padding == "SAME":
            return convolution(x, w, "SAME")
        assert padding == "VALID"
        y = torch.nn.functional.conv3d(xv.permute(0, 4, 1, 2, 3), wv.permute(4, 3, 0, 1, 2))
        return T(y.permute(0, 2, 3, 4, 1))

    nn.convolution, nn.conv3d_transpose, nn.dropout, nn.conv3d = convolution, conv3d_transpose, dropout, conv3d
    nn.relu = lambda x: T(torch.relu(_unwrap(x)))
    nn.leaky_relu = lambda x, alpha=0.2: T(torch.nn.functional.leaky_relu(_unwrap(x), alpha))
    nn.softmax = lambda x, name=None: T(torch.softmax(_unwrap(x), dim=-1))
    nn.softmax_cross_entropy_with_logits = lambda labels=None, logits=None: T(
        -(_unwrap(labels) * torch.log_softmax(_unwrap(logits), dim=-1)).sum(dim=-1))
    tf.nn = nn

    # ---- tf.layers ----------------------------------------------------------------------------------------------
    layers = types.ModuleType("tensorflow.layers")

    def batch_normalization(x, momentum=0.99, epsilon=0.001, center=True, scale=True, training=False):
        """tf.layers.batch_normalization: a fresh BatchNormalization layer per call, variable scope
        `batch_normalization[_k]` uniquified inside the enclosing scope; build order gamma, beta, moving_mean,
        moving_variance; training mode: batch mean and biased variance over all axes but the last, UPDATE_OPS
        moving <- moving * momentum + batch * (1 - momentum)."""
        xv = _unwrap(x)
        C = int(xv.shape[-1])
        with variable_scope(None, default_name="batch_normalization"):
            gamma = get_variable("gamma", shape=(C,)) if scale else None
            beta = get_variable("beta", shape=(C,)) if center else None
            mm = get_variable("moving_mean", shape=(C,), trainable=False)
            mv = get_variable("moving_variance", shape=(C,), trainable=False)
            scope = st.path()
        axes = tuple(range(xv.dim() - 1))
        if bool(training):
            mean = xv.mean(dim=axes)
            var = ((xv - mean) ** 2).mean(dim=axes)
            st.bn_updates[scope + "/moving_mean"] = (mm.v * momentum + mean * (1 - momentum)).detach()
            st.bn_updates[scope + "/moving_variance"] = (mv.v * momentum + var * (1 - momentum)).detach()
        else:
            mean, var = mm.v, mv.v
        y = (xv - mean) / torch.sqrt(var + epsilon)
        if gamma is not None:
            y = y * gamma.v
        if beta is not None:
            y = y + beta.v
        return T(y)

    layers.batch_normalization = batch_normalization
    tf.layers = layers
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.nn"] = nn
    sys.modules["tensorflow.layers"] = layers
    return tf


def uninstall():
    for k in ("tensorflow", "tensorflow.nn", "tensorflow.layers", "networks", "layers2", "VNet", "Layers", "attention", "OutputModule"):
        sys.modules.pop(k, None)
