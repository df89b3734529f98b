"""Definition-level check of the TensorFlow op semantics the oracle encodes ("TF semantics the restatement
must encode").  TensorFlow itself cannot run here, and the oracle as well as the TF-API stand-in of the reference pin
(tests/tf1_shim.py) sit on torch's convolution routines - a shared misreading (filter layout, flip, which side the SAME
padding goes to, the transposed filter's channel order) would go unnoticed between them.  The functions below restate
each op straight from its TensorFlow-1.15 documentation formula with explicit NumPy index arithmetic and no library
convolution, and the oracle's ops are held to them in fp64:

  tf.nn.convolution / conv3d   out[b,i,j,k,q] = sum_{di,dj,dk,c} in[b, s*i+di-p0, s*j+dj-p1, s*k+dk-p2, c] * f[di,dj,dk,c,q]
                               SAME: out = ceil(in/s), pad_total = max((out-1)*s + k - in, 0), before = pad_total // 2
  tf.nn.conv3d_transpose       "the transpose (gradient) of conv3d": <conv3d(u, f), y> == <u, conv3d_transpose(y, f)>
                               with f = [kd,kh,kw, C_out_of_the_transpose, C_in_of_the_transpose]
  tf.nn.moments / tf.layers.batch_normalization(training=True)   mean and *biased* variance over all axes but the last,
                               epsilon inside the square root, moving <- moving*m + batch*(1-m)
  tf.nn.softmax, tf.one_hot (out-of-range row = 0), tf.argmax (first maximum), tf.nn.dropout (keep where u >= rate)
"""
import math

import numpy as np
import torch

from oracle import ref_vnet as R


def _conv_by_definition(x, f, stride):
    """x [B,D,H,W,C], f [kd,kh,kw,C,Q]; SAME padding by the documented rule; cross-correlation, no flip."""
    B, D, H, W, C = x.shape
    kd, kh, kw, _, Q = f.shape
    out_sz, before = [], []
    for size, k in ((D, kd), (H, kh), (W, kw)):
        o = -(-size // stride)
        total = max((o - 1) * stride + k - size, 0)
        out_sz.append(o)
        before.append(total // 2)
    y = np.zeros((B, out_sz[0], out_sz[1], out_sz[2], Q))
    for i in range(out_sz[0]):
        for j in range(out_sz[1]):
            for k in range(out_sz[2]):
                for di in range(kd):
                    for dj in range(kh):
                        for dk in range(kw):
                            a, b, c = stride * i + di - before[0], stride * j + dj - before[1], stride * k + dk - before[2]
                            if 0 <= a < D and 0 <= b < H and 0 <= c < W:      # outside: the zero padding
                                y[:, i, j, k, :] += x[:, a, b, c, :] @ f[di, dj, dk]
    return y


def test_convolution_is_cross_correlation_with_same_padding_split_low_first():
    rng = np.random.default_rng(0)
    for dims, k, stride, cin, cout in (((4, 5, 6), 5, 1, 2, 3), ((3, 4, 5), 3, 1, 3, 2), ((4, 6, 2), 2, 2, 2, 4),
                                       ((5, 3, 4), 2, 2, 1, 2), ((3, 3, 3), 1, 1, 4, 2)):
        x = rng.normal(size=(2,) + dims + (cin,))
        f = rng.normal(size=(k, k, k, cin, cout))
        b = rng.normal(size=(cout,))
        got = R.conv_same(torch.from_numpy(x), torch.from_numpy(f), torch.from_numpy(b), stride).numpy()
        want = _conv_by_definition(x, f, stride) + b
        assert got.shape == want.shape, (dims, k, stride)
        assert np.abs(got - want).max() < 1e-12, (dims, k, stride)


def test_same_padding_of_an_odd_extent_goes_to_the_far_end():
    """k = 2, stride 2 on an odd extent: pad_total = 1, before = 0 - the single zero plane is appended, not prepended
    (SURVEY H7).  An impulse in the last input plane must reach the last output plane through filter tap 0."""
    x = np.zeros((1, 5, 2, 2, 1))
    x[0, 4] = 1.0
    f = np.zeros((2, 2, 2, 1, 1))
    f[0, 0, 0] = 1.0            # tap 0 only
    got = R.conv_same(torch.from_numpy(x), torch.from_numpy(f), torch.zeros(1, dtype=torch.float64), 2).numpy()
    assert got.shape == (1, 3, 1, 1, 1) and got[0, 2, 0, 0, 0] == 1.0 and got[0, :2].sum() == 0.0
    assert R._same_pads(5, 2, 2) == (0, 1) and R._same_pads(6, 5, 1) == (2, 2) and R._same_pads(7, 3, 2) == (1, 1)


def test_conv3d_transpose_is_the_adjoint_of_the_strided_convolution():
    """layers2.py:65-74,88-94: filter [2,2,2,c,2c] read by conv3d_transpose as [kd,kh,kw,out,in]; as a forward conv3d
    filter it maps c (fine) -> 2c (coarse) with stride 2.  The adjoint identity pins layout and tap order at once, and
    the scatter form of SURVEY a3 follows from it."""
    rng = np.random.default_rng(1)
    c_fine, c_coarse = 3, 5
    f = rng.normal(size=(2, 2, 2, c_fine, c_coarse))
    u = rng.normal(size=(2, 4, 6, 2, c_fine))          # fine grid
    y = rng.normal(size=(2, 2, 3, 1, c_coarse))        # coarse grid
    zero = torch.zeros(c_fine, dtype=torch.float64)
    up = R.deconv_k2s2(torch.from_numpy(y), torch.from_numpy(f), zero, (4, 6, 2)).numpy()
    lhs = (_conv_by_definition(u, f, 2) * y).sum()
    rhs = (u * up).sum()
    assert abs(lhs - rhs) < 1e-10 * max(1.0, abs(lhs))
    # scatter form: out[n,2i+a,2j+b,2k+d,co] = sum_ci x[n,i,j,k,ci] * w[a,b,d,co,ci]
    want = np.zeros_like(up)
    for i in range(2):
        for j in range(3):
            for a in range(2):
                for b in range(2):
                    for d in range(2):
                        want[:, 2 * i + a, 2 * j + b, d, :] = y[:, i, j, 0, :] @ f[a, b, d].T
    assert np.abs(up - want).max() < 1e-12
    bias = rng.normal(size=(c_fine,))                  # layers2.py:90: the bias has filter[-2] = c entries
    up_b = R.deconv_k2s2(torch.from_numpy(y), torch.from_numpy(f), torch.from_numpy(bias), (4, 6, 2)).numpy()
    assert np.abs(up_b - (want + bias)).max() < 1e-12


def test_training_mode_batch_norm_uses_biased_moments_and_epsilon_under_the_root():
    rng = np.random.default_rng(2)
    x = rng.normal(2.0, 3.0, size=(2, 3, 2, 4, 3))
    gamma, beta = rng.normal(size=3), rng.normal(size=3)
    mm, mv = rng.normal(size=3), rng.uniform(0.5, 2.0, size=3)
    scope = "s"
    p = {scope + "/batch_normalization/" + k: torch.from_numpy(v)
         for k, v in (("gamma", gamma), ("beta", beta), ("moving_mean", mm), ("moving_variance", mv))}
    ctx = R._Ctx(p, None)
    got = ctx.bn(torch.from_numpy(x), scope).numpy()
    flat = x.reshape(-1, 3)
    n = flat.shape[0]
    mean = flat.sum(0) / n
    var = ((flat - mean) ** 2).sum(0) / n                                   # biased: / n, not / (n - 1)
    want = (x - mean) / np.sqrt(var + 1e-3) * gamma + beta                  # epsilon = 1e-3 (networks.py:259), inside
    assert np.abs(got - want).max() < 1e-12
    assert np.abs(ctx.updates[scope + "/batch_normalization/moving_mean"].numpy() - (0.99 * mm + 0.01 * mean)).max() < 1e-14
    assert np.abs(ctx.updates[scope + "/batch_normalization/moving_variance"].numpy() - (0.99 * mv + 0.01 * var)).max() < 1e-14
    unbiased = var * n / (n - 1)
    assert np.abs(got - ((x - mean) / np.sqrt(unbiased + 1e-3) * gamma + beta)).max() > 1e-4   # the test can tell them apart


def test_prelu_softmax_one_hot_argmax_dropout_by_definition():
    rng = np.random.default_rng(3)
    x = rng.normal(size=(2, 2, 2, 2, 3))
    alpha = np.array([0.1, -0.3, 2.0])
    want = np.maximum(0, x) + alpha * np.minimum(0, x)                        # layers2.py:97-99
    assert np.abs(R.prelu(torch.from_numpy(x), torch.from_numpy(alpha)).numpy() - want).max() == 0.0
    # softmax over the last axis; argmax = lowest index among equal maxima; one_hot of an out-of-range label is all zero
    logits = np.array([[[[[1.0, 3.0, 3.0], [0.0, 0.0, 0.0]]]]])
    e = np.exp(logits - logits.max(-1, keepdims=True))
    labels = np.array([[[[1, 7]]]], np.int32)
    terms = R.dice_terms(torch.from_numpy(logits), torch.from_numpy(labels), "sorensen").numpy()   # [N,K,(I,L,R)]
    sm = e / e.sum(-1, keepdims=True)
    onehot = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 0.0]])                     # label 7 of 3 classes: no class
    assert np.abs(terms[0, :, 0] - (sm[0, 0, 0] * onehot).sum(0)).max() < 1e-15
    assert np.abs(terms[0, :, 1] - sm[0, 0, 0].sum(0)).max() < 1e-15
    assert np.abs(terms[0, :, 2] - onehot.sum(0)).max() == 0.0
    assert R.predict(torch.from_numpy(logits)).numpy().tolist() == [[[[1, 0]]]]
    # dropout: x * keep / (1 - rate), keep = (u >= rate); rate 0 is the identity
    keep = (rng.uniform(size=x.shape) >= 0.25).astype(np.float64)
    got = R._dropout(torch.from_numpy(x), 0.25, {"k": torch.from_numpy(keep)}, "k").numpy()
    assert np.abs(got - x * keep / 0.75).max() < 1e-15
    assert R._dropout(torch.from_numpy(x), 0.0, None, "k") is not None and \
        np.array_equal(R._dropout(torch.from_numpy(x), 0.0, None, "k").numpy(), x)


def test_adam_and_learning_rate_by_the_documented_formulas():
    """tf.train.AdamOptimizer docs: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t); m_t = b1 m + (1 - b1) g;
    v_t = b2 v + (1 - b2) g^2; var -= lr_t * m_t / (sqrt(v_t) + eps).  exponential_decay: lr0 * rate^(step / steps)."""
    rng = np.random.default_rng(4)
    p, g = rng.normal(size=5), rng.normal(size=5)
    m, v = np.zeros(5), np.zeros(5)
    tp, tm, tv = torch.from_numpy(p), torch.from_numpy(m), torch.from_numpy(v)
    for t in (1, 2, 3):
        lr = 1e-2 * 0.99 ** ((t - 1) / 100.0)
        assert abs(R.learning_rate(1e-2, t - 1, 100.0, 0.99) - lr) < 1e-18
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        p = p - lr * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m / (np.sqrt(v) + 1e-8)
        tp, tm, tv = R.adam_update(tp, torch.from_numpy(g), tm, tv, t, lr)
        assert np.abs(tp.numpy() - p).max() < 1e-15
