"""Bodies of the per-op hook tests, shared by the CPU-emulation twins (tests/test_engine_emul.py) and the
GPU tests (tests/test_gpu_parity.py): each takes a loaded library and compares one vnb_op_* call with the oracle."""
import ctypes as C

import numpy as np
import torch

from oracle import ref_vnet as R
from tests.helpers import rel_err
from vnet_tensorflow_b200 import _ffi

_ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)


def check_k2_ops(lib, precision, cf, cc, dims, n=2, tol=None):
    """2x2x2 stride-2 down convolution, transposed up convolution and their filter gradient (layers2.py:65-94)."""
    prec = _ffi.PRECISIONS[precision]
    tol = tol or (1e-5 if precision == "fp32" else 3e-5)
    rng = np.random.default_rng(11)
    dc, hc, wc = dims
    fine = rng.normal(0, 1, (n, 2 * dc, 2 * hc, 2 * wc, cf)).astype(np.float32)
    coarse = rng.normal(0, 1, (n, dc, hc, wc, cc)).astype(np.float32)
    w = rng.normal(0, 0.2, (2, 2, 2, cf, cc)).astype(np.float32)
    b_c = rng.normal(0, 1, (cc,)).astype(np.float32)
    b_f = rng.normal(0, 1, (cf,)).astype(np.float32)
    # op 0: down convolution (fine -> coarse), tf.nn.convolution stride 2
    xt = torch.from_numpy(fine).double().requires_grad_(True)
    wt = torch.from_numpy(w).double().requires_grad_(True)
    y_ref = R.conv_same(xt, wt, torch.from_numpy(b_c).double(), stride=2)
    out = np.empty_like(coarse)
    lib.check(lib.vnb_op_k2(0, prec, 0, _ptr(fine), None, _ptr(w), _ptr(b_c), _ptr(out), n, dc, hc, wc, cf, cc))
    assert rel_err(out, y_ref.detach().numpy()) < tol
    # op 2: its filter gradient against autograd
    y_ref.backward(torch.from_numpy(coarse).double())
    dw = np.empty_like(w)
    lib.check(lib.vnb_op_k2(0, prec, 2, _ptr(fine), _ptr(coarse), None, None, _ptr(dw), n, dc, hc, wc, cf, cc))
    assert rel_err(dw, wt.grad.numpy()) < 5 * tol
    # op 1: transposed convolution (coarse -> fine) with the [2,2,2,out,in] filter of layers2.py:92 = the down conv's dgrad
    up_ref = R.deconv_k2s2(torch.from_numpy(coarse).double(), torch.from_numpy(w).double(), torch.from_numpy(b_f).double(),
                           (2 * dc, 2 * hc, 2 * wc))
    out_f = np.empty_like(fine)
    lib.check(lib.vnb_op_k2(0, prec, 1, None, _ptr(coarse), _ptr(w), _ptr(b_f), _ptr(out_f), n, dc, hc, wc, cf, cc))
    assert rel_err(out_f, up_ref.numpy()) < tol
    lib.check(lib.vnb_op_k2(0, prec, 1, None, _ptr(coarse), _ptr(w), None, _ptr(out_f), n, dc, hc, wc, cf, cc))
    assert rel_err(out_f, xt.grad.numpy()) < tol           # without bias: exactly the input gradient of op 0


def check_bn_ops(lib, voxels, c, with_alpha=True):
    """Training-mode batch norm (+ PReLU) forward and backward (networks.py:319, layers2.py:97-99)."""
    rng = np.random.default_rng(5)
    z = rng.normal(0.3, 2.0, (voxels, c)).astype(np.float32)
    gamma = rng.uniform(0.5, 1.5, c).astype(np.float32)
    beta = rng.normal(0, 0.5, c).astype(np.float32)
    alpha = rng.uniform(0.05, 0.3, c).astype(np.float32) if with_alpha else None
    dy = rng.normal(0, 1, (voxels, c)).astype(np.float32)
    zt = torch.from_numpy(z).double().requires_grad_(True)
    gt, bt = torch.from_numpy(gamma).double().requires_grad_(True), torch.from_numpy(beta).double().requires_grad_(True)
    mu, var = zt.mean(0), zt.var(0, unbiased=False)
    yhat = gt * (zt - mu) / torch.sqrt(var + 1e-3) + bt
    if with_alpha:
        at = torch.from_numpy(alpha).double().requires_grad_(True)
        y_ref = R.prelu(yhat, at)
    else:
        y_ref = yhat
    y_ref.backward(torch.from_numpy(dy).double())
    y = np.empty_like(z)
    mean, v = np.empty(c, np.float64), np.empty(c, np.float64)
    lib.check(lib.vnb_op_bn_fwd(0, _ptr(z), _ptr(gamma), _ptr(beta), _ptr(alpha), _ptr(y), _ptr(mean), _ptr(v), voxels, c))
    assert rel_err(y, y_ref.detach().numpy()) < 2e-5
    assert rel_err(mean, mu.detach().numpy()) < 1e-5 and rel_err(v, var.detach().numpy()) < 1e-5
    dz, dg, db, da = np.empty_like(z), np.empty(c, np.float32), np.empty(c, np.float32), np.empty(c, np.float32)
    lib.check(lib.vnb_op_bn_bwd(0, _ptr(z), _ptr(dy), _ptr(gamma), _ptr(beta), _ptr(alpha), _ptr(dz), _ptr(dg), _ptr(db),
                                _ptr(da) if with_alpha else None, voxels, c))
    assert rel_err(dz, zt.grad.numpy()) < 1e-4
    assert rel_err(dg, gt.grad.numpy()) < 1e-4 and rel_err(db, bt.grad.numpy()) < 1e-4
    if with_alpha:
        assert rel_err(da, at.grad.numpy()) < 1e-4


def check_softmax_dice_ops(lib, loss, weights, alpha, n=2, voxels=4096, k=3):
    """softmax / one-hot / loss zoo / argmax (model.py:26-92,447,477,495-568) forward and backward."""
    rng = np.random.default_rng(9)
    logits = rng.normal(0, 2, (n, voxels, 1, 1, k)).astype(np.float32)
    labels = rng.integers(0, k, (n, voxels, 1, 1)).astype(np.int32)
    lt = torch.from_numpy(logits).double().requires_grad_(True)
    ref = R.loss_from_logits(lt, torch.from_numpy(labels), loss, weights, alpha)
    ref.backward()
    w = np.asarray(weights if len(weights) else [1.0] * k, np.float32)
    out = C.c_float()
    sm, am = np.empty((n, voxels, k), np.float32), np.empty((n, voxels), np.int64)
    terms = np.empty((n, k, 4), np.float64)
    code = _ffi.LOSSES[loss]
    lib.check(lib.vnb_op_softmax_dice_fwd(0, _ptr(logits), _ptr(labels), n, voxels, k, code, _ptr(w), alpha, C.byref(out), _ptr(sm),
                                          _ptr(am), _ptr(terms)))
    assert abs(out.value - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    assert rel_err(sm, torch.softmax(lt.detach().reshape(n, voxels, k), -1).numpy()) < 1e-5
    assert np.array_equal(am, np.argmax(logits.reshape(n, voxels, k), -1))
    assert np.allclose(terms[..., 2].sum(1), voxels)          # one-hot mass
    dl = np.empty((n, voxels, k), np.float32)
    lib.check(lib.vnb_op_softmax_dice_bwd(0, _ptr(logits), _ptr(labels), n, voxels, k, code, _ptr(w), alpha, _ptr(dl)))
    assert rel_err(dl, lt.grad.numpy().reshape(n, voxels, k)) < 2e-4


def check_adam_op(lib, count=10007):
    """tf.train.AdamOptimizer in TF's epsilon-hat form (model.py:652), three steps with a decaying learning rate."""
    rng = np.random.default_rng(2)
    p = rng.normal(0, 1, count).astype(np.float32)
    m, v = np.zeros(count, np.float32), np.zeros(count, np.float32)
    pt, mt, vt = torch.from_numpy(p.copy()), torch.zeros(count), torch.zeros(count)
    for t in range(1, 4):
        g = rng.normal(0, 0.1, count).astype(np.float32)
        lr = R.learning_rate(1e-2, t - 1, 2.0, 0.5)
        pt, mt, vt = R.adam_update(pt, torch.from_numpy(g), mt, vt, t, lr)
        lib.check(lib.vnb_op_adam(0, _ptr(p), _ptr(g), _ptr(m), _ptr(v), count, lr, t))
    # v: the kernel forms (1 - beta2) in float32 like TensorFlow's apply_adam (0.00100005), the oracle in double (0.001)
    assert rel_err(p, pt.numpy()) < 1e-6 and rel_err(m, mt.numpy()) < 1e-6 and rel_err(v, vt.numpy()) < 5e-5
