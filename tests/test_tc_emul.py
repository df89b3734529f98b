"""The tcgen05 implicit-GEMM convolution kernel on the CPU: compiled with the emulation shim, its TMA
boxes, mbarrier pipeline, UMMA descriptors, TMEM double buffering and kw-folding epilogue run against
the probe-validated model of the hardware (tests/emul/emul_sm100.h) and are compared with torch."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import ref_vnet as R
from tests.helpers import analytically_zero, engine_for, perturbed_params, rel_err
from vnet_tensorflow_b200.synthetic import synth_batch


def _bf16(x):
    return torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()


@pytest.mark.parametrize("cin,cout,dims,n,prec", [
    (16, 16, (2, 4, 32), 1, 2),     # CT=16, T=3 tiles, SW32
    (16, 16, (2, 7, 128), 1, 1),    # full 128-wide lines, partial last h-block, bf16x3
    (32, 32, (2, 4, 16), 2, 1),     # CT=32, SW64, two k-steps per stage
    (32, 16, (3, 8, 16), 1, 2),     # two output slices in dgrad
    (64, 32, (4, 8, 8), 1, 1),      # box spanning two d-planes
    (16, 16, (2, 5, 24), 1, 1),     # W does not divide 128: five whole lines per MMA tile, tile stride 120 rows
    (32, 32, (3, 4, 48), 1, 2),     # two lines per tile (96 of 128 rows used)
    (16, 32, (2, 3, 192), 1, 1),    # W > 128: two haloed 96-wide segments per line
    (32, 16, (2, 4, 160), 1, 2),    # W > 128, CT = 32 forward / 16 backward
    (16, 16, (3, 5, 16), 1, 1),     # odd H: the box (5 lines x 4 planes) leaves MMA rows unused, last d-block partial
    (32, 16, (2, 6, 64), 2, 1),     # fprop: two 16-channel k-chunks accumulate; dgrad 16 -> 32: two slices (column kernel), bf16x3
    (16, 16, (19, 3, 128), 1, 1),   # column kernel: 19 planes walked with five accumulators in flight, slot ring wraps, bf16x3
    (16, 16, (40, 2, 64), 1, 2),    # column kernel: two 20-plane segments per column (halo planes re-loaded), two lines per tile
])
def test_conv5_tc_kernel_matches_torch(emul_lib, cin, cout, dims, n, prec):
    rng = np.random.default_rng(7)
    x = rng.normal(0, 1, (n,) + dims + (cin,)).astype(np.float32)
    w = rng.normal(0, 0.05, (5, 5, 5, cin, cout)).astype(np.float32)
    b = rng.normal(0, 1, (cout,)).astype(np.float32)
    r = rng.normal(0, 1, (n,) + dims + (cout,)).astype(np.float32)
    dy = rng.normal(0, 1, (n,) + dims + (cout,)).astype(np.float32)
    xr, wr, dyr = (_bf16(x), _bf16(w), _bf16(dy)) if prec == 2 else (x, w, dy)  # bf16 mode: exact vs rounded inputs
    tol = 1e-6 if prec == 2 else 3e-5
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    y_ref = (R.conv_same(torch.from_numpy(xr).double(), torch.from_numpy(wr).double(), torch.from_numpy(b).double())
             + torch.from_numpy(r).double()).numpy()
    y = np.empty_like(dy)
    emul_lib.check(emul_lib.vnb_op_conv5_fprop(0, prec, ptr(x), ptr(w), ptr(b), ptr(r), ptr(y), n, *dims, cin, cout))
    assert rel_err(y, y_ref) < tol
    xt = torch.from_numpy(x).double().requires_grad_(True)
    R.conv_same(xt, torch.from_numpy(wr).double(), torch.zeros(cout).double()).backward(torch.from_numpy(dyr).double())
    dx = np.empty_like(x)
    emul_lib.check(emul_lib.vnb_op_conv5_dgrad(0, prec, ptr(dy), ptr(w), ptr(dx), n, *dims, cin, cout))
    assert rel_err(dx, xt.grad.numpy()) < tol


def test_unsupported_shapes_are_rejected_not_miscomputed(emul_lib):
    x = np.zeros((1, 3, 5, 16, 8), np.float32)
    w = np.zeros((5, 5, 5, 8, 16), np.float32)
    y = np.zeros((1, 3, 5, 16, 16), np.float32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = emul_lib.vnb_op_conv5_fprop(0, 1, ptr(x), ptr(w), None, None, ptr(y), 1, 3, 5, 16, 8, 16)   # Cin = 8
    assert rc == -1 and b"not supported" in emul_lib.vnb_last_error()


def test_engine_bf16x3_matches_oracle(emul_lib):
    """Whole network with tensor-core convolutions where the geometry allows (levels 1-2 here) and the
    fp32 kernels elsewhere: concat inputs, residual epilogue, split dgrad outputs, accumulate flags."""
    spec = R.VNetSpec(num_classes=2, in_channels=1, num_channels=16, num_levels=2, num_convolutions=(1, 2), bottom_convolutions=2)
    P, N = 16, 1
    params = perturbed_params(spec)
    img, lab = synth_batch(0, N, P, 1, 2)
    eng = engine_for(spec, P, N, "weighted_sorensen", (0.1, 1.0), emul_lib, precision="bf16x3")
    eng.set_params(params)
    l = eng.forward_backward(img, lab)
    lo, lg, go, _ = R.loss_and_grads(params, img, lab, spec, "weighted_sorensen", (0.1, 1.0))
    logits, _, am = eng.forward(img)
    assert abs(l - float(lo)) < 2e-6
    assert rel_err(logits, lg.numpy()) < 1e-4          # north_star bound is 1e-3
    assert int((am != R.predict(lg).numpy()).sum()) == 0
    g = eng.get_grads()
    for k, v in g.items():                              # conditioning floor, see test_gpu_parity.GRAD_TOL
        if analytically_zero(k, spec):
            continue
        ref = go[k].numpy().astype(np.float64)
        assert np.sqrt(((v - ref) ** 2).sum()) <= 5e-2 * max(np.sqrt((ref ** 2).sum()), 1e-9), k
    eng.close()


def test_engine_bf16x3_short_batch_after_a_full_one(emul_lib):
    """Tensor-core plans are built for max_batch; a shorter batch (last batch of an epoch) only shrinks the item count
    and the statistics count - nothing of the earlier full batch may leak into it (TMA boxes, split-K partials,
    bf16 copies of patch 1 are still in memory)."""
    spec = R.VNetSpec(num_classes=2, in_channels=1, num_channels=16, num_levels=2, num_convolutions=(1, 2), bottom_convolutions=2)
    P = 16
    params = perturbed_params(spec)
    eng = engine_for(spec, P, 2, "weighted_sorensen", (0.1, 1.0), emul_lib, precision="bf16x3")
    eng.set_params(params)
    img2, lab2 = synth_batch(0, 2, P, 1, 2)
    l2 = eng.forward_backward(img2, lab2)
    lo2 = R.loss_and_grads(params, img2, lab2, spec, "weighted_sorensen", (0.1, 1.0))[0]
    assert abs(l2 - float(lo2)) < 2e-6
    img, lab = synth_batch(5, 1, P, 1, 2)
    l = eng.forward_backward(img, lab)
    lo, lg, go, _ = R.loss_and_grads(params, img, lab, spec, "weighted_sorensen", (0.1, 1.0))
    assert abs(l - float(lo)) < 2e-6
    g = eng.get_grads()
    for k, v in g.items():
        if analytically_zero(k, spec):
            continue
        ref = go[k].numpy().astype(np.float64)
        assert np.sqrt(((v - ref) ** 2).sum()) <= 5e-2 * max(np.sqrt((ref ** 2).sum()), 1e-9), k
    logits, _, am = eng.forward(img)
    assert rel_err(logits, lg.numpy()) < 1e-4 and int((am != R.predict(lg).numpy()).sum()) == 0
    eng.close()


@pytest.mark.parametrize("cin,cout,dims,n,prec", [
    (16, 16, (3, 6, 16), 1, 2),     # one (ci,co) pair, split-K over slabs
    (16, 16, (2, 5, 32), 2, 1),     # odd H (partial last slab), bf16x3
    (32, 16, (2, 4, 64), 1, 2),     # two ci chunks
    (16, 32, (3, 8, 8), 1, 1),      # W = 8: one K step spans two lines
    (16, 16, (2, 3, 128), 1, 2),    # full-width lines
    (16, 16, (2, 4, 24), 1, 1),     # W not a multiple of the K step: boxes rounded to 32, TMA zero fill past the line
    (16, 16, (2, 3, 12), 1, 2),     # W = 12 -> one K step of 16
    (16, 16, (1, 3, 192), 1, 2),    # W > 128
    (32, 32, (2, 4, 128), 1, 1),    # N = 160 (two dZ chunks), bf16x3: lines cut into two 64-voxel K segments, two plane-offset groups
    (64, 16, (3, 6, 32), 2, 1),     # Cout = 16, Cin = 64: roles swapped (P = dZ, Q = two X chunks per CTA), mirrored taps
    (16, 64, (6, 5, 16), 1, 2),     # two Q chunk groups, odd H, D > plane-offset group size
    (32, 32, (2, 8, 8), 1, 1),      # W = 8 with two interleaved chunks: K rows 8..15 come from the next line
    (128, 128, (3, 8, 8), 1, 1),    # deep-level kernel (wgrad_deep.cuh): per-tap GEMM, one K tile per plane, bf16x3
    (128, 256, (2, 4, 16), 2, 2),   # deep-level kernel: N = 256 accumulators, four-line K tiles, bf16
])
def test_wgrad5_tc_kernel_matches_torch(emul_lib, cin, cout, dims, n, prec):
    """MN-major overlapping-atom tap folding (kw on M, kh on N, kd on five TMEM accumulators)."""
    rng = np.random.default_rng(3)
    x = rng.normal(0, 1, (n,) + dims + (cin,)).astype(np.float32)
    dy = rng.normal(0, 1, (n,) + dims + (cout,)).astype(np.float32)
    xr, dyr = (_bf16(x), _bf16(dy)) if prec == 2 else (x, dy)
    wt = torch.zeros(5, 5, 5, cin, cout, dtype=torch.float64, requires_grad=True)
    R.conv_same(torch.from_numpy(xr).double(), wt, torch.zeros(cout).double()).backward(torch.from_numpy(dyr).double())
    dw = np.empty((5, 5, 5, cin, cout), np.float32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    emul_lib.check(emul_lib.vnb_op_conv5_wgrad(0, prec, ptr(x), ptr(dy), ptr(dw), n, *dims, cin, cout))
    assert rel_err(dw, wt.grad.numpy()) < (1e-6 if prec == 2 else 3e-5)


def test_wgrad5_deep_kernel_split_k_matches_torch(emul_lib, monkeypatch):
    """wgrad_deep.cuh with K splits (partials + fixed-order reduce) and a second ci block."""
    monkeypatch.setenv("VNB_WD_KSPLIT", "2")
    cin, cout, dims, n = 256, 128, (3, 8, 8), 1
    rng = np.random.default_rng(4)
    x = rng.normal(0, 1, (n,) + dims + (cin,)).astype(np.float32)
    dy = rng.normal(0, 1, (n,) + dims + (cout,)).astype(np.float32)
    wt = torch.zeros(5, 5, 5, cin, cout, dtype=torch.float64, requires_grad=True)
    R.conv_same(torch.from_numpy(_bf16(x)).double(), wt, torch.zeros(cout).double()).backward(torch.from_numpy(_bf16(dy)).double())
    dw = np.empty((5, 5, 5, cin, cout), np.float32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    emul_lib.check(emul_lib.vnb_op_conv5_wgrad(0, 2, ptr(x), ptr(dy), ptr(dw), n, *dims, cin, cout))
    assert rel_err(dw, wt.grad.numpy()) < 1e-6


def test_multimodal_input_runs_on_tensor_cores_with_padded_channels(emul_lib):
    """in_channels = 3 (not a multiple of 16): the input convolution and its filter gradient run on the
    tensor-core kernels over zero-padded bf16 copies of the image (BASELINE config #3 shape)."""
    spec = R.VNetSpec(num_classes=3, in_channels=3, num_channels=16, num_levels=1, num_convolutions=(1,), bottom_convolutions=1)
    P, N = 16, 1
    params = perturbed_params(spec)
    img, lab = synth_batch(4, N, P, 3, 3)
    eng = engine_for(spec, P, N, "weighted_sorensen", (0.1, 0.5, 1.0), emul_lib, precision="bf16x3")
    eng.set_params(params)
    l = eng.forward_backward(img, lab)
    lo, lg, go, _ = R.loss_and_grads(params, img, lab, spec, "weighted_sorensen", (0.1, 0.5, 1.0))
    logits, _, am = eng.forward(img)
    assert abs(l - float(lo)) < 5e-6
    assert rel_err(logits, lg.numpy()) < 1e-4
    k = "vnet/input_layer/weights"
    g = eng.get_grads()[k]
    ref = go[k].numpy()
    assert g.shape == (5, 5, 5, 3, 16)
    assert np.sqrt(((g - ref) ** 2).sum()) <= 2e-2 * np.sqrt((ref ** 2).sum())
    eng.close()


@pytest.mark.parametrize("cin,cout,dims,n,prec", [
    (16, 16, (2, 4, 32), 1, 2),     # CT=16, three tiles, streaming + resident
    (64, 64, (3, 8, 16), 1, 1),     # the attention-module shape: 64 -> 64, CT=32, two slices, two k-chunks, bf16x3
    (32, 64, (2, 5, 64), 2, 2),     # partial last h-block
    (2, 64, (3, 4, 8), 1, 0),       # first module conv (K -> 64): exact fp32 kernel
])
def test_conv3_kernels_match_torch(emul_lib, cin, cout, dims, n, prec):
    """3^3 SAME convolution of the attention / output modules (attention.py:63-92): forward, input gradient,
    filter gradient -- tensor-core kernel with KS = 3 (fprop / dgrad) and the fp32 kernels."""
    import torch.nn.functional as F
    rng = np.random.default_rng(11)
    x = rng.normal(0, 1, (n,) + dims + (cin,)).astype(np.float32)
    w = rng.normal(0, 0.1, (3, 3, 3, cin, cout)).astype(np.float32)
    b = rng.normal(0, 1, (cout,)).astype(np.float32)
    dy = rng.normal(0, 1, (n,) + dims + (cout,)).astype(np.float32)
    xr, wr, dyr = (_bf16(x), _bf16(w), _bf16(dy)) if prec == 2 else (x, w, dy)
    tol = 1e-6 if prec == 2 else 3e-5
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)

    def conv(xt, wt):
        return F.conv3d(xt.permute(0, 4, 1, 2, 3), wt.permute(4, 3, 0, 1, 2), padding=1).permute(0, 2, 3, 4, 1)

    xt = torch.from_numpy(xr).double().requires_grad_(True)
    wt = torch.from_numpy(wr).double().requires_grad_(True)
    y_ref = conv(xt, wt) + torch.from_numpy(b).double()
    y = np.empty_like(dy)
    emul_lib.check(emul_lib.vnb_op_conv3_fprop(0, prec, ptr(x), ptr(w), ptr(b), None, ptr(y), n, *dims, cin, cout))
    assert rel_err(y, y_ref.detach().numpy()) < tol
    y_ref.backward(torch.from_numpy(dyr).double())
    dx = np.empty_like(x)
    emul_lib.check(emul_lib.vnb_op_conv3_dgrad(0, prec, ptr(dy), ptr(w), ptr(dx), n, *dims, cin, cout))
    assert rel_err(dx, xt.grad.numpy()) < tol
    if True:
        dw = np.empty_like(w)
        emul_lib.check(emul_lib.vnb_op_conv3_wgrad(0, prec, ptr(x), ptr(dy), ptr(dw), n, *dims, cin, cout))
        assert rel_err(dw, wt.grad.numpy()) < (3e-5 if prec != 2 else 1e-5)


def test_attention_engine_bf16x3_matches_oracle(emul_lib):
    """Gated network (row a15) with the module 3^3 convolutions on the tensor-core kernel (fprop + dgrad),
    fp32 filter gradients, inference-mode batch norms, gate and attention loss."""
    from tests.helpers import perturbed_attention_params
    from vnet_tensorflow_b200.synthetic import synth_patch
    spec = R.VNetSpec(num_classes=2, in_channels=1, num_channels=16, num_levels=1, num_convolutions=(1,),
                      bottom_convolutions=1, flavour="legacy")
    P, N, nch = 8, 1, 16
    params = perturbed_attention_params(spec, nch)
    im, lb, dm = (a[None] for a in synth_patch(3, P, 1, 2))
    eng = engine_for(spec, P, N, "jaccard", (), emul_lib, precision="bf16x3", attention=True, attention_loss="l2",
                     module_channels=nch)
    eng.set_params(params)
    eng.set_distmap(dm)
    tot, seg, att, out, go, _ = R.attention_loss_and_grads(params, im, lb, dm, spec, "jaccard", "l2")
    l = eng.forward_backward(im, lb)
    logits, _, am = eng.forward(im)
    assert abs(l - float(tot)) < 1e-4 * max(1.0, abs(float(tot)))
    assert rel_err(logits, out["logits_output"].numpy()) < 1e-4
    assert int((am != R.predict(out["logits_output"]).numpy()).sum()) == 0
    g = eng.get_grads()
    for k, v in g.items():
        if analytically_zero(k, spec):
            continue
        ref = go[k].numpy().astype(np.float64)
        assert np.sqrt(((v - ref) ** 2).sum()) <= 5e-2 * max(np.sqrt((ref ** 2).sum()), 1e-9), k
    eng.close()


def test_engine_patch_24_runs_on_tensor_cores(emul_lib):
    """PatchShape not a power of two (W = 24 and 12 along the levels, as in BASELINE config #5's 192 -> 96 -> 48 -> 24
    -> 12): every 5^3 convolution still takes the tensor-core path (lines that do not divide the 128-row MMA tile)."""
    spec = R.VNetSpec(num_classes=2, in_channels=1, num_channels=16, num_levels=1, num_convolutions=(2,), bottom_convolutions=1)
    P, N = 24, 1
    params = perturbed_params(spec)
    img, lab = synth_batch(2, N, P, 1, 2)
    eng = engine_for(spec, P, N, "weighted_sorensen", (0.1, 1.0), emul_lib, precision="bf16x3")
    eng.set_params(params)
    l = eng.forward_backward(img, lab)
    lo, lg, go, _ = R.loss_and_grads(params, img, lab, spec, "weighted_sorensen", (0.1, 1.0))
    logits, _, am = eng.forward(img)
    assert abs(l - float(lo)) < 2e-6
    assert rel_err(logits, lg.numpy()) < 1e-4
    assert int((am != R.predict(lg).numpy()).sum()) == 0
    for k, v in eng.get_grads().items():
        if analytically_zero(k, spec):
            continue
        ref = go[k].numpy().astype(np.float64)
        assert np.sqrt(((v - ref) ** 2).sum()) <= 5e-2 * max(np.sqrt((ref ** 2).sum()), 1e-9), k
    eng.close()
