"""Parity tests proper (run on the B200 with -m gpu): the CUDA engine, called through the C ABI,
against the CPU oracle on the same seeded inputs and against the committed golden fixtures.

Tolerances (BASELINE.json north_star): logits within 1e-3 relative fp32 (relative to the logit scale),
bit-exact argmax label volumes, Dice terms/loss to fp32 rounding.  The fp32 and bf16x3 precisions are
held to that bar; bf16 is the documented reduced-precision fast mode (looser bound, measured)."""
import numpy as np
import pytest
import torch

from oracle import ref_vnet as R
from tests.helpers import (CASES, analytically_zero, argmax_parity, assert_argmax_parity, engine_for, load_golden,
                           perturbed_params, rel_err)
from vnet_tensorflow_b200.synthetic import synth_batch

pytestmark = pytest.mark.gpu

LOGIT_TOL = {"fp32": 2e-4, "bf16x3": 1e-3, "bf16": 8e-2}
# Gradients of this network are ill-conditioned in the max norm: a 1e-5 relative perturbation of one
# layer's weights moves the *exact* (fp64) gradients by ~3e-2 while the logits move by 1e-5, because a
# handful of PReLU inputs cross zero (measured with the oracle, see DESIGN.md "conditioning").  The
# fp32 path (same engine, exact-fp32 convolutions) is therefore held to a tight bound and pins the
# engine logic; the tensor-core precisions get a bound at the conditioning floor, and their kernels
# are pinned tightly per op in test_conv5_ops_match_torch.
GRAD_TOL = {"fp32": 5e-3, "bf16x3": 8e-2, "bf16": 0.5}


def _report(line):
    """Measured parity numbers of the full-size tests: printed, and appended to gpurun_out/parity_report.txt when that
    directory exists (the file is copied to profiles/ as round evidence)."""
    import os
    print(line)
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_report.txt"), "a") as f:
            f.write(line + "\n")


def _check_grads(eng, grads_ref, spec, tol, l2=False):
    g = eng.get_grads()
    scale = max(float(np.abs(v).max()) for v in grads_ref.values())
    worst = 0.0
    for k, v in g.items():
        ref = grads_ref[k]
        if analytically_zero(k, spec):
            assert np.abs(v).max() <= 1e-5 * scale + 1e-12, k
            continue
        if l2:  # per-tensor relative L2 error: robust against single PReLU sign flips
            err = np.sqrt(((v.astype(np.float64) - ref) ** 2).sum()) / max(np.sqrt((ref.astype(np.float64) ** 2).sum()), 1e-12)
        else:
            err = np.abs(v - ref).max() / max(np.abs(ref).max(), 1e-3 * scale)
        worst = max(worst, err)
        assert err <= tol, (k, err)
    return worst


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("name", ["tiny_m1_k2", "tiny_m2_k3", "default_m1_k2_p32"])
def test_golden_fixtures(gpu_lib, name, precision):
    kw, P, N, loss, weights = CASES[name]
    spec = R.VNetSpec(**kw)
    gold = load_golden(name)
    img, lab = synth_batch(0, N, P, spec.in_channels, spec.num_classes)
    eng = engine_for(spec, P, N, loss, weights, gpu_lib, precision=precision)
    eng.set_params(perturbed_params(spec))
    logits, _, argmax = eng.forward(img)
    assert rel_err(logits, gold["logits"]) < LOGIT_TOL[precision]
    flips = (argmax != gold["argmax"])
    if precision == "bf16":
        assert flips.mean() < 0.02
    else:  # bit-exact label volume (north_star) up to ties inside the rounding band of the logits
        assert_argmax_parity(argmax, logits, gold["logits"], gold["argmax"])
    l, terms = eng.loss(img, lab, want_terms=True)
    assert abs(l - float(gold["loss"])) < (5e-3 if precision == "bf16" else 5e-5)
    if precision != "bf16":
        assert rel_err(terms[..., :3], gold["dice_terms"]) < 2e-4
    eng.forward_backward(img, lab)
    g = eng.get_grads()
    scale = max(float(v) for k, v in gold.items() if k.startswith("gnorm/"))
    # default_m1_k2_p32 normalises the bottom level over 2^3 = 8 voxels per channel: batch norm over so
    # few samples amplifies fp32 rounding differences (summation order) by orders of magnitude, so its
    # element-wise gradient samples are only held to a few percent; the tiny_* nets keep the tight bound.
    sample_tol = GRAD_TOL[precision] * (20.0 if name.startswith("default") else 1.0)
    for k, v in g.items():
        if analytically_zero(k, spec):
            continue
        gn = float(np.sqrt((v.astype(np.float64) ** 2).sum()))
        ref = float(gold["gnorm/" + k])
        assert abs(gn - ref) <= sample_tol * max(ref, 1e-3 * scale), (k, gn, ref)
        stride = max(1, v.size // 64)
        assert np.abs(v.reshape(-1)[::stride][:64] - gold["gsample/" + k]).max() <= sample_tol * max(
            np.abs(gold["gsample/" + k]).max(), 1e-3 * scale), k
    eng.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_config1_64cube_forward_dice_matches_oracle(gpu_lib, precision):
    """BASELINE config #1: single 64^3, 1-modality, 2-class patch, batch 1, forward + weighted Dice."""
    spec = R.VNetSpec(num_classes=2, in_channels=1)
    params = R.init_params(spec, 42)
    img, lab = synth_batch(0, 1, 64, 1, 2)
    eng = engine_for(spec, 64, 1, "weighted_sorensen", (0.1, 1.0), gpu_lib, precision=precision)
    eng.set_params(params)
    logits, softmax, argmax = eng.forward(img)
    p = R.to_torch(params)
    with torch.no_grad():
        lo, _ = R.forward(p, torch.from_numpy(img), spec)
        loss_o = R.loss_from_logits(lo, torch.from_numpy(lab), "weighted_sorensen", (0.1, 1.0))
    ref = lo.numpy()
    err = rel_err(logits, ref)
    assert err < LOGIT_TOL[precision], err
    ref_arg = R.predict(lo).numpy()
    nflip = assert_argmax_parity(argmax, logits, ref, ref_arg)
    # hard Dice from the label volumes (integer TP / FP): equal up to the voxels tied inside the rounding band
    for mine, theirs in ((argmax, ref_arg),):
        assert abs(int(((mine == 1) & (lab == 1)).sum()) - int(((theirs == 1) & (lab == 1)).sum())) <= nflip
        assert abs(int(((mine == 1) & (lab != 1)).sum()) - int(((theirs == 1) & (lab != 1)).sum())) <= nflip
    assert abs(eng.loss(img, lab) - float(loss_o)) < 5e-5
    eng.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_gradients_match_oracle_default_net(gpu_lib, precision):
    spec = R.VNetSpec(num_classes=2, in_channels=1)
    params = perturbed_params(spec)
    img, lab = synth_batch(5, 2, 32, 1, 2)
    eng = engine_for(spec, 32, 2, "weighted_sorensen", (0.1, 1.0), gpu_lib, precision=precision)
    eng.set_params(params)
    l = eng.forward_backward(img, lab, update_moving_stats=True)
    lo, _, go, upd = R.loss_and_grads(params, img, lab, spec, "weighted_sorensen", (0.1, 1.0))
    assert abs(l - float(lo)) < 5e-5
    # 32^3 through 4 levels leaves 2^3 voxels x 2 samples per channel at the bottom: batch norm over 16
    # values amplifies rounding noise and single PReLU sign flips dominate the max norm, so this net is
    # checked in the per-tensor relative L2 norm; the tiny nets keep the tight element-wise bound
    _check_grads(eng, {k: v.numpy() for k, v in go.items()}, spec, max(GRAD_TOL[precision], 3e-2), l2=True)
    for k, u in upd.items():
        assert np.abs(eng.get_param(k) - u.numpy()).max() < 1e-3 * max(1.0, float(u.abs().max())), k
    eng.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
def test_multimodal_four_class_net_matches_oracle(gpu_lib, precision):
    """BASELINE config #3 shape (4 modalities, 4 classes) at a size the oracle finishes in seconds.  In the
    tensor-core precisions the 4-channel input convolution runs on zero-padded 16-channel bf16 copies."""
    spec = R.VNetSpec(num_classes=4, in_channels=4)
    params = perturbed_params(spec)
    img, lab = synth_batch(2, 1, 32, 4, 4)
    eng = engine_for(spec, 32, 1, "weighted_sorensen", (0.01, 0.1, 0.5, 1.0), gpu_lib, precision=precision)
    eng.set_params(params)
    l = eng.forward_backward(img, lab)
    lo, lg, go, _ = R.loss_and_grads(params, img, lab, spec, "weighted_sorensen", (0.01, 0.1, 0.5, 1.0))
    assert abs(l - float(lo)) < (5e-3 if precision == "bf16" else 5e-5)
    logits, _, _ = eng.forward(img)
    assert rel_err(logits, lg.numpy()) < LOGIT_TOL[precision]
    if precision != "bf16":  # single-pass bf16 on this ill-conditioned 32^3 net: gradients only checked to be finite
        _check_grads(eng, {k: v.numpy() for k, v in go.items()}, spec, max(GRAD_TOL[precision], 3e-2), l2=True)
    else:
        assert all(np.isfinite(v).all() for v in eng.get_grads().values())
    eng.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_legacy_vnet_py_flavour(gpu_lib, precision):
    """the VNet.py graph that train.py builds (3 levels, (1,2,2), bottom 3, prelu)."""
    spec = R.VNetSpec(num_classes=2, in_channels=1, num_levels=3, num_convolutions=(1, 2, 2), bottom_convolutions=3,
                      flavour="legacy")
    params = perturbed_params(spec)
    img, lab = synth_batch(1, 2, 32, 1, 2)
    eng = engine_for(spec, 32, 2, "jaccard", (), gpu_lib, precision=precision)
    assert list(eng.variables()) == [n for n, _, _ in R.param_specs(spec)]
    eng.set_params(params)
    l = eng.forward_backward(img, lab, update_moving_stats=True)
    lo, lg, go, upd = R.loss_and_grads(params, img, lab, spec, "jaccard", ())
    logits, _, am = eng.forward(img)
    assert abs(l - float(lo)) < 5e-5
    assert rel_err(logits, lg.numpy()) < LOGIT_TOL[precision]
    assert_argmax_parity(am, logits, lg.numpy())
    _check_grads(eng, {k: v.numpy() for k, v in go.items()}, spec, max(GRAD_TOL[precision], 3e-2), l2=True)
    eng.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_three_training_steps_follow_the_oracle(gpu_lib, precision):
    spec = R.VNetSpec(num_classes=2, in_channels=1, num_channels=16, num_levels=2, num_convolutions=(1, 2), bottom_convolutions=2)
    params = perturbed_params(spec)
    state = R.TrainState(params={k: v.copy() for k, v in params.items()})
    eng = engine_for(spec, 16, 2, "weighted_sorensen", (0.1, 1.0), gpu_lib, precision=precision, learning_rate=1e-3)
    eng.set_params(params)
    for step in range(3):
        img, lab = synth_batch(step, 2, 16, 1, 2)
        lo, _, _ = R.train_step(state, img, lab, spec, "weighted_sorensen", (0.1, 1.0), lr0=1e-3)
        le = eng.train_step(img, lab)
        assert abs(le - lo) < 2e-3, (step, le, lo)
    assert eng.global_step == 3
    eng.close()


@pytest.mark.parametrize("cin,cout,dims", [(16, 16, (8, 8, 16)), (32, 16, (4, 8, 32)), (64, 32, (8, 8, 8)),
                                           (4, 16, (6, 5, 9)), (128, 128, (4, 4, 4)),
                                           (16, 16, (4, 6, 24)), (32, 32, (2, 4, 48)), (64, 64, (2, 2, 96)),
                                           (16, 32, (2, 3, 192)), (32, 16, (3, 4, 160)), (128, 128, (2, 12, 12)),
                                           (128, 128, (4, 8, 8)), (256, 128, (3, 8, 16)), (128, 256, (8, 8, 8))])
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_conv5_ops_match_torch(gpu_lib, cin, cout, dims, precision):
    """Per-op hooks: 5^3 SAME convolution forward, input gradient and filter gradient."""
    import ctypes as C
    if precision != "fp32" and (cin % 16 or cout % 16 or dims[2] < 8):
        pytest.skip("shape outside the tensor-core kernels' domain (channels % 16, W >= 8): the engine runs "
                    "such layers on the fp32 kernels")
    rng = np.random.default_rng(7)
    n = 2
    x = rng.normal(0, 1, (n,) + dims + (cin,)).astype(np.float32)
    w = rng.normal(0, 0.05, (5, 5, 5, cin, cout)).astype(np.float32)
    b = rng.normal(0, 1, (cout,)).astype(np.float32)
    r = rng.normal(0, 1, (n,) + dims + (cout,)).astype(np.float32)
    xt = torch.from_numpy(x).requires_grad_(True)
    wt = torch.from_numpy(w).requires_grad_(True)
    y_ref = R.conv_same(xt, wt, torch.from_numpy(b)) + torch.from_numpy(r)
    dy = rng.normal(0, 1, y_ref.shape).astype(np.float32)
    y_ref.backward(torch.from_numpy(dy))
    prec = {"fp32": 0, "bf16x3": 1}[precision]
    tol = 1e-5 if precision == "fp32" else 3e-5
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    y = np.empty_like(dy)
    gpu_lib.check(gpu_lib.vnb_op_conv5_fprop(0, prec, ptr(x), ptr(w), ptr(b), ptr(r), ptr(y), n, *dims, cin, cout))
    assert rel_err(y, y_ref.detach().numpy()) < tol
    dx = np.empty_like(x)
    gpu_lib.check(gpu_lib.vnb_op_conv5_dgrad(0, prec, ptr(dy), ptr(w), ptr(dx), n, *dims, cin, cout))
    assert rel_err(dx, xt.grad.numpy()) < tol
    dw = np.empty_like(w)
    gpu_lib.check(gpu_lib.vnb_op_conv5_wgrad(0, prec, ptr(x), ptr(dy), ptr(dw), n, *dims, cin, cout))
    assert rel_err(dw, wt.grad.numpy()) < 5 * tol


@pytest.mark.parametrize("cin,cout,dims", [(64, 64, (8, 8, 32)), (16, 64, (4, 6, 16)), (64, 32, (3, 8, 128)), (2, 64, (5, 6, 7)),
                                           (64, 64, (2, 3, 192)), (64, 64, (3, 5, 24))])
@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
def test_conv3_ops_match_torch(gpu_lib, cin, cout, dims, precision):
    """3^3 SAME convolution of the attention / output modules (attention.py:63-92): forward, input and filter gradient."""
    import ctypes as C
    import torch.nn.functional as F
    if precision != "fp32" and (cin % 16 or cout % 16 or dims[2] < 8):
        pytest.skip("shape outside the tensor-core kernels' domain: the engine runs such layers on the fp32 kernels")
    rng = np.random.default_rng(9)
    n = 2
    x = rng.normal(0, 1, (n,) + dims + (cin,)).astype(np.float32)
    w = rng.normal(0, 0.1, (3, 3, 3, cin, cout)).astype(np.float32)
    b = rng.normal(0, 1, (cout,)).astype(np.float32)
    dy = rng.normal(0, 1, (n,) + dims + (cout,)).astype(np.float32)
    rnd = (lambda a: torch.from_numpy(a).to(torch.bfloat16).to(torch.float64)) if precision == "bf16" else (lambda a: torch.from_numpy(a).double())
    xt, wt = rnd(x).requires_grad_(True), rnd(w).requires_grad_(True)
    y_ref = F.conv3d(xt.permute(0, 4, 1, 2, 3), wt.permute(4, 3, 0, 1, 2), padding=1).permute(0, 2, 3, 4, 1) + torch.from_numpy(b).double()
    y_ref.backward(rnd(dy))
    prec = {"fp32": 0, "bf16x3": 1, "bf16": 2}[precision]
    tol = {"fp32": 1e-5, "bf16x3": 3e-5, "bf16": 2e-6}[precision]   # bf16: exact products of the rounded operands
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    y = np.empty_like(dy)
    gpu_lib.check(gpu_lib.vnb_op_conv3_fprop(0, prec, ptr(x), ptr(w), ptr(b), None, ptr(y), n, *dims, cin, cout))
    assert rel_err(y, y_ref.detach().numpy()) < tol
    dx = np.empty_like(x)
    gpu_lib.check(gpu_lib.vnb_op_conv3_dgrad(0, prec, ptr(dy), ptr(w), ptr(dx), n, *dims, cin, cout))
    assert rel_err(dx, xt.grad.numpy()) < tol
    dw = np.empty_like(w)
    gpu_lib.check(gpu_lib.vnb_op_conv3_wgrad(0, prec, ptr(x), ptr(dy), ptr(dw), n, *dims, cin, cout))
    assert rel_err(dw, wt.grad.numpy()) < 5 * tol


def test_full_size_properties_128cube(gpu_lib):
    """BASELINE config #2 size (128^3, batch 2): size-independent properties instead of the oracle --
    determinism of the forward, bias invariance of the logits, Dice of a perfect prediction,
    and a loss that decreases over optimiser steps."""
    spec = R.VNetSpec(num_classes=2, in_channels=1)
    params = R.init_params(spec, 42)
    img, lab = synth_batch(0, 2, 128, 1, 2)
    eng = engine_for(spec, 128, 2, "weighted_sorensen", (0.1, 1.0), gpu_lib, precision="bf16x3", learning_rate=1e-3)
    eng.set_params(params)
    a, _, am = eng.forward(img, want_softmax=False)
    b, _, _ = eng.forward(img, want_softmax=False, want_argmax=False)
    assert np.array_equal(a, b)
    rng = np.random.default_rng(0)
    for k, (shape, _) in eng.variables().items():
        if k.endswith("/biases"):
            eng.set_param(k, rng.normal(0, 2, shape).astype(np.float32))
    c, _, _ = eng.forward(img, want_softmax=False, want_argmax=False)
    assert rel_err(c, a) < 1e-3
    l0, terms = eng.loss(img, lab, want_terms=True)
    assert np.allclose(terms[..., 2].sum(1), 128 ** 3)  # one-hot mass = voxel count per sample
    assert np.allclose(terms[..., 1].sum(1), 128 ** 3, rtol=1e-4)  # softmax mass
    losses = [eng.train_step(img, lab) for _ in range(4)]
    assert losses[-1] < losses[0]
    eng.close()


def _attention_case(spec, P, N, nch, weight_scale):
    from tests.helpers import perturbed_attention_params
    from vnet_tensorflow_b200.synthetic import synth_patch
    params = perturbed_attention_params(spec, nch, weight_scale=weight_scale)
    samples = [synth_patch(1234 + 100000 * i, P, spec.in_channels, spec.num_classes) for i in range(N)]
    img, lab, dm = (np.stack([s[j] for s in samples], 0) for j in range(3))
    return params, img, lab, dm


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("flavour,K,M,loss,att_loss", [("legacy", 2, 1, "jaccard", "l2"),
                                                       ("legacy", 2, 1, "sorensen_fg", "abs"),
                                                       ("networks", 3, 2, "weighted_sorensen", "l2")])
def test_attention_gating_path(gpu_lib, flavour, K, M, loss, att_loss, precision):
    """SURVEY §8 row a15 (BASELINE config #5): V-Net -> AttentionModule (64 ch) -> (1 + softmax) * logits ->
    OutputModule (64 ch), Dice + attention loss, against the oracle restatement of train.py:281-312,351-418."""
    convs = (1, 2, 2) if flavour == "legacy" else (1, 2)
    spec = R.VNetSpec(num_classes=K, in_channels=M, num_channels=16, num_levels=len(convs), num_convolutions=convs,
                      bottom_convolutions=2, flavour=flavour)
    P, N, nch = 32, 2, 64
    weights = (0.01, 0.1, 1.0) if "weighted" in loss else ()
    params, img, lab, dm = _attention_case(spec, P, N, nch, weight_scale=0.25)
    eng = engine_for(spec, P, N, loss, weights, gpu_lib, precision=precision, attention=True, attention_loss=att_loss)
    assert list(eng.variables()) == [n for n, _, _ in R.attention_param_specs(spec, nch)]
    eng.set_params(params)
    tot, seg, att, out, go, upd = R.attention_loss_and_grads(params, img, lab, dm, spec, loss, att_loss, weights=weights)
    logits, _, am = eng.forward(img)
    tol = LOGIT_TOL[precision]
    assert rel_err(logits, out["logits_output"].numpy()) < tol
    assert rel_err(eng.softmax_attention(N), out["softmax_attention"].numpy()) < tol
    if precision == "bf16":
        assert (am != R.predict(out["logits_output"]).numpy()).mean() < 0.02
    else:
        assert_argmax_parity(am, logits, out["logits_output"].numpy())
    eng.set_distmap(dm)
    l = eng.forward_backward(img, lab, update_moving_stats=True)
    t3 = eng.losses()
    ltol = 5e-2 if precision == "bf16" else 2e-4
    assert abs(l - float(tot)) < ltol * max(1.0, abs(float(tot)))
    assert abs(t3[1] - float(seg)) < ltol and abs(t3[2] - float(att)) < ltol * max(1.0, abs(float(att)))
    if precision == "bf16":
        assert all(np.isfinite(v).all() for v in eng.get_grads().values())
    else:
        _check_grads(eng, {k: v.numpy() for k, v in go.items()}, spec, max(GRAD_TOL[precision], 3e-2), l2=True)
    eng.close()


def test_attention_path_trains(gpu_lib):
    """Three Adam steps of the gated network follow the oracle's trajectory (losses, fp32 path)."""
    spec = R.VNetSpec(num_classes=2, in_channels=1, num_channels=16, num_levels=2, num_convolutions=(1, 2),
                      bottom_convolutions=1, flavour="legacy")
    P, N, nch = 16, 2, 64
    params, img, lab, dm = _attention_case(spec, P, N, nch, weight_scale=0.25)
    eng = engine_for(spec, P, N, "jaccard", (), gpu_lib, precision="fp32", attention=True, attention_loss="l2",
                     learning_rate=1e-3)
    eng.set_params(params)
    eng.set_distmap(dm)
    p = {k: v.copy() for k, v in params.items()}
    m = {k: torch.zeros(v.shape) for k, v in p.items()}
    v2 = {k: torch.zeros(v.shape) for k, v in p.items()}
    for step in range(3):
        tot, _, _, _, go, upd = R.attention_loss_and_grads(p, img, lab, dm, spec, "jaccard", "l2")
        le = eng.train_step(img, lab)
        assert abs(le - float(tot)) < 2e-3 * max(1.0, abs(float(tot))), (step, le, float(tot))
        lr = R.learning_rate(1e-3, step, 100.0, 0.99)
        for k, g in go.items():
            pk, m[k], v2[k] = R.adam_update(torch.from_numpy(p[k]), g, m[k], v2[k], step + 1, lr)
            p[k] = pk.numpy()
        for k, u in upd.items():
            p[k] = u.numpy()
    assert eng.global_step == 3
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_sliding_window_evaluation_matches_the_reference_loop(gpu_lib, precision):
    """vnb_evaluate_volume (window gather, accumulation and argmax on the device) against the oracle restatement of
    model.py:866-937 fed by vnb_forward: identical windows / batches / order of additions -> bit-exact."""
    from oracle import ref_eval
    from vnet_tensorflow_b200.engine import VNetEngine
    spec = R.VNetSpec(num_classes=3, in_channels=2, num_channels=16, num_levels=2, num_convolutions=(1, 2), bottom_convolutions=1)
    P, B = (16, 16, 16), 3
    eng = VNetEngine(num_classes=3, in_channels=2, patch_shape=P, max_batch=B, num_channels=16, num_levels=2,
                     num_convolutions=(1, 2), bottom_convolutions=1, precision=precision, library=gpu_lib)
    eng.set_params(R.init_params(spec, 7))
    rng = np.random.default_rng(3)
    vol = rng.uniform(0, 255, (37, 16, 29, 2)).astype(np.float32)   # ragged: clamped last windows, a unit axis
    stride = (9, 8, 16)
    lab, sums, wgt = eng.evaluate_volume(vol, stride, B)
    lab_o, sums_o, wgt_o = ref_eval.evaluate_volume(vol, P, stride, B, 3, lambda x: eng.forward(x, want_logits=False, want_argmax=False)[1])
    assert np.array_equal(wgt, wgt_o) and wgt.min() >= 1
    assert np.array_equal(sums, sums_o)
    assert np.array_equal(lab, lab_o) and lab.dtype == np.int64
    np.testing.assert_allclose(sums.sum(-1), wgt, rtol=1e-5)   # softmax rows sum to one per covering window
    with pytest.raises(Exception):
        eng.evaluate_volume(vol[:8], stride, B)                 # smaller than the patch: the caller must pad
    eng.close()


def test_every_unit_against_the_oracle_taps(gpu_lib):
    """vnb_read_tensor around vnb_forward_backward on the exact-fp32 path: every unit's batch-norm input, output
    activation and dL/d(BN input) against the oracle's taps (the per-op view of the fused BN / PReLU / loss passes)."""
    spec = R.VNetSpec(**CASES["tiny_m1_k2"][0])   # 16 channels, two levels: the shapes of the golden-fixture case
    P, N = 16, 2
    params = perturbed_params(spec, 3)
    img, lab = synth_batch(1, N, P, 1, 2)
    eng = engine_for(spec, P, N, "weighted_sorensen", (0.1, 1.0), gpu_lib)
    eng.set_params(params)
    loss = eng.forward_backward(img, lab, dropout_rate=0.0)
    col = {}
    tp = R.to_torch(params, torch.float64, requires_grad=True)
    logits, _ = R.forward(tp, torch.from_numpy(img).double(), spec, collect=col)
    for t in col.values():
        if t.requires_grad:
            t.retain_grad()
    ref_loss = R.loss_from_logits(logits, torch.from_numpy(lab), "weighted_sorensen", (0.1, 1.0))
    ref_loss.backward()
    assert abs(loss - float(ref_loss.detach())) < 1e-5
    worst = {"z": 0.0, "a": 0.0, "dz": 0.0}
    for name, t in col.items():
        scope, _, kind = name.partition(":")
        shape = tuple(t.shape)
        if kind == "bn_in":
            if scope == "vnet/input_layer":
                continue
            worst["z"] = max(worst["z"], rel_err(eng.read_tensor(scope, 2, N, shape[-1], shape[1:4]), t.detach().numpy()))
            dz, ref = eng.read_tensor(scope, 1, N, shape[-1], shape[1:4]).astype(np.float64), t.grad.numpy()
            # relative L2: a PReLU input within fp32 rounding of zero may take the other branch at a single voxel
            worst["dz"] = max(worst["dz"], float(np.sqrt(((dz - ref) ** 2).sum()) / max(np.sqrt((ref ** 2).sum()), 1e-30)))
        else:
            worst["a"] = max(worst["a"], rel_err(eng.read_tensor(scope, 0, N, shape[-1], shape[1:4]), t.detach().numpy()))
    print("per-unit worst relative errors:", worst)
    assert worst["z"] < LOGIT_TOL["fp32"] and worst["a"] < LOGIT_TOL["fp32"] and worst["dz"] < GRAD_TOL["fp32"]
    eng.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_short_batch_after_a_full_one(gpu_lib, precision):
    """A handle created for max_batch = 2 serves a full batch and then one patch (the last batch of an epoch is short):
    statistics, loss and gradients of the second call belong to its single patch.  CPU twin of the tensor-core form:
    tests/test_tc_emul.py::test_engine_bf16x3_short_batch_after_a_full_one.  (Added after the round-1 GPU budget was
    spent: first run on hardware is the round-end GPU suite.)"""
    spec = R.VNetSpec(**CASES["tiny_m1_k2"][0])
    P = 16
    params = perturbed_params(spec)
    eng = engine_for(spec, P, 2, "weighted_sorensen", (0.1, 1.0), gpu_lib, precision=precision)
    eng.set_params(params)
    img2, lab2 = synth_batch(0, 2, P, 1, 2)
    l2 = eng.forward_backward(img2, lab2)
    lo2 = R.loss_and_grads(params, img2, lab2, spec, "weighted_sorensen", (0.1, 1.0))[0]
    assert abs(l2 - float(lo2)) < 5e-5
    img, lab = synth_batch(5, 1, P, 1, 2)
    l = eng.forward_backward(img, lab)
    lo, lg, go, _ = R.loss_and_grads(params, img, lab, spec, "weighted_sorensen", (0.1, 1.0))
    assert abs(l - float(lo)) < 5e-5
    _check_grads(eng, {k: v.numpy() for k, v in go.items()}, spec, GRAD_TOL[precision], l2=True)
    logits, _, argmax = eng.forward(img)
    assert logits.shape[0] == 1 and rel_err(logits, lg.numpy()) < LOGIT_TOL[precision]
    assert_argmax_parity(argmax, logits, lg.numpy())
    eng.close()


def test_device_resident_batches_give_identical_results(gpu_lib):
    """SURVEY 8(b) '*_dev' variants: images / labels (and forward outputs) in caller-owned CUDA memory take the same
    kernels as host buffers - logits, argmax, loss and the trained weights are bit-identical; wrong dtype / device is
    refused before anything reaches the C ABI.  (Added after the round-1 GPU budget was spent: first run on hardware is
    the round-end GPU suite.)"""
    import ctypes as C
    spec = R.VNetSpec(**CASES["tiny_m1_k2"][0])
    P, N = 16, 2
    img, lab = synth_batch(3, N, P, 1, 2)
    img_d, lab_d = torch.from_numpy(img).cuda(), torch.from_numpy(lab).cuda()
    results = []
    for images, labels in ((img, lab), (img_d, lab_d)):
        eng = engine_for(spec, P, N, "weighted_sorensen", (0.1, 1.0), gpu_lib, precision="bf16x3")
        eng.set_params(perturbed_params(spec))
        logits, _, argmax = eng.forward(images)
        loss = eng.loss(images, labels)
        step_loss = eng.train_step(images, labels, dropout_rate=0.0)
        results.append((logits, argmax, loss, step_loss, eng.get_param("vnet/encoder/level_1/conv_1/weights")))
        if images is img_d:   # outputs straight into CUDA memory through the C ABI (after the step: new weights)
            ref_logits, _, ref_argmax = eng.forward(img)
            out = torch.empty((N, P, P, P, 2), dtype=torch.float32, device="cuda")
            am = torch.empty((N, P, P, P), dtype=torch.int64, device="cuda")
            gpu_lib.check(gpu_lib.vnb_forward(eng._h, C.c_void_p(img_d.data_ptr()), N, C.c_void_p(out.data_ptr()), None,
                                              C.c_void_p(am.data_ptr())))
            assert np.array_equal(out.cpu().numpy(), ref_logits) and np.array_equal(am.cpu().numpy(), ref_argmax)
            with pytest.raises(ValueError):
                eng.forward(img_d.double())
            with pytest.raises(ValueError):
                eng.loss(img_d, lab_d.long())
        eng.close()
    host, dev = results
    assert np.array_equal(host[0], dev[0]) and np.array_equal(host[1], dev[1])
    assert host[2] == dev[2] and host[3] == dev[3]
    # the 2x2x2 / 1x1x1 filter gradients combine their splits with fp32 atomics: last-bit freedom only
    assert np.abs(host[4] - dev[4]).max() <= 1e-6 * np.abs(host[4]).max()


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_step_metrics_match_the_reference_metric_block(gpu_lib, precision):
    """vnb_read_metrics on the device (model.py:586-626): confusion counts bit-exact against the engine's argmax, AUC
    histograms bit-exact against the engine's softmax at the 200 tf.metrics.auc thresholds, derived scalars equal to the
    literal restatement of the reference block.  CPU twin: tests/test_engine_emul.py.  (Added after the round-1 GPU
    budget was spent: first run on hardware is the round-end GPU suite.)"""
    from oracle import ref_metrics as RM
    from vnet_tensorflow_b200 import metrics as M
    spec = R.VNetSpec(num_classes=3, in_channels=2, num_channels=16, num_levels=2, num_convolutions=(1, 2), bottom_convolutions=1)
    P, N, K = 32, 2, 3
    eng = engine_for(spec, P, N, "weighted_sorensen", (0.1, 0.5, 1.0), gpu_lib, precision=precision)
    eng.set_params(perturbed_params(spec))
    img, lab = synth_batch(2, N, P, 2, K)
    lab = lab.copy()
    lab[0, 0, 0, :3] = (K, -1, K + 5)           # labels outside [0, K): all-zero one-hot rows
    eng.forward_backward(img, lab)
    cm, hist = eng.metric_counts(N)
    logits, softmax, argmax = eng.forward(img)
    want_cm = np.zeros_like(cm)
    np.add.at(want_cm, (np.where((lab >= 0) & (lab < K), lab, K).reshape(-1), argmax.reshape(-1)), 1)
    assert np.array_equal(cm, want_cm) and int(cm.sum()) == lab.size
    thr = M.auc_thresholds()
    want_hist = np.zeros_like(hist)
    for c in range(1, K):
        bins = np.searchsorted(thr, softmax[..., c].reshape(-1), side="left")      # number of thresholds below p
        np.add.at(want_hist[c], ((lab.reshape(-1) == c).astype(int), bins), 1)
    assert np.array_equal(hist, want_hist)
    got = M.step_metrics(cm, hist, [0, 1, 2])
    ref = RM.step_metrics(logits, lab, softmax, [0, 1, 2])
    for k in got:
        same = got[k] == ref[k] or (np.isnan(got[k]) and np.isnan(ref[k]))
        assert same or (k.startswith("auc_") and abs(got[k] - ref[k]) < 1e-6), (k, got[k], ref[k])
    # and against the oracle's own forward pass: same hard counts wherever the argmax agrees (it does, bit-exactly)
    lg_o = R.forward(R.to_torch(perturbed_params(spec)), torch.from_numpy(img), spec)[0]
    assert_argmax_parity(argmax, logits, lg_o.numpy())
    eng.close()


def test_staged_input_steps_equal_direct_steps(gpu_lib):
    """vnb_stage_batch / vnb_train_step_staged on the device: copy stream, events and the device-to-device commit give
    the losses and weights of vnb_train_step bit for bit, with the next batch staged while a step runs.  CPU twin:
    tests/test_engine_emul.py.  (Added after the round-1 GPU budget was spent: first run on hardware is the round-end
    GPU suite.)"""
    spec = R.VNetSpec(**CASES["tiny_m1_k2"][0])
    P = 16
    params = perturbed_params(spec)
    batches = [synth_batch(s, n, P, 1, 2) for s, n in ((0, 2), (1, 2), (2, 1), (3, 2))]
    direct = engine_for(spec, P, 2, "weighted_sorensen", (0.1, 1.0), gpu_lib, precision="bf16x3")
    direct.set_params(params)
    want = [direct.train_step(img, lab, 0.0, seed=i) for i, (img, lab) in enumerate(batches)]
    eng = engine_for(spec, P, 2, "weighted_sorensen", (0.1, 1.0), gpu_lib, precision="bf16x3")
    eng.set_params(params)
    pins = [(eng.pinned_array((2, P, P, P, 1), np.float32), eng.pinned_array((2, P, P, P), np.int32)) for _ in range(2)]

    def stage(i):
        img, lab = batches[i]
        n = img.shape[0]
        pi, pl = pins[i % 2]
        pi[:n], pl[:n] = img, lab
        eng.stage_batch(pi[:n], pl[:n])

    got = []
    stage(0)
    for i in range(len(batches)):
        eng.train_step_staged(0.0, seed=i, want_loss=False)
        if i + 1 < len(batches):
            stage(i + 1)
        got.append(eng.last_loss())
    # the first step is bit-identical; from the second on the 2x2x2 / 1x1x1 filter gradients (fp32 atomics over their
    # voxel splits) leave the two engines last-bit freedom, so the trajectories are compared to rounding
    assert got[0] == want[0]
    assert max(abs(a - b) for a, b in zip(got, want)) < 1e-4
    k = "vnet/encoder/level_1/conv_1/weights"
    a, b = eng.get_param(k).astype(np.float64), direct.get_param(k).astype(np.float64)
    assert np.sqrt(((a - b) ** 2).sum()) <= 1e-3 * np.sqrt((b ** 2).sum())
    assert eng.global_step == direct.global_step == len(batches)
    eng.close()
    direct.close()


# ---------------------------------------------------------------------------------------------------------------------
# Parity at the BASELINE sizes: the oracle runs a 128^3 x 2 optimiser step in seconds on the GPU
# box's host cores, so the benchmarked configuration itself is compared, not a scaled-down stand-in.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
def test_benchmarked_config2_128cube_batch2_matches_oracle(gpu_lib, precision):
    """BASELINE configs[1] (the benchmarked workload): 128^3, 1 modality, 2 classes, batch 2, seed-42 weights.
    north_star bars: logits within 1e-3 relative, argmax label volume and hard-Dice counts bit-exact, plus |dloss| and
    the step-1 gradients per tensor in relative L2 (the same function bench.py prints as its `parity` block)."""
    import bench
    spec = R.VNetSpec(num_classes=2, in_channels=1)
    params = R.init_params(spec, 42)
    img, lab = synth_batch(0, 2, 128, 1, 2)
    lo, lg, go, _ = R.loss_and_grads(params, img, lab, spec, "weighted_sorensen", (0.1, 1.0))
    eng = engine_for(spec, 128, 2, "weighted_sorensen", (0.1, 1.0), gpu_lib, precision=precision)
    rep = bench.engine_parity(eng, params, img, lab, float(lo), lg.numpy(), {k: v.numpy() for k, v in go.items()}, precision)
    print("parity[%s]:" % precision, {k: v for k, v in rep.items() if k not in ("hard_dice", "against")})
    assert rep["logits_max_rel_err"] <= (2e-4 if precision == "fp32" else 1e-3)
    assert rep["argmax_mismatches_outside_rounding_band"] == 0
    assert rep["argmax_mismatches"] <= 1e-4 * rep["voxels"]
    assert rep["hard_dice_max_count_diff"] <= rep["argmax_mismatches"]
    if precision == "bf16x3":
        # the reference's own ambiguity: the oracle evaluated in fp64 against itself in fp32 on the same tensors
        with torch.no_grad():
            lg64, _ = R.forward(R.to_torch(params, torch.float64), torch.from_numpy(img).double(), spec)
        n64, out64, band64 = argmax_parity(np.argmax(lg.numpy(), -1), lg.numpy(), lg64.numpy())
        _report("oracle fp32 vs oracle fp64 on the same tensors: %d argmax mismatches (%d outside its rounding band %.3g); "
                "engine[bf16x3] vs oracle fp32: %d" % (n64, out64, band64, rep["argmax_mismatches"]))
    assert rep["abs_loss_diff"] < 5e-5
    # gradients: per-tensor relative L2 at the conditioning floor of this network (a 1e-5 perturbation of one
    # layer's weights moves the exact gradients by ~3e-2 in the max norm); the worst tensors are batch-norm betas of the deep
    # levels, the first convolution's filter gradient is held tightly
    assert rep["grad_rel_l2_worst"] < (2e-2 if precision == "fp32" else 3e-2), rep["grad_rel_l2_worst_tensor"]
    assert rep["grad_rel_l2_first_conv"] < (5e-4 if precision == "fp32" else 2e-3)
    _report("config2 128^3 x2 [%s]: %s" % (precision, {k: v for k, v in rep.items() if k not in ("hard_dice", "against")}))
    eng.close()


def test_config3_128cube_four_modalities_four_classes_matches_oracle(gpu_lib):
    """BASELINE configs[2]: 128^3, 4 modalities, 4 classes (BraTS-shaped), batch 2.  The configuration is quoted in
    bf16; the parity anchor at this size is the fp32-grade bf16x3 mode (bit-exact argmax), and the single-pass bf16
    numbers are measured and recorded, with the bound the mode is documented to meet."""
    spec = R.VNetSpec(num_classes=4, in_channels=4)
    weights = (0.01, 0.1, 0.5, 1.0)
    params = R.init_params(spec, 42)
    img, lab = synth_batch(0, 2, 128, 4, 4)
    p = R.to_torch(params)
    with torch.no_grad():
        lg, _ = R.forward(p, torch.from_numpy(img), spec)
        loss_o = float(R.loss_from_logits(lg, torch.from_numpy(lab), "weighted_sorensen", weights))
    ref = lg.numpy()
    ref_arg = R.predict(lg).numpy()
    for precision in ("bf16x3", "bf16"):
        eng = engine_for(spec, 128, 2, "weighted_sorensen", weights, gpu_lib, precision=precision)
        eng.set_params(params)
        logits, _, argmax = eng.forward(img, want_softmax=False)
        err = rel_err(logits, ref)
        flips = int((argmax != ref_arg).sum())
        dl = abs(eng.loss(img, lab) - loss_o)
        _report("config3 128^3 [%s]: logits rel err %.3e, argmax flips %d of %d, |dloss| %.3e" % (precision, err, flips, ref_arg.size, dl))
        if precision == "bf16x3":
            assert err <= 1e-3 and dl < 5e-5
            # four classes on seed-42 weights: three near-equal runner-up logits per voxel, ~1e-4 of the voxels are tied
            # inside the rounding band (measured 440 of 4.2 M); none may lie outside it
            assert_argmax_parity(argmax, logits, ref, ref_arg, max_frac=5e-4)
        else:
            # single-pass bf16 is the documented reduced-precision mode, not a parity mode: measured 7.6e-2 relative on the
            # logits and 4.1 % label flips on these seed-42 weights, whose four logits are near-tied at most voxels
            assert err <= 0.1 and flips <= 0.1 * ref_arg.size and dl < 5e-3
        eng.close()


def test_config5_192cube_attention_forward_matches_oracle(gpu_lib):
    """BASELINE configs[4]: one 192^3, 2-modality, 3-class patch through V-Net -> AttentionModule -> gating -> OutputModule
    (forward only at this size: the oracle's 64-channel 3^3 modules at 192^3 take ~a minute on the host cores)."""
    from tests.helpers import perturbed_attention_params
    from vnet_tensorflow_b200.synthetic import synth_patch
    spec = R.VNetSpec(num_classes=3, in_channels=2)
    nch = 64
    params = perturbed_attention_params(spec, nch, weight_scale=0.25)
    im, lb, dm = synth_patch(1234, 192, 2, 3)
    img = im[None]
    with torch.no_grad():
        out = R.attention_forward(R.to_torch(params), torch.from_numpy(img), spec)
    out = out[0] if isinstance(out, tuple) else out
    ref = out["logits_output"].numpy()
    ref_arg = R.predict(out["logits_output"]).numpy()
    for precision in ("bf16x3", "bf16"):
        eng = engine_for(spec, 192, 1, "weighted_sorensen", (0.01, 0.1, 1.0), gpu_lib, precision=precision, attention=True,
                         attention_loss="l2")
        eng.set_params(params)
        logits, _, argmax = eng.forward(img, want_softmax=False)
        err = rel_err(logits, ref)
        flips = int((argmax != ref_arg).sum())
        _report("config5 192^3 attention [%s]: logits rel err %.3e, argmax flips %d of %d" % (precision, err, flips, ref_arg.size))
        if precision == "bf16x3":
            assert err <= 1e-3
            assert_argmax_parity(argmax, logits, ref, ref_arg, max_frac=5e-4)
            assert rel_err(eng.softmax_attention(1), out["softmax_attention"].numpy()) <= 1e-3
        else:
            assert err <= 0.1 and flips <= 0.1 * ref_arg.size
        eng.close()


# ---------------------------------------------------------------------------------------------------------------------
# GPU twins of what round 1 only ran on the CPU emulation: dropout, the remaining losses, the
# remaining optimisers, synchronised batch norm over NCCL.
# ---------------------------------------------------------------------------------------------------------------------
SPEC16 = dict(num_classes=2, in_channels=1, num_channels=16, num_levels=2, num_convolutions=(1, 2), bottom_convolutions=2)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_dropout_with_the_exported_mask_matches_oracle(gpu_lib, precision):
    """SURVEY 8 row a9 (networks.py:321,339,349,363): tf.nn.dropout inside every convolution unit.  The engine's
    counter-based keep-mask is exported through the activations (a == 0 exactly where dropped) and injected into the
    oracle, which must then give the same loss and gradients; same seed -> bit-identical step, other seed -> other mask."""
    spec = R.VNetSpec(**SPEC16)
    P, N, rate, seed = 16, 2, 0.3, 1234
    params = perturbed_params(spec)
    img, lab = synth_batch(0, N, P, 1, 2)
    eng = engine_for(spec, P, N, "sorensen", (), gpu_lib, precision=precision)
    eng.set_params(params)
    l1 = eng.forward_backward(img, lab, dropout_rate=rate, seed=seed)
    g1 = eng.get_grads()
    l2 = eng.forward_backward(img, lab, dropout_rate=rate, seed=seed)
    assert l1 == l2
    k5 = [k for k in g1 if k.endswith("/weights") and "convolution/" not in k and "output_layer" not in k]
    g2 = eng.get_grads()
    if precision == "fp32":   # the exact-fp32 reference kernels combine their voxel splits with atomics: equal to rounding
        assert all(np.abs(g1[k] - g2[k]).max() <= 1e-5 * np.abs(g1[k]).max() for k in k5)
    else:                     # tensor-core 5^3 filter gradients: fixed-order split-K reduction, bit-identical (the 4^3 bottom
        assert all(np.array_equal(g1[k], g2[k]) for k in k5 if "/level_" in k)   # level runs on the fp32 kernels: W < 8)
        assert all(np.abs(g1[k] - g2[k]).max() <= 1e-5 * np.abs(g1[k]).max() for k in k5)
    assert eng.forward_backward(img, lab, dropout_rate=rate, seed=seed + 1) != l1
    eng.forward_backward(img, lab, dropout_rate=rate, seed=seed)
    masks = {}
    d1, d2, d3 = (P,) * 3, (P // 2,) * 3, (P // 4,) * 3
    for name, c, sp in [("vnet/encoder/level_1/conv_1", 16, d1), ("vnet/encoder/level_2/conv_1", 32, d2),
                        ("vnet/encoder/level_2/conv_2", 32, d2), ("vnet/bottom_level/conv_1", 64, d3),
                        ("vnet/bottom_level/conv_2", 64, d3), ("vnet/decoder/level_2/conv_1", 32, d2),
                        ("vnet/decoder/level_2/conv_2", 32, d2), ("vnet/decoder/level_1/conv_1", 16, d1)]:
        a = eng.read_tensor(name, 0, N, c, sp)
        masks[name] = torch.from_numpy((a != 0).astype(np.float32))
        assert 0.6 < float(masks[name].mean()) < 0.8          # keep probability 0.7
    lo, _, go, _ = R.loss_and_grads(params, img, lab, spec, "sorensen", (), dropout_rate=rate, masks=masks)
    assert abs(float(lo) - l1) < 5e-5
    _check_grads(eng, {k: v.numpy() for k, v in go.items()}, spec, GRAD_TOL[precision], l2=True)
    eng.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("loss,weights,alpha", [("xent", (), 1.0), ("weighted_xent", (0.2, 1.0), 1.0), ("sorensen", (), 1.0),
                                                ("jaccard", (), 1.0), ("weighted_jaccard", (0.3, 1.0), 1.0),
                                                ("mixed_sorensen", (), 0.7), ("mixed_weighted_sorensen", (0.1, 1.0), 1.0),
                                                ("mixed_jaccard", (), 1.3), ("mixed_weighted_jaccard", (0.2, 1.0), 1.5)])
def test_loss_zoo_matches_oracle(gpu_lib, loss, weights, alpha, precision):
    """every Loss.Name of model.py:495-560 forward and backward on the device."""
    spec = R.VNetSpec(**SPEC16)
    P, N = 16, 2
    params = perturbed_params(spec)
    img, lab = synth_batch(3, N, P, 1, 2)
    eng = engine_for(spec, P, N, loss, weights, gpu_lib, precision=precision, loss_alpha=alpha)
    eng.set_params(params)
    lo, lg, go, _ = R.loss_and_grads(params, img, lab, spec, loss, weights, alpha)
    l = eng.forward_backward(img, lab)
    assert abs(l - float(lo)) < 5e-5 * max(1.0, abs(float(lo)))
    if loss.startswith("mixed"):
        d, x = eng.loss_parts()
        assert abs(d + x - l) < 1e-5 * max(1.0, abs(l))
    _check_grads(eng, {k: v.numpy() for k, v in go.items()}, spec, GRAD_TOL[precision], l2=True)
    eng.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("optimizer", ["SGD", "Momentum", "NesterovMomentum", "Adam"])
def test_optimizers_follow_the_oracle(gpu_lib, optimizer, precision):
    """model.py:649-658: three steps of every Optimizer.Name with a decaying learning rate; losses and weights."""
    spec = R.VNetSpec(**SPEC16)
    P, N = 16, 2
    params = perturbed_params(spec)
    state = R.TrainState(params={k: v.copy() for k, v in params.items()})
    eng = engine_for(spec, P, N, "weighted_sorensen", (0.1, 1.0), gpu_lib, precision=precision, optimizer=optimizer,
                     learning_rate=1e-3, decay_factor=0.5, decay_steps=2.0)
    eng.set_params(params)
    for step in range(3):
        img, lab = synth_batch(step, N, P, 1, 2)
        lo, _, _ = R.train_step(state, img, lab, spec, "weighted_sorensen", (0.1, 1.0), lr0=1e-3, decay_steps=2.0,
                                decay_factor=0.5, optimizer=optimizer)
        le = eng.train_step(img, lab)
        assert abs(le - lo) < 2e-3, (step, le, lo)
    k = "vnet/encoder/level_2/conv_1/weights"
    a, b = eng.get_param(k).astype(np.float64), state.params[k].astype(np.float64)
    if optimizer != "Adam":   # Adam's sign-like update amplifies last-bit gradient differences near zero; covered by the losses
        assert np.sqrt(((a - b) ** 2).sum()) <= 1e-3 * np.sqrt((b ** 2).sum())
    assert eng.global_step == 3
    eng.close()


def _run_torchrun(script_args, nproc, timeout=600):
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", "29533"] + script_args
    return subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_sync_bn_over_nccl_two_gpus_equal_one_device_at_the_global_batch(gpu_lib, precision):
    """2 ranks x 1 patch with vnb_comm_sync_bn == 1 device x 2 patches (loss, averaged gradients, moving
    statistics, logits).  Needs two GPUs (gpurun --gpus 2); skipped on a one-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _run_torchrun(["tools/sync_bn_check.py", "--precision", precision, "--patch", "32"], 2)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("SYNC_BN_CHECK")]
    assert line and "ok=True" in line[-1], r.stdout[-2000:]
    print(line[-1])


# ---- per-op hooks on the device ------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("cf,cc,dims", [(16, 32, (8, 8, 8)), (32, 64, (4, 6, 8)), (64, 128, (3, 4, 4)), (128, 256, (2, 2, 4)),
                                        (16, 32, (5, 3, 7))])
def test_k2_stride2_ops_match_torch(gpu_lib, cf, cc, dims, precision):
    """SURVEY 8 rows a2 / a3 per op: the 2^3 stride-2 down convolution, the transposed up convolution (= its input
    gradient) and their filter gradient, against torch (layers2.py:65-94)."""
    from tests import op_hook_cases as H
    H.check_k2_ops(gpu_lib, precision, cf, cc, dims)


def test_bn_softmax_dice_adam_ops_match_oracle(gpu_lib):
    from tests import op_hook_cases as H
    H.check_bn_ops(gpu_lib, 32768, 16)
    H.check_bn_ops(gpu_lib, 4096, 256, with_alpha=False)
    for loss, w, a, k in (("weighted_sorensen", (0.1, 1.0), 1.0, 2), ("jaccard", (), 1.0, 3), ("xent", (), 1.0, 2),
                          ("weighted_xent", (0.2, 0.3, 1.0), 1.0, 3), ("mixed_weighted_jaccard", (0.2, 0.3, 1.0), 1.5, 3),
                          ("mixed_sorensen", (), 0.7, 4)):
        H.check_softmax_dice_ops(gpu_lib, loss, w, a, voxels=32768, k=k)
    H.check_adam_op(gpu_lib, 1 << 20)
