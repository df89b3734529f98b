"""Kernel + engine logic on the CPU: the product's .cu sources are compiled with tests/emul/cuda_emul.h
(fiber-per-thread CUDA emulation) and the resulting engine is compared with the oracle.  These tests
exercise exactly the code that runs on the GPU in VNB_PREC_FP32 mode (index arithmetic, reductions,
BN chains, backward ordering, optimiser), without a GPU."""
import numpy as np
import pytest
import torch

from oracle import ref_vnet as R
from tests.helpers import analytically_zero, engine_for, perturbed_params, rel_err
from vnet_tensorflow_b200 import _ffi
from vnet_tensorflow_b200.synthetic import synth_batch

SPEC_A = R.VNetSpec(num_classes=2, in_channels=1, num_channels=4, num_levels=2, num_convolutions=(1, 2), bottom_convolutions=2)
SPEC_B = R.VNetSpec(num_classes=3, in_channels=2, num_channels=4, num_levels=2, num_convolutions=(3, 1), bottom_convolutions=1)
SPEC_C = R.VNetSpec(num_classes=4, in_channels=1, num_channels=8, num_levels=1, num_convolutions=(4,), bottom_convolutions=1)
# wide channels on a tiny volume: 64 / 128 channels put a whole warp (and more) of channel quads into the per-channel reductions
SPEC_W = R.VNetSpec(num_classes=2, in_channels=1, num_channels=64, num_levels=1, num_convolutions=(1,), bottom_convolutions=1)


def _grad_check(eng, grads_o, spec, tol):
    g = eng.get_grads()
    scale = max(float(np.abs(v.numpy()).max()) for v in grads_o.values())
    for k, v in g.items():
        ref = grads_o[k].numpy()
        if analytically_zero(k, spec):
            assert np.abs(v).max() <= 1e-6 * scale + 1e-12, k
            assert np.abs(ref).max() <= 1e-4 * scale, k  # the oracle only has rounding noise here
            continue
        assert np.abs(v - ref).max() <= tol * max(np.abs(ref).max(), 1e-3 * scale), k


@pytest.mark.parametrize("spec,N,loss,weights", [
    (SPEC_A, 2, "weighted_sorensen", (0.1, 1.0)),
    (SPEC_B, 1, "mixed_weighted_jaccard", (0.1, 0.5, 1.0)),
    (SPEC_C, 2, "jaccard", ()),
    (SPEC_A, 1, "xent", ()),
    (SPEC_B, 2, "weighted_xent", (0.2, 0.3, 1.0)),
    (SPEC_A, 2, "mixed_sorensen", ()),
    (SPEC_A, 2, "sorensen", ()),
    (SPEC_W, 2, "weighted_sorensen", (0.3, 1.0)),
])
def test_forward_loss_backward_match_oracle(emul_lib, spec, N, loss, weights):
    P = 8
    params = perturbed_params(spec)
    img, lab = synth_batch(0, N, P, spec.in_channels, spec.num_classes)
    eng = engine_for(spec, P, N, loss, weights, emul_lib)
    assert list(eng.variables()) == [n for n, _, _ in R.param_specs(spec)]
    eng.set_params(params)
    loss_o, logits_o, grads_o, upd = R.loss_and_grads(params, img, lab, spec, loss, weights)
    logits, softmax, argmax = eng.forward(img)
    assert rel_err(logits, logits_o.numpy()) < 2e-5
    assert rel_err(softmax, torch.softmax(logits_o, -1).numpy()) < 2e-5
    assert int((argmax != R.predict(logits_o).numpy()).sum()) == 0
    l, terms = eng.loss(img, lab, want_terms=True)
    assert abs(l - float(loss_o)) < 2e-6
    kind = "jaccard" if "jaccard" in loss else "sorensen"
    assert rel_err(terms[..., :3], R.dice_terms(logits_o, torch.from_numpy(lab), kind).numpy()) < 1e-5
    l2 = eng.forward_backward(img, lab, update_moving_stats=True)
    assert abs(l2 - float(loss_o)) < 2e-6
    if spec is SPEC_W:   # batch norm over 128 values per channel at the bottom amplifies fp32 rounding: per-tensor relative L2
        for k, v in eng.get_grads().items():
            ref = grads_o[k].numpy().astype(np.float64)
            if not analytically_zero(k, spec):
                assert np.sqrt(((v - ref) ** 2).sum()) <= 5e-3 * np.sqrt((ref ** 2).sum()), k
    else:
        _grad_check(eng, grads_o, spec, 2e-4)
    for k, u in upd.items():  # UPDATE_OPS (model.py:665-666)
        assert np.abs(eng.get_param(k) - u.numpy()).max() < 1e-4 * max(1.0, float(u.abs().max())), k
    eng.close()


def test_inference_does_not_touch_moving_statistics(emul_lib):
    eng = engine_for(SPEC_A, 8, 1, "sorensen", (), emul_lib)
    eng.set_params(perturbed_params(SPEC_A))
    img, lab = synth_batch(0, 1, 8, 1, 2)
    before = eng.get_param("vnet/output_layer/batch_normalization/moving_mean")
    eng.forward(img)
    eng.loss(img, lab)
    assert np.array_equal(before, eng.get_param("vnet/output_layer/batch_normalization/moving_mean"))
    eng.close()


@pytest.mark.parametrize("optimizer", ["Adam", "SGD", "Momentum", "NesterovMomentum"])
def test_training_trajectory_matches_oracle(emul_lib, optimizer):
    """Three optimiser steps (model.py:641-666): lr decay, TF-form Adam, moving stats, global_step."""
    spec, P, N = SPEC_A, 8, 2
    params = perturbed_params(spec)
    state = R.TrainState(params={k: v.copy() for k, v in params.items()})
    eng = engine_for(spec, P, N, "weighted_sorensen", (0.1, 1.0), emul_lib, optimizer=optimizer,
                     learning_rate=1e-2, decay_factor=0.5, decay_steps=2.0)
    eng.set_params(params)
    for step in range(3):
        img, lab = synth_batch(step, N, P, 1, 2)
        lo, _, _ = R.train_step(state, img, lab, spec, "weighted_sorensen", (0.1, 1.0), lr0=1e-2, decay_steps=2.0,
                                decay_factor=0.5, optimizer=optimizer)
        le = eng.train_step(img, lab)
        assert abs(le - lo) < 5e-5, (step, le, lo)
    assert eng.global_step == 3
    for k in eng.variables():
        if analytically_zero(k, spec):
            continue  # Adam turns rounding noise into +-lr walks on these
        a, b = eng.get_param(k), state.params[k]
        assert np.abs(a - b).max() <= 5e-3 * max(np.abs(b).max(), 1e-2), k
    eng.close()


def test_dropout_mask_is_reproducible_and_matches_oracle_with_injected_mask(emul_lib):
    spec, P, N, rate, seed = SPEC_A, 8, 1, 0.3, 1234
    params = perturbed_params(spec)
    img, lab = synth_batch(0, N, P, 1, 2)
    eng = engine_for(spec, P, N, "sorensen", (), emul_lib)
    eng.set_params(params)
    l1 = eng.forward_backward(img, lab, dropout_rate=rate, seed=seed)
    g1 = eng.get_grads()
    l2 = eng.forward_backward(img, lab, dropout_rate=rate, seed=seed)
    assert l1 == l2
    assert all(np.array_equal(g1[k], v) for k, v in eng.get_grads().items())
    l3 = eng.forward_backward(img, lab, dropout_rate=rate, seed=seed + 1)
    assert l3 != l1
    # recover the keep-masks from the activations (a == 0 exactly where dropped) and replay in the oracle
    eng.forward_backward(img, lab, dropout_rate=rate, seed=seed)
    masks = {}
    dims = {1: (8, 8, 8), 2: (4, 4, 4)}
    for name, c, sp in [("vnet/encoder/level_1/conv_1", 4, dims[1]), ("vnet/encoder/level_2/conv_1", 8, dims[2]),
                        ("vnet/encoder/level_2/conv_2", 8, dims[2]), ("vnet/bottom_level/conv_1", 16, (2, 2, 2)),
                        ("vnet/bottom_level/conv_2", 16, (2, 2, 2)), ("vnet/decoder/level_2/conv_1", 8, dims[2]),
                        ("vnet/decoder/level_2/conv_2", 8, dims[2]), ("vnet/decoder/level_1/conv_1", 4, dims[1])]:
        a = eng.read_tensor(name, 0, N, c, sp)
        masks[name] = torch.from_numpy((a != 0).astype(np.float32))
        keep = float(masks[name].mean())
        assert 0.45 < keep < 0.95
    lo, _, go, _ = R.loss_and_grads(params, img, lab, spec, "sorensen", (), dropout_rate=rate, masks=masks)
    assert abs(float(lo) - l1) < 5e-6
    _grad_check(eng, go, spec, 5e-4)
    eng.close()


def test_abi_error_behaviour(emul_lib):
    """Errors come back as negative status + message, never as exceptions across the C boundary."""
    import ctypes as C
    cfg = _ffi.VnbConfig()
    h = C.c_void_p()
    rc = emul_lib.vnb_create(C.byref(cfg), 0, C.byref(h))  # all-zero config
    assert rc == -1 and b"BottomConvolutions" in emul_lib.vnb_last_error()
    with pytest.raises(AssertionError):
        engine_for(SPEC_A, 8, 1, "weighted_sorensen", (1.0,), emul_lib)  # model.py:71 length assert
    with pytest.raises(SystemExit):
        engine_for(SPEC_A, 8, 1, "not_a_loss", (), emul_lib)             # model.py:559-560
    with pytest.raises(_ffi.VnbError):
        engine_for(SPEC_A, 10, 1, "sorensen", (), emul_lib)              # 10 is not divisible by 2^NumLevels
    eng = engine_for(SPEC_A, 8, 1, "sorensen", (), emul_lib)
    with pytest.raises(ValueError):
        eng.forward(np.zeros((2, 8, 8, 8, 1), np.float32))               # batch > max_batch
    with pytest.raises(KeyError):
        eng.get_param("vnet/nope")
    rc = emul_lib.vnb_get_param(eng._h, b"vnet/nope", None, 0)
    assert rc == -1
    rc = emul_lib.vnb_forward(None, None, 1, None, None, None)
    assert rc == -1 and b"handle" in emul_lib.vnb_last_error()
    eng.close()


def test_sixteen_channel_net_tight_gradients(emul_lib):
    """16-channel network in fp32 mode: exercises the vectorised BN passes and the tiled 2x2x2 kernels
    (which need channel multiples of 16) against tight oracle bounds."""
    spec = R.VNetSpec(num_classes=2, in_channels=1, num_channels=16, num_levels=2, num_convolutions=(1, 2), bottom_convolutions=1)
    P, N = 8, 2
    params = perturbed_params(spec)
    img, lab = synth_batch(1, N, P, 1, 2)
    eng = engine_for(spec, P, N, "weighted_sorensen", (0.1, 1.0), emul_lib)
    eng.set_params(params)
    l = eng.forward_backward(img, lab, update_moving_stats=True)
    lo, lg, go, upd = R.loss_and_grads(params, img, lab, spec, "weighted_sorensen", (0.1, 1.0))
    logits, _, _ = eng.forward(img)
    assert abs(l - float(lo)) < 2e-6
    assert rel_err(logits, lg.numpy()) < 2e-5
    _grad_check(eng, go, spec, 5e-4)
    eng.close()


@pytest.mark.parametrize("convs,bottom,loss", [((1, 2, 2), 3, "jaccard"), ((2, 1), 1, "weighted_sorensen")])
def test_legacy_vnet_py_flavour_matches_oracle(emul_lib, convs, bottom, loss):
    """VNet.py graph (two BNs per conv, true residual added between them) as used by
    train.py:271-279 (num_levels=3, (1,2,2), bottom 3, prelu)."""
    spec = R.VNetSpec(num_classes=2, in_channels=1, num_channels=4, num_levels=len(convs), num_convolutions=convs,
                      bottom_convolutions=bottom, flavour="legacy")
    P, N = (16 if len(convs) == 3 else 8), 2   # keep >= 2^3 voxels at the bottom level (batch-norm conditioning)
    weights = (0.1, 1.0) if "weighted" in loss else ()
    params = perturbed_params(spec)
    img, lab = synth_batch(0, N, P, 1, 2)
    eng = engine_for(spec, P, N, loss, weights, emul_lib)
    assert list(eng.variables()) == [n for n, _, _ in R.param_specs(spec)]
    eng.set_params(params)
    lo, lg, go, upd = R.loss_and_grads(params, img, lab, spec, loss, weights)
    logits, _, am = eng.forward(img)
    assert rel_err(logits, lg.numpy()) < 2e-5
    assert int((am != R.predict(lg).numpy()).sum()) == 0
    l = eng.forward_backward(img, lab, update_moving_stats=True)
    assert abs(l - float(lo)) < 2e-6
    _grad_check(eng, go, spec, 3e-4)
    for k, u in upd.items():
        assert np.abs(eng.get_param(k) - u.numpy()).max() < 1e-4 * max(1.0, float(u.abs().max())), k
    eng.close()


@pytest.mark.parametrize("flavour,K,loss,att_loss", [
    ("legacy", 2, "jaccard", "l2"),        # train.py defaults: VNet.py graph, --loss_function jaccard, --attention_loss_function l2
    ("legacy", 2, "sorensen_fg", "abs"),   # train.py:373-377 + :394-398
    ("networks", 3, "weighted_sorensen", "l2"),  # BASELINE config #5 shape: 2 modalities, 3 classes, model.py loss
])
def test_attention_gating_path_matches_oracle(emul_lib, flavour, K, loss, att_loss):
    """V-Net -> AttentionModule -> (1 + softmax) gating -> OutputModule, Dice + attention loss
    (train.py:281-312,351-418; attention.py:83-114; OutputModule.py:83-114)."""
    from tests.helpers import perturbed_attention_params
    from vnet_tensorflow_b200.synthetic import synth_patch
    M = 2 if K == 3 else 1
    spec = R.VNetSpec(num_classes=K, in_channels=M, num_channels=4, num_levels=2, num_convolutions=(1, 2),
                      bottom_convolutions=1, flavour=flavour)
    P, N, nch = 8, 2, 8
    weights = (0.01, 0.1, 1.0) if "weighted" in loss else ()
    params = perturbed_attention_params(spec, nch)
    samples = [synth_patch(1234 + 100000 * i, P, M, K) for i in range(N)]
    img, lab, dm = (np.stack([s[j] for s in samples], 0) for j in range(3))
    eng = engine_for(spec, P, N, loss, weights, emul_lib, attention=True, attention_loss=att_loss, module_channels=nch)
    assert list(eng.variables()) == [n for n, _, _ in R.attention_param_specs(spec, nch)]
    eng.set_params(params)
    tot, seg, att, out, go, upd = R.attention_loss_and_grads(params, img, lab, dm, spec, loss, att_loss, weights=weights)
    logits, softmax, am = eng.forward(img)
    assert rel_err(logits, out["logits_output"].numpy()) < 3e-5
    assert int((am != R.predict(out["logits_output"]).numpy()).sum()) == 0
    assert rel_err(eng.softmax_attention(N), out["softmax_attention"].numpy()) < 3e-5
    assert rel_err(eng.read_tensor("masked_vnet", 0, N, K, (P, P, P)), out["logits_masked"].numpy()) < 3e-5
    with pytest.raises(_ffi.VnbError):  # attention loss configured but no distance map fed
        eng.loss(img, lab)
    eng.set_distmap(dm)
    l = eng.forward_backward(img, lab, update_moving_stats=True)
    t3 = eng.losses()
    assert abs(l - float(tot)) < 3e-6 * max(1.0, abs(float(tot)))
    assert abs(t3[1] - float(seg)) < 3e-6 and abs(t3[2] - float(att)) < 3e-6 * max(1.0, abs(float(att)))
    g = eng.get_grads()
    assert set(g) == set(go)
    scale = max(float(np.abs(v.numpy()).max()) for v in go.values())
    for k, v in g.items():
        ref = go[k].numpy()
        if analytically_zero(k, spec):
            assert np.abs(v).max() <= 1e-6 * scale + 1e-12, k
            continue
        assert np.abs(v - ref).max() <= 3e-4 * max(np.abs(ref).max(), 1e-3 * scale), k
    for k, u in upd.items():  # only the V-Net's batch norms run UPDATE_OPS; the modules' moving statistics stay put
        assert np.abs(eng.get_param(k) - u.numpy()).max() < 1e-4 * max(1.0, float(u.abs().max())), k
    for k in params:
        if k.startswith(("AttentionModule/", "output/")) and k.endswith(("moving_mean", "moving_variance")):
            assert np.array_equal(eng.get_param(k), params[k]), k
    eng.close()


def test_attention_path_training_steps_follow_oracle(emul_lib):
    """Adam over the gated network: V-Net, attention and output module variables all move as in the oracle."""
    from tests.helpers import perturbed_attention_params
    from vnet_tensorflow_b200.synthetic import synth_patch
    spec = R.VNetSpec(num_classes=2, in_channels=1, num_channels=4, num_levels=1, num_convolutions=(1,),
                      bottom_convolutions=1, flavour="legacy")
    P, N, nch = 8, 1, 4
    params = perturbed_attention_params(spec, nch)
    im, lb, dm = (a[None] for a in synth_patch(7, P, 1, 2))
    eng = engine_for(spec, P, N, "jaccard", (), emul_lib, attention=True, attention_loss="l2", module_channels=nch,
                     learning_rate=1e-3)
    eng.set_params(params)
    eng.set_distmap(dm)
    p = {k: v.copy() for k, v in params.items()}
    m = {k: torch.zeros(v.shape) for k, v in p.items()}
    v2 = {k: torch.zeros(v.shape) for k, v in p.items()}
    for step in range(2):
        tot, _, _, _, go, upd = R.attention_loss_and_grads(p, im, lb, dm, spec, "jaccard", "l2")
        le = eng.train_step(im, lb)
        assert abs(le - float(tot)) < 1e-4 * max(1.0, abs(float(tot))), (step, le, float(tot))
        lr = R.learning_rate(1e-3, step, 100.0, 0.99)
        for k, g in go.items():
            pk, m[k], v2[k] = R.adam_update(torch.from_numpy(p[k]), g, m[k], v2[k], step + 1, lr)
            p[k] = pk.numpy()
        for k, u in upd.items():
            p[k] = u.numpy()
    for k in ("attention/AttentionModule/encoder/Variable", "output/output/output/Variable_1", "output/encoder/batch_normalization_8/gamma"):
        assert np.abs(eng.get_param(k) - p[k]).max() < 2e-4, k  # two Adam steps move each weight by ~2e-3
    eng.close()


@pytest.mark.parametrize("spec,N", [(SPEC_A, 2), (SPEC_B, 1)])
def test_every_unit_against_the_oracle_taps(emul_lib, spec, N):
    """The per-op view the fused engine offers (vnb_read_tensor around vnb_forward_backward): for every unit the
    tensor its first batch norm normalises (convolution output, plus the block input where networks.py:318 adds it),
    the unit's output activation after the BN chain / PReLU, and dL/d(that BN input) - against the oracle's taps.
    Covers layers2.convolution / down_convolution / up_convolution, the batch-norm chains, PReLU and, at the output
    layer, the softmax-Dice backward (dL/dlogits reaches the head's batch norm)."""
    P = 8
    params = perturbed_params(spec, 3)
    img, lab = synth_batch(1, N, P, spec.in_channels, spec.num_classes)
    weights = (0.1, 0.5, 1.0)[-spec.num_classes:]
    eng = engine_for(spec, P, N, "weighted_sorensen", weights, emul_lib)
    eng.set_params(params)
    loss = eng.forward_backward(img, lab, dropout_rate=0.0)
    col = {}
    tp = R.to_torch(params, torch.float64, requires_grad=True)
    logits, _ = R.forward(tp, torch.from_numpy(img).double(), spec, collect=col)
    for t in col.values():
        if t.requires_grad:  # (the tiled image in front of the input batch norm is a constant)
            t.retain_grad()
    ref_loss = R.loss_from_logits(logits, torch.from_numpy(lab), "weighted_sorensen", weights)
    ref_loss.backward()
    assert abs(loss - float(ref_loss.detach())) < 2e-6
    seen = {"act": 0, "bn_in": 0}
    for name, t in col.items():
        scope, _, kind = name.partition(":")
        shape = tuple(t.shape)
        if kind == "bn_in":
            if scope == "vnet/input_layer" and spec.in_channels == 1:
                continue  # the tiled image itself (networks.py:258): no convolution unit in front of this batch norm
            z = eng.read_tensor(scope, 2, N, shape[-1], shape[1:4])
            dz = eng.read_tensor(scope, 1, N, shape[-1], shape[1:4])
            assert rel_err(z, t.detach().numpy()) < 5e-6, name
            assert rel_err(dz, t.grad.numpy()) < 2e-5, name
            seen["bn_in"] += 1
        else:
            a = eng.read_tensor(scope, 0, N, shape[-1], shape[1:4])
            assert rel_err(a, t.detach().numpy()) < 5e-6, name
            seen["act"] += 1
    n_units = sum(1 for k in eng.variables() if k.endswith("/weights"))
    assert seen["bn_in"] >= n_units - 2 * spec.num_levels and seen["act"] >= n_units  # up-convolutions tap outputs only
    eng.close()


def test_profile_records_name_every_convolution_launch(emul_lib):
    """vnb_profile_count / vnb_profile_launch (the per-layer roofline table of bench.py --per-layer): one
    forward + backward pass yields one labelled record per 5^3 convolution and pass, whose algorithmic FLOPs are
    2 * 125 * Cin * Cout * voxels, and the per-class totals agree with vnb_profile_read."""
    spec, P, N = SPEC_A, 8, 2
    eng = engine_for(spec, P, N, "weighted_sorensen", (0.1, 1.0), emul_lib)
    eng.set_params(perturbed_params(spec))
    img, lab = synth_batch(0, N, P, spec.in_channels, spec.num_classes)
    assert eng.profile_launches() == []
    eng.profile_enable(True)
    eng.forward_backward(img, lab)
    recs = eng.profile_launches()
    conv5 = [n[:-len("/weights")] for n, s, k in R.param_specs(spec) if k == "weights" and tuple(s[:3]) == (5, 5, 5)]
    by_pass = {p: [r for r in recs if r[0].split(" ")[1] == p] for p in ("fprop", "dgrad", "wgrad")}
    assert sorted(r[0].split(" ")[0] for r in by_pass["fprop"]) == sorted(conv5)
    assert sorted(r[0].split(" ")[0] for r in by_pass["wgrad"]) == sorted(conv5)
    assert len(by_pass["dgrad"]) == len(conv5)       # M = 1: the input convolution feeds a batch norm, so it has a dgrad
    shapes = {n[:-len("/weights")]: s for n, s, k in R.param_specs(spec) if k == "weights"}
    for label, cls, ms, flops in recs:
        scope, pas, chans, at = label.split(" ")
        cin, cout = (int(v) for v in chans.split("->"))
        d, h, w = (int(v) for v in at[1:].split("x"))
        assert (cin, cout) == tuple(shapes[scope][3:5])
        assert flops == 2.0 * 125 * cin * cout * N * d * h * w
        assert cls == (1 if pas == "wgrad" else 0) and ms == 0.0      # no device clock in the emulation build
    for cls in (0, 1):
        _, n, fl = eng.profile_read(cls)
        assert n == sum(r[1] == cls for r in recs) and fl == sum(r[3] for r in recs if r[1] == cls)
    eng.profile_enable(False)
    assert eng.profile_launches() == []
    eng.close()


def test_ragged_batches_below_max_batch(emul_lib):
    """A handle created for max_batch = 3 serves batches of 3, 1 and 2 in turn (the last batch of an epoch is short,
    model.py:716-748 feeds whatever the dataset yields): batch statistics, loss mean over n and gradients belong to
    the n patches of the call, nothing of an earlier, larger batch leaks in; n = 0 and n > max_batch are errors."""
    spec, P = SPEC_A, 8
    params = perturbed_params(spec)
    eng = engine_for(spec, P, 3, "weighted_sorensen", (0.1, 1.0), emul_lib)
    eng.set_params(params)
    for seed, n in ((0, 3), (1, 1), (2, 2)):
        img, lab = synth_batch(seed, n, P, spec.in_channels, spec.num_classes)
        loss_o, logits_o, grads_o, _ = R.loss_and_grads(params, img, lab, spec, "weighted_sorensen", (0.1, 1.0))
        loss = eng.forward_backward(img, lab)
        assert abs(loss - float(loss_o)) < 2e-6, n
        _grad_check(eng, grads_o, spec, 2e-4)
        logits, _, argmax = eng.forward(img)
        assert logits.shape[0] == n and rel_err(logits, logits_o.numpy()) < 2e-5
        assert int((argmax != R.predict(logits_o).numpy()).sum()) == 0
    with pytest.raises(ValueError):
        eng.forward(np.zeros((4, P, P, P, 1), np.float32))
    with pytest.raises((ValueError, _ffi.VnbError)):
        eng.forward(np.zeros((0, P, P, P, 1), np.float32))
    eng.close()


@pytest.mark.parametrize("spec,N", [(SPEC_A, 2), (SPEC_B, 1)])
def test_step_metrics_match_the_reference_metric_block(emul_lib, spec, N):
    """vnb_read_metrics (model.py:586-626, part of summary_op in the training sess.run): the confusion counts are
    bit-exact against argmax / labels, the AUC histograms bit-exact against the engine's own softmax at the 200
    tf.metrics.auc thresholds, and the derived scalars equal the literal restatement of the reference block -
    including labels outside [0, K) (all-zero one-hot row) and a class that never occurs (0/0 = NaN like tf.divide)."""
    from oracle import ref_metrics as RM
    from vnet_tensorflow_b200 import metrics as M
    P, K = 8, spec.num_classes
    eng = engine_for(spec, P, N, "sorensen", (), emul_lib)
    eng.set_params(perturbed_params(spec))
    img, lab = synth_batch(2, N, P, spec.in_channels, K)
    lab = lab.copy()
    lab[0, 0, 0, :3] = (K, -1, K + 5)           # no class
    if K == 3:
        lab[lab == 2] = 0                       # class 2 never occurs as a label
    with pytest.raises(_ffi.VnbError):
        eng.metric_counts(N)                    # no call with labels yet
    eng.forward_backward(img, lab)
    cm, hist = eng.metric_counts(N)
    logits, softmax, argmax = eng.forward(img)
    with pytest.raises(_ffi.VnbError):
        eng.metric_counts(N)                    # vnb_forward dropped the pairing of logits and labels
    want_cm = np.zeros_like(cm)
    for t, p in zip(lab.reshape(-1), argmax.reshape(-1)):
        want_cm[t if 0 <= t < K else K, p] += 1
    assert np.array_equal(cm, want_cm) and int(cm.sum()) == lab.size
    thr = M.auc_thresholds()
    want_hist = np.zeros_like(hist)
    for c in range(1, K):
        bins = (softmax[..., c].reshape(-1)[:, None] > thr[None, :]).sum(1)
        np.add.at(want_hist[c], ((lab.reshape(-1) == c).astype(int), bins), 1)
    assert np.array_equal(hist, want_hist)
    got = M.step_metrics(cm, hist, list(range(K)))
    ref = RM.step_metrics(logits, lab, softmax, list(range(K)))
    assert list(got) == list(ref)
    for k in got:
        same = got[k] == ref[k] or (np.isnan(got[k]) and np.isnan(ref[k]))
        assert same or (k.startswith("auc_") and abs(got[k] - ref[k]) < 1e-6), (k, got[k], ref[k])
    if K == 3:
        assert np.isnan(got["sensitivity_2"]) and got["true_positives_2"] == 0
    eng.loss(img, lab)
    cm2, hist2 = eng.metric_counts(N, want_auc=False)
    assert hist2 is None and np.array_equal(cm2, cm)
    eng.close()


def test_staged_input_steps_equal_direct_steps(emul_lib):
    """vnb_stage_batch / vnb_train_step_staged (the feed_dict copy taken off the critical path) is the same arithmetic
    as vnb_train_step: three steps with a one-batch look-ahead from page-locked arrays give bit-identical losses and
    weights; a staged step without a staged batch is an error; a short last batch is staged like any other."""
    spec, P = SPEC_A, 8
    params = perturbed_params(spec)
    batches = [synth_batch(s, n, P, spec.in_channels, spec.num_classes) for s, n in ((0, 2), (1, 2), (2, 1))]
    direct = engine_for(spec, P, 2, "weighted_sorensen", (0.1, 1.0), emul_lib)
    direct.set_params(params)
    want = [direct.train_step(img, lab, 0.0, seed=i) for i, (img, lab) in enumerate(batches)]
    eng = engine_for(spec, P, 2, "weighted_sorensen", (0.1, 1.0), emul_lib)
    eng.set_params(params)
    with pytest.raises(_ffi.VnbError):
        eng.train_step_staged()
    pin_img = [eng.pinned_array((2, P, P, P, 1), np.float32) for _ in range(2)]
    pin_lab = [eng.pinned_array((2, P, P, P), np.int32) for _ in range(2)]

    def stage(i):
        img, lab = batches[i]
        n = img.shape[0]
        pin_img[i % 2][:n], pin_lab[i % 2][:n] = img, lab
        eng.stage_batch(pin_img[i % 2][:n], pin_lab[i % 2][:n])

    got = []
    stage(0)
    for i in range(len(batches)):
        eng.train_step_staged(0.0, seed=i, want_loss=False)
        if i + 1 < len(batches):
            stage(i + 1)              # overlaps step i on the device
        got.append(eng.last_loss())
    assert got == want
    for k in ("vnet/encoder/level_1/conv_1/weights", "vnet/output_layer/weights", "vnet/decoder/level_1/up_convolution/weights"):
        assert np.array_equal(eng.get_param(k), direct.get_param(k)), k
    assert eng.global_step == direct.global_step == 3
    cm, _ = eng.metric_counts(1, want_auc=False)          # the labels of the staged short batch are the current ones
    assert int(cm.sum()) == P ** 3
    eng.close()
    direct.close()


class _FakeCudaTensor:
    """Stand-in for a torch CUDA tensor over NumPy memory: in the emulation build a 'device pointer' is a host pointer,
    so VNetEngine's device-tensor path (dtype / contiguity / device checks, raw data_ptr hand-over) runs on the CPU."""
    __module__ = "torch"
    is_cuda = True

    class _Dev:
        index = 0

    def __init__(self, array, torch_dtype):
        self._a = np.ascontiguousarray(array)
        self.dtype, self.device = torch_dtype, self._Dev()
        self.shape, self.ndim = self._a.shape, self._a.ndim

    def is_contiguous(self):
        return True

    def data_ptr(self):
        return self._a.ctypes.data

    def __getitem__(self, idx):
        return _FakeCudaTensor(self._a[idx], self.dtype)


def test_device_tensor_inputs_take_the_raw_pointer_path(emul_lib, monkeypatch):
    """CPU twin of the GPU test test_device_resident_batches_give_identical_results: same results as host arrays,
    [N,X,Y,Z,1] labels accepted as a view, wrong dtype / wrong GPU refused before the C ABI is reached."""
    class _Stream:
        def synchronize(self):
            pass
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stream())
    spec, P, N = SPEC_A, 8, 2
    img, lab = synth_batch(4, N, P, 1, 2)
    eng = engine_for(spec, P, N, "weighted_sorensen", (0.1, 1.0), emul_lib)
    eng.set_params(perturbed_params(spec))
    want_logits = eng.forward(img)[0]
    want_loss = eng.loss(img, lab)
    d_img, d_lab = _FakeCudaTensor(img, torch.float32), _FakeCudaTensor(lab[..., None], torch.int32)
    assert np.array_equal(eng.forward(d_img)[0], want_logits)
    assert eng.loss(d_img, d_lab) == want_loss
    assert eng.forward_backward(d_img, d_lab) == want_loss
    assert eng.stage_batch(d_img, d_lab) == N and eng.train_step_staged(want_loss=True) == want_loss
    with pytest.raises(ValueError):
        eng.forward(_FakeCudaTensor(img.astype(np.float64), torch.float64))
    with pytest.raises(ValueError):
        eng.loss(d_img, _FakeCudaTensor(lab.astype(np.int64), torch.int64))
    other = _FakeCudaTensor(img, torch.float32)
    other.device.index = 1      # a tensor on another GPU than the handle's
    with pytest.raises(ValueError):
        eng.forward(other)
    eng.close()


@pytest.mark.parametrize("loss,weights,alpha", [("mixed_sorensen", (), 0.7), ("mixed_weighted_jaccard", (0.2, 1.0), 1.5),
                                                ("weighted_sorensen", (0.1, 1.0), 1.0), ("xent", (), 1.0)])
def test_loss_parts_are_the_two_summaries_of_the_mixed_losses(emul_lib, loss, weights, alpha):
    """vnb_read_loss_parts: '1.dice' = 1 - dice and '2.regularized_xent' = Loss.Alpha * xent (model.py:529-530)."""
    spec, P, N = SPEC_A, 8, 2
    params = perturbed_params(spec)
    img, lab = synth_batch(1, N, P, 1, 2)
    eng = engine_for(spec, P, N, loss, weights, emul_lib, loss_alpha=alpha)
    eng.set_params(params)
    total = eng.loss(img, lab)
    dice_part, xent_part = eng.loss_parts()
    logits = R.forward(R.to_torch(params), torch.from_numpy(img), spec)[0]
    want_total = float(R.loss_from_logits(logits, torch.from_numpy(lab), loss, weights, alpha))
    assert abs(total - want_total) < 2e-6 and abs(dice_part + xent_part - total) < 1e-6
    if loss == "xent":
        assert dice_part == 0.0
    else:
        kind = "sorensen" if "sorensen" in loss else "jaccard"
        onehot = torch.nn.functional.one_hot(torch.from_numpy(lab).long(), 2).float()
        d = R.dice_coe(torch.softmax(logits, -1), onehot, loss_type=kind, weights=tuple(weights) if "weighted" in loss else ())
        assert abs(dice_part - (1.0 - float(d))) < 2e-6
        if not loss.startswith("mixed"):
            assert abs(xent_part) < 1e-6
    eng.close()


# ---- per-op hooks: CPU twins of the GPU tests in tests/test_gpu_parity.py -----------------------------------
def test_op_hooks_k2_bn_softmax_dice_adam(emul_lib):
    from tests import op_hook_cases as H
    H.check_k2_ops(emul_lib, "fp32", 16, 32, (3, 4, 5))
    H.check_k2_ops(emul_lib, "bf16x3", 16, 32, (2, 4, 8))       # tcgen05 / TMA kernels (k2_tc.cuh) on the CPU model of the async units
    H.check_k2_ops(emul_lib, "bf16x3", 32, 64, (3, 5, 12))      # grid that no 128-voxel box tiles exactly: zero-filled loads, clipped stores
    H.check_k2_ops(emul_lib, "bf16x3", 64, 128, (2, 2, 4), n=1)  # scatter in two 256-column blocks, filter gradient in 64-column blocks
    H.check_k2_ops(emul_lib, "bf16x3", 16, 24, (2, 4, 8))       # CC % 32 != 0: the mma.sync 3xTF32 tiles (CPU model of the warp MMA)
    H.check_k2_ops(emul_lib, "fp32", 4, 6, (2, 3, 2), n=1)      # channels outside the tiled kernels' domain
    H.check_bn_ops(emul_lib, 4096, 16)
    H.check_bn_ops(emul_lib, 1000, 128, with_alpha=False)
    H.check_softmax_dice_ops(emul_lib, "weighted_sorensen", (0.1, 0.5, 1.0), 1.0)
    H.check_softmax_dice_ops(emul_lib, "mixed_weighted_jaccard", (0.2, 0.3, 1.0), 1.5)
    H.check_softmax_dice_ops(emul_lib, "xent", (), 1.0, k=2)
    H.check_adam_op(emul_lib)
