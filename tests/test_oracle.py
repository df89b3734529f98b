"""CPU oracle self-checks: the oracle is test infrastructure, so it is pinned by
(i) fp64 finite differences, (ii) invariances that hold by construction of the reference graph,
(iii) analytic Dice cases, (iv) the committed golden fixtures (regression guard)."""
import numpy as np
import pytest
import torch

from oracle import ref_vnet as R
from tests.helpers import CASES, load_golden, perturbed_params
from vnet_tensorflow_b200.synthetic import synth_batch

TINY = R.VNetSpec(num_classes=2, in_channels=1, num_channels=4, num_levels=2, num_convolutions=(1, 2), bottom_convolutions=2)
TINY3 = R.VNetSpec(num_classes=3, in_channels=2, num_channels=4, num_levels=2, num_convolutions=(3, 1), bottom_convolutions=1)


def test_variable_inventory_default_net():
    spec = R.VNetSpec()
    specs = R.param_specs(spec)
    names = [n for n, _, _ in specs]
    assert len(set(names)) == len(names)
    conv_params = sum(int(np.prod(s)) for n, s, k in specs if k == "weights")
    assert conv_params == 43_925_152 + 0 or abs(conv_params - 43.93e6) < 0.02e6  # 43.93 M
    assert sum(k == "weights" for _, _, k in specs) == 30
    assert sum(k == "gamma" for _, _, k in specs) == 38
    assert sum(k == "alpha" for _, _, k in specs) == 29
    assert "vnet/decoder/level_1/conv_1/batch_normalization_2/gamma" in names
    assert dict((n, s) for n, s, _ in specs)["vnet/decoder/level_3/up_convolution/weights"] == (2, 2, 2, 64, 128)


@pytest.mark.parametrize("spec,loss,weights", [(TINY, "weighted_sorensen", (0.1, 1.0)), (TINY3, "mixed_jaccard", (0.2, 0.5, 1.0))])
def test_gradients_match_fp64_finite_differences(spec, loss, weights):
    p = perturbed_params(spec)
    img, lab = synth_batch(3, 2, 8, spec.in_channels, spec.num_classes)
    _, _, g, _ = R.loss_and_grads(p, img, lab, spec, loss, weights, dtype=torch.float64)

    def L(pp):
        P = R.to_torch(pp, torch.float64)
        lg, _ = R.forward(P, torch.from_numpy(img).double(), spec)
        return float(R.loss_from_logits(lg, torch.from_numpy(lab), loss, weights))

    rng = np.random.default_rng(0)
    names = [n for n in g if n.endswith(("weights", "gamma", "alpha"))]
    for n in names[::2]:
        idx = tuple(int(rng.integers(0, s)) for s in p[n].shape)
        pp = {k: v.astype(np.float64).copy() for k, v in p.items()}
        eps = 1e-7  # small step: PReLU kinks make the loss only piecewise smooth
        pp[n][idx] += eps
        lp = L(pp)
        pp[n][idx] -= 2 * eps
        lm = L(pp)
        fd, an = (lp - lm) / (2 * eps), float(g[n][idx])
        assert abs(fd - an) <= 2e-5 * max(abs(fd), abs(an)) + 1e-8, (n, fd, an)


def test_logits_invariant_to_conv_biases_and_last_weight_scale():
    """every conv is followed by a batch-statistics BN, so biases cannot change the logits;
    the output BN also removes any positive rescaling of the 1x1 weights."""
    p = perturbed_params(TINY)
    img, _ = synth_batch(1, 2, 8, 1, 2)
    x = torch.from_numpy(img).double()
    base, _ = R.forward(R.to_torch(p, torch.float64), x, TINY)
    q = {k: v.copy() for k, v in p.items()}
    rng = np.random.default_rng(5)
    for k in q:
        if k.endswith("biases"):
            q[k] = rng.normal(0, 3, q[k].shape).astype(np.float32)
    out, _ = R.forward(R.to_torch(q, torch.float64), x, TINY)
    assert float((out - base).abs().max()) < 1e-9
    q["vnet/output_layer/weights"] = q["vnet/output_layer/weights"] * 7.5
    out, _ = R.forward(R.to_torch(q, torch.float64), x, TINY)
    # epsilon = 1e-3 sits inside the sqrt, so the rescaling is only removed up to O(eps / var)
    assert float((out - base).abs().max()) < 5e-2


def test_same_padding_and_shapes():
    p = R.to_torch(R.init_params(TINY))
    x = torch.zeros(1, 8, 8, 8, 1)
    logits, upd = R.forward(p, x + 1.0, TINY)
    assert tuple(logits.shape) == (1, 8, 8, 8, 2)
    assert len(upd) == 2 * sum(1 for n, _, k in R.param_specs(TINY) if k == "gamma")


def test_decoder_quirk_is_x_plus_bn_x():
    """networks.py:358-360: layer_input is overwritten with BN(x), so the 'residual' is x + BN(x)."""
    spec = TINY
    p = R.to_torch(perturbed_params(spec), torch.float64)
    img, _ = synth_batch(2, 1, 8, 1, 2)
    col = {}
    R.forward(p, torch.from_numpy(img).double(), spec, collect=col)
    sc = "vnet/decoder/level_2/conv_2"
    x_in = col["vnet/decoder/level_2/conv_1"]
    z = R.conv_same(x_in, p[sc + "/weights"], p[sc + "/biases"])
    cx = R._Ctx(p, None)
    want = R.prelu(cx.bn(z + cx.bn(z, sc, 0), sc, 1), p[sc + "/alpha"])
    assert float((want - col[sc]).abs().max()) < 1e-12


def test_dice_known_answers():
    s = 1e-5
    t = torch.zeros(1, 4, 4, 4, 2)
    t[..., 0] = 1
    assert abs(float(R.dice_coe(t, t, "sorensen")) - np.mean([(2 * 64 + s) / (128 + s), 1.0])) < 1e-7  # class 1 empty/empty -> 1
    o = torch.zeros_like(t)
    o[..., 1] = 1  # disjoint
    d = float(R.dice_coe(o, t, "jaccard"))
    assert abs(d - s / (64 + s)) < 1e-9
    w = (0.1, 1.0)  # weighted form adds `smooth` once per class (model.py:74)
    num = 2 * (0.1 * 64) + 2 * s
    den = 0.1 * 128 + 2 * s
    assert abs(float(R.dice_coe(t, t, "sorensen", weights=w)) - num / den) < 1e-7


def test_argmax_tie_and_onehot_out_of_range():
    lg = torch.zeros(1, 1, 1, 2, 3)
    lg[0, 0, 0, 1, 2] = 1.0
    assert R.predict(lg).flatten().tolist() == [0, 2]
    lab = torch.tensor([[[[5, 1]]]], dtype=torch.int32)  # 5 is out of range -> all-zero one-hot
    terms = R.dice_terms(lg, lab)
    assert float(terms[0, :, 2].sum()) == 1.0


def test_lr_schedule_and_adam_match_tf_forms():
    assert abs(R.learning_rate(1e-2, 250, 100, 0.99) - 1e-2 * 0.99 ** 2.5) < 1e-12
    p, g = torch.tensor([1.0]), torch.tensor([0.5])
    p1, m1, v1 = R.adam_update(p, g, torch.zeros(1), torch.zeros(1), 1, 1e-2)
    lr_t = 1e-2 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert abs(float(p1) - (1.0 - lr_t * 0.05 / (np.sqrt(0.00025) + 1e-8))) < 1e-7


def test_window_starts_last_window_clamped():
    assert R.window_starts(300, 128, 64) == [0, 64, 128, 172]
    assert R.window_starts(128, 128, 64) == [0]


@pytest.mark.parametrize("name", ["tiny_m1_k2", "tiny_m2_k3"])
def test_oracle_reproduces_golden(name):
    kw, P, N, loss, weights = CASES[name]
    spec = R.VNetSpec(**kw)
    gold = load_golden(name)
    img, lab = synth_batch(0, N, P, spec.in_channels, spec.num_classes)
    l, logits, grads, _ = R.loss_and_grads(perturbed_params(spec), img, lab, spec, loss, weights)
    assert abs(float(l) - float(gold["loss"])) < 2e-6
    assert np.abs(logits.numpy() - gold["logits"]).max() < 2e-4
    for k, g in grads.items():
        gn = float(np.sqrt((g.numpy().astype(np.float64) ** 2).sum()))
        assert abs(gn - float(gold["gnorm/" + k])) <= 2e-3 * max(gn, 1e-6) + 1e-7, k


def test_golden_manifest_matches_the_committed_fixtures():
    """tests/golden/MANIFEST.json: every .npz is listed with its SHA-256 and generator script, and
    nothing listed is missing - a fixture regenerated without updating the manifest (or the reverse) fails here."""
    import hashlib
    import json
    import os

    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with open(os.path.join(gold, "MANIFEST.json")) as f:
        manifest = json.load(f)
    on_disk = sorted(f for f in os.listdir(gold) if f.endswith(".npz"))
    assert on_disk == sorted(manifest)
    for name, entry in manifest.items():
        with open(os.path.join(gold, name), "rb") as f:
            blob = f.read()
        assert hashlib.sha256(blob).hexdigest() == entry["sha256"], name
        assert len(blob) == entry["bytes"]
        assert os.path.exists(os.path.join(gold, entry["generator"]))
