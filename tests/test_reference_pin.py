"""The oracle against golden vectors computed by the reference's own code.

tests/golden/ref_*.npz were written by tests/golden/make_reference_fixtures.py, which imports the reference's
`networks.py` / `layers2.py` unmodified and compiles `dice_coe` out of its `model.py`, and runs them over the eager
TF-1 API stand-in tests/tf1_shim.py (see its docstring: the wiring of the reference graph is pinned, TensorFlow's
kernels are not).  The oracle restatement must reproduce variable inventory and creation order, logits, loss,
every gradient and the UPDATE_OPS moving statistics of that run."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_vnet as R
from tests.golden.make_golden import perturbed_params
from tests.golden.make_reference_fixtures import CASES, REFERENCE, grad_digest
from vnet_tensorflow_b200.synthetic import synth_batch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _check_gradients(grads, ref, tol):
    """Every trainable variable: gradient norm, sum and strided samples against the reference run's digest."""
    assert sorted(k for k in ref if k.startswith("gnorm/")) == sorted("gnorm/" + k for k in grads)
    gscale = max(float(ref[k]) for k in ref if k.startswith("gnorm/"))
    sscale = max(np.abs(ref[k]).max() for k in ref if k.startswith("gsample/"))
    for k, g in grads.items():
        norm, total, sample = grad_digest(g)
        assert abs(norm - float(ref["gnorm/" + k])) <= tol * gscale, k
        assert abs(total - float(ref["gsum/" + k])) <= tol * gscale * np.sqrt(g.numel()), k
        np.testing.assert_allclose(sample, ref["gsample/" + k], rtol=0, atol=tol * sscale, err_msg=k)


def _loss_name(loss_type, weights):
    return ("weighted_" if len(weights) else "") + loss_type


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_the_reference_code(name):
    kw, M, P, N, loss_type, weights = CASES[name]
    spec = R.VNetSpec(in_channels=M, flavour="legacy" if name.startswith("legacy_") else "networks", **kw)
    params = perturbed_params(spec)
    img, lab = synth_batch(0, N, P, M, kw["num_classes"])
    with np.load(os.path.join(GOLDEN, "ref_%s.npz" % name)) as z:
        ref = {k: z[k] for k in z.files}
    # the variables the reference code creates, in its creation order = the oracle's parameter inventory
    assert list(ref["variable_names"]) == list(params.keys())
    trainable = {n: bool(t) for n, t in zip(ref["variable_names"], ref["trainable"])}
    assert all(trainable[n] == (not n.endswith(("moving_mean", "moving_variance"))) for n in trainable)
    loss, logits, grads, updates = R.loss_and_grads(params, img, lab, spec, _loss_name(loss_type, weights), weights,
                                                    dtype=torch.float64)
    np.testing.assert_allclose(logits.numpy(), ref["logits"], rtol=0, atol=1e-9 * np.abs(ref["logits"]).max())
    # the reference casts Loss.Weights to float32 (model.py:73); the float64 oracle keeps them in double: 1e-8 relative
    assert abs(float(loss) - float(ref["loss"])) < 1e-8
    _check_gradients(grads, ref, 2e-7)
    for k, u in updates.items():
        np.testing.assert_allclose(u.numpy(), ref["moving/" + k], rtol=1e-6, atol=1e-7, err_msg=k)
    assert sorted(k for k in ref if k.startswith("moving/")) == sorted("moving/" + k for k in updates)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree only exists in the build container")
def test_fixtures_are_what_the_reference_code_computes_today():
    """Provenance guard: re-run the reference's code (one case) and compare with the committed fixture."""
    from tests.golden.make_reference_fixtures import run_reference
    name = "m1_k2_c12"
    kw, M, P, N, loss_type, weights = CASES[name]
    _, _, _, _, out = run_reference(kw, M, P, N, loss_type, weights, legacy=name.startswith("legacy_"))
    with np.load(os.path.join(GOLDEN, "ref_%s.npz" % name)) as z:
        assert list(z["variable_names"]) == list(out["variable_names"])
        np.testing.assert_array_equal(z["logits"], out["logits"])
        assert float(z["loss"]) == float(out["loss"])


def test_attention_path_matches_the_reference_modules():
    """VNet.py + attention.py + OutputModule.py executed as they are (composition of train.py:281-312 restated in the
    generator, see run_reference_attention): every named tensor, both losses and all gradients of the oracle."""
    from tests.golden.make_reference_fixtures import ATTENTION_CASE as c
    from vnet_tensorflow_b200.synthetic import synth_patch
    spec = R.VNetSpec(num_classes=2, in_channels=1, flavour="legacy", num_channels=c["num_channels"], num_levels=c["num_levels"],
                      num_convolutions=c["num_convolutions"], bottom_convolutions=c["bottom_convolutions"])
    params = R.init_attention_params(spec)
    im, lb, dm = synth_patch(77, c["P"], 1, 2)
    with np.load(os.path.join(GOLDEN, "ref_attention_k2.npz")) as z:
        ref = {k: z[k] for k in z.files}
    assert list(ref["variable_names"]) == list(params.keys())   # incl. attention/AttentionModule/encoder/Variable_k
    total, seg, att, out, grads, _ = R.attention_loss_and_grads(params, im[None], lb[None], dm[None], spec, loss="jaccard",
                                                                att_loss="l2", dtype=torch.float64)
    for k in ("logits_vnet", "softmax_attention", "logits_masked", "logits_output"):
        np.testing.assert_allclose(out[k].numpy(), ref[k], rtol=0, atol=1e-9 * max(1.0, np.abs(ref[k]).max()), err_msg=k)
    # train.py's dice_coe casts its sums to float32 (train.py:143): 1e-7 relative on the segmentation term
    assert abs(float(att) - float(ref["att_loss"])) < 1e-9 * float(ref["att_loss"])
    assert abs(float(seg) - float(ref["seg_loss"])) < 1e-6
    assert abs(float(total) - float(ref["total_loss"])) < 1e-6
    _check_gradients(grads, ref, 2e-6)


@pytest.mark.parametrize("name", ["m1_k2_c12", "legacy_m1_k2_c12"])
def test_cuda_engine_sources_match_the_reference_code(emul_lib, name):
    """The product's kernel + engine sources (compiled for the CPU, tests/emul) against the vectors computed by the
    reference's own code -- the same comparison the GPU suite makes against the oracle, one link shorter."""
    from vnet_tensorflow_b200.engine import VNetEngine
    kw, M, P, N, loss_type, weights = CASES[name]
    legacy = name.startswith("legacy_")
    spec = R.VNetSpec(in_channels=M, flavour="legacy" if legacy else "networks", **kw)
    params = perturbed_params(spec)
    img, lab = synth_batch(0, N, P, M, kw["num_classes"])
    with np.load(os.path.join(GOLDEN, "ref_%s.npz" % name)) as z:
        ref = {k: z[k] for k in z.files}
    eng = VNetEngine(num_classes=kw["num_classes"], in_channels=M, patch_shape=(P, P, P), max_batch=N,
                     num_channels=kw["num_channels"], num_levels=kw["num_levels"], num_convolutions=kw["num_convolutions"],
                     bottom_convolutions=kw["bottom_convolutions"], precision="fp32", loss=_loss_name(loss_type, weights),
                     loss_weights=weights, flavour="legacy" if legacy else "networks", library=emul_lib)
    assert list(eng.variables().keys()) == list(ref["variable_names"])   # TF names, TF creation order
    eng.set_params(params)
    loss = eng.forward_backward(img, lab)
    logits, _, _ = eng.forward(img)
    assert np.abs(logits - ref["logits"]).max() <= 2e-4 * np.abs(ref["logits"]).max()
    assert abs(loss - float(ref["loss"])) < 1e-5
    grads = eng.get_grads()
    gscale = max(float(ref[k]) for k in ref if k.startswith("gnorm/"))
    for k, g in grads.items():
        if k.endswith("biases"):   # analytically zero (every convolution feeds a batch-statistics norm)
            continue
        norm = float(np.sqrt((g.astype(np.float64) ** 2).sum()))
        assert abs(norm - float(ref["gnorm/" + k])) <= 2e-3 * gscale, k
    eng.close()


def test_loss_functions_match_model_py():
    """model.py:26-92 (`dice_coe` in its four forms, `weighted_softmax_cross_entropy_with_logits`) compiled from the
    reference file and run on random inputs, against the oracle's restatements."""
    with np.load(os.path.join(GOLDEN, "ref_losses.npz")) as z:
        ref = {k: z[k] for k in z.files}
    logits = torch.from_numpy(ref["logits"])
    lab = torch.from_numpy(ref["labels"])
    w = tuple(float(x) for x in ref["weights"])
    for name in ("sorensen", "jaccard", "weighted_sorensen", "weighted_jaccard"):
        loss = R.loss_from_logits(logits, lab, name, w)
        assert abs((1.0 - float(loss)) - float(ref["dice_" + name])) < 1e-8, name
    assert abs(float(R.loss_from_logits(logits, lab, "weighted_xent", w)) - float(ref["weighted_xent"])) < 1e-8
    # layers2.py:4-30 initialisers: uniform in +-sqrt(6 / (patch volume * (Cin + Cout))), zeros for the biases
    for key, shape in (("xavier_5x5x5x16x32_absmax", (5, 5, 5, 16, 32)), ("xavier_2x2x2x32x16_absmax", (2, 2, 2, 32, 16))):
        lim = np.abs(R.xavier_uniform(shape, np.random.Generator(np.random.PCG64(0)))).max()
        bound = np.sqrt(6.0 / (np.prod(shape[:3]) * (shape[3] + shape[4])))
        assert lim <= bound + 1e-7 and float(ref[key]) <= bound + 1e-7
        assert lim > 0.999 * bound and float(ref[key]) > 0.999 * bound
    assert ref["constant_init"].shape == (7,) and not ref["constant_init"].any()
