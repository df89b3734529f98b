"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN CODE (run in the build container, where
/root/reference exists; the fixtures it writes travel with the repo).

`/root/reference/networks.py` and `layers2.py` are imported unmodified and `dice_coe` is cut out of
`/root/reference/model.py` (lines 26-85; model.py itself cannot be imported: SimpleITK).  They run over
`tests/tf1_shim.py`, an eager stand-in for the ~25 TensorFlow-1 symbols they touch -- see its docstring for what
this pins (the graph the reference builds: op order, scopes, the x + BN(x) quirk, dead batch norms, variable names and
creation order, and through torch autograd the gradients of exactly that graph) and what it cannot (TensorFlow's
own kernels).  Nothing is copied from the reference into the repo: only tensors it computed.

Per case `tests/golden/ref_<name>.npz` holds: variable names in the reference's creation order, logits (float64),
softmax-Dice loss through the reference's dice_coe, a digest of the gradient of every trainable variable (L2 norm,
sum, 512 strided entries) and the UPDATE_OPS moving statistics.  Inputs and parameters are regenerated from seeds by the test.
Run:  python tests/golden/make_reference_fixtures.py
"""
import ast
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_vnet as R  # noqa: E402  (only for the seeded parameter / patch generators)
from tests import tf1_shim  # noqa: E402
from tests.golden.make_golden import perturbed_params  # noqa: E402
from vnet_tensorflow_b200.synthetic import synth_batch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("VNB_REFERENCE_DIR", "/root/reference")

CASES = {
    # name: (VNet kwargs, in_channels, P, N, dice loss_type, Loss.Weights or ()); names starting with "legacy_" run
    # VNet.py / Layers.py (the graph train.py:271-279 builds), the others networks.py / layers2.py (main.py path)
    "m1_k2_c12": (dict(num_classes=2, num_channels=16, num_levels=2, num_convolutions=(1, 2), bottom_convolutions=2),
                  1, 8, 2, "sorensen", (0.1, 1.0)),
    "m2_k3_c31": (dict(num_classes=3, num_channels=16, num_levels=2, num_convolutions=(3, 1), bottom_convolutions=1),
                  2, 8, 1, "jaccard", (0.01, 0.1, 1.0)),
    "m1_k2_c123": (dict(num_classes=2, num_channels=16, num_levels=3, num_convolutions=(1, 2, 3), bottom_convolutions=3),
                   1, 8, 2, "sorensen", ()),
    "legacy_m1_k2_c12": (dict(num_classes=2, num_channels=16, num_levels=2, num_convolutions=(1, 2), bottom_convolutions=2),
                         1, 8, 2, "jaccard", ()),
    "legacy_m2_k3_c21": (dict(num_classes=3, num_channels=16, num_levels=2, num_convolutions=(2, 1), bottom_convolutions=3),
                         2, 8, 1, "sorensen", (0.01, 0.1, 1.0)),
}


def grad_digest(g):
    """Compact fingerprint of one gradient tensor: L2 norm, sum (float64) and up to 512 evenly strided entries
    (float32) -- keeps the fixtures small without weakening what a wiring error would show."""
    a = g.detach().numpy().astype(np.float64).reshape(-1)
    step = max(1, a.size // 512)
    return np.float64(np.sqrt((a ** 2).sum())), np.float64(a.sum()), a[::step][:512].astype(np.float32)


def reference_function(tf, name, module="model.py"):
    """A top-level function exactly as written in the reference file, compiled on its own (the files themselves
    cannot be imported: SimpleITK / NiftiDataset imports at module level)."""
    src = open(os.path.join(REFERENCE, module)).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name][0]
    ns = {"tf": tf}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), os.path.join(REFERENCE, module), "exec"), ns)
    return ns[name]


def reference_dice_coe(tf, module="model.py"):
    """`dice_coe` of model.py:26-85 (or train.py:100-149)."""
    return reference_function(tf, "dice_coe", module)


def run_reference_losses(dtype=torch.float64):
    """model.py:26-92 on random inputs: dice_coe in all four forms and weighted_softmax_cross_entropy_with_logits."""
    rng = np.random.default_rng(5)
    K, shape = 3, (2, 5, 4, 3)
    logits = rng.normal(0, 2, shape + (K,))
    labels = rng.integers(0, K, shape)
    weights = [0.01, 0.1, 1.0]
    tf1_shim.uninstall()
    tf = tf1_shim.install({}, dtype)
    sm = tf.nn.softmax(tf1_shim.T(torch.from_numpy(logits).to(dtype)))
    oh = tf.one_hot(tf1_shim.T(torch.from_numpy(labels)), K)
    dice = reference_dice_coe(tf)
    out = {"logits": logits, "labels": labels, "weights": np.asarray(weights)}
    for kind in ("sorensen", "jaccard"):
        out["dice_" + kind] = np.float64(dice(sm, oh, loss_type=kind).v)
        out["dice_weighted_" + kind] = np.float64(dice(sm, oh, loss_type=kind, weights=list(weights)).v)
    # layers2.py:4-30: the reference's own initialisers (NumPy global RNG, seeded here only to make the file stable)
    sys.path.insert(0, REFERENCE)
    try:
        layers2 = importlib.import_module("layers2")
        np.random.seed(0)
        out["xavier_5x5x5x16x32"] = layers2.xavier_initializer_convolution([5, 5, 5, 16, 32]).astype(np.float32)[::7, 0, 0, 0, 0]
        out["xavier_5x5x5x16x32_absmax"] = np.float64(np.abs(layers2.xavier_initializer_convolution([5, 5, 5, 16, 32])).max())
        out["xavier_2x2x2x32x16_absmax"] = np.float64(np.abs(layers2.xavier_initializer_convolution([2, 2, 2, 32, 16])).max())
        out["constant_init"] = layers2.constant_initializer(0, shape=7)
    finally:
        sys.path.remove(REFERENCE)
    wx = reference_function(tf, "weighted_softmax_cross_entropy_with_logits")
    out["weighted_xent"] = np.float64(wx(oh, tf1_shim.T(torch.from_numpy(logits).to(dtype)), weights).v)
    tf1_shim.uninstall()
    return out


def run_reference(kw, in_channels, P, N, loss_type, weights, dtype=torch.float64, legacy=False):
    spec = R.VNetSpec(in_channels=in_channels, flavour="legacy" if legacy else "networks", **kw)
    params = perturbed_params(spec)
    img, lab = synth_batch(0, N, P, in_channels, kw["num_classes"])
    tf1_shim.uninstall()
    tf = tf1_shim.install(params, dtype)
    sys.path.insert(0, REFERENCE)
    try:
        if legacy:
            vnet_py = importlib.import_module("VNet")
            net = vnet_py.VNet(keep_prob=1.0, activation_fn="prelu", **kw)                        # train.py:271-278
            logits = net.network_fn(tf1_shim.T(torch.from_numpy(img).to(dtype)))               # train.py:279
        else:
            networks = importlib.import_module("networks")
            net = networks.VNet(dropout_rate=0.0, is_training=True, activation_fn="prelu", **kw)   # model.py:428-438
            logits = net.GetNetwork(tf1_shim.T(torch.from_numpy(img).to(dtype)))               # model.py:444
        softmax = tf.nn.softmax(logits)                                                       # model.py:447
        onehot = tf.one_hot(tf1_shim.T(torch.from_numpy(lab)), kw["num_classes"])           # model.py:477
        dice = reference_dice_coe(tf)(softmax, onehot, loss_type=loss_type, weights=list(weights))   # model.py:499-522
        loss = 1.0 - dice
    finally:
        sys.path.remove(REFERENCE)
    st = tf.STATE
    names = list(st.created)
    train = [n for n in names if st.trainable[n]]
    grads = torch.autograd.grad(loss.v, [st.created[n] for n in train], allow_unused=True)
    out = {"variable_names": np.array(names), "trainable": np.array([st.trainable[n] for n in names]),
           "logits": logits.v.detach().numpy(), "loss": np.float64(loss.v.detach())}
    for n, g in zip(train, grads):
        out["gnorm/" + n], out["gsum/" + n], out["gsample/" + n] = grad_digest(torch.zeros_like(st.created[n]) if g is None else g)
    for n, u in st.bn_updates.items():
        out["moving/" + n] = u.numpy().astype(np.float32)
    tf1_shim.uninstall()
    return spec, params, img, lab, out


ATTENTION_CASE = dict(num_channels=16, num_levels=2, num_convolutions=(1, 2), bottom_convolutions=1, P=8, N=1)


def run_reference_attention(dtype=torch.float64):
    """The attention-gated network of train.py:269-312.  `VNet.py`, `attention.py` and `OutputModule.py` run as they
    are; the few lines of train.py that compose them sit inside its `train()` function (train.py cannot be imported:
    top-level `import NiftiDataset3D`), so the composition below restates train.py:281-312 (modules under
    tf.name_scope("attention") / ("output"), `logits_masked = (1 + softmax_attention) * logits_vnet`),
    train.py:538-540 (both modules are fed train_phase False), train.py:378-382 (`--loss_function jaccard` through
    train.py's own dice_coe, lines 100-149) and train.py:387-393,415 (l2 attention loss, total = att + seg)."""
    c = ATTENTION_CASE
    kw = dict(num_classes=2, num_channels=c["num_channels"], num_levels=c["num_levels"], num_convolutions=c["num_convolutions"],
              bottom_convolutions=c["bottom_convolutions"])
    spec = R.VNetSpec(in_channels=1, flavour="legacy", **kw)
    params = R.init_attention_params(spec)
    from vnet_tensorflow_b200.synthetic import synth_patch
    im, lb, dm = synth_patch(77, c["P"], 1, 2)
    img, lab, dist = im[None], lb[None], dm[None]
    tf1_shim.uninstall()
    tf = tf1_shim.install(params, dtype)
    sys.path.insert(0, REFERENCE)
    try:
        vnet_py = importlib.import_module("VNet")
        attention = importlib.import_module("attention")
        output_module = importlib.import_module("OutputModule")
        net = vnet_py.VNet(keep_prob=1.0, activation_fn="prelu", **kw)
        logits_vnet = net.network_fn(tf1_shim.T(torch.from_numpy(img).to(dtype)))
        with tf.name_scope("attention"):
            att = attention.AttentionModule(num_classes=2, is_training=True, activation_fn="relu", keep_prob=1.0)
            att.train_phase.value = False
            logits_attention = att.GetNetwork(logits_vnet)
            softmax_attention = tf.nn.softmax(logits_attention)
        with tf.name_scope("masked_vnet"):
            logits_masked = (1 + softmax_attention) * logits_vnet
        with tf.name_scope("output"):
            om = output_module.OutputModule(num_classes=2, is_training=True, activation_fn="relu", keep_prob=1.0)
            om.train_phase.value = False
            logits_output = om.GetNetwork(logits_masked)
        softmax_op = tf.nn.softmax(logits_output)
        onehot = tf.cast(tf.one_hot(tf1_shim.T(torch.from_numpy(lab)), 2), tf.float32)
        jaccard = reference_dice_coe(tf, "train.py")(softmax_op, onehot, loss_type="jaccard", axis=[1, 2, 3])
        loss_op = 1.0 - jaccard
        d1 = torch.from_numpy(dist).to(dtype)
        att_loss_op = tf1_shim.T((torch.square(softmax_attention.v[..., 1] - d1) * 100).mean())
        total = att_loss_op + loss_op
    finally:
        sys.path.remove(REFERENCE)
    st = tf.STATE
    names = list(st.created)
    train = [n for n in names if st.trainable[n]]
    grads = torch.autograd.grad(total.v, [st.created[n] for n in train], allow_unused=True)
    out = {"variable_names": np.array(names), "logits_vnet": logits_vnet.v.detach().numpy(),
           "softmax_attention": softmax_attention.v.detach().numpy(), "logits_masked": logits_masked.v.detach().numpy(),
           "logits_output": logits_output.v.detach().numpy(), "total_loss": np.float64(total.v.detach()),
           "seg_loss": np.float64(loss_op.v.detach()), "att_loss": np.float64(att_loss_op.v.detach())}
    for n, g in zip(train, grads):
        out["gnorm/" + n], out["gsum/" + n], out["gsample/" + n] = grad_digest(torch.zeros_like(st.created[n]) if g is None else g)
    tf1_shim.uninstall()
    return spec, params, img, lab, dist, out


def main():
    np.savez_compressed(os.path.join(HERE, "ref_losses.npz"), **run_reference_losses())
    spec, params, img, lab, dist, out = run_reference_attention()
    np.savez_compressed(os.path.join(HERE, "ref_attention_k2.npz"), **out)
    print("attention_k2 variables", len(out["variable_names"]), "total loss", float(out["total_loss"]))
    for name, (kw, M, P, N, loss_type, weights) in CASES.items():
        spec, params, img, lab, out = run_reference(kw, M, P, N, loss_type, weights, legacy=name.startswith("legacy_"))
        np.savez_compressed(os.path.join(HERE, "ref_" + name + ".npz"), **out)
        print(name, "variables", len(out["variable_names"]), "loss", float(out["loss"]), "logits", out["logits"].shape)


if __name__ == "__main__":
    main()
