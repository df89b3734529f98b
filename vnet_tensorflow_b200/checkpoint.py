"""Checkpoint save / restore with the reference's conventions (model.py:676-709,758-764,803-809).

The reference writes TF bundle files `CheckpointDir/checkpoint-<global_step>.{meta,index,data-*}` and a
`checkpoint-latest` state file through tf.train.Saver.  The native format here is one `.npz` per checkpoint
under the *same names and variable keys* (`vnet/encoder/level_1/conv_1/weights`, ..., Adam slots
`<var>/Adam`, `<var>/Adam_1`, `global_step`, `start_epoch`) with the `checkpoint-latest` pointer +
`CheckpointPath` semantics kept.  TF bundles themselves are read and written through tf_bundle.py
(`format="tf"` / `"both"`, `TrainingSetting.CheckpointFormat`): `restore` takes either kind of prefix, so a
run of the reference can be resumed or evaluated here, and `export_tf` writes the `.index` / `.data` pair
the reference's Saver restores (its `.meta` graph file stays TensorFlow's to write).
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from . import _ffi, tf_bundle

LATEST = "checkpoint-latest"
FORMATS = ("npz", "tf", "both")

# The optimizer is built under tf.name_scope("training") (model.py:646): slot variables come from get_variable
# under the primary's own scope (`<var>/Adam`), the two Adam accumulators are plain tf.Variables and take the scope.
OPTIMIZER_SCOPE = "training/"
_BETA1, _BETA2 = 0.9, 0.999  # tf.train.AdamOptimizer defaults, model.py:652


def _optimizer_name(engine) -> str:
    code = int(engine.cfg.optimizer)
    return next(k for k, v in _ffi.OPTIMIZERS.items() if v == code)


def _npz_arrays(engine, global_step: int, start_epoch: int):
    arrays = {}
    for name, (_, trainable) in engine.variables().items():
        arrays[name] = engine.get_param(name)
        if trainable:
            arrays[name + "/Adam"] = engine.get_param(name, _ffi.SLOT_ADAM_M)
            arrays[name + "/Adam_1"] = engine.get_param(name, _ffi.SLOT_ADAM_V)
    arrays["global_step"] = np.asarray(global_step, np.int64)
    arrays["start_epoch"] = np.asarray([start_epoch], np.int32)
    return arrays


def tf_variables(engine, global_step: int, start_epoch: int = 0):
    """Every variable tf.train.Saver() would save for the reference's training graph, by TF name: the model
    variables, the optimizer's slots (tf.train.{Adam,Momentum}Optimizer), `global_step` (int64 scalar,
    model.py:299) and `start_epoch` (int32 [1], model.py:668)."""
    opt = _optimizer_name(engine)
    out = {}
    for name, (_, trainable) in engine.variables().items():
        out[name] = engine.get_param(name)
        if not trainable:
            continue
        if opt == "Adam":
            out[name + "/Adam"] = engine.get_param(name, _ffi.SLOT_ADAM_M)
            out[name + "/Adam_1"] = engine.get_param(name, _ffi.SLOT_ADAM_V)
        elif opt in ("Momentum", "NesterovMomentum"):
            out[name + "/Momentum"] = engine.get_param(name, _ffi.SLOT_ADAM_M)
    if opt == "Adam":  # beta^t accumulators: initialised to beta, multiplied by beta after every step
        out[OPTIMIZER_SCOPE + "beta1_power"] = np.asarray(_BETA1 ** (int(global_step) + 1), np.float32)
        out[OPTIMIZER_SCOPE + "beta2_power"] = np.asarray(_BETA2 ** (int(global_step) + 1), np.float32)
    out["global_step"] = np.asarray(global_step, np.int64)
    out["start_epoch"] = np.asarray([start_epoch], np.int32)
    return out


def export_tf(engine, prefix: str, global_step: int, start_epoch: int = 0) -> str:
    """Write `<prefix>.index` + `<prefix>.data-00000-of-00001` for the reference's Saver (model.py:696-699)."""
    tf_bundle.write_bundle(prefix, tf_variables(engine, global_step, start_epoch))
    return prefix


def import_tf(engine, prefix: str):
    """Load a checkpoint written by the reference (model.py:758-764,803-809). Returns (global_step, start_epoch).

    Model variables must all be present, as Saver.restore demands; optimizer slots, `global_step` and
    `start_epoch` are taken when present (a checkpoint of another optimizer still evaluates)."""
    reader = tf_bundle.BundleReader(prefix)
    missing = [n for n in engine.variables() if n not in reader]
    if missing:
        raise KeyError("Key %s not found in checkpoint %s (%d of %d variables missing)"
                       % (missing[0], prefix, len(missing), len(engine.variables())))
    opt = _optimizer_name(engine)
    for name, (shape, trainable) in engine.variables().items():
        value = reader.get_tensor(name)
        if tuple(value.shape) != tuple(shape):
            raise ValueError("%s: checkpoint shape %s, graph shape %s" % (name, value.shape, shape))
        engine.set_param(name, value)
        if not trainable:
            continue
        if opt == "Adam" and name + "/Adam" in reader and name + "/Adam_1" in reader:
            engine.set_param(name, reader.get_tensor(name + "/Adam"), _ffi.SLOT_ADAM_M)
            engine.set_param(name, reader.get_tensor(name + "/Adam_1"), _ffi.SLOT_ADAM_V)
        elif opt in ("Momentum", "NesterovMomentum") and name + "/Momentum" in reader:
            engine.set_param(name, reader.get_tensor(name + "/Momentum"), _ffi.SLOT_ADAM_M)
    step = int(reader.get_tensor("global_step")) if "global_step" in reader else 0
    epoch = int(reader.get_tensor("start_epoch").reshape(-1)[0]) if "start_epoch" in reader else 0
    engine.global_step = step
    return step, epoch


MAX_TO_KEEP = 5                 # tf.train.Saver default max_to_keep (model.py:676 passes none)
KEEP_EVERY_N_HOURS = 5.0        # model.py:676: tf.train.Saver(keep_checkpoint_every_n_hours=5)
_SUFFIXES = (".npz", ".index", ".data-00000-of-00001", ".meta")


def _checkpoint_files(prefix: str):
    return [prefix + s for s in _SUFFIXES if os.path.exists(prefix + s)]


def _read_state(ckpt_dir: str):
    """(latest prefix basename, recent basenames oldest first, timestamp of the last checkpoint kept for good)."""
    latest_name, recent, kept_at = None, [], None
    p = os.path.join(ckpt_dir, LATEST)
    if os.path.exists(p):
        with open(p) as f:
            for line in f:
                key, _, val = line.partition(":")
                val = val.strip().strip('"')
                if key == "model_checkpoint_path":
                    latest_name = val
                elif key == "all_model_checkpoint_paths":
                    recent.append(val)
                elif key == "last_preserved_timestamp":
                    kept_at = float(val)
    return latest_name, recent, kept_at


def save(engine, ckpt_dir: str, global_step: int, start_epoch: int = 0, format: str = "npz",
         max_to_keep: int = MAX_TO_KEEP, keep_every_n_hours: float = KEEP_EVERY_N_HOURS) -> str:
    """One checkpoint per call, with tf.train.Saver's retention (model.py:676,696-699): the last `max_to_keep` are kept,
    older ones are deleted unless `keep_every_n_hours` have passed since the last one that was kept for good.  Files are
    written under a temporary name and renamed, so a crash never leaves a truncated checkpoint behind the pointer."""
    import time
    if format not in FORMATS:
        raise ValueError("CheckpointFormat must be one of %s" % (FORMATS,))
    os.makedirs(ckpt_dir, exist_ok=True)
    name = "checkpoint-%d" % global_step
    prefix = os.path.join(ckpt_dir, name)
    if format in ("npz", "both"):
        tmp = prefix + ".tmp.npz"
        np.savez(tmp, **_npz_arrays(engine, global_step, start_epoch))
        os.replace(tmp, prefix + ".npz")
    if format in ("tf", "both"):
        tmp_prefix = prefix + ".tmp"
        export_tf(engine, tmp_prefix, global_step, start_epoch)
        for suffix in (".data-00000-of-00001", ".index"):   # data first: the index is what marks the bundle complete
            os.replace(tmp_prefix + suffix, prefix + suffix)
    _, recent, kept_at = _read_state(ckpt_dir)
    now = time.time()
    if kept_at is None:
        kept_at = now
    recent = [r for r in recent if r != name] + [name]
    while max_to_keep and len(recent) > max_to_keep:
        old = recent.pop(0)
        files = _checkpoint_files(os.path.join(ckpt_dir, old))
        stamp = max([os.path.getmtime(f) for f in files], default=now)
        if keep_every_n_hours and stamp - kept_at >= keep_every_n_hours * 3600.0:
            kept_at = stamp          # this one stays on disk, like Saver's keep_checkpoint_every_n_hours
            continue
        for f in files:
            os.remove(f)
    tmp = os.path.join(ckpt_dir, LATEST + ".tmp")
    with open(tmp, "w") as f:  # same role and keys as tf's CheckpointState file
        f.write('model_checkpoint_path: "%s"\n' % name)
        for r in recent:
            f.write('all_model_checkpoint_paths: "%s"\n' % r)
        f.write("last_preserved_timestamp: %r\n" % kept_at)
    os.replace(tmp, os.path.join(ckpt_dir, LATEST))
    return prefix


def latest(ckpt_dir: str) -> Optional[str]:
    p = os.path.join(ckpt_dir, LATEST)
    if not os.path.exists(p):
        return None
    with open(p) as f:
        line = f.readline()
    return os.path.join(ckpt_dir, line.split('"')[1])


def restore(engine, prefix: str):
    """prefix: path without extension, as in EvaluationSetting.CheckpointPath. Returns (global_step, start_epoch).
    Takes the native `.npz` when it exists, else a TensorFlow bundle of that prefix."""
    path = prefix if prefix.endswith(".npz") else prefix + ".npz"
    if not os.path.exists(path):
        for cand in (prefix, prefix[:-len(".index")] if prefix.endswith(".index") else None):
            if cand and tf_bundle.is_bundle(cand):
                return import_tf(engine, cand)
    with np.load(path) as z:
        for name, (_, trainable) in engine.variables().items():
            engine.set_param(name, z[name])
            if trainable and name + "/Adam" in z:
                engine.set_param(name, z[name + "/Adam"], _ffi.SLOT_ADAM_M)
                engine.set_param(name, z[name + "/Adam_1"], _ffi.SLOT_ADAM_V)
        step = int(z["global_step"])
        epoch = int(z["start_epoch"][0]) if "start_epoch" in z else 0
    engine.global_step = step
    return step, epoch


def export_binary(engine, path: str) -> str:
    """Flat binary dump of the model variables for the native driver (cxx/vnb_infer.cpp), the role of the
    reference's meta_to_pb.py (frozen `nWeights` assigns consumed by cxx/tf_inference.cpp:110-144):
    "VNBW" u32 version=1 u32 count, then per variable u32 name length, name (TF variable name), u32 ndim,
    i64 dims[ndim], float32 data."""
    import struct
    names = list(engine.variables().items())
    with open(path, "wb") as f:
        f.write(b"VNBW" + struct.pack("<II", 1, len(names)))
        for name, _ in names:
            a = np.ascontiguousarray(engine.get_param(name), np.float32)
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)) + nb + struct.pack("<I", a.ndim) + struct.pack("<%dq" % a.ndim, *a.shape))
            f.write(a.tobytes())
    return path
