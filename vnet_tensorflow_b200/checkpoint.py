"""Checkpoint save / restore with the reference's conventions (model.py:676-709,758-764,803-809).

The reference writes TF bundle files `CheckpointDir/checkpoint-<global_step>.{meta,index,data-*}` and a
`checkpoint-latest` state file through tf.train.Saver.  The TF bundle format cannot be produced without
TensorFlow, so this module stores one `.npz` per checkpoint under the *same names and variable keys*
(`vnet/encoder/level_1/conv_1/weights`, ..., Adam slots `<var>/Adam`, `<var>/Adam_1`, `global_step`,
`start_epoch`) and keeps the `checkpoint-latest` pointer + `CheckpointPath` semantics.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from . import _ffi

LATEST = "checkpoint-latest"


def save(engine, ckpt_dir: str, global_step: int, start_epoch: int = 0) -> str:
    os.makedirs(ckpt_dir, exist_ok=True)
    prefix = os.path.join(ckpt_dir, "checkpoint-%d" % global_step)
    arrays = {}
    for name, (_, trainable) in engine.variables().items():
        arrays[name] = engine.get_param(name)
        if trainable:
            arrays[name + "/Adam"] = engine.get_param(name, _ffi.SLOT_ADAM_M)
            arrays[name + "/Adam_1"] = engine.get_param(name, _ffi.SLOT_ADAM_V)
    arrays["global_step"] = np.asarray(global_step, np.int64)
    arrays["start_epoch"] = np.asarray([start_epoch], np.int32)
    np.savez(prefix + ".npz", **arrays)
    with open(os.path.join(ckpt_dir, LATEST), "w") as f:  # same role as tf's latest_filename
        f.write('model_checkpoint_path: "%s"\n' % os.path.basename(prefix))
    return prefix


def latest(ckpt_dir: str) -> Optional[str]:
    p = os.path.join(ckpt_dir, LATEST)
    if not os.path.exists(p):
        return None
    with open(p) as f:
        line = f.readline()
    return os.path.join(ckpt_dir, line.split('"')[1])


def restore(engine, prefix: str):
    """prefix: path without extension, as in EvaluationSetting.CheckpointPath. Returns (global_step, start_epoch)."""
    path = prefix if prefix.endswith(".npz") else prefix + ".npz"
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
