"""ctypes binding of libvnet_b200.so (include/vnet_b200.h) -- the only bridge between the Python host
code and the CUDA engine.  There is no CPU fallback: if the shared library is missing or no B200 is
visible, loading / `vnb_create` raise, they never degrade to another implementation.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libvnet_b200.so")

PRECISIONS = {"fp32": 0, "bf16x3": 1, "bf16": 2}
AUC_BINS = 201   # VNB_AUC_BINS: tf.metrics.auc has 200 thresholds, a probability lies above 0..200 of them
LOSSES = {
    "xent": 0, "weighted_xent": 1, "sorensen": 2, "weighted_sorensen": 3, "jaccard": 4,
    "weighted_jaccard": 5, "mixed_sorensen": 6, "mixed_weighted_sorensen": 7, "mixed_jaccard": 8,
    "mixed_weighted_jaccard": 9,
    "sorensen_fg": 10,  # legacy train.py --loss_function sorensen (foreground channel only)
}
ATTENTION_LOSSES = {None: 0, "none": 0, "l2": 1, "abs": 2}
OPTIMIZERS = {"Adam": 0, "SGD": 1, "Momentum": 2, "NesterovMomentum": 3}
SLOT_VALUE, SLOT_GRAD, SLOT_ADAM_M, SLOT_ADAM_V = 0, 1, 2, 3
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_int, C.c_void_p)  # vnb_allreduce_fn


class VnbConfig(C.Structure):
    _fields_ = [
        ("in_channels", C.c_int32), ("num_classes", C.c_int32), ("num_channels", C.c_int32),
        ("num_levels", C.c_int32), ("num_convolutions", C.c_int32 * 8), ("bottom_convolutions", C.c_int32),
        ("patch_shape", C.c_int32 * 3), ("max_batch", C.c_int32), ("precision", C.c_int32),
        ("loss", C.c_int32), ("loss_weights", C.c_float * 8), ("loss_alpha", C.c_float),
        ("optimizer", C.c_int32), ("learning_rate", C.c_float), ("decay_factor", C.c_float),
        ("decay_steps", C.c_float), ("momentum", C.c_float), ("graph_flavour", C.c_int32),
        ("attention", C.c_int32), ("attention_loss", C.c_int32), ("module_channels", C.c_int32),
    ]


class VnbError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("libvnet_b200 error %d: %s" % (code, message))
        self.code = code


# every exported symbol of include/vnet_b200.h with its prototype (tests check the .so exports all)
_PROTOTYPES = {
    "vnb_last_error": (C.c_char_p, []),
    "vnb_version": (C.c_char_p, []),
    "vnb_create": (C.c_int, [C.POINTER(VnbConfig), C.c_int, C.POINTER(C.c_void_p)]),
    "vnb_destroy": (C.c_int, [C.c_void_p]),
    "vnb_num_params": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "vnb_param_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int),
                                 C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "vnb_set_param": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]),
    "vnb_get_param": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]),
    "vnb_set_slot": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_size_t]),
    "vnb_get_slot": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_size_t]),
    "vnb_get_step": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "vnb_set_step": (C.c_int, [C.c_void_p, C.c_int64]),
    "vnb_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vnb_evaluate_volume": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "vnb_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_void_p]),
    "vnb_train_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_uint64,
                                 C.POINTER(C.c_float)]),
    "vnb_forward_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_uint64,
                                       C.c_int, C.POINTER(C.c_float)]),
    "vnb_apply_gradients": (C.c_int, [C.c_void_p]),
    "vnb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "vnb_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "vnb_comm_world": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vnb_comm_sync_bn": (C.c_int, [C.c_void_p, C.c_int]),
    "vnb_set_stats_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "vnb_upload_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "vnb_train_step_resident": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_uint64]),
    "vnb_event_record": (C.c_int, [C.c_void_p, C.c_int]),
    "vnb_event_elapsed_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "vnb_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "vnb_profile_read": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "vnb_stage_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "vnb_train_step_staged": (C.c_int, [C.c_void_p, C.c_float, C.c_uint64, C.POINTER(C.c_float)]),
    "vnb_host_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "vnb_host_free": (C.c_int, [C.c_void_p]),
    "vnb_read_loss_parts": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "vnb_read_metrics": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "vnb_profile_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "vnb_profile_launch": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.c_char_p, C.c_size_t]),
    "vnb_set_distmap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "vnb_read_losses": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "vnb_read_softmax_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "vnb_sync": (C.c_int, [C.c_void_p]),
    "vnb_gpu_launches": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "vnb_read_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int]),
    "vnb_op_conv5_fprop": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p] + [C.c_int] * 6),
    "vnb_op_conv5_dgrad": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6),
    "vnb_op_conv5_wgrad": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6),
    "vnb_op_conv3_fprop": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p] + [C.c_int] * 6),
    "vnb_op_conv3_dgrad": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6),
    "vnb_op_conv3_wgrad": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6),
    "vnb_op_k2": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6),
    "vnb_op_bn_fwd": (C.c_int, [C.c_int] + [C.c_void_p] * 7 + [C.c_longlong, C.c_int]),
    "vnb_op_bn_bwd": (C.c_int, [C.c_int] + [C.c_void_p] * 9 + [C.c_longlong, C.c_int]),
    "vnb_op_softmax_dice_fwd": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_void_p,
                                          C.c_float, C.POINTER(C.c_float), C.c_void_p, C.c_void_p, C.c_void_p]),
    "vnb_op_softmax_dice_bwd": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_void_p,
                                          C.c_float, C.c_void_p]),
    "vnb_op_adam": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_longlong]),
}
EXPORTED_SYMBOLS = tuple(_PROTOTYPES)


class Library:
    """A loaded libvnet_b200 shared object with typed entry points."""

    def __init__(self, path: Optional[str] = None):
        self.path = path or os.environ.get("VNB_LIBRARY", DEFAULT_LIB)
        if not os.path.exists(self.path):
            raise FileNotFoundError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % self.path)
        self.cdll = C.CDLL(self.path)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(self.cdll, name)  # AttributeError here = the .so does not match the header
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def check(self, rc: int):
        if rc != 0:
            raise VnbError(rc, (self.vnb_last_error() or b"").decode("utf-8", "replace"))

    def version(self) -> str:
        return self.vnb_version().decode()


_default: Optional[Library] = None


def default_library() -> Library:
    global _default
    if _default is None:
        _default = Library()
    return _default
