"""Python host object over one libvnet_b200 handle (one handle <-> one GPU <-> one host thread).

`VNetEngine` owns the opaque C handle and exposes the steps the reference performs with
`sess.run(...)`: inference (model.py:914-917), loss-only test step (model.py:784-789) and the
training step (model.py:743-748).  All arrays cross as C-contiguous NumPy buffers in the reference's
layouts (NDHWC images, int32 label indices).
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np

from . import _ffi


class _DevBuf:
    """A torch CUDA tensor seen as a raw device pointer (keeps the tensor alive for the duration of the call)."""

    def __init__(self, t):
        self.t = t
        self.shape = tuple(t.shape)
        self.ndim = t.ndim
        self.ptr = C.c_void_p(t.data_ptr())


def _ptr(a):
    if a is None:
        return None
    return a.ptr if isinstance(a, _DevBuf) else a.ctypes.data_as(C.c_void_p)


class VNetEngine:
    def __init__(self, *, num_classes: int, in_channels: int = 1, patch_shape: Sequence[int] = (64, 64, 64),
                 max_batch: int = 1, num_channels: int = 16, num_levels: int = 4,
                 num_convolutions: Sequence[int] = (1, 2, 3, 3), bottom_convolutions: int = 3,
                 precision: str = "fp32", loss: str = "weighted_sorensen", loss_weights: Sequence[float] = (),
                 loss_alpha: float = 1.0, optimizer: str = "Adam", learning_rate: float = 1e-2,
                 decay_factor: float = 0.99, decay_steps: float = 100.0, momentum: float = 0.9, flavour: str = "networks", device: int = 0,
                 attention: bool = False, attention_loss: Optional[str] = None, module_channels: int = 64,
                 library: Optional[_ffi.Library] = None):
        self.lib = library or _ffi.default_library()
        if precision not in _ffi.PRECISIONS:
            raise ValueError("Precision must be one of %s" % sorted(_ffi.PRECISIONS))
        if loss not in _ffi.LOSSES:
            raise SystemExit("Invalid loss function")  # model.py:559-560
        if optimizer not in _ffi.OPTIMIZERS:
            raise SystemExit("Invalid optimizer")  # model.py:657-658
        if len(num_convolutions) != num_levels:
            raise AssertionError("num_levels == len(num_convolutions)")  # networks.py:228
        if len(patch_shape) != 3:
            raise ValueError("only 3-D PatchShape is accelerated (the 2-D branches are out of scope)")
        cfg = _ffi.VnbConfig()
        cfg.in_channels, cfg.num_classes = in_channels, num_classes
        cfg.num_channels, cfg.num_levels = num_channels, num_levels
        for i, n in enumerate(num_convolutions):
            cfg.num_convolutions[i] = int(n)
        cfg.bottom_convolutions = bottom_convolutions
        for i in range(3):
            cfg.patch_shape[i] = int(patch_shape[i])
        cfg.max_batch = max_batch
        cfg.precision = _ffi.PRECISIONS[precision]
        cfg.loss = _ffi.LOSSES[loss]
        w = list(loss_weights) if len(loss_weights) else [1.0] * num_classes
        if "weighted" in loss and len(w) != num_classes:
            raise AssertionError("Length of DICE weight is {}, should be {}".format(len(w), num_classes))  # model.py:71
        for i in range(8):
            cfg.loss_weights[i] = float(w[i]) if i < len(w) else 1.0
        cfg.loss_alpha = loss_alpha
        cfg.optimizer = _ffi.OPTIMIZERS[optimizer]
        cfg.learning_rate, cfg.decay_factor, cfg.decay_steps = learning_rate, decay_factor, decay_steps
        cfg.momentum = momentum
        if flavour not in ("networks", "legacy"):
            raise ValueError("flavour must be 'networks' (networks.VNet) or 'legacy' (VNet.py)")
        cfg.graph_flavour = 1 if flavour == "legacy" else 0
        if attention_loss not in _ffi.ATTENTION_LOSSES:
            raise SystemExit("Invalid loss function")  # train.py:399-400
        cfg.attention = 1 if attention else 0
        cfg.attention_loss = _ffi.ATTENTION_LOSSES[attention_loss]
        cfg.module_channels = int(module_channels)
        self.attention = bool(attention)
        self.cfg = cfg
        self.num_classes, self.in_channels = num_classes, in_channels
        self.patch_shape = tuple(int(p) for p in patch_shape)
        self.max_batch = max_batch
        self.precision = precision
        self.device = int(device)
        self._h = C.c_void_p()
        self.lib.check(self.lib.vnb_create(C.byref(cfg), device, C.byref(self._h)))
        self._specs = self._read_specs()

    # ---- lifetime -----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.vnb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- variables ----------------------------------------------------------------------------
    def _read_specs(self):
        n = C.c_int()
        self.lib.check(self.lib.vnb_num_params(self._h, C.byref(n)))
        specs = OrderedDict()
        for i in range(n.value):
            name, ndim, tr = C.c_char_p(), C.c_int(), C.c_int()
            dims = (C.c_int64 * 5)()
            self.lib.check(self.lib.vnb_param_info(self._h, i, C.byref(name), C.byref(ndim), dims, C.byref(tr)))
            specs[name.value.decode()] = (tuple(int(dims[k]) for k in range(ndim.value)), bool(tr.value))
        return specs

    def variables(self) -> "OrderedDict[str, Tuple[Tuple[int, ...], bool]]":
        """TF variable name -> (shape, trainable), in creation order."""
        return self._specs

    def set_param(self, name: str, value: np.ndarray, slot: int = _ffi.SLOT_VALUE):
        shape, _ = self._specs[name]
        a = np.ascontiguousarray(value, dtype=np.float32)
        if tuple(a.shape) != shape:
            raise ValueError("%s: expected shape %s, got %s" % (name, shape, a.shape))
        self.lib.check(self.lib.vnb_set_slot(self._h, name.encode(), slot, _ptr(a), a.nbytes))

    def get_param(self, name: str, slot: int = _ffi.SLOT_VALUE) -> np.ndarray:
        shape, _ = self._specs[name]
        a = np.empty(shape, np.float32)
        self.lib.check(self.lib.vnb_get_slot(self._h, name.encode(), slot, _ptr(a), a.nbytes))
        return a

    def set_params(self, params: Dict[str, np.ndarray]):
        for k, v in params.items():
            self.set_param(k, v)

    def get_params(self, names: Optional[Iterable[str]] = None) -> "OrderedDict[str, np.ndarray]":
        return OrderedDict((k, self.get_param(k)) for k in (names or self._specs))

    def get_grads(self) -> "OrderedDict[str, np.ndarray]":
        return OrderedDict((k, self.get_param(k, _ffi.SLOT_GRAD)) for k, (_, tr) in self._specs.items() if tr)

    @property
    def global_step(self) -> int:
        s = C.c_int64()
        self.lib.check(self.lib.vnb_get_step(self._h, C.byref(s)))
        return s.value

    @global_step.setter
    def global_step(self, v: int):
        self.lib.check(self.lib.vnb_set_step(self._h, int(v)))

    # ---- attention path (train.py:281-312,383-418) ---------------------------------------------
    def set_distmap(self, distmap: np.ndarray):
        """Distance map in [0,1] fed to the attention loss, [N,X,Y,Z] (or [N,X,Y,Z,1]) float32."""
        if self._is_device_tensor(distmap):
            a = self._device_tensor(distmap[..., 0] if distmap.ndim == 5 and distmap.shape[-1] == 1 else distmap,
                                    "float32", "distmap")
        else:
            a = np.ascontiguousarray(distmap, dtype=np.float32)
            if a.ndim == 5 and a.shape[-1] == 1:
                a = np.ascontiguousarray(a[..., 0])
        if a.ndim != 4 or tuple(a.shape[1:]) != self.patch_shape:
            raise ValueError("distmap must be [N,%d,%d,%d] float32, got %s" % (self.patch_shape + (tuple(a.shape),)))
        self.lib.check(self.lib.vnb_set_distmap(self._h, _ptr(a), a.shape[0]))

    def losses(self) -> Tuple[float, float, float]:
        """(total, segmentation, attention) loss of the last loss / training call."""
        out = (C.c_float * 3)()
        self.lib.check(self.lib.vnb_read_losses(self._h, out))
        return float(out[0]), float(out[1]), float(out[2])

    def loss_parts(self) -> Tuple[float, float]:
        """('1.dice' = 1 - dice, '2.regularized_xent' = Loss.Alpha * cross entropy) of the last loss / training call,
        the two scalars the reference logs next to the total of a mixed loss (model.py:529-530)."""
        out = (C.c_float * 2)()
        self.lib.check(self.lib.vnb_read_loss_parts(self._h, out))
        return float(out[0]), float(out[1])

    def softmax_attention(self, n: int) -> np.ndarray:
        a = np.empty((n,) + self.patch_shape + (self.num_classes,), np.float32)
        self.lib.check(self.lib.vnb_read_softmax_attention(self._h, _ptr(a), a.nbytes, n))
        return a

    # ---- steps --------------------------------------------------------------------------------
    @staticmethod
    def _is_device_tensor(x) -> bool:
        return type(x).__module__.split(".")[0] == "torch" and bool(getattr(x, "is_cuda", False))

    def _device_tensor(self, t, dtype_name: str, what: str):
        """A torch CUDA tensor handed straight to the C ABI (no host bounce): right dtype, contiguous, on the handle's
        GPU; the producing stream is synchronised here because the engine copies on its own stream."""
        import torch
        if t.dtype != getattr(torch, dtype_name) or not t.is_contiguous():
            raise ValueError("%s on the device must be a contiguous %s tensor" % (what, dtype_name))
        if t.device.index != self.device:
            raise ValueError("%s lives on cuda:%s, the engine on cuda:%d" % (what, t.device.index, self.device))
        torch.cuda.current_stream(t.device).synchronize()
        return _DevBuf(t)

    def _check_images(self, images):
        dev = self._is_device_tensor(images)
        a = self._device_tensor(images, "float32", "images") if dev else np.ascontiguousarray(images, dtype=np.float32)
        want = self.patch_shape + (self.in_channels,)
        if a.ndim != 5 or tuple(a.shape[1:]) != want:
            raise ValueError("images must be [N,%d,%d,%d,%d] float32, got %s" % (want + (tuple(a.shape),)))
        if not 1 <= a.shape[0] <= self.max_batch:
            raise ValueError("batch %d outside [1, %d]" % (a.shape[0], self.max_batch))
        return a

    def _check_labels(self, labels, n: int):
        if self._is_device_tensor(labels):
            if labels.ndim == 5 and labels.shape[-1] == 1:
                labels = labels[..., 0]      # a view: still contiguous
            l = self._device_tensor(labels, "int32", "labels")
        else:
            l = np.ascontiguousarray(labels, dtype=np.int32)
            if l.ndim == 5 and l.shape[-1] == 1:  # model.py:741 feeds [...,np.newaxis]
                l = np.ascontiguousarray(l[..., 0])
        if tuple(l.shape) != (n,) + self.patch_shape:
            raise ValueError("labels must be [N,X,Y,Z] int32 class indices, got %s" % (tuple(l.shape),))
        return l

    def forward(self, images, want_logits=True, want_softmax=True, want_argmax=True):
        """sess.run(['predicted_label/prediction:0','softmax:0']) of model.py:914-917 (+ logits)."""
        a = self._check_images(images)
        n = a.shape[0]
        shp = (n,) + self.patch_shape
        logits = np.empty(shp + (self.num_classes,), np.float32) if want_logits else None
        softmax = np.empty(shp + (self.num_classes,), np.float32) if want_softmax else None
        argmax = np.empty(shp, np.int64) if want_argmax else None
        self.lib.check(self.lib.vnb_forward(self._h, _ptr(a), n, _ptr(logits), _ptr(softmax), _ptr(argmax)))
        return logits, softmax, argmax

    def evaluate_volume(self, volume, stride, batch=1, want_sums=True, want_weight=True):
        """Window loop of model.py:866-937 on the device: `volume` [X,Y,Z,M] (every extent >= the patch extent).
        Returns (label int64 [X,Y,Z], softmax sums [X,Y,Z,K] or None, weight [X,Y,Z] or None)."""
        v = np.ascontiguousarray(volume, dtype=np.float32)
        if v.ndim != 4 or v.shape[3] != self.in_channels:
            raise ValueError("volume must be [X,Y,Z,%d], got %r" % (self.in_channels, v.shape))
        dims = (C.c_int32 * 3)(*v.shape[:3])
        st = (C.c_int32 * 3)(*[int(x) for x in stride])
        label = np.empty(v.shape[:3], np.int64)
        sums = np.empty(v.shape[:3] + (self.num_classes,), np.float32) if want_sums else None
        weight = np.empty(v.shape[:3], np.float32) if want_weight else None
        self.lib.check(self.lib.vnb_evaluate_volume(self._h, _ptr(v), dims, st, int(batch), _ptr(label), _ptr(sums), _ptr(weight)))
        return label, sums, weight

    def loss(self, images, labels, want_terms=False):
        a = self._check_images(images)
        l = self._check_labels(labels, a.shape[0])
        out = C.c_float()
        terms = np.empty((a.shape[0], self.num_classes, 4), np.float64) if want_terms else None
        self.lib.check(self.lib.vnb_loss(self._h, _ptr(a), _ptr(l), a.shape[0], C.byref(out), _ptr(terms)))
        return (out.value, terms) if want_terms else out.value

    def train_step(self, images, labels, dropout_rate=0.0, seed=0, want_loss=True):
        a = self._check_images(images)
        l = self._check_labels(labels, a.shape[0])
        out = C.c_float()
        self.lib.check(self.lib.vnb_train_step(self._h, _ptr(a), _ptr(l), a.shape[0], float(dropout_rate),
                                               int(seed), C.byref(out) if want_loss else None))
        return out.value if want_loss else None

    def forward_backward(self, images, labels, dropout_rate=0.0, seed=0, update_moving_stats=False, want_loss=True):
        a = self._check_images(images)
        l = self._check_labels(labels, a.shape[0])
        out = C.c_float()
        self.lib.check(self.lib.vnb_forward_backward(self._h, _ptr(a), _ptr(l), a.shape[0], float(dropout_rate),
                                                     int(seed), int(bool(update_moving_stats)),
                                                     C.byref(out) if want_loss else None))
        return out.value if want_loss else None

    def metric_counts(self, n: int, want_auc: bool = True):
        """Integer counts behind the tf.metrics block of model.py:586-626 for the last batch that came with labels:
        (confusion uint64 [K+1, K] (label row, argmax column; row K = labels outside [0, K)),
         auc_hist uint64 [K, 2, 201] or None).  `metrics.step_metrics` turns them into the reference's scalars."""
        cm = np.zeros((self.num_classes + 1, self.num_classes), np.uint64)
        hist = np.zeros((self.num_classes, 2, _ffi.AUC_BINS), np.uint64) if want_auc else None
        self.lib.check(self.lib.vnb_read_metrics(self._h, int(n), _ptr(cm), _ptr(hist)))
        return cm, hist

    def apply_gradients(self):
        self.lib.check(self.lib.vnb_apply_gradients(self._h))

    def upload_batch(self, images, labels):
        a = self._check_images(images)
        l = self._check_labels(labels, a.shape[0])
        self.lib.check(self.lib.vnb_upload_batch(self._h, _ptr(a), _ptr(l), a.shape[0]))
        return a.shape[0]

    def stage_batch(self, images, labels):
        """Enqueue the copy of the NEXT batch on the copy stream (overlaps the running step when the arrays are
        page-locked, see pinned_array); pair with train_step_staged."""
        a = self._check_images(images)
        l = self._check_labels(labels, a.shape[0])
        self.lib.check(self.lib.vnb_stage_batch(self._h, _ptr(a), _ptr(l), a.shape[0]))
        self._staged = (a, l)        # keep the host buffers alive until the staged step has consumed them
        return a.shape[0]

    def train_step_staged(self, dropout_rate=0.0, seed=0, want_loss=True):
        out = C.c_float()
        self.lib.check(self.lib.vnb_train_step_staged(self._h, float(dropout_rate), int(seed),
                                                      C.byref(out) if want_loss else None))
        return out.value if want_loss else None

    def last_loss(self) -> float:
        """Loss of the last step issued with want_loss=False (synchronises)."""
        return self.losses()[0]

    def pinned_array(self, shape, dtype) -> np.ndarray:
        """NumPy array over page-locked host memory (cudaMallocHost), freed with the array."""
        dt = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dt.itemsize
        p = C.c_void_p()
        self.lib.check(self.lib.vnb_host_alloc(self._h, nbytes, C.byref(p)))
        lib = self.lib

        class _Owner:
            def __init__(self, ptr):
                self.ptr = ptr
                self.buf = (C.c_char * max(nbytes, 1)).from_address(ptr.value)

            def __del__(self):
                lib.vnb_host_free(self.ptr)

        owner = _Owner(p)
        arr = np.frombuffer(owner.buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
        self.__dict__.setdefault("_pinned_owners", []).append(owner)   # freed with the engine object, not before its views
        return arr

    def train_step_resident(self, n, dropout_rate=0.0, seed=0):
        self.lib.check(self.lib.vnb_train_step_resident(self._h, int(n), float(dropout_rate), int(seed)))

    def event_record(self, which: int):
        self.lib.check(self.lib.vnb_event_record(self._h, which))

    def event_elapsed_ms(self) -> float:
        ms = C.c_float()
        self.lib.check(self.lib.vnb_event_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    def profile_enable(self, on: bool):
        self.lib.check(self.lib.vnb_profile_enable(self._h, int(on)))

    def profile_read(self, kernel_class: int):
        ms, n, fl = C.c_double(), C.c_int64(), C.c_double()
        self.lib.check(self.lib.vnb_profile_read(self._h, kernel_class, C.byref(ms), C.byref(n), C.byref(fl)))
        return ms.value, n.value, fl.value

    def profile_launches(self):
        """Every profiled launch since profile_enable(True): (label "scope pass Cin->Cout @DxHxW", class, ms, FLOPs)."""
        n = C.c_int64()
        self.lib.check(self.lib.vnb_profile_count(self._h, C.byref(n)))
        out = []
        buf = C.create_string_buffer(256)
        for i in range(n.value):
            cls, ms, fl = C.c_int(), C.c_double(), C.c_double()
            self.lib.check(self.lib.vnb_profile_launch(self._h, i, C.byref(cls), C.byref(ms), C.byref(fl), buf, len(buf)))
            out.append((buf.value.decode(), cls.value, ms.value, fl.value))
        return out

    def sync(self):
        self.lib.check(self.lib.vnb_sync(self._h))

    def gpu_launches(self) -> int:
        c = C.c_int64()
        self.lib.check(self.lib.vnb_gpu_launches(self._h, C.byref(c)))
        return c.value

    def read_tensor(self, scope: str, kind: int, n: int, channels: int, spatial: Sequence[int]) -> np.ndarray:
        a = np.empty((n,) + tuple(spatial) + (channels,), np.float32)
        self.lib.check(self.lib.vnb_read_tensor(self._h, scope.encode(), kind, _ptr(a), a.nbytes, n))
        return a

    # ---- data parallel ------------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self.lib.check(self.lib.vnb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, rank: int, world: int, unique_id: bytes):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self.lib.check(self.lib.vnb_comm_init(self._h, rank, world, buf))

    def comm_sync_bn(self, on: bool = True):
        """Synchronised batch norm over the engine's NCCL communicator (collective: every rank calls it)."""
        self.lib.check(self.lib.vnb_comm_sync_bn(self._h, 1 if on else 0))

    def set_stats_allreduce(self, fn, world: int):
        """Synchronised batch norm with the exchange done by the caller: `fn(values)` sums a float64 NumPy row in
        place over `world` ranks (None switches back to local statistics)."""
        if fn is None:
            self._stats_cb = None
            self.lib.check(self.lib.vnb_set_stats_allreduce(self._h, None, None, 1))
            return

        def trampoline(ptr, n, _user):
            try:
                fn(np.ctypeslib.as_array(ptr, shape=(n,)))
                return 0
            except Exception:  # never unwind through the C frames
                import traceback
                traceback.print_exc()
                return 1
        self._stats_cb = _ffi.ALLREDUCE_FN(trampoline)  # keep the thunk alive as long as the engine uses it
        self.lib.check(self.lib.vnb_set_stats_allreduce(self._h, C.cast(self._stats_cb, C.c_void_p), None, int(world)))
