"""Drop-in for the reference orchestrator `model.image2label` (model.py:169-1243): same constructor
(`image2label(sess, config)`, `sess` ignored), same `train()` / `evaluate()` entry points, same config
keys, checkpoint naming and stdout lines -- with the TF1 graph replaced by the B200 engine.

Mapped reference code:
  read_config            model.py:185-245   -> vnet_tensorflow_b200.config (tolerant)
  build_model_graph      model.py:297-630   -> VNetEngine construction + variable initialisation
  dataset_iterator       model.py:267-295   -> NiftiDataset3D.NiftiDataset(...).get_dataset() + shuffle(3) + batch
  train                  model.py:632-815   -> epoch / step loop, LogInterval checkpoints, TestStep test loss
  evaluate_single_3D     model.py:817-977   -> sliding windows, softmax accumulation, argmax, optional LCC / volume threshold
  evaluate               model.py:1131-1243 -> per-case loop, NIfTI outputs
TensorBoard image/metric summaries (model.py:314-334,449-463,570-626) are out of scope; the loss and learning-rate
scalars go to `LogDir/{train,test}/scalars.jsonl` and, under the reference's tags, to TensorBoard event files there
(events.py).
"""
from __future__ import annotations

import datetime
import json
import math
import os
import random
import shutil
import sys
from typing import List

import numpy as np

from . import checkpoint, config as config_mod, events, metrics, nifti
from .engine import VNetEngine
from .init import initialize
from .pipeline import NiftiDataset3D


def _now():
    return datetime.datetime.now()


def postprocess_label(label, spacing, lcc=False, volume_threshold=0.0):
    """model.py:1217-1223 with ExtractLargestConnectedComponents (:142-167) and volume_threshold (:117-140): face-
    connected components of the non-zero voxels; `lcc` keeps the one of largest physical size (first on ties),
    `volume_threshold` > 0 then keeps components whose physical size (mm^3) is strictly larger.  Both return a
    0/1 uint8 mask as the reference's filters do - class values do not survive them.  One deviation: a volume
    without foreground stays empty (the reference thresholds at label 0 there and returns all ones)."""
    if not lcc and not volume_threshold > 0:
        return label
    from scipy import ndimage
    voxel = float(np.prod(spacing))
    for which in ("lcc", "volume"):
        if (which == "lcc" and not lcc) or (which == "volume" and not volume_threshold > 0):
            continue
        comp, n = ndimage.label(np.asarray(label) != 0)
        if n == 0:
            return np.zeros(np.shape(label), np.uint8)
        sizes = ndimage.sum(np.ones(comp.shape, np.int64), comp, range(1, n + 1)) * voxel
        keep = np.zeros(n + 1, bool)
        if which == "lcc":
            keep[1 + int(np.argmax(sizes))] = True   # argmax = first maximum, as the strict `>` scan
        else:
            keep[1:] = sizes > volume_threshold
        label = keep[comp].astype(np.uint8)
    return label


class image2label(object):
    def __init__(self, sess, config, device: int = 0, library=None):
        self.sess = sess  # kept for signature compatibility (model.py:170); unused
        self.config = config
        self.device = device
        self.library = library
        self.engine = None
        self.epoches = 999999999999999999

    # ---- configuration ------------------------------------------------------------------------
    def read_config(self):
        print("{}: Reading configuration file...".format(_now()))
        self.cfg = config_mod.from_dict(self.config) if isinstance(self.config, dict) else self.config
        for k, v in vars(self.cfg).items():
            setattr(self, k, v)
        self.dimension = len(self.cfg.patch_shape)
        if self.dimension != 3:
            sys.exit("Only the 3-D V-Net path is accelerated (PatchShape must have 3 entries)")
        print("{}: Reading configuration file complete".format(_now()))

    def _transforms(self, yaml_path, phase):
        """model.py:341-402: instantiate NiftiDataset3D transforms by name from the pipeline YAML."""
        if not yaml_path or not os.path.exists(yaml_path):
            # the reference opens the file unconditionally (model.py:341-343): a missing pipeline is an error, not an
            # empty one -- training without normalisation / resampling / cropping would only fail later, or silently.
            # Synthetic runs have no files to transform and may leave it out.
            if self.cfg.synthetic:
                return []
            raise FileNotFoundError("pipeline YAML not found: {!r} (set TrainingSetting.Pipeline / "
                                    "EvaluationSetting.Pipeline, or Synthetic for a dry run)".format(yaml_path))
        import yaml
        with open(yaml_path) as f:
            spec = yaml.safe_load(f)
        out = []
        for t in (spec.get("preprocess", {}).get(phase, {}) or {}).get("3D", []) or []:
            cls = getattr(NiftiDataset3D, t["name"], None)
            if cls is None:   # the reference's getattr raises on an unknown class name (model.py:352-356)
                raise AttributeError("pipeline {}: unknown transform {!r} (module pipeline.NiftiDataset3D)"
                                     .format(yaml_path, t["name"]))
            out.append(cls(**(t.get("variables") or {})))
        return out

    def dataset_iterator(self, data_dir, transforms, train=True, pinned_ring=None):
        """model.py:267-295: dataset -> shuffle(buffer 3) -> batch(drop_remainder).  `pinned_ring` (optional): list of
        (images, labels) page-locked arrays of one batch each; batches are then collated straight into them, in turn,
        by the prefetch thread, so that the asynchronous host-to-device copy of the staged step needs no extra copy."""
        usable = (not self.cfg.synthetic) and os.path.isdir(data_dir) and self._has_nifti(data_dir)
        if usable:
            ds = NiftiDataset3D.NiftiDataset(data_dir=data_dir, image_filenames=self.image_filenames,
                                             label_filename=self.label_filename, transforms=transforms, train=train,
                                             labels=self.label_classes).get_dataset(num_parallel_calls=self.cfg.data_workers)
        else:
            print("{}: no readable NIfTI data in {} -- using synthetic patches".format(_now(), data_dir))
            ds = NiftiDataset3D.SyntheticDataset(self.patch_shape, self.input_channel_num, self.output_channel_num,
                                                 size=4 * self.batch_size, seed=0 if train else 10 ** 6).get_dataset()

        def batches():
            return NiftiDataset3D.prefetch_iter(_batches(), depth=2)

        def _batches():
            buf: List = []
            it = iter(ds)
            pending = []
            produced = 0
            while True:
                while len(buf) < 3:  # tf.data shuffle(buffer_size=3)
                    try:
                        buf.append(next(it))
                    except StopIteration:
                        break
                if not buf:
                    break
                pending.append(buf.pop(random.randrange(len(buf))))
                if len(pending) == self.batch_size:
                    if pinned_ring:
                        img_out, lab_out = pinned_ring[produced % len(pinned_ring)]
                        np.stack([p[0] for p in pending], 0, out=img_out)
                        np.stack([np.asarray(p[1]).reshape(lab_out.shape[1:]) for p in pending], 0, out=lab_out)
                        yield img_out, lab_out
                    else:
                        yield np.stack([p[0] for p in pending], 0), np.stack([p[1] for p in pending], 0)
                    produced += 1
                    pending = []
        return batches

    def _has_nifti(self, data_dir):
        for case in sorted(os.listdir(data_dir)):
            p = os.path.join(data_dir, case, self.image_filenames[0])
            if os.path.exists(p):
                try:
                    nifti.read(p)
                    return True
                except Exception:
                    return False
        return False

    def build_model_graph(self, max_batch=None):
        print("{}: Start to build model graph...".format(_now()))
        self.engine = VNetEngine(
            num_classes=self.output_channel_num, in_channels=self.input_channel_num, patch_shape=self.patch_shape,
            max_batch=max_batch or max(self.batch_size, self.evaluate_batch), num_channels=self.num_channel,
            num_levels=self.num_levels, num_convolutions=self.num_convolutions, bottom_convolutions=self.bottom_convolutions,
            precision=self.precision, loss=self.loss_name, loss_weights=self.loss_weights, loss_alpha=self.loss_alpha,
            optimizer=self.optimizer_name, learning_rate=self.initial_learning_rate, decay_factor=self.decay_factor,
            decay_steps=self.decay_steps, momentum=self.momentum, device=self.device, library=self.library)
        initialize(self.engine)  # tf.initializers.global_variables(), model.py:673
        print("{}: Build graph complete".format(_now()))

    # ---- training -----------------------------------------------------------------------------
    def _log(self, which, step, **scalars):
        d = os.path.join(self.log_dir, which)
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "scalars.jsonl"), "a") as f:
            # 0/0 metrics (a class absent from the batch) are NaN as in the reference's summaries; JSON has no NaN
            f.write(json.dumps(dict(step=step, **{k: (None if isinstance(v, float) and math.isnan(v) else v)
                                                  for k, v in scalars.items()})) + "\n")
        # the same scalars under the reference's summary tags, for `tensorboard --logdir LogDir` (model.py:562,644)
        writers = self.__dict__.setdefault("_event_writers", {})
        if which not in writers:
            writers[which] = events.EventFileWriter(d)
        tags = {"total_loss": "loss/0.total_loss", "learning_rate": "learning_rate"}
        writers[which].add_scalars(step, {tags.get(k, k): v for k, v in scalars.items()})

    def _step_metrics(self, n):
        """The tf.metrics block of summary_op (model.py:586-626), fetched with every training / test step
        (model.py:743-748,784-789): accuracy and, per foreground class, sensitivity / specificity / dice / auc under
        the reference's summary tags.  The counts come from the device (vnb_read_metrics)."""
        cm, hist = self.engine.metric_counts(n)
        scalars = metrics.step_metrics(cm, hist, self.label_classes)
        keep = ("accuracy", "sensitivity_", "specificity_", "dice_", "auc_")
        out = {"metrics/" + k: float(v) for k, v in scalars.items() if k.startswith(keep)}
        if self.loss_name.startswith("mixed"):      # model.py:529-530: the two summands of a mixed loss
            out["loss/1.dice"], out["loss/2.regularized_xent"] = self.engine.loss_parts()
        return out

    def _learning_rate(self, step):
        """tf.train.exponential_decay(lr0, global_step, decay_steps, decay_factor, staircase=False), model.py:641-643."""
        return self.initial_learning_rate * self.decay_factor ** (step / self.decay_steps)

    def train(self):
        print("{}: VNet Tensorflow training start...".format(_now()))
        self.read_config()
        self.build_model_graph()
        # page-locked batch buffers for the staged input copy: 2 queued by the prefetch thread + 1 collated and waiting
        # for a queue slot + 1 in flight to the device + 1 whose step is running, + 1 spare
        ring = [(self.engine.pinned_array((self.batch_size,) + tuple(self.patch_shape) + (self.input_channel_num,), np.float32),
                 self.engine.pinned_array((self.batch_size,) + tuple(self.patch_shape), np.int32)) for _ in range(6)]
        train_batches = self.dataset_iterator(self.train_data_dir, self._transforms(self.training_pipeline, "train"), True,
                                              pinned_ring=ring)
        test_batches = self.dataset_iterator(self.test_data_dir, self._transforms(self.training_pipeline, "test"), True) if self.testing else None
        start_epoch = 0
        print("{}: Start training...".format(_now()))
        if not self.restore_training:  # model.py:679-688: wipe log / checkpoint dirs
            for d in (self.log_dir, self.ckpt_dir):
                if os.path.exists(d):
                    shutil.rmtree(d)
                os.makedirs(d)
        else:
            latest = checkpoint.latest(self.ckpt_dir)
            if latest is not None:
                print("{}: Last checkpoint found at {}, loading...".format(_now(), self.ckpt_dir))
                _, start_epoch = checkpoint.restore(self.engine, latest)
            print("{}: Last checkpoint epoch: {}".format(_now(), start_epoch))
            print("{}: Last checkpoint global step: {}".format(_now(), self.engine.global_step))
        test_iter = iter(test_batches()) if test_batches else None
        for epoch in range(start_epoch, self.epoches):
            print("{}: Epoch {} starts...".format(_now(), epoch + 1))
            loss_sum, count = 0.0, 0
            # model.py:736-748 with the feed_dict copy taken off the critical path: the next batch is staged (copied
            # to the device on a copy stream) while the step on the current one runs; same arithmetic, same order
            batch_iter = iter(train_batches())
            staged = next(batch_iter, None)
            if staged is not None:
                self.engine.stage_batch(*staged)
            while staged is not None:
                image = staged[0]
                if self.engine.global_step > self.max_itr:
                    sys.exit("{}: Reach maximum iteration steps, training abort.".format(_now()))
                self.engine.train_step_staged(self.dropout_rate, seed=self.engine.global_step, want_loss=False)
                staged = next(batch_iter, None)
                if staged is not None:
                    self.engine.stage_batch(*staged)
                loss = self.engine.last_loss()
                print('{}: Segmentation training loss: {}'.format(_now(), str(loss)))
                loss_sum += loss
                count += 1
                step = self.engine.global_step
                self._log("train", step, total_loss=loss, learning_rate=self._learning_rate(step - 1),
                          **self._step_metrics(len(image)))
                if step % self.log_interval == 0:
                    print("{}: Saving checkpoint of step {} at {}...".format(_now(), step, self.ckpt_dir))
                    checkpoint.save(self.engine, self.ckpt_dir, step, epoch, self.cfg.checkpoint_format)
                if self.testing and step % self.test_step == 0:
                    try:
                        timg, tlab = next(test_iter)
                    except StopIteration:
                        test_iter = iter(test_batches())
                        timg, tlab = next(test_iter)
                    tloss = self.engine.loss(timg, tlab)  # dropout 0, batch statistics, no update (model.py:784-789)
                    print('{}: Segmentation testing loss: {}'.format(_now(), str(tloss)))
                    self._log("test", step, total_loss=tloss, **self._step_metrics(len(timg)))
            print("{}: Training of epoch {} complete, epoch loss: {}".format(_now(), epoch + 1, loss_sum / max(count, 1)))
            print("{}: Saving checkpoint of epoch {} at {}...".format(_now(), epoch + 1, self.ckpt_dir))
            checkpoint.save(self.engine, self.ckpt_dir, self.engine.global_step, epoch + 1, self.cfg.checkpoint_format)
            print("{}: Saving checkpoint succeed".format(_now()))
        for w in self.__dict__.pop("_event_writers", {}).values():  # model.py:812-815
            w.close()

    # ---- evaluation ---------------------------------------------------------------------------
    def evaluate_single_3D(self, images_np: np.ndarray):
        """model.py:866-937 on an [X,Y,Z,M] array: returns (label int64 [X,Y,Z], softmax sums [X,Y,Z,K], weight)."""
        P, S = self.patch_shape, self.evaluate_stride
        dims = images_np.shape[:3]
        pads = [(0, max(p - d, 0)) for d, p in zip(dims, P)] + [(0, 0)]
        if any(p[1] for p in pads):
            images_np = np.pad(images_np, pads)
        # window loop, softmax accumulation and final argmax run on the device (vnb_evaluate_volume)
        label_np, softmax_np, weight_np = self.engine.evaluate_volume(images_np, S, self.evaluate_batch)
        crop = tuple(slice(0, d) for d in dims)
        return label_np[crop], softmax_np[crop], weight_np[crop]

    def evaluate(self):
        self.read_config()
        self.build_model_graph(max_batch=self.evaluate_batch)
        checkpoint.restore(self.engine, self.checkpoint_path)  # model.py:1138-1139
        transforms = self._transforms(self.evaluate_pipeline, "evaluate")
        print("{}: Start evaluation...".format(_now()))
        for case in sorted(os.listdir(self.evaluate_data_dir)):
            case_dir = os.path.join(self.evaluate_data_dir, case)
            if not os.path.isdir(case_dir):
                continue
            print("{}: Evaluating {}...".format(_now(), case))
            images = [nifti.read(os.path.join(case_dir, ch)) for ch in self.evaluate_image_filenames]
            sample = {'image': images, 'label': nifti.Image(np.zeros(images[0].GetSize(), np.int32), images[0].spacing, images[0].origin)}
            for t in transforms:
                sample = t(sample)
            arr = np.stack([np.asarray(im.array, np.float32) for im in sample['image']], -1)
            label_np, softmax_np, weight_np = self.evaluate_single_3D(arr)
            # back to the input image's grid (model.py:957-975): nearest for the label, linear for the probabilities
            ref, orig = sample['image'][0], images[0]
            same_grid = (ref.GetSize() == orig.GetSize() and np.allclose(ref.spacing, orig.spacing) and np.allclose(ref.origin, orig.origin))

            def back(arr, order):
                im = nifti.Image(arr, ref.spacing, ref.origin, ref.direction)
                if not same_grid:
                    im = NiftiDataset3D.resample_image(im, orig.spacing, orig.GetSize(), orig.origin, order)
                return nifti.Image(im.array, orig.spacing, orig.origin, orig.direction)
            if not same_grid:
                print("{}: Resampling label back to original image space...".format(_now()))
            # the reference writes the argmax class *index* (model.py:934,945); `MapLabelValues` writes the
            # SegmentationClasses value of that index instead
            out = label_np.astype(np.int32)
            if self.cfg.evaluate_map_label_values:
                out = np.asarray(self.label_classes, np.int32)[label_np]
            out = back(out, 0).array
            out = postprocess_label(out, orig.spacing, self.evaluate_lcc, self.evaluate_volume_threshold)
            nifti.write(os.path.join(case_dir, self.evaluate_label_filename), nifti.Image(out, orig.spacing, orig.origin, orig.direction))
            if self.evaluate_probability_output:  # model.py:935-937,1234-1243
                prob = softmax_np / np.maximum(weight_np[..., None], 1.0)
                for c, value in enumerate(self.label_classes):
                    name = self.evaluate_probability_filename.replace(".nii", "_%s.nii" % value, 1)
                    nifti.write(os.path.join(case_dir, name), back(prob[..., c].astype(np.float32), 1))
        print("{}: Evaluation complete".format(_now()))
