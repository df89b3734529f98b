"""Patch interface of the reference's data layer (pipeline/NiftiDataset3D.py), without SimpleITK/tf.data.

Kept: the class / constructor names that the YAML pipelines resolve by `getattr(NiftiDataset3D, name)
(**variables)` (model.py:341-402), the sample dict {'image': [img per modality], 'label': img}, and the
output contract of `NiftiDataset.get_dataset()`: (image float32 [X,Y,Z,M], label int32 [X,Y,Z]) with
labels remapped to class *indices* (NiftiDataset3D.py:119-137,150-165).  Images are vnet_tensorflow_b200.
nifti.Image objects (NumPy array[x,y,z] + spacing/origin).  The resampling transforms use trilinear /
nearest interpolation from SciPy instead of ITK's B-spline (SURVEY.md "next" row N1); accelerating this
CPU stage is out of the hot path's scope.
"""
from __future__ import annotations

import os
import random
from typing import List, Optional, Sequence

import numpy as np

from .. import nifti


class NiftiDataset(object):
    """NiftiDataset3D.py:10-165.  get_dataset() returns a re-iterable over (image, label) patches."""

    def __init__(self, data_dir='', image_filenames='', label_filename='', transforms=None, train=False, labels=[0, 1]):
        self.data_dir = data_dir
        self.image_filenames = image_filenames
        self.label_filename = label_filename
        self.transforms = transforms
        self.train = train
        self.labels = labels

    def read_image(self, path):
        return nifti.read(path)

    def case_dirs(self) -> List[str]:
        return [os.path.join(self.data_dir, c) for c in sorted(os.listdir(self.data_dir))]

    def input_parser(self, case_dir):
        images = [self.read_image(os.path.join(case_dir, ch)) for ch in self.image_filenames]
        for im in images[1:]:  # NiftiDataset3D.py:66-82: all modalities must share the grid
            if im.GetSize() != images[0].GetSize():
                raise Exception("Header info inconsistent: {}".format(case_dir))
        if self.train:
            label = self.read_image(os.path.join(case_dir, self.label_filename))
        else:  # NiftiDataset3D.py:104-111: empty label
            label = nifti.Image(np.zeros(images[0].GetSize(), np.int32), images[0].spacing, images[0].origin)
        sample = {'image': images, 'label': label}
        if self.transforms:
            for transform in self.transforms:
                sample = transform(sample)
        label_np = np.asarray(sample['label'].array)
        remapped = np.zeros(label_np.shape, np.int32)  # NiftiDataset3D.py:119-137: label value -> class index
        for idx, value in enumerate(self.labels):
            remapped[label_np == value] = idx
        image_np = np.stack([np.asarray(im.array, np.float32) for im in sample['image']], axis=-1)
        return image_np.astype(np.float32), remapped

    def get_dataset(self, num_parallel_calls=1, prefetch=None):
        """NiftiDataset3D.py:39-55.  The reference maps `input_parser` through tf.py_func with
        num_parallel_calls=1 (one SimpleITK pipeline at a time); here `num_parallel_calls` cases are parsed
        concurrently (NumPy / SciPy release the GIL in the resampling and filtering kernels) and handed out in the
        original order, at most `prefetch` cases ahead of the consumer."""
        ds = self

        class _Iterable:
            def __iter__(self_inner):
                return parallel_map(ds.input_parser, ds.case_dirs(), num_parallel_calls, prefetch)

            def __len__(self_inner):
                return len(ds.case_dirs())

        self.dataset = _Iterable()
        self.data_size = len(self.case_dirs())
        return self.dataset


def parallel_map(fn, items, workers=1, depth=None):
    """Order-preserving map over `items` on a thread pool with a bounded look-ahead window."""
    workers = max(1, int(workers))
    if workers == 1:
        for item in items:
            yield fn(item)
        return
    from collections import deque
    from concurrent.futures import ThreadPoolExecutor
    depth = max(workers, int(depth)) if depth else 2 * workers
    pending = deque()
    ex = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="vnb-data")
    try:
        for item in items:
            pending.append(ex.submit(fn, item))
            if len(pending) >= depth:
                yield pending.popleft().result()
        while pending:
            yield pending.popleft().result()
    finally:
        for f in pending:
            f.cancel()
        ex.shutdown(wait=True)


def prefetch_iter(iterable, depth=2):
    """Runs `iterable` on a background thread, `depth` items ahead: batch assembly and file I/O of the next step
    overlap the (GIL-free) engine call of the current one."""
    import queue
    import threading
    q = queue.Queue(maxsize=max(1, int(depth)))
    end, stop = object(), threading.Event()

    def work():
        try:
            for item in iterable:
                while not stop.is_set():
                    try:
                        q.put(item, timeout=0.1)
                        break
                    except queue.Full:
                        continue
                if stop.is_set():
                    return
            q.put(end)
        except BaseException as e:  # re-raised in the consumer
            q.put(e)

    t = threading.Thread(target=work, name="vnb-prefetch", daemon=True)
    t.start()
    try:
        while True:
            item = q.get()
            if item is end:
                return
            if isinstance(item, BaseException):
                raise item
            yield item
    finally:
        stop.set()


class SyntheticDataset(object):
    """Same output contract, generated patches (vnet_tensorflow_b200.synthetic) -- used when the data
    directories hold no readable NIfTI files (the reference's own data/ are git-LFS stubs)."""

    def __init__(self, patch_shape, modalities, classes, size=8, seed=0):
        self.patch_shape, self.modalities, self.classes, self.size, self.seed = tuple(patch_shape), modalities, classes, size, seed

    def get_dataset(self):
        from ..synthetic import synth_patch
        ds = self

        class _Iterable:
            def __iter__(self_inner):
                for i in range(ds.size):
                    im, lb, _ = synth_patch(ds.seed + i, min(ds.patch_shape), ds.modalities, ds.classes, ds.patch_shape)
                    yield im, lb

            def __len__(self_inner):
                return ds.size

        return _Iterable()


# ---- transforms (callables sample -> sample), constructor signatures as in the reference -------------
def _map(sample, fn_img, fn_lbl=None):
    out = {'image': [fn_img(im) for im in sample['image']], 'label': sample['label']}
    if fn_lbl is not None:
        out['label'] = fn_lbl(sample['label'])
    return out


class StatisticalNormalization(object):
    """NiftiDataset3D.py:210-254: clamp to mean +- sigma*std, rescale to 0..255."""

    def __init__(self, sigma, pre_norm=False):
        self.name = 'StatisticalNormalization'
        self.sigma, self.pre_norm = sigma, pre_norm

    def __call__(self, sample):
        def norm(im):
            a = np.asarray(im.array, np.float32)
            if self.pre_norm:
                a = (a - a.mean()) / max(a.std(), 1e-6)
            lo, hi = a.mean() - self.sigma * a.std(), a.mean() + self.sigma * a.std()
            a = (np.clip(a, lo, hi) - lo) / max(hi - lo, 1e-6) * 255.0
            return nifti.Image(a.astype(np.float32), im.spacing, im.origin, im.direction)
        return _map(sample, norm)


class ManualNormalization(object):
    """NiftiDataset3D.py:285-308: window [windowMin, windowMax] -> 0..255."""

    def __init__(self, windowMin, windowMax):
        self.name = 'ManualNormalization'
        self.windowMin, self.windowMax = float(windowMin), float(windowMax)

    def __call__(self, sample):
        def norm(im):
            a = (np.clip(np.asarray(im.array, np.float32), self.windowMin, self.windowMax) - self.windowMin) \
                / max(self.windowMax - self.windowMin, 1e-6) * 255.0
            return nifti.Image(a.astype(np.float32), im.spacing, im.origin, im.direction)
        return _map(sample, norm)


class Normalization(StatisticalNormalization):
    """NiftiDataset3D.py:167-185 (0..255 rescale of the full range)."""

    def __init__(self):
        self.name = 'Normalization'

    def __call__(self, sample):
        def norm(im):
            a = np.asarray(im.array, np.float32)
            a = (a - a.min()) / max(a.max() - a.min(), 1e-6) * 255.0
            return nifti.Image(a, im.spacing, im.origin, im.direction)
        return _map(sample, norm)


class Resample(object):
    """NiftiDataset3D.py:345-398: resample to `voxel_size` (linear for images, nearest for labels)."""

    def __init__(self, voxel_size):
        self.name = 'Resample'
        self.voxel_size = (voxel_size,) * 3 if isinstance(voxel_size, (int, float)) else tuple(voxel_size)

    def _res(self, im, order):
        from scipy import ndimage
        zoom = [s / v for s, v in zip(im.spacing, self.voxel_size)]
        a = ndimage.zoom(np.asarray(im.array), zoom, order=order, mode="nearest")
        return nifti.Image(a.astype(im.array.dtype), tuple(float(v) for v in self.voxel_size), im.origin, im.direction)

    def __call__(self, sample):
        return _map(sample, lambda im: self._res(im, 1), lambda lb: self._res(lb, 0))


class Padding(object):
    """NiftiDataset3D.py:400-456: zero-pad symmetric up to at least `output_size`."""

    def __init__(self, output_size):
        self.name = 'Padding'
        self.output_size = (output_size,) * 3 if isinstance(output_size, int) else tuple(output_size)

    def _pad(self, im):
        a = np.asarray(im.array)
        pads = [(max(o - s, 0) // 2, max(o - s, 0) - max(o - s, 0) // 2) for s, o in zip(a.shape, self.output_size)]
        return nifti.Image(np.pad(a, pads), im.spacing, im.origin, im.direction)

    def __call__(self, sample):
        return _map(sample, self._pad, self._pad)


class RandomCrop(object):
    """NiftiDataset3D.py:458-551: random crop of `output_size`; with probability 1-drop_ratio the crop
    must contain at least `min_pixel` foreground voxels."""

    def __init__(self, output_size, drop_ratio=0.1, min_pixel=1):
        self.name = 'Random Crop'
        self.output_size = (output_size,) * 3 if isinstance(output_size, int) else tuple(output_size)
        self.drop_ratio, self.min_pixel = drop_ratio, min_pixel

    def _starts(self, shape, centre=None, jitter=None):
        out = []
        for i, (s, o) in enumerate(zip(shape, self.output_size)):
            hi = s - o
            if centre is None:
                out.append(random.randint(0, hi) if hi > 0 else 0)
            else:
                c = int(centre[i] + random.uniform(-jitter, jitter)) - o // 2
                out.append(min(max(c, 0), hi) if hi > 0 else 0)
        return out

    def _crop(self, im, st):
        sl = tuple(slice(s, s + o) for s, o in zip(st, self.output_size))
        return nifti.Image(np.asarray(im.array)[sl], im.spacing, im.origin, im.direction)

    def __call__(self, sample):
        lab = np.asarray(sample['label'].array)
        st = self._starts(lab.shape)
        for _ in range(50):
            st = self._starts(lab.shape)
            sl = tuple(slice(s, s + o) for s, o in zip(st, self.output_size))
            if (lab[sl] > 0).sum() >= self.min_pixel or random.random() < self.drop_ratio:
                break
        return _map(sample, lambda im: self._crop(im, st), lambda lb: self._crop(lb, st))


class ConfidenceCrop2(RandomCrop):
    """NiftiDataset3D.py:661-793: crop centred on a random connected foreground component with
    probability `probability`, jittered by `rand_range`; random crop otherwise."""

    def __init__(self, output_size, rand_range=3, probability=0.5, random_empty_region=False):
        super().__init__(output_size)
        self.name = 'Confidence Crop 2'
        self.rand_range, self.probability = rand_range, probability

    def __call__(self, sample):
        lab = np.asarray(sample['label'].array)
        st = self._starts(lab.shape)
        if random.random() <= self.probability and (lab > 0).any():
            from scipy import ndimage
            comp, n = ndimage.label(lab > 0)
            k = random.randint(1, n)
            centre = [float(c) for c in ndimage.center_of_mass(comp == k)]
            st = self._starts(lab.shape, centre, self.rand_range)
        return _map(sample, lambda im: self._crop(im, st), lambda lb: self._crop(lb, st))


class RandomNoise(object):
    """NiftiDataset3D.py:553-572: additive Gaussian noise on the images."""

    def __init__(self, std=0.1):
        self.name = 'Random Noise'
        self.std = std

    def __call__(self, sample):
        def noise(im):
            a = np.asarray(im.array, np.float32)
            return nifti.Image(a + np.random.normal(0, self.std, a.shape).astype(np.float32), im.spacing, im.origin, im.direction)
        return _map(sample, noise)


class RandomFlip(object):
    """NiftiDataset3D.py:187-208."""

    def __init__(self, axes=[False, False, False]):
        self.name = 'Flip'
        self.axes = axes

    def __call__(self, sample):
        flips = [ax for ax, on in enumerate(self.axes) if on and random.random() > 0.5]
        f = lambda im: nifti.Image(np.flip(np.asarray(im.array), flips).copy() if flips else im.array, im.spacing, im.origin, im.direction)
        return _map(sample, f, f)


class Invert(object):
    """NiftiDataset3D.py:330-343."""

    def __init__(self):
        self.name = 'Invert'

    def __call__(self, sample):
        return _map(sample, lambda im: nifti.Image(255.0 - np.asarray(im.array, np.float32), im.spacing, im.origin, im.direction))
