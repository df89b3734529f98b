"""Patch interface of the reference's data layer (pipeline/NiftiDataset3D.py), without SimpleITK/tf.data.

Kept: the class / constructor names that the YAML pipelines resolve by `getattr(NiftiDataset3D, name)
(**variables)` (model.py:341-402), the sample dict {'image': [img per modality], 'label': img}, and the
output contract of `NiftiDataset.get_dataset()`: (image float32 [X,Y,Z,M], label int32 [X,Y,Z]) with
labels remapped to class *indices* (NiftiDataset3D.py:119-137,150-165).  Images are vnet_tensorflow_b200.
nifti.Image objects (NumPy array[x,y,z] + spacing/origin).  Resampling follows sitk.ResampleImageFilter's
index arithmetic (`resample_image`: linear / nearest, zero outside the input) in NumPy; accelerating this CPU
stage is out of the hot path's scope (SURVEY.md "next" row N1).
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
        skip = (".DS_Store", "@eaDir")  # NiftiDataset3D.py:41-45
        return [os.path.join(self.data_dir, c) for c in sorted(os.listdir(self.data_dir)) if c not in skip]

    @staticmethod
    def _same_header(a, b):
        return (a.GetSize() == b.GetSize(), tuple(a.GetSpacing()) == tuple(b.GetSpacing()),
                tuple(a.GetDirection()) == tuple(b.GetDirection()))

    def input_parser(self, case_dir):
        images = []
        for ch in self.image_filenames:
            try:
                images.append(self.read_image(os.path.join(case_dir, ch)))
            except Exception:
                raise Exception("Error loading image: {}".format(os.path.join(case_dir, ch)))
        for ch, im in zip(self.image_filenames, images):  # NiftiDataset3D.py:72-86: all modalities share the grid
            same = self._same_header(im, images[0])
            if not all(same):
                raise Exception('Header info inconsistent: {}\nSame size: {}\nSame spacing: {}\nSame direction: {}'
                                .format(os.path.join(case_dir, ch), *same))
        ref = images[0]
        label = nifti.Image(np.zeros(ref.GetSize(), np.uint8), ref.spacing, ref.origin, ref.direction)
        if self.train:
            path = os.path.join(case_dir, self.label_filename)
            try:
                label_ = self.read_image(path)
            except Exception:
                raise Exception("Error loading label: {}".format(path))
            same = self._same_header(label_, ref)
            if not all(same):
                raise Exception('Header info inconsistent: {}\nSame size: {}\nSame spacing: {}\nSame direction: {}'
                                .format(path, *same))
            # NiftiDataset3D.py:119-137: label value -> class index *before* the transforms see it; values that are
            # not in `labels` become background
            src = np.asarray(label_.array)
            remapped = np.zeros(src.shape, np.uint8)
            for idx, value in enumerate(self.labels):
                remapped[src == value] = idx
            label = nifti.Image(remapped, ref.spacing, ref.origin, ref.direction)
        sample = {'image': images, 'label': label}
        if self.transforms:
            for transform in self.transforms:
                sample = transform(sample)
        image_np = np.stack([np.asarray(im.array, np.float32) for im in sample['image']], axis=-1)
        return image_np.astype(np.float32), np.asarray(sample['label'].array).astype(np.int32)

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


def resample_image(im, spacing, size, origin, order):
    """sitk.ResampleImageFilter with an identity transform, as the reference configures it (output spacing / size /
    origin given, direction kept; NiftiDataset3D.py:378-396, model.py:957-975): output voxel j of an axis sits at
    origin + j*spacing and reads the input at continuous index (origin + j*spacing - im.origin) / im.spacing, linearly
    (order 1, neighbours clamped to the edge) or at the nearest voxel (order 0, halves round up); positions outside
    [-0.5, n-0.5) give 0.  Axes are independent because the direction matrix is shared."""
    a = np.asarray(im.array)
    out = a.astype(np.float32) if order else a
    for axis in range(3):
        n = a.shape[axis]
        c = (origin[axis] + np.arange(size[axis], dtype=np.float64) * spacing[axis] - im.origin[axis]) / im.spacing[axis]
        inside = (c >= -0.5) & (c < n - 0.5)
        shape = [1, 1, 1]
        shape[axis] = -1
        if order:
            i0 = np.floor(c).astype(np.int64)
            f = (c - i0).astype(np.float32).reshape(shape)
            lo, hi = np.clip(i0, 0, n - 1), np.clip(i0 + 1, 0, n - 1)
            out = np.take(out, lo, axis) * (1 - f) + np.take(out, hi, axis) * f
        else:
            out = np.take(out, np.clip(np.floor(c + 0.5).astype(np.int64), 0, n - 1), axis)
        out = out * inside.reshape(shape).astype(out.dtype)
    return nifti.Image(out.astype(np.float32 if order else a.dtype), tuple(float(v) for v in spacing),
                       tuple(float(v) for v in origin), im.direction)


def _like(im, array):
    return nifti.Image(array, im.spacing, im.origin, im.direction)


def _cast(values, dtype):
    """static_cast<pixel type>: truncation toward zero for integer images (after clamping to the type's range, as
    NumPy leaves an out-of-range cast undefined), rounding to nearest for float32."""
    dtype = np.dtype(dtype)
    if dtype.kind in "iu":
        info = np.iinfo(dtype)
        return np.trunc(np.clip(values, info.min, info.max)).astype(dtype)
    return np.asarray(values).astype(dtype)


def intensity_window(a, window_min, window_max, out_min=0.0, out_max=255.0):
    """sitk.IntensityWindowingImageFilter: below / above the window -> out_min / out_max, inside x * scale + shift in
    double precision, cast back to the image's own pixel type; the window bounds are first cast to that type, which
    truncates them for integer images (CT in int16 comes out as whole numbers 0..255)."""
    a = np.asarray(a)
    if a.dtype.kind in "iu":
        lo, hi = int(_cast(window_min, a.dtype)), int(_cast(window_max, a.dtype))
    else:
        lo, hi = float(a.dtype.type(window_min)), float(a.dtype.type(window_max))
    scale = (out_max - out_min) / (hi - lo) if hi != lo else 0.0
    v = a.astype(np.float64) * scale + (out_min - lo * scale)
    v = np.where(a < lo, out_min, np.where(a > hi, out_max, v))
    return _cast(v, a.dtype)


class StatisticalNormalization(object):
    """NiftiDataset3D.py:210-254: window mean +- sigma * std (StatisticsImageFilter: the N-1 estimate), clamped to the
    pixel type's range, mapped to 0..255; `pre_norm` first standardises the image (NormalizeImageFilter, float)."""

    def __init__(self, sigma, pre_norm=False):
        self.name = 'StatisticalNormalization'
        assert isinstance(sigma, float)
        self.sigma, self.pre_norm = sigma, pre_norm

    def __call__(self, sample):
        def norm(im):
            a = np.asarray(im.array)
            if self.pre_norm:
                d = a.astype(np.float64)
                a = ((d - d.mean()) / max(d.std(ddof=1), 1e-30)).astype(np.float32)
            d = a.astype(np.float64)
            mean, std = d.mean(), (d.std(ddof=1) if d.size > 1 else 0.0)
            info = np.iinfo(a.dtype) if a.dtype.kind in "iu" else np.finfo(a.dtype)
            hi = min(mean + self.sigma * std, float(info.max))
            lo = max(mean - self.sigma * std, float(info.min))
            return _like(im, intensity_window(a, lo, hi))
        return _map(sample, norm)


class ManualNormalization(object):
    """NiftiDataset3D.py:285-308: window [windowMin, windowMax] -> 0..255."""

    def __init__(self, windowMin, windowMax):
        self.name = 'ManualNormalization'
        assert isinstance(windowMax, (int, float)) and isinstance(windowMin, (int, float))
        self.windowMin, self.windowMax = float(windowMin), float(windowMax)

    def __call__(self, sample):
        return _map(sample, lambda im: _like(im, intensity_window(im.array, self.windowMin, self.windowMax)))


class Normalization(object):
    """NiftiDataset3D.py:167-185 (RescaleIntensityImageFilter): the full intensity range -> 0..255, in the image's
    pixel type.  Applied per modality (the reference hands the filter the modality list itself)."""

    def __init__(self):
        self.name = 'Normalization'

    def __call__(self, sample):
        def norm(im):
            a = np.asarray(im.array)
            lo, hi = float(a.min()), float(a.max())
            scale = 255.0 / (hi - lo) if hi != lo else 0.0
            return _like(im, _cast(a.astype(np.float64) * scale - lo * scale, a.dtype))
        return _map(sample, norm)


class Resample(object):
    """NiftiDataset3D.py:345-398: resample to `voxel_size` on the grid size ceil(extent / voxel_size) from the same
    origin (linear for images, nearest for the label)."""

    def __init__(self, voxel_size):
        self.name = 'Resample'
        self.voxel_size = (voxel_size,) * 3 if isinstance(voxel_size, (int, float)) else tuple(voxel_size)

    def _res(self, im, order, like=None):
        like = like or im
        size = [int(np.ceil(sp * n / v)) for sp, n, v in zip(like.spacing, like.GetSize(), self.voxel_size)]
        return resample_image(im, self.voxel_size, size, im.origin, order)

    def __call__(self, sample):
        last = sample['image'][-1]  # the label reuses the resampler of the last modality (NiftiDataset3D.py:391-396)
        return _map(sample, lambda im: self._res(im, 1), lambda lb: self._res(lb, 0, last))


class Padding(object):
    """NiftiDataset3D.py:400-456: when any axis is shorter than `output_size`, resample onto the larger grid from
    the same origin - zeros appended at the far end of the short axes, voxel (0,0,0) stays where it was."""

    def __init__(self, output_size):
        self.name = 'Padding'
        self.output_size = (output_size,) * 3 if isinstance(output_size, int) else tuple(output_size)
        assert all(i > 0 for i in self.output_size)

    def _pad(self, im):
        a = np.asarray(im.array)
        pads = [(0, max(o - s, 0)) for s, o in zip(a.shape, self.output_size)]
        return nifti.Image(np.pad(a, pads), im.spacing, im.origin, im.direction)

    def __call__(self, sample):
        return _map(sample, self._pad, self._pad)


class RandomCrop(object):
    """NiftiDataset3D.py:458-551: random crop of `output_size`, redrawn until it holds at least `min_pixel` foreground
    voxels or, with probability `drop_ratio` per failed draw, is accepted anyway."""

    MAX_DRAWS = 100000  # the reference loops for ever on a label-free case with drop_ratio 0

    def __init__(self, output_size, drop_ratio=0.1, min_pixel=1):
        self.name = 'Random Crop'
        assert isinstance(output_size, (int, tuple, list))
        self.output_size = (output_size,) * 3 if isinstance(output_size, int) else tuple(output_size)
        assert len(self.output_size) == 3
        assert isinstance(drop_ratio, (int, float))
        if not 0 <= drop_ratio <= 1:
            raise RuntimeError('Drop ratio should be between 0 and 1')
        assert isinstance(min_pixel, int)
        if min_pixel < 0:
            raise RuntimeError('Min label pixel count should be integer larger than 0')
        self.drop_ratio, self.min_pixel = drop_ratio, min_pixel

    def drop(self, probability):
        return random.random() <= probability

    def _crop(self, im, st):
        sl = tuple(slice(s, s + o) for s, o in zip(st, self.output_size))
        origin = tuple(o + k * sp for o, k, sp in zip(im.origin, st, im.spacing))  # RegionOfInterest keeps physical positions
        return nifti.Image(np.asarray(im.array)[sl], im.spacing, origin, im.direction)

    def _crop_sample(self, sample, st):
        return _map(sample, lambda im: self._crop(im, st), lambda lb: self._crop(lb, st))

    def __call__(self, sample):
        fg = np.asarray(sample['label'].array) >= 1
        old, new = fg.shape, self.output_size
        for _ in range(self.MAX_DRAWS):
            # np.random.randint's upper bound is exclusive: the last admissible start is never drawn (as in the reference)
            st = [0 if old[i] <= new[i] else int(np.random.randint(0, old[i] - new[i])) for i in range(3)]
            sl = tuple(slice(s, s + o) for s, o in zip(st, new))
            if fg[sl].sum() >= self.min_pixel or self.drop(self.drop_ratio):
                break
        return self._crop_sample(sample, st)


class ConfidenceCrop2(RandomCrop):
    """NiftiDataset3D.py:661-793: with probability `probability` (in tenths: int(10p) ones against int(10(1-p)) zeros)
    a crop centred on the bounding box of a random connected foreground component, shifted per axis by a uniform
    integer in [-rand_range, rand_range] and clamped into the volume; otherwise a random region (`random_empty_region`:
    redrawn until it holds no foreground)."""

    def __init__(self, output_size, rand_range=3, probability=0.5, random_empty_region=False):
        super().__init__(output_size)
        self.name = 'Confidence Crop 2'
        assert isinstance(rand_range, (int, tuple, list))
        self.rand_range = (rand_range,) * 3 if isinstance(rand_range, int) else tuple(rand_range)
        assert len(self.rand_range) == 3 and all(r >= 0 for r in self.rand_range)
        assert isinstance(probability, float) and 0 <= probability <= 1
        self.probability = probability
        assert isinstance(random_empty_region, bool)
        self.random_empty_region = random_empty_region

    def _random_index(self, shape):
        # range(0, size - crop - 1): the reference raises on a one-voxel margin; that case starts at 0 here
        return [0 if shape[i] - self.output_size[i] - 1 <= 0 else random.choice(range(0, shape[i] - self.output_size[i] - 1))
                for i in range(3)]

    def RandomRegion(self, sample):
        return self._crop_sample(sample, self._random_index(np.asarray(sample['label'].array).shape))

    def RandomEmptyRegion(self, sample):
        lab = np.asarray(sample['label'].array)
        for _ in range(self.MAX_DRAWS):
            st = self._random_index(lab.shape)
            if lab[tuple(slice(s, s + o) for s, o in zip(st, self.output_size))].sum() < 1:
                break
        return self._crop_sample(sample, st)

    def __call__(self, sample):
        lab = np.asarray(sample['label'].array).astype(np.int16)
        choices = [0] * int(10 * (1 - self.probability)) + [1] * int(10 * self.probability)
        positive = random.choice(choices)
        n = 0
        if positive:
            from scipy import ndimage
            comp, n = ndimage.label(lab != 0)
        if not positive or n == 0:
            return self.RandomEmptyRegion(sample) if self.random_empty_region else self.RandomRegion(sample)
        selected = random.choice(range(0, n)) + 1
        box = ndimage.find_objects(comp)[selected - 1]
        index = [0, 0, 0]
        for i in range(3):
            extent = box[i].stop - box[i].start
            index[i] = box[i].start + int(extent / 2) - int(self.output_size[i] / 2) + \
                random.choice(range(-1 * self.rand_range[i], self.rand_range[i] + 1))
            if lab.shape[i] - index[i] - 1 < self.output_size[i]:
                index[i] = lab.shape[i] - self.output_size[i] - 1
            if index[i] < 0:
                index[i] = 0
        return self._crop_sample(sample, index)


class RandomNoise(object):
    """NiftiDataset3D.py:553-572 (AdditiveGaussianNoiseImageFilter): x + N(0, sigma) on every modality, clamped and
    cast back to the pixel type."""

    def __init__(self, sigma=5):
        self.name = 'Random Noise'
        self.sigma = sigma

    def __call__(self, sample):
        def noise(im):
            a = np.asarray(im.array)
            return _like(im, _cast(a.astype(np.float64) + np.random.normal(0, self.sigma, a.shape), a.dtype))
        return _map(sample, noise)


class RandomFlip(object):
    """NiftiDataset3D.py:187-208: one coin per sample; heads flips all the axes marked in `axes` together."""

    def __init__(self, axes):
        self.name = 'Flip'
        assert len(axes) > 0 and len(axes) <= 3
        self.axes = axes

    def __call__(self, sample):
        if not np.random.randint(2, size=1)[0]:
            return sample
        flips = [ax for ax, on in enumerate(self.axes) if on]
        f = lambda im: _like(im, np.flip(np.asarray(im.array), flips).copy() if flips else im.array)
        return _map(sample, f, f)


class Invert(object):
    """NiftiDataset3D.py:330-343 (InvertIntensityImageFilter, maximum 255): 255 - x in the image's pixel type."""

    def __init__(self):
        self.name = 'Invert'

    def __call__(self, sample):
        return _map(sample, lambda im: _like(im, _cast(255.0 - np.asarray(im.array).astype(np.float64), np.asarray(im.array).dtype)))


class ExtremumNormalization(object):
    """NiftiDataset3D.py:256-283: window [min + percent*range, min + (1-percent)*range] of each modality -> 0..255,
    values outside the window clamped (IntensityWindowingImageFilter)."""

    def __init__(self, percent=0.05):
        self.name = 'ExtremumNormalization'
        assert isinstance(percent, float)
        self.percent = percent

    def __call__(self, sample):
        def norm(im):
            a = np.asarray(im.array, np.float32)
            lo_all, hi_all = float(a.min()), float(a.max())
            lo = (hi_all - lo_all) * self.percent + lo_all
            hi = (hi_all - lo_all) * (1 - self.percent) + lo_all
            a = (np.clip(a, lo, hi) - lo) / max(hi - lo, 1e-6) * 255.0
            return nifti.Image(a.astype(np.float32), im.spacing, im.origin, im.direction)
        return _map(sample, norm)


class Reorient(object):
    """NiftiDataset3D.py:310-328 (PermuteAxesImageFilter): output axis i is input axis order[i]; spacing and
    origin follow their axes.  Applied to every modality and to the label."""

    def __init__(self, order):
        self.name = 'Reoreient'
        assert isinstance(order, (tuple, list)) and len(order) == 3 and sorted(order) == [0, 1, 2]
        self.order = tuple(int(o) for o in order)

    def __call__(self, sample):
        def perm(im):
            o = self.order
            return nifti.Image(np.ascontiguousarray(np.transpose(np.asarray(im.array), o)), tuple(im.spacing[k] for k in o),
                               tuple(im.origin[k] for k in o), im.direction)
        return _map(sample, perm, perm)


class ConfidenceCrop(RandomCrop):
    """NiftiDataset3D.py:574-659: crop of `output_size` around the centroid of a randomly chosen connected
    foreground component (the volume's first-octant centre when there is none), shifted per axis by a rounded
    N(0, sigma * size / 2) offset that is redrawn until the crop lies inside the volume."""

    def __init__(self, output_size, sigma=2.5):
        super().__init__(output_size)
        self.name = 'Confidence Crop'
        assert isinstance(sigma, (float, tuple, list))
        self.sigma = (sigma,) * 3 if isinstance(sigma, float) else tuple(sigma)
        assert len(self.sigma) == 3 and all(s >= 0 for s in self.sigma)

    def NormalOffset(self, size, sigma):
        s = np.random.normal(0, size * sigma / 2, 100)
        return int(round(random.choice(s)))

    def __call__(self, sample):
        lab = np.asarray(sample['label'].array)
        size = self.output_size
        if any(s < o for s, o in zip(lab.shape, size)):
            raise ValueError("ConfidenceCrop: volume %s is smaller than the crop %s (pad first)" % (lab.shape, tuple(size)))
        from scipy import ndimage
        comp, n = ndimage.label(lab.astype(np.int8) != 0)
        if n == 0:
            centroid = [int(o / 2) for o in size]
        else:
            k = random.randint(1, n)
            centroid = [int(round(c)) for c in ndimage.center_of_mass(comp == k)]
        start = [0, 0, 0]
        for i in range(3):
            if centroid[i] < size[i] / 2:
                centroid[i] = int(size[i] / 2)
            elif lab.shape[i] - centroid[i] < size[i] / 2:
                centroid[i] = lab.shape[i] - int(size[i] / 2) - 1
            if self.sigma[i] == 0 or lab.shape[i] == size[i]:   # a zero offset is the only one that can be accepted
                start[i] = min(max(centroid[i] - int(size[i] / 2), 0), lab.shape[i] - size[i])
                continue
            while True:
                start[i] = centroid[i] + self.NormalOffset(size[i], self.sigma[i]) - int(size[i] / 2)
                if start[i] >= 0 and start[i] + size[i] - 1 <= lab.shape[i] - 1:
                    break
        return _map(sample, lambda im: self._crop(im, start), lambda lb: self._crop(lb, start))


class BSplineDeformation(object):
    """NiftiDataset3D.py:795-835: free-form deformation by a cubic B-spline over a 10x10x10 mesh of the image domain
    (13^3 control points per displacement component, ITK's BSplineTransform layout: grid spacing = extent / 10, first
    control point one grid step before the origin) whose coefficients are uniform in [0, randomness) physical units;
    output(x) = input(x + d(x)).  Images are interpolated linearly as sitk.Resample's default does; labels with
    nearest neighbour, which keeps them class values (the reference resamples them linearly and truncates)."""

    MESH, ORDER = 10, 3

    def __init__(self, randomness=10):
        self.name = 'BSpline Deformation'
        assert isinstance(randomness, (int, float))
        if randomness > 0:
            self.randomness = randomness
        else:
            raise RuntimeError('Randomness should be non zero values')

    def displacement(self, shape, spacing, coeffs):
        """d[c][x,y,z] in voxels of axis c: the B-spline with control coefficients `coeffs` [3,13,13,13] (physical units)."""
        from scipy import ndimage
        # continuous control-grid index of voxel i along an axis: position i*spacing over grid step extent/MESH, plus
        # the (ORDER-1)/2 = 1 control point that sits before the domain origin
        u = [np.arange(n, dtype=np.float64) * self.MESH / n + (self.ORDER - 1) / 2 for n in shape]
        grid = np.meshgrid(*u, indexing="ij")
        return [ndimage.map_coordinates(coeffs[c], grid, order=3, prefilter=False, mode="nearest") / spacing[c]
                for c in range(3)]

    def __call__(self, sample):
        from scipy import ndimage
        ref = sample['image'][0]
        shape = np.asarray(ref.array).shape
        n_ctrl = self.MESH + self.ORDER
        coeffs = np.random.random(3 * n_ctrl ** 3).reshape(3, n_ctrl, n_ctrl, n_ctrl) * self.randomness
        # ITK stores the parameters x-fastest per component; the draw is i.i.d., so the order has no effect here
        d = self.displacement(shape, ref.spacing, coeffs)
        idx = np.meshgrid(*[np.arange(n, dtype=np.float64) for n in shape], indexing="ij")
        coords = [i + dd for i, dd in zip(idx, d)]

        def warp(im, order):
            a = np.asarray(im.array)
            out = ndimage.map_coordinates(a.astype(np.float32) if order else a, coords, order=order, mode="constant", cval=0)
            return nifti.Image(out.astype(a.dtype), im.spacing, im.origin, im.direction)
        return _map(sample, lambda im: warp(im, 1), lambda lb: warp(lb, 0))
