"""Deterministic synthetic patches in the NiftiDataset3D patch format.

The reference's data (`data/*/image.nii`) are git-LFS stubs and SimpleITK is unavailable, so every
test and benchmark feeds patches generated here.  The format is exactly what
`pipeline/NiftiDataset3D.py:150-165` hands to the graph: image float32 [X,Y,Z,M] with intensities in
0..255 (`NiftiDataset3D.py:236-238`), label int32 [X,Y,Z] holding class *indices*
(`NiftiDataset3D.py:119-137`).  Batches are stacked to [N,X,Y,Z,M] / [N,X,Y,Z] (model.py:311-312).
"""
from __future__ import annotations

import numpy as np

NESTED_RADII = (1.0 / 4.0, 1.0 / 6.0, 1.0 / 10.0)


def _box3(a: np.ndarray) -> np.ndarray:
    """3-tap box filter along each spatial axis with edge replication (keeps the value range)."""
    for ax in range(3):
        p = np.concatenate([a.take([0], axis=ax), a, a.take([-1], axis=ax)], axis=ax)
        n = a.shape[ax]
        a = (p.take(range(0, n), axis=ax) + p.take(range(1, n + 1), axis=ax)
             + p.take(range(2, n + 2), axis=ax)) / 3.0
    return a


def synth_patch(seed: int, patch: int, modalities: int = 1, classes: int = 2, shape=None):
    """One (image [X,Y,Z,M] f32, label [X,Y,Z] i32, distmap [X,Y,Z] f32) sample."""
    rng = np.random.Generator(np.random.PCG64(seed))
    dims = tuple(shape) if shape is not None else (patch, patch, patch)
    P = min(dims)
    centre = np.array([d / 2.0 for d in dims]) + rng.uniform(-P / 8.0, P / 8.0, size=3)
    grid = np.meshgrid(*[np.arange(d, dtype=np.float32) for d in dims], indexing="ij")
    dist = np.sqrt(sum((g - c) ** 2 for g, c in zip(grid, centre))).astype(np.float32)
    label = np.zeros(dims, np.int32)
    for k in range(1, classes):
        # nested spheres: class k lives inside class k-1 (BraTS-like nesting)
        radius = P * NESTED_RADII[k - 1] if k <= len(NESTED_RADII) else P * NESTED_RADII[-1] / (k - 2)
        label[dist < radius] = k
    image = np.empty(dims + (modalities,), np.float32)
    for m in range(modalities):
        bg = _box3(rng.uniform(0.0, 255.0, size=dims).astype(np.float32))
        fg = 64.0 * (m + 1) / modalities
        image[..., m] = np.clip(bg + fg * (label > 0), 0.0, 255.0)
    surf = np.abs(dist - P * NESTED_RADII[0])
    distmap = np.exp(-(surf ** 2) / (2.0 * (P / 8.0) ** 2)).astype(np.float32)
    return image, label, distmap


def synth_batch(step: int, batch: int, patch: int, modalities: int = 1, classes: int = 2,
                rank: int = 0, shape=None):
    """Batch for optimiser step `step` on data-parallel rank `rank`:
    seeds 1234 + 1000*rank + step (+ 100000*i for batch element i)."""
    imgs, labs = [], []
    for i in range(batch):
        im, lb, _ = synth_patch(1234 + 1000 * rank + step + 100000 * i, patch, modalities, classes, shape)
        imgs.append(im)
        labs.append(lb)
    return np.stack(imgs, 0), np.stack(labs, 0)
