"""Minimal NIfTI-1 reader / writer (.nii, .nii.gz) in NumPy.

The reference does all image I/O through SimpleITK (pipeline/NiftiDataset3D.py:60-117, model.py:1191-1243),
which is not installable here.  This module covers what the patch interface needs: single-file NIfTI-1,
the common scalar dtypes, spacing / origin / direction from the qform (quaternion + qfac) or the sform rows,
data returned in SimpleITK's `GetArrayFromImage(...).transpose(2,1,0)` index order, i.e. array[x, y, z]
(NiftiDataset3D.py:150-158).

Orientation.  NIfTI stores voxel -> world maps in RAS+ coordinates; ITK / SimpleITK images live in LPS+.  Like ITK's
NiftiImageIO, `read` converts: origin and direction rows x, y are negated, so `Image.origin` / `Image.direction`
are what `sitk.ReadImage(...).GetOrigin() / GetDirection()` return (direction row-major, columns = axes), and `write`
converts back and stores both a qform (code 1) and the equivalent sform rows, so that a label volume written with the
input's geometry overlays the input in any viewer (model.py:1191-1243 copies origin, spacing and direction).
"""
from __future__ import annotations

import gzip
import struct
from dataclasses import dataclass, field
from typing import Tuple

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
           768: np.uint32}
_CODES = {np.dtype(v).str[1:]: k for k, v in _DTYPES.items()}


@dataclass
class Image:
    """array[x, y, z] plus the geometry SimpleITK images carry."""
    array: np.ndarray
    spacing: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    origin: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    direction: Tuple[float, ...] = field(default_factory=lambda: (1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0))

    def GetSize(self):
        return tuple(int(s) for s in self.array.shape)

    def GetSpacing(self):
        return self.spacing

    def GetOrigin(self):
        return self.origin

    def GetDirection(self):
        return self.direction


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def _quat_to_matrix(b: float, c: float, d: float):
    """NIfTI-1 quaternion (b, c, d; a = sqrt(1 - b^2 - c^2 - d^2)) -> proper rotation matrix (nifti1.h)."""
    a2 = 1.0 - (b * b + c * c + d * d)
    if a2 < 1e-7:   # 180 degree rotation: renormalise (b, c, d), a = 0
        n = 1.0 / np.sqrt(b * b + c * c + d * d)
        b, c, d, a = b * n, c * n, d * n, 0.0
    else:
        a = float(np.sqrt(a2))
    return np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                     [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                     [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]], np.float64)


def _matrix_to_quat(R: np.ndarray):
    """Proper rotation matrix -> (b, c, d) with a >= 0 (nifti_mat44_to_quatern's branches)."""
    a = R[0, 0] + R[1, 1] + R[2, 2] + 1.0
    if a > 0.5:
        a = 0.5 * np.sqrt(a)
        b = 0.25 * (R[2, 1] - R[1, 2]) / a
        c = 0.25 * (R[0, 2] - R[2, 0]) / a
        d = 0.25 * (R[1, 0] - R[0, 1]) / a
    else:
        xd, yd, zd = 1.0 + R[0, 0] - (R[1, 1] + R[2, 2]), 1.0 + R[1, 1] - (R[0, 0] + R[2, 2]), 1.0 + R[2, 2] - (R[0, 0] + R[1, 1])
        if xd > 1.0:
            b = 0.5 * np.sqrt(xd)
            c = 0.25 * (R[0, 1] + R[1, 0]) / b
            d = 0.25 * (R[0, 2] + R[2, 0]) / b
            a = 0.25 * (R[2, 1] - R[1, 2]) / b
        elif yd > 1.0:
            c = 0.5 * np.sqrt(yd)
            b = 0.25 * (R[0, 1] + R[1, 0]) / c
            d = 0.25 * (R[1, 2] + R[2, 1]) / c
            a = 0.25 * (R[0, 2] - R[2, 0]) / c
        else:
            d = 0.5 * np.sqrt(zd)
            b = 0.25 * (R[0, 2] + R[2, 0]) / d
            c = 0.25 * (R[1, 2] + R[2, 1]) / d
            a = 0.25 * (R[1, 0] - R[0, 1]) / d
        if a < 0.0:
            b, c, d = -b, -c, -d
    return float(b), float(c), float(d)


_RAS_TO_LPS = np.diag([-1.0, -1.0, 1.0])


def _geometry(raw: bytes, end: str, pixdim):
    """(origin, direction) in ITK's LPS convention from the header's qform (preferred, as ITK does) or sform."""
    qform_code, sform_code = struct.unpack(end + "2h", raw[252:256])
    spacing = np.array([abs(p) if p != 0 else 1.0 for p in pixdim[1:4]], np.float64)
    if qform_code > 0:
        b, c, d = struct.unpack(end + "3f", raw[256:268])
        off = np.array(struct.unpack(end + "3f", raw[268:280]), np.float64)
        R = _quat_to_matrix(b, c, d)
        if pixdim[0] < 0:   # qfac = -1: the third axis is flipped
            R[:, 2] = -R[:, 2]
    elif sform_code > 0:
        rows = np.array([struct.unpack(end + "4f", raw[280 + 16 * i:296 + 16 * i]) for i in range(3)], np.float64)
        off = rows[:, 3]
        M = rows[:, :3]
        norms = np.sqrt((M ** 2).sum(0))
        norms[norms == 0] = 1.0
        R = M / norms
        spacing = norms
    else:   # NIfTI "method 1": no orientation information, axes taken as given
        return (0.0, 0.0, 0.0), (1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0), tuple(float(x) for x in spacing)
    R = _RAS_TO_LPS @ R
    off = _RAS_TO_LPS @ off
    return tuple(float(x) for x in off), tuple(float(x) for x in R.reshape(-1)), tuple(float(x) for x in spacing)


def read(path: str) -> Image:
    with _open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 352:
        raise ValueError("%s: not a NIfTI-1 file (%d bytes; git-LFS pointer?)" % (path, len(raw)))
    end = "<" if struct.unpack("<i", raw[:4])[0] == 348 else ">"
    if struct.unpack(end + "i", raw[:4])[0] != 348:
        raise ValueError("%s: bad NIfTI-1 header" % path)
    dim = struct.unpack(end + "8h", raw[40:56])
    datatype = struct.unpack(end + "h", raw[70:72])[0]
    pixdim = struct.unpack(end + "8f", raw[76:108])
    vox_offset = int(struct.unpack(end + "f", raw[108:112])[0])
    slope, inter = struct.unpack(end + "2f", raw[112:120])
    if datatype not in _DTYPES:
        raise ValueError("%s: unsupported NIfTI datatype %d" % (path, datatype))
    shape = tuple(int(d) for d in dim[1:1 + max(3, dim[0])])[:3]
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(end)
    n = int(np.prod(shape))
    data = np.frombuffer(raw, dt, count=n, offset=max(vox_offset, 352)).reshape(shape, order="F")
    arr = np.array(data, dtype=dt.newbyteorder("="))
    if slope not in (0.0, 1.0) or inter != 0.0:
        arr = arr.astype(np.float32) * (slope if slope != 0 else 1.0) + inter
    origin, direction, spacing = _geometry(raw, end, pixdim)
    return Image(arr, spacing, origin, direction)


def write(path: str, image: Image):
    arr = np.asarray(image.array)
    code = _CODES.get(arr.dtype.str[1:])
    if code is None:
        arr = arr.astype(np.float32)
        code = 16
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    dims = [3] + list(arr.shape) + [1] * 4
    struct.pack_into("<8h", hdr, 40, *dims)
    struct.pack_into("<h", hdr, 70, code)
    struct.pack_into("<h", hdr, 72, arr.dtype.itemsize * 8)
    # geometry: ITK (LPS) direction / origin back to NIfTI's RAS qform + sform
    spacing = [float(s) for s in image.spacing]
    R = _RAS_TO_LPS @ np.asarray(image.direction, np.float64).reshape(3, 3)
    off = _RAS_TO_LPS @ np.asarray(image.origin, np.float64)
    qfac = 1.0
    if np.linalg.det(R) < 0:   # improper: flip the third axis and record it in qfac
        R = R.copy()
        R[:, 2] = -R[:, 2]
        qfac = -1.0
    b, c, d = _matrix_to_quat(R)
    struct.pack_into("<8f", hdr, 76, qfac, *spacing, 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<2f", hdr, 112, 1.0, 0.0)
    struct.pack_into("<2h", hdr, 252, 1, 1)  # qform_code, sform_code (scanner anatomical)
    struct.pack_into("<3f", hdr, 256, b, c, d)
    struct.pack_into("<3f", hdr, 268, *[float(o) for o in off])
    Rs = R.copy()
    Rs[:, 2] *= qfac
    for i in range(3):
        struct.pack_into("<4f", hdr, 280 + 16 * i, *[float(Rs[i, k] * spacing[k]) for k in range(3)], float(off[i]))
    hdr[344:348] = b"n+1\0"
    with _open(path, "wb") as f:
        f.write(bytes(hdr) + b"\0\0\0\0" + np.asfortranarray(arr).tobytes(order="F"))
