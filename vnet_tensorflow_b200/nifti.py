"""Minimal NIfTI-1 reader / writer (.nii, .nii.gz) in NumPy.

The reference does all image I/O through SimpleITK (pipeline/NiftiDataset3D.py:60-117, model.py:1191-1243),
which is not installable here.  This module covers what the patch interface needs: single-file NIfTI-1,
the common scalar dtypes, spacing and origin from pixdim / qoffset, data returned in SimpleITK's
`GetArrayFromImage(...).transpose(2,1,0)` index order, i.e. array[x, y, z] (NiftiDataset3D.py:150-158).
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
    qoff = struct.unpack(end + "3f", raw[268:280])
    if datatype not in _DTYPES:
        raise ValueError("%s: unsupported NIfTI datatype %d" % (path, datatype))
    shape = tuple(int(d) for d in dim[1:1 + max(3, dim[0])])[:3]
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(end)
    n = int(np.prod(shape))
    data = np.frombuffer(raw, dt, count=n, offset=max(vox_offset, 352)).reshape(shape, order="F")
    arr = np.array(data, dtype=dt.newbyteorder("="))
    if slope not in (0.0, 1.0) or inter != 0.0:
        arr = arr.astype(np.float32) * (slope if slope != 0 else 1.0) + inter
    return Image(arr, tuple(float(p) for p in pixdim[1:4]), tuple(float(q) for q in qoff))


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
    struct.pack_into("<8f", hdr, 76, 1.0, *[float(s) for s in image.spacing], 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<2f", hdr, 112, 1.0, 0.0)
    struct.pack_into("<h", hdr, 252, 1)  # qform_code
    struct.pack_into("<3f", hdr, 268, *[float(o) for o in image.origin])
    hdr[344:348] = b"n+1\0"
    with _open(path, "wb") as f:
        f.write(bytes(hdr) + b"\0\0\0\0" + np.asfortranarray(arr).tobytes(order="F"))
