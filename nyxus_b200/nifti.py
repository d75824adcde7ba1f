"""Minimal NIfTI-1 reader / writer (single-file .nii and .nii.gz) for the Nyxus3D entry points.

The reference decodes volumes with its vendored niftilib through NiftiLoader (src/nyx/raw_nifti.h:189-330): integer voxel
types are handed over as unsigned 32-bit grey values.  File decoding is outside the B200 hot path; this
reader covers the voxel types that map to unsigned grey values without a floating-point dynamic-range remap:
uint8 / int16 / uint16 / int32 / uint32 and float32 / float64 volumes whose values are non-negative integers."""
import gzip
import struct

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16, 768: np.uint32}
_CODES = {np.dtype(v).str[1:]: k for k, v in _DTYPES.items()}


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def read(path):
    """Returns the volume as [D, H, W] (first time frame), in the file's voxel type."""
    with _open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 352:
        raise IOError(f"{path}: not a NIfTI-1 file")
    end = "<"
    if struct.unpack("<i", raw[0:4])[0] != 348:
        if struct.unpack(">i", raw[0:4])[0] != 348:
            raise IOError(f"{path}: bad NIfTI-1 header size")
        end = ">"
    dim = struct.unpack(end + "8h", raw[40:56])
    datatype, bitpix = struct.unpack(end + "hh", raw[70:74])
    vox_offset = int(struct.unpack(end + "f", raw[108:112])[0])
    scl_slope, scl_inter = struct.unpack(end + "ff", raw[112:120])
    magic = raw[344:348]
    if magic[:3] not in (b"n+1",):
        raise IOError(f"{path}: only single-file NIfTI-1 (magic n+1) is supported")
    if datatype not in _DTYPES:
        raise IOError(f"{path}: unsupported NIfTI data type {datatype}")
    nd = dim[0]
    W, H = dim[1], dim[2]
    D = dim[3] if nd >= 3 else 1
    T = dim[4] if nd >= 4 else 1
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(end)
    n = W * H * D * max(T, 1)
    a = np.frombuffer(raw, dtype=dt, count=n, offset=max(vox_offset, 352)).reshape(max(T, 1), D, H, W)[0]
    if scl_slope not in (0.0, 1.0) or (scl_slope != 0.0 and scl_inter != 0.0):
        a = a.astype(np.float64) * scl_slope + scl_inter
    return a


def to_grey(a, what="intensity"):
    """voxels -> uint32 grey values (uint16 kept as is)"""
    if a.dtype == np.uint16 or a.dtype == np.uint32:
        return np.ascontiguousarray(a.astype(a.dtype.newbyteorder("=")))
    if np.issubdtype(a.dtype, np.floating):
        if not np.all(np.isfinite(a)) or np.any(a < 0) or np.any(a != np.floor(a)):
            raise ValueError(f"{what} volume holds non-integer or negative floating-point voxels: the floating-point "
                             "dynamic-range remap of the reference's loaders is outside the path implemented by nyxus_b200")
    elif np.any(a < 0):
        raise ValueError(f"{what} volume holds negative voxels (Hounsfield-type scale): outside the path implemented by nyxus_b200")
    return np.ascontiguousarray(a.astype(np.uint32))


def write(path, vol):
    """vol: [D, H, W] array of a supported voxel type"""
    vol = np.ascontiguousarray(vol)
    code = _CODES[vol.dtype.str[1:]]
    D, H, W = vol.shape
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, W, H, D, 1, 1, 1, 1)
    struct.pack_into("<hh", hdr, 70, code, vol.dtype.itemsize * 8)
    struct.pack_into("<8f", hdr, 76, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<ff", hdr, 112, 1.0, 0.0)
    hdr[344:348] = b"n+1\0"
    with _open(path, "wb") as f:
        f.write(bytes(hdr) + b"\0\0\0\0" + vol.astype(vol.dtype.newbyteorder("<")).tobytes())
