"""Minimal numpy reader/writer of mantaflow .uni grid files (ref source/fileio.cpp:36-43, 608-660, 834-1002):
gzip stream = 4-byte magic ('MNT2' 3D, 'M4T2' 4D) + 288-byte UniHeader [+ int32 dimT for 4D] + raw payload,
x fastest.  Used by tests/bench to exchange grids with the reference and with the `manta` host module."""
import gzip
import struct

import numpy as np

_HEAD = struct.Struct("<6i256sQ")  # dimX dimY dimZ gridType elementType bytesPerElement info timestamp


def write_uni(name, a, info=b"flof-b200 synthetic"):
    a = np.ascontiguousarray(a)
    if a.dtype == np.int32:
        grid_type3, grid_type4, elem_type, bpe = 2, 2, 0, 4
        vec = False
    else:
        a = a.astype(np.float32, copy=False)
        vec = a.ndim in (4, 5) and a.shape[-1] in (3, 4) and a.ndim - 1 in (3, 4) and (a.ndim == 5 or a.shape[-1] == 3)
        bpe = 4 * (a.shape[-1] if vec else 1)
        elem_type = 2 if vec else 1
        grid_type3 = 4 if vec else 1
        grid_type4 = (8 if a.shape[-1] == 4 else 4) if vec else 1
    spatial = a.shape[:-1] if vec else a.shape
    with gzip.open(name, "wb", compresslevel=1) as f:
        if len(spatial) == 4:
            nt, nz, ny, nx = spatial
            f.write(b"M4T2")
            f.write(_HEAD.pack(nx, ny, nz, grid_type4, elem_type, bpe, info, 0))
            f.write(struct.pack("<i", nt))
        else:
            nz, ny, nx = spatial
            f.write(b"MNT2")
            f.write(_HEAD.pack(nx, ny, nz, grid_type3, elem_type, bpe, info, 0))
        f.write(a.tobytes())


def read_uni(name):
    with gzip.open(name, "rb") as f:
        magic = f.read(4)
        nx, ny, nz, _gt, _et, bpe, _info, _ts = _HEAD.unpack(f.read(_HEAD.size))
        if magic == b"M4T2":
            nt = struct.unpack("<i", f.read(4))[0]
            shape = (nt, nz, ny, nx)
        elif magic == b"MNT2":
            shape = (nz, ny, nx)
        else:
            raise ValueError("%s: unsupported uni magic %r" % (name, magic))
        comps = bpe // 4
        data = np.frombuffer(f.read(), dtype=np.int32 if _et == 0 else np.float32)
    if comps > 1:
        shape = shape + (comps,)
    return data.reshape(shape).copy()
