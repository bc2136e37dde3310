"""Flat binary container for a compiled scene ("model blob").

The blob is what crosses the C-ABI (`ge_create(model_blob, nbytes, ...)`, see
include/grasp_engine.h).  It replaces the in-memory `mjModel` that the reference obtains from
`mujoco_py.load_model_from_path` (reference: gym_grasper/controller/MujocoController.py:33).

Layout (little endian):
    char     magic[8]   = "GEBLOB01"
    int64    n_entries
    int64    total_bytes
    entry[n] : char name[32]; int32 dtype; int32 ndim; int64 shape[4]; int64 offset; int64 nbytes
    ...data, every array 16-byte aligned (so that rows can be bulk-copied into shared memory)...
dtype codes: 0=float64 1=int32 2=uint8 3=float32
"""
import struct

import numpy as np

MAGIC = b"GEBLOB01"
_DT = {np.dtype("float64"): 0, np.dtype("int32"): 1, np.dtype("uint8"): 2, np.dtype("float32"): 3}
_DT_INV = {v: k for k, v in _DT.items()}
_ENTRY = struct.Struct("<32sii4qqq")
_HEAD = struct.Struct("<8sqq")


def pack(arrays):
    """dict name -> ndarray  =>  bytes"""
    names = list(arrays)
    n = len(names)
    off = _HEAD.size + n * _ENTRY.size
    off = (off + 15) & ~15
    entries, chunks = [], []
    for k in names:
        a = np.ascontiguousarray(arrays[k])
        if a.dtype not in _DT:
            if np.issubdtype(a.dtype, np.integer):
                a = a.astype(np.int32)
            else:
                a = a.astype(np.float64)
        if a.ndim == 0:
            a = a.reshape(1)
        assert a.ndim <= 4 and len(k) < 32, k
        shape = list(a.shape) + [0] * (4 - a.ndim)
        nb = a.nbytes
        entries.append(_ENTRY.pack(k.encode(), _DT[a.dtype], a.ndim, *shape, off, nb))
        pad = (-nb) & 15
        chunks.append(a.tobytes() + b"\0" * pad)
        off += nb + pad
    head = _HEAD.pack(MAGIC, n, off)
    body = head + b"".join(entries)
    body += b"\0" * ((-len(body)) & 15)
    out = body + b"".join(chunks)
    assert len(out) == off
    return out


def unpack(buf):
    """bytes => dict name -> ndarray (copies)"""
    magic, n, total = _HEAD.unpack_from(buf, 0)
    if magic != MAGIC:
        raise ValueError("not a grasp-engine model blob")
    out = {}
    for i in range(n):
        name, dt, ndim, s0, s1, s2, s3, off, nb = _ENTRY.unpack_from(buf, _HEAD.size + i * _ENTRY.size)
        name = name.rstrip(b"\0").decode()
        shape = (s0, s1, s2, s3)[:ndim]
        out[name] = np.frombuffer(buf, dtype=_DT_INV[dt], count=int(np.prod(shape)) if ndim else 1,
                                  offset=off).reshape(shape).copy()
    return out
