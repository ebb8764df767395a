"""On-disk form of descriptors / key points (reference include/SerializationUtils.h:76-153, the primitive stream of the boost
text / binary archives used by System::SaveAtlas / LoadAtlas, src/System.cc:1339-1475) over the C ABI."""
import ctypes as C

import numpy as np

from .capi import KP_DTYPE, OrbxError, lib, ptr


def _chk(n):
    if n < 0:
        raise OrbxError(int(n), lib().orbx_last_error().decode("utf-8", "replace"))
    return int(n)


def serialize_descriptors(desc, text=False):
    """serializeMatrix(ar, mDescriptors): bytes of the archive's primitive stream."""
    d = np.asarray(desc, np.uint8)
    if d.ndim == 1:
        d = d.reshape(1, -1)
    if d.size == 0 or d.strides[1] != 1:
        d = np.ascontiguousarray(d)
    step = d.strides[0] if d.shape[0] > 1 else d.shape[1]
    n = _chk(lib().orbx_serialize_matrix_u8(int(text), ptr(d), d.shape[0], d.shape[1], step, None, 0))
    out = np.zeros(n, np.uint8)
    _chk(lib().orbx_serialize_matrix_u8(int(text), ptr(d), d.shape[0], d.shape[1], step, ptr(out), n))
    return out.tobytes()


def deserialize_descriptors(buf, text=False):
    """Returns (matrix, bytes consumed)."""
    src = np.frombuffer(buf, np.uint8)
    rows, cols = C.c_int(0), C.c_int(0)
    _chk(lib().orbx_deserialize_matrix_u8(int(text), ptr(src), len(src), C.byref(rows), C.byref(cols), None, 0, 0))
    out = np.zeros((rows.value, cols.value), np.uint8)
    used = _chk(lib().orbx_deserialize_matrix_u8(int(text), ptr(src), len(src), C.byref(rows), C.byref(cols), ptr(out), max(cols.value, 1),
                                                 out.size))
    return out, used


def serialize_keypoints(kps, text=False):
    k = np.ascontiguousarray(kps, KP_DTYPE)
    n = _chk(lib().orbx_serialize_keypoints(int(text), ptr(k), len(k), None, 0))
    out = np.zeros(n, np.uint8)
    _chk(lib().orbx_serialize_keypoints(int(text), ptr(k), len(k), ptr(out), n))
    return out.tobytes()


def deserialize_keypoints(buf, text=False):
    src = np.frombuffer(buf, np.uint8)
    n = C.c_int(0)
    _chk(lib().orbx_deserialize_keypoints(int(text), ptr(src), len(src), C.byref(n), None, 0))
    out = np.zeros(n.value, KP_DTYPE)
    used = _chk(lib().orbx_deserialize_keypoints(int(text), ptr(src), len(src), C.byref(n), ptr(out), n.value))
    return out, used
