"""SURVEY.md §8(f)4, second half: descriptor rows and key points as serializeMatrix / serializeVectorKeyPoints
(reference include/SerializationUtils.h:76-153) put them into the boost archives of System::SaveAtlas (src/System.cc:1339-1475).
Host code, no GPU.  boost is not in this image, so the streams are checked against hand-built expectations of boost's
primitive encoding (binary archive: native little-endian bytes; text archive: one space before every token, unsigned char as a
decimal number, float with 9 significant digits) and by round trips."""
import struct

import numpy as np
import pytest

import wut_cuda_orb_slam3_b200 as orbx
from wut_cuda_orb_slam3_b200 import io, synth
from wut_cuda_orb_slam3_b200.capi import KP_DTYPE, OrbxError


def test_binary_matrix_stream_layout():
    d = synth.descriptors(3, 5)
    s = io.serialize_descriptors(d, text=False)
    # int cols, int rows, int type (CV_8UC1 = 0), bool continuous, then rows * cols raw bytes (SerializationUtils.h:79-92)
    assert s == struct.pack("<iii?", 32, 5, 0, True) + d.tobytes()
    back, used = io.deserialize_descriptors(s, text=False)
    assert used == len(s) and np.array_equal(back, d)


def test_text_matrix_stream_layout():
    d = np.array([[0, 7, 255], [16, 1, 200]], np.uint8)
    s = io.serialize_descriptors(d, text=True)
    assert s == b" 3 2 0 1 0 7 255 16 1 200"
    back, used = io.deserialize_descriptors(s, text=True)
    assert used == len(s) and np.array_equal(back, d)
    # boost starts a new line now and then: any white space separates tokens
    back2, _ = io.deserialize_descriptors(b"3 2 0 1\n0 7 255\n16 1 200\n", text=True)
    assert np.array_equal(back2, d)


@pytest.mark.parametrize("text", [False, True])
def test_matrix_round_trip_rows_strided_and_empty(text):
    big = synth.descriptors(4, 40).reshape(20, 64)
    view = big[:, 16:48]                                   # non-continuous: rows are written one by one (:93-98), same bytes
    s = io.serialize_descriptors(view, text=text)
    back, used = io.deserialize_descriptors(s + b" 99", text=text)       # trailing archive content is left alone
    assert used == len(s) and np.array_equal(back, view)
    if not text:
        assert s[12] == 0                                  # continuous flag of the strided view
    one = synth.descriptors(5, 1)                          # MapPoint::mDescriptor: a 1 x 32 row
    b1, _ = io.deserialize_descriptors(io.serialize_descriptors(one, text=text), text=text)
    assert np.array_equal(b1, one)
    e, _ = io.deserialize_descriptors(io.serialize_descriptors(np.zeros((0, 32), np.uint8), text=text), text=text)
    assert e.shape == (0, 32)


def test_keypoint_stream_layout_and_round_trip():
    k = np.zeros(3, KP_DTYPE)
    k["x"] = [10.5, 300.25, 751.0]; k["y"] = [20.0, 7.125, 479.0]; k["size"] = [31, 37, 111]
    k["angle"] = [0.0, 359.98337, 123.456]; k["response"] = [20, 7, 255]; k["octave"] = [0, 1, 7]; k["class_id"] = -1
    s = io.serialize_keypoints(k, text=False)
    # NumEl, then per key point: angle, response, size, pt.x, pt.y, class_id, octave (SerializationUtils.h:140-147)
    exp = struct.pack("<i", 3)
    for r in k:
        exp += struct.pack("<fffffii", r["angle"], r["response"], r["size"], r["x"], r["y"], r["class_id"], r["octave"])
    assert s == exp
    back, used = io.deserialize_keypoints(s, text=False)
    assert used == len(s) and back.tobytes() == k.tobytes()
    t = io.serialize_keypoints(k, text=True)
    assert t.startswith(b" 3 0.000000000e+00 2.000000000e+01 3.100000000e+01 1.050000000e+01 2.000000000e+01 -1 0 ")
    back, used = io.deserialize_keypoints(t, text=True)                   # 9 significant digits restore every float32 exactly
    assert used == len(t) and back.tobytes() == k.tobytes()
    rng = np.random.default_rng(0)
    r = np.zeros(500, KP_DTYPE)
    for f in ("x", "y", "size", "angle", "response"):
        r[f] = rng.random(500).astype(np.float32) * np.float32(1000)
    r["octave"] = rng.integers(0, 8, 500); r["class_id"] = -1
    for text in (False, True):
        b, _ = io.deserialize_keypoints(io.serialize_keypoints(r, text=text), text=text)
        assert b.tobytes() == r.tobytes()


def test_truncated_streams_are_rejected():
    d = synth.descriptors(6, 4)
    for text in (False, True):
        s = io.serialize_descriptors(d, text=text)
        with pytest.raises(OrbxError):
            io.deserialize_descriptors(s[:len(s) // 2], text=text)
        with pytest.raises(OrbxError):
            io.deserialize_descriptors(s[:3], text=text)
    with pytest.raises(OrbxError):
        io.deserialize_descriptors(struct.pack("<iii?", 32, 1, 5, True) + bytes(32), text=False)   # not CV_8U
