"""The C++ drop-in adapter (csrc/adapter/ORBextractor.h: same class/signatures as the reference's include/ORBextractor.h)
compiled against a cv stub, run as Frame::ExtractORB would, must produce exactly what the C ABI returns through ctypes."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "adapter_test.cpp")
EXE = os.path.join(ROOT, "build", "adapter_test")


def build_exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "tests", "cpp", "cv_stub"), "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(ROOT, "wut_cuda_orb_slam3_b200", "csrc", "adapter"), "-o", EXE, SRC,
                           "-L" + os.path.join(ROOT, "wut_cuda_orb_slam3_b200"), "-lorbx",
                           "-Wl,-rpath," + os.path.join(ROOT, "wut_cuda_orb_slam3_b200")])


def fnv(h, data):
    for b in bytes(data):
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_adapter_compiles_against_header():
    """CPU: the adapter + include/orbx.h compile and link (no compute call)."""
    build_exe()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_adapter_matches_c_abi():
    import wut_cuda_orb_slam3_b200 as orbx
    from wut_cuda_orb_slam3_b200 import synth
    build_exe()
    out = subprocess.check_output([EXE, "3"], text=True)
    m = re.search(r"mono=(-?\d+) n=(\d+) levels=(\d+) scale1=([\d.]+) kphash=([0-9a-f]+) pyrhash=([0-9a-f]+) empty=(-?\d+) dist01=(-?\d+)", out)
    assert m, out
    img = synth.image(3, 752, 480)
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    nm, kps, desc = ex(img, None, (0, 1000))
    h = fnv(1469598103934665603, kps.tobytes())
    h = fnv(h, desc.tobytes())
    hp = 1469598103934665603
    for l in range(8):
        hp = fnv(hp, ex.pyramid_level(l).tobytes())
    assert int(m.group(1)) == nm and int(m.group(2)) == len(kps) and int(m.group(3)) == 8
    assert abs(float(m.group(4)) - 1.2000000477) < 1e-7
    assert m.group(5) == "%016x" % h
    assert m.group(6) == "%016x" % hp
    assert int(m.group(7)) == -1                      # empty image -> -1, like the reference
    assert int(m.group(8)) == orbx.ORBmatcher.DescriptorDistance(desc[0], desc[1])


MSRC = os.path.join(ROOT, "tests", "cpp", "matcher_adapter_test.cpp")
MEXE = os.path.join(ROOT, "build", "matcher_adapter_test")


def build_matcher_exe():
    os.makedirs(os.path.dirname(MEXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "tests", "cpp", "cv_stub"), "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(ROOT, "wut_cuda_orb_slam3_b200", "csrc", "adapter"), "-o", MEXE, MSRC,
                           "-L" + os.path.join(ROOT, "wut_cuda_orb_slam3_b200"), "-lorbx",
                           "-Wl,-rpath," + os.path.join(ROOT, "wut_cuda_orb_slam3_b200")])


def test_matcher_adapter_compiles():
    """CPU: csrc/adapter/ORBmatcherProjection.h compiles against Frame / MapPoint stand-ins and links (no compute call)."""
    build_matcher_exe()
    assert os.path.exists(MEXE)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
def test_matcher_adapter_matches_oracle(oracle, tmp_path, mode):
    """The C++ adapter, driven like Tracking::SearchLocalPoints / TrackWithMotionModel, against the oracle's sequential loop."""
    from tests.proj_synth import SCALE, make_frame, make_points
    build_matcher_exe()
    rng = np.random.default_rng(500 + mode)
    kp, desc, ur, occ, bounds = make_frame(rng, 900, crowd=6)
    P = make_points(rng, kp, desc, ur, 1500, dup_frac=0.5, max_flip=90)
    b = [np.float32(v) for v in bounds]
    bg = np.array(b + [np.float32(64) / (b[2] - b[0]), np.float32(48) / (b[3] - b[1])], np.float32)
    th, ratio, mbf = 3.0, 0.8, 40.0
    fl = np.array(list(bg) + [15.0 if mode else th, mbf if mode else ratio, 1.0, 20.0], np.float32)
    scene = tmp_path / "scene.bin"; result = tmp_path / "result.bin"
    with open(scene, "wb") as f:
        f.write(np.array([len(kp), len(P["x"]), len(SCALE), mode], np.int32).tobytes()); f.write(fl.tobytes())
        f.write(kp.tobytes()); f.write(desc.tobytes()); f.write(ur.tobytes()); f.write(occ.tobytes()); f.write(SCALE.tobytes())
        fa, fb = (P["valid"], np.zeros_like(P["valid"])) if mode else (P["in_view"], P["bad"])
        f.write(fa.tobytes()); f.write(fb.tobytes())
        for a in (P["x"], P["y"], P["xr"], P["view_cos"], P["invz"] if mode else P["depth"], P["angle"]):
            f.write(np.ascontiguousarray(a, np.float32).tobytes())
        f.write(P["level"].astype(np.int32).tobytes()); f.write(P["n_obs"].astype(np.int32).tobytes()); f.write(P["desc"].tobytes())
    subprocess.check_call([MEXE, str(scene), str(result)])
    raw = np.fromfile(result, np.int32)
    nm, got = int(raw[0]), raw[1:]
    if mode == 0:
        want, wnm = oracle.search_by_projection_map(kp, desc, ur, occ, bg, SCALE, P["in_view"], P["bad"], P["x"], P["y"], P["xr"], P["view_cos"],
                                                    P["depth"], P["level"], P["n_obs"], P["desc"], th=th, far=True, th_far=20.0, nnratio=ratio)
    else:
        want, wnm = oracle.search_by_projection_last(kp, desc, ur, occ, bg, SCALE, mbf, P["valid"], P["x"], P["y"], P["invz"], P["level"],
                                                     P["angle"], P["n_obs"], P["desc"], 15.0, False, False, True)
    assert nm == wnm and nm > 50 and np.array_equal(got, want)


CSRC = os.path.join(ROOT, "tests", "cpp", "orbmatcher_class_test.cpp")
CEXE = os.path.join(ROOT, "build", "orbmatcher_class_test")


def build_class_exe():
    os.makedirs(os.path.dirname(CEXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "tests", "cpp", "cv_stub"), "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(ROOT, "wut_cuda_orb_slam3_b200", "csrc", "adapter"), "-o", CEXE, CSRC,
                           "-L" + os.path.join(ROOT, "wut_cuda_orb_slam3_b200"), "-lorbx",
                           "-Wl,-rpath," + os.path.join(ROOT, "wut_cuda_orb_slam3_b200")])


def test_orbmatcher_class_compiles():
    """CPU: csrc/adapter/ORBmatcher.h (class ORBmatcher with the reference's declarations, include/ORBmatcher.h:36-60) compiles
    against Frame / KeyFrame / MapPoint stand-ins, all four search overloads instantiated, and links."""
    build_class_exe()
    assert os.path.exists(CEXE)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
def test_orbmatcher_class_search_by_bow(oracle, tmp_path, mode):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) / (KeyFrame*, KeyFrame*, ...) through the C++ class (std::map feature
    vectors, MapPoint* outputs) against the oracle's restatement of src/ORBmatcher1.cc:225-427 / src/ORBmatcher2.cc:36-171."""
    import wut_cuda_orb_slam3_b200 as orbx
    from tests import bow_synth
    build_class_exe()
    parent, desc, weights = bow_synth.make_vocab(11, 10, 4)
    voc = orbx.ORBVocabulary(parent, desc, weights, 10, 4)
    P = bow_synth.make_pair(1300 + mode, desc, parent, 1500, 1600)
    _, fva = voc.transform(P["desc_a"], 2)
    _, fvb = voc.transform(P["desc_b"], 2)
    ratio, check_ori = 0.75, True
    rng = np.random.default_rng(3)
    # bit 0: the key frame holds a map point for the feature, bit 1: that map point isBad()
    va = (P["valid_a"] | ((rng.random(len(P["valid_a"])) < 0.05).astype(np.uint8) << 1)).astype(np.uint8)
    vb = (P["valid_b"] | ((rng.random(len(P["valid_b"])) < 0.05).astype(np.uint8) << 1)).astype(np.uint8)
    ok_a = ((va & 1) != 0) & ((va & 2) == 0); ok_b = ((vb & 1) != 0) & ((vb & 2) == 0)
    scene = tmp_path / "scene.bin"; result = tmp_path / "result.bin"

    def fv_bytes(fv):
        nodes, off, idx = fv.node_ids, fv.offsets, fv.indices
        return np.array([len(nodes)], np.int32).tobytes() + nodes.astype(np.uint32).tobytes() + off.astype(np.int32).tobytes() + idx.astype(np.uint32).tobytes()
    with open(scene, "wb") as f:
        f.write(np.array([mode, len(P["desc_a"]), len(P["desc_b"]), int(check_ori)], np.int32).tobytes())
        f.write(np.array([ratio], np.float32).tobytes())
        f.write(P["desc_a"].tobytes()); f.write(P["desc_b"].tobytes())
        f.write(P["angle_a"].astype(np.float32).tobytes()); f.write(P["angle_b"].astype(np.float32).tobytes())
        f.write(va.tobytes()); f.write(vb.tobytes())
        f.write(fv_bytes(fva)); f.write(fv_bytes(fvb))
    subprocess.check_call([CEXE, str(scene), str(result), "instantiate"])
    raw = np.fromfile(result, np.int32)
    nm, d01, got = int(raw[0]), int(raw[1]), raw[2:]
    fv_t = lambda fv: (fv.node_ids, fv.offsets, fv.indices)
    if mode == 0:
        want, wnm = oracle.search_by_bow_kf_frame(P["desc_a"], P["angle_a"], ok_a.astype(np.uint8), fv_t(fva), P["desc_b"], P["angle_b"], fv_t(fvb),
                                                  -1, ratio, check_ori)
    else:
        want, wnm = oracle.search_by_bow_kf_kf(P["desc_a"], P["angle_a"], ok_a.astype(np.uint8), fv_t(fva), P["desc_b"], P["angle_b"],
                                               ok_b.astype(np.uint8), fv_t(fvb), ratio, check_ori)
    assert nm == wnm and nm > 50 and np.array_equal(got, want)
    assert d01 == oracle.hamming(P["desc_a"][0], P["desc_a"][1], swar=True)
