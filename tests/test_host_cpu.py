"""CPU-only checks of the host logic and the C-ABI surface (no GPU, no compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import wut_cuda_orb_slam3_b200 as orbx
from wut_cuda_orb_slam3_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    """Every function declared in include/orbx.h is exported by liborbx.so and bound in capi.SIGNATURES."""
    hdr = open(os.path.join(ROOT, "include", "orbx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(orbx_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    L = C.CDLL(capi.SO_PATH)
    for n in names:
        assert hasattr(L, n), n
    assert names == set(capi.SIGNATURES), names ^ set(capi.SIGNATURES)


def test_keypoint_layout_is_cv_keypoint():
    assert capi.KP_DTYPE.itemsize == 28
    assert [capi.KP_DTYPE.fields[f][1] for f in ("x", "y", "size", "angle", "response", "octave", "class_id")] == [0, 4, 8, 12, 16, 20, 24]


@pytest.mark.parametrize("nf,sf,nl", [(1000, 1.2, 8), (1200, 1.2, 8), (2000, 1.2, 8), (5000, 1.2, 8), (800, 1.5, 4), (600, 1.1, 12)])
def test_tables_match_oracle(oracle, nf, sf, nl):
    t = orbx.compute_tables(nf, sf, nl)
    o = oracle.tables(nf, sf, nl)
    for k in ("scale", "inv", "sigma2", "invsigma2", "nfeat"):
        assert np.array_equal(t[k], o[k]), k
    if (nf, sf, nl) == (1000, 1.2, 8):
        assert t["nfeat"].tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    if (nf, sf, nl) == (2000, 1.2, 8):
        assert t["nfeat"].tolist() == [434, 362, 302, 251, 209, 175, 145, 122]


def test_invalid_arguments_are_reported_not_fatal():
    L = capi.lib()
    assert L.orbx_compute_tables(1000, 1.2, 0, None, None, None, None, None) == capi.ORBX_ERR_INVALID_ARG
    assert b"bad extractor parameters" in L.orbx_last_error()
    assert L.orbx_compute_tables(1000, 0.9, 8, None, None, None, None, None) == capi.ORBX_ERR_INVALID_ARG


def test_no_device_is_an_error_not_a_fallback():
    if capi.lib().orbx_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(orbx.OrbxError) as e:
        orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    assert e.value.code == capi.ORBX_ERR_NO_DEVICE
    with pytest.raises(orbx.OrbxError):
        orbx.ORBmatcher().knn2(np.zeros((2, 32), np.uint8), np.zeros((2, 32), np.uint8))


def test_host_descriptor_distance(oracle):
    rng = np.random.default_rng(0)
    for _ in range(100):
        a = rng.integers(0, 256, 32, dtype=np.uint8); b = rng.integers(0, 256, 32, dtype=np.uint8)
        d = orbx.ORBmatcher.DescriptorDistance(a, b)
        assert d == oracle.hamming(a, b, swar=True) == oracle.hamming(a, b) == int(np.unpackbits(a ^ b).sum())
    z = np.zeros(32, np.uint8)
    assert orbx.ORBmatcher.DescriptorDistance(z, z) == 0 and orbx.ORBmatcher.DescriptorDistance(z, ~z) == 256


def test_ratio_test_modes(oracle):
    m = orbx.ORBmatcher(0.6)
    dist = np.array([[10, 20], [30, 40], [50, 90], [51, 200], [0, 0], [49, 82]], np.int32)
    for mode, th in ((0, 50), (1, 50)):
        acc = m.ratio_test(dist, mode=mode, th_low=th)
        for i, (d1, d2) in enumerate(dist):
            exp = (d1 <= th if mode == 0 else d1 < th) and np.float32(d1) < np.float32(0.6) * np.float32(d2)
            assert acc[i] == bool(exp)


def test_synth_generator_is_deterministic_and_textured():
    a = synth.image(5, 752, 480); b = synth.image(5, 752, 480); c = synth.image(6, 752, 480)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert 20 < a.std() < 80
    l, r = synth.image(9, 320, 240, view=0), synth.image(9, 320, 240, view=1)
    assert not np.array_equal(l, r)
    d = synth.descriptors(1, 100); q = synth.descriptors(1, 100, is_query=True, ndb=100, plant_every=2)
    assert d.shape == (100, 32) and q.shape == (100, 32)
    bits = np.unpackbits(d).mean()
    assert 0.45 < bits < 0.55


def test_rotation_consistency_matches_oracle(oracle):
    rng = np.random.default_rng(1)
    for n in (0, 1, 5, 200, 3000):
        a = rng.uniform(0, 360, n).astype(np.float32)
        b = (a - rng.choice([0.0, 30.0, 61.0, 200.0], n, p=[0.6, 0.25, 0.1, 0.05]) + rng.normal(0, 3, n)).astype(np.float32) % np.float32(360)
        keep = orbx.rotation_consistency(a, b)
        assert np.array_equal(keep, oracle.rotation_consistency(a, b))
        if n >= 200:
            assert keep.any() and not keep.all()


def test_distinctive_descriptor_matches_oracle(oracle):
    rng = np.random.default_rng(2)
    for n in (1, 2, 3, 8, 25):
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        d = np.stack([base ^ np.packbits(rng.random(256) < p) for p in rng.uniform(0.0, 0.2, n)])
        assert orbx.distinctive_descriptor(d) == oracle.distinctive_descriptor(d)
