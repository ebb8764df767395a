"""CUDA path (through the C ABI) directly against the REFERENCE's own compiled functions (oracle/_ref/libref.so, built in the
dev container by oracle/build_ref.sh and shipped with the snapshot) — no oracle code between the two, except where an
OpenCV-owned primitive has to produce the input.  Skipped when the library did not travel."""
import numpy as np
import pytest

import wut_cuda_orb_slam3_b200 as orbx
from tests import bow_synth, ref_cases, ref_lib
from tests.proj_synth import SCALE, make_frame, make_points
from wut_cuda_orb_slam3_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref/libref.so not present")]


@pytest.fixture(scope="module")
def ref():
    return ref_lib.load()


def test_tables_vs_reference_ctor(ref):
    for nf, sf, nl in [(1000, 1.2, 8), (1200, 1.2, 8), (2000, 1.2, 8), (500, 1.5, 5)]:
        t = orbx.compute_tables(nf, sf, nl); r = ref.tables(nf, sf, nl)
        for k in ("scale", "inv", "sigma2", "invsigma2"):
            assert t[k].tobytes() == r[k].tobytes()
        assert np.array_equal(t["nfeat"], r["nfeat"])


def test_octree_kernel_vs_reference_distribute_octree(ref):
    for case in range(120):
        xs, ys, sc, w, h, N = ref_cases.octree_inputs(case)
        idx = orbx.distribute_octree(xs, ys, sc, 16, 16 + w, 16, 16 + h, N)
        rx, ry, rs = ref.distribute_octree(xs, ys, sc, 16, 16 + w, 16, 16 + h, N)
        assert len(idx) == len(rx) and np.array_equal(xs[idx], rx) and np.array_equal(ys[idx], ry) and np.array_equal(sc[idx], rs), case


@pytest.mark.parametrize("cols,rows,seed", [(752, 480, 1), (331, 277, 7)])
def test_fast_cells_vs_reference_tile_calc(ref, cols, rows, seed):
    img = synth.image(seed, cols, rows)
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    ex(img, None, (0, 0))
    x, y, s = ex.candidates(0)
    rx, ry, rs = ref.tile_calc_keypoints(img, 1000, 20, 7)
    assert np.array_equal(x, rx) and np.array_equal(y, ry) and np.array_equal(s, rs)


def test_extract_levels_vs_reference_octree_and_descriptor(ref):
    """Whole extraction: per level, the retained key points must be DistributeOctTree(tileCalcKeypoints(level)) and the
    descriptors computeOrbDescriptor on the blurred level, all by the reference's code."""
    img = synth.image(5, 752, 480)
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    ex(img, None, (0, 0))
    nfeat = orbx.compute_tables(1000, 1.2, 8)["nfeat"]
    same = total = 0
    for l in range(8):
        lvl = ex.pyramid_level(l)
        h, w = lvl.shape
        rx, ry, rs = ref.tile_calc_keypoints(lvl, 1000, 20, 7)
        ox, oy, osc = ref.distribute_octree(rx, ry, rs, 16, w - 16, 16, h - 16, int(nfeat[l]), l)
        kps, desc = ex.level_keypoints(l)
        assert len(kps) == len(ox)
        assert np.array_equal(kps["x"], (ox + 16).astype(np.float32)) and np.array_equal(kps["y"], (oy + 16).astype(np.float32))
        assert np.array_equal(kps["response"], osc.astype(np.float32))
        blurred = ex.blurred_level(l)
        for i in range(0, len(kps), 3):
            d = ref.descriptor(blurred, int(kps["x"][i]), int(kps["y"][i]), float(kps["angle"][i]))
            same += int(np.array_equal(d, desc[i])); total += 1
    assert same == total            # same angle in, same bits out


def test_stereo_vs_reference_compute_stereo_matches(ref):
    left = synth.image(31, 752, 480, view=0); right = synth.image(31, 752, 480, view=1)
    exL = orbx.ORBextractor(1200, 1.2, 8, 20, 7); exR = orbx.ORBextractor(1200, 1.2, 8, 20, 7)
    _, kL, dL = exL(left, None, (0, 0)); _, kR, dR = exR(right, None, (0, 0))
    mbf, mb = np.float32(47.9), np.float32(0.11)
    u, d = orbx.compute_stereo_matches(exL, exR, kL, dL, kR, dR, mbf, np.float32(mbf / mb))
    pyrL = [exL.pyramid_level(l, with_border=True) for l in range(8)]; pyrR = [exR.pyramid_level(l, with_border=True) for l in range(8)]
    t = orbx.compute_tables(1200, 1.2, 8)
    ru, rd = ref.compute_stereo_matches(kL, dL, kR, dR, pyrL, pyrR, t["scale"], t["inv"], mb, mbf)
    assert (ru >= 0).sum() > 50 and u.tobytes() == ru.tobytes() and d.tobytes() == rd.tobytes()


def test_projection_searches_vs_reference(ref):
    rng = np.random.default_rng(42)
    kp, desc, ur, occ, bounds = make_frame(rng, 900, crowd=8)
    P = make_points(rng, kp, desc, ur, 2200, dup_frac=0.6, max_flip=100)
    fv = orbx.FrameView(kp, desc, SCALE, bounds, u_right=ur, occupied=occ)
    bg = fv.bounds_grid()
    got, nm = orbx.search_by_projection_map(fv, P["in_view"], P["bad"], P["x"], P["y"], P["xr"], P["view_cos"], P["depth"], P["level"], P["n_obs"],
                                            P["desc"], th=4.0, far_points=True, th_far_points=25.0, nnratio=0.8)
    want, wnm = ref.search_by_projection_map(kp, desc, ur, occ, bg, SCALE, P["in_view"], P["bad"], P["x"], P["y"], P["xr"], P["view_cos"], P["depth"],
                                             P["level"], P["n_obs"], P["desc"], th=4.0, far=True, th_far=25.0, nnratio=0.8)
    assert nm == wnm and nm > 50 and np.array_equal(got, want)
    for tz in (0.0, 2.0, -2.0):
        want, wnm, invz = ref.search_by_projection_last(kp, desc, ur, occ, bg, SCALE, 40.0, P["valid"], P["x"], P["y"], P["depth"], P["level"],
                                                        P["angle"], P["n_obs"], P["desc"], 15.0, last_tz=tz)
        got, nm = orbx.search_by_projection_last(fv, 40.0, P["valid"], P["x"], P["y"], invz, P["level"], P["angle"], P["n_obs"], P["desc"], 15.0,
                                                 forward=tz > 1.0, backward=-tz > 1.0)
        assert nm == wnm and nm > 50 and np.array_equal(got, want)
    found = (rng.random(len(P["x"])) < 0.1).astype(np.uint8)
    want, wnm, d3 = ref.search_by_projection_kf(kp, desc, occ, bg, SCALE, P["valid"], found, P["x"], P["y"], P["depth"], P["min_dist"] * 30,
                                                P["max_dist"] * 30, P["level"], P["angle"], P["desc"], 10.0, 100)
    fv2 = orbx.FrameView(kp, desc, SCALE, bounds, occupied=occ)
    got, nm = orbx.search_by_projection_kf(fv2, P["valid"] & (1 - found), P["x"], P["y"], d3, P["min_dist"] * 30, P["max_dist"] * 30, P["level"],
                                           P["angle"], P["desc"], 10.0, 100)
    assert nm == wnm and nm > 50 and np.array_equal(got, want)


def test_search_by_bow_vs_reference(ref, oracle):
    k, L = 10, 4
    parent, vdesc, weights = bow_synth.make_vocab(601, k, L)
    voc = orbx.ORBVocabulary(parent, vdesc, weights, k, L)
    B = bow_synth.make_pair(602, vdesc, parent, 1500, 1400)
    _, fva = voc.transform(B["desc_a"], 4); _, fvb = voc.transform(B["desc_b"], 4)
    t = lambda fv: (fv.node_ids, fv.offsets, fv.indices)   # noqa: E731
    mb, ma, nm = orbx.search_by_bow(B["desc_a"], B["angle_a"], B["valid_a"], fva, B["desc_b"], B["angle_b"], fvb, nnratio=0.7)
    want, wnm = ref.search_by_bow_kf_frame(B["desc_a"], B["angle_a"], B["valid_a"], t(fva), B["desc_b"], B["angle_b"], t(fvb), -1, 0.7, True)
    assert nm == wnm and nm > 50 and np.array_equal(mb, want)
    mb, ma, nm = orbx.search_by_bow(B["desc_a"], B["angle_a"], B["valid_a"], fva, B["desc_b"], B["angle_b"], fvb, valid_b=B["valid_b"], kf_kf=True,
                                    nnratio=0.8)
    want, wnm = ref.search_by_bow_kf_kf(B["desc_a"], B["angle_a"], B["valid_a"], t(fva), B["desc_b"], B["angle_b"], B["valid_b"], t(fvb), 0.8, True)
    assert nm == wnm and nm > 50 and np.array_equal(ma, want)


def test_bow_transform_vs_dbow2(ref, tmp_path):
    """The CUDA vocabulary descent + host BowVector / FeatureVector assembly against DBoW2's own compiled transform, both
    loading the same ORBvoc-format file (no trailing newline: see tests/test_ref_pin.py)."""
    k, L = 10, 4
    parent, vdesc, weights = bow_synth.make_vocab(611, k, L)
    path = tmp_path / "voc.txt"
    bow_synth.write_vocab_text(path, parent, vdesc, weights, k, L, 0, 0)
    txt = open(path).read().rstrip()
    open(path, "w").write(txt)
    rv = ref_lib.RefVocabulary(ref, path)
    voc = orbx.ORBVocabulary.loadFromTextFile(path)
    feats = bow_synth.make_features(612, vdesc, parent, 3000)
    for a, b in zip(voc.transform_features(feats, 4), rv.transform_features(feats, 4)):
        assert np.array_equal(a, b)
    (ids, vals), fv = voc.transform(feats, 4)
    (rid, rval), rfv = rv.transform(feats, 4)
    assert np.array_equal(ids, rid) and vals.tobytes() == rval.tobytes()
    for a, b in zip((fv.node_ids, fv.offsets, fv.indices), rfv):
        assert np.array_equal(a, b)
    (ids2, vals2), _ = voc.transform(bow_synth.make_features(613, vdesc, parent, 2500), 4)
    assert voc.score((ids, vals), (ids2, vals2)) == rv.score((rid, rval), (ids2, vals2))
