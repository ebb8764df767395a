"""Known-answer cases shared by tests/golden/make_ref_golden.py (run against the REFERENCE's compiled code, oracle/_ref) and
tests/test_golden.py (run against the oracle, anywhere): seeded inputs -> CRC32 of the outputs."""
import zlib

import numpy as np

from tests import bow_synth
from tests.proj_synth import SCALE, make_frame, make_points
from wut_cuda_orb_slam3_b200 import synth


def crc(*arrays):
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    return c & 0xFFFFFFFF


def bounds_grid(bounds):
    b = [np.float32(x) for x in bounds]
    return np.array(b + [np.float32(64) / (b[2] - b[0]), np.float32(48) / (b[3] - b[1])], np.float32)


def octree_inputs(case):
    rng = np.random.default_rng(1000 + case)
    w, h = [(720, 448), (595, 368), (178, 102), (1209, 344)][case % 4]
    n = int(rng.integers(50, 3000))
    if case % 3 == 2:
        xs = (rng.integers(0, w // 8, n) * 8) % w; ys = (rng.integers(0, h // 8, n) * 8) % h
    else:
        xs = rng.integers(0, w, n); ys = rng.integers(0, h, n)
    pts = np.unique(np.stack([ys, xs], 1), axis=0)
    sc = rng.integers(7, 60, len(pts))
    return pts[:, 1].astype(np.int32), pts[:, 0].astype(np.int32), sc.astype(np.int32), w, h, int(rng.choice([217, 60, 434, 1000]))


def run_all(oracle, ref=None):
    """Returns {name: crc}.  With ref=None every case runs on the oracle; otherwise on the reference library (the oracle still
    supplies the OpenCV-owned inputs: images, blurred levels, extracted key points)."""
    out = {}
    use_ref = ref is not None
    # constructor tables
    for nf, sf, nl in [(1000, 1.2, 8), (2000, 1.2, 8), (500, 1.5, 5)]:
        t = (ref if use_ref else oracle).tables(nf, sf, nl)
        out["tables_%d_%g_%d" % (nf, sf, nl)] = crc(t["scale"], t["inv"], t["sigma2"], t["invsigma2"], t["nfeat"], t["umax"])
    # DistributeOctTree
    for case in range(40):
        xs, ys, sc, w, h, N = octree_inputs(case)
        if use_ref:
            rx, ry, rs = ref.distribute_octree(xs, ys, sc, 16, 16 + w, 16, 16 + h, N)
        else:
            idx = oracle.octree(xs, ys, sc, 16, 16 + w, 16, 16 + h, N)
            rx, ry, rs = xs[idx], ys[idx], sc[idx]
        out["octree_%d" % case] = crc(rx, ry, rs)
    # the CPU cell loop
    for cols, rows, seed in [(752, 480, 1), (210, 134, 3), (331, 277, 7)]:
        img = synth.image(seed, cols, rows)
        r = ref.tile_calc_keypoints(img, 1000, 20, 7) if use_ref else oracle.cell_fast(img, 20, 7)
        out["cells_%dx%d" % (cols, rows)] = crc(*r)
    # computeOrbDescriptor
    rng = np.random.default_rng(9)
    blurred = oracle.blur(synth.image(11, 400, 300))
    ds = []
    for _ in range(300):
        x, y, ang = int(rng.integers(19, 381)), int(rng.integers(19, 281)), float(np.float32(rng.uniform(0, 360)))
        ds.append((ref if use_ref else oracle).descriptor(blurred, x, y, ang))
    out["descriptors"] = crc(np.stack(ds))
    # operator() placement
    img = synth.image(21, 752, 480)
    ex = oracle.extractor(1000, 1.2, 8, 20, 7)
    kps, desc, n_mono = ex.extract(img, (300, 450))
    if use_ref:
        lv = [ex.level_keypoints(l) for l in range(8)]
        kps, desc, n_mono = ref.pack([k for k, _ in lv], [d for _, d in lv], ex.tables["scale"], (300, 450), len(kps))
    out["pack"] = crc(kps, desc, np.int32(n_mono))
    # ComputeStereoMatches
    left = synth.image(33, 400, 300, view=0); right = synth.image(33, 400, 300, view=1)
    oL = oracle.extractor(500, 1.2, 8, 20, 7); oR = oracle.extractor(500, 1.2, 8, 20, 7)
    kL, dL, _ = oL.extract(left, (0, 0)); kR, dR, _ = oR.extract(right, (0, 0))
    pyrL = [oL.pyramid_level(l, with_border=True) for l in range(8)]; pyrR = [oR.pyramid_level(l, with_border=True) for l in range(8)]
    mbf, mb = np.float32(47.9), np.float32(0.11)
    if use_ref:
        u, d = ref.compute_stereo_matches(kL, dL, kR, dR, pyrL, pyrR, oL.tables["scale"], oL.tables["inv"], mb, mbf)
    else:
        u, d, _ = oracle.stereo_match_raw(kL, dL, kR, dR, pyrL, pyrR, oL.tables["scale"], oL.tables["inv"], mbf, np.float32(mbf / mb))
    out["stereo"] = crc(u, d)
    # frame grid + SearchByProjection (Frame, MapPoints)
    rng = np.random.default_rng(42)
    kp, desc, ur, occ, bounds = make_frame(rng, 800, crowd=10)
    bg = bounds_grid(bounds)
    P = make_points(rng, kp, desc, ur, 2000, dup_frac=0.7, max_flip=100)
    impl = ref if use_ref else oracle
    out["grid"] = crc(*impl.assign_features_to_grid(kp, bg))
    m, nm = impl.search_by_projection_map(kp, desc, ur, occ, bg, SCALE, P["in_view"], P["bad"], P["x"], P["y"], P["xr"], P["view_cos"], P["depth"],
                                          P["level"], P["n_obs"], P["desc"], th=5.0, far=True, th_far=20.0, nnratio=0.8)
    out["projection_map"] = crc(m, np.int32(nm))
    # SearchByBoW
    k, L = 4, 3
    parent, vdesc, weights = bow_synth.make_vocab(71, k, L)
    voc = oracle.vocabulary(parent, vdesc, weights, k, L)
    B = bow_synth.make_pair(171, vdesc, parent, 600, 560)
    _, fva = voc.transform(B["desc_a"], 2); _, fvb = voc.transform(B["desc_b"], 2)
    m, nm = impl.search_by_bow_kf_frame(B["desc_a"], B["angle_a"], B["valid_a"], fva, B["desc_b"], B["angle_b"], fvb, -1, 0.75, True)
    out["bow_kf_frame"] = crc(m, np.int32(nm))
    m, nm = impl.search_by_bow_kf_kf(B["desc_a"], B["angle_a"], B["valid_a"], fva, B["desc_b"], B["angle_b"], B["valid_b"], fvb, 0.8, True)
    out["bow_kf_kf"] = crc(m, np.int32(nm))
    # SearchForTriangulation (F12 / epipole supplied)
    from tests import test_bow_cpu
    kpa, kpb, F, ep, scale, sigma2, fa, fb, sa, sb = test_bow_cpu.tri_inputs(B, 91)
    m, nm = impl.search_for_triangulation(kpa, B["desc_a"], fa, sa, fva, kpb, B["desc_b"], fb, sb, fvb, F, ep, scale, sigma2, False, False, True)
    out["triangulation"] = crc(m, np.int32(nm))
    # ComputeDistinctiveDescriptors
    rng = np.random.default_rng(17)
    best = []
    for n in (1, 2, 5, 20, 61):
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        d = np.stack([base ^ (rng.integers(0, 256, 32, dtype=np.uint8) & rng.integers(0, 256, 32, dtype=np.uint8)) for _ in range(n)])
        best.append(impl.distinctive_descriptor(d))
    out["distinctive"] = crc(np.array(best, np.int32))
    # DBoW2 transform + L1 score (the reference side loads the vocabulary with DBoW2's own loadFromTextFile)
    import os
    import tempfile
    if use_ref:
        from tests import ref_lib
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "voc.txt")
            bow_synth.write_vocab_text(path, parent, vdesc, weights, k, L, 0, 0)
            txt = open(path).read().rstrip()
            open(path, "w").write(txt)
            v = ref_lib.RefVocabulary(ref, path)
            (ids, vals), fv = v.transform(B["desc_a"], 2)
            (ids2, vals2), _ = v.transform(B["desc_b"], 2)
            sc = v.score((ids, vals), (ids2, vals2))
    else:
        (ids, vals), fv = voc.transform(B["desc_a"], 2)
        (ids2, vals2), _ = voc.transform(B["desc_b"], 2)
        sc = oracle.bow_score_l1((ids, vals), (ids2, vals2))
    out["dbow2_transform"] = crc(ids, vals, *fv, np.float64(sc))
    return out
