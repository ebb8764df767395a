"""GPU parity tests: the CUDA path, called through the C ABI (liborbx.so), against the CPU oracle on the same seeded
inputs.  Bit-exact for pyramid / blur / FAST candidates / octree keypoints / match indices; orientation within 1e-3
degrees; >= 99.5 % of descriptors bit-identical (BASELINE.json north_star)."""
import numpy as np
import pytest

import wut_cuda_orb_slam3_b200 as orbx
from wut_cuda_orb_slam3_b200 import synth

pytestmark = pytest.mark.gpu

ANGLE_TOL_DEG = 1e-3
DESC_IDENTICAL_MIN = 0.995

CONFIGS = [
    # cols, rows, nfeatures, seed        (BASELINE.json configs[0..3] shapes + a small one)
    (752, 480, 1000, 101),
    (752, 480, 1200, 102),
    (1241, 376, 2000, 103),
    (1280, 720, 2000, 104),
    (160, 120, 300, 105),
    (331, 277, 500, 106),
]


def compare_full(oracle, ex, oex, img, lapping):
    nm_o = None
    kps_o, desc_o, nm_o = oex.extract(img, lapping)
    nm, kps, desc = ex(img, None, lapping)
    assert nm == nm_o
    assert len(kps) == len(kps_o)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(kps[f], kps_o[f]), f
    dang = np.abs(kps["angle"].astype(np.float64) - kps_o["angle"].astype(np.float64))
    dang = np.minimum(dang, 360.0 - dang)
    assert dang.max(initial=0.0) <= ANGLE_TOL_DEG
    same = (desc == desc_o).all(axis=1)
    frac = float(same.mean()) if len(same) else 1.0
    assert frac >= DESC_IDENTICAL_MIN, frac
    return kps, desc, frac, float(dang.max(initial=0.0))


@pytest.mark.parametrize("cols,rows,nfeatures,seed", CONFIGS)
def test_extract_stage_by_stage(oracle, cols, rows, nfeatures, seed):
    img = synth.image(seed, cols, rows)
    ex = orbx.ORBextractor(nfeatures, 1.2, 8, 20, 7)
    oex = oracle.extractor(nfeatures, 1.2, 8, 20, 7)
    kps, desc, frac_same, max_dang = compare_full(oracle, ex, oex, img, (0, 0))
    assert len(kps) > nfeatures // 3
    for level in range(8):
        # pyramid (bordered buffer, bit-exact)
        assert np.array_equal(ex.pyramid_level(level, with_border=True), oex.pyramid_level(level, with_border=True)), level
        # blurred level (bit-exact)
        assert np.array_equal(ex.blurred_level(level), oex.blurred_level(level)), level
        # FAST candidates handed to the octree: same list, same order
        xs, ys, sc = ex.candidates(level)
        oxs, oys, osc = oex.candidates(level)
        assert np.array_equal(xs, oxs) and np.array_equal(ys, oys) and np.array_equal(sc, osc), level
        # octree output (+border) in the reference's list order
        lk, ld = ex.level_keypoints(level)
        olk, old = oex.level_keypoints(level)
        assert len(lk) == len(olk)
        for f in ("x", "y", "response", "octave", "size"):
            assert np.array_equal(lk[f], olk[f]), (level, f)
    # this build aims for exact equality, not just the tolerance
    assert max_dang == 0.0
    assert frac_same == 1.0


@pytest.mark.parametrize("lapping", [(0, 1000), (300, 500), (0, 0), (2000, 3000)])
def test_lapping_area_packing(oracle, lapping):
    img = synth.image(7, 752, 480)
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    oex = oracle.extractor(1000, 1.2, 8, 20, 7)
    compare_full(oracle, ex, oex, img, lapping)


def test_mono_ini_extractor_5x_features(oracle):
    img = synth.image(8, 752, 480)
    ex = orbx.ORBextractor(5000, 1.2, 8, 20, 7)     # Tracking uses 5*nFeatures until initialised (src/Tracking2.cc:413-416)
    oex = oracle.extractor(5000, 1.2, 8, 20, 7)
    compare_full(oracle, ex, oex, img, (0, 1000))


def test_other_parameters(oracle):
    img = synth.image(9, 640, 480)
    for (nf, sf, nl, ini, mn) in [(800, 1.5, 4, 30, 10), (600, 1.1, 12, 15, 5), (400, 2.0, 3, 20, 7)]:
        ex = orbx.ORBextractor(nf, sf, nl, ini, mn)
        oex = oracle.extractor(nf, sf, nl, ini, mn)
        compare_full(oracle, ex, oex, img, (0, 0))


@pytest.mark.parametrize("cols,rows,nfeatures,nlevels", [(1920, 1080, 3000, 8), (480, 752, 1000, 8), (97, 83, 200, 8), (64, 64, 100, 5),
                                                         (2048, 1536, 4000, 10)])
def test_unusual_shapes(oracle, cols, rows, nfeatures, nlevels):
    """Full-HD and larger, portrait (nIni = 1), and images so small that the deep levels have no FAST window at all."""
    img = synth.image(55, cols, rows)
    ex = orbx.ORBextractor(nfeatures, 1.2, nlevels, 20, 7)
    oex = oracle.extractor(nfeatures, 1.2, nlevels, 20, 7)
    compare_full(oracle, ex, oex, img, (0, 0))
    for level in range(nlevels):
        assert np.array_equal(ex.pyramid_level(level, with_border=True), oex.pyramid_level(level, with_border=True)), level


def test_flat_and_noise_images(oracle):
    ex = orbx.ORBextractor(500, 1.2, 8, 20, 7)
    oex = oracle.extractor(500, 1.2, 8, 20, 7)
    flat = np.full((240, 320), 128, np.uint8)
    nm, kps, desc = ex(flat, None, (0, 0))
    assert nm == 0 and len(kps) == 0
    noise = np.random.default_rng(0).integers(0, 256, (240, 320), dtype=np.uint8)
    compare_full(oracle, ex, oex, noise, (0, 0))
    # empty image: reference returns -1 (src/ORBextractor.cc:1231-1232)
    assert ex(np.zeros((0, 0), np.uint8), None, (0, 0))[0] == -1


def test_strided_input_and_shape_change(oracle):
    big = synth.image(10, 800, 500)
    view = big[10:490, 20:772]                      # 752x480 view with row stride 800
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    oex = oracle.extractor(1000, 1.2, 8, 20, 7)
    compare_full(oracle, ex, oex, np.ascontiguousarray(view), (0, 0))
    nm, kps, desc = ex(view, None, (0, 0))
    kps_o, desc_o, nm_o = oex.extract(np.ascontiguousarray(view), (0, 0))
    assert np.array_equal(desc, desc_o) and np.array_equal(kps["x"], kps_o["x"])
    compare_full(oracle, ex, oex, synth.image(11, 400, 300), (0, 0))   # same handle, new shape


def test_batch_equals_single_and_device_api(oracle):
    import torch
    F, cols, rows = 6, 752, 480
    imgs = np.stack([synth.image(200 + f, cols, rows) for f in range(F)])
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_batch=4)       # forces 2 chunks / 2 slots in the host batch API
    nm, n, kps, desc = ex.extract_batch(imgs, (0, 0))
    single = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    oex = oracle.extractor(1000, 1.2, 8, 20, 7)
    for f in range(F):
        nm1, k1, d1 = single(imgs[f], None, (0, 0))
        assert nm[f] == nm1 and n[f] == len(k1)
        assert np.array_equal(kps[f, :n[f]], k1) and np.array_equal(desc[f, :n[f]], d1)
        ko, do, nmo = oex.extract(imgs[f], (0, 0))
        assert np.array_equal(k1["x"], ko["x"]) and np.array_equal(d1, do)
    # device-resident inputs generated on the GPU must equal the host generator, and the device API the host API
    dev = torch.device("cuda:0")
    d_img = torch.empty((F, rows, cols), dtype=torch.uint8, device=dev)
    synth.images_device(d_img, 200, F, cols, rows, cols, rows * cols)
    torch.cuda.synchronize()
    assert np.array_equal(d_img.cpu().numpy(), imgs)
    cap = ex.max_keypoints()
    d_kps = torch.zeros((F, cap, 7), dtype=torch.float32, device=dev)
    d_desc = torch.zeros((F, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(F, dtype=torch.int32, device=dev)
    d_nm = torch.zeros(F, dtype=torch.int32, device=dev)
    ex.extract_batch_device(d_img, F, rows, cols, cols, rows * cols, d_kps, d_desc, cap, d_n, d_nm)
    ex.sync()
    assert np.array_equal(d_n.cpu().numpy(), n) and np.array_equal(d_nm.cpu().numpy(), nm)
    hk = d_kps.cpu().numpy().view(np.uint8).reshape(F, cap, 28)
    for f in range(F):
        assert np.array_equal(hk[f, :n[f]].reshape(-1), kps[f, :n[f]].view(np.uint8).reshape(-1))
        assert np.array_equal(d_desc[f, :n[f]].cpu().numpy(), desc[f, :n[f]])


def random_candidates(rng, n, w, h, clustered):
    if clustered:
        cx, cy = rng.integers(0, w, 6), rng.integers(0, h, 6)
        k = rng.integers(0, 6, n)
        xs = np.clip(cx[k] + rng.integers(-12, 13, n), 0, w - 1)
        ys = np.clip(cy[k] + rng.integers(-12, 13, n), 0, h - 1)
    else:
        xs, ys = rng.integers(0, w, n), rng.integers(0, h, n)
    pts = np.unique(np.stack([ys // 38 * 1000 + xs // 36, ys, xs], 1), axis=0)     # unique pixels, cell-major-ish order
    return pts[:, 2].astype(np.int32), pts[:, 1].astype(np.int32)


def test_octree_standalone_random(oracle):
    rng = np.random.default_rng(42)
    cases = 0
    for (w, h) in [(720, 448), (1209, 344), (178, 102), (595, 368), (88, 88), (300, 100)]:
        for n in (0, 1, 2, 5, 40, 300, 1500, 6000):
            for N in (1, 7, 60, 217, 434):
                for clustered in (False, True):
                    xs, ys = random_candidates(rng, n, w, h, clustered)
                    sc = rng.integers(7, 60, len(xs)).astype(np.int32)       # few distinct scores -> many response ties
                    ref = oracle.octree(xs, ys, sc, 16, 16 + w, 16, 16 + h, N)
                    got = orbx.distribute_octree(xs, ys, sc, 16, 16 + w, 16, 16 + h, N)
                    assert np.array_equal(ref, got), (w, h, n, N, clustered)
                    cases += 1
    assert cases > 400


def test_knn2_matches_oracle():
    from tests import oracle_lib
    o = oracle_lib.load()
    m = orbx.ORBmatcher(0.7)
    for (nq, ndb, seed) in [(1, 1, 1), (3, 2, 2), (64, 1, 3), (700, 4000, 4), (1000, 20000, 5), (513, 70001, 6)]:
        db = synth.descriptors(seed, ndb)
        q = synth.descriptors(seed, nq, is_query=True, ndb=ndb, plant_every=3)
        if ndb > 10:
            db[5] = db[3]; db[ndb - 1] = db[3]                       # ties on distance -> lower index must win
            q[0] = db[3]
        idx, dist = m.knn2(q, db)
        ridx, rdist = o.knn2(q, db)
        assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist), (nq, ndb)
    # empty database
    idx, dist = m.knn2(synth.descriptors(1, 4), np.zeros((0, 32), np.uint8))
    assert (idx == -1).all() and (dist == np.iinfo(np.int32).max).all()
    # ratio tests (src/ORBmatcher1.cc:329-333, src/ORBmatcher2.cc:120, src/Frame.cc:1181)
    acc0 = m.ratio_test(dist=rdist, mode=0, th_low=50, ratio=0.7)
    acc2 = m.ratio_test(dist=rdist, mode=2, ratio=0.7)
    for i in range(len(rdist)):
        assert acc0[i] == o.ratio_accept(rdist[i, 0], rdist[i, 1], 0.7, 50)
        assert acc2[i] == o.ratio_accept(rdist[i, 0], rdist[i, 1], 0.7, -1)
    assert acc0.any() and not acc0.all()


def test_descriptor_distance_host(oracle):
    rng = np.random.default_rng(3)
    for _ in range(200):
        a = rng.integers(0, 256, 32, dtype=np.uint8); b = rng.integers(0, 256, 32, dtype=np.uint8)
        assert orbx.ORBmatcher.DescriptorDistance(a, b) == oracle.hamming(a, b, swar=True)


@pytest.mark.parametrize("cols,rows,nfeatures", [(752, 480, 1200), (1241, 376, 2000)])
def test_stereo_matches(oracle, cols, rows, nfeatures):
    left = synth.image(31, cols, rows, view=0)
    right = synth.image(31, cols, rows, view=1)
    exL = orbx.ORBextractor(nfeatures, 1.2, 8, 20, 7); exR = orbx.ORBextractor(nfeatures, 1.2, 8, 20, 7)
    oL = oracle.extractor(nfeatures, 1.2, 8, 20, 7); oR = oracle.extractor(nfeatures, 1.2, 8, 20, 7)
    _, kL, dL = exL(left, None, (0, 0)); _, kR, dR = exR(right, None, (0, 0))
    okL, odL, _ = oL.extract(left, (0, 0)); okR, odR, _ = oR.extract(right, (0, 0))
    assert np.array_equal(dL, odL) and np.array_equal(dR, odR)
    bf, max_d = 47.9, 435.0 * 47.9 / 47.9 if False else 47.9 / 0.11
    u, d = orbx.compute_stereo_matches(exL, exR, kL, dL, kR, dR, bf, max_d)
    ou, od, kept = oracle.stereo_match(oL, oR, okL, odL, okR, odR, bf, max_d)
    assert np.array_equal(u, ou) and np.array_equal(d, od)
    assert kept > 50 and (u >= 0).sum() == kept


@pytest.mark.parametrize("ch,rgb", [(3, True), (3, False), (4, True), (4, False)])
def test_color_input_cvt_gray(oracle, ch, rgb):
    """Tracking::GrabImage* image prep (src/Tracking2.cc:289-316): cvtColor on the device in front of operator()."""
    rng = np.random.default_rng(90 + ch)
    for (cols, rows) in [(752, 480), (331, 277), (161, 123)]:
        base = synth.image(70 + ch, cols, rows).astype(np.int32)
        color = np.clip(base[..., None] + rng.integers(-40, 41, (rows, cols, ch)), 0, 255).astype(np.uint8)
        gray = oracle.cvt_gray(color, rgb)
        ex = orbx.ORBextractor(600, 1.2, 8, 20, 7)
        nm, kps, desc = ex.extract_color(color, rgb)
        assert np.array_equal(ex.pyramid_level(0), gray)                     # mImGray
        ex2 = orbx.ORBextractor(600, 1.2, 8, 20, 7)
        nm2, kps2, desc2 = ex2(gray)
        assert nm == nm2 and np.array_equal(kps, kps2) and np.array_equal(desc, desc2)
        oex = oracle.extractor(600, 1.2, 8, 20, 7)
        kps_o, desc_o, nm_o = oex.extract(gray, (0, 0))
        assert nm == nm_o and np.array_equal(kps["x"], kps_o["x"]) and np.array_equal(kps["y"], kps_o["y"])
    # strided (non-contiguous rows) colour view
    big = rng.integers(0, 256, (200, 300, ch), dtype=np.uint8)
    view = big[10:150, 20:260]
    ex = orbx.ORBextractor(300, 1.2, 8, 20, 7)
    nm, kps, desc = ex.extract_color(view, rgb)
    assert np.array_equal(ex.pyramid_level(0), oracle.cvt_gray(np.ascontiguousarray(view), rgb))


def test_host_batch_serial_pipeline_matches_single_frames(oracle):
    """orbx_extract_batch with >= 128-frame chunks takes the three-stream serial pipeline (H2D / compute / D2H streams, slots
    reused through events): every frame of a 700-frame call must equal the single-frame result, bit for bit, also on a second
    call that reuses the slots, and a sample of frames must equal the oracle."""
    cols, rows, F = 200, 150, 700
    imgs = np.stack([synth.image(4000 + f, cols, rows) for f in range(F)])
    ex = orbx.ORBextractor(300, 1.2, 6, 20, 7, max_cols=cols, max_rows=rows, max_batch=128)
    single = orbx.ORBextractor(300, 1.2, 6, 20, 7)
    for rep in range(2):
        nm, n, kps, desc = ex.extract_batch(imgs if rep == 0 else imgs[::-1].copy())
        order = range(F) if rep == 0 else range(F - 1, -1, -1)
        for slot, f in enumerate(order):
            if f % 7 and rep:          # second pass: a sample is enough
                continue
            m1, k1, d1 = single(imgs[f])
            assert nm[slot] == m1 and n[slot] == len(k1), (rep, f)
            assert kps[slot, :n[slot]].tobytes() == k1.tobytes() and np.array_equal(desc[slot, :n[slot]], d1), (rep, f)
    oex = oracle.extractor(300, 1.2, 6, 20, 7)
    for f in (0, 333, 699):
        ok, od, onm = oex.extract(imgs[f], (0, 0))
        slot = F - 1 - f
        assert n[slot] == len(ok) and np.array_equal(kps[slot, :n[slot]]["x"], ok["x"]) and np.array_equal(kps[slot, :n[slot]]["y"], ok["y"])


@pytest.mark.parametrize("cols,rows,nfeatures,nlevels", [(1300, 100, 100, 4), (1390, 84, 60, 3)])
def test_wide_image_small_nfeatures_capacity(oracle, cols, rows, nfeatures, nlevels):
    """DistributeOctTree splits all nIni = round(width / height) roots before it compares with N (src/ORBextractor.cc:589-608,
    635-698): a strip image returns up to 4 * nIni keypoints per level, far more than mnFeaturesPerLevel + 3.  Outputs sized by
    orbx_max_keypoints_for(rows, cols) must hold them on a fresh handle — single frame and host batch."""
    imgs = np.stack([synth.image(4000 + f, cols, rows) for f in range(3)])
    oex = oracle.extractor(nfeatures, 1.2, nlevels, 20, 7)
    ex = orbx.ORBextractor(nfeatures, 1.2, nlevels, 20, 7)
    assert ex.max_keypoints(rows, cols) > ex.max_keypoints()        # fresh handle: the shape-free figure is too small
    kps, _, _, _ = compare_full(oracle, ex, oex, imgs[0], (0, 0))
    assert len(kps) > nfeatures + 3 * nlevels                       # the case the old bound could not hold
    assert ex.max_keypoints() == ex.max_keypoints(rows, cols)       # configured handle agrees with the shape-only bound
    ex.close()
    ex = orbx.ORBextractor(nfeatures, 1.2, nlevels, 20, 7)
    nmv, nv, kb, db = ex.extract_batch(imgs)
    for f in range(3):
        ko, do, nmo = oex.extract(imgs[f], (0, 0))
        assert int(nmv[f]) == nmo and int(nv[f]) == len(ko)
        for fld in ("x", "y", "size", "response", "octave"):
            assert np.array_equal(kb[f][:nv[f]][fld], ko[fld]), fld
        assert np.array_equal(db[f][:nv[f]], do)
    ex.close()
