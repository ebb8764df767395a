"""Pins the oracle's restatement of the OpenCV-owned primitives against cv2 (the same third-party arithmetic the
reference calls: cv::resize, copyMakeBorder, FAST, GaussianBlur, fastAtan2, BFMatcher.knnMatch, ORB descriptors).
CPU only.  SURVEY.md §8(c)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from wut_cuda_orb_slam3_b200 import synth  # host generator only (no GPU needed)

SHAPES = [(752, 480), (1241, 376), (1280, 720), (160, 120)]


def level_sizes(oracle, cols, rows, nlevels=8, sf=1.2):
    inv = oracle.tables(1000, sf, nlevels)["inv"]
    return [(int(np.rint(np.float32(cols) * inv[l])), int(np.rint(np.float32(rows) * inv[l]))) for l in range(nlevels)]


@pytest.mark.parametrize("cols,rows", SHAPES)
def test_resize_chain_matches_cv2(oracle, cols, rows):
    img = synth.image(7, cols, rows)
    prev = img
    for (w, h) in level_sizes(oracle, cols, rows)[1:]:
        ref = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
        got = oracle.resize(prev, w, h)
        assert np.array_equal(ref, got), (w, h, int((ref != got).sum()))
        prev = ref


def test_resize_noise_and_odd_ratios(oracle):
    rng = np.random.default_rng(0)
    for (sw, sh, dw, dh) in [(97, 61, 81, 51), (300, 200, 250, 167), (64, 64, 32, 32), (101, 77, 50, 38), (40, 30, 33, 25)]:
        src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
        got = oracle.resize(src, dw, dh)
        assert np.array_equal(ref, got), (sw, sh, dw, dh)


def test_border_reflect101(oracle):
    rng = np.random.default_rng(1)
    for (w, h) in [(64, 48), (210, 134), (23, 21)]:
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        ref = cv2.copyMakeBorder(src, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
        assert np.array_equal(ref, oracle.make_border(src, 19))


@pytest.mark.parametrize("threshold", [20, 7])
def test_fast_matches_cv2(oracle, threshold):
    det = cv2.FastFeatureDetector_create(threshold=threshold, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    rng = np.random.default_rng(2)
    imgs = [synth.image(3, 160, 120), synth.image(4, 80, 60), rng.integers(0, 256, (50, 47), dtype=np.uint8),
            synth.image(5, 42, 44), np.zeros((30, 30), np.uint8), synth.image(6, 9, 12)]
    total = 0
    for img in imgs:
        kps = det.detect(img)
        ref = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kps]
        xs, ys, sc = oracle.fast9(img, threshold, True)
        got = list(zip(xs.tolist(), ys.tolist(), sc.tolist()))
        assert ref == got
        for k in kps:
            assert k.size == 7 and k.angle == -1 and k.octave == 0 and k.class_id == -1
        total += len(ref)
    assert total > 100


def test_fast_no_nms_matches_cv2(oracle):
    det = cv2.FastFeatureDetector_create(threshold=12, nonmaxSuppression=False, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    img = synth.image(8, 120, 90)
    ref = sorted((int(k.pt[1]), int(k.pt[0])) for k in det.detect(img))
    xs, ys, sc = oracle.fast9(img, 12, False)
    assert ref == sorted(zip(ys.tolist(), xs.tolist()))


def cell_fast_cv2(img, ini_th=20, min_th=7):
    """The reference's tileCalcKeypoints loop (src/ORBextractor.cc:867-950) with the real cv2 FAST per cell."""
    h, w = img.shape
    min_b = 16
    max_bx, max_by = w - 16, h - 16
    width, height = np.float32(max_bx - min_b), np.float32(max_by - min_b)
    n_cols, n_rows = int(width / np.float32(35)), int(height / np.float32(35))
    w_cell, h_cell = int(np.ceil(width / n_cols)), int(np.ceil(height / n_rows))
    det_i = cv2.FastFeatureDetector_create(threshold=ini_th, nonmaxSuppression=True)
    det_m = cv2.FastFeatureDetector_create(threshold=min_th, nonmaxSuppression=True)
    out = []
    for i in range(n_rows):
        ini_y = min_b + i * h_cell
        max_y = ini_y + h_cell + 6
        if ini_y >= max_by - 3:
            continue
        max_y = min(max_y, max_by)
        for j in range(n_cols):
            ini_x = min_b + j * w_cell
            max_x = ini_x + w_cell + 6
            if ini_x >= max_bx - 6:
                continue
            max_x = min(max_x, max_bx)
            roi = np.ascontiguousarray(img[ini_y:max_y, ini_x:max_x])
            kps = det_i.detect(roi)
            if not kps:
                kps = det_m.detect(roi)
            out += [(int(k.pt[0]) + j * w_cell, int(k.pt[1]) + i * h_cell, int(k.response)) for k in kps]
    return out


@pytest.mark.parametrize("cols,rows,seed", [(752, 480, 1), (363, 231, 2), (210, 134, 3), (346, 105, 4), (1280, 720, 5)])
def test_cell_fast_matches_cv2_cell_loop(oracle, cols, rows, seed):
    img = synth.image(seed, cols, rows)
    ref = cell_fast_cv2(img)
    xs, ys, sc = oracle.cell_fast(img, 20, 7)
    assert ref == list(zip(xs.tolist(), ys.tolist(), sc.tolist()))
    assert len(ref) > 10
    if cols * rows >= 752 * 480:
        assert any(s < 20 for (_, _, s) in ref), "generator must exercise the minThFAST retry path"


def test_gaussian_blur(oracle):
    rng = np.random.default_rng(3)
    imgs = [synth.image(9, 210, 134), synth.image(10, 752, 480), rng.integers(0, 256, (61, 97), dtype=np.uint8),
            rng.integers(0, 256, (8, 9), dtype=np.uint8), np.full((20, 20), 255, np.uint8)]
    for img in imgs:
        ref = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(ref, oracle.blur(img))


def test_fast_atan2(oracle):
    rng = np.random.default_rng(4)
    pts = rng.integers(-2_000_000, 2_000_000, (20000, 2))
    pts[:50] = rng.integers(-3, 4, (50, 2))
    for y, x in pts:
        assert oracle.fast_atan2(y, x) == cv2.fastAtan2(float(y), float(x))
    assert oracle.fast_atan2(0, 0) == 0.0


def test_knn2_matches_bfmatcher(oracle):
    rng = np.random.default_rng(5)
    db = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    q = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    q[:20] = db[rng.integers(0, 300, 20)]          # exact hits -> distance 0
    db[100] = db[7]; db[200] = db[7]; q[3] = db[7]  # three-way tie
    q[21] = db[5]; q[21, 0] ^= 0x0F
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    ref = bf.knnMatch(q, db, k=2)
    idx, dist = oracle.knn2(q, db)
    idx_s, dist_s = oracle.knn2(q, db, swar=True)
    assert np.array_equal(idx, idx_s) and np.array_equal(dist, dist_s)
    for i, m in enumerate(ref):
        assert [m[0].trainIdx, m[1].trainIdx] == idx[i].tolist()
        assert [int(m[0].distance), int(m[1].distance)] == dist[i].tolist()
    assert idx[3].tolist() == [7, 100]


def test_descriptor_and_angle_match_cv2_orb(oracle):
    """cv2.ORB.compute on a single level (scale 1) uses the same blur, pattern and rounding as computeOrbDescriptor."""
    img = synth.image(11, 320, 240)
    xs, ys, sc = oracle.cell_fast(img, 20, 7)
    xs, ys = xs + 16, ys + 16
    sel = np.arange(0, len(xs), max(1, len(xs) // 300))
    angles = [oracle.ic_angle(img, int(xs[i]), int(ys[i])) for i in sel]
    kps = [cv2.KeyPoint(float(xs[i]), float(ys[i]), 31.0, float(a), float(sc[i]), 0, -1) for i, a in zip(sel, angles)]
    orb = cv2.ORB_create(nfeatures=len(kps), scaleFactor=1.2, nlevels=1, edgeThreshold=19, firstLevel=0, WTA_K=2, patchSize=31)
    kps2, ref = orb.compute(img, kps)
    assert len(kps2) == len(kps)
    blurred = oracle.blur(img)
    # cv2's ORB blurs a *sub-matrix* of its pyramid, which sends cv::GaussianBlur down the float sepFilter2D path instead
    # of the bit-exact fixed-point path taken for the reference's `mvImagePyramid[level].clone()` (src/ORBextractor.cc:1270),
    # so a few comparisons between nearly equal pixels flip.  Pattern / rotation / rounding errors would flip dozens of bits.
    nbits = []
    for k, r in zip(kps2, ref):
        got = oracle.descriptor(blurred, int(k.pt[0]), int(k.pt[1]), k.angle)
        nbits.append(int(np.unpackbits(got ^ r).sum()))
    nbits = np.array(nbits)
    assert nbits.max() <= 3 and (nbits == 0).mean() > 0.6, np.bincount(nbits)


@pytest.mark.parametrize("cols,rows,seed", [(752, 480, 200), (1241, 376, 201), (331, 277, 202), (160, 120, 203)])
def test_ic_angle_pinned_against_cv2_orb_icangles(oracle, cols, rows, seed):
    """SURVEY.md §8 row A7.  The fork deleted the CPU IC_Angle (only `using cv::fastAtan2` survives at src/ORBextractor.cc:85)
    and its OpenCL kernel has no reduction (src/OpenCL/Kernel/Angle.cl:55-60), so the reference holds nothing to pin the
    orientation against.  ORB-SLAM's IC_Angle is OpenCV's ICAngles (same umax table, same u/v summation, same fastAtan2), and
    cv2.ORB with ONE level runs ICAngles on the image itself at the integer FAST positions: the oracle's IC_Angle must give
    the same float32, bit for bit, at every keypoint cv2 detects (two FAST thresholds: ~40 000 keypoints over the shapes)."""
    img = synth.image(seed, cols, rows)
    total = 0
    for fast_th in (20, 7):
        orb = cv2.ORB_create(nfeatures=20000, scaleFactor=1.2, nlevels=1, edgeThreshold=19, firstLevel=0, WTA_K=2, patchSize=31,
                             fastThreshold=fast_th)
        kps = orb.detect(img, None)
        assert len(kps) > 100
        for kp in kps:
            x, y = kp.pt
            assert x == int(x) and y == int(y)                 # level 0 of a one-level pyramid: integer FAST positions
            assert 19 <= x < cols - 19 and 19 <= y < rows - 19
            got = np.float32(oracle.ic_angle(img, int(x), int(y)))
            assert got == np.float32(kp.angle), (x, y, got, kp.angle)
            total += 1
    assert total > 500


def test_descriptor_against_numpy_restatement(oracle):
    """Independent numpy float32 restatement of computeOrbDescriptor on cv2.GaussianBlur output."""
    pat = np.array([int(v) for l in open("wut_cuda_orb_slam3_b200/csrc/brief_pattern.inc") if not l.startswith("//")
                    for v in l.strip().strip(",").split(",")], np.int32).reshape(512, 2)
    img = synth.image(13, 300, 200)
    bl = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    rng = np.random.default_rng(6)
    px, py = pat[:, 0].astype(np.float32), pat[:, 1].astype(np.float32)
    for _ in range(200):
        x, y = int(rng.integers(19, 281)), int(rng.integers(19, 181))
        ang_deg = np.float32(rng.uniform(0, 360))
        ang = ang_deg * np.float32(np.pi / np.float32(180.0))
        a, b = np.float32(np.cos(np.float64(ang))), np.float32(np.sin(np.float64(ang)))
        iy = np.rint(px * b + py * a).astype(int)
        ix = np.rint(px * a - py * b).astype(int)
        v = bl[y + iy, x + ix].astype(int)
        bits = (v[0::2] < v[1::2]).astype(np.uint8)
        ref = np.packbits(bits.reshape(32, 8)[:, ::-1], axis=1).reshape(32)
        assert np.array_equal(ref, oracle.descriptor(bl, x, y, ang_deg))


def test_ic_angle_against_numpy_moments(oracle):
    img = synth.image(12, 200, 150)
    umax = oracle.tables(1000, 1.2, 8)["umax"]
    assert umax.tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    for (x, y) in [(40, 40), (100, 75), (180, 130), (19, 19)]:
        m10 = m01 = 0
        for v in range(-15, 16):
            d = umax[abs(v)]
            row = img[y + v, x - d:x + d + 1].astype(np.int64)
            u = np.arange(-d, d + 1)
            m10 += int((u * row).sum()); m01 += int(v * row.sum())
        assert oracle.ic_angle(img, x, y) == cv2.fastAtan2(float(m01), float(m10))


@pytest.mark.parametrize("code,ch,rgb", [(cv2.COLOR_RGB2GRAY, 3, True), (cv2.COLOR_BGR2GRAY, 3, False),
                                         (cv2.COLOR_RGBA2GRAY, 4, True), (cv2.COLOR_BGRA2GRAY, 4, False)])
def test_cvt_gray_bit_exact(oracle, code, ch, rgb):
    """cv::cvtColor to grey as Tracking::GrabImage* calls it (reference src/Tracking2.cc:289-316)."""
    rng = np.random.default_rng(5)
    for (h, w) in [(480, 752), (37, 53), (1, 1), (3, 4)]:
        img = rng.integers(0, 256, (h, w, ch), dtype=np.uint8)
        assert np.array_equal(oracle.cvt_gray(img, rgb), cv2.cvtColor(img, code).reshape(h, w))
    ramp = np.stack(np.meshgrid(np.arange(256), np.arange(256), indexing="ij"), -1).astype(np.uint8)   # every (c0, c1) pair
    for c2 in (0, 1, 127, 255):
        img = np.concatenate([ramp, np.full((256, 256, ch - 2), c2, np.uint8)], -1)
        assert np.array_equal(oracle.cvt_gray(img, rgb), cv2.cvtColor(img, code))


def _camera(rng, trial):
    K = np.array([rng.uniform(300, 900), rng.uniform(300, 900), rng.uniform(300, 400), rng.uniform(200, 280)], np.float32)
    nd = [4, 5, 8][trial % 3]
    d = np.zeros(nd, np.float32)
    d[0] = rng.uniform(-0.4, 0.1); d[1] = rng.uniform(-0.1, 0.2); d[2:4] = rng.uniform(-1e-3, 1e-3, 2)
    if nd >= 5:
        d[4] = rng.uniform(-0.05, 0.05)
    if nd == 8:
        d[5:8] = rng.uniform(-0.05, 0.05, 3)
    P = K.copy() if trial % 2 == 0 else np.array([rng.uniform(300, 900), rng.uniform(300, 900), rng.uniform(300, 400), rng.uniform(200, 280)], np.float32)
    return K, d, P


def _mat(k):
    return np.array([[k[0], 0, k[2]], [0, k[1], k[3]], [0, 0, 1]], np.float32)


def test_undistort_points_bit_exact():
    """cv::undistortPoints as called by Frame::UndistortKeyPoints (src/Frame.cc:797): pinhole + radtan, R = I, P = mK."""
    from tests import oracle_lib
    oracle = oracle_lib.load()
    rng = np.random.default_rng(0)
    for trial in range(24):
        K, d, P = _camera(rng, trial)
        pts = np.stack([rng.uniform(-20, 772, 1500), rng.uniform(-20, 500, 1500)], 1).astype(np.float32)
        want = cv2.undistortPoints(pts.reshape(-1, 1, 2), _mat(K), d.reshape(1, -1), None, _mat(P)).reshape(-1, 2)
        got = oracle.undistort_points(pts, K.astype(np.float64), d.astype(np.float64), P.astype(np.float64))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), trial


def _maps(rng, sw, sh, dw, dh, shift=0.0, noise=3.0):
    yy, xx = np.mgrid[0:dh, 0:dw].astype(np.float32)
    mx = (xx * sw / dw + rng.normal(0, noise, (dh, dw)) + 5 * np.sin(yy / 30) - shift).astype(np.float32)
    my = (yy * sh / dh + rng.normal(0, noise, (dh, dw)) + shift).astype(np.float32)
    return mx, my


@pytest.mark.parametrize("case", range(8))
def test_remap_linear_bit_exact(case):
    """cv::remap(INTER_LINEAR, CV_32FC1 maps, BORDER_CONSTANT 0) as called by System::TrackStereo (src/System.cc:259-260)."""
    from tests import oracle_lib
    oracle = oracle_lib.load()
    rng = np.random.default_rng(100 + case)
    sh, sw = [(480, 752), (376, 1241), (61, 47), (480, 752)][case % 4]
    dh, dw = [(480, 752), (300, 500), (70, 90), (481, 750)][case % 4]
    src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    if case == 6:      # identity and exact half-pixel positions: the saturated (0,0) weight and round-half-to-even of the map
        yy, xx = np.mgrid[0:dh, 0:dw].astype(np.float32)
        mx = (xx + np.float32(0.5) * (yy.astype(np.int32) % 2)).astype(np.float32)
        my = (yy + np.float32(1.0 / 64) * (xx.astype(np.int32) % 3)).astype(np.float32)
    else:
        mx, my = _maps(rng, sw, sh, dw, dh, shift=20.0 if case >= 4 else 0.0)
    want = cv2.remap(src, mx, my, cv2.INTER_LINEAR)
    got = oracle.remap(src, mx, my)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("src,dst", [((480, 752), (400, 627)), ((376, 1241), (480, 752)), ((100, 90), (37, 201)), ((480, 640), (960, 1280)),
                                     ((48, 64), (96, 128)), ((480, 752), (240, 376)), ((60, 80), (61, 79))])
def test_resize_to_new_size_bit_exact(oracle, src, dst):
    """cv::resize(im, out, newImSize) of System::TrackStereo (src/System.cc:261-263), up- and down-scales (rows are not clamped
    like columns: the first/last rows of an up-scale blend a row with itself)."""
    rng = np.random.default_rng(400)
    img = rng.integers(0, 256, src, dtype=np.uint8)
    assert np.array_equal(oracle.resize(img, dst[1], dst[0]), cv2.resize(img, (dst[1], dst[0])))
