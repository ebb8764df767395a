"""GPU parity tests of the preparation steps either side of the extractor, bit-exact against the oracle
(which tests/test_oracle_cv2.py pins against cv2)."""
import numpy as np
import pytest
import torch

import wut_cuda_orb_slam3_b200 as orbx
from tests.oracle_lib import KP_DTYPE
from tests.test_oracle_cv2 import _camera, _maps

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("trial", range(9))
def test_undistort_keypoints_matches_oracle(oracle, trial):
    rng = np.random.default_rng(trial)
    K, d, P = _camera(rng, trial)
    n = [1000, 1, 3333][trial % 3]
    kps = np.zeros(n, KP_DTYPE)
    kps["x"] = rng.uniform(0, 752, n); kps["y"] = rng.uniform(0, 480, n)
    kps["octave"] = rng.integers(0, 8, n); kps["angle"] = rng.uniform(0, 360, n); kps["response"] = rng.uniform(7, 200, n)
    kps["size"] = 31; kps["class_id"] = -1
    got = orbx.undistort_keypoints(kps, K, d, P)
    want = oracle.undistort_keypoints(kps, K, d, P)
    assert got.tobytes() == want.tobytes()
    # mDistCoef[0] == 0: plain copy (src/Frame.cc:779-783)
    d0 = d.copy(); d0[0] = 0
    assert orbx.undistort_keypoints(kps, K, d0, P).tobytes() == kps.tobytes()


@pytest.mark.parametrize("case", range(6))
def test_remap_matches_oracle(oracle, case):
    rng = np.random.default_rng(200 + case)
    sh, sw = [(480, 752), (376, 1241), (61, 47)][case % 3]
    dh, dw = [(480, 752), (300, 501), (70, 90)][case % 3]
    src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    mx, my = _maps(rng, sw, sh, dw, dh, shift=25.0 if case >= 3 else 0.0)
    r = orbx.Rectifier(mx, my)
    got = r.remap(src)
    assert np.array_equal(got, oracle.remap(src, mx, my))
    # a strided view as input and a second image through the same rectifier
    big = rng.integers(0, 256, (sh, sw + 13), dtype=np.uint8)
    assert np.array_equal(r.remap(big[:, 5:5 + sw]), oracle.remap(np.ascontiguousarray(big[:, 5:5 + sw]), mx, my))
    r.close()


def test_remap_device_batch_feeds_extractor(oracle):
    """Rectify a batch on the device and extract from the rectified frames without leaving the GPU."""
    rng = np.random.default_rng(300)
    h, w, nf = 480, 752, 3
    imgs = np.stack([orbx.synth.image(900 + f, w, h) for f in range(nf)])
    mx, my = _maps(rng, w, h, w, h, noise=0.0)
    r = orbx.Rectifier(mx, my)
    d_src = torch.from_numpy(imgs).cuda()
    d_dst = torch.zeros((nf, h, w), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    r.remap_device(d_src, h, w, w, h * w, nf, d_dst, w, h * w)
    torch.cuda.synchronize()
    rect = d_dst.cpu().numpy()
    for f in range(nf):
        assert np.array_equal(rect[f], oracle.remap(imgs[f], mx, my))
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_cols=w, max_rows=h, max_batch=nf)
    oex = oracle.extractor(1000, 1.2, 8, 20, 7)
    cap = ex.max_keypoints()
    d_kps = torch.zeros((nf, cap, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((nf, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(nf, dtype=torch.int32, device="cuda"); d_nm = torch.zeros(nf, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ex.extract_batch_device(d_dst, nf, h, w, w, h * w, d_kps, d_desc, cap, d_n, d_nm)
    ex.sync()
    n = d_n.cpu().numpy()
    hk = d_kps.cpu().numpy().view(np.uint8).reshape(nf, cap, 28)
    for f in range(nf):
        okp, odesc, _ = oex.extract(rect[f])
        assert n[f] == len(okp) and n[f] > 500
        kp = hk[f, :n[f]].copy().view(KP_DTYPE).reshape(-1)
        assert np.array_equal(kp["x"], okp["x"]) and np.array_equal(kp["y"], okp["y"]) and np.array_equal(kp["octave"], okp["octave"])
        assert (d_desc[f, :n[f]].cpu().numpy() != odesc).any(axis=1).mean() <= 0.005


@pytest.mark.parametrize("src,dst", [((480, 752), (400, 627)), ((480, 752), (240, 376)), ((376, 1241), (480, 752)), ((100, 90), (37, 201)),
                                     ((480, 640), (960, 1280))])
def test_resize_matches_oracle(oracle, src, dst):
    """cv::resize(im, out, newImSize) (src/System.cc:261-263): the oracle's resize is pinned against cv2 in tests/test_oracle_cv2.py."""
    rng = np.random.default_rng(400)
    img = rng.integers(0, 256, src, dtype=np.uint8)
    r = orbx.Rectifier(resize=(src[0], src[1], dst[0], dst[1]))
    got = r.remap(img)
    want = oracle.resize(img, dst[1], dst[0])
    assert np.array_equal(got, want)
    with pytest.raises(orbx.OrbxError):
        r.remap(img[:-1])
    r.close()


@pytest.mark.parametrize("case", range(5))
def test_remap_device_batch_tiled_matches_oracle(oracle, case):
    """>= 8 resident frames take the tiled shared-memory kernel: smooth maps, maps leaving the image, noisy maps (window too
    large -> gather fallback inside the kernel), odd destination sizes."""
    rng = np.random.default_rng(700 + case)
    sh, sw = [(480, 752), (376, 1248), (480, 752), (128, 160), (480, 752)][case]
    dh, dw = [(480, 752), (300, 501), (481, 750), (70, 90), (240, 376)][case]
    nf = [9, 8, 17, 33, 16][case]
    src = rng.integers(0, 256, (nf, sh, sw), dtype=np.uint8)
    noise = [0.0, 0.3, 0.0, 40.0, 0.0][case]
    mx, my = _maps(rng, sw, sh, dw, dh, shift=[0.0, 0.0, 30.0, 0.0, -12.0][case], noise=noise)
    r = orbx.Rectifier(mx, my)
    d_src = torch.from_numpy(src).cuda()
    d_dst = torch.full((nf, dh, dw), 7, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    r.remap_device(d_src, sh, sw, sw, sh * sw, nf, d_dst, dw, dh * dw)
    torch.cuda.synchronize()
    got = d_dst.cpu().numpy()
    for f in range(nf):
        assert np.array_equal(got[f], oracle.remap(src[f], mx, my)), (case, f)
    r.close()


@pytest.mark.parametrize("src,dst,nf", [((480, 752), (400, 627), 9), ((480, 752), (240, 376), 8), ((376, 1248), (480, 752), 17), ((96, 80), (37, 201), 33),
                                        ((480, 640), (960, 1280), 8), ((480, 752), (30, 47), 12)])
def test_resize_device_batch_tiled_matches_oracle(oracle, src, dst, nf):
    """>= 8 resident frames take the tiled cp.async kernel: down-scales, the exact-2x area path, up-scales, odd sizes and a
    16x down-scale whose source window does not fit the ring (global fallback inside the kernel)."""
    rng = np.random.default_rng(800 + nf)
    imgs = rng.integers(0, 256, (nf,) + src, dtype=np.uint8)
    r = orbx.Rectifier(resize=(src[0], src[1], dst[0], dst[1]))
    d_src = torch.from_numpy(imgs).cuda()
    d_dst = torch.full((nf,) + dst, 9, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    r.remap_device(d_src, src[0], src[1], src[1], src[0] * src[1], nf, d_dst, dst[1], dst[0] * dst[1])
    torch.cuda.synchronize()
    got = d_dst.cpu().numpy()
    for f in range(nf):
        assert np.array_equal(got[f], oracle.resize(imgs[f], dst[1], dst[0])), f
    r.close()
