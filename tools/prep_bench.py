"""Device throughput of the rectification / resize kernels on resident batches (CUDA events on the launching stream) and the
per-call latency of the projection-guided searches and UndistortKeyPoints (host arrays in, results out)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import wut_cuda_orb_slam3_b200 as orbx
from tests.proj_synth import SCALE, make_frame, make_points
from wut_cuda_orb_slam3_b200 import synth


def rect_maps(w, h):
    """A radial-distortion-like rectification map (smooth, sub-pixel everywhere), what cv::initUndistortRectifyMap produces."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    xn, yn = (xx - w / 2) / (w / 2), (yy - h / 2) / (h / 2)
    r2 = xn * xn + yn * yn
    return (xx + 6.0 * xn * r2 + 0.37).astype(np.float32), (yy + 6.0 * yn * r2 - 0.21).astype(np.float32)


def time_device(fn, reps=20):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            fn(st)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        st.synchronize()
        e0.record(st)
        for _ in range(reps):
            fn(st)
        e1.record(st)
        st.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    out = {}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6555.8))
    w, h, nf = 752, 480, 512                                    # 185 MB in + 185 MB out per launch: larger than L2
    d_src = torch.empty((nf, h, w), dtype=torch.uint8, device="cuda")
    synth.images_device(d_src, 100, nf, w, h, w, h * w)
    d_dst = torch.empty_like(d_src)
    mx, my = rect_maps(w, h)
    rect = orbx.Rectifier(mx, my)
    ms = time_device(lambda st: rect.remap_device(d_src, h, w, w, h * w, nf, d_dst, w, h * w, stream=st.cuda_stream))
    out["remap_752x480_x512"] = dict(ms=ms, frames_per_s=nf / ms * 1e3, algorithmic_GBps=2 * nf * w * h / ms / 1e6,
                                     frac_of_hbm_peak=2 * nf * w * h / ms / 1e6 / hbm)
    rs = orbx.Rectifier(resize=(h, w, 400, 627))
    d_small = torch.empty((nf, 400, 627), dtype=torch.uint8, device="cuda")
    ms = time_device(lambda st: rs.remap_device(d_src, h, w, w, h * w, nf, d_small, 627, 400 * 627, stream=st.cuda_stream))
    out["resize_752x480_to_627x400_x512"] = dict(ms=ms, frames_per_s=nf / ms * 1e3, algorithmic_GBps=nf * (w * h + 627 * 400) / ms / 1e6,
                                                 frac_of_hbm_peak=nf * (w * h + 627 * 400) / ms / 1e6 / hbm)
    # projection searches: per-call latency through the C ABI
    rng = np.random.default_rng(71)
    kp, desc, ur, occ, bounds = make_frame(rng, 1200)
    P = make_points(rng, kp, desc, ur, 2000)
    fv = orbx.FrameView(kp, desc, SCALE, bounds, u_right=ur, occupied=occ)

    def lat(fn, reps=200):
        for _ in range(10):
            fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e6

    out["search_by_projection_map_1200x2000_us"] = lat(lambda: orbx.search_by_projection_map(
        fv, P["in_view"], P["bad"], P["x"], P["y"], P["xr"], P["view_cos"], P["depth"], P["level"], P["n_obs"], P["desc"]))
    out["search_by_projection_map_rounds"] = orbx.projection_rounds()
    out["search_by_projection_last_1200x2000_us"] = lat(lambda: orbx.search_by_projection_last(
        fv, 40.0, P["valid"], P["x"], P["y"], P["invz"], P["level"], P["angle"], P["n_obs"], P["desc"], 15.0))
    out["search_by_projection_last_rounds"] = orbx.projection_rounds()
    # a well-tracked frame: 1000 map points, each aimed at its own feature (few competing claims)
    U = make_points(rng, kp, desc, ur, 1000, unique=True, max_flip=50)
    out["search_by_projection_map_1200x1000_unique_us"] = lat(lambda: orbx.search_by_projection_map(
        fv, U["in_view"], U["bad"], U["x"], U["y"], U["xr"], U["view_cos"], U["depth"], U["level"], U["n_obs"], U["desc"]))
    out["search_by_projection_map_unique_rounds"] = orbx.projection_rounds()
    out["search_by_projection_last_1200x1000_unique_us"] = lat(lambda: orbx.search_by_projection_last(
        fv, 40.0, U["valid"], U["x"], U["y"], U["invz"], U["level"], U["angle"], U["n_obs"], U["desc"], 15.0))
    out["search_by_projection_last_unique_rounds"] = orbx.projection_rounds()
    kps = np.zeros(1200, orbx.KP_DTYPE); kps["x"] = rng.uniform(0, 752, 1200); kps["y"] = rng.uniform(0, 480, 1200)
    out["undistort_keypoints_1200_us"] = lat(lambda: orbx.undistort_keypoints(kps, [458.6, 457.3, 367.2, 248.4], [-0.283, 0.074, 1.9e-4, 1.8e-5]))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
