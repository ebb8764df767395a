"""GPU parity tests of the frame grid and the projection-guided searches: the CUDA path through the C ABI
against the CPU oracle's sequential restatement, bit-exact (match indices and counts)."""
import time

import numpy as np
import pytest

import wut_cuda_orb_slam3_b200 as orbx
from tests.proj_synth import SCALE, make_frame, make_points

pytestmark = pytest.mark.gpu


def view(kp, desc, ur, occ, bounds):
    return orbx.FrameView(kp, desc, SCALE, bounds, u_right=ur, occupied=occ)


@pytest.mark.parametrize("n,seed", [(1500, 11), (37, 12), (1, 13), (6000, 14)])
def test_grid_and_area_match_oracle(oracle, n, seed):
    rng = np.random.default_rng(seed)
    kp, desc, ur, occ, bounds = make_frame(rng, n)
    fv = view(kp, desc, ur, occ, bounds)
    cs, items = fv.assign_features_to_grid()
    ocs, oitems = oracle.assign_features_to_grid(kp, fv.bounds_grid())
    assert np.array_equal(cs, ocs) and np.array_equal(items, oitems)
    for t in range(40):
        x, y = float(rng.uniform(-30, 780)), float(rng.uniform(-30, 510))
        r = float(rng.choice([2.5, 4.0, 10.0, 30.0, 80.0, 900.0]))
        lv = int(rng.integers(0, 8))
        mn, mx = [(-1, -1), (lv, -1), (0, lv), (lv - 1, lv + 1), (lv - 1, lv)][t % 5]
        got = fv.get_features_in_area(x, y, r, mn, mx)
        want = oracle.get_features_in_area(kp, fv.bounds_grid(), x, y, r, mn, mx)
        assert np.array_equal(got, want), (t, x, y, r, mn, mx)


@pytest.mark.parametrize("n,n_pts,th,crowd,dup,seed,mono", [
    (1000, 1500, 1.0, 0, 0.3, 21, False),       # TrackLocalMap-like
    (1200, 3000, 3.0, 0, 0.5, 22, False),       # th = 3 (after relocalisation)
    (800, 2000, 5.0, 10, 0.7, 23, False),       # crowded: long dependence chains
    (1000, 1000, 1.0, 0, 0.3, 24, True),        # monocular: no mvuRight
    (5, 50, 10.0, 1, 0.9, 25, False),           # everyone fights over five features
    (300, 0, 1.0, 0, 0.0, 26, False),           # no map points
])
def test_search_map_matches_oracle(oracle, n, n_pts, th, crowd, dup, seed, mono):
    rng = np.random.default_rng(seed)
    kp, desc, ur, occ, bounds = make_frame(rng, n, crowd=crowd)
    if mono:
        ur = None
    P = make_points(rng, kp, desc, np.full(n, -1.0, np.float32) if ur is None else ur, n_pts, dup_frac=dup, max_flip=90)
    fv = view(kp, desc, ur, occ, bounds)
    for far in (False, True):
        for ratio in (0.8, 0.6):
            got, nm = orbx.search_by_projection_map(fv, P["in_view"], P["bad"], P["x"], P["y"], P["xr"], P["view_cos"], P["depth"], P["level"],
                                                    P["n_obs"], P["desc"], th=th, far_points=far, th_far_points=20.0, nnratio=ratio)
            want, wnm = oracle.search_by_projection_map(kp, desc, ur, occ, fv.bounds_grid(), SCALE, P["in_view"], P["bad"], P["x"], P["y"],
                                                        P["xr"], P["view_cos"], P["depth"], P["level"], P["n_obs"], P["desc"], th=th, far=far,
                                                        th_far=20.0, nnratio=ratio)
            assert nm == wnm, (far, ratio, orbx.projection_rounds())
            assert np.array_equal(got, want)
    if n_pts >= 1000:
        assert nm > 100


@pytest.mark.parametrize("n,n_last,th,crowd,dup,seed,mono", [
    (1200, 1200, 15.0, 0, 0.1, 31, False),      # stereo TrackWithMotionModel (th = 15)
    (1000, 1000, 7.0, 0, 0.1, 32, True),        # monocular th = 7
    (1000, 1500, 30.0, 8, 0.6, 33, False),      # 2 * th retry on a crowded frame
    (3, 40, 30.0, 1, 0.9, 34, False),
])
def test_search_last_matches_oracle(oracle, n, n_last, th, crowd, dup, seed, mono):
    rng = np.random.default_rng(seed)
    kp, desc, ur, occ, bounds = make_frame(rng, n, crowd=crowd)
    if mono:
        ur = None
    P = make_points(rng, kp, desc, np.full(n, -1.0, np.float32) if ur is None else ur, n_last, dup_frac=dup, jitter=4.0)
    fv = view(kp, desc, ur, occ, bounds)
    for fwd, bwd in ((0, 0), (1, 0), (0, 1)):
        for ori in (True, False):
            got, nm = orbx.search_by_projection_last(fv, 40.0, P["valid"], P["x"], P["y"], P["invz"], P["level"], P["angle"], P["n_obs"],
                                                     P["desc"], th, fwd, bwd, ori)
            want, wnm = oracle.search_by_projection_last(kp, desc, ur, occ, fv.bounds_grid(), SCALE, 40.0, P["valid"], P["x"], P["y"],
                                                         P["invz"], P["level"], P["angle"], P["n_obs"], P["desc"], th, fwd, bwd, ori)
            assert nm == wnm, (fwd, bwd, ori, orbx.projection_rounds())
            assert np.array_equal(got, want)


@pytest.mark.parametrize("n,n_kf,th,orb_dist,seed", [(1000, 800, 10.0, 100, 41), (1000, 800, 3.0, 64, 42), (200, 2000, 10.0, 100, 43)])
def test_search_kf_matches_oracle(oracle, n, n_kf, th, orb_dist, seed):
    rng = np.random.default_rng(seed)
    kp, desc, ur, occ, bounds = make_frame(rng, n, occupied_frac=0.4)
    P = make_points(rng, kp, desc, ur, n_kf, dup_frac=0.4)
    fv = view(kp, desc, None, occ, bounds)
    for ori in (True, False):
        got, nm = orbx.search_by_projection_kf(fv, P["valid"], P["x"], P["y"], P["dist3d"], P["min_dist"], P["max_dist"], P["level"],
                                               P["angle"], P["desc"], th, orb_dist, ori)
        want, wnm = oracle.search_by_projection_kf(kp, desc, occ, fv.bounds_grid(), SCALE, P["valid"], P["x"], P["y"], P["dist3d"],
                                                   P["min_dist"], P["max_dist"], P["level"], P["angle"], P["desc"], th, orb_dist, ori)
        assert nm == wnm and np.array_equal(got, want)


def test_random_crowded_scenes(oracle):
    """Many small crowded scenes: out-of-order decisions must never leak into earlier map points (versioned taken state)."""
    for seed in range(100, 160):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(20, 400)); n_pts = int(rng.integers(50, 900))
        kp, desc, ur, occ, bounds = make_frame(rng, n, crowd=int(rng.integers(1, 6)), occupied_frac=0.05)
        # few distinct descriptors => many equal distances, second-best candidates shared between map points
        desc = desc[rng.integers(0, max(2, n // 8), n)]
        P = make_points(rng, kp, desc, ur, n_pts, dup_frac=0.8, max_flip=110, jitter=5.0)
        fv = view(kp, desc, ur, occ, bounds)
        th = float(rng.choice([1.0, 3.0, 8.0])); ratio = float(rng.choice([0.6, 0.8, 0.9]))
        got, nm = orbx.search_by_projection_map(fv, P["in_view"], P["bad"], P["x"], P["y"], P["xr"], P["view_cos"], P["depth"], P["level"],
                                                P["n_obs"], P["desc"], th=th, nnratio=ratio)
        want, wnm = oracle.search_by_projection_map(kp, desc, ur, occ, fv.bounds_grid(), SCALE, P["in_view"], P["bad"], P["x"], P["y"],
                                                    P["xr"], P["view_cos"], P["depth"], P["level"], P["n_obs"], P["desc"], th=th, nnratio=ratio)
        assert nm == wnm and np.array_equal(got, want), (seed, orbx.projection_rounds())
        got, nm = orbx.search_by_projection_last(fv, 40.0, P["valid"], P["x"], P["y"], P["invz"], P["level"], P["angle"], P["n_obs"], P["desc"],
                                                 th * 5, 0, 0, True)
        want, wnm = oracle.search_by_projection_last(kp, desc, ur, occ, fv.bounds_grid(), SCALE, 40.0, P["valid"], P["x"], P["y"], P["invz"],
                                                     P["level"], P["angle"], P["n_obs"], P["desc"], th * 5, 0, 0, True)
        assert nm == wnm and np.array_equal(got, want), (seed, orbx.projection_rounds())


def test_worst_case_chain_terminates(oracle):
    """Identical descriptors everywhere: every map point claims every feature in its window, so the speculation degenerates to
    one decision per round — the result must still equal the sequential loop."""
    rng = np.random.default_rng(51)
    kp, desc, ur, occ, bounds = make_frame(rng, 64, crowd=1, occupied_frac=0.0, outside_frac=0.0)
    desc[:] = desc[0]
    kp["octave"] = 2
    P = make_points(rng, kp, desc, ur, 200, dup_frac=0.9, max_flip=0, level_slop=False)
    P["n_obs"][:] = 1
    fv = view(kp, desc, None, occ, bounds)
    got, nm = orbx.search_by_projection_last(fv, 40.0, P["valid"], P["x"], P["y"], P["invz"], P["level"], P["angle"], P["n_obs"], P["desc"],
                                             30.0, 0, 0, False)
    want, wnm = oracle.search_by_projection_last(kp, desc, None, occ, fv.bounds_grid(), SCALE, 40.0, P["valid"], P["x"], P["y"], P["invz"],
                                                 P["level"], P["angle"], P["n_obs"], P["desc"], 30.0, 0, 0, False)
    assert nm == wnm and np.array_equal(got, want)
    assert orbx.projection_rounds() >= 10


def test_bad_arguments():
    rng = np.random.default_rng(61)
    kp, desc, ur, occ, bounds = make_frame(rng, 50)
    P = make_points(rng, kp, desc, ur, 20)
    fv = view(kp, desc, ur, occ, bounds)
    lvl = P["level"].copy(); lvl[P["in_view"].astype(bool) & ~P["bad"].astype(bool)] = 9
    with pytest.raises(orbx.OrbxError):
        orbx.search_by_projection_map(fv, P["in_view"], P["bad"], P["x"], P["y"], P["xr"], P["view_cos"], P["depth"], lvl, P["n_obs"], P["desc"])
    with pytest.raises(orbx.OrbxError):
        orbx.FrameView(kp, desc, SCALE, (0.0, 0.0, -1.0, 480.0)).assign_features_to_grid()     # negative grid element size


def test_latency_report(oracle):
    """Not a parity test: prints the per-call latency of a TrackLocalMap-sized search next to the oracle's (visible with -s)."""
    rng = np.random.default_rng(71)
    kp, desc, ur, occ, bounds = make_frame(rng, 1200)
    P = make_points(rng, kp, desc, ur, 2000)
    fv = view(kp, desc, ur, occ, bounds)
    args = (fv, P["in_view"], P["bad"], P["x"], P["y"], P["xr"], P["view_cos"], P["depth"], P["level"], P["n_obs"], P["desc"])
    for _ in range(5):
        orbx.search_by_projection_map(*args)
    t0 = time.perf_counter()
    for _ in range(50):
        orbx.search_by_projection_map(*args)
    gpu_us = (time.perf_counter() - t0) / 50 * 1e6
    t0 = time.perf_counter()
    for _ in range(20):
        oracle.search_by_projection_map(kp, desc, ur, occ, fv.bounds_grid(), SCALE, P["in_view"], P["bad"], P["x"], P["y"], P["xr"], P["view_cos"],
                                        P["depth"], P["level"], P["n_obs"], P["desc"])
    cpu_us = (time.perf_counter() - t0) / 20 * 1e6
    print("search_by_projection_map 1200 features x 2000 map points: GPU %.0f us / call (rounds %d), CPU oracle %.0f us"
          % (gpu_us, orbx.projection_rounds(), cpu_us))
