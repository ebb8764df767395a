"""Oracle of the frame grid / projection-guided searches against an independent brute-force Python restatement (CPU)."""
import numpy as np
import pytest

from tests import oracle_lib
from tests.proj_synth import SCALE, make_frame, make_points


@pytest.fixture(scope="module")
def oracle():
    return oracle_lib.load()


def bounds_grid(bounds):
    b = [np.float32(x) for x in bounds]
    return np.array(b + [np.float32(64) / (b[2] - b[0]), np.float32(48) / (b[3] - b[1])], np.float32)


def py_cells(kp, bg):
    """PosInGrid (src/Frame.cc:755-766) in float32; C round() = half away from zero."""
    def rnd(v):
        return np.where(v >= 0, np.floor(v + np.float32(0.5)), np.ceil(v - np.float32(0.5)))
    vx = (kp["x"] - bg[0]) * bg[4]; vy = (kp["y"] - bg[1]) * bg[5]
    # floor(v + 0.5) in float32 can differ from round() when v + 0.5 rounds up; use float64 on the exact float32 product
    px = rnd(vx.astype(np.float64)).astype(np.int64); py = rnd(vy.astype(np.float64)).astype(np.int64)
    ok = (px >= 0) & (px < 64) & (py >= 0) & (py < 48)
    return np.where(ok, px * 48 + py, -1)


def py_area(kp, cells, bg, x, y, r, min_level, max_level):
    """GetFeaturesInArea by brute force over all features, ordered by (cell id, index)."""
    x, y, r = np.float32(x), np.float32(y), np.float32(r)
    x0 = max(0, int(np.floor((x - bg[0] - r) * bg[4]))); x1 = min(63, int(np.ceil((x - bg[0] + r) * bg[4])))
    y0 = max(0, int(np.floor((y - bg[1] - r) * bg[5]))); y1 = min(47, int(np.ceil((y - bg[1] + r) * bg[5])))
    if x0 >= 64 or x1 < 0 or y0 >= 48 or y1 < 0:
        return []
    out = []
    check = min_level > 0 or max_level >= 0
    for i in range(len(kp)):
        c = cells[i]
        if c < 0 or not (x0 <= c // 48 <= x1 and y0 <= c % 48 <= y1):
            continue
        if check and (kp["octave"][i] < min_level or (max_level >= 0 and kp["octave"][i] > max_level)):
            continue
        if abs(np.float32(kp["x"][i] - x)) < r and abs(np.float32(kp["y"][i] - y)) < r:
            out.append((c, i))
    return [i for _, i in sorted(out)]


def hamming(a, b):
    return int(np.unpackbits(a ^ b).sum())


def test_grid_matches_bruteforce(oracle):
    rng = np.random.default_rng(1)
    kp, desc, ur, occ, bounds = make_frame(rng, 1500)
    bg = bounds_grid(bounds)
    cs, items = oracle.assign_features_to_grid(kp, bg)
    cells = py_cells(kp, bg)
    assert cs[-1] == (cells >= 0).sum() and len(items) == cs[-1]
    for c in range(64 * 48):
        assert list(items[cs[c]:cs[c + 1]]) == list(np.nonzero(cells == c)[0])


def test_features_in_area_matches_bruteforce(oracle):
    rng = np.random.default_rng(2)
    kp, desc, ur, occ, bounds = make_frame(rng, 1200)
    bg = bounds_grid(bounds)
    cells = py_cells(kp, bg)
    for t in range(200):
        x, y = rng.uniform(-30, 780), rng.uniform(-30, 510)
        r = float(rng.choice([2.5, 4.0, 10.0, 30.0, 80.0]))
        lv = int(rng.integers(0, 8))
        mn, mx = [(-1, -1), (lv, -1), (0, lv), (lv - 1, lv + 1), (lv - 1, lv)][t % 5]
        got = oracle.get_features_in_area(kp, bg, x, y, r, mn, mx)
        assert list(got) == py_area(kp, cells, bg, x, y, r, mn, mx), (t, x, y, r, mn, mx)


def py_search_map(kp, desc, ur, occ, bg, P, th, nnratio):
    """Sequential restatement of src/ORBmatcher1.cc:45-215 over py_area (independent of the oracle's grid code)."""
    cells = py_cells(kp, bg)
    taken = occ.astype(bool).copy()
    match = np.full(len(kp), -1, np.int32)
    nm = 0
    for i in range(len(P["x"])):
        if not P["in_view"][i] or P["bad"][i]:
            continue
        lvl = int(P["level"][i])
        r = np.float32(2.5) if np.float64(P["view_cos"][i]) > 0.998 else np.float32(4.0)
        if th != 1.0:
            r = np.float32(r * np.float32(th))
        rr = np.float32(r * SCALE[lvl])
        best, best2, bl, bl2, bi = 256, 256, -1, -1, -1
        for idx in py_area(kp, cells, bg, P["x"][i], P["y"][i], rr, lvl - 1, lvl):
            if taken[idx]:
                continue
            if ur[idx] > 0 and abs(np.float32(P["xr"][i] - ur[idx])) > rr:
                continue
            d = hamming(P["desc"][i], desc[idx])
            if d < best:
                best2, bl2 = best, bl
                best, bl, bi = d, int(kp["octave"][idx]), idx
            elif d < best2:
                best2, bl2 = d, int(kp["octave"][idx])
        if best <= 100:
            if bl == bl2 and np.float32(best) > np.float32(nnratio) * np.float32(best2):
                continue
            match[bi] = i
            nm += 1
            taken[bi] = P["n_obs"][i] > 0
    return match, nm


@pytest.mark.parametrize("seed,th,crowd", [(3, 1.0, 0), (4, 3.0, 0), (5, 5.0, 12)])
def test_search_map_matches_python(oracle, seed, th, crowd):
    rng = np.random.default_rng(seed)
    kp, desc, ur, occ, bounds = make_frame(rng, 400, crowd=crowd)
    bg = bounds_grid(bounds)
    P = make_points(rng, kp, desc, ur, 300, dup_frac=0.5)
    got, nm = oracle.search_by_projection_map(kp, desc, ur, occ, bg, SCALE, P["in_view"], P["bad"], P["x"], P["y"], P["xr"], P["view_cos"],
                                              P["depth"], P["level"], P["n_obs"], P["desc"], th=th, nnratio=0.8)
    want, wnm = py_search_map(kp, desc, ur, occ, bg, P, th, 0.8)
    assert nm == wnm and nm > 20
    assert np.array_equal(got, want)


def test_search_last_and_kf_run_and_are_consistent(oracle):
    """1-NN variants: every accepted match is within the threshold, inside the window and on an un-occupied feature."""
    rng = np.random.default_rng(7)
    kp, desc, ur, occ, bounds = make_frame(rng, 600)
    bg = bounds_grid(bounds)
    P = make_points(rng, kp, desc, ur, 500)
    for fwd, bwd in ((0, 0), (1, 0), (0, 1)):
        m, nm = oracle.search_by_projection_last(kp, desc, ur, occ, bg, SCALE, 40.0, P["valid"], P["x"], P["y"], P["invz"], P["level"],
                                                 P["angle"], P["n_obs"], P["desc"], 15.0, fwd, bwd, True)
        assert nm == (m >= 0).sum() or (P["n_obs"] == 0).any()
        for f in np.nonzero(m >= 0)[0]:
            i = m[f]
            assert not occ[f] and hamming(P["desc"][i], desc[f]) <= 100
            assert abs(kp["x"][f] - P["x"][i]) < 15.0 * SCALE[P["level"][i]]
    m, nm = oracle.search_by_projection_kf(kp, desc, occ, bg, SCALE, P["valid"], P["x"], P["y"], P["dist3d"], P["min_dist"], P["max_dist"],
                                           P["level"], P["angle"], P["desc"], 10.0, 64, True)
    assert nm == (m >= 0).sum() and nm > 50
    for f in np.nonzero(m >= 0)[0]:
        assert not occ[f] and hamming(P["desc"][m[f]], desc[f]) <= 64
