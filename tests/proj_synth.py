"""Synthetic tracking scenes for the projection-guided searches (tests only): a frame of key points on the 64x48 grid and
map points projected near them, with controllable crowding so that several map points compete for the same feature."""
import numpy as np

from tests.oracle_lib import KP_DTYPE

SCALE = np.cumprod(np.array([1.0] + [1.2] * 7, np.float32)).astype(np.float32)     # float32 chain, src/ORBextractor.cc:418-425


def flip_bits(rng, d, k):
    d = d.copy()
    if k > 0:
        pos = rng.choice(256, size=k, replace=False)
        for p in pos:
            d[p >> 3] ^= np.uint8(1 << (p & 7))
    return d


def make_frame(rng, n, w=752.0, h=480.0, stereo_frac=0.6, occupied_frac=0.1, outside_frac=0.03, crowd=0):
    """Key points (undistorted: a few fall outside the image bounds), descriptors, mvuRight, occupied flags."""
    kp = np.zeros(n, KP_DTYPE)
    if crowd:
        centres = rng.uniform([40, 40], [w - 40, h - 40], size=(crowd, 2))
        pts = centres[rng.integers(0, crowd, n)] + rng.normal(0, 6.0, size=(n, 2))
    else:
        pts = rng.uniform([0, 0], [w, h], size=(n, 2))
    out = rng.random(n) < outside_frac
    pts[out] += rng.choice([-1, 1], size=(out.sum(), 2)) * rng.uniform(0, 30, size=(out.sum(), 2)) + np.where(rng.random((out.sum(), 2)) < 0.5, -w, w) * 0.0
    pts[out, 0] = np.where(rng.random(out.sum()) < 0.5, -rng.uniform(0, 12, out.sum()), w + rng.uniform(0, 12, out.sum()))
    # some coordinates on exact half-cell positions to exercise round()
    kp["x"] = pts[:, 0].astype(np.float32); kp["y"] = pts[:, 1].astype(np.float32)
    snap = rng.random(n) < 0.05
    kp["x"][snap] = np.float32(w / 64.0) * (rng.integers(0, 64, snap.sum()) + 0.5)
    kp["octave"] = rng.integers(0, 8, n)
    kp["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    kp["size"] = 31.0; kp["response"] = 50.0; kp["class_id"] = -1
    desc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    ur = np.where(rng.random(n) < stereo_frac, kp["x"] - rng.uniform(1, 40, n), -1.0).astype(np.float32)
    occ = (rng.random(n) < occupied_frac).astype(np.uint8)
    bounds = (0.0, 0.0, w, h)
    return kp, desc, ur, occ, bounds


def make_points(rng, kp, desc, ur, n_pts, jitter=2.0, max_flip=70, dup_frac=0.3, level_slop=True, unique=False):
    """Map points aimed at frame features: position jitter, descriptor bit flips, several points per feature (dup_frac), or
    every point at its own feature (unique=True, n_pts <= len(kp): what a well-tracked frame looks like)."""
    n = len(kp)
    if unique:
        tgt = rng.permutation(n)[:n_pts]
    else:
        base = rng.integers(0, n, max(1, int(n_pts * (1 - dup_frac))))
        tgt = np.concatenate([base, rng.choice(base, n_pts - len(base))]) if n_pts > len(base) else base[:n_pts]
        rng.shuffle(tgt)
    P = {}
    P["target"] = tgt
    P["x"] = (kp["x"][tgt] + rng.normal(0, jitter, n_pts)).astype(np.float32)
    P["y"] = (kp["y"][tgt] + rng.normal(0, jitter, n_pts)).astype(np.float32)
    P["xr"] = np.where(ur[tgt] > 0, ur[tgt] + rng.normal(0, jitter * 1.5, n_pts), P["x"] - 10.0).astype(np.float32)
    lvl = kp["octave"][tgt].astype(np.int32)
    if level_slop:
        lvl = np.clip(lvl + rng.integers(-1, 2, n_pts), 0, 7)
    P["level"] = lvl.astype(np.int32)
    P["desc"] = np.stack([flip_bits(rng, desc[t], int(rng.integers(0, max_flip + 1))) for t in tgt]) if n_pts else np.zeros((0, 32), np.uint8)
    P["angle"] = ((kp["angle"][tgt] + rng.choice([0.0, 0.0, 0.0, 45.0, 180.0], n_pts) + rng.normal(0, 4, n_pts)) % 360).astype(np.float32)
    P["n_obs"] = np.where(rng.random(n_pts) < 0.15, 0, rng.integers(1, 6, n_pts)).astype(np.int32)
    P["view_cos"] = np.where(rng.random(n_pts) < 0.5, 0.9995, 0.9).astype(np.float32)
    P["depth"] = rng.uniform(0.5, 40.0, n_pts).astype(np.float32)
    P["in_view"] = (rng.random(n_pts) < 0.9).astype(np.uint8)
    P["bad"] = (rng.random(n_pts) < 0.05).astype(np.uint8)
    P["valid"] = (rng.random(n_pts) < 0.9).astype(np.uint8)
    P["invz"] = np.where(rng.random(n_pts) < 0.03, -0.1, 1.0 / P["depth"]).astype(np.float32)
    P["dist3d"] = P["depth"]
    P["min_dist"] = (P["depth"] * np.where(rng.random(n_pts) < 0.05, 1.2, 0.5)).astype(np.float32)
    P["max_dist"] = (P["depth"] * np.where(rng.random(n_pts) < 0.05, 0.8, 2.0)).astype(np.float32)
    return P
