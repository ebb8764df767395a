#!/usr/bin/env python3
"""Generates tests/golden/orb_golden.json from the CPU oracle (the reference ships no golden vectors for this path and
cannot be built/imported here — SURVEY.md §4, §8(c) — so these known-answer vectors pin the ORACLE against drift and give
the GPU tests a fixture that does not pass through oracle code at test time).  Run from the repo root:
    python tests/golden/make_golden.py
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import oracle_lib  # noqa: E402
from wut_cuda_orb_slam3_b200 import synth  # noqa: E402

CASES = [dict(cols=160, rows=120, nfeatures=300, seed=105, lapping=[0, 0]),
         dict(cols=752, rows=480, nfeatures=1000, seed=101, lapping=[0, 1000]),
         dict(cols=1241, rows=376, nfeatures=2000, seed=103, lapping=[0, 0])]


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def main():
    o = oracle_lib.load()
    out = {"generator": "tests/golden/make_golden.py (CPU oracle, oracle/orb_oracle.cpp)", "cases": []}
    for c in CASES:
        img = synth.image(c["seed"], c["cols"], c["rows"])
        ex = o.extractor(c["nfeatures"], 1.2, 8, 20, 7)
        kps, desc, nm = ex.extract(img, tuple(c["lapping"]))
        rec = dict(c)
        rec.update(image_crc=crc(img), n=len(kps), n_mono=nm, keypoints_crc=crc(kps), descriptors_crc=crc(desc), levels=[])
        for l in range(8):
            xs, ys, sc = ex.candidates(l)
            lk, ld = ex.level_keypoints(l)
            rec["levels"].append(dict(size=list(ex.level_size(l)), pyramid_crc=crc(ex.pyramid_level(l, with_border=True)),
                                      blur_crc=crc(ex.blurred_level(l)), n_candidates=len(xs),
                                      candidates_crc=crc(np.stack([xs, ys, sc], 1).astype(np.int32)), n_keypoints=len(lk),
                                      keypoints_xy_crc=crc(np.stack([lk["x"], lk["y"], lk["response"]], 1).astype(np.float32)),
                                      angles_crc=crc(lk["angle"]), desc_crc=crc(ld)))
        if c["cols"] == 160:   # full small case inline
            rec["keypoints"] = [[float(k["x"]), float(k["y"]), float(k["size"]), float(k["angle"]), float(k["response"]), int(k["octave"])] for k in kps[:40]]
            rec["descriptors_hex"] = [bytes(d).hex() for d in desc[:40]]
        out["cases"].append(rec)
    db = synth.descriptors(21, 4000)
    q = synth.descriptors(21, 1000, is_query=True, ndb=4000, plant_every=3)
    idx, dist = o.knn2(q, db)
    out["knn2"] = dict(seed=21, nq=1000, ndb=4000, plant_every=3, idx_crc=crc(idx), dist_crc=crc(dist), first=[idx[:8].tolist(), dist[:8].tolist()])
    with open(os.path.join(ROOT, "tests", "golden", "orb_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
