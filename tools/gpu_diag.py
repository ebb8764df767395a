"""Stage-by-stage GPU-vs-oracle diagnostic (prints, never asserts)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wut_cuda_orb_slam3_b200 as orbx
from wut_cuda_orb_slam3_b200 import synth
from tests import oracle_lib

o = oracle_lib.load()
cfgs = [(752, 480, 1000, 101), (160, 120, 300, 105), (1241, 376, 2000, 103)]
if len(sys.argv) > 1:
    cfgs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for (cols, rows, nf, seed) in cfgs:
    print("=== %dx%d nf=%d seed=%d" % (cols, rows, nf, seed))
    img = synth.image(seed, cols, rows)
    ex = orbx.ORBextractor(nf, 1.2, 8, 20, 7)
    oex = o.extractor(nf, 1.2, 8, 20, 7)
    nm, kps, desc = ex(img, None, (0, 0))
    okps, odesc, onm = oex.extract(img, (0, 0))
    print("n", len(kps), len(okps), "nmono", nm, onm)
    for level in range(8):
        a = ex.pyramid_level(level, with_border=True); b = oex.pyramid_level(level, with_border=True)
        pyr_bad = int((a != b).sum()) if a.shape == b.shape else -1
        a = ex.blurred_level(level); b = oex.blurred_level(level)
        blur_bad = int((a != b).sum()) if a.shape == b.shape else -1
        xs, ys, sc = ex.candidates(level); oxs, oys, osc = oex.candidates(level)
        cand_same = len(xs) == len(oxs) and np.array_equal(xs, oxs) and np.array_equal(ys, oys) and np.array_equal(sc, osc)
        sg = set(zip(xs.tolist(), ys.tolist(), sc.tolist())); so = set(zip(oxs.tolist(), oys.tolist(), osc.tolist()))
        lk, ld = ex.level_keypoints(level); olk, old = oex.level_keypoints(level)
        kp_same = len(lk) == len(olk) and np.array_equal(lk["x"], olk["x"]) and np.array_equal(lk["y"], olk["y"]) and np.array_equal(lk["response"], olk["response"])
        skg = set(zip(lk["x"].tolist(), lk["y"].tolist())); sko = set(zip(olk["x"].tolist(), olk["y"].tolist()))
        ang_bad = -1; desc_bad = -1
        if kp_same:
            ang_bad = float(np.abs(lk["angle"] - olk["angle"]).max(initial=0)); desc_bad = int((ld != old).any(axis=1).sum())
        print(" L%d pyr_bad=%d blur_bad=%d cand n=%d/%d same=%s (only_gpu=%d only_ref=%d) kps n=%d/%d same=%s (set diff %d/%d) ang_maxdiff=%s desc_bad=%s" % (
            level, pyr_bad, blur_bad, len(xs), len(oxs), cand_same, len(sg - so), len(so - sg), len(lk), len(olk), kp_same, len(skg - sko), len(sko - skg), ang_bad, desc_bad))
        if not kp_same and cand_same and len(lk) == len(olk):
            first = next(i for i in range(len(lk)) if (lk["x"][i], lk["y"][i]) != (olk["x"][i], olk["y"][i]))
            print("    first order diff at", first, "of", len(lk))
    if len(kps) == len(okps):
        for f in ("x", "y", "size", "response", "octave", "class_id", "angle"):
            print("  final", f, "mismatches", int((kps[f] != okps[f]).sum()))
        print("  final desc rows differing", int((desc != odesc).any(axis=1).sum()))
