"""Randomised parity sweep (GPU vs CPU oracle): random shapes, feature counts, level counts, scale factors and FAST thresholds,
single frames and small batches, with every kernel variant.  usage: fuzz_parity.py [n_cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wut_cuda_orb_slam3_b200 as orbx
from wut_cuda_orb_slam3_b200 import synth
from tests import oracle_lib



def run(n_cases, seed0):
  rng = np.random.default_rng(seed0)
  o = oracle_lib.load()
  bad = 0
  for case in range(n_cases):
      cols, rows = int(rng.integers(90, 1400)), int(rng.integers(80, 800))
      if cols < rows * 0.6:
          cols, rows = rows, cols                       # keep nIni >= 1 (the reference divides by zero otherwise)
      nf = int(rng.integers(100, 3000)); nl = int(rng.integers(3, 9)); sf = float(rng.choice([1.2, 1.2, 1.15, 1.3, 1.5]))
      while min(cols, rows) / sf ** (nl - 1) < 45:     # the reference needs every level to hold its 16-px FAST window border
          nl -= 1
      ini, mn = int(rng.integers(10, 40)), int(rng.integers(3, 10))
      seed = int(rng.integers(0, 1 << 30))
      nb = int(rng.choice([1, 1, 9, 17]))
      lap = [(0, 0), (0, 1000), (cols // 3, 2 * cols // 3)][int(rng.integers(0, 3))]     # stereo / mono / fisheye-style lapping areas
      imgs = np.stack([synth.image(seed + f, cols, rows) for f in range(nb)])
      ex = orbx.ORBextractor(nf, sf, nl, ini, mn)
      oex = o.extractor(nf, sf, nl, ini, mn)
      if nb == 1:
          nm, kps, desc = ex(imgs[0], None, lap)
          res = [(nm, kps, desc)]
      else:
          nmv, nv, kb, db = ex.extract_batch(imgs, lap)
          res = [(int(nmv[f]), kb[f][:nv[f]], db[f][:nv[f]]) for f in range(nb)]
      for f in range(nb):
          ko, do, nmo = oex.extract(imgs[f], lap)
          nm, kps, desc = res[f]
          ok = nm == nmo and len(kps) == len(ko)
          if ok:
              for fld in ("x", "y", "size", "response", "octave"):
                  ok = ok and np.array_equal(kps[fld], ko[fld])
              dang = np.abs(kps["angle"] - ko["angle"]); dang = np.minimum(dang, 360 - dang)
              ok = ok and dang.max(initial=0) <= 1e-3 and (len(desc) == 0 or (desc == do).all(axis=1).mean() >= 0.995)
          if not ok:
              bad += 1
              print("MISMATCH case", case, dict(cols=cols, rows=rows, nf=nf, nl=nl, sf=sf, ini=ini, mn=mn, seed=seed, nb=nb, lap=lap, frame=f), flush=True)
      ex.close()
  print("fuzz: %d cases, %d mismatching frames (env %s)" % (n_cases, bad, {k: v for k, v in os.environ.items() if k.startswith("ORBX_")}))
  return bad


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 30, int(sys.argv[2]) if len(sys.argv) > 2 else 1) else 0)
