"""How many distinct 32-byte DRAM sectors do orientation + descriptors have to touch?  (CPU, oracle only.)
For one synthetic 752x480 / 1000-feature frame: every keypoint reads its umax disc on the un-blurred level (IC_Angle,
Angle.cl:24-53) and the 512 rotated pattern samples on the blurred level (computeOrbDescriptor, src/ORBextractor.cc:105-149).
With the product's buffer layout (bordered pyramid rows of align16(32 + w + 19) bytes, blurred rows of align16(w)) the union of
the sectors these reads fall into is the least traffic the stage can cause when a frame's levels are no longer in L2 (a
512-frame batch is ~1.1 GB of pyramid + blur) — SURVEY.md §8(d)'s 36 KB per frame counts outputs only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import oracle_lib

o = oracle_lib.load()
img = o.synth_image(1000, 752, 480)
ex = o.extractor(1000, 1.2, 8, 20, 7)
kps, desc, nm = ex.extract(img, (0, 0))
umax = [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
import re
_src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "wut_cuda_orb_slam3_b200", "csrc", "brief_pattern.inc")).read()
_src = "\n".join(ln.split("//")[0] for ln in _src.splitlines())
pat = np.array([int(v) for v in re.findall(r"-?\d+", _src)], np.int32).reshape(-1, 2)
assert pat.shape == (512, 2)
tot_pyr = tot_blur = 0
bytes_pyr = bytes_blur = 0
for level in range(8):
    lk, _ = ex.level_keypoints(level)
    w, h = ex.level_size(level)
    pitch = (32 + w + 19 + 15) // 16 * 16
    bpitch = (w + 15) // 16 * 16
    bytes_pyr += pitch * (h + 38); bytes_blur += bpitch * h
    sp, sb = set(), set()
    for k in lk:
        x, y, ang = int(k["x"]), int(k["y"]), float(k["angle"])
        for v in range(-15, 16):
            u = umax[abs(v)]
            row = (y + v + 19) * pitch + 32
            a0, a1 = (row + x - u) // 32, (row + x + u) // 32
            sp.update(range(a0, a1 + 1))
        a = np.float32(np.cos(np.float32(ang) * np.float32(np.pi / 180.0))); b = np.float32(np.sin(np.float32(ang) * np.float32(np.pi / 180.0)))
        px = pat[:, 0].astype(np.float32); py = pat[:, 1].astype(np.float32)
        iy = np.rint(px * b + py * a).astype(np.int64); ix = np.rint(px * a - py * b).astype(np.int64)
        sb.update((((y + iy) * bpitch + x + ix) // 32).tolist())
    tot_pyr += len(sp); tot_blur += len(sb)
print("keypoints %d; distinct 32-byte sectors: pyramid %d (%.2f MB of %.2f MB), blurred %d (%.2f MB of %.2f MB); union traffic %.2f MB per frame"
      % (len(kps), tot_pyr, tot_pyr * 32 / 1e6, bytes_pyr / 1e6, tot_blur, tot_blur * 32 / 1e6, bytes_blur / 1e6, (tot_pyr + tot_blur) * 32 / 1e6))
