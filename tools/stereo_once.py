"""A few orbx_extract_stereo calls on the EuRoC-shape pair (for ncu launch lists of the single-frame / stereo path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wut_cuda_orb_slam3_b200 as orbx
from wut_cuda_orb_slam3_b200 import synth
L = synth.image(31, 752, 480, view=0, max_disp=40); R = synth.image(31, 752, 480, view=1, max_disp=40)
exL = orbx.ORBextractor(1200, 1.2, 8, 20, 7); exR = orbx.ORBextractor(1200, 1.2, 8, 20, 7)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    r = orbx.extract_stereo(exL, exR, L, R, 47.9, 435.2)
print(len(r[0]), len(r[2]), int((r[4] >= 0).sum()))
