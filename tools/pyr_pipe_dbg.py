import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wut_cuda_orb_slam3_b200 as orbx
from wut_cuda_orb_slam3_b200 import synth
cols, rows, F = 200, 150, 128
imgs = np.stack([synth.image(4000 + f, cols, rows) for f in range(F)])
ex = orbx.ORBextractor(300, 1.2, 6, 20, 7, max_cols=cols, max_rows=rows, max_batch=128)
nm, n, kps, desc = ex.extract_batch(imgs)
print("ok", n[:8])
