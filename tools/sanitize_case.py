"""Small end-to-end exercise for compute-sanitizer (memcheck / racecheck): two image shapes, batch API, knn2, stereo, octree."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wut_cuda_orb_slam3_b200 as orbx
from wut_cuda_orb_slam3_b200 import synth

for (cols, rows, nf) in [(752, 480, 1000), (331, 277, 500), (160, 120, 300)]:
    ex = orbx.ORBextractor(nf, 1.2, 8, 20, 7)
    img = synth.image(3, cols, rows)
    nm, kps, desc = ex(img, None, (0, 0))
    print(cols, rows, len(kps))
imgs = np.stack([synth.image(10 + f, 640, 480) for f in range(5)])
ex = orbx.ORBextractor(800, 1.2, 8, 20, 7, max_batch=2)
nm, n, kps, desc = ex.extract_batch(imgs)
print("batch", n.tolist())
L = synth.image(31, 752, 480, view=0); R = synth.image(31, 752, 480, view=1)
exL = orbx.ORBextractor(1200, 1.2, 8, 20, 7); exR = orbx.ORBextractor(1200, 1.2, 8, 20, 7)
_, kL, dL = exL(L, None, (0, 0)); _, kR, dR = exR(R, None, (0, 0))
u, d = orbx.compute_stereo_matches(exL, exR, kL, dL, kR, dR, 47.9, 435.0)
print("stereo matches", int((u >= 0).sum()))
db = synth.descriptors(5, 30000); q = synth.descriptors(5, 700, is_query=True, ndb=30000, plant_every=2)
idx, dist = orbx.ORBmatcher().knn2(q, db)
print("knn2", idx[:2].tolist())
rng = np.random.default_rng(0)
pts = np.unique(np.stack([rng.integers(0, 400, 3000), rng.integers(0, 700, 3000)], 1), axis=0)
out = orbx.distribute_octree(pts[:, 1].astype(np.int32), pts[:, 0].astype(np.int32), rng.integers(7, 50, len(pts)).astype(np.int32), 16, 716, 16, 416, 200)
print("octree", len(out))
