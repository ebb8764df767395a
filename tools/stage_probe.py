"""Per-stage device times of the resident batch path for an arbitrary shape.  usage: stage_probe.py cols rows nfeatures batch"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wut_cuda_orb_slam3_b200 as orbx
from wut_cuda_orb_slam3_b200 import synth
COLS, ROWS, NF, B = [int(x) for x in sys.argv[1:5]]
dev = torch.device("cuda:0")
d_img = torch.empty((B, ROWS, COLS), dtype=torch.uint8, device=dev)
synth.images_device(d_img, 9000, B, COLS, ROWS, COLS, ROWS * COLS, device=0)
ex = orbx.ORBextractor(NF, 1.2, 8, 20, 7, device=0, max_cols=COLS, max_rows=ROWS, max_batch=B)
cap = ex.max_keypoints()
d_kps = torch.zeros((B, cap, 7), dtype=torch.float32, device=dev); d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
d_n = torch.zeros(B, dtype=torch.int32, device=dev); d_nm = torch.zeros(B, dtype=torch.int32, device=dev)
ts = torch.cuda.Stream(device=dev); torch.cuda.synchronize(); torch.cuda.set_stream(ts)
def step():
    ex.extract_batch_device(d_img, B, ROWS, COLS, COLS, ROWS * COLS, d_kps, d_desc, cap, d_n, d_nm, (0, 0), stream=ts.cuda_stream)
for _ in range(3): step()
torch.cuda.synchronize()
ex.profile_begin()
for _ in range(5): step()
torch.cuda.synchronize()
st, n = ex.profile_end()
tot = sum(st.values()) / n
print("%dx%d nf=%d batch=%d: %.3f ms/step = %.0f frames/s;" % (COLS, ROWS, NF, B, tot, B / tot * 1e3), {k: round(v / n, 3) for k, v in st.items()},
      "kp/frame %.0f" % d_n.float().mean().item())
