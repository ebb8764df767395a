"""Times the resident batch path (orbx_extract_batch_device) without per-stage events.  usage: resident_probe.py [batch]
(Measured with this probe and rejected: running the blur on a side stream beside FAST + octree — 5.125 vs 5.130 ms at 512 frames,
0.951 vs 0.961 ms at 64 — and splitting a batch over 2-6 streams — 5.51 vs 5.38 ms: the kernels fill the machine one at a time.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wut_cuda_orb_slam3_b200 as orbx
from wut_cuda_orb_slam3_b200 import synth

B, COLS, ROWS = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 752, 480
dev = torch.device("cuda:0")
d_img = torch.empty((B, ROWS, COLS), dtype=torch.uint8, device=dev)
synth.images_device(d_img, 1000, B, COLS, ROWS, COLS, ROWS * COLS, device=0)
ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, device=0, max_cols=COLS, max_rows=ROWS, max_batch=B)
cap = ex.max_keypoints()
d_kps = torch.zeros((B, cap, 7), dtype=torch.float32, device=dev); d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
d_n = torch.zeros(B, dtype=torch.int32, device=dev); d_nm = torch.zeros(B, dtype=torch.int32, device=dev)
ts = torch.cuda.Stream(device=dev); torch.cuda.synchronize(); torch.cuda.set_stream(ts)
def step():
    ex.extract_batch_device(d_img, B, ROWS, COLS, COLS, ROWS * COLS, d_kps, d_desc, cap, d_n, d_nm, (0, 0), stream=ts.cuda_stream)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("B=%d: %.3f ms/step, %.0f frames/s, checksum %d" % (B, ms, B / ms * 1e3, int(d_n.sum().item()) + int(d_desc.sum().item())))
