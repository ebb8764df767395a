"""Prints the per-chunk timeline of orbx_extract_batch (ORBX_TRACE_BATCH=1) for one 512-frame call."""
import ctypes as C, os, sys, time
os.environ["ORBX_TRACE_BATCH"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wut_cuda_orb_slam3_b200 as orbx
from wut_cuda_orb_slam3_b200 import synth
from wut_cuda_orb_slam3_b200.capi import lib, ptr, check
B, COLS, ROWS = int(os.environ.get("B", 512)), 752, 480
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 64
d_img = torch.empty((B, ROWS, COLS), dtype=torch.uint8, device="cuda")
synth.images_device(d_img, 1000, B, COLS, ROWS, COLS, ROWS * COLS, device=0)
h_img = torch.empty((B, ROWS, COLS), dtype=torch.uint8).pin_memory(); h_img.copy_(d_img)
ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, device=0, max_cols=COLS, max_rows=ROWS, max_batch=chunk)
cap = ex.max_keypoints()
h_kps = torch.empty((B, cap, 7), dtype=torch.float32).pin_memory(); h_desc = torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory()
h_n = torch.empty(B, dtype=torch.int32).pin_memory(); h_nm = torch.empty(B, dtype=torch.int32).pin_memory()
ptrs = (C.c_void_p * B)(*[h_img.data_ptr() + f * ROWS * COLS for f in range(B)])
N = int(os.environ.get("CALLS", 4))
for i in range(N):
    if i == N - 1: sys.stderr.write("---- traced call ----\n")
    t0 = time.perf_counter()
    check(lib().orbx_extract_batch(ex._h, ptrs, B, ROWS, COLS, COLS, 0, 0, ptr(h_kps), ptr(h_desc), cap, ptr(h_n), ptr(h_nm)))
    dt = time.perf_counter() - t0
    sys.stdout.write("call %d: %.3f ms (%.0f frames/s)\n" % (i, dt * 1e3, B / dt))
sys.stderr.write("wall %.3f ms\n" % (dt * 1e3))
