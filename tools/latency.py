"""Single-frame latency of orbx_extract (host image in, keypoints + descriptors out), the SLAM tracking use case."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import wut_cuda_orb_slam3_b200 as orbx
from wut_cuda_orb_slam3_b200 import synth
from wut_cuda_orb_slam3_b200.capi import lib, ptr, check

def measure(cols, rows, nf, reps=200, mirror=False):
    img_t = torch.empty((rows, cols), dtype=torch.uint8).pin_memory()
    img_t.copy_(torch.from_numpy(synth.image(77, cols, rows)))
    ex = orbx.ORBextractor(nf, 1.2, 8, 20, 7, max_cols=cols, max_rows=rows, max_batch=1)
    if mirror:
        ex.set_pyramid_mirror(True)
    cap = ex.max_keypoints()
    kps = torch.empty((cap, 7), dtype=torch.float32).pin_memory(); desc = torch.empty((cap, 32), dtype=torch.uint8).pin_memory()
    n = C.c_int(0); nm = C.c_int(0)
    def call():
        check(lib().orbx_extract(ex._h, ptr(img_t), rows, cols, cols, 0, 0, ptr(kps), ptr(desc), cap, C.byref(n), C.byref(nm)))
    for _ in range(20): call()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    return float(np.median(ts)), float(np.percentile(ts, 95)), n.value

def measure_stereo(cols, rows, nf, reps=200):
    """Frame.cc:124-143 in one call: two extractions on two streams + ComputeStereoMatches on the device-resident features."""
    L = torch.empty((rows, cols), dtype=torch.uint8).pin_memory(); R = torch.empty((rows, cols), dtype=torch.uint8).pin_memory()
    L.copy_(torch.from_numpy(synth.image(31, cols, rows, view=0, max_disp=40))); R.copy_(torch.from_numpy(synth.image(31, cols, rows, view=1, max_disp=40)))
    exL = orbx.ORBextractor(nf, 1.2, 8, 20, 7, max_cols=cols, max_rows=rows, max_batch=1)
    exR = orbx.ORBextractor(nf, 1.2, 8, 20, 7, max_cols=cols, max_rows=rows, max_batch=1)
    cap = exL.max_keypoints(rows, cols)
    kL = torch.empty((cap, 7), dtype=torch.float32).pin_memory(); kR = torch.empty((cap, 7), dtype=torch.float32).pin_memory()
    dL = torch.empty((cap, 32), dtype=torch.uint8).pin_memory(); dR = torch.empty((cap, 32), dtype=torch.uint8).pin_memory()
    u = torch.empty(cap, dtype=torch.float32).pin_memory(); d = torch.empty(cap, dtype=torch.float32).pin_memory()
    nL = C.c_int(0); nR = C.c_int(0)
    def call():
        check(lib().orbx_extract_stereo(exL._h, exR._h, ptr(L), ptr(R), rows, cols, cols, ptr(kL), ptr(dL), C.byref(nL), ptr(kR), ptr(dR), C.byref(nR),
                                        cap, 47.9, 435.2, ptr(u), ptr(d)))
    for _ in range(20): call()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    return float(np.median(ts)), float(np.percentile(ts, 95)), nL.value, nR.value, int((u[:nL.value] >= 0).sum())

if __name__ == "__main__":
    for (c, r, nf) in [(752, 480, 1000), (1241, 376, 2000), (1280, 720, 2000)]:
        med, p95, n = measure(c, r, nf)
        print("%dx%d nfeatures=%d: median %.1f us, p95 %.1f us, %d keypoints" % (c, r, nf, med, p95, n))
    for (c, r, nf) in [(752, 480, 1000), (1280, 720, 2000)]:
        med, p95, n = measure(c, r, nf, mirror=True)
        print("%dx%d nfeatures=%d with the host mirror of mvImagePyramid (8 bordered levels to pinned memory during the call): median %.1f us, p95 %.1f us" % (c, r, nf, med, p95))
    for (c, r, nf) in [(752, 480, 1200), (1241, 376, 2000)]:
        med, p95, nl, nr, m = measure_stereo(c, r, nf)
        print("stereo pair %dx%d nfeatures=%d (orbx_extract_stereo: 2 extractions + ComputeStereoMatches, host images in, everything out): "
              "median %.1f us, p95 %.1f us, %d | %d keypoints, %d stereo matches" % (c, r, nf, med, p95, nl, nr, m))
    # per-stage device time of the single-frame path (serial launch order; the production path forks per level)
    img = synth.image(77, 752, 480)
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_cols=752, max_rows=480, max_batch=1)
    for _ in range(10): ex(img)
    ex.profile_begin()
    for _ in range(50): ex(img)
    st, nchunks = ex.profile_end()
    print("single-frame device time per stage (us):", {k: round(1e3 * v / nchunks, 1) for k, v in st.items()}, "sum %.1f" % (1e3 * sum(st.values()) / nchunks))
