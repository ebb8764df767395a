import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wut_cuda_orb_slam3_b200 as orbx
from wut_cuda_orb_slam3_b200 import synth
from wut_cuda_orb_slam3_b200.capi import lib, ptr, check
L = lib()
L.orbx_debug_octree_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p]
img = synth.image(1000, 752, 480)
ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
ex(img, None, (0, 0))
names = ["codes+count+scan", "bin scatter", "roots", "phase1 sweeps", "introsort", "phase2 rest", "retain", "#sweeps", "#p2 rounds", "n", "m max"]
# NB: the batch path inside orbx_debug_octree_timing (ORBX_DBG_REPLICAS) still uses 256-thread CTAs; a single frame gets 1024
for level in (0, 3, 7):
    out = np.zeros(16, np.int64)
    check(L.orbx_debug_octree_timing(ex._h, ptr(img), 480, 752, 752, level, ptr(out)))
    tot = out[:7].sum()
    print("level", level, "total cycles", tot, "= %.1f us @1.9GHz" % (tot / 1.9e3))
    for i, nme in enumerate(names):
        print("   %-14s %10d %s" % (nme, out[i], ("%.1f%%" % (100.0 * out[i] / tot)) if i < 7 else ""))
    print("   first phase in detail (cycles): zero + cell counts %d, cell-offset scan %d, gather %d, bins %d, bin scan %d" % tuple(out[11:16]))
