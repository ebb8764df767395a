"""The C++ drop-in adapter (csrc/adapter/ORBextractor.h: same class/signatures as the reference's include/ORBextractor.h)
compiled against a cv stub, run as Frame::ExtractORB would, must produce exactly what the C ABI returns through ctypes."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "adapter_test.cpp")
EXE = os.path.join(ROOT, "build", "adapter_test")


def build_exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "tests", "cpp", "cv_stub"), "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(ROOT, "wut_cuda_orb_slam3_b200", "csrc", "adapter"), "-o", EXE, SRC,
                           "-L" + os.path.join(ROOT, "wut_cuda_orb_slam3_b200"), "-lorbx",
                           "-Wl,-rpath," + os.path.join(ROOT, "wut_cuda_orb_slam3_b200")])


def fnv(h, data):
    for b in bytes(data):
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_adapter_compiles_against_header():
    """CPU: the adapter + include/orbx.h compile and link (no compute call)."""
    build_exe()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_adapter_matches_c_abi():
    import wut_cuda_orb_slam3_b200 as orbx
    from wut_cuda_orb_slam3_b200 import synth
    build_exe()
    out = subprocess.check_output([EXE, "3"], text=True)
    m = re.search(r"mono=(-?\d+) n=(\d+) levels=(\d+) scale1=([\d.]+) kphash=([0-9a-f]+) pyrhash=([0-9a-f]+) empty=(-?\d+) dist01=(-?\d+)", out)
    assert m, out
    img = synth.image(3, 752, 480)
    ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    nm, kps, desc = ex(img, None, (0, 1000))
    h = fnv(1469598103934665603, kps.tobytes())
    h = fnv(h, desc.tobytes())
    hp = 1469598103934665603
    for l in range(8):
        hp = fnv(hp, ex.pyramid_level(l).tobytes())
    assert int(m.group(1)) == nm and int(m.group(2)) == len(kps) and int(m.group(3)) == 8
    assert abs(float(m.group(4)) - 1.2000000477) < 1e-7
    assert m.group(5) == "%016x" % h
    assert m.group(6) == "%016x" % hp
    assert int(m.group(7)) == -1                      # empty image -> -1, like the reference
    assert int(m.group(8)) == orbx.ORBmatcher.DescriptorDistance(desc[0], desc[1])
