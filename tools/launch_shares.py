#!/usr/bin/env python3
"""Compares the per-kernel shares of an ncu launch list (gpu__time_duration.sum, --clock-control none) with the per-stage
CUDA-event shares bench.py measured inside its timed region.  A step = the launches from pyr_level0* to pack_kernel; only
steps whose octree grid covers `frames` frames are counted.  usage: launch_shares.py launches.csv bench.json frames"""
import csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, im, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Grid Size")
bench = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
frames = int(sys.argv[3])
stage_of = [("pyr_", "pyramid"), ("fast_cells", "fast"), ("blur_", "blur"), ("octree_kernel", "octree"), ("orient_describe", "orient_describe"),
            ("pack_kernel", "pack")]
steps, cur, ok = [], None, False
for r in rows:
    if r is hdr or r[im] != "gpu__time_duration.sum":
        continue
    name = r[ik]
    us = float(r[iv].replace(",", "")) / 1e3
    if "pyr_level0" in name:
        cur, ok = [], False
    if cur is None:
        continue
    cur.append((name, us))
    if "octree_kernel" in name:
        ok = int(r[ig].strip("()").split(",")[0]) == frames
    if "pack_kernel" in name:
        if ok:
            steps.append(cur)
        cur = None
n = max(len(steps), 1)
per_kernel, per_stage = {}, {}
for st in steps:
    for name, us in st:
        short = name.split("(")[0].replace("orbx::", "").replace("void ", "")
        per_kernel.setdefault(short, []).append(us)
        for key, sname in stage_of:
            if key in name:
                per_stage[sname] = per_stage.get(sname, 0.0) + us
print("ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches); %d complete step(s) of %d frames" % (len(steps), frames))
for k, v in sorted(per_kernel.items(), key=lambda kv: -sum(kv[1])):
    print("%-40s launches/step=%4.1f mean_us=%9.1f total_us/step=%9.1f" % (k, len(v) / n, sum(v) / len(v), sum(v) / n))
tot = sum(per_stage.values()) / n
btot = sum(s["ms_per_step"] for s in bench["extra"]["stages"].values()) * 1e3
print("\n%-16s %12s %10s %14s %12s" % ("stage", "ncu_us/step", "ncu_share", "bench_us/step", "bench_share"))
for _, st in stage_of:
    a = per_stage.get(st, 0.0) / n
    b = bench["extra"]["stages"][st]["ms_per_step"] * 1e3
    print("%-16s %12.1f %10.3f %14.1f %12.3f" % (st, a, a / tot, b, b / btot))
