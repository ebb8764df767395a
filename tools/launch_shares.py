#!/usr/bin/env python3
"""Compares the per-kernel shares of an ncu launch list (gpu__time_duration.sum, --clock-control none) with the per-stage
CUDA-event shares bench.py measured inside its timed region.  usage: launch_shares.py launches.csv bench.json frames_per_launch"""
import csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
ig = hdr.index("Grid Size") if "Grid Size" in hdr else None
bench = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
frames = int(sys.argv[3])
stage_of = [("pyr_", "pyramid"), ("fast_cells", "fast"), ("blur_kernel", "blur"), ("octree_kernel", "octree"), ("orient_describe", "orient_describe"),
            ("pack_kernel", "pack")]
per_kernel, per_stage, steps = {}, {}, 0
for r in rows:
    if r is hdr or r[im] != "gpu__time_duration.sum":
        continue
    name = r[ik]
    if ig is not None:
        g = [int(x) for x in r[ig].strip("()").split(",")]
        if g[-1] != frames and g[1] != frames and g[0] != frames:      # keep only launches over the full batch
            continue
    us = float(r[iv].replace(",", "")) / 1e3
    short = name.split("(")[0].replace("orbx::", "").replace("void ", "")
    per_kernel.setdefault(short, []).append(us)
    for key, st in stage_of:
        if key in name:
            per_stage[st] = per_stage.get(st, 0.0) + us
    if "pack_kernel" in name:
        steps += 1
steps = max(steps, 1)
print("ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches); %d step(s) of %d frames captured" % (steps, frames))
for k, v in sorted(per_kernel.items(), key=lambda kv: -sum(kv[1])):
    print("%-48s launches=%3d mean_us=%9.1f total_us/step=%9.1f" % (k, len(v), sum(v) / len(v), sum(v) / steps))
tot = sum(per_stage.values()) / steps
btot = sum(s["ms_per_step"] for s in bench["extra"]["stages"].values()) * 1e3
print("\n%-16s %12s %10s %14s %12s" % ("stage", "ncu_us/step", "ncu_share", "bench_us/step", "bench_share"))
for _, st in stage_of:
    n = per_stage.get(st, 0.0) / steps
    b = bench["extra"]["stages"][st]["ms_per_step"] * 1e3
    print("%-16s %12.1f %10.3f %14.1f %12.3f" % (st, n, n / tot, b, b / btot))
