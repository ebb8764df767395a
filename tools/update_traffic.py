#!/usr/bin/env python3
"""profiles/traffic.json from an ncu csv with dram__bytes_read.sum / dram__bytes_write.sum of the FAST launches of ONE bench step
(the warp-per-cell kernel is launched once per group of pyramid levels).  usage: update_traffic.py dram.csv frames source-note"""
import csv, json, os, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
iv, im, iu = hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Metric Unit")
mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = wr = 0.0
n = 0
for r in rows:
    if r is hdr:
        continue
    v = float(r[iv].replace(",", "")) * mult.get(r[iu], 1.0)
    if r[im] == "dram__bytes_read.sum":
        rd += v; n += 1
    elif r[im] == "dram__bytes_write.sum":
        wr += v
frames = int(sys.argv[2])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(root, "profiles", "traffic.json")
t = json.load(open(path)) if os.path.exists(path) else {}
t["fast"] = {"dram_bytes_per_frame": (rd + wr) / frames, "dram_read": rd, "dram_write": wr, "frames": frames, "launches": n, "source": sys.argv[3]}
json.dump(t, open(path, "w"), indent=1)
print(t["fast"])
