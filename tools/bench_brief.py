"""Reads bench.py's JSON line from stdin and prints the few numbers an A/B run is about."""
import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    st = {k: round(v["ms_per_step"], 3) for k, v in d.get("extra", {}).get("stages", {}).items()}
    c4 = d.get("cfg4") or {}
    print(sys.argv[1] if len(sys.argv) > 1 else "", "resident %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], st,
          "cfg4", c4.get("frames_per_s"), (c4.get("e2e") or {}).get("frames_per_s"), "checksum", c4.get("checksum"), "matches_n1", c4.get("checksum_matches_n1"))
