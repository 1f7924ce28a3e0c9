#!/usr/bin/env python
"""Print the headline numbers of a bench.py JSON line read from stdin: tools/bench_summary.py [label]"""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
label = sys.argv[1] if len(sys.argv) > 1 else ""
print(label, "ms/frame %.4f  fps %.1f  e2e fps %.1f" % (d["ms_per_step"], d["fps"], d["e2e"]["fps"]),
      {k: round(v["launch_ms_avg"], 4) for k, v in d.get("stages", {}).items() if v["launch_ms_avg"]})
if "stages_serial" in d:
    print("  serial:", {k: round(v["launch_ms_avg"], 4) for k, v in d["stages_serial"].items() if v["launch_ms_avg"]})
if "me_ms_by_temporal_layer" in d:
    print("  ME ms by temporal layer (serial):", {k: round(v, 4) for k, v in d["me_ms_by_temporal_layer"].items()})
