#!/usr/bin/env python3
"""Print the key numbers of a bench.py JSON line (last line of the given log)."""
import json
import sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4g %s | ms/step %.2f | stages %s | e2e %.4g | verified %s | roofline %.3f | launches %s" % (
    d["value"], d["unit"], d["ms_per_step"], d.get("stage_ms"), d.get("e2e", {}).get("value", 0),
    d.get("results_verified"), d.get("roofline", {}).get("frac", 0), d.get("gpu_launches")))
if d.get("cpu_baseline"):
    print("cpu_baseline", d["cpu_baseline"])
