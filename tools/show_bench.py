#!/usr/bin/env python
"""Prints the interesting keys of bench.py JSON lines read from stdin."""
import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    out = {k: d.get(k) for k in ("impl", "n_gpus", "ms_per_step", "value", "tiles_per_sec", "gpu_launches") if k in d}
    if "e2e" in d and d["e2e"]: out["e2e"] = d["e2e"].get("value"); out["e2e_ms"] = d["e2e"].get("ms_per_step")
    if "stage_ms" in d: out["stage_ms"] = {k: round(v, 3) for k, v in d["stage_ms"].items()}
    if "cpu_baseline" in d and d["cpu_baseline"]: out["cpu"] = d["cpu_baseline"].get("value")
    if "roofline" in d: out["frac"] = d["roofline"].get("frac")
    if "clocks" in d: out["clocks"] = d["clocks"]
    print(json.dumps(out))
