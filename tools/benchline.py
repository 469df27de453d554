"""Prints the interesting fields of one bench.py JSON line read from stdin (for gpurun one-liners)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
src = open(sys.argv[2]).read() if len(sys.argv) > 2 else sys.stdin.read()
line = [l for l in src.strip().splitlines() if l.startswith("{")][-1]
j = json.loads(line)
r = j.get("roofline") or {}
print(tag, "ms/step", round(j["ms_per_step"], 4), "value %.4g" % j["value"], "e2e", (j.get("e2e") or {}).get("value"),
      "frac", r.get("frac"), "kernels_ms", r.get("kernels_ms"))
