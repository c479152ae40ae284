#!/usr/bin/env python
"""Print the headline of a bench.py JSON line file: value, ms per step, per-op breakdown."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"], 4), {k: round(v, 3) for k, v in d.get("breakdown_ms_per_step", {}).items()})
