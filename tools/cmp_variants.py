#!/usr/bin/env python
"""Side-by-side %HBM of bk_bench JSON outputs: tools/cmp_variants.py gpurun_out/bk_*.json"""
import json, sys, os
files = sys.argv[1:]
data = {}
names = []
for f in files:
    n = os.path.basename(f).replace("bk_", "").replace(".json", "")
    names.append(n)
    for r in json.load(open(f))["rows"]:
        data.setdefault((r["kernel"], r["p"]), {})[n] = r
print(f"{'kern':>5}{'p':>3}" + "".join(f"{n[:11]:>12}" for n in names))
for key in sorted(data):
    print(f"{key[0]:>5}{key[1]:>3}" + "".join(f"{100*data[key][n]['frac']:>12.1f}" if n in data[key] else f"{'-':>12}" for n in names))
