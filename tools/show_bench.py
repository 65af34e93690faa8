#!/usr/bin/env python
"""One-line summary of bench.py JSON outputs: python tools/show_bench.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    txt = [l for l in open(f) if l.startswith("{")]
    d = json.loads(txt[-1])
    st = {k: round(v["ms_per_step"], 3) for k, v in d.get("stages", {}).items()}
    print(f, "n_gpus", d["n_gpus"], "pairs/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None,
          "full", round(d["full_registration"]["value"], 1) if d.get("full_registration") else None, "gt", d.get("gt_check"),
          "frac", round(d["roofline"]["frac"], 3) if d.get("roofline", {}).get("frac") else None, st)
