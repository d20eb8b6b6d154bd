#!/usr/bin/env python
"""gpurun_out/scale_sweep_<route>_n<N>_<tag>.json (tools/scale_round.sh) -> profiles/scaling_<round>.json
    python tools/scale_collect.py r02_v23 r02 "note ..." """
import glob
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
out, base = {}, None
files = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "scale_sweep_*_n*_%s.json" % tag)),
               key=lambda f: (0 if "_auto_" in f else 1, int(re.search(r"_n(\d+)_", f).group(1))))
for f in files:
    route, n = re.search(r"scale_sweep_(.+)_n(\d+)_", os.path.basename(f)).groups()
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except (IndexError, ValueError):
        continue
    if route == "auto" and int(n) == 1:
        base = d["value"]
    key = "%s_n%s" % ("peer" if route == "auto" else route, n)
    out[key] = {"n_gpus": d["n_gpus"], "value": d["value"], "ms_per_step": d["ms_per_step"], "scaling": d["scaling"], "steps": d["steps"],
                "warmup": d["warmup"], "clocks": d["clocks"], "e2e_value": d["e2e"]["value"], "e2e_ms_per_step": d["e2e"]["ms_per_step"],
                "kernel_ms": d["roofline"]["kernel_ms"], "kernel_ms_per_rank": d["roofline"].get("kernel_ms_per_rank"),
                "exchange": d["config"]["exchange"], "api": d["config"]["api"]}
for k, v in out.items():
    if base:
        v["efficiency_vs_n1"] = v["value"] / (base * v["n_gpus"])
out["note"] = note
json.dump(out, open(os.path.join(ROOT, "profiles", "scaling_%s.json" % rnd), "w"), indent=1)
for k, v in out.items():
    if isinstance(v, dict):
        print("%-16s %8.1f Gpairs/s  %.4f ms  eff %.3f" % (k, v["value"] / 1e9, v["ms_per_step"], v.get("efficiency_vs_n1", 0)))
