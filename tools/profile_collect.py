#!/usr/bin/env python
"""Turns the scratch evidence of tools/profile_round.sh (gpurun_out/*_<tag>.*) into the tracked summaries under
profiles/: bench lines, the ncu launch list, per-kernel counter summaries + detail CSVs, stall and hot-spot tables and
traffic.json (dram bytes per launch of the dominant kernels, read by bench.py's roofline object).
    python tools/profile_collect.py r01_v10"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for a, b in (("bench_%s.json" % tag, "bench_%s.json" % tag), ("bench_ref_%s.json" % tag, "bench_%s_reference_arm.json" % tag),
             ("launches_%s.csv" % tag, "ncu_launches_%s.csv" % tag)):
    if os.path.isfile(os.path.join(src, a)):
        shutil.copy(os.path.join(src, a), os.path.join(dst, b))
tpath = os.path.join(dst, "traffic.json")
traffic = {k: v for k, v in (json.load(open(tpath)) if os.path.isfile(tpath) else {}).items() if k != "source"}
kernels = ["aligned", "aligned5", "assign", "sweep", "loss", "nms", "nms_pipeline", "assigner", "headloss", "gdloss", "format", "approx"]
for k in kernels:
    rep = os.path.join(src, "prof_%s_%s.ncu-rep" % (k, tag))
    if not os.path.isfile(rep):
        continue
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    tmp = os.path.join(src, "_raw_%s.csv" % k)
    open(tmp, "w").write(raw)
    stalls = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_stalls.py"), tmp], capture_output=True, text=True).stdout
    open(os.path.join(dst, "ncu_%s_%s_summary.txt" % (k, tag)), "w").write(summ + "\nwarp stalls (cycles per issued instruction)\n" + stalls)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(dst, "ncu_%s_%s_details.csv" % (k, tag)), "w").write(det)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    r = rows[2]
    rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    units = rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic[k] = float(r[rd].replace(",", "")) * scale.get(units[rd], 1.0) + float(r[wr].replace(",", "")) * scale.get(units[wr], 1.0)
    kern = {"aligned": "k_iou_aligned2", "aligned5": "k_iou_aligned2", "assign": "k_iou_rows32", "sweep": "k_iou_pairwise2",
            "loss": "k_loss", "nms": "k_nms", "nms_pipeline": "k_nms", "assigner": "k_iou_pairwise2", "headloss": "k_decode_loss", "gdloss": "k_obb_loss", "format": "k_box_format4", "approx": "k_approx_aligned4"}[k]
    if k in ("aligned", "assign", "sweep", "headloss"):
        hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hotspots.py"), rep,
                              os.path.join(ROOT, "sph_retina_b200", "_lib", "libsphk.so"), kern, "--top", "45"], capture_output=True, text=True).stdout
        open(os.path.join(dst, "ncu_%s_%s_hotspots.txt" % (k, tag)), "w").write(hot)
# roofline_ncu.json: what bench.py's roofline object quotes from the committed capture of the dominant kernel
rpath = os.path.join(dst, "roofline_ncu.json")
roof = json.load(open(rpath)) if os.path.isfile(rpath) else {}
for k in ("sweep", "assign", "aligned"):
    spath = os.path.join(dst, "ncu_%s_%s_summary.txt" % (k, tag))
    if not os.path.isfile(spath) or k not in traffic:
        continue
    vals = {}
    for line in open(spath):
        parts = line.split()
        if len(parts) >= 2 and ("__" in parts[0]):
            try:
                vals[parts[0]] = float(parts[1].replace(",", ""))
                if parts[0] == "gpu__time_duration.sum" and len(parts) > 2:
                    vals[parts[0]] *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(parts[2], 1.0)      # -> microseconds
            except ValueError:
                pass
    el, act = vals.get("sm__cycles_elapsed.max"), vals.get("sm__cycles_active.avg")
    roof[k] = {
        "dram_bytes_per_launch": traffic[k],
        "counters": {
            "issue_active_pct_of_active_cycles": vals.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "sm_active_frac_of_elapsed": (act / el) if el and act else None,
            "pipe_fma_pct": vals.get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "pipe_alu_pct": vals.get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
            "pipe_xu_pct": vals.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
            "pipe_lsu_pct": vals.get("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": vals.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "warp_instructions": vals.get("smsp__inst_executed.sum"),
            "threads_per_instruction": vals.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "registers_per_thread": vals.get("launch__registers_per_thread"),
            "kernel_us_under_ncu": vals.get("gpu__time_duration.sum"),
            "shared_bank_conflicts": vals.get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        },
        "source": "profiles/ncu_%s_%s_summary.txt (ncu --set full --clock-control none, one launch)" % (k, tag),
    }
json.dump(roof, open(rpath, "w"), indent=1)
traffic["source"] = ("ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, one launch "
                     "(profiles/ncu_*_%s_details.csv)" % tag)
json.dump(traffic, open(os.path.join(dst, "traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
