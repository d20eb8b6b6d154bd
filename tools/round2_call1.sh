#!/bin/bash
# Runs ON THE GPU BOX (gpurun, 1 GPU): GPU parity suite, the default bench, the reference arm, and A/B lines of the
# round-2 kernel changes through the instrumented twin library (SPHK_PROBE_LIB=_lib/libsphk_tuning.so).
#   tools/round2_call1.sh <tag>
set -u
tag=${1:-r02_v1}
out=gpurun_out; mkdir -p $out
python __graft_entry__.py > $out/build_$tag.log 2>&1 || { echo "build failed"; tail -20 $out/build_$tag.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu_$tag.log
tail -5 $out/pytest_gpu_$tag.log
timeout 600 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err || { echo "bench failed"; tail -20 $out/bench_$tag.err; }
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err || echo "reference arm failed"
timeout 300 python bench.py --workload assign --no-extras --no-cpu > $out/bench_assign_$tag.json 2> $out/bench_assign_$tag.err || echo "assign bench failed"
tun=$PWD/sph_retina_b200/_lib/libsphk_tuning.so
ab() {   # name, workload, env...
  name=$1; wl=$2; shift; shift
  env SPHK_PROBE_LIB=$tun "$@" timeout 300 python bench.py --workload $wl --no-extras --no-cpu --no-e2e > $out/ab_${name}_$tag.json 2> $out/ab_${name}_$tag.err || echo "ab $name failed"
}
ab assign_base assign SPHK_X=0
ab assign_noboxcull assign SPHK_NO_BOXCULL=1
ab assign_nopdl assign SPHK_NO_PDL=1
ab sweep_base sweep SPHK_X=0
ab sweep_noboxcull sweep SPHK_NO_BOXCULL=1
ab sweep_tr16 sweep SPHK_TR=16
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/*_$tag.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print("%-40s %8.2f Gpairs/s  %.4f ms/step  e2e %s  kernel_ms %s frac %s early %s" % (
            f.split("/")[-1], d["value"] / 1e9, d["ms_per_step"], (d.get("e2e") or {}).get("value"),
            r.get("kernel_ms"), r.get("frac"), r.get("early_out_rate")))
    except Exception as e:
        print(f, "unreadable", e)
PY
