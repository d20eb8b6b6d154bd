#!/bin/bash
# Runs ON THE GPU BOX: A/B lines of bench.py through instrumented twins of the library.
#   tools/ab.sh <tag> <workload> name=lib[,ENV=VAL...] ...     lib: file under sph_retina_b200/_lib ("-" = the product library)
set -u
tag=$1; wl=$2; shift; shift
out=gpurun_out; mkdir -p $out
for spec in "$@"; do
  name=${spec%%=*}; rest=${spec#*=}
  lib=${rest%%,*}; envs=""
  [ "$rest" != "$lib" ] && envs=$(echo "${rest#*,}" | tr ',' ' ')
  if [ "$lib" = "-" ]; then libenv=""; else libenv="SPHK_PROBE_LIB=$PWD/sph_retina_b200/_lib/$lib"; fi
  env $libenv $envs timeout 300 python bench.py --workload $wl --no-extras --no-cpu --no-e2e > $out/ab_${wl}_${name}_$tag.json 2> $out/ab_${wl}_${name}_$tag.err || { echo "ab $name failed"; tail -3 $out/ab_${wl}_${name}_$tag.err; }
done
python - "$tag" "$wl" <<'PY'
import json, glob, sys
tag, wl = sys.argv[1:3]
for f in sorted(glob.glob("gpurun_out/ab_%s_*_%s.json" % (wl, tag))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d["roofline"]
        print("%-44s %8.2f Gpairs/s  %.4f ms/step  kernel_ms %.4f frac %.3f early %.4f dense %.1f G" % (
            f.split("/")[-1], d["value"] / 1e9, d["ms_per_step"], r["kernel_ms"], r["frac"], r["early_out_rate"], r["dense_pairs_per_s"] / 1e9))
    except Exception as e:
        print(f, "unreadable", e)
PY
