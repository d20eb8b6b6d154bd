#!/bin/bash
# Runs ON THE GPU BOX (gpurun --gpus N): the per-N bench lines of both sharded workloads into gpurun_out/.
#   tools/scale_round.sh <tag> <n1> [n2 ...]        e.g.  r01_v13 1 2 4 8
set -u
tag=$1; shift
out=gpurun_out; mkdir -p $out
port=29600
for wl in assign sweep; do
  for n in "$@"; do
    port=$((port + 1))
    extra="--no-extras --no-cpu"; [ $wl = sweep ] && extra="$extra --workload sweep --steps 10 --warmup 3"
    if [ $n = 1 ]; then python bench.py --gpus 1 $extra > $out/scale_${wl}_n${n}_$tag.json 2> $out/scale_${wl}_n${n}_$tag.err
    else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n $extra \
           > $out/scale_${wl}_n${n}_$tag.json 2> $out/scale_${wl}_n${n}_$tag.err; fi
    python - <<PY
import json
try:
    d = json.loads(open("$out/scale_${wl}_n${n}_$tag.json").read().strip().splitlines()[-1])
    print("$wl", d["n_gpus"], "%.1f Gpairs/s" % (d["value"] / 1e9), "%.3f ms" % d["ms_per_step"], d["clocks"]["reasons"])
except Exception as e:
    print("$wl", $n, "FAILED", e)
PY
  done
done
