#!/bin/bash
# Runs ON THE GPU BOX (gpurun --gpus N): the per-N bench lines of the default (row-sharded sweep) workload into
# gpurun_out/, launched exactly as the driver launches them.
#   tools/scale_round.sh <tag> <exchange: auto|peer|nccl> <n1> [n2 ...]        e.g.  r02_v4 auto 1 2
set -u
tag=$1; ex=$2; shift; shift
out=gpurun_out; mkdir -p $out
port=29600
for n in "$@"; do
  port=$((port + 1))
  extra="--no-extras --no-cpu --exchange $ex"
  f=$out/scale_sweep_${ex}_n${n}_$tag
  if [ $n = 1 ]; then timeout 300 python bench.py --gpus 1 $extra > $f.json 2> $f.err
  else NCCL_DEBUG=${NCCL_DEBUG:-WARN} timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n $extra \
         > $f.json 2> $f.err; fi
  [ -s $f.json ] || tail -30 $f.err
  grep -i "warn\|symmetric" $f.err | head -5
done
python - "$tag" "$ex" "$@" <<'PY'
import json, sys
tag, ex, ns = sys.argv[1], sys.argv[2], sys.argv[3:]
base = None
for n in ns:
    try:
        d = json.loads(open("gpurun_out/scale_sweep_%s_n%s_%s.json" % (ex, n, tag)).read().strip().splitlines()[-1])
        if base is None:
            base = (int(n), d["value"], d["e2e"]["value"])
        eff = d["value"] / (base[1] * int(n) / base[0])
        eff_e = d["e2e"]["value"] / (base[2] * int(n) / base[0])
        print("sweep[%s] N=%s  %.1f Gpairs/s  %.4f ms/step  eff %.3f | e2e %.1f Gpairs/s  %.4f ms  eff %.3f | kernel_ms %.4f  clocks %s %s | %s" % (
            ex, n, d["value"] / 1e9, d["ms_per_step"], eff, d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"], eff_e,
            d["roofline"]["kernel_ms"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["config"]["exchange"][:12]))
    except Exception as e:
        print("sweep N=%s FAILED %r" % (n, e))
PY
