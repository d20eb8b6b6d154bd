#!/bin/bash
# Runs ON THE GPU BOX (gpurun): the evidence set of one round into gpurun_out/ -- plain bench first, then the ncu launch
# list of the same command, then one --set full capture per hot kernel (each only after its command ran clean without ncu).
#   tools/profile_round.sh <tag> [bench|nobench] [kernel ...]     e.g.  r01_v10 bench aligned assign sweep
# (gpurun copies back at most 64 MiB: a full capture is ~8 MB, so split the kernels over two calls)
set -u
tag=${1:-r01}
mode=${2:-bench}
shift; shift
kernels=${*:-aligned aligned5 assign sweep loss gdloss nms nms_pipeline assigner headloss format approx}
out=gpurun_out
mkdir -p $out
if [ "$mode" = bench ]; then
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err || { echo "bench failed"; tail -5 $out/bench_$tag.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err || echo "reference arm failed"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu > $out/ncu_launches_$tag.log 2>&1 || echo "launch list failed"
fi
for k in $kernels; do
    python tools/run_kernel.py $k > $out/plain_$k.log 2>&1 || { echo "plain $k failed"; continue; }
    ncu --set full --import-source on --clock-control none \
        -k regex:"k_iou_aligned2|k_iou_pairwise2|k_iou_rows32|k_loss|k_nms|k_decode_loss|k_obb_loss|k_box_format|k_approx_aligned4" -s 3 -c 1 -o $out/prof_${k}_$tag -f \
        python tools/run_kernel.py $k > $out/ncu_$k.log 2>&1 || echo "ncu $k failed"
done
ls -la $out | grep $tag
