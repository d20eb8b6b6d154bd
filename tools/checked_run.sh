#!/bin/bash
# Runs ON THE GPU BOX: the bounds-checked twin of the library (-DSPHK_CHECKED: the invariants of the shared-memory rings
# and the index range of every compacted store trap) under every kernel on small awkward shapes, the whole GPU parity
# suite and one full-size sweep.  compute-sanitizer is closed on this pool; this is the check "of our own".
#   tools/checked_run.sh <tag>
set -u
tag=${1:-r02}
out=gpurun_out/checked_$tag.txt
mkdir -p gpurun_out
L=sph_retina_b200/_lib/libsphk_checked.so
[ -f $L ] || python -m sph_retina_b200.build --checked
{
echo "== library: $L (-DSPHK_CHECKED), $(date -u +%FT%TZ)"
echo "== tools/sanitize_small.py"; SPHK_PROBE_LIB=$L python tools/sanitize_small.py 2>&1 | tail -3
echo "== pytest -m gpu"; SPHK_PROBE_LIB=$L python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== tools/time_sweep.py (1,048,576 x 1,024, both orientations)"; SPHK_PROBE_LIB=$L python tools/time_sweep.py 2>&1 | tail -3
echo "== tools/run_kernel.py aligned / aligned5 / assign / headloss"
for k in aligned aligned5 assign headloss; do SPHK_PROBE_LIB=$L python tools/run_kernel.py $k 2>&1 | tail -1; echo "$k rc=$?"; done
} > $out 2>&1
grep -c "SPHK_CHECK failed" $out | sed 's/^/SPHK_CHECK failures: /' | tee -a $out
tail -25 $out
