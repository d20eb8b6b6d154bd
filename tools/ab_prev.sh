#!/bin/bash
# Runs ON THE GPU BOX: A/B of the product library against sph_retina_b200/_lib/libsphk_prev.so (a copy of the library of the
# previous commit, same ABI) on the same GPU: sweep both orientations, per-image assign, aligned 1M.   tools/ab_prev.sh
P=sph_retina_b200/_lib/libsphk_prev.so
for rep in 1 2; do
echo "--- new"; python tools/time_sweep.py; python tools/time_assign.py 2>&1 | head -2; SIZES=1000000,16000000 python tools/time_aligned.py 2>&1 | tail -4
echo "--- prev"; SPHK_PROBE_LIB=$P python tools/time_sweep.py; SPHK_PROBE_LIB=$P python tools/time_assign.py 2>&1 | head -2; SPHK_PROBE_LIB=$P SIZES=1000000,16000000 python tools/time_aligned.py 2>&1 | tail -4
done
