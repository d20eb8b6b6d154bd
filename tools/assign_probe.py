#!/usr/bin/env python
"""Where does the time of ONE SphOverlaps2D call of the assign workload go?  Times back-to-back calls (CUDA events, no L2
flush) for several anchor counts (fixed cost vs slope) and prints one line per setting.  Knobs come from the environment
(SPHK_TR, SPHK_PROBE) because the library reads them at load time:
    python tools/assign_probe.py [--reps 200]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sph_retina_b200 import synthetic as S  # noqa: E402
from sph_retina_b200.sphdet.iou import SphOverlaps2D, sph_max_overlaps  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=200)
a = ap.parse_args()
dev = torch.device("cuda:0")
anchors = S.retina_anchors().to(dev)
gts = [S.generate_boxes(32, alpha_range=(5, 120), beta_range=(5, 120), box="rbfov", seed=100 + i).to(dev) for i in range(16)]
calc = SphOverlaps2D('sph2pob_efficient_iou', 5)


def timed(fn, reps):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


tag = "TR=%s PROBE=%s" % (os.environ.get("SPHK_TR", "-"), os.environ.get("SPHK_PROBE", "-"))
for mult in (1, 2, 4):
    anc = anchors.repeat(mult, 1)
    outs = [None]

    def call(i=0):
        outs[0] = calc(gts[i % 16], anc)
    us = timed(call, a.reps)
    print("%s  matrix  C=%7d  %.2f us/call  %.1f Gpairs/s" % (tag, anc.size(0), us, 32 * anc.size(0) / us / 1e3))

    def call2(i=0):
        outs[0] = sph_max_overlaps(gts[i % 16], anc)
    us = timed(call2, a.reps)
    print("%s  fusedmax C=%7d  %.2f us/call  %.1f Gpairs/s" % (tag, anc.size(0), us, 32 * anc.size(0) / us / 1e3))
