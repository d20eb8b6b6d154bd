#!/usr/bin/env python
"""Device time of the per-image assignment calls (CUDA events), for tuning."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sph_retina_b200 import synthetic as S
from sph_retina_b200.sphdet.iou import SphOverlaps2D, sph_max_overlaps
gts, anchors = S.assignment_batch(); gts, anchors = gts.cuda(), anchors.cuda()
calc = SphOverlaps2D('sph2pob_efficient_iou', 5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def run(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ms = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return statistics.median(ms)
ms = run(lambda: [calc(gts[i], anchors) for i in range(16)])
print("SPHK_TR=%s  16 calls: %.3f ms  -> %.1f Gpairs/s" % (os.environ.get("SPHK_TR", "auto"), ms, 16 * 32 * 98208 / ms / 1e6))
ms1 = run(lambda: calc(gts[0], anchors))
print("   single call: %.1f us" % (ms1 * 1e3))
ms2 = run(lambda: [sph_max_overlaps(gts[i], anchors) for i in range(16)])
print("   fused max x16: %.3f ms" % ms2)
