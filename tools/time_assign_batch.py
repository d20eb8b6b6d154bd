#!/usr/bin/env python
"""Device time of the batched assignment paths (k_iou_pairwise2 with the box-frame test): all 16 images' GT in one
SphOverlaps2D call, the fused per-image max / argmax, and SphMaxIoUAssigner.assign_batch."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sph_retina_b200 import synthetic as S
from sph_retina_b200.sphdet.iou import SphOverlaps2D, sph_max_overlaps
from sph_retina_b200.sphdet.assigners import SphMaxIoUAssigner
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
pairs = 16 * 32 * 98208
ms = run(lambda: calc(gts.view(-1, 5), anchors))
print("one call [512 x 98208]: %.4f ms  %.1f Gpairs/s" % (ms, pairs / ms / 1e6))
ms = run(lambda: sph_max_overlaps(gts.view(-1, 5), anchors))
print("fused max [512 x 98208]: %.4f ms  %.1f Gpairs/s" % (ms, pairs / ms / 1e6))
asg = SphMaxIoUAssigner(0.5, 0.4, min_pos_iou=0, iou_calculator=dict(type='SphOverlaps2D', backend='sph2pob_efficient_iou', box_version=5))
gl = [gts[i] for i in range(16)]
ms = run(lambda: asg.assign_batch(anchors, gl, None), n=10)
print("assign_batch 16 images: %.4f ms  %.1f Gpairs/s" % (ms, pairs / ms / 1e6))
