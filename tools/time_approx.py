#!/usr/bin/env python
"""Device time of the aligned Sph-IoU / FoV-IoU siblings at several sizes, L2 flushed between calls (HBM GB/s = 36 B/pair)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sph_retina_b200 import synthetic as S
from sph_retina_b200.sphdet.iou import fov_iou, sph_iou
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def run(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ms = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return statistics.median(ms)
for n in [int(v) for v in os.environ.get("SIZES", "1000000,4000000,16000000").split(",")]:
    b1 = S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box="bfov", seed=0).cuda()
    b2 = S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box="bfov", seed=1).cuda()
    for name, fn in (("sph_iou", sph_iou), ("fov_iou", fov_iou)):
        ms = run(lambda: fn(b1, b2, is_aligned=True))
        print("%s %9d pairs: %8.1f us  %6.1f Gpairs/s  %6.0f GB/s" % (name, n, ms * 1e3, n / ms / 1e6, n * 36 / ms / 1e6))
