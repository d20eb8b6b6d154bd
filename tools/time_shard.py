#!/usr/bin/env python
"""Kernel time of one rank's share of the sweep at 1/N of the anchors (k_box_pre + k_iou_pairwise2 into a key block),
L2 flushed, CUDA events:   python tools/time_shard.py [N ...]      (A/B: SPHK_PROBE_LIB=..._tuning.so SPHK_NO_TAIL=1)"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sph_retina_b200 import _native, synthetic as S  # noqa: E402

dev = "cuda:0"
A = S.generate_boxes(1 << 20, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=0).to(dev)
G = S.generate_boxes(1024, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=1).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n in [int(a) for a in sys.argv[1:]] or [1, 8]:
    rows = (1 << 20) // n
    a = A[:rows].contiguous()
    rk = torch.zeros(rows, dtype=torch.int64, device=dev)
    ck = torch.zeros(1024, dtype=torch.int64, device=dev)
    ms = []
    for it in range(15):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _native.iou_pairwise_keys("sph2pob_efficient", a, G, row_keys_out=rk, col_keys_out=ck)
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ms.append(e0.elapsed_time(e1))
    print("1/%d of the sweep (%d x 1024): %.4f ms median, min %.4f  -> %.1f Gpairs/s; x%d = %.4f ms" % (
        n, rows, statistics.median(ms), min(ms), rows * 1024 / statistics.median(ms) / 1e6, n, statistics.median(ms) * n))
