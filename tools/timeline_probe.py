#!/usr/bin/env python
"""CTA timeline of k_iou_pairwise2 on the assign workload (one image), from an instrumented build of the library:
    nvcc ... -DSPHK_TIMELINE -o sph_retina_b200/_lib/libsphk_tl.so ; SPHK_PROBE_LIB=that python tools/timeline_probe.py
Prints when CTAs start / end relative to the first start, per-SM busy time and the tail."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sph_retina_b200 import _native  # noqa: E402
from sph_retina_b200 import synthetic as S  # noqa: E402
from sph_retina_b200.sphdet.iou import SphOverlaps2D  # noqa: E402

dev = torch.device("cuda:0")
anchors = S.retina_anchors().to(dev)
gt = S.generate_boxes(32, alpha_range=(5, 120), beta_range=(5, 120), box="rbfov", seed=100).to(dev)
calc = SphOverlaps2D('sph2pob_efficient_iou', 5)
for _ in range(20):
    m = calc(gt, anchors)
torch.cuda.synchronize()
n = int(os.environ.get("NCTA", 1536))
t = np.zeros(2 * n, np.uint64)
sm = np.zeros(n, np.uint32)
rc = _native.lib.sphk_debug_timeline(t.ctypes.data_as(ctypes.c_void_p), sm.ctypes.data_as(ctypes.c_void_p), n)
assert rc == 0, rc
t = t.reshape(n, 2).astype(np.int64)
t0 = t[:, 0].min()
st, en = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
print("CTAs %d  first start 0  last start %.2f us  first end %.2f  last end %.2f us" % (n, st.max(), en.min(), en.max()))
dur = en - st
print("CTA duration us: min %.2f  median %.2f  mean %.2f  p90 %.2f  max %.2f" % (dur.min(), np.median(dur), dur.mean(), np.percentile(dur, 90), dur.max()))
for q in (0.25, 0.5, 0.75, 0.9, 0.95, 0.99, 1.0):
    print("  %3.0f %% of the CTAs ended by %.2f us" % (q * 100, np.quantile(en, q)))
sms = np.unique(sm)
last = np.array([en[sm == k].max() for k in sms])
busy = np.array([dur[sm == k].sum() for k in sms])
cnt = np.array([(sm == k).sum() for k in sms])
print("SMs %d  CTAs/SM min %d max %d | SM last-end: min %.2f median %.2f max %.2f | sum of CTA durations per SM: min %.1f median %.1f max %.1f" % (
    len(sms), cnt.min(), cnt.max(), last.min(), np.median(last), last.max(), busy.min(), np.median(busy), busy.max()))
# which CTAs end last?
order = np.argsort(-en)[:12]
gx = 4
for b in order:
    print("  late CTA %4d (row tile %d, col tile idx %d)  start %.2f  dur %.2f  end %.2f  sm %d" % (b, b % gx, b // gx, st[b], dur[b], en[b], sm[b]))
# concurrency over time
grid = np.linspace(0, en.max(), 40)
conc = [(int(((st <= x) & (en > x)).sum())) for x in grid]
print("resident CTAs over time:", conc)
