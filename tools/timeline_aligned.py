#!/usr/bin/env python
"""Per-warp timeline of k_iou_aligned2 on BASELINE configs[0] (1 M aligned BFoV pairs) from the instrumented build:
    python -m sph_retina_b200.build --timeline ; SPHK_PROBE_LIB=sph_retina_b200/_lib/libsphk_tl.so python tools/timeline_aligned.py
Prints when warps start, finish scanning and end (relative to the first start), and what the last ones were doing."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sph_retina_b200 import _native  # noqa: E402
from sph_retina_b200 import synthetic as S  # noqa: E402
from sph_retina_b200.sphdet.iou import sph2pob_efficient_iou  # noqa: E402

dev = torch.device("cuda:0")
n = int(os.environ.get("PAIRS", 1_000_000))
box = os.environ.get("BOX", "bfov")
b1 = S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box=box, seed=0).to(dev)
b2 = S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box=box, seed=1).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5):
    flush.zero_()
    sph2pob_efficient_iou(b1, b2, is_aligned=True)
torch.cuda.synchronize()
flush.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
sph2pob_efficient_iou(b1, b2, is_aligned=True)
e1.record()
torch.cuda.synchronize()
print("event-timed call: %.2f us" % (e0.elapsed_time(e1) * 1e3))
sm, _, _ = _native.device_info()
nw = min(4 * sm, (n // 32 + 31) // 32) * 8
t = np.zeros(3 * nw, np.uint64)
slow = np.zeros(nw, np.uint32)
lib = _native.lib
lib.sphk_debug_timeline_warps.restype = ctypes.c_int
rc = lib.sphk_debug_timeline_warps(t.ctypes.data_as(ctypes.c_void_p), slow.ctypes.data_as(ctypes.c_void_p), nw)
assert rc == 0, rc
t = t.reshape(nw, 3).astype(np.int64)
t0 = t[:, 0].min()
st, sc, en = ((t[:, k] - t0) / 1e3 for k in range(3))
print("warps %d | start: median %.2f p99 %.2f max %.2f us | scan end: median %.2f max %.2f | end: median %.2f p90 %.2f p99 %.2f max %.2f us" % (
    nw, np.median(st), np.percentile(st, 99), st.max(), np.median(sc), sc.max(), np.median(en), np.percentile(en, 90), np.percentile(en, 99), en.max()))
print("warp life us: median %.2f  p90 %.2f  max %.2f | after the last scan round: median %.2f p90 %.2f max %.2f" % (
    np.median(en - st), np.percentile(en - st, 90), (en - st).max(), np.median(en - sc), np.percentile(en - sc, 90), (en - sc).max()))
print("warps with reference-order pairs: %d (pairs %d); their end: %s" % ((slow > 0).sum(), slow.sum(), np.round(np.sort(en[slow > 0])[-8:], 2)))
order = np.argsort(-en)[:10]
for w in order:
    print("  late warp %5d (CTA %4d)  start %.2f  scan end %.2f  end %.2f  slow %d" % (w, w // 8, st[w], sc[w], en[w], slow[w]))
grid = np.linspace(0, en.max(), 30)
print("running warps over time:", [int(((st <= x) & (en > x)).sum()) for x in grid])
