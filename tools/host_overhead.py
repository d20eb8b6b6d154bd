#!/usr/bin/env python
"""Host-side cost of one SphOverlaps2D call (the per-image call MaxIoUAssigner makes), piece by piece."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sph_retina_b200 import _native as N  # noqa: E402
from sph_retina_b200 import synthetic as S  # noqa: E402
from sph_retina_b200.sphdet.iou import SphOverlaps2D, sph2pob_efficient_iou  # noqa: E402

gts, anchors = S.assignment_batch()
gts, anchors = gts.cuda(), anchors.cuda()
calc = SphOverlaps2D("sph2pob_efficient_iou", 5)
g0 = gts[0]


def bench(name, fn, n=2000):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    dt = (time.perf_counter() - t) / n * 1e6
    torch.cuda.synchronize()
    print("%-52s %7.2f us" % (name, dt))


small = anchors[:256].contiguous()
bench("ctypes: sphk_abi_version()", lambda: N.lib.sphk_abi_version())
bench("ctypes: sphk_iou_pairwise with R = 0 (18 args, no launch)",
      lambda: N.lib.sphk_iou_pairwise(0, g0.data_ptr(), 0, small.data_ptr(), 256, 5, 0, 0, 0, None, 256, None, None, None, None, 0, 0, None, 0))
bench("torch.empty((32, 98208))", lambda: torch.empty((32, 98208), dtype=torch.float32, device="cuda"))
bench("gts[i]", lambda: gts[3])
bench("slice [..., :5] x2", lambda: (g0[..., :5], anchors[..., :5]))
bench("_native._boxes x2", lambda: (N._boxes(g0, "a"), N._boxes(anchors, "b")))
bench("_native._workspace", lambda: N._workspace(g0.device, 104 * (32 + 98208) + 32))
bench("_native._on_device enter/exit", lambda: N._on_device(g0.device).__enter__())
bench("_native.iou_pairwise(32 x 256) (2 tiny launches)", lambda: N.iou_pairwise("sph2pob_efficient", g0, small))
bench("sph2pob_efficient_iou(32 x 256)", lambda: sph2pob_efficient_iou(g0, small))
bench("SphOverlaps2D(32 x 256)", lambda: calc(g0, small))
bench("SphOverlaps2D(32 x 98208) issue rate", lambda: calc(g0, anchors), n=400)
