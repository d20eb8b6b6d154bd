"""Times the secondary IoU backends (naive / unbiased / legacy) on configs[0]'s 1 M aligned pairs and unbiased NMS.
Run on the GPU box:  python tools/time_backends.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sph_retina_b200 import synthetic as S  # noqa: E402
from sph_retina_b200.sphdet.iou import naive_iou, sph2pob_efficient_iou, sph2pob_legacy_iou, unbiased_iou  # noqa: E402
from sph_retina_b200.sphdet.bbox.nms import sph_nms_image_blocks  # noqa: E402


def quick(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(iters):
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


dev = torch.device("cuda:0")
n = 1_000_000
out = {}
for box in ("bfov", "rbfov"):
    b1 = S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box=box, seed=0).to(dev)
    b2 = S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box=box, seed=1).to(dev)
    fns = [("efficient", sph2pob_efficient_iou), ("naive", naive_iou), ("unbiased", unbiased_iou)]
    if box == "bfov":
        fns.append(("legacy", sph2pob_legacy_iou))
    for name, fn in fns:
        ms = quick(lambda: fn(b1, b2, is_aligned=True))
        out["aligned_1M_%s_%s" % (name, box)] = {"ms": ms, "Gpairs_per_s": n / ms / 1e6}
    near = (b1 + torch.randn_like(b1) * 5).clamp(min=1)
    near[:, 0].clamp_(0, 360); near[:, 1].clamp_(0, 180); near[:, 2:4].clamp_(max=179)
    ms = quick(lambda: unbiased_iou(b1, near, is_aligned=True))
    out["aligned_1M_unbiased_%s_overlapping_pairs" % box] = {"ms": ms, "Gpairs_per_s": n / ms / 1e6}
boxes, scores, labels, image_ids = (t.to(dev) for t in S.nms_batch(64, 1000, 80))
for calc in ("sph2pob_efficient", "naive_iou", "unbiased_iou"):
    ms = quick(lambda: sph_nms_image_blocks(boxes, scores, labels, 64, 80, 0.5, 100, iou_calculator=calc))
    out["nms_64x1000x80_" + calc] = {"ms": ms}
print(json.dumps(out, indent=1))
