#!/usr/bin/env python
"""Launches one hot-path kernel a few times on BASELINE-shaped synthetic input (for ncu captures).
    python tools/run_kernel.py aligned|aligned5|loss|gdloss|nms|nms_agnostic|nms_pipeline|sweep|assign|assigner|headloss|format|approx [--iters 5]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sph_retina_b200 import synthetic as S  # noqa: E402
from sph_retina_b200.sphdet.bbox.nms import sph_batched_nms_images  # noqa: E402
from sph_retina_b200.sphdet.iou import SphOverlaps2D, sph2pob_efficient_iou, sph_max_overlaps  # noqa: E402
from sph_retina_b200.sphdet.losses import Sph2PobIoULoss  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("which")
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
dev = "cuda:0"
if a.which in ("aligned", "aligned5"):
    box = "bfov" if a.which == "aligned" else "rbfov"
    b1 = S.generate_boxes(1_000_000, alpha_range=(1, 100), beta_range=(1, 100), box=box, seed=0).to(dev)
    b2 = S.generate_boxes(1_000_000, alpha_range=(1, 100), beta_range=(1, 100), box=box, seed=1).to(dev)
    fn = lambda: sph2pob_efficient_iou(b1, b2, is_aligned=True)
elif a.which == "loss":
    p, t = (x.to(dev) for x in S.loss_pairs(200_000))
    L = Sph2PobIoULoss(reduction="sum")

    def fn():
        q = p.detach().requires_grad_(True)
        L(q, t).backward()
elif a.which in ("nms", "nms_agnostic"):
    boxes, scores, labels, image_ids = (x.to(dev) for x in S.nms_batch(64, 1000, 80))
    if a.which == "nms_agnostic":
        labels = torch.zeros_like(labels)
    fn = lambda: sph_batched_nms_images(boxes, scores, labels, image_ids, 0.5)
elif a.which == "nms_pipeline":
    from sph_retina_b200.sphdet.bbox.nms import sph_nms_image_blocks
    boxes, scores, labels, image_ids = (x.to(dev) for x in S.nms_batch(64, 1000, 80))
    fn = lambda: sph_nms_image_blocks(boxes, scores, labels, 64, 80, 0.5, 100)
elif a.which == "sweep":
    A = S.generate_boxes(1 << 20, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=0).to(dev)
    G = S.generate_boxes(1024, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=1).to(dev)
    fn = lambda: sph_max_overlaps(A, G)
elif a.which == "headloss":
    from sph_retina_b200.sphdet.bbox.coder import DeltaXYWHASphBBoxCoder
    from sph_retina_b200.sphdet.losses import Sph2PobDecodedIoULoss
    anchors, deltas, target, weight = (t.to(dev) for t in S.head_loss_batch(16))
    coder = DeltaXYWHASphBBoxCoder(target_stds=(0.1, 0.1, 0.2, 0.2, 0.1))
    LD = Sph2PobDecodedIoULoss()
    npos = float((weight[:, 0] > 0).sum())

    def fn():
        d = deltas.detach().requires_grad_(True)
        LD.forward_decoded(coder, anchors, d, target, weight, avg_factor=npos).backward()
elif a.which == "gdloss":
    from sph_retina_b200.sphdet.losses import Sph2PobGDLoss
    pred, target = (t.to(dev) for t in S.loss_pairs(200_000))
    LG = Sph2PobGDLoss("kld", reduction="sum")

    def fn():
        p = pred.detach().requires_grad_(True)
        LG(p, target).backward()
elif a.which == "approx":
    from sph_retina_b200.sphdet.iou import sph_iou
    c1 = S.generate_boxes(16_000_000, alpha_range=(1, 100), beta_range=(1, 100), box="bfov", seed=0).to(dev)
    c2 = S.generate_boxes(16_000_000, alpha_range=(1, 100), beta_range=(1, 100), box="bfov", seed=1).to(dev)
    fn = lambda: sph_iou(c1, c2, is_aligned=True)
elif a.which == "format":
    from sph_retina_b200.sphdet.bbox.box_formator import Sph2PlanarBoxTransform
    sph = S.generate_boxes(1 << 24, alpha_range=(1, 100), beta_range=(1, 100), box="bfov", seed=0).to(dev)
    T = Sph2PlanarBoxTransform('sph2pix', 4)
    fn = lambda: T(sph, (512, 1024))
elif a.which == "assigner":
    from sph_retina_b200.sphdet.assigners import SphMaxIoUAssigner
    gts, anchors = S.assignment_batch()
    gts, anchors = gts.to(dev), anchors.to(dev)
    asg = SphMaxIoUAssigner(0.5, 0.3, min_pos_iou=0.0, iou_calculator=SphOverlaps2D('sph2pob_efficient_iou', 5))
    gl = [gts[i] for i in range(16)]
    fn = lambda: asg.assign_batch(anchors, gl)
else:
    gts, anchors = S.assignment_batch()
    gts, anchors = gts.to(dev), anchors.to(dev)
    calc = SphOverlaps2D('sph2pob_efficient_iou', 5)
    fn = lambda: [calc(gts[i], anchors) for i in range(16)]
for _ in range(a.iters):
    fn()
torch.cuda.synchronize()
print("ok", a.which)
