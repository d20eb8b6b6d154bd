#!/usr/bin/env python
"""Every kernel once on small, awkward shapes -- meant to run under compute-sanitizer
(memcheck / racecheck / synccheck, one tool per GPU call)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sph_retina_b200 import synthetic as S  # noqa: E402
from sph_retina_b200.sphdet.bbox.nms import SphNMS, sph_batched_nms_images  # noqa: E402
from sph_retina_b200.sphdet.iou import (fov_iou, naive_iou, sph2pob_efficient_iou, sph2pob_legacy_iou,  # noqa: E402
                                        sph2pob_standard_iou, sph_iou, sph_max_overlaps, unbiased_iou)
from sph_retina_b200.sphdet.bbox.nms import sph_nms_image_blocks  # noqa: E402
from sph_retina_b200.sphdet.losses import Sph2PobIoULoss  # noqa: E402

dev = "cuda:0"
for box, D in (("bfov", 4), ("rbfov", 5)):
    rows = S.generate_boxes(77, alpha_range=(5, 120), beta_range=(5, 120), box=box, seed=1).to(dev)
    cols = S.generate_boxes(1000, alpha_range=(5, 120), beta_range=(5, 120), box=box, seed=2).to(dev)
    cols[3] = rows[5]                       # identical pair -> reference-order path
    cols[::50, 0] = 0.0                     # clamped centres
    for fn in (sph2pob_efficient_iou, sph2pob_standard_iou):
        m = fn(rows, cols)
        m2 = fn(cols, rows)                 # many row tiles, few column tiles
        a = fn(cols[:999], cols.roll(1, 0)[:999], is_aligned=True)
        assert torch.isfinite(m).all() and torch.isfinite(m2).all() and torch.isfinite(a).all()
    r = sph_max_overlaps(rows, cols, return_matrix=True)
    r = sph_max_overlaps(cols, rows)
    for fn in (naive_iou, unbiased_iou) + ((sph2pob_legacy_iou,) if D == 4 else ()):      # the one-pair-per-thread kinds
        m = fn(rows, cols)
        a = fn(cols[:999], cols.roll(1, 0)[:999], is_aligned=True)
        assert m.shape == (77, 1000) and a.shape == (999,)
    sph_max_overlaps(rows, cols, backend='unbiased_iou')
    if D == 4:
        sph_iou(rows, cols); fov_iou(rows, cols); sph_iou(cols[:500], cols[500:], is_aligned=True)
    p, t = S.loss_pairs(700, box=box)
    for mode in ("iou", "ciou"):
        q = p.to(dev).requires_grad_(True)
        Sph2PobIoULoss(mode=mode)(q, t.to(dev)).backward()
        assert torch.isfinite(q.grad).all()
    boxes = (S.generate_boxes(100, alpha_range=(5, 60), beta_range=(5, 60), box=box, seed=3).repeat(5, 1) + torch.randn(500, D)).clamp(min=1).to(dev)
    scores = torch.rand(500, device=dev)
    SphNMS()(boxes, scores, torch.randint(0, 4, (500,), device=dev), dict(iou_threshold=0.5))
    SphNMS()(boxes, scores, torch.zeros(500, dtype=torch.long, device=dev), dict(iou_threshold=0.5))
    for calc in ('naive_iou', 'unbiased_iou'):
        SphNMS(calc)(boxes, scores, torch.randint(0, 4, (500,), device=dev), dict(iou_threshold=0.5))
        sph_nms_image_blocks(boxes, scores, torch.randint(0, 4, (500,), device=dev), 5, 4, 0.5, 50, iou_calculator=calc)
    sph_batched_nms_images(boxes, scores, torch.randint(0, 4, (500,), device=dev), torch.randint(0, 3, (500,), device=dev), 0.5)
torch.cuda.synchronize()
print("sanitize_small ok")
