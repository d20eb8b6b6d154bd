"""``SphOverlaps2D`` / ``sph_overlaps`` -- the drop-in boundary used by ``MaxIoUAssigner``
(sphdet/iou/sph_iou_calculator.py:8-113; mmdet/core/bbox/assigners/max_iou_assigner.py:65,113)."""
from __future__ import annotations

import torch

from ..registry import IOU_CALCULATORS
from .sph_iou_api import (fov_iou, naive_iou, sph2pob_efficient_iou, sph2pob_legacy_iou, sph2pob_standard_iou, sph_iou,
                          unbiased_iou)

# backends of the reference that have a CUDA kernel here; the others are out of this path's scope
_BACKENDS = {
    'sph2pob_standard_iou': sph2pob_standard_iou,
    'sph2pob_efficient_iou': sph2pob_efficient_iou,
    'fov_iou': fov_iou,
    'sph_iou': sph_iou,
    'naive_iou': naive_iou,
    'unbiased_iou': unbiased_iou,
    'sph2pob_legacy_iou': sph2pob_legacy_iou,
}
_REFERENCE_BACKENDS = ['unbiased_iou', 'sph2pob_standard_iou', 'sph2pob_legacy_iou', 'sph2pob_efficient_iou',
                       'naive_iou', 'fov_iou', 'sph_iou', 'kent_iou']


@IOU_CALCULATORS.register_module()
class SphOverlaps2D(object):
    """2D Overlaps Calculator for spherical boxes (sph_iou_calculator.py:8-51).

    Signature and defaults are the reference's, default backend 'unbiased_iou' included (a CPU numpy routine there, a
    double-precision kernel here); the spherical configs pass backend='sph2pob_efficient_iou'
    (configs/_base_/models/sph_rotated_retinanet_r50_fpn.py:18-19).  Of the reference's eight backend names only 'kent_iou'
    (the Kent-distribution code, out of scope) has no kernel."""

    def __init__(self, backend='unbiased_iou', box_version=4):
        self.backend = backend
        self.box_version = box_version

    def __call__(self, bboxes1, bboxes2, mode='iou', is_aligned=False):
        assert bboxes1.size(-1) in [0, 4, 5, 6]
        assert bboxes2.size(-1) in [0, 4, 5, 6]
        bv = self.box_version
        if bboxes1.size(-1) != bv:          # (a slice of the full width is the tensor itself: skip the view, ~2 us each)
            bboxes1 = bboxes1[..., :bv]
        if bboxes2.size(-1) != bv:
            bboxes2 = bboxes2[..., :bv]
        # the kernels never record autograd history (the reference wraps this call in torch.no_grad())
        return sph_overlaps(bboxes1, bboxes2, mode, is_aligned, self.backend)

    def __repr__(self):
        return self.__class__.__name__ + '()'


def sph_overlaps(bboxes1, bboxes2, mode='iou', is_aligned=False, backend='unbiased_iou'):
    """sph_iou_calculator.py:58-113: (m, n) overlaps, or (m,) when ``is_aligned``."""
    assert mode in ['iou', 'iof']
    assert backend in _REFERENCE_BACKENDS
    rows, cols = bboxes1.size(0), bboxes2.size(0)
    if rows * cols == 0:
        return bboxes1.new_zeros((rows, 1)) if is_aligned else bboxes1.new_zeros((rows, cols))
    fn = _BACKENDS.get(backend)
    if fn is None:
        raise NotImplementedError("backend %r is outside the B200 hot path (no CUDA kernel, no fallback); "
                                  "available: %s" % (backend, sorted(_BACKENDS)))
    return fn(bboxes1, bboxes2, mode, is_aligned)
