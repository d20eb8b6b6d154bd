"""IoU API with the reference's names and signatures (sphdet/iou/sph_iou_api.py:88-177).

Each function is the whole reference pipeline -- pair expansion, jiter_spherical_bboxes, the
Sph2Pob transform, jiter_rotated_bboxes, rotated-box IoU, view, clamp (sph_iou_api.py:48-86) --
executed by ONE CUDA kernel launch (no expansion is materialised, no intermediate reaches HBM).
Inputs are never modified (tests/test_all_ious.py:322-331)."""
from __future__ import annotations

import torch

from ... import _native

__all__ = ["sph2pob_standard_iou", "sph2pob_efficient_iou", "sph2pob_legacy_iou", "fov_iou", "sph_iou", "naive_iou", "unbiased_iou"]


def _empty(bboxes1, rows, cols, is_aligned):
    # sph_iou_api.py:56-57 returns an uninitialised tensor of this shape; zeros here
    return bboxes1.new_zeros((rows, 1)) if is_aligned else bboxes1.new_zeros((rows, cols))


def _run(kind, bboxes1, bboxes2, mode, is_aligned, edge, angle="equator"):
    rows, cols = bboxes1.size(0), bboxes2.size(0)
    if rows * cols == 0:
        return _empty(bboxes1, rows, cols, is_aligned)
    with torch.no_grad():
        if is_aligned:
            assert rows == cols
            out = _native.iou_aligned(kind, bboxes1, bboxes2, mode, edge, angle)
        else:
            out = _native.iou_pairwise(kind, bboxes1, bboxes2, mode, edge, angle=angle)[0]
    return out if out.dtype == bboxes1.dtype else out.to(bboxes1.dtype)


def _sph2pob_iou(kind, bboxes1, bboxes2, mode, is_aligned, calculator, rbb_edge, rbb_angle):
    # sph_iou_api.py:49-51
    assert mode in ['iou', 'iof']
    assert calculator in ['common', 'diff']
    assert rbb_edge in ['arc', 'chord', 'tangent']
    assert rbb_angle in ['equator', 'project']      # sph2pob_efficient.py:27
    return _run(kind, bboxes1, bboxes2, mode, is_aligned, rbb_edge, rbb_angle)


def sph2pob_standard_iou(bboxes1, bboxes2, mode='iou', is_aligned=False, calculator='common', rbb_edge='arc',
                         rbb_angle='equator'):
    """sphdet/iou/sph_iou_api.py:94-95."""
    return _sph2pob_iou("sph2pob_standard", bboxes1, bboxes2, mode, is_aligned, calculator, rbb_edge, rbb_angle)


def sph2pob_efficient_iou(bboxes1, bboxes2, mode='iou', is_aligned=False, calculator='common', rbb_edge='arc',
                          rbb_angle='equator'):
    """sphdet/iou/sph_iou_api.py:97-98."""
    return _sph2pob_iou("sph2pob_efficient", bboxes1, bboxes2, mode, is_aligned, calculator, rbb_edge, rbb_angle)


def sph2pob_legacy_iou(bboxes1, bboxes2, mode='iou', is_aligned=False, calculator='common', rbb_edge='arc'):
    """sphdet/iou/sph_iou_api.py:91-92: the hand-crafted first version of the transform (sph2pob_legacy.py:8-31).  BFoV
    only, as in the reference (``torch.chunk(box, 4)`` at sph2pob_legacy.py:52-53 fails on five columns)."""
    if bboxes1.size(0) * bboxes2.size(0) != 0 and (bboxes1.size(-1) != 4 or bboxes2.size(-1) != 4):
        raise ValueError("sph2pob_legacy_iou takes BFoV boxes [n, 4] (sph2pob_legacy.py:52-53)")
    return _sph2pob_iou("sph2pob_legacy", bboxes1, bboxes2, mode, is_aligned, calculator, rbb_edge, 'equator')


def sph_iou(bboxes1, bboxes2, mode='iou', is_aligned=False, calculator='diff'):
    """sphdet/iou/sph_iou_api.py:130-151 (+ approximate_ious.py:3-25).  BFoV only."""
    assert mode in ['iou']
    return _run("sph", bboxes1, bboxes2, mode, is_aligned, "arc")


def fov_iou(bboxes1, bboxes2, mode='iou', is_aligned=False, calculator='diff'):
    """sphdet/iou/sph_iou_api.py:156-177 (+ approximate_ious.py:28-55).  BFoV only."""
    assert mode in ['iou']
    return _run("fov", bboxes1, bboxes2, mode, is_aligned, "arc")


def naive_iou(bboxes1, bboxes2, mode='iou', is_aligned=False, box_formator='sph2pix'):
    """sphdet/iou/sph_iou_api.py:181-198: the boxes read as planar boxes of the 512 x 1024 equirectangular image
    (``Sph2PlanarBoxTransform('sph2pix')``) and mmcv's planar IoU (``bbox_overlaps`` for BFoV, ``box_iou_rotated`` for
    RBFoV), no jitter, no clamp -- the calculator the reference's indoor360 configs give to the test-time NMS."""
    assert mode in ['iou']
    if box_formator != 'sph2pix':
        raise NotImplementedError("naive_iou: only box_formator='sph2pix' (the reference's default) has a kernel")
    return _run("naive", bboxes1, bboxes2, mode, is_aligned, "arc")


def unbiased_iou(bboxes1, bboxes2, mode='iou', is_aligned=False):
    """sphdet/iou/sph_iou_api.py:103-125: the exact spherical IoU (jiter_spherical_bboxes, then the Unbiased-IoU classes of
    unbiased_iou_bfov.py / unbiased_iou_rbfov.py, clamp) -- the default backend of ``SphOverlaps2D`` and the calculator the
    reference's pandora configs give to the test-time NMS.  The reference runs it in numpy on the CPU (40 s per million
    pairs); here it is one kernel launch, evaluated in double precision per pair."""
    assert mode in ['iou']
    return _run("unbiased", bboxes1, bboxes2, mode, is_aligned, "arc")
