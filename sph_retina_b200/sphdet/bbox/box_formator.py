"""Box format conversions either side of the IoU path, with the reference's names and conventions
(sphdet/bbox/box_formator.py:17-200).  Every function is ONE launch of ``k_box_format`` behind ``sphk_box_format`` (a row in,
a row out); the Kent transforms of the reference (``Planar2KentTransform``, ``SphBox2KentTransform``) are outside this path.

The in-place clamps ``climp_spherical_boxes`` / ``climp_rotated_boxes`` (:144-163) are the clamp halves of the two jitters,
which the IoU kernels apply themselves; they are not exported."""
from __future__ import annotations

import torch

from ... import _native

__all__ = ['xyxy2xywh', 'xywh2xyxy', 'obb2hbb_wywh', 'obb2hbb_xyxy', 'bfov2rbfov', 'geo2sph', 'sph2geo', 'is_valid_boxes',
           'Sph2PlanarBoxTransform', 'Planar2SphBoxTransform', 'bbox2roi']


def _cols(boxes, n, what):
    assert boxes.dim() == 2 and boxes.size(1) == n, "%s takes [n, %d] boxes, got %s" % (what, n, tuple(boxes.shape))


def xyxy2xywh(boxes):
    """box_formator.py:17-23."""
    _cols(boxes, 4, 'xyxy2xywh')
    return _native.box_format('xyxy2xywh', boxes, 4)


def xywh2xyxy(boxes):
    """box_formator.py:25-31."""
    _cols(boxes, 4, 'xywh2xyxy')
    return _native.box_format('xywh2xyxy', boxes, 4)


def obb2hbb_wywh(obb):
    """box_formator.py:33-50: the axis-aligned box (cx, cy, w', h') around an oriented box (cx, cy, w, h, a rad)."""
    _cols(obb, 5, 'obb2hbb_wywh')
    return _native.box_format('obb2hbb_xywh', obb, 4)


def obb2hbb_xyxy(obb):
    """box_formator.py:52-55."""
    _cols(obb, 5, 'obb2hbb_xyxy')
    return _native.box_format('obb2hbb_xyxy', obb, 4)


def bfov2rbfov(bfovs):
    """box_formator.py:57-61: gamma = 0 appended."""
    _cols(bfovs, 4, 'bfov2rbfov')
    return _native.box_format('bfov2rbfov', bfovs, 5)


def geo2sph(boxes):
    """box_formator.py:64-68: (lon, lat, ...) -> (theta = lon + 180, phi = 90 - lat, ...)."""
    return _native.box_format('geo2sph', boxes, boxes.size(1))


def sph2geo(boxes):
    """box_formator.py:70-74."""
    return _native.box_format('sph2geo', boxes, boxes.size(1))


def is_valid_boxes(boxes, mode='sph', need_raise=False):
    """box_formator.py:120-141 (a host-side range check: min / max reductions, no kernel of this package)."""
    import math
    try:
        if mode == 'sph':
            assert boxes.size(-1) in [4, 5]
            lo, hi = boxes[:, :4].min(dim=0)[0], boxes[:, :4].max(dim=0)[0]
            assert bool((lo >= 0).all()) and bool((hi <= boxes.new_tensor([360., 180., 360., 180.])).all())
        elif mode == 'obb':
            assert boxes.size(-1) == 5
            lo, hi = boxes[:, 2:].min(dim=0)[0], boxes[:, 2:].max(dim=0)[0]
            assert bool((lo >= boxes.new_tensor([0., 0., -math.pi / 2])).all())
            assert bool((hi <= boxes.new_tensor([math.pi, math.pi, math.pi / 2])).all())
    except AssertionError as e:
        if need_raise:
            raise e
        return False
    return True


class Sph2PlanarBoxTransform:
    """box_formator.py:166-182: BFoV -> xyxy of the equirectangular image, RBFoV -> (x, y, w, h, -gamma rad)."""

    def __init__(self, mode='sph2pix', box_version=4):
        assert mode in ['sph2pix', 'sph2tan']
        assert box_version in [4, 5]
        self.box_version = box_version
        self.mode = mode

    def __call__(self, boxes, img_size=(512, 1024), box_version=None):
        box_version = self.box_version if box_version is None else box_version
        _cols(boxes, box_version, 'Sph2PlanarBoxTransform')
        return _native.box_format('sph2planar_pix' if self.mode == 'sph2pix' else 'sph2planar_tan', boxes, box_version, img_size)


class Planar2SphBoxTransform:
    """box_formator.py:185-200: xyxy of the image -> BFoV, or RBFoV with gamma = 0."""

    def __init__(self, mode='sph2pix', box_version=4):
        assert mode in ['sph2pix', 'pix2sph', 'sph2tan', 'tan2sph']
        assert box_version in [4, 5]
        self.box_version = box_version
        self.mode = mode

    def __call__(self, boxes, img_size=(512, 1024), box_version=None):
        box_version = self.box_version if box_version is None else box_version
        _cols(boxes, 4, 'Planar2SphBoxTransform')
        fmt = 'planar2sph_pix' if self.mode in ['sph2pix', 'pix2sph'] else 'planar2sph_tan'
        return _native.box_format(fmt, boxes, box_version, img_size)


def bbox2roi(bbox_list, box_version=4):
    """box_formator.py:229-244: [batch_ind, box] rows of a list of per-image boxes (a concatenation: plain torch)."""
    rois_list = []
    for img_id, bboxes in enumerate(bbox_list):
        if bboxes.size(0) > 0:
            img_inds = bboxes.new_full((bboxes.size(0), 1), img_id)
            rois = torch.cat([img_inds, bboxes[:, :box_version]], dim=-1)
        else:
            rois = bboxes.new_zeros((0, box_version + 1))
        rois_list.append(rois)
    return torch.cat(rois_list, 0)
