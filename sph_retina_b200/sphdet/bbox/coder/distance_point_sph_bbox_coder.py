"""The anchor-free heads' box coder with the reference's class name, constructor arguments and encode / decode
signatures (sphdet/bbox/coder/distance_point_sph_bbox_coder.py:8-69; ``bbox2distance`` :130-162, ``distance2bbox``
:72-127), registered in mmdet's ``BBOX_CODERS``: a spherical box <-> the (left, top, right, bottom) pixel distances
from a point of the equirectangular image.

Both directions cross the spherical / planar boundary through the box-format kernel of ``sphdet.bbox.box_formator``
(one launch of ``sphk_box_format``: ``Sph2PlanarBoxTransform('sph2pix')`` for encode, ``Planar2SphBoxTransform`` for
decode -- the reference's fp32 expressions operation by operation, so the boxes come out bit-identical); the four
additions / subtractions and the clamps on the pixel side are the reference's own torch expressions.  The format
kernel is linear in the box, so decode stays differentiable (FCOS computes its IoU loss on decoded boxes): the backward of
the kernel call is its transposed constant map."""
from __future__ import annotations

import torch

from .... import _native
from ...registry import BBOX_CODERS


class _Xyxy2Bfov(torch.autograd.Function):
    """box_formator.py:85-92 after :17-23: (x1, y1, x2, y2) -> (theta, phi, alpha, beta) = ((x1 + x2) / 2 / W * 360,
    (y1 + y2) / 2 / H * 180, (x2 - x1) / W * 360, (y2 - y1) / H * 180)."""

    @staticmethod
    def forward(ctx, xyxy, img_shape):
        ctx.kx, ctx.ky = 360.0 / float(img_shape[1]), 180.0 / float(img_shape[0])
        return _native.box_format('planar2sph_pix', xyxy, 4, img_shape)

    @staticmethod
    def backward(ctx, g):
        gt, gp, ga, gb = g.unbind(-1)
        kx, ky = ctx.kx, ctx.ky
        return torch.stack([gt * (0.5 * kx) - ga * kx, gp * (0.5 * ky) - gb * ky,
                            gt * (0.5 * kx) + ga * kx, gp * (0.5 * ky) + gb * ky], -1), None


def distance2bbox(points, distance, max_shape=None, img_shape=(512, 1024)):
    """distance_point_sph_bbox_coder.py:72-127.  points [N, 2] (x, y) pixels, distance [N, 4 | 5] (left, top, right,
    bottom[, gamma]); ``max_shape`` (H, W) clamps the pixel box first.  [N, 4 | 5] spherical boxes; a fifth column is
    handed through.  (The reference's batched [B, N, .] branch ends in a ``torch.chunk(., 4, dim=1)`` over the N axis
    and cannot run; two-dimensional input only.)"""
    assert points.dim() == 2 and distance.dim() == 2, 'distance2bbox: [N, 2] points and [N, 4 | 5] distances'
    x1 = points[..., 0] - distance[..., 0]
    y1 = points[..., 1] - distance[..., 1]
    x2 = points[..., 0] + distance[..., 2]
    y2 = points[..., 1] + distance[..., 3]
    if max_shape is not None:
        h, w = float(max_shape[0]), float(max_shape[1])
        x1, x2 = x1.clamp(min=0, max=w), x2.clamp(min=0, max=w)
        y1, y2 = y1.clamp(min=0, max=h), y2.clamp(min=0, max=h)
    xyxy = torch.stack([x1, y1, x2, y2], -1)
    if xyxy.size(0) == 0:
        bboxes = xyxy
    else:
        bboxes = _Xyxy2Bfov.apply(xyxy, tuple(img_shape))
    if distance.size(-1) == 5:
        bboxes = torch.cat([bboxes, distance[..., [-1]]], dim=-1)
    return bboxes


def bbox2distance(points, bbox, max_dis=None, eps=0.1, img_shape=(512, 1024)):
    """distance_point_sph_bbox_coder.py:130-162.  bbox [N, 4 | 5] spherical; [N, 4 | 5] distances (left, top, right,
    bottom[, gamma]), clamped to [0, max_dis - eps] when ``max_dis`` is given.  Targets: not differentiated."""
    angle = bbox[:, -1] if bbox.size(-1) == 5 else None
    sph = bbox[:, :4]
    if sph.size(0) == 0:
        xyxy = sph
    else:
        xyxy = _native.box_format('sph2planar_pix', sph, 4, tuple(img_shape))
    left = points[:, 0] - xyxy[:, 0]
    top = points[:, 1] - xyxy[:, 1]
    right = xyxy[:, 2] - points[:, 0]
    bottom = xyxy[:, 3] - points[:, 1]
    cols = [left, top, right, bottom]
    if max_dis is not None:
        cols = [c.clamp(min=0, max=max_dis - eps) for c in cols]
    if angle is not None:
        cols.append(angle)
    return torch.stack(cols, -1)


@BBOX_CODERS.register_module()
class DistancePointSphBBoxCoder:
    """distance_point_sph_bbox_coder.py:8-69."""

    def __init__(self, clip_border=True, box_version=4, img_shape=None):
        self.clip_border = clip_border
        self.box_version = box_version
        self.img_shape = img_shape

    def encode(self, points, gt_bboxes, max_dis=None, eps=0.1, img_shape=(512, 1024)):
        assert points.size(0) == gt_bboxes.size(0)
        assert points.size(-1) == 2
        assert gt_bboxes.size(-1) == self.box_version
        img_shape = self.img_shape if self.img_shape else img_shape
        return bbox2distance(points, gt_bboxes, max_dis, eps, img_shape)

    def decode(self, points, pred_bboxes, max_shape=None, img_shape=(512, 1024)):
        assert points.size(0) == pred_bboxes.size(0)
        assert points.size(-1) == 2
        assert pred_bboxes.size(-1) == self.box_version
        if self.clip_border is False:
            max_shape = None
        img_shape = self.img_shape if self.img_shape else img_shape
        return distance2bbox(points, pred_bboxes, max_shape, img_shape)
