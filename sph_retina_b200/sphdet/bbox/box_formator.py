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


# ---- autograd through the conversions ------------------------------------------------------------------------------
# In the reference these functions are plain torch expressions, so gradients flow through them (obb2hbb_xyxy sits inside
# the differentiable GIoU / DIoU / CIoU epilogue, sph2pob_iou_loss.py:142-194).  The kernel call itself is opaque to
# autograd; each differentiable format therefore gets an autograd.Function whose forward is the one kernel launch:
#   * the affine formats (every pixel / degree rescaling and column shuffle): y = x M + c, backward = grad M^T, with M
#     read off the kernel itself once per (format, image size) by pushing the unit rows through it;
#   * obb2hbb_*: the analytic derivative with torch's sub-gradient conventions for |cos a|, |sin a| (sign(0) = 0);
#   * the tangent-plane formats (sph2tan) have no backward here: asking for one raises instead of returning zeros.
_AFFINE = {'xyxy2xywh', 'xywh2xyxy', 'bfov2rbfov', 'geo2sph', 'sph2geo', 'sph2planar_pix', 'planar2sph_pix'}
_jacobians = {}


def _affine_matrix(fmt, d_in, d_out, img_size, device):
    key = (fmt, d_in, d_out, float(img_size[0]), float(img_size[1]), str(device))
    M = _jacobians.get(key)
    if M is None:
        # the map is affine (no clamp, no wrap in these formats): rows of the Jacobian from steps of 64 off a base
        # point (a power of two, so that the division is exact and the fp32 rounding of the outputs weighs 64 x less)
        base = torch.tensor([[128.0, 64.0, 128.0, 64.0, 16.0][:d_in]], device=device)
        probe = torch.cat([base, base + 64.0 * torch.eye(d_in, device=device)])
        out = _native.box_format(fmt, probe, d_out, img_size).double()
        M = ((out[1:] - out[:1]) / 64.0).float()          # [d_in, d_out]
        _jacobians[key] = M
    return M


class _AffineFormat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, boxes, fmt, d_out, img_size):
        ctx.args = (fmt, boxes.size(1), d_out, img_size)
        return _native.box_format(fmt, boxes, d_out, img_size)

    @staticmethod
    def backward(ctx, grad):
        fmt, d_in, d_out, img_size = ctx.args
        M = _affine_matrix(fmt, d_in, d_out, img_size, grad.device)
        return (grad.float() @ M.t()).to(grad.dtype), None, None, None


class _Obb2Hbb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, obb, fmt):
        ctx.save_for_backward(obb)
        ctx.fmt = fmt
        return _native.box_format(fmt, obb, 4)

    @staticmethod
    def backward(ctx, grad):
        (obb,) = ctx.saved_tensors
        w, h, a = obb[:, 2], obb[:, 3], obb[:, 4]
        c, s = torch.cos(a), torch.sin(a)
        if ctx.fmt == 'obb2hbb_xyxy':          # (cx - W/2, cy - H/2, cx + W/2, cy + H/2)
            g_cx, g_cy = grad[:, 0] + grad[:, 2], grad[:, 1] + grad[:, 3]
            g_W, g_H = 0.5 * (grad[:, 2] - grad[:, 0]), 0.5 * (grad[:, 3] - grad[:, 1])
        else:                                  # (cx, cy, W, H)
            g_cx, g_cy, g_W, g_H = grad[:, 0], grad[:, 1], grad[:, 2], grad[:, 3]
        # W = |cos a| w + |sin a| h,  H = |sin a| w + |cos a| h
        dc, ds = -torch.sign(c) * s, torch.sign(s) * c          # d|cos a|/da, d|sin a|/da
        g_w = g_W * c.abs() + g_H * s.abs()
        g_h = g_W * s.abs() + g_H * c.abs()
        g_a = g_W * (dc * w + ds * h) + g_H * (ds * w + dc * h)
        return torch.stack([g_cx, g_cy, g_w, g_h, g_a], dim=1), None


def _format(fmt, boxes, d_out, img_size=(512, 1024)):
    """One launch of k_box_format; differentiable where the reference's torch expression is."""
    if boxes.requires_grad and torch.is_grad_enabled():
        if fmt in _AFFINE:
            return _AffineFormat.apply(boxes, fmt, d_out, tuple(img_size))
        if fmt in ('obb2hbb_xywh', 'obb2hbb_xyxy'):
            return _Obb2Hbb.apply(boxes, fmt)
        raise NotImplementedError("box format %r has no backward on this path: detach the boxes, or convert them before "
                                  "they require grad (silently returning a tensor cut off from autograd would zero the "
                                  "gradients without an error)" % fmt)
    return _native.box_format(fmt, boxes, d_out, img_size)


def xyxy2xywh(boxes):
    """box_formator.py:17-23."""
    _cols(boxes, 4, 'xyxy2xywh')
    return _format('xyxy2xywh', boxes, 4)


def xywh2xyxy(boxes):
    """box_formator.py:25-31."""
    _cols(boxes, 4, 'xywh2xyxy')
    return _format('xywh2xyxy', boxes, 4)


def obb2hbb_wywh(obb):
    """box_formator.py:33-50: the axis-aligned box (cx, cy, w', h') around an oriented box (cx, cy, w, h, a rad)."""
    _cols(obb, 5, 'obb2hbb_wywh')
    return _format('obb2hbb_xywh', obb, 4)


def obb2hbb_xyxy(obb):
    """box_formator.py:52-55."""
    _cols(obb, 5, 'obb2hbb_xyxy')
    return _format('obb2hbb_xyxy', obb, 4)


def bfov2rbfov(bfovs):
    """box_formator.py:57-61: gamma = 0 appended."""
    _cols(bfovs, 4, 'bfov2rbfov')
    return _format('bfov2rbfov', bfovs, 5)


def geo2sph(boxes):
    """box_formator.py:64-68: (lon, lat, ...) -> (theta = lon + 180, phi = 90 - lat, ...)."""
    return _format('geo2sph', boxes, boxes.size(1))


def sph2geo(boxes):
    """box_formator.py:70-74."""
    return _format('sph2geo', boxes, boxes.size(1))


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
        return _format('sph2planar_pix' if self.mode == 'sph2pix' else 'sph2planar_tan', boxes, box_version, img_size)


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
        return _format(fmt, boxes, box_version, img_size)


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
