"""The spherical delta box coders with the reference's class names, constructor arguments and encode / decode
signatures (sphdet/bbox/coder/delta_xywh_sph_bbox_coder.py:10-115 for BFoV (theta, phi, alpha, beta),
delta_xywha_rsph_bbox_coder.py:10-115 for RBFoV (+ gamma)), registered in mmdet's ``BBOX_CODERS``.

Reference: ~25 eager kernels per decode (repeat, reshape, clamps, exp, cat, four in-place clamps); here one launch
(``sphk_coder_decode`` / ``sphk_coder_encode``), differentiable w.r.t. the deltas (``sphk_coder_decode_bwd``: zero
where one of the decode's clamps is active, exactly torch.clamp's backward).  The head's training step does not
need the decoded boxes at all: see ``sphdet.losses.Sph2PobDecodedIoULoss``."""
from __future__ import annotations

import torch

from .... import _native
from ...registry import BBOX_CODERS


class _Decode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rois, deltas, kw):
        ctx.save_for_backward(rois.detach(), deltas.detach())
        ctx.kw = kw
        ctx.in_dtype = deltas.dtype
        return _native.coder_decode(rois.detach(), deltas.detach(), **kw).to(deltas.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        rois, deltas = ctx.saved_tensors
        return None, _native.coder_decode(rois, deltas, grad_out=grad_out, **ctx.kw).to(ctx.in_dtype), None


def delta2bbox(rois, deltas, means=None, stds=None, max_shape=None, wh_ratio_clip=16 / 1000, clip_border=True,
               add_ctr_clamp=False, ctr_clamp=32):
    """delta_xywh_sph_bbox_coder.py:165-262 / delta_xywha_rsph_bbox_coder.py:167-268.  rois [N, D]; deltas [N, D] or
    [N, num_classes * D] (every class decoded against the same roi, :238 / :244).  ``max_shape`` is accepted and
    ignored like in the reference (its border clamp is hard-wired to the spherical ranges, :255-259)."""
    D = rois.size(-1)
    if deltas.size(0) == 0:
        return deltas
    n, cols = deltas.size(0), deltas.size(1)
    if cols % D != 0:
        raise ValueError("deltas with %d columns do not hold whole %d-column boxes" % (cols, D))
    kw = dict(means=means, stds=stds, wh_ratio_clip=wh_ratio_clip, clip_border=clip_border, add_ctr_clamp=add_ctr_clamp,
              ctr_clamp=ctr_clamp)
    if cols != D:
        rois = rois.repeat(1, cols // D).reshape(-1, D)
        deltas = deltas.reshape(-1, D)
    if deltas.requires_grad and torch.is_grad_enabled():
        out = _Decode.apply(rois, deltas, kw)
    else:
        out = _native.coder_decode(rois, deltas, **kw).to(deltas.dtype)
    return out.reshape(n, -1)


def bbox2delta(proposals, gt, means=None, stds=None):
    """delta_xywh_sph_bbox_coder.py:117-162 / delta_xywha_rsph_bbox_coder.py:117-164 (float32 like the reference,
    which casts its inputs)."""
    assert proposals.size() == gt.size()
    lead = proposals.shape[:-1]
    D = proposals.size(-1)
    out = _native.coder_encode(proposals.reshape(-1, D), gt.reshape(-1, D), means=means, stds=stds)
    return out.reshape(*lead, D)


class _DeltaSphCoderBase:
    box_version = 4

    def __init__(self, target_means=None, target_stds=None, clip_border=True, add_ctr_clamp=False, ctr_clamp=32):
        D = self.box_version
        self.means = tuple(target_means) if target_means is not None else (0.0,) * D
        self.stds = tuple(target_stds) if target_stds is not None else (1.0,) * D
        self.clip_border = clip_border
        self.add_ctr_clamp = add_ctr_clamp
        self.ctr_clamp = ctr_clamp

    def encode(self, bboxes, gt_bboxes):
        assert bboxes.size(0) == gt_bboxes.size(0)
        assert bboxes.size(-1) == gt_bboxes.size(-1) == self.box_version
        return bbox2delta(bboxes, gt_bboxes, self.means, self.stds)

    def decode(self, bboxes, pred_bboxes, max_shape=None, wh_ratio_clip=16 / 1000):
        assert pred_bboxes.size(0) == bboxes.size(0)
        if pred_bboxes.ndim == 3:
            assert pred_bboxes.size(1) == bboxes.size(1)
            # the reference raises here too (its onnx/batched branch ends in `raise NotImplemented(...)`, :105)
            raise NotImplementedError('omnx function is not implement!')
        return delta2bbox(bboxes, pred_bboxes, self.means, self.stds, max_shape, wh_ratio_clip, self.clip_border,
                          self.add_ctr_clamp, self.ctr_clamp)

    def kernel_kwargs(self, wh_ratio_clip=16 / 1000):
        """The decode parameters in the form the fused decode + loss kernel takes them."""
        return dict(means=self.means, stds=self.stds, wh_ratio_clip=wh_ratio_clip, clip_border=self.clip_border,
                    add_ctr_clamp=self.add_ctr_clamp, ctr_clamp=self.ctr_clamp)


@BBOX_CODERS.register_module()
class DeltaXYWHSphBBoxCoder(_DeltaSphCoderBase):
    """sphdet/bbox/coder/delta_xywh_sph_bbox_coder.py:10-115 -- (theta, phi, alpha, beta) in degrees."""
    box_version = 4


@BBOX_CODERS.register_module()
class DeltaXYWHASphBBoxCoder(_DeltaSphCoderBase):
    """sphdet/bbox/coder/delta_xywha_rsph_bbox_coder.py:10-115 -- (theta, phi, alpha, beta, gamma) in degrees; the
    gamma delta is in radians (:153, :252)."""
    box_version = 5
