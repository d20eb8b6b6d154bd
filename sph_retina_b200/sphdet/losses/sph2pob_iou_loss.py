"""Sph2Pob IoU losses with the reference's class names, constructor arguments and forward signature
(sphdet/losses/sph2pob_iou_loss.py:16-58,199-296; sphdet/losses/sph2pob_transform.py:11-37).

Reference dataflow: clone -> jiter_spherical_bboxes -> sph2pob_standard -> jiter_rotated_bboxes ->
diff_iou_rotated_2d -> clamp -> 1 - iou [-> GIoU/DIoU/CIoU epilogue] -> weight_reduce_loss, all as
eager autograd ops.  Here:
  * mode 'iou': ONE kernel launch computes the IoU of every pair AND d(iou)/d(pred), d(iou)/d(target)
    in registers (recompute-free: the backward pass only scales the stored per-row gradients by the
    incoming d(loss)/d(iou));
  * modes 'giou'/'diou'/'ciou': the OBBs come from sphk_obb_fwd (backward sphk_obb_bwd), the rotated
    IoU from sphk_riou_fwd_bwd, and the cheap enclosing-box epilogue stays torch autograd on the OBBs
    exactly as sph2pob_iou_loss.py:142-194."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ... import _native
from ..registry import LOSSES


# ------------------------------------------------------------------------------------------------
# autograd bridges to the kernels
# ------------------------------------------------------------------------------------------------
class _Sph2PobIoU(torch.autograd.Function):
    """iou[n] = clamp(rotated_iou(sph2pob_standard(jitter(pred, target)))) with analytic gradients."""

    @staticmethod
    def forward(ctx, pred, target):
        need_p, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if need_p or need_t:
            iou, gp, gt = _native.loss_fwd_bwd(pred.detach(), target.detach(), None, need_p, need_t)
            ctx.save_for_backward(*[g for g in (gp, gt) if g is not None])
            ctx.have = (need_p, need_t)
        else:
            iou, _, _ = _native.loss_fwd_bwd(pred.detach(), target.detach())
            ctx.have = (False, False)
        ctx.in_dtypes = (pred.dtype, target.dtype)
        return iou.to(pred.dtype)

    @staticmethod
    def backward(ctx, grad_iou):
        saved = list(ctx.saved_tensors)
        gp = saved.pop(0) if ctx.have[0] else None
        gt = saved.pop(0) if ctx.have[1] else None
        g = grad_iou.float().unsqueeze(1)
        return (None if gp is None else (gp * g).to(ctx.in_dtypes[0]),
                None if gt is None else (gt * g).to(ctx.in_dtypes[1]))


class _Sph2PobReducedLoss(torch.autograd.Function):
    """scale * sum_i w_i (1 - iou_i) and its gradients from ONE kernel launch (mode 'iou', reduction mean / sum)."""

    @staticmethod
    def forward(ctx, pred, target, weight, scale):
        need_p, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        # the kernel writes the finished scalar: the forward of a training step is this one launch, no reduction op
        total, gp, gt = _native.loss_reduce_total(pred.detach(), target.detach(), None if weight is None else weight.detach(),
                                                  scale, need_p, need_t)
        ctx.save_for_backward(*[g for g in (gp, gt) if g is not None])
        ctx.have = (need_p, need_t)
        ctx.in_dtypes = (pred.dtype, target.dtype)
        return total if pred.dtype == torch.float32 else total.to(pred.dtype)

    @staticmethod
    def backward(ctx, grad_loss):
        saved = list(ctx.saved_tensors)
        gp = saved.pop(0) if ctx.have[0] else None
        gt = saved.pop(0) if ctx.have[1] else None
        g = grad_loss if grad_loss.dtype == torch.float32 else grad_loss.float()
        # the stored gradients are already d(loss)/d(box) for an upstream gradient of 1: one multiply per operand
        return (None if gp is None else (gp * g).to(ctx.in_dtypes[0]),
                None if gt is None else (gt * g).to(ctx.in_dtypes[1]), None, None)


class _Sph2PobObbs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, kind):
        o1, o2 = _native.obb_fwd(kind, pred.detach(), target.detach())
        ctx.save_for_backward(pred.detach(), target.detach())
        ctx.kind = kind
        ctx.in_dtypes = (pred.dtype, target.dtype)
        return o1.to(pred.dtype), o2.to(pred.dtype)

    @staticmethod
    def backward(ctx, g1, g2):
        pred, target = ctx.saved_tensors
        gb1, gb2 = _native.obb_bwd(ctx.kind, pred, target, g1, g2, want1=ctx.needs_input_grad[0],
                                   want2=ctx.needs_input_grad[1])
        return (None if gb1 is None else gb1.to(ctx.in_dtypes[0]),
                None if gb2 is None else gb2.to(ctx.in_dtypes[1]), None)


class _RotatedIoU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, o1, o2):
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if need1 or need2:
            iou, g1, g2 = _native.riou_fwd_bwd(o1.detach(), o2.detach(), None, need1, need2)
            ctx.save_for_backward(*[g for g in (g1, g2) if g is not None])
        else:
            iou, _, _ = _native.riou_fwd_bwd(o1.detach(), o2.detach())
        ctx.have = (need1, need2)
        ctx.in_dtype = o1.dtype
        return iou.to(o1.dtype)

    @staticmethod
    def backward(ctx, grad_iou):
        saved = list(ctx.saved_tensors)
        g1 = saved.pop(0) if ctx.have[0] else None
        g2 = saved.pop(0) if ctx.have[1] else None
        g = grad_iou.float().unsqueeze(1)
        return (None if g1 is None else (g1 * g).to(ctx.in_dtype), None if g2 is None else (g2 * g).to(ctx.in_dtype))


def sph2pob_iou(pred, target):
    """Differentiable clamped Sph2Pob-standard IoU of aligned spherical boxes [n, 4|5] (degrees)."""
    return _Sph2PobIoU.apply(pred, target)


def sph2pob_obbs(pred, target, transform='sph2pob_standard'):
    """The planar OBBs (x, y, w, h, angle rad) after jitter -> transform -> jitter
    (sph2pob_transform.py:26-30), differentiable."""
    return _Sph2PobObbs.apply(pred, target, transform)


def rotated_iou(obb1, obb2):
    """Differentiable rotated-box IoU of aligned OBBs [n, 5] (diff_iou_rotated_2d semantics), clamped to [0, 1]."""
    return _RotatedIoU.apply(obb1, obb2)


# ------------------------------------------------------------------------------------------------
# reduction: mmdet/models/losses/utils.py (weight_reduce_loss)
# ------------------------------------------------------------------------------------------------
def _weight_reduce_loss(loss, weight=None, reduction='mean', avg_factor=None):
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        if reduction == 'mean':
            return loss.mean()
        if reduction == 'sum':
            return loss.sum()
        return loss
    if reduction == 'mean':
        return loss.sum() / (avg_factor + torch.finfo(torch.float32).eps)
    if reduction != 'none':
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss


def _obb2hbb_xyxy(obb):
    """sphdet/bbox/box_formator.py:34-54."""
    w, h, a = obb[:, 2], obb[:, 3], obb[:, 4]
    cosa, sina = torch.cos(a).abs(), torch.sin(a).abs()
    hw, hh = (cosa * w + sina * h) / 2, (sina * w + cosa * h) / 2
    return torch.stack((obb[:, 0] - hw, obb[:, 1] - hh, obb[:, 0] + hw, obb[:, 1] + hh), -1)


def _obb_epilogue(ious, pred, target, mode, eps):
    """sph2pob_iou_loss.py:142-194 on the transformed OBBs."""
    hbb_pred, hbb_target = _obb2hbb_xyxy(pred), _obb2hbb_xyxy(target)
    enclose_wh = (torch.max(hbb_pred[:, 2:], hbb_target[:, 2:]) -
                  torch.min(hbb_pred[:, :2], hbb_target[:, :2])).clamp(min=0)
    if mode == 'giou':
        inter_wh = (torch.min(hbb_pred[:, 2:], hbb_target[:, 2:]) -
                    torch.max(hbb_pred[:, :2], hbb_target[:, :2])).clamp(min=0)
        area_enclose = enclose_wh[:, 0] * enclose_wh[:, 1]
        area_union = pred[:, 2] * pred[:, 3] + target[:, 2] * target[:, 3] - inter_wh[:, 0] * inter_wh[:, 1]
        area_ratio = (area_enclose - area_union) / (area_enclose + eps)
        return 1 - (ious - area_ratio.clamp(min=0, max=1.0))
    c2 = enclose_wh[:, 0] ** 2 + enclose_wh[:, 1] ** 2 + eps
    rho2 = (target[:, 0] - pred[:, 0]) ** 2 + (target[:, 1] - pred[:, 1]) ** 2
    if mode == 'diou':
        return 1 - (ious - (rho2 / c2).clamp(min=0, max=1.0))
    factor = 4 / math.pi ** 2
    v = factor * torch.pow(torch.atan(target[:, 2] / (target[:, 3] + eps)) - torch.atan(pred[:, 2] / (pred[:, 3] + eps)), 2)
    with torch.no_grad():
        alpha = (ious > 0.5).float() * v / (1 - ious + v + eps)
    if mode == 'ciou':
        return 1 - (ious - ((rho2 / c2).clamp(min=0, max=1.0) + alpha * v))
    raise NotImplementedError('Not supported version of iou-based loss.')


def _elementwise_loss(pred, target, mode, eps, transform='sph2pob_standard'):
    if mode == 'iou' and transform == 'sph2pob_standard':
        return 1 - sph2pob_iou(pred, target)
    o1, o2 = sph2pob_obbs(pred, target, transform)
    ious = rotated_iou(o1, o2)
    if mode == 'iou':
        return 1 - ious
    return _obb_epilogue(ious, o1, o2, mode, eps)


class _SphLossBase(nn.Module):
    _transform = 'sph2pob_standard'

    def __init__(self, mode='iou', eps=1e-6, reduction='mean', loss_weight=1.0):
        super().__init__()
        assert mode in ['iou', 'giou', 'diou', 'ciou']
        self.mode = mode
        self.eps = eps
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        box_version = target.size(-1)
        # sph2pob_transform.py:32-34: a 2-D BFoV weight gains the column the OBB angle would use
        if weight is not None and weight.dim() > 1 and box_version == 4:
            weight = torch.cat([weight, weight.mean(-1, keepdim=True)], dim=-1)
        # sph2pob_iou_loss.py:36-39: nothing positive -> a zero that still hangs off `pred`
        if weight is not None and not torch.any(weight > 0):
            return pred.sum() * 0
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        if weight is not None and weight.dim() > 1:
            assert weight.shape == (pred.size(0), 5)    # :46 weight.shape == obb pred.shape
            weight = weight.mean(-1)
        if (self.mode == 'iou' and self._transform == 'sph2pob_standard' and reduction in ('mean', 'sum') and pred.is_cuda
                and not isinstance(avg_factor, torch.Tensor)):     # a tensor avg_factor would need a host sync for the scale
            # fused: elementwise loss, weights, reduction and both gradients in one kernel launch
            if avg_factor is None:
                scale = self.loss_weight / max(pred.size(0), 1) if reduction == 'mean' else self.loss_weight
            elif reduction == 'mean':
                scale = self.loss_weight / (avg_factor + torch.finfo(torch.float32).eps)
            else:
                raise ValueError('avg_factor can not be used with reduction="sum"')
            if pred.size(0) > 0:
                return _Sph2PobReducedLoss.apply(pred, target, weight, float(scale))
        loss = _elementwise_loss(pred, target, self.mode, self.eps, self._transform)
        return self.loss_weight * _weight_reduce_loss(loss, weight, reduction, avg_factor)


class _DecodedReducedLoss(torch.autograd.Function):
    """scale * sum_i w_i (1 - iou(decode(anchor_i, delta_i), target_i)) and d/d(deltas) from ONE kernel launch."""

    @staticmethod
    def forward(ctx, anchors, deltas, target, weight, scale, coder_kw):
        need = ctx.needs_input_grad[1]
        partial, grad = _native.decode_loss_reduce(anchors.detach(), deltas.detach(), target.detach(),
                                                   None if weight is None else weight.detach(), scale, want_grad=need, **coder_kw)
        if need:
            ctx.save_for_backward(grad)
        ctx.in_dtype = deltas.dtype
        return (partial.sum() * scale).to(deltas.dtype)

    @staticmethod
    def backward(ctx, grad_loss):
        (grad,) = ctx.saved_tensors if ctx.needs_input_grad[1] else (None,)
        return None, (None if grad is None else (grad * grad_loss.float()).to(ctx.in_dtype)), None, None, None, None


class OBBIoULoss(_SphLossBase):
    """Name kept for parity with sph2pob_iou_loss.py:16; here it already includes the Sph2Pob transform
    (the reference applies it through the ``Sph2PobTransfrom`` class decorator)."""


@LOSSES.register_module()
class Sph2PobIoULoss(_SphLossBase):
    """sphdet/losses/sph2pob_iou_loss.py:220-236 -- ``Sph2PobTransfrom()(OBBIoULoss)``.

    pred / target: spherical boxes [n, 4|5] in degrees; weight: None, [n] or [n, box_version]."""


@LOSSES.register_module()
class SphIoULossLegacy(nn.Module):
    """sphdet/losses/sph2pob_iou_loss.py:199-216 -- ``Sph2PobTransfrom()(RotatedIoULoss)``: mmrotate 0.3.2's rotated IoU
    loss (``-log(iou)``, ``1 - iou`` or ``1 - iou^2`` of ``diff_iou_rotated_2d(...).clamp(min=eps)``) on the Sph2Pob OBBs of
    the pair.  The IoU and its two gradients come from the same single launch as ``Sph2PobIoULoss`` (``sphk_loss_fwd_bwd``);
    the scalar map and the reduction are torch ops on the [n] vector.

    pred / target: spherical boxes [n, 4|5] in degrees; weight: None, [n] or [n, box_version]."""

    def __init__(self, linear=False, eps=1e-6, reduction='mean', loss_weight=1.0, mode='log'):
        super().__init__()
        assert mode in ('linear', 'square', 'log')
        self.mode = 'linear' if linear else mode
        self.linear = linear
        self.eps = eps
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        box_version = target.size(-1)
        if weight is not None and weight.dim() > 1 and box_version == 4:       # sph2pob_transform.py:32-34
            weight = torch.cat([weight, weight.mean(-1, keepdim=True)], dim=-1)
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        if weight is not None and not torch.any(weight > 0) and reduction != 'none':
            return pred.sum() * 0                                              # rotated_iou_loss.py: (pred * weight).sum()
        if weight is not None and weight.dim() > 1:
            assert weight.shape == (pred.size(0), 5)
            weight = weight.mean(-1)
        ious = _Sph2PobIoU.apply(pred, target).clamp(min=self.eps)
        if self.mode == 'linear':
            loss = 1 - ious
        elif self.mode == 'square':
            loss = 1 - ious ** 2
        else:
            loss = -ious.log()
        return self.loss_weight * _weight_reduce_loss(loss, weight, reduction, avg_factor)


@LOSSES.register_module()
class SphIoULoss(_SphLossBase):
    """sphdet/losses/sph2pob_iou_loss.py:239-296.  The reference class is non-functional as shipped
    (its calculator='diff' branch is dead code); this one computes what it was written to compute for
    iou_calculator='sph2pob_standard', mode='iou'."""

    def __init__(self, mode='iou', iou_calculator='sph2pob_standard', eps=1e-6, reduction='mean', loss_weight=1.0):
        assert iou_calculator in ['sph2pob_standard', 'sph', 'fov']
        if iou_calculator != 'sph2pob_standard':
            raise NotImplementedError("SphIoULoss: only 'sph2pob_standard' is differentiable on this path")
        if mode != 'iou':
            raise NotImplementedError("SphIoULoss: the reference returns None for mode != 'iou'")
        super().__init__(mode, eps, reduction, loss_weight)


@LOSSES.register_module()
class Sph2PobDecodedIoULoss(Sph2PobIoULoss):
    """``bbox_coder.decode`` + ``Sph2PobIoULoss`` of the head's regression branch in one step (SURVEY.md 8f row 2;
    sphdet/models/heads/sph_retina_head.py:252-265 with reg_decoded_bbox=True):

        loss_bbox = self.loss_bbox.forward_decoded(self.bbox_coder, anchors, bbox_pred, bbox_targets, bbox_weights,
                                                   avg_factor=num_total_samples)

    equals ``self.loss_bbox(self.bbox_coder.decode(anchors, bbox_pred), bbox_targets, bbox_weights, avg_factor=...)``.
    The head calls it on ALL anchors of the batch with zero weights for the negatives; for mode 'iou' one kernel reads
    the weights, skips the zero rows and runs decode, transform, IoU and the whole backward for the others.  The plain
    ``forward(pred, target, ...)`` of ``Sph2PobIoULoss`` is inherited unchanged."""

    def forward_decoded(self, bbox_coder, anchors, bbox_pred, target, weight=None, avg_factor=None, reduction_override=None,
                        wh_ratio_clip=16 / 1000):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        D = target.size(-1)
        fused = (self.mode == 'iou' and reduction in ('mean', 'sum') and bbox_pred.is_cuda and bbox_pred.dim() == 2
                 and bbox_pred.size(1) == D and bbox_pred.size(0) > 0 and not isinstance(avg_factor, torch.Tensor)
                 and (weight is None or weight.dim() == 1 or weight.size(1) == D))
        if not fused:
            return self.forward(bbox_coder.decode(anchors, bbox_pred, wh_ratio_clip=wh_ratio_clip), target, weight,
                                avg_factor=avg_factor, reduction_override=reduction_override)
        if avg_factor is None:
            scale = self.loss_weight / bbox_pred.size(0) if reduction == 'mean' else self.loss_weight
        elif reduction == 'mean':
            scale = self.loss_weight / (avg_factor + torch.finfo(torch.float32).eps)
        else:
            raise ValueError('avg_factor can not be used with reduction="sum"')
        # (no "any positive weight" test: that is a host sync; with all-zero weights the kernel returns exactly the
        # zero loss and zero gradient the reference's early-out produces, sph2pob_iou_loss.py:36-39)
        return _DecodedReducedLoss.apply(anchors, bbox_pred, target, weight, float(scale), bbox_coder.kernel_kwargs(wh_ratio_clip))
