"""Shared host side of the losses that sit on the Sph2Pob OBBs (Sph2PobGDLoss, Sph2PobKFLoss, Sph2PobL1Loss): the
``Sph2PobTransfrom`` decorator of the reference (sphdet/losses/sph2pob_transform.py:11-37) and the weight / reduction
contract of mmdet's ``weighted_loss`` (mmdet/models/losses/utils.py), on top of ONE kernel launch per call
(``sphk_obb_loss``: jitter -> transform -> jitter -> row loss -> full backward)."""
from __future__ import annotations

import torch

from ... import _native


class _ObbLossReduced(torch.autograd.Function):
    """scale * sum(weight * loss) and both gradients from one launch (reductions 'mean' / 'sum')."""

    @staticmethod
    def forward(ctx, pred, target, weight, scale, kind, cfg):
        need_p, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        # the kernel writes the finished scalar: the forward of a training step is this one launch, no reduction op
        total, gp, gt = _native.obb_loss_total(kind, pred.detach(), target.detach(), None if weight is None else weight.detach(),
                                               scale, want_grad_pred=need_p, want_grad_target=need_t, **cfg)
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
        return (None if gp is None else (gp * g).to(ctx.in_dtypes[0]),
                None if gt is None else (gt * g).to(ctx.in_dtypes[1]), None, None, None, None)


class _ObbLossElementwise(torch.autograd.Function):
    """Unweighted elementwise loss [n] (GD / KF) or [n, 5] (L1); the backward is a second launch of the same kernel
    with the incoming gradient as its upstream (nothing but the inputs is kept)."""

    @staticmethod
    def forward(ctx, pred, target, kind, cfg):
        loss, _, _, _ = _native.obb_loss(kind, pred.detach(), target.detach(), None, 1.0, want_loss=True, **cfg)
        ctx.save_for_backward(pred.detach(), target.detach())
        ctx.kind, ctx.cfg = kind, cfg
        ctx.in_dtypes = (pred.dtype, target.dtype)
        return loss.to(pred.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        pred, target = ctx.saved_tensors
        need_p, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        _, _, gp, gt = _native.obb_loss(ctx.kind, pred, target, grad_out.float(), 1.0, want_loss=False, want_grad_pred=need_p,
                                        want_grad_target=need_t, **ctx.cfg)
        return (None if gp is None else gp.to(ctx.in_dtypes[0]), None if gt is None else gt.to(ctx.in_dtypes[1]), None, None)


def _weight_reduce_loss(loss, weight=None, reduction='mean', avg_factor=None):
    """mmdet/models/losses/utils.py: weight_reduce_loss."""
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


def widen_bfov_weight(weight, box_version):
    """sph2pob_transform.py:32-34: a 2-D BFoV weight gains the column the OBB angle uses (the row mean)."""
    if weight is not None and weight.dim() > 1 and box_version == 4:
        weight = torch.cat([weight, weight.mean(-1, keepdim=True)], dim=-1)
    return weight


def obb_loss_forward(kind, cfg, columns, pred, target, weight, avg_factor, reduction, loss_weight):
    """``loss_weight * weighted_loss(row loss)(obbs(pred), obbs(target), weight, reduction, avg_factor)``.

    ``weight`` must already have the shape the reference's loss would multiply with: [n] for the one-column losses,
    [n, 5] for L1 (anything else goes through torch broadcasting on the elementwise path and fails where the
    reference fails).  All-zero weights need no special case: the kernel skips those rows, which gives the exact zero
    (with zero gradients) of the reference's early-outs without a host sync."""
    n = pred.size(0)
    fused = (reduction in ('mean', 'sum') and n > 0 and not isinstance(avg_factor, torch.Tensor)
             and (weight is None or tuple(weight.shape) == ((n,) if columns == 1 else (n, columns))))
    if fused:
        if avg_factor is None:
            scale = loss_weight / (n * columns) if reduction == 'mean' else loss_weight
        elif reduction == 'mean':
            scale = loss_weight / (avg_factor + torch.finfo(torch.float32).eps)
        else:
            raise ValueError('avg_factor can not be used with reduction="sum"')
        return _ObbLossReduced.apply(pred, target, weight, float(scale), kind, cfg)
    if n == 0:
        # no rows: a zero that still hangs off `pred` (mmdet's l1_loss does exactly this; the mean of an empty tensor the
        # GD / KF classes would produce is NaN and is not reproduced), or the empty elementwise tensor
        zero = pred.sum() * 0
        return zero if reduction != 'none' else zero + pred.new_zeros((0, columns) if columns > 1 else (0,))
    loss = _ObbLossElementwise.apply(pred, target, kind, cfg)
    return loss_weight * _weight_reduce_loss(loss, weight, reduction, avg_factor)
