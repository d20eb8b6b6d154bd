"""TEST INFRASTRUCTURE ONLY -- restatement (torch, dtype-generic) of the two mmrotate loss classes the reference
subclasses for its "other" Sph2Pob losses:

    sphdet/losses/sph2pob_gd_loss.py:2,9    class Sph2PobGDLoss(GDLoss)      from mmrotate.models.losses import GDLoss
    sphdet/losses/sph2pob_kf_loss.py:2,10   class Sph2PobKFLoss(KFLoss)      from mmrotate.models.losses import KFLoss
    sphdet/losses/sph2pob_iou_loss.py:8,201 class SphIoULossLegacy(RotatedIoULoss)  from mmrotate.models.losses import RotatedIoULoss

PARITY UNPINNED at the mmrotate boundary: mmrotate (pinned to 0.3.2 by the reference's README.md:95,102) is a
third-party package that is neither vendored in the reference tree nor installable offline.  What follows restates the
published algorithm of mmrotate 0.3.2 --

    mmrotate/models/losses/gaussian_dist_loss.py   xy_wh_r_2_xy_sigma, postprocess, gwd_loss, kld_loss, jd_loss,
                                                   kld_symmax_loss, kld_symmin_loss (all @weighted_loss), class GDLoss
    mmrotate/models/losses/kf_iou_loss.py          xy_wh_r_2_xy_sigma, kfiou_loss (@weighted_loss), class KFLoss

-- every clamp, default and early-out included, and parity is anchored on the reference's own call sites: the
``Sph2PobTransfrom`` decorator (sphdet/losses/sph2pob_transform.py:11-37, present and pinned), the swapped decoded boxes of
``Sph2PobKFLoss.forward`` (sph2pob_kf_loss.py:26) and the calls of the reference's tests
(tests/test_sph_iou_loss.py:112-135: ``Sph2PobGDLoss(loss_type='kld', reduction='none')(pt, gt)`` + backward).
``oracle/ref_harness.py`` plugs these classes in as ``mmrotate.models.losses.{GDLoss,KFLoss}`` so that the golden
vectors go through the reference's real subclasses and decorator.

Nothing in the product package imports this file.
"""
from __future__ import annotations

from copy import deepcopy

import torch
from torch import nn


# mmdet/models/losses/utils.py (vendored in the reference: reduce_loss, weight_reduce_loss, weighted_loss)
def weight_reduce_loss(loss, weight=None, reduction="mean", avg_factor=None):
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        if reduction == "mean":
            return loss.mean()
        if reduction == "sum":
            return loss.sum()
        return loss
    if reduction == "mean":
        eps = torch.finfo(torch.float32).eps
        return loss.sum() / (avg_factor + eps)
    if reduction != "none":
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss


def weighted_loss(loss_func):
    def wrapper(pred, target, weight=None, reduction="mean", avg_factor=None, **kwargs):
        loss = loss_func(pred, target, **kwargs)
        return weight_reduce_loss(loss, weight, reduction, avg_factor)
    wrapper.__name__ = loss_func.__name__
    return wrapper


# ---- gaussian_dist_loss.py -----------------------------------------------------------------------------------------
def xy_wh_r_2_xy_sigma(xywhr):
    """OBB (x, y, w, h, r) -> (centre, covariance R diag(w/2, h/2)^2 R^T); w, h clamped to [1e-7, 1e7]."""
    _shape = xywhr.shape
    assert _shape[-1] == 5
    xy = xywhr[..., :2]
    wh = xywhr[..., 2:4].clamp(min=1e-7, max=1e7).reshape(-1, 2)
    r = xywhr[..., 4]
    cos_r = torch.cos(r)
    sin_r = torch.sin(r)
    R = torch.stack((cos_r, -sin_r, sin_r, cos_r), dim=-1).reshape(-1, 2, 2)
    S = 0.5 * torch.diag_embed(wh)
    sigma = R.bmm(S.square()).bmm(R.permute(0, 2, 1)).reshape(_shape[:-1] + (2, 2))
    return xy, sigma


def postprocess(distance, fun="log1p", tau=1.0):
    if fun == "log1p":
        distance = torch.log1p(distance)
    elif fun == "sqrt":
        distance = torch.sqrt(distance.clamp(1e-7))
    elif fun == "none":
        pass
    else:
        raise ValueError(f"Invalid non-linear function {fun}")
    if tau >= 1.0:
        return 1 - 1 / (tau + distance)
    return distance


def _det2(S):
    return S.det()


def _gwd(pred, target, fun="log1p", tau=1.0, alpha=1.0, normalize=True):
    xy_p, Sigma_p = pred
    xy_t, Sigma_t = target
    xy_distance = (xy_p - xy_t).square().sum(dim=-1)
    whr_distance = Sigma_p.diagonal(dim1=-2, dim2=-1).sum(dim=-1)
    whr_distance = whr_distance + Sigma_t.diagonal(dim1=-2, dim2=-1).sum(dim=-1)
    _t_tr = (Sigma_p.bmm(Sigma_t)).diagonal(dim1=-2, dim2=-1).sum(dim=-1)
    _t_det_sqrt = (_det2(Sigma_p) * _det2(Sigma_t)).clamp(1e-7).sqrt()
    whr_distance = whr_distance + (-2) * ((_t_tr + 2 * _t_det_sqrt).clamp(1e-7).sqrt())
    distance = (xy_distance + alpha * alpha * whr_distance).clamp(1e-7).sqrt()
    if normalize:
        scale = 2 * (_t_det_sqrt.clamp(1e-7).sqrt().clamp(1e-7).sqrt()).clamp(1e-7)
        distance = distance / scale
    return postprocess(distance, fun=fun, tau=tau)


def _kld(pred, target, fun="log1p", tau=1.0, alpha=1.0, sqrt=True):
    xy_p, Sigma_p = pred
    xy_t, Sigma_t = target
    _shape = xy_p.shape
    xy_p = xy_p.reshape(-1, 2)
    xy_t = xy_t.reshape(-1, 2)
    Sigma_p = Sigma_p.reshape(-1, 2, 2)
    Sigma_t = Sigma_t.reshape(-1, 2, 2)
    Sigma_p_inv = torch.stack((Sigma_p[..., 1, 1], -Sigma_p[..., 0, 1], -Sigma_p[..., 1, 0], Sigma_p[..., 0, 0]),
                              dim=-1).reshape(-1, 2, 2)
    Sigma_p_inv = Sigma_p_inv / _det2(Sigma_p).unsqueeze(-1).unsqueeze(-1)
    dxy = (xy_p - xy_t).unsqueeze(-1)
    xy_distance = 0.5 * dxy.permute(0, 2, 1).bmm(Sigma_p_inv).bmm(dxy).view(-1)
    whr_distance = 0.5 * Sigma_p_inv.bmm(Sigma_t).diagonal(dim1=-2, dim2=-1).sum(dim=-1)
    Sigma_p_det_log = _det2(Sigma_p).log()
    Sigma_t_det_log = _det2(Sigma_t).log()
    whr_distance = whr_distance + 0.5 * (Sigma_p_det_log - Sigma_t_det_log)
    whr_distance = whr_distance - 1
    distance = (xy_distance / (alpha * alpha) + whr_distance)
    if sqrt:
        distance = distance.clamp(1e-7).sqrt()
    distance = distance.reshape(_shape[:-1])
    return postprocess(distance, fun=fun, tau=tau)


gwd_loss = weighted_loss(_gwd)
kld_loss = weighted_loss(_kld)


@weighted_loss
def jd_loss(pred, target, fun="log1p", tau=1.0, alpha=1.0, sqrt=True):
    jd = kld_loss(pred, target, fun="none", tau=0, alpha=alpha, sqrt=False, reduction="none")
    jd = jd + kld_loss(target, pred, fun="none", tau=0, alpha=alpha, sqrt=False, reduction="none")
    jd = jd * 0.5
    if sqrt:
        jd = jd.clamp(1e-7).sqrt()
    return postprocess(jd, fun=fun, tau=tau)


@weighted_loss
def kld_symmax_loss(pred, target, fun="log1p", tau=1.0, alpha=1.0, sqrt=True):
    kld_pt = kld_loss(pred, target, fun="none", tau=0, alpha=alpha, sqrt=sqrt, reduction="none")
    kld_tp = kld_loss(target, pred, fun="none", tau=0, alpha=alpha, sqrt=sqrt, reduction="none")
    return postprocess(torch.max(kld_pt, kld_tp), fun=fun, tau=tau)


@weighted_loss
def kld_symmin_loss(pred, target, fun="log1p", tau=1.0, alpha=1.0, sqrt=True):
    kld_pt = kld_loss(pred, target, fun="none", tau=0, alpha=alpha, sqrt=sqrt, reduction="none")
    kld_tp = kld_loss(target, pred, fun="none", tau=0, alpha=alpha, sqrt=sqrt, reduction="none")
    return postprocess(torch.min(kld_pt, kld_tp), fun=fun, tau=tau)


class GDLoss(nn.Module):
    BAG_GD_LOSS = {"gwd": gwd_loss, "kld": kld_loss, "jd": jd_loss, "kld_symmax": kld_symmax_loss,
                   "kld_symmin": kld_symmin_loss}
    BAG_PREP = {"xy_wh_r": xy_wh_r_2_xy_sigma}

    def __init__(self, loss_type, representation="xy_wh_r", fun="log1p", tau=0.0, alpha=1.0, reduction="mean",
                 loss_weight=1.0, **kwargs):
        super().__init__()
        assert reduction in ["none", "sum", "mean"]
        assert fun in ["log1p", "none", "sqrt"]
        assert loss_type in self.BAG_GD_LOSS
        self.loss = self.BAG_GD_LOSS[loss_type]
        self.preprocess = self.BAG_PREP[representation]
        self.fun = fun
        self.tau = tau
        self.alpha = alpha
        self.reduction = reduction
        self.loss_weight = loss_weight
        self.kwargs = kwargs

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        assert reduction_override in (None, "none", "mean", "sum")
        reduction = reduction_override if reduction_override else self.reduction
        if (weight is not None) and (not torch.any(weight > 0)) and (reduction != "none"):
            return (pred * weight).sum()
        if weight is not None and weight.dim() > 1:
            assert weight.shape == pred.shape
            weight = weight.mean(-1)
        _kwargs = deepcopy(self.kwargs)
        _kwargs.update(kwargs)
        pred = self.preprocess(pred)
        target = self.preprocess(target)
        return self.loss(pred, target, fun=self.fun, tau=self.tau, alpha=self.alpha, weight=weight, avg_factor=avg_factor,
                         reduction=reduction, **_kwargs) * self.loss_weight


# ---- kf_iou_loss.py --------------------------------------------------------------------------------------------------
@weighted_loss
def kfiou_loss(pred, target, pred_decode=None, targets_decode=None, fun=None, beta=1.0 / 9.0, eps=1e-6):
    xy_p = pred[:, :2]
    xy_t = target[:, :2]
    _, Sigma_p = xy_wh_r_2_xy_sigma(pred_decode)
    _, Sigma_t = xy_wh_r_2_xy_sigma(targets_decode)
    diff = torch.abs(xy_p - xy_t)
    xy_loss = torch.where(diff < beta, 0.5 * diff * diff / beta, diff - 0.5 * beta).sum(dim=-1)
    Vb_p = 4 * _det2(Sigma_p).sqrt()
    Vb_t = 4 * _det2(Sigma_t).sqrt()
    K = Sigma_p.bmm((Sigma_p + Sigma_t).inverse())
    Sigma = Sigma_p - K.bmm(Sigma_p)
    Vb = 4 * _det2(Sigma).sqrt()
    Vb = torch.where(torch.isnan(Vb), torch.full_like(Vb, 0), Vb)
    KFIoU = Vb / (Vb_p + Vb_t - Vb + eps)
    if fun == "ln":
        kf_loss = -torch.log(KFIoU + eps)
    elif fun == "exp":
        kf_loss = torch.exp(1 - KFIoU) - 1
    else:
        kf_loss = 1 - KFIoU
    return (xy_loss + kf_loss).clamp(0)


class KFLoss(nn.Module):
    def __init__(self, fun="none", reduction="mean", loss_weight=1.0, **kwargs):
        super().__init__()
        assert reduction in ["none", "sum", "mean"]
        assert fun in ["none", "ln", "exp"]
        self.fun = fun
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, pred_decode=None, targets_decode=None,
                reduction_override=None, **kwargs):
        assert reduction_override in (None, "none", "mean", "sum")
        reduction = reduction_override if reduction_override else self.reduction
        if (weight is not None) and (not torch.any(weight > 0)) and (reduction != "none"):
            return (pred * weight).sum()
        if weight is not None and weight.dim() > 1:
            assert weight.shape == pred.shape
            weight = weight.mean(-1)
        return kfiou_loss(pred, target, fun=self.fun, weight=weight, avg_factor=avg_factor, pred_decode=pred_decode,
                          targets_decode=targets_decode, reduction=reduction, **kwargs) * self.loss_weight


# ---- rotated_iou_loss.py (mmrotate 0.3.2) ---------------------------------------------------------------------------
# The base class of the reference's SphIoULossLegacy (sphdet/losses/sph2pob_iou_loss.py:199-216:
# ``Sph2PobTransfrom()(RotatedIoULoss)``).  ``iou_fn``: mmcv.ops.diff_iou_rotated_2d at run time (the harness installs the
# reference's vendored copy under that name).
@weighted_loss
def rotated_iou_loss(pred, target, linear=False, mode="log", eps=1e-6):
    assert mode in ["linear", "square", "log"]
    if linear:
        mode = "linear"
    from mmcv.ops import diff_iou_rotated_2d
    ious = diff_iou_rotated_2d(pred.unsqueeze(0), target.unsqueeze(0))
    ious = ious.squeeze(0).clamp(min=eps)
    if mode == "linear":
        return 1 - ious
    if mode == "square":
        return 1 - ious ** 2
    return -ious.log()


class RotatedIoULoss(nn.Module):
    def __init__(self, linear=False, eps=1e-6, reduction="mean", loss_weight=1.0, mode="log"):
        super().__init__()
        assert mode in ("linear", "square", "log")
        if linear:
            mode = "linear"
        self.mode = mode
        self.linear = linear
        self.eps = eps
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        assert reduction_override in (None, "none", "mean", "sum")
        reduction = reduction_override if reduction_override else self.reduction
        if (weight is not None) and (not torch.any(weight > 0)) and (reduction != "none"):
            return (pred * weight).sum()
        if weight is not None and weight.dim() > 1:
            assert weight.shape == pred.shape
            weight = weight.mean(-1)
        return self.loss_weight * rotated_iou_loss(pred, target, weight, mode=self.mode, eps=self.eps, reduction=reduction,
                                                   avg_factor=avg_factor, **kwargs)
