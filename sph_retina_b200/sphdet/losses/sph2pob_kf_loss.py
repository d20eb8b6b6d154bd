"""Sph2PobKFLoss -- sphdet/losses/sph2pob_kf_loss.py:8-26: ``Sph2PobTransfrom()(mmrotate KFLoss)``.

Smooth-L1 on the OBB centres + (1 - KFIoU) of the two Gaussians, with the reference's call convention kept as it is:
``pred_decode=target, targets_decode=pred`` (:26).  One launch of ``sphk_obb_loss`` for loss and gradients."""
from __future__ import annotations

import torch.nn as nn

from ..registry import LOSSES
from ._obb_loss import obb_loss_forward, widen_bfov_weight

_FUN = {'none': 0, 'ln': 1, 'exp': 2}


@LOSSES.register_module()
class Sph2PobKFLoss(nn.Module):
    """pred / target: spherical boxes [n, 4|5] in degrees; weight: None, [n] or [n, box_version]."""

    def __init__(self, fun='none', reduction='mean', loss_weight=1.0, **kwargs):
        super().__init__()
        assert reduction in ['none', 'sum', 'mean']
        assert fun in ['none', 'ln', 'exp']
        self.fun = fun
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, beta=1.0 / 9.0, eps=1e-6):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        weight = widen_bfov_weight(weight, target.size(-1))
        if weight is not None and weight.dim() > 1:
            assert weight.shape == (pred.size(0), 5)
            weight = weight.mean(-1)
        cfg = dict(fun=_FUN[self.fun], beta=float(beta), eps=float(eps))
        return obb_loss_forward('kfiou', cfg, 1, pred, target, weight, avg_factor, reduction, self.loss_weight)
