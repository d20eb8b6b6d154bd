"""Sph2PobL1Loss -- sphdet/losses/sph2pob_l1_loss.py:9-94: ``Sph2PobTransfrom()(mmdet L1Loss)`` on the OBB deltas.

``encode=True``: loss = |bbox2delta(pred OBB, target OBB)| (``swap`` exchanges the roles), five columns per row
(dx, dy, dw, dh, da); ``encode=False``: |pred OBB - target OBB|.  (The reference's constructor stops in
``pdb.set_trace()`` at :25; that line is not reproduced.)"""
from __future__ import annotations

import torch.nn as nn

from ..registry import LOSSES
from ._obb_loss import obb_loss_forward, widen_bfov_weight


@LOSSES.register_module()
class Sph2PobL1Loss(nn.Module):
    """pred / target: spherical boxes [n, 4|5] in degrees; weight: None or [n, box_version] (one weight per loss column:
    a BFoV weight is widened by its row mean for the angle column)."""

    def __init__(self, encode=True, swap=False, angle_modifier='original', reduction='mean', loss_weight=1.0):
        assert angle_modifier in ['original', 'modulus']
        super().__init__()
        self.encode = encode
        self.swap = swap
        self.angle_modifier = angle_modifier
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        weight = widen_bfov_weight(weight, target.size(-1))
        flags = (1 if self.encode else 0) | (2 if self.swap else 0) | (4 if self.angle_modifier == 'modulus' else 0)
        return obb_loss_forward('l1', dict(flags=flags), 5, pred, target, weight, avg_factor, reduction, self.loss_weight)
