from .sph2pob_gd_loss import Sph2PobGDLoss
from .sph2pob_iou_loss import (OBBIoULoss, Sph2PobDecodedIoULoss, Sph2PobIoULoss, SphIoULoss, SphIoULossLegacy, rotated_iou, sph2pob_iou,
                               sph2pob_obbs)
from .sph2pob_kf_loss import Sph2PobKFLoss
from .sph2pob_l1_loss import Sph2PobL1Loss

__all__ = ['Sph2PobIoULoss', 'Sph2PobDecodedIoULoss', 'SphIoULoss', 'SphIoULossLegacy', 'OBBIoULoss', 'Sph2PobGDLoss', 'Sph2PobKFLoss',
           'Sph2PobL1Loss', 'sph2pob_iou', 'sph2pob_obbs', 'rotated_iou']


# sphdet/losses/__init__.py:1 -- ``from mmdet.models.losses import L1Loss as SphL1Loss``: mmdet's plain L1 on the encoded
# deltas, no spherical arithmetic in it.  With mmdet installed the alias is mmdet's class; without it, the same few torch
# expressions (mmdet/models/losses/smooth_l1_loss.py:36-53,103-146) so that the import keeps resolving.
try:
    from mmdet.models.losses import L1Loss as SphL1Loss
except Exception:
    import torch as _torch

    from .sph2pob_iou_loss import _weight_reduce_loss

    class SphL1Loss(_torch.nn.Module):
        def __init__(self, reduction='mean', loss_weight=1.0):
            super().__init__()
            self.reduction = reduction
            self.loss_weight = loss_weight

        def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
            assert reduction_override in (None, 'none', 'mean', 'sum')
            reduction = reduction_override if reduction_override else self.reduction
            if target.numel() == 0:
                loss = pred.sum() * 0
            else:
                assert pred.size() == target.size()
                loss = _torch.abs(pred - target)
            return self.loss_weight * _weight_reduce_loss(loss, weight, reduction, avg_factor)

__all__.append('SphL1Loss')
