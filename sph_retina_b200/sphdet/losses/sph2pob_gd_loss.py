"""Sph2PobGDLoss -- sphdet/losses/sph2pob_gd_loss.py:7-26: ``Sph2PobTransfrom()(mmrotate GDLoss)``.

Spherical boxes -> planar OBBs (jitter, sph2pob_standard, jitter) -> 2-D Gaussians -> GWD / KLD / JD / symmetric KLD
distance -> fun / tau post-processing -> weighted reduction.  Constructor and forward arguments are those of mmrotate
0.3.2's ``GDLoss`` (the reference's README pins that version); the whole chain, gradients included, is one launch of
``sphk_obb_loss``."""
from __future__ import annotations

from copy import deepcopy

import torch.nn as nn

from ..registry import LOSSES
from ._obb_loss import obb_loss_forward, widen_bfov_weight

_FUN = {'none': 0, 'log1p': 1, 'sqrt': 2}
# keyword each distance accepts beyond fun / tau / alpha (gaussian_dist_loss.py: gwd_loss(normalize), kld family (sqrt))
_OPTION = {'gwd': 'normalize', 'kld': 'sqrt', 'jd': 'sqrt', 'kld_symmax': 'sqrt', 'kld_symmin': 'sqrt'}


@LOSSES.register_module()
class Sph2PobGDLoss(nn.Module):
    """pred / target: spherical boxes [n, 4|5] in degrees; weight: None, [n] or [n, box_version]."""

    BAG_GD_LOSS = tuple(_OPTION)

    def __init__(self, loss_type, representation='xy_wh_r', fun='log1p', tau=0.0, alpha=1.0, reduction='mean',
                 loss_weight=1.0, **kwargs):
        super().__init__()
        assert reduction in ['none', 'sum', 'mean']
        assert fun in ['log1p', 'none', 'sqrt']
        assert loss_type in self.BAG_GD_LOSS
        if representation != 'xy_wh_r':
            raise NotImplementedError("Sph2PobGDLoss: the Sph2Pob transform yields (x, y, w, h, r) boxes; representation "
                                      "%r does not apply" % (representation,))
        self.loss_type = loss_type
        self.fun = fun
        self.tau = tau
        self.alpha = alpha
        self.reduction = reduction
        self.loss_weight = loss_weight
        self.kwargs = kwargs

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        weight = widen_bfov_weight(weight, target.size(-1))
        if weight is not None and weight.dim() > 1:
            assert weight.shape == (pred.size(0), 5)         # GDLoss.forward: weight.shape == (OBB) pred.shape
            weight = weight.mean(-1)
        _kwargs = deepcopy(self.kwargs)
        _kwargs.update(kwargs)
        option = _kwargs.pop(_OPTION[self.loss_type], True)
        if _kwargs:      # the distance functions take no other keyword: same TypeError as calling them would raise
            raise TypeError("%s_loss() got an unexpected keyword argument %r" % (self.loss_type, sorted(_kwargs)[0]))
        cfg = dict(fun=_FUN[self.fun], flags=int(bool(option)), tau=float(self.tau), alpha=float(self.alpha))
        return obb_loss_forward(self.loss_type, cfg, 1, pred, target, weight, avg_factor, reduction, self.loss_weight)
