#!/usr/bin/env python
"""Device time of the GD / KF / L1 losses on the Sph2Pob OBBs, 200 k RBFoV pairs: forward + backward eager, and the
kernel alone (forward call of the autograd function = one k_obb_loss launch + the partial sum)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sph_retina_b200 import synthetic as S, _native
from sph_retina_b200.sphdet.losses import Sph2PobGDLoss, Sph2PobKFLoss, Sph2PobL1Loss
pred, target = (x.cuda() for x in S.loss_pairs(200_000))
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def run(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ms = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return statistics.median(ms)
for name, L in (("gwd", Sph2PobGDLoss("gwd", reduction="sum")), ("kld", Sph2PobGDLoss("kld", reduction="sum")),
                ("kfiou", Sph2PobKFLoss(reduction="sum")), ("l1", Sph2PobL1Loss(reduction="sum"))):
    def fb():
        p = pred.detach().requires_grad_(True)
        L(p, target).backward()
    def fwd_only():
        with torch.no_grad():
            L(pred, target)
    p = pred.detach().requires_grad_(True)
    print("%-6s fwd+bwd eager %.1f us   forward (no grad) %.1f us" % (name, run(fb) * 1e3, run(fwd_only) * 1e3))
