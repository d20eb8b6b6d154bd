"""Seeded synthetic inputs of the BASELINE.json configurations.

``generate_boxes`` reproduces the reference's canonical generator for dtype='float'
(tests/utils/generate_data.py:10-42: uniform theta [0,360), phi [0,180), alpha/beta in the given
ranges, gamma [-90,90), degrees, CPU generator).  ``retina_anchors`` reproduces what the RetinaNet
anchor generator + pixel->sphere conversion emit for a 512x1024 equirectangular image
(mmdet/core/anchor/anchor_generator.py:98-101,169-192; sphdet/bbox/box_formator.py:85-92)."""
from __future__ import annotations

import torch


def generate_boxes(num, theta_range=(0, 360), phi_range=(0, 180), alpha_range=(1, 180), beta_range=(1, 180),
                   gamma_range=(-90, 90), box="bfov", seed=None):
    if seed is not None:
        torch.manual_seed(seed)
    u = torch.rand((num, 5))
    rng = [theta_range, phi_range, alpha_range, beta_range, gamma_range]
    cols = [u[:, k] * (r[1] - r[0]) + r[0] for k, r in enumerate(rng)]
    return torch.stack(cols[:4] if box == "bfov" else cols, dim=1).contiguous()


def retina_anchors(height=512, width=1024, strides=(8, 16, 32, 64, 128), octave_base_scale=4, scales_per_octave=3,
                   ratios=(0.5, 1.0, 2.0), box="rbfov"):
    """[sum_l 9*(H/s_l)*(W/s_l), 5] = 98,208 anchors for the defaults: (theta, phi, alpha, beta, gamma=0)."""
    out = []
    for s in strides:
        fh, fw = height // s, width // s
        ys, xs = torch.meshgrid(torch.arange(fh) * s, torch.arange(fw) * s, indexing="ij")
        base = []
        for r in ratios:
            for i in range(scales_per_octave):
                sc = octave_base_scale * 2 ** (i / scales_per_octave)
                base.append((s * sc / r ** 0.5, s * sc * r ** 0.5))
        base = torch.tensor(base)
        n = base.size(0)
        cx = xs.reshape(-1, 1).float().expand(-1, n)
        cy = ys.reshape(-1, 1).float().expand(-1, n)
        w = base[:, 0][None, :].expand_as(cx)
        h = base[:, 1][None, :].expand_as(cx)
        cols = [cx / width * 360, cy / height * 180, w / width * 360, h / height * 180]
        if box == "rbfov":
            cols.append(torch.zeros_like(cx))
        out.append(torch.stack(cols, dim=-1).reshape(-1, len(cols)))
    return torch.cat(out).float().contiguous()


def assignment_batch(images=16, gts_per_image=32, seed0=100):
    """Config #2: 16 images x 32 random RBFoV GT (seeds 100..115) against the shared anchor grid."""
    gts = [generate_boxes(gts_per_image, alpha_range=(5, 120), beta_range=(5, 120), gamma_range=(-90, 90), box="rbfov",
                          seed=seed0 + i) for i in range(images)]
    return torch.stack(gts), retina_anchors()


def loss_pairs(n=200_000, box="rbfov", seed=0):
    """Config #3: PANDORA-style positives, pred = clamp(target + N(0, sigma), min=1)."""
    target = generate_boxes(n, alpha_range=(5, 100), beta_range=(5, 100), box=box, seed=seed)
    torch.manual_seed(seed + 1)
    sigma = torch.tensor([6.0, 6.0, 6.0, 6.0, 10.0])[:target.size(1)]
    pred = (target + torch.randn_like(target) * sigma).clamp(min=1)
    return pred.contiguous(), target


def nms_batch(images=64, per_image=1000, classes=80, box="bfov", seed=0):
    """Config #4: per image 200 seed boxes + 4 noisy copies each (so suppression happens), random scores/labels."""
    torch.manual_seed(seed)
    n_seed = per_image // 5
    D = 4 if box == "bfov" else 5
    seeds = generate_boxes(images * n_seed, alpha_range=(5, 60), beta_range=(5, 60), box=box).view(images, n_seed, D)
    boxes = (seeds.repeat(1, 5, 1) + torch.randn(images, n_seed * 5, D) * 2).clamp(min=1)
    scores = torch.rand(images * per_image)
    labels = torch.randint(0, classes, (images * per_image,))
    image_ids = torch.arange(images).repeat_interleave(per_image)
    return boxes.reshape(-1, D).contiguous(), scores, labels, image_ids


def head_loss_batch(images=16, pos_fraction=0.0125, stds=(0.1, 0.1, 0.2, 0.2, 0.1), seed=7):
    """The regression branch of the head's loss on a whole batch: every anchor of every image
    [images * 98,208, 5] with its delta prediction, target and 2-D weight (sph_retina_head.py:252-265).  ~1.25 % of the
    anchors are positives (config #2 measured 1,229 of 98,208): their target is a box near the anchor, their deltas the
    encoding of that target plus noise (a mid-training prediction, IoU ~ 0.6); negatives carry zero targets / weights."""
    torch.manual_seed(seed)
    anchors = retina_anchors().repeat(images, 1)
    n = anchors.size(0)
    pos = torch.rand(n) < pos_fraction
    a = anchors[pos]
    a = torch.cat([a[:, :2], a[:, 2:4].clamp(max=170.0), a[:, 4:]], dim=1)
    k = int(pos.sum())
    target = torch.zeros(n, 5)
    tgt = a + torch.randn(k, 5) * torch.tensor([0.15, 0.15, 0.2, 0.2, 0.0]) * torch.cat([a[:, 2:4], a[:, 2:4], a[:, :1]], 1)
    tgt[:, 4] = (torch.rand(k) - 0.5) * 90
    tgt[:, 0] = tgt[:, 0].clamp(1, 359)
    tgt[:, 1] = tgt[:, 1].clamp(1, 179)
    tgt[:, 2:4] = tgt[:, 2:4].clamp(1, 175)
    target[pos] = tgt
    std = torch.tensor(stds)
    enc = torch.stack([(tgt[:, 0] - a[:, 0]) / a[:, 2], (tgt[:, 1] - a[:, 1]) / a[:, 3], torch.log(tgt[:, 2] / a[:, 2]),
                       torch.log(tgt[:, 3] / a[:, 3]), torch.deg2rad(tgt[:, 4] - a[:, 4])], dim=1) / std
    deltas = torch.randn(n, 5) * 0.1
    deltas[pos] = enc + torch.randn(k, 5) * 0.8
    weight = pos.float()[:, None].expand(n, 5).contiguous()
    return anchors.contiguous(), deltas.contiguous(), target.contiguous(), weight
