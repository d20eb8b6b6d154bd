"""Training targets of the spherical anchor heads for a whole batch (the step between the assigner and
``loss_single``): ``AnchorHead.get_targets`` / ``_get_targets_single`` (mmdet/models/dense_heads/anchor_head.py:202-299,
301-399) with the ``PseudoSampler`` RetinaNet uses (mmdet/core/bbox/samplers/pseudo_sampler.py) and
``allowed_border = -1`` (every anchor is valid, as in the reference's configs).

The reference runs this image by image: assigner (K x N matrix, three reductions, a Python loop over the GTs), sampler
(two ``nonzero`` with host syncs), four scatter writes, ``bbox_coder.encode`` of the positives.  Here the whole batch is
two C-ABI calls -- ``sphk_max_iou_assign`` and ``sphk_anchor_targets`` -- and the only host synchronisation is reading
the per-image positive / negative counts (mmdet reads ``pos_inds.numel()`` at the same place)."""
from __future__ import annotations

import torch

from .... import _native
from ...assigners import SphMaxIoUAssigner


def get_targets_batch(anchors, gt_bboxes_list, gt_labels_list, assigner, num_classes, bbox_coder=None, reg_decoded_bbox=True,
                      pos_weight=-1, sync_counts=True):
    """Targets of all images of a step that share the anchor list.

    anchors [N, box_version]; gt_bboxes_list / gt_labels_list: one tensor per image (gt_labels_list may be None: RPN);
    assigner: a ``SphMaxIoUAssigner`` on a Sph2Pob calculator; bbox_coder: needed when ``reg_decoded_bbox`` is False.
    Returns ``(labels [B, N] int64, label_weights [B, N], bbox_targets [B, N, D], bbox_weights [B, N, D],
    num_total_pos, num_total_neg)`` where the totals follow anchor_head.py:379-380 (``sum(max(count, 1))`` over the
    images) as python ints, or, with ``sync_counts=False``, the raw ``counts [B, 2]`` device tensor twice (no sync)."""
    assert isinstance(assigner, SphMaxIoUAssigner)
    bv = assigner.iou_calculator.box_version
    boxes = anchors[..., :bv]
    counts = [int(g.size(0)) for g in gt_bboxes_list]
    offsets = [0]
    for c in counts:
        offsets.append(offsets[-1] + c)
    nonempty = [g[..., :bv] for g in gt_bboxes_list if g.size(0) > 0]
    gts = torch.cat(nonempty) if len(nonempty) > 1 else (nonempty[0] if nonempty else None)
    labels_cat = None
    if gt_labels_list is not None and offsets[-1] > 0:
        ll = [l for l, c in zip(gt_labels_list, counts) if c > 0]
        labels_cat = torch.cat(ll) if len(ll) > 1 else ll[0]
    gt_inds, _, _ = assigner._assign_batch_raw(boxes, gts, offsets)
    means = stds = None
    if not reg_decoded_bbox:
        assert bbox_coder is not None, "reg_decoded_bbox=False needs the bbox_coder (targets are its encode())"
        means, stds = bbox_coder.means, bbox_coder.stds
    with torch.no_grad():
        labels, label_weights, bbox_targets, bbox_weights, cnt = _native.anchor_targets(
            gt_inds, boxes, gts, labels_cat, offsets, num_classes, pos_weight, reg_decoded_bbox, means, stds)
    if not sync_counts:
        return labels, label_weights, bbox_targets, bbox_weights, cnt, cnt
    host = cnt.cpu()                                    # the one host sync of the step
    num_total_pos = int(host[:, 0].clamp(min=1).sum())
    num_total_neg = int(host[:, 1].clamp(min=1).sum())
    return labels, label_weights, bbox_targets, bbox_weights, num_total_pos, num_total_neg
