"""``SphMaxIoUAssigner`` -- ``MaxIoUAssigner`` for spherical boxes without the K x N overlaps matrix
(the "next" row 1 of SURVEY.md 8f).

Same constructor, ``assign`` signature, thresholds and result as
``mmdet/core/bbox/assigners/max_iou_assigner.py:47-220``.  The reference materialises
``overlaps = iou_calculator(gt_bboxes, bboxes)`` ([K, N] fp32) and then runs ``max(dim=0)``,
``max(dim=1)`` and a Python loop over the K ground truths with one N-wide equality scan each.
Here:

  pass 1  the N x M kernel with fused row/column max+argmax          -> per-GT and per-anchor (max, argmax)
  pass 2  the same kernel in tie mode (only for ``gt_max_assign_all``) -> per anchor, the last GT whose
          row maximum it ties (the result of the reference's ascending ``for i in range(num_gts)`` loop)
  epilogue kernel: thresholds, low-quality override, labels

all inside ONE C-ABI call (``sphk_max_iou_assign``), for one image (``assign``) or for all images of a step
that share the anchor list (``assign_batch``).  The matrix is never written.
No host synchronisation happens anywhere in ``assign``.  An ignore region (``ignore_iof_thr > 0`` with
``gt_bboxes_ignore``) takes the matrix path, still on the GPU through the same calculator.  The mmdet quirk
that a GT whose best overlap is exactly 0 (with ``min_pos_iou <= 0``) grabs *every* anchor is reproduced."""
from __future__ import annotations

import torch

from ... import _native
from ..iou.sph_iou_calculator import SphOverlaps2D
from ..registry import build_iou_calculator

_KINDS = {'sph2pob_standard_iou': 'sph2pob_standard', 'sph2pob_efficient_iou': 'sph2pob_efficient'}


class AssignResult:
    """Minimal stand-in for mmdet's AssignResult (same attribute names) when mmdet is not importable."""

    def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
        self.num_gts, self.gt_inds, self.max_overlaps, self.labels = num_gts, gt_inds, max_overlaps, labels

    @property
    def num_preds(self):
        return len(self.gt_inds)


try:
    from mmdet.core.bbox.assigners.assign_result import AssignResult as _ResultClass
except Exception:
    _ResultClass = AssignResult


def _result(num_gts, gt_inds, max_overlaps, labels):
    return _ResultClass(num_gts, gt_inds, max_overlaps, labels=labels)


class SphMaxIoUAssigner:
    def __init__(self, pos_iou_thr, neg_iou_thr, min_pos_iou=.0, gt_max_assign_all=True, ignore_iof_thr=-1,
                 ignore_wrt_candidates=True, match_low_quality=True, gpu_assign_thr=-1,
                 iou_calculator=dict(type='SphOverlaps2D', backend='sph2pob_efficient_iou', box_version=4)):
        self.pos_iou_thr = pos_iou_thr
        self.neg_iou_thr = neg_iou_thr
        self.min_pos_iou = min_pos_iou
        self.gt_max_assign_all = gt_max_assign_all
        self.ignore_iof_thr = ignore_iof_thr
        self.ignore_wrt_candidates = ignore_wrt_candidates
        self.gpu_assign_thr = gpu_assign_thr          # accepted, ignored: there is no CPU path
        self.match_low_quality = match_low_quality
        self.iou_calculator = iou_calculator if isinstance(iou_calculator, SphOverlaps2D) else build_iou_calculator(iou_calculator)

    # ------------------------------------------------------------------------------------------
    def assign(self, bboxes, gt_bboxes, gt_bboxes_ignore=None, gt_labels=None):
        calc = self.iou_calculator
        has_ignore = (self.ignore_iof_thr > 0 and gt_bboxes_ignore is not None and gt_bboxes_ignore.numel() > 0
                      and bboxes.numel() > 0)
        fused_ok = isinstance(calc, SphOverlaps2D) and calc.backend in _KINDS and not has_ignore
        if fused_ok and bboxes.size(0) > 0:
            return self._assign_fused(bboxes, gt_bboxes, gt_labels)
        overlaps = calc(gt_bboxes, bboxes)
        if has_ignore:
            if self.ignore_wrt_candidates:
                ignore_max = calc(bboxes, gt_bboxes_ignore, mode='iof').max(dim=1)[0]
            else:
                ignore_max = calc(gt_bboxes_ignore, bboxes, mode='iof').max(dim=0)[0]
            overlaps[:, ignore_max > self.ignore_iof_thr] = -1
        return self.assign_wrt_overlaps(overlaps, gt_labels)

    # ------------------------------------------------------------------------------------------
    def _thresholds(self, max_overlaps, argmax_overlaps):
        """Steps 1-3 of max_iou_assigner.py:147-190 as select ops (no boolean-mask indexing, hence no host sync)."""
        assigned = torch.full_like(argmax_overlaps, -1)
        if isinstance(self.neg_iou_thr, float):
            neg = (max_overlaps >= 0) & (max_overlaps < self.neg_iou_thr)
        else:
            assert isinstance(self.neg_iou_thr, tuple) and len(self.neg_iou_thr) == 2
            neg = (max_overlaps >= self.neg_iou_thr[0]) & (max_overlaps < self.neg_iou_thr[1])
        assigned = torch.where(neg, torch.zeros_like(assigned), assigned)
        return torch.where(max_overlaps >= self.pos_iou_thr, argmax_overlaps + 1, assigned)

    @staticmethod
    def _labels(assigned, gt_labels):
        if gt_labels is None:
            return None
        picked = gt_labels.to(assigned.dtype)[(assigned - 1).clamp(min=0)]
        return torch.where(assigned > 0, picked, torch.full_like(assigned, -1))

    def _neg_range(self):
        if isinstance(self.neg_iou_thr, float):
            return 0.0, self.neg_iou_thr
        assert isinstance(self.neg_iou_thr, tuple) and len(self.neg_iou_thr) == 2
        return float(self.neg_iou_thr[0]), float(self.neg_iou_thr[1])

    def _assign_fused(self, bboxes, gt_bboxes, gt_labels):
        return self.assign_batch(bboxes, [gt_bboxes], None if gt_labels is None else [gt_labels])[0]

    def _assign_batch_raw(self, boxes, gts, offsets, labels=None):
        """(gt_inds [B, N], max_overlaps [B, N], labels [B, N] | None) for the concatenated GT lists `gts` with image
        offsets `offsets` (python list of B + 1 ints): the tensors of sphk_max_iou_assign before they are cut per image."""
        lo, hi = self._neg_range()
        with torch.no_grad():
            return _native.max_iou_assign(_KINDS[self.iou_calculator.backend], gts, offsets, boxes, self.pos_iou_thr, lo, hi,
                                          self.min_pos_iou, self.gt_max_assign_all, self.match_low_quality, labels)

    def assign_batch(self, bboxes, gt_bboxes_list, gt_labels_list=None):
        """All images of a step in one go when they share the anchor list (RetinaNet: SURVEY.md 3.1, the reference
        loops over the images in Python, mmdet/models/dense_heads/anchor_head.py:368-377).  One C-ABI call
        (``sphk_max_iou_assign``: 5-6 kernel launches for the whole batch), no K x N matrix, no host sync.
        Returns a list of AssignResult, one per image."""
        calc = self.iou_calculator
        assert isinstance(calc, SphOverlaps2D) and calc.backend in _KINDS, "fused assignment needs a Sph2Pob calculator"
        bv = calc.box_version
        boxes = bboxes[..., :bv]
        counts = [int(g.size(0)) for g in gt_bboxes_list]
        offsets = [0]
        for c in counts:
            offsets.append(offsets[-1] + c)
        nonempty = [g[..., :bv] for g in gt_bboxes_list if g.size(0) > 0]
        gts = torch.cat(nonempty) if len(nonempty) > 1 else (nonempty[0] if nonempty else None)
        labels = None
        if gt_labels_list is not None and offsets[-1] > 0:
            ll = [l for l, c in zip(gt_labels_list, counts) if c > 0]
            labels = torch.cat(ll) if len(ll) > 1 else ll[0]
        gt_inds, max_overlaps, out_labels = self._assign_batch_raw(boxes, gts, offsets, labels)
        res = []
        for b, k in enumerate(counts):
            lab = None
            if gt_labels_list is not None:
                lab = out_labels[b] if out_labels is not None else gt_inds.new_full((gt_inds.size(1),), -1)
            res.append(_result(k, gt_inds[b], max_overlaps[b], lab))
        return res

    # ------------------------------------------------------------------------------------------
    def assign_wrt_overlaps(self, overlaps, gt_labels=None):
        """max_iou_assigner.py:135-220 on a materialised matrix, vectorised (no Python loop over the GTs)."""
        num_gts, num_bboxes = overlaps.size(0), overlaps.size(1)
        assigned = overlaps.new_full((num_bboxes,), -1, dtype=torch.long)
        if num_gts == 0 or num_bboxes == 0:
            max_overlaps = overlaps.new_zeros((num_bboxes,))
            if num_gts == 0:
                assigned[:] = 0
            labels = None if gt_labels is None else overlaps.new_full((num_bboxes,), -1, dtype=torch.long)
            return _result(num_gts, assigned, max_overlaps, labels)
        max_overlaps, argmax_overlaps = overlaps.max(dim=0)
        gt_max, gt_arg = overlaps.max(dim=1)
        assigned = self._thresholds(max_overlaps, argmax_overlaps)
        if self.match_low_quality:
            valid = gt_max >= self.min_pos_iou
            idx = torch.arange(1, num_gts + 1, device=overlaps.device)
            if self.gt_max_assign_all:
                hit = (overlaps == gt_max[:, None]) & valid[:, None]
                last = (hit.long() * idx[:, None]).max(dim=0)[0]
            else:
                last = torch.zeros(num_bboxes, dtype=torch.long, device=overlaps.device)
                last.scatter_reduce_(0, gt_arg, torch.where(valid, idx, torch.zeros_like(idx)), reduce='amax', include_self=True)
            assigned = torch.where(last > 0, last, assigned)
        return _result(num_gts, assigned, max_overlaps, self._labels(assigned, gt_labels))


try:  # register next to mmdet's own assigners when mmdet is importable
    from mmdet.core.bbox.builder import BBOX_ASSIGNERS
    BBOX_ASSIGNERS.register_module()(SphMaxIoUAssigner)
except Exception:
    pass
