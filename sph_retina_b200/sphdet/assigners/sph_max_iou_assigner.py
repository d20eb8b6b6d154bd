"""``SphMaxIoUAssigner`` -- ``MaxIoUAssigner`` for spherical boxes without the K x N overlaps matrix
(the "next" row 1 of SURVEY.md 8f).

Same constructor, ``assign`` signature, thresholds and result as
``mmdet/core/bbox/assigners/max_iou_assigner.py:47-220``.  The reference materialises
``overlaps = iou_calculator(gt_bboxes, bboxes)`` ([K, N] fp32) and then runs ``max(dim=0)``,
``max(dim=1)`` and a Python loop over the K ground truths with one N-wide equality scan each.
Here:

  pass 1  ``sphk_iou_pairwise`` with fused row/column max+argmax     -> per-GT and per-anchor (max, argmax)
  pass 2  ``sphk_iou_pairwise_ties`` (only for ``gt_max_assign_all``) -> per anchor, the last GT whose
          row maximum it ties (the result of the reference's ascending ``for i in range(num_gts)`` loop)

plus a handful of N-long elementwise torch ops for the thresholds.  The matrix is never written.
Corner cases that need the matrix semantics exactly (an ignore region, or a GT whose best overlap is
exactly 0 while ``min_pos_iou <= 0`` -- mmdet then assigns *every* zero-overlap anchor to it) take the
matrix path, still on the GPU through the same calculator."""
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


def _result(num_gts, gt_inds, max_overlaps, labels):
    try:
        from mmdet.core.bbox.assigners.assign_result import AssignResult as MM
        return MM(num_gts, gt_inds, max_overlaps, labels=labels)
    except Exception:
        return AssignResult(num_gts, gt_inds, max_overlaps, labels)


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
        if fused_ok and gt_bboxes.size(0) > 0 and bboxes.size(0) > 0:
            res = self._assign_fused(bboxes, gt_bboxes, gt_labels)
            if res is not None:
                return res
        overlaps = calc(gt_bboxes, bboxes)
        if has_ignore:
            if self.ignore_wrt_candidates:
                ignore_max = calc(bboxes, gt_bboxes_ignore, mode='iof').max(dim=1)[0]
            else:
                ignore_max = calc(gt_bboxes_ignore, bboxes, mode='iof').max(dim=0)[0]
            overlaps[:, ignore_max > self.ignore_iof_thr] = -1
        return self.assign_wrt_overlaps(overlaps, gt_labels)

    # ------------------------------------------------------------------------------------------
    def _thresholds(self, assigned, max_overlaps, argmax_overlaps):
        """Steps 2 and 3 of max_iou_assigner.py:178-190."""
        if isinstance(self.neg_iou_thr, float):
            assigned[(max_overlaps >= 0) & (max_overlaps < self.neg_iou_thr)] = 0
        elif isinstance(self.neg_iou_thr, tuple):
            assert len(self.neg_iou_thr) == 2
            assigned[(max_overlaps >= self.neg_iou_thr[0]) & (max_overlaps < self.neg_iou_thr[1])] = 0
        pos = max_overlaps >= self.pos_iou_thr
        assigned[pos] = argmax_overlaps[pos] + 1

    @staticmethod
    def _labels(assigned, gt_labels):
        if gt_labels is None:
            return None
        labels = assigned.new_full((assigned.numel(),), -1)
        pos = assigned > 0
        labels[pos] = gt_labels[assigned[pos] - 1]
        return labels

    def _assign_fused(self, bboxes, gt_bboxes, gt_labels):
        calc = self.iou_calculator
        kind = _KINDS[calc.backend]
        gts, boxes = gt_bboxes[..., :calc.box_version], bboxes[..., :calc.box_version]
        with torch.no_grad():
            _, (gt_max, gt_arg), (max_overlaps, argmax) = _native.iou_pairwise(
                kind, gts, boxes, want_matrix=False, want_row_max=True, want_col_max=True)
            assigned = torch.full((boxes.size(0),), -1, dtype=torch.long, device=boxes.device)
            self._thresholds(assigned, max_overlaps, argmax.long())
            if self.match_low_quality:
                valid = gt_max >= self.min_pos_iou
                if self.gt_max_assign_all:
                    # a GT whose best overlap is exactly 0 ties with every zero entry of its row: matrix semantics
                    if bool((valid & (gt_max <= 0)).any()):
                        return None
                    target = torch.where(valid, gt_max, torch.full_like(gt_max, -1.0))
                    tie = _native.iou_pairwise_ties(kind, gts, boxes, target).long()
                    assigned = torch.where(tie > 0, tie, assigned)
                else:
                    # assigned[gt_argmax[i]] = i + 1 for ascending i: the largest i wins on duplicates
                    idx = torch.arange(1, gts.size(0) + 1, device=boxes.device)
                    idx = torch.where(valid, idx, torch.zeros_like(idx))
                    last = torch.zeros(boxes.size(0), dtype=torch.long, device=boxes.device)
                    last.scatter_reduce_(0, gt_arg.long(), idx, reduce='amax', include_self=True)
                    assigned = torch.where(last > 0, last, assigned)
        return _result(gts.size(0), assigned, max_overlaps, self._labels(assigned, gt_labels))

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
        self._thresholds(assigned, max_overlaps, argmax_overlaps)
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
