"""Fused overlaps + max/argmax: what ``MaxIoUAssigner.assign_wrt_overlaps`` reduces the matrix to
(mmdet/core/bbox/assigners/max_iou_assigner.py:173-176: ``overlaps.max(dim=0)`` per anchor and
``overlaps.max(dim=1)`` per GT) without writing the K x N matrix to HBM."""
from __future__ import annotations

import torch

from ... import _native

_KINDS = {'sph2pob_standard_iou': 'sph2pob_standard', 'sph2pob_efficient_iou': 'sph2pob_efficient',
          'fov_iou': 'fov', 'sph_iou': 'sph', 'naive_iou': 'naive', 'unbiased_iou': 'unbiased',
          'sph2pob_legacy_iou': 'sph2pob_legacy'}


def sph_max_overlaps(bboxes1, bboxes2, backend='sph2pob_efficient_iou', mode='iou', box_version=None,
                     row_base=0, col_base=0, return_matrix=False):
    """For overlaps = backend(bboxes1, bboxes2) of shape (m, n) returns
    ``(row_max[m], row_argmax[m], col_max[n], col_argmax[n])`` (+ the matrix if asked), ties to the
    lowest index; ``row_base``/``col_base`` offset the reported indices (shards of a larger matrix)."""
    if box_version is not None:
        bboxes1, bboxes2 = bboxes1[..., :box_version], bboxes2[..., :box_version]
    with torch.no_grad():
        mat, rowm, colm = _native.iou_pairwise(_KINDS[backend], bboxes1, bboxes2, mode, "arc", want_matrix=return_matrix,
                                               want_row_max=True, want_col_max=True, row_base=row_base, col_base=col_base)
    res = (rowm[0], rowm[1].long(), colm[0], colm[1].long())
    return res + (mat,) if return_matrix else res
