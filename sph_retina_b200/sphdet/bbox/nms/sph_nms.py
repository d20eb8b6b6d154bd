"""Spherical NMS with the reference's interface (sphdet/bbox/nms/sph_nms.py:7-74).

The reference runs a Python ``while`` loop per class with one IoU-API call and a ``nonzero`` host
sync per kept box.  Here all (image, class) segments go through ONE kernel launch
(``sphk_nms_batched``: warp-ballot suppression words + in-warp greedy scan); the host side only
sorts and slices."""
from __future__ import annotations

import torch

from .... import _native


_LABEL_CAP = 1024     # class ids the single-image fast path of sph_batched_nms covers (LVIS has 1203: general path)


def _desc_score_key(scores):
    """int64 in [0, 2**32): ascending key order == descending float32 score (IEEE bit trick, no sort yet)."""
    b = scores.float().contiguous().view(torch.int32)
    m = b ^ ((b >> 31) & 0x7FFFFFFF)            # signed-int order == float order
    return (~m).to(torch.int64) + 0x80000000    # reversed, shifted into [0, 2**32)


def _segments(scores, seg_ids, desc=None):
    """One sort on the composite key (segment ascending, score descending): order, offsets, longest segment.
    (The reference sorts per class with torch.argsort(descending=True), sph_nms.py:65; equal scores are in
    unspecified order there as well.)"""
    key = (seg_ids.long() << 32) | (_desc_score_key(scores) if desc is None else desc)
    key, order = torch.sort(key)
    _, counts = torch.unique_consecutive(key >> 32, return_counts=True)
    offsets = torch.zeros(counts.numel() + 1, dtype=torch.int32, device=scores.device)
    offsets[1:] = counts.cumsum(0)
    return order, offsets, int(counts.max().item())


def _keep_indices(boxes, scores, seg_ids, iou_threshold, desc=None, iou_calculator='sph2pob_efficient'):
    """Indices (into boxes) that survive the per-segment greedy NMS, unordered."""
    order, offsets, longest = _segments(scores, seg_ids, desc)
    flags = _native.nms_batched(boxes, order, offsets, longest, iou_threshold, iou_calculator=iou_calculator)
    return order[flags == 1]            # (no segment is refused here: the kernel is sized for the longest one)


def sph_batched_nms(boxes, scores, idxs, nms_cfg, iou_calculator='sph2pob_efficient', class_agnostic=False):
    """sph_nms.py:22-60.  Returns ``(dets[K, D+1], keep[K])``, kept boxes sorted by descending score."""
    if nms_cfg is None:
        raise ValueError("nms_cfg is None: the reference aborts the process here (sph_nms.py:24-29)")
    nms_cfg_ = nms_cfg.copy()
    # the reference pops these and never uses them (:33-37): NMS is ALWAYS per label
    nms_cfg_.pop('class_agnostic', class_agnostic)
    nms_cfg_.pop('type', 'nms')
    nms_cfg_.pop('split_thr', 10000)
    iou_threshold = nms_cfg_.pop('iou_threshold', 0.5)
    max_num = min(nms_cfg_.pop('max_num', boxes.shape[0]), boxes.shape[0])
    assert boxes.size(1) in [4, 5]
    if boxes.size(0) == 0:
        return torch.cat([boxes, scores[:, None]], -1), boxes.new_zeros((0,), dtype=torch.long)
    if boxes.size(0) <= 16384 and boxes.is_cuda:
        # one image = one block of the device-side pipeline (sort, suppression, ordering: three launches); labels are
        # taken to lie in [0, 1024) -- an image that breaks this reports count -1 and goes through the general path
        idx, count = _native.nms_images(boxes, scores, idxs, 1, _LABEL_CAP, iou_threshold, max(int(max_num), 1),
                                        iou_calculator=iou_calculator)
        n = int(count)                                       # the host synchronisation the caller needs anyway
        if n >= 0:
            keep = idx[0, :min(n, max_num)].long()
            return torch.cat([boxes[keep], scores[keep, None].to(boxes.dtype)], -1), keep
    keep = _keep_indices(boxes, scores, idxs, iou_threshold, iou_calculator=iou_calculator)
    keep = keep.sort()[0]                                   # :49 nonzero() order
    kept_scores, inds = scores[keep].sort(descending=True)  # :51
    keep = keep[inds][:max_num]
    dets = torch.cat([boxes[keep], kept_scores[:max_num, None]], -1)
    return dets, keep


def sph_batched_nms_images(boxes, scores, labels, image_ids, iou_threshold=0.5, num_images=None, num_classes=None,
                           max_per_segment=None, valid=None, iou_calculator='sph2pob_efficient'):
    """Test-time batch: one launch over every (image, class) segment of a whole batch (labels < 2**20,
    image ids < 2**11).  Returns the kept indices (into boxes), grouped by image and score-descending inside an image.

    With the three hints (batch size, number of classes, an upper bound of the boxes per (image, class) segment, e.g.
    nms_pre) the segment table is built densely on the device and the only host synchronisation left is the final
    compaction of the kept indices.  ``valid`` (bool [M], needs the hints): boxes to leave out -- padded candidate lists
    go in as they are; the invalid ones are parked behind the last segment and never looked at.
    A label outside [0, num_classes) raises; a segment longer than ``max_per_segment`` (which the kernel refuses) sends the
    call through the exact-length path instead of dropping its boxes."""
    desc = _desc_score_key(scores)
    if num_images is None or num_classes is None or max_per_segment is None:
        if valid is not None:
            raise ValueError("valid= needs num_images, num_classes and max_per_segment")
        seg = (image_ids.long() << 20) | labels.long()
        keep = _keep_indices(boxes, scores, seg, iou_threshold, desc, iou_calculator)
    else:
        nseg = int(num_images) * int(num_classes)
        seg = image_ids.long() * int(num_classes) + labels.long()
        if valid is not None:
            seg = torch.where(valid, seg, seg.new_full((), nseg))
        key, order = torch.sort((seg << 32) | desc)
        counts = torch.bincount(seg, minlength=nseg + 1)[:nseg]
        offsets = torch.zeros(counts.numel() + 1, dtype=torch.int32, device=scores.device)
        offsets[1:] = counts.cumsum(0)
        typical = max(1, (2 * boxes.size(0)) // max(1, counts.numel()))     # twice the mean segment length
        flags = _native.nms_batched(boxes, order, offsets, int(max_per_segment), iou_threshold, typical, iou_calculator)
        # one more scalar next to the compaction's synchronisation: labels in range, no segment refused (flag 0xFF)
        in_range = (labels >= 0) & (labels < int(num_classes))
        if valid is not None:
            in_range = in_range | ~valid
        problems = torch.stack([(~in_range).any(), (flags == 0xFF).any()])
        keep = order[flags == 1]
        bad_label, refused = problems.tolist()
        if bad_label:
            raise ValueError("sph_batched_nms_images: a label lies outside [0, num_classes = %d)" % int(num_classes))
        if refused:        # the hint was too small: same result through the exact-length path, nothing is dropped
            sel = torch.arange(boxes.size(0), device=boxes.device) if valid is None else valid.nonzero(as_tuple=False).view(-1)
            seg = (image_ids[sel].long() << 20) | labels[sel].long()
            keep = sel[_keep_indices(boxes[sel], scores[sel], seg, iou_threshold, desc[sel], iou_calculator)]
    return keep[torch.argsort((image_ids[keep].long() << 32) | desc[keep])]


def sph_nms_image_blocks(boxes, scores, labels, num_images, num_classes, iou_threshold=0.5, max_per_img=None, valid=None,
                         iou_calculator='sph2pob_efficient'):
    """Test-time NMS of a whole batch whose candidates come as `num_images` equal, contiguous blocks (what the head's
    post-processing emits): sort, per-(image, class) suppression and the per-image score ordering all run on the device
    (``sphk_nms_images``: three launches, no host synchronisation).  Returns ``(idx [num_images, max_per_img] int32,
    count [num_images] int32)``: ``idx[b, :count[b]]`` are the kept boxes of image b, score-descending -- the same
    boxes ``SphNMS`` keeps for that image (equal scores are ordered by index here, unspecified in the reference).
    ``count[b] == -1`` flags an image the pipeline does not cover (a label outside [0, num_classes), or more than 4096
    candidates of one class): run that batch through :func:`sph_batched_nms_images` instead."""
    per_image = boxes.size(0) // max(1, num_images)
    max_out = per_image if max_per_img is None else min(int(max_per_img), per_image)
    return _native.nms_images(boxes, scores, labels, int(num_images), int(num_classes), iou_threshold, max(max_out, 1), valid,
                              iou_calculator=iou_calculator)


class SphNMS:
    """sph_nms.py:7-19.  All three calculators of the reference run in the same NMS kernel: 'sph2pob_efficient' (the
    default), 'naive_iou' (the indoor360 configs: planar IoU of the sph2pix boxes) and 'unbiased_iou' (the pandora configs:
    the exact spherical IoU, CPU numpy in the reference, a double-precision device function here)."""

    def __init__(self, iou_calculator='sph2pob_efficient'):
        if iou_calculator in ('sph2pob_efficient', 'naive_iou', 'unbiased_iou'):
            self.iou_calculator = iou_calculator
        else:
            raise NotImplementedError('Not supported iou_calculator.')

    def __call__(self, boxes, scores, idxs, nms_cfg, class_agnostic=False):
        return sph_batched_nms(boxes, scores, idxs, nms_cfg, self.iou_calculator, class_agnostic)


class PlanarNMS:
    """sphdet/bbox/nms/planar_nms.py:7-18 -- what the head uses when ``test_cfg.iou_calculator == 'planar'``
    (sph_retina_head.py:89-90): the boxes read as planar (x1, y1, x2, y2) boxes of the 512 x 1024 image
    (``Sph2PlanarBoxTransform('sph2pix')``) and mmcv's ``batched_nms`` of type 'nms' on them.  That is the suppression rule
    of ``SphNMS('naive_iou')`` for BFoV boxes (IoU = inter / (a1 + a2 - inter), suppress iff IoU > thr -- a NaN IoU of
    two zero-area boxes keeps the box, as mmcv's ``inter > thr * union`` does, where SphNMS would drop it), class-agnostic by
    default (planar_nms.py:11) -- so it runs in the same NMS kernel with one segment per image, or one per label when
    ``class_agnostic`` is False (mmcv's coordinate-offset trick separates the classes exactly)."""

    def __init__(self, box_formator='sph2pix'):
        if box_formator != 'sph2pix':
            raise NotImplementedError("PlanarNMS: only box_formator='sph2pix' (the reference's default) has a kernel")
        self.box_formator = box_formator

    def __call__(self, boxes, scores, idxs, nms_cfg, class_agnostic=True):
        if nms_cfg is None:                                   # mmcv batched_nms: no NMS, sorted by score
            scores, inds = scores.sort(descending=True)
            return torch.cat([boxes[inds], scores[:, None]], -1), inds
        cfg = nms_cfg.copy()
        class_agnostic = cfg.pop('class_agnostic', class_agnostic)
        if cfg.pop('type', 'nms') != 'nms':
            raise NotImplementedError("PlanarNMS: only mmcv's hard 'nms' has a kernel here")
        if boxes.size(-1) != 4:
            raise NotImplementedError("PlanarNMS: mmcv's nms takes (x1, y1, x2, y2): BFoV boxes only")
        cfg.pop('split_thr', None)                            # (a batching detail of mmcv: same result either way)
        thr = cfg.pop('iou_threshold', 0.5)
        max_num = cfg.pop('max_num', -1)
        score_thr = cfg.pop('score_threshold', 0)
        index = None
        if score_thr > 0:                                     # mmcv nms: boxes at or below the score threshold drop out first
            index = (scores > score_thr).nonzero(as_tuple=False).view(-1)
            boxes, scores, idxs = boxes[index], scores[index], idxs[index]
        labels = torch.zeros_like(idxs) if class_agnostic else idxs
        inner = dict(iou_threshold=thr)
        if max_num > 0:
            inner['max_num'] = max_num
        dets, keep = sph_batched_nms(boxes, scores, labels, inner, 'planar')      # naive IoU, mmcv's `>` rule (NaN keeps)
        return dets, (keep if index is None else index[keep])


def multiclass_nms(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1, score_factors=None,
                   return_inds=False, nms_op=None, box_version=4):
    """sphdet/bbox/nms/utils.py:6-95 (R-CNN heads) with ``nms_op`` defaulting to :class:`SphNMS`."""
    nms_op = SphNMS() if nms_op is None else nms_op
    num_classes = multi_scores.size(1) - 1
    if multi_bboxes.shape[1] > box_version:
        bboxes = multi_bboxes.view(multi_scores.size(0), -1, box_version)
    else:
        bboxes = multi_bboxes[:, None].expand(multi_scores.size(0), num_classes, box_version)
    scores = multi_scores[:, :-1]
    labels = torch.arange(num_classes, dtype=torch.long, device=scores.device).view(1, -1).expand_as(scores)
    bboxes, scores, labels = bboxes.reshape(-1, box_version), scores.reshape(-1), labels.reshape(-1)
    valid_mask = scores > score_thr
    if score_factors is not None:
        scores = scores * score_factors.view(-1, 1).expand(multi_scores.size(0), num_classes).reshape(-1)
    inds = valid_mask.nonzero(as_tuple=False).squeeze(1)
    bboxes, scores, labels = bboxes[inds], scores[inds], labels[inds]
    if bboxes.numel() == 0:
        dets = torch.cat([bboxes, scores[:, None]], -1)
        return (dets, labels, inds) if return_inds else (dets, labels)
    dets, keep = nms_op(bboxes, scores, labels, nms_cfg)
    if max_num > 0:
        dets, keep = dets[:max_num], keep[:max_num]
    return (dets, labels[keep], inds[keep]) if return_inds else (dets, labels[keep])
