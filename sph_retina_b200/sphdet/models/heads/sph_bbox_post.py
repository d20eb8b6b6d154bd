"""Test-time box post-processing of the spherical RetinaNet head:
``SphRetinaHead._get_bboxes_single`` + ``_bbox_post_process`` (sphdet/models/heads/sph_retina_head.py:35-212), i.e. per
level sigmoid -> ``filter_scores_and_topk`` (mmdet/core/utils/misc.py:104-165) -> ``bbox_coder.decode``, then concat,
``SphNMS`` and the ``max_per_img`` cut.

The reference does this image by image, level by level: ~60 eager kernels, two host syncs per level (``nonzero``), and
the Python NMS loop per image.  ``get_bboxes_batch`` takes the head outputs of the whole batch: per level ONE ``topk``
over every image, one decode launch and the device-side NMS pipeline (``sphk_nms_images``: sort, suppression, per-image
ordering in three launches) for all images x levels x classes; the only host synchronisation is reading the per-image
counts of the kept detections.  ``get_bboxes_single`` keeps the reference's
per-image contract on top of the same kernels."""
from __future__ import annotations

import torch

from ...bbox.nms.sph_nms import PlanarNMS, SphNMS, sph_batched_nms_images, sph_nms_image_blocks


def filter_scores_and_topk(scores, score_thr, topk, results=None):
    """mmdet/core/utils/misc.py:104-165, verbatim semantics: keep scores > score_thr, the `topk` largest of them,
    returned score-descending as (scores, labels, anchor_idxs, filtered_results)."""
    valid_mask = scores > score_thr
    scores = scores[valid_mask]
    valid_idxs = torch.nonzero(valid_mask)
    num_topk = min(topk, valid_idxs.size(0))
    scores, idxs = scores.sort(descending=True)
    scores = scores[:num_topk]
    topk_idxs = valid_idxs[idxs[:num_topk]]
    keep_idxs, labels = topk_idxs.unbind(dim=1)
    filtered_results = None
    if results is not None:
        if isinstance(results, dict):
            filtered_results = {k: v[keep_idxs] for k, v in results.items()}
        elif isinstance(results, list):
            filtered_results = [result[keep_idxs] for result in results]
        elif isinstance(results, torch.Tensor):
            filtered_results = results[keep_idxs]
        else:
            raise NotImplementedError(f'Only supports dict or list or Tensor, but get {type(results)}.')
    return scores, labels, keep_idxs, filtered_results


def _cfg_get(cfg, key, default=None):
    return cfg.get(key, default) if hasattr(cfg, "get") else getattr(cfg, key, default)


def get_bboxes_single(cls_score_list, bbox_pred_list, mlvl_priors, bbox_coder, cfg, box_version=4, use_sigmoid_cls=True,
                      with_nms=True, score_factor_list=None, img_shape=None):
    """One image, the reference's contract (sph_retina_head.py:103-212 then :35-101): cls_score_list[l] is
    [A * C, H, W], bbox_pred_list[l] is [A * box_version, H, W], mlvl_priors[l] is [H * W * A, box_version].
    Returns (det_bboxes [K, box_version + 1], det_labels [K]) -- or (bboxes, scores, labels) when with_nms is False.

    The anchor-free heads share the block (sph_fcos_head.py:196-321): ``score_factor_list[l]`` ([1, H, W] centerness
    logits) multiplies the kept scores by its sigmoid before the NMS (:225-227), ``mlvl_priors[l]`` are then the [H * W, 2]
    points of ``DistancePointSphBBoxCoder`` and ``img_shape`` (H, W) is handed to ``decode`` as ``max_shape`` (:312-313)."""
    nms_pre = _cfg_get(cfg, 'nms_pre', -1)
    score_thr = _cfg_get(cfg, 'score_thr', 0.0)
    mlvl_bboxes, mlvl_scores, mlvl_labels, mlvl_factors = [], [], [], []
    if score_factor_list is None or score_factor_list[0] is None:
        score_factor_list = [None] * len(cls_score_list)
    decode_kw = {} if img_shape is None else dict(max_shape=img_shape)
    for cls_score, bbox_pred, priors, score_factor in zip(cls_score_list, bbox_pred_list, mlvl_priors, score_factor_list):
        assert cls_score.size()[-2:] == bbox_pred.size()[-2:]
        bbox_pred = bbox_pred.permute(1, 2, 0).reshape(-1, box_version)
        num_cls = cls_score.size(0) // (bbox_pred.size(0) // (cls_score.size(-1) * cls_score.size(-2)))
        cls_score = cls_score.permute(1, 2, 0).reshape(-1, num_cls)
        scores = cls_score.sigmoid() if use_sigmoid_cls else cls_score.softmax(-1)[:, :-1]
        topk = nms_pre if nms_pre > 0 else scores.numel()
        scores, labels, keep_idxs, filtered = filter_scores_and_topk(scores, score_thr, topk, dict(bbox_pred=bbox_pred, priors=priors))
        mlvl_bboxes.append(bbox_coder.decode(filtered['priors'], filtered['bbox_pred'], **decode_kw))
        mlvl_scores.append(scores)
        mlvl_labels.append(labels)
        if score_factor is not None:
            mlvl_factors.append(score_factor.permute(1, 2, 0).reshape(-1).sigmoid()[keep_idxs])
    bboxes, scores, labels = torch.cat(mlvl_bboxes), torch.cat(mlvl_scores), torch.cat(mlvl_labels)
    if mlvl_factors:
        scores = scores * torch.cat(mlvl_factors)
    if not with_nms:
        return bboxes, scores, labels
    if bboxes.numel() == 0:
        return torch.cat([bboxes, scores[:, None]], -1), labels
    if _cfg_get(cfg, 'iou_calculator', 'sph2pob_efficient') == 'planar':          # sph_retina_head.py:89-90
        nms = PlanarNMS(box_formator=_cfg_get(cfg, 'box_formator', 'sph2pix'))
    else:
        nms = SphNMS(iou_calculator=_cfg_get(cfg, 'iou_calculator', 'sph2pob_efficient'))
    det_bboxes, keep = nms(bboxes, scores, labels, _cfg_get(cfg, 'nms'))
    max_per_img = _cfg_get(cfg, 'max_per_img', det_bboxes.size(0))
    return det_bboxes[:max_per_img], labels[keep][:max_per_img]


def get_bboxes_batch(cls_scores, bbox_preds, mlvl_priors, bbox_coder, cfg, box_version=4):
    """The whole batch at once (sigmoid classification heads).  cls_scores[l]: [B, A * C, H, W]; bbox_preds[l]:
    [B, A * box_version, H, W]; mlvl_priors[l]: [H * W * A, box_version].  Returns a list of B
    (det_bboxes [K_b, box_version + 1], det_labels [K_b]) equal to ``get_bboxes_single`` image by image (up to the
    order of equal scores, which is unspecified in the reference as well)."""
    nms_pre = _cfg_get(cfg, 'nms_pre', -1)
    score_thr = float(_cfg_get(cfg, 'score_thr', 0.0))
    nms_cfg = _cfg_get(cfg, 'nms')
    if nms_cfg is None:
        raise ValueError("test_cfg.nms is None: the reference aborts the process here (sph_nms.py:24-29)")
    unknown = set(nms_cfg) - {'type', 'iou_threshold', 'max_num', 'class_agnostic', 'split_thr'}
    if unknown:
        raise ValueError("get_bboxes_batch: unsupported test_cfg.nms keys %s" % sorted(unknown))
    iou_thr = float(nms_cfg.get('iou_threshold', 0.5))
    iou_calc = _cfg_get(cfg, 'iou_calculator', 'sph2pob_efficient')      # test_cfg.iou_calculator (sph_retina_head.py:89-90)
    if iou_calc == 'planar':                   # class-agnostic planar NMS per image: the per-image contract covers it
        return [get_bboxes_single([c[b] for c in cls_scores], [p[b] for p in bbox_preds], mlvl_priors, bbox_coder, cfg,
                                  box_version=box_version) for b in range(cls_scores[0].size(0))]
    SphNMS(iou_calc)                                                      # same refusals as the per-image path
    B, D = cls_scores[0].size(0), box_version
    sc, lb, dl, pr = [], [], [], []
    num_cls = None
    for cls_score, bbox_pred, priors in zip(cls_scores, bbox_preds, mlvl_priors):
        n = priors.size(0)
        deltas = bbox_pred.permute(0, 2, 3, 1).reshape(B, n, D)
        num_cls = cls_score.size(1) * cls_score.size(2) * cls_score.size(3) // n
        scores = cls_score.permute(0, 2, 3, 1).reshape(B, n * num_cls).sigmoid()
        k = min(nms_pre, n * num_cls) if nms_pre > 0 else n * num_cls
        # topk of the admissible scores; the inadmissible ones sink to -1 and are marked invalid below
        top, idx = torch.where(scores > score_thr, scores, scores.new_full((), -1.0)).topk(k, dim=1)
        anchor = idx // num_cls
        sc.append(top)
        lb.append(idx - anchor * num_cls)
        dl.append(deltas.gather(1, anchor[..., None].expand(-1, -1, D)))
        pr.append(priors[anchor])
    scores, labels = torch.cat(sc, 1), torch.cat(lb, 1)                    # [B, K]
    K = scores.size(1)
    boxes = bbox_coder.decode(torch.cat(pr, 1).reshape(-1, D), torch.cat(dl, 1).reshape(-1, D))     # one launch
    scores, labels = scores.reshape(-1), labels.reshape(-1)
    # per-image cap: nms.max_num is applied by sph_batched_nms (sph_nms.py:46-53) before the head's max_per_img
    max_per_img = min(int(_cfg_get(cfg, 'max_per_img', K)), int(nms_cfg.get('max_num', K)), K)
    counts = None
    if K <= 16384:
        # sort, suppression and per-image ordering on the device; the one host synchronisation is reading the counts
        idx, count = sph_nms_image_blocks(boxes, scores, labels, B, num_cls, iou_thr, max_per_img, valid=scores > score_thr,
                                          iou_calculator=iou_calc)
        counts = count.tolist()
        if min(counts) >= 0:
            sels = [idx[b, :counts[b]].long() for b in range(B)]
        else:
            counts = None          # an (image, class) segment beyond the device pipeline's sorting buffer: general path
    if counts is None:
        image_ids = torch.arange(B, device=scores.device).repeat_interleave(K)
        keep = sph_batched_nms_images(boxes, scores, labels, image_ids, iou_thr, num_images=B, num_classes=num_cls,
                                      max_per_segment=K, valid=scores > score_thr, iou_calculator=iou_calc)
        counts = torch.bincount(image_ids[keep], minlength=B).tolist()
        sels, start = [], 0
        for b in range(B):
            sels.append(keep[start:start + min(counts[b], max_per_img)])
            start += counts[b]
    return [(torch.cat([boxes[sel], scores[sel, None]], -1), labels[sel]) for sel in sels]
