from .sph_bbox_post import filter_scores_and_topk, get_bboxes_batch, get_bboxes_single  # noqa: F401
