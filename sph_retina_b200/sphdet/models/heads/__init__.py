from .sph_bbox_post import filter_scores_and_topk, get_bboxes_batch, get_bboxes_single  # noqa: F401
from .sph_targets import get_targets_batch  # noqa: F401
