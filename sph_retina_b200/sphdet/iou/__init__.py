from .sph_iou_api import (fov_iou, naive_iou, sph2pob_efficient_iou, sph2pob_legacy_iou, sph2pob_standard_iou, sph_iou,
                          unbiased_iou)
from .sph_iou_calculator import SphOverlaps2D, sph_overlaps
from .assign import sph_max_overlaps

__all__ = ['SphOverlaps2D', 'sph_overlaps', 'sph2pob_standard_iou', 'sph2pob_efficient_iou', 'sph2pob_legacy_iou', 'fov_iou', 'sph_iou',
           'naive_iou', 'unbiased_iou', 'sph_max_overlaps']
