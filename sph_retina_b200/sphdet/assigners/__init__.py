from .sph_max_iou_assigner import AssignResult, SphMaxIoUAssigner

__all__ = ['SphMaxIoUAssigner', 'AssignResult']
