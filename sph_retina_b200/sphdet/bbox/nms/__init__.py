from .sph_nms import SphNMS, multiclass_nms, sph_batched_nms, sph_batched_nms_images

__all__ = ['SphNMS', 'multiclass_nms', 'sph_batched_nms', 'sph_batched_nms_images']
