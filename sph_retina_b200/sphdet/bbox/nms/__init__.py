from .sph_nms import SphNMS, multiclass_nms, sph_batched_nms, sph_batched_nms_images, sph_nms_image_blocks

__all__ = ['SphNMS', 'multiclass_nms', 'sph_batched_nms', 'sph_batched_nms_images', 'sph_nms_image_blocks']
