from .sph_nms import PlanarNMS, SphNMS, multiclass_nms, sph_batched_nms, sph_batched_nms_images, sph_nms_image_blocks

__all__ = ['SphNMS', 'PlanarNMS', 'multiclass_nms', 'sph_batched_nms', 'sph_batched_nms_images', 'sph_nms_image_blocks']
