"""sph_retina_b200 -- B200-native (sm_100a) spherical-box IoU hot path of ManuelVeras/sph-retina.

Public surface mirrors the reference's (``sphdet.iou``, ``sphdet.losses``, ``sphdet.bbox.nms``):

    from sph_retina_b200.sphdet.iou import SphOverlaps2D, sph2pob_efficient_iou, sph_iou, fov_iou
    from sph_retina_b200.sphdet.losses import Sph2PobIoULoss
    from sph_retina_b200.sphdet.bbox.nms import SphNMS

All arithmetic runs in the hand-written CUDA kernels of ``csrc/`` behind the C ABI of
``include/sphk.h``.  Importing any ``sphdet`` sub-package loads ``_lib/libsphk.so`` and raises
ImportError if it has not been built (``python -m sph_retina_b200.build``): there is no fallback."""
__version__ = "0.1.0"
