from .sph2pob_iou_loss import OBBIoULoss, Sph2PobIoULoss, SphIoULoss, sph2pob_iou, sph2pob_obbs, rotated_iou

__all__ = ['Sph2PobIoULoss', 'SphIoULoss', 'OBBIoULoss', 'sph2pob_iou', 'sph2pob_obbs', 'rotated_iou']
