from .sph2pob_iou_loss import (OBBIoULoss, Sph2PobDecodedIoULoss, Sph2PobIoULoss, SphIoULoss, rotated_iou, sph2pob_iou,
                               sph2pob_obbs)

__all__ = ['Sph2PobIoULoss', 'Sph2PobDecodedIoULoss', 'SphIoULoss', 'OBBIoULoss', 'sph2pob_iou', 'sph2pob_obbs', 'rotated_iou']
