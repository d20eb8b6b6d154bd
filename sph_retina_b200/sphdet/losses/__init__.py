from .sph2pob_gd_loss import Sph2PobGDLoss
from .sph2pob_iou_loss import (OBBIoULoss, Sph2PobDecodedIoULoss, Sph2PobIoULoss, SphIoULoss, SphIoULossLegacy, rotated_iou, sph2pob_iou,
                               sph2pob_obbs)
from .sph2pob_kf_loss import Sph2PobKFLoss
from .sph2pob_l1_loss import Sph2PobL1Loss

__all__ = ['Sph2PobIoULoss', 'Sph2PobDecodedIoULoss', 'SphIoULoss', 'SphIoULossLegacy', 'OBBIoULoss', 'Sph2PobGDLoss', 'Sph2PobKFLoss',
           'Sph2PobL1Loss', 'sph2pob_iou', 'sph2pob_obbs', 'rotated_iou']
