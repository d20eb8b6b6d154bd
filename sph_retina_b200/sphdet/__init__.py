"""Mirror of the reference's ``sphdet`` package for the IoU hot path (iou, losses, bbox.nms)."""
