from .delta_sph_bbox_coder import (DeltaXYWHASphBBoxCoder, DeltaXYWHSphBBoxCoder, bbox2delta, delta2bbox)  # noqa: F401
from .distance_point_sph_bbox_coder import DistancePointSphBBoxCoder, bbox2distance, distance2bbox  # noqa: F401
