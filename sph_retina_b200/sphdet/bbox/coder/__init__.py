from .delta_sph_bbox_coder import (DeltaXYWHASphBBoxCoder, DeltaXYWHSphBBoxCoder, bbox2delta, delta2bbox)  # noqa: F401
