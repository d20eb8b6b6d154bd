"""One launch set of the unbiased_iou aligned kernel on configs[0]'s boxes (for ncu):  python tools/run_unbiased.py [bfov|rbfov]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sph_retina_b200 import synthetic as S  # noqa: E402
from sph_retina_b200.sphdet.iou import unbiased_iou  # noqa: E402

box = sys.argv[1] if len(sys.argv) > 1 else "bfov"
dev = torch.device("cuda:0")
n = 1_000_000
b1 = S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box=box, seed=0).to(dev)
b2 = S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box=box, seed=1).to(dev)
for _ in range(5):
    out = unbiased_iou(b1, b2, is_aligned=True)
torch.cuda.synchronize()
print(float(out.mean()))
