"""Launches the device NMS pipeline on the two shapes of interest (for `ncu --metrics gpu__time_duration.sum`)."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from sph_retina_b200 import synthetic as S, _native
b1 = [t.cuda() for t in S.nms_batch(1, 5000, 80)]
b64 = [t.cuda() for t in S.nms_batch(64, 1000, 80)]
for _ in range(3):
    _native.nms_images(b1[0], b1[1], b1[2], 1, 1024, 0.5, 100)
    _native.nms_images(b64[0], b64[1], b64[2], 64, 80, 0.5, 100)
torch.cuda.synchronize()
