import os, sys, statistics
sys.path.insert(0, '/root/repo')
import torch
from sph_retina_b200 import synthetic as S
from sph_retina_b200.sphdet.iou import sph_max_overlaps
A = S.generate_boxes(1 << 20, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=0).cuda()
G = S.generate_boxes(1024, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=1).cuda()
def run(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ms = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return statistics.median(ms)
ms = run(lambda: sph_max_overlaps(A, G))
print("sweep %.3f ms %.1f Gpairs/s" % (ms, (1 << 30) / ms / 1e6))
ms = run(lambda: sph_max_overlaps(G, A))
print("sweep transposed %.3f ms %.1f Gpairs/s" % (ms, (1 << 30) / ms / 1e6))
