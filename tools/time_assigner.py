import sys, os, statistics, torch
sys.path.insert(0, os.getcwd())
from sph_retina_b200 import synthetic as S
from sph_retina_b200.sphdet.assigners import SphMaxIoUAssigner
dev = torch.device("cuda:0")
gts, anchors = S.assignment_batch(images=16)
gts, anchors = gts.to(dev), anchors.to(dev)
A = SphMaxIoUAssigner(0.5, 0.4, min_pos_iou=0, iou_calculator=dict(type='SphOverlaps2D', backend='sph2pob_efficient_iou', box_version=5))
def run(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ms = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return statistics.median(ms)
print("assign_batch 16 img: %.3f ms" % run(lambda: A.assign_batch(anchors, list(gts))))
print("assign per image x16: %.3f ms" % run(lambda: [A.assign(anchors, gts[i]) for i in range(16)]))
