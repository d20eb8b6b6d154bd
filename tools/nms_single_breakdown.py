"""Where the time of one SphNMS call (5000 candidates, 80 classes) goes: device pipeline vs Python around it."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
from sph_retina_b200 import synthetic as S, _native
from sph_retina_b200.sphdet.bbox.nms import SphNMS
boxes, scores, labels, _ = (t.cuda() for t in S.nms_batch(1, 5000, 80))
def t(name, fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); print("%-46s %.1f us" % (name, (time.perf_counter() - t0) / n * 1e6))
def ev(name, fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ms = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    print("%-46s %.1f us (device, median)" % (name, sorted(ms)[len(ms) // 2] * 1e3))
nms = SphNMS()
t("SphNMS call, wall", lambda: nms(boxes, scores, labels, dict(iou_threshold=0.5, max_num=100)))
t("nms_images only, wall (async)", lambda: _native.nms_images(boxes, scores, labels, 1, 1024, 0.5, 100))
ev("nms_images only", lambda: _native.nms_images(boxes, scores, labels, 1, 1024, 0.5, 100))
ev("nms_images, 80 classes declared", lambda: _native.nms_images(boxes, scores, labels, 1, 80, 0.5, 100))
idx, count = _native.nms_images(boxes, scores, labels, 1, 1024, 0.5, 100)
t("int(count)", lambda: int(count))
n = int(count)
t("keep + dets assembly", lambda: (lambda keep: torch.cat([boxes[keep], scores[keep, None]], -1))(idx[0, :n].long()))
