import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
from sph_retina_b200 import synthetic as S, _native
from sph_retina_b200.sphdet.bbox.nms import sph_batched_nms_images
from sph_retina_b200.sphdet.bbox.nms.sph_nms import _desc_score_key, _segments
boxes, scores, labels, image_ids = (x.cuda() for x in S.nms_batch(64, 1000, 80))
def t(name, fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); print("%-40s %.1f us" % (name, (time.perf_counter()-t0)/n*1e6))
seg = (image_ids.long() << 20) | labels.long()
t("seg id", lambda: (image_ids.long() << 20) | labels.long())
t("desc key", lambda: _desc_score_key(scores))
key = (seg << 32) | _desc_score_key(scores)
t("sort int64 64000", lambda: torch.sort(key))
ks, order = torch.sort(key)
t("unique_consecutive", lambda: torch.unique_consecutive(ks >> 32, return_counts=True))
t("_segments total", lambda: _segments(scores, seg))
order, offsets, longest = _segments(scores, seg)
t("nms kernel call", lambda: _native.nms_batched(boxes, order, offsets, longest, 0.5))
flags = _native.nms_batched(boxes, order, offsets, longest, 0.5)
t("order[flags.bool()]", lambda: order[flags.bool()])
keep = order[flags.bool()]
t("final sort", lambda: keep[torch.argsort((image_ids[keep].long() << 32) | _desc_score_key(scores[keep]))])
t("total", lambda: sph_batched_nms_images(boxes, scores, labels, image_ids, 0.5))
