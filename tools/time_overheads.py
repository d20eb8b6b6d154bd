#!/usr/bin/env python
"""What the CUDA-event bracket itself costs (empty kernel), and the aligned kernel with / without an L2 flush,
single calls vs back-to-back calls over rotating input sets (footprint > L2)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sph_retina_b200 import synthetic as S, _native as N
from sph_retina_b200.sphdet.iou import sph2pob_efficient_iou, sph_iou
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
sink = torch.empty(256, device='cuda')
def run(fn, n=20, do_flush=True):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ms = []
    for _ in range(n):
        if do_flush: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return statistics.median(ms) * 1e3
st = torch.cuda.current_stream().cuda_stream
empty = lambda: N.lib.sphk_probe_fp32(1, 1, sink.data_ptr(), st)
print("empty kernel, flush   : %6.1f us" % run(empty))
print("empty kernel, no flush: %6.1f us" % run(empty, do_flush=False))
print("nothing, flush        : %6.1f us" % run(lambda: None))
n = int(os.environ.get("N", 1000000))
K = 8
sets = [(S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box="bfov", seed=2 * i).cuda(),
         S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box="bfov", seed=2 * i + 1).cuda()) for i in range(K)]
b1, b2 = sets[0]
for name, f in (("sph2pob", sph2pob_efficient_iou), ("sph_iou", sph_iou)):
    print("%s 1 call, flush      : %6.1f us" % (name, run(lambda: f(b1, b2, is_aligned=True))))
    print("%s 1 call, no flush   : %6.1f us" % (name, run(lambda: f(b1, b2, is_aligned=True), do_flush=False)))
    t = run(lambda: [f(a, b, is_aligned=True) for a, b in sets], do_flush=False)
    print("%s %d calls rotating   : %6.1f us per call (%d MB footprint)  %.1f Gpairs/s" % (name, K, t / K, K * n * 36 >> 20, n / (t / K) / 1e3))
    t = run(lambda: [f(a, b, is_aligned=True) for a, b in sets], do_flush=True)
    print("%s %d calls rot+flush  : %6.1f us per call" % (name, K, t / K))
# CUDA-graph replay: no host launch cost between the calls
for name, f in (("sph2pob", sph2pob_efficient_iou), ("sph_iou", sph_iou)):
    for reps in (1, 8):
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for a, b in sets[:reps]: f(a, b, is_aligned=True)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            outs = [f(a, b, is_aligned=True) for a, b in sets[:reps]]
        for fl in (True, False):
            t = run(g.replay, do_flush=fl)
            print("%s graph of %d call(s), flush=%d: %6.1f us per call  %.1f Gpairs/s" % (name, reps, fl, t / reps, n / (t / reps) / 1e3))
