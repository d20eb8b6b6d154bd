#!/usr/bin/env python
"""bench.py -- throughput of the spherical-box IoU hot path (BASELINE.json metric: Sph2Pob-IoU pairs/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sweep|assign] [--impl reference]

Default workload = BASELINE.json configs[4] ("sweep"), the N x M configuration the metric is quoted on at 1/2/4/8 B200:
pairwise Sph2Pob-efficient IoU of 1,048,576 x 1,024 random RBFoV boxes (1.07 G pairs) with the fused per-anchor and
per-GT max/argmax -- the reduction MaxIoUAssigner applies to the matrix (mmdet/core/bbox/assigners/max_iou_assigner.py:
173-176) -- through ``sph_retina_b200.sharded.sharded_max_overlaps``.  One "step" = the whole sweep.  With N > 1
(torchrun, one rank per GPU) the anchors are row-sharded over the ranks, the 1,024 GT replicated, and the timed step
contains the exchange of the packed keys (pushed into the peers' symmetric buffers over NVLink by the compute kernel, or
one NCCL all_gather where symmetric memory is unavailable) and the unpack launch: STRONG scaling, total work fixed.  N = 1
runs the very same code path (the exchange degenerates to a view).  Timing: W untimed steps, synchronize + barrier, then
K timed steps, each with its own CUDA-event pair on the launching stream (the 256 MB L2 flush between steps sits outside
the spans), synchronize + barrier; the sum over the K steps, maximum over the ranks.  With N > 1 two further untimed steps
are enqueued between the barrier and the first timed step, so that every host thread is ahead of its GPU before the
first timed handshake between the GPUs (a rank leaving the barrier late would otherwise be timed by all the others).  ``--workload assign`` is configs[1] (16 per-image
``SphOverlaps2D`` calls of 32 GT x 98,208 anchors, full matrices written; replicas, weak scaling); it is also timed
briefly under ``other_configs`` of the default run, as are configs[0], [2], [3].

One JSON line on stdout (rank 0).  ``--impl reference`` times the reference's own CPU implementation of the same call --
its unmodified Python files, copied by ``__graft_entry__.build()`` into the git-ignored ``oracle/_ref/`` and imported
through ``oracle/ref_harness.py`` (the absent mmcv op bound to the reference's vendored ``diff_iou_rotated.py``); if
``oracle/_ref`` is missing, the torch restatement ``oracle/sph_oracle.py`` ("port") -- on the host cores, each step a
bounded row slice of the sweep.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sph2pob_iou_pairs_per_s"
UNIT = "pairs/s"
IMAGES, GTS, FLOP_PER_PAIR = 16, 32, 512.0      # W = 512 flop per Sph2Pob-IoU pair
SWEEP_ANCHORS, SWEEP_GTS = 1 << 20, 1024        # BASELINE.json configs[4]
REF_ROWS_PER_STEP = 4096                         # reference arm: anchors per step (x 1024 GT = 4.2 M pairs, ~1 s of CPU)
SWEEP_WORKLOAD = ("sweep: 1,048,576 x 1,024 RBFoV Sph2Pob-efficient IoU (BASELINE configs[4]), anchors row-sharded over the GPUs, "
                  "fused per-anchor and per-GT max/argmax, exchange of the packed keys + unpack launch in the timed region")
HBM_FALLBACK_GBS = 6650.0                       # B200_PROFILING.md fallback if MEASURED_PEAKS.json is absent


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE line (the JSON).  Libraries print there too (NCCL's version banner):
# route fd 1 to stderr for the whole run and keep the real stdout for the final line.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


# ---------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            return json.load(open(path)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": HBM_FALLBACK_GBS, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md), sampled through NVML from a
    background thread every ~2 ms (the timed region lasts tens of ms: too short for `nvidia-smi -lms`)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    _CHILD = r"""
import json, signal, sys, time
import pynvml as nv
nv.nvmlInit()
hs = [nv.nvmlDeviceGetHandleByIndex(int(a)) for a in sys.argv[1:]]
sm, power, mask, stop = [], [], 0, [False]
signal.signal(signal.SIGTERM, lambda *a: stop.__setitem__(0, True))
print("ready", flush=True)
while not stop[0]:
    for h in hs:
        try:
            sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
        except Exception:
            pass
    time.sleep(0.002)
print(json.dumps({"sm": sm, "mask": mask, "power": power,
                  "max": float(nv.nvmlDeviceGetMaxClockInfo(hs[0], nv.NVML_CLOCK_SM)) if hs else None}), flush=True)
"""

    def __init__(self, gpu_indices, external=False):
        """gpu_indices: the local GPUs to watch.  In a multi-rank run ONE rank watches all of them, from a child PROCESS
        (external=True): a sampler thread in every rank competes with that rank's launch thread for the interpreter lock
        and with the other ranks for the host cores, which the 32 us-per-call assign workload can feel."""
        self.external, self.child, self.gpu_indices = external and len(gpu_indices) > 0, None, list(gpu_indices)
        self.thread, self.stop_flag = None, False
        self.sm, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self.hs = []
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            for g in gpu_indices:
                phys = int(vis.split(",")[g]) if vis and vis.split(",")[g].isdigit() else g
                self.hs.append(pynvml.nvmlDeviceGetHandleByIndex(phys))
            self.nv = pynvml if self.hs else None
            if self.hs:
                self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.hs[0], pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            for h in self.hs:
                try:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
            time.sleep(0.002)

    def _physical(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        return [str(int(vis.split(",")[g]) if vis and vis.split(",")[g].isdigit() else g) for g in self.gpu_indices]

    def start(self):
        if self.external:
            import subprocess
            try:
                self.child = subprocess.Popen([sys.executable, "-c", self._CHILD] + self._physical(), stdout=subprocess.PIPE, text=True)
                if self.child.stdout.readline().strip() != "ready":
                    self.child = None
            except Exception:
                self.child = None
            return
        if self.nv is None:
            return
        import threading
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.external:
            if self.child is None:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
            self.child.terminate()
            try:
                d = json.loads(self.child.communicate(timeout=10)[0].strip().splitlines()[-1])
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml sampler failed"]}
            return {"sm_mhz": statistics.median(d["sm"]) if d["sm"] else None, "sm_max_mhz": d["max"],
                    "reasons": sorted(n for b, n in self.REASONS.items() if d["mask"] & b), "samples": len(d["sm"]),
                    "power_w_max": max(d["power"]) if d["power"] else None, "gpus_watched": len(self.gpu_indices)}
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag = True
        self.thread.join(timeout=2)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None}


LEAD_IN = 2      # multi-GPU runs: untimed steps between the barrier and the first timed step (time_steps)


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class L2Flusher:
    def __init__(self, torch, dev):
        self.buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # 2x the 126 MB L2

    def __call__(self):
        self.buf.zero_()


def time_steps(torch, step, steps, warmup, flush, barrier, lead_in=0):
    """W untimed + K timed steps; each timed step has its own CUDA-event pair on the launching stream, the L2
    flush between steps sits outside the timed spans.  Returns per-step ms.
    lead_in: untimed steps enqueued AFTER the barrier, back to back with the timed ones.  A multi-rank step ends with a
    handshake between the GPUs, so a rank whose host thread leaves the barrier a millisecond late would make every other
    GPU wait inside its first timed step; after the lead-in steps every host is several steps ahead of its GPU and the
    timed spans hold device work only."""
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    barrier()
    for _ in range(lead_in):
        flush()
        step()
    evs = []
    for _ in range(steps):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    barrier()
    return [a.elapsed_time(b) for a, b in evs]


def quick(torch, fn, iters=10, warmup=3, flush=None):
    ms = time_steps(torch, fn, iters, warmup, flush or (lambda: None), lambda: None)
    return statistics.median(ms)


# ---------------------------------------------------------------------------------------------------
# CPU side: the reference's own implementation (oracle/_ref) or its restatement as the reported baseline / reference arm
# ---------------------------------------------------------------------------------------------------
def sweep_inputs():
    """BASELINE configs[4]: A = generate_boxes(1,048,576, alpha/beta (1,100), 'rbfov') seed 0 as bboxes1
    (rows), G = generate_boxes(1,024, same) seed 1 as bboxes2."""
    from sph_retina_b200 import synthetic as S
    A = S.generate_boxes(SWEEP_ANCHORS, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=0)
    G = S.generate_boxes(SWEEP_GTS, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=1)
    return A, G


def load_cpu_reference():
    """(callable(b1, b2) -> [R, C] IoU matrix on the CPU, kind, description).  kind "reference": the reference's own
    sph2pob_efficient_iou from the files under oracle/_ref (copied there by __graft_entry__.build() in the build
    container; /root/reference itself does not exist on the GPU box).  kind "port": oracle/sph_oracle.py."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    ref_root = os.path.join(ROOT, "oracle", "_ref")
    if os.path.isfile(os.path.join(ref_root, "sphdet", "iou", "sph_iou_api.py")):
        try:
            os.environ["SPH_REFERENCE_ROOT"] = ref_root
            import ref_harness
            ref_harness.REFERENCE_ROOT = ref_root
            ns = ref_harness.load_reference()
            return (lambda b1, b2: ns.sph2pob_efficient_iou(b1, b2, is_aligned=False), "reference",
                    "the reference's own sphdet/iou/sph_iou_api.py::sph2pob_efficient_iou (files under oracle/_ref, unmodified; "
                    "mmcv.ops.box_iou_rotated bound to the reference's vendored sphdet/iou/diff_iou_rotated.py), PyTorch eager, "
                    "fp32, torch threads = cores")
        except Exception as e:      # fall through to the port, and say why
            log("bench.py: oracle/_ref present but not importable (%r); using the port" % (e,))
    import sph_oracle as O
    return (lambda b1, b2: O.sph2pob_iou(b1, b2, "efficient"), "port",
            "oracle/sph_oracle.py: the reference's PyTorch-eager algorithm restated, fp32, torch threads = cores")


def cpu_sweep_slice(fn, A, G, row0, rows):
    """One bounded sample of the sweep on the CPU: IoU(A[row0:row0+rows], G) and the two max/argmax reductions."""
    m = fn(A[row0:row0 + rows], G)
    return m.max(dim=1), m.max(dim=0)


def cpu_c_port_sweep(A, G, rows=16384):
    """Second CPU line: float64 C restatement with OpenMP (oracle/sph_oracle.c) on a row slice of the sweep."""
    import ctypes
    import numpy as np
    try:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
        lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libsph_oracle.so"))
    except Exception:  # pragma: no cover
        return None
    fp, dp = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
    r, c = np.ascontiguousarray(A[:rows].numpy()), np.ascontiguousarray(G.numpy())
    out = np.empty((r.shape[0], c.shape[0]))
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        lib.sph_oracle_iou_pairwise(0, r.ctypes.data_as(fp), ctypes.c_long(r.shape[0]), c.ctypes.data_as(fp),
                                    ctypes.c_long(c.shape[0]), r.shape[1], 0, 0, out.ctypes.data_as(dp))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return out.size / best


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    A, G = sweep_inputs()
    fn, kind, what = load_cpu_reference()
    rows = REF_ROWS_PER_STEP
    windows = SWEEP_ANCHORS // rows
    pairs_per_step = rows * SWEEP_GTS
    with torch.no_grad():
        for i in range(args.warmup):
            cpu_sweep_slice(fn, A, G, (i % windows) * rows, rows)
        t0 = time.perf_counter()
        for i in range(args.steps):
            cpu_sweep_slice(fn, A, G, ((args.warmup + i) % windows) * rows, rows)
        dt = time.perf_counter() - t0
    value = pairs_per_step * args.steps / dt
    sample = ("per step: %d consecutive anchors (a different window each step) x %d GT = %d pairs of the 1,048,576 x 1,024 sweep, "
              "IoU matrix + max/argmax over both axes" % (rows, SWEEP_GTS, pairs_per_step))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": SWEEP_WORKLOAD, "api": "sph2pob_efficient_iou(A, G, is_aligned=False) (CPU, PyTorch eager)",
                   "pairs_per_step": SWEEP_ANCHORS * SWEEP_GTS, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "what": what},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------
def fp32_peak(torch, native, dev):
    sm, _, _ = native.device_info()
    blocks, iters = sm * 16, 1 << 14
    native.probe_fp32(blocks, 256, dev)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        flop = native.probe_fp32(blocks, iters, dev)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flop / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def other_configs(torch, native, dev, flush, hbm_peak_gbs=6544.3, fp32_peak_tf=67.4):
    """The remaining BASELINE.json configurations, timed briefly on one GPU (kernel time, inputs resident)."""
    from sph_retina_b200 import synthetic as S
    from sph_retina_b200.sphdet.bbox.nms import sph_batched_nms_images
    from sph_retina_b200.sphdet.iou import SphOverlaps2D, fov_iou, sph2pob_efficient_iou, sph_iou
    from sph_retina_b200.sphdet.losses import Sph2PobIoULoss
    out = {}
    n = 1_000_000
    for box in ("bfov", "rbfov"):
        b1 = S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box=box, seed=0).to(dev)
        b2 = S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box=box, seed=1).to(dev)
        ms = quick(torch, lambda: sph2pob_efficient_iou(b1, b2, is_aligned=True), iters=20, flush=flush)
        native.set_dense(True)
        ms_dense = quick(torch, lambda: sph2pob_efficient_iou(b1, b2, is_aligned=True), flush=flush)
        native.set_dense(False)
        # configs[0] under the bench's own timing: ONE call, L2 flushed before it, CUDA events around it
        out["aligned_1M_%s" % box] = {"ms": ms, "pairs_per_s": n / ms * 1e3,
                                      "fp32_frac": n * FLOP_PER_PAIR / (ms * 1e-3) / 1e12 / fp32_peak_tf,
                                      "dense_ms": ms_dense, "dense_pairs_per_s": n / ms_dense * 1e3,
                                      "dense_fp32_frac": n * FLOP_PER_PAIR / (ms_dense * 1e-3) / 1e12 / fp32_peak_tf,
                                      "hbm_gbs": n * (36 if box == "bfov" else 44) / ms / 1e6,
                                      "timing": "single call, L2 flushed (256 MB write) before every call, CUDA events, median of 20"}
        if box == "bfov":
            for name, fn in (("sph_iou", sph_iou), ("fov_iou", fov_iou)):
                ms = quick(torch, lambda: fn(b1, b2, is_aligned=True), flush=flush)
                out["aligned_1M_%s" % name] = {"ms": ms, "pairs_per_s": n / ms * 1e3, "hbm_gbs": n * 36 / ms / 1e6}
    # the other two calculators of the reference on configs[0]'s boxes: naive_iou (planar IoU of the sph2pix boxes) and
    # unbiased_iou (exact spherical IoU, double precision per pair; CPU numpy in the reference, ~40 s per million pairs)
    from sph_retina_b200.sphdet.iou import naive_iou, unbiased_iou
    for box in ("bfov", "rbfov"):
        u1 = S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box=box, seed=0).to(dev)
        u2 = S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box=box, seed=1).to(dev)
        for name, fn in (("naive_iou", naive_iou), ("unbiased_iou", unbiased_iou)):
            ms = quick(torch, lambda: fn(u1, u2, is_aligned=True), iters=5, flush=flush)
            out["aligned_1M_%s_%s" % (name, box)] = {"ms": ms, "pairs_per_s": n / ms * 1e3}
    del u1, u2
    # the format conversions either side of the path (sphdet/bbox/box_formator.py), one launch each, 16 M boxes (HBM-bound)
    from sph_retina_b200.sphdet.bbox.box_formator import Planar2SphBoxTransform, Sph2PlanarBoxTransform
    nb = 16_000_000
    f4 = S.generate_boxes(nb, alpha_range=(1, 100), beta_range=(1, 100), box="bfov", seed=3).to(dev)
    to_planar, to_sph = Sph2PlanarBoxTransform('sph2pix', 4), Planar2SphBoxTransform('pix2sph', 4)
    ms = quick(torch, lambda: to_planar(f4), iters=5, flush=flush)
    pl = to_planar(f4)
    ms2 = quick(torch, lambda: to_sph(pl), iters=5, flush=flush)
    out["box_format_16M_bfov"] = {"sph2planar_ms": ms, "planar2sph_ms": ms2, "hbm_gbs_sph2planar": nb * 32 / ms / 1e6,
                                  "hbm_frac_of_measured_peak": nb * 32 / ms / 1e6 / hbm_peak_gbs}
    del f4, pl
    # the same calls at 16 M pairs (inputs 512 MB, larger than L2): the throughput once the ~10 us of launch + ramp are amortised
    n16 = 16_000_000
    c1 = S.generate_boxes(n16, alpha_range=(1, 100), beta_range=(1, 100), box="bfov", seed=0).to(dev)
    c2 = S.generate_boxes(n16, alpha_range=(1, 100), beta_range=(1, 100), box="bfov", seed=1).to(dev)
    ms = quick(torch, lambda: sph2pob_efficient_iou(c1, c2, is_aligned=True), iters=5, flush=flush)
    out["aligned_16M_bfov"] = {"ms": ms, "pairs_per_s": n16 / ms * 1e3}
    ms = quick(torch, lambda: sph_iou(c1, c2, is_aligned=True), iters=5, flush=flush)
    out["aligned_16M_sph_iou"] = {"ms": ms, "pairs_per_s": n16 / ms * 1e3, "hbm_gbs": n16 * 36 / ms / 1e6,
                                  "hbm_frac_of_measured_peak": n16 * 36 / ms / 1e6 / hbm_peak_gbs}
    del c1, c2
    pred, target = S.loss_pairs(200_000)
    pred, target = pred.to(dev), target.to(dev)
    L = Sph2PobIoULoss(mode="iou", reduction="sum")

    def fwd_bwd():
        p = pred.detach().requires_grad_(True)
        L(p, target).backward()
    with torch.no_grad():
        ms_f = quick(torch, lambda: L(pred, target), flush=flush)
    ms_fb = quick(torch, fwd_bwd, flush=flush)
    out["loss_200k_rbfov"] = {"fwd_ms": ms_f, "fwd_bwd_ms": ms_fb, "pairs_per_s_fwd_bwd": 200_000 / ms_fb * 1e3}
    # the same forward + backward captured in a CUDA graph (the step a training loop replays): what the device needs for
    # it once the ~10 eager autograd / allocator steps of the host are out of the way
    def graph_ms(loss_module):
        p_static = pred.detach().clone().requires_grad_(True)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                p_static.grad = None
                loss_module(p_static, target).backward()
        torch.cuda.current_stream().wait_stream(side)
        p_static.grad = None
        gl = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gl):
            loss_static = loss_module(p_static, target)
            loss_static.backward()
        ms_ = quick(torch, gl.replay, flush=flush)
        del gl
        return ms_
    try:
        ms_g = graph_ms(L)
        out["loss_200k_rbfov"].update({"fwd_bwd_graph_ms": ms_g, "pairs_per_s_fwd_bwd_graph": 200_000 / ms_g * 1e3,
                                       "fp32_frac_graph": 200_000 * 3 * FLOP_PER_PAIR / (ms_g * 1e-3) / 1e12 / fp32_peak_tf})
    except Exception as e:      # capture is an extra: never lose the line over it
        out["loss_200k_rbfov"]["fwd_bwd_graph_error"] = repr(e)
    # the other losses on the same OBBs: one launch each for loss + both gradients
    from sph_retina_b200.sphdet.losses import Sph2PobGDLoss, Sph2PobKFLoss, Sph2PobL1Loss
    other = {}
    for name, Lo in (("gwd", Sph2PobGDLoss("gwd", reduction="sum")), ("kld", Sph2PobGDLoss("kld", reduction="sum")),
                     ("kfiou", Sph2PobKFLoss(reduction="sum")), ("l1", Sph2PobL1Loss(reduction="sum"))):
        def fb(Lo=Lo):
            p = pred.detach().requires_grad_(True)
            Lo(p, target).backward()
        ms = quick(torch, fb, flush=flush)
        other[name] = {"fwd_bwd_ms": ms, "pairs_per_s_fwd_bwd": 200_000 / ms * 1e3}
        try:        # the device's share of the step: the same forward + backward replayed from a CUDA graph
            other[name]["fwd_bwd_graph_ms"] = graph_ms(Lo)
        except Exception as e:
            other[name]["fwd_bwd_graph_error"] = repr(e)
    out["other_losses_200k_rbfov"] = other
    # training targets of the head for the 16 images (assign -> PseudoSampler -> labels / weights / box targets): two C-ABI calls
    from sph_retina_b200.sphdet.assigners import SphMaxIoUAssigner
    from sph_retina_b200.sphdet.models.heads import get_targets_batch
    gts16, anc = S.assignment_batch(images=16)
    gts16, anc = gts16.to(dev), anc.to(dev)
    glist = list(gts16)
    llist = [torch.randint(0, 80, (g.size(0),), device=dev) for g in glist]
    asg = SphMaxIoUAssigner(0.5, 0.4, min_pos_iou=0, iou_calculator=dict(type='SphOverlaps2D', backend='sph2pob_efficient_iou', box_version=5))
    ms = quick(torch, lambda: get_targets_batch(anc, glist, llist, asg, 80, sync_counts=False), flush=flush)
    out["train_targets_16img"] = {"ms": ms, "anchors_per_s": 16 * anc.size(0) / ms * 1e3,
                                  "what": "SphMaxIoUAssigner + PseudoSampler + anchor_head._get_targets_single for 16 images x 98208 "
                                          "anchors x 32 GT: sphk_max_iou_assign + sphk_anchor_targets, no host sync"}
    boxes, scores, labels, image_ids = (t.to(dev) for t in S.nms_batch(64, 1000, 80))
    ms = quick(torch, lambda: sph_batched_nms_images(boxes, scores, labels, image_ids, 0.5), iters=5, flush=flush)
    ms_h = quick(torch, lambda: sph_batched_nms_images(boxes, scores, labels, image_ids, 0.5, num_images=64, num_classes=80,
                                                       max_per_segment=1000), iters=5, flush=flush)
    out["nms_64img_1000box_80cls"] = {"ms": ms, "images_per_s": 64 / ms * 1e3, "ms_with_shape_hints": ms_h,
                                      "images_per_s_with_shape_hints": 64 / ms_h * 1e3}
    from sph_retina_b200.sphdet.bbox.nms import sph_nms_image_blocks
    ms_d = quick(torch, lambda: sph_nms_image_blocks(boxes, scores, labels, 64, 80, 0.5, 100), iters=5, flush=flush)
    out["nms_64img_1000box_80cls"].update({"ms_device_pipeline": ms_d, "images_per_s_device_pipeline": 64 / ms_d * 1e3,
                                           "device_pipeline": "sphk_nms_images: per-image sort, per-segment NMS and per-image "
                                                              "ordering in three launches, top-100 per image, no host sync"})
    for calc_name in ("naive_iou", "unbiased_iou"):        # the calculators the reference's indoor360 / pandora configs give SphNMS
        ms_k = quick(torch, lambda: sph_nms_image_blocks(boxes, scores, labels, 64, 80, 0.5, 100, iou_calculator=calc_name), iters=5, flush=flush)
        out["nms_64img_1000box_80cls"]["ms_device_pipeline_" + calc_name] = ms_k
    ms = quick(torch, lambda: sph_batched_nms_images(boxes, scores, torch.zeros_like(labels), image_ids, 0.5), iters=5, flush=flush)
    zl = torch.zeros_like(labels)
    ms_d = quick(torch, lambda: sph_nms_image_blocks(boxes, scores, zl, 64, 1, 0.5, 100), iters=5, flush=flush)
    out["nms_64img_1000box_class_agnostic"] = {"ms": ms, "images_per_s": 64 / ms * 1e3, "ms_device_pipeline": ms_d,
                                               "images_per_s_device_pipeline": 64 / ms_d * 1e3}
    # configs[1] again, but all 16 images' GT in ONE call (legal whenever the anchors are shared by the images, as in
    # RetinaNet: SURVEY.md 3.1 "same anchors for every image"): [16*32, 98208] in two launches instead of 32
    gts, anchors = S.assignment_batch(IMAGES, GTS)
    gts, anchors = gts.to(dev), anchors.to(dev)
    calc = SphOverlaps2D('sph2pob_efficient_iou', 5)
    ms = quick(torch, lambda: calc(gts.view(-1, 5), anchors).view(IMAGES, GTS, -1), flush=flush)
    out["assign_16img_one_call"] = {"ms": ms, "pairs_per_s": IMAGES * GTS * anchors.size(0) / ms * 1e3}
    # configs[1] the drop-in way: one SphOverlaps2D call per image, as MaxIoUAssigner makes them (round 1's headline)
    ms = quick(torch, lambda: [calc(gts[i], anchors) for i in range(IMAGES)], iters=20, flush=flush)
    out["assign_16img_per_image_calls"] = {"ms": ms, "pairs_per_s": IMAGES * GTS * anchors.size(0) / ms * 1e3,
                                           "fp32_frac": IMAGES * GTS * anchors.size(0) * FLOP_PER_PAIR / (ms * 1e-3) / 1e12 / fp32_peak_tf,
                                           "what": "16 x k_iou_rows32 (one launch per call, chained by programmatic dependent launch)"}
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            [calc(gts[i], anchors) for i in range(IMAGES)]
        torch.cuda.current_stream().wait_stream(side)
        ga = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ga):
            keep_a = [calc(gts[i], anchors) for i in range(IMAGES)]
        ms = quick(torch, ga.replay, iters=20, flush=flush)
        out["assign_16img_per_image_calls"].update({"graph_ms": ms, "graph_pairs_per_s": IMAGES * GTS * anchors.size(0) / ms * 1e3})
        del ga, keep_a
    except Exception as e:
        out["assign_16img_per_image_calls"]["graph_error"] = repr(e)
    # the headline step (one call per image) with the 16 calls alternating between two CUDA streams: the drain of one
    # launch (its last CTAs, ~a quarter of a 31 us kernel) is covered by the start of the next
    s2 = [torch.cuda.Stream(), torch.cuda.Stream()]

    def two_streams():
        cur = torch.cuda.current_stream()
        for st in s2:
            st.wait_stream(cur)
        keep = []
        for i in range(IMAGES):
            with torch.cuda.stream(s2[i & 1]):
                keep.append(calc(gts[i], anchors))
        for st in s2:
            cur.wait_stream(st)
        return keep
    ms = quick(torch, two_streams, flush=flush)
    out["assign_16img_per_image_calls_two_streams"] = {"ms": ms, "pairs_per_s": IMAGES * GTS * anchors.size(0) / ms * 1e3}
    # the consumer of configs[1]: MaxIoUAssigner(pos 0.5, neg 0.3, min_pos 0) per image.  (a) the drop-in way:
    # matrix from the calculator + assign_wrt_overlaps on it; (b) SphMaxIoUAssigner: no matrix, two fused passes
    from sph_retina_b200.sphdet.assigners import SphMaxIoUAssigner
    asg = SphMaxIoUAssigner(0.5, 0.3, min_pos_iou=0.0, iou_calculator=calc)
    labels = torch.randint(0, 37, (IMAGES, GTS), device=dev)
    ms_a = quick(torch, lambda: [asg.assign_wrt_overlaps(calc(gts[i], anchors), labels[i]) for i in range(IMAGES)], iters=5, flush=flush)
    ms_b = quick(torch, lambda: [asg.assign(anchors, gts[i], gt_labels=labels[i]) for i in range(IMAGES)], iters=5, flush=flush)
    gl, ll = [gts[i] for i in range(IMAGES)], [labels[i] for i in range(IMAGES)]
    ms_c = quick(torch, lambda: asg.assign_batch(anchors, gl, ll), iters=5, flush=flush)
    out["assigner_16img"] = {"matrix_then_assign_ms": ms_a, "fused_per_image_ms": ms_b, "fused_batch_ms": ms_c,
                             "pairs_per_s_fused_batch": IMAGES * GTS * anchors.size(0) / ms_c * 1e3}
    # configs[0] as a stream of calls: 8 input sets (288 MB > L2, nothing re-used between calls), the 8 calls captured in
    # one CUDA graph -- no host launch gaps, no dirty L2 lines from a flush write
    sets = [(S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box="bfov", seed=10 + 2 * i).to(dev),
             S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box="bfov", seed=11 + 2 * i).to(dev)) for i in range(8)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for a, b in sets:
            sph2pob_efficient_iou(a, b, is_aligned=True)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        keep = [sph2pob_efficient_iou(a, b, is_aligned=True) for a, b in sets]
    ms = quick(torch, graph.replay) / len(sets)
    out["aligned_1M_bfov_stream_of_calls"] = {"ms_per_call": ms, "pairs_per_s": n / ms * 1e3,
                                              "how": "8 input sets (288 MB, larger than L2), 8 calls in one CUDA graph, no flush"}
    del graph, keep, sets
    # SURVEY.md 8f row 2 -- the regression branch of the head's loss on a whole batch: bbox_coder.decode on all
    # 16 x 98,208 anchors + Sph2PobIoULoss(weight = 0 for the negatives) + backward to the deltas
    from sph_retina_b200.sphdet.bbox.coder import DeltaXYWHASphBBoxCoder
    from sph_retina_b200.sphdet.losses import Sph2PobDecodedIoULoss
    anchors_b, deltas, target_b, weight = (t.to(dev) for t in S.head_loss_batch(IMAGES))
    coder = DeltaXYWHASphBBoxCoder(target_stds=(0.1, 0.1, 0.2, 0.2, 0.1))
    LD = Sph2PobDecodedIoULoss(mode="iou", loss_weight=1.0)
    npos = float((weight[:, 0] > 0).sum())

    def fused():
        d = deltas.detach().requires_grad_(True)
        LD.forward_decoded(coder, anchors_b, d, target_b, weight, avg_factor=npos).backward()

    def two_step():
        d = deltas.detach().requires_grad_(True)
        L_ = Sph2PobIoULoss(mode="iou", loss_weight=1.0)
        L_(coder.decode(anchors_b, d), target_b, weight, avg_factor=npos).backward()
    ms_fused, ms_two = quick(torch, fused, iters=5, flush=flush), quick(torch, two_step, iters=5, flush=flush)
    out["head_loss_16img"] = {"rows": anchors_b.size(0), "positives": int(npos), "fused_fwd_bwd_ms": ms_fused,
                              "decode_then_loss_fwd_bwd_ms": ms_two, "rows_per_s_fused": anchors_b.size(0) / ms_fused * 1e3}
    return out


def ncu_reference(workload):
    """Counter figures of the dominant kernel from the committed ncu capture (profiles/roofline_ncu.json, written by
    tools/profile_collect.py from the `ncu --set full` report of the same command)."""
    path = os.path.join(ROOT, "profiles", "roofline_ncu.json")
    try:
        return json.load(open(path)).get(workload)
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device. The product path is CUDA-only (no CPU fallback); "
                         "use --impl reference for the CPU arm.")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # one slice of the host cores per rank: the launch thread of a rank is then never queued behind another rank's
        # threads (torchrun does not pin)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = len(cores) // int(os.environ.get("LOCAL_WORLD_SIZE", world))
            if per >= 1:
                os.sched_setaffinity(0, cores[local_rank * per:(local_rank + 1) * per])
        except (AttributeError, OSError, ValueError):
            pass
        dist.init_process_group("nccl", device_id=dev)
    from sph_retina_b200 import _native as native
    from sph_retina_b200 import synthetic as S
    from sph_retina_b200.sphdet.iou import SphOverlaps2D, sph_max_overlaps

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    flush = L2Flusher(torch, dev)
    peaks, peaks_src = measured_peaks()
    result = {}
    if args.workload == "assign":
        gts_h, anchors_h = S.assignment_batch(IMAGES, GTS)
        if world > 1:       # every rank gets its own images (weak scaling, replicas: SURVEY.md 8e)
            gts_h = torch.stack([S.generate_boxes(GTS, alpha_range=(5, 120), beta_range=(5, 120), box="rbfov",
                                                  seed=100 + rank * IMAGES + i) for i in range(IMAGES)])
        gts, anchors = gts_h.to(dev), anchors_h.to(dev)
        calc = SphOverlaps2D('sph2pob_efficient_iou', 5)
        pairs_per_step = IMAGES * GTS * anchors.size(0)

        def step():
            return [calc(gts[i], anchors) for i in range(IMAGES)]

        def kernel_call():
            return calc(gts[0], anchors)
        kernel_name, kernels_per_step, pairs_per_kernel = "k_iou_rows32<5>", IMAGES, GTS * anchors.size(0)
        rows_k, cols_k = gts[0], anchors
        scaling = "weak"
        total_pairs_per_step = pairs_per_step * world
        hbm_bytes_per_kernel = GTS * anchors.size(0) * 4 + (GTS + anchors.size(0)) * 5 * 4
        workload = ("assign: pairwise Sph2Pob-efficient IoU, 32 RBFoV GT x 98208 anchors (512x1024, 9/loc), batch 16 "
                    "per GPU (replicas), full [32 x 98208] fp32 matrix written per image (BASELINE configs[1])")
    else:
        from sph_retina_b200.sharded import block_capacity, key_block, shard_bounds, sharded_max_overlaps
        A_h, G_h = sweep_inputs()
        lo, hi = shard_bounds(SWEEP_ANCHORS, world, rank)
        A_loc_h = A_h[lo:hi].contiguous()
        A_loc, G = A_loc_h.to(dev), G_h.to(dev)
        pairs_per_step = (hi - lo) * SWEEP_GTS
        total_pairs_per_step = SWEEP_ANCHORS * SWEEP_GTS

        def step():
            return sharded_max_overlaps(A_loc, G, SWEEP_ANCHORS, lo, anchors_are='bboxes1', exchange=args.exchange)

        _blk = key_block(SWEEP_ANCHORS, SWEEP_GTS, world, dev, fresh=True)
        _cap = block_capacity(SWEEP_ANCHORS, world)

        def kernel_call():      # the step without the exchange: k_box_pre + k_iou_pairwise2 into a communication block
            return native.iou_pairwise_keys("sph2pob_efficient", A_loc, G, row_base=lo, row_keys_out=_blk[:hi - lo], col_keys_out=_blk[_cap:])
        kernel_name, kernels_per_step, pairs_per_kernel = "k_iou_pairwise2<5,32,2> (RBFoV, 32-row tiles, separating-axis stage)", 1, pairs_per_step
        rows_k, cols_k = A_loc, G
        scaling = "strong"
        hbm_bytes_per_kernel = (A_loc.size(0) + SWEEP_GTS) * (5 * 4 + 8)
        workload = SWEEP_WORKLOAD

    # ---- timed region (device time, per-step events, L2 flushed between steps) --------------------
    # rank 0 watches every local GPU of the run; the other ranks do not run a sampler thread
    sampler = ClockSampler(list(range(int(os.environ.get("LOCAL_WORLD_SIZE", world)))) if rank == 0 else [], external=world > 1)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    l0 = native.launches
    sampler.start()
    ms = time_steps(torch, step, args.steps, 0, flush, barrier, lead_in=LEAD_IN if world > 1 else 0)
    clocks = sampler.stop()
    gpu_launches = native.launches - l0
    total_ms = max_over_ranks(sum(ms))
    ms_per_step = total_ms / args.steps
    value = total_pairs_per_step / (ms_per_step * 1e-3)

    # ---- the dominant kernel against the roofline that bounds it: FP32 / ALU instruction issue -------------------
    # (W = 512 flop per pair for ALL pairs, early-outs included; the early-out rate and the dense
    # throughput are printed next to it; the HBM view -- 8 B of keys per anchor, or 4 B per pair for the matrix -- is
    # in roofline_hbm and says nothing about this kernel)
    kernel_ms = statistics.median(time_steps(torch, kernel_call, 10, 3, flush, lambda: None))
    kernel_ms_ranks = [kernel_ms]
    if world > 1:       # the step ends when the slowest rank's kernel does: how far apart are they?
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = kernel_ms
        dist.all_reduce(t)
        kernel_ms_ranks = [float(x) for x in t.tolist()]
    if rank == 0:
        peak_tf = fp32_peak(torch, native, dev)
        tf = pairs_per_kernel * FLOP_PER_PAIR / (kernel_ms * 1e-3) / 1e12
        live = native.prefilter_live_pairs(rows_k, cols_k)
        native.set_dense(True)
        dense_ms = statistics.median(time_steps(torch, kernel_call, 5, 2, flush, lambda: None))
        native.set_dense(False)
        prof = ncu_reference(args.workload) or {}
        nominal = 148 * 128 * 2 * (peaks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
        result["roofline"] = {
            "bound": "fp32", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
            "traffic": prof.get("dram_bytes_per_launch"), "kernel": kernel_name, "kernel_ms": kernel_ms,
            "kernel_ms_per_rank": kernel_ms_ranks,
            "kernels_per_step": kernels_per_step, "kernel_share_of_step": kernel_ms * kernels_per_step / statistics.mean(ms),
            "flop_per_pair": FLOP_PER_PAIR, "pairs_per_launch": pairs_per_kernel,
            "peak_source": "FMA-chain probe (sphk_probe_fp32) on this GPU in this run; nominal 148 SM x 128 lanes x 2 x %.0f MHz = %.1f"
                           % (peaks.get("sm_max_mhz") or 1965.0, nominal),
            "peak_nominal": nominal, "frac_of_nominal": tf / nominal,
            "early_out_rate": 1.0 - live / float(pairs_per_kernel), "live_pairs": live,
            "dense_kernel_ms": dense_ms, "dense_pairs_per_s": pairs_per_kernel / (dense_ms * 1e-3),
            "dense_frac": pairs_per_kernel * FLOP_PER_PAIR / (dense_ms * 1e-3) / 1e12 / peak_tf,
            "ncu": prof.get("counters"), "ncu_source": prof.get("source"),
            "note": "achieved = ALL pairs of one launch x 512 flop (SURVEY.md 8d) / kernel_ms; kernel_ms = median of 10 launches of the "
                    "kernel alone (k_box_pre included), CUDA events on the launching stream, L2 flushed; early_out_rate = pairs the "
                    "prefilter proves disjoint (exact 0 in the reference); dense_* = the same launch with the early-outs disabled "
                    "(sphk_set_dense); ncu = issue-slot / pipe utilisation of the committed `ncu --set full` capture"}
        hbm_achieved = hbm_bytes_per_kernel / (kernel_ms * 1e-3) / 1e9
        result["roofline_hbm"] = {"bound": "hbm", "achieved": hbm_achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                  "frac": hbm_achieved / peaks["hbm_gbs"], "peak_source": peaks_src,
                                  "algorithmic_bytes_per_launch": hbm_bytes_per_kernel, "traffic": prof.get("dram_bytes_per_launch"),
                                  "note": "not the bound of this kernel (boxes in, 8 B of packed keys per box out)"}

    # ---- end to end: host buffers in, host result out, through the public API ----------------------
    e2e = None
    if args.no_e2e:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "note": "skipped (--no-e2e)"}
    elif args.workload == "sweep":
        n_loc = hi - lo
        A_pin, G_pin = A_loc_h.pin_memory(), G_h.pin_memory()
        A_d, G_d = torch.empty_like(A_pin, device=dev), torch.empty_like(G_pin, device=dev)
        amax_pin = torch.empty(n_loc, dtype=torch.float32).pin_memory()
        aarg_pin = torch.empty(n_loc, dtype=torch.int64).pin_memory()
        gmax_pin = torch.empty(SWEEP_GTS, dtype=torch.float32).pin_memory()
        garg_pin = torch.empty(SWEEP_GTS, dtype=torch.int64).pin_memory()

        from sph_retina_b200.sharded import HostSweep
        hs = HostSweep(SWEEP_ANCHORS, n_loc, SWEEP_GTS, 5, dev, exchange=args.exchange, **({'max_chunks': args.e2e_chunks, 'min_chunk_rows': 1} if args.e2e_chunks else {}))

        def e2e_step():
            # every rank: its shard of the anchors and the GT in pinned host memory -> HostSweep (H2D, kernels and D2H
            # pipelined over row chunks; per-GT keys exchanged between the ranks) -> the per-anchor result of its shard and
            # the global per-GT result in pinned host memory: across the ranks the host holds the whole result exactly once
            hs(A_pin, G_pin, lo, amax_pin, aarg_pin, gmax_pin, garg_pin)
        e2e_ms = time_steps(torch, e2e_step, max(5, args.steps // 2), 3, flush, barrier, lead_in=LEAD_IN if world > 1 else 0)
        t = max_over_ranks(statistics.mean(e2e_ms))
        e2e = {"value": total_pairs_per_step / (t * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": SWEEP_ANCHORS * 20 + world * SWEEP_GTS * 20,
               "d2h_bytes_per_step": SWEEP_ANCHORS * 12 + world * SWEEP_GTS * 12,
               "ms_per_step": t, "api": "sph_retina_b200.sharded.HostSweep(...)(anchors_pinned, gts_pinned, offset, out...)",
               "chunks_per_rank": hs.chunks,
               "copies": "per rank and step: H2D of its anchor shard (n/N x 20 B) + the GT (20 kB) from pinned memory; D2H of the "
                         "(max fp32, argmax int64) of its anchors + the global per-GT result into pinned memory; copies run on "
                         "a copy stream under the kernels of the neighbouring row chunks; byte counts are whole-job sums over the ranks"}
        if world == 1 and rank == 0:
            # the same result without the fused reduction: the full 4.3 GB matrix does not leave the device in the
            # reference either (MaxIoUAssigner reduces it on the GPU); listed for scale
            e2e["matrix_bytes_not_moved"] = SWEEP_ANCHORS * SWEEP_GTS * 4
    else:
        gts_pin, anchors_pin = gts_h.pin_memory(), anchors_h.pin_memory()
        out_pin = torch.empty((IMAGES, GTS, anchors_h.size(0)), dtype=torch.float32).pin_memory()
        gts_d, anchors_d = torch.empty_like(gts_pin, device=dev), torch.empty_like(anchors_pin, device=dev)
        side = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))

        def e2e_step():
            # H2D of this step's inputs on the launching stream, then the 16 calculator calls alternate between two
            # side streams, each call followed by the D2H of its matrix on its own stream: the kernel of image i + 1 runs
            # while the copy engine drains image i.  The step is bound by the 201 MB D2H over PCIe.
            gts_d.copy_(gts_pin, non_blocking=True)
            anchors_d.copy_(anchors_pin, non_blocking=True)
            cur = torch.cuda.current_stream(dev)
            for st in side:
                st.wait_stream(cur)
            for i in range(IMAGES):
                with torch.cuda.stream(side[i & 1]):
                    out_pin[i].copy_(calc(gts_d[i], anchors_d), non_blocking=True)
            for st in side:
                cur.wait_stream(st)
        e2e_ms = time_steps(torch, e2e_step, max(3, args.steps // 2), 2, flush, barrier)
        t = max_over_ranks(statistics.mean(e2e_ms))
        e2e = {"value": total_pairs_per_step / (t * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": gts_pin.numel() * 4 + anchors_pin.numel() * 4, "d2h_bytes_per_step": out_pin.numel() * 4,
               "ms_per_step": t, "api": "SphOverlaps2D('sph2pob_efficient_iou', 5)(gt, anchors) x 16, pinned host in/out",
               "streams": "calls alternate between 2 CUDA streams (compute of one image overlaps the D2H of the previous one)"}

    exchange_route_name = "none (replicas)"
    if args.workload == "sweep":
        from sph_retina_b200.sharded import exchange_route
        exchange_route_name = exchange_route()
        if exchange_route_name == "peer" and args.exchange in ("nccl", "peer-pull"):
            exchange_route_name = args.exchange
        exchange_route_name = {"peer": "peer: the compute kernel stores its anchors' keys into the peers' symmetric buffers over NVLink "
                                       "while it runs (sphk_iou_pairwise_keys_push); flag handshake + local unpack in one launch "
                                       "(sphk_unpack_peer_keys); no collective",
                               "peer-pull": "peer-pull: keys read from the owners' symmetric buffers over NVLink inside the unpack launch "
                                            "(sphk_unpack_peer_keys), no collective", "nccl": "nccl: one all_gather_into_tensor + unpack launch",
                               "single": "single GPU: the unpack launch reads the local block"}[exchange_route_name]
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload,
                       "api": ("sharded_max_overlaps -> %s (Sph2Pob-efficient, box_version 5)" % (
                                   "sphk_iou_pairwise_keys + sphk_unpack_gathered_keys" if world == 1 or exchange_route_name.startswith("nccl")
                                   else ("sphk_iou_pairwise_keys_push + sphk_unpack_peer_keys" if exchange_route_name.startswith("peer:")
                                         else "sphk_iou_pairwise_keys + sphk_unpack_peer_keys")))
                              if args.workload == "sweep" else "SphOverlaps2D('sph2pob_efficient_iou', box_version=5)",
                       "pairs_per_step": total_pairs_per_step,
                       "l2": "256 MB written between timed steps (L2 flush)",
                       "timing": "CUDA events per step on the launching stream, sum over K steps, max over ranks"
                                 + ("; %d untimed lead-in steps follow the barrier (host threads ahead of the GPUs before the first "
                                    "timed handshake)" % LEAD_IN if world > 1 else ""),
                       "exchange": exchange_route_name,
                       "collectives_per_step": (1 if exchange_route_name == "nccl" else 0)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches,
        }
        line.update(result)
        if world == 1 and not args.no_cpu and args.workload == "sweep":
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            fn, kind, what = load_cpu_reference()
            rows, best = 8192, None
            with torch.no_grad():
                for k in range(2):
                    t0 = time.perf_counter()
                    cpu_sweep_slice(fn, A_h, G_h, k * rows, rows)
                    dt = time.perf_counter() - t0
                    best = dt if best is None else min(best, dt)
            line["cpu_baseline"] = {"value": rows * SWEEP_GTS / best, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": "%d anchors x %d GT = %.1f M pairs of the sweep (IoU matrix + both max/argmax), best of 2 "
                                              "windows, %.1f s" % (rows, SWEEP_GTS, rows * SWEEP_GTS / 1e6, best),
                                    "what": what}
            c = cpu_c_port_sweep(A_h, G_h)
            if c:
                line["cpu_baseline_c_openmp"] = {"value": c, "unit": UNIT, "cores": cores, "kind": "port",
                                                 "sample": "16384 anchors x 1024 GT (16.8 M pairs), best of 3",
                                                 "what": "oracle/sph_oracle.c: float64 scalar restatement, OpenMP"}
        if world == 1 and not args.no_extras:
            try:
                line["other_configs"] = other_configs(torch, native, dev, flush, peaks["hbm_gbs"], line["roofline"]["peak"])
            except Exception as e:   # the headline line must still be printed
                line["other_configs"] = {"error": repr(e)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sweep", choices=["sweep", "assign"])
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "peer-pull", "nccl"],
                    help="sweep, N > 1: how the ranks' keys are exchanged (sph_retina_b200/sharded.py)")
    ap.add_argument("--e2e-chunks", type=int, default=0, help="sweep e2e leg: row chunks per rank of the copy / compute pipeline (0 = HostSweep's default)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs and keep the run short (ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
