"""TEST INFRASTRUCTURE ONLY -- freeze golden vectors from the REAL reference.

Run in the build container (needs /root/reference):

    python oracle/make_golden.py

Writes tests/golden/*.npz.  Inputs are seeded; outputs come from the
reference's own Python code (oracle/ref_harness.py explains the two stubs)
run in float32 ("as shipped") and in float64 ("truth").  The fixtures travel
to the GPU box; the reference does not.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _ref64(fn, *tensors, **kw):
    with rh.float64_mode():
        return fn(*[t.double() for t in tensors], **kw)


def _np(t):
    return t.detach().cpu().numpy()


def make_pairs(R, n, box, seed):
    """Random pairs + the hard cases the reference's jitters exist for."""
    D = 4 if box == "bfov" else 5
    torch.manual_seed(seed)
    b1 = R.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), dtype="float", box=box)
    b2 = R.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), dtype="float", box=box)
    q = n // 8
    sig = torch.tensor([6, 6, 6, 6, 10.0])[:D]
    # near-coincident (loss style), three noise scales
    b2[0:q] = (b1[0:q] + torch.randn(q, D) * sig).clamp(min=1)
    b2[q:2 * q] = (b1[q:2 * q] + torch.randn(q, D) * sig * 0.1).clamp(min=1)
    b2[2 * q:2 * q + 64] = (b1[2 * q:2 * q + 64] + torch.randn(64, D) * 1e-3).clamp(min=1)
    # identical boxes, integer-valued boxes (jitter masks fire), one shared column
    b2[2 * q + 64:2 * q + 128] = b1[2 * q + 64:2 * q + 128]
    k = 2 * q + 128
    bi = R.generate_boxes(128, alpha_range=(1, 100), beta_range=(1, 100), dtype="int", box=box)
    bj = R.generate_boxes(128, alpha_range=(1, 100), beta_range=(1, 100), dtype="int", box=box)
    b1[k:k + 128], b2[k:k + 128] = bi, bj
    k += 128
    b2[k:k + 64, 2] = b1[k:k + 64, 2]          # same alpha only
    k += 64
    b2[k:k + 64, :2] = b1[k:k + 64, :2]        # same centre, different size
    k += 64
    # poles, seams, oversize and zero boxes, antipodes
    special = torch.tensor([
        [0.0, 0.0, 10, 10, 0], [360.0, 180.0, 10, 10, 0], [0.0, 90.0, 404, 120, 0], [180.0, 90.0, 200, 200, 30],
        [0.0, 0.0, 0, 0, 0], [359.9999, 0.0001, 1, 1, -90], [10.0, 90.0, 20, 20, 0], [190.0, 90.0, 20, 20, 0],
        [45.0, 45.0, 30, 60, 45], [45.0, 45.0, 60, 30, -45], [90.0, 1e-3, 50, 50, 10], [270.0, 1e-3, 50, 50, 80],
    ])[:, :D]
    m = special.size(0)
    b1[k:k + m] = special
    b2[k:k + m] = special.roll(1, dims=0)
    b1[k + m:k + 2 * m] = special
    b2[k + m:k + 2 * m] = special
    return b1.contiguous(), b2.contiguous()


def golden_aligned(R):
    for box in ("bfov", "rbfov"):
        b1, b2 = make_pairs(R, 4096, box, seed=11 if box == "bfov" else 12)
        out = dict(b1=_np(b1), b2=_np(b2))
        for tr in ("efficient", "standard"):
            fn = getattr(R, "sph2pob_%s_iou" % tr)
            for mode in ("iou", "iof"):
                out["%s_%s_f32" % (tr, mode)] = _np(fn(b1, b2, mode=mode, is_aligned=True))
                out["%s_%s_f64" % (tr, mode)] = _np(_ref64(fn, b1, b2, mode=mode, is_aligned=True))
            for edge in ("chord", "tangent"):
                out["%s_%s_f64" % (tr, edge)] = _np(_ref64(fn, b1, b2, is_aligned=True, rbb_edge=edge))
                out["%s_%s_f32" % (tr, edge)] = _np(fn(b1, b2, is_aligned=True, rbb_edge=edge))
            out["%s_project_f64" % tr] = _np(_ref64(fn, b1, b2, is_aligned=True, rbb_angle="project"))
            out["%s_project_f32" % tr] = _np(fn(b1, b2, is_aligned=True, rbb_angle="project"))
        if box == "bfov":
            for k in ("sph", "fov"):
                fn = getattr(R, "%s_iou" % k)
                out["%s_f32" % k] = _np(fn(b1, b2, is_aligned=True))
                out["%s_f64" % k] = _np(_ref64(fn, b1, b2, is_aligned=True))
        # the OBBs after transform + both jitters (first 512 pairs), for unit-level checks
        for tr, tf in (("efficient", R.eff.sph2pob_efficient), ("standard", R.std.sph2pob_standard)):
            with rh.float64_mode():
                j1, j2 = R.jiter_spherical_bboxes(b1[:512].double().clone(), b2[:512].double().clone())
                o1, o2 = tf(j1, j2, rbb_angle_version="rad")
                o1, o2 = R.jiter_rotated_bboxes(o1, o2)
            out["%s_obb1_f64" % tr], out["%s_obb2_f64" % tr] = _np(o1), _np(o2)
        np.savez_compressed(os.path.join(OUT, "aligned_%s.npz" % box), **out)
        print("aligned", box, {k: v.shape for k, v in out.items() if k.endswith("iou_f64")})


def golden_kat(R):
    """The 7 hand-picked pairs of tests/test_all_ious.py:243-261 (printed there, frozen here)."""
    g1 = torch.tensor([[40, 50, 35, 55], [30, 60, 60, 60], [50, -78, 25, 46], [30, 75, 30, 60],
                       [40, 70, 25, 30], [30, 75, 30, 30], [30, 60, 60, 60]]).float()
    g2 = torch.tensor([[35, 20, 37, 50], [55, 40, 60, 60], [30, -75, 26, 45], [60, 40, 60, 60],
                       [60, 85, 30, 30], [60, 55, 40, 50], [60, 60, 60, 60]]).float()
    b1, b2 = R.box_formator.geo2sph(g1), R.box_formator.geo2sph(g2)
    out = dict(b1=_np(b1), b2=_np(b2))
    for name in ("sph2pob_efficient_iou", "sph2pob_standard_iou", "sph_iou", "fov_iou"):
        out[name] = _np(getattr(R, name)(b1, b2, is_aligned=True))
    # tests/test_nms.py:6-27 fixture
    boxes = torch.tensor([[20, 40, 30, 30], [20, 40, 30, 30], [22, 38, 32, 28], [60, 60, 10, 10], [60, 60, 10, 10],
                          [60, 60, 10, 10], [60, 60, 10, 10], [30, 10, 10, 10], [30, 45, 45, 45], [80, 20, 66, 66]]).float()
    scores = torch.tensor([0.9, 0.8, 0.7, 0.6, 0.5, 0.85, 0.75, 0.65, 0.4, 0.3])
    idxs = torch.tensor([1, 1, 1, 1, 1, 2, 2, 2, 3, 3])
    dets, keep = R.SphNMS("sph2pob_efficient")(boxes, scores, idxs, dict(type="nms", iou_threshold=0.5))
    out.update(nms_boxes=_np(boxes), nms_scores=_np(scores), nms_idxs=_np(idxs), nms_dets=_np(dets), nms_keep=_np(keep))
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **out)
    print("kat", out["sph2pob_efficient_iou"], out["nms_keep"])


def anchors_512x1024():
    """mmdet AnchorGenerator(octave_base_scale=4, scales_per_octave=3, ratios=[.5,1,2],
    strides=[8..128]) on 512x1024 -> (theta, phi, alpha, beta, 0); SURVEY.md 8(d) config #2
    (mmdet/core/anchor/anchor_generator.py:98-101,169-192; sphdet/bbox/box_formator.py:85-92)."""
    H, W = 512, 1024
    out = []
    for s in (8, 16, 32, 64, 128):
        fh, fw = H // s, W // s
        ys, xs = torch.meshgrid(torch.arange(fh) * s, torch.arange(fw) * s, indexing="ij")
        base = []
        for r in (0.5, 1.0, 2.0):
            for i in range(3):
                sc = 4 * 2 ** (i / 3)
                base.append((s * sc / r ** 0.5, s * sc * r ** 0.5))
        base = torch.tensor(base)                       # [9,2] (w,h): ratio-major like mmdet
        cx = xs.reshape(-1, 1).float().expand(-1, 9)
        cy = ys.reshape(-1, 1).float().expand(-1, 9)
        w = base[:, 0][None, :].expand_as(cx)
        h = base[:, 1][None, :].expand_as(cx)
        a = torch.stack([cx / W * 360, cy / H * 180, w / W * 360, h / H * 180, torch.zeros_like(cx)], dim=-1)
        out.append(a.reshape(-1, 5))
    return torch.cat(out).float().contiguous()


def golden_pairwise(R):
    out = {}
    for box in ("bfov", "rbfov"):
        torch.manual_seed(21)
        rows = R.generate_boxes(48, alpha_range=(5, 120), beta_range=(5, 120), dtype="float", box=box)
        cols = R.generate_boxes(160, alpha_range=(5, 120), beta_range=(5, 120), dtype="float", box=box)
        cols[:16] = rows[:16]                                  # exact duplicates across the two sets
        cols[16:32] = (rows[16:32] + torch.randn_like(rows[16:32])).clamp(min=1)
        out["%s_rows" % box], out["%s_cols" % box] = _np(rows), _np(cols)
        fn = R.sph2pob_efficient_iou
        out["%s_rc_f32" % box] = _np(fn(rows, cols))
        out["%s_rc_f64" % box] = _np(_ref64(fn, rows, cols))
        out["%s_cr_f64" % box] = _np(_ref64(fn, cols, rows))
        out["%s_cr_f32" % box] = _np(fn(cols, rows))
    # config #2 orientation through the registry class: GT rows x anchor cols (a strided anchor sample)
    anc = anchors_512x1024()[::37].contiguous()
    torch.manual_seed(100)
    gt = R.generate_boxes(32, alpha_range=(5, 120), beta_range=(5, 120), gamma_range=(-90, 90), dtype="float", box="rbfov")
    calc = R.SphOverlaps2D("sph2pob_efficient_iou", 5)
    out["assign_gt"], out["assign_anchors"] = _np(gt), _np(anc)
    out["assign_f32"] = _np(calc(gt, anc))
    with rh.float64_mode():
        out["assign_f64"] = _np(calc(gt.double(), anc.double()))
    np.savez_compressed(os.path.join(OUT, "pairwise.npz"), **out)
    print("pairwise", out["assign_f64"].shape, float((out["assign_f64"] > 0).mean()))


def golden_loss(R):
    for box in ("bfov", "rbfov"):
        D = 4 if box == "bfov" else 5
        n = 2048
        torch.manual_seed(31)
        t = R.generate_boxes(n, alpha_range=(5, 100), beta_range=(5, 100), dtype="float", box=box)
        p = (t + torch.randn(n, D) * torch.tensor([6, 6, 6, 6, 10.0])[:D]).clamp(min=1)
        p[:32] = t[:32]                                  # identical rows
        t[32:64] = 0                                     # negatives carry all-zero targets (sph_retina_head.py:252-265)
        p[64:96] = (t[64:96] + torch.randn(32, D) * 0.05).clamp(min=1)
        w1 = (torch.rand(n) > 0.3).float()
        w2 = torch.rand(n, D)
        out = dict(pred=_np(p), target=_np(t), w1=_np(w1), w2=_np(w2))
        for mode in ("iou", "giou", "diou", "ciou"):
            for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
                L = R.Sph2PobIoULoss(mode=mode, reduction="sum")
                pp = p.to(dt).clone().requires_grad_(True)
                tt = t.to(dt).clone().requires_grad_(True)
                if dt == torch.float64:
                    with rh.float64_mode():
                        el = L(pp, tt, reduction_override="none")
                        el.sum().backward()
                else:
                    el = L(pp, tt, reduction_override="none")
                    el.sum().backward()
                out["%s_loss_%s" % (mode, tag)] = _np(el)
                out["%s_gpred_%s" % (mode, tag)] = _np(pp.grad)
                out["%s_gtarget_%s" % (mode, tag)] = _np(tt.grad)
        # reduction / weight / avg_factor combinations (float64 scalars)
        with rh.float64_mode():
            L = R.Sph2PobIoULoss(mode="iou", loss_weight=2.0)
            pd, td = p.double(), t.double()
            out["red_mean"] = _np(L(pd, td))
            out["red_w1_avg"] = _np(L(pd, td, w1.double(), avg_factor=123.0))
            out["red_w2"] = _np(L(pd, td, w2.double()))
            out["red_w1_sum"] = _np(L(pd, td, w1.double(), reduction_override="sum"))
            out["red_zero_w"] = _np(L(pd, td, torch.zeros(n).double()))
        np.savez_compressed(os.path.join(OUT, "loss_%s.npz" % box), **out)
        print("loss", box, float(out["iou_loss_f64"].mean()))


OTHER_LOSS_VARIANTS = {
    # name: (class name, constructor kwargs)            -- SURVEY.md 8f row 3
    "gwd": ("Sph2PobGDLoss", dict(loss_type="gwd")),
    "kld": ("Sph2PobGDLoss", dict(loss_type="kld")),
    "jd": ("Sph2PobGDLoss", dict(loss_type="jd")),
    "kld_symmax": ("Sph2PobGDLoss", dict(loss_type="kld_symmax")),
    "kld_symmin": ("Sph2PobGDLoss", dict(loss_type="kld_symmin")),
    "gwd_sqrt_tau1_raw": ("Sph2PobGDLoss", dict(loss_type="gwd", fun="sqrt", tau=1.0, normalize=False)),
    "gwd_alpha2_tau2": ("Sph2PobGDLoss", dict(loss_type="gwd", alpha=2.0, tau=2.0)),
    "kld_none_tau1_nosqrt": ("Sph2PobGDLoss", dict(loss_type="kld", fun="none", tau=1.0, sqrt=False)),
    "jd_nosqrt": ("Sph2PobGDLoss", dict(loss_type="jd", sqrt=False, alpha=0.5)),
    "kf": ("Sph2PobKFLoss", dict()),
    "kf_ln": ("Sph2PobKFLoss", dict(fun="ln")),
    "kf_exp": ("Sph2PobKFLoss", dict(fun="exp")),
    "l1": ("Sph2PobL1Loss", dict()),
    "l1_swap": ("Sph2PobL1Loss", dict(swap=True)),
    "l1_modulus": ("Sph2PobL1Loss", dict(angle_modifier="modulus")),
    "l1_plain": ("Sph2PobL1Loss", dict(encode=False)),
}


def golden_other_losses(R):
    """Sph2PobGDLoss / Sph2PobKFLoss / Sph2PobL1Loss through the reference's own subclasses and Sph2PobTransfrom decorator
    (GDLoss / KFLoss underneath = oracle/mmrotate_losses.py, the absent mmrotate 0.3.2)."""
    for box in ("bfov", "rbfov"):
        D = 4 if box == "bfov" else 5
        n = 1024
        torch.manual_seed(53)
        t = R.generate_boxes(n, alpha_range=(5, 100), beta_range=(5, 100), dtype="float", box=box)
        p = (t + torch.randn(n, D) * torch.tensor([6, 6, 6, 6, 10.0])[:D]).clamp(min=1)
        p[:16] = t[:16]                                  # identical rows
        t[16:32] = 0                                     # the head's negatives carry all-zero targets
        p[32:64] = (t[32:64] + torch.randn(32, D) * 0.05).clamp(min=1)
        w1 = (torch.rand(n) > 0.3).float()
        w2 = torch.rand(n, D) * (torch.rand(n, 1) > 0.3).float()
        out = dict(pred=_np(p), target=_np(t), w1=_np(w1), w2=_np(w2), variants_json=np.array(json.dumps(OTHER_LOSS_VARIANTS)))
        for name, (cls, kw) in OTHER_LOSS_VARIANTS.items():
            for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
                L = getattr(R, cls)(reduction="sum", **kw)
                pp = p.to(dt).clone().requires_grad_(True)
                tt = t.to(dt).clone().requires_grad_(True)
                if dt == torch.float64:
                    with rh.float64_mode():
                        el = L(pp, tt, reduction_override="none")
                        el.sum().backward()
                else:
                    el = L(pp, tt, reduction_override="none")
                    el.sum().backward()
                out["%s_loss_%s" % (name, tag)] = _np(el)
                out["%s_gpred_%s" % (name, tag)] = _np(pp.grad)
                out["%s_gtarget_%s" % (name, tag)] = _np(tt.grad)
            with rh.float64_mode():
                L = getattr(R, cls)(loss_weight=2.0, **kw)
                pd, td = p.double(), t.double()
                out[name + "_red_mean"] = _np(L(pd, td))
                out[name + "_red_w2_avg"] = _np(L(pd, td, w2.double(), avg_factor=77.0))
                out[name + "_red_w2_sum"] = _np(L(pd, td, w2.double(), reduction_override="sum"))
                if cls != "Sph2PobL1Loss":                # a 1-D weight cannot broadcast against the [n, 5] L1 loss
                    out[name + "_red_w1_avg"] = _np(L(pd, td, w1.double(), avg_factor=123.0))
        np.savez_compressed(os.path.join(OUT, "other_losses_%s.npz" % box), **out)
        print("other losses", box, {k: float(out[k + "_loss_f64"].mean()) for k in OTHER_LOSS_VARIANTS})


def golden_naive(R):
    """naive_iou (sph_iou_api.py:181-198) through the reference's own function and box formator (mmcv's two planar ops
    underneath are stand-ins: oracle/ref_harness.py), aligned + N x M, and SphNMS('naive_iou') keep sets."""
    out = {}
    for box in ("bfov", "rbfov"):
        torch.manual_seed(61)
        n = 2048
        b1 = R.generate_boxes(n, alpha_range=(1, 120), beta_range=(1, 120), dtype="float", box=box)
        b2 = (b1 + torch.randn_like(b1) * torch.tensor([8, 8, 8, 8, 15.0])[:b1.size(1)]).clamp(min=0.5)
        b2[:64] = R.generate_boxes(64, alpha_range=(1, 120), beta_range=(1, 120), dtype="float", box=box)   # unrelated pairs
        b2[64:96] = b1[64:96]                                                                                # identical pairs
        out[box + "_b1"], out[box + "_b2"] = _np(b1), _np(b2)
        out[box + "_aligned_f32"] = _np(R.naive_iou(b1, b2, is_aligned=True))
        out[box + "_aligned_f64"] = _np(_ref64(R.naive_iou, b1, b2, is_aligned=True))
        rows, cols = b1[:37], b2[:301]
        out[box + "_rc_f32"] = _np(R.naive_iou(rows, cols))
        out[box + "_rc_f64"] = _np(_ref64(R.naive_iou, rows, cols))
        torch.manual_seed(62)
        seeds = R.generate_boxes(80, alpha_range=(5, 60), beta_range=(5, 60), dtype="float", box=box)
        boxes = (seeds.repeat(5, 1) + torch.randn(400, seeds.size(1)) * 2).clamp(min=1)
        boxes[350:400] = boxes[300:350]
        scores, idxs = torch.rand(400), torch.randint(0, 6, (400,))
        for thr in (0.3, 0.5):
            dets, keep = R.SphNMS("naive_iou")(boxes, scores, idxs, dict(type="nms", iou_threshold=thr, max_num=150))
            out["%s_keep_thr%d" % (box, int(thr * 10))] = _np(keep)
        out["%s_pair_iou_f64" % box] = _np(_ref64(R.naive_iou, boxes, boxes)).astype(np.float32)
        out["%s_boxes" % box], out["%s_scores" % box], out["%s_idxs" % box] = _np(boxes), _np(scores), _np(idxs)
    np.savez_compressed(os.path.join(OUT, "naive.npz"), **out)
    print("naive", {k: (v.shape, float(np.nanmean(v))) for k, v in out.items() if "aligned_f64" in k or "keep" in k})


def golden_unbiased(R):
    """unbiased_iou (sph_iou_api.py:103-125; numpy classes unbiased_iou_bfov.Sph / unbiased_iou_rbfov.Sph) run on float64
    copies of the float32 boxes ("truth": the method needs double, unbiased_iou_bfov.py:187) and as shipped on float32,
    aligned + N x M, and SphNMS('unbiased_iou') keep sets from the float64 run."""
    out = {}
    for box in ("bfov", "rbfov"):
        torch.manual_seed(71)
        n = 1536
        b1 = R.generate_boxes(n, alpha_range=(1, 120), beta_range=(1, 120), dtype="float", box=box)
        b2 = (b1 + torch.randn_like(b1) * torch.tensor([8, 8, 8, 8, 15.0])[:b1.size(1)]).clamp(min=0.5)
        b2[:, 0].clamp_(0, 360); b2[:, 1].clamp_(0, 180); b2[:, 2:4].clamp_(max=179)
        b2[:64] = R.generate_boxes(64, alpha_range=(1, 120), beta_range=(1, 120), dtype="float", box=box)   # unrelated pairs
        b2[64:96] = b1[64:96]                                                                                # identical pairs
        b2[96:128, 2:4] = b1[96:128, 2:4] * 0.4                                                              # nested boxes
        b2[96:128, :2] = b1[96:128, :2] + 0.5
        out[box + "_b1"], out[box + "_b2"] = _np(b1), _np(b2)
        out[box + "_aligned_f32"] = _np(R.unbiased_iou(b1, b2, is_aligned=True))
        out[box + "_aligned_f64"] = _np(R.unbiased_iou(b1.double(), b2.double(), is_aligned=True)).astype(np.float64)
        rows, cols = b1[:23], b2[:201]
        out[box + "_rc_f32"] = _np(R.unbiased_iou(rows, cols))
        out[box + "_rc_f64"] = _np(R.unbiased_iou(rows.double(), cols.double())).astype(np.float64)
        torch.manual_seed(72)
        seeds = R.generate_boxes(60, alpha_range=(5, 60), beta_range=(5, 60), dtype="float", box=box)
        boxes = (seeds.repeat(5, 1) + torch.randn(300, seeds.size(1)) * 2).clamp(min=1)
        boxes[:, 1].clamp_(1, 179)
        scores, idxs = torch.rand(300), torch.randint(0, 5, (300,))
        for thr in (0.3, 0.5):
            _, keep = R.SphNMS("unbiased_iou")(boxes.double(), scores.double(), idxs, dict(type="nms", iou_threshold=thr, max_num=120))
            out["%s_keep_thr%d" % (box, int(thr * 10))] = _np(keep)
        out["%s_pair_iou_f64" % box] = _np(R.unbiased_iou(boxes.double(), boxes.double())).astype(np.float32)
        out["%s_boxes" % box], out["%s_scores" % box], out["%s_idxs" % box] = _np(boxes), _np(scores), _np(idxs)
    np.savez_compressed(os.path.join(OUT, "unbiased.npz"), **out)
    print("unbiased", {k: (v.shape, float(np.nanmean(v))) for k, v in out.items() if "aligned_f" in k or "keep" in k})


def golden_legacy(R):
    """sph2pob_legacy_iou (sph_iou_api.py:91-92; BFoV only): the pair set of golden_aligned, every mode / edge option, an
    R x C call, and the 7 hand-picked pairs of tests/test_all_ious.py:243-261."""
    b1, b2 = make_pairs(R, 4096, "bfov", seed=13)
    fn = R.sph2pob_legacy_iou
    out = dict(b1=_np(b1), b2=_np(b2))
    for mode in ("iou", "iof"):
        out["%s_f32" % mode] = _np(fn(b1, b2, mode=mode, is_aligned=True))
        out["%s_f64" % mode] = _np(_ref64(fn, b1, b2, mode=mode, is_aligned=True))
    for edge in ("chord", "tangent"):
        out["%s_f32" % edge] = _np(fn(b1, b2, is_aligned=True, rbb_edge=edge))
        out["%s_f64" % edge] = _np(_ref64(fn, b1, b2, is_aligned=True, rbb_edge=edge))
    out["rc_f32"] = _np(fn(b1[:29], b2[:333]))
    out["rc_f64"] = _np(_ref64(fn, b1[:29], b2[:333]))
    g1 = torch.tensor([[40, 50, 35, 55], [30, 60, 60, 60], [50, -78, 25, 46], [30, 75, 30, 60],
                       [40, 70, 25, 30], [30, 75, 30, 30], [30, 60, 60, 60]]).float()
    g2 = torch.tensor([[35, 20, 37, 50], [55, 40, 60, 60], [30, -75, 26, 45], [60, 40, 60, 60],
                       [60, 85, 30, 30], [60, 55, 40, 50], [60, 60, 60, 60]]).float()
    k1, k2 = R.box_formator.geo2sph(g1), R.box_formator.geo2sph(g2)
    out.update(kat_b1=_np(k1), kat_b2=_np(k2), kat_iou=_np(fn(k1, k2, is_aligned=True)))
    np.savez_compressed(os.path.join(OUT, "legacy.npz"), **out)
    print("legacy", out["kat_iou"], float(np.nanmean(out["iou_f64"])), int(np.isnan(out["iou_f64"]).sum()), int(np.isnan(out["iou_f32"]).sum()))


def golden_legacy_loss(R):
    """SphIoULossLegacy = Sph2PobTransfrom()(mmrotate RotatedIoULoss) (sph2pob_iou_loss.py:199-216) on the rows of golden_loss:
    per-row losses and both gradients for the three modes, fp32 and fp64 runs, plus the reductions."""
    out = {}
    for box in ("bfov", "rbfov"):
        g = np.load(os.path.join(OUT, "loss_%s.npz" % box))
        p, t, w1, w2 = (torch.from_numpy(g[k]) for k in ("pred", "target", "w1", "w2"))
        for mode in ("log", "linear", "square"):
            for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
                L = R.SphIoULossLegacy(mode=mode, reduction="sum")
                pp, tt = p.to(dt).clone().requires_grad_(True), t.to(dt).clone().requires_grad_(True)
                if dt == torch.float64:
                    with rh.float64_mode():
                        el = L(pp, tt, reduction_override="none")
                        el.sum().backward()
                else:
                    el = L(pp, tt, reduction_override="none")
                    el.sum().backward()
                out["%s_%s_loss_%s" % (box, mode, tag)] = _np(el)
                out["%s_%s_gpred_%s" % (box, mode, tag)] = _np(pp.grad)
                out["%s_%s_gtarget_%s" % (box, mode, tag)] = _np(tt.grad)
        with rh.float64_mode():
            L = R.SphIoULossLegacy(loss_weight=2.0)
            pd, td = p.double(), t.double()
            out[box + "_red_mean"] = _np(L(pd, td))
            out[box + "_red_w1_avg"] = _np(L(pd, td, w1.double(), avg_factor=123.0))
            out[box + "_red_w2"] = _np(L(pd, td, w2.double()))
            out[box + "_red_w1_sum"] = _np(L(pd, td, w1.double(), reduction_override="sum"))
            out[box + "_red_zero_w"] = _np(L(pd, td, torch.zeros_like(pd)))      # (a 1-D zero weight fails to broadcast in mmrotate)
            out[box + "_red_linear"] = _np(R.SphIoULossLegacy(linear=True)(pd, td, w1.double()))
        print("legacy loss", box, float(out[box + "_log_loss_f64"].mean()), float(out[box + "_red_mean"]))
    np.savez_compressed(os.path.join(OUT, "legacy_loss.npz"), **out)


def golden_planar_nms(R):
    """PlanarNMS (planar_nms.py:7-18; test_cfg.iou_calculator = 'planar') on the BFoV boxes of golden_naive: the reference's
    class on top of the restated mmcv batched_nms; class-agnostic (its default) and per class, two thresholds."""
    g = np.load(os.path.join(OUT, "naive.npz"))
    boxes, scores, idxs = (torch.from_numpy(g["bfov_" + k]) for k in ("boxes", "scores", "idxs"))
    out = {}
    for thr in (0.3, 0.5):
        for tag, kw in (("agnostic", {}), ("per_class", dict(class_agnostic=False))):
            dets, keep = R.PlanarNMS()(boxes.clone(), scores.clone(), idxs, dict(type="nms", iou_threshold=thr), **kw)
            out["keep_%s_thr%d" % (tag, int(thr * 10))] = _np(keep)
            out["dets_%s_thr%d" % (tag, int(thr * 10))] = _np(dets)
    dets, keep = R.PlanarNMS()(boxes.clone(), scores.clone(), idxs, dict(type="nms", iou_threshold=0.5, max_num=40, score_threshold=0.2))
    out["keep_max40_score02"] = _np(keep)
    np.savez_compressed(os.path.join(OUT, "planar_nms.npz"), **out)
    print("planar nms", {k: len(v) for k, v in out.items() if k.startswith("keep")})


def golden_box_format(R):
    """sphdet/bbox/box_formator.py through the reference's own functions and classes (fp32, as the pipeline runs them)."""
    bf = R.box_formator
    torch.manual_seed(81)
    sph4 = R.generate_boxes(1500, alpha_range=(0.5, 179), beta_range=(0.5, 179), dtype="float", box="bfov")
    sph5 = R.generate_boxes(1500, alpha_range=(0.5, 179), beta_range=(0.5, 179), dtype="float", box="rbfov")
    xywh = torch.rand(1500, 4) * torch.tensor([1024., 512., 300., 200.])
    xyxy = bf.xywh2xyxy(xywh)
    obb = torch.cat([torch.randn(1500, 2), torch.rand(1500, 2) * 3, (torch.rand(1500, 1) - 0.5) * 7], dim=1)
    geo = torch.cat([torch.rand(1500, 1) * 360 - 180, torch.rand(1500, 1) * 180 - 90, torch.rand(1500, 3) * 90], dim=1)
    out = dict(sph4=_np(sph4), sph5=_np(sph5), xywh=_np(xywh), xyxy=_np(xyxy), obb=_np(obb), geo=_np(geo))
    out["xyxy2xywh"] = _np(bf.xyxy2xywh(xyxy)); out["xywh2xyxy"] = _np(xyxy)
    out["obb2hbb_wywh"] = _np(bf.obb2hbb_wywh(obb)); out["obb2hbb_xyxy"] = _np(bf.obb2hbb_xyxy(obb))
    out["bfov2rbfov"] = _np(bf.bfov2rbfov(sph4))
    out["geo2sph_5"] = _np(bf.geo2sph(geo)); out["geo2sph_4"] = _np(bf.geo2sph(geo[:, :4].contiguous()))
    out["sph2geo_5"] = _np(bf.sph2geo(sph5)); out["sph2geo_4"] = _np(bf.sph2geo(sph4))
    for mode in ("sph2pix", "sph2tan"):
        for size in ((512, 1024), (960, 1920)):
            tag = "%s_%d" % (mode, size[0])
            out["planar4_" + tag] = _np(bf.Sph2PlanarBoxTransform(mode, 4)(sph4, size))
            out["planar5_" + tag] = _np(bf.Sph2PlanarBoxTransform(mode, 5)(sph5, size))
            out["back4_" + tag] = _np(bf.Planar2SphBoxTransform(mode, 4)(xyxy, size))
            out["back5_" + tag] = _np(bf.Planar2SphBoxTransform(mode, 5)(xyxy, size))
    np.savez_compressed(os.path.join(OUT, "box_format.npz"), **out)
    print("box_format", len(out))


def golden_nms(R):
    out = {}
    for box in ("bfov", "rbfov"):
        torch.manual_seed(41)
        seeds = R.generate_boxes(80, alpha_range=(5, 60), beta_range=(5, 60), dtype="float", box=box)
        boxes = (seeds.repeat(5, 1) + torch.randn(400, seeds.size(1)) * 2).clamp(min=1)
        boxes[350:400] = boxes[300:350]                  # exact duplicates
        scores = torch.rand(400)
        idxs = torch.randint(0, 6, (400,))
        for thr in (0.3, 0.5):
            dets, keep = R.SphNMS("sph2pob_efficient")(boxes, scores, idxs, dict(type="nms", iou_threshold=thr, max_num=150))
            out["%s_keep_thr%d" % (box, int(thr * 10))] = _np(keep)
            out["%s_dets_thr%d" % (box, int(thr * 10))] = _np(dets)
        dets, keep = R.SphNMS("sph2pob_efficient")(boxes, scores, idxs, dict(iou_threshold=0.5), class_agnostic=True)
        out["%s_keep_agnostic" % box] = _np(keep)
        # the decision margin: |IoU - thr| of every ordered pair within a label, to know which fixtures are tie-free
        iou = _ref64(R.sph2pob_efficient_iou, boxes, boxes)
        out["%s_pair_iou_f64" % box] = _np(iou).astype(np.float32)
        out["%s_boxes" % box], out["%s_scores" % box], out["%s_idxs" % box] = _np(boxes), _np(scores), _np(idxs)
    np.savez_compressed(os.path.join(OUT, "nms.npz"), **out)
    print("nms", {k: v.shape for k, v in out.items() if "keep" in k})


def golden_nms_cfg4(R):
    """BASELINE configs[3] at its own shape: ONE image of 1,000 candidates (200 seed boxes + 4 noisy copies each, as
    sph_retina_b200/synthetic.py::nms_batch builds them) with labels in [0, 80), thresholds 0.3 and 0.5, through the
    reference's SphNMS -- and the R-CNN heads' wrapper ``multiclass_nms`` (sphdet/bbox/nms/utils.py:6-95) on top of it.
    The seed is advanced until no same-label pair sits within 1e-5 of either threshold (float64 run), so that the keep
    lists are decided well inside fp32 resolution."""
    out = {}
    for box in ("bfov", "rbfov"):
        D = 4 if box == "bfov" else 5
        seed = 400
        while True:
            torch.manual_seed(seed)
            seeds = R.generate_boxes(200, alpha_range=(5, 60), beta_range=(5, 60), dtype="float", box=box)
            boxes = (seeds.repeat(5, 1) + torch.randn(1000, D) * 2).clamp(min=1)
            scores = torch.rand(1000)
            idxs = torch.randint(0, 80, (1000,))
            iou = _ref64(R.sph2pob_efficient_iou, boxes, boxes)
            same = (idxs[:, None] == idxs[None, :]) & ~torch.eye(1000, dtype=torch.bool)
            margin = min(float((iou[same] - t).abs().min()) for t in (0.3, 0.5))
            if margin > 1e-5:
                break
            seed += 1
        out["%s_seed" % box], out["%s_margin" % box] = np.int64(seed), np.float64(margin)
        out["%s_boxes" % box], out["%s_scores" % box], out["%s_idxs" % box] = _np(boxes), _np(scores), _np(idxs)
        for thr in (0.3, 0.5):
            dets, keep = R.SphNMS("sph2pob_efficient")(boxes, scores, idxs, dict(type="nms", iou_threshold=thr))
            out["%s_keep_thr%d" % (box, int(thr * 10))] = _np(keep)
            out["%s_dets_thr%d" % (box, int(thr * 10))] = _np(dets)
        dets, keep = R.SphNMS("sph2pob_efficient")(boxes, scores, idxs, dict(type="nms", iou_threshold=0.5, max_num=100))
        out["%s_keep_thr5_max100" % box] = _np(keep)
        # multiclass_nms: n proposals x C classes, class-specific boxes [n, C * D], scores [n, C + 1] (last = background)
        torch.manual_seed(seed + 1000)
        n, C = 120, 8
        mb = (R.generate_boxes(n, alpha_range=(5, 60), beta_range=(5, 60), dtype="float", box=box)[:, None, :]
              + torch.randn(n, C, D) * 1.5).clamp(min=1)
        mb[60:] = mb[:60] + torch.randn(60, C, D) * 1.0           # near-duplicates: suppression happens inside a class
        mb = mb.clamp(min=1).reshape(n, C * D).contiguous()
        ms = torch.rand(n, C + 1)
        fac = torch.rand(n) * 0.5 + 0.5
        for tag, kw in (("plain", {}), ("factors", dict(score_factors=fac))):
            dets, labels, inds = R.multiclass_nms(mb, ms, 0.3, dict(type="nms", iou_threshold=0.5), max_num=100, return_inds=True,
                                                  nms_op=R.SphNMS("sph2pob_efficient"), box_version=D, **kw)
            out["%s_mc_%s_dets" % (box, tag)], out["%s_mc_%s_labels" % (box, tag)], out["%s_mc_%s_inds" % (box, tag)] = _np(dets), _np(labels), _np(inds)
        shared = mb[:, :D].contiguous()                            # class-agnostic boxes [n, D]
        dets, labels, inds = R.multiclass_nms(shared, ms, 0.3, dict(type="nms", iou_threshold=0.5), max_num=-1, return_inds=True,
                                              nms_op=R.SphNMS("sph2pob_efficient"), box_version=D)
        out["%s_mc_shared_dets" % box], out["%s_mc_shared_labels" % box], out["%s_mc_shared_inds" % box] = _np(dets), _np(labels), _np(inds)
        out["%s_mc_bboxes" % box], out["%s_mc_scores" % box], out["%s_mc_factors" % box] = _np(mb), _np(ms), _np(fac)
        # decision margin of the multiclass fixture (same class, both boxes above the score threshold)
        flat = mb.view(n, C, D)
        mm = 1.0
        for c in range(C):
            v = ms[:, c] > 0.3
            if int(v.sum()) > 1:
                i2 = _ref64(R.sph2pob_efficient_iou, flat[v, c], flat[v, c])
                off = ~torch.eye(int(v.sum()), dtype=torch.bool)
                mm = min(mm, float((i2[off] - 0.5).abs().min()))
        out["%s_mc_margin" % box] = np.float64(mm)
        print("nms cfg4", box, "seed", seed, "margin %.2e" % margin, "mc margin %.2e" % mm,
              {k: v.shape for k, v in out.items() if k.startswith(box) and ("keep" in k or "inds" in k)})
    np.savez_compressed(os.path.join(OUT, "nms_cfg4.npz"), **out)


def golden_coder(R):
    """Box coders and the decode -> Sph2PobIoULoss step of the head (reg_decoded_bbox=True): anchors from the real
    anchor grid, deltas that trigger every clamp, ~10 % positive rows (2-D weights as the head passes them)."""
    out = {}
    A = anchors_512x1024()[::37].contiguous()
    n = A.size(0)
    for box, D, cls in (("bfov", 4, R.DeltaXYWHSphBBoxCoder), ("rbfov", 5, R.DeltaXYWHASphBBoxCoder)):
        torch.manual_seed(51 + D)
        anchors = A[:, :D].clone()
        if D == 5:
            anchors[:, 4] = (torch.rand(n) - 0.5) * 60
        means = (0.01, -0.02, 0.03, 0.0, 0.02)[:D]
        stds = (0.1, 0.1, 0.2, 0.2, 0.1)[:D]
        deltas = torch.randn(n, D) * 1.5
        deltas[::13] *= 12                                 # size / border / gamma clamps fire
        target = R.generate_boxes(n, alpha_range=(5, 100), beta_range=(5, 100), dtype="float", box=box)
        pos = torch.rand(n) < 0.1
        # positives: targets near the anchors (what the assigner produces), deltas near the encoding of the target
        target[pos] = (anchors[pos] + torch.randn(int(pos.sum()), D) * torch.tensor([4, 4, 6, 6, 10.0])[:D]).clamp(min=1)
        target[~pos] = 0
        weight = pos.float()[:, None].expand(n, D).contiguous()
        out.update({box + "_anchors": _np(anchors), box + "_deltas": _np(deltas), box + "_target": _np(target),
                    box + "_weight": _np(weight), box + "_means": np.array(means), box + "_stds": np.array(stds)})
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            for cname, kw in (("plain", {}), ("norm", dict(target_means=means, target_stds=stds)),
                              ("ctr", dict(target_means=means, target_stds=stds, add_ctr_clamp=True, ctr_clamp=8)),
                              ("noclip", dict(target_stds=stds, clip_border=False))):
                coder = cls(**kw)
                dec = coder.decode(anchors.to(dt), deltas.to(dt))
                out["%s_%s_decode_%s" % (box, cname, tag)] = _np(dec)
                out["%s_%s_encode_%s" % (box, cname, tag)] = _np(coder.encode(anchors.to(dt), dec))
            # the head's regression loss: decode -> Sph2PobIoULoss(pred, target, weight, avg_factor = #positives)
            coder = cls(target_means=means, target_stds=stds)
            # deltas of the positives: the encoding of the target plus noise (a realistic mid-training prediction)
            d2 = deltas.clone()
            d2[pos] = coder.encode(anchors[pos], target[pos]) + torch.randn(int(pos.sum()), D) * 0.5
            out[box + "_loss_deltas"] = _np(d2)
            for mode in ("iou", "ciou"):
                L = R.Sph2PobIoULoss(mode=mode, loss_weight=1.5)
                dd = d2.to(dt).clone().requires_grad_(True)
                if dt == torch.float64:
                    with rh.float64_mode():
                        loss = L(coder.decode(anchors.to(dt), dd), target.to(dt), weight.to(dt), avg_factor=float(pos.sum()))
                        loss.backward()
                else:
                    loss = L(coder.decode(anchors.to(dt), dd), target.to(dt), weight.to(dt), avg_factor=float(pos.sum()))
                    loss.backward()
                out["%s_%s_loss_%s" % (box, mode, tag)] = _np(loss)
                out["%s_%s_gdeltas_%s" % (box, mode, tag)] = _np(dd.grad)
        print("coder", box, n, int(pos.sum()), float(out[box + "_iou_loss_f64"]))
    np.savez_compressed(os.path.join(OUT, "coder.npz"), **out)


def golden_distance_coder(R):
    """DistancePointSphBBoxCoder (sphdet/bbox/coder/distance_point_sph_bbox_coder.py): points of the FCOS strides on the
    512 x 1024 image, distances that leave the image on some rows; decode with and without the border clamp, 4 and 5
    columns, the gradient of a weighted sum of the decoded boxes w.r.t. the distances, encode of the decoded boxes with
    and without max_dis, a second image size."""
    torch.manual_seed(11)
    out = {}
    n = 1000
    for tag, shape in (("512", (512, 1024)), ("960", (960, 1920))):
        H, W = shape
        stride = torch.tensor([8., 16., 32., 64., 128.])[torch.randint(0, 5, (n,))]
        pts = torch.stack([torch.floor(torch.rand(n) * W / stride) * stride + torch.floor(stride / 2),
                           torch.floor(torch.rand(n) * H / stride) * stride + torch.floor(stride / 2)], 1)
        dist = torch.rand(n, 4) * stride[:, None] * 6
        dist[:100] *= 8                                        # far outside the image: the clamp is active
        dist[100:130] = 0                                      # zero-size boxes
        gamma = torch.rand(n, 1) * 180 - 90
        wsum = torch.randn(n, 4)
        out["points_" + tag], out["dist_" + tag], out["gamma_" + tag], out["w_" + tag] = _np(pts), _np(dist), _np(gamma), _np(wsum)
        for D in (4, 5):
            coder = R.DistancePointSphBBoxCoder(box_version=D, img_shape=shape if tag == "960" else None)
            d = dist if D == 4 else torch.cat([dist, gamma], 1)
            for clip, ms in (("clip", shape), ("noclip", None)):
                dd = d.clone().requires_grad_(True)
                dec = coder.decode(pts, dd, max_shape=ms, img_shape=shape)
                (dec[:, :4] * wsum).sum().backward()
                out["decode%d_%s_%s" % (D, clip, tag)] = _np(dec)
                out["decode%d_%s_%s_grad" % (D, clip, tag)] = _np(dd.grad)
                for md_tag, md in (("nomax", None), ("max", 64.0)):
                    out["encode%d_%s_%s_%s" % (D, clip, md_tag, tag)] = _np(coder.encode(pts, dec.detach(), max_dis=md, img_shape=shape))
        nob = R.DistancePointSphBBoxCoder(clip_border=False)
        out["decode4_clip_border_false_" + tag] = _np(nob.decode(pts, dist, max_shape=shape, img_shape=shape))
    np.savez_compressed(os.path.join(OUT, "distance_coder.npz"), **out)
    print("distance coder", len(out), "arrays")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    R = rh.load_reference()
    if len(sys.argv) > 1:                 # regenerate the named fixtures only:  python oracle/make_golden.py nms_cfg4
        for name in sys.argv[1:]:
            globals()["golden_" + name](R)
        sys.exit(0)
    golden_kat(R)
    golden_aligned(R)
    golden_pairwise(R)
    golden_loss(R)
    golden_nms(R)
    golden_nms_cfg4(R)
    golden_coder(R)
    golden_other_losses(R)
    golden_naive(R)
    golden_unbiased(R)
    golden_legacy(R)
    golden_legacy_loss(R)
    golden_planar_nms(R)
    golden_box_format(R)
    golden_distance_coder(R)
