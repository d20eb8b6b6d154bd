"""The oracle (oracle/sph_oracle.py torch restatement, oracle/sph_oracle.c float64 C restatement)
against the golden vectors frozen from the REAL reference (oracle/make_golden.py)."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, other_loss_variants

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sph_oracle as O  # noqa: E402


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
@pytest.mark.parametrize("tr", ["efficient", "standard"])
def test_restatement_aligned_f64(box, tr):
    g = load_golden("aligned_" + box)
    b1, b2 = torch.from_numpy(g["b1"]).double(), torch.from_numpy(g["b2"]).double()
    for key, kw in (("iou", {}), ("iof", dict(mode="iof")), ("chord", dict(rbb_edge="chord")),
                    ("tangent", dict(rbb_edge="tangent")), ("project", dict(rbb_angle="project"))):
        got = O.sph2pob_iou(b1, b2, tr, is_aligned=True, **kw).numpy()
        want = g["%s_%s_f64" % (tr, key)]
        assert np.abs(got - want).max() < 1e-9, (box, tr, key)


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_restatement_aligned_f32_is_the_shipped_reference(box):
    g = load_golden("aligned_" + box)
    b1, b2 = torch.from_numpy(g["b1"]), torch.from_numpy(g["b2"])
    for tr in ("efficient", "standard"):
        got = O.sph2pob_iou(b1, b2, tr, is_aligned=True).numpy()
        want = g["%s_iou_f32" % tr]
        # fp32 runs differ in op order (bmm vs matmul); the bulk must agree tightly
        assert np.median(np.abs(got - want)) < 1e-6
        assert (np.abs(got - want) > 1e-3).mean() < 0.03


def test_restatement_obbs():
    for box in ("bfov", "rbfov"):
        g = load_golden("aligned_" + box)
        b1, b2 = torch.from_numpy(g["b1"][:512]).double(), torch.from_numpy(g["b2"][:512]).double()
        for tr, fn in (("efficient", O.sph2pob_efficient), ("standard", O.sph2pob_standard)):
            j1, j2 = O.jitter_spherical(b1, b2)
            o1, o2 = O.jitter_rotated(*fn(j1, j2))
            assert np.abs(o1.numpy() - g["%s_obb1_f64" % tr]).max() < 1e-9
            assert np.abs(o2.numpy() - g["%s_obb2_f64" % tr]).max() < 1e-9


def test_restatement_sph_fov():
    g = load_golden("aligned_bfov")
    b1, b2 = torch.from_numpy(g["b1"]).double(), torch.from_numpy(g["b2"]).double()
    for k in ("sph", "fov"):
        got = O.approx_iou(b1, b2, k, is_aligned=True).numpy()
        assert np.abs(got - g[k + "_f64"]).max() < 1e-9


def test_known_answers():
    """The 7 pairs printed by the reference's tests/test_all_ious.py:243-284 and the NMS fixture of
    tests/test_nms.py:6-27 (values frozen from the reference; SURVEY.md 8c lists them)."""
    g = load_golden("kat")
    b1, b2 = torch.from_numpy(g["b1"]), torch.from_numpy(g["b2"])
    want = [0.233804, 0.334934, 0.617307, 0.135759, 0.283414, 0.203681, 0.554219]
    np.testing.assert_allclose(g["sph2pob_efficient_iou"], want, atol=2e-6)
    np.testing.assert_allclose(O.sph2pob_iou(b1, b2, "efficient", is_aligned=True).numpy(), want, atol=5e-6)
    np.testing.assert_allclose(O.sph2pob_iou(b1, b2, "standard", is_aligned=True).numpy(), g["sph2pob_standard_iou"], atol=5e-6)
    np.testing.assert_allclose(O.approx_iou(b1, b2, "sph", is_aligned=True).numpy(), g["sph_iou"], atol=2e-6)
    np.testing.assert_allclose(O.approx_iou(b1, b2, "fov", is_aligned=True).numpy(), g["fov_iou"], atol=2e-6)
    dets, keep = O.nms_batched(torch.from_numpy(g["nms_boxes"]), torch.from_numpy(g["nms_scores"]),
                               torch.from_numpy(g["nms_idxs"]), 0.5)
    assert keep.tolist() == g["nms_keep"].tolist() == [0, 5, 7, 3, 8, 9]
    np.testing.assert_allclose(dets.numpy(), g["nms_dets"], atol=1e-6)


def test_restatement_pairwise():
    g = load_golden("pairwise")
    for box in ("bfov", "rbfov"):
        rows, cols = torch.from_numpy(g[box + "_rows"]).double(), torch.from_numpy(g[box + "_cols"]).double()
        assert np.abs(O.sph2pob_iou(rows, cols, "efficient").numpy() - g[box + "_rc_f64"]).max() < 1e-9
        assert np.abs(O.sph2pob_iou(cols, rows, "efficient").numpy() - g[box + "_cr_f64"]).max() < 1e-9
    gt, anc = torch.from_numpy(g["assign_gt"]).double(), torch.from_numpy(g["assign_anchors"]).double()
    assert np.abs(O.sph2pob_iou(gt, anc, "efficient").numpy() - g["assign_f64"]).max() < 1e-9


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_restatement_loss_and_grads(box):
    g = load_golden("loss_" + box)
    for mode in ("iou", "giou", "diou", "ciou"):
        p = torch.from_numpy(g["pred"]).double().requires_grad_(True)
        t = torch.from_numpy(g["target"]).double().requires_grad_(True)
        el = O.sph2pob_iou_loss_elementwise(p, t, mode=mode)
        el.sum().backward()
        assert np.abs(el.detach().numpy() - g[mode + "_loss_f64"]).max() < 1e-9, mode
        for got, key in ((p.grad, "gpred"), (t.grad, "gtarget")):
            want = g["%s_%s_f64" % (mode, key)]
            assert np.abs(got.numpy() - want).max() <= 1e-7 * max(1.0, np.abs(want).max()), (mode, key)
    p, t = torch.from_numpy(g["pred"]).double(), torch.from_numpy(g["target"]).double()
    w1, w2 = torch.from_numpy(g["w1"]).double(), torch.from_numpy(g["w2"]).double()
    n = p.size(0)
    np.testing.assert_allclose(O.sph2pob_iou_loss(p, t, loss_weight=2.0).item(), g["red_mean"], rtol=1e-10)
    np.testing.assert_allclose(O.sph2pob_iou_loss(p, t, w1, avg_factor=123.0, loss_weight=2.0).item(), g["red_w1_avg"], rtol=1e-10)
    np.testing.assert_allclose(O.sph2pob_iou_loss(p, t, w2, loss_weight=2.0).item(), g["red_w2"], rtol=1e-10)
    np.testing.assert_allclose(O.sph2pob_iou_loss(p, t, w1, reduction="sum", loss_weight=2.0).item(), g["red_w1_sum"], rtol=1e-10)
    np.testing.assert_allclose(O.sph2pob_iou_loss(p, t, torch.zeros(n).double(), loss_weight=2.0).item(), g["red_zero_w"], atol=1e-12)


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_restatement_other_losses(box):
    """Sph2PobGDLoss / Sph2PobKFLoss / Sph2PobL1Loss: the restatement against the reference's
    own subclasses + decorator run in float64 (the mmrotate base classes are restated, see oracle/mmrotate_losses.py)."""
    g = load_golden("other_losses_" + box)
    fns = {"Sph2PobGDLoss": O.sph2pob_gd_loss, "Sph2PobKFLoss": O.sph2pob_kf_loss, "Sph2PobL1Loss": O.sph2pob_l1_loss}
    w1, w2 = torch.from_numpy(g["w1"]).double(), torch.from_numpy(g["w2"]).double()
    for name, (cls, kw) in other_loss_variants(g).items():
        fn = fns[cls]
        p = torch.from_numpy(g["pred"]).double().requires_grad_(True)
        t = torch.from_numpy(g["target"]).double().requires_grad_(True)
        el = fn(p, t, reduction="none", **kw)
        el.sum().backward()
        want = g[name + "_loss_f64"]
        assert np.abs(el.detach().numpy() - want).max() <= 1e-9 * max(1.0, np.abs(want).max()), name
        for got, key in ((p.grad, "gpred"), (t.grad, "gtarget")):
            want = g["%s_%s_f64" % (name, key)]
            assert np.abs(got.numpy() - want).max() <= 1e-7 * max(1.0, np.abs(want).max()), (name, key)
        p, t = p.detach(), t.detach()
        np.testing.assert_allclose(fn(p, t, loss_weight=2.0, **kw).item(), g[name + "_red_mean"], rtol=1e-9)
        np.testing.assert_allclose(fn(p, t, w2, avg_factor=77.0, loss_weight=2.0, **kw).item(), g[name + "_red_w2_avg"], rtol=1e-9)
        np.testing.assert_allclose(fn(p, t, w2, reduction_override="sum", loss_weight=2.0, **kw).item(), g[name + "_red_w2_sum"], rtol=1e-9)
        if cls != "Sph2PobL1Loss":
            np.testing.assert_allclose(fn(p, t, w1, avg_factor=123.0, loss_weight=2.0, **kw).item(), g[name + "_red_w1_avg"], rtol=1e-9)


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_restatement_nms(box):
    g = load_golden("nms")
    boxes, scores, idxs = (torch.from_numpy(g["%s_%s" % (box, k)]) for k in ("boxes", "scores", "idxs"))
    for thr, tag in ((0.3, "thr3"), (0.5, "thr5")):
        dets, keep = O.nms_batched(boxes, scores, idxs, thr, max_num=150)
        assert keep.tolist() == g["%s_keep_%s" % (box, tag)].tolist()
        np.testing.assert_allclose(dets.numpy(), g["%s_dets_%s" % (box, tag)], atol=1e-6)
    # the reference pops `class_agnostic` and ignores it (sph_nms.py:33): still per label
    _, keep = O.nms_batched(boxes, scores, idxs, 0.5, class_agnostic=True)
    assert keep.tolist() == g[box + "_keep_agnostic"].tolist()


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_restatement_nms_at_the_test_time_shape_and_multiclass_nms(box):
    """BASELINE configs[3] at its own shape (one image: 1,000 candidates, labels in [0, 80), thresholds 0.3 / 0.5; no
    same-label pair within 1e-5 of a threshold in the float64 run -- tests/golden/nms_cfg4.npz records the margin) and the
    R-CNN wrapper multiclass_nms (sphdet/bbox/nms/utils.py:6-95), both from the reference's own classes."""
    g = load_golden("nms_cfg4")
    assert float(g[box + "_margin"]) > 1e-5 and float(g[box + "_mc_margin"]) > 1e-5
    boxes, scores, idxs = (torch.from_numpy(g["%s_%s" % (box, k)]) for k in ("boxes", "scores", "idxs"))
    assert boxes.shape[0] == 1000 and int(idxs.max()) < 80
    for thr, tag in ((0.3, "thr3"), (0.5, "thr5")):
        dets, keep = O.nms_batched(boxes, scores, idxs, thr)
        assert keep.tolist() == g["%s_keep_%s" % (box, tag)].tolist()
        np.testing.assert_allclose(dets.numpy(), g["%s_dets_%s" % (box, tag)], atol=1e-6)
    _, keep = O.nms_batched(boxes, scores, idxs, 0.5, max_num=100)
    assert keep.tolist() == g[box + "_keep_thr5_max100"].tolist()
    D = boxes.size(1)
    mb, ms, fac = (torch.from_numpy(g["%s_mc_%s" % (box, k)]) for k in ("bboxes", "scores", "factors"))
    for tag, kw in (("plain", {}), ("factors", dict(score_factors=fac))):
        dets, labels, inds = O.multiclass_nms(mb, ms, 0.3, 0.5, max_num=100, box_version=D, **kw)
        assert inds.tolist() == g["%s_mc_%s_inds" % (box, tag)].tolist() and labels.tolist() == g["%s_mc_%s_labels" % (box, tag)].tolist()
        np.testing.assert_allclose(dets.numpy(), g["%s_mc_%s_dets" % (box, tag)], atol=1e-6)
    dets, labels, inds = O.multiclass_nms(mb[:, :D].contiguous(), ms, 0.3, 0.5, box_version=D)
    assert inds.tolist() == g[box + "_mc_shared_inds"].tolist() and labels.tolist() == g[box + "_mc_shared_labels"].tolist()


CODER_VARIANTS = (("plain", {}), ("norm", "norm"), ("ctr", "ctr"), ("noclip", "noclip"))


def coder_kwargs(g, box, name):
    means, stds = tuple(g[box + "_means"].tolist()), tuple(g[box + "_stds"].tolist())
    return {"plain": {}, "norm": dict(means=means, stds=stds),
            "ctr": dict(means=means, stds=stds, add_ctr_clamp=True, ctr_clamp=8),
            "noclip": dict(stds=stds, clip_border=False)}[name]


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_restatement_coder(box):
    """delta2bbox / bbox2delta and the head's decode -> Sph2PobIoULoss step against the reference's own coder classes."""
    g = load_golden("coder")
    anchors, deltas = torch.from_numpy(g[box + "_anchors"]).double(), torch.from_numpy(g[box + "_deltas"]).double()
    for name, _ in CODER_VARIANTS:
        kw = coder_kwargs(g, box, name)
        dec = O.delta2bbox(anchors, deltas, **kw)
        assert np.abs(dec.numpy() - g["%s_%s_decode_f64" % (box, name)]).max() < 1e-9, (box, name)
        enc = O.bbox2delta(anchors, dec, means=kw.get("means"), stds=kw.get("stds"))
        np.testing.assert_allclose(enc.numpy(), g["%s_%s_encode_f64" % (box, name)], rtol=1e-5, atol=1e-5)   # fp32 inside
        dec32 = O.delta2bbox(anchors.float(), deltas.float(), **kw).numpy()
        np.testing.assert_allclose(dec32, g["%s_%s_decode_f32" % (box, name)], rtol=2e-6, atol=1e-6)
    target, weight = torch.from_numpy(g[box + "_target"]).double(), torch.from_numpy(g[box + "_weight"]).double()
    kw = coder_kwargs(g, box, "norm")
    for mode in ("iou", "ciou"):
        d = torch.from_numpy(g[box + "_loss_deltas"]).double().requires_grad_(True)
        loss = O.decode_iou_loss(anchors, d, target, weight, avg_factor=float((weight[:, 0] > 0).sum()), mode=mode,
                                 loss_weight=1.5, **kw)
        loss.backward()
        assert abs(float(loss.detach()) - float(g["%s_%s_loss_f64" % (box, mode)])) < 1e-9
        assert np.abs(d.grad.numpy() - g["%s_%s_gdeltas_f64" % (box, mode)]).max() < 1e-9


# ---- the float64 C restatement (exact polygon clipping) ----------------------------------------
def _c_aligned(lib, kind, b1, b2, mode=0, edge=0):
    b1, b2 = np.ascontiguousarray(b1, np.float32), np.ascontiguousarray(b2, np.float32)
    P, D = b1.shape
    out = np.empty(P, np.float64)
    fp, dp = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
    lib.sph_oracle_iou_aligned(kind, b1.ctypes.data_as(fp), b2.ctypes.data_as(fp), ctypes.c_long(P), D, mode, edge,
                               out.ctypes.data_as(dp), None)
    return out


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_c_oracle_aligned(c_oracle, box):
    """Exact clipping vs the reference's vertex-sort rotated IoU: they agree to 1e-6 except on the few
    degenerate pairs where the vendored algorithm itself is inexact (coincident edges)."""
    g = load_golden("aligned_" + box)
    for kind, tr in ((0, "efficient"), (1, "standard")):
        for key, mode, edge in (("iou", 0, 0), ("iof", 1, 0), ("chord", 0, 1), ("tangent", 0, 2)):
            got = _c_aligned(c_oracle, kind, g["b1"], g["b2"], mode, edge)
            err = np.abs(got - g["%s_%s_f64" % (tr, key)])
            assert np.median(err) < 1e-9
            assert (err > 1e-6).mean() < 5e-3, (box, tr, key, (err > 1e-6).sum())


def test_restatement_get_targets_single_hand_case():
    """anchor_head.py:254-285 on a case small enough to check by hand."""
    anchors = torch.tensor([[10., 20, 30, 40], [50, 60, 10, 10], [90, 90, 20, 20], [0, 0, 5, 5]])
    gts = torch.tensor([[12., 22, 28, 36], [88, 92, 22, 18]])
    gt_inds = torch.tensor([1, 0, 2, -1])
    labels, lw, bt, bw, npos, nneg = O.get_targets_single(anchors, gts, torch.tensor([7, 3]), gt_inds, 80)
    assert labels.tolist() == [7, 80, 3, 80] and lw.tolist() == [1, 1, 1, 0] and (npos, nneg) == (2, 1)
    assert torch.equal(bt[0], gts[0]) and torch.equal(bt[2], gts[1]) and not bt[1].any() and not bt[3].any()
    assert bw.tolist() == [[1] * 4, [0] * 4, [1] * 4, [0] * 4]
    enc = O.get_targets_single(anchors, gts, None, gt_inds, 1, reg_decoded_bbox=False, pos_weight=2.0, stds=[0.1, 0.1, 0.2, 0.2])
    assert enc[0].tolist() == [0, 1, 0, 1] and enc[1].tolist() == [2, 1, 2, 0]
    np.testing.assert_allclose(enc[2][0].numpy(), [(12 - 10) / 30 / 0.1, (22 - 20) / 40 / 0.1, np.log(28 / 30) / 0.2, np.log(36 / 40) / 0.2], rtol=1e-5)


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_restatement_naive_iou_and_its_nms(box):
    """naive_iou (sph_iou_api.py:181-198) and SphNMS('naive_iou') against the reference's own functions."""
    g = load_golden("naive")
    b1, b2 = torch.from_numpy(g[box + "_b1"]), torch.from_numpy(g[box + "_b2"])
    got = O.naive_iou(b1.double(), b2.double(), is_aligned=True).numpy()
    np.testing.assert_allclose(got, g[box + "_aligned_f64"], atol=1e-9, equal_nan=True)
    got = O.naive_iou(b1[:37].double(), b2[:301].double()).numpy()
    np.testing.assert_allclose(got, g[box + "_rc_f64"], atol=1e-9, equal_nan=True)
    np.testing.assert_allclose(O.naive_iou(b1, b2, is_aligned=True).numpy(), g[box + "_aligned_f32"], atol=1e-6, equal_nan=True)
    boxes, scores, idxs = (torch.from_numpy(g["%s_%s" % (box, k)]) for k in ("boxes", "scores", "idxs"))
    for thr in (0.3, 0.5):
        _, keep = O.nms_batched(boxes, scores, idxs, thr, max_num=150, iou_fn=lambda a, b: O.naive_iou(a, b))
        assert keep.tolist() == g["%s_keep_thr%d" % (box, int(thr * 10))].tolist()


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_restatement_unbiased_iou_and_its_nms(box):
    """unbiased_iou (sph_iou_api.py:103-125) against the reference's numpy classes fed with float64 copies of the boxes."""
    g = load_golden("unbiased")
    b1, b2 = torch.from_numpy(g[box + "_b1"]), torch.from_numpy(g[box + "_b2"])
    got = O.unbiased_iou(b1, b2, is_aligned=True).numpy()
    err = np.abs(got - g[box + "_aligned_f64"])
    assert (err > 1e-6).sum() <= 1, np.where(err > 1e-6)[0]          # (a degenerate pair may land on the other side of the 5e-9 test)
    err = np.abs(O.unbiased_iou(b1[:23], b2[:201]).numpy() - g[box + "_rc_f64"])
    assert (err > 1e-6).sum() <= 1
    boxes, scores, idxs = (torch.from_numpy(g["%s_%s" % (box, k)]) for k in ("boxes", "scores", "idxs"))
    for thr in (0.3, 0.5):
        _, keep = O.nms_batched(boxes, scores, idxs, thr, max_num=120, iou_fn=lambda a, b: O.unbiased_iou(a, b))
        assert keep.tolist() == g["%s_keep_thr%d" % (box, int(thr * 10))].tolist()


def test_restatement_sph2pob_legacy():
    """oracle.sph2pob_legacy (sph2pob_legacy.py:8-31) behind the common pipeline against the reference's sph2pob_legacy_iou."""
    g = load_golden("legacy")
    b1, b2 = torch.from_numpy(g["b1"]), torch.from_numpy(g["b2"])
    for key, kw in (("iou", {}), ("iof", dict(mode="iof")), ("chord", dict(rbb_edge="chord")), ("tangent", dict(rbb_edge="tangent"))):
        got = O.sph2pob_iou(b1.double(), b2.double(), transform="legacy", is_aligned=True, **kw).numpy()
        np.testing.assert_allclose(got, g[key + "_f64"], atol=1e-8)
        got = O.sph2pob_iou(b1, b2, transform="legacy", is_aligned=True, **kw).numpy()
        np.testing.assert_allclose(got, g[key + "_f32"], atol=1e-6)
    got = O.sph2pob_iou(b1[:29].double(), b2[:333].double(), transform="legacy").numpy()
    np.testing.assert_allclose(got, g["rc_f64"], atol=1e-8)
    with pytest.raises(AssertionError):
        O.sph2pob_legacy(torch.zeros(2, 5), torch.zeros(2, 5))


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_restatement_legacy_iou_loss(box):
    """oracle.sph_iou_loss_legacy_elementwise against the reference's SphIoULossLegacy (float64 run): losses and gradients."""
    g, base = load_golden("legacy_loss"), load_golden("loss_" + box)
    for mode in ("log", "linear", "square"):
        p = torch.from_numpy(base["pred"]).double().requires_grad_(True)
        t = torch.from_numpy(base["target"]).double().requires_grad_(True)
        el = O.sph_iou_loss_legacy_elementwise(p, t, mode)
        el.sum().backward()
        key = "%s_%s_" % (box, mode)
        np.testing.assert_allclose(el.detach().numpy(), g[key + "loss_f64"], atol=1e-8)
        np.testing.assert_allclose(p.grad.numpy(), g[key + "gpred_f64"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(t.grad.numpy(), g[key + "gtarget_f64"], rtol=1e-6, atol=1e-7)


def test_restatement_planar_nms():
    """oracle.planar_nms (PlanarNMS over the restated mmcv batched_nms) against the reference's class, and against the
    independent route: per class it is the greedy NMS of SphNMS('naive_iou') on the same boxes."""
    g, base = load_golden("planar_nms"), load_golden("naive")
    boxes, scores, idxs = (torch.from_numpy(base["bfov_" + k]) for k in ("boxes", "scores", "idxs"))
    for thr in (0.3, 0.5):
        for tag, kw in (("agnostic", {}), ("per_class", dict(class_agnostic=False))):
            dets, keep = O.planar_nms(boxes, scores, idxs, dict(type="nms", iou_threshold=thr), **kw)
            assert keep.tolist() == g["keep_%s_thr%d" % (tag, int(thr * 10))].tolist()
            np.testing.assert_array_equal(dets.numpy(), g["dets_%s_thr%d" % (tag, int(thr * 10))])
        _, keep = O.nms_batched(boxes, scores, idxs, thr, iou_fn=lambda a, b: O.naive_iou(a, b))
        assert keep.tolist() == g["keep_per_class_thr%d" % int(thr * 10)].tolist()
    _, keep = O.planar_nms(boxes, scores, idxs, dict(type="nms", iou_threshold=0.5, max_num=40, score_threshold=0.2))
    assert keep.tolist() == g["keep_max40_score02"].tolist()


def test_restatement_distance_point_coder():
    """distance2bbox / bbox2distance against the reference's DistancePointSphBBoxCoder, bit for bit."""
    g = load_golden("distance_coder")
    T = lambda k: torch.from_numpy(g[k])
    for tag, shape in (("512", (512, 1024)), ("960", (960, 1920))):
        pts = T("points_" + tag)
        for D in (4, 5):
            d = T("dist_" + tag) if D == 4 else torch.cat([T("dist_" + tag), T("gamma_" + tag)], 1)
            for clip, ms in (("clip", shape), ("noclip", None)):
                dec = O.distance2bbox(pts, d, ms, shape)
                assert torch.equal(dec, T("decode%d_%s_%s" % (D, clip, tag)))
                for md_tag, md in (("nomax", None), ("max", 64.0)):
                    assert torch.equal(O.bbox2distance(pts, dec, md, 0.1, shape), T("encode%d_%s_%s_%s" % (D, clip, md_tag, tag)))
