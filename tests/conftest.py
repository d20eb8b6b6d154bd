import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


# ---- parity report -------------------------------------------------------------------------------------------------
# Every numeric comparison of the suite leaves a record here (per test: sample size, max / median error against the
# reference's float64 run, how many elements exceed the tolerance against the float64 run AND against the reference's
# own float32 run, and how many elements pass only through each of the criterion's allowances).  At the end of the
# session the records are written to gpurun_out/parity_report_{gpu,cpu}.json; the GPU file of the round is committed
# as profiles/parity_r02.json.  Nothing here changes a verdict: it makes the verdicts readable.
_PARITY = {"tests": {}}
_CURRENT = [None]

PARITY_NOTES = {
    "criterion": "IoU: |kernel - reference run in float64| <= 1e-5, OR no farther from it than the reference's own float32 run is "
                 "on that element (SURVEY.md 8c: the fp32 reference's noise floor against its fp64 run is 1.2e-5..1.2e-3). "
                 "Gradients: row-wise relative L2 error <= 1e-4 against fp64 autograd, or no worse than the fp32 autograd row. "
                 "NMS / assignment: index sets equal.",
    "allowances": {
        "fp32_clause": "elements with error > tol that pass because the fp32 reference is at least as far from the fp64 run",
        "degenerate_pairs": "pairs with a box below 0.06 deg (jitter_2 inflates it to a 2.5e-4 rad speck whose position relative to "
                            "the other box's edge is decided below fp32 resolution): checked at 1e-3",
        "quirk_zone": "pairs constructed inside jitter_2's eps windows: a size / angle difference within fp32 rounding of eps takes the "
                      "other branch than the float64 run in ANY fp32 evaluation (the reference-order path included): <= 12 of "
                      "120,000 pairs may exceed 1e-5 (none 1e-3)",
    },
    "kAcosLo": "csrc/sphk_math.cuh:56-58 confines acos-derived angles to [4.47213602e-4, pi - 4.47213602e-4] = acos(1 - 1e-7) as "
               "the reference's FLOAT64 run clamps them. The reference as shipped in float32 clamps the cosine at fl32(1 - 1e-7) = "
               "0.99999988, i.e. at 4.88e-4 rad. Inside that zone (centres or tangent directions closer than 4.9e-4 rad to "
               "(anti)parallel: near-coincident loss / NMS duplicates) the kernels track the float64 run, not the float32 one: a "
               "documented deviation of up to 4.1e-5 rad in one planar angle / the centre distance, IoU effect <= 2e-5 on such pairs, "
               "always on the side of the float64 truth.",
}


@pytest.fixture(autouse=True)
def _parity_scope(request):
    _CURRENT[0] = request.node.nodeid
    yield
    _CURRENT[0] = None


def parity_record(check, **fields):
    """Append one comparison record to the running test's entry of the parity report."""
    if _CURRENT[0] is None:
        return
    rec = {"check": check}
    for k, v in fields.items():
        if isinstance(v, (np.floating, np.integer)):
            v = v.item()
        rec[k] = v
    _PARITY["tests"].setdefault(_CURRENT[0], []).append(rec)


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY["tests"]:
        return
    import json
    try:
        import torch
        on_gpu = torch.cuda.is_available()
    except Exception:
        on_gpu = False
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    totals = {"comparisons": 0, "elements": 0, "n_gt_tol_vs_fp64": 0, "n_gt_tol_vs_fp32_ref": 0, "fp32_ref_n_gt_tol": 0,
              "n_passed_only_via_fp32_clause": 0, "n_rescued_degenerate": 0, "n_quirk_zone_gt_1e-5": 0, "n_fail_before_allowances": 0}
    for recs in _PARITY["tests"].values():
        for r in recs:
            totals["comparisons"] += 1
            totals["elements"] += int(r.get("n", 0))
            for k in list(totals)[2:]:
                totals[k] += int(r.get(k, 0))
    doc = {"device": "cuda" if on_gpu else "cpu (hostsim build of the device headers / oracle)", "exit_status": int(exitstatus),
           "notes": PARITY_NOTES, "totals": totals, "tests": _PARITY["tests"]}
    with open(os.path.join(out_dir, "parity_report_%s.json" % ("gpu" if on_gpu else "cpu")), "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def _system_gxx():
    for c in ("/usr/bin/g++", "g++"):
        if os.path.isfile(c) or c == "g++":
            return c


@pytest.fixture(scope="session")
def hostsim():
    """g++ build of the DEVICE arithmetic headers (csrc/sphk_math.cuh, sphk_grad.cuh): lets the no-GPU
    suite check the kernels' maths.  Test infrastructure only; the product never loads it."""
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    out_dir = os.path.join(ROOT, "tests", "hostsim", "_build")
    out = os.path.join(out_dir, "libhostsim.so")
    deps = [src] + [os.path.join(ROOT, "sph_retina_b200", "csrc", f) for f in ("sphk_math.cuh", "sphk_fast.cuh", "sphk_grad.cuh", "sphk_coder.cuh", "sphk_obbloss.cuh")]
    if not os.path.isfile(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call([_system_gxx(), "-O2", "-fPIC", "-shared", "-DSPHK_WITH_GRAD", "-ffp-contract=fast",
                               "-mfma", "-o", out, src, "-lm"])
    lib = ctypes.CDLL(out)
    return lib


@pytest.fixture(scope="session")
def c_oracle():
    """oracle/_build/libsph_oracle.so (float64 C restatement, OpenMP)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libsph_oracle.so"))
    return lib


def within(kernel, truth64, ref32=None, tol=1e-5, slack=1.0, what=None):
    """The parity criterion of SURVEY.md 8(c): |kernel - fp64 reference| <= tol, OR the kernel is at
    least as close to the fp64 run as the reference's own fp32 run is on that element."""
    err = np.abs(np.asarray(kernel, np.float64) - np.asarray(truth64, np.float64))
    ok = err <= tol
    rec = {"n": int(err.size), "tol": tol, "max_err_vs_fp64": float(err.max()) if err.size else 0.0,
           "median_err_vs_fp64": float(np.median(err)) if err.size else 0.0, "n_gt_tol_vs_fp64": int((err > tol).sum())}
    if ref32 is not None:
        ref_err = np.abs(np.asarray(ref32, np.float64) - np.asarray(truth64, np.float64))
        via32 = (~ok) & (err <= slack * ref_err)
        ok |= err <= slack * ref_err
        d32 = np.abs(np.asarray(kernel, np.float64) - np.asarray(ref32, np.float64))
        rec.update({"fp32_ref_n_gt_tol": int((ref_err > tol).sum()), "fp32_ref_max_err_vs_fp64": float(ref_err.max()) if err.size else 0.0,
                    "n_passed_only_via_fp32_clause": int(via32.sum()), "max_diff_vs_fp32_ref": float(d32.max()) if err.size else 0.0,
                    "n_gt_tol_vs_fp32_ref": int((d32 > tol).sum())})
    rec["n_fail_before_allowances"] = int((~ok).sum())
    parity_record(what or "within", **rec)
    return ok, err


def allow_degenerate(ok, err, b1, b2, tol=1e-3):
    """The zero-size-box allowance (see degenerate_pairs), with the number of elements it rescues on record."""
    deg = degenerate_pairs(b1, b2)
    if deg.shape != ok.shape:
        deg = np.broadcast_to(deg.reshape(-1, *([1] * (ok.ndim - 1))), ok.shape) if deg.size == ok.shape[0] else deg.reshape(ok.shape)
    rescued = (~ok) & deg & (err < tol)
    parity_record("degenerate_pairs", n=int(deg.sum()), n_rescued_degenerate=int(rescued.sum()), tol=tol)
    return ok | (deg & (err < tol))


def degenerate_pairs(b1, b2, min_fov_deg=0.06):
    """Pairs with a (near) zero-size box: jitter_2 inflates it to a ~2.5e-4 rad square, and whether that
    speck straddles the other box's edge is decided below fp32 resolution of the O(1) coordinates.
    IoF of such a pair is ill-conditioned in ANY fp32 evaluation; these rows are checked at 1e-3."""
    b1, b2 = np.asarray(b1), np.asarray(b2)
    return (np.minimum(b1[:, 2:4].min(axis=1), b2[:, 2:4].min(axis=1)) < min_fov_deg)


def grad_rows_ok(got, truth, ref32, live_rows, tol=1e-4):
    """Row-wise gradient parity: relative L2 error <= tol against the reference's fp64 autograd, or no
    worse than the reference's own fp32 autograd on that row.  Returns (ok[rows], rel[rows], rel32[rows])."""
    got, truth, ref32 = (np.asarray(a, np.float64) for a in (got, truth, ref32))
    den = np.sqrt((truth ** 2).sum(axis=1))
    rel = np.sqrt(((got - truth) ** 2).sum(axis=1)) / np.maximum(den, 1e-30)
    rel32 = np.sqrt(((ref32 - truth) ** 2).sum(axis=1)) / np.maximum(den, 1e-30)
    live = np.asarray(live_rows, bool) & (den > 1e-12)
    good = ((rel <= tol) | (rel <= rel32))[live]
    parity_record("grad_rows", n=int(live.sum()), tol=tol, median_rel_vs_fp64=float(np.median(rel[live])) if live.any() else 0.0,
                  max_rel_vs_fp64=float(rel[live].max()) if live.any() else 0.0, n_gt_tol_vs_fp64=int((rel[live] > tol).sum()),
                  fp32_ref_n_gt_tol=int((rel32[live] > tol).sum()), fp32_ref_median_rel=float(np.median(rel32[live])) if live.any() else 0.0,
                  n_passed_only_via_fp32_clause=int(((rel > tol) & (rel <= rel32))[live].sum()), n_fail_before_allowances=int((~good).sum()))
    return good, rel[live], rel32[live]


def other_loss_variants(g):
    """{name: (reference class name, constructor kwargs)} stored inside tests/golden/other_losses_*.npz."""
    import json
    return {k: (v[0], v[1]) for k, v in json.loads(str(g["variants_json"])).items()}


def other_loss_kernel_args(cls, kw):
    """(loss kind, fun, flags, tau, alpha) of sphk_obb_loss for a reference loss class + constructor kwargs."""
    if cls == "Sph2PobGDLoss":
        kind = {"gwd": 0, "kld": 1, "jd": 2, "kld_symmax": 3, "kld_symmin": 4}[kw["loss_type"]]
        opt = kw.get("normalize", True) if kind == 0 else kw.get("sqrt", True)
        return kind, {"none": 0, "log1p": 1, "sqrt": 2}[kw.get("fun", "log1p")], int(opt), kw.get("tau", 0.0), kw.get("alpha", 1.0)
    if cls == "Sph2PobKFLoss":
        return 5, {"none": 0, "ln": 1, "exp": 2}[kw.get("fun", "none")], 0, 0.0, 1.0
    flags = (1 if kw.get("encode", True) else 0) | (2 if kw.get("swap", False) else 0) | \
        (4 if kw.get("angle_modifier", "original") == "modulus" else 0)
    return 6, 0, flags, 0.0, 1.0


def check_other_loss(name, g, loss, gpred, gtarget, identical_rows=16):
    """Parity of one Sph2PobGD/KF/L1 variant against the golden run (SURVEY.md 8c criterion).
    Loss: 1e-5 (relative to max(1, |truth|): l1_swap divides by jitter-sized widths and reaches 1e4) or no worse than
    the fp32 reference.  Gradient rows: 1e-4 relative or no worse than the fp32 reference's row, on >= 99 % of the
    rows; the rows whose pred == target (3-eps apart after jitter_1, d(angle)/d(centre) ~ 1e5 with cancelling signs)
    are conditioned beyond fp32 and only counted in a looser overall bound."""
    t64, r32 = g[name + "_loss_f64"], g[name + "_loss_f32"]
    loss = np.asarray(loss, np.float64).reshape(t64.shape)
    err = np.abs(loss - t64) / np.maximum(1.0, np.abs(t64))
    err32 = np.abs(r32 - t64) / np.maximum(1.0, np.abs(t64))
    ok = (err <= 1e-5) | (err <= err32)
    assert ok.mean() > 0.999, (name, "loss", np.where(~ok)[0][:10], err[~ok][:10])
    assert (err > 1e-5).sum() <= max(2, 0.5 * (err32 > 1e-5).sum()), (name, (err > 1e-5).sum(), (err32 > 1e-5).sum())
    for got, key in ((gpred, "gpred"), (gtarget, "gtarget")):
        truth, ref32 = g["%s_%s_f64" % (name, key)], g["%s_%s_f32" % (name, key)]
        live = np.ones(len(truth), bool)
        good, rel, rel32 = grad_rows_ok(got, truth, ref32, live)
        assert np.isfinite(np.asarray(got)).all(), (name, key)
        assert good.mean() > 0.97, (name, key, (~good).sum())
        live[:identical_rows] = False
        good, rel, rel32 = grad_rows_ok(got, truth, ref32, live)
        assert good.mean() > 0.99, (name, key, (~good).sum())
        assert np.median(rel) < 3e-6, (name, key, np.median(rel))
        assert (rel > 1e-4).sum() < 0.5 * (rel32 > 1e-4).sum(), (name, key)


BOX_FORMAT_CASES = [   # (format id, input key, d_out, golden key, image size, exact)
    (0, "xyxy", 4, "xyxy2xywh", (512, 1024), True), (1, "xywh", 4, "xywh2xyxy", (512, 1024), True),
    (2, "obb", 4, "obb2hbb_wywh", (512, 1024), False), (3, "obb", 4, "obb2hbb_xyxy", (512, 1024), False),
    (4, "sph4", 5, "bfov2rbfov", (512, 1024), True),
    (5, "geo", 5, "geo2sph_5", (512, 1024), True), (6, "sph5", 5, "sph2geo_5", (512, 1024), True), (6, "sph4", 4, "sph2geo_4", (512, 1024), True),
    (11, "sph4", 4, "planar4_sph2pix_512", (512, 1024), True), (11, "sph5", 5, "planar5_sph2pix_512", (512, 1024), True),
    (11, "sph4", 4, "planar4_sph2pix_960", (960, 1920), True), (11, "sph5", 5, "planar5_sph2pix_960", (960, 1920), True),
    (12, "sph4", 4, "planar4_sph2tan_512", (512, 1024), False), (12, "sph5", 5, "planar5_sph2tan_960", (960, 1920), False),
    (13, "xyxy", 4, "back4_sph2pix_512", (512, 1024), True), (13, "xyxy", 5, "back5_sph2pix_960", (960, 1920), True),
    (14, "xyxy", 4, "back4_sph2tan_512", (512, 1024), False), (14, "xyxy", 5, "back5_sph2tan_960", (960, 1920), False),
]


def check_distance_coder(coder_mod, dev):
    """DistancePointSphBBoxCoder against tests/golden/distance_coder.npz (the reference's class through
    oracle/make_golden.py::golden_distance_coder): decode bit for bit with / without the border clamp, 4 and 5 columns, two
    image sizes; its gradient w.r.t. the distances; encode with / without max_dis; clip_border=False; empty input."""
    import torch
    g = load_golden("distance_coder")
    T = lambda k: torch.from_numpy(g[k]).to(dev)
    for tag, shape in (("512", (512, 1024)), ("960", (960, 1920))):
        pts, dist, gamma, wsum = T("points_" + tag), T("dist_" + tag), T("gamma_" + tag), T("w_" + tag)
        for D in (4, 5):
            coder = coder_mod.DistancePointSphBBoxCoder(box_version=D, img_shape=shape if tag == "960" else None)
            d = dist if D == 4 else torch.cat([dist, gamma], 1)
            for clip, ms in (("clip", shape), ("noclip", None)):
                dd = d.clone().requires_grad_(True)
                dec = coder.decode(pts, dd, max_shape=ms, img_shape=shape)
                assert dec.shape == (len(pts), D)
                assert torch.equal(dec.detach(), T("decode%d_%s_%s" % (D, clip, tag))), (D, clip, tag)
                (dec[:, :4] * wsum).sum().backward()
                want = T("decode%d_%s_%s_grad" % (D, clip, tag))
                assert float((dd.grad - want).abs().max()) <= 1e-6 * float(want.abs().max()), (D, clip, tag)
                assert torch.equal(dd.grad == 0, want == 0)                  # clamped rows: exactly no gradient
                for md_tag, md in (("nomax", None), ("max", 64.0)):
                    enc = coder.encode(pts, dec.detach(), max_dis=md, img_shape=shape)
                    assert torch.equal(enc, T("encode%d_%s_%s_%s" % (D, clip, md_tag, tag))), (D, clip, md_tag, tag)
        nob = coder_mod.DistancePointSphBBoxCoder(clip_border=False)
        assert torch.equal(nob.decode(pts, dist, max_shape=shape, img_shape=shape), T("decode4_clip_border_false_" + tag))
    coder = coder_mod.DistancePointSphBBoxCoder()
    e2, e4 = torch.zeros(0, 2, device=dev), torch.zeros(0, 4, device=dev)
    assert coder.decode(e2, e4, max_shape=(512, 1024)).shape == (0, 4) and coder.encode(e2, e4).shape == (0, 4)
    import pytest
    with pytest.raises(AssertionError):
        coder.decode(pts, torch.cat([dist, gamma], 1))                       # box_version 4 coder, five columns


def check_anchor_free_post_processing(post_mod, coder_mod, O, dev, with_nms):
    """get_bboxes_single with score factors and the point coder (sph_fcos_head.py:196-321) against the literal chain on
    the oracle's restatements: per level sigmoid -> filter_scores_and_topk -> distance2bbox(max_shape=img_shape), then
    scores x sigmoid(centerness) and, with_nms, the greedy class-wise NMS of the oracle."""
    import torch
    torch.manual_seed(5)
    H, W, C, D = 512, 1024, 5, 4
    cls, reg, ctr, pts = [], [], [], []
    for s in (32, 64, 128):
        h, w = H // s, W // s
        cls.append(torch.randn(C, h, w) - 1.0)
        reg.append(torch.rand(D, h, w) * s * 5)
        ctr.append(torch.randn(1, h, w))
        ys, xs = torch.meshgrid(torch.arange(h) * s + s // 2, torch.arange(w) * s + s // 2, indexing="ij")
        pts.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1).float())
    cfg = dict(nms_pre=200, score_thr=0.05, nms=dict(iou_threshold=0.5), max_per_img=50, iou_calculator='sph2pob_efficient')
    coder = coder_mod.DistancePointSphBBoxCoder()
    to = lambda xs: [x.to(dev) for x in xs]
    clip, noclip, raw, fac, lbs = [], [], [], [], []
    for c, r, f, p in zip(cls, reg, ctr, pts):
        r = r.permute(1, 2, 0).reshape(-1, D)
        sc = c.permute(1, 2, 0).reshape(-1, C).sigmoid()
        valid = sc > cfg["score_thr"]
        vs, vi = sc[valid], torch.nonzero(valid)
        vs, order = vs.sort(descending=True)                                # distinct random scores: the order is determined
        keep, labels = vi[order[:min(cfg["nms_pre"], vi.size(0))]].unbind(dim=1)
        clip.append(O.distance2bbox(p[keep], r[keep], (H, W), (H, W)))
        noclip.append(O.distance2bbox(p[keep], r[keep], None, (H, W)))
        raw.append(vs[:len(keep)])
        fac.append(f.permute(1, 2, 0).reshape(-1).sigmoid()[keep])
        lbs.append(labels)
    clip, noclip, raw, fac, lbs = torch.cat(clip), torch.cat(noclip), torch.cat(raw), torch.cat(fac), torch.cat(lbs)
    assert len(clip) > 300 and not torch.equal(clip, noclip)               # all three levels contribute, the clamp is active
    b, sc, lb = post_mod.get_bboxes_single(to(cls), to(reg), to(pts), coder, cfg, box_version=D, with_nms=False,
                                           score_factor_list=to(ctr), img_shape=(H, W))
    assert torch.equal(b.cpu(), clip) and torch.equal(lb.cpu(), lbs)
    assert float((sc.cpu() - raw * fac).abs().max()) < 1e-6
    b0, sc0, lb0 = b, sc, lb
    # without the two arguments the call is the anchor-based heads' one: no factor, decode without max_shape
    b, sc, lb = post_mod.get_bboxes_single(to(cls), to(reg), to(pts), coder, cfg, box_version=D, with_nms=False)
    assert torch.equal(b.cpu(), noclip) and torch.equal(lb.cpu(), lbs) and float((sc.cpu() - raw).abs().max()) < 1e-6
    b, sc, lb = post_mod.get_bboxes_single(to(cls), to(reg), to(pts), coder, cfg, box_version=D, with_nms=False,
                                           score_factor_list=[None] * 3)
    assert float((sc.cpu() - raw).abs().max()) < 1e-6
    if with_nms:
        dets, labels = post_mod.get_bboxes_single(to(cls), to(reg), to(pts), coder, cfg, box_version=D,
                                                  score_factor_list=to(ctr), img_shape=(H, W))
        # the NMS itself is pinned to the oracle elsewhere (test_nms_*): here, that the block hands it these candidates
        wd, wk = post_mod.SphNMS('sph2pob_efficient')(b0.to(dev), sc0.to(dev), lb0.to(dev), cfg["nms"])
        assert dets.shape[1] == D + 1 and 0 < len(dets) == min(len(wk), cfg["max_per_img"])
        assert torch.equal(labels, lb0.to(dev)[wk][:cfg["max_per_img"]]) and torch.equal(dets, wd[:cfg["max_per_img"]])
        assert bool((dets[:-1, -1] >= dets[1:, -1]).all())
