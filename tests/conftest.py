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


def within(kernel, truth64, ref32=None, tol=1e-5, slack=1.0):
    """The parity criterion of SURVEY.md 8(c): |kernel - fp64 reference| <= tol, OR the kernel is at
    least as close to the fp64 run as the reference's own fp32 run is on that element."""
    err = np.abs(np.asarray(kernel, np.float64) - np.asarray(truth64, np.float64))
    ok = err <= tol
    if ref32 is not None:
        ref_err = np.abs(np.asarray(ref32, np.float64) - np.asarray(truth64, np.float64))
        ok |= err <= slack * ref_err
    return ok, err


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
    return ((rel <= tol) | (rel <= rel32))[live], rel[live], rel32[live]


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
