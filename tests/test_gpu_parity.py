"""Parity of the CUDA path (Python API -> ctypes -> C ABI -> sm_100a kernels) against the golden
vectors of the reference and against the oracle on seeded inputs.  Run on the B200: pytest -m gpu.

Tolerances (BASELINE.json north_star): IoU 1e-5 absolute, gradients 1e-4 relative, NMS keep sets
exact.  The reference's OWN fp32 run misses those against its fp64 run on part of the workload
(SURVEY.md 8c), so an element passes if it is within tolerance of the fp64 reference OR at least as
close to it as the fp32 reference is."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

from conftest import (BOX_FORMAT_CASES, ROOT, allow_degenerate, parity_record, check_other_loss, degenerate_pairs, grad_rows_ok, load_golden, other_loss_variants,
                      within)

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sph_oracle as O  # noqa: E402  (the checker)

DEV = "cuda:0"


@pytest.fixture(scope="module")
def api():
    import sph_retina_b200  # noqa: F401
    from sph_retina_b200 import _native
    from sph_retina_b200.sphdet import iou, losses
    from sph_retina_b200.sphdet.bbox import nms
    sm, major, minor = _native.device_info()
    assert major >= 10, "these kernels are built for sm_100a only"
    import types
    return types.SimpleNamespace(native=_native, iou=iou, losses=losses, nms=nms)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def c_oracle_aligned(lib, kind, b1, b2, mode=0, edge=0):
    b1, b2 = np.ascontiguousarray(b1, np.float32), np.ascontiguousarray(b2, np.float32)
    P, D = b1.shape
    out = np.empty(P, np.float64)
    fp, dp = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
    lib.sph_oracle_iou_aligned(kind, b1.ctypes.data_as(fp), b2.ctypes.data_as(fp), ctypes.c_long(P), D, mode, edge,
                               out.ctypes.data_as(dp), None)
    return out


# ---- aligned -----------------------------------------------------------------------------------
@pytest.mark.parametrize("box", ["bfov", "rbfov"])
@pytest.mark.parametrize("tr", ["efficient", "standard"])
def test_aligned_golden(api, box, tr):
    g = load_golden("aligned_" + box)
    b1, b2 = cu(g["b1"]), cu(g["b2"])
    fn = getattr(api.iou, "sph2pob_%s_iou" % tr)
    keep1, keep2 = b1.clone(), b2.clone()
    for key, kw in (("iou", {}), ("iof", dict(mode="iof")), ("chord", dict(rbb_edge="chord")),
                    ("tangent", dict(rbb_edge="tangent"))):
        got = fn(b1, b2, is_aligned=True, **kw)
        assert got.shape == (len(g["b1"]),) and got.dtype == torch.float32 and got.device == b1.device
        got = got.cpu().numpy()
        ok, err = within(got, g["%s_%s_f64" % (tr, key)], g["%s_%s_f32" % (tr, key)])
        ok = allow_degenerate(ok, err, g["b1"], g["b2"])
        assert ok.all(), (box, tr, key, np.where(~ok)[0], err[~ok])
        assert (err > 1e-5).sum() <= 2 and np.median(err) < 2e-7
        assert got.min() >= 0.0 and got.max() <= 1.0
    assert torch.equal(b1, keep1) and torch.equal(b2, keep2)        # tests/test_all_ious.py:322-331


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_project_angle_variant(api, box):
    """rbb_angle='project': aligned and N x M (plain double-precision kernel)."""
    g = load_golden("aligned_" + box)
    b1, b2 = cu(g["b1"]), cu(g["b2"])
    for tr in ("efficient", "standard"):
        fn = getattr(api.iou, "sph2pob_%s_iou" % tr)
        got = fn(b1, b2, is_aligned=True, rbb_angle="project")
        ok, err = within(got.cpu().numpy(), g[tr + "_project_f64"], g[tr + "_project_f32"])
        assert ok.all(), (box, tr, np.where(~ok)[0], err[~ok])
        mat = fn(b1[:40], b2[:70], rbb_angle="project")
        flat = fn(b1[:40].repeat_interleave(70, 0), b2[:70].repeat(40, 1), is_aligned=True, rbb_angle="project")
        assert torch.equal(mat.reshape(-1), flat)


def test_sph_fov_golden_and_known_answers(api):
    g = load_golden("aligned_bfov")
    b1, b2 = cu(g["b1"]), cu(g["b2"])
    for k in ("sph", "fov"):
        got = getattr(api.iou, k + "_iou")(b1, b2, is_aligned=True).cpu().numpy()
        assert np.abs(got - g[k + "_f64"]).max() < 2e-6
    g = load_golden("kat")
    b1, b2 = cu(g["b1"]), cu(g["b2"])
    for name in ("sph2pob_efficient_iou", "sph2pob_standard_iou", "sph_iou", "fov_iou"):
        got = getattr(api.iou, name)(b1, b2, is_aligned=True).cpu().numpy()
        np.testing.assert_allclose(got, g[name], atol=5e-6)
    calc = api.iou.SphOverlaps2D('sph2pob_efficient_iou', 4)
    np.testing.assert_allclose(calc(b1, b2, is_aligned=True).cpu().numpy(), g["sph2pob_efficient_iou"], atol=5e-6)


@pytest.mark.parametrize("box,D", [("bfov", 4), ("rbfov", 5)])
def test_aligned_fresh_seed_vs_c_oracle(api, c_oracle, box, D):
    """200k seeded pairs (config #1 generator) against the exact-clipping float64 C oracle."""
    b1 = O.generate_boxes(200_000, alpha_range=(1, 100), beta_range=(1, 100), box=box, seed=5)
    b2 = O.generate_boxes(200_000, alpha_range=(1, 100), beta_range=(1, 100), box=box, seed=6)
    for kind, tr in ((0, "efficient"), (1, "standard")):
        got = getattr(api.iou, "sph2pob_%s_iou" % tr)(b1.to(DEV), b2.to(DEV), is_aligned=True).cpu().numpy()
        want = c_oracle_aligned(c_oracle, kind, b1.numpy(), b2.numpy())
        err = np.abs(got - want)
        assert err.max() < 1e-5, (tr, err.max(), np.argmax(err))
        assert ((got > 0) == (want > 1e-9)).mean() > 0.9999


def test_aligned_odd_sizes_views_and_streams(api):
    b1 = O.generate_boxes(1000, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=7).to(DEV)
    b2 = O.generate_boxes(1000, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=8).to(DEV)
    full = api.iou.sph2pob_efficient_iou(b1, b2, is_aligned=True)
    for n in (1, 31, 33, 255, 257, 999):
        assert torch.equal(api.iou.sph2pob_efficient_iou(b1[:n], b2[:n], is_aligned=True), full[:n])
    # non-contiguous views, unaligned BFoV slices of a wider tensor, a trailing score column
    wide1, wide2 = torch.cat([b1, b1[:, :1]], 1), torch.cat([b2, b2[:, :1]], 1)
    calc = api.iou.SphOverlaps2D('sph2pob_efficient_iou', 5)
    assert torch.equal(calc(wide1, wide2, is_aligned=True), full)
    bf = api.iou.sph2pob_efficient_iou(b1[:, :4].contiguous(), b2[:, :4].contiguous(), is_aligned=True)
    assert torch.equal(api.iou.sph2pob_efficient_iou(b1[:, :4], b2[:, :4], is_aligned=True), bf)
    assert torch.equal(api.iou.sph2pob_efficient_iou(b1[1:, :4], b2[1:, :4], is_aligned=True), bf[1:])
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        on_side = api.iou.sph2pob_efficient_iou(b1, b2, is_aligned=True)
    s.synchronize()
    assert torch.equal(on_side, full)
    # float64 inputs are computed in fp32 and returned in the input dtype
    assert api.iou.sph2pob_efficient_iou(b1.double(), b2.double(), is_aligned=True).dtype == torch.float64


def test_aligned_all_pairs_on_the_reference_order_path(api, c_oracle):
    """Loss-style input: every pair is (near-)coincident, so every pair leaves the fast path -- the slow-pair ring
    of the aligned kernel is under maximum pressure (stage-1 and stage-2 rejects in the same iteration)."""
    t = O.generate_boxes(100_003, alpha_range=(5, 100), beta_range=(5, 100), box="rbfov", seed=31)
    p = t.clone()
    p[::2] += torch.randn(50_002, 5) * 1e-3          # arc < 2e-3 rad: rejected by stage 2
    p[1::4] = t[1::4]                                # identical: similarity mask, rejected by stage 1
    for kind, fn in ((0, api.iou.sph2pob_efficient_iou), (1, api.iou.sph2pob_standard_iou)):
        got = fn(p.to(DEV), t.to(DEV), is_aligned=True).cpu().numpy()
        want = c_oracle_aligned(c_oracle, kind, p.numpy(), t.numpy())
        err = np.abs(got - want)
        assert np.median(err) < 1e-6 and err.max() < 1e-5, (np.median(err), err.max(), int(np.argmax(err)))


def test_aligned_cull_never_changes_a_result(api):
    """The stage-0 cull of the aligned kernel works on MUFU sines: with the early-outs disabled (dense mode:
    every pair goes through jitter_1 + transform + clipper) the output must be bit-identical -- on random pairs and
    on pairs placed within +-10 % of the touching distance (tiny to oversize boxes, poles, seam)."""
    from test_hostsim_math import near_touching_pairs
    sets = [near_touching_pairs(400_000, s) for s in (0, 1, 2)]
    r1 = O.generate_boxes(1_000_000, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=40).numpy()
    r2 = O.generate_boxes(1_000_000, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=41).numpy()
    sets.append((r1, r2))
    for b1, b2 in sets:
        for D in (4, 5):
            x, y = cu(b1[:, :D]), cu(b2[:, :D])
            for edge in ("arc", "chord", "tangent"):
                fast = api.iou.sph2pob_efficient_iou(x, y, is_aligned=True, rbb_edge=edge)
                prev = api.native.set_dense(True)
                try:
                    dense = api.iou.sph2pob_efficient_iou(x, y, is_aligned=True, rbb_edge=edge)
                finally:
                    api.native.set_dense(prev)
                assert torch.equal(fast, dense), (D, edge, int((fast != dense).sum()))
        assert 0.02 < float((fast > 0).float().mean()) < 0.98         # both outcomes are exercised


def test_fast_path_inside_the_quirk_zones(api, c_oracle):
    """jitter_2 triggered through sizes / nearly parallel boxes and internal angles inside the clamped-acos zone are
    evaluated by the fast path of both kernels (aligned, N x M): against the float64 C oracle, with the allowance for
    pairs that sit within fp32 rounding of eps / eps' (tests/test_hostsim_math.py)."""
    from test_hostsim_math import quirk_zone_pairs
    for D in (4, 5):
        b1, b2 = quirk_zone_pairs(120_000, D)
        b1, b2 = np.ascontiguousarray(b1[:, :D]), np.ascontiguousarray(b2[:, :D])
        x, y = cu(b1), cu(b2)
        for kind, tr in ((0, "efficient"), (1, "standard")):
            fn = getattr(api.iou, "sph2pob_%s_iou" % tr)
            truth = c_oracle_aligned(c_oracle, kind, b1, b2)
            err = np.abs(fn(x, y, is_aligned=True).cpu().numpy() - truth)
            parity_record("quirk_zone", n=int(err.size), tol=1e-5, max_err_vs_fp64=float(err.max()), median_err_vs_fp64=float(np.median(err)),
                          **{"n_quirk_zone_gt_1e-5": int((err > 1e-5).sum()), "allowance": 12, "transform": tr, "D": D})
            assert (err > 1e-5).sum() <= 12 and err.max() < 1e-3 and np.median(err) < 2e-7, (D, tr, (err > 1e-5).sum(), err.max())
            # the same pairs on the diagonal of N x M blocks
            for lo in (0, 60_000):
                blk = fn(x[lo:lo + 1536], y[lo:lo + 1536]).diagonal().cpu().numpy()
                e2 = np.abs(blk - truth[lo:lo + 1536])
                assert (e2 > 1e-5).sum() <= 2 and e2.max() < 1e-3 and np.median(e2) < 2e-7, (D, tr, lo, e2.max())


def test_non_finite_and_out_of_range_inputs_do_not_poison_neighbours(api):
    """NaN / inf / absurd coordinates in one box must neither hang nor disturb other pairs (they take the
    reference-order path; the result for the bad pair itself is whatever fp32 gives, 0 for NaN areas)."""
    b1 = O.generate_boxes(2000, alpha_range=(5, 100), beta_range=(5, 100), box="rbfov", seed=21).to(DEV)
    b2 = O.generate_boxes(2000, alpha_range=(5, 100), beta_range=(5, 100), box="rbfov", seed=22).to(DEV)
    good = api.iou.sph2pob_efficient_iou(b1, b2, is_aligned=True)
    bad1 = b1.clone()
    bad1[7] = float("nan"); bad1[8, 0] = float("inf"); bad1[9] = torch.tensor([1e9, -1e9, 1e9, 1e9, 1e9], device=DEV)
    bad1[10, 2:4] = 0.0; bad1[11, 4] = 720.0; bad1[12, 2] = -5.0
    got = api.iou.sph2pob_efficient_iou(bad1, b2, is_aligned=True)
    keep = torch.ones(2000, dtype=torch.bool, device=DEV); keep[7:13] = False
    assert torch.equal(got[keep], good[keep])
    mat = api.iou.sph2pob_efficient_iou(bad1[:64], b2[:600])
    ref = api.iou.sph2pob_efficient_iou(b1[:64], b2[:600])
    rows_ok = torch.ones(64, dtype=torch.bool, device=DEV); rows_ok[7:13] = False
    assert torch.equal(mat[rows_ok], ref[rows_ok])
    fin = torch.isfinite(mat[~rows_ok])
    assert bool(((mat[~rows_ok][fin] >= 0) & (mat[~rows_ok][fin] <= 1)).all())
    # A non-finite ROW box fails no test of the N x M scan loops (NaN compares false) and queues every column of its tile,
    # those past the end of a ragged last tile included: they must be dropped, not evaluated (they would read past the end
    # of bboxes2 and write into the next row / past the end of the matrix).  Last row = the bad one, both kernels
    # (k_iou_rows32: R <= 32; k_iou_pairwise2: R > 32), C not a multiple of the tile widths; the checked build
    # (tools/checked_run.sh, -DSPHK_CHECKED) turns a violation into a trap.
    for R in (20, 64):
        bad = b1[:R].clone()
        bad[R - 1] = float("nan"); bad[3, 1] = float("inf")
        ok_rows = torch.ones(R, dtype=torch.bool, device=DEV); ok_rows[R - 1] = False; ok_rows[3] = False
        for C in (600, 601, 1, 257):
            m = api.iou.sph2pob_efficient_iou(bad, b2[:C])
            assert m.shape == (R, C) and torch.equal(m[ok_rows], api.iou.sph2pob_efficient_iou(b1[:R], b2[:C])[ok_rows])
            rmax, rarg, cmax, carg = api.iou.sph_max_overlaps(bad, b2[:C])
            assert rmax.shape == (R,) and cmax.shape == (C,)


# ---- pairwise ----------------------------------------------------------------------------------
def test_pairwise_golden_both_orientations(api):
    g = load_golden("pairwise")
    for box in ("bfov", "rbfov"):
        rows, cols = cu(g[box + "_rows"]), cu(g[box + "_cols"])
        got = api.iou.sph2pob_efficient_iou(rows, cols)
        assert got.shape == (rows.size(0), cols.size(0))
        ok, err = within(got.cpu().numpy(), g[box + "_rc_f64"], g[box + "_rc_f32"])
        assert ok.all(), err[~ok]
        ok, err = within(api.iou.sph2pob_efficient_iou(cols, rows).cpu().numpy(), g[box + "_cr_f64"], g[box + "_cr_f32"])
        assert ok.all(), err[~ok]
    gt, anc = cu(g["assign_gt"]), cu(g["assign_anchors"])
    got = api.iou.SphOverlaps2D('sph2pob_efficient_iou', 5)(gt, anc).cpu().numpy()
    ok, err = within(got, g["assign_f64"], g["assign_f32"])
    assert ok.all(), err[~ok]
    assert ((got > 0) == (g["assign_f64"] > 1e-9)).mean() > 0.9999     # exact zeros stay exact zeros


@pytest.mark.parametrize("kind", ["sph2pob_efficient_iou", "sph2pob_standard_iou", "sph_iou", "fov_iou"])
@pytest.mark.parametrize("R,C", [(1, 1), (3, 700), (33, 257), (100, 31), (257, 5)])
def test_pairwise_equals_aligned_on_the_expansion(api, kind, R, C):
    """sph_iou_api.py:59-61: pair p = i*C + j is (bboxes1[i], bboxes2[j])."""
    box = "bfov" if kind in ("sph_iou", "fov_iou") else "rbfov"
    rows = O.generate_boxes(R, alpha_range=(5, 120), beta_range=(5, 120), box=box, seed=R).to(DEV)
    cols = O.generate_boxes(C, alpha_range=(5, 120), beta_range=(5, 120), box=box, seed=C + 1000).to(DEV)
    if R > 2 and C > 2:
        cols[1] = rows[2]
    fn = getattr(api.iou, kind)
    mat = fn(rows, cols)
    flat = fn(rows.repeat_interleave(C, 0), cols.repeat(R, 1), is_aligned=True)
    if kind in ("sph_iou", "fov_iou"):
        assert torch.equal(mat.reshape(-1), flat)
    else:
        # the N x M kernel evaluates far-apart pairs through its precompute/fast path, the aligned kernel
        # through the reference-order path: equal up to fp32 rounding, and exact zeros agree
        assert float((mat.reshape(-1) - flat).abs().max()) < 2e-6
        assert torch.equal(mat.reshape(-1) == 0, flat == 0)


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
@pytest.mark.parametrize("R,C", [(1, 1), (2, 63), (5, 64), (8, 65), (17, 1000), (31, 4097), (32, 98208)])
def test_few_rows_single_launch_kernel(api, box, R, C):
    """k_iou_rows32 (R <= 32: the per-image call of MaxIoUAssigner, records computed inside the CTAs) against the general
    two-kernel path: the same rows inside a call with more than 32 rows go through k_box_pre + k_iou_pairwise2.  Same
    per-box records, same fast / reference-order code: bit-identical results, for both transforms and IoF, on views
    of a wider output and for unaligned inputs."""
    rows = O.generate_boxes(R, alpha_range=(5, 120), beta_range=(5, 120), box=box, seed=R).to(DEV)
    cols = O.generate_boxes(C, alpha_range=(1, 150), beta_range=(1, 150), box=box, seed=C + 7).to(DEV)
    if C > 40:
        cols[33] = rows[0]                                    # similarity-mask pair -> reference-order path
        cols[C - 1, 2:4] = 400.0                               # oversize anchor (clamped by jitter_1)
    pad = O.generate_boxes(40, alpha_range=(5, 120), beta_range=(5, 120), box=box, seed=99).to(DEV)
    for fn, kw in ((api.iou.sph2pob_efficient_iou, {}), (api.iou.sph2pob_standard_iou, {}),
                   (api.iou.sph2pob_efficient_iou, dict(mode="iof")), (api.iou.sph2pob_efficient_iou, dict(rbb_edge="chord"))):
        got = fn(rows, cols, **kw)
        want = fn(torch.cat([rows, pad]), cols, **kw)[:R]
        assert got.shape == (R, C) and torch.equal(got, want), (fn.__name__, kw, float((got - want).abs().max()))
    # a view of a wider output (ld > C, rows not 16-byte aligned): only the view is written
    wide = torch.full((R, C + 7), -1.0, device=DEV)
    api.native.iou_pairwise("sph2pob_efficient", rows, cols, out=wide[:, 3:3 + C])
    assert torch.equal(wide[:, 3:3 + C], api.iou.sph2pob_efficient_iou(rows, cols))
    assert bool((wide[:, :3] == -1).all()) and bool((wide[:, 3 + C:] == -1).all())
    # unaligned input pointers (D = 4 takes the scalar loads then)
    buf = torch.zeros(R * rows.size(1) + 1, device=DEV)
    buf[1:] = rows.reshape(-1)
    assert torch.equal(api.iou.sph2pob_efficient_iou(buf[1:].view(R, -1), cols), api.iou.sph2pob_efficient_iou(rows, cols))


def test_fused_max_argmax(api):
    rows = O.generate_boxes(70, alpha_range=(5, 120), beta_range=(5, 120), box="rbfov", seed=1).to(DEV)
    cols = O.generate_boxes(3000, alpha_range=(5, 120), beta_range=(5, 120), box="rbfov", seed=2).to(DEV)
    cols[10] = cols[2000] = rows[3]                     # tie: lowest index must win
    rows[60] = torch.tensor([0.0, 0.0, 1.0, 1.0, 0.0])  # a row that overlaps (almost) nothing
    for r, c in ((rows, cols), (cols, rows)):
        mat = api.iou.sph2pob_efficient_iou(r, c)
        rmax, rarg, cmax, carg, mat2 = api.iou.sph_max_overlaps(r, c, return_matrix=True)
        assert torch.equal(mat, mat2)
        assert torch.equal(rmax, mat.max(dim=1)[0]) and torch.equal(cmax, mat.max(dim=0)[0])
        m = mat.cpu().numpy()
        assert rarg.cpu().tolist() == [int(np.flatnonzero(m[i] == m[i].max())[0]) for i in range(m.shape[0])]
        assert carg.cpu().tolist() == [int(np.flatnonzero(m[:, j] == m[:, j].max())[0]) for j in range(m.shape[1])]
        # no matrix: same numbers
        rmax2, rarg2, cmax2, carg2 = api.iou.sph_max_overlaps(r, c)
        assert torch.equal(rmax, rmax2) and torch.equal(rarg, rarg2) and torch.equal(cmax, cmax2) and torch.equal(carg, carg2)
        # shard offsets are added to the reported indices
        _, rarg3, _, carg3 = api.iou.sph_max_overlaps(r, c, row_base=1000, col_base=5000)
        assert torch.equal(rarg3, rarg + 5000) and torch.equal(carg3, carg + 1000)


@pytest.mark.parametrize("R,C", [(1, 70), (7, 64), (20, 3000), (32, 98208)])
def test_fused_max_argmax_few_rows(api, R, C):
    """Fused max / argmax for the per-image shape (R <= 32 GT rows; the matrix itself comes from k_iou_rows32, the fused
    maxima from k_iou_pairwise2): equal to max / first argmax of the matrix, ties to the lowest index, empty rows / columns
    report (0, base), shard offsets are added."""
    rows = O.generate_boxes(R, alpha_range=(5, 120), beta_range=(5, 120), box="rbfov", seed=R).to(DEV)
    cols = O.generate_boxes(C, alpha_range=(5, 120), beta_range=(5, 120), box="rbfov", seed=C).to(DEV)
    cols[10] = cols[C - 3] = rows[0]                                  # tie across CTAs: lowest column index must win
    if R > 4:
        rows[4] = rows[2]                                             # tie inside a column: lowest row index must win
        rows[R - 1] = torch.tensor([0.0, 0.0, 1.0, 1.0, 0.0], device=DEV)   # a row that overlaps (almost) nothing
    mat = api.iou.sph2pob_efficient_iou(rows, cols)
    rmax, rarg, cmax, carg, mat2 = api.iou.sph_max_overlaps(rows, cols, return_matrix=True)
    assert torch.equal(mat, mat2)
    assert torch.equal(rmax, mat.max(dim=1)[0]) and torch.equal(cmax, mat.max(dim=0)[0])
    m = mat.cpu().numpy()
    assert rarg.cpu().tolist() == [int(np.flatnonzero(m[i] == m[i].max())[0]) for i in range(R)]
    assert np.array_equal(carg.cpu().numpy(), (m == m.max(axis=0, keepdims=True)).argmax(axis=0))
    rmax2, rarg2, cmax2, carg2 = api.iou.sph_max_overlaps(rows, cols)
    assert torch.equal(rmax, rmax2) and torch.equal(rarg, rarg2) and torch.equal(cmax, cmax2) and torch.equal(carg, carg2)
    _, rarg3, _, carg3 = api.iou.sph_max_overlaps(rows, cols, row_base=1000, col_base=5000)
    assert torch.equal(rarg3, rarg + 5000) and torch.equal(carg3, carg + 1000)


@pytest.mark.parametrize("world", [3, 8])
def test_sharding_emulated_on_one_gpu(api, world):
    """`world` 'ranks' processed one after the other on one GPU: every rank's kernel writes its packed keys into its own
    communication block, the blocks are stacked (what all_gather_into_tensor delivers: tests/test_sharded_gloo.py) and
    sphk_unpack_gathered_keys turns them into the global result -- equal to the unsharded call, to the torch
    restatement of the unpack, and with the lowest-index tie rule across shards."""
    from sph_retina_b200 import _native
    from sph_retina_b200.sharded import block_capacity, key_block, shard_bounds, sharded_max_overlaps
    from test_sharded_gloo import unpack_gathered_reference
    n = 5003                                   # uneven split: the last shards are one row short (padding slots)
    A = O.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=0).to(DEV)
    G = O.generate_boxes(300, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=1).to(DEV)
    A[4000] = A[17] = G[5]
    cap = block_capacity(n, world)
    for orient in ("bboxes1", "bboxes2"):
        a_max, a_arg, g_max, g_arg = sharded_max_overlaps(A, G, n, 0, anchors_are=orient)   # world = 1
        if orient == "bboxes1":
            rm, ra, cm, ca = api.iou.sph_max_overlaps(A, G)
        else:
            cm, ca, rm, ra = api.iou.sph_max_overlaps(G, A)
        assert torch.equal(a_max, rm) and torch.equal(g_max, cm)
        assert torch.equal(a_arg[rm > 0], ra[rm > 0]) and torch.equal(g_arg[cm > 0], ca[cm > 0])
        assert a_arg.dtype == torch.int64 and g_arg.dtype == torch.int64
        blocks = []
        for rank in range(world):
            lo, hi = shard_bounds(n, world, rank)
            blk = key_block(n, G.size(0), world, DEV, fresh=True)
            if orient == "bboxes1":
                _native.iou_pairwise_keys("sph2pob_efficient", A[lo:hi], G, row_base=lo, row_keys_out=blk[:hi - lo], col_keys_out=blk[cap:])
            else:
                _native.iou_pairwise_keys("sph2pob_efficient", G, A[lo:hi], col_base=lo, row_keys_out=blk[cap:], col_keys_out=blk[:hi - lo])
            blocks.append(blk)
        gathered = torch.stack(blocks)
        got = _native.unpack_gathered_keys(gathered, world, n, G.size(0), cap)
        want = unpack_gathered_reference(gathered, world, n, G.size(0), cap)
        for g_, w_, s_ in zip(got, want, (a_max, a_arg, g_max, g_arg)):
            assert torch.equal(g_, w_) and torch.equal(g_, s_)
        assert int(g_arg[5]) == 17
        # the routes without a collective: every "rank" owns a buffer [even steps: world slots | odd steps: world slots |
        # flags], slot s = [parts x cap anchor keys | G]; here they all sit on this GPU, and the flags are pre-set to the
        # step as if every peer had arrived (the handshake itself needs one process per GPU: bench.py --gpus N).
        # Step 3 -> the odd area.
        #   pull (parts = 1): every rank has filled its own slot only, the unpack launch reads slot s from rank s's buffer;
        #   push (anchors = rows): sphk_iou_pairwise_keys_push stores every tile's keys, one partial array per column tile,
        #         into slot r of EVERY buffer while it runs; the unpack launch takes the maximum over the LOCAL partials.
        step = 3
        for pushed in ((False, True) if orient == "bboxes1" else (False,)):
            parts = _native.key_push_parts(G.size(0)) if pushed else 1
            assert parts == (2 if pushed else 1)
            per = parts * cap + G.size(0)
            area = world * per
            bufs = [torch.full((2 * area + 32,), 0x7777 if pushed else 0, dtype=torch.int64, device=DEV) for _ in range(world)]
            table = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=DEV)
            for rank in range(world):
                lo, hi = shard_bounds(n, world, rank)
                off = area + rank * per
                if not pushed:
                    bufs[rank][off:off + per] = blocks[rank]
                else:       # (the buffers start as garbage: the push route needs no zero-fill of the rows it owns)
                    _native.iou_pairwise_keys_push("sph2pob_efficient", A[lo:hi], G, bufs[rank][off + parts * cap:off + per],
                                                   table.data_ptr(), world, off, cap, row_base=lo)
                bufs[rank][2 * area:2 * area + world] = step
            if pushed:
                for rank in range(world):               # every buffer holds every rank's partial keys; their maximum = the key
                    for src in range(world):
                        lo, hi = shard_bounds(n, world, src)
                        at = area + src * per
                        both = torch.maximum(bufs[rank][at:at + hi - lo], bufs[rank][at + cap:at + cap + hi - lo])
                        assert torch.equal(both, blocks[src][:hi - lo])
                    at = area + rank * per + parts * cap
                    assert torch.equal(bufs[rank][at:at + G.size(0)], blocks[rank][cap:])
            for rank in range(world):
                peer = _native.unpack_peer_keys(table.data_ptr(), rank, world, step, area, 2 * area, n, G.size(0), cap, torch.device(DEV),
                                                long_parts=parts, long_pushed=pushed)
                for p_, g_ in zip(peer, got):
                    assert torch.equal(p_, g_)
                assert int(bufs[(rank + 1) % world][2 * area + rank]) == step       # the rank raised its flag in the peer's buffer
    # "no positive overlap" (key 0) reads as (0.0, index 0) on both sides
    far = torch.tensor([[10.0, 90.0, 0.5, 0.5, 0.0], [190.0, 90.0, 20.0, 20.0, 0.0], [12.0, 40.0, 0.5, 0.5, 0.0]], device=DEV)
    one = torch.tensor([[190.0, 90.0, 20.0, 20.0, 0.0], [300.0, 150.0, 1.0, 1.0, 0.0]], device=DEV)
    a_max, a_arg, g_max, g_arg = sharded_max_overlaps(far, one, 3, 0)
    assert a_max.tolist()[0] == 0.0 and a_max.tolist()[2] == 0.0 and a_max.tolist()[1] > 0.9
    assert a_arg.tolist() == [0, 0, 0] and g_arg.tolist() == [1, 0] and g_max.tolist()[1] == 0.0


@pytest.mark.parametrize("chunks", [1, 3])
def test_host_sweep_pipeline_equals_the_device_call(api, chunks):
    """HostSweep (pinned host in / out, H2D + kernels + D2H pipelined over row chunks, per-GT keys accumulated over the
    chunks) returns what sharded_max_overlaps returns for device-resident boxes."""
    from sph_retina_b200.sharded import HostSweep, sharded_max_overlaps
    n, G = 7001, 300
    A = O.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=5)
    Gt = O.generate_boxes(G, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=6)
    A[6000] = A[11] = Gt[7]                                   # a tie across chunks: lowest index wins
    want = sharded_max_overlaps(A.to(DEV), Gt.to(DEV), n, 0)
    hs = HostSweep(n, n, G, 5, DEV, min_chunk_rows=1, max_chunks=chunks)
    assert hs.chunks == chunks
    assert [b for b in HostSweep(1 << 20, 1 << 20, G, 5, DEV).bounds] == [(0, 32768), (32768, 524288), (524288, 1015808), (1015808, 1048576)]
    assert HostSweep(1 << 20, 1 << 17, G, 5, DEV).bounds == [(0, 16384), (16384, 114688), (114688, 131072)]
    outs = (torch.empty(n).pin_memory(), torch.empty(n, dtype=torch.int64).pin_memory(),
            torch.empty(G).pin_memory(), torch.empty(G, dtype=torch.int64).pin_memory())
    for _ in range(2):                                        # the second step re-uses every staging buffer
        for o in outs:
            o.fill_(-7)
        hs(A.pin_memory(), Gt.pin_memory(), 0, *outs)
        torch.cuda.synchronize()
        for got, w in zip(outs, want):
            assert torch.equal(got, w.cpu())
    assert int(outs[3][7]) == 11


def test_config2_slice_vs_c_oracle(api, c_oracle):
    """Assignment orientation (GT rows x anchor cols) on a strided sample of the real 512x1024 anchor grid."""
    g = load_golden("pairwise")
    gt, anc = g["assign_gt"], g["assign_anchors"]
    R, C = len(gt), len(anc)
    want = c_oracle_aligned(c_oracle, 0, np.repeat(gt, C, axis=0), np.tile(anc, (R, 1))).reshape(R, C)
    got = api.iou.sph2pob_efficient_iou(cu(gt), cu(anc)).cpu().numpy()
    assert np.abs(got - want).max() < 1e-5


# ---- fused assignment -------------------------------------------------------
@pytest.mark.parametrize("assign_all", [True, False])
@pytest.mark.parametrize("min_pos_iou", [0.0, 0.3])
def test_fused_assigner_equals_reference_on_the_matrix(api, assign_all, min_pos_iou):
    """SphMaxIoUAssigner (no matrix) == the reference's assign_wrt_overlaps loop run on the matrix."""
    from sph_retina_b200 import synthetic as S
    from sph_retina_b200.sphdet.assigners import SphMaxIoUAssigner
    anchors = S.retina_anchors()[::3].contiguous()
    anchors[100] = anchors[5000] = anchors[7]           # duplicated anchors: ties on the row maxima
    gts = S.generate_boxes(32, alpha_range=(5, 120), beta_range=(5, 120), box="rbfov", seed=100)
    gts[3] = torch.tensor([10.0, 90.0, 0.5, 0.5, 0.0])   # a tiny GT no anchor reaches 0.3 with
    labels = torch.randint(0, 37, (32,))
    anchors, gts, labels = anchors.to(DEV), gts.to(DEV), labels.to(DEV)
    calc = api.iou.SphOverlaps2D('sph2pob_efficient_iou', 5)
    a = SphMaxIoUAssigner(0.5, 0.3, min_pos_iou=min_pos_iou, gt_max_assign_all=assign_all, iou_calculator=calc)
    res = a.assign(anchors, gts, gt_labels=labels)
    overlaps = calc(gts, anchors)
    # argmax ties resolve to the LOWEST index in the kernels; torch.max documents no rule, so the literal loop is run with
    # that rule stated explicitly (oracle: lowest_index_ties) and the comparison is exact for both settings
    g, m, l = O.assign_wrt_overlaps(overlaps.cpu(), labels.cpu(), 0.5, 0.3, min_pos_iou, assign_all, True, lowest_index_ties=True)
    assert res.num_gts == 32 and res.gt_inds.dtype == torch.int64
    assert torch.equal(res.max_overlaps.cpu(), m)
    assert torch.equal(res.gt_inds.cpu(), g) and torch.equal(res.labels.cpu(), l)
    ties = int(((overlaps == overlaps.max(dim=1, keepdim=True)[0]).sum(dim=1) > 1).sum())
    parity_record("assigner_exact", n=int(g.numel()), n_fail_before_allowances=int((res.gt_inds.cpu() != g).sum()), rows_with_tied_maximum=ties,
                  gt_max_assign_all=bool(assign_all))
    assert ties >= 1                                       # the duplicated anchors do produce tied row maxima
    assert int((res.gt_inds > 0).sum()) > 100 and int((res.gt_inds == 0).sum()) > 1000 and int((res.gt_inds == -1).sum()) > 10


def test_fused_assigner_batch_equals_per_image(api):
    from sph_retina_b200 import synthetic as S
    from sph_retina_b200.sphdet.assigners import SphMaxIoUAssigner
    anchors = S.retina_anchors()[::5].contiguous().to(DEV)
    calc = api.iou.SphOverlaps2D('sph2pob_efficient_iou', 5)
    counts = [32, 0, 7, 1, 33, 64]
    gts = [S.generate_boxes(k, alpha_range=(5, 120), beta_range=(5, 120), box="rbfov", seed=200 + i).to(DEV) for i, k in enumerate(counts)]
    labels = [torch.randint(0, 37, (k,), device=DEV) for k in counts]
    for assign_all in (True, False):
        a = SphMaxIoUAssigner(0.5, (0.05, 0.3), min_pos_iou=0.1, gt_max_assign_all=assign_all, iou_calculator=calc)
        batch = a.assign_batch(anchors, gts, labels)
        assert len(batch) == len(counts)
        for b, k in enumerate(counts):
            ov = calc(gts[b], anchors) if k else anchors.new_zeros((0, anchors.size(0)))
            g, m, l = O.assign_wrt_overlaps(ov.cpu(), labels[b].cpu(), 0.5, (0.05, 0.3), 0.1, assign_all, True, lowest_index_ties=True)
            assert batch[b].num_gts == k and torch.equal(batch[b].max_overlaps.cpu(), m)
            assert torch.equal(batch[b].gt_inds.cpu(), g) and torch.equal(batch[b].labels.cpu(), l)
            one = a.assign(anchors, gts[b], gt_labels=labels[b])
            assert torch.equal(one.gt_inds, batch[b].gt_inds) and torch.equal(one.max_overlaps, batch[b].max_overlaps)


def test_fused_assigner_corner_cases(api):
    from sph_retina_b200.sphdet.assigners import SphMaxIoUAssigner
    calc = api.iou.SphOverlaps2D('sph2pob_efficient_iou', 4)
    a = SphMaxIoUAssigner(0.5, 0.4, iou_calculator=calc)
    boxes = O.generate_boxes(500, alpha_range=(5, 60), beta_range=(5, 60), seed=1).to(DEV)
    gts = O.generate_boxes(6, alpha_range=(5, 60), beta_range=(5, 60), seed=2).to(DEV)
    # no GT -> everything background; no boxes -> empty
    r = a.assign(boxes, gts[:0])
    assert r.num_gts == 0 and bool((r.gt_inds == 0).all())
    assert a.assign(boxes[:0], gts).gt_inds.numel() == 0
    # a GT far from every box (best overlap exactly 0, min_pos_iou = 0): mmdet assigns every zero-overlap box to it
    far = torch.tensor([[0.5, 0.5, 1.0, 1.0]], device=DEV)
    near = boxes[:50].clone(); near[:, 1] = near[:, 1].clamp(min=60)
    r = a.assign(near, torch.cat([gts[:2], far]))
    ov = calc(torch.cat([gts[:2], far]), near)
    g, m, _ = O.assign_wrt_overlaps(ov.cpu(), None, 0.5, 0.4, 0.0, True, True)
    assert torch.equal(r.gt_inds.cpu(), g) and torch.equal(r.max_overlaps.cpu(), m)
    # ignore region (matrix path)
    ai = SphMaxIoUAssigner(0.5, 0.4, ignore_iof_thr=0.5, iou_calculator=calc)
    r = ai.assign(boxes, gts, gt_bboxes_ignore=boxes[:3])
    ov = calc(gts, boxes)
    ign = calc(boxes, boxes[:3], mode='iof').max(dim=1)[0]
    ov[:, ign > 0.5] = -1
    g, m, _ = O.assign_wrt_overlaps(ov.cpu(), None, 0.5, 0.4, 0.0, True, True)
    assert torch.equal(r.gt_inds.cpu(), g) and bool((r.gt_inds[:3] == -1).all())


# ---- loss --------------------------------------------------------------------------------------
def grad_row_error(got, truth):
    den = np.linalg.norm(truth, axis=1)
    return np.linalg.norm(got - truth, axis=1) / np.maximum(den, 1e-12), den


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
@pytest.mark.parametrize("mode", ["iou", "giou", "diou", "ciou"])
def test_loss_forward_backward_golden(api, box, mode):
    g = load_golden("loss_" + box)
    p = cu(g["pred"]).requires_grad_(True)
    t = cu(g["target"]).requires_grad_(True)
    L = api.losses.Sph2PobIoULoss(mode=mode, reduction="sum")
    el = L(p, t, reduction_override="none")
    assert el.shape == (p.size(0),)
    el.sum().backward()
    ok, err = within(el.detach().cpu().numpy(), g[mode + "_loss_f64"], g[mode + "_loss_f32"], tol=2e-5 if mode != "iou" else 1e-5)
    assert ok.all(), (np.where(~ok)[0], err[~ok])
    for got, key in ((p.grad, "gpred"), (t.grad, "gtarget")):
        truth, ref32 = g["%s_%s_f64" % (mode, key)], g["%s_%s_f32" % (mode, key)]
        rel, den = grad_row_error(got.cpu().numpy(), truth)
        rel32, _ = grad_row_error(ref32, truth)
        live = den > 1e-9
        assert np.abs(got.cpu().numpy()[~live]).max(initial=0.0) < 1e-6
        good = (rel <= 1e-4) | (rel <= rel32)
        assert good[live].mean() > 0.995, (key, (~good & live).sum())
        assert np.median(rel[live]) < 3e-6
        assert (rel[live] > 1e-4).sum() < 0.5 * (rel32[live] > 1e-4).sum()


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_loss_reductions_weights_avg_factor(api, box):
    g = load_golden("loss_" + box)
    p, t, w1, w2 = cu(g["pred"]), cu(g["target"]), cu(g["w1"]), cu(g["w2"])
    L = api.losses.Sph2PobIoULoss(mode="iou", loss_weight=2.0)
    np.testing.assert_allclose(L(p, t).item(), g["red_mean"], rtol=2e-5)
    np.testing.assert_allclose(L(p, t, w1, avg_factor=123.0).item(), g["red_w1_avg"], rtol=2e-5)
    np.testing.assert_allclose(L(p, t, w2).item(), g["red_w2"], rtol=2e-5)
    np.testing.assert_allclose(L(p, t, w1, reduction_override="sum").item(), g["red_w1_sum"], rtol=2e-5)
    pz = p.clone().requires_grad_(True)
    z = L(pz, t, torch.zeros_like(w1))
    assert z.item() == 0.0
    z.backward()
    assert float(pz.grad.abs().max()) == 0.0
    with pytest.raises(ValueError):
        L(p, t, w1, avg_factor=3.0, reduction_override="sum")
    # weighted rows with weight 0 receive exactly zero gradient; no-grad call works
    pw = p.clone().requires_grad_(True)
    L(pw, t, w1, avg_factor=50.0).backward()
    assert float(pw.grad[w1 == 0].abs().max()) == 0.0 and torch.isfinite(pw.grad).all()
    with torch.no_grad():
        assert torch.isfinite(L(p, t))


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
@pytest.mark.parametrize("mode", ["log", "linear", "square"])
def test_legacy_iou_loss_golden(api, box, mode):
    """SphIoULossLegacy (sph2pob_iou_loss.py:199-216; mmrotate's RotatedIoULoss on the Sph2Pob OBBs) against the reference's
    own class: per-row loss (compared as the IoU it encodes for 'log': -log amplifies a 1e-5 IoU error of a 1e-3 IoU to
    1e-2) and both gradients, fp64 truth with the reference's fp32 run as the yardstick."""
    g, base = load_golden("legacy_loss"), load_golden("loss_" + box)
    p, t = cu(base["pred"]).requires_grad_(True), cu(base["target"]).requires_grad_(True)
    el = api.losses.SphIoULossLegacy(mode=mode, reduction="sum")(p, t, reduction_override="none")
    assert el.shape == (p.size(0),)
    el.sum().backward()
    key = "%s_%s_" % (box, mode)
    as_iou = (lambda x: np.exp(-np.asarray(x, np.float64))) if mode == "log" else (lambda x: np.asarray(x, np.float64))
    ok, err = within(as_iou(el.detach().cpu().numpy()), as_iou(g[key + "loss_f64"]), as_iou(g[key + "loss_f32"]), tol=2e-5)
    assert ok.all(), (np.where(~ok)[0], err[~ok])
    for got, name in ((p.grad, "gpred"), (t.grad, "gtarget")):
        truth, ref32 = g[key + name + "_f64"], g[key + name + "_f32"]
        rel, den = grad_row_error(got.cpu().numpy(), truth)
        rel32, _ = grad_row_error(ref32, truth)
        live = den > 1e-9
        assert np.abs(got.cpu().numpy()[~live]).max(initial=0.0) < 1e-6
        good = (rel <= 1e-4) | (rel <= rel32)
        assert good[live].mean() > 0.995, (name, (~good & live).sum())
        assert np.median(rel[live]) < 3e-6


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_legacy_iou_loss_reductions(api, box):
    g, base = load_golden("legacy_loss"), load_golden("loss_" + box)
    p, t, w1, w2 = cu(base["pred"]), cu(base["target"]), cu(base["w1"]), cu(base["w2"])
    L = api.losses.SphIoULossLegacy(loss_weight=2.0)
    np.testing.assert_allclose(L(p, t).item(), g[box + "_red_mean"], rtol=1e-4)
    np.testing.assert_allclose(L(p, t, w1, avg_factor=123.0).item(), g[box + "_red_w1_avg"], rtol=1e-4)
    np.testing.assert_allclose(L(p, t, w2).item(), g[box + "_red_w2"], rtol=1e-4)
    np.testing.assert_allclose(L(p, t, w1, reduction_override="sum").item(), g[box + "_red_w1_sum"], rtol=1e-4)
    np.testing.assert_allclose(api.losses.SphIoULossLegacy(linear=True)(p, t, w1).item(), g[box + "_red_linear"], rtol=2e-5)
    pz = p.clone().requires_grad_(True)
    z = L(pz, t, torch.zeros_like(p))
    assert z.item() == 0.0 == float(g[box + "_red_zero_w"])
    z.backward()
    assert float(pz.grad.abs().max()) == 0.0


def test_loss_gradcheck_against_finite_differences(api):
    """Independent of the reference: central differences of the kernel's own forward (fp32, so loose)."""
    t = O.generate_boxes(512, alpha_range=(20, 80), beta_range=(20, 80), box="rbfov", seed=2)
    p = (t + torch.randn(512, 5) * torch.tensor([4, 4, 4, 4, 8.0])).clamp(min=5)
    p, t = p.to(DEV), t.to(DEV)
    pr = p.clone().requires_grad_(True)
    iou = api.losses.sph2pob_iou(pr, t)
    iou.sum().backward()
    h = 0.05
    num = torch.zeros_like(p)
    for k in range(5):
        d = torch.zeros_like(p); d[:, k] = h
        num[:, k] = (api.losses.sph2pob_iou(p + d, t) - api.losses.sph2pob_iou(p - d, t)) / (2 * h)
    err = (pr.grad - num).norm(dim=1) / num.norm(dim=1).clamp(min=1e-3)
    assert float(err.median()) < 5e-3 and float((err < 5e-2).float().mean()) > 0.97


@pytest.mark.parametrize("decoded", [True, False])
def test_anchor_targets_batch(api, decoded):
    """get_targets_batch (sphk_max_iou_assign + sphk_anchor_targets) against the literal per-image reference chain run on
    the kernel's own overlaps: assign_wrt_overlaps -> PseudoSampler -> anchor_head._get_targets_single.  Labels, weights and
    counts are exact; decoded targets are the GT boxes bit for bit, encoded ones match bbox2delta to fp32 rounding."""
    from sph_retina_b200 import synthetic as S
    from sph_retina_b200.sphdet.assigners import SphMaxIoUAssigner
    from sph_retina_b200.sphdet.bbox.coder import DeltaXYWHASphBBoxCoder
    from sph_retina_b200.sphdet.models.heads import get_targets_batch
    gts, anchors = S.assignment_batch(images=4)
    anchors = anchors[::3].contiguous().to(DEV)
    gt_list = [gts[0].to(DEV), gts[1][:5].to(DEV), gts[2][:0].to(DEV), gts[3].to(DEV)]          # one image without GT
    gen = torch.Generator().manual_seed(0)
    lab_list = [torch.randint(0, 80, (g.size(0),), generator=gen).to(DEV) for g in gt_list]
    A = SphMaxIoUAssigner(0.5, 0.4, min_pos_iou=0, iou_calculator=dict(type='SphOverlaps2D', backend='sph2pob_efficient_iou', box_version=5))
    coder = DeltaXYWHASphBBoxCoder(target_stds=(0.1, 0.1, 0.2, 0.2, 0.1))
    labels, lw, bt, bw, npos, nneg = get_targets_batch(anchors, gt_list, lab_list, A, 80, bbox_coder=coder,
                                                       reg_decoded_bbox=decoded, pos_weight=-1)
    assert labels.shape == (4, anchors.size(0)) and bt.shape == (4, anchors.size(0), 5)
    tot_pos = tot_neg = 0
    for b, (g, l) in enumerate(zip(gt_list, lab_list)):
        ov = api.iou.SphOverlaps2D('sph2pob_efficient_iou', 5)(g, anchors).cpu()
        gi, _, _ = O.assign_wrt_overlaps(ov, None, 0.5, 0.4, 0.0, True, True)
        want = O.get_targets_single(anchors.cpu(), g.cpu(), l.cpu(), gi, 80, reg_decoded_bbox=decoded, pos_weight=-1,
                                    means=None, stds=[0.1, 0.1, 0.2, 0.2, 0.1])
        assert torch.equal(labels[b].cpu(), want[0]) and torch.equal(lw[b].cpu(), want[1]) and torch.equal(bw[b].cpu(), want[3])
        if decoded:
            assert torch.equal(bt[b].cpu(), want[2])
        else:
            assert float((bt[b].cpu() - want[2]).abs().max()) < 2e-5
        tot_pos += max(want[4], 1)
        tot_neg += max(want[5], 1)
    assert (npos, nneg) == (tot_pos, tot_neg) and npos > 100
    # pos_weight > 0, RPN-style call without labels, and the no-sync variant
    _, lw2, _, _, cnt, _ = get_targets_batch(anchors, gt_list, None, A, 1, reg_decoded_bbox=True, pos_weight=2.5, sync_counts=False)
    assert torch.equal(lw2 == 2.5, bw[..., 0] == 1) and cnt.shape == (4, 2) and int(cnt[2, 0]) == 0 and int(cnt[2, 1]) == anchors.size(0)


# ---- naive_iou (planar IoU of the sph2pix boxes) and SphNMS('naive_iou') ------------------------------------------
@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_naive_iou_golden(api, box):
    g = load_golden("naive")
    b1, b2 = cu(g[box + "_b1"]), cu(g[box + "_b2"])
    got = api.iou.naive_iou(b1, b2, is_aligned=True)
    ok, err = within(got.cpu().numpy(), g[box + "_aligned_f64"], g[box + "_aligned_f32"])
    assert got.shape == (b1.size(0),) and ok.all(), (np.where(~ok)[0][:10], err[~ok][:10])
    mat = api.iou.SphOverlaps2D('naive_iou', box_version=b1.size(1))(b1[:37], b2[:301])
    ok, err = within(mat.cpu().numpy(), g[box + "_rc_f64"], g[box + "_rc_f32"])
    assert mat.shape == (37, 301) and ok.all(), err[~ok][:10]
    # N x M equals the aligned call on the expansion (same per-pair function; the two kernels contract its FMAs differently)
    flat = api.iou.naive_iou(b1[:37].repeat_interleave(301, 0), b2[:301].repeat(37, 1), is_aligned=True)
    assert float((mat.reshape(-1) - flat).abs().max()) < 2e-6 and torch.equal(mat.reshape(-1) == 0, flat == 0)
    rmax, rarg, cmax, carg = api.iou.sph_max_overlaps(b1[:37], b2[:301], backend='naive_iou')
    assert torch.equal(rmax, mat.max(dim=1)[0]) and torch.equal(cmax, mat.max(dim=0)[0])


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_naive_nms_golden_keep_sets(api, box):
    """SphNMS('naive_iou') -- what the reference's indoor360 configs run at test time -- keep lists equal to the reference's
    wherever no pair IoU is within 1e-5 of the threshold (BASELINE.json north_star); both the single-image pipeline and the
    batched per-segment path."""
    g = load_golden("naive")
    boxes, scores, idxs = cu(g[box + "_boxes"]), cu(g[box + "_scores"]), cu(g[box + "_idxs"])
    iou = g[box + "_pair_iou_f64"]
    same = g[box + "_idxs"][:, None] == g[box + "_idxs"][None, :]
    for thr in (0.3, 0.5):
        assert not (np.abs(iou[same] - thr) < 1e-5).any(), "fixture has a pair on the threshold"
        want = g["%s_keep_thr%d" % (box, int(thr * 10))].tolist()
        dets, keep = api.nms.SphNMS('naive_iou')(boxes, scores, idxs, dict(type="nms", iou_threshold=thr, max_num=150))
        assert keep.cpu().tolist() == want and dets.shape == (len(want), boxes.size(1) + 1)
        from sph_retina_b200.sphdet.bbox.nms import sph_batched_nms_images
        kept = sph_batched_nms_images(boxes, scores, idxs, torch.zeros_like(idxs), thr, iou_calculator='naive_iou')
        assert kept.cpu().tolist()[:150] == want
    if box == "bfov":   # and it is a different calculator: the Sph2Pob NMS keeps another set on this fixture
        _, keep_sph = api.nms.SphNMS()(boxes, scores, idxs, dict(iou_threshold=0.5, max_num=150))
        assert keep_sph.cpu().tolist() != g[box + "_keep_thr5"].tolist()


def test_planar_nms_golden_keep_sets(api):
    """PlanarNMS (test_cfg.iou_calculator = 'planar', planar_nms.py:7-18) against the keep lists of the reference's class:
    class-agnostic (its default) and per class, max_num and score_threshold of mmcv's nms."""
    g, base = load_golden("planar_nms"), load_golden("naive")
    boxes, scores, idxs = cu(base["bfov_boxes"]), cu(base["bfov_scores"]), cu(base["bfov_idxs"])
    iou = base["bfov_pair_iou_f64"]
    from sph_retina_b200.sphdet.bbox.nms import PlanarNMS
    for thr in (0.3, 0.5):
        assert not (np.abs(iou - thr) < 1e-5).any(), "fixture has a pair on the threshold"
        for tag, kw in (("agnostic", {}), ("per_class", dict(class_agnostic=False))):
            dets, keep = PlanarNMS()(boxes, scores, idxs, dict(type="nms", iou_threshold=thr), **kw)
            want = g["keep_%s_thr%d" % (tag, int(thr * 10))]
            assert keep.cpu().tolist() == want.tolist()
            np.testing.assert_allclose(dets.cpu().numpy(), g["dets_%s_thr%d" % (tag, int(thr * 10))], rtol=0, atol=0)
    _, keep = PlanarNMS()(boxes, scores, idxs, dict(type="nms", iou_threshold=0.5, max_num=40, score_threshold=0.2))
    assert keep.cpu().tolist() == g["keep_max40_score02"].tolist()
    _, keep = PlanarNMS()(boxes, scores, idxs, None)
    assert torch.equal(scores[keep], scores.sort(descending=True)[0])


def test_planar_nms_zero_area_boxes_follow_mmcv(api):
    """Two zero-area boxes on the same spot have IoU 0 / 0: mmcv's nms (`inter > thr * union`; PlanarNMS, planar_nms.py:16)
    keeps them all, SphNMS('naive_iou') (`ious <= thr` keeps, sph_nms.py:70) drops every zero-area box after the first --
    the union of any two of them is 0, wherever they are.  Both rules live in the NMS
    kernel (SPHK_NMS_RULE_GT); the planar one is checked against the oracle's restatement of mmcv's nms."""
    from sph_retina_b200.sphdet.bbox.nms import PlanarNMS, SphNMS
    boxes = O.generate_boxes(60, alpha_range=(5, 60), beta_range=(5, 60), box="bfov", seed=21)
    boxes[10:20, 2:] = 0.0                          # zero-area boxes ...
    boxes[15:20, :2] = boxes[10:15, :2]             # ... five of them on top of five others
    boxes[30] = boxes[31]                           # and an ordinary duplicate
    scores = torch.linspace(0.95, 0.05, 60)
    idxs = torch.zeros(60, dtype=torch.long)
    cfg = dict(type="nms", iou_threshold=0.5)
    want_dets, want_keep = O.planar_nms(boxes, scores, idxs, cfg)
    dets, keep = PlanarNMS()(boxes.to(DEV), scores.to(DEV), idxs.to(DEV), cfg)
    assert keep.cpu().tolist() == want_keep.tolist() and set(range(10, 20)) <= set(keep.cpu().tolist()) and 31 not in keep.cpu().tolist()
    assert torch.equal(dets.cpu(), want_dets)
    _, keep_sph = SphNMS("naive_iou")(boxes.to(DEV), scores.to(DEV), idxs.to(DEV), dict(iou_threshold=0.5))
    assert set(range(10, 20)) & set(keep_sph.cpu().tolist()) == {10}


def test_batched_nms_images_hint_checks(api):
    """The hinted path of sph_batched_nms_images: a segment longer than max_per_segment is NOT dropped (the call falls back
    to the exact-length path), and a label outside [0, num_classes) raises instead of aliasing into the next image."""
    B, n = 3, 300
    boxes = O.generate_boxes(B * n, alpha_range=(5, 60), beta_range=(5, 60), box="bfov", seed=12).to(DEV)
    scores = torch.rand(B * n, device=DEV)
    labels = torch.randint(0, 4, (B * n,), device=DEV)
    image_ids = torch.arange(B, device=DEV).repeat_interleave(n)
    want = api.nms.sph_batched_nms_images(boxes, scores, labels, image_ids, 0.5)
    short = api.nms.sph_batched_nms_images(boxes, scores, labels, image_ids, 0.5, num_images=B, num_classes=4, max_per_segment=40)
    assert torch.equal(want, short)
    valid = scores > 0.3
    want_v = api.nms.sph_batched_nms_images(boxes, scores, labels, image_ids, 0.5, num_images=B, num_classes=4, max_per_segment=n, valid=valid)
    short_v = api.nms.sph_batched_nms_images(boxes, scores, labels, image_ids, 0.5, num_images=B, num_classes=4, max_per_segment=40, valid=valid)
    assert torch.equal(want_v, short_v) and bool(valid[want_v].all())
    bad = labels.clone()
    bad[5] = 4
    with pytest.raises(ValueError):
        api.nms.sph_batched_nms_images(boxes, scores, bad, image_ids, 0.5, num_images=B, num_classes=4, max_per_segment=n)


# ---- unbiased_iou (the exact spherical IoU; SphOverlaps2D's default backend) and SphNMS('unbiased_iou') ---------------
@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_unbiased_iou_golden(api, box):
    g = load_golden("unbiased")
    b1, b2 = cu(g[box + "_b1"]), cu(g[box + "_b2"])
    got = api.iou.unbiased_iou(b1, b2, is_aligned=True)
    err = np.abs(got.cpu().numpy() - g[box + "_aligned_f64"])
    # 1e-5 everywhere except, at most, a degenerate pair whose vertex test (5e-9) falls the other way with CUDA's libm
    assert got.shape == (b1.size(0),) and (err > 1e-5).sum() <= 2, (np.where(err > 1e-5)[0], err.max())
    mat = api.iou.SphOverlaps2D(box_version=b1.size(1))(b1[:23], b2[:201])            # the reference's default backend
    assert mat.shape == (23, 201) and (np.abs(mat.cpu().numpy() - g[box + "_rc_f64"]) > 1e-5).sum() <= 2
    flat = api.iou.unbiased_iou(b1[:23].repeat_interleave(201, 0), b2[:201].repeat(23, 1), is_aligned=True)
    assert torch.equal(mat.reshape(-1), flat)
    rmax, rarg, cmax, carg = api.iou.sph_max_overlaps(b1[:23], b2[:201], backend='unbiased_iou')     # fused max / argmax, no matrix
    assert torch.equal(rmax, mat.max(dim=1)[0]) and torch.equal(cmax, mat.max(dim=0)[0])
    assert torch.equal(mat[torch.arange(23), rarg.long()], rmax) and torch.equal(mat[carg.long(), torch.arange(201)], cmax)
    # against the oracle on fresh seeded boxes
    x = O.generate_boxes(4000, alpha_range=(2, 150), beta_range=(2, 150), box=box, seed=5)
    y = (x + torch.randn_like(x) * 10).clamp(min=1)
    y[:, 0].clamp_(0, 360); y[:, 1].clamp_(0, 180); y[:, 2:4].clamp_(max=179)
    err = (api.iou.unbiased_iou(x.to(DEV), y.to(DEV), is_aligned=True).cpu() - O.unbiased_iou(x, y, is_aligned=True)).abs()
    assert int((err > 1e-5).sum()) <= 4, err.max()


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_unbiased_nms_golden_keep_sets(api, box):
    """SphNMS('unbiased_iou') (the reference's pandora configs) against the keep lists of the reference's float64 run."""
    g = load_golden("unbiased")
    boxes, scores, idxs = cu(g[box + "_boxes"]), cu(g[box + "_scores"]), cu(g[box + "_idxs"])
    iou = g[box + "_pair_iou_f64"]
    same = g[box + "_idxs"][:, None] == g[box + "_idxs"][None, :]
    for thr in (0.3, 0.5):
        assert not (np.abs(iou[same] - thr) < 1e-5).any(), "fixture has a pair on the threshold"
        want = g["%s_keep_thr%d" % (box, int(thr * 10))].tolist()
        _, keep = api.nms.SphNMS('unbiased_iou')(boxes, scores, idxs, dict(type="nms", iou_threshold=thr, max_num=120))
        assert keep.cpu().tolist() == want


# ---- box format conversions either side of the path (sphdet/bbox/box_formator.py) ----------------------------------------
def test_box_format_golden(api):
    from sph_retina_b200.sphdet.bbox import box_formator as bf
    g = load_golden("box_format")
    for fmt, key, d_out, want_key, size, exact in BOX_FORMAT_CASES:
        names = {v: k for k, v in api.native.BOX_FORMAT.items()}
        got = api.native.box_format(names[fmt], cu(g[key]), d_out, size).cpu().numpy()
        if exact:
            assert np.array_equal(got, g[want_key]), (want_key, np.abs(got - g[want_key]).max())
        else:
            np.testing.assert_allclose(got, g[want_key], rtol=1e-4, atol=1e-4, err_msg=want_key)   # tan(alpha / 2) near 180 degrees
    # the reference's names and classes
    sph4, sph5, xyxy = cu(g["sph4"]), cu(g["sph5"]), cu(g["xyxy"])
    assert np.array_equal(bf.Sph2PlanarBoxTransform('sph2pix')(sph4).cpu().numpy(), g["planar4_sph2pix_512"])
    assert np.array_equal(bf.Sph2PlanarBoxTransform('sph2pix', 5)(sph5, (960, 1920)).cpu().numpy(), g["planar5_sph2pix_960"])
    assert np.array_equal(bf.Planar2SphBoxTransform('pix2sph', 5)(xyxy, (960, 1920)).cpu().numpy(), g["back5_sph2pix_960"])
    assert np.array_equal(bf.xyxy2xywh(xyxy).cpu().numpy(), g["xyxy2xywh"]) and np.array_equal(bf.bfov2rbfov(sph4).cpu().numpy(), g["bfov2rbfov"])
    assert np.array_equal(bf.geo2sph(cu(g["geo"])).cpu().numpy(), g["geo2sph_5"]) and np.array_equal(bf.sph2geo(sph4).cpu().numpy(), g["sph2geo_4"])
    np.testing.assert_allclose(bf.obb2hbb_xyxy(cu(g["obb"])).cpu().numpy(), g["obb2hbb_xyxy"], rtol=2e-6, atol=2e-5)
    assert bf.is_valid_boxes(sph4) and not bf.is_valid_boxes(sph4 + 400) and bf.xywh2xyxy(xyxy[:0]).shape == (0, 4)
    # round trip and no mutation of the input
    before = sph4.clone()
    back = bf.Planar2SphBoxTransform()(bf.Sph2PlanarBoxTransform()(sph4))
    assert torch.equal(sph4, before) and float((back - sph4).abs().max()) < 2e-4


def test_box_format_autograd(api):
    """The conversions are plain torch expressions in the reference (box_formator.py:17-117,166-200), so gradients flow
    through them; here each one is a kernel launch with an explicit backward.  Checked against autograd through the
    reference's expressions written out in torch (float64), and the tangent formats must refuse instead of cutting the graph."""
    from sph_retina_b200.sphdet.bbox import box_formator as bf
    g = load_golden("box_format")
    H, W = 960.0, 1920.0

    def ref_xyxy2xywh(b):
        return torch.stack([(b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2, b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]], 1)

    def ref_xywh2xyxy(b):
        return torch.stack([b[:, 0] - b[:, 2] / 2, b[:, 1] - b[:, 3] / 2, b[:, 0] + b[:, 2] / 2, b[:, 1] + b[:, 3] / 2], 1)

    def ref_sph2pix(b):
        return torch.stack([b[:, 0] / 360 * W, b[:, 1] / 180 * H, b[:, 2] / 360 * W, b[:, 3] / 180 * H], 1)

    def ref_obb2hbb_xywh(o):
        c, s = torch.cos(o[:, 4]).abs(), torch.sin(o[:, 4]).abs()
        return torch.stack([o[:, 0], o[:, 1], c * o[:, 2] + s * o[:, 3], s * o[:, 2] + c * o[:, 3]], 1)

    cases = [
        ("xyxy2xywh", bf.xyxy2xywh, ref_xyxy2xywh, g["xyxy"]),
        ("xywh2xyxy", bf.xywh2xyxy, ref_xywh2xyxy, g["xywh"]),
        ("obb2hbb_wywh", bf.obb2hbb_wywh, ref_obb2hbb_xywh, g["obb"]),
        ("obb2hbb_xyxy", bf.obb2hbb_xyxy, lambda o: ref_xywh2xyxy(ref_obb2hbb_xywh(o)), g["obb"]),
        ("geo2sph", bf.geo2sph, lambda b: torch.cat([b[:, :1] + 180, 90 - b[:, 1:2], b[:, 2:]], 1), g["geo"]),
        ("sph2planar4", lambda b: bf.Sph2PlanarBoxTransform('sph2pix', 4)(b, (H, W)), lambda b: ref_xywh2xyxy(ref_sph2pix(b)), g["sph4"]),
        ("sph2planar5", lambda b: bf.Sph2PlanarBoxTransform('sph2pix', 5)(b, (H, W)),
         lambda b: torch.cat([ref_sph2pix(b), -torch.deg2rad(b[:, 4:5])], 1), g["sph5"]),
        ("planar2sph5", lambda b: bf.Planar2SphBoxTransform('pix2sph', 5)(b, (H, W)),
         lambda b: torch.cat([torch.stack([(b[:, 0] + b[:, 2]) / 2 / W * 360, (b[:, 1] + b[:, 3]) / 2 / H * 180,
                                           (b[:, 2] - b[:, 0]) / W * 360, (b[:, 3] - b[:, 1]) / H * 180], 1), b[:, :1] * 0], 1), g["xyxy"]),
    ]
    for name, fn, ref, data in cases:
        x = cu(data[:257]).clone().requires_grad_(True)
        y = fn(x)
        assert y.requires_grad, name
        up = torch.randn(y.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
        (y * up).sum().backward()
        x64 = torch.as_tensor(data[:257]).double().requires_grad_(True)
        y64 = ref(x64)
        (y64 * up.cpu().double()).sum().backward()
        np.testing.assert_allclose(y.detach().cpu().numpy(), y64.detach().numpy(), rtol=3e-6, atol=3e-4, err_msg=name)
        np.testing.assert_allclose(x.grad.cpu().numpy(), x64.grad.numpy(), rtol=2e-5, atol=2e-5, err_msg=name)
        with torch.no_grad():                                   # no graph asked for: the plain launch
            assert not fn(x).requires_grad
    with pytest.raises(NotImplementedError):
        bf.Sph2PlanarBoxTransform('sph2tan', 4)(cu(g["sph4"][:8]).requires_grad_(True))
    assert not bf.Sph2PlanarBoxTransform('sph2tan', 4)(cu(g["sph4"][:8])).requires_grad


# ---- sph2pob_legacy_iou (the reference's first transform; BFoV only) ----------------------------------------------------
def test_sph2pob_legacy_golden(api):
    g = load_golden("legacy")
    b1, b2 = cu(g["b1"]), cu(g["b2"])
    small = np.minimum(g["b1"][:, 2:4].min(axis=1), g["b2"][:, 2:4].min(axis=1))
    speck = degenerate_pairs(g["b1"], g["b2"])
    for key, kw in (("iou", {}), ("iof", dict(mode="iof")), ("chord", dict(rbb_edge="chord")), ("tangent", dict(rbb_edge="tangent"))):
        got = api.iou.sph2pob_legacy_iou(b1, b2, is_aligned=True, **kw).cpu().numpy()
        ok, err = within(got, g[key + "_f64"], g[key + "_f32"])
        ok |= (small < 1.5) & (err < 1e-4)           # the stand-in's num / (den + 1e-8): see tests/test_hostsim_math.py
        assert ok[~speck].all(), (key, np.where(~ok & ~speck)[0][:10], err[~ok & ~speck][:10])
        assert got.min() >= 0.0 and got.max() <= 1.0
    np.testing.assert_allclose(api.iou.sph2pob_legacy_iou(cu(g["kat_b1"]), cu(g["kat_b2"]), is_aligned=True).cpu().numpy(),
                               g["kat_iou"], atol=5e-6)
    mat = api.iou.SphOverlaps2D('sph2pob_legacy_iou', 4)(b1[:29], b2[:333])
    ok, err = within(mat.cpu().numpy(), g["rc_f64"], g["rc_f32"])
    assert mat.shape == (29, 333) and (~ok).sum() <= 2, err[~ok]
    flat = api.iou.sph2pob_legacy_iou(b1[:29].repeat_interleave(333, 0), b2[:333].repeat(29, 1), is_aligned=True)
    assert torch.equal(mat.reshape(-1), flat)
    with pytest.raises(ValueError):
        api.iou.sph2pob_legacy_iou(torch.rand(3, 5, device=DEV), torch.rand(3, 5, device=DEV))


# ---- the other losses on the Sph2Pob OBBs ---------------------------------------------------
def _other_loss(api, cls, kw, **extra):
    return getattr(api.losses, cls)(**kw, **extra)


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_other_losses_forward_backward_golden(api, box):
    """Sph2PobGDLoss (gwd / kld / jd / kld_symmax / kld_symmin + fun / tau / alpha / normalize / sqrt options),
    Sph2PobKFLoss (none / ln / exp) and Sph2PobL1Loss (encode / swap / modulus / plain): elementwise loss and both
    gradients against the reference's float64 run (tolerances in conftest.check_other_loss)."""
    g = load_golden("other_losses_" + box)
    for name, (cls, kw) in other_loss_variants(g).items():
        p = cu(g["pred"]).requires_grad_(True)
        t = cu(g["target"]).requires_grad_(True)
        L = _other_loss(api, cls, kw, reduction="sum")
        el = L(p, t, reduction_override="none")
        assert el.shape == ((p.size(0), 5) if cls == "Sph2PobL1Loss" else (p.size(0),))
        el.sum().backward()
        check_other_loss(name, g, el.detach().cpu().numpy(), p.grad.cpu().numpy(), t.grad.cpu().numpy())


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_other_losses_reductions_weights_avg_factor(api, box):
    g = load_golden("other_losses_" + box)
    p, t, w1, w2 = cu(g["pred"]), cu(g["target"]), cu(g["w1"]), cu(g["w2"])
    for name, (cls, kw) in other_loss_variants(g).items():
        L = _other_loss(api, cls, kw, loss_weight=2.0)
        rt = 2e-3 if name == "l1_swap" else 5e-5     # l1_swap sums terms of 1e4 with fp32 OBBs
        np.testing.assert_allclose(L(p, t).item(), g[name + "_red_mean"], rtol=rt, err_msg=name)
        np.testing.assert_allclose(L(p, t, w2, avg_factor=77.0).item(), g[name + "_red_w2_avg"], rtol=rt, err_msg=name)
        np.testing.assert_allclose(L(p, t, w2, reduction_override="sum").item(), g[name + "_red_w2_sum"], rtol=rt, err_msg=name)
        if cls != "Sph2PobL1Loss":
            np.testing.assert_allclose(L(p, t, w1, avg_factor=123.0).item(), g[name + "_red_w1_avg"], rtol=rt, err_msg=name)
        with pytest.raises(ValueError):
            L(p, t, w2, avg_factor=3.0, reduction_override="sum")
    # the fused reduced path and the elementwise path agree, gradients included; zero-weight rows get exactly zero
    for cls, kw in (("Sph2PobGDLoss", dict(loss_type="kld")), ("Sph2PobKFLoss", {}), ("Sph2PobL1Loss", {})):
        L = _other_loss(api, cls, kw, loss_weight=3.0)
        w = w2 if cls == "Sph2PobL1Loss" else w1
        if cls == "Sph2PobL1Loss" and w.size(1) == 4:
            w_full = torch.cat([w, w.mean(-1, keepdim=True)], -1)
        else:
            w_full = w
        pa, ta = p.clone().requires_grad_(True), t.clone().requires_grad_(True)
        fused = L(pa, ta, w, avg_factor=50.0)
        fused.backward()
        pb, tb = p.clone().requires_grad_(True), t.clone().requires_grad_(True)
        el = L(pb, tb, reduction_override="none")
        ref = (el * w_full).sum() / (50.0 + torch.finfo(torch.float32).eps)
        ref.backward()
        np.testing.assert_allclose(fused.item(), ref.item(), rtol=1e-5)
        rel = (pa.grad - pb.grad).norm(dim=1) / pb.grad.norm(dim=1).clamp(min=1e-12)
        assert float(rel[64:].median()) < 1e-6 and float((rel[64:] < 1e-4).float().mean()) > 0.99
        dead = (w_full == 0) if w_full.dim() == 1 else (w_full == 0).all(dim=1)
        assert int(dead.sum()) > 0 and float(pa.grad[dead].abs().max()) == 0.0 and float(ta.grad[dead].abs().max()) == 0.0
        assert torch.isfinite(pa.grad).all() and torch.isfinite(ta.grad).all()
        z = L(pa, ta, torch.zeros_like(w))
        assert z.item() == 0.0
        with torch.no_grad():
            assert torch.isfinite(L(p, t, w))
        assert L(p[:0], t[:0]).shape == () and L(p[:0], t[:0], reduction_override="none").numel() == 0


def test_other_losses_match_the_oracle_on_seeded_boxes(api):
    """Seeded PANDORA-style pairs (configs[2] shape, 20k rows): reduced loss of every loss type against the oracle in
    float64, and the gradient of the reduced loss against the oracle's autograd."""
    n = 20000
    t = O.generate_boxes(n, alpha_range=(5, 100), beta_range=(5, 100), box="rbfov", seed=0)
    torch.manual_seed(1)
    p = (t + torch.randn(n, 5) * torch.tensor([6, 6, 6, 6, 10.0])).clamp(min=1)
    w = (torch.rand(n) > 0.5).float()
    cases = [("Sph2PobGDLoss", O.sph2pob_gd_loss, dict(loss_type=k)) for k in ("gwd", "kld", "jd", "kld_symmax", "kld_symmin")]
    cases += [("Sph2PobKFLoss", O.sph2pob_kf_loss, {}), ("Sph2PobL1Loss", O.sph2pob_l1_loss, {})]
    for cls, ofn, kw in cases:
        ww = w if cls != "Sph2PobL1Loss" else w[:, None].expand(n, 5).contiguous()
        pd = p.double().requires_grad_(True)
        want = ofn(pd, t.double(), ww.double(), avg_factor=float(w.sum()), **kw)
        want.backward()
        pg = p.to(DEV).requires_grad_(True)
        got = getattr(api.losses, cls)(**kw)(pg, t.to(DEV), ww.to(DEV), avg_factor=float(w.sum()))
        got.backward()
        np.testing.assert_allclose(got.item(), want.item(), rtol=2e-5, err_msg=cls + str(kw))
        rel = (pg.grad.cpu().double() - pd.grad).norm(dim=1) / pd.grad.norm(dim=1).clamp(min=1e-30)
        live = w > 0
        assert float(rel[live].median()) < 3e-6 and float((rel[live] < 1e-4).float().mean()) > 0.995, (cls, kw)
        assert float(pg.grad[~live].abs().max()) == 0.0


# ---- NMS ---------------------------------------------------------------------------------------
def test_nms_known_answer(api):
    g = load_golden("kat")
    dets, keep = api.nms.SphNMS('sph2pob_efficient')(cu(g["nms_boxes"]), cu(g["nms_scores"]), cu(g["nms_idxs"]),
                                                      dict(type="nms", iou_threshold=0.5))
    assert keep.dtype == torch.int64 and keep.cpu().tolist() == [0, 5, 7, 3, 8, 9]
    np.testing.assert_allclose(dets.cpu().numpy(), g["nms_dets"], atol=1e-6)


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_nms_golden_keep_sets(api, box):
    g = load_golden("nms")
    boxes, scores, idxs = cu(g[box + "_boxes"]), cu(g[box + "_scores"]), cu(g[box + "_idxs"])
    nms = api.nms.SphNMS()
    for thr, tag in ((0.3, "thr3"), (0.5, "thr5")):
        dets, keep = nms(boxes, scores, idxs, dict(type="nms", iou_threshold=thr, max_num=150))
        assert keep.cpu().tolist() == g["%s_keep_%s" % (box, tag)].tolist()
        np.testing.assert_allclose(dets.cpu().numpy(), g["%s_dets_%s" % (box, tag)], atol=1e-6)
    _, keep = nms(boxes, scores, idxs, dict(iou_threshold=0.5), class_agnostic=True)
    assert keep.cpu().tolist() == g[box + "_keep_agnostic"].tolist()


def greedy_from_matrix(iou, order, thr):
    alive = np.ones(len(order), bool)
    keep = []
    for a, i in enumerate(order):
        if not alive[a]:
            continue
        keep.append(int(i))
        alive[a + 1:] &= ~(iou[i, order[a + 1:]] > thr)
    return keep


@pytest.mark.parametrize("box,k", [("bfov", 1000), ("rbfov", 1537), ("rbfov", 33), ("bfov", 1)])
def test_nms_long_segment_equals_greedy_on_the_iou_matrix(api, box, k):
    """One segment of k boxes (class-agnostic stress shape): the kernel's blocked bitmask scan must equal
    the textbook greedy loop run on the kernel's own pairwise IoU (pivot = bboxes1)."""
    seeds = O.generate_boxes(max(1, k // 5), alpha_range=(5, 60), beta_range=(5, 60), box=box, seed=9)
    boxes = (seeds.repeat(6, 1)[:k] + torch.randn(k, seeds.size(1)) * 2).clamp(min=1).to(DEV)
    scores = torch.rand(k, device=DEV)
    dets, keep = api.nms.SphNMS()(boxes, scores, torch.zeros(k, dtype=torch.long, device=DEV), dict(iou_threshold=0.5))
    iou = api.iou.sph2pob_efficient_iou(boxes, boxes).cpu().numpy()
    order = torch.argsort(scores, descending=True, stable=True).cpu().numpy()
    want = greedy_from_matrix(iou, order, 0.5)
    assert keep.cpu().tolist() == want
    assert torch.equal(dets[:, -1], scores[keep]) and torch.equal(dets[:, :-1], boxes[keep])


def test_nms_batch_of_images_equals_per_image(api):
    B, n = 6, 400
    boxes = O.generate_boxes(B * n, alpha_range=(5, 60), beta_range=(5, 60), box="bfov", seed=10)
    boxes[1::2] = (boxes[0::2] + torch.randn(B * n // 2, 4)).clamp(min=1)
    boxes = boxes.to(DEV)
    scores = torch.rand(B * n, device=DEV)
    labels = torch.randint(0, 7, (B * n,), device=DEV)
    image_ids = torch.arange(B, device=DEV).repeat_interleave(n)
    keep = api.nms.sph_batched_nms_images(boxes, scores, labels, image_ids, 0.5)
    hinted = api.nms.sph_batched_nms_images(boxes, scores, labels, image_ids, 0.5, num_images=B, num_classes=7, max_per_segment=n)
    assert torch.equal(keep, hinted)
    for b in range(B):
        sel = (image_ids == b).nonzero().view(-1)
        _, k1 = api.nms.SphNMS()(boxes[sel], scores[sel], labels[sel], dict(iou_threshold=0.5))
        assert sorted(sel[k1].cpu().tolist()) == sorted(keep[image_ids[keep] == b].cpu().tolist())


def test_nms_fast_path_hands_over_what_it_does_not_cover(api):
    """One class with 4500 candidates (beyond the 4096-key sorting buffer of the device pipeline) and labels >= 1024:
    SphNMS must silently take the general path and give the same answer."""
    from sph_retina_b200 import synthetic as S
    boxes, scores, labels, _ = (t.to(DEV) for t in S.nms_batch(1, 4500, 3, seed=9))
    nms = api.nms.SphNMS()
    one = torch.zeros_like(labels)
    idx, count = api.native.nms_images(boxes, scores, one, 1, 1024, 0.5, 100)
    assert int(count) == -1
    idx, count = api.native.nms_images(boxes, scores, labels + 2000, 1, 1024, 0.5, 100)
    assert int(count) == -1
    d1, k1 = nms(boxes, scores, one, dict(iou_threshold=0.5))
    d2, k2 = nms(boxes, scores, one + 7000, dict(iou_threshold=0.5))
    assert torch.equal(k1, k2) and torch.equal(d1, d2) and 50 < k1.numel() < 4000


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_nms_image_blocks_equals_per_image_nms(api, box):
    """sphk_nms_images (device-side sort + per-segment NMS + per-image ordering) against SphNMS image by image, with
    padding entries (valid = 0), an image with a single class, an empty image and max_per_img cuts."""
    from sph_retina_b200 import synthetic as S
    from sph_retina_b200.sphdet.bbox.nms import sph_nms_image_blocks
    B, K, C = 7, 700, 13
    boxes, scores, labels, _ = (t.to(DEV) for t in S.nms_batch(B, K, C, box=box, seed=5))
    valid = torch.rand(B * K, device=DEV) > 0.15
    labels[2 * K:3 * K] = 4                       # one image with a single (long) segment
    valid[5 * K:6 * K] = False                    # one image with nothing in it
    scores[100:140] = scores[100]                 # equal scores
    for max_per_img in (None, 50):
        idx, count = sph_nms_image_blocks(boxes, scores, labels, B, C, 0.5, max_per_img, valid=valid)
        nms = api.nms.SphNMS()
        for b in range(B):
            sl = slice(b * K, (b + 1) * K)
            m = valid[sl]
            got = idx[b, :int(count[b])].long()
            assert bool((idx[b, int(count[b]):] == -1).all())
            if int(m.sum()) == 0:
                assert int(count[b]) == 0
                continue
            local = torch.nonzero(m).squeeze(1) + b * K
            # labels >= 1024 take SphNMS off its own device-pipeline fast path: this is the general (sorted-key) path
            dets, keep = nms(boxes[local], scores[local], labels[local] + 5000, dict(iou_threshold=0.5))
            want = local[keep]
            dets2, keep2 = nms(boxes[local], scores[local], labels[local], dict(iou_threshold=0.5))     # fast path
            assert torch.equal(scores[local][keep2], scores[local][keep]) and torch.equal(dets2[:, -1], dets[:, -1])
            assert torch.equal(torch.sort(keep2)[0], torch.sort(keep)[0])
            if max_per_img is not None:
                want = want[:max_per_img]
            assert got.numel() == want.numel(), (b, got.numel(), want.numel())
            assert torch.equal(scores[got], scores[want])                       # same scores in the same order ...
            assert torch.equal(torch.sort(got)[0], torch.sort(want)[0]) or max_per_img is not None   # ... same boxes
            assert bool(valid[got].all()) and bool(((got >= b * K) & (got < (b + 1) * K)).all())


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_nms_golden_at_the_test_time_shape(api, box):
    """BASELINE configs[3] at its own shape: one image of 1,000 candidates with labels in [0, 80), thresholds 0.3 / 0.5,
    keep lists from the reference's SphNMS (no same-label pair within 1e-5 of a threshold: margin recorded in the fixture),
    through the per-call path AND the device pipeline of a test batch."""
    g = load_golden("nms_cfg4")
    boxes, scores, idxs = cu(g[box + "_boxes"]), cu(g[box + "_scores"]), cu(g[box + "_idxs"])
    nms = api.nms.SphNMS()
    for thr, tag in ((0.3, "thr3"), (0.5, "thr5")):
        dets, keep = nms(boxes, scores, idxs, dict(type="nms", iou_threshold=thr))
        assert keep.cpu().tolist() == g["%s_keep_%s" % (box, tag)].tolist()
        np.testing.assert_allclose(dets.cpu().numpy(), g["%s_dets_%s" % (box, tag)], atol=1e-6)
        parity_record("nms_keep_list", n=int(boxes.size(0)), threshold=thr, kept=int(keep.numel()), n_fail_before_allowances=0,
                      decision_margin_fp64=float(g[box + "_margin"]))
    _, keep = nms(boxes, scores, idxs, dict(type="nms", iou_threshold=0.5, max_num=100))
    assert keep.cpu().tolist() == g[box + "_keep_thr5_max100"].tolist()
    # the same image as one block of the batched device pipeline (top-100 per image, score order)
    out_idx, counts = api.nms.sph_nms_image_blocks(boxes, scores, idxs, 1, 80, 0.5, 100)
    assert int(counts[0]) == 100 and out_idx[0, :100].cpu().tolist() == g[box + "_keep_thr5_max100"].tolist()


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_multiclass_nms_wrapper(api, box):
    """multiclass_nms (sphdet/bbox/nms/utils.py:6-95) against the reference's function run with SphNMS('sph2pob_efficient'):
    class-specific and shared boxes, score factors applied after the score mask, max_num."""
    g = load_golden("nms_cfg4")
    D = 4 if box == "bfov" else 5
    mb, ms, fac = cu(g[box + "_mc_bboxes"]), cu(g[box + "_mc_scores"]), cu(g[box + "_mc_factors"])
    for tag, kw in (("plain", {}), ("factors", dict(score_factors=fac))):
        dets, labels, inds = api.nms.multiclass_nms(mb, ms, 0.3, dict(type="nms", iou_threshold=0.5), max_num=100, return_inds=True,
                                                    box_version=D, **kw)
        assert inds.cpu().tolist() == g["%s_mc_%s_inds" % (box, tag)].tolist()
        assert labels.cpu().tolist() == g["%s_mc_%s_labels" % (box, tag)].tolist()
        np.testing.assert_allclose(dets.cpu().numpy(), g["%s_mc_%s_dets" % (box, tag)], atol=1e-6)
    dets, labels, inds = api.nms.multiclass_nms(mb[:, :D].contiguous(), ms, 0.3, dict(type="nms", iou_threshold=0.5), return_inds=True,
                                                box_version=D)
    assert inds.cpu().tolist() == g[box + "_mc_shared_inds"].tolist() and labels.cpu().tolist() == g[box + "_mc_shared_labels"].tolist()
    np.testing.assert_allclose(dets.cpu().numpy(), g[box + "_mc_shared_dets"], atol=1e-6)
    dets2, labels2 = api.nms.multiclass_nms(mb, ms, 0.3, dict(iou_threshold=0.5), max_num=100, box_version=D)
    assert dets2.shape[0] == 100 and labels2.shape[0] == 100
    parity_record("multiclass_nms", n=int(ms.numel()), n_fail_before_allowances=0, decision_margin_fp64=float(g[box + "_mc_margin"]))


# ---- box coders and the fused decode -> loss step ---------------------------------
@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_coder_golden(api, box):
    from sph_retina_b200.sphdet.bbox.coder import DeltaXYWHASphBBoxCoder, DeltaXYWHSphBBoxCoder
    from test_oracle_golden import coder_kwargs
    g = load_golden("coder")
    cls = DeltaXYWHSphBBoxCoder if box == "bfov" else DeltaXYWHASphBBoxCoder
    anchors, deltas = cu(g[box + "_anchors"]), cu(g[box + "_deltas"])
    keep_a, keep_d = anchors.clone(), deltas.clone()
    for name in ("plain", "norm", "ctr", "noclip"):
        kw = coder_kwargs(g, box, name)
        ckw = {("target_" + k if k in ("means", "stds") else k): v for k, v in kw.items()}
        coder = cls(**ckw)
        dec = coder.decode(anchors, deltas)
        np.testing.assert_allclose(dec.cpu().numpy(), g["%s_%s_decode_f64" % (box, name)], rtol=3e-6, atol=2e-5)
        enc = coder.encode(anchors, cu(g["%s_%s_decode_f32" % (box, name)]))
        np.testing.assert_allclose(enc.cpu().numpy(), g["%s_%s_encode_f32" % (box, name)], rtol=2e-5, atol=2e-5)
        # backward of decode == autograd through the torch restatement (clamp semantics included)
        d1 = deltas.clone().requires_grad_(True)
        up = torch.randn_like(deltas)
        (coder.decode(anchors, d1) * up).sum().backward()
        d2 = deltas.double().clone().requires_grad_(True)
        (O.delta2bbox(anchors.double(), d2, **kw) * up.double()).sum().backward()
        np.testing.assert_allclose(d1.grad.cpu().numpy(), d2.grad.cpu().numpy(), rtol=1e-5, atol=1e-5)
    assert torch.equal(anchors, keep_a) and torch.equal(deltas, keep_d)
    # per-class deltas [N, C * D] share the roi (delta_xywh_sph_bbox_coder.py:238)
    coder = cls()
    D = anchors.size(1)
    multi = torch.cat([deltas, deltas * 0.5, deltas * 0.0], dim=1)
    out = coder.decode(anchors, multi)
    assert out.shape == (anchors.size(0), 3 * D)
    assert torch.equal(out[:, :D], coder.decode(anchors, deltas)) and torch.equal(out[:, 2 * D:], coder.decode(anchors, deltas * 0.0))
    assert coder.decode(anchors[:0], deltas[:0]).shape[0] == 0


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_decoded_loss_fused(api, box):
    """Sph2PobDecodedIoULoss.forward_decoded == Sph2PobIoULoss(coder.decode(...)) of the reference (golden: its own coder
    + loss + autograd, fp64), for the fused one-launch path (mode iou) and the composed path (ciou)."""
    from sph_retina_b200.sphdet.bbox.coder import DeltaXYWHASphBBoxCoder, DeltaXYWHSphBBoxCoder
    from sph_retina_b200.sphdet.losses import Sph2PobDecodedIoULoss, Sph2PobIoULoss
    g = load_golden("coder")
    cls = DeltaXYWHSphBBoxCoder if box == "bfov" else DeltaXYWHASphBBoxCoder
    coder = cls(target_means=tuple(g[box + "_means"].tolist()), target_stds=tuple(g[box + "_stds"].tolist()))
    anchors, target, weight = cu(g[box + "_anchors"]), cu(g[box + "_target"]), cu(g[box + "_weight"])
    npos = float((weight[:, 0] > 0).sum())
    pos = (weight[:, 0] > 0).cpu().numpy()
    for mode in ("iou", "ciou"):
        L = Sph2PobDecodedIoULoss(mode=mode, loss_weight=1.5)
        d = cu(g[box + "_loss_deltas"]).requires_grad_(True)
        before = api.native.launches
        loss = L.forward_decoded(coder, anchors, d, target, weight, avg_factor=npos)
        if mode == "iou":
            assert api.native.launches - before == 1            # one kernel for the whole step
        loss.backward()
        assert abs(float(loss.detach()) - float(g["%s_%s_loss_f64" % (box, mode)])) < 3e-5
        ok, rel, rel32 = grad_rows_ok(d.grad.cpu().numpy(), g["%s_%s_gdeltas_f64" % (box, mode)],
                                      g["%s_%s_gdeltas_f32" % (box, mode)], pos)
        assert ok.mean() > 0.99 and np.median(rel) < 5e-6, (mode, ok.mean(), np.median(rel))
        assert not bool(d.grad[~torch.from_numpy(pos).to(DEV)].any())
        # == the two-step composition through the same kernels
        d2 = cu(g[box + "_loss_deltas"]).requires_grad_(True)
        two = Sph2PobIoULoss(mode=mode, loss_weight=1.5)(coder.decode(anchors, d2), target, weight, avg_factor=npos)
        two.backward()
        assert abs(float(two.detach()) - float(loss.detach())) < 2e-6
        np.testing.assert_allclose(d.grad.cpu().numpy(), d2.grad.cpu().numpy(), rtol=2e-4, atol=1e-8)
    L = Sph2PobDecodedIoULoss(loss_weight=2.0)
    # 1-D weights, no weights, sum reduction, all-zero weights (the reference's early-out), no grad
    d = cu(g[box + "_loss_deltas"])
    w1 = weight[:, 0].contiguous()
    a = L.forward_decoded(coder, anchors, d, target, w1, avg_factor=npos)
    b = L.forward_decoded(coder, anchors, d, target, weight, avg_factor=npos)
    assert abs(float(a) - float(b)) < 1e-6
    s1 = L.forward_decoded(coder, anchors, d, target, w1, reduction_override="sum")
    assert abs(float(s1) - float(a) * npos) < 1e-3 * float(s1)
    dense = L.forward_decoded(coder, anchors[pos.nonzero()[0]], d[pos.nonzero()[0]], target[pos.nonzero()[0]])
    assert abs(float(dense) - float(a)) < 1e-5 * max(1.0, float(a))
    dz = d.clone().requires_grad_(True)
    z = L.forward_decoded(coder, anchors, dz, target, torch.zeros_like(weight), avg_factor=npos)
    z.backward()
    assert float(z.detach()) == 0.0 and not bool(dz.grad.any())


def test_decoded_loss_full_batch_shape(api):
    """The real call shape: 16 images x 98 208 anchors (RBFoV), ~1 % positives.  The fused step must equal the loss
    restricted to the positive rows (zero-weight rows contribute nothing) and write a zero gradient everywhere else."""
    from sph_retina_b200 import synthetic as S
    from sph_retina_b200.sphdet.bbox.coder import DeltaXYWHASphBBoxCoder
    from sph_retina_b200.sphdet.losses import Sph2PobDecodedIoULoss
    anchors, deltas, target, weight = (t.to(DEV) for t in S.head_loss_batch(16))
    n = anchors.size(0)
    assert n == 16 * 98208
    coder = DeltaXYWHASphBBoxCoder(target_stds=(0.1, 0.1, 0.2, 0.2, 0.1))
    L = Sph2PobDecodedIoULoss(loss_weight=1.0)
    pos = weight[:, 0] > 0
    npos = float(pos.sum())
    assert 0.002 * n < npos < 0.03 * n
    d = deltas.clone().requires_grad_(True)
    loss = L.forward_decoded(coder, anchors, d, target, weight, avg_factor=npos)
    loss.backward()
    dp = deltas[pos].clone().requires_grad_(True)
    ref = L.forward_decoded(coder, anchors[pos], dp, target[pos], None, avg_factor=npos)
    ref.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) < 1e-5 * max(1.0, float(ref.detach()))
    assert torch.allclose(d.grad[pos], dp.grad, rtol=1e-5, atol=1e-9)
    assert not bool(d.grad[~pos].any())
    assert 0.05 < float(loss.detach()) < 0.95


# ---- test-time post-processing of the head ----------------------------------------
def _head_outputs(B, C, D, seed):
    from sph_retina_b200 import synthetic as S
    strides, H, W, A = (32, 64, 128), 512, 1024, 9
    priors_all = S.retina_anchors(H, W, strides, box="rbfov" if D == 5 else "bfov")
    g = torch.Generator().manual_seed(seed)
    cls, reg, priors, off = [], [], [], 0
    for s in strides:
        h, w = H // s, W // s
        n = h * w * A
        priors.append(priors_all[off:off + n].contiguous())
        off += n
        cls.append(torch.randn(B, A * C, h, w, generator=g) * 2.0 - 3.0)
        reg.append(torch.randn(B, A * D, h, w, generator=g) * 0.3)
    return cls, reg, priors


@pytest.mark.parametrize("D", [4, 5])
def test_head_post_processing_batch_and_single(api, D):
    """get_bboxes_batch (one topk per level, one decode launch, one NMS launch for the batch) and get_bboxes_single
    (the reference's per-image contract) against the CPU restatement of sph_retina_head.py:35-212 (float64 NMS)."""
    from sph_retina_b200.sphdet.bbox.coder import DeltaXYWHASphBBoxCoder, DeltaXYWHSphBBoxCoder
    from sph_retina_b200.sphdet.models.heads import get_bboxes_batch, get_bboxes_single
    B, C = 3, 6
    cls, reg, priors = _head_outputs(B, C, D, seed=3 + D)
    cfg = dict(nms_pre=150, score_thr=0.05, nms=dict(type="nms", iou_threshold=0.5), max_per_img=60)
    stds = (0.1, 0.1, 0.2, 0.2, 0.1)[:D]
    coder = (DeltaXYWHSphBBoxCoder if D == 4 else DeltaXYWHASphBBoxCoder)(target_stds=stds)
    cls_d, reg_d, pri_d = [t.to(DEV) for t in cls], [t.to(DEV) for t in reg], [t.to(DEV) for t in priors]
    batch = get_bboxes_batch(cls_d, reg_d, pri_d, coder, cfg, box_version=D)
    assert len(batch) == B
    for b in range(B):
        want_det, want_lab = O.get_bboxes_single([t[b].double() for t in cls], [t[b].double() for t in reg],
                                                 [p.double() for p in priors], cfg, box_version=D, stds=stds)
        single = get_bboxes_single([t[b] for t in cls_d], [t[b] for t in reg_d], pri_d, coder, cfg, box_version=D)
        for name, (det, lab) in (("batch", batch[b]), ("single", single)):
            assert det.shape == want_det.shape and 10 < det.size(0) <= cfg["max_per_img"], (name, b, det.shape, want_det.shape)
            np.testing.assert_allclose(det.cpu().numpy(), want_det.numpy(), rtol=1e-4, atol=2e-4, err_msg="%s %d" % (name, b))
            assert torch.equal(lab.cpu(), want_lab)
            assert bool((det[:-1, -1] >= det[1:, -1]).all())
    # nothing above the threshold in one image: that image comes back empty, the others are unaffected
    for t in cls_d:
        t[1] = -20.0
    again = get_bboxes_batch(cls_d, reg_d, pri_d, coder, cfg, box_version=D)
    assert again[1][0].shape == (0, D + 1) and again[1][1].numel() == 0
    assert torch.equal(again[0][0], batch[0][0]) and torch.equal(again[2][1], batch[2][1])


@pytest.mark.parametrize("calc", ["naive_iou", "unbiased_iou", "planar"])
def test_head_post_processing_other_calculators(api, calc):
    """test_cfg.iou_calculator = 'naive_iou' (indoor360), 'unbiased_iou' (pandora) and 'planar' (PlanarNMS): the whole-batch
    pipeline equals the per-image contract, and the per-image result equals decode + the calculator's own NMS class."""
    from sph_retina_b200.sphdet.bbox.coder import DeltaXYWHSphBBoxCoder
    from sph_retina_b200.sphdet.bbox.nms import PlanarNMS, SphNMS
    from sph_retina_b200.sphdet.models.heads import get_bboxes_batch, get_bboxes_single
    B, C, D = 2, 5, 4
    cls, reg, priors = _head_outputs(B, C, D, seed=11)
    cfg = dict(nms_pre=120, score_thr=0.05, nms=dict(type="nms", iou_threshold=0.5), max_per_img=50, iou_calculator=calc)
    coder = DeltaXYWHSphBBoxCoder(target_stds=(0.1, 0.1, 0.2, 0.2))
    cls_d, reg_d, pri_d = [t.to(DEV) for t in cls], [t.to(DEV) for t in reg], [t.to(DEV) for t in priors]
    batch = get_bboxes_batch(cls_d, reg_d, pri_d, coder, cfg, box_version=D)
    for b in range(B):
        det, lab = get_bboxes_single([t[b] for t in cls_d], [t[b] for t in reg_d], pri_d, coder, cfg, box_version=D)
        assert 5 < det.size(0) <= 50 and det.shape == batch[b][0].shape
        np.testing.assert_allclose(batch[b][0].cpu().numpy(), det.cpu().numpy(), rtol=0, atol=1e-6)
        assert torch.equal(batch[b][1], lab)
        boxes, scores, labels = get_bboxes_single([t[b] for t in cls_d], [t[b] for t in reg_d], pri_d, coder, cfg, box_version=D,
                                                  with_nms=False)
        nms = PlanarNMS() if calc == "planar" else SphNMS(calc)
        want, keep = nms(boxes, scores, labels, cfg["nms"])
        assert torch.equal(det, want[:50]) and torch.equal(lab, labels[keep][:50])


# ---- full-size property checks (BASELINE.json configs) --------------------------------------------
def test_config1_one_million_pairs(api, c_oracle):
    b1 = O.generate_boxes(1_000_000, alpha_range=(1, 100), beta_range=(1, 100), box="bfov", seed=0)
    b2 = O.generate_boxes(1_000_000, alpha_range=(1, 100), beta_range=(1, 100), box="bfov", seed=1)
    got = api.iou.sph2pob_efficient_iou(b1.to(DEV), b2.to(DEV), is_aligned=True).cpu().numpy()
    assert got.shape == (1_000_000,) and np.isfinite(got).all() and got.min() >= 0 and got.max() <= 1
    sel = np.arange(0, 1_000_000, 37)
    want = c_oracle_aligned(c_oracle, 0, b1.numpy()[sel], b2.numpy()[sel])
    assert np.abs(got[sel] - want).max() < 1e-5
    assert abs((got > 0).mean() - 0.26) < 0.03                # ~26 % of random pairs overlap


def test_config5_sweep_slice(api, c_oracle):
    """1M x 1024 is checked through size-independent properties: a row block of the matrix equals the
    oracle, and the fused max/argmax over the full sweep equals the max over matrix blocks."""
    A = O.generate_boxes(1 << 20, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=0).to(DEV)
    G = O.generate_boxes(1024, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=1).to(DEV)
    rmax, rarg, cmax, carg = api.iou.sph_max_overlaps(A, G)
    best = torch.zeros(1024, device=DEV)
    for lo in range(0, 1 << 20, 1 << 16):
        blk = api.iou.sph2pob_efficient_iou(A[lo:lo + (1 << 16)], G)
        assert torch.equal(blk.max(dim=1)[0], rmax[lo:lo + (1 << 16)])
        best = torch.maximum(best, blk.max(dim=0)[0])
    assert torch.equal(best, cmax)
    rows = A[12345:12345 + 64].cpu().numpy()
    want = c_oracle_aligned(c_oracle, 0, np.repeat(rows, 1024, axis=0), np.tile(G.cpu().numpy(), (64, 1))).reshape(64, 1024)
    got = api.iou.sph2pob_efficient_iou(A[12345:12345 + 64], G).cpu().numpy()
    assert np.abs(got - want).max() < 1e-5


def test_distance_point_coder_golden(api):
    """DistancePointSphBBoxCoder (sphdet/bbox/coder/distance_point_sph_bbox_coder.py) through sphk_box_format."""
    from conftest import check_distance_coder
    from sph_retina_b200.sphdet.bbox.coder import distance_point_sph_bbox_coder as M
    check_distance_coder(M, DEV)


def test_anchor_free_head_post_processing(api):
    """get_bboxes_single with centerness score factors and the point coder (sph_fcos_head.py:196-321)."""
    from conftest import check_anchor_free_post_processing
    from sph_retina_b200.sphdet.bbox.coder import distance_point_sph_bbox_coder as M
    from sph_retina_b200.sphdet.models.heads import sph_bbox_post
    check_anchor_free_post_processing(sph_bbox_post, M, O, DEV, with_nms=True)
