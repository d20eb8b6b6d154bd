"""The DEVICE arithmetic (sph_retina_b200/csrc/sphk_math.cuh, sphk_grad.cuh) compiled for the host
by g++ (tests/hostsim) against the golden vectors of the reference.  Lets the no-GPU suite catch
maths regressions; the GPU suite (tests/test_gpu_parity.py) repeats the checks through the C ABI.

Parity criterion: |kernel - reference fp64| <= 1e-5, or the kernel is at least as
close to the fp64 run as the reference's own fp32 run on that element."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from conftest import (BOX_FORMAT_CASES, allow_degenerate, check_other_loss, degenerate_pairs, grad_rows_ok, load_golden, other_loss_kernel_args,
                      other_loss_variants, within)

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sph_oracle as O  # noqa: E402  (the checker the device maths is compared with where no golden vector exists)

fp = ctypes.POINTER(ctypes.c_float)


def hs_aligned(lib, kind, b1, b2, mode=0, edge=0):
    b1, b2 = np.ascontiguousarray(b1, np.float32), np.ascontiguousarray(b2, np.float32)
    P, D = b1.shape
    out = np.empty(P, np.float32)
    lib.hostsim_iou_aligned(kind, b1.ctypes.data_as(fp), b2.ctypes.data_as(fp), ctypes.c_long(P), D, mode, edge,
                            out.ctypes.data_as(fp))
    return out


def hs_fast(lib, fn, kind, b1, b2, mode=0, edge=0):
    """csrc/sphk_fast.cuh on the host: fn = hostsim_iou_aligned_v2 (N x M formulation: precompute, prefilter, fast
    path) or hostsim_iou_aligned_v3 (aligned formulation: job + clip stages).  Returns (iou, path) with
    path 0 = culled as disjoint, 1 = fast path, 2 = reference-order path."""
    b1, b2 = np.ascontiguousarray(b1, np.float32), np.ascontiguousarray(b2, np.float32)
    P, D = b1.shape
    out, path = np.empty(P, np.float32), np.empty(P, np.uint8)
    getattr(lib, fn)(kind, b1.ctypes.data_as(fp), b2.ctypes.data_as(fp), ctypes.c_long(P), D, mode, edge,
                     out.ctypes.data_as(fp), path.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)))
    return out, path


def hs_loss(lib, pred, target, grad_iou):
    p, t = np.ascontiguousarray(pred, np.float32), np.ascontiguousarray(target, np.float32)
    n, D = p.shape
    gi = np.ascontiguousarray(grad_iou, np.float32)
    iou, g1, g2 = np.empty(n, np.float32), np.empty((n, D), np.float32), np.empty((n, D), np.float32)
    lib.hostsim_loss_fwd_bwd(p.ctypes.data_as(fp), t.ctypes.data_as(fp), gi.ctypes.data_as(fp), ctypes.c_long(n), D,
                             iou.ctypes.data_as(fp), g1.ctypes.data_as(fp), g2.ctypes.data_as(fp))
    return iou, g1, g2


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
@pytest.mark.parametrize("kind,tr", [(0, "efficient"), (1, "standard")])
def test_aligned_iou(hostsim, box, kind, tr):
    g = load_golden("aligned_" + box)
    for key, mode, edge in (("iou", 0, 0), ("iof", 1, 0), ("chord", 0, 1), ("tangent", 0, 2)):
        got = hs_aligned(hostsim, kind, g["b1"], g["b2"], mode, edge)
        ok, err = within(got, g["%s_%s_f64" % (tr, key)], g["%s_%s_f32" % (tr, key)])
        ok = allow_degenerate(ok, err, g["b1"], g["b2"])
        assert ok.all(), (box, tr, key, np.where(~ok)[0], err[~ok])
        assert (err > 1e-5).sum() <= 2 and np.median(err) < 2e-7
        assert got.min() >= 0.0 and got.max() <= 1.0


def test_sph_fov(hostsim):
    g = load_golden("aligned_bfov")
    for kind, k in ((2, "sph"), (3, "fov")):
        got = hs_aligned(hostsim, kind, g["b1"], g["b2"])
        assert np.abs(got - g[k + "_f64"]).max() < 2e-6


def test_approx_identity_path_is_bit_identical(hostsim):
    """approx_iou_pair skips the hi + lo jitter bookkeeping where jitter_1 is the identity; it must return the bits of
    the general form on every pair: random boxes, boxes across the seam, and boxes on / next to the clamp zones and
    within eps of each other (where the shortcut must NOT be taken: path 0)."""
    rng = np.random.default_rng(7)
    n = 400_000
    def boxes(n):
        return np.stack([rng.uniform(0, 360, n), rng.uniform(0, 180, n), rng.uniform(1, 100, n), rng.uniform(1, 100, n)], 1).astype(np.float32)
    b1, b2 = boxes(n), boxes(n)
    # wrap cases: thetas on opposite sides of the seam
    b1[:50_000, 0] = rng.uniform(0, 30, 50_000); b2[:50_000, 0] = rng.uniform(330, 360, 50_000)
    b1[50_000:100_000, 0] = rng.uniform(330, 360, 50_000); b2[50_000:100_000, 0] = rng.uniform(0, 30, 50_000)
    # clamp zones and range ends of every column, both roles
    eps = np.float32(1e-4 * 1.2345678)
    ends = {0: 360.0, 1: 180.0, 2: 180.0, 3: 180.0}
    k = 100_000
    for col, end in ends.items():
        for b in (b1, b2):
            idx = rng.integers(k, 2 * k, 3000)
            b[idx, col] = rng.choice([0.0, 1e-5, eps, 2 * eps, 3 * eps, np.nextafter(eps, np.float32(1)), np.nextafter(2 * eps, np.float32(0)),
                                      end, end - eps, end - 2 * eps, end - 3 * eps, np.nextafter(np.float32(end), np.float32(0)), -0.0, -1.0, end + 1],
                                     3000).astype(np.float32)
    # near-equal columns (the similarity mask): copies, copies moved by less / more than eps, integer-valued boxes
    j = slice(2 * k, 2 * k + 60_000)
    b2[j] = b1[j]
    b2[2 * k + 20_000:2 * k + 40_000] += rng.uniform(-2, 2, (20_000, 4)).astype(np.float32) * eps
    b1[2 * k + 40_000:2 * k + 60_000] = np.round(b1[2 * k + 40_000:2 * k + 60_000])
    b2[2 * k + 40_000:2 * k + 60_000] = np.round(b1[2 * k + 40_000:2 * k + 60_000] + rng.integers(-3, 4, (20_000, 4)))
    # non-finite columns pass through both forms alike
    b1[-20:, 1] = np.nan; b2[-40:-20, 2] = np.inf
    u8 = ctypes.POINTER(ctypes.c_ubyte)
    for kind in (2, 3):
        got = hs_aligned(hostsim, kind, b1, b2)
        ref, path = np.empty(n, np.float32), np.empty(n, np.uint8)
        hostsim.hostsim_approx_general(kind, b1.ctypes.data_as(fp), b2.ctypes.data_as(fp), ctypes.c_long(n), ref.ctypes.data_as(fp),
                                       path.ctypes.data_as(u8))
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), int((got.view(np.uint32) != ref.view(np.uint32)).sum())
        assert path[:k].mean() > 0.99 and 0 < path[k:].mean() < 1          # both branches exercised
        assert path[2 * k:2 * k + 20_000].sum() == 0                        # identical boxes never take the shortcut


def test_known_answers(hostsim):
    g = load_golden("kat")
    for kind, name in ((0, "sph2pob_efficient_iou"), (1, "sph2pob_standard_iou"), (2, "sph_iou"), (3, "fov_iou")):
        np.testing.assert_allclose(hs_aligned(hostsim, kind, g["b1"], g["b2"]), g[name], atol=5e-6)


def test_pairwise_orientation(hostsim):
    """IoU(b1, b2) != IoU(b2, b1) at the 1e-4 level (role-asymmetric jitters): both orders are pinned."""
    g = load_golden("pairwise")
    for box in ("bfov", "rbfov"):
        rows, cols = g[box + "_rows"], g[box + "_cols"]
        R, C = len(rows), len(cols)
        b1, b2 = np.repeat(rows, C, axis=0), np.tile(cols, (R, 1))
        ok, err = within(hs_aligned(hostsim, 0, b1, b2).reshape(R, C), g[box + "_rc_f64"], g[box + "_rc_f32"])
        assert ok.all(), err[~ok]
        ok, err = within(hs_aligned(hostsim, 0, b2, b1).reshape(R, C).T, g[box + "_cr_f64"], g[box + "_cr_f32"])
        assert ok.all(), err[~ok]


def grad_row_error(got, truth):
    den = np.linalg.norm(truth, axis=1)
    return np.linalg.norm(got - truth, axis=1) / np.maximum(den, 1e-12), den


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_loss_gradients(hostsim, box):
    """Gradients: 1e-4 relative (row-wise L2) against the reference's fp64 autograd, or no worse than
    the reference's own fp32 autograd on that row (4-6 % of its rows are beyond 1e-4)."""
    g = load_golden("loss_" + box)
    n = len(g["pred"])
    iou, g1, g2 = hs_loss(hostsim, g["pred"], g["target"], -np.ones(n, np.float32))   # loss = 1 - iou
    ok, err = within(1 - iou, g["iou_loss_f64"], g["iou_loss_f32"])
    assert ok.all(), err[~ok]
    for got, key in ((g1, "gpred"), (g2, "gtarget")):
        truth, ref32 = g["iou_%s_f64" % key], g["iou_%s_f32" % key]
        rel, den = grad_row_error(got, truth)
        rel32, _ = grad_row_error(ref32, truth)
        live = den > 1e-9
        assert np.abs(got[~live]).max(initial=0.0) < 1e-6          # zero rows stay zero
        good = (rel <= 1e-4) | (rel <= rel32)
        assert good[live].mean() > 0.995, (key, (~good & live).sum())
        assert np.median(rel[live]) < 2e-6
        # the kernel is (much) closer to the fp64 truth than the fp32 reference is
        assert (rel[live] > 1e-4).sum() < 0.5 * (rel32[live] > 1e-4).sum()


@pytest.mark.parametrize("fn", ["hostsim_iou_aligned_v2", "hostsim_iou_aligned_v3"])
@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_fast_formulations_golden(hostsim, fn, box):
    """The precompute / prefilter / fast-path formulation used by the N x M and aligned kernels."""
    g = load_golden("aligned_" + box)
    for kind, tr in ((0, "efficient"), (1, "standard")):
        for key, mode, edge in (("iou", 0, 0), ("iof", 1, 0), ("chord", 0, 1), ("tangent", 0, 2)):
            got, path = hs_fast(hostsim, fn, kind, g["b1"], g["b2"], mode, edge)
            truth = g["%s_%s_f64" % (tr, key)]
            ok, err = within(got, truth, g["%s_%s_f32" % (tr, key)])
            ok = allow_degenerate(ok, err, g["b1"], g["b2"])
            assert ok.all(), (fn, box, tr, key, np.where(~ok)[0], err[~ok])
            assert not (((path == 0) | (path == 4)) & (truth > 0)).any()    # a culled pair is exactly 0 in the reference
            assert (path == 1).sum() > 1000                       # the fast path is what is being tested


@pytest.mark.parametrize("fn", ["hostsim_iou_aligned_v2", "hostsim_iou_aligned_v3"])
def test_fast_formulations_equal_reference_order_path(hostsim, fn):
    """Fast path vs the reference-order path of the same header on random, thin and huge boxes."""
    rng = np.random.RandomState(3)
    n = 200_000
    def boxes():
        b = np.stack([rng.uniform(0, 360, n), rng.uniform(0, 180, n), rng.uniform(0.5, 179, n), rng.uniform(0.5, 179, n),
                      rng.uniform(-90, 90, n)], axis=1).astype(np.float32)
        b[::3, 3] = rng.uniform(0.3, 3, len(b[::3]))         # thin boxes
        b[::7, 2] = rng.uniform(170, 400, len(b[::7]))       # oversize (clamped to 180 - eps)
        b[::11, 0] = 0.0                                     # on the seam (clamped to eps / 2 eps)
        b[::13, 1] = 0.0
        return b
    b1, b2 = boxes(), boxes()
    for kind in (0, 1):
        want = hs_aligned(hostsim, kind, b1, b2)
        got, path = hs_fast(hostsim, fn, kind, b1, b2)
        assert np.abs(got - want).max() < 5e-6
        assert not (((path == 0) | (path == 4)) & (want > 0)).any()
        assert (path == 1).mean() > 0.25


def quirk_zone_pairs(n, seed):
    """Overlapping pairs that sit inside the zones where a reference quirk is active WITHOUT jitter_1's similarity
    mask: sizes within jitter_2's eps (rad), boxes nearly parallel (|a1 - a2| < eps, eps'), internal angles inside
    the clamped-acos zone (centres on one meridian / one latitude circle near the equator), and mixtures."""
    rng = np.random.RandomState(seed)
    t1, p1 = rng.uniform(5, 355, n), rng.uniform(20, 160, n)
    a1, b1 = rng.uniform(5, 90, n), rng.uniform(5, 90, n)
    a2, b2 = rng.uniform(5, 90, n), rng.uniform(5, 90, n)
    off = rng.uniform(0.15, 0.6, n) * np.minimum(np.hypot(a1, b1), np.hypot(a2, b2))
    brg = rng.uniform(0, 2 * np.pi, n)
    g1, g2 = rng.uniform(-90, 90, n), rng.uniform(-90, 90, n)
    k = np.arange(n) % 6
    sgn = rng.choice([-1.0, 1.0], n)
    small = sgn * np.exp(rng.uniform(np.log(2e-4), np.log(6e-3), n))          # deg: > jitter_1's eps, < jitter_2's eps (rad)
    a2 = np.where(k == 0, a1 + small, a2)                                     # widths within eps
    b2 = np.where(k == 1, b1 + small, b2)                                     # heights within eps
    brg = np.where(k == 2, np.round(brg / np.pi) * np.pi + small * 0.05, brg)      # same meridian: |sin a| tiny
    p1 = np.where(k == 3, 90 + small, p1)                                     # both on the equator ...
    brg = np.where(k == 3, np.pi / 2 + small * 0.01, brg)                     # ... east-west: angle at the other clamp end
    dp, dt = -off * np.cos(brg), off * np.sin(brg) / np.maximum(np.sin(np.radians(p1)), 0.2)
    t2, p2 = t1 + dt, np.clip(p1 + dp, 1, 179)
    # nearly parallel planar boxes: gamma_2 chosen so that (a_g - gamma_1) - (a_p - gamma_2) = delta, |delta| up to 3e-3 rad
    # (the internal angles from the tangent-plane bearings, sphk_math.cuh).  Where gamma_2 has to be wrapped into
    # (-179, 179) the unwrapped angles of 'efficient' end up 2 pi apart while 'standard' still sees them parallel.
    tg, pg, tp, pp = (np.radians(v.astype(np.float32).astype(np.float64)) for v in (t1, p1, t2 % 360.0, p2))
    hth = np.sin(0.5 * (tp - tg)) ** 2
    ag = np.arctan2(np.sin(pp - pg) - 2 * np.cos(pg) * np.sin(pp) * hth, -np.sin(pp) * np.sin(tp - tg))
    ap = np.arctan2(np.sin(pp - pg) + 2 * np.cos(pp) * np.sin(pg) * hth, -np.sin(pg) * np.sin(tp - tg))
    delta = sgn * np.exp(rng.uniform(np.log(1e-5), np.log(3e-3), n))
    want = g1 - np.degrees(ag - ap) + np.degrees(delta)
    want = (want + 180.0) % 360.0 - 180.0
    g2 = np.where((k >= 4) & (np.abs(want) < 178.0), want, g2)
    B1 = np.stack([t1, p1, a1, b1, g1], 1).astype(np.float32)
    B2 = np.stack([t2 % 360.0, p2, a2, b2, g2], 1).astype(np.float32)
    return B1, B2


def near_touching_pairs(n, seed):
    """Pairs whose centre distance sits within a few per cent of the sum of the circumradii (where a cull test
    decides), for small, medium and oversize boxes, all over the sphere incl. poles and the seam."""
    rng = np.random.RandomState(seed)
    size = np.exp(rng.uniform(np.log(0.02), np.log(150.0), (n, 4)))            # alpha, beta of both boxes (deg)
    size[::9] = rng.uniform(170, 400, (len(size[::9]), 4))
    t1, p1 = rng.uniform(0, 360, n), np.degrees(np.arccos(rng.uniform(-1, 1, n)))
    p1[::17] = rng.uniform(0, 0.5, len(p1[::17])); p1[5::17] = 180 - rng.uniform(0, 0.5, len(p1[5::17]))
    r = 0.5 * (np.hypot(size[:, 0], size[:, 1]) + np.hypot(size[:, 2], size[:, 3]))
    dist = np.radians(np.minimum(r * rng.uniform(0.9, 1.1, n), 179.9))          # great-circle distance to box 2
    brg = rng.uniform(0, 2 * np.pi, n)
    lat1 = np.radians(90 - p1)
    lat2 = np.arcsin(np.clip(np.sin(lat1) * np.cos(dist) + np.cos(lat1) * np.sin(dist) * np.cos(brg), -1, 1))
    dlon = np.arctan2(np.sin(brg) * np.sin(dist) * np.cos(lat1), np.cos(dist) - np.sin(lat1) * np.sin(lat2))
    t2, p2 = (t1 + np.degrees(dlon)) % 360.0, 90 - np.degrees(lat2)
    g = rng.uniform(-90, 90, (n, 2))
    b1 = np.stack([t1, p1, size[:, 0], size[:, 1], g[:, 0]], 1).astype(np.float32)
    b2 = np.stack([t2, p2, size[:, 2], size[:, 3], g[:, 1]], 1).astype(np.float32)
    return b1, b2


def near_box_edge_pairs(n, seed):
    """Pairs whose second centre sits within a few per cent of the border of box 1 grown by box 2's circumradius,
    measured along box 1's own axes (where the box-frame prefilter decides): thin, square and oversize boxes, every
    gamma incl. the +-179..180 zone, poles and the seam, integer-valued copies (jitter_1's similarity mask)."""
    rng = np.random.RandomState(seed)
    size = np.exp(rng.uniform(np.log(1.0), np.log(120.0), (n, 4)))         # (the test's margin is 0.48 deg: tiny boxes never fire)
    size[::9] = rng.uniform(100, 400, (len(size[::9]), 4))
    size[4::9] = np.exp(rng.uniform(np.log(0.02), np.log(1.0), (len(size[4::9]), 4)))
    t1, p1 = rng.uniform(0, 360, n), np.degrees(np.arccos(rng.uniform(-1, 1, n)))
    p1[::17] = rng.uniform(0, 0.5, len(p1[::17])); p1[5::17] = 180 - rng.uniform(0, 0.5, len(p1[5::17]))
    g = rng.uniform(-90, 90, (n, 2))
    g[::23] = rng.choice([-1.0, 1.0], (len(g[::23]), 2)) * rng.uniform(178.5, 180, (len(g[::23]), 2))
    r2 = 0.5 * np.hypot(size[:, 2], size[:, 3])
    # target offset of centre 2 in box 1's frame (deg): on the grown border along x or y, +-10 %, anywhere along the other
    along_x = rng.rand(n) < 0.5
    border = np.where(along_x, 0.5 * size[:, 0], 0.5 * size[:, 1]) + r2
    main = border * rng.uniform(0.9, 1.15, n) * rng.choice([-1.0, 1.0], n)
    other = rng.uniform(-1.2, 1.2, n) * (np.where(along_x, 0.5 * size[:, 1], 0.5 * size[:, 0]) + r2)
    fx, fy = np.where(along_x, main, other), np.where(along_x, other, main)
    dist = np.radians(np.minimum(np.hypot(fx, fy), 179.5))
    # bearing: frame axes are (east, south) turned by gamma -> compass bearing from north, clockwise
    gam = np.radians(g[:, 0])
    east = fx * np.cos(gam) + fy * np.sin(gam)
    south = -fx * np.sin(gam) + fy * np.cos(gam)
    brg = np.arctan2(east, -south)
    lat1 = np.radians(90 - p1)
    lat2 = np.arcsin(np.clip(np.sin(lat1) * np.cos(dist) + np.cos(lat1) * np.sin(dist) * np.cos(brg), -1, 1))
    dlon = np.arctan2(np.sin(brg) * np.sin(dist) * np.cos(lat1), np.cos(dist) - np.sin(lat1) * np.sin(lat2))
    t2, p2 = (t1 + np.degrees(dlon)) % 360.0, 90 - np.degrees(lat2)
    b1 = np.stack([t1, p1, size[:, 0], size[:, 1], g[:, 0]], 1).astype(np.float32)
    b2 = np.stack([t2, p2, size[:, 2], size[:, 3], g[:, 1]], 1).astype(np.float32)
    k = np.arange(n) % 29 == 0                       # integer-valued boxes sharing a column: jitter_1's mask fires
    b1[k] = np.round(b1[k]); b2[k] = np.round(b2[k]); b2[k, 2] = b1[k, 2]
    return b1, b2


def _sphere_step(t1, p1, dist, brg):
    """(theta, phi) in degrees reached from (t1, p1) after `dist` radians along compass bearing `brg`."""
    lat1 = np.radians(90 - p1)
    lat2 = np.arcsin(np.clip(np.sin(lat1) * np.cos(dist) + np.cos(lat1) * np.sin(dist) * np.cos(brg), -1, 1))
    dlon = np.arctan2(np.sin(brg) * np.sin(dist) * np.cos(lat1), np.cos(dist) - np.sin(lat1) * np.sin(lat2))
    return (t1 + np.degrees(dlon)) % 360.0, 90 - np.degrees(lat2)


def _frames(t, p, g):
    """centre unit vector and the planar frame axes (width, height) of a box as 3-D vectors (sphk_fast.cuh: box_pre)."""
    t, p, g = np.radians(t), np.radians(p), np.radians(g)
    st, ct, sp, cp, sg, cg = np.sin(t), np.cos(t), np.sin(p), np.cos(p), np.sin(g), np.cos(g)
    u = np.stack([sp * ct, sp * st, cp], 1)
    east = np.stack([-st, ct, 0 * st], 1)
    south = np.stack([cp * ct, cp * st, -sp], 1)
    return u, cg[:, None] * east - sg[:, None] * south, sg[:, None] * east + cg[:, None] * south


def near_sat_border_pairs(n, seed, spread=0.06):
    """Pairs whose planar boxes (arc edges) just touch or just miss each other, i.e. sit within a few per cent of the
    border the separating-axis prefilter decides on: corner-to-edge and edge-to-edge contacts at every relative
    rotation, thin / square / oversize boxes, poles and the seam.  The touching distance along a bearing is found by a
    short fixed-point iteration on the float64 planar model (the relative rotation depends on the distance); the pairs
    are then placed at 0.95 .. 1 + spread times that distance (the float64 separating-axis verdict of this model agrees
    with "reference IoU == 0" on 99.8 % of them)."""
    rng = np.random.RandomState(seed)
    size = np.exp(rng.uniform(np.log(1.0), np.log(120.0), (n, 4)))
    size[::9] = rng.uniform(100, 400, (len(size[::9]), 4))
    size[4::9] = np.exp(rng.uniform(np.log(0.02), np.log(3.0), (len(size[4::9]), 4)))
    t1, p1 = rng.uniform(0, 360, n), np.degrees(np.arccos(rng.uniform(-1, 1, n)))
    p1[::17] = rng.uniform(0, 0.5, len(p1[::17])); p1[5::17] = 180 - rng.uniform(0, 0.5, len(p1[5::17]))
    g = rng.uniform(-90, 90, (n, 2))
    g[::23] = rng.choice([-1.0, 1.0], (len(g[::23]), 2)) * rng.uniform(178.5, 180, (len(g[::23]), 2))
    g[7::23] = np.round(g[7::23])                                   # parallel / perpendicular boxes now and then
    brg = rng.uniform(0, 2 * np.pi, n)
    hs = np.radians(np.minimum(size, 180.0)) * 0.5                  # half sizes hw1 hh1 hw2 hh2 (rad)
    dist = np.minimum(0.5 * (np.hypot(hs[:, 0], hs[:, 1]) + np.hypot(hs[:, 2], hs[:, 3])), 3.0)
    for _ in range(4):
        t2, p2 = _sphere_step(t1, p1, dist, brg)
        u1, e1, f1 = _frames(t1, p1, g[:, 0])
        u2, e2, f2 = _frames(t2, p2, g[:, 1])
        A, B = (u2 * e1).sum(1), (u2 * f1).sum(1)
        C, Dd = (u1 * e2).sum(1), (u1 * f2).sum(1)
        S2 = np.maximum(A * A + B * B, 1e-12)
        cr, sr = np.abs(C * A + Dd * B) / S2, np.abs(Dd * A - C * B) / S2
        S = np.sqrt(S2)
        ext = np.stack([hs[:, 0] + hs[:, 2] * cr + hs[:, 3] * sr, hs[:, 1] + hs[:, 2] * sr + hs[:, 3] * cr,
                        hs[:, 2] + hs[:, 0] * cr + hs[:, 1] * sr, hs[:, 3] + hs[:, 0] * sr + hs[:, 1] * cr], 1)
        proj = np.maximum(np.abs(np.stack([A, B, C, Dd], 1)) / S[:, None], 1e-9)
        dist = np.clip((ext / proj).min(1), 1e-3, 3.1)
    t2, p2 = _sphere_step(t1, p1, np.clip(dist * rng.uniform(0.95, 1.0 + spread, n), 1e-3, 3.13), brg)
    b1 = np.stack([t1, p1, size[:, 0], size[:, 1], g[:, 0]], 1).astype(np.float32)
    b2 = np.stack([t2, p2, size[:, 2], size[:, 3], g[:, 1]], 1).astype(np.float32)
    return b1, b2


def test_arc_from_hav_polynomial_against_float64(hostsim):
    """arc_from_hav (z = min(hav, 1 - hav), rsqrt + one Newton step, degree-6 polynomial; csrc/sphk_math.cuh) against
    2 asin(sqrt(hav)) in float64: uniform and log-uniform hav, the ends of the range, values a rounding error above 1."""
    rng = np.random.RandomState(0)
    hav = np.concatenate([rng.uniform(0, 1, 1_000_000), 10.0 ** rng.uniform(-12, 0, 500_000), 1 - 10.0 ** rng.uniform(-8, 0, 500_000),
                          [0.0, 1.0, 0.5, np.nextafter(np.float32(0.5), np.float32(1)), 1.0000001, 0.25, 0.75]]).astype(np.float32)
    arc = np.empty_like(hav)
    hostsim.hostsim_arc_from_hav(hav.ctypes.data_as(fp), ctypes.c_long(len(hav)), arc.ctypes.data_as(fp))
    want = 2 * np.arcsin(np.sqrt(np.minimum(hav.astype(np.float64), 1.0)))
    ulp = np.spacing(np.maximum(want, 1e-30).astype(np.float32)).astype(np.float64)
    err = np.abs(arc.astype(np.float64) - want) / ulp
    big = hav > 1e-20                       # (below: the 1e-30 floor of z; such arcs are far inside the clamped-acos zone anyway)
    assert err[big].max() <= 2.0 and err[big].mean() < 0.5, (err[big].max(), err[big].mean())
    assert arc[hav == 0].max() < 1e-14 and abs(float(arc[hav >= 1].min()) - np.pi) < 3e-7
    nan = np.array([np.nan], np.float32)
    hostsim.hostsim_arc_from_hav(nan.ctypes.data_as(fp), ctypes.c_long(1), nan.ctypes.data_as(fp))
    assert np.isnan(nan[0])


def test_half_difference_terms_are_bit_identical_to_sincos_deg(hostsim):
    """pair_job takes sin^2(x) and sin(2 x) of the half differences from sin2_and_sin_double_deg; the aligned path takes
    them from sincos_deg (s * s, 2 s c).  Both kernels must see the same bits (test_pairwise_equals_aligned_on_the_expansion
    checks that on the device; this is the host twin over every quadrant, the multiples of 45 and 90 degrees, and -0)."""
    rng = np.random.RandomState(1)
    x = np.concatenate([rng.uniform(-180, 180, 1_000_000), np.arange(-180, 181, 15, dtype=np.float64), [0.0, -0.0, 1e-30, -1e-30, 44.999996, 45.000004, 134.99999, 179.99998]]).astype(np.float32)
    n = len(x)
    out = [np.empty(n, np.float32) for _ in range(4)]
    hostsim.hostsim_half_angle_terms(x.ctypes.data_as(fp), ctypes.c_long(n), *[o.ctypes.data_as(fp) for o in out])
    s2, sin2x, s2_ref, sin2x_ref = out
    assert np.array_equal(s2.view(np.uint32), s2_ref.view(np.uint32))
    assert np.array_equal(sin2x.view(np.uint32), sin2x_ref.view(np.uint32))
    assert np.abs(sin2x.astype(np.float64) - np.sin(np.radians(2 * x.astype(np.float64)))).max() < 3e-7


def test_separating_axis_cull_is_conservative(hostsim):
    """The separating-axis prefilter of the N x M kernels (sphk_fast.cuh: pre_sat_disjoint) may only fire where the
    reference-order path, run WITHOUT any early-out, returns exactly 0 -- both transforms, both modes, every edge option,
    BFoV and RBFoV -- on pairs concentrated at the border it decides on (gaps of a few milliradians, inside its margins)
    and on pairs up to 1.6 x the touching distance apart, of which it has to decide most of the disjoint ones."""
    ub = ctypes.POINTER(ctypes.c_ubyte)
    fired = total = disjoint = 0
    for D in (4, 5):
        for seed in (0, 1, 2, 3):
            if seed < 2:
                b1, b2 = near_sat_border_pairs(250_000, seed + 10 * D, spread=0.06 if seed == 0 else 0.6)
            elif seed == 2:
                b1, b2 = near_box_edge_pairs(250_000, 3 + D)
            else:
                b1, b2 = near_touching_pairs(250_000, 11 + D)
            b1, b2 = np.ascontiguousarray(b1[:, :D]), np.ascontiguousarray(b2[:, :D])
            n = len(b1)
            dense, cull = np.empty(n, np.float32), np.empty(n, np.uint8)
            for kind, mode, edge in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 2)):
                hostsim.hostsim_prefilter(kind, b1.ctypes.data_as(fp), b2.ctypes.data_as(fp), ctypes.c_long(n), D, mode, edge,
                                          dense.ctypes.data_as(fp), cull.ctypes.data_as(ub))
                sat = (cull & 4) != 0
                bad = sat & (dense != 0)
                assert not bad.any(), (D, seed, kind, mode, edge, int(bad.sum()), np.where(bad)[0][:5], dense[bad][:5])
                if seed == 1 and edge == 0 and kind == 0 and mode == 0:
                    fired += int(sat.sum()); total += n; disjoint += int((dense == 0).sum())
    # of the pairs that ARE disjoint in the wide set (0.95 .. 1.6 x the touching distance), the test proves most
    # (a ninth of the set are boxes under 3 degrees, smaller than the margins, another ninth oversize boxes that never cull)
    assert fired > 0.5 * disjoint and disjoint > 0.5 * total, (fired, disjoint, total)


def test_separating_axis_cull_rate_on_the_sweep_distribution(hostsim):
    """On the box distribution of the bench's sweep (random RBFoV boxes, 1-100 degrees) the circle test leaves ~40 % of
    the pairs alive and the separating-axis test ends about a third of those: what goes on to the clipper is within a
    few per cent of the pairs that really overlap.  (The figures DESIGN.md and bench.py's early_out_rate quote.)"""
    from sph_retina_b200 import synthetic as S
    ub = ctypes.POINTER(ctypes.c_ubyte)
    n = 400_000
    b1 = np.ascontiguousarray(S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=0).numpy())
    b2 = np.ascontiguousarray(S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=1).numpy())
    dense, cull = np.empty(n, np.float32), np.empty(n, np.uint8)
    hostsim.hostsim_prefilter(0, b1.ctypes.data_as(fp), b2.ctypes.data_as(fp), ctypes.c_long(n), 5, 0, 0, dense.ctypes.data_as(fp),
                              cull.ctypes.data_as(ub))
    circle_live = (cull & 1) == 0
    sat = (cull & 4) != 0
    assert not (((cull & 5) != 0) & (dense != 0)).any()                 # neither test ever fires on an overlapping pair
    live = circle_live & ~sat
    positive = dense > 0
    assert 0.38 < circle_live.mean() < 0.43, circle_live.mean()
    assert 0.29 < (circle_live & sat).sum() / circle_live.sum() < 0.38
    assert 0.25 < positive.mean() < 0.28 and live.mean() - positive.mean() < 0.015, (live.mean(), positive.mean())


def test_box_frame_cull_is_conservative(hostsim):
    """The second prefilter test of the N x M scan loops (sphk_fast.cuh: pre_outside_box) may only fire where the
    reference-order path, run WITHOUT any early-out, returns exactly 0 -- for both transforms, every edge option and
    both modes; and it has to decide a good share of the near-border pairs the circle test leaves alive."""
    ub = ctypes.POINTER(ctypes.c_ubyte)
    fired = 0
    for D in (4, 5):
        for seed in (0, 1, 2):
            if seed < 2:
                b1, b2 = near_box_edge_pairs(250_000, seed + 10 * D)
            else:
                b1, b2 = near_touching_pairs(250_000, 7 + D)
            b1, b2 = np.ascontiguousarray(b1[:, :D]), np.ascontiguousarray(b2[:, :D])
            n = len(b1)
            dense, cull = np.empty(n, np.float32), np.empty(n, np.uint8)
            for kind, mode, edge in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 2)):
                hostsim.hostsim_prefilter(kind, b1.ctypes.data_as(fp), b2.ctypes.data_as(fp), ctypes.c_long(n), D, mode, edge,
                                          dense.ctypes.data_as(fp), cull.ctypes.data_as(ub))
                bad = (cull != 0) & (dense != 0)
                assert not bad.any(), (D, seed, kind, mode, edge, int(bad.sum()), np.where(bad)[0][:5], dense[bad][:5], cull[bad][:5])
                if seed < 2 and edge == 0:
                    only_box = ((cull & 3) == 2)
                    fired += int(only_box.sum())
                    assert only_box.mean() > 0.05, only_box.mean()       # it decides pairs the circle test cannot
    assert fired > 50_000


def test_stage0_cull_is_conservative(hostsim):
    """pair_far_apart (the approximate-trigonometry cull every aligned pair goes through) may only fire where the
    exact dead test of stage 1 fires too, with room for the MUFU sine error (1e-6 per sine, 6 sines: the GPU suite
    measures the real thing), and must never fire on out-of-range centres."""
    ub = ctypes.POINTER(ctypes.c_ubyte)
    for D in (4, 5):
        for seed in (0, 1):
            b1, b2 = near_touching_pairs(300_000, seed)
            b1, b2 = np.ascontiguousarray(b1[:, :D]), np.ascontiguousarray(b2[:, :D])
            n = len(b1)
            far, dead1, hav = np.empty(n, np.uint8), np.empty(n, np.uint8), np.empty(n, np.float32)
            for edge in (0, 1, 2):
                hostsim.hostsim_stage0(b1.ctypes.data_as(fp), b2.ctypes.data_as(fp), ctypes.c_long(n), D, edge,
                                       far.ctypes.data_as(ub), dead1.ctypes.data_as(ub), hav.ctypes.data_as(fp))
                assert not (far.astype(bool) & ~dead1.astype(bool)).any()
                if edge == 2:
                    assert not far.any()
                else:
                    assert far.sum() > 0.5 * dead1.sum()        # it decides (tiny boxes fall inside the haversine slack)
            # out-of-range / non-finite centres are never culled
            bad = b1.copy()
            bad[::4, 0] = 360.5; bad[1::4, 1] = -0.25; bad[2::4, 0] = np.nan; bad[3::4, 1] = 180.25
            hostsim.hostsim_stage0(bad.ctypes.data_as(fp), b2.ctypes.data_as(fp), ctypes.c_long(n), D, 0,
                                   far.ctypes.data_as(ub), dead1.ctypes.data_as(ub), hav.ctypes.data_as(fp))
            assert not far.any()


@pytest.mark.parametrize("fn", ["hostsim_iou_aligned_v2", "hostsim_iou_aligned_v3"])
def test_fast_path_inside_the_quirk_zones(hostsim, c_oracle, fn):
    """jitter_2 (size and angle triggers) and the clamped-acos zone are handled by the fast path itself: it must
    agree with the reference-order path of the same header and with the float64 C oracle there."""
    from test_oracle_golden import _c_aligned as c_aligned
    for D in (4, 5):
        b1, b2 = quirk_zone_pairs(120_000, D)
        b1, b2 = np.ascontiguousarray(b1[:, :D]), np.ascontiguousarray(b2[:, :D])
        for kind in (0, 1):
            want = hs_aligned(hostsim, kind, b1, b2)
            got, path = hs_fast(hostsim, fn, kind, b1, b2)
            truth = c_aligned(c_oracle, kind, b1, b2)
            assert (path == 1).mean() > 0.9, (path == 1).mean()             # ... and it really is the fast path
            e_fast, e_slow = np.abs(got - truth), np.abs(want - truth)
            # A pair whose |a1 - a2| (or size difference) lies within fp32 rounding of eps / eps' takes the other
            # branch of jitter_2 than the float64 run -- in ANY fp32 evaluation, the reference-order path included
            # (a few pairs in 1e5 here, a ~1e-4 step): both paths are held to the same allowance.
            for e in (e_fast, e_slow):
                assert (e > 1e-5).sum() <= 12 and e.max() < 1e-3, (fn, D, kind, (e > 1e-5).sum(), e.max(), int(np.argmax(e)))
            assert np.median(e_fast) < 2e-7
            assert (truth > 0.02).mean() > 0.9


def masked_or_clamped_pairs(n, seed):
    """Overlapping pairs on which jitter_1 is more than the lower clamp: one column shared exactly or within eps (the
    similarity mask: both boxes are shifted, sph_iou_api.py:246-251), integer-valued boxes, centres at or next to the
    upper end of their range (theta -> 360, phi -> 180: the clamp value 360 - eps is not a float)."""
    rng = np.random.RandomState(seed)
    t1, p1 = rng.uniform(5, 355, n), rng.uniform(15, 165, n)
    a1, b1, a2, b2 = (rng.uniform(5, 90, n) for _ in range(4))
    off = rng.uniform(0.1, 0.6, n) * np.minimum(np.hypot(a1, b1), np.hypot(a2, b2))
    brg = rng.uniform(0, 2 * np.pi, n)
    t2 = t1 + off * np.sin(brg) / np.maximum(np.sin(np.radians(p1)), 0.2)
    p2 = np.clip(p1 - off * np.cos(brg), 1, 179)
    g1, g2 = rng.uniform(-90, 90, n), rng.uniform(-90, 90, n)
    B1 = np.stack([t1, p1, a1, b1, g1], 1).astype(np.float32)
    B2 = np.stack([t2 % 360.0, p2, a2, b2, g2], 1).astype(np.float32)
    k = np.arange(n) % 8
    tiny = (rng.uniform(-1, 1, n) * 1.2e-4).astype(np.float32)
    for col in range(5):                                  # one column equal / within eps
        sel = k == col
        B2[sel, col] = B1[sel, col] + np.where(rng.rand(sel.sum()) < 0.5, 0.0, tiny[sel])
    sel = k == 5                                          # integer-valued annotations
    B1[sel] = np.round(B1[sel]); B2[sel] = np.round(B2[sel])
    sel = k == 6                                          # theta at the upper end (box 1, box 2, both)
    up = 360.0 - np.float32(10.0) ** rng.uniform(-5.5, -3.0, sel.sum()).astype(np.float32)
    which = rng.randint(0, 3, sel.sum())
    B1[sel, 0] = np.where(which != 1, up, B1[sel, 0]); B2[sel, 0] = np.where(which != 0, np.minimum(up + 3e-5, 360.0), B2[sel, 0])
    B2[sel, 0] = np.where(which == 0, 360.0 - rng.uniform(2, 30, sel.sum()), B2[sel, 0])
    B1[sel, 0] = np.where(which == 1, 360.0 - rng.uniform(2, 30, sel.sum()), B1[sel, 0])
    sel = k == 7                                          # phi at the upper end: both boxes next to the south pole
    B1[sel, 1] = 180.0 - np.float32(10.0) ** rng.uniform(-5.5, -3.0, sel.sum()).astype(np.float32)
    B2[sel, 1] = 180.0 - rng.uniform(2, 25, sel.sum())
    return B1, B2


def test_general_stage1_for_masked_and_clamped_pairs(hostsim, c_oracle):
    """The aligned kernel keeps pairs with an active similarity mask or an upper-end clamp on its batch path
    (pair_stage1_general: the hi + lo arithmetic of the reference-order path feeding the common stage 2 + clipper)
    instead of sending each to ~1500 instructions of reference-order code on one lane.  Same result as that path and
    as the float64 C oracle."""
    from test_oracle_golden import _c_aligned as c_aligned
    for D in (4, 5):
        b1, b2 = masked_or_clamped_pairs(160_000, 20 + D)
        b1, b2 = np.ascontiguousarray(b1[:, :D]), np.ascontiguousarray(b2[:, :D])
        for kind in (0, 1):
            want = hs_aligned(hostsim, kind, b1, b2)
            got, path = hs_fast(hostsim, "hostsim_iou_aligned_v3", kind, b1, b2)
            truth = c_aligned(c_oracle, kind, b1, b2)
            assert (path == 5).mean() > 0.5, np.bincount(path, minlength=6) / len(path)     # it is what is being tested
            gen = path == 5
            e_fast, e_slow = np.abs(got - truth)[gen], np.abs(want - truth)[gen]
            assert (truth[gen] > 0.02).mean() > 0.8
            # both paths are held to the same allowance (pairs whose size / angle difference sits within fp32 rounding of
            # jitter_2's eps take the other branch than the float64 run in ANY fp32 evaluation)
            for e in (e_fast, e_slow):
                assert (e > 1e-5).sum() <= 12 and e.max() < 1e-3, (D, kind, (e > 1e-5).sum(), e.max())
            assert np.median(e_fast) < 3e-7 and np.abs(got - want)[gen].max() < 1e-3
            assert (np.abs(got - want)[gen] > 5e-6).sum() <= 12


def _f5(v):
    return (ctypes.c_float * 5)(*(list(v) + [0.0] * (5 - len(v)))) if v is not None else None


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_coder_and_decode_loss(hostsim, box):
    """csrc/sphk_coder.cuh on the host: decode / encode of every coder variant against the reference's coder classes,
    and the fused decode -> loss step (value and d/d(deltas)) against the reference's autograd."""
    from test_oracle_golden import coder_kwargs
    g = load_golden("coder")
    anchors, deltas = np.ascontiguousarray(g[box + "_anchors"]), np.ascontiguousarray(g[box + "_deltas"])
    n, D = anchors.shape
    for name in ("plain", "norm", "ctr", "noclip"):
        kw = coder_kwargs(g, box, name)
        out = np.empty((n, D), np.float32)
        hostsim.hostsim_coder_decode(anchors.ctypes.data_as(fp), deltas.ctypes.data_as(fp), ctypes.c_long(n), D,
                                     _f5(kw.get("means")), _f5(kw.get("stds")), ctypes.c_float(16 / 1000),
                                     int(kw.get("clip_border", True)), int(kw.get("add_ctr_clamp", False)),
                                     ctypes.c_float(kw.get("ctr_clamp", 32)), out.ctypes.data_as(fp))
        want = g["%s_%s_decode_f64" % (box, name)]
        np.testing.assert_allclose(out, want, rtol=3e-6, atol=2e-5)
        dec = np.ascontiguousarray(g["%s_%s_decode_f32" % (box, name)])
        enc = np.empty((n, D), np.float32)
        hostsim.hostsim_coder_encode(anchors.ctypes.data_as(fp), dec.ctypes.data_as(fp), ctypes.c_long(n), D,
                                     _f5(kw.get("means")), _f5(kw.get("stds")), enc.ctypes.data_as(fp))
        np.testing.assert_allclose(enc, g["%s_%s_encode_f32" % (box, name)], rtol=2e-5, atol=2e-5)
    kw = coder_kwargs(g, box, "norm")
    d2, target = np.ascontiguousarray(g[box + "_loss_deltas"]), np.ascontiguousarray(g[box + "_target"])
    w = np.ascontiguousarray(g[box + "_weight"].mean(axis=1).astype(np.float32))
    scale = 1.5 / (float((w > 0).sum()) + float(np.finfo(np.float32).eps))
    gd = np.empty((n, D), np.float32)
    hostsim.hostsim_decode_loss.restype = ctypes.c_double
    total = hostsim.hostsim_decode_loss(anchors.ctypes.data_as(fp), d2.ctypes.data_as(fp), target.ctypes.data_as(fp),
                                        w.ctypes.data_as(fp), ctypes.c_long(n), D, _f5(kw["means"]), _f5(kw["stds"]),
                                        ctypes.c_float(16 / 1000), 1, 0, ctypes.c_float(32), ctypes.c_float(scale),
                                        gd.ctypes.data_as(fp))
    assert abs(total * scale - float(g[box + "_iou_loss_f64"])) < 2e-5
    ok, rel, rel32 = grad_rows_ok(gd, g[box + "_iou_gdeltas_f64"], g[box + "_iou_gdeltas_f32"], w > 0)
    assert ok.mean() > 0.99 and np.median(rel) < 3e-6, (ok.mean(), np.median(rel))
    assert (rel > 1e-4).sum() <= 0.5 * (rel32 > 1e-4).sum() + 2
    assert not gd[w == 0].any()


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_project_angle_variant(hostsim, box):
    """rbb_angle='project' (sph2pob_efficient.py:92-93, sph2pob_standard.py:99-100)."""
    g = load_golden("aligned_" + box)
    b1, b2 = np.ascontiguousarray(g["b1"], np.float32), np.ascontiguousarray(g["b2"], np.float32)
    P, D = b1.shape
    for kind, tr in ((0, "efficient"), (1, "standard")):
        out = np.empty(P, np.float32)
        hostsim.hostsim_iou_aligned_project(kind, b1.ctypes.data_as(fp), b2.ctypes.data_as(fp), ctypes.c_long(P), D, 0, 0,
                                            out.ctypes.data_as(fp))
        ok, err = within(out, g[tr + "_project_f64"], g[tr + "_project_f32"])
        assert ok.all(), (box, tr, np.where(~ok)[0], err[~ok])
        assert (err > 1e-5).sum() <= 2


def hs_obb_loss(lib, cls, kw, pred, target, up=None, use_double=1):
    kind, fun, flags, tau, alpha = other_loss_kernel_args(cls, kw)
    p, t = np.ascontiguousarray(pred, np.float32), np.ascontiguousarray(target, np.float32)
    n, D = p.shape
    L = 5 if kind == 6 else 1
    loss, g1, g2 = np.empty((n, L), np.float32), np.empty((n, D), np.float32), np.empty((n, D), np.float32)
    up_cols = 0
    if up is not None:
        up = np.ascontiguousarray(up, np.float32)
        up_cols = 1 if up.ndim == 1 else up.shape[1]
    cf = ctypes.c_float
    lib.hostsim_obb_loss(kind, fun, flags, cf(tau), cf(alpha), cf(1.0 / 9.0), cf(1e-6), 1, p.ctypes.data_as(fp), t.ctypes.data_as(fp),
                         up.ctypes.data_as(fp) if up is not None else None, up_cols, ctypes.c_long(n), D, use_double,
                         loss.ctypes.data_as(fp), g1.ctypes.data_as(fp), g2.ctypes.data_as(fp))
    return (loss if L > 1 else loss[:, 0]), g1, g2


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_other_losses_golden(hostsim, box):
    """Sph2PobGDLoss / KFLoss / L1Loss rows (csrc/sphk_obbloss.cuh: dual-number row losses in double, as the kernel is
    built, chained through the fp32 transform backward) against the reference's float64 run."""
    g = load_golden("other_losses_" + box)
    for name, (cls, kw) in other_loss_variants(g).items():
        loss, g1, g2 = hs_obb_loss(hostsim, cls, kw, g["pred"], g["target"])
        check_other_loss(name, g, loss, g1, g2)


def test_other_losses_upstream_is_linear(hostsim):
    """The per-row / per-element upstream (weights or incoming gradient) scales the gradients and nothing else."""
    g = load_golden("other_losses_rbfov")
    rng = np.random.default_rng(0)
    n = len(g["pred"])
    for cls, kw, cols in (("Sph2PobGDLoss", dict(loss_type="kld"), 1), ("Sph2PobKFLoss", {}, 1), ("Sph2PobL1Loss", {}, 5)):
        base = hs_obb_loss(hostsim, cls, kw, g["pred"], g["target"])
        w = rng.uniform(0.5, 2.0, size=n).astype(np.float32)
        got = hs_obb_loss(hostsim, cls, kw, g["pred"], g["target"], up=w)
        np.testing.assert_array_equal(got[0], base[0])
        for a, b in ((got[1], base[1] * w[:, None]), (got[2], base[2] * w[:, None])):
            rel, _ = grad_row_error(a[64:], b[64:])          # (rows 0-63: near-identical boxes, cancelling 1e5-sized terms)
            assert np.median(rel) < 1e-6 and (rel < 1e-4).mean() > 0.99
        if cols == 5:       # one column at a time adds up to all columns
            acc = np.zeros_like(base[1], dtype=np.float64)
            for k in range(5):
                u = np.zeros((n, 5), np.float32)
                u[:, k] = 1.0
                acc += hs_obb_loss(hostsim, cls, kw, g["pred"], g["target"], up=u)[1]
            rel, _ = grad_row_error(acc[64:], base[1][64:].astype(np.float64))
            assert np.median(rel) < 1e-6 and (rel < 1e-4).mean() > 0.99


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_naive_iou_golden(hostsim, box):
    """naive_iou_pair (planar IoU of the sph2pix boxes; the clipper for RBFoV) against the reference's float64 run."""
    g = load_golden("naive")
    got = hs_aligned(hostsim, 4, g[box + "_b1"], g[box + "_b2"])
    ok, err = within(got, g[box + "_aligned_f64"], g[box + "_aligned_f32"])
    assert ok.all(), (np.where(~ok)[0][:10], err[~ok][:10])
    assert np.nanmax(err) < 5e-6


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_unbiased_iou_golden(hostsim, box):
    """unbiased_iou_pair (the exact spherical IoU, double precision) against the reference's numpy classes run on float64
    copies of the boxes: identical, nested, unrelated and overlapping pairs.  The reference's own float32 run is off by up
    to 1.0 on 2-3 % of these pairs (its validity test works at 5e-9)."""
    g = load_golden("unbiased")
    got = hs_aligned(hostsim, 5, g[box + "_b1"], g[box + "_b2"])
    err = np.abs(got - g[box + "_aligned_f64"])
    err32 = np.abs(g[box + "_aligned_f32"] - g[box + "_aligned_f64"])
    assert err.max() < 1e-6, (np.where(err >= 1e-6)[0][:10], err.max())
    assert (err32 > 1e-5).sum() > 10


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_unbiased_iou_clear_rejection_is_exact(hostsim, box):
    """unbiased_iou_pair rejects most of the 40 candidate vertices on c . N_k alone (no square root, no division); that
    shortcut must never change a result: random unrelated pairs (mostly disjoint), wide boxes up to the 180-degree clamp,
    boxes at the poles, across the 0/360 seam and touching side by side, against the restatement (which evaluates
    round(V . N, 8) >= 0 for all 40 candidates of every pair)."""
    n = 6000
    b1 = O.generate_boxes(n, alpha_range=(1, 179.9), beta_range=(1, 179.9), box=box, seed=31)
    b2 = O.generate_boxes(n, alpha_range=(1, 179.9), beta_range=(1, 179.9), box=box, seed=32)
    b1[:1500, 2:4] *= 0.1; b2[:1500, 2:4] *= 0.1                        # small boxes: nearly all disjoint
    b1[1500:1700, 1] = torch.tensor([0.0, 180.0]).repeat(100)          # poles
    b2[1700:1900, 0] = (b1[1700:1900, 0] + 359.0) % 360                  # neighbours across the seam
    b2[1700:1900, 1] = b1[1700:1900, 1]
    b2[1900:2100] = b1[1900:2100]; b2[1900:2100, 0] = (b1[1900:2100, 0] + b1[1900:2100, 2]) % 360   # touching side by side
    want = O.unbiased_iou(b1, b2, is_aligned=True).numpy()
    got = hs_aligned(hostsim, 5, b1.numpy(), b2.numpy())
    err = np.abs(got - want)
    assert err.max() < 1e-6, (np.where(err >= 1e-6)[0][:10], err.max())
    assert (want < 1e-6).mean() > 0.2 and (want > 0.01).mean() > 0.2      # both regimes are exercised


def test_sph2pob_legacy_golden(hostsim):
    """sph2pob_legacy_iou_pair (kind 6, BFoV) against the reference's own sph2pob_legacy_iou: modes, edges, the 7 known
    answers of tests/test_all_ious.py.  Criterion of SURVEY.md 8(c): within 1e-5 of the reference's fp64 run, or no
    further from it than the reference's fp32 run (whose clamped acos near 1 is good to ~2e-4 rad only)."""
    g = load_golden("legacy")
    for key, mode, edge in (("iou", 0, 0), ("iof", 1, 0), ("chord", 0, 1), ("tangent", 0, 2)):
        got = hs_aligned(hostsim, 6, g["b1"], g["b2"], mode, edge)
        ok, err = within(got, g[key + "_f64"], g[key + "_f32"])
        # The golden values go through the reference's vendored diff_iou_rotated_2d (the stand-in for mmcv's box_iou_rotated,
        # SURVEY.md 8c), whose crossing parameter is num / (den + 1e-8) with den ~ w * h: boxes of ~1 degree are 3e-5 off
        # the exact intersection there, and the <0.06 degree specks jitter_2 inflates to 2.5e-4 rad (den ~ 1e-8, which the
        # legacy transform -- unlike efficient / standard -- leaves overlapping) are off by a factor.  The kernel clips exactly.
        small = np.minimum(g["b1"][:, 2:4].min(axis=1), g["b2"][:, 2:4].min(axis=1))
        ok |= (small < 1.5) & (err < 1e-4)
        speck = degenerate_pairs(g["b1"], g["b2"])
        assert ok[~speck].all(), (key, np.where(~ok & ~speck)[0][:10], err[~ok & ~speck][:10])
        assert (err[~speck] > 1e-5).sum() <= 4 and np.median(err) < 1e-7
        assert np.nanmin(got) >= 0.0 and np.nanmax(got) <= 1.0
    # the zero-size pair (both boxes become eps-sized, then 2.5e-4 / 1.25e-4 rad squares whose corners overlap by 6.5e-5 x
    # 6.3e-5 after jitter_2's shifts): exact IoU = 4.08e-9 / 7.48e-8
    i = int(np.where((g["b1"] == 0).all(axis=1) & (g["b2"] == 0).all(axis=1))[0][0])
    assert abs(hs_aligned(hostsim, 6, g["b1"][i:i + 1], g["b2"][i:i + 1])[0] - 0.0545) < 2e-3
    np.testing.assert_allclose(hs_aligned(hostsim, 6, g["kat_b1"], g["kat_b2"]), g["kat_iou"], atol=5e-6)
    np.testing.assert_allclose(g["kat_iou"], [0.232635, 0.333749, 0.617413, 0.138463, 0.286415, 0.203624, 0.554315], atol=2e-6)


@pytest.mark.parametrize("box", ["bfov", "rbfov"])
def test_unbiased_iou_near_coincident_pairs(hostsim, box):
    """The vertex test of unbiased_iou_pair works on t_k = c . N_k against 0.5e-8 (|c| + delta) instead of the reference's
    round(V . N_k, 8) >= 0 on the normalised vector: the two can only differ where the reference itself is decided by
    rounding noise.  Identical, integer-valued and slightly perturbed pairs (where boundary circles nearly coincide) against
    the restatement, which follows the reference's arithmetic."""
    n = 3000
    for scale in (0.0, 1e-4, 1e-2, 1.0):
        b1 = O.generate_boxes(n, alpha_range=(0.5, 120), beta_range=(0.5, 120), box=box, seed=7)
        b2 = (b1 + torch.randn_like(b1) * scale).clamp(min=0.2)
        b2[:, 0].clamp_(0, 360); b2[:, 1].clamp_(0, 180); b2[:, 2:4].clamp_(max=179.5)
        b1[:500], b2[:500] = b1[:500].round(), b2[:500].round()          # integer annotations: jitter_1's mask fires
        want = O.unbiased_iou(b1, b2, is_aligned=True).numpy()
        err = np.abs(hs_aligned(hostsim, 5, b1.numpy(), b2.numpy()) - want)
        assert (err > 1e-6).sum() <= 1, (scale, np.where(err > 1e-6)[0][:10], err.max())


def test_box_format_golden(hostsim):
    """box_format_row (csrc/sphk_format.cuh) against the reference's box_formator functions / classes: the pure-arithmetic
    formats bit for bit, the ones through tan / atan / sin / cos to 1e-4 (libm against torch's vectorised kernels)."""
    g = load_golden("box_format")
    for fmt, key, d_out, want_key, size, exact in BOX_FORMAT_CASES:
        x = np.ascontiguousarray(g[key], np.float32)
        if fmt == 5 and d_out == 4:
            x = np.ascontiguousarray(x[:, :4])
        out = np.empty((len(x), d_out), np.float32)
        hostsim.hostsim_box_format(fmt, x.ctypes.data_as(fp), ctypes.c_long(len(x)), x.shape[1], d_out, ctypes.c_float(size[0]),
                                   ctypes.c_float(size[1]), out.ctypes.data_as(fp))
        want = g[want_key]
        if exact:
            assert np.array_equal(out, want), (want_key, np.abs(out - want).max())
        else:
            np.testing.assert_allclose(out, want, rtol=1e-4, atol=1e-4, err_msg=want_key)   # tan(alpha / 2) near alpha = 180 amplifies an ulp 200 x


def test_distance_point_coder_golden(hostsim, monkeypatch):
    """The product's DistancePointSphBBoxCoder host code with the box-format launch replaced by the same row function built
    for the host (hostsim_box_format = csrc/sphk_format.cuh): what the GPU suite checks through the C ABI, on the CPU."""
    from conftest import check_distance_coder
    from sph_retina_b200 import _native
    from sph_retina_b200.sphdet.bbox.coder import distance_point_sph_bbox_coder as M

    def host_box_format(fmt, boxes, d_out, img_size=(512, 1024)):
        x = np.ascontiguousarray(boxes.detach().numpy(), np.float32)
        out = np.empty((len(x), d_out), np.float32)
        hostsim.hostsim_box_format(_native.BOX_FORMAT[fmt], x.ctypes.data_as(fp), ctypes.c_long(len(x)), x.shape[1], d_out,
                                   ctypes.c_float(img_size[0]), ctypes.c_float(img_size[1]), out.ctypes.data_as(fp))
        return torch.from_numpy(out)
    monkeypatch.setattr(_native, "box_format", host_box_format)
    check_distance_coder(M, "cpu")
    # the anchor-free heads' test-time block up to the NMS (which has no host twin: the GPU suite runs it)
    from conftest import check_anchor_free_post_processing
    from sph_retina_b200.sphdet.models.heads import sph_bbox_post
    check_anchor_free_post_processing(sph_bbox_post, M, O, "cpu", with_nms=False)
