"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch, dtype-generic) of the
reference's spherical-box IoU hot path.  It is the *checker* for the CUDA
kernels; nothing in the product package imports it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may call it.

Parity status: PINNED against the reference's own Python code run in this
container (``oracle/ref_harness.py`` + ``oracle/make_golden.py`` ->
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them on any
box).  The one third-party piece, ``mmcv.ops.box_iou_rotated`` /
``diff_iou_rotated_2d`` (mmcv-full 1.6.0, not vendored, not installable
offline), is UNPINNED at the mmcv boundary; it is replaced by a restatement of
the reference's own vendored ``sphdet/iou/diff_iou_rotated.py`` (which the
reference's test ``tests/test_sph_iou_loss.py:21-34`` asserts is within 1e-6
mean of the mmcv kernel).

Run it with float64 inputs for "the truth" and float32 inputs for "the
reference as shipped" (SURVEY.md section 8c).  Every function cites the
reference lines it follows (paths relative to the reference root).
"""
from __future__ import annotations

import math

import torch

# sph_iou_api.py:223,245  eps = 1e-4 * 1.2345678 ; :232 eps = 1e-3 * 1.2345678
EPS_SMALL = 1e-4 * 1.2345678
EPS_ANGLE = 1e-3 * 1.2345678
# sph2pob_efficient.py:205 / sph2pob_standard.py:214  clamp(-1+1e-7, 1-1e-7)
COS_LO, COS_HI = -1 + 1e-7, 1 - 1e-7


# --------------------------------------------------------------------------- #
# jitters (sph_iou_api.py:222-260)
# --------------------------------------------------------------------------- #
def jitter_spherical(b1: torch.Tensor, b2: torch.Tensor):
    """sph_iou_api.py:244-260 (jiter_spherical_bboxes), out-of-place.

    Gradient dataflow is the reference's: identity through the masked shift,
    zero where a clamp is active."""
    e = EPS_SMALL
    near = ((b1 - b2).abs() < e).any(dim=1, keepdim=True)
    b1 = torch.where(near, b1 - 2 * e, b1)
    b2 = torch.where(near, b2 + e, b2)
    full, half = 360, 180
    c1 = [b1[:, 0].clamp(2 * e, full - e)] + [b1[:, k].clamp(2 * e, half - e) for k in (1, 2, 3)]
    c2 = [b2[:, 0].clamp(e, full - 2 * e)] + [b2[:, k].clamp(e, half - 2 * e) for k in (1, 2, 3)]
    if b1.size(1) == 5:
        # :256-258 -- b2's gamma is clamped twice, b1's never (reference quirk, kept)
        g2 = b2[:, 4].clamp(-full + e, full - 2 * e).clamp(-full + 2 * e, full - e)
        c1.append(b1[:, 4])
        c2.append(g2)
    return torch.stack(c1, dim=1), torch.stack(c2, dim=1)


def jitter_rotated(o1: torch.Tensor, o2: torch.Tensor):
    """sph_iou_api.py:222-242 (jiter_rotated_bboxes), out-of-place."""
    e = EPS_SMALL
    cols = [0, 2, 3, 4]
    near = ((o1[:, cols] - o2[:, cols]).abs() < e).any(dim=1, keepdim=True)
    add1 = torch.tensor([e, e, 2 * e, 2 * e, e], dtype=o1.dtype)
    add2 = torch.tensor([2 * e, 2 * e, e, e, 5 * e], dtype=o1.dtype)
    o1 = torch.where(near, o1 + add1, o1)
    o2 = torch.where(near, o2 + add2, o2)
    ea = EPS_ANGLE
    same_angle = (o1[:, 4] - o2[:, 4]).abs() < ea
    a1 = torch.where(same_angle, o1[:, 4] + ea, o1[:, 4])
    a2 = torch.where(same_angle, o2[:, 4] + 2 * ea, o2[:, 4])
    pi = math.pi
    w1 = o1[:, 2:4].clamp(min=2 * ea / 10)
    w2 = o2[:, 2:4].clamp(min=ea / 10)
    a1 = a1.clamp(-2 * pi + 2 * ea, 2 * pi - ea)
    a2 = a2.clamp(-2 * pi + ea, 2 * pi - 2 * ea)
    o1 = torch.cat([o1[:, 0:2], w1, a1[:, None]], dim=1)
    o2 = torch.cat([o2[:, 0:2], w2, a2[:, None]], dim=1)
    return o1, o2


# --------------------------------------------------------------------------- #
# small vector helpers shared by both transforms
# --------------------------------------------------------------------------- #
def _unit(v):
    # F.normalize(dim=1): v / max(|v|, 1e-12)
    return v / v.norm(dim=1, keepdim=True).clamp_min(1e-12)


def _angle_between(a, b):
    """sph2pob_efficient.py:192-208: |acos(clamp(a^ . b^))| in radians."""
    c = (_unit(a) * _unit(b)).sum(dim=1).clamp(COS_LO, COS_HI)
    return torch.acos(c).abs()


def _turn_sign(a, b, ref):
    """sph2pob_efficient.py:211-226: +1 where (a x b) . ref < 0 else -1 (zero -> -1)."""
    crit = (torch.linalg.cross(a, b, dim=1) * ref).sum(dim=1) < 0
    one = torch.ones((), dtype=a.dtype)
    return torch.where(crit, one, -one)


def _centre_and_tangent(theta, phi):
    """sph2pob_efficient.py:111-162: unit centre c and the southward tangent d."""
    st, ct, sp, cp = torch.sin(theta), torch.cos(theta), torch.sin(phi), torch.cos(phi)
    c = torch.stack([sp * ct, sp * st, cp], dim=1)
    d = torch.stack([cp * ct, cp * st, -sp], dim=1)
    return c, d


def _edge(fov, mode):
    """sph2pob_efficient.py:100-108."""
    if mode == "arc":
        return fov
    if mode == "tangent":
        return 2 * torch.tan(fov / 2)
    if mode == "chord":
        return 2 * torch.sin(fov / 2)
    raise NotImplementedError(mode)


# --------------------------------------------------------------------------- #
# Sph2Pob-efficient (sph2pob_efficient.py:9-73), output angles in radians
# --------------------------------------------------------------------------- #
def sph2pob_efficient(b1, b2, rbb_edge="arc", rbb_angle="equator"):
    assert rbb_edge in ("arc", "chord", "tangent") and rbb_angle in ("equator", "project")
    r1, r2 = torch.deg2rad(b1), torch.deg2rad(b2)
    cg, dg = _centre_and_tangent(r1[:, 0], r1[:, 1])
    cp, dp = _centre_and_tangent(r2[:, 0], r2[:, 1])
    z = torch.linalg.cross(cg, cp, dim=1)          # :49
    ref = (cg + cp) / 2                            # :50
    arc = _angle_between(cg, cp)                   # :51

    def internal(d):                               # :81-97
        if rbb_angle == "project":
            d = torch.cat([torch.zeros_like(d[:, :1]), d[:, 1:]], dim=1)
        return _angle_between(d, z) * _turn_sign(z, d, ref)

    ag, ap = internal(dg), internal(dp)
    if b1.size(1) == 5 and b2.size(1) == 5:        # :55-57
        ag = ag - r1[:, 4]
        ap = ap - r2[:, 4]
    zero = torch.zeros_like(arc)
    o1 = torch.stack([zero, zero, _edge(r1[:, 2], rbb_edge), _edge(r1[:, 3], rbb_edge), ag], dim=1)
    o2 = torch.stack([arc, zero, _edge(r2[:, 2], rbb_edge), _edge(r2[:, 3], rbb_edge), ap], dim=1)
    return o1, o2


# --------------------------------------------------------------------------- #
# Sph2Pob-legacy (sph2pob_legacy.py:8-31), BFoV only, output angles in radians
# --------------------------------------------------------------------------- #
def sph2pob_legacy(b1, b2, rbb_edge="arc", rbb_angle=None):
    """The hand-crafted first version of the transform ('the calculated internal angle is not accurate',
    sph2pob_legacy.py:10-11): both centres moved to the equator keeping their great-circle distance and their latitude
    difference (:38-82), the box angle measured between the meridian tangent at the box and the one at the pair's
    mid-longitude (:99-129).  ``torch.chunk(box, 4)`` (:52-53) makes it BFoV-only."""
    assert rbb_edge in ("arc", "chord", "tangent")
    assert b1.size(1) == 4 and b2.size(1) == 4, "sph2pob_legacy takes BFoV boxes (torch.chunk(., 4) at sph2pob_legacy.py:52)"
    pi = math.pi
    far = (b1[:, 0] - b2[:, 0]).abs() > 180                                    # standardize_spherical_box :224-244
    tg = torch.where(far, torch.remainder(b1[:, 0] + 180, 360), b1[:, 0])
    tp = torch.where(far, torch.remainder(b2[:, 0] + 180, 360), b2[:, 0])
    r1 = torch.deg2rad(torch.stack([tg, b1[:, 1], b1[:, 2], b1[:, 3]], dim=1))
    r2 = torch.deg2rad(torch.stack([tp, b2[:, 1], b2[:, 2], b2[:, 3]], dim=1))
    # transform_position :38-82 ('convention' radians :203-221: theta - pi, pi/2 - phi)
    theta_g, phi_g, theta_p, phi_p = r1[:, 0] - pi, pi / 2 - r1[:, 1], r2[:, 0] - pi, pi / 2 - r2[:, 1]
    phi_i = (phi_g + phi_p) / 2
    phi_g_, phi_p_ = phi_g - phi_i, phi_p - phi_i
    d_phi, d_theta = (phi_g - phi_p).abs(), (theta_g - theta_p).abs()
    L = 2 * torch.arcsin(torch.sqrt(torch.sin(d_phi / 2) ** 2 + torch.cos(phi_g) * torch.cos(phi_p) * torch.sin(d_theta / 2) ** 2))
    d_theta_ = (2 * torch.arcsin(torch.sqrt((torch.sin(L / 2) ** 2 - torch.sin(d_phi / 2) ** 2)
                                            / (torch.cos(phi_g_) * torch.cos(phi_p_))))).abs()
    one = torch.ones((), dtype=b1.dtype)
    theta_p_ = d_theta_ * torch.where(theta_p > theta_g, one, -one)

    # transfrom_anlge :99-129 ('math' radians = plain deg2rad)
    def internal(theta, phi, theta_ref):
        d_box = torch.stack([torch.cos(phi) * torch.cos(theta), torch.cos(phi) * torch.sin(theta), -torch.sin(phi)], dim=1)
        d_ref = torch.stack([torch.cos(phi) * torch.cos(theta_ref), torch.cos(phi) * torch.sin(theta_ref), -torch.sin(phi)], dim=1)
        ang = torch.acos((_unit(d_box) * _unit(d_ref)).sum(dim=1).clamp(COS_LO, COS_HI)) / pi * 180          # degrees, :197-200
        ang = ang.abs()
        keep = ((theta >= theta_ref) & (phi < pi / 2)) | ((theta <= theta_ref) & (phi > pi / 2))
        return torch.deg2rad(torch.where(keep, ang, -ang))                                                    # :275-277
    theta_mid = (r1[:, 0] + r2[:, 0]) / 2
    ag, ap = internal(r1[:, 0], r1[:, 1], theta_mid), internal(r2[:, 0], r2[:, 1], theta_mid)
    zero = torch.zeros_like(ag)
    o1 = torch.stack([zero, phi_g_, _edge(r1[:, 2], rbb_edge), _edge(r1[:, 3], rbb_edge), ag], dim=1)
    o2 = torch.stack([theta_p_, phi_p_, _edge(r2[:, 2], rbb_edge), _edge(r2[:, 3], rbb_edge), ap], dim=1)
    return o1, o2


# --------------------------------------------------------------------------- #
# Sph2Pob-standard (sph2pob_standard.py:8-80), output angles in radians
# --------------------------------------------------------------------------- #
def _frame_from_angles(theta, phi):
    """sph2pob_standard.py:238-261: rows (look, down, right)."""
    st, ct, sp, cp = torch.sin(theta), torch.cos(theta), torch.sin(phi), torch.cos(phi)
    zero = torch.zeros_like(theta)
    look = torch.stack([sp * ct, sp * st, cp], dim=1)
    down = torch.stack([cp * ct, cp * st, -sp], dim=1)
    right = torch.stack([st, -ct, zero], dim=1)
    return torch.stack([look, down, right], dim=1)   # [P,3,3], row-major


def _pair_frame(cg, cp, theta_mid, phi_mid):
    """sph2pob_standard.py:264-297 (compute_rotate_matrix_auto)."""
    look = _unit(cg + cp)
    right = _unit(cp - cg)
    up = torch.linalg.cross(look, right, dim=1)
    R = torch.stack([look, right, up], dim=1)
    normal = ((cg - cp).abs().sum(dim=1) > 1e-8)[:, None, None]
    return torch.where(normal, R, _frame_from_angles(theta_mid, phi_mid))


def _rotate_tangent_by_gamma(theta, phi, gamma, d):
    """sph2pob_standard.py:47-54,300-314: T^T Rx(-gamma) T applied to d."""
    T = _frame_from_angles(theta, phi)
    g = -gamma
    sg, cg_ = torch.sin(g), torch.cos(g)
    one, zero = torch.ones_like(g), torch.zeros_like(g)
    Rx = torch.stack([torch.stack([one, zero, zero], dim=1),
                      torch.stack([zero, cg_, -sg], dim=1),
                      torch.stack([zero, sg, cg_], dim=1)], dim=1)
    M = T.transpose(1, 2) @ (Rx @ T)
    return (M @ d[:, :, None]).squeeze(-1)


def _angle_between_deg(a, b):
    # sph2pob_standard.py:202-217: same clamp, but returned in degrees
    c = (_unit(a) * _unit(b)).sum(dim=1).clamp(COS_LO, COS_HI)
    return (torch.acos(c) / torch.pi * 180).abs()


def sph2pob_standard(b1, b2, rbb_edge="arc", rbb_angle="equator"):
    r1, r2 = torch.deg2rad(b1), torch.deg2rad(b2)
    theta_mid, phi_mid = (r1[:, 0] + r2[:, 0]) / 2, (r1[:, 1] + r2[:, 1]) / 2
    cg, dg = _centre_and_tangent(r1[:, 0], r1[:, 1])
    cp, dp = _centre_and_tangent(r2[:, 0], r2[:, 1])
    R = _pair_frame(cg, cp, theta_mid, phi_mid)
    if b1.size(1) == 5:                               # :47 (tests b1 only, b2 assumed alike)
        dg = _rotate_tangent_by_gamma(r1[:, 0], r1[:, 1], r1[:, 4], dg)
        dp = _rotate_tangent_by_gamma(r2[:, 0], r2[:, 1], r2[:, 4], dp)
    rot = lambda v: (R @ v[:, :, None]).squeeze(-1)
    cg, cp, dg, dp = rot(cg), rot(cp), rot(dg), rot(dp)
    ex = torch.tensor([1.0, 0.0, 0.0], dtype=b1.dtype).expand_as(cg)
    ez = torch.tensor([0.0, 0.0, 1.0], dtype=b1.dtype).expand_as(cg)

    def internal(d):                                  # :88-108, degrees
        if rbb_angle == "project":
            d = torch.cat([torch.zeros_like(d[:, :1]), d[:, 1:]], dim=1)
        return _angle_between_deg(d, ez) * _turn_sign(ez, d, ex)

    def to_sph(c):                                    # :175-199, degrees then deg2rad
        phi = _angle_between_deg(c, ez)
        cxy = torch.cat([c[:, :2], torch.zeros_like(c[:, :1])], dim=1)
        theta = _angle_between_deg(cxy, ex) * _turn_sign(ex, cxy, -ez)
        return torch.deg2rad(theta), torch.deg2rad(phi)

    ag, ap = torch.deg2rad(internal(dg)), torch.deg2rad(internal(dp))   # :342-364 ('rad')
    tg, pg = to_sph(cg)
    tp, pp = to_sph(cp)
    o1 = torch.stack([tg, pg, _edge(r1[:, 2], rbb_edge), _edge(r1[:, 3], rbb_edge), ag], dim=1)
    o2 = torch.stack([tp, pp, _edge(r2[:, 2], rbb_edge), _edge(r2[:, 3], rbb_edge), ap], dim=1)
    return o1, o2


# --------------------------------------------------------------------------- #
# rotated IoU, restating sphdet/iou/diff_iou_rotated.py:20-343
# --------------------------------------------------------------------------- #
_TINY = 1e-8   # diff_iou_rotated.py:17 EPSILON


def obb_corners(o):
    """diff_iou_rotated.py:297-322: corners (+,+),(-,+),(-,-),(+,-) rotated CCW by the angle."""
    sx = torch.tensor([0.5, -0.5, -0.5, 0.5], dtype=o.dtype)
    sy = torch.tensor([0.5, 0.5, -0.5, -0.5], dtype=o.dtype)
    lx, ly = sx * o[:, 2:3], sy * o[:, 3:4]
    s, c = torch.sin(o[:, 4:5]), torch.cos(o[:, 4:5])
    return torch.stack([lx * c - ly * s + o[:, 0:1], lx * s + ly * c + o[:, 1:2]], dim=-1)  # [P,4,2]


def _edge_crossings(q1, q2):
    """diff_iou_rotated.py:20-60: 4x4 segment crossings, strict interior on both."""
    a0, a1 = q1[:, :, None, :], q1.roll(-1, dims=1)[:, :, None, :]
    b0, b1 = q2[:, None, :, :], q2.roll(-1, dims=1)[:, None, :, :]
    da, db = a0 - a1, b0 - b1
    den = da[..., 0] * db[..., 1] - da[..., 1] * db[..., 0]
    ab = a0 - b0
    tn = ab[..., 0] * db[..., 1] - ab[..., 1] * db[..., 0]
    un = da[..., 0] * ab[..., 1] - da[..., 1] * ab[..., 0]
    par = den == 0
    t = torch.where(par, torch.full_like(den, -1.0), tn / den)
    u = torch.where(par, torch.full_like(den, -1.0), -un / den)
    ok = (t > 0) & (t < 1) & (u > 0) & (u < 1)
    ts = tn / (den + _TINY)
    pts = a0 + ts[..., None] * (a1 - a0)
    return pts * ok[..., None].to(pts.dtype), ok


def _corners_inside(q, box):
    """diff_iou_rotated.py:63-89: closed test via projections on two box edges."""
    a, b, d = box[:, 0:1], box[:, 1:2], box[:, 3:4]
    ab, ad, am = b - a, d - a, q - a
    pab, pad = (ab * am).sum(-1), (ad * am).sum(-1)
    return (pab >= 0) & (pab <= (ab * ab).sum(-1)) & (pad >= 0) & (pad <= (ad * ad).sum(-1))


def rotated_intersection_area(q1, q2):
    """diff_iou_rotated.py:278-295 on corner tensors [P,4,2]."""
    P = q1.size(0)
    pts, ok = _edge_crossings(q1, q2)
    in12, in21 = _corners_inside(q1, q2), _corners_inside(q2, q1)
    with torch.no_grad():
        # :196-223 coincident corners: keep box-1's copy only
        same = (q1[:, :, None, :] == q2[:, None, :, :]).all(-1)   # [P,4(i),4(j)]
        in12 = in12 | same.any(dim=2)
        in21 = in21 & ~same.any(dim=1)
    verts = torch.cat([q1, q2, pts.reshape(P, 16, 2)], dim=1)          # [P,24,2]
    mask = torch.cat([in12, in21, ok.reshape(P, 16)], dim=1)            # [P,24]
    with torch.no_grad():
        n = mask.sum(dim=1)
        mean = (verts * mask[..., None].to(verts.dtype)).sum(dim=1, keepdim=True) / n[:, None, None]
        rel = verts - mean
        x = torch.where(mask, rel[..., 0], torch.full_like(rel[..., 0], -1e6))
        y = torch.where(mask, rel[..., 1], torch.full_like(rel[..., 1], 1e-6))
        order = torch.argsort(torch.atan2(y, x), dim=-1)
        order.scatter_(1, n[:, None].clamp(max=23), order[:, :1].clone())   # close the loop
        order = order[:, :9]
        keep = (torch.arange(9)[None, :] < (n[:, None] + 1)) & (n[:, None] >= 3)
    poly = torch.gather(verts, 1, order[..., None].expand(-1, -1, 2)) * keep[..., None].to(verts.dtype)
    cross = poly[:, :-1, 0] * poly[:, 1:, 1] - poly[:, :-1, 1] * poly[:, 1:, 0]
    return cross.sum(dim=1).abs() / 2


def rotated_iou(o1, o2, mode="iou"):
    """diff_iou_rotated.py:325-343 (+ 'iof' as mmcv.ops.box_iou_rotated defines it)."""
    inter = rotated_intersection_area(obb_corners(o1), obb_corners(o2))
    a1, a2 = o1[:, 2] * o1[:, 3], o2[:, 2] * o2[:, 3]
    return inter / (a1 + a2 - inter) if mode == "iou" else inter / a1


# --------------------------------------------------------------------------- #
# public IoU functions (sph_iou_api.py:48-177)
# --------------------------------------------------------------------------- #
def _expand(b1, b2, is_aligned):
    """sph_iou_api.py:59-64: pair p = i*C + j = (b1[i], b2[j])."""
    if is_aligned:
        assert b1.size(0) == b2.size(0)
        return b1, b2
    R, C = b1.size(0), b2.size(0)
    return b1.repeat_interleave(C, dim=0), b2.repeat(R, 1)


def sph2pob_iou(b1, b2, transform="efficient", mode="iou", is_aligned=False,
                rbb_edge="arc", rbb_angle="equator", chunk=1 << 18):
    """sph_iou_api.py:48-98 (_sph2pob_iou_auxiliary with the efficient/standard transform)."""
    assert mode in ("iou", "iof")
    R, C = b1.size(0), b2.size(0)
    if R * C == 0:
        return b1.new_zeros((R, 1)) if is_aligned else b1.new_zeros((R, C))
    fn = {"efficient": sph2pob_efficient, "standard": sph2pob_standard, "legacy": sph2pob_legacy}[transform]
    e1, e2 = _expand(b1, b2, is_aligned)
    outs = []
    for s in range(0, e1.size(0), chunk):
        j1, j2 = jitter_spherical(e1[s:s + chunk], e2[s:s + chunk])
        o1, o2 = fn(j1, j2, rbb_edge=rbb_edge, rbb_angle=rbb_angle)
        o1, o2 = jitter_rotated(o1, o2)
        outs.append(rotated_iou(o1, o2, mode).clamp(0, 1))
    out = torch.cat(outs)
    return out if is_aligned else out.view(R, C)


def _wrap_far_apart(b1, b2):
    """approximate_ious.py:60-81: where |dtheta| > 180 shift both thetas by 180 mod 360."""
    far = (b1[:, 0] - b2[:, 0]).abs() > 180
    t1 = torch.where(far, (b1[:, 0] + 180) % 360, b1[:, 0])
    t2 = torch.where(far, (b2[:, 0] + 180) % 360, b2[:, 0])
    return torch.cat([t1[:, None], b1[:, 1:]], 1), torch.cat([t2[:, None], b2[:, 1:]], 1)


def _to_convention(b):
    """approximate_ious.py:83-100: radians with theta-pi, pi/2-phi."""
    r = torch.deg2rad(b)
    return r[:, 0] - torch.pi, torch.pi / 2 - r[:, 1], r[:, 2], r[:, 3]


def _approx_iou(b1, b2, kind):
    """approximate_ious.py:3-55 (sph_iou_aligned / fov_iou_aligned)."""
    b1, b2 = _wrap_far_apart(b1, b2)
    tg, pg, ag, bg = _to_convention(b1)
    tp, pp, ap, bp = _to_convention(b2)
    if kind == "sph":
        lo = torch.max(tg - ag / 2, tp - ap / 2)
        hi = torch.min(tg + ag / 2, tp + ap / 2)
    else:
        delta = (tp - tg) * torch.cos((pg + pp) / 2)
        lo = torch.max(-ag / 2, delta - ap / 2)
        hi = torch.min(ag / 2, delta + ap / 2)
    plo = torch.max(pg - bg / 2, pp - bp / 2)
    phi_ = torch.min(pg + bg / 2, pp + bp / 2)
    inter = (hi - lo).clamp(min=0) * (phi_ - plo).clamp(min=0)
    return inter / (ag * bg + ap * bp - inter + 1e-8)


def approx_iou(b1, b2, kind="sph", is_aligned=False):
    """sph_iou_api.py:130-177 (sph_iou / fov_iou wrappers), BFoV only."""
    assert kind in ("sph", "fov")
    R, C = b1.size(0), b2.size(0)
    if R * C == 0:
        return b1.new_zeros((R, 1)) if is_aligned else b1.new_zeros((R, C))
    e1, e2 = _expand(b1, b2, is_aligned)
    j1, j2 = jitter_spherical(e1, e2)
    out = _approx_iou(j1, j2, kind).clamp(0, 1)
    return out if is_aligned else out.view(R, C)


# --------------------------------------------------------------------------- #
# Sph2Pob loss (sph2pob_transform.py:24-35 + sph2pob_iou_loss.py:25-58,104-196)
# --------------------------------------------------------------------------- #
def obb_to_hbb(o):
    """sphdet/bbox/box_formator.py obb2hbb_xyxy: enclosing axis-aligned box of an OBB."""
    c, s = torch.cos(o[:, 4]).abs(), torch.sin(o[:, 4]).abs()
    w = o[:, 2] * c + o[:, 3] * s
    h = o[:, 2] * s + o[:, 3] * c
    return torch.stack([o[:, 0] - w / 2, o[:, 1] - h / 2, o[:, 0] + w / 2, o[:, 1] + h / 2], dim=1)


def loss_obbs(pred, target):
    """sph2pob_transform.py:26-30: jitter -> standard transform -> jitter."""
    j1, j2 = jitter_spherical(pred, target)
    o1, o2 = sph2pob_standard(j1, j2)
    return jitter_rotated(o1, o2)


def sph2pob_iou_loss_elementwise(pred, target, mode="iou", eps=1e-6):
    """sph2pob_iou_loss.py:104-196 on top of the Sph2Pob transform; per-row loss."""
    o1, o2 = loss_obbs(pred, target)
    ious = rotated_iou(o1, o2).clamp(0, 1)
    if mode == "iou":
        return 1 - ious
    h1, h2 = obb_to_hbb(o1), obb_to_hbb(o2)
    enc = (torch.max(h1[:, 2:], h2[:, 2:]) - torch.min(h1[:, :2], h2[:, :2])).clamp(min=0)
    if mode == "giou":
        iw = (torch.min(h1[:, 2:], h2[:, 2:]) - torch.max(h1[:, :2], h2[:, :2])).clamp(min=0)
        a_enc = enc[:, 0] * enc[:, 1]
        union = o1[:, 2] * o1[:, 3] + o2[:, 2] * o2[:, 3] - iw[:, 0] * iw[:, 1]
        return 1 - (ious - ((a_enc - union) / (a_enc + eps)).clamp(0, 1))
    c2 = enc[:, 0] ** 2 + enc[:, 1] ** 2 + eps
    rho2 = (o2[:, 0] - o1[:, 0]) ** 2 + (o2[:, 1] - o1[:, 1]) ** 2
    if mode == "diou":
        return 1 - (ious - (rho2 / c2).clamp(0, 1))
    v = 4 / math.pi ** 2 * (torch.atan(o2[:, 2] / (o2[:, 3] + eps)) - torch.atan(o1[:, 2] / (o1[:, 3] + eps))) ** 2
    with torch.no_grad():
        alpha = (ious > 0.5).to(v.dtype) * v / (1 - ious + v + eps)
    if mode == "ciou":
        return 1 - (ious - ((rho2 / c2).clamp(0, 1) + alpha * v))
    raise NotImplementedError(mode)


def sph_iou_loss_legacy_elementwise(pred, target, mode="log", eps=1e-6):
    """SphIoULossLegacy (sph2pob_iou_loss.py:199-216): mmrotate 0.3.2 rotated_iou_loss on the Sph2Pob OBBs; per-row loss."""
    o1, o2 = loss_obbs(pred, target)
    ious = rotated_iou(o1, o2).clamp(min=eps)
    return {"linear": 1 - ious, "square": 1 - ious ** 2, "log": -ious.log()}[mode]


def sph2pob_iou_loss(pred, target, weight=None, avg_factor=None, mode="iou", eps=1e-6,
                     reduction="mean", loss_weight=1.0):
    """Sph2PobIoULoss.forward: sph2pob_transform.py:24-35, sph2pob_iou_loss.py:25-58,
    mmdet/models/losses/utils.py weight_reduce_loss."""
    if weight is not None and weight.dim() > 1 and target.size(-1) == 4:
        weight = torch.cat([weight, weight.mean(-1, keepdim=True)], dim=-1)
    if weight is not None and not torch.any(weight > 0):
        o1, _ = loss_obbs(pred, target)
        w = weight.unsqueeze(1) if o1.dim() == weight.dim() + 1 else weight
        return (o1 * w).sum()
    if weight is not None and weight.dim() > 1:
        weight = weight.mean(-1)
    loss = sph2pob_iou_loss_elementwise(pred, target, mode=mode, eps=eps)
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        loss = loss.mean() if reduction == "mean" else loss.sum() if reduction == "sum" else loss
    elif reduction == "mean":
        loss = loss.sum() / (avg_factor + torch.finfo(torch.float32).eps)
    elif reduction != "none":
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss_weight * loss


# --------------------------------------------------------------------------- #
# naive_iou (sph_iou_api.py:181-198)
# --------------------------------------------------------------------------- #
def mmcv_bbox_overlaps(b1, b2, mode="iou", aligned=False, offset=0):
    """mmcv-full 1.6.0 ``bbox_overlaps`` (mmcv/ops/csrc/common/cuda/bbox_overlaps_cuda_kernel.cuh; the package is not
    in the reference tree -- parity UNPINNED at that boundary, published kernel restated): xyxy boxes,
    inter / max(area1 + area2 - inter, offset) ('iou') or inter / max(area1, offset) ('iof')."""
    if not aligned:
        R, C = b1.size(0), b2.size(0)
        return mmcv_bbox_overlaps(b1.repeat_interleave(C, 0), b2.repeat(R, 1), mode, True, offset).view(R, C)
    a1 = (b1[:, 2] - b1[:, 0] + offset) * (b1[:, 3] - b1[:, 1] + offset)
    a2 = (b2[:, 2] - b2[:, 0] + offset) * (b2[:, 3] - b2[:, 1] + offset)
    w = (torch.min(b1[:, 2], b2[:, 2]) - torch.max(b1[:, 0], b2[:, 0]) + offset).clamp(min=0)
    h = (torch.min(b1[:, 3], b2[:, 3]) - torch.max(b1[:, 1], b2[:, 1]) + offset).clamp(min=0)
    inter = w * h
    base = (a1 + a2 - inter) if mode == "iou" else a1
    return inter / base.clamp(min=offset)


def mmcv_nms(boxes, scores, iou_threshold, offset=0, score_threshold=0, max_num=-1):
    """mmcv-full 1.6.0 ``nms`` (mmcv/ops/nms.py: NMSop.forward + nms; kernel nms_cuda_kernel.cuh -- parity UNPINNED at that
    boundary, published algorithm restated): drop scores <= score_threshold, visit by descending score, suppress a box iff
    inter / (a1 + a2 - inter) > iou_threshold against a kept one.  Returns (dets [k, 5] score-descending, keep [k])."""
    index = None
    if score_threshold > 0:
        index = (scores > score_threshold).nonzero(as_tuple=False).view(-1)
        boxes, scores = boxes[index], scores[index]
    order = torch.argsort(scores, descending=True, stable=True)
    b = boxes[order]
    area = (b[:, 2] - b[:, 0] + offset) * (b[:, 3] - b[:, 1] + offset)
    alive = torch.ones(b.size(0), dtype=torch.bool)
    kept = []
    for i in range(b.size(0)):
        if not alive[i]:
            continue
        kept.append(i)
        w = (torch.min(b[i, 2], b[i + 1:, 2]) - torch.max(b[i, 0], b[i + 1:, 0]) + offset).clamp(min=0)
        h = (torch.min(b[i, 3], b[i + 1:, 3]) - torch.max(b[i, 1], b[i + 1:, 1]) + offset).clamp(min=0)
        inter = w * h
        alive[i + 1:] &= ~(inter / (area[i] + area[i + 1:] - inter) > iou_threshold)
    keep = order[torch.tensor(kept, dtype=torch.long)]
    if max_num > 0:
        keep = keep[:max_num]
    dets = torch.cat([boxes[keep], scores[keep, None]], dim=1)
    return dets, (keep if index is None else index[keep])


def mmcv_batched_nms(boxes, scores, idxs, nms_cfg, class_agnostic=False):
    """mmcv-full 1.6.0 ``batched_nms`` (mmcv/ops/nms.py) for nms_cfg type 'nms' below split_thr: classes are separated by
    shifting every box by label * (max coordinate + 1)."""
    if nms_cfg is None:
        scores, inds = scores.sort(descending=True)
        return torch.cat([boxes[inds], scores[:, None]], -1), inds
    cfg = dict(nms_cfg)
    class_agnostic = cfg.pop("class_agnostic", class_agnostic)
    if class_agnostic:
        boxes_for_nms = boxes
    else:
        boxes_for_nms = boxes + (idxs.to(boxes) * (boxes.max() + 1))[:, None]
    assert cfg.pop("type", "nms") == "nms"
    assert boxes.size(0) < cfg.pop("split_thr", 10000), "the per-class loop of mmcv above split_thr is not restated"
    dets, keep = mmcv_nms(boxes_for_nms, scores, **cfg)
    return torch.cat([boxes[keep], dets[:, -1:]], -1), keep


def planar_nms(boxes, scores, idxs, nms_cfg, class_agnostic=True):
    """PlanarNMS('sph2pix') (sphdet/bbox/nms/planar_nms.py:7-18)."""
    dets_, keep = mmcv_batched_nms(sph2pix(boxes), scores, idxs, nms_cfg, class_agnostic)
    return torch.cat([boxes[keep], dets_[:, -1:]], dim=-1), keep


def sph2pix(boxes, img_size=(512, 1024)):
    """Sph2PlanarBoxTransform('sph2pix') (box_formator.py:79-87,176-193): xyxy for BFoV, (x, y, w, h, -gamma rad) for RBFoV."""
    img_h, img_w = img_size
    x, y = (boxes[:, 0] / 360) * img_w, (boxes[:, 1] / 180) * img_h
    w, h = (boxes[:, 2] / 360) * img_w, (boxes[:, 3] / 180) * img_h
    if boxes.size(1) == 4:
        return torch.stack([x - w / 2, y - h / 2, x + w / 2, y + h / 2], dim=1)
    return torch.stack([x, y, w, h, -torch.deg2rad(boxes[:, 4])], dim=1)


def naive_iou(b1, b2, mode="iou", is_aligned=False):
    """sph_iou_api.py:181-198: planar IoU of the sph2pix boxes (mmcv bbox_overlaps / box_iou_rotated), no jitter, no clamp."""
    assert mode == "iou"
    rows, cols = b1.size(0), b2.size(0)
    if rows * cols == 0:
        return b1.new_zeros((rows, 1)) if is_aligned else b1.new_zeros((rows, cols))
    p1, p2 = sph2pix(b1), sph2pix(b2)
    if b1.size(1) == 4:
        return mmcv_bbox_overlaps(p1, p2, mode, is_aligned)
    if not is_aligned:
        return rotated_iou(p1.repeat_interleave(cols, 0), p2.repeat(rows, 1), mode).view(rows, cols)
    return rotated_iou(p1, p2, mode)


# --------------------------------------------------------------------------- #
# unbiased_iou (sph_iou_api.py:103-125; unbiased_iou_bfov.py:10-204, unbiased_iou_rbfov.py:4-181)
# --------------------------------------------------------------------------- #
def _unbiased_normals(box, rotated):
    """getNormal: inward unit normals (left, right, up, down) of the four boundary circles, [4, n, 3]; RBFoV: rotated
    about the view axis by gamma (roll_T)."""
    theta, phi, fx, fy = box[:, 0:1], box[:, 1:2], box[:, 2:3] / 2, box[:, 3:4] / 2
    look = torch.cat([torch.sin(phi) * torch.cos(theta), torch.sin(phi) * torch.sin(theta), torch.cos(phi)], dim=1)
    right = torch.cat([-torch.sin(theta), torch.cos(theta), torch.zeros_like(theta)], dim=1)
    up = torch.cat([-torch.cos(phi) * torch.cos(theta), -torch.cos(phi) * torch.sin(theta), torch.sin(phi)], dim=1)
    N = [-torch.cos(fx) * right + torch.sin(fx) * look, torch.cos(fx) * right + torch.sin(fx) * look,
         -torch.cos(fy) * up + torch.sin(fy) * look, torch.cos(fy) * up + torch.sin(fy) * look]
    if rotated:
        g = box[:, 4]
        c, s_, k = torch.cos(g), torch.sin(g), 1 - torch.cos(g)
        nx, ny, nz = look[:, 0], look[:, 1], look[:, 2]
        M = torch.stack([torch.stack([nx * nx * k + c, nx * ny * k - nz * s_, nx * nz * k + ny * s_], -1),
                         torch.stack([nx * ny * k + nz * s_, ny * ny * k + c, ny * nz * k - nx * s_], -1),
                         torch.stack([nx * nz * k - ny * s_, ny * nz * k + nx * s_, nz * nz * k + c], -1)], -2)
        N = [torch.einsum("nij,nj->ni", M, v) for v in N]
    return torch.stack(N)


def unbiased_iou(b1, b2, is_aligned=False):
    """Restatement of the reference's Unbiased-IoU pipeline (vectorised; float64 inside, float32 out, like the numpy
    classes fed with double boxes): jitter, 40 candidate vertices, round(V . N, 8) >= 0 for the eight normals, area from
    the interior angles of the counted vertices (duplicates are not merged: the DFS clean-up is disabled in the reference),
    the two different closing formulas of the BFoV / RBFoV files, clamp."""
    rows, cols = b1.size(0), b2.size(0)
    if rows * cols == 0:
        return b1.new_zeros((rows, 1)) if is_aligned else b1.new_zeros((rows, cols))
    D = b1.size(1)
    if not is_aligned:
        b1, b2 = b1.repeat_interleave(cols, 0), b2.repeat(rows, 1)
    j1, j2 = jitter_spherical(b1.double().clone(), b2.double().clone())
    j1, j2 = torch.deg2rad(j1), torch.deg2rad(j2)
    N1, N2 = _unbiased_normals(j1, D == 5), _unbiased_normals(j2, D == 5)
    N = torch.cat([N1, N2])                                       # [8, n, 3]
    cand, e0, e1 = [], [], []
    for Nb in (N1, N2):
        for a, b in ((0, 2), (3, 0), (2, 1), (1, 3)):             # left x up, down x left, up x right, right x down
            c = torch.cross(Nb[a], Nb[b], dim=-1)
            cand.append(c / c.norm(dim=1, keepdim=True)); e0.append(Nb[a]); e1.append(Nb[b])
    for i in range(4):
        for j in range(4):
            c = torch.cross(N1[i], N2[j], dim=-1)
            v = c / (c.norm(dim=1, keepdim=True) + 1e-10)
            cand += [v, -v]; e0 += [N1[i], N2[j]]; e1 += [N2[j], N1[i]]
    V, E0, E1 = torch.stack(cand), torch.stack(e0), torch.stack(e1)          # [40, n, 3]
    dots = torch.einsum("vni,kni->vnk", V, N)
    valid = (torch.round(dots * 1e8) >= 0).all(dim=2)                        # [40, n]
    ang = torch.arccos((-(E0 * E1).sum(-1)).clamp(-1, 1))
    count = valid.sum(0)
    inter = (ang * valid).sum(0) - (count - 2).double() * math.pi       # (int tensor * float would round pi to float32)
    inter = torch.where(count == 0, torch.zeros_like(inter), inter)

    def area(fx, fy):
        return 4 * torch.arccos(-torch.sin(fx / 2) * torch.sin(fy / 2)) - 2 * math.pi
    a1, a2 = area(j1[:, 2], j1[:, 3]), area(j2[:, 2], j2[:, 3])
    iou = (inter + 1e-8) / (a1 + a2 - (inter + 1e-8)) if D == 4 else inter / (a1 + a2 - inter + 1e-8)
    iou = iou.float().clamp(0, 1)
    return iou if is_aligned else iou.view(rows, cols)


# ---- the other Sph2Pob losses --------------------------------------------------------------------
def _decorated(pred, target, weight):
    """Sph2PobTransfrom.new_forward (sph2pob_transform.py:24-35): OBBs of the pair + the widened BFoV weight."""
    if weight is not None and weight.dim() > 1 and target.size(-1) == 4:
        weight = torch.cat([weight, weight.mean(-1, keepdim=True)], dim=-1)
    o1, o2 = loss_obbs(pred, target)
    return o1, o2, weight


def sph2pob_gd_loss(pred, target, weight=None, avg_factor=None, reduction_override=None, loss_type="gwd", fun="log1p",
                    tau=0.0, alpha=1.0, reduction="mean", loss_weight=1.0, **kwargs):
    """Sph2PobGDLoss (sph2pob_gd_loss.py:7-26) = decorator + mmrotate 0.3.2 GDLoss (oracle/mmrotate_losses.py)."""
    from mmrotate_losses import GDLoss
    o1, o2, weight = _decorated(pred, target, weight)
    return GDLoss(loss_type, fun=fun, tau=tau, alpha=alpha, reduction=reduction, loss_weight=loss_weight, **kwargs)(
        o1, o2, weight, avg_factor=avg_factor, reduction_override=reduction_override)


def sph2pob_kf_loss(pred, target, weight=None, avg_factor=None, reduction_override=None, fun="none", reduction="mean",
                    loss_weight=1.0):
    """Sph2PobKFLoss (sph2pob_kf_loss.py:8-26): mmrotate KFLoss with pred_decode=target OBB, targets_decode=pred OBB."""
    from mmrotate_losses import KFLoss
    o1, o2, weight = _decorated(pred, target, weight)
    return KFLoss(fun=fun, reduction=reduction, loss_weight=loss_weight)(
        o1, o2, weight, avg_factor=avg_factor, pred_decode=o2, targets_decode=o1, reduction_override=reduction_override)


def _obb_bbox2delta(proposals, gt, angle_modifier="original", eps=1e-7):
    """sph2pob_l1_loss.py:40-87 (means 0, stds 1; forced to float32 at :68-69)."""
    proposals, gt = proposals.float(), gt.float()
    px, py, pw, ph, pa = proposals.unbind(dim=-1)
    gx, gy, gw, gh, ga = gt.unbind(dim=-1)
    pw, ph = pw.clip(min=eps), ph.clip(min=eps)
    gw, gh = gw.clip(min=eps), gh.clip(min=eps)
    wrap = (lambda a: a) if angle_modifier == "original" else (lambda a: (a + torch.pi) % torch.pi)
    return torch.stack([(gx - px) / pw, (gy - py) / ph, torch.log(gw / pw), torch.log(gh / ph),
                        (wrap(ga) - wrap(pa)) / torch.pi], dim=-1)


def sph2pob_l1_loss(pred, target, weight=None, avg_factor=None, reduction_override=None, encode=True, swap=False,
                    angle_modifier="original", reduction="mean", loss_weight=1.0):
    """Sph2PobL1Loss (sph2pob_l1_loss.py:9-38) on mmdet's L1Loss (mmdet/models/losses/smooth_l1_loss.py:36-52,107-146)."""
    o1, o2, weight = _decorated(pred, target, weight)
    if encode:
        o1 = _obb_bbox2delta(o2, o1, angle_modifier) if swap else _obb_bbox2delta(o1, o2, angle_modifier)
        o2 = torch.zeros_like(o2)
    red = reduction_override if reduction_override else reduction
    if o2.numel() == 0:
        return loss_weight * (o1.sum() * 0)
    loss = torch.abs(o1 - o2)
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        loss = loss.mean() if red == "mean" else loss.sum() if red == "sum" else loss
    elif red == "mean":
        loss = loss.sum() / (avg_factor + torch.finfo(torch.float32).eps)
    elif red != "none":
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss_weight * loss


# ---- bbox coders (sphdet/bbox/coder/delta_xywh_sph_bbox_coder.py:117-262, delta_xywha_rsph_bbox_coder.py:117-268)
def bbox2delta(proposals, gt, means=None, stds=None):
    """(d_theta, d_phi, d_alpha, d_beta[, d_gamma]) of gt w.r.t. proposals: centre offsets in units of the proposal
    size, log size ratios, gamma difference in radians; then (x - mean) / std.  Sizes clipped at 1e-7.
    The reference casts both inputs to float32 (:139-140), whatever comes in."""
    D = proposals.size(-1)
    eps = 1e-7
    proposals, gt = proposals.float(), gt.float()
    px, py = proposals[..., 0], proposals[..., 1]
    pw, ph = proposals[..., 2].clamp(min=eps), proposals[..., 3].clamp(min=eps)
    gx, gy = gt[..., 0], gt[..., 1]
    gw, gh = gt[..., 2].clamp(min=eps), gt[..., 3].clamp(min=eps)
    cols = [(gx - px) / pw, (gy - py) / ph, torch.log(gw / pw), torch.log(gh / ph)]
    if D == 5:
        cols.append(torch.deg2rad(gt[..., 4] - proposals[..., 4]))
    deltas = torch.stack(cols, dim=-1)
    means = deltas.new_tensor(means if means is not None else [0.0] * D)
    stds = deltas.new_tensor(stds if stds is not None else [1.0] * D)
    return (deltas - means) / stds


def delta2bbox(rois, deltas, means=None, stds=None, wh_ratio_clip=16 / 1000, clip_border=True, add_ctr_clamp=False,
               ctr_clamp=32):
    """Inverse of bbox2delta with the reference's clamps: |d_size| <= |log(wh_ratio_clip)| (only from above with
    add_ctr_clamp, which also clamps the centre shift to +-ctr_clamp), then -- clip_border -- theta into
    [1e-7, 360 - 1e-7], phi / alpha / beta into [1e-7, 180 - 1e-7], gamma into [-90 + 1e-7, 90 - 1e-7]."""
    D = rois.size(-1)
    eps = 1e-7
    means = deltas.new_tensor(means if means is not None else [0.0] * D)
    stds = deltas.new_tensor(stds if stds is not None else [1.0] * D)
    d = deltas * stds + means
    dxy_wh = rois[:, 2:4] * d[:, :2]
    max_ratio = abs(math.log(wh_ratio_clip))
    if add_ctr_clamp:
        dxy_wh = dxy_wh.clamp(min=-ctr_clamp, max=ctr_clamp)
        dwh = d[:, 2:4].clamp(max=max_ratio)
    else:
        dwh = d[:, 2:4].clamp(min=-max_ratio, max=max_ratio)
    gxy = rois[:, :2] + dxy_wh
    gwh = rois[:, 2:4] * dwh.exp()
    cols = [gxy, gwh]
    if D == 5:
        cols.append((rois[:, 4] + torch.rad2deg(d[:, 4]))[:, None])
    b = torch.cat(cols, dim=-1)
    if clip_border:
        lo = [eps, eps, eps, eps, -90 + eps][:D]
        hi = [360 - eps, 180 - eps, 180 - eps, 180 - eps, 90 - eps][:D]
        b = torch.stack([b[:, k].clamp(min=lo[k], max=hi[k]) for k in range(D)], dim=-1)
    return b


# ---- the anchor-free heads' coder (sphdet/bbox/coder/distance_point_sph_bbox_coder.py:72-162) ----------------------------
def distance2bbox(points, distance, max_shape=None, img_shape=(512, 1024)):
    """:72-127 (two-dimensional input): pixel box from the point and the four distances, optional border clamp, then
    xyxy2xywh (box_formator.py:17-23) and _pix2sph_box_transform (:85-92); a fifth column is handed through."""
    x1, y1 = points[:, 0] - distance[:, 0], points[:, 1] - distance[:, 1]
    x2, y2 = points[:, 0] + distance[:, 2], points[:, 1] + distance[:, 3]
    if max_shape is not None:
        x1, x2 = x1.clamp(min=0, max=max_shape[1]), x2.clamp(min=0, max=max_shape[1])
        y1, y2 = y1.clamp(min=0, max=max_shape[0]), y2.clamp(min=0, max=max_shape[0])
    img_h, img_w = img_shape
    x, y, w, h = (x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1
    cols = [(x / img_w) * 360, (y / img_h) * 180, (w / img_w) * 360, (h / img_h) * 180]
    if distance.size(-1) == 5:
        cols.append(distance[:, 4])
    return torch.stack(cols, dim=-1)


def bbox2distance(points, bbox, max_dis=None, eps=0.1, img_shape=(512, 1024)):
    """:130-162: the box's sph2pix corners (sph2pix above) measured from the point, clamped to [0, max_dis - eps]."""
    xyxy = sph2pix(bbox[:, :4], img_shape)
    cols = [points[:, 0] - xyxy[:, 0], points[:, 1] - xyxy[:, 1], xyxy[:, 2] - points[:, 0], xyxy[:, 3] - points[:, 1]]
    if max_dis is not None:
        cols = [c.clamp(min=0, max=max_dis - eps) for c in cols]
    if bbox.size(-1) == 5:
        cols.append(bbox[:, 4])
    return torch.stack(cols, dim=-1)


def decode_iou_loss(anchors, deltas, target, weight=None, avg_factor=None, mode="iou", reduction="mean", loss_weight=1.0,
                    **coder):
    """What the head does with reg_decoded_bbox=True (sphdet/models/heads/sph_retina_head.py:252-265):
    bbox_coder.decode(anchors, bbox_pred) -> Sph2PobIoULoss(pred, target, weight, avg_factor)."""
    return sph2pob_iou_loss(delta2bbox(anchors, deltas, **coder), target, weight=weight, avg_factor=avg_factor, mode=mode,
                            reduction=reduction, loss_weight=loss_weight)


def get_bboxes_single(cls_score_list, bbox_pred_list, mlvl_priors, cfg, box_version=4, **coder):
    """SphRetinaHead._get_bboxes_single + _bbox_post_process for one image (sph_retina_head.py:35-212) with sigmoid
    classification, mmdet's filter_scores_and_topk (mmdet/core/utils/misc.py:143-152), the coder restatement above
    and the greedy NMS below.  Returns (det_bboxes [K, D + 1], det_labels [K])."""
    bbs, scs, lbs = [], [], []
    for cls_score, bbox_pred, priors in zip(cls_score_list, bbox_pred_list, mlvl_priors):
        bbox_pred = bbox_pred.permute(1, 2, 0).reshape(-1, box_version)
        num_cls = cls_score.size(0) * cls_score.size(1) * cls_score.size(2) // bbox_pred.size(0)
        scores = cls_score.permute(1, 2, 0).reshape(-1, num_cls).sigmoid()
        valid = scores > cfg["score_thr"]
        vs, vi = scores[valid], torch.nonzero(valid)
        k = min(cfg["nms_pre"], vi.size(0))
        vs, order = vs.sort(descending=True)
        keep_idxs, labels = vi[order[:k]].unbind(dim=1)
        bbs.append(delta2bbox(priors[keep_idxs], bbox_pred[keep_idxs], **coder))
        scs.append(vs[:k])
        lbs.append(labels)
    bboxes, scores, labels = torch.cat(bbs), torch.cat(scs), torch.cat(lbs)
    if bboxes.numel() == 0:
        return torch.cat([bboxes, scores[:, None]], -1), labels
    dets, keep = nms_batched(bboxes, scores, labels, cfg["nms"]["iou_threshold"])
    return dets[:cfg["max_per_img"]], labels[keep][:cfg["max_per_img"]]


# --------------------------------------------------------------------------- #
# spherical NMS (sphdet/bbox/nms/sph_nms.py:22-74)
# --------------------------------------------------------------------------- #
def nms_single(boxes, scores, thr, iou_fn):
    """sph_nms.py:62-74: greedy; pivot is bboxes1, the survivors are bboxes2; suppress iff IoU > thr."""
    order = torch.argsort(scores, descending=True)
    keep = []
    while order.numel() > 0:
        keep.append(int(order[0]))
        if order.numel() == 1:
            break
        iou = iou_fn(boxes[order[0]][None, :], boxes[order[1:]]).reshape(-1)
        order = order[1:][iou <= thr]
    return torch.tensor(keep, dtype=torch.long)


def nms_batched(boxes, scores, idxs, iou_threshold=0.5, max_num=None, class_agnostic=False, iou_fn=None):
    """sph_nms.py:22-60: per label greedy NMS, union, sort by score, truncate, append score column.
    `class_agnostic` is accepted and IGNORED exactly as in the reference (:33 pops it, :44 loops over
    torch.unique(idxs) regardless)."""
    if iou_fn is None:
        iou_fn = lambda a, b: sph2pob_iou(a, b, "efficient")
    max_num = boxes.size(0) if max_num is None else min(max_num, boxes.size(0))
    kept = torch.zeros(boxes.size(0), dtype=torch.bool)
    labels = idxs
    for lab in torch.unique(labels):
        sel = (labels == lab).nonzero().view(-1)
        kept[sel[nms_single(boxes[sel], scores[sel], iou_threshold, iou_fn)]] = True
    keep = kept.nonzero().view(-1)
    s, o = scores[keep].sort(descending=True)
    keep = keep[o][:max_num]
    return torch.cat([boxes[keep], s[:max_num, None]], dim=-1), keep


# --------------------------------------------------------------------------- #
# MaxIoUAssigner.assign_wrt_overlaps (mmdet/core/bbox/assigners/max_iou_assigner.py:135-220)
# --------------------------------------------------------------------------- #
def multiclass_nms(multi_bboxes, multi_scores, score_thr, iou_threshold, max_num=-1, score_factors=None, box_version=4,
                   nms_max_num=None):
    """sphdet/bbox/nms/utils.py:6-95 (the R-CNN heads' wrapper) restated on top of nms_batched (= SphNMS, sph_nms.py:22-74):
    class-specific boxes [n, C * D] or shared boxes [n, D]; scores [n, C + 1] with the background in the last column;
    candidates are the (proposal, class) pairs with score > score_thr (:55), score factors are applied AFTER that mask
    (:58-63); SphNMS per class; the first max_num detections (:88-90).  Returns (dets, labels, inds into the flattened
    [n * C] (proposal, class) grid)."""
    num_classes = multi_scores.size(1) - 1
    if multi_bboxes.shape[1] > box_version:
        bboxes = multi_bboxes.view(multi_scores.size(0), -1, box_version)
    else:
        bboxes = multi_bboxes[:, None].expand(multi_scores.size(0), num_classes, box_version)
    scores = multi_scores[:, :-1]
    labels = torch.arange(num_classes, dtype=torch.long).view(1, -1).expand_as(scores)
    bboxes, scores, labels = bboxes.reshape(-1, box_version), scores.reshape(-1), labels.reshape(-1)
    valid = scores > score_thr
    if score_factors is not None:
        scores = scores * score_factors.view(-1, 1).expand(multi_scores.size(0), num_classes).reshape(-1)
    inds = valid.nonzero(as_tuple=False).squeeze(1)
    bboxes, scores, labels = bboxes[inds], scores[inds], labels[inds]
    if bboxes.numel() == 0:
        return torch.cat([bboxes, scores[:, None]], -1), labels, inds
    dets, keep = nms_batched(bboxes, scores, labels, iou_threshold, max_num=nms_max_num)
    if max_num > 0:
        dets, keep = dets[:max_num], keep[:max_num]
    return dets, labels[keep], inds[keep]


def assign_wrt_overlaps(overlaps, gt_labels=None, pos_iou_thr=0.5, neg_iou_thr=0.4, min_pos_iou=0.0,
                        gt_max_assign_all=True, match_low_quality=True, lowest_index_ties=False):
    """Literal restatement, Python loop over the GTs included.  Returns (gt_inds, max_overlaps, labels)."""
    num_gts, num_bboxes = overlaps.size(0), overlaps.size(1)
    assigned = overlaps.new_full((num_bboxes,), -1, dtype=torch.long)
    if num_gts == 0 or num_bboxes == 0:
        if num_gts == 0:
            assigned[:] = 0
        labels = None if gt_labels is None else overlaps.new_full((num_bboxes,), -1, dtype=torch.long)
        return assigned, overlaps.new_zeros((num_bboxes,)), labels
    max_overlaps, argmax_overlaps = overlaps.max(dim=0)
    gt_max_overlaps, gt_argmax_overlaps = overlaps.max(dim=1)
    if lowest_index_ties:
        # torch.max documents no tie rule; the kernels resolve ties to the LOWEST index: state that rule explicitly so
        # that the comparison is exact (argmax of the first maximal entry along each axis)
        argmax_overlaps = (overlaps == max_overlaps[None, :]).to(torch.uint8).argmax(dim=0)
        gt_argmax_overlaps = (overlaps == gt_max_overlaps[:, None]).to(torch.uint8).argmax(dim=1)
    if isinstance(neg_iou_thr, float):
        assigned[(max_overlaps >= 0) & (max_overlaps < neg_iou_thr)] = 0
    else:
        assigned[(max_overlaps >= neg_iou_thr[0]) & (max_overlaps < neg_iou_thr[1])] = 0
    pos = max_overlaps >= pos_iou_thr
    assigned[pos] = argmax_overlaps[pos] + 1
    if match_low_quality:
        for i in range(num_gts):
            if gt_max_overlaps[i] >= min_pos_iou:
                if gt_max_assign_all:
                    assigned[overlaps[i, :] == gt_max_overlaps[i]] = i + 1
                else:
                    assigned[gt_argmax_overlaps[i]] = i + 1
    labels = None
    if gt_labels is not None:
        labels = assigned.new_full((num_bboxes,), -1)
        p = assigned > 0
        labels[p] = gt_labels[assigned[p] - 1]
    return assigned, max_overlaps, labels


def get_targets_single(anchors, gt_bboxes, gt_labels, gt_inds, num_classes, reg_decoded_bbox=True, pos_weight=-1,
                       means=None, stds=None):
    """mmdet/models/dense_heads/anchor_head.py:254-285 (_get_targets_single after the assigner, every anchor valid) with the
    PseudoSampler (mmdet/core/bbox/samplers/pseudo_sampler.py:33-39: pos = nonzero(gt_inds > 0), neg = nonzero(gt_inds == 0)).
    Returns (labels, label_weights, bbox_targets, bbox_weights, num_pos, num_neg)."""
    n = anchors.size(0)
    bbox_targets = torch.zeros_like(anchors, dtype=torch.float)
    bbox_weights = torch.zeros_like(anchors)
    labels = anchors.new_full((n,), num_classes, dtype=torch.long)
    label_weights = anchors.new_zeros(n, dtype=torch.float)
    pos_inds = torch.nonzero(gt_inds > 0, as_tuple=False).squeeze(-1).unique()
    neg_inds = torch.nonzero(gt_inds == 0, as_tuple=False).squeeze(-1).unique()
    if len(pos_inds) > 0:
        pos_gt = gt_bboxes[gt_inds[pos_inds] - 1]
        pos_bbox_targets = pos_gt if reg_decoded_bbox else bbox2delta(anchors[pos_inds], pos_gt, means, stds)
        bbox_targets[pos_inds, :] = pos_bbox_targets.float()
        bbox_weights[pos_inds, :] = 1.0
        labels[pos_inds] = 0 if gt_labels is None else gt_labels[gt_inds[pos_inds] - 1]
        label_weights[pos_inds] = 1.0 if pos_weight <= 0 else pos_weight
    if len(neg_inds) > 0:
        label_weights[neg_inds] = 1.0
    return labels, label_weights, bbox_targets, bbox_weights, len(pos_inds), len(neg_inds)


# --------------------------------------------------------------------------- #
# synthetic inputs (tests/utils/generate_data.py:10-42, dtype='float' branch)
# --------------------------------------------------------------------------- #
def generate_boxes(num, theta_range=(0, 360), phi_range=(0, 180), alpha_range=(1, 180),
                   beta_range=(1, 180), gamma_range=(-90, 90), box="bfov", seed=None):
    if seed is not None:
        torch.manual_seed(seed)
    u = torch.rand((num, 5))
    rng = [theta_range, phi_range, alpha_range, beta_range, gamma_range]
    cols = [u[:, k] * (r[1] - r[0]) + r[0] for k, r in enumerate(rng)]
    return torch.stack(cols[:4] if box == "bfov" else cols, dim=1)
