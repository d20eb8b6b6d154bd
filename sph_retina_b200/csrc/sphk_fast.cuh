// sphk_fast.cuh -- the N x M formulation of the Sph2Pob IoU pair: per-box precompute, a conservative
// "cannot touch" prefilter, and a trig-free fast path for the pairs that survive it.
//
// Why: in the reference every pair pays the whole pipeline (sph_iou_api.py:48-86).  In the N x M
// workloads 70-80 % of the pairs are far apart and their IoU is exactly 0, and most of
// the per-pair transcendental work depends on one box only.  So
//   * BoxRec / BoxCull hold, per box and per role, everything that does not depend on the partner, for
//     the case that jitter_1's similarity mask is false (sph_iou_api.py:246-251: the mask is the only
//     pair-dependent part of jitter_1; with it false each box is merely clamped);
//   * pre_disjoint() proves from 7 FMAs that the two planar boxes cannot touch whatever the jitters
//     do (centre distance > sum of circumradii + margin), in which case both rotated-IoU
//     implementations of the reference return exactly 0;
//   * pair_fast() evaluates a surviving pair without atan2/sincos: the internal angles are only ever
//     used through their sine and cosine, which are the normalised (n, m) components themselves.
// Whenever a reference quirk could be active (similarity mask, acos clamp zone, jitter_2, size
// clamps, gamma beyond +-179 deg) pair_fast() declines and the caller runs the reference-order code of
// sphk_math.cuh (sph2pob_iou_pair).  Both paths share riou_core().
#pragma once
#include <string.h>

#include "sphk_math.cuh"

namespace sphk {

// What the fast path needs about one box (valid when jitter_1's mask is false for the pair).
// 64 bytes = 4 x float4.
struct BoxRec {
    float t, p, tj, pj;   // raw theta, phi (deg) for the similarity mask; jittered (clamped) theta, phi for the geometry
    float sp, cp, w, h;   // sin / cos phi; planar edge lengths (rad) before jitter_2
    float sg, cg, a, b;   // sin / cos gamma (0, 1 for BFoV); raw alpha, beta (deg) for the similarity mask
    float g, flag, pad0, pad1;  // raw gamma (deg); flag != 0: reference-order path only
};
// Prefilter operands of one box.  64 bytes = 4 x float4.
//   circle test (pre_disjoint):      centre unit vector u, cos / sin of r = circumradius of the planar box + jitter margin
//   box-frame test (pre_outside_box): the box's own planar axes as 3-D tangent vectors e (width axis) and f (height
//                                     axis) at its centre, its half sizes + margin, and r itself
struct BoxCull {
    float ux, uy, uz, rc;    // centre unit vector; cos(r)
    float rs, bias, r, pad0; // sin(r); bias: -1e-6 (rounding margin) or -10 (box too large to ever be culled); r (rad)
    float ex, ey, ez, hwm;   // width axis; half width + kBoxMargin (rad), 10 = this box never culls by its frame
    float fx, fy, fz, hhm;   // height axis; half height + kBoxMargin
};
constexpr int kBoxRecFloats = 16, kBoxCullFloats = 16;
// What may move between the records and the jittered OBBs the reference clips (sph_iou_api.py:222-260): centres by
// <= 3.5e-4 rad (jitter_2) + 4.5e-4 (acos clamp of the arc), half sizes by <= 1.3e-4, the frame of box 1 by
// <= eps + eps' + the acos clamp of its angle = 1.9e-3 rad, i.e. the partner's centre by <= pi * 1.9e-3 = 6e-3.
constexpr float kBoxMargin = 8e-3f;

// Record of one box in its role (1 = bboxes1 / rows, 2 = bboxes2 / columns); the jittered box is returned too.
SPHK_HD JitBox box_rec(const RawBox& x, int role, int D, int edge, BoxRec* rec) {
    const JitBox j = (role == 1) ? jitter1_role1(x, false, D) : jitter1_role2(x, false, D);
    sincos_deg(j.p_hi, j.p_lo, &rec->sp, &rec->cp);
    rec->t = x.t; rec->p = x.p; rec->a = x.a; rec->b = x.b; rec->g = (D == 5) ? x.g : 0.0f;
    rec->tj = j.t_hi + j.t_lo; rec->pj = j.p_hi + j.p_lo;   // exact unless clamped at the upper end (flagged below)
    rec->pad0 = 0.0f; rec->pad1 = 0.0f;
    rec->w = edge_len_deg(j.a, edge);
    rec->h = edge_len_deg(j.b, edge);
    rec->sg = 0.0f; rec->cg = 1.0f;
    // slow-path-only boxes: theta or phi clamped at the UPPER end (360 - eps is not a float: the hi + lo
    // pair of the reference-order path is needed; a clamp at the lower end gives tj = the offset,
    // exactly), sizes in reach of jitter_2's minimum, gamma in reach of jitter_2's +-2 pi clamp, NaNs.
    // A clamped alpha / beta is fine: a, b stay raw, w, h come from the clamped values.
    bool slow = (j.t_hi != x.t && j.t_hi != 0.0f) || (j.p_hi != x.p && j.p_hi != 0.0f) || !(x.t == x.t) || !(x.p == x.p);
    slow = slow || !(rec->w >= 2.0f * kMinWh1) || !(rec->h >= 2.0f * kMinWh1);
    if (D == 5) {
        sincos_deg(j.g, 0.0f, &rec->sg, &rec->cg);
        slow = slow || !(fabsf(x.g) <= 179.0f);
    }
    rec->flag = slow ? 1.0f : 0.0f;
    return j;
}

// circumradius of the planar box plus the jitter margin: jitter_1 moves w, h by <= 4.3e-6 rad, jitter_2 by
// <= 2.5e-4 rad and the centre distance by <= 1.3e-4 rad: 4e-4 per box covers all of it
SPHK_HD float cull_radius(float w, float h) { return fmaf(0.5f * sqrtf(fmaf(w, w, h * h)), 1.0001f, 4e-4f); }

SPHK_HD void box_pre(const RawBox& x, int role, int D, int edge, BoxRec* rec, BoxCull* cull) {
    const JitBox j = box_rec(x, role, D, edge, rec);
    float st, ct;
    sincos_deg(j.t_hi, j.t_lo, &st, &ct);
    cull->ux = rec->sp * ct; cull->uy = rec->sp * st; cull->uz = rec->cp;
    const float r = cull_radius(rec->w, rec->h);
    if (r < 1.55f) {               // r_g + r_p < pi: cos is monotone over [0, r_g + r_p]
        sincos_f(r, &cull->rs, &cull->rc);
        cull->bias = -1e-6f;
    } else {
        cull->rs = 0.0f; cull->rc = 0.0f; cull->bias = -10.0f;
    }
    cull->r = (r == r) ? r : 10.0f;
    cull->pad0 = 0.0f;
    // The planar box's axes in 3-D.  With east = (-sin t, cos t, 0) and south = (cos p cos t, cos p sin t, -sin p) at the
    // centre, the Sph2Pob frame of the box (internal angle minus gamma, sph2pob_efficient.py:53-57) has its width
    // axis along cos(g) east - sin(g) south and its height axis along sin(g) east + cos(g) south: the partner's
    // centre sits at arc * (cos, sin)(bearing) in that frame, and u_partner . axis = sin(arc) * (cos, sin)(bearing).
    const float sx = rec->cp * ct, sy = rec->cp * st, sz = -rec->sp;
    cull->ex = fmaf(rec->cg, -st, -rec->sg * sx); cull->ey = fmaf(rec->cg, ct, -rec->sg * sy); cull->ez = -rec->sg * sz;
    cull->fx = fmaf(rec->sg, -st, rec->cg * sx);  cull->fy = fmaf(rec->sg, ct, rec->cg * sy);  cull->fz = rec->cg * sz;
    const bool frame_ok = rec->flag == 0.0f;       // reference-order-only boxes keep the circle test alone
    cull->hwm = frame_ok ? fmaf(0.5f, rec->w, kBoxMargin) : 10.0f;
    cull->hhm = frame_ok ? fmaf(0.5f, rec->h, kBoxMargin) : 10.0f;
}

// true: the planar boxes are disjoint for every jitter outcome (centre distance > sum of the circumradii
// + margins) -> IoU is exactly 0 in the reference.  cos(arc) < cos(r_g + r_p), 7 FMA-class instructions.
SPHK_HD bool pre_disjoint(const BoxCull& g, const BoxCull& p) {
    const float dot = fmaf(g.ux, p.ux, fmaf(g.uy, p.uy, g.uz * p.uz));
    const float thr = fmaf(g.rc, p.rc, fmaf(-g.rs, p.rs, g.bias + p.bias));
    return dot < thr;
}

// true: the centre of box p lies outside box g grown by r_p (+ margin) along one of g's own axes -> the planar boxes
// cannot touch -> IoU is exactly 0 in the reference.  In g's planar frame p sits at arc * (c, s) with (c, s) the unit
// bearing; a = u_p . axis_g = sin(arc) * c.  For T = half size + margin + r_p <= pi/2: sin is concave on [0, pi], so
// sin(arc |c|) >= sin(arc) |c| = |a|, hence arc |c| >= asin|a| >= |a| (1 + a^2 / 6); if that lower bound exceeds T the
// centre is farther than T along the axis.  For T > pi/2 the bound (<= 7/6) can never exceed it: no special case.
// 6 FMA + 2 x 4 instructions; half sizes of 10 (flagged boxes) or r = 10 (NaN sizes) switch the test off.
SPHK_HD bool pre_outside_box(float ex, float ey, float ez, float hwm, float fx, float fy, float fz, float hhm,
                             float pux, float puy, float puz, float pr) {
    const float a = fmaf(ex, pux, fmaf(ey, puy, ez * puz));
    const float b = fmaf(fx, pux, fmaf(fy, puy, fz * puz));
    const float la = fabsf(a) * fmaf(a * a, 0.16666667f, 1.0f);
    const float lb = fabsf(b) * fmaf(b * b, 0.16666667f, 1.0f);
    return (la > hwm + pr) | (lb > hhm + pr);
}
SPHK_HD bool pre_outside_box(const BoxCull& g, const BoxCull& p) {
    return pre_outside_box(g.ex, g.ey, g.ez, g.hwm, g.fx, g.fy, g.fz, g.hhm, p.ux, p.uy, p.uz, p.r);
}

// Separating-axis test on the two planar boxes, from the per-box records alone (no arc, no atan2, no sincos).
// Sph2Pob lays the boxes out along the great circle through their centres: in the frame of box g the centre of p sits
// at arc * (cos a1, -sin a1), and p is turned by r = a2 - a1 against g, with a1 / a2 the bearings of the partner seen
// from g / p in the boxes' own frames (internal angle minus gamma).  With the frame axes e (width) and f (height) of the
// records as 3-D tangent vectors,
//     A = u_p . e_g = -S cos a1,   B = u_p . f_g = S sin a1,   C = u_g . e_p = S cos a2,   Dd = u_g . f_p = -S sin a2,
// S = sin(arc): |cos r| = |C A + Dd B| / S^2, |sin r| = |Dd A - C B| / S^2, and the four candidate axes are the boxes'
// own: along e_g the centres are arc |A| / S apart and the boxes reach hw_g + hw_p |cos r| + hh_p |sin r|, and so on.
// The arc enters through a series LOWER bound from cos(arc) = u_g . u_p (arc = 2 asin(x), x^2 = (1 - cos) / 2,
// asin(x) >= x (1 + x^2 / 6 + 3 x^4 / 40)), so a reported gap is a real one.  Margins: the half sizes in the records carry
// kBoxMargin, kSatMargin more covers what jitter_1 / jitter_2 and the acos clamps can turn both frames by (2 x 1.9e-3 rad
// x (hw + hh <= pi) of extent, pi x 1.9e-3 of centre offset).  Never fires for S < 0.01 (near-coincident or antipodal
// centres: the bearings are ill-conditioned there), for flagged boxes (half sizes 10) or NaNs.
// true => the planar boxes are disjoint for every jitter outcome => IoU is exactly 0 in the reference.  ~60 instructions,
// run by the N x M kernels on the pairs the circle test leaves alive when the operands are of similar size (1 M x 1,024
// random boxes: a third of them end here -- 40.5 % -> 27.0 % of all pairs go on to the clipper, 26.2 % overlap).
constexpr float kSatMargin = 1.2e-2f;
SPHK_HD float rcp_approx(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
SPHK_HD bool pre_sat_disjoint(float gux, float guy, float guz, float gex, float gey, float gez, float ghw, float gfx, float gfy,
                              float gfz, float ghh, float pux, float puy, float puz, float pex, float pey, float pez, float phw,
                              float pfx, float pfy, float pfz, float phh) {
    const float dot = fmaf(gux, pux, fmaf(guy, puy, guz * puz));
    const float A = fmaf(gex, pux, fmaf(gey, puy, gez * puz)), B = fmaf(gfx, pux, fmaf(gfy, puy, gfz * puz));
    const float C = fmaf(pex, gux, fmaf(pey, guy, pez * guz)), Dd = fmaf(pfx, gux, fmaf(pfy, guy, pfz * guz));
    const float S2 = fmaf(A, A, B * B);
    const float x2 = fmaxf(fmaf(-0.5f, dot, 0.5f), 1e-12f);
    const float x = x2 * rsqrt_f(x2);
    const float arc = (x + x) * fmaf(x2, fmaf(x2, 0.075f, 0.16666667f), 1.0f);
    const float rs2 = rcp_approx(S2);
    const float q = arc * rsqrt_f(S2);                                  // arc / S
    const float X = fabsf(fmaf(C, A, Dd * B)) * rs2, Y = fabsf(fmaf(Dd, A, -C * B)) * rs2;   // |cos r|, |sin r|
    const bool t1 = q * fabsf(A) > fmaf(phw, X, fmaf(phh, Y, ghw + kSatMargin));
    const bool t2 = q * fabsf(B) > fmaf(phw, Y, fmaf(phh, X, ghh + kSatMargin));
    const bool t3 = q * fabsf(C) > fmaf(ghw, X, fmaf(ghh, Y, phw + kSatMargin));
    const bool t4 = q * fabsf(Dd) > fmaf(ghw, Y, fmaf(ghh, X, phh + kSatMargin));
    return (S2 > 1e-4f) & (t1 | t2 | t3 | t4);
}
SPHK_HD bool pre_sat_disjoint(const BoxCull& g, const BoxCull& p) {
    return pre_sat_disjoint(g.ux, g.uy, g.uz, g.ex, g.ey, g.ez, g.hwm, g.fx, g.fy, g.fz, g.hhm, p.ux, p.uy, p.uz, p.ex, p.ey, p.ez,
                            p.hwm, p.fx, p.fy, p.fz, p.hhm);
}

// The clipping job a pair boils down to: box 2 (w2 x h2, centre (px, py), rotated by r) against the
// axis-aligned box 1 (w1 x h1) at the origin.
struct ClipJob {
    float px, py, cr, sr, w1, h1, w2, h2;
};
enum { JOB_READY = 0, JOB_DEAD = 1, JOB_SLOW = 2 };

// sin / cos of the ends of the reference's clamped-acos range (kAcosLo = acos(1 - 1e-7))
constexpr float kSinLo = 4.47213587e-4f, kCosLo = 0.9999999f;

// |angle| confined to [kAcosLo, pi - kAcosLo] on its (sin, cos); the sign of the angle is the sign of the sine,
// zero counting as negative (sph2pob_efficient.py:224-225, sph2pob_standard.py:220-235)
SPHK_HD void clamp_axis(float& s, float& c) {
    if (fabsf(s) < kSinLo) {
        s = (s > 0.0f) ? kSinLo : -kSinLo;
        c = (c < 0.0f) ? -kCosLo : kCosLo;
    }
}

// Common tail of the fast transform.  In: the arc (outside the acos clamp zone), the internal angles
// a_g = atan2(sag, cag), a_p = atan2(sap, cap) as unit vectors, sin / cos of gamma (0, 1 for BFoV) and gamma itself in
// degrees, the planar sizes.  Applies the clamped-acos range of the reference (before gamma in 'efficient', after it
// in 'standard') and jitter_2 (sph_iou_api.py:222-242) in the rotated frame:
//   * its mask can only fire through the sizes or the angles here (|x1 - x2| = arc >= 2e-3 > eps); then box 1 moves by
//     (eps, eps), grows by 2 eps and turns by eps; box 2 moves by (2 eps, 2 eps), grows by eps and turns by 5 eps;
//   * angles matter only when the boxes are nearly parallel (cos r > 0, |sin r| < 2.5e-3 covers eps' + 4 eps): that rare
//     lane recovers the angles themselves (unwrapped, as the reference holds them) and replays jitter_2 on them.
// The size and angle clamps of jitter_2 cannot be active: sizes near the minimum and |gamma| > 179 deg never get here.
SPHK_HD void finish_job(float arc, float sag, float cag, float sap, float cap, float sg1, float cg1, float sg2, float cg2,
                        float g1_deg, float g2_deg, float w1, float h1, float w2, float h2, int D, int kind, ClipJob* job) {
    const bool standard = (kind == KIND_SPH2POB_STANDARD);
    if (!standard) { clamp_axis(sag, cag); clamp_axis(sap, cap); }
    float c1 = cag, s1 = sag, c2 = cap, s2 = sap;
    if (D == 5) {       // a - gamma
        c1 = fmaf(cag, cg1, sag * sg1); s1 = fmaf(sag, cg1, -cag * sg1);
        c2 = fmaf(cap, cg2, sap * sg2); s2 = fmaf(sap, cg2, -cap * sg2);
    }
    if (standard) { clamp_axis(s1, c1); clamp_axis(s2, c2); }
    float cr = fmaf(c2, c1, s2 * s1), sr = fmaf(s2, c1, -c2 * s1);         // r = a2 - a1
    bool m = (fabsf(w1 - w2) < kEps) | (fabsf(h1 - h2) < kEps);
    float dx = arc, dy = 0.0f;
    if (cr > 0.0f && fabsf(sr) < 2.5e-3f) {
        // nearly parallel: the angles themselves, as the reference-order path holds them
        float a1, a2;
        if (standard) {
            a1 = atan2f(s1, c1); a2 = atan2f(s2, c2);                       // wrapped after gamma, then clamped
        } else {
            a1 = atan2f(sag, cag); a2 = atan2f(sap, cap);
            if (D == 5) { a1 -= g1_deg * kDeg2Rad; a2 -= g2_deg * kDeg2Rad; }   // unwrapped (sph2pob_efficient.py:55-57)
        }
        m = m | (fabsf(a1 - a2) < kEps);
        if (m) { a1 += kEps; a2 += kEps5; }
        if (fabsf(a1 - a2) < kEpsA) { a1 += kEpsA; a2 += kEpsA2; }
        sincos_f(a1, &s1, &c1);
        sincos_f(a2 - a1, &sr, &cr);
    } else if (m) {
        // turn box 1 by eps and the relative angle by 4 eps
        const float ce = 0.99999999237f, se = 1.2345678e-4f, c4 = 0.9999998781f, s4 = 4.9382710e-4f;
        const float t1 = fmaf(c1, ce, -s1 * se), u1 = fmaf(s1, ce, c1 * se);
        const float tr = fmaf(cr, c4, -sr * s4), ur = fmaf(sr, c4, cr * s4);
        c1 = t1; s1 = u1; cr = tr; sr = ur;
    }
    if (m) { w1 += kEps2; h1 += kEps2; w2 += kEps; h2 += kEps; dx = arc + kEps; dy = kEps; }
    job->sr = (sr == 0.0f) ? 1e-30f : sr;
    job->cr = (cr == 0.0f) ? 1e-30f : cr;
    // centre of box 2 in the frame of box 1: R(-a1) (dx, dy)   (same for both transforms)
    job->px = fmaf(c1, dx, s1 * dy); job->py = fmaf(-s1, dx, c1 * dy);
    job->w1 = w1; job->h1 = h1; job->w2 = w2; job->h2 = h2;
}

// Transform stage of the fast path.  JOB_READY: *job is set.  JOB_DEAD (only with cull = true): the planar
// boxes cannot touch, IoU is exactly 0.  JOB_SLOW: a reference quirk may be active, run sph2pob_iou_pair.
SPHK_HD int pair_job(const BoxRec& g, const BoxRec& p, int D, int kind, bool cull, ClipJob* job) {
    if (g.flag != 0.0f || p.flag != 0.0f) return JOB_SLOW;
    // jitter_1's similarity mask (sph_iou_api.py:246-247) on the raw coordinates
    bool sim = (fabsf(g.t - p.t) < kEps) | (fabsf(g.p - p.p) < kEps) | (fabsf(g.a - p.a) < kEps) | (fabsf(g.b - p.b) < kEps);
    if (D == 5) sim = sim | (fabsf(g.g - p.g) < kEps);
    if (sim) return JOB_SLOW;
    float hth, sin_dth, hph, sin_dph;       // sin^2(dtheta / 2), sin(dtheta), sin^2(dphi / 2), sin(dphi)
    sin2_and_sin_double_deg(0.5f * (p.tj - g.tj), &hth, &sin_dth);
    sin2_and_sin_double_deg(0.5f * (p.pj - g.pj), &hph, &sin_dph);
    const float hav = fmaf(g.sp * p.sp, hth, hph);
    const float arc = arc_from_hav(hav);
    // outside: the acos clamp zone of the arc and jitter_2's |x1 - x2| < eps
    if (!(arc > 2e-3f && arc < 3.14f)) return JOB_SLOW;
    if (cull && arc > cull_radius(g.w, g.h) + cull_radius(p.w, p.h)) return JOB_DEAD;
    const float ng = fmaf(-2.0f * g.cp * p.sp, hth, sin_dph), mg = -p.sp * sin_dth;
    const float np = fmaf(2.0f * p.cp * g.sp, hth, sin_dph), mp = -g.sp * sin_dth;
    const float S2g = fmaf(ng, ng, mg * mg), S2p = fmaf(np, np, mp * mp);   // both sin^2(arc)
    if (!(S2g > 1e-30f && S2p > 1e-30f)) return JOB_SLOW;
    const float ig = rsqrt_f(S2g), ip = rsqrt_f(S2p);
    // a_g = atan2(ng, mg), a_p = atan2(np, mp): only their sine and cosine are ever used
    finish_job(arc, ng * ig, mg * ig, np * ip, mp * ip, g.sg, g.cg, p.sg, p.cg, g.g, p.g, g.w, g.h, p.w, p.h, D, kind, job);
    return JOB_READY;
}

// Clipping stage: IoU / IoF of a job, with the final clamp(0, 1) of sph_iou_api.py:86.
SPHK_HD float clip_job_iou(const ClipJob& j, int mode) {
    EdgeClip E;
    const float A1 = j.w1 * j.h1, A2 = j.w2 * j.h2;
    float I = riou_core(j.px, j.py, j.cr, j.sr, j.w1, j.h1, j.w2, j.h2, &E);
    I = fminf(fmaxf(I, 0.0f), fminf(A1, A2));
    if (!(I == I)) I = 0.0f;
    const float den = (mode == MODE_IOF) ? A1 : (A1 + A2 - I);
    return clampf(I / den, 0.0f, 1.0f);
}

// Fast path of one pair.  Returns true and sets *out when it applies; false = run the reference-order path.
SPHK_HD bool pair_fast(const BoxRec& g, const BoxRec& p, int D, int kind, int mode, float* out) {
    ClipJob job;
    if (pair_job(g, p, D, kind, false, &job) != JOB_READY) return false;
    *out = clip_job_iou(job, mode);
    return true;
}

// ---- aligned pairs: two-stage evaluation (no per-box reuse, so nothing is precomputed) ---------------
// Stage 1 runs for every pair: jitter_1 (similarity mask + clamps), three degree-domain sincos, hav, and a
// conservative dead test that needs neither asin nor sqrt of hav: a series lower bound of
// arc = 2 asin(sqrt(hav)) is compared with r_g + r_p in squared form.  Stage 2 (survivors only, warp-compacted by the kernel)
// finishes the transform (arc, internal angles as sin/cos, gamma, jitter_2 checks) into a ClipJob.
struct PairS1 {
    float hav, sdt, cdt, sdp, cdp;   // sin^2(arc/2); sin/cos(dtheta/2); sin/cos(dphi/2)
    float s1, c1, s2, c2;            // sin/cos phi of box 1 and box 2
    float w1, h1, w2, h2;            // planar edges (rad)
    float g1, g2;                    // gamma (deg)
};

SPHK_HD int pair_stage1(const RawBox& x, const RawBox& y, int D, int edge, bool cull, PairS1* s) {
    bool sim = (fabsf(x.t - y.t) < kEps) | (fabsf(x.p - y.p) < kEps) | (fabsf(x.a - y.a) < kEps) | (fabsf(x.b - y.b) < kEps);
    if (D == 5) sim = sim | (fabsf(x.g - y.g) < kEps);
    // jitter_1 clamps of the unshifted boxes (sph_iou_api.py:249-258).  A centre within 1e-3 deg of the upper
    // range end needs the hi + lo arithmetic of the reference-order path: flagged slow (as are NaNs).
    bool slow = sim | !(x.t <= 359.999f) | !(y.t <= 359.999f) | !(x.p <= 179.999f) | !(y.p <= 179.999f);
    const float t1 = fmaxf(x.t, kEps2), p1 = fmaxf(x.p, kEps2);
    const float t2 = fmaxf(y.t, kEps), p2 = fmaxf(y.p, kEps);
    const float a1 = clampf(x.a, kEps2, (float)(180.0 - SPHK_EPS_D)), b1 = clampf(x.b, kEps2, (float)(180.0 - SPHK_EPS_D));
    const float a2 = clampf(y.a, kEps, (float)(180.0 - 2 * SPHK_EPS_D)), b2 = clampf(y.b, kEps, (float)(180.0 - 2 * SPHK_EPS_D));
    s->w1 = edge_len_deg(a1, edge); s->h1 = edge_len_deg(b1, edge);
    s->w2 = edge_len_deg(a2, edge); s->h2 = edge_len_deg(b2, edge);
    slow = slow | !(s->w1 >= 2.0f * kMinWh1) | !(s->h1 >= 2.0f * kMinWh1) | !(s->w2 >= 2.0f * kMinWh1) | !(s->h2 >= 2.0f * kMinWh1);
    s->g1 = 0.0f; s->g2 = 0.0f;
    if (D == 5) {
        s->g1 = x.g; s->g2 = y.g;
        slow = slow | !(fabsf(x.g) <= 179.0f) | !(fabsf(y.g) <= 179.0f);
    }
    if (slow) return JOB_SLOW;
    sincos_deg(0.5f * (t2 - t1), 0.0f, &s->sdt, &s->cdt);
    sincos_deg(0.5f * (p2 - p1), 0.0f, &s->sdp, &s->cdp);
    sincos_deg(p1, 0.0f, &s->s1, &s->c1);
    // phi_2 = phi_1 + dphi: angle addition instead of a fourth sincos
    const float sd = 2.0f * s->sdp * s->cdp, cd = fmaf(-2.0f * s->sdp, s->sdp, 1.0f);
    s->s2 = fmaf(s->s1, cd, s->c1 * sd);
    s->c2 = fmaf(s->c1, cd, -s->s1 * sd);
    s->hav = fmaf(s->s1 * s->s2, s->sdt * s->sdt, s->sdp * s->sdp);
    if (cull) {
        // (r_g + r_p)^2 with r = circumradius + margin; the rsqrt-based sqrt is within 2 ulp, far inside the margin
        const float q1 = fmaf(s->w1, s->w1, s->h1 * s->h1), q2 = fmaf(s->w2, s->w2, s->h2 * s->h2);
        const float R = fmaf(0.5f * (q1 * rsqrt_f(q1) + q2 * rsqrt_f(q2)), 1.0002f, 8e-4f);
        // asin(x) >= x (1 + x^2/6 + 3 x^4/40) on [0, 1]: a lower bound of arc/2 from hav = x^2 alone
        const float k = fmaf(s->hav, fmaf(s->hav, 0.075f, 0.16666667f), 1.0f);
        if (4.0f * s->hav * k * k > R * R) return JOB_DEAD;
    }
    return JOB_READY;
}

// Stage 1 for the pairs pair_stage1() declines because jitter_1 is NOT the plain lower clamp there: its similarity mask
// fires (one column of the two boxes within eps: the boxes are shifted apart by -2 eps / +eps first) or a centre is
// clamped at the upper end of its range (360 - eps is not a float).  Same outputs, but through the hi + lo coordinates of
// jitter1_role1 / jitter1_role2 (sphk_math.cuh) -- the arithmetic of the reference-order path -- so that the pair can
// go on through pair_stage2() and the clipper with everybody else instead of ~1500 instructions of reference-order
// code on one lane (1e-5 of random pairs: at one million pairs that lane was the tail of the whole launch).
// Still JOB_SLOW: NaNs, sizes in reach of jitter_2's minimum, |gamma| > 179 deg; pair_stage2() adds the acos clamp zones.
SPHK_HD int pair_stage1_general(const RawBox& x, const RawBox& y, int D, int edge, bool cull, PairS1* s) {
    const bool m = jitter1_mask(x, y, D);
    const JitBox g = jitter1_role1(x, m, D), p = jitter1_role2(y, m, D);
    bool slow = !(x.t == x.t) | !(x.p == x.p) | !(y.t == y.t) | !(y.p == y.p);
    s->w1 = edge_len_deg(g.a, edge); s->h1 = edge_len_deg(g.b, edge);
    s->w2 = edge_len_deg(p.a, edge); s->h2 = edge_len_deg(p.b, edge);
    slow = slow | !(s->w1 >= 2.0f * kMinWh1) | !(s->h1 >= 2.0f * kMinWh1) | !(s->w2 >= 2.0f * kMinWh1) | !(s->h2 >= 2.0f * kMinWh1);
    s->g1 = 0.0f; s->g2 = 0.0f;
    if (D == 5) {
        s->g1 = g.g; s->g2 = p.g;
        slow = slow | !(fabsf(x.g) <= 179.0f) | !(fabsf(y.g) <= 179.0f);
    }
    if (slow) return JOB_SLOW;
    sincos_deg(0.5f * (p.t_hi - g.t_hi), 0.5f * (p.t_lo - g.t_lo), &s->sdt, &s->cdt);
    sincos_deg(0.5f * (p.p_hi - g.p_hi), 0.5f * (p.p_lo - g.p_lo), &s->sdp, &s->cdp);
    sincos_deg(g.p_hi, g.p_lo, &s->s1, &s->c1);
    sincos_deg(p.p_hi, p.p_lo, &s->s2, &s->c2);
    s->hav = fmaf(s->s1 * s->s2, s->sdt * s->sdt, s->sdp * s->sdp);
    if (cull) {
        const float q1 = fmaf(s->w1, s->w1, s->h1 * s->h1), q2 = fmaf(s->w2, s->w2, s->h2 * s->h2);
        const float R = fmaf(0.5f * (q1 * rsqrt_f(q1) + q2 * rsqrt_f(q2)), 1.0002f, 8e-4f);
        const float k = fmaf(s->hav, fmaf(s->hav, 0.075f, 0.16666667f), 1.0f);
        if (4.0f * s->hav * k * k > R * R) return JOB_DEAD;
    }
    return JOB_READY;
}

// ---- stage 0 of the aligned kernel: "cannot touch" from approximate trigonometry ----------------------
// The same dead test as pair_stage1() (series lower bound of the arc against r_g + r_p), but on a haversine built
// from four MUFU sines (sin.approx: absolute error < 1e-6 on [-pi, pi]) and on the RAW sizes, at a fraction of the
// instructions: it runs for every pair, everything exact runs only for the survivors.  Conservative by construction:
//   * hav is lowered by kHavSlack = 2e-5 >= 3 x (6 sine errors);
//   * raw alpha / beta (deg -> rad) can only exceed the clamped, edge-converted planar sizes for 'arc' and 'chord'
//     (2 sin(a/2) <= a; jitter_1's lower clamp adds <= 4.3e-6 rad, inside the 8e-4 margin); 'tangent' is not culled;
//   * centres outside [0, 360] x [0, 180] (which jitter_1 would clamp), -0.0 and NaNs are never culled: one unsigned
//     compare on the bit patterns (non-negative floats order like unsigned integers);
//   * NaN / inf sizes make the comparison false.
// true => pair_stage1(cull) would return JOB_DEAD as well, and the reference's IoU is exactly 0.
constexpr float kHavSlack = 2e-5f;
SPHK_HD float sin_approx(float x) {
#if defined(__CUDA_ARCH__)
    return __sinf(x);
#else
    return sinf(x);
#endif
}
SPHK_HD uint32_t f32_bits(float x) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(x);
#else
    uint32_t u;
    memcpy(&u, &x, 4);
    return u;
#endif
}
SPHK_HD bool pair_far_apart(const RawBox& x, const RawBox& y, int edge) {
    const uint32_t ta = f32_bits(x.t), tb = f32_bits(y.t), pa = f32_bits(x.p + x.p), pb = f32_bits(y.p + y.p);
    const uint32_t m1 = ta > tb ? ta : tb, m2 = pa > pb ? pa : pb;
    if ((m1 > m2 ? m1 : m2) > 0x43B40000u || edge == EDGE_TANGENT) return false;       // 360.0f
    const float kh = (float)(SPHK_PI_D / 360.0);
    const float sdt = sin_approx((y.t - x.t) * kh), sdp = sin_approx((y.p - x.p) * kh);
    const float s1 = sin_approx(x.p * kDeg2Rad), s2 = sin_approx(y.p * kDeg2Rad);
    const float hav = fmaf(s1 * s2, sdt * sdt, sdp * sdp) - kHavSlack;
    const float q1 = fmaf(x.a, x.a, x.b * x.b), q2 = fmaf(y.a, y.a, y.b * y.b);
    const float R = fmaf((q1 * rsqrt_f(q1) + q2 * rsqrt_f(q2)) * (0.5f * kDeg2Rad), 1.0002f, 8e-4f);
    const float k = fmaf(hav, fmaf(hav, 0.075f, 0.16666667f), 1.0f);
    return 4.0f * hav * k * k > R * R;
}

SPHK_HD int pair_stage2(const PairS1& s, int D, int kind, ClipJob* job) {
    const float hth = s.sdt * s.sdt;
    const float sin_dth = 2.0f * s.sdt * s.cdt, sin_dph = 2.0f * s.sdp * s.cdp;
    const float arc = arc_from_hav(s.hav);
    if (!(arc > 2e-3f && arc < 3.14f)) return JOB_SLOW;
    const float ng = fmaf(-2.0f * s.c1 * s.s2, hth, sin_dph), mg = -s.s2 * sin_dth;
    const float np = fmaf(2.0f * s.c2 * s.s1, hth, sin_dph), mp = -s.s1 * sin_dth;
    const float S2g = fmaf(ng, ng, mg * mg), S2p = fmaf(np, np, mp * mp);
    if (!(S2g > 1e-30f && S2p > 1e-30f)) return JOB_SLOW;
    const float ig = rsqrt_f(S2g), ip = rsqrt_f(S2p);
    float sg1 = 0.0f, cg1 = 1.0f, sg2 = 0.0f, cg2 = 1.0f;
    if (D == 5) {
        sincos_deg(s.g1, 0.0f, &sg1, &cg1);
        sincos_deg(s.g2, 0.0f, &sg2, &cg2);
    }
    finish_job(arc, ng * ig, mg * ig, np * ip, mp * ip, sg1, cg1, sg2, cg2, s.g1, s.g2, s.w1, s.h1, s.w2, s.h2, D, kind, job);
    return JOB_READY;
}

}  // namespace sphk
