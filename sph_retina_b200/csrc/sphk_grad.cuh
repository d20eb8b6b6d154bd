// sphk_grad.cuh -- analytic backward of the Sph2Pob IoU pair pipeline (fp32, in registers).
//
// Reference dataflow being reproduced (SURVEY.md 8a note 9): autograd through
//   clone -> jiter_spherical_bboxes -> sph2pob_standard -> jiter_rotated_bboxes
//         -> diff_iou_rotated_2d -> clamp(0,1)
// (sphdet/losses/sph2pob_transform.py:24-35, sphdet/losses/sph2pob_iou_loss.py:122,
//  sphdet/iou/diff_iou_rotated.py:325-343).  Gradients pass unchanged through the jitter
// offsets, are zero where a clamp is active, and masks / signs / vertex order are constants.
//
// Nothing is saved by the forward: the backward recomputes the pair (it is cheaper than one
// round trip of the intermediates through HBM) and differentiates it in closed form:
//   * intersection area: dI = sum over boundary pieces of (piece length) x (normal velocity),
//     (Reynolds transport) -- it needs exactly the clipped intervals the forward already has;
//   * arc and the two internal angles: spherical-trig derivatives of the tangent-plane bearings.
#pragma once
#include "sphk_math.cuh"

namespace sphk {

// d(IoU or IoF)/d(obb1, obb2) scaled by the upstream gradient `giou`.
// g1 / g2 receive (x, y, w, h, a).  Returns the clamped IoU.
SPHK_HD float riou_grad(const ObbPair& o, int mode, float giou, float* g1, float* g2) {
    RiouGeom G;
    EdgeClip E;
    const float A1 = o.w1 * o.h1, A2 = o.w2 * o.h2;
    const float Iraw = riou_intersection(o, &G, &E);
    float I = fminf(fmaxf(Iraw, 0.0f), fminf(A1, A2));
    if (!(I == I)) I = 0.0f;
    const float U = (mode == MODE_IOF) ? A1 : (A1 + A2 - I);
    const float iou_raw = I / U;
    const float iou = clampf(iou_raw, 0.0f, 1.0f);
    const bool live = (giou != 0.0f) && (iou_raw >= 0.0f) && (iou_raw <= 1.0f) && (I > 0.0f);
    if (!live) {
        for (int k = 0; k < 5; ++k) { g1[k] = 0.0f; g2[k] = 0.0f; }
        return iou;
    }
    // --- pieces of box 2's edges (lengths and mid coordinates in box 2's own axes)
    const float d0 = fmaxf(E.t1[0] - E.t0[0], 0.0f), d1 = fmaxf(E.t1[1] - E.t0[1], 0.0f);
    const float d2 = fmaxf(E.t1[2] - E.t0[2], 0.0f), d3 = fmaxf(E.t1[3] - E.t0[3], 0.0f);
    const float l0 = o.w2 * d0, l1 = o.h2 * d1, l2 = o.w2 * d2, l3 = o.h2 * d3;
    const float hw2 = 0.5f * o.w2, hh2 = 0.5f * o.h2;
    const float m0 = hw2 - o.w2 * 0.5f * (E.t0[0] + E.t1[0]);    // u-coordinate of the piece on the +v side
    const float m1 = hh2 - o.h2 * 0.5f * (E.t0[1] + E.t1[1]);    // v-coordinate, -u side
    const float m2 = -hw2 + o.w2 * 0.5f * (E.t0[2] + E.t1[2]);   // u-coordinate, -v side
    const float m3 = -hh2 + o.h2 * 0.5f * (E.t0[3] + E.t1[3]);   // v-coordinate, +u side
    // dI/d(centre of box 2), frame-1 components: (l3 - l1) u^ + (l0 - l2) v^
    const float gu = l3 - l1, gv = l0 - l2;
    const float gpx = gu * G.cr - gv * G.sr, gpy = gu * G.sr + gv * G.cr;
    // to world axes: R(a1)
    const float gx2 = G.c1 * gpx - G.s1 * gpy, gy2 = G.s1 * gpx + G.c1 * gpy;
    const float gw2 = 0.5f * (l1 + l3), gh2 = 0.5f * (l0 + l2);
    const float ga2 = (d0 > 0.0f ? l0 * m0 : 0.0f) + (d1 > 0.0f ? l1 * m1 : 0.0f) - (d2 > 0.0f ? l2 * m2 : 0.0f) -
                      (d3 > 0.0f ? l3 * m3 : 0.0f);
    // --- pieces of box 1's sides: top(0) left(1) bottom(2) right(3)
    const float L0 = fmaxf(E.hi[0] - E.lo[0], 0.0f), L1 = fmaxf(E.hi[1] - E.lo[1], 0.0f);
    const float L2 = fmaxf(E.hi[2] - E.lo[2], 0.0f), L3 = fmaxf(E.hi[3] - E.lo[3], 0.0f);
    const float c0 = 0.5f * (E.hi[0] + E.lo[0]), c1 = 0.5f * (E.hi[1] + E.lo[1]);
    const float c2 = 0.5f * (E.hi[2] + E.lo[2]), c3 = 0.5f * (E.hi[3] + E.lo[3]);
    const float gw1 = 0.5f * (L1 + L3), gh1 = 0.5f * (L0 + L2);
    const float ga1 = (L0 > 0.0f ? L0 * c0 : 0.0f) + (L1 > 0.0f ? L1 * c1 : 0.0f) - (L2 > 0.0f ? L2 * c2 : 0.0f) -
                      (L3 > 0.0f ? L3 * c3 : 0.0f);
    // --- IoU = I / U
    float kI, kA1, kA2;   // d iou / dI, dA1, dA2
    if (mode == MODE_IOF) {
        kI = 1.0f / A1; kA1 = -I / (A1 * A1); kA2 = 0.0f;
    } else {
        const float iu2 = 1.0f / (U * U);
        kI = (A1 + A2) * iu2; kA1 = -I * iu2; kA2 = kA1;
    }
    kI *= giou; kA1 *= giou; kA2 *= giou;
    g1[0] = -kI * gx2; g1[1] = -kI * gy2;
    g1[2] = kI * gw1 + kA1 * o.h1; g1[3] = kI * gh1 + kA1 * o.w1; g1[4] = kI * ga1;
    g2[0] = kI * gx2; g2[1] = kI * gy2;
    g2[2] = kI * gw2 + kA2 * o.h2; g2[3] = kI * gh2 + kA2 * o.w2; g2[4] = kI * ga2;
    return iou;
}

// backward of jitter_2: zero the components whose clamp was active
SPHK_HD void jitter2_grad(uint32_t pass, float* g1, float* g2) {
    if (!(pass & 1u)) g1[2] = 0.0f;
    if (!(pass & 2u)) g1[3] = 0.0f;
    if (!(pass & 4u)) g1[4] = 0.0f;
    if (!(pass & 8u)) g2[2] = 0.0f;
    if (!(pass & 16u)) g2[3] = 0.0f;
    if (!(pass & 32u)) g2[4] = 0.0f;
}

// backward of the transform (+ jitter_1): OBB gradients -> gradients w.r.t. the raw boxes in DEGREES.
// gb1/gb2 receive (theta, phi, alpha, beta, gamma).
SPHK_HD void xform_grad(int kind, const JitBox& g, const JitBox& p, int D, int edge, const XformAux& aux,
                        const float* go1, const float* go2, float* gb1, float* gb2) {
    const SphGeom q = sph_geom(g, p);
    // gradient w.r.t. arc and the two pre-gamma angles
    float g_arc, g_ag = go1[4], g_ap = go2[4];
    if (kind == KIND_SPH2POB_STANDARD) g_arc = aux.degenerate ? 0.0f : 0.5f * (go2[0] - go1[0]);
    else g_arc = go2[0];
    if (aux.arc_clamped) g_arc = 0.0f;
    float g_gam1 = -g_ag, g_gam2 = -g_ap;   // angle = a - gamma
    if (aux.ag_clamped) { g_ag = 0.0f; if (kind == KIND_SPH2POB_STANDARD) g_gam1 = 0.0f; }
    if (aux.ap_clamped) { g_ap = 0.0f; if (kind == KIND_SPH2POB_STANDARD) g_gam2 = 0.0f; }
    if (aux.degenerate) { g_ag = 0.0f; g_ap = 0.0f; }

    const float S2 = fmaxf(q.ng * q.ng + q.mg * q.mg, 1e-30f);     // sin^2(arc)
    const float iS = 1.0f / sqrtf(S2), iS2 = 1.0f / S2;
    const float cos_dth = 1.0f - 2.0f * q.sdt * q.sdt, sin_dth = 2.0f * q.sdt * q.cdt;
    const float cos_arc = 1.0f - 2.0f * q.hav;
    const float s1 = q.s1, c1 = q.c1, s2 = q.s2, c2 = q.c2;
    // arc
    const float arc_tg = s1 * q.mg * iS, arc_pg = -q.ng * iS;
    const float arc_tp = -s2 * q.mp * iS, arc_pp = q.np * iS;
    // a_g = atan2(N, M): N = c1 s2 cos(dth) - s1 c2, M = -s2 sin(dth)
    const float N = q.ng, M = q.mg;
    const float ag_tg = (M * (c1 * s2 * sin_dth) - N * (s2 * cos_dth)) * iS2;
    const float ag_tp = -ag_tg;
    const float ag_pg = -M * cos_arc * iS2;
    const float ag_pp = (M * (c1 * c2 * cos_dth + s1 * s2) + N * (c2 * sin_dth)) * iS2;
    // a_p = atan2(N', M'): N' = s2 c1 - c2 s1 cos(dth), M' = -s1 sin(dth)
    const float Np = q.np, Mp = q.mp;
    const float ap_tg = (Mp * (-c2 * s1 * sin_dth) - Np * (s1 * cos_dth)) * iS2;
    const float ap_tp = -ap_tg;
    const float ap_pg = (Mp * (-s1 * s2 - c1 * c2 * cos_dth) + Np * (c1 * sin_dth)) * iS2;
    const float ap_pp = Mp * cos_arc * iS2;

    const float k = kDeg2Rad;
    float t1 = g_arc * arc_tg + g_ag * ag_tg + g_ap * ap_tg;
    float f1 = g_arc * arc_pg + g_ag * ag_pg + g_ap * ap_pg;
    float t2 = g_arc * arc_tp + g_ag * ag_tp + g_ap * ap_tp;
    float f2 = g_arc * arc_pp + g_ag * ag_pp + g_ap * ap_pp;
    gb1[0] = (g.pass & 1u) ? t1 * k : 0.0f;
    gb1[1] = (g.pass & 2u) ? f1 * k : 0.0f;
    gb1[2] = (g.pass & 4u) ? go1[2] * edge_len_grad_deg(g.a, edge) * k : 0.0f;
    gb1[3] = (g.pass & 8u) ? go1[3] * edge_len_grad_deg(g.b, edge) * k : 0.0f;
    gb1[4] = (D == 5 && (g.pass & 16u)) ? g_gam1 * k : 0.0f;
    gb2[0] = (p.pass & 1u) ? t2 * k : 0.0f;
    gb2[1] = (p.pass & 2u) ? f2 * k : 0.0f;
    gb2[2] = (p.pass & 4u) ? go2[2] * edge_len_grad_deg(p.a, edge) * k : 0.0f;
    gb2[3] = (p.pass & 8u) ? go2[3] * edge_len_grad_deg(p.b, edge) * k : 0.0f;
    gb2[4] = (D == 5 && (p.pass & 16u)) ? g_gam2 * k : 0.0f;
}

// Full pair: IoU (mode 'iou') and, scaled by giou = d(total)/d(iou), the gradients w.r.t. both raw boxes.
SPHK_HD float sph2pob_iou_pair_grad(const RawBox& b1, const RawBox& b2, int D, int kind, int edge, float giou,
                                    float* gb1, float* gb2) {
    const bool m = jitter1_mask(b1, b2, D);
    const JitBox g = jitter1_role1(b1, m, D), p = jitter1_role2(b2, m, D);
    XformAux aux;
    ObbPair o = (kind == KIND_SPH2POB_STANDARD) ? sph2pob_standard(g, p, D, edge, &aux)
                                                : sph2pob_efficient(g, p, D, edge, &aux);
    const uint32_t pass2 = jitter2(o);
    float go1[5], go2[5];
    if (obb_disjoint(o)) {
        for (int k = 0; k < 5; ++k) { gb1[k] = 0.0f; gb2[k] = 0.0f; }
        return 0.0f;
    }
    const float iou = riou_grad(o, MODE_IOU, giou, go1, go2);
    jitter2_grad(pass2, go1, go2);
    xform_grad(kind, g, p, D, edge, aux, go1, go2, gb1, gb2);
    return iou;
}

}  // namespace sphk
