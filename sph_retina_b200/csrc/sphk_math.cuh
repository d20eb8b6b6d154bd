// sphk_math.cuh -- per-pair arithmetic of the spherical-box IoU hot path (fp32).
//
// One box pair is processed entirely in registers:
//   jitter_1 -> Sph2Pob transform (efficient | standard) -> jitter_2 -> rotated-box IoU
// following the reference semantics
//   sphdet/iou/sph_iou_api.py:48-86,222-260      (pair pipeline + both jitters)
//   sphdet/iou/sph2pob_efficient.py:9-73         (efficient transform)
//   sphdet/iou/sph2pob_standard.py:8-80          (standard transform, used by the loss)
//   sphdet/iou/diff_iou_rotated.py:325-343       (rotated IoU = area(A n B) / union)
//   sphdet/iou/approximate_ious.py:3-55          (Sph-IoU / FoV-IoU)
//
// It is NOT a transcription of those files.  The reference builds 3-D vectors, a
// cross product and acos() of normalised dot products, which is ill-conditioned in
// fp32 for near-coincident boxes (its own fp32 run is off by up to 4e-3 from its
// fp64 run there, tests/golden).  Here the same quantities are obtained from
// closed forms in the angle differences (haversine arc, tangent-plane bearings)
// and the intersection area from a branch-free boundary integral, so that the fp32
// kernel tracks the float64 run of the reference to ~1e-6.  DESIGN.md derives the
// identities.
//
// The header compiles with nvcc (device) and with g++ (tests/hostsim, a CPU build of
// the very same arithmetic used only by the no-GPU unit tests).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SPHK_HD __host__ __device__ __forceinline__
#else
#define SPHK_HD inline
#endif

namespace sphk {

// ---- enums shared with the C ABI (include/sphk.h) -------------------------------------
enum Kind { KIND_SPH2POB_EFFICIENT = 0, KIND_SPH2POB_STANDARD = 1, KIND_SPH = 2, KIND_FOV = 3, KIND_NAIVE = 4, KIND_UNBIASED = 5,
            KIND_SPH2POB_LEGACY = 6 };
enum Mode { MODE_IOU = 0, MODE_IOF = 1 };
enum Edge { EDGE_ARC = 0, EDGE_CHORD = 1, EDGE_TANGENT = 2 };

// ---- reference constants (python doubles rounded to fp32 exactly as torch does) -------
#define SPHK_EPS_D (1e-4 * 1.2345678)   // sph_iou_api.py:223,245
#define SPHK_EPSA_D (1e-3 * 1.2345678)  // sph_iou_api.py:232
#define SPHK_PI_D 3.14159265358979323846
constexpr float kEps = (float)SPHK_EPS_D;
constexpr float kEps2 = (float)(2 * SPHK_EPS_D);
constexpr float kEps3 = (float)(3 * SPHK_EPS_D);
constexpr float kEps5 = (float)(5 * SPHK_EPS_D);
constexpr float kEpsA = (float)SPHK_EPSA_D;
constexpr float kEpsA2 = (float)(2 * SPHK_EPSA_D);
constexpr float kPi = (float)SPHK_PI_D;
constexpr float kHalfPi = (float)(SPHK_PI_D / 2);
constexpr float kTwoPi = (float)(2 * SPHK_PI_D);
constexpr float kDeg2Rad = (float)(SPHK_PI_D / 180.0);
// acos(clamp(x, -1+1e-7, 1-1e-7)) can never leave [kAcosLo, pi-kAcosLo]
// (sph2pob_efficient.py:205).  Value of the float64 run: acos(1-1e-7) = 4.47213602e-4.
constexpr float kAcosLo = 4.47213602e-4f;
constexpr float kAcosHi = (float)(SPHK_PI_D - 4.47213602e-4);
// jitter_2 clamps (sph_iou_api.py:237-240)
constexpr float kMinWh1 = (float)(2 * SPHK_EPSA_D / 10);
constexpr float kMinWh2 = (float)(SPHK_EPSA_D / 10);
constexpr float kA1Lo = (float)(-2 * SPHK_PI_D + 2 * SPHK_EPSA_D);
constexpr float kA1Hi = (float)(2 * SPHK_PI_D - SPHK_EPSA_D);
constexpr float kA2Lo = (float)(-2 * SPHK_PI_D + SPHK_EPSA_D);
constexpr float kA2Hi = (float)(2 * SPHK_PI_D - 2 * SPHK_EPSA_D);

SPHK_HD float fmin2(float a, float b) { return fminf(a, b); }
SPHK_HD float fmax2(float a, float b) { return fmaxf(a, b); }
SPHK_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// ---- raw box and its jittered, role-specific form -------------------------------------
struct RawBox {
    float t, p, a, b, g;  // theta, phi, alpha, beta, gamma in degrees (g = 0 for BFoV)
};

// A jittered coordinate is kept as hi + lo (hi: the fp32 input or a range end, lo: the
// jitter offset) so that the angle DIFFERENCES of near-identical boxes stay exact
// (fl(300 - 2eps) alone is already off by 8% of the 3eps separation jitter_1 creates).
struct JitBox {
    float t_hi, t_lo, p_hi, p_lo;  // degrees
    float a, b, g;                 // degrees
    uint32_t pass;                 // bit k: column k not clamped (gradient passes)
};

// sph_iou_api.py:246-247: similar_mask = any_k |b1[k]-b2[k]| < eps over all D columns
SPHK_HD bool jitter1_mask(const RawBox& x, const RawBox& y, int D) {
    bool m = (fabsf(x.t - y.t) < kEps) | (fabsf(x.p - y.p) < kEps) | (fabsf(x.a - y.a) < kEps) |
             (fabsf(x.b - y.b) < kEps);
    if (D == 5) m = m | (fabsf(x.g - y.g) < kEps);
    return m;
}

// one coordinate: value = raw + shift, clamped to [lo_end + lo_off, hi_end + hi_off]
// (the bounds are "range end + small offset" so they are carried exactly too).
SPHK_HD void jit_coord(float raw, float shift, float lo_off, float hi_end, float hi_off, float& hi, float& lo,
                       bool& pass) {
    const bool below = (raw + shift) < lo_off;                // lower range end is 0
    const bool above = ((raw - hi_end) + shift) > hi_off;     // raw - hi_end is exact near the end
    hi = below ? 0.0f : (above ? hi_end : raw);
    lo = below ? lo_off : (above ? hi_off : shift);
    pass = !(below | above);
}

// sph_iou_api.py:249-258, role of bboxes1 (rows / gt / pred / NMS pivot)
SPHK_HD JitBox jitter1_role1(const RawBox& x, bool m, int D) {
    const float s = m ? -kEps2 : 0.0f;
    JitBox o;
    bool p0, p1, p2, p3;
    float ah, al, bh, bl;
    jit_coord(x.t, s, kEps2, 360.0f, -kEps, o.t_hi, o.t_lo, p0);
    jit_coord(x.p, s, kEps2, 180.0f, -kEps, o.p_hi, o.p_lo, p1);
    jit_coord(x.a, s, kEps2, 180.0f, -kEps, ah, al, p2);
    jit_coord(x.b, s, kEps2, 180.0f, -kEps, bh, bl, p3);
    o.a = ah + al;
    o.b = bh + bl;
    o.g = (D == 5) ? x.g + s : 0.0f;  // bboxes1's gamma is never clamped (:256-258)
    o.pass = (uint32_t)p0 | ((uint32_t)p1 << 1) | ((uint32_t)p2 << 2) | ((uint32_t)p3 << 3) | (1u << 4);
    return o;
}

// role of bboxes2 (cols / anchors / target / NMS candidates)
SPHK_HD JitBox jitter1_role2(const RawBox& x, bool m, int D) {
    const float s = m ? kEps : 0.0f;
    JitBox o;
    bool p0, p1, p2, p3, p4 = true;
    float ah, al, bh, bl;
    jit_coord(x.t, s, kEps, 360.0f, -kEps2, o.t_hi, o.t_lo, p0);
    jit_coord(x.p, s, kEps, 180.0f, -kEps2, o.p_hi, o.p_lo, p1);
    jit_coord(x.a, s, kEps, 180.0f, -kEps2, ah, al, p2);
    jit_coord(x.b, s, kEps, 180.0f, -kEps2, bh, bl, p3);
    o.a = ah + al;
    o.b = bh + bl;
    o.g = 0.0f;
    if (D == 5) {
        // clamped twice: net [-360+2eps, 360-2eps]
        const float v = x.g + s;
        const bool below = v < (-360.0f + kEps2), above = ((x.g - 360.0f) + s) > -kEps2;
        o.g = below ? (float)(-360.0 + 2 * SPHK_EPS_D) : (above ? (float)(360.0 - 2 * SPHK_EPS_D) : v);
        p4 = !(below | above);
    }
    o.pass = (uint32_t)p0 | ((uint32_t)p1 << 1) | ((uint32_t)p2 << 2) | ((uint32_t)p3 << 3) | ((uint32_t)p4 << 4);
    return o;
}

// ---- planar oriented box pair (what the transform hands to the rotated IoU) ----------
struct ObbPair {
    float x1, y1, w1, h1, a1;
    float x2, y2, w2, h2, a2;
};

// Spherical part shared by both transforms.  With g = box 1 centre, p = box 2 centre,
// (d, e) = (south, east) unit tangents, dth = theta_p - theta_g, dph = phi_p - phi_g:
//   hav   = sin^2(arc/2) = sin^2(dph/2) + sin(phi_g) sin(phi_p) sin^2(dth/2)
//   p.e_g = sin(phi_p) sin(dth)              p.d_g =  sin(dph) - 2 cos(phi_g) sin(phi_p) sin^2(dth/2)
//   g.e_p = -sin(phi_g) sin(dth)             g.d_p = -sin(dph) - 2 cos(phi_p) sin(phi_g) sin^2(dth/2)
// and the reference's signed internal angles are (sph2pob_efficient.py:81-97 rewritten
// in the tangent basis; z = g x p normalised):
//   a_g = atan2( p.d_g, -p.e_g )             a_p = atan2( -g.d_p,  g.e_p )
struct SphGeom {
    float hav;             // sin^2(arc/2), unclamped
    float ng, mg, np, mp;  // a_g = atan2(ng, mg), a_p = atan2(np, mp)   (|.| ~ sin(arc))
    float s1, c1, s2, c2;  // sin/cos phi of both boxes
    float sdt, cdt;        // sin/cos(dth/2)
    float sdp, cdp;        // sin/cos(dph/2)
};

SPHK_HD void sincos_f(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
    sincosf(x, s, c);
#else
    *s = sinf(x);
    *c = cosf(x);
#endif
}

// sin/cos of an angle given in DEGREES as hi + lo.  The range reduction to [-45,45] happens in
// degrees, where it is exact in fp32 (hi - 90k loses no bits), so sin AND cos keep full relative
// accuracy next to every multiple of 90 degrees -- which radians-based sincosf cannot deliver
// (fl(pi/2) is already 4e-8 off) and which near-antipodal / near-equator pairs need.
// Kernels on [-pi/4, pi/4]: single-precision minimax polynomials (cephes sinf/cosf), ~1 ulp.
SPHK_HD void sincos_deg(float hi, float lo, float* s, float* c) {
    const float k = rintf(hi * (1.0f / 90.0f));
    const float y = fmaf(-90.0f, k, hi) + lo;      // exact reduction, then the small offset
    const float r = y * kDeg2Rad;
    const float z = r * r;
    const float ps = fmaf(fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f) * z, r, r);
    const float pc = fmaf(fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f),
                          z * z, fmaf(-0.5f, z, 1.0f));
    const int q = ((int)k) & 3;
    const float a = (q & 1) ? pc : ps;             // |sin| source
    const float b = (q & 1) ? ps : pc;             // |cos| source
    *s = (q & 2) ? -a : a;
    *c = ((q + 1) & 2) ? -b : b;
}

// sin^2(x) and sin(2 x) of an angle x given in degrees: what the haversine formulation needs of the HALF differences of
// the centres (hav = s1 s2 sin^2(dt/2) + sin^2(dp/2); sin dt, sin dp for the bearings).  Same reduction and the same
// polynomials as sincos_deg, bit-identical to (s * s, (2 s) * c) of its outputs, without assembling the signs of s and c
// separately: sign(s) sign(c) is negative exactly in the odd quadrants.
SPHK_HD void sin2_and_sin_double_deg(float x, float* s2, float* sin2x) {
    const float k = rintf(x * (1.0f / 90.0f));
    const float y = fmaf(-90.0f, k, x) + 0.0f;
    const float r = y * kDeg2Rad;
    const float z = r * r;
    const float ps = fmaf(fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f) * z, r, r);
    const float pc = fmaf(fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f),
                          z * z, fmaf(-0.5f, z, 1.0f));
    const bool odd = (((int)k) & 1) != 0;
    const float a = odd ? pc : ps, b = odd ? ps : pc;
    *s2 = a * a;
    const float t = (2.0f * a) * b;
    *sin2x = odd ? -t : t;
}

// planar edge length of a field-of-view angle (degrees in), sph2pob_efficient.py:100-108.
// chord/tangent go through the degree-domain sincos so that 2 tan(alpha/2) stays accurate for
// alpha -> 180 (oversize anchors are clamped to 180 - eps by jitter_1).
SPHK_HD float edge_len_deg(float fov_deg, int edge) {
    if (edge == EDGE_ARC) return fov_deg * kDeg2Rad;
    float s, c;
    sincos_deg(0.5f * fov_deg, 0.0f, &s, &c);
    return edge == EDGE_CHORD ? 2.0f * s : 2.0f * s / c;
}
// d(edge length)/d(fov in radians)
SPHK_HD float edge_len_grad_deg(float fov_deg, int edge) {
    if (edge == EDGE_ARC) return 1.0f;
    float s, c;
    sincos_deg(0.5f * fov_deg, 0.0f, &s, &c);
    return edge == EDGE_CHORD ? c : 1.0f / (c * c);
}

SPHK_HD SphGeom sph_geom(const JitBox& g, const JitBox& p) {
    SphGeom q;
    const float dth_hi = p.t_hi - g.t_hi, dth_lo = p.t_lo - g.t_lo;   // theta_p - theta_g (deg)
    const float dph_hi = p.p_hi - g.p_hi, dph_lo = p.p_lo - g.p_lo;
    sincos_deg(g.p_hi, g.p_lo, &q.s1, &q.c1);
    sincos_deg(p.p_hi, p.p_lo, &q.s2, &q.c2);
    sincos_deg(0.5f * dth_hi, 0.5f * dth_lo, &q.sdt, &q.cdt);
    sincos_deg(0.5f * dph_hi, 0.5f * dph_lo, &q.sdp, &q.cdp);
    const float hth = q.sdt * q.sdt;
    const float sin_dth = 2.0f * q.sdt * q.cdt;
    const float sin_dph = 2.0f * q.sdp * q.cdp;
    q.hav = fmaf(q.s1 * q.s2, hth, q.sdp * q.sdp);
    q.ng = fmaf(-2.0f * q.c1 * q.s2, hth, sin_dph);        //  p.d_g
    q.mg = -q.s2 * sin_dth;                                // -p.e_g
    q.np = fmaf(2.0f * q.c2 * q.s1, hth, sin_dph);         // -g.d_p
    q.mp = -q.s1 * sin_dth;                                //  g.e_p
    return q;
}

// |acos(clamp(.))| * sign with the reference's conventions: magnitude confined to
// [kAcosLo, kAcosHi]; sign +1 iff the "numerator" is > 0, else -1 (zero -> -1,
// sph2pob_efficient.py:224-225).
SPHK_HD float signed_clamped_angle(float n, float m, bool* clamped) {
    float mag = atan2f(fabsf(n), m);
    if (n == 0.0f && m == 0.0f) mag = kHalfPi;  // z = 0: F.normalize gives 0 -> acos(0)
    const float c = clampf(mag, kAcosLo, kAcosHi);
    *clamped = (c != mag);
    return n > 0.0f ? c : -c;
}

// 1/sqrt(x) to ~1 ulp (MUFU.RSQ): used to normalise (cos, sin) pairs and for the conservative cull radius, where
// a 2e-7 relative error is immaterial (rsqrtf() without fast-math adds a dozen instructions of special-case handling)
SPHK_HD float rsqrt_f(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / sqrtf(x);
#endif
}

// arc = 2 asin(sqrt(hav)) for hav = sin^2(arc / 2) in [0, 1], without sqrtf / asinf (IEEE square root + the library
// arcsine cost 38 instructions per pair, 4 % of the N x M kernels): with z = min(hav, 1 - hav) <= 1/2 (1 - hav is exact
// for hav > 1/2) and s = sqrt(z) from MUFU.RSQ + one Newton step,
//     asin(s) = s + s z P(z)      (P: degree-6 minimax fit of (asin(sqrt z) / sqrt z - 1) / z on [0, 1/2], 5e-8 relative),
// arc = 2 asin(s) for hav <= 1/2, pi - 2 asin(s) above (pi as hi + lo).  Measured against float64 on 3 M values
// (uniform and log-uniform hav): <= 1.75 ulp, 0.38 ulp on average.  hav a rounding error above 1 gives pi, as
// min(sqrt(hav), 1) did; NaN stays NaN.
SPHK_HD float arc_from_hav(float hav) {
    const bool big = hav > 0.5f;
    float z = big ? 1.0f - hav : hav;
    z = (z < 1e-30f) ? 1e-30f : z;
    const float y = rsqrt_f(z);
    const float s0 = z * y;
    const float s = fmaf(fmaf(-s0, s0, z), 0.5f * y, s0);
    float p = 0.12372123607510499f;
    p = fmaf(p, z, -0.11530392646856973f);
    p = fmaf(p, z, 0.09340074622705977f);
    p = fmaf(p, z, 0.010430741612056323f);
    p = fmaf(p, z, 0.04762428807115189f);
    p = fmaf(p, z, 0.07478349097117638f);
    p = fmaf(p, z, 0.16667234492246794f);
    const float a = fmaf(s * z, p, s);
    const float two = a + a;
    return big ? (3.14159274f - two) + -8.74227766e-8f : two;
}

struct XformAux {     // what the backward pass needs to know about active clamps
    bool arc_clamped, ag_clamped, ap_clamped, degenerate;
};

// Sph2Pob-efficient: sph2pob_efficient.py:9-73 (rbb_angle='equator').
SPHK_HD ObbPair sph2pob_efficient(const JitBox& g, const JitBox& p, int D, int edge, XformAux* aux) {
    const SphGeom q = sph_geom(g, p);
    ObbPair o;
    const float arc = arc_from_hav(q.hav);
    const float arc_c = clampf(arc, kAcosLo, kAcosHi);
    aux->arc_clamped = (arc_c != arc);
    aux->degenerate = false;
    o.a1 = signed_clamped_angle(q.ng, q.mg, &aux->ag_clamped);
    o.a2 = signed_clamped_angle(q.np, q.mp, &aux->ap_clamped);
    if (D == 5) {  // :55-57  angle -= gamma (unwrapped)
        o.a1 -= g.g * kDeg2Rad;
        o.a2 -= p.g * kDeg2Rad;
    }
    o.x1 = 0.0f; o.y1 = 0.0f; o.x2 = arc_c; o.y2 = 0.0f;
    o.w1 = edge_len_deg(g.a, edge); o.h1 = edge_len_deg(g.b, edge);
    o.w2 = edge_len_deg(p.a, edge); o.h2 = edge_len_deg(p.b, edge);
    return o;
}

// wrap to (-pi, pi]
SPHK_HD float wrap_pi(float a) {
    if (a > kPi) a -= kTwoPi;
    if (a < -kPi) a += kTwoPi;
    return a;
}

// Sph2Pob-standard: sph2pob_standard.py:8-80.  Geometrically the pair frame
// (look, right, up) puts both centres on the equator at -/+ arc/2 and `up` equals the
// efficient transform's z, so the same tangent-plane bearings apply; gamma rotates the
// tangent BEFORE the clamped acos (sph2pob_standard.py:47-54), hence the wrap here.
SPHK_HD ObbPair sph2pob_standard(const JitBox& g, const JitBox& p, int D, int edge, XformAux* aux) {
    const SphGeom q = sph_geom(g, p);
    ObbPair o;
    const float half = asinf(fminf(sqrtf(q.hav), 1.0f));   // arc/2 in [0, pi/2]
    const float half_c = fmaxf(half, kAcosLo);
    aux->arc_clamped = (half_c != half);
    // sph2pob_standard.py:291: centres closer than 1e-8 (L1) use the mid-angle frame instead
    aux->degenerate = q.hav < 2.5e-17f;
    float ag = atan2f(q.ng, q.mg), ap = atan2f(q.np, q.mp);
    if (aux->degenerate) { ag = kHalfPi; ap = kHalfPi; }
    if (D == 5) {
        ag = wrap_pi(ag - g.g * kDeg2Rad);
        ap = wrap_pi(ap - p.g * kDeg2Rad);
    }
    const float mg = clampf(fabsf(ag), kAcosLo, kAcosHi), mp = clampf(fabsf(ap), kAcosLo, kAcosHi);
    aux->ag_clamped = (mg != fabsf(ag));
    aux->ap_clamped = (mp != fabsf(ap));
    o.a1 = ag > 0.0f ? mg : -mg;
    o.a2 = ap > 0.0f ? mp : -mp;
    o.x1 = -half_c; o.x2 = aux->degenerate ? -half_c : half_c;
    o.y1 = kHalfPi; o.y2 = kHalfPi;
    o.w1 = edge_len_deg(g.a, edge); o.h1 = edge_len_deg(g.b, edge);
    o.w2 = edge_len_deg(p.a, edge); o.h2 = edge_len_deg(p.b, edge);
    return o;
}

// sph_iou_api.py:222-242 (jiter_rotated_bboxes).  `pass` bits (for the backward): 0 w1, 1 h1,
// 2 a1, 3 w2, 4 h2, 5 a2 -- set when the corresponding clamp is inactive.
SPHK_HD uint32_t jitter2(ObbPair& o) {
    const bool m = (fabsf(o.x1 - o.x2) < kEps) | (fabsf(o.w1 - o.w2) < kEps) | (fabsf(o.h1 - o.h2) < kEps) |
                   (fabsf(o.a1 - o.a2) < kEps);
    if (m) {
        o.x1 += kEps;  o.y1 += kEps;  o.w1 += kEps2; o.h1 += kEps2; o.a1 += kEps;
        o.x2 += kEps2; o.y2 += kEps2; o.w2 += kEps;  o.h2 += kEps;  o.a2 += kEps5;
    }
    if (fabsf(o.a1 - o.a2) < kEpsA) {
        o.a1 += kEpsA;
        o.a2 += kEpsA2;
    }
    uint32_t pass = 0;
    pass |= (o.w1 >= kMinWh1) ? 1u : 0u;        pass |= (o.h1 >= kMinWh1) ? 2u : 0u;
    pass |= (o.a1 >= kA1Lo && o.a1 <= kA1Hi) ? 4u : 0u;
    pass |= (o.w2 >= kMinWh2) ? 8u : 0u;        pass |= (o.h2 >= kMinWh2) ? 16u : 0u;
    pass |= (o.a2 >= kA2Lo && o.a2 <= kA2Hi) ? 32u : 0u;
    o.w1 = fmaxf(o.w1, kMinWh1); o.h1 = fmaxf(o.h1, kMinWh1);
    o.w2 = fmaxf(o.w2, kMinWh2); o.h2 = fmaxf(o.h2, kMinWh2);
    o.a1 = clampf(o.a1, kA1Lo, kA1Hi);
    o.a2 = clampf(o.a2, kA2Lo, kA2Hi);
    return pass;
}

// ---- rotated-box intersection: boundary integral, no polygon is ever materialised -----
//
// area(A n B) = 1/2 * closed integral of (x - o) x dx over the boundary of A n B, which consists
// of the parts of B's edges inside A plus the parts of A's sides inside B.  Everything is done
// in the frame of box 1 (A = [-hw1,hw1] x [-hh1,hh1], o = its centre):
//   * an edge of B is a segment P + t d, t in [0,1]; Liang-Barsky against A yields the inside
//     interval [t0,t1] without branches; contribution 1/2 (t1-t0) P x d;
//   * a side of A (say y = +hh1) is inside B between the largest lower and the smallest upper
//     bound imposed by B's four edge lines -- and those bounds are the SAME crossing points
//     P + t d already found for B's edges.  Sharing them is what keeps fp32 accurate for
//     nearly parallel boxes: a crossing of two almost parallel lines is ill-conditioned along
//     the lines (error ~ ulp / sin r) but as long as both incident boundary pieces use the
//     identical point the area error stays O(ulp) (what a polygon clipper gets for free).
// A division by zero cannot occur: sin/cos of the relative angle are nudged off exact 0.
//
// Corner order / rotation convention = diff_iou_rotated.py:297-322 (CCW by the angle,
// mmcv's clockwise=True in image coordinates).
struct EdgeClip {
    float t0[4], t1[4];  // edges of box 2 (order: +v side, -u side, -v side, +u side), parameter interval
    float lo[4], hi[4];  // sides of box 1 (order: top y=+hh1, left x=-hw1, bottom, right): coordinate interval
};

// reciprocal to ~1 ulp: MUFU.RCP + one Newton step (the IEEE-rounded __frcp_rn costs 4x as much and the
// clipper only needs the crossing parameters to fp32 rounding accuracy)
SPHK_HD float rcp_f(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
#else
    return 1.0f / x;
#endif
}

struct RiouGeom {  // everything the forward (and the backward) needs about one OBB pair
    float px, py;        // centre of box 2 in the frame of box 1
    float cr, sr;        // cos/sin of r = a2 - a1
    float c1, s1;        // cos/sin of a1
};

// one edge of box 2: start corner (Px,Py), direction (dx,dy) with reciprocals (ix,iy).
// Returns the clipped parameter interval and the four line crossings with the sides of box 1.
struct EdgeX {
    float x_top, x_bot, y_left, y_right;
};
SPHK_HD EdgeX clip_edge(float Px, float Py, float dx, float dy, float ix, float iy, float hw, float hh, float* t0,
                        float* t1) {
    const float txm = (-hw - Px) * ix, txp = (hw - Px) * ix;
    const float tym = (-hh - Py) * iy, typ = (hh - Py) * iy;
    const float lo = fmaxf(fmaxf(fminf(txm, txp), fminf(tym, typ)), 0.0f);
    const float hi = fminf(fminf(fmaxf(txm, txp), fmaxf(tym, typ)), 1.0f);
    *t0 = lo;
    *t1 = hi;
    EdgeX e;
    e.x_top = fmaf(typ, dx, Px);
    e.x_bot = fmaf(tym, dx, Px);
    e.y_left = fmaf(txm, dy, Py);
    e.y_right = fmaf(txp, dy, Py);
    return e;
}

// Intersection area of A = [-w1/2,w1/2] x [-h1/2,h1/2] with the box of size (w2,h2) centred at
// (px,py) and rotated by r (cr = cos r, sr = sin r, neither exactly 0), everything in A's frame.
SPHK_HD float riou_core(float px, float py, float cr, float sr, float w1, float h1, float w2, float h2, EdgeClip* E) {
    struct { float w1, h1, w2, h2; } o = {w1, h1, w2, h2};
    const float hw1 = 0.5f * o.w1, hh1 = 0.5f * o.h1, hw2 = 0.5f * o.w2, hh2 = 0.5f * o.h2;
    const float icr = rcp_f(cr), isr = rcp_f(sr);
    const float iw2 = rcp_f(o.w2), ih2 = rcp_f(o.h2);

    // box 2 in frame 1: axes u = (cr, sr), v = (-sr, cr); corners CCW:
    // c0 = p + hw2 u + hh2 v, c1 = p - hw2 u + hh2 v, c2 = p - hw2 u - hh2 v, c3 = p + hw2 u - hh2 v
    const float ux = cr * hw2, uy = sr * hw2, vx = -sr * hh2, vy = cr * hh2;
    const float c0x = px + ux + vx, c0y = py + uy + vy;
    const float c1x = px - ux + vx, c1y = py - uy + vy;
    const float c2x = px - ux - vx, c2y = py - uy - vy;
    const float c3x = px + ux - vx, c3y = py + uy - vy;
    // edge directions: e0 = -w2 u^ (from c0), e1 = -h2 v^ (from c1), e2 = +w2 u^ (from c2), e3 = +h2 v^ (from c3)
    const float dux = o.w2 * cr, duy = o.w2 * sr;     // w2 u^
    const float dvx = -o.h2 * sr, dvy = o.h2 * cr;    // h2 v^
    const float iux = icr * iw2, iuy = isr * iw2;     // 1/dux, 1/duy
    const float ivx = -isr * ih2, ivy = icr * ih2;    // 1/dvx, 1/dvy
    const EdgeX x0 = clip_edge(c0x, c0y, -dux, -duy, -iux, -iuy, hw1, hh1, &E->t0[0], &E->t1[0]);
    const EdgeX x1 = clip_edge(c1x, c1y, -dvx, -dvy, -ivx, -ivy, hw1, hh1, &E->t0[1], &E->t1[1]);
    const EdgeX x2 = clip_edge(c2x, c2y, dux, duy, iux, iuy, hw1, hh1, &E->t0[2], &E->t1[2]);
    const EdgeX x3 = clip_edge(c3x, c3y, dvx, dvy, ivx, ivy, hw1, hh1, &E->t0[3], &E->t1[3]);
    // P x d of a CCW edge = |d| * (signed distance from o to the edge line along the outward
    // normal); p.u^ = pxv, p.v^ = -pxu
    const float pxu = px * sr - py * cr;
    const float pxv = px * cr + py * sr;
    const float k0 = o.w2 * (hh2 - pxu);   // edge on the +v side
    const float k1 = o.h2 * (hw2 - pxv);   // edge on the -u side
    const float k2 = o.w2 * (hh2 + pxu);   // edge on the -v side
    const float k3 = o.h2 * (hw2 + pxv);   // edge on the +u side
    float area2 = 0.0f;  // twice the area
    area2 = fmaf(fmaxf(E->t1[0] - E->t0[0], 0.0f), k0, area2);
    area2 = fmaf(fmaxf(E->t1[1] - E->t0[1], 0.0f), k1, area2);
    area2 = fmaf(fmaxf(E->t1[2] - E->t0[2], 0.0f), k2, area2);
    area2 = fmaf(fmaxf(E->t1[3] - E->t0[3], 0.0f), k3, area2);

    // sides of box 1.  "inside B" = left of every CCW edge line of B:  d.x (Y - Py) - d.y (X - Px) >= 0.
    // On a horizontal side (coordinate X) an edge with d.y > 0 gives an upper bound, d.y < 0 a lower one;
    // on a vertical side (coordinate Y) d.x > 0 gives a lower bound, d.x < 0 an upper one.
    // e0/e2 have opposite directions, and so have e1/e3: one of each pair bounds from above.
    const bool up_pos = duy > 0.0f;   // e2 (d = +w2 u^) has d.y > 0
    const bool vp_pos = dvy > 0.0f;   // e3 (d = +h2 v^) has d.y > 0
    const bool ux_pos = dux > 0.0f;   // e2 has d.x > 0
    const bool vx_pos = dvx > 0.0f;   // e3 has d.x > 0
    {   // top: y = +hh1
        const float ub = fminf(up_pos ? x2.x_top : x0.x_top, vp_pos ? x3.x_top : x1.x_top);
        const float lb = fmaxf(up_pos ? x0.x_top : x2.x_top, vp_pos ? x1.x_top : x3.x_top);
        E->lo[0] = fmaxf(lb, -hw1); E->hi[0] = fminf(ub, hw1);
    }
    {   // bottom: y = -hh1
        const float ub = fminf(up_pos ? x2.x_bot : x0.x_bot, vp_pos ? x3.x_bot : x1.x_bot);
        const float lb = fmaxf(up_pos ? x0.x_bot : x2.x_bot, vp_pos ? x1.x_bot : x3.x_bot);
        E->lo[2] = fmaxf(lb, -hw1); E->hi[2] = fminf(ub, hw1);
    }
    {   // left: x = -hw1
        const float lb = fmaxf(ux_pos ? x2.y_left : x0.y_left, vx_pos ? x3.y_left : x1.y_left);
        const float ub = fminf(ux_pos ? x0.y_left : x2.y_left, vx_pos ? x1.y_left : x3.y_left);
        E->lo[1] = fmaxf(lb, -hh1); E->hi[1] = fminf(ub, hh1);
    }
    {   // right: x = +hw1
        const float lb = fmaxf(ux_pos ? x2.y_right : x0.y_right, vx_pos ? x3.y_right : x1.y_right);
        const float ub = fminf(ux_pos ? x0.y_right : x2.y_right, vx_pos ? x1.y_right : x3.y_right);
        E->lo[3] = fmaxf(lb, -hh1); E->hi[3] = fminf(ub, hh1);
    }
    const float lh = fmaxf(E->hi[0] - E->lo[0], 0.0f) + fmaxf(E->hi[2] - E->lo[2], 0.0f);  // horizontal sides
    const float lv = fmaxf(E->hi[1] - E->lo[1], 0.0f) + fmaxf(E->hi[3] - E->lo[3], 0.0f);  // vertical sides
    area2 = fmaf(lh, hh1, area2);
    area2 = fmaf(lv, hw1, area2);
    return 0.5f * area2;
}

SPHK_HD float riou_intersection(const ObbPair& o, RiouGeom* G, EdgeClip* E) {
    float s1, c1, sr, cr;
    sincos_f(o.a1, &s1, &c1);
    sincos_f(o.a2 - o.a1, &sr, &cr);
    sr = (sr == 0.0f) ? 1e-30f : sr;
    cr = (cr == 0.0f) ? 1e-30f : cr;
    const float dx = o.x2 - o.x1, dy = o.y2 - o.y1;
    const float px = c1 * dx + s1 * dy, py = -s1 * dx + c1 * dy;      // R(-a1) (dx,dy)
    G->px = px; G->py = py; G->cr = cr; G->sr = sr; G->c1 = c1; G->s1 = s1;
    return riou_core(px, py, cr, sr, o.w1, o.h1, o.w2, o.h2, E);
}

// IoU / IoF of an oriented box pair + the final clamp(0,1) of sph_iou_api.py:86.
SPHK_HD float riou_value(const ObbPair& o, int mode) {
    RiouGeom G;
    EdgeClip E;
    const float A1 = o.w1 * o.h1, A2 = o.w2 * o.h2;
    float I = riou_intersection(o, &G, &E);
    I = fminf(fmaxf(I, 0.0f), fminf(A1, A2));
    if (!(I == I)) I = 0.0f;
    const float den = (mode == MODE_IOF) ? A1 : (A1 + A2 - I);
    return clampf(I / den, 0.0f, 1.0f);
}

// Cheap, conservative disjointness test on the jittered OBBs: centres further apart than
// the sum of the circumradii -> intersection is exactly 0 in the reference too
// (both mmcv's kernel and diff_iou_rotated return 0 for disjoint rectangles).
SPHK_HD bool obb_disjoint(const ObbPair& o) {
    const float dx = o.x2 - o.x1, dy = o.y2 - o.y1;
    const float r1 = o.w1 * o.w1 + o.h1 * o.h1, r2 = o.w2 * o.w2 + o.h2 * o.h2;
    // (|c| > (sqrt(r1)+sqrt(r2))/2)  with a 1e-5 relative safety margin
    const float rs = 0.5f * (sqrtf(r1) + sqrtf(r2));
    return (dx * dx + dy * dy) > rs * rs * 1.00002f;
}

// ---- full Sph2Pob IoU of one pair ------------------------------------------------------
SPHK_HD float sph2pob_iou_pair(const RawBox& b1, const RawBox& b2, int D, int kind, int mode, int edge,
                               bool dense = false) {
    const bool m = jitter1_mask(b1, b2, D);
    const JitBox g = jitter1_role1(b1, m, D), p = jitter1_role2(b2, m, D);
    XformAux aux;
    ObbPair o = (kind == KIND_SPH2POB_STANDARD) ? sph2pob_standard(g, p, D, edge, &aux)
                                                : sph2pob_efficient(g, p, D, edge, &aux);
    jitter2(o);
    if (!dense && obb_disjoint(o)) return 0.0f;   // dense: measurement only (sphk_set_dense)
    return riou_value(o, mode);
}

// ---- rbb_angle = 'project' (sph2pob_efficient.py:92-93, sph2pob_standard.py:99-100) -----------------------
// The ablation variant that zeroes the x component of the box tangent before measuring the internal angle.
// It has no closed form in the tangent basis, so it is evaluated with explicit 3-D vectors, in double
// precision (the option is rare; correctness over speed), then handed to the same jitter_2 + clipper.
struct V3 { double x, y, z; };
SPHK_HD V3 v3(double x, double y, double z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
SPHK_HD V3 cross3(const V3& a, const V3& b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
SPHK_HD double dot3(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SPHK_HD V3 unit3(const V3& a) {                 // F.normalize: v / max(|v|, 1e-12)
    double n = sqrt(dot3(a, a));
    n = n < 1e-12 ? 1e-12 : n;
    return v3(a.x / n, a.y / n, a.z / n);
}
SPHK_HD double clamped_angle(const V3& a, const V3& b) {   // |acos(clamp(a^ . b^, -1 + 1e-7, 1 - 1e-7))|
    double c = dot3(unit3(a), unit3(b));
    c = c < -1.0 + 1e-7 ? -1.0 + 1e-7 : (c > 1.0 - 1e-7 ? 1.0 - 1e-7 : c);
    return fabs(acos(c));
}
SPHK_HD double turn_sign(const V3& a, const V3& b, const V3& ref) { return dot3(cross3(a, b), ref) < 0.0 ? 1.0 : -1.0; }
SPHK_HD V3 rot3(const V3 R[3], const V3& v) { return v3(dot3(R[0], v), dot3(R[1], v), dot3(R[2], v)); }

SPHK_HD ObbPair sph2pob_project(const JitBox& g, const JitBox& p, int D, int kind, int edge) {
    const double k = SPHK_PI_D / 180.0;
    const double tg = ((double)g.t_hi + (double)g.t_lo) * k, pg = ((double)g.p_hi + (double)g.p_lo) * k;
    const double tp = ((double)p.t_hi + (double)p.t_lo) * k, pp = ((double)p.p_hi + (double)p.p_lo) * k;
    V3 cg = v3(sin(pg) * cos(tg), sin(pg) * sin(tg), cos(pg)), dg = v3(cos(pg) * cos(tg), cos(pg) * sin(tg), -sin(pg));
    V3 cp = v3(sin(pp) * cos(tp), sin(pp) * sin(tp), cos(pp)), dp = v3(cos(pp) * cos(tp), cos(pp) * sin(tp), -sin(pp));
    ObbPair o;
    double a1, a2;
    if (kind == KIND_SPH2POB_STANDARD) {
        V3 R[3];
        if (fabs(cg.x - cp.x) + fabs(cg.y - cp.y) + fabs(cg.z - cp.z) > 1e-8) {          // sph2pob_standard.py:264-297
            R[0] = unit3(v3(cg.x + cp.x, cg.y + cp.y, cg.z + cp.z));
            R[1] = unit3(v3(cp.x - cg.x, cp.y - cg.y, cp.z - cg.z));
            R[2] = cross3(R[0], R[1]);
        } else {
            const double tm = 0.5 * (tg + tp), pm = 0.5 * (pg + pp);
            R[0] = v3(sin(pm) * cos(tm), sin(pm) * sin(tm), cos(pm));
            R[1] = v3(cos(pm) * cos(tm), cos(pm) * sin(tm), -sin(pm));
            R[2] = v3(sin(tm), -cos(tm), 0.0);
        }
        if (D == 5) {   // :47-54 rotate the tangent about the centre by -gamma: d cos(gamma) + e sin(gamma) in the local frame
            const double gg = (double)g.g * k, gp = (double)p.g * k;
            const V3 eg = v3(sin(tg), -cos(tg), 0.0), ep = v3(sin(tp), -cos(tp), 0.0);   // third row of the local frame
            dg = v3(dg.x * cos(gg) - eg.x * sin(gg), dg.y * cos(gg) - eg.y * sin(gg), dg.z * cos(gg) - eg.z * sin(gg));
            dp = v3(dp.x * cos(gp) - ep.x * sin(gp), dp.y * cos(gp) - ep.y * sin(gp), dp.z * cos(gp) - ep.z * sin(gp));
        }
        const V3 ex = v3(1, 0, 0), ez = v3(0, 0, 1), nez = v3(0, 0, -1);
        const V3 c1 = rot3(R, cg), c2 = rot3(R, cp);
        V3 d1 = rot3(R, dg), d2 = rot3(R, dp);
        d1.x = 0.0; d2.x = 0.0;                                                          // 'project'
        a1 = clamped_angle(d1, ez) * turn_sign(ez, d1, ex);
        a2 = clamped_angle(d2, ez) * turn_sign(ez, d2, ex);
        const V3 c1xy = v3(c1.x, c1.y, 0.0), c2xy = v3(c2.x, c2.y, 0.0);
        o.x1 = (float)(clamped_angle(c1xy, ex) * turn_sign(ex, c1xy, nez)); o.y1 = (float)clamped_angle(c1, ez);
        o.x2 = (float)(clamped_angle(c2xy, ex) * turn_sign(ex, c2xy, nez)); o.y2 = (float)clamped_angle(c2, ez);
    } else {
        const V3 z = cross3(cg, cp), ref = v3(0.5 * (cg.x + cp.x), 0.5 * (cg.y + cp.y), 0.5 * (cg.z + cp.z));
        dg.x = 0.0; dp.x = 0.0;                                                          // 'project'
        a1 = clamped_angle(dg, z) * turn_sign(z, dg, ref);
        a2 = clamped_angle(dp, z) * turn_sign(z, dp, ref);
        if (D == 5) { a1 -= (double)g.g * k; a2 -= (double)p.g * k; }
        o.x1 = 0.0f; o.y1 = 0.0f; o.x2 = (float)clamped_angle(cg, cp); o.y2 = 0.0f;
    }
    o.a1 = (float)a1; o.a2 = (float)a2;
    o.w1 = edge_len_deg(g.a, edge); o.h1 = edge_len_deg(g.b, edge);
    o.w2 = edge_len_deg(p.a, edge); o.h2 = edge_len_deg(p.b, edge);
    return o;
}

SPHK_HD float sph2pob_iou_pair_project(const RawBox& b1, const RawBox& b2, int D, int kind, int mode, int edge) {
    const bool m = jitter1_mask(b1, b2, D);
    const JitBox g = jitter1_role1(b1, m, D), p = jitter1_role2(b2, m, D);
    ObbPair o = sph2pob_project(g, p, D, kind, edge);
    jitter2(o);
    if (obb_disjoint(o)) return 0.0f;
    return riou_value(o, mode);
}

// ---- Sph2Pob-legacy (sph2pob_legacy.py:8-31 behind sph2pob_legacy_iou, sph_iou_api.py:91-92) ------------------------
// The reference's first, hand-crafted transform (BFoV only: torch.chunk(box, 4) at :52-53).  Position (:38-82): both
// centres go to the equator keeping their latitudes relative to the mid-latitude and their great-circle distance L
// (haversine), which fixes the longitude difference there.  Angle (:99-129): between the meridian tangent at the box
// centre and the one at (mid-longitude of the pair, same latitude), clamped acos, sign from a quadrant rule.  Like the
// 'project' variant it is a rarely used option: explicit formulas in double precision, then the common jitter_2 + clipper.
SPHK_HD double legacy_internal_angle(double th, double ph, double th_mid) {
    const double cp = cos(ph), sp = sin(ph);
    double c = cp * cp * cos(th - th_mid) + sp * sp;              // d(th, ph) . d(th_mid, ph), both unit
    c = c < -1.0 + 1e-7 ? -1.0 + 1e-7 : (c > 1.0 - 1e-7 ? 1.0 - 1e-7 : c);
    const double ang = fabs(acos(c));
    const double half_pi = 0.5 * SPHK_PI_D;
    const bool keep = ((th >= th_mid) && (ph < half_pi)) || ((th <= th_mid) && (ph > half_pi));      // :127-128
    return keep ? ang : -ang;
}

SPHK_HD ObbPair sph2pob_legacy(const JitBox& g, const JitBox& p, int edge) {
    const double k = SPHK_PI_D / 180.0;
    double tg = (double)g.t_hi + (double)g.t_lo, tp = (double)p.t_hi + (double)p.t_lo;          // degrees
    if (fabs(tg - tp) > 180.0) { tg = fmod(tg + 180.0, 360.0); tp = fmod(tp + 180.0, 360.0); }    // :224-244
    const double th_g = tg * k, th_p = tp * k;
    const double ph_g = ((double)g.p_hi + (double)g.p_lo) * k, ph_p = ((double)p.p_hi + (double)p.p_lo) * k;
    // 'convention' radians (:203-221): longitude theta - pi (only differences are used), latitude pi/2 - phi
    const double lat_g = 0.5 * SPHK_PI_D - ph_g, lat_p = 0.5 * SPHK_PI_D - ph_p, lat_i = 0.5 * (lat_g + lat_p);
    const double yg = lat_g - lat_i, yp = lat_p - lat_i;
    const double sh = sin(0.5 * fabs(lat_g - lat_p)), st = sin(0.5 * fabs(th_g - th_p));
    const double L = 2.0 * asin(sqrt(sh * sh + cos(lat_g) * cos(lat_p) * st * st));
    const double sL = sin(0.5 * L);
    const double dth = fabs(2.0 * asin(sqrt((sL * sL - sh * sh) / (cos(yg) * cos(yp)))));
    const double th_mid = 0.5 * (th_g + th_p);
    ObbPair o;
    o.x1 = 0.0f; o.y1 = (float)yg;
    o.x2 = (float)((th_p > th_g) ? dth : -dth); o.y2 = (float)yp;
    o.a1 = (float)legacy_internal_angle(th_g, ph_g, th_mid);
    o.a2 = (float)legacy_internal_angle(th_p, ph_p, th_mid);
    o.w1 = edge_len_deg(g.a, edge); o.h1 = edge_len_deg(g.b, edge);
    o.w2 = edge_len_deg(p.a, edge); o.h2 = edge_len_deg(p.b, edge);
    return o;
}

SPHK_HD float sph2pob_legacy_iou_pair(const RawBox& b1, const RawBox& b2, int mode, int edge) {
    const bool m = jitter1_mask(b1, b2, 4);
    const JitBox g = jitter1_role1(b1, m, 4), p = jitter1_role2(b2, m, 4);
    ObbPair o = sph2pob_legacy(g, p, edge);
    jitter2(o);
    if (obb_disjoint(o)) return 0.0f;
    return riou_value(o, mode);
}

// ---- Sph-IoU / FoV-IoU (approximate_ious.py:3-55 behind sph_iou_api.py:130-177) -------
// Everything after jitter_1: dt = theta_p - theta_g (degrees, already moved by -/+360 where |dtheta| > 180),
// pgd / ppd = the two phi (degrees), dpd = phi_p - phi_g (degrees), the four extents in degrees.
// A product that feeds a sum of two products stays a rounded product (mul_keep): which of the two the compiler would
// contract into an FMA otherwise depends on the code around the inlined copy, and the aligned and N x M kernels must
// return the same bits for the same pair.
#if defined(__CUDA_ARCH__)
SPHK_HD float mul_keep(float a, float b) { return __fmul_rn(a, b); }
#else
SPHK_HD float mul_keep(float a, float b) { return a * b; }
#endif

SPHK_HD float approx_iou_tail(float dt, float pgd, float ppd, float dpd, float gad, float gbd, float pad, float pbd,
                              int kind) {
    const float pg = kHalfPi - pgd * kDeg2Rad, pp = kHalfPi - ppd * kDeg2Rad;
    const float dphi = mul_keep(-dpd, kDeg2Rad);   // pp - pg
    const float ag = mul_keep(gad, kDeg2Rad), bg = mul_keep(gbd, kDeg2Rad), ap = mul_keep(pad, kDeg2Rad), bp = mul_keep(pbd, kDeg2Rad);
    const float dtr = mul_keep(dt, kDeg2Rad);
    // overlaps are translation invariant: measure everything relative to box 1's centre
    float lo, hi;
    if (kind == KIND_SPH) {
        lo = fmaxf(-0.5f * ag, dtr - 0.5f * ap);
        hi = fminf(0.5f * ag, dtr + 0.5f * ap);
    } else {
        const float delta = mul_keep(dtr, cosf(0.5f * (pg + pp)));
        lo = fmaxf(-0.5f * ag, delta - 0.5f * ap);
        hi = fminf(0.5f * ag, delta + 0.5f * ap);
    }
    const float plo = fmaxf(-0.5f * bg, dphi - 0.5f * bp);
    const float phi = fminf(0.5f * bg, dphi + 0.5f * bp);
    const float inter = mul_keep(fmaxf(hi - lo, 0.0f), fmaxf(phi - plo, 0.0f));
    const float iou = inter / (fmaf(ag, bg, mul_keep(ap, bp)) - inter + 1e-8f);
    return clampf(iou, 0.0f, 1.0f);
}

// The pair as jitter_1 leaves it in general: shifted where two columns are within eps, clamped at the range ends.
SPHK_HD float approx_iou_pair_general(const RawBox& b1, const RawBox& b2, int kind) {
    const bool m = jitter1_mask(b1, b2, 4);
    const JitBox g = jitter1_role1(b1, m, 4), p = jitter1_role2(b2, m, 4);
    const float dlo = p.t_lo - g.t_lo;
    float dt = (p.t_hi - g.t_hi) + dlo;                     // theta_p - theta_g, exact for close boxes
    // approximate_ious.py:60-81: where |dtheta| > 180 both thetas become (theta+180) mod 360,
    // i.e. the difference moves by -/+360.  (360 - hi) is exact, so the small result is too.
    if (dt > 180.0f) dt = -(g.t_hi + (360.0f - p.t_hi)) + dlo;
    else if (dt < -180.0f) dt = (p.t_hi + (360.0f - g.t_hi)) + dlo;
    return approx_iou_tail(dt, g.p_hi + g.p_lo, p.p_hi + p.p_lo, (p.p_hi - g.p_hi) + (p.p_lo - g.p_lo), g.a, g.b, p.a, p.b,
                           kind);
}

// True when jitter_1 is the identity on this pair: no two columns within eps (no shift) and no coordinate inside
// the clamp zone of either role (jit_coord with shift 0: below = raw < lo_off, above = (raw - end) > hi_off; both
// are monotonic in raw, so the four columns are tested through their min / max).  A NaN column fails every
// comparison in jit_coord and is skipped by fminf / fmaxf here: it passes through unclamped on both paths.
SPHK_HD bool jitter1_is_identity4(const RawBox& x, const RawBox& y) {
    const bool m = jitter1_mask(x, y, 4);
    const float xlo = fminf(fminf(x.t, x.p), fminf(x.a, x.b)), ylo = fminf(fminf(y.t, y.p), fminf(y.a, y.b));
    const float xhi = fmaxf(fmaxf(x.p, x.a), x.b), yhi = fmaxf(fmaxf(y.p, y.a), y.b);
    const bool c1 = (xlo < kEps2) | ((x.t - 360.0f) > -kEps) | ((xhi - 180.0f) > -kEps);     // role of bboxes1
    const bool c2 = (ylo < kEps) | ((y.t - 360.0f) > -kEps2) | ((yhi - 180.0f) > -kEps2);     // role of bboxes2
    return !(m | c1 | c2);
}

// Nearly every pair of a real workload is untouched by jitter_1; the hi + lo bookkeeping of the general form is
// ~110 of its ~180 instructions.  With every lo = 0 and every hi = raw the general expressions reduce to the ones
// below up to the sign of a zero, which no later operation observes: the two paths return the same bits
// (tests/test_hostsim_math.py::test_approx_identity_path_is_bit_identical).
SPHK_HD float approx_iou_pair(const RawBox& b1, const RawBox& b2, int kind) {
    if (!jitter1_is_identity4(b1, b2)) return approx_iou_pair_general(b1, b2, kind);
    float dt = b2.t - b1.t;
    if (dt > 180.0f) dt = -(b1.t + (360.0f - b2.t));
    else if (dt < -180.0f) dt = b2.t + (360.0f - b1.t);
    return approx_iou_tail(dt, b1.p, b2.p, b2.p - b1.p, b1.a, b1.b, b2.a, b2.b, kind);
}

// ---- naive_iou (sph_iou_api.py:181-198) -------------------------------------------------------------------------
// The spherical box read as a planar box of the 512 x 1024 equirectangular image (Sph2PlanarBoxTransform 'sph2pix',
// box_formator.py:79-87,176-193: x = theta / 360 * W, y = phi / 180 * H, w = alpha / 360 * W, h = beta / 180 * H) and mmcv's
// planar IoU on it: bbox_overlaps on (x1, y1, x2, y2) for BFoV, box_iou_rotated on (x, y, w, h, -gamma rad) for RBFoV.
// No jitter and no clamp (what the reference's indoor360 configs use for the test-time NMS).  mmcv-full 1.6.0 is not in
// the reference tree; its published kernels are restated: bbox_overlaps_cuda_kernel.cuh (offset 0: inter / max(union, 0))
// and box_iou_rotated_utils.hpp (0 if an area is below 1e-14, else inter / (a1 + a2 - inter)).
SPHK_HD float naive_iou_pair(const RawBox& b1, const RawBox& b2, int D, int mode) {
    const float W = 1024.0f, H = 512.0f;
    const float x1 = (b1.t / 360.0f) * W, y1 = (b1.p / 180.0f) * H, w1 = (b1.a / 360.0f) * W, h1 = (b1.b / 180.0f) * H;
    const float x2 = (b2.t / 360.0f) * W, y2 = (b2.p / 180.0f) * H, w2 = (b2.a / 360.0f) * W, h2 = (b2.b / 180.0f) * H;
    if (D == 4) {
        const float l1 = x1 - w1 / 2, t1 = y1 - h1 / 2, r1 = x1 + w1 / 2, q1 = y1 + h1 / 2;
        const float l2 = x2 - w2 / 2, t2 = y2 - h2 / 2, r2 = x2 + w2 / 2, q2 = y2 + h2 / 2;
        const float a1 = (r1 - l1) * (q1 - t1), a2 = (r2 - l2) * (q2 - t2);
        const float iw = fmaxf(fminf(r1, r2) - fmaxf(l1, l2), 0.0f), ih = fmaxf(fminf(q1, q2) - fmaxf(t1, t2), 0.0f);
        const float inter = iw * ih;
        const float base = (mode == MODE_IOF) ? fmaxf(a1, 0.0f) : fmaxf(a1 + a2 - inter, 0.0f);
        return inter / base;
    }
    ObbPair o;
    o.x1 = x1; o.y1 = y1; o.w1 = w1; o.h1 = h1; o.a1 = -b1.g * kDeg2Rad;
    o.x2 = x2; o.y2 = y2; o.w2 = w2; o.h2 = h2; o.a2 = -b2.g * kDeg2Rad;
    if (w1 * h1 < 1e-14f || w2 * h2 < 1e-14f) return 0.0f;
    if (obb_disjoint(o)) return 0.0f;
    return riou_value(o, mode);
}

// ---- unbiased_iou (sph_iou_api.py:103-125): the exact spherical IoU ------------------------------------------------
// sphdet/iou/unbiased_iou_bfov.py:10-204 (BFoV) and unbiased_iou_rbfov.py:4-181 (RBFoV), as the reference runs them: each
// box is the intersection of four hemispheres (inward normals N_left/right/up/down, rotated about the view axis by gamma
// for RBFoV); the candidate vertices of the intersection polygon are the 4 + 4 box corners and +-(N_i x N'_j) for the 16
// pairs of boundary circles of the two boxes; a candidate counts if round(V . N, 8) >= 0 for all eight normals; the
// area is sum over the counted vertices of acos(-E0 . E1) (the interior angle between the two circles meeting there)
// minus (count - 2) pi -- no ordering of the vertices is needed, and duplicates are NOT merged (the DFS clean-up is
// switched off in the reference: unbiased_iou_bfov.py:176).  All of it in double: the method subtracts O(1) angles to
// get areas of 1e-4 sr and its own validity test works at 5e-9.
struct D3 { double x, y, z; };
SPHK_HD D3 d3(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
SPHK_HD D3 d3_cross(const D3& a, const D3& b) { return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
SPHK_HD double d3_dot(const D3& a, const D3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SPHK_HD D3 d3_lin(double a, const D3& u, double b, const D3& v) { return d3(a * u.x + b * v.x, a * u.y + b * v.y, a * u.z + b * v.z); }

// normals in the reference's order: left, right, up, down.  theta, phi, fx, fy, gamma in radians.
SPHK_HD void unbiased_normals(double theta, double phi, double fx, double fy, double gamma, bool rotated, D3* N) {
    const double st = sin(theta), ct = cos(theta), sp = sin(phi), cp = cos(phi);
    const D3 look = d3(sp * ct, sp * st, cp), right = d3(-st, ct, 0.0), up = d3(-cp * ct, -cp * st, sp);
    const double sx = sin(0.5 * fx), cx = cos(0.5 * fx), sy = sin(0.5 * fy), cy = cos(0.5 * fy);
    N[0] = d3_lin(-cx, right, sx, look);
    N[1] = d3_lin(cx, right, sx, look);
    N[2] = d3_lin(-cy, up, sy, look);
    N[3] = d3_lin(cy, up, sy, look);
    if (rotated) {      // roll_T (unbiased_iou_rbfov.py:10-33): Rodrigues rotation about the view axis by gamma
        const double cg = cos(gamma), sg = sin(gamma), k = 1.0 - cg;
        const double nx = look.x, ny = look.y, nz = look.z;
        const double m11 = nx * nx * k + cg, m12 = nx * ny * k - nz * sg, m13 = nx * nz * k + ny * sg;
        const double m21 = nx * ny * k + nz * sg, m22 = ny * ny * k + cg, m23 = ny * nz * k - nx * sg;
        const double m31 = nx * nz * k - ny * sg, m32 = ny * nz * k + nx * sg, m33 = nz * nz * k + cg;
        for (int i = 0; i < 4; ++i) {
            const D3 v = N[i];
            N[i] = d3(m11 * v.x + m12 * v.y + m13 * v.z, m21 * v.x + m22 * v.y + m23 * v.z, m31 * v.x + m32 * v.y + m33 * v.z);
        }
    }
}

// A candidate V = c / (|c| + delta) (c = E0 x E1; delta = 0 for a box corner, 1e-10 for a crossing) and its antipode -V.
// The reference counts V iff round(V . N_k, 8) >= 0 for the eight normals, i.e. rint(1e8 V . N_k) >= 0, i.e.
// 1e8 V . N_k >= -0.5 (half-to-even: -0.5 rounds to -0, which passes), i.e. with t_k = c . N_k
//     t_k >= -0.5e-8 (|c| + delta)        for all k;
// c -> -c negates every product exactly, so -V counts iff t_k <= +0.5e-8 (|c| + delta) for all k.  The same t_k serve
// both, and neither the division nor the normalised vector is needed.  (t_k / (|c| + delta) and the reference's
// fl(V) . N_k differ by rounding at 1e-16 relative, which matters only if a product sits within 1e-15 of the threshold;
// where boundary circles are parallel to 1e-7 rad the decision is noise in the reference too -- its sin / cos differ from
// CUDA's in the last place.)  Nearly all candidates fail by a wide margin, which the squares show without the square
// root: t_k < 0 and t_k^2 > 1e-16 |c|^2 + 2e-36 implies -t_k > 0.7e-8 |c| + 1e-18 > 0.5e-8 (|c| + 1e-10).
// A[4], B[4]: the normals of the two boxes (any order: the test is an AND over all eight).
// Returns bit 0: V counts, bit 1: -V counts.
// (The comparisons are written as sign bits of a sum, OR-ed together as integers: a chain of `x < y` on doubles against
// one threshold is turned by the compiler into a running fmin / fmax, which costs ~9 instructions per element on sm_100a.)
SPHK_HD int d_hi(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    long long b;
    memcpy(&b, &x, sizeof b);
    return (int)(b >> 32);
#endif
}

SPHK_HD int unbiased_candidate(const D3& c, double delta, const D3* A, const D3* B, bool want_neg) {
    const double n2 = d3_dot(c, c), thr = 1e-16 * n2 + 2e-36;
    double t[8];
    int below = 0, above = 0;               // sign bit set: some t_k |t_k| < -thr  /  some t_k |t_k| > thr
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        t[k] = d3_dot(c, A[k]);
        t[4 + k] = d3_dot(c, B[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        below |= d_hi(fma(t[k], fabs(t[k]), thr));
        above |= d_hi(fma(-t[k], fabs(t[k]), thr));
    }
    bool pos = below >= 0, neg = want_neg && above >= 0;
    if (!(pos || neg)) return 0;
    const double bound = 0.5e-8 * (sqrt(n2) + delta);
    below = 0; above = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        below |= d_hi(t[k] + bound);        // t_k < -bound
        above |= d_hi(bound - t[k]);        // t_k >  bound
    }
    pos = pos && below >= 0;
    neg = neg && above >= 0;
    return (pos ? 1 : 0) | (neg ? 2 : 0);
}

SPHK_HD void d3_rotate4(D3* Q) { const D3 t = Q[0]; Q[0] = Q[1]; Q[1] = Q[2]; Q[2] = Q[3]; Q[3] = t; }

SPHK_HD double unbiased_angle(double minus_cos) {      // interArea: arccos(clip(-E0 . E1, -1, 1))
    return acos(minus_cos < -1.0 ? -1.0 : (minus_cos > 1.0 ? 1.0 : minus_cos));
}

SPHK_HD float unbiased_iou_pair(const RawBox& b1, const RawBox& b2, int D) {
    // jiter_spherical_bboxes (sph_iou_api.py:244-260) in double, on the float32 box values: with nearly coincident
    // boundary circles the 5e-9 validity test decides which vertices count, so the jittered values must be those of the
    // reference's float64 run to the last bit, not the fp32-rounded ones of JitBox
    const double eps = SPHK_EPS_D;
    const double u[5] = {b1.t, b1.p, b1.a, b1.b, b1.g}, v[5] = {b2.t, b2.p, b2.a, b2.b, b2.g};
    bool m = false;
#pragma unroll
    for (int k = 0; k < 5; ++k) m = m || (k < D && fabs(u[k] - v[k]) < eps);
    double x[5], y[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) { x[k] = (k < D) ? (m ? u[k] - 2.0 * eps : u[k]) : 0.0; y[k] = (k < D) ? (m ? v[k] + eps : v[k]) : 0.0; }
    x[0] = fmin(fmax(x[0], 2.0 * eps), 360.0 - eps);
    y[0] = fmin(fmax(y[0], eps), 360.0 - 2.0 * eps);
#pragma unroll
    for (int k = 1; k < 4; ++k) { x[k] = fmin(fmax(x[k], 2.0 * eps), 180.0 - eps); y[k] = fmin(fmax(y[k], eps), 180.0 - 2.0 * eps); }
    if (D == 5) y[4] = fmin(fmax(y[4], -360.0 + 2.0 * eps), 360.0 - 2.0 * eps);       // both clamps of :257-258 hit bboxes2
    const double d2r = SPHK_PI_D / 180.0;
    const double fx1 = x[2] * d2r, fy1 = x[3] * d2r, fx2 = y[2] * d2r, fy2 = y[3] * d2r;
    // The interior angle at a box's own corner is the same at all four: -(N_a . N_b) = -sin(fx/2) sin(fy/2) for the four
    // pairs of getNormal (roll_T is a rotation), and the box area is 4 of them minus 2 pi (Sph.area).
    const double corner1 = unbiased_angle(-sin(0.5 * fx1) * sin(0.5 * fy1)), corner2 = unbiased_angle(-sin(0.5 * fx2) * sin(0.5 * fy2));
    const double a1 = 4.0 * corner1 - 2.0 * SPHK_PI_D, a2 = 4.0 * corner2 - 2.0 * SPHK_PI_D;
    // normals kept in cyclic order (left, up, right, down) so that corner k is Q[k] x Q[k+1]: the reference's four
    // (left x up, down x left, up x right, right x down).  The candidate loops are NOT unrolled -- each iteration works on
    // element 0 (and 1) and then rotates the array, so the arrays stay in registers with one copy of the test in the
    // instruction stream (the unrolled form was 24 copies and stalled on instruction fetch).
    D3 A[4], B[4];
    {
        D3 N[4];
        unbiased_normals(x[0] * d2r, x[1] * d2r, fx1, fy1, x[4] * d2r, D == 5, N);
        A[0] = N[0]; A[1] = N[2]; A[2] = N[1]; A[3] = N[3];
        unbiased_normals(y[0] * d2r, y[1] * d2r, fx2, fy2, y[4] * d2r, D == 5, N);
        B[0] = N[0]; B[1] = N[2]; B[2] = N[1]; B[3] = N[3];
    }
    double sum = 0.0;
    int count = 0;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {                       // corners of box 1
        if (unbiased_candidate(d3_cross(A[0], A[1]), 0.0, A, B, false)) { sum += corner1; count += 1; }
        d3_rotate4(A);
    }
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {                       // corners of box 2
        if (unbiased_candidate(d3_cross(B[0], B[1]), 0.0, A, B, false)) { sum += corner2; count += 1; }
        d3_rotate4(B);
    }
    // crossings of a boundary circle of box 1 with one of box 2: +-(N_i x N'_j) / (|.| + 1e-10)
#pragma unroll 1
    for (int ij = 0; ij < 16; ++ij) {
        const int hit = unbiased_candidate(d3_cross(A[0], B[0]), 1e-10, A, B, true);
        if (hit) {
            const double ang = unbiased_angle(-d3_dot(A[0], B[0]));
            sum += (hit == 3) ? 2.0 * ang : ang;
            count += (hit == 3) ? 2 : 1;
        }
        d3_rotate4(B);
        if ((ij & 3) == 3) d3_rotate4(A);
    }
    const double inter = (count == 0) ? 0.0 : sum - (double)(count - 2) * SPHK_PI_D;
    // the two files end differently (unbiased_iou_bfov.py:199 / unbiased_iou_rbfov.py:175)
    const double iou = (D == 4) ? (inter + 1e-8) / (a1 + a2 - (inter + 1e-8)) : inter / (a1 + a2 - inter + 1e-8);
    return clampf((float)iou, 0.0f, 1.0f);        // .float() then clamp(0, 1) (sph_iou_api.py:125)
}

}  // namespace sphk
