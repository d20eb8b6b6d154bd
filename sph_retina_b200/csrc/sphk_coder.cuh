// sphk_coder.cuh -- the spherical delta box coders of the reference, per box, in registers:
//   sphdet/bbox/coder/delta_xywh_sph_bbox_coder.py:117-262    (BFoV,  D = 4)
//   sphdet/bbox/coder/delta_xywha_rsph_bbox_coder.py:117-268  (RBFoV, D = 5)
// decode (delta2bbox) with every clamp of the reference and a record of which clamps were active (the backward
// pass needs it: torch's clamp passes the gradient on the closed interval and zeroes it outside), its chain rule,
// and encode (bbox2delta).  Compiles for the device and, for the CPU test-suite, for the host (tests/hostsim).
#pragma once
#include "sphk_math.cuh"

namespace sphk {

struct CoderParams {
    float mean[5], stdv[5];
    float max_ratio;        // |log(wh_ratio_clip)|
    float ctr_clamp;        // add_ctr_clamp: |centre shift| <= ctr_clamp, size deltas clamped from above only
    int clip_border, add_ctr_clamp;
};

constexpr float kCoderEps = 1e-7f;
constexpr float kRad2Deg = (float)(180.0 / SPHK_PI_D);

SPHK_HD float exp_f(float x) {
#if defined(__CUDA_ARCH__)
    return expf(x);
#else
    return (float)exp((double)x);
#endif
}

// box = delta2bbox(roi, delta).  pass bit k set: d(box[k]) / d(delta[k]) is not cut by a clamp.
// jac[k] = d(box[k]) / d(delta[k]) where it passes (the coder is diagonal: component k depends on delta[k] only).
SPHK_HD RawBox coder_decode(const RawBox& roi, const float* delta, int D, const CoderParams& cp, uint32_t* pass, float* jac) {
    float d[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) d[k] = (k < D) ? fmaf(delta[k], cp.stdv[k], cp.mean[k]) : 0.0f;
    uint32_t ps = 0x1Fu;
    float sx = roi.a * d[0], sy = roi.b * d[1];
    float dw = d[2], dh = d[3];
    if (cp.add_ctr_clamp) {
        if (!(sx >= -cp.ctr_clamp && sx <= cp.ctr_clamp)) ps &= ~1u;
        if (!(sy >= -cp.ctr_clamp && sy <= cp.ctr_clamp)) ps &= ~2u;
        sx = clampf(sx, -cp.ctr_clamp, cp.ctr_clamp); sy = clampf(sy, -cp.ctr_clamp, cp.ctr_clamp);
        if (!(dw <= cp.max_ratio)) ps &= ~4u;
        if (!(dh <= cp.max_ratio)) ps &= ~8u;
        dw = fminf(dw, cp.max_ratio); dh = fminf(dh, cp.max_ratio);
    } else {
        if (!(dw >= -cp.max_ratio && dw <= cp.max_ratio)) ps &= ~4u;
        if (!(dh >= -cp.max_ratio && dh <= cp.max_ratio)) ps &= ~8u;
        dw = clampf(dw, -cp.max_ratio, cp.max_ratio); dh = clampf(dh, -cp.max_ratio, cp.max_ratio);
    }
    RawBox b;
    b.t = roi.t + sx; b.p = roi.p + sy;
    b.a = roi.a * exp_f(dw); b.b = roi.b * exp_f(dh);
    b.g = (D == 5) ? roi.g + kRad2Deg * d[4] : 0.0f;
    jac[0] = roi.a * cp.stdv[0]; jac[1] = roi.b * cp.stdv[1];
    jac[2] = b.a * cp.stdv[2]; jac[3] = b.b * cp.stdv[3];
    jac[4] = (D == 5) ? kRad2Deg * cp.stdv[4] : 0.0f;
    if (cp.clip_border) {
        const float lo = kCoderEps, t_hi = 360.0f - kCoderEps, p_hi = 180.0f - kCoderEps;     // fp32: 360, 180
        if (!(b.t >= lo && b.t <= t_hi)) ps &= ~1u;
        if (!(b.p >= lo && b.p <= p_hi)) ps &= ~2u;
        if (!(b.a >= lo && b.a <= p_hi)) ps &= ~4u;
        if (!(b.b >= lo && b.b <= p_hi)) ps &= ~8u;
        b.t = clampf(b.t, lo, t_hi); b.p = clampf(b.p, lo, p_hi);
        b.a = clampf(b.a, lo, p_hi); b.b = clampf(b.b, lo, p_hi);
        if (D == 5) {
            const float g_lo = -90.0f + kCoderEps, g_hi = 90.0f - kCoderEps;
            if (!(b.g >= g_lo && b.g <= g_hi)) ps &= ~16u;
            b.g = clampf(b.g, g_lo, g_hi);
        }
    }
    *pass = ps;
    return b;
}

// d(loss)/d(delta) from d(loss)/d(box)
SPHK_HD void coder_decode_grad(const float* gbox, uint32_t pass, const float* jac, int D, float* gdelta) {
#pragma unroll
    for (int k = 0; k < 5; ++k) gdelta[k] = (k < D && ((pass >> k) & 1u)) ? gbox[k] * jac[k] : 0.0f;
}

// delta = bbox2delta(proposal, gt)
SPHK_HD void coder_encode(const RawBox& pr, const RawBox& gt, int D, const CoderParams& cp, float* delta) {
    const float pw = fmaxf(pr.a, kCoderEps), ph = fmaxf(pr.b, kCoderEps);
    const float gw = fmaxf(gt.a, kCoderEps), gh = fmaxf(gt.b, kCoderEps);
    float d[5];
    d[0] = (gt.t - pr.t) / pw; d[1] = (gt.p - pr.p) / ph;
    d[2] = logf(gw / pw); d[3] = logf(gh / ph);
    d[4] = (D == 5) ? (gt.g - pr.g) * kDeg2Rad : 0.0f;
#pragma unroll
    for (int k = 0; k < 5; ++k) delta[k] = (k < D) ? (d[k] - cp.mean[k]) / cp.stdv[k] : 0.0f;
}

}  // namespace sphk
