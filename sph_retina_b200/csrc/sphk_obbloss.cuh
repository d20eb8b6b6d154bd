// sphk_obbloss.cuh -- the other regression losses the reference hangs on the Sph2Pob OBBs:
//
//   Sph2PobGDLoss  sphdet/losses/sph2pob_gd_loss.py:7-26  = Sph2PobTransfrom()(mmrotate GDLoss): gwd / kld / jd /
//                  kld_symmax / kld_symmin on the 2-D Gaussians of the two planar boxes
//   Sph2PobKFLoss  sphdet/losses/sph2pob_kf_loss.py:8-26  = Sph2PobTransfrom()(mmrotate KFLoss), called with
//                  pred_decode=target, targets_decode=pred (:26 -- the swap is the reference's and is kept)
//   Sph2PobL1Loss  sphdet/losses/sph2pob_l1_loss.py:9-94  = L1 on bbox2delta(pred OBB, target OBB) (or on the OBBs)
//
// GDLoss / KFLoss live in mmrotate 0.3.2 (README.md:95), which is NOT part of the reference tree: the formulas below
// restate its published models/losses/gaussian_dist_loss.py (xy_wh_r_2_xy_sigma, gwd_loss, kld_loss, jd_loss,
// kld_symmax_loss, kld_symmin_loss, postprocess) and models/losses/kf_iou_loss.py (kfiou_loss), clamps included.
//
// One row = one (pred, target) pair.  The loss of a row is a short scalar program in the ten OBB parameters
// (x, y, w, h, a of both boxes); it is evaluated ONCE on forward-mode dual numbers carrying all ten partials, so the
// value and d(loss)/d(OBBs) come out of the same registers and the clamps / max / where / abs take exactly the
// autograd sub-gradients torch uses (clamp: passes where lo <= x <= hi; abs: sign, 0 at 0; max/min: split on ties;
// where(isnan): 0).  The caller chains d/d(OBBs) through jitter_2 and the transform (sphk_grad.cuh) in the same thread.
#pragma once
#include "sphk_math.cuh"

namespace sphk {

enum LossKind { LOSS_GWD = 0, LOSS_KLD = 1, LOSS_JD = 2, LOSS_KLD_SYMMAX = 3, LOSS_KLD_SYMMIN = 4, LOSS_KFIOU = 5, LOSS_L1 = 6 };
enum LossFun { FUN_NONE = 0, FUN_LOG1P = 1, FUN_SQRT = 2 };          // GDLoss fun;  KFLoss: 0 none, 1 ln, 2 exp
enum LossFlag { LF_NORMALIZE_OR_SQRT = 1,                           // gwd: normalize;  kld family: sqrt
                LF_L1_ENCODE = 1, LF_L1_SWAP = 2, LF_L1_MODULUS = 4 };

struct LossParams {
    int kind, fun, flags;
    float tau, alpha;    // GDLoss
    float beta, eps;     // KFLoss (beta = 1/9, eps = 1e-6)
};

SPHK_HD int loss_columns(int kind) { return kind == LOSS_L1 ? 5 : 1; }

// ---- forward-mode dual numbers --------------------------------------------------------------------------------
constexpr int kNV = 10;   // x1 y1 w1 h1 a1 x2 y2 w2 h2 a2
template <typename T>
struct Dual {
    T v;
    T d[kNV];
};

template <typename T> SPHK_HD Dual<T> dconst(T c) {
    Dual<T> r; r.v = c;
#pragma unroll
    for (int k = 0; k < kNV; ++k) r.d[k] = (T)0;
    return r;
}
template <typename T> SPHK_HD Dual<T> dvar(T x, int idx) {
    Dual<T> r = dconst<T>(x);
    r.d[idx] = (T)1;
    return r;
}
// r = f(a) with f'(a) = fa
template <typename T> SPHK_HD Dual<T> dchain(const Dual<T>& a, T fv, T fa) {
    Dual<T> r; r.v = fv;
#pragma unroll
    for (int k = 0; k < kNV; ++k) r.d[k] = fa * a.d[k];
    return r;
}
// r = f(a, b)
template <typename T> SPHK_HD Dual<T> dchain2(const Dual<T>& a, const Dual<T>& b, T fv, T fa, T fb) {
    Dual<T> r; r.v = fv;
#pragma unroll
    for (int k = 0; k < kNV; ++k) r.d[k] = fa * a.d[k] + fb * b.d[k];
    return r;
}
template <typename T> SPHK_HD Dual<T> operator+(const Dual<T>& a, const Dual<T>& b) { return dchain2(a, b, a.v + b.v, (T)1, (T)1); }
template <typename T> SPHK_HD Dual<T> operator-(const Dual<T>& a, const Dual<T>& b) { return dchain2(a, b, a.v - b.v, (T)1, (T)-1); }
template <typename T> SPHK_HD Dual<T> operator*(const Dual<T>& a, const Dual<T>& b) { return dchain2(a, b, a.v * b.v, b.v, a.v); }
template <typename T> SPHK_HD Dual<T> operator/(const Dual<T>& a, const Dual<T>& b) {
    const T q = a.v / b.v;
    return dchain2(a, b, q, (T)1 / b.v, -q / b.v);
}
template <typename T> SPHK_HD Dual<T> operator-(const Dual<T>& a) { return dchain(a, -a.v, (T)-1); }
template <typename T> SPHK_HD Dual<T> operator+(const Dual<T>& a, T c) { return dchain(a, a.v + c, (T)1); }
template <typename T> SPHK_HD Dual<T> operator-(const Dual<T>& a, T c) { return dchain(a, a.v - c, (T)1); }
template <typename T> SPHK_HD Dual<T> operator*(const Dual<T>& a, T c) { return dchain(a, a.v * c, c); }
template <typename T> SPHK_HD Dual<T> operator*(T c, const Dual<T>& a) { return dchain(a, a.v * c, c); }
template <typename T> SPHK_HD Dual<T> operator/(const Dual<T>& a, T c) { return dchain(a, a.v / c, (T)1 / c); }
template <typename T> SPHK_HD Dual<T> operator-(T c, const Dual<T>& a) { return dchain(a, c - a.v, (T)-1); }
template <typename T> SPHK_HD Dual<T> operator/(T c, const Dual<T>& a) {
    const T q = c / a.v;
    return dchain(a, q, -q / a.v);
}

SPHK_HD float m_sqrt(float x) { return sqrtf(x); }     SPHK_HD double m_sqrt(double x) { return sqrt(x); }
SPHK_HD float m_log(float x) { return logf(x); }       SPHK_HD double m_log(double x) { return log(x); }
SPHK_HD float m_log1p(float x) { return log1pf(x); }   SPHK_HD double m_log1p(double x) { return log1p(x); }
SPHK_HD float m_exp(float x) { return expf(x); }       SPHK_HD double m_exp(double x) { return exp(x); }
SPHK_HD float m_sin(float x) { return sinf(x); }       SPHK_HD double m_sin(double x) { return sin(x); }
SPHK_HD float m_cos(float x) { return cosf(x); }       SPHK_HD double m_cos(double x) { return cos(x); }
SPHK_HD float m_floor(float x) { return floorf(x); }   SPHK_HD double m_floor(double x) { return floor(x); }

template <typename T> SPHK_HD Dual<T> dsqrt(const Dual<T>& a) { const T s = m_sqrt(a.v); return dchain(a, s, (T)0.5 / s); }
template <typename T> SPHK_HD Dual<T> dlog(const Dual<T>& a) { return dchain(a, m_log(a.v), (T)1 / a.v); }
template <typename T> SPHK_HD Dual<T> dlog1p(const Dual<T>& a) { return dchain(a, m_log1p(a.v), (T)1 / ((T)1 + a.v)); }
template <typename T> SPHK_HD Dual<T> dexp(const Dual<T>& a) { const T e = m_exp(a.v); return dchain(a, e, e); }
template <typename T> SPHK_HD Dual<T> dabs(const Dual<T>& a) {
    const T av = a.v;
    return dchain(a, (av < (T)0) ? -av : av, (av > (T)0) ? (T)1 : ((av < (T)0) ? (T)-1 : (T)0));
}
// torch.clamp(min=lo[, max=hi]): NaN stays NaN; the gradient passes where lo <= x <= hi
template <typename T> SPHK_HD Dual<T> dclamp_min(const Dual<T>& a, T lo) {
    return dchain(a, (a.v < lo) ? lo : a.v, (a.v >= lo) ? (T)1 : (T)0);
}
template <typename T> SPHK_HD Dual<T> dclamp(const Dual<T>& a, T lo, T hi) {
    return dchain(a, (a.v < lo) ? lo : ((a.v > hi) ? hi : a.v), (a.v >= lo && a.v <= hi) ? (T)1 : (T)0);
}
// torch.max / torch.min of two tensors: equal values share the gradient
template <typename T> SPHK_HD Dual<T> dmax(const Dual<T>& a, const Dual<T>& b) {
    const T wa = (a.v > b.v) ? (T)1 : ((a.v == b.v) ? (T)0.5 : (T)0);
    return dchain2(a, b, (a.v > b.v) ? a.v : b.v, wa, (T)1 - wa);
}
template <typename T> SPHK_HD Dual<T> dmin(const Dual<T>& a, const Dual<T>& b) {
    const T wa = (a.v < b.v) ? (T)1 : ((a.v == b.v) ? (T)0.5 : (T)0);
    return dchain2(a, b, (a.v < b.v) ? a.v : b.v, wa, (T)1 - wa);
}

// ---- mmrotate gaussian_dist_loss.py: xy_wh_r_2_xy_sigma -----------------------------------------------------------
template <typename T>
struct Sigma2 { Dual<T> s11, s12, s22; };

template <typename T>
SPHK_HD Sigma2<T> obb_sigma(const Dual<T>& w, const Dual<T>& h, const Dual<T>& r) {
    const Dual<T> wc = dclamp(w, (T)1e-7, (T)1e7), hc = dclamp(h, (T)1e-7, (T)1e7);
    const Dual<T> a = wc * wc * (T)0.25, b = hc * hc * (T)0.25;          // (0.5 diag(wh))^2
    const T cv = m_cos(r.v), sv = m_sin(r.v);
    const Dual<T> c = dchain(r, cv, -sv), s = dchain(r, sv, cv);
    Sigma2<T> S;
    S.s11 = c * c * a + s * s * b;
    S.s12 = c * s * (a - b);
    S.s22 = s * s * a + c * c * b;
    return S;
}
template <typename T> SPHK_HD Dual<T> sigma_det(const Sigma2<T>& S) { return S.s11 * S.s22 - S.s12 * S.s12; }

template <typename T>
SPHK_HD Dual<T> gd_postprocess(Dual<T> distance, int fun, T tau) {
    if (fun == FUN_LOG1P) distance = dlog1p(distance);
    else if (fun == FUN_SQRT) distance = dsqrt(dclamp_min(distance, (T)1e-7));
    if (tau >= (T)1) return (T)1 - (T)1 / (distance + tau);
    return distance;
}

template <typename T>
SPHK_HD Dual<T> gwd_distance(const Dual<T>& dx, const Dual<T>& dy, const Sigma2<T>& P, const Sigma2<T>& Q, T alpha, bool normalize) {
    const Dual<T> xy = dx * dx + dy * dy;
    Dual<T> whr = (P.s11 + P.s22) + (Q.s11 + Q.s22);
    const Dual<T> t_tr = P.s11 * Q.s11 + (T)2 * (P.s12 * Q.s12) + P.s22 * Q.s22;       // trace(Sigma_p Sigma_t)
    const Dual<T> t_det_sqrt = dsqrt(dclamp_min(sigma_det(P) * sigma_det(Q), (T)1e-7));
    whr = whr + (T)-2 * dsqrt(dclamp_min(t_tr + (T)2 * t_det_sqrt, (T)1e-7));
    Dual<T> distance = dsqrt(dclamp_min(xy + (alpha * alpha) * whr, (T)1e-7));
    if (normalize) {
        const Dual<T> scale = (T)2 * dclamp_min(dsqrt(dclamp_min(dsqrt(dclamp_min(t_det_sqrt, (T)1e-7)), (T)1e-7)), (T)1e-7);
        distance = distance / scale;
    }
    return distance;
}

// KL(N_p || N_t) as mmrotate's kld_loss writes it (the inverse is taken of Sigma_p)
template <typename T>
SPHK_HD Dual<T> kld_distance(const Dual<T>& dx, const Dual<T>& dy, const Sigma2<T>& P, const Sigma2<T>& Q, T alpha, bool sqrt_) {
    const Dual<T> det_p = sigma_det(P), det_t = sigma_det(Q);
    const Dual<T> i11 = P.s22 / det_p, i12 = -P.s12 / det_p, i22 = P.s11 / det_p;
    const Dual<T> xy = (T)0.5 * (dx * dx * i11 + (T)2 * (dx * dy * i12) + dy * dy * i22);
    Dual<T> whr = (T)0.5 * (i11 * Q.s11 + (T)2 * (i12 * Q.s12) + i22 * Q.s22);
    whr = whr + (T)0.5 * (dlog(det_p) - dlog(det_t));
    whr = whr - (T)1;
    Dual<T> distance = xy / (alpha * alpha) + whr;
    if (sqrt_) distance = dsqrt(dclamp_min(distance, (T)1e-7));
    return distance;
}

template <typename T>
SPHK_HD Dual<T> gd_loss_row(const Dual<T>* q, const LossParams& P) {
    // q[0..4] = pred OBB (x y w h a), q[5..9] = target OBB
    const Sigma2<T> Sp = obb_sigma(q[2], q[3], q[4]), St = obb_sigma(q[7], q[8], q[9]);
    const Dual<T> dx = q[0] - q[5], dy = q[1] - q[6];
    const T alpha = (T)P.alpha;
    const bool opt = (P.flags & LF_NORMALIZE_OR_SQRT) != 0;
    Dual<T> dist;
    if (P.kind == LOSS_GWD) {
        dist = gwd_distance(dx, dy, Sp, St, alpha, opt);
    } else if (P.kind == LOSS_KLD) {
        dist = kld_distance(dx, dy, Sp, St, alpha, opt);
    } else {
        const Dual<T> mdx = -dx, mdy = -dy;
        if (P.kind == LOSS_JD) {
            dist = (kld_distance(dx, dy, Sp, St, alpha, false) + kld_distance(mdx, mdy, St, Sp, alpha, false)) * (T)0.5;
            if (opt) dist = dsqrt(dclamp_min(dist, (T)1e-7));
        } else {
            const Dual<T> a = kld_distance(dx, dy, Sp, St, alpha, opt), b = kld_distance(mdx, mdy, St, Sp, alpha, opt);
            dist = (P.kind == LOSS_KLD_SYMMAX) ? dmax(a, b) : dmin(a, b);
        }
    }
    return gd_postprocess(dist, P.fun, (T)P.tau);
}

// mmrotate kf_iou_loss.py: kfiou_loss(pred, target, pred_decode, targets_decode).  The reference's subclass passes
// pred_decode = TARGET OBB and targets_decode = PRED OBB (sph2pob_kf_loss.py:26).
template <typename T>
SPHK_HD Dual<T> kf_loss_row(const Dual<T>* q, const LossParams& P) {
    const T beta = (T)P.beta, eps = (T)P.eps;
    Dual<T> xy_loss = dconst<T>((T)0);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const Dual<T> diff = dabs(q[k] - q[5 + k]);
        xy_loss = xy_loss + ((diff.v < beta) ? ((T)0.5 * diff * diff / beta) : (diff - (T)0.5 * beta));
    }
    const Sigma2<T> Sp = obb_sigma(q[7], q[8], q[9]);      // "pred_decode"    = target OBB
    const Sigma2<T> St = obb_sigma(q[2], q[3], q[4]);      // "targets_decode" = pred OBB
    const Dual<T> Vb_p = (T)4 * dsqrt(sigma_det(Sp)), Vb_t = (T)4 * dsqrt(sigma_det(St));
    // K = Sigma_p (Sigma_p + Sigma_t)^-1 ;  Sigma = Sigma_p - K Sigma_p
    const Dual<T> m11 = Sp.s11 + St.s11, m12 = Sp.s12 + St.s12, m22 = Sp.s22 + St.s22;
    const Dual<T> md = m11 * m22 - m12 * m12;
    const Dual<T> n11 = m22 / md, n12 = -m12 / md, n22 = m11 / md;
    const Dual<T> k11 = Sp.s11 * n11 + Sp.s12 * n12, k12 = Sp.s11 * n12 + Sp.s12 * n22;
    const Dual<T> k21 = Sp.s12 * n11 + Sp.s22 * n12, k22 = Sp.s12 * n12 + Sp.s22 * n22;
    const Dual<T> e11 = Sp.s11 - (k11 * Sp.s11 + k12 * Sp.s12), e12 = Sp.s12 - (k11 * Sp.s12 + k12 * Sp.s22);
    const Dual<T> e21 = Sp.s12 - (k21 * Sp.s11 + k22 * Sp.s12), e22 = Sp.s22 - (k21 * Sp.s12 + k22 * Sp.s22);
    Dual<T> Vb = (T)4 * dsqrt(e11 * e22 - e12 * e21);
    if (!(Vb.v == Vb.v)) Vb = dconst<T>((T)0);             // torch.where(isnan(Vb), 0, Vb)
    const Dual<T> kfiou = Vb / (Vb_p + Vb_t - Vb + eps);
    Dual<T> kf;
    if (P.fun == 1) kf = -dlog(kfiou + eps);
    else if (P.fun == 2) kf = dexp((T)1 - kfiou) - (T)1;
    else kf = (T)1 - kfiou;
    return dclamp_min(xy_loss + kf, (T)0);
}

// One row of a GD / KF loss: returns the loss, adds up * d(loss)/d(OBBs) to go1 / go2 (x y w h a each).
template <typename T>
SPHK_HD float obb_scalar_loss_row(const ObbPair& o, const LossParams& P, float up, float* go1, float* go2) {
    Dual<T> q[kNV];
    const float raw[kNV] = {o.x1, o.y1, o.w1, o.h1, o.a1, o.x2, o.y2, o.w2, o.h2, o.a2};
#pragma unroll
    for (int k = 0; k < kNV; ++k) q[k] = dvar<T>((T)raw[k], k);
    const Dual<T> L = (P.kind == LOSS_KFIOU) ? kf_loss_row<T>(q, P) : gd_loss_row<T>(q, P);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        go1[k] += up * (float)L.d[k];
        go2[k] += up * (float)L.d[5 + k];
    }
    return (float)L.v;
}

// Sph2PobL1Loss (sph2pob_l1_loss.py:27-87): five columns per row.  ell[5] receives |delta_j| (encode) or
// |pred_j - target_j|; go1 / go2 accumulate sum_j up[j] * d(ell_j)/d(OBBs).  All arithmetic in fp32, as the
// reference's bbox2delta forces with .float() (:68-69).
SPHK_HD float l1_sign(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }
SPHK_HD float l1_wrap(float a, bool modulus) {
    if (!modulus) return a;
    const float s = a + kPi;                  // (angle + pi) % pi, python modulo (sign of the divisor)
    return s - kPi * floorf(s / kPi);
}
SPHK_HD void obb_l1_loss_row(const ObbPair& o, const LossParams& P, const float* up, float* ell, float* go1, float* go2) {
    const float a[5] = {o.x1, o.y1, o.w1, o.h1, o.a1}, b[5] = {o.x2, o.y2, o.w2, o.h2, o.a2};
    if (!(P.flags & LF_L1_ENCODE)) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const float d = a[k] - b[k];
            ell[k] = fabsf(d);
            go1[k] += up[k] * l1_sign(d);
            go2[k] -= up[k] * l1_sign(d);
        }
        return;
    }
    // bbox2delta(proposals, gt): proposals = pred OBB unless swap
    const bool swap = (P.flags & LF_L1_SWAP) != 0;
    const float* p = swap ? b : a;
    const float* g = swap ? a : b;
    float* gp = swap ? go2 : go1;
    float* gg = swap ? go1 : go2;
    const bool pw_pass = p[2] >= 1e-7f, ph_pass = p[3] >= 1e-7f, gw_pass = g[2] >= 1e-7f, gh_pass = g[3] >= 1e-7f;
    const float pw = fmaxf(p[2], 1e-7f), ph = fmaxf(p[3], 1e-7f), gw = fmaxf(g[2], 1e-7f), gh = fmaxf(g[3], 1e-7f);
    const float dx = (g[0] - p[0]) / pw, dy = (g[1] - p[1]) / ph;
    const float dw = logf(gw / pw), dh = logf(gh / ph);
    const bool modulus = (P.flags & LF_L1_MODULUS) != 0;
    const float da = (l1_wrap(g[4], modulus) - l1_wrap(p[4], modulus)) / kPi;
    ell[0] = fabsf(dx); ell[1] = fabsf(dy); ell[2] = fabsf(dw); ell[3] = fabsf(dh); ell[4] = fabsf(da);
    const float sx = up[0] * l1_sign(dx), sy = up[1] * l1_sign(dy), sw = up[2] * l1_sign(dw), sh = up[3] * l1_sign(dh),
                sa = up[4] * l1_sign(da);
    gg[0] += sx / pw;  gp[0] -= sx / pw;
    gg[1] += sy / ph;  gp[1] -= sy / ph;
    if (pw_pass) gp[2] += -sx * dx / pw - sw / pw;
    if (ph_pass) gp[3] += -sy * dy / ph - sh / ph;
    if (gw_pass) gg[2] += sw / gw;
    if (gh_pass) gg[3] += sh / gh;
    gg[4] += sa / kPi;  gp[4] -= sa / kPi;
}

}  // namespace sphk
