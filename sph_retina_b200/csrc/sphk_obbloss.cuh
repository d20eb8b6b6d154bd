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

// ---- forward-mode dual numbers with a compile-time sparsity mask ---------------------------------------------
// Dual<T, M> carries the value and d/d(variable k) for the variables k named in the bit mask M only (x1 y1 w1 h1 a1 x2 y2 w2
// h2 a2 = bits 0..9).  The mask is part of the type and propagates through every operation (union of the operands'
// masks), so the covariance of the predicted box carries three partials, the centre offsets two, and only the last few
// operations of a loss carry all ten: the partials that are structurally zero are neither stored nor multiplied (the first
// version carried all ten through everything: 216 registers and ~3 x the FP64 work).  Entries of d[] outside M are never
// written or read; with the loops unrolled they do not exist.
constexpr int kNV = 10;
constexpr unsigned kAllVars = (1u << kNV) - 1u;
template <typename T, unsigned M>
struct Dual {
    static constexpr unsigned mask = M;
    T v;
    T d[kNV];
};

template <typename T> SPHK_HD Dual<T, 0u> dconst(T c) {
    Dual<T, 0u> r; r.v = c;
    return r;
}
template <int IDX, typename T> SPHK_HD Dual<T, (1u << IDX)> dvar(T x) {
    Dual<T, (1u << IDX)> r; r.v = x;
    r.d[IDX] = (T)1;
    return r;
}
// the same number with more (zero) partials named: where branches of different sparsity meet
template <unsigned MW, typename T, unsigned M> SPHK_HD Dual<T, MW> dwiden(const Dual<T, M>& a) {
    static_assert((M & ~MW) == 0u, "dwiden: the wider mask must contain the narrower one");
    Dual<T, MW> r; r.v = a.v;
#pragma unroll
    for (int k = 0; k < kNV; ++k)
        if ((MW >> k) & 1u) r.d[k] = ((M >> k) & 1u) ? a.d[k] : (T)0;
    return r;
}
// r = f(a) with f'(a) = fa
template <typename T, unsigned M> SPHK_HD Dual<T, M> dchain(const Dual<T, M>& a, T fv, T fa) {
    Dual<T, M> r; r.v = fv;
#pragma unroll
    for (int k = 0; k < kNV; ++k)
        if ((M >> k) & 1u) r.d[k] = fa * a.d[k];
    return r;
}
// r = f(a, b)
template <typename T, unsigned MA, unsigned MB>
SPHK_HD Dual<T, (MA | MB)> dchain2(const Dual<T, MA>& a, const Dual<T, MB>& b, T fv, T fa, T fb) {
    Dual<T, (MA | MB)> r; r.v = fv;
#pragma unroll
    for (int k = 0; k < kNV; ++k) {
        const bool ia = (MA >> k) & 1u, ib = (MB >> k) & 1u;
        if (ia && ib) r.d[k] = fa * a.d[k] + fb * b.d[k];
        else if (ia) r.d[k] = fa * a.d[k];
        else if (ib) r.d[k] = fb * b.d[k];
    }
    return r;
}
template <typename T, unsigned MA, unsigned MB> SPHK_HD Dual<T, (MA | MB)> operator+(const Dual<T, MA>& a, const Dual<T, MB>& b) { return dchain2(a, b, a.v + b.v, (T)1, (T)1); }
template <typename T, unsigned MA, unsigned MB> SPHK_HD Dual<T, (MA | MB)> operator-(const Dual<T, MA>& a, const Dual<T, MB>& b) { return dchain2(a, b, a.v - b.v, (T)1, (T)-1); }
template <typename T, unsigned MA, unsigned MB> SPHK_HD Dual<T, (MA | MB)> operator*(const Dual<T, MA>& a, const Dual<T, MB>& b) { return dchain2(a, b, a.v * b.v, b.v, a.v); }
template <typename T, unsigned MA, unsigned MB> SPHK_HD Dual<T, (MA | MB)> operator/(const Dual<T, MA>& a, const Dual<T, MB>& b) {
    const T q = a.v / b.v;
    return dchain2(a, b, q, (T)1 / b.v, -q / b.v);
}
template <typename T, unsigned M> SPHK_HD Dual<T, M> operator-(const Dual<T, M>& a) { return dchain(a, -a.v, (T)-1); }
template <typename T, unsigned M> SPHK_HD Dual<T, M> operator+(const Dual<T, M>& a, T c) { return dchain(a, a.v + c, (T)1); }
template <typename T, unsigned M> SPHK_HD Dual<T, M> operator-(const Dual<T, M>& a, T c) { return dchain(a, a.v - c, (T)1); }
template <typename T, unsigned M> SPHK_HD Dual<T, M> operator*(const Dual<T, M>& a, T c) { return dchain(a, a.v * c, c); }
template <typename T, unsigned M> SPHK_HD Dual<T, M> operator*(T c, const Dual<T, M>& a) { return dchain(a, a.v * c, c); }
template <typename T, unsigned M> SPHK_HD Dual<T, M> operator/(const Dual<T, M>& a, T c) { return dchain(a, a.v / c, (T)1 / c); }
template <typename T, unsigned M> SPHK_HD Dual<T, M> operator-(T c, const Dual<T, M>& a) { return dchain(a, c - a.v, (T)-1); }
template <typename T, unsigned M> SPHK_HD Dual<T, M> operator/(T c, const Dual<T, M>& a) {
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

template <typename T, unsigned M> SPHK_HD Dual<T, M> dsqrt(const Dual<T, M>& a) { const T s = m_sqrt(a.v); return dchain(a, s, (T)0.5 / s); }
template <typename T, unsigned M> SPHK_HD Dual<T, M> dlog(const Dual<T, M>& a) { return dchain(a, m_log(a.v), (T)1 / a.v); }
template <typename T, unsigned M> SPHK_HD Dual<T, M> dlog1p(const Dual<T, M>& a) { return dchain(a, m_log1p(a.v), (T)1 / ((T)1 + a.v)); }
template <typename T, unsigned M> SPHK_HD Dual<T, M> dexp(const Dual<T, M>& a) { const T e = m_exp(a.v); return dchain(a, e, e); }
template <typename T, unsigned M> SPHK_HD Dual<T, M> dabs(const Dual<T, M>& a) {
    const T av = a.v;
    return dchain(a, (av < (T)0) ? -av : av, (av > (T)0) ? (T)1 : ((av < (T)0) ? (T)-1 : (T)0));
}
// torch.clamp(min=lo[, max=hi]): NaN stays NaN; the gradient passes where lo <= x <= hi
template <typename T, unsigned M> SPHK_HD Dual<T, M> dclamp_min(const Dual<T, M>& a, T lo) {
    return dchain(a, (a.v < lo) ? lo : a.v, (a.v >= lo) ? (T)1 : (T)0);
}
template <typename T, unsigned M> SPHK_HD Dual<T, M> dclamp(const Dual<T, M>& a, T lo, T hi) {
    return dchain(a, (a.v < lo) ? lo : ((a.v > hi) ? hi : a.v), (a.v >= lo && a.v <= hi) ? (T)1 : (T)0);
}
// torch.max / torch.min of two tensors: equal values share the gradient
template <typename T, unsigned MA, unsigned MB> SPHK_HD Dual<T, (MA | MB)> dmax(const Dual<T, MA>& a, const Dual<T, MB>& b) {
    const T wa = (a.v > b.v) ? (T)1 : ((a.v == b.v) ? (T)0.5 : (T)0);
    return dchain2(a, b, (a.v > b.v) ? a.v : b.v, wa, (T)1 - wa);
}
template <typename T, unsigned MA, unsigned MB> SPHK_HD Dual<T, (MA | MB)> dmin(const Dual<T, MA>& a, const Dual<T, MB>& b) {
    const T wa = (a.v < b.v) ? (T)1 : ((a.v == b.v) ? (T)0.5 : (T)0);
    return dchain2(a, b, (a.v < b.v) ? a.v : b.v, wa, (T)1 - wa);
}

// ---- mmrotate gaussian_dist_loss.py: xy_wh_r_2_xy_sigma -----------------------------------------------------------
template <typename T, unsigned M>
struct Sigma2 { Dual<T, M> s11, s12, s22; };

template <typename T, unsigned MW, unsigned MH, unsigned MR>
SPHK_HD Sigma2<T, (MW | MH | MR)> obb_sigma(const Dual<T, MW>& w, const Dual<T, MH>& h, const Dual<T, MR>& r) {
    const auto wc = dclamp(w, (T)1e-7, (T)1e7);
    const auto hc = dclamp(h, (T)1e-7, (T)1e7);
    const auto a = wc * wc * (T)0.25;
    const auto b = hc * hc * (T)0.25;          // (0.5 diag(wh))^2
    const T cv = m_cos(r.v), sv = m_sin(r.v);
    const auto c = dchain(r, cv, -sv);
    const auto s = dchain(r, sv, cv);
    Sigma2<T, (MW | MH | MR)> S;
    S.s11 = c * c * a + s * s * b;
    S.s12 = c * s * (a - b);
    S.s22 = s * s * a + c * c * b;
    return S;
}
template <typename T, unsigned M> SPHK_HD Dual<T, M> sigma_det(const Sigma2<T, M>& S) { return S.s11 * S.s22 - S.s12 * S.s12; }

template <typename T, unsigned M>
SPHK_HD Dual<T, M> gd_postprocess(Dual<T, M> distance, int fun, T tau) {
    if (fun == FUN_LOG1P) distance = dlog1p(distance);
    else if (fun == FUN_SQRT) distance = dsqrt(dclamp_min(distance, (T)1e-7));
    if (tau >= (T)1) return (T)1 - (T)1 / (distance + tau);
    return distance;
}

template <typename T, unsigned MX, unsigned MY, unsigned MP, unsigned MQ>
SPHK_HD Dual<T, (MX | MY | MP | MQ)> gwd_distance(const Dual<T, MX>& dx, const Dual<T, MY>& dy, const Sigma2<T, MP>& P,
                                                  const Sigma2<T, MQ>& Q, T alpha, bool normalize) {
    const auto xy = dx * dx + dy * dy;
    const auto whr0 = (P.s11 + P.s22) + (Q.s11 + Q.s22);
    const auto t_tr = P.s11 * Q.s11 + (T)2 * (P.s12 * Q.s12) + P.s22 * Q.s22;       // trace(Sigma_p Sigma_t)
    const auto t_det_sqrt = dsqrt(dclamp_min(sigma_det(P) * sigma_det(Q), (T)1e-7));
    const auto whr = whr0 + (T)-2 * dsqrt(dclamp_min(t_tr + (T)2 * t_det_sqrt, (T)1e-7));
    auto distance = dsqrt(dclamp_min(xy + (alpha * alpha) * whr, (T)1e-7));
    if (normalize) {
        const auto scale = (T)2 * dclamp_min(dsqrt(dclamp_min(dsqrt(dclamp_min(t_det_sqrt, (T)1e-7)), (T)1e-7)), (T)1e-7);
        distance = distance / scale;
    }
    return distance;
}

// KL(N_p || N_t) as mmrotate's kld_loss writes it (the inverse is taken of Sigma_p)
template <typename T, unsigned MX, unsigned MY, unsigned MP, unsigned MQ>
SPHK_HD Dual<T, (MX | MY | MP | MQ)> kld_distance(const Dual<T, MX>& dx, const Dual<T, MY>& dy, const Sigma2<T, MP>& P,
                                                  const Sigma2<T, MQ>& Q, T alpha, bool sqrt_) {
    const auto det_p = sigma_det(P);
    const auto det_t = sigma_det(Q);
    const auto i11 = P.s22 / det_p;
    const auto i12 = -P.s12 / det_p;
    const auto i22 = P.s11 / det_p;
    const auto xy = (T)0.5 * (dx * dx * i11 + (T)2 * (dx * dy * i12) + dy * dy * i22);
    const auto whr0 = (T)0.5 * (i11 * Q.s11 + (T)2 * (i12 * Q.s12) + i22 * Q.s22);
    const auto whr = whr0 + (T)0.5 * (dlog(det_p) - dlog(det_t)) - (T)1;
    auto distance = xy / (alpha * alpha) + whr;
    if (sqrt_) distance = dsqrt(dclamp_min(distance, (T)1e-7));
    return distance;
}

// the ten OBB parameters as dual variables, each with its own partial
template <typename T>
struct ObbVars {
    Dual<T, (1u << 0)> x1; Dual<T, (1u << 1)> y1; Dual<T, (1u << 2)> w1; Dual<T, (1u << 3)> h1; Dual<T, (1u << 4)> a1;   // pred OBB
    Dual<T, (1u << 5)> x2; Dual<T, (1u << 6)> y2; Dual<T, (1u << 7)> w2; Dual<T, (1u << 8)> h2; Dual<T, (1u << 9)> a2;   // target OBB
};

template <typename T>
SPHK_HD Dual<T, kAllVars> gd_loss_row(const ObbVars<T>& q, const LossParams& P) {
    const auto Sp = obb_sigma(q.w1, q.h1, q.a1);
    const auto St = obb_sigma(q.w2, q.h2, q.a2);
    const auto dx = q.x1 - q.x2;
    const auto dy = q.y1 - q.y2;
    const T alpha = (T)P.alpha;
    const bool opt = (P.flags & LF_NORMALIZE_OR_SQRT) != 0;
    Dual<T, kAllVars> dist;
    if (P.kind == LOSS_GWD) {
        dist = gwd_distance(dx, dy, Sp, St, alpha, opt);
    } else if (P.kind == LOSS_KLD) {
        dist = kld_distance(dx, dy, Sp, St, alpha, opt);
    } else {
        const auto mdx = -dx;
        const auto mdy = -dy;
        if (P.kind == LOSS_JD) {
            dist = (kld_distance(dx, dy, Sp, St, alpha, false) + kld_distance(mdx, mdy, St, Sp, alpha, false)) * (T)0.5;
            if (opt) dist = dsqrt(dclamp_min(dist, (T)1e-7));
        } else {
            const auto a = kld_distance(dx, dy, Sp, St, alpha, opt);
            const auto b = kld_distance(mdx, mdy, St, Sp, alpha, opt);
            dist = (P.kind == LOSS_KLD_SYMMAX) ? dmax(a, b) : dmin(a, b);
        }
    }
    return gd_postprocess(dist, P.fun, (T)P.tau);
}

// smooth-L1 of one centre coordinate as mmrotate's kfiou_loss writes it
template <typename T, unsigned MA, unsigned MB>
SPHK_HD Dual<T, (MA | MB)> kf_center_term(const Dual<T, MA>& a, const Dual<T, MB>& b, T beta) {
    const auto diff = dabs(a - b);
    return (diff.v < beta) ? ((T)0.5 * diff * diff / beta) : (diff - (T)0.5 * beta);
}

// mmrotate kf_iou_loss.py: kfiou_loss(pred, target, pred_decode, targets_decode).  The reference's subclass passes
// pred_decode = TARGET OBB and targets_decode = PRED OBB (sph2pob_kf_loss.py:26).
template <typename T>
SPHK_HD Dual<T, kAllVars> kf_loss_row(const ObbVars<T>& q, const LossParams& P) {
    const T beta = (T)P.beta, eps = (T)P.eps;
    const auto xy_loss = kf_center_term(q.x1, q.x2, beta) + kf_center_term(q.y1, q.y2, beta);
    const auto Sp = obb_sigma(q.w2, q.h2, q.a2);      // "pred_decode"    = target OBB
    const auto St = obb_sigma(q.w1, q.h1, q.a1);      // "targets_decode" = pred OBB
    const auto Vb_p = (T)4 * dsqrt(sigma_det(Sp));
    const auto Vb_t = (T)4 * dsqrt(sigma_det(St));
    // K = Sigma_p (Sigma_p + Sigma_t)^-1 ;  Sigma = Sigma_p - K Sigma_p
    const auto m11 = Sp.s11 + St.s11;
    const auto m12 = Sp.s12 + St.s12;
    const auto m22 = Sp.s22 + St.s22;
    const auto md = m11 * m22 - m12 * m12;
    const auto n11 = m22 / md;
    const auto n12 = -m12 / md;
    const auto n22 = m11 / md;
    const auto k11 = Sp.s11 * n11 + Sp.s12 * n12;
    const auto k12 = Sp.s11 * n12 + Sp.s12 * n22;
    const auto k21 = Sp.s12 * n11 + Sp.s22 * n12;
    const auto k22 = Sp.s12 * n12 + Sp.s22 * n22;
    const auto e11 = Sp.s11 - (k11 * Sp.s11 + k12 * Sp.s12);
    const auto e12 = Sp.s12 - (k11 * Sp.s12 + k12 * Sp.s22);
    const auto e21 = Sp.s12 - (k21 * Sp.s11 + k22 * Sp.s12);
    const auto e22 = Sp.s22 - (k21 * Sp.s12 + k22 * Sp.s22);
    auto Vb = (T)4 * dsqrt(e11 * e22 - e12 * e21);
    if (!(Vb.v == Vb.v)) Vb = dwiden<decltype(Vb)::mask>(dconst<T>((T)0));       // torch.where(isnan(Vb), 0, Vb)
    const auto kfiou = Vb / (Vb_p + Vb_t - Vb + eps);
    const auto k1 = (T)1 - kfiou;
    auto kf = k1;
    if (P.fun == 1) kf = -dlog(kfiou + eps);
    else if (P.fun == 2) kf = dexp(k1) - (T)1;
    return dclamp_min(xy_loss + kf, (T)0);
}

// One row of a GD / KF loss: returns the loss, adds up * d(loss)/d(OBBs) to go1 / go2 (x y w h a each).
template <typename T>
SPHK_HD float obb_scalar_loss_row(const ObbPair& o, const LossParams& P, float up, float* go1, float* go2) {
    ObbVars<T> q;
    q.x1 = dvar<0>((T)o.x1); q.y1 = dvar<1>((T)o.y1); q.w1 = dvar<2>((T)o.w1); q.h1 = dvar<3>((T)o.h1); q.a1 = dvar<4>((T)o.a1);
    q.x2 = dvar<5>((T)o.x2); q.y2 = dvar<6>((T)o.y2); q.w2 = dvar<7>((T)o.w2); q.h2 = dvar<8>((T)o.h2); q.a2 = dvar<9>((T)o.a2);
    const Dual<T, kAllVars> L = (P.kind == LOSS_KFIOU) ? kf_loss_row<T>(q, P) : gd_loss_row<T>(q, P);
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
