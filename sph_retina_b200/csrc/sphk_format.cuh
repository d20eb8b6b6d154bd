// Box format conversions either side of the IoU path (sphdet/bbox/box_formator.py): one row in, one row out.
// The arithmetic follows the reference's torch expressions operation by operation (each product / quotient rounded on
// its own -- no FMA contraction), so the pure-arithmetic modes are bit-identical to it in fp32.
#pragma once
#include "sphk_math.cuh"

namespace sphk {

enum BoxFormat {
    FMT_XYXY2XYWH = 0,       // box_formator.py:17-23
    FMT_XYWH2XYXY = 1,       // :25-31
    FMT_OBB2HBB_XYWH = 2,    // :33-50  obb2hbb_wywh   [n, 5] -> [n, 4]
    FMT_OBB2HBB_XYXY = 3,    // :52-55                 [n, 5] -> [n, 4]
    FMT_BFOV2RBFOV = 4,      // :57-61                 [n, 4] -> [n, 5]
    FMT_GEO2SPH = 5,         // :64-68  theta + 180, 90 - phi, the other columns copied (4 or 5 columns)
    FMT_SPH2GEO = 6,         // :70-74
    FMT_SPH2PIX = 7,         // :77-84  (theta, phi, alpha, beta) -> (x, y, w, h) of the H x W image
    FMT_PIX2SPH = 8,         // :86-93
    FMT_SPH2TAN = 9,         // :99-107
    FMT_TAN2SPH = 10,        // :109-117
    FMT_SPH2PLANAR_PIX = 11, // Sph2PlanarBoxTransform('sph2pix') :166-182: [n, 4] -> xyxy, [n, 5] -> (x, y, w, h, -gamma rad)
    FMT_SPH2PLANAR_TAN = 12, // Sph2PlanarBoxTransform('sph2tan')
    FMT_PLANAR2SPH_PIX = 13, // Planar2SphBoxTransform('sph2pix' | 'pix2sph') :185-200: xyxy -> bfov [n, 4] or rbfov [n, 5] (gamma 0)
    FMT_PLANAR2SPH_TAN = 14, // Planar2SphBoxTransform('sph2tan' | 'tan2sph')
    FMT_COUNT = 15
};

// products / quotients / sums that must not be contracted into an FMA (torch evaluates them as separate kernels' ops)
#if defined(__CUDA_ARCH__)
SPHK_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
SPHK_HD float f_div(float a, float b) { return __fdiv_rn(a, b); }
SPHK_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
#else
SPHK_HD float f_mul(float a, float b) { volatile float r = a * b; return r; }
SPHK_HD float f_div(float a, float b) { volatile float r = a / b; return r; }
SPHK_HD float f_add(float a, float b) { volatile float r = a + b; return r; }
#endif
SPHK_HD float f_sub(float a, float b) { return f_add(a, -b); }

SPHK_HD int box_format_cols_out(int fmt, int d_in) {
    switch (fmt) {
        case FMT_OBB2HBB_XYWH: case FMT_OBB2HBB_XYXY: return 4;
        case FMT_BFOV2RBFOV: return 5;
        case FMT_GEO2SPH: case FMT_SPH2GEO: case FMT_SPH2PLANAR_PIX: case FMT_SPH2PLANAR_TAN: return d_in;
        default: return 4;       // (FMT_PLANAR2SPH_* with box_version 5 is 5: the caller passes d_out)
    }
}
SPHK_HD int box_format_cols_in(int fmt, int d_out) {
    switch (fmt) {
        case FMT_OBB2HBB_XYWH: case FMT_OBB2HBB_XYXY: return 5;
        case FMT_GEO2SPH: case FMT_SPH2GEO: case FMT_SPH2PLANAR_PIX: case FMT_SPH2PLANAR_TAN: return d_out;
        default: return 4;
    }
}

SPHK_HD void fmt_xyxy2xywh(const float* i, float* o) {
    o[0] = f_div(f_add(i[0], i[2]), 2.0f); o[1] = f_div(f_add(i[1], i[3]), 2.0f);
    o[2] = f_sub(i[2], i[0]); o[3] = f_sub(i[3], i[1]);
}
SPHK_HD void fmt_xywh2xyxy(const float* i, float* o) {
    const float hw = f_div(i[2], 2.0f), hh = f_div(i[3], 2.0f);
    o[0] = f_sub(i[0], hw); o[1] = f_sub(i[1], hh); o[2] = f_add(i[0], hw); o[3] = f_add(i[1], hh);
}
SPHK_HD void fmt_sph2pix(const float* i, float* o, float H, float W) {
    o[0] = f_mul(f_div(i[0], 360.0f), W); o[1] = f_mul(f_div(i[1], 180.0f), H);
    o[2] = f_mul(f_div(i[2], 360.0f), W); o[3] = f_mul(f_div(i[3], 180.0f), H);
}
SPHK_HD void fmt_pix2sph(const float* i, float* o, float H, float W) {
    o[0] = f_mul(f_div(i[0], W), 360.0f); o[1] = f_mul(f_div(i[1], H), 180.0f);
    o[2] = f_mul(f_div(i[2], W), 360.0f); o[3] = f_mul(f_div(i[3], H), 180.0f);
}
// torch.deg2rad / rad2deg multiply by the float32 constants pi/180 and 180/pi
SPHK_HD void fmt_sph2tan(const float* i, float* o, float H, float W) {
    const float R2 = (float)((double)W / SPHK_PI_D);
    o[0] = f_mul(f_div(i[0], 360.0f), W); o[1] = f_mul(f_div(i[1], 180.0f), H);
    o[2] = f_mul(R2, tanf(f_div(f_mul(i[2], (float)(SPHK_PI_D / 180.0)), 2.0f)));
    o[3] = f_mul(R2, tanf(f_div(f_mul(i[3], (float)(SPHK_PI_D / 180.0)), 2.0f)));
}
SPHK_HD void fmt_tan2sph(const float* i, float* o, float H, float W) {
    const float R2 = (float)((double)W / SPHK_PI_D);
    o[0] = f_mul(f_div(i[0], W), 360.0f); o[1] = f_mul(f_div(i[1], H), 180.0f);
    o[2] = f_mul(f_mul(2.0f, atanf(f_div(i[2], R2))), (float)(180.0 / SPHK_PI_D));
    o[3] = f_mul(f_mul(2.0f, atanf(f_div(i[3], R2))), (float)(180.0 / SPHK_PI_D));
}

// in: d_in columns, out: d_out columns (both <= 5)
SPHK_HD void box_format_row(int fmt, const float* in, int d_in, float* out, int d_out, float H, float W) {
    float t[4];
    switch (fmt) {
        case FMT_XYXY2XYWH: fmt_xyxy2xywh(in, out); break;
        case FMT_XYWH2XYXY: fmt_xywh2xyxy(in, out); break;
        case FMT_OBB2HBB_XYWH:
        case FMT_OBB2HBB_XYXY: {
            const float c = fabsf(cosf(in[4])), s = fabsf(sinf(in[4]));
            t[0] = in[0]; t[1] = in[1];
            t[2] = f_add(f_mul(c, in[2]), f_mul(s, in[3]));
            t[3] = f_add(f_mul(s, in[2]), f_mul(c, in[3]));
            if (fmt == FMT_OBB2HBB_XYXY) fmt_xywh2xyxy(t, out);
            else { out[0] = t[0]; out[1] = t[1]; out[2] = t[2]; out[3] = t[3]; }
            break;
        }
        case FMT_BFOV2RBFOV: out[0] = in[0]; out[1] = in[1]; out[2] = in[2]; out[3] = in[3]; out[4] = 0.0f; break;
        case FMT_GEO2SPH:
        case FMT_SPH2GEO:
            out[0] = (fmt == FMT_GEO2SPH) ? f_add(in[0], 180.0f) : f_sub(in[0], 180.0f);
            out[1] = f_sub(90.0f, in[1]);
            for (int k = 2; k < d_in; ++k) out[k] = in[k];
            break;
        case FMT_SPH2PIX: fmt_sph2pix(in, out, H, W); break;
        case FMT_PIX2SPH: fmt_pix2sph(in, out, H, W); break;
        case FMT_SPH2TAN: fmt_sph2tan(in, out, H, W); break;
        case FMT_TAN2SPH: fmt_tan2sph(in, out, H, W); break;
        case FMT_SPH2PLANAR_PIX:
        case FMT_SPH2PLANAR_TAN:
            if (fmt == FMT_SPH2PLANAR_PIX) fmt_sph2pix(in, t, H, W); else fmt_sph2tan(in, t, H, W);
            if (d_in == 4) fmt_xywh2xyxy(t, out);
            else { out[0] = t[0]; out[1] = t[1]; out[2] = t[2]; out[3] = t[3]; out[4] = -f_mul(in[4], (float)(SPHK_PI_D / 180.0)); }
            break;
        case FMT_PLANAR2SPH_PIX:
        case FMT_PLANAR2SPH_TAN:
            fmt_xyxy2xywh(in, t);
            if (fmt == FMT_PLANAR2SPH_PIX) fmt_pix2sph(t, out, H, W); else fmt_tan2sph(t, out, H, W);
            if (d_out == 5) out[4] = 0.0f;
            break;
        default: break;
    }
}

}  // namespace sphk
