// TEST INFRASTRUCTURE ONLY.  A g++ build of sph_retina_b200/csrc/sphk_math.cuh so that the
// no-GPU test-suite can check the *device arithmetic* (same source, host libm) against the
// golden vectors.  The product package never loads this library: the product path is the CUDA
// C-ABI library and fails loudly without a GPU.
#include "../../sph_retina_b200/csrc/sphk_math.cuh"
#include "../../sph_retina_b200/csrc/sphk_fast.cuh"
#include "../../sph_retina_b200/csrc/sphk_coder.cuh"
#include "../../sph_retina_b200/csrc/sphk_format.cuh"
#ifdef SPHK_WITH_GRAD
#include "../../sph_retina_b200/csrc/sphk_grad.cuh"
#include "../../sph_retina_b200/csrc/sphk_obbloss.cuh"
#endif

using namespace sphk;

static inline RawBox load_box(const float* b, long i, int D) {
    RawBox r;
    r.t = b[i * D + 0]; r.p = b[i * D + 1]; r.a = b[i * D + 2]; r.b = b[i * D + 3];
    r.g = (D == 5) ? b[i * D + 4] : 0.0f;
    return r;
}

extern "C" {

void hostsim_iou_aligned(int kind, const float* b1, const float* b2, long P, int D, int mode, int edge, float* out) {
    for (long i = 0; i < P; ++i) {
        const RawBox x = load_box(b1, i, D), y = load_box(b2, i, D);
        out[i] = (kind == KIND_SPH || kind == KIND_FOV) ? approx_iou_pair(x, y, kind)
                 : (kind == KIND_NAIVE)                 ? naive_iou_pair(x, y, D, mode)
                 : (kind == KIND_UNBIASED)              ? unbiased_iou_pair(x, y, D)
                 : (kind == KIND_SPH2POB_LEGACY)        ? sph2pob_legacy_iou_pair(x, y, mode, edge)
                                                        : sph2pob_iou_pair(x, y, D, kind, mode, edge);
    }
}

// Sph-IoU / FoV-IoU through the general jitter_1 form only (hi + lo bookkeeping for every pair) and, in path[i], whether
// approx_iou_pair takes its identity shortcut on that pair -- to check that the shortcut returns the same bits.
void hostsim_approx_general(int kind, const float* b1, const float* b2, long P, float* out, unsigned char* path) {
    for (long i = 0; i < P; ++i) {
        const RawBox x = load_box(b1, i, 4), y = load_box(b2, i, 4);
        out[i] = approx_iou_pair_general(x, y, kind);
        path[i] = jitter1_is_identity4(x, y) ? 1 : 0;
    }
}

// The N x M formulation (csrc/sphk_fast.cuh): per-box precompute, prefilter, fast path with fallback.
// path[i] (optional): 0 = circle prefilter said disjoint, 4 = box-frame prefilter did, 1 = fast path, 2 = reference-order path.
void hostsim_iou_aligned_v2(int kind, const float* b1, const float* b2, long P, int D, int mode, int edge, float* out,
                            unsigned char* path) {
    for (long i = 0; i < P; ++i) {
        const RawBox x = load_box(b1, i, D), y = load_box(b2, i, D);
        BoxRec gr, pr;
        BoxCull gc, pc;
        box_pre(x, 1, D, edge, &gr, &gc);
        box_pre(y, 2, D, edge, &pr, &pc);
        if (pre_disjoint(gc, pc)) { out[i] = 0.0f; if (path) path[i] = 0; continue; }
        if (pre_outside_box(gc, pc)) { out[i] = 0.0f; if (path) path[i] = 4; continue; }   // row-frame box test of the scan loops
        float v;
        if (pair_fast(gr, pr, D, kind, mode, &v)) { out[i] = v; if (path) path[i] = 1; continue; }
        out[i] = sph2pob_iou_pair(x, y, D, kind, mode, edge);
        if (path) path[i] = 2;
    }
}

// Scalar building blocks of the fast path: arc = 2 asin(sqrt(hav)) (polynomial, no sqrtf / asinf), and the half-difference
// terms of pair_job next to what sincos_deg gives for the same angle (they must be the same bits).
void hostsim_arc_from_hav(const float* hav, long n, float* arc) {
    for (long i = 0; i < n; ++i) arc[i] = arc_from_hav(hav[i]);
}
void hostsim_half_angle_terms(const float* x_deg, long n, float* s2, float* sin2x, float* s2_ref, float* sin2x_ref) {
    for (long i = 0; i < n; ++i) {
        sin2_and_sin_double_deg(x_deg[i], &s2[i], &sin2x[i]);
        float s, c;
        sincos_deg(x_deg[i], 0.0f, &s, &c);
        s2_ref[i] = s * s;
        sin2x_ref[i] = 2.0f * s * c;
    }
}

// Prefilter verdicts next to the reference-order IoU computed WITHOUT any early-out (dense = true): cull[i] bit 0 = circle
// test, bit 1 = box-frame test (row = box 1), bit 2 = separating-axis test.  A culled pair must have dense IoU exactly 0.
void hostsim_prefilter(int kind, const float* b1, const float* b2, long P, int D, int mode, int edge, float* dense_iou,
                       unsigned char* cull) {
    for (long i = 0; i < P; ++i) {
        const RawBox x = load_box(b1, i, D), y = load_box(b2, i, D);
        BoxRec gr, pr;
        BoxCull gc, pc;
        box_pre(x, 1, D, edge, &gr, &gc);
        box_pre(y, 2, D, edge, &pr, &pc);
        cull[i] = (pre_disjoint(gc, pc) ? 1 : 0) | (pre_outside_box(gc, pc) ? 2 : 0) | (pre_sat_disjoint(gc, pc) ? 4 : 0);
        dense_iou[i] = sph2pob_iou_pair(x, y, D, kind, mode, edge, true);
    }
}

// The aligned formulation (k_iou_aligned2): stage 0 (approximate cull) for every pair; jitter_1 + transform +
// clipper for the survivors.  path: 0 = culled by stage 0, 3 = culled by the exact test of stage 1, 1 = fast,
// 5 = fast through the hi + lo stage 1 (similarity mask / upper clamp), 2 = reference-order path.
void hostsim_iou_aligned_v3(int kind, const float* b1, const float* b2, long P, int D, int mode, int edge, float* out,
                            unsigned char* path) {
    for (long i = 0; i < P; ++i) {
        const RawBox x = load_box(b1, i, D), y = load_box(b2, i, D);
        if (pair_far_apart(x, y, edge)) { out[i] = 0.0f; if (path) path[i] = 0; continue; }
        PairS1 s1;
        ClipJob job;
        int st = pair_stage1(x, y, D, edge, true, &s1);
        bool general = false;
        if (st == JOB_SLOW) { st = pair_stage1_general(x, y, D, edge, true, &s1); general = true; }
        if (st == JOB_READY) st = pair_stage2(s1, D, kind, &job);
        if (path) path[i] = (st == JOB_DEAD) ? 3 : (st == JOB_READY ? (general ? 5 : 1) : 2);
        out[i] = (st == JOB_DEAD) ? 0.0f : (st == JOB_READY ? clip_job_iou(job, mode) : sph2pob_iou_pair(x, y, D, kind, mode, edge));
    }
}

// Stage 0 against the exact tests: far[i] = pair_far_apart, dead1[i] = pair_stage1(cull) == JOB_DEAD or slow-path
// pair whose jittered OBBs are disjoint (obb_disjoint), hav[i] = the exact haversine of stage 1 (NaN when slow).
void hostsim_stage0(const float* b1, const float* b2, long P, int D, int edge, unsigned char* far, unsigned char* dead1,
                    float* hav) {
    for (long i = 0; i < P; ++i) {
        const RawBox x = load_box(b1, i, D), y = load_box(b2, i, D);
        far[i] = pair_far_apart(x, y, edge) ? 1 : 0;
        PairS1 s1;
        const int st = pair_stage1(x, y, D, edge, true, &s1);
        hav[i] = (st == JOB_SLOW) ? NAN : s1.hav;
        if (st == JOB_SLOW) {
            const bool m = jitter1_mask(x, y, D);
            const JitBox g = jitter1_role1(x, m, D), p = jitter1_role2(y, m, D);
            XformAux aux;
            ObbPair o = sph2pob_efficient(g, p, D, edge, &aux);
            jitter2(o);
            dead1[i] = obb_disjoint(o) ? 1 : 0;
        } else {
            dead1[i] = (st == JOB_DEAD) ? 1 : 0;
        }
    }
}

void hostsim_iou_aligned_project(int kind, const float* b1, const float* b2, long P, int D, int mode, int edge, float* out) {
    for (long i = 0; i < P; ++i)
        out[i] = sph2pob_iou_pair_project(load_box(b1, i, D), load_box(b2, i, D), D, kind, mode, edge);
}

// OBBs after transform + both jitters: out [P,10] = (x1,y1,w1,h1,a1,x2,y2,w2,h2,a2)
void hostsim_obbs(int kind, const float* b1, const float* b2, long P, int D, int edge, float* out) {
    for (long i = 0; i < P; ++i) {
        const RawBox x = load_box(b1, i, D), y = load_box(b2, i, D);
        const bool m = jitter1_mask(x, y, D);
        const JitBox g = jitter1_role1(x, m, D), p = jitter1_role2(y, m, D);
        XformAux aux;
        ObbPair o = (kind == KIND_SPH2POB_STANDARD) ? sph2pob_standard(g, p, D, edge, &aux)
                                                    : sph2pob_efficient(g, p, D, edge, &aux);
        jitter2(o);
        float* r = out + i * 10;
        r[0] = o.x1; r[1] = o.y1; r[2] = o.w1; r[3] = o.h1; r[4] = o.a1;
        r[5] = o.x2; r[6] = o.y2; r[7] = o.w2; r[8] = o.h2; r[9] = o.a2;
    }
}

static CoderParams make_coder(int D, const float* means, const float* stds, float wh_ratio_clip, int clip_border,
                              int add_ctr_clamp, float ctr_clamp) {
    CoderParams cp;
    for (int k = 0; k < 5; ++k) { cp.mean[k] = (means && k < D) ? means[k] : 0.0f; cp.stdv[k] = (stds && k < D) ? stds[k] : 1.0f; }
    cp.max_ratio = fabsf(logf(wh_ratio_clip));
    cp.ctr_clamp = ctr_clamp; cp.clip_border = clip_border; cp.add_ctr_clamp = add_ctr_clamp;
    return cp;
}

// box coders (csrc/sphk_coder.cuh): out = delta2bbox(rois, deltas) / bbox2delta(rois, gt)
void hostsim_coder_decode(const float* rois, const float* deltas, long n, int D, const float* means, const float* stds,
                          float wh_ratio_clip, int clip_border, int add_ctr_clamp, float ctr_clamp, float* out) {
    const CoderParams cp = make_coder(D, means, stds, wh_ratio_clip, clip_border, add_ctr_clamp, ctr_clamp);
    for (long i = 0; i < n; ++i) {
        uint32_t pass; float jac[5];
        const RawBox b = coder_decode(load_box(rois, i, D), deltas + i * D, D, cp, &pass, jac);
        const float v[5] = {b.t, b.p, b.a, b.b, b.g};
        for (int k = 0; k < D; ++k) out[i * D + k] = v[k];
    }
}
void hostsim_coder_encode(const float* rois, const float* gt, long n, int D, const float* means, const float* stds, float* out) {
    const CoderParams cp = make_coder(D, means, stds, 1.0f, 0, 0, 0.0f);
    for (long i = 0; i < n; ++i) {
        float v[5];
        coder_encode(load_box(rois, i, D), load_box(gt, i, D), D, cp, v);
        for (int k = 0; k < D; ++k) out[i * D + k] = v[k];
    }
}

#ifdef SPHK_WITH_GRAD
// fused loss forward/backward (standard transform): iou[P], g1[P,D], g2[P,D] for upstream grad_iou[P]
void hostsim_loss_fwd_bwd(const float* b1, const float* b2, const float* grad_iou, long P, int D, float* iou,
                          float* g1, float* g2) {
    for (long i = 0; i < P; ++i) {
        const RawBox x = load_box(b1, i, D), y = load_box(b2, i, D);
        float ga[5], gb[5];
        iou[i] = sph2pob_iou_pair_grad(x, y, D, KIND_SPH2POB_STANDARD, EDGE_ARC, grad_iou[i], ga, gb);
        for (int k = 0; k < D; ++k) { g1[i * D + k] = ga[k]; g2[i * D + k] = gb[k]; }
    }
}

// the head's decode -> Sph2PobIoULoss step (k_decode_loss): returns sum_i w_i (1 - iou_i), gd = d(scale * that)/d(deltas)
double hostsim_decode_loss(const float* anchors, const float* deltas, const float* target, const float* weight, long n, int D,
                           const float* means, const float* stds, float wh_ratio_clip, int clip_border, int add_ctr_clamp,
                           float ctr_clamp, float scale, float* gd) {
    const CoderParams cp = make_coder(D, means, stds, wh_ratio_clip, clip_border, add_ctr_clamp, ctr_clamp);
    double total = 0.0;
    for (long i = 0; i < n; ++i) {
        const float w = weight ? weight[i] : 1.0f;
        for (int k = 0; k < D; ++k) gd[i * D + k] = 0.0f;
        if (w == 0.0f) continue;
        uint32_t pass; float jac[5], g1[5], g2[5], g[5];
        const RawBox pred = coder_decode(load_box(anchors, i, D), deltas + i * D, D, cp, &pass, jac);
        const float iou = sph2pob_iou_pair_grad(pred, load_box(target, i, D), D, KIND_SPH2POB_STANDARD, EDGE_ARC, -w * scale, g1, g2);
        coder_decode_grad(g1, pass, jac, D, g);
        for (int k = 0; k < D; ++k) gd[i * D + k] = g[k];
        total += (double)w * (1.0 - (double)iou);
    }
    return total;
}

// GD / KF / L1 losses on the Sph2Pob OBBs (k_obb_loss): loss[n, L], g1/g2[n, D] = sum_j up[i, j] * d(loss_ij)/d(box);
// use_double selects the arithmetic type of the dual numbers (the kernel is built with float)
void hostsim_obb_loss(int kind, int fun, int flags, float tau, float alpha, float beta, float eps, int xkind, const float* b1,
                      const float* b2, const float* up, int up_cols, long n, int D, int use_double, float* loss, float* g1,
                      float* g2) {
    LossParams lp;
    lp.kind = kind; lp.fun = fun; lp.flags = flags; lp.tau = tau; lp.alpha = alpha; lp.beta = beta; lp.eps = eps;
    const int L = loss_columns(kind);
    for (long i = 0; i < n; ++i) {
        const RawBox x = load_box(b1, i, D), y = load_box(b2, i, D);
        const bool m = jitter1_mask(x, y, D);
        const JitBox g = jitter1_role1(x, m, D), p = jitter1_role2(y, m, D);
        XformAux aux;
        ObbPair o = (xkind == KIND_SPH2POB_STANDARD) ? sph2pob_standard(g, p, D, EDGE_ARC, &aux)
                                                     : sph2pob_efficient(g, p, D, EDGE_ARC, &aux);
        const uint32_t pass2 = jitter2(o);
        float u[5], go1[5] = {0, 0, 0, 0, 0}, go2[5] = {0, 0, 0, 0, 0}, gb1[5], gb2[5];
        for (int k = 0; k < 5; ++k) u[k] = up ? up[i * up_cols + (up_cols > 1 ? k : 0)] : 1.0f;
        if (kind == LOSS_L1) {
            obb_l1_loss_row(o, lp, u, loss + i * 5, go1, go2);
        } else {
            loss[i] = use_double ? obb_scalar_loss_row<double>(o, lp, u[0], go1, go2)
                                 : obb_scalar_loss_row<float>(o, lp, u[0], go1, go2);
        }
        (void)L;
        jitter2_grad(pass2, go1, go2);
        xform_grad(xkind, g, p, D, EDGE_ARC, aux, go1, go2, gb1, gb2);
        for (int k = 0; k < D; ++k) { g1[i * D + k] = gb1[k]; g2[i * D + k] = gb2[k]; }
    }
}

#endif
void hostsim_box_format(int fmt, const float* in, long n, int d_in, int d_out, float img_h, float img_w, float* out) {
    for (long i = 0; i < n; ++i) {
        float a[5], b[5];
        for (int k = 0; k < d_in; ++k) a[k] = in[i * d_in + k];
        box_format_row(fmt, a, d_in, b, d_out, img_h, img_w);
        for (int k = 0; k < d_out; ++k) out[i * d_out + k] = b[k];
    }
}

}  // extern "C"
