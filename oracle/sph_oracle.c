/* TEST INFRASTRUCTURE ONLY -- scalar float64 C restatement of the Sph2Pob-IoU pair
 * pipeline (jitter_1 -> transform -> jitter_2 -> rotated IoU -> clamp), OpenMP over pairs.
 *
 * Two jobs:
 *   1. second opinion for the rotated-box IoU: the intersection is computed by exact
 *      Sutherland-Hodgman polygon clipping, i.e. the geometric quantity that
 *      mmcv.ops.box_iou_rotated (call site sphdet/iou/sph_iou_api.py:79; source not in the
 *      reference tree) evaluates, independent of both the vendored vertex-sort algorithm
 *      (sphdet/iou/diff_iou_rotated.py, restated in oracle/sph_oracle.py) and of the CUDA
 *      kernel's boundary-integral formulation;
 *   2. the "compiled CPU" baseline line of bench.py (cpu_baseline.kind = "port").
 *
 * The transforms follow the reference's 3-D vector formulation
 * (sph2pob_efficient.py:9-73, sph2pob_standard.py:8-80), NOT the kernel's closed forms.
 * Parity status: pinned through tests/test_oracle_golden.py (agrees with the golden vectors
 * frozen from the reference wherever the vendored rotated IoU is itself exact).
 *
 * Build: oracle/Makefile -> oracle/_build/libsph_oracle.so
 */
#include <math.h>
#include <string.h>

#define EPS (1e-4 * 1.2345678)
#define EPSA (1e-3 * 1.2345678)
#define PI 3.14159265358979323846

static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* sph_iou_api.py:244-260 */
static void jitter_spherical(double* a, double* b, int D) {
    int near = 0;
    for (int k = 0; k < D; ++k) near |= fabs(a[k] - b[k]) < EPS;
    if (near)
        for (int k = 0; k < D; ++k) { a[k] -= 2 * EPS; b[k] += EPS; }
    a[0] = clampd(a[0], 2 * EPS, 360 - EPS);
    b[0] = clampd(b[0], EPS, 360 - 2 * EPS);
    for (int k = 1; k < 4; ++k) {
        a[k] = clampd(a[k], 2 * EPS, 180 - EPS);
        b[k] = clampd(b[k], EPS, 180 - 2 * EPS);
    }
    if (D == 5) {
        b[4] = clampd(b[4], -360 + EPS, 360 - 2 * EPS);
        b[4] = clampd(b[4], -360 + 2 * EPS, 360 - EPS);
    }
}

/* sph_iou_api.py:222-242 ; o = (x, y, w, h, a) */
static void jitter_rotated(double* o1, double* o2) {
    static const int cols[4] = {0, 2, 3, 4};
    static const double add1[5] = {EPS, EPS, 2 * EPS, 2 * EPS, EPS};
    static const double add2[5] = {2 * EPS, 2 * EPS, EPS, EPS, 5 * EPS};
    int near = 0;
    for (int k = 0; k < 4; ++k) near |= fabs(o1[cols[k]] - o2[cols[k]]) < EPS;
    if (near)
        for (int k = 0; k < 5; ++k) { o1[k] += add1[k]; o2[k] += add2[k]; }
    if (fabs(o1[4] - o2[4]) < EPSA) { o1[4] += EPSA; o2[4] += 2 * EPSA; }
    for (int k = 2; k < 4; ++k) {
        if (o1[k] < 2 * EPSA / 10) o1[k] = 2 * EPSA / 10;
        if (o2[k] < EPSA / 10) o2[k] = EPSA / 10;
    }
    o1[4] = clampd(o1[4], -2 * PI + 2 * EPSA, 2 * PI - EPSA);
    o2[4] = clampd(o2[4], -2 * PI + EPSA, 2 * PI - 2 * EPSA);
}

static void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void unit3(const double* a, double* u) {
    double n = sqrt(dot3(a, a));
    if (n < 1e-12) n = 1e-12;
    u[0] = a[0] / n; u[1] = a[1] / n; u[2] = a[2] / n;
}
/* sph2pob_efficient.py:192-208 */
static double angle_between(const double* a, const double* b) {
    double ua[3], ub[3];
    unit3(a, ua); unit3(b, ub);
    return fabs(acos(clampd(dot3(ua, ub), -1 + 1e-7, 1 - 1e-7)));
}
/* sph2pob_efficient.py:211-226 */
static double turn_sign(const double* a, const double* b, const double* ref) {
    double c[3];
    cross3(a, b, c);
    return dot3(c, ref) < 0 ? 1.0 : -1.0;
}
static void centre_tangent(double th, double ph, double* c, double* d) {
    c[0] = sin(ph) * cos(th); c[1] = sin(ph) * sin(th); c[2] = cos(ph);
    d[0] = cos(ph) * cos(th); d[1] = cos(ph) * sin(th); d[2] = -sin(ph);
}
static double edge_len(double fov, int edge) {
    return edge == 1 ? 2 * sin(fov / 2) : edge == 2 ? 2 * tan(fov / 2) : fov;
}

/* sph2pob_efficient.py:9-73 (rbb_angle = 'equator') */
static void sph2pob_efficient(const double* b1, const double* b2, int D, int edge, double* o1, double* o2) {
    const double r = PI / 180;
    double cg[3], dg[3], cp[3], dp[3], z[3], ref[3];
    centre_tangent(b1[0] * r, b1[1] * r, cg, dg);
    centre_tangent(b2[0] * r, b2[1] * r, cp, dp);
    cross3(cg, cp, z);
    for (int k = 0; k < 3; ++k) ref[k] = (cg[k] + cp[k]) / 2;
    const double arc = angle_between(cg, cp);
    double ag = angle_between(dg, z) * turn_sign(z, dg, ref);
    double ap = angle_between(dp, z) * turn_sign(z, dp, ref);
    if (D == 5) { ag -= b1[4] * r; ap -= b2[4] * r; }
    o1[0] = 0; o1[1] = 0; o1[2] = edge_len(b1[2] * r, edge); o1[3] = edge_len(b1[3] * r, edge); o1[4] = ag;
    o2[0] = arc; o2[1] = 0; o2[2] = edge_len(b2[2] * r, edge); o2[3] = edge_len(b2[3] * r, edge); o2[4] = ap;
}

static void matvec(const double R[3][3], const double* v, double* o) {
    for (int i = 0; i < 3; ++i) o[i] = R[i][0] * v[0] + R[i][1] * v[1] + R[i][2] * v[2];
}
/* sph2pob_standard.py:238-261: rows (look, down, right) */
static void frame_from_angles(double th, double ph, double R[3][3]) {
    R[0][0] = sin(ph) * cos(th); R[0][1] = sin(ph) * sin(th); R[0][2] = cos(ph);
    R[1][0] = cos(ph) * cos(th); R[1][1] = cos(ph) * sin(th); R[1][2] = -sin(ph);
    R[2][0] = sin(th); R[2][1] = -cos(th); R[2][2] = 0;
}
/* sph2pob_standard.py:300-314 with gamma := -gamma, applied to d */
static void rotate_tangent(double th, double ph, double gamma, double* d) {
    double T[3][3], t[3], u[3];
    frame_from_angles(th, ph, T);
    matvec(T, d, t);
    const double g = -gamma, s = sin(g), c = cos(g);
    u[0] = t[0]; u[1] = c * t[1] - s * t[2]; u[2] = s * t[1] + c * t[2];
    for (int i = 0; i < 3; ++i) d[i] = T[0][i] * u[0] + T[1][i] * u[1] + T[2][i] * u[2];
}
/* sph2pob_standard.py:8-80 (rbb_angle = 'equator', output angle in rad) */
static void sph2pob_standard(const double* b1, const double* b2, int D, int edge, double* o1, double* o2) {
    const double r = PI / 180;
    double cg[3], dg[3], cp[3], dp[3], R[3][3], v[3];
    centre_tangent(b1[0] * r, b1[1] * r, cg, dg);
    centre_tangent(b2[0] * r, b2[1] * r, cp, dp);
    double l1 = 0;
    for (int k = 0; k < 3; ++k) l1 += fabs(cg[k] - cp[k]);
    if (l1 > 1e-8) {
        double s[3], t[3];
        for (int k = 0; k < 3; ++k) { s[k] = cg[k] + cp[k]; t[k] = cp[k] - cg[k]; }
        unit3(s, R[0]); unit3(t, R[1]); cross3(R[0], R[1], R[2]);
    } else {
        frame_from_angles((b1[0] + b2[0]) * r / 2, (b1[1] + b2[1]) * r / 2, R);
    }
    if (D == 5) {
        rotate_tangent(b1[0] * r, b1[1] * r, b1[4] * r, dg);
        rotate_tangent(b2[0] * r, b2[1] * r, b2[4] * r, dp);
    }
    const double ex[3] = {1, 0, 0}, ez[3] = {0, 0, 1}, nez[3] = {0, 0, -1};
    const double* boxes[2] = {b1, b2};
    double* cs[2] = {cg, cp};
    double* ds[2] = {dg, dp};
    double* os[2] = {o1, o2};
    for (int i = 0; i < 2; ++i) {
        double c[3], d[3], cxy[3];
        matvec(R, cs[i], c); matvec(R, ds[i], d);
        const double ang = angle_between(d, ez) * turn_sign(ez, d, ex);
        const double phi = angle_between(c, ez);
        cxy[0] = c[0]; cxy[1] = c[1]; cxy[2] = 0;
        const double theta = angle_between(cxy, ex) * turn_sign(ex, cxy, nez);
        (void)v;
        os[i][0] = theta; os[i][1] = phi;
        os[i][2] = edge_len(boxes[i][2] * r, edge); os[i][3] = edge_len(boxes[i][3] * r, edge);
        os[i][4] = ang;
    }
}

/* corners as diff_iou_rotated.py:297-322 (CCW by the angle) */
static void corners(const double* o, double q[4][2]) {
    static const double sx[4] = {0.5, -0.5, -0.5, 0.5}, sy[4] = {0.5, 0.5, -0.5, -0.5};
    const double s = sin(o[4]), c = cos(o[4]);
    for (int k = 0; k < 4; ++k) {
        const double lx = sx[k] * o[2], ly = sy[k] * o[3];
        q[k][0] = lx * c - ly * s + o[0];
        q[k][1] = lx * s + ly * c + o[1];
    }
}

/* exact convex clip of quad B by the 4 half-planes of quad A (both CCW), area by shoelace */
static double clip_area(double A[4][2], double B[4][2]) {
    double poly[16][2], tmp[16][2];
    int n = 4;
    memcpy(poly, B, sizeof(double) * 8);
    for (int e = 0; e < 4 && n > 0; ++e) {
        const double ax = A[e][0], ay = A[e][1];
        const double ex = A[(e + 1) & 3][0] - ax, ey = A[(e + 1) & 3][1] - ay;
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const double* P = poly[i];
            const double* Q = poly[(i + 1) % n];
            const double sp = ex * (P[1] - ay) - ey * (P[0] - ax);
            const double sq = ex * (Q[1] - ay) - ey * (Q[0] - ax);
            if (sp >= 0) { tmp[m][0] = P[0]; tmp[m][1] = P[1]; ++m; }
            if ((sp >= 0) != (sq >= 0)) {
                const double t = sp / (sp - sq);
                tmp[m][0] = P[0] + t * (Q[0] - P[0]);
                tmp[m][1] = P[1] + t * (Q[1] - P[1]);
                ++m;
            }
        }
        n = m;
        memcpy(poly, tmp, sizeof(double) * 2 * (size_t)n);
    }
    double s = 0;
    for (int i = 0; i < n; ++i) {
        const double* P = poly[i];
        const double* Q = poly[(i + 1) % n];
        s += P[0] * Q[1] - P[1] * Q[0];
    }
    return fabs(s) / 2;
}

static double rotated_iou(const double* o1, const double* o2, int mode) {
    double A[4][2], B[4][2];
    corners(o1, A); corners(o2, B);
    const double inter = clip_area(A, B);
    const double a1 = o1[2] * o1[3], a2 = o2[2] * o2[3];
    return mode == 0 ? inter / (a1 + a2 - inter) : inter / a1;
}

/* kind: 0 efficient, 1 standard.  b1,b2: [P,D] float32 degrees.  out: [P] float64.
 * obb (optional): [P,10] float64 = both OBBs after jitter_2. */
void sph_oracle_iou_aligned(int kind, const float* b1, const float* b2, long P, int D, int mode, int edge,
                            double* out, double* obb) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < P; ++i) {
        double a[5] = {0, 0, 0, 0, 0}, b[5] = {0, 0, 0, 0, 0}, o1[5], o2[5];
        for (int k = 0; k < D; ++k) { a[k] = b1[i * D + k]; b[k] = b2[i * D + k]; }
        jitter_spherical(a, b, D);
        if (kind == 0) sph2pob_efficient(a, b, D, edge, o1, o2);
        else sph2pob_standard(a, b, D, edge, o1, o2);
        jitter_rotated(o1, o2);
        if (obb) { memcpy(obb + i * 10, o1, 40); memcpy(obb + i * 10 + 5, o2, 40); }
        out[i] = clampd(rotated_iou(o1, o2, mode), 0, 1);
    }
}

/* rows [R,D] x cols [C,D] -> out [R,C] (pair = (rows[i], cols[j]), sph_iou_api.py:59-61) */
void sph_oracle_iou_pairwise(int kind, const float* rows, long R, const float* cols, long C, int D, int mode,
                             int edge, double* out) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < R; ++i)
        for (long j = 0; j < C; ++j) {
            double a[5] = {0, 0, 0, 0, 0}, b[5] = {0, 0, 0, 0, 0}, o1[5], o2[5];
            for (int k = 0; k < D; ++k) { a[k] = rows[i * D + k]; b[k] = cols[j * D + k]; }
            jitter_spherical(a, b, D);
            if (kind == 0) sph2pob_efficient(a, b, D, edge, o1, o2);
            else sph2pob_standard(a, b, D, edge, o1, o2);
            jitter_rotated(o1, o2);
            out[i * C + j] = clampd(rotated_iou(o1, o2, mode), 0, 1);
        }
}

/* plain rotated IoU of given OBBs [P,5] x2 (float64) */
void sph_oracle_rotated_iou(const double* o1, const double* o2, long P, int mode, double* out) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < P; ++i) out[i] = rotated_iou(o1 + i * 5, o2 + i * 5, mode);
}
