// geom.cuh -- device math shared by the preprocess / binning kernels.
#pragma once
#include "common.cuh"

namespace gsb {

// camera as the kernels see it.  The four small tensors stay on the device (they are torch
// tensors on the caller's side: gaussian_renderer/__init__.py:36-50); CamArgs is the by-value
// kernel argument, CamParams the per-block shared-memory copy.
struct CamArgs {
    const float *view, *proj, *campos, *bg;
    float tanfovx, tanfovy, focal_x, focal_y, scale_modifier;
    int W, H, gx, gy;
    int sh_degree, sh_coeffs, antialiasing;
};

struct CamParams {
    float view[16];
    float proj[16];
    float campos[3];
    float bg[3];
    float tanfovx, tanfovy, focal_x, focal_y, scale_modifier;
    int W, H, gx, gy;
    int sh_degree, sh_coeffs, antialiasing;
};

// cooperative load by the first 38 threads of the block; caller must __syncthreads() afterwards
__device__ __forceinline__ void load_cam(const CamArgs &a, CamParams &c) {
    const int t = threadIdx.x;
    if (t < 16) c.view[t] = __ldg(a.view + t);
    else if (t < 32) c.proj[t - 16] = __ldg(a.proj + t - 16);
    else if (t < 35) c.campos[t - 32] = __ldg(a.campos + t - 32);
    else if (t < 38) c.bg[t - 35] = __ldg(a.bg + t - 35);
    if (t == 0) {
        c.tanfovx = a.tanfovx; c.tanfovy = a.tanfovy; c.focal_x = a.focal_x; c.focal_y = a.focal_y;
        c.scale_modifier = a.scale_modifier; c.W = a.W; c.H = a.H; c.gx = a.gx; c.gy = a.gy;
        c.sh_degree = a.sh_degree; c.sh_coeffs = a.sh_coeffs; c.antialiasing = a.antialiasing;
    }
}

__device__ __forceinline__ void quat_to_R(const float4 q, float R[9]) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - r * z); R[2] = 2.f * (x * z + r * y);
    R[3] = 2.f * (x * y + r * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - r * x);
    R[6] = 2.f * (x * z - r * y); R[7] = 2.f * (y * z + r * x); R[8] = 1.f - 2.f * (x * x + y * y);
}

// Sigma = (R S)(R S)^T packed xx,xy,xz,yy,yz,zz.  Formula: /root/reference/utils/general_utils.py:78-110,
// /root/reference/scene/gaussian_model.py:33-37 (quaternion used as given, see oracle/torch_oracle.py).
__device__ __forceinline__ void cov3d_from_scale_rot(const float sx, const float sy, const float sz, const float mod,
                                                     const float4 q, float c6[6]) {
    float R[9];
    quat_to_R(q, R);
    const float s[3] = {mod * sx, mod * sy, mod * sz};
    float L[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) L[3 * i + k] = R[3 * i + k] * s[k];
    c6[0] = L[0] * L[0] + L[1] * L[1] + L[2] * L[2];
    c6[1] = L[0] * L[3] + L[1] * L[4] + L[2] * L[5];
    c6[2] = L[0] * L[6] + L[1] * L[7] + L[2] * L[8];
    c6[3] = L[3] * L[3] + L[4] * L[4] + L[5] * L[5];
    c6[4] = L[3] * L[6] + L[4] * L[7] + L[5] * L[8];
    c6[5] = L[6] * L[6] + L[7] * L[7] + L[8] * L[8];
}

struct Cov2D {
    float a, b, c;         // undilated 2D covariance
    float M0[3], M1[3];    // rows of J * R_view
    float tx, ty, tz;      // (clamped) view-space position
    bool cx, cy;           // clamp flags
};

__device__ __forceinline__ void cov2d(const CamParams &cam, const float px, const float py, const float pz,
                                      const float c6[6], Cov2D &o) {
    const float *v = cam.view;
    const float t0 = v[0] * px + v[4] * py + v[8] * pz + v[12];
    const float t1 = v[1] * px + v[5] * py + v[9] * pz + v[13];
    const float t2 = v[2] * px + v[6] * py + v[10] * pz + v[14];
    const float limx = FRUSTUM_CLAMP * cam.tanfovx, limy = FRUSTUM_CLAMP * cam.tanfovy;
    const float txtz = t0 / t2, tytz = t1 / t2;
    o.cx = (txtz < -limx) || (txtz > limx);
    o.cy = (tytz < -limy) || (tytz > limy);
    const float tx = fminf(limx, fmaxf(-limx, txtz)) * t2;
    const float ty = fminf(limy, fmaxf(-limy, tytz)) * t2;
    o.tx = tx; o.ty = ty; o.tz = t2;
    const float J00 = cam.focal_x / t2, J02 = -(cam.focal_x * tx) / (t2 * t2);
    const float J11 = cam.focal_y / t2, J12 = -(cam.focal_y * ty) / (t2 * t2);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        o.M0[c] = J00 * v[4 * c + 0] + J02 * v[4 * c + 2];
        o.M1[c] = J11 * v[4 * c + 1] + J12 * v[4 * c + 2];
    }
    const float S0[3] = {c6[0], c6[1], c6[2]}, S1[3] = {c6[1], c6[3], c6[4]}, S2[3] = {c6[2], c6[4], c6[5]};
    const float SM0[3] = {S0[0] * o.M0[0] + S0[1] * o.M0[1] + S0[2] * o.M0[2],
                          S1[0] * o.M0[0] + S1[1] * o.M0[1] + S1[2] * o.M0[2],
                          S2[0] * o.M0[0] + S2[1] * o.M0[1] + S2[2] * o.M0[2]};
    const float SM1[3] = {S0[0] * o.M1[0] + S0[1] * o.M1[1] + S0[2] * o.M1[2],
                          S1[0] * o.M1[0] + S1[1] * o.M1[1] + S1[2] * o.M1[2],
                          S2[0] * o.M1[0] + S2[1] * o.M1[1] + S2[2] * o.M1[2]};
    o.a = o.M0[0] * SM0[0] + o.M0[1] * SM0[1] + o.M0[2] * SM0[2];
    o.b = o.M0[0] * SM1[0] + o.M0[1] * SM1[1] + o.M0[2] * SM1[2];
    o.c = o.M1[0] * SM1[0] + o.M1[1] * SM1[1] + o.M1[2] * SM1[2];
}

// SH basis (degree <= 3) at a unit direction: /root/reference/utils/sh_utils.py:57-112
__device__ __forceinline__ void sh_basis(const int deg, const float x, const float y, const float z, float b[16]) {
    b[0] = SH_C0;
    if (deg < 1) return;
    b[1] = -SH_C1 * y; b[2] = SH_C1 * z; b[3] = -SH_C1 * x;
    if (deg < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = SH_C2[0] * xy; b[5] = SH_C2[1] * yz; b[6] = SH_C2[2] * (2.0f * zz - xx - yy);
    b[7] = SH_C2[3] * xz; b[8] = SH_C2[4] * (xx - yy);
    if (deg < 3) return;
    b[9] = SH_C3[0] * y * (3.0f * xx - yy);
    b[10] = SH_C3[1] * xy * z;
    b[11] = SH_C3[2] * y * (4.0f * zz - xx - yy);
    b[12] = SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    b[13] = SH_C3[4] * x * (4.0f * zz - xx - yy);
    b[14] = SH_C3[5] * z * (xx - yy);
    b[15] = SH_C3[6] * x * (xx - 3.0f * yy);
}

// ------------------------------------------------------------------------------------------
// Exact tile culling.  A gaussian can only change a pixel where
//     alpha = min(0.99, o * exp(-q/2)) >= 1/255   <=>   q(d) = A dx^2 + 2B dx dy + C dy^2 <= 2 ln(255 o),
// so inside the reference's 3-sigma tile rectangle only the tiles that meet that ellipse are
// binned.  Dropped tiles would have been skipped pixel by pixel by the blend loop anyway, so the
// image is unchanged while the instance count D (and with it sort + blend work) shrinks.
//
// Per tile row the hit columns are one contiguous span (convexity), found in O(1): the ellipse's
// x-extent over the band dy in [y0,y1] is attained at the band point closest to the ellipse's
// extreme point (x_r(dy) is concave).  All arithmetic is pinned (round-to-nearest intrinsics, MUFU approximations
// through inline PTX: common.cuh) so that the counting pass (preprocess) and the emitting pass (binning) see
// bit-identical spans.
// ------------------------------------------------------------------------------------------
struct CullGeom {
    float cx, cy;      // pixel-space mean
    float A, B, C;     // conic
    float lim;         // padded 2 ln(255 o); < 0 means "never visible"
    int rx0, ry0, rx1, ry1;  // reference tile rectangle (exclusive max)
};

__device__ __forceinline__ float cull_limit(const float opacity) {
    if (!(opacity >= ALPHA_MIN)) return -1.0f;
    const float tau2 = __fmul_rn(2.0f, __logf(__fmul_rn(255.0f, opacity)));
    // log error and blend-loop rounding are covered by the padding
    return __fmaf_rn(fmaxf(tau2, 0.0f), 1.002f, 0.03f);
}

// tile-row range [ty0, ty1) of the ellipse inside the rectangle
__device__ __forceinline__ void cull_rows(const CullGeom &g, int &ty0, int &ty1) {
    ty0 = g.ry0; ty1 = g.ry0;
    if (g.lim < 0.0f) return;
    const float detc = __fmaf_rn(g.A, g.C, -__fmul_rn(g.B, g.B));
    if (!(detc > 0.0f)) { ty1 = g.ry1; return; }
    const float ey = __fmaf_rn(sqrt_apx(div_apx(__fmul_rn(g.lim, g.A), detc)), 1.002f, 0.05f);
    // rows whose pixel-centre band [16t, 16t+15] meets [cy-ey, cy+ey]
    const float lo = div_apx(__fsub_rn(__fsub_rn(g.cy, ey), 15.0f), 16.0f);
    const float hi = div_apx(__fadd_rn(g.cy, ey), 16.0f);
    const int a = (int)fminf(fmaxf(ceilf(lo), -1.0f), 65536.0f);
    const int b = (int)fminf(fmaxf(floorf(hi), -1.0f), 65536.0f) + 1;
    ty0 = max(g.ry0, a);
    ty1 = min(g.ry1, b);
    if (ty1 < ty0) ty1 = ty0;
}

// column span [tx0, tx1) of tile row ty
__device__ __forceinline__ void cull_span(const CullGeom &g, const int ty, int &tx0, int &tx1) {
    tx0 = g.rx0; tx1 = g.rx1;
    const float detc = __fmaf_rn(g.A, g.C, -__fmul_rn(g.B, g.B));
    if (!(detc > 0.0f)) return;
    const float ey = sqrt_apx(div_apx(__fmul_rn(g.lim, g.A), detc));
    const float ex = sqrt_apx(div_apx(__fmul_rn(g.lim, g.C), detc));
    const float y0 = __fsub_rn((float)(ty * TILE), g.cy), y1 = __fadd_rn(y0, 15.0f);
    const float ylo = fmaxf(y0, -ey), yhi = fminf(y1, ey);
    const float m = __fmaf_rn(ex, 0.002f, 0.05f);
    if (ylo > yhi) {
        // band misses the ellipse by less than the row padding: keep the nearest point's span
        const float yc = (y0 > 0.0f) ? ey : -ey;
        const float xc = div_apx(-__fmul_rn(g.B, yc), g.A);
        const float l = div_apx(__fsub_rn(__fsub_rn(__fadd_rn(g.cx, xc), m), 15.0f), 16.0f);
        const float h = div_apx(__fadd_rn(__fadd_rn(g.cx, xc), m), 16.0f);
        tx0 = max(g.rx0, (int)fminf(fmaxf(ceilf(l), -1.0f), 65536.0f));
        tx1 = min(g.rx1, (int)fminf(fmaxf(floorf(h), -1.0f), 65536.0f) + 1);
        if (tx1 < tx0) tx1 = tx0;
        return;
    }
    // extreme points of the ellipse: x = +-ex at dy = -+(B/C) ex
    const float dyR = __fmul_rn(div_apx(-g.B, g.C), ex);
    const float yr = fminf(fmaxf(dyR, ylo), yhi);
    const float yl = fminf(fmaxf(-dyR, ylo), yhi);
    const float Alim = __fmul_rn(g.A, g.lim);
    const float dr = fmaxf(__fmaf_rn(-detc, __fmul_rn(yr, yr), Alim), 0.0f);
    const float dl = fmaxf(__fmaf_rn(-detc, __fmul_rn(yl, yl), Alim), 0.0f);
    const float xr = div_apx(__fadd_rn(-__fmul_rn(g.B, yr), sqrt_apx(dr)), g.A);
    const float xl = div_apx(__fsub_rn(-__fmul_rn(g.B, yl), sqrt_apx(dl)), g.A);
    const float l = div_apx(__fsub_rn(__fsub_rn(__fadd_rn(g.cx, xl), m), 15.0f), 16.0f);
    const float h = div_apx(__fadd_rn(__fadd_rn(g.cx, xr), m), 16.0f);
    tx0 = max(g.rx0, (int)fminf(fmaxf(ceilf(l), -1.0f), 65536.0f));
    tx1 = min(g.rx1, (int)fminf(fmaxf(floorf(h), -1.0f), 65536.0f) + 1);
    if (tx1 < tx0) tx1 = tx0;
}

__device__ __forceinline__ uint32_t cull_count(const CullGeom &g) {
    int ty0, ty1;
    cull_rows(g, ty0, ty1);
    uint32_t n = 0;
    for (int ty = ty0; ty < ty1; ++ty) {
        int tx0, tx1;
        cull_span(g, ty, tx0, tx1);
        n += (uint32_t)(tx1 - tx0);
    }
    return n;
}

}  // namespace gsb
