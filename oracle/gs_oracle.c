/*
 * gs_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, not product code) for the differentiable
 * 3D-Gaussian-splatting rasterizer: forward + hand-derived backward, float32, OpenMP over
 * gaussians / tiles.  Built by oracle/Makefile into oracle/libgs_oracle.so and driven from
 * oracle/c_oracle.py through ctypes.
 *
 * PARITY STATUS: "parity unpinned" for the splatting rules.  The reference implementation of
 * this path (submodules/diff-gaussian-rasterization @ 59f5f77e, branch dr_aa:
 * cuda_rasterizer/{forward,backward,rasterizer_impl}.cu, rasterize_points.cu) is an EMPTY
 * directory in /root/reference, so no file:line can be cited for it.  What is cited:
 *   - SH basis / constants ........ /root/reference/utils/sh_utils.py:26-112
 *   - colour rule (+0.5, clamp 0) . /root/reference/gaussian_renderer/__init__.py:76-80
 *   - quaternion -> R, Sigma=LL^T . /root/reference/utils/general_utils.py:78-110,
 *                                   /root/reference/scene/gaussian_model.py:33-37
 *   - matrix conventions .......... /root/reference/scene/cameras.py:86-89 (transposed 4x4)
 *   - outputs (image, radii, inverse depth) /root/reference/gaussian_renderer/__init__.py:91-126
 * The splatting constants (0.2 near cull, 1.3 clamp, 0.3 dilation, 3-sigma radius, 1/255,
 * 0.99, 1e-4, 16x16 tiles) are restated from the published 3DGS algorithm and are the same
 * UNVERIFIED_VS_REFERENCE set as oracle/torch_oracle.py, which validates this file's
 * backward through torch.autograd (tests/test_oracle.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16
#define NEAR_CULL 0.2f
#define FRUSTUM_CLAMP 1.3f
#define DILATION 0.3f
#define AA_FLOOR 0.000025f
#define ALPHA_MIN (1.0f / 255.0f)
#define ALPHA_MAX 0.99f
#define T_STOP 0.0001f

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

typedef struct {
    int P, M, deg, W, H, gx, gy, aa, ty0, ty1;
    int has_sh, has_cov_precomp;
    float tanfovx, tanfovy, scale_modifier;
    float view[16], proj[16], campos[3], bg[3];
    /* inputs kept by pointer (caller keeps them alive until gso_free) */
    const float *means, *shs, *colors_pre, *opac, *scales, *rots, *cov_pre;
    /* per gaussian */
    float *depth, *xy, *conic_o, *rgb, *cov3d;
    int *radii, *rect;
    uint8_t *clamped;
    /* binning */
    int64_t D;
    int64_t *tile_start; /* gx*gy+1 */
    int *list;           /* D sorted gaussian ids */
    /* per pixel */
    float *final_T;
    int *n_contrib;
} GsoState;

static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* basis values for degree <= 3 at unit direction (x,y,z) -- utils/sh_utils.py:57-112 */
static void sh_basis(int deg, float x, float y, float z, float *b) {
    b[0] = SH_C0;
    if (deg < 1) return;
    b[1] = -SH_C1 * y; b[2] = SH_C1 * z; b[3] = -SH_C1 * x;
    if (deg < 2) return;
    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
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

/* d basis / d(x,y,z) */
static void sh_basis_grad(int deg, float x, float y, float z, float *bx, float *by, float *bz) {
    for (int i = 0; i < 16; ++i) bx[i] = by[i] = bz[i] = 0.0f;
    if (deg < 1) return;
    by[1] = -SH_C1; bz[2] = SH_C1; bx[3] = -SH_C1;
    if (deg < 2) return;
    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    bx[4] = SH_C2[0] * y; by[4] = SH_C2[0] * x;
    by[5] = SH_C2[1] * z; bz[5] = SH_C2[1] * y;
    bx[6] = SH_C2[2] * -2.0f * x; by[6] = SH_C2[2] * -2.0f * y; bz[6] = SH_C2[2] * 4.0f * z;
    bx[7] = SH_C2[3] * z; bz[7] = SH_C2[3] * x;
    bx[8] = SH_C2[4] * 2.0f * x; by[8] = SH_C2[4] * -2.0f * y;
    if (deg < 3) return;
    bx[9] = SH_C3[0] * 6.0f * xy;            by[9] = SH_C3[0] * (3.0f * xx - 3.0f * yy);
    bx[10] = SH_C3[1] * yz;                  by[10] = SH_C3[1] * xz;  bz[10] = SH_C3[1] * xy;
    bx[11] = SH_C3[2] * -2.0f * xy;          by[11] = SH_C3[2] * (4.0f * zz - xx - 3.0f * yy);
    bz[11] = SH_C3[2] * 8.0f * yz;
    bx[12] = SH_C3[3] * -6.0f * xz;          by[12] = SH_C3[3] * -6.0f * yz;
    bz[12] = SH_C3[3] * (6.0f * zz - 3.0f * xx - 3.0f * yy);
    bx[13] = SH_C3[4] * (4.0f * zz - 3.0f * xx - yy); by[13] = SH_C3[4] * -2.0f * xy;
    bz[13] = SH_C3[4] * 8.0f * xz;
    bx[14] = SH_C3[5] * 2.0f * xz;           by[14] = SH_C3[5] * -2.0f * yz;
    bz[14] = SH_C3[5] * (xx - yy);
    bx[15] = SH_C3[6] * (3.0f * xx - 3.0f * yy); by[15] = SH_C3[6] * -6.0f * xy;
}

static void quat_to_R(const float *q, float R[9]) {
    float r = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - r * z); R[2] = 2.f * (x * z + r * y);
    R[3] = 2.f * (x * y + r * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - r * x);
    R[6] = 2.f * (x * z - r * y); R[7] = 2.f * (y * z + r * x); R[8] = 1.f - 2.f * (x * x + y * y);
}

/* Sigma = (R S)(R S)^T, packed xx,xy,xz,yy,yz,zz */
static void cov3d_from_scale_rot(const float *s, float mod, const float *q, float *c6) {
    float R[9]; quat_to_R(q, R);
    float L[9];
    for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) L[3 * i + k] = R[3 * i + k] * (mod * s[k]);
    float S[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        float a = 0.f; for (int k = 0; k < 3; ++k) a += L[3 * i + k] * L[3 * j + k]; S[3 * i + j] = a;
    }
    c6[0] = S[0]; c6[1] = S[1]; c6[2] = S[2]; c6[3] = S[4]; c6[4] = S[5]; c6[5] = S[8];
}

typedef struct { float a, b, c; float M0[3], M1[3]; float tx, ty, tz; int cx, cy; float fx, fy; } Cov2D;

static void cov2d(const float *view, const float *p, const float *c6, float fx, float fy,
                  float tanx, float tany, Cov2D *o) {
    float t0 = view[0] * p[0] + view[4] * p[1] + view[8] * p[2] + view[12];
    float t1 = view[1] * p[0] + view[5] * p[1] + view[9] * p[2] + view[13];
    float t2 = view[2] * p[0] + view[6] * p[1] + view[10] * p[2] + view[14];
    float limx = FRUSTUM_CLAMP * tanx, limy = FRUSTUM_CLAMP * tany;
    float txtz = t0 / t2, tytz = t1 / t2;
    o->cx = (txtz < -limx) || (txtz > limx);
    o->cy = (tytz < -limy) || (tytz > limy);
    float tx = clampf(txtz, -limx, limx) * t2, ty = clampf(tytz, -limy, limy) * t2;
    o->tx = tx; o->ty = ty; o->tz = t2; o->fx = fx; o->fy = fy;
    float J00 = fx / t2, J02 = -(fx * tx) / (t2 * t2), J11 = fy / t2, J12 = -(fy * ty) / (t2 * t2);
    /* Rv[r][c] = view[4c + r] */
    for (int c = 0; c < 3; ++c) {
        o->M0[c] = J00 * view[4 * c + 0] + J02 * view[4 * c + 2];
        o->M1[c] = J11 * view[4 * c + 1] + J12 * view[4 * c + 2];
    }
    float S[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
    float SM0[3], SM1[3];
    for (int i = 0; i < 3; ++i) {
        SM0[i] = S[3 * i] * o->M0[0] + S[3 * i + 1] * o->M0[1] + S[3 * i + 2] * o->M0[2];
        SM1[i] = S[3 * i] * o->M1[0] + S[3 * i + 1] * o->M1[1] + S[3 * i + 2] * o->M1[2];
    }
    o->a = o->M0[0] * SM0[0] + o->M0[1] * SM0[1] + o->M0[2] * SM0[2];
    o->b = o->M0[0] * SM1[0] + o->M0[1] * SM1[1] + o->M0[2] * SM1[2];
    o->c = o->M1[0] * SM1[0] + o->M1[1] * SM1[1] + o->M1[2] * SM1[2];
}

typedef struct { uint32_t key; int id; } DepthKey;
static int cmp_depth(const void *pa, const void *pb) {
    const DepthKey *a = (const DepthKey *)pa, *b = (const DepthKey *)pb;
    if (a->key != b->key) return a->key < b->key ? -1 : 1;
    return a->id < b->id ? -1 : (a->id > b->id ? 1 : 0);
}

int64_t gso_num_rendered(const GsoState *s) { return s->D; }
const float *gso_final_T(const GsoState *s) { return s->final_T; }
const int *gso_n_contrib(const GsoState *s) { return s->n_contrib; }
const int *gso_rect(const GsoState *s) { return s->rect; }
int gso_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the timed CPU legs of bench.py ask for the cores explicitly */
void gso_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void gso_free(GsoState *s) {
    if (!s) return;
    free(s->depth); free(s->xy); free(s->conic_o); free(s->rgb); free(s->cov3d); free(s->radii);
    free(s->rect); free(s->clamped); free(s->tile_start); free(s->list); free(s->final_T);
    free(s->n_contrib); free(s);
}

/* Forward.  tile rows [ty0, ty1) are rendered (ty1 <= 0 means all); pixels outside keep 0. */
GsoState *gso_forward(int P, int M, int deg, const float *means, const float *shs,
                      const float *colors_pre, const float *opac, const float *scales,
                      const float *rots, const float *cov_pre, float scale_modifier,
                      const float *view, const float *proj, const float *campos, const float *bg,
                      int W, int H, float tanfovx, float tanfovy, int antialiasing, int ty0, int ty1,
                      float *out_color, int *out_radii, float *out_invdepth) {
    GsoState *s = (GsoState *)calloc(1, sizeof(GsoState));
    s->P = P; s->M = M; s->deg = deg; s->W = W; s->H = H; s->aa = antialiasing;
    s->gx = (W + TILE - 1) / TILE; s->gy = (H + TILE - 1) / TILE;
    if (ty1 <= 0 || ty1 > s->gy) ty1 = s->gy;
    if (ty0 < 0) ty0 = 0;
    s->ty0 = ty0; s->ty1 = ty1;
    s->tanfovx = tanfovx; s->tanfovy = tanfovy; s->scale_modifier = scale_modifier;
    memcpy(s->view, view, 64); memcpy(s->proj, proj, 64); memcpy(s->campos, campos, 12); memcpy(s->bg, bg, 12);
    s->means = means; s->shs = shs; s->colors_pre = colors_pre; s->opac = opac; s->scales = scales;
    s->rots = rots; s->cov_pre = cov_pre; s->has_sh = colors_pre == NULL; s->has_cov_precomp = cov_pre != NULL;
    size_t Pn = P > 0 ? (size_t)P : 1;
    s->depth = (float *)calloc(Pn, 4); s->xy = (float *)calloc(Pn * 2, 4);
    s->conic_o = (float *)calloc(Pn * 4, 4); s->rgb = (float *)calloc(Pn * 3, 4);
    s->cov3d = (float *)calloc(Pn * 6, 4); s->radii = (int *)calloc(Pn, 4);
    s->rect = (int *)calloc(Pn * 4, 4); s->clamped = (uint8_t *)calloc(Pn, 1);
    const float fx = W / (2.0f * tanfovx), fy = H / (2.0f * tanfovy);
    const int gx = s->gx, gy = s->gy;

#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        const float *p = means + 3 * i;
        float tz = view[2] * p[0] + view[6] * p[1] + view[10] * p[2] + view[14];
        if (tz <= NEAR_CULL) continue;
        float hx = proj[0] * p[0] + proj[4] * p[1] + proj[8] * p[2] + proj[12];
        float hy = proj[1] * p[0] + proj[5] * p[1] + proj[9] * p[2] + proj[13];
        float hw = proj[3] * p[0] + proj[7] * p[1] + proj[11] * p[2] + proj[15];
        float pw = 1.0f / (hw + 0.0000001f);
        float ndcx = hx * pw, ndcy = hy * pw;
        float *c6 = s->cov3d + 6 * i;
        if (cov_pre) memcpy(c6, cov_pre + 6 * i, 24);
        else cov3d_from_scale_rot(scales + 3 * i, scale_modifier, rots + 4 * i, c6);
        Cov2D cv; cov2d(view, p, c6, fx, fy, tanfovx, tanfovy, &cv);
        float det0 = cv.a * cv.c - cv.b * cv.b;
        float a = cv.a + DILATION, c = cv.c + DILATION, b = cv.b;
        float det = a * c - b * b;
        float hscale = 1.0f;
        if (antialiasing) hscale = sqrtf(fmaxf(AA_FLOOR, det0 / det));
        if (det == 0.0f) continue;
        float det_inv = 1.0f / det;
        float mid = 0.5f * (a + c);
        float root = sqrtf(fmaxf(0.1f, mid * mid - det));
        float lam = fmaxf(mid + root, mid - root);
        float radius = ceilf(3.0f * sqrtf(lam));
        float px = ((ndcx + 1.0f) * W - 1.0f) * 0.5f, py = ((ndcy + 1.0f) * H - 1.0f) * 0.5f;
        int x0 = (int)((px - radius) / TILE), y0 = (int)((py - radius) / TILE);
        int x1 = (int)((px + radius + TILE - 1) / TILE), y1 = (int)((py + radius + TILE - 1) / TILE);
        x0 = x0 < 0 ? 0 : (x0 > gx ? gx : x0); x1 = x1 < 0 ? 0 : (x1 > gx ? gx : x1);
        y0 = y0 < 0 ? 0 : (y0 > gy ? gy : y0); y1 = y1 < 0 ? 0 : (y1 > gy ? gy : y1);
        if ((x1 - x0) * (y1 - y0) == 0) continue;
        if (colors_pre) {
            for (int ch = 0; ch < 3; ++ch) s->rgb[3 * i + ch] = colors_pre[3 * i + ch];
        } else {
            float dx = p[0] - campos[0], dy = p[1] - campos[1], dz = p[2] - campos[2];
            float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
            float bas[16]; sh_basis(deg, dx * inv, dy * inv, dz * inv, bas);
            int nb = (deg + 1) * (deg + 1);
            uint8_t cl = 0;
            for (int ch = 0; ch < 3; ++ch) {
                float v = 0.f;
                for (int k = 0; k < nb; ++k) v += bas[k] * shs[((size_t)i * M + k) * 3 + ch];
                v += 0.5f;
                if (v < 0.f) { cl |= (uint8_t)(1 << ch); v = 0.f; }
                s->rgb[3 * i + ch] = v;
            }
            s->clamped[i] = cl;
        }
        s->depth[i] = tz; s->radii[i] = (int)radius;
        s->xy[2 * i] = px; s->xy[2 * i + 1] = py;
        s->conic_o[4 * i] = c * det_inv; s->conic_o[4 * i + 1] = -b * det_inv; s->conic_o[4 * i + 2] = a * det_inv;
        s->conic_o[4 * i + 3] = opac[i] * hscale;
        s->rect[4 * i] = x0; s->rect[4 * i + 1] = y0; s->rect[4 * i + 2] = x1; s->rect[4 * i + 3] = y1;
    }
    if (out_radii) memcpy(out_radii, s->radii, (size_t)P * 4);

    /* binning: (depth bits, id) sort, then per-tile lists in that order */
    int nvis = 0;
    DepthKey *keys = (DepthKey *)malloc(Pn * sizeof(DepthKey));
    for (int i = 0; i < P; ++i) if (s->radii[i] > 0) {
        uint32_t u; memcpy(&u, &s->depth[i], 4); keys[nvis].key = u; keys[nvis].id = i; ++nvis;
    }
    qsort(keys, (size_t)nvis, sizeof(DepthKey), cmp_depth);
    int nt = gx * gy;
    s->tile_start = (int64_t *)calloc((size_t)nt + 1, 8);
    for (int k = 0; k < nvis; ++k) {
        const int *r = s->rect + 4 * keys[k].id;
        for (int y = r[1]; y < r[3]; ++y) for (int x = r[0]; x < r[2]; ++x) s->tile_start[y * gx + x + 1]++;
    }
    for (int t = 0; t < nt; ++t) s->tile_start[t + 1] += s->tile_start[t];
    s->D = s->tile_start[nt];
    s->list = (int *)malloc((size_t)(s->D > 0 ? s->D : 1) * 4);
    int64_t *cursor = (int64_t *)malloc((size_t)nt * 8);
    memcpy(cursor, s->tile_start, (size_t)nt * 8);
    for (int k = 0; k < nvis; ++k) {
        const int *r = s->rect + 4 * keys[k].id;
        for (int y = r[1]; y < r[3]; ++y) for (int x = r[0]; x < r[2]; ++x) s->list[cursor[y * gx + x]++] = keys[k].id;
    }
    free(cursor); free(keys);

    /* render */
    s->final_T = (float *)calloc((size_t)W * H, 4); s->n_contrib = (int *)calloc((size_t)W * H, 4);
    memset(out_color, 0, (size_t)3 * W * H * 4);
    if (out_invdepth) memset(out_invdepth, 0, (size_t)W * H * 4);
    const int nrows = ty1 - ty0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int tt = 0; tt < nrows * gx; ++tt) {
        int tyi = ty0 + tt / gx, txi = tt % gx, t = tyi * gx + txi;
        int64_t b0 = s->tile_start[t], b1 = s->tile_start[t + 1];
        for (int ly = 0; ly < TILE; ++ly) for (int lx = 0; lx < TILE; ++lx) {
            int x = txi * TILE + lx, y = tyi * TILE + ly;
            if (x >= W || y >= H) continue;
            float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f;
            int contributor = 0, last = 0;
            for (int64_t k = b0; k < b1; ++k) {
                int g = s->list[k];
                ++contributor;
                float dx = s->xy[2 * g] - (float)x, dy = s->xy[2 * g + 1] - (float)y;
                const float *co = s->conic_o + 4 * g;
                float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                if (power > 0.0f) continue;
                float alpha = fminf(ALPHA_MAX, co[3] * expf(power));
                if (alpha < ALPHA_MIN) continue;
                float test_T = T * (1.0f - alpha);
                if (test_T < T_STOP) break;
                float w = alpha * T;
                C0 += s->rgb[3 * g] * w; C1 += s->rgb[3 * g + 1] * w; C2 += s->rgb[3 * g + 2] * w;
                Dp += (1.0f / s->depth[g]) * w;
                T = test_T; last = contributor;
            }
            size_t pid = (size_t)y * W + x;
            s->final_T[pid] = T; s->n_contrib[pid] = last;
            out_color[pid] = C0 + T * bg[0];
            out_color[(size_t)W * H + pid] = C1 + T * bg[1];
            out_color[(size_t)2 * W * H + pid] = C2 + T * bg[2];
            if (out_invdepth) out_invdepth[pid] = Dp;
        }
    }
    return s;
}

/* Backward.  Any output pointer may be NULL.  Outputs are overwritten. */
void gso_backward(GsoState *s, const float *dL_dcolor, const float *dL_dinvdepth,
                  float *g_means3D, float *g_means2D, float *g_shs, float *g_colors,
                  float *g_opac, float *g_scales, float *g_rots, float *g_cov3D) {
    const int P = s->P, W = s->W, H = s->H, gx = s->gx, M = s->M;
    size_t Pn = P > 0 ? (size_t)P : 1;
    /* per-gaussian accumulators (double so that summation order does not show at float precision):
       0,1 mean2D (ndc-scaled)  2,3,4 conic A,B,C  5 opacity(w)  6,7,8 rgb  9 invdepth */
    double *acc = (double *)calloc(Pn * 10, 8);
    const int nrows = s->ty1 - s->ty0;
#pragma omp parallel
    {
        double *loc = NULL; size_t loc_cap = 0;
#pragma omp for schedule(dynamic, 1)
        for (int tt = 0; tt < nrows * gx; ++tt) {
            int tyi = s->ty0 + tt / gx, txi = tt % gx, t = tyi * gx + txi;
            int64_t b0 = s->tile_start[t], b1 = s->tile_start[t + 1];
            size_t n = (size_t)(b1 - b0);
            if (n == 0) continue;
            if (n > loc_cap) { free(loc); loc = (double *)malloc(n * 10 * 8); loc_cap = n; }
            memset(loc, 0, n * 10 * 8);
            for (int ly = 0; ly < TILE; ++ly) for (int lx = 0; lx < TILE; ++lx) {
                int x = txi * TILE + lx, y = tyi * TILE + ly;
                if (x >= W || y >= H) continue;
                size_t pid = (size_t)y * W + x;
                const float T_final = s->final_T[pid];
                float T = T_final;
                int last = s->n_contrib[pid];
                float dLp[3] = {dL_dcolor[pid], dL_dcolor[(size_t)W * H + pid], dL_dcolor[(size_t)2 * W * H + pid]};
                float dLd = dL_dinvdepth ? dL_dinvdepth[pid] : 0.0f;
                float accum[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0.f;
                float accum_d = 0.f, last_d = 0.f;
                float bg_dot = s->bg[0] * dLp[0] + s->bg[1] * dLp[1] + s->bg[2] * dLp[2];
                for (int64_t k = b0 + last - 1; k >= b0; --k) {
                    int g = s->list[k];
                    float dx = s->xy[2 * g] - (float)x, dy = s->xy[2 * g + 1] - (float)y;
                    const float *co = s->conic_o + 4 * g;
                    float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0.0f) continue;
                    float G = expf(power);
                    float alpha = fminf(ALPHA_MAX, co[3] * G);
                    if (alpha < ALPHA_MIN) continue;
                    T = T / (1.0f - alpha);
                    float w = alpha * T;
                    double *L = loc + (size_t)(k - b0) * 10;
                    float dL_dalpha = 0.f;
                    for (int ch = 0; ch < 3; ++ch) {
                        float c = s->rgb[3 * g + ch];
                        accum[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum[ch];
                        last_color[ch] = c;
                        dL_dalpha += (c - accum[ch]) * dLp[ch];
                        L[6 + ch] += w * dLp[ch];
                    }
                    float invd = 1.0f / s->depth[g];
                    accum_d = last_alpha * last_d + (1.f - last_alpha) * accum_d;
                    last_d = invd;
                    dL_dalpha += (invd - accum_d) * dLd;
                    L[9] += w * dLd;
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                    float dL_dpow = co[3] * G * dL_dalpha;
                    L[0] += dL_dpow * (-co[0] * dx - co[1] * dy) * (0.5f * W);
                    L[1] += dL_dpow * (-co[2] * dy - co[1] * dx) * (0.5f * H);
                    L[2] += dL_dpow * (-0.5f * dx * dx);
                    L[3] += dL_dpow * (-dx * dy);
                    L[4] += dL_dpow * (-0.5f * dy * dy);
                    L[5] += G * dL_dalpha;
                }
            }
            for (size_t k = 0; k < n; ++k) {
                int g = s->list[b0 + k];
                for (int c = 0; c < 10; ++c) {
                    double v = loc[k * 10 + c];
                    if (v != 0.0) {
#pragma omp atomic
                        acc[(size_t)g * 10 + c] += v;
                    }
                }
            }
        }
        free(loc);
    }

    const float fx = W / (2.0f * s->tanfovx), fy = H / (2.0f * s->tanfovy);
    const float *view = s->view, *proj = s->proj;
    const int nb = (s->deg + 1) * (s->deg + 1);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        float gm[3] = {0, 0, 0};
        if (g_means2D) { g_means2D[3 * i] = 0; g_means2D[3 * i + 1] = 0; g_means2D[3 * i + 2] = 0; }
        if (g_opac) g_opac[i] = 0;
        if (g_colors) for (int c = 0; c < 3; ++c) g_colors[3 * i + c] = 0;
        if (g_shs) for (int k = 0; k < 3 * M; ++k) g_shs[(size_t)i * 3 * M + k] = 0;
        if (g_scales) for (int c = 0; c < 3; ++c) g_scales[3 * i + c] = 0;
        if (g_rots) for (int c = 0; c < 4; ++c) g_rots[4 * i + c] = 0;
        if (g_cov3D) for (int c = 0; c < 6; ++c) g_cov3D[6 * i + c] = 0;
        if (g_means3D) for (int c = 0; c < 3; ++c) g_means3D[3 * i + c] = 0;
        if (s->radii[i] <= 0) continue;
        const double *A = acc + (size_t)i * 10;
        const float d_m2x = (float)A[0], d_m2y = (float)A[1];
        const float dA = (float)A[2], dB = (float)A[3], dC = (float)A[4];
        const float d_w = (float)A[5];
        const float d_rgb[3] = {(float)A[6], (float)A[7], (float)A[8]};
        const float d_invd = (float)A[9];
        const float *p = s->means + 3 * i;
        if (g_means2D) { g_means2D[3 * i] = d_m2x; g_means2D[3 * i + 1] = d_m2y; }

        /* colour */
        if (s->has_sh) {
            float dx = p[0] - s->campos[0], dy = p[1] - s->campos[1], dz = p[2] - s->campos[2];
            float len = sqrtf(dx * dx + dy * dy + dz * dz), inv = 1.0f / len;
            float ux = dx * inv, uy = dy * inv, uz = dz * inv;
            float bas[16], bx[16], by[16], bz[16];
            sh_basis(s->deg, ux, uy, uz, bas); sh_basis_grad(s->deg, ux, uy, uz, bx, by, bz);
            float ddir[3] = {0, 0, 0};
            for (int ch = 0; ch < 3; ++ch) {
                float d = (s->clamped[i] >> ch) & 1 ? 0.0f : d_rgb[ch];
                for (int k = 0; k < nb; ++k) {
                    float c = s->shs[((size_t)i * M + k) * 3 + ch];
                    if (g_shs) g_shs[((size_t)i * M + k) * 3 + ch] = bas[k] * d;
                    ddir[0] += bx[k] * c * d; ddir[1] += by[k] * c * d; ddir[2] += bz[k] * c * d;
                }
            }
            float dot = ux * ddir[0] + uy * ddir[1] + uz * ddir[2];
            gm[0] += (ddir[0] - ux * dot) * inv; gm[1] += (ddir[1] - uy * dot) * inv; gm[2] += (ddir[2] - uz * dot) * inv;
        } else if (g_colors) {
            for (int ch = 0; ch < 3; ++ch) g_colors[3 * i + ch] = d_rgb[ch];
        }

        /* conic -> dilated cov2D (a,b,c) */
        const float *c6 = s->cov3d + 6 * i;
        Cov2D cv; cov2d(view, p, c6, fx, fy, s->tanfovx, s->tanfovy, &cv);
        float a0 = cv.a, b = cv.b, c0 = cv.c;
        float a = a0 + DILATION, c = c0 + DILATION;
        float det = a * c - b * b;
        float dinv2 = 1.0f / (det * det + 0.0000001f);
        float dL_da = dinv2 * (-c * c * dA + b * c * dB - b * b * dC);
        float dL_dc = dinv2 * (-b * b * dA + a * b * dB - a * a * dC);
        float dL_db = dinv2 * (2.f * b * c * dA - (det + 2.f * b * b) * dB + 2.f * a * b * dC);
        float d_opacity = d_w;
        if (s->aa) {
            float det0 = a0 * c0 - b * b;
            float ratio = det0 / det;
            float hs = sqrtf(fmaxf(AA_FLOOR, ratio));
            d_opacity = d_w * hs;
            if (ratio > AA_FLOOR) {
                float d_ratio = d_w * s->opac[i] / (2.0f * hs);
                dL_da += d_ratio * (c0 / det - det0 * c / (det * det));
                dL_dc += d_ratio * (a0 / det - det0 * a / (det * det));
                dL_db += d_ratio * (-2.f * b / det + det0 * 2.f * b / (det * det));
            }
        }
        if (g_opac) g_opac[i] = d_opacity;

        /* cov2D -> Sigma (6) and M */
        const float *M0 = cv.M0, *M1 = cv.M1;
        float dS[6];
        dS[0] = M0[0] * M0[0] * dL_da + M0[0] * M1[0] * dL_db + M1[0] * M1[0] * dL_dc;
        dS[3] = M0[1] * M0[1] * dL_da + M0[1] * M1[1] * dL_db + M1[1] * M1[1] * dL_dc;
        dS[5] = M0[2] * M0[2] * dL_da + M0[2] * M1[2] * dL_db + M1[2] * M1[2] * dL_dc;
        dS[1] = 2.f * M0[0] * M0[1] * dL_da + (M0[0] * M1[1] + M0[1] * M1[0]) * dL_db + 2.f * M1[0] * M1[1] * dL_dc;
        dS[2] = 2.f * M0[0] * M0[2] * dL_da + (M0[0] * M1[2] + M0[2] * M1[0]) * dL_db + 2.f * M1[0] * M1[2] * dL_dc;
        dS[4] = 2.f * M0[1] * M0[2] * dL_da + (M0[1] * M1[2] + M0[2] * M1[1]) * dL_db + 2.f * M1[1] * M1[2] * dL_dc;
        float S[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
        float SM0[3], SM1[3], dM0[3], dM1[3];
        for (int r = 0; r < 3; ++r) {
            SM0[r] = S[3 * r] * M0[0] + S[3 * r + 1] * M0[1] + S[3 * r + 2] * M0[2];
            SM1[r] = S[3 * r] * M1[0] + S[3 * r + 1] * M1[1] + S[3 * r + 2] * M1[2];
        }
        for (int r = 0; r < 3; ++r) {
            dM0[r] = 2.f * dL_da * SM0[r] + dL_db * SM1[r];
            dM1[r] = 2.f * dL_dc * SM1[r] + dL_db * SM0[r];
        }
        float dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
        for (int cc = 0; cc < 3; ++cc) {
            dJ00 += dM0[cc] * view[4 * cc + 0]; dJ02 += dM0[cc] * view[4 * cc + 2];
            dJ11 += dM1[cc] * view[4 * cc + 1]; dJ12 += dM1[cc] * view[4 * cc + 2];
        }
        float tz = cv.tz, tz2 = 1.0f / (tz * tz), tz3 = tz2 / tz;
        float dtx = cv.cx ? 0.0f : -fx * tz2 * dJ02;
        float dty = cv.cy ? 0.0f : -fy * tz2 * dJ12;
        float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + 2.f * fx * cv.tx * tz3 * dJ02 + 2.f * fy * cv.ty * tz3 * dJ12;
        dtz -= d_invd * tz2;
        for (int cc = 0; cc < 3; ++cc)
            gm[cc] += dtx * view[4 * cc + 0] + dty * view[4 * cc + 1] + dtz * view[4 * cc + 2];

        /* 2D mean (ndc) -> 3D mean */
        {
            float hx = proj[0] * p[0] + proj[4] * p[1] + proj[8] * p[2] + proj[12];
            float hy = proj[1] * p[0] + proj[5] * p[1] + proj[9] * p[2] + proj[13];
            float hw = proj[3] * p[0] + proj[7] * p[1] + proj[11] * p[2] + proj[15];
            float pw = 1.0f / (hw + 0.0000001f);
            float mul1 = hx * pw * pw, mul2 = hy * pw * pw;
            for (int cc = 0; cc < 3; ++cc)
                gm[cc] += (proj[4 * cc + 0] * pw - proj[4 * cc + 3] * mul1) * d_m2x
                        + (proj[4 * cc + 1] * pw - proj[4 * cc + 3] * mul2) * d_m2y;
        }
        if (g_means3D) for (int cc = 0; cc < 3; ++cc) g_means3D[3 * i + cc] = gm[cc];

        /* Sigma -> scale / rotation */
        if (s->has_cov_precomp) {
            if (g_cov3D) for (int k = 0; k < 6; ++k) g_cov3D[6 * i + k] = dS[k];
        } else {
            const float *q = s->rots + 4 * i, *sc = s->scales + 3 * i;
            float mod = s->scale_modifier;
            float R[9]; quat_to_R(q, R);
            float sp[3] = {mod * sc[0], mod * sc[1], mod * sc[2]};
            /* G symmetric: diag = dS_ii, off = 0.5 * dS_ij ; dL/dL = 2 G L, L = R diag(sp) */
            float Gm[9] = {dS[0], 0.5f * dS[1], 0.5f * dS[2], 0.5f * dS[1], dS[3], 0.5f * dS[4], 0.5f * dS[2], 0.5f * dS[4], dS[5]};
            float dLm[9];
            for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) {
                float v = 0.f;
                for (int j = 0; j < 3; ++j) v += Gm[3 * r + j] * (R[3 * j + k] * sp[k]);
                dLm[3 * r + k] = 2.f * v;
            }
            float dR[9];
            for (int k = 0; k < 3; ++k) {
                float ds = 0.f;
                for (int r = 0; r < 3; ++r) { ds += dLm[3 * r + k] * R[3 * r + k]; dR[3 * r + k] = dLm[3 * r + k] * sp[k]; }
                if (g_scales) g_scales[3 * i + k] = mod * ds;
            }
            if (g_rots) {
                float r = q[0], x = q[1], y = q[2], z = q[3];
                g_rots[4 * i + 0] = 2.f * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
                g_rots[4 * i + 1] = 2.f * (y * dR[1] + z * dR[2] + y * dR[3] - 2.f * x * dR[4] - r * dR[5] + z * dR[6] + r * dR[7] - 2.f * x * dR[8]);
                g_rots[4 * i + 2] = 2.f * (-2.f * y * dR[0] + x * dR[1] + r * dR[2] + x * dR[3] + z * dR[5] - r * dR[6] + z * dR[7] - 2.f * y * dR[8]);
                g_rots[4 * i + 3] = 2.f * (-2.f * z * dR[0] - r * dR[1] + x * dR[2] + r * dR[3] - 2.f * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
            }
        }
    }
    free(acc);
}
