/*
 * grpg_oracle.c -- CPU restatement of the reference rasterizer.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product path (gaussianrpg_b200/) never does.
 *
 * What it restates (paths relative to /root/reference/submodules/diff-gaussian-rasterization):
 *   forward  per Gaussian : cuda_rasterizer/forward.cu:155-256 (preprocessCUDA) with
 *                            auxiliary.h:139-164 (in_frustum), forward.cu:118-152 (computeCov3D),
 *                            forward.cu:74-113 (computeCov2D), forward.cu:20-71 (computeColorFromSH),
 *                            auxiliary.h:41-56 (ndc2Pix, getRect)
 *   binning               : rasterizer_impl.cu:70-111 (duplicateWithKeys), :303-311 (stable sort
 *                            on (tile<<32 | depth bits)), :116-138 (identifyTileRanges)
 *   forward blend          : forward.cu:340-467 (renderCUDA)
 *   backward blend         : backward.cu:415-641 (renderCUDA)
 *   backward per Gaussian  : backward.cu:144-274 (computeCov2DCUDA), :346-412 (preprocessCUDA),
 *                            :20-139 (computeColorFromSH), :278-341 (computeCov3D)
 *
 * Rounding.  All arithmetic is IEEE binary32 with round-to-nearest; the file must be compiled
 * with -ffp-contract=off so the compiler never fuses on its own.  Where the reference build
 * (nvcc 12.9, sm_100a, default -fmad=true) contracts a multiply-add on a value that decides a
 * tile/key index, the same fused operation is written explicitly with fmaf(); the contraction
 * pattern was read from the SASS of that build: a sum of products p0 + p1 + p2 [+ c] evaluates
 * as fma(p2, fma(p0, mul(p1))) [+ c], with the exceptions noted inline (shared quaternion
 * products).  Division, sqrt and reciprocal are correctly rounded on both sides.  exp() differs:
 * the GPU uses ex2.approx inside expf (<= 2 ulp), here glibc's expf -- colours agree to ~1e-6,
 * not bit-exactly.  Parity status: pinned against outputs of the reference extension itself
 * (tests/golden/*.npz, generated on a B200 by tests/golden/make_golden.py).
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BLOCK_X 16
#define BLOCK_Y 16
#define BLOCK_SIZE 256

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float fmn(float a, float b) { return fminf(a, b); }
static inline float fmx(float a, float b) { return fmaxf(a, b); }
/* CUDA float->int conversions saturate and map NaN to 0 */
static inline int f2i_rz(float f) {
    if (isnan(f)) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (-2147483647 - 1);
    return (int)f;
}
static inline int f2i_ru(float f) { return f2i_rz(ceilf(f)); }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* matrix[r]*x + matrix[4+r]*y + matrix[8+r]*z + matrix[12+r]  (auxiliary.h:58-76) */
static inline float xform(const float* m, int r, float x, float y, float z) {
    return fmaf(z, m[8 + r], fmaf(x, m[r], y * m[4 + r])) + m[12 + r];
}
static inline float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
    return fmaf(a2, b2, fmaf(a0, b0, a1 * b1));
}

/* forward.cu:118-152 */
static void compute_cov3d(const float* scale, float mod, const float* rot, float* cov3D) {
    const float sx = scale[0] * mod, sy = scale[1] * mod, sz = scale[2] * mod;
    const float r = rot[0], x = rot[1], y = rot[2], z = rot[3]; /* not normalised, forward.cu:127 */
    /* shared products keep their own rounding in the reference build */
    const float xz = x * z, rx = r * x, rz = r * z, yy = y * y, zz = z * z;
    const float R00 = 1.0f - 2.0f * (yy + zz);
    const float R01 = 2.0f * fmaf(x, y, -rz);
    const float R02 = 2.0f * fmaf(r, y, xz);
    const float R10 = 2.0f * fmaf(x, y, rz);
    const float R11 = 1.0f - 2.0f * fmaf(x, x, zz);
    const float R12 = 2.0f * fmaf(y, z, -rx);
    const float R20 = 2.0f * fmaf(-r, y, xz);
    const float R21 = 2.0f * fmaf(y, z, rx);
    const float R22 = 1.0f - 2.0f * fmaf(x, x, yy);
    /* M = S * R (glm, column-major): M[c][r] = s_r * R[c][r] */
    const float M[3][3] = {{sx * R00, sy * R01, sz * R02}, {sx * R10, sy * R11, sz * R12}, {sx * R20, sy * R21, sz * R22}};
    /* Sigma = transpose(M) * M : Sigma[c][r] = sum_k M[r][k] M[c][k] */
    cov3D[0] = dot3(M[0][0], M[0][0], M[0][1], M[0][1], M[0][2], M[0][2]);
    cov3D[1] = dot3(M[1][0], M[0][0], M[1][1], M[0][1], M[1][2], M[0][2]);
    cov3D[2] = dot3(M[2][0], M[0][0], M[2][1], M[0][1], M[2][2], M[0][2]);
    cov3D[3] = dot3(M[1][0], M[1][0], M[1][1], M[1][1], M[1][2], M[1][2]);
    cov3D[4] = dot3(M[2][0], M[1][0], M[2][1], M[1][1], M[2][2], M[1][2]);
    cov3D[5] = dot3(M[2][0], M[2][0], M[2][1], M[2][1], M[2][2], M[2][2]);
}

/* forward.cu:74-113; returns (a,b,c) and the intermediates T (2x3) for the backward */
static void compute_cov2d(float x, float y, float z, float fx, float fy, float tanx, float tany, const float* cov3D,
                          const float* v, float* abc, float* T0, float* T1) {
    float tx = xform(v, 0, x, y, z), ty = xform(v, 1, x, y, z);
    const float tz = xform(v, 2, x, y, z);
    const float limx = 1.3f * tanx, limy = 1.3f * tany;
    const float cx = fmn(limx, fmx(-limx, tx / tz)), cy = fmn(limy, fmx(-limy, ty / tz));
    /* t.x = clamp * t.z is only used negated and multiplied: -(fx * t.x) / (tz*tz) */
    const float tz2 = tz * tz;
    const float J00 = fx / tz, J02 = ((tz * -cx) * fx) / tz2;
    const float J11 = fy / tz, J12 = ((tz * -cy) * fy) / tz2;
    for (int r = 0; r < 3; ++r) {
        /* T = W * J with W[0][r]=v[4r], W[1][r]=v[4r+1], W[2][r]=v[4r+2]; zero entries of J drop out */
        T0[r] = fmaf(v[4 * r + 2], J02, v[4 * r] * J00);
        T1[r] = fmaf(v[4 * r + 2], J12, v[4 * r + 1] * J11);
    }
    const float B[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
    float X0[3], X1[3]; /* X[c][r] for r = 0,1 */
    for (int c = 0; c < 3; ++c) {
        X0[c] = dot3(T0[0], B[c][0], T0[1], B[c][1], T0[2], B[c][2]);
        X1[c] = dot3(T1[0], B[c][0], T1[1], B[c][1], T1[2], B[c][2]);
    }
    abc[0] = dot3(X0[0], T0[0], X0[1], T0[1], X0[2], T0[2]) + 0.3f;
    abc[1] = dot3(X1[0], T0[0], X1[1], T0[1], X1[2], T0[2]);
    abc[2] = dot3(X1[0], T1[0], X1[1], T1[1], X1[2], T1[2]) + 0.3f;
}

/* forward.cu:20-71 */
static void sh_to_rgb(int deg, const float* sh, const float* p, const float* campos, float* rgb, uint8_t* clamped) {
    const float dx = p[0] - campos[0], dy = p[1] - campos[1], dz = p[2] - campos[2];
    const float len = sqrtf(fmaf(dz, dz, fmaf(dx, dx, dy * dy)));
    const float x = dx / len, y = dy / len, z = dz / len;
    float res[3];
    for (int ch = 0; ch < 3; ++ch) res[ch] = SH_C0 * sh[ch];
    if (deg > 0) {
        const float c1y = SH_C1 * y, c1z = SH_C1 * z, c1x = SH_C1 * x;
        for (int ch = 0; ch < 3; ++ch)
            res[ch] = fmaf(-c1x, sh[9 + ch], fmaf(c1z, sh[6 + ch], fmaf(-c1y, sh[3 + ch], res[ch])));
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            const float k4 = SH_C2[0] * xy, k5 = SH_C2[1] * yz, k6 = SH_C2[2] * ((2.0f * zz - xx) - yy);
            const float k7 = SH_C2[3] * xz, k8 = SH_C2[4] * (xx - yy);
            for (int ch = 0; ch < 3; ++ch) {
                float r = fmaf(k4, sh[12 + ch], res[ch]);
                r = fmaf(k5, sh[15 + ch], r);
                r = fmaf(k6, sh[18 + ch], r);
                r = fmaf(k7, sh[21 + ch], r);
                res[ch] = fmaf(k8, sh[24 + ch], r);
            }
            if (deg > 2) {
                const float k9 = (SH_C3[0] * y) * fmaf(3.0f, xx, -yy);
                const float k10 = (SH_C3[1] * xy) * z;
                const float f4 = fmaf(4.0f, zz, -xx) - yy;
                const float k11 = (SH_C3[2] * y) * f4;
                const float k12 = (SH_C3[3] * z) * fmaf(-3.0f, yy, fmaf(-3.0f, xx, 2.0f * zz));
                const float k13 = (SH_C3[4] * x) * f4;
                const float k14 = (SH_C3[5] * z) * (xx - yy);
                const float k15 = (SH_C3[6] * x) * fmaf(-3.0f, yy, xx);
                for (int ch = 0; ch < 3; ++ch) {
                    float r = fmaf(k9, sh[27 + ch], res[ch]);
                    r = fmaf(k10, sh[30 + ch], r);
                    r = fmaf(k11, sh[33 + ch], r);
                    r = fmaf(k12, sh[36 + ch], r);
                    r = fmaf(k13, sh[39 + ch], r);
                    r = fmaf(k14, sh[42 + ch], r);
                    res[ch] = fmaf(k15, sh[45 + ch], r);
                }
            }
        }
    }
    for (int ch = 0; ch < 3; ++ch) {
        const float r = res[ch] + 0.5f;
        clamped[ch] = r < 0.0f;
        rgb[ch] = r < 0.0f ? 0.0f : r;
    }
}

typedef struct {
    uint64_t key;
    uint32_t val;
    uint32_t seq;
} inst_t;

static int inst_cmp(const void* a, const void* b) {
    const inst_t* x = (const inst_t*)a;
    const inst_t* y = (const inst_t*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq ? 1 : 0); /* stable, as the LSD radix sort */
}

/* ---- forward, stage 1: per Gaussian ------------------------------------------------------ */
void oracle_preprocess(int P, int D, int M, const float* means3D, const float* scales, float scale_modifier,
                       const float* rotations, const float* opacities, const float* shs, const float* cov3D_precomp,
                       const float* colors_precomp, const float* view, const float* proj, const float* campos, int W,
                       int H, float tan_fovx, float tan_fovy,
                       /* outputs, all sized per Gaussian */
                       int* radii, float* means2D /*[P,2]*/, float* depths, float* cov3Ds /*[P,6]*/, float* rgb /*[P,3]*/,
                       float* conic_opacity /*[P,4]*/, uint32_t* tiles_touched, uint8_t* clamped /*[P,3]*/,
                       uint32_t* rects /*[P,4] xmin,ymin,xmax,ymax*/) {
    const float focal_y = H / (2.0f * tan_fovy), focal_x = W / (2.0f * tan_fovx);
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    for (int i = 0; i < P; ++i) {
        radii[i] = 0;
        tiles_touched[i] = 0;
        const float x = means3D[3 * i], y = means3D[3 * i + 1], z = means3D[3 * i + 2];
        const float vz = xform(view, 2, x, y, z);
        if (vz <= 0.2f) continue; /* auxiliary.h:154 */
        const float hx = xform(proj, 0, x, y, z), hy = xform(proj, 1, x, y, z), hw = xform(proj, 3, x, y, z);
        const float p_w = 1.0f / (hw + 0.0000001f);
        const float ndc_x = hx * p_w, ndc_y = hy * p_w;
        const float* cov3D;
        if (cov3D_precomp) cov3D = cov3D_precomp + 6 * (size_t)i;
        else {
            compute_cov3d(scales + 3 * i, scale_modifier, rotations + 4 * i, cov3Ds + 6 * (size_t)i);
            cov3D = cov3Ds + 6 * (size_t)i;
        }
        float abc[3], T0[3], T1[3];
        compute_cov2d(x, y, z, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, view, abc, T0, T1);
        const float a = abc[0], b = abc[1], c = abc[2];
        const float det = fmaf(a, c, -(b * b));
        if (det == 0.0f) continue;
        const float det_inv = 1.0f / det;
        const float conic[3] = {c * det_inv, b * -det_inv, a * det_inv};
        const float mid = 0.5f * (a + c);
        const float s = sqrtf(fmx(0.1f, fmaf(mid, mid, -det)));
        const float lam = fmx(mid + s, mid - s);
        const int my_radius = f2i_ru(3.0f * sqrtf(lam));
        /* ndc2Pix in double (auxiliary.h:41-44): ((v + 1.0) * S - 1.0) * 0.5, the multiply-subtract fused */
        const float px = (float)(fma((double)ndc_x + 1.0, (double)W, -1.0) * 0.5);
        const float py = (float)(fma((double)ndc_y + 1.0, (double)H, -1.0) * 0.5);
        const float rf = (float)my_radius;
        const int x0 = imin(gx, imax(0, f2i_rz((px - rf) * 0.0625f)));
        const int y0 = imin(gy, imax(0, f2i_rz((py - rf) * 0.0625f)));
        const int x1 = imin(gx, imax(0, f2i_rz((((px + rf) + 16.0f) - 1.0f) * 0.0625f)));
        const int y1 = imin(gy, imax(0, f2i_rz((((py + rf) + 16.0f) - 1.0f) * 0.0625f)));
        if ((x1 - x0) * (y1 - y0) == 0) continue;
        if (!colors_precomp) sh_to_rgb(D, shs + (size_t)i * M * 3, means3D + 3 * i, campos, rgb + 3 * i, clamped + 3 * i);
        depths[i] = vz;
        radii[i] = my_radius;
        means2D[2 * i] = px;
        means2D[2 * i + 1] = py;
        conic_opacity[4 * i] = conic[0];
        conic_opacity[4 * i + 1] = conic[1];
        conic_opacity[4 * i + 2] = conic[2];
        conic_opacity[4 * i + 3] = opacities[i];
        tiles_touched[i] = (uint32_t)((x1 - x0) * (y1 - y0));
        rects[4 * i] = x0; rects[4 * i + 1] = y0; rects[4 * i + 2] = x1; rects[4 * i + 3] = y1;
    }
}

/* ---- forward, stage 2: instances, order, ranges -------------------------------------------- */
/* returns R; keys/point_list must hold sum(tiles_touched) entries */
long long oracle_binning(int P, int W, int H, const int* radii, const float* depths, const uint32_t* rects,
                         uint64_t* keys, uint32_t* point_list, uint32_t* ranges /*[tiles,2]*/) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    long long R = 0;
    for (int i = 0; i < P; ++i)
        if (radii[i] > 0) R += (long long)(rects[4 * i + 2] - rects[4 * i]) * (rects[4 * i + 3] - rects[4 * i + 1]);
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy);
    if (R == 0) return 0;
    inst_t* inst = (inst_t*)malloc(sizeof(inst_t) * (size_t)R);
    size_t n = 0;
    for (int i = 0; i < P; ++i) {
        if (radii[i] <= 0) continue;
        for (uint32_t ty = rects[4 * i + 1]; ty < rects[4 * i + 3]; ++ty)
            for (uint32_t tx = rects[4 * i]; tx < rects[4 * i + 2]; ++tx) {
                inst[n].key = ((uint64_t)(ty * (uint32_t)gx + tx) << 32) | f2u(depths[i]);
                inst[n].val = (uint32_t)i;
                inst[n].seq = (uint32_t)n;
                ++n;
            }
    }
    qsort(inst, (size_t)R, sizeof(inst_t), inst_cmp);
    for (long long k = 0; k < R; ++k) {
        keys[k] = inst[k].key;
        point_list[k] = inst[k].val;
        const uint32_t cur = (uint32_t)(inst[k].key >> 32);
        if (k == 0) ranges[2 * cur] = 0;
        else {
            const uint32_t prev = (uint32_t)(inst[k - 1].key >> 32);
            if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)k; ranges[2 * cur] = (uint32_t)k; }
        }
        if (k == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
    }
    free(inst);
    return R;
}

static inline float blend_power(const float* con, float dx, float dy) {
    /* -0.5f * (A dx dx + C dy dy) - B dx dy as the reference build evaluates it */
    const float tc = dy * (dy * con[2]);
    const float tb = dy * (dx * con[1]);
    return fmaf(fmaf(dx, dx * con[0], tc), -0.5f, -tb);
}

/* ---- forward, stage 3: per-tile blend; tile_y0/tile_y1 bound the tile rows rendered ---------- */
void oracle_blend_forward(int W, int H, int S, int tile_y0, int tile_y1, const uint32_t* ranges,
                          const uint32_t* point_list, const float* means2D, const float* features /*[P,3]*/,
                          const float* depths, const float* semantics /*[P,S]*/, const float* conic_opacity,
                          const float* bg, float* out_color, float* out_depth, float* out_alpha, float* out_semantic,
                          uint32_t* n_contrib) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X;
    const size_t hw = (size_t)H * W;
    for (int ty = tile_y0; ty < tile_y1; ++ty)
        for (int tx = 0; tx < gx; ++tx) {
            const uint32_t r0 = ranges[2 * (ty * gx + tx)], r1 = ranges[2 * (ty * gx + tx) + 1];
            for (int ly = 0; ly < BLOCK_Y; ++ly)
                for (int lx = 0; lx < BLOCK_X; ++lx) {
                    const int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
                    if (px >= W || py >= H) continue;
                    const size_t pid = (size_t)py * W + px;
                    float T = 1.0f, C[3] = {0, 0, 0}, weight = 0.0f, Dp = 0.0f;
                    uint32_t contributor = 0, last = 0;
                    for (int ch = 0; ch < S; ++ch) out_semantic[ch * hw + pid] = 0.0f;
                    for (uint32_t k = r0; k < r1; ++k) {
                        ++contributor;
                        const uint32_t id = point_list[k];
                        const float dx = means2D[2 * id] - (float)px, dy = means2D[2 * id + 1] - (float)py;
                        const float* con = conic_opacity + 4 * (size_t)id;
                        const float power = blend_power(con, dx, dy);
                        if (power > 0.0f) continue;
                        const float alpha = fmn(0.99f, con[3] * expf(power));
                        if (alpha < 1.0f / 255.0f) continue;
                        const float test_T = T * (1.0f - alpha);
                        if (test_T < 0.0001f) break; /* pixel done, this instance not blended */
                        for (int ch = 0; ch < 3; ++ch) C[ch] = fmaf(T, alpha * features[3 * (size_t)id + ch], C[ch]);
                        for (int ch = 0; ch < S; ++ch)
                            out_semantic[ch * hw + pid] =
                                fmaf(T, alpha * semantics[(size_t)id * S + ch], out_semantic[ch * hw + pid]);
                        weight = fmaf(T, alpha, weight);
                        Dp = fmaf(T, alpha * depths[id], Dp);
                        T = test_T;
                        last = contributor;
                    }
                    n_contrib[pid] = last;
                    for (int ch = 0; ch < 3; ++ch) out_color[ch * hw + pid] = fmaf(bg[ch], T, C[ch]);
                    out_alpha[pid] = weight;
                    out_depth[pid] = Dp;
                }
        }
}

/* ---- backward blend (backward.cu:415-641): accumulates into zero-initialised outputs -------- */
void oracle_blend_backward(int W, int H, int S, const uint32_t* ranges, const uint32_t* point_list, const float* bg,
                           const float* means2D, const float* conic_opacity, const float* colors, const float* depths,
                           const float* semantics, const float* alphas, const uint32_t* n_contrib,
                           const float* dL_dpixels, const float* dL_dpixel_depths, const float* dL_dalphas,
                           const float* dL_dpixel_semantics, double* dL_dmean2D /*[P,3]*/, double* dL_dconic /*[P,4]*/,
                           double* dL_dopacity, double* dL_dcolors /*[P,3]*/, double* dL_ddepths,
                           double* dL_dsemantics /*[P,S]*/) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const size_t hw = (size_t)H * W;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    float* acc_sem = (float*)calloc((size_t)(S > 0 ? S : 1) * 3, sizeof(float));
    float* last_sem = acc_sem + (S > 0 ? S : 1);
    float* dsem = last_sem + (S > 0 ? S : 1);
    for (int ty = 0; ty < gy; ++ty)
        for (int tx = 0; tx < gx; ++tx) {
            const uint32_t r0 = ranges[2 * (ty * gx + tx)], r1 = ranges[2 * (ty * gx + tx) + 1];
            for (int ly = 0; ly < BLOCK_Y; ++ly)
                for (int lx = 0; lx < BLOCK_X; ++lx) {
                    const int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
                    if (px >= W || py >= H) continue;
                    const size_t pid = (size_t)py * W + px;
                    const float T_final = 1.0f - alphas[pid];
                    float T = T_final;
                    const uint32_t last_contributor = n_contrib[pid];
                    float accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0};
                    float accum_depth = 0, last_depth = 0, accum_alpha = 0, last_alpha = 0;
                    const float dpix[3] = {dL_dpixels[pid], dL_dpixels[hw + pid], dL_dpixels[2 * hw + pid]};
                    const float dpd = dL_dpixel_depths[pid], dpa = dL_dalphas[pid];
                    for (int ch = 0; ch < S; ++ch) {
                        acc_sem[ch] = 0; last_sem[ch] = 0;
                        dsem[ch] = dL_dpixel_semantics[ch * hw + pid];
                    }
                    float bg_dot = 0;
                    for (int ch = 0; ch < 3; ++ch) bg_dot += bg[ch] * dpix[ch];
                    for (uint32_t k = r1; k-- > r0;) {
                        if (k - r0 >= last_contributor) continue;
                        const uint32_t id = point_list[k];
                        const float dx = means2D[2 * id] - (float)px, dy = means2D[2 * id + 1] - (float)py;
                        const float* con = conic_opacity + 4 * (size_t)id;
                        const float power = blend_power(con, dx, dy);
                        if (power > 0.0f) continue;
                        const float G = expf(power);
                        const float alpha = fmn(0.99f, con[3] * G);
                        if (alpha < 1.0f / 255.0f) continue;
                        T = T / (1.0f - alpha);
                        const float w = alpha * T;
                        float dL_dopa = 0.0f;
                        for (int ch = 0; ch < 3; ++ch) {
                            const float c = colors[3 * (size_t)id + ch];
                            accum_rec[ch] = last_alpha * last_color[ch] + (1.0f - last_alpha) * accum_rec[ch];
                            last_color[ch] = c;
                            dL_dopa += (c - accum_rec[ch]) * dpix[ch];
                            dL_dcolors[3 * (size_t)id + ch] += (double)(w * dpix[ch]);
                        }
                        for (int ch = 0; ch < S; ++ch) {
                            const float sv = semantics[(size_t)id * S + ch];
                            acc_sem[ch] = last_alpha * last_sem[ch] + (1.0f - last_alpha) * acc_sem[ch];
                            last_sem[ch] = sv;
                            dL_dopa += (sv - acc_sem[ch]) * dsem[ch];
                            dL_dsemantics[(size_t)id * S + ch] += (double)(w * dsem[ch]);
                        }
                        const float cd = depths[id];
                        accum_depth = last_alpha * last_depth + (1.0f - last_alpha) * accum_depth;
                        last_depth = cd;
                        dL_dopa += (cd - accum_depth) * dpd;
                        dL_ddepths[id] += (double)(w * dpd);
                        accum_alpha = last_alpha + (1.0f - last_alpha) * accum_alpha;
                        dL_dopa += (1.0f - accum_alpha) * dpa;
                        dL_dopa *= T;
                        last_alpha = alpha;
                        dL_dopa += (-T_final / (1.0f - alpha)) * bg_dot;
                        const float dL_dG = con[3] * dL_dopa;
                        const float gdx = G * dx, gdy = G * dy;
                        const float dG_ddelx = -gdx * con[0] - gdy * con[1];
                        const float dG_ddely = -gdy * con[2] - gdx * con[1];
                        const float gx_ = dL_dG * dG_ddelx * ddelx_dx, gy_ = dL_dG * dG_ddely * ddely_dy;
                        dL_dmean2D[3 * (size_t)id] += (double)gx_;
                        dL_dmean2D[3 * (size_t)id + 1] += (double)gy_;
                        dL_dmean2D[3 * (size_t)id + 2] += (double)(fabsf(gx_) + fabsf(gy_));
                        dL_dconic[4 * (size_t)id] += (double)(-0.5f * gdx * dx * dL_dG);
                        dL_dconic[4 * (size_t)id + 1] += (double)(-0.5f * gdx * dy * dL_dG);
                        dL_dconic[4 * (size_t)id + 3] += (double)(-0.5f * gdy * dy * dL_dG);
                        dL_dopacity[id] += (double)(G * dL_dopa);
                    }
                }
        }
    free(acc_sem);
}

/* ---- backward per Gaussian (backward.cu:144-274 then :346-412) ------------------------------ */
void oracle_preprocess_backward(int P, int D, int M, const float* means3D, const int* radii, const float* shs,
                                const uint8_t* clamped /*[P,3]*/, const float* scales, const float* rotations,
                                float scale_modifier, const float* cov3Ds, const float* view, const float* proj,
                                int W, int H, float tan_fovx, float tan_fovy, const float* campos,
                                const float* dL_dmean2D /*[P,3]*/, const float* dL_dconic /*[P,4]*/,
                                const float* dL_dcolor /*[P,3]*/, const float* dL_ddepth,
                                float* dL_dmean3D /*[P,3]*/, float* dL_dcov3D /*[P,6]*/, float* dL_dsh /*[P,M,3]*/,
                                float* dL_dscale /*[P,3]*/, float* dL_drot /*[P,4]*/) {
    const float h_y = H / (2.0f * tan_fovy), h_x = W / (2.0f * tan_fovx);
    for (int i = 0; i < P; ++i) {
        if (!(radii[i] > 0)) continue;
        const float mx = means3D[3 * i], my = means3D[3 * i + 1], mz = means3D[3 * i + 2];
        const float* cov3D = cov3Ds + 6 * (size_t)i;
        const float dcon[3] = {dL_dconic[4 * i], dL_dconic[4 * i + 1], dL_dconic[4 * i + 3]};
        float tx = view[0] * mx + view[4] * my + view[8] * mz + view[12];
        float ty = view[1] * mx + view[5] * my + view[9] * mz + view[13];
        const float tz = view[2] * mx + view[6] * my + view[10] * mz + view[14];
        const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
        const float txtz = tx / tz, tytz = ty / tz;
        tx = fmn(limx, fmx(-limx, txtz)) * tz;
        ty = fmn(limy, fmx(-limy, tytz)) * tz;
        const float xm = (txtz < -limx || txtz > limx) ? 0.0f : 1.0f;
        const float ym = (tytz < -limy || tytz > limy) ? 0.0f : 1.0f;
        /* glm: J columns (J00,0,J02),(0,J11,J12),0 ; W[c][r]; T = W*J */
        const float J00 = h_x / tz, J02 = -(h_x * tx) / (tz * tz), J11 = h_y / tz, J12 = -(h_y * ty) / (tz * tz);
        float T[2][3];
        for (int r = 0; r < 3; ++r) {
            T[0][r] = view[4 * r] * J00 + view[4 * r + 2] * J02;
            T[1][r] = view[4 * r + 1] * J11 + view[4 * r + 2] * J12;
        }
        const float V[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
        float TV[2][3];
        for (int q = 0; q < 2; ++q)
            for (int k = 0; k < 3; ++k) TV[q][k] = T[q][0] * V[k][0] + T[q][1] * V[k][1] + T[q][2] * V[k][2];
        const float a = TV[0][0] * T[0][0] + TV[0][1] * T[0][1] + TV[0][2] * T[0][2] + 0.3f;
        const float b = TV[0][0] * T[1][0] + TV[0][1] * T[1][1] + TV[0][2] * T[1][2];
        const float c = TV[1][0] * T[1][0] + TV[1][1] * T[1][1] + TV[1][2] * T[1][2] + 0.3f;
        const float denom = a * c - b * b;
        float da = 0, db = 0, dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float* dcov = dL_dcov3D + 6 * (size_t)i;
        if (denom2inv != 0) {
            da = denom2inv * (-c * c * dcon[0] + 2 * b * c * dcon[1] + (denom - a * c) * dcon[2]);
            dc = denom2inv * (-a * a * dcon[2] + 2 * a * b * dcon[1] + (denom - a * c) * dcon[0]);
            db = denom2inv * 2 * (b * c * dcon[0] - (denom + 2 * b * b) * dcon[1] + a * b * dcon[2]);
            dcov[0] = T[0][0] * T[0][0] * da + T[0][0] * T[1][0] * db + T[1][0] * T[1][0] * dc;
            dcov[3] = T[0][1] * T[0][1] * da + T[0][1] * T[1][1] * db + T[1][1] * T[1][1] * dc;
            dcov[5] = T[0][2] * T[0][2] * da + T[0][2] * T[1][2] * db + T[1][2] * T[1][2] * dc;
            dcov[1] = 2 * T[0][0] * T[0][1] * da + (T[0][0] * T[1][1] + T[0][1] * T[1][0]) * db + 2 * T[1][0] * T[1][1] * dc;
            dcov[2] = 2 * T[0][0] * T[0][2] * da + (T[0][0] * T[1][2] + T[0][2] * T[1][0]) * db + 2 * T[1][0] * T[1][2] * dc;
            dcov[4] = 2 * T[0][2] * T[0][1] * da + (T[0][1] * T[1][2] + T[0][2] * T[1][1]) * db + 2 * T[1][1] * T[1][2] * dc;
        } else
            for (int k = 0; k < 6; ++k) dcov[k] = 0;
        const float dT00 = 2 * TV[0][0] * da + TV[1][0] * db, dT01 = 2 * TV[0][1] * da + TV[1][1] * db,
                    dT02 = 2 * TV[0][2] * da + TV[1][2] * db;
        const float dT10 = 2 * TV[1][0] * dc + TV[0][0] * db, dT11 = 2 * TV[1][1] * dc + TV[0][1] * db,
                    dT12 = 2 * TV[1][2] * dc + TV[0][2] * db;
        const float dJ00 = view[0] * dT00 + view[4] * dT01 + view[8] * dT02;
        const float dJ02 = view[2] * dT00 + view[6] * dT01 + view[10] * dT02;
        const float dJ11 = view[1] * dT10 + view[5] * dT11 + view[9] * dT12;
        const float dJ12 = view[2] * dT10 + view[6] * dT11 + view[10] * dT12;
        const float tzi = 1.0f / tz, tz2 = tzi * tzi, tz3 = tz2 * tzi;
        const float dtx = xm * -h_x * tz2 * dJ02, dty = ym * -h_y * tz2 * dJ12;
        const float dtz = -h_x * tz2 * dJ00 - h_y * tz2 * dJ11 + (2 * h_x * tx) * tz3 * dJ02 + (2 * h_y * ty) * tz3 * dJ12;
        float dm[3] = {view[0] * dtx + view[1] * dty + view[2] * dtz, view[4] * dtx + view[5] * dty + view[6] * dtz,
                       view[8] * dtx + view[9] * dty + view[10] * dtz};
        /* backward.cu:371-389 */
        const float m_w = 1.0f / ((proj[3] * mx + proj[7] * my + proj[11] * mz + proj[15]) + 0.0000001f);
        const float mul1 = (proj[0] * mx + proj[4] * my + proj[8] * mz + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * mx + proj[5] * my + proj[9] * mz + proj[13]) * m_w * m_w;
        const float g2x = dL_dmean2D[3 * i], g2y = dL_dmean2D[3 * i + 1];
        dm[0] += (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
        dm[1] += (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
        dm[2] += (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;
        /* backward.cu:391-403 */
        const float mul3 = view[2] * mx + view[6] * my + view[10] * mz + view[14];
        dm[0] += (view[2] - view[3] * mul3) * dL_ddepth[i];
        dm[1] += (view[6] - view[7] * mul3) * dL_ddepth[i];
        dm[2] += (view[10] - view[11] * mul3) * dL_ddepth[i];
        if (shs) { /* backward.cu:20-139 */
            const float* sh = shs + (size_t)i * M * 3;
            float* dsh = dL_dsh + (size_t)i * M * 3;
            float dRGB[3];
            for (int ch = 0; ch < 3; ++ch) dRGB[ch] = clamped[3 * i + ch] ? 0.0f : dL_dcolor[3 * i + ch];
            const float ox = mx - campos[0], oy = my - campos[1], oz = mz - campos[2];
            const float len = sqrtf(ox * ox + oy * oy + oz * oz);
            const float x = ox / len, y = oy / len, z = oz / len;
            float dx_[3] = {0, 0, 0}, dy_[3] = {0, 0, 0}, dz_[3] = {0, 0, 0};
            for (int ch = 0; ch < 3; ++ch) dsh[ch] = SH_C0 * dRGB[ch];
            if (D > 0) {
                for (int ch = 0; ch < 3; ++ch) {
                    dsh[3 + ch] = -SH_C1 * y * dRGB[ch];
                    dsh[6 + ch] = SH_C1 * z * dRGB[ch];
                    dsh[9 + ch] = -SH_C1 * x * dRGB[ch];
                    dx_[ch] = -SH_C1 * sh[9 + ch];
                    dy_[ch] = -SH_C1 * sh[3 + ch];
                    dz_[ch] = SH_C1 * sh[6 + ch];
                }
                if (D > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    const float bs[5] = {SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2.f * zz - xx - yy), SH_C2[3] * xz,
                                         SH_C2[4] * (xx - yy)};
                    for (int ch = 0; ch < 3; ++ch) {
                        for (int k = 0; k < 5; ++k) dsh[3 * (4 + k) + ch] = bs[k] * dRGB[ch];
                        dx_[ch] += SH_C2[0] * y * sh[12 + ch] + SH_C2[2] * 2.f * -x * sh[18 + ch] + SH_C2[3] * z * sh[21 + ch] +
                                   SH_C2[4] * 2.f * x * sh[24 + ch];
                        dy_[ch] += SH_C2[0] * x * sh[12 + ch] + SH_C2[1] * z * sh[15 + ch] + SH_C2[2] * 2.f * -y * sh[18 + ch] +
                                   SH_C2[4] * 2.f * -y * sh[24 + ch];
                        dz_[ch] += SH_C2[1] * y * sh[15 + ch] + SH_C2[2] * 2.f * 2.f * z * sh[18 + ch] + SH_C2[3] * x * sh[21 + ch];
                    }
                    if (D > 2) {
                        const float b3[7] = {SH_C3[0] * y * (3.f * xx - yy), SH_C3[1] * xy * z,
                                             SH_C3[2] * y * (4.f * zz - xx - yy),
                                             SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy),
                                             SH_C3[4] * x * (4.f * zz - xx - yy), SH_C3[5] * z * (xx - yy),
                                             SH_C3[6] * x * (xx - 3.f * yy)};
                        for (int ch = 0; ch < 3; ++ch) {
                            for (int k = 0; k < 7; ++k) dsh[3 * (9 + k) + ch] = b3[k] * dRGB[ch];
                            dx_[ch] += SH_C3[0] * sh[27 + ch] * 3.f * 2.f * xy + SH_C3[1] * sh[30 + ch] * yz +
                                       SH_C3[2] * sh[33 + ch] * -2.f * xy + SH_C3[3] * sh[36 + ch] * -3.f * 2.f * xz +
                                       SH_C3[4] * sh[39 + ch] * (-3.f * xx + 4.f * zz - yy) +
                                       SH_C3[5] * sh[42 + ch] * 2.f * xz + SH_C3[6] * sh[45 + ch] * 3.f * (xx - yy);
                            dy_[ch] += SH_C3[0] * sh[27 + ch] * 3.f * (xx - yy) + SH_C3[1] * sh[30 + ch] * xz +
                                       SH_C3[2] * sh[33 + ch] * (-3.f * yy + 4.f * zz - xx) +
                                       SH_C3[3] * sh[36 + ch] * -3.f * 2.f * yz + SH_C3[4] * sh[39 + ch] * -2.f * xy +
                                       SH_C3[5] * sh[42 + ch] * -2.f * yz + SH_C3[6] * sh[45 + ch] * -3.f * 2.f * xy;
                            dz_[ch] += SH_C3[1] * sh[30 + ch] * xy + SH_C3[2] * sh[33 + ch] * 4.f * 2.f * yz +
                                       SH_C3[3] * sh[36 + ch] * 3.f * (2.f * zz - xx - yy) +
                                       SH_C3[4] * sh[39 + ch] * 4.f * 2.f * xz + SH_C3[5] * sh[42 + ch] * (xx - yy);
                        }
                    }
                }
            }
            const float ddx = dx_[0] * dRGB[0] + dx_[1] * dRGB[1] + dx_[2] * dRGB[2];
            const float ddy = dy_[0] * dRGB[0] + dy_[1] * dRGB[1] + dy_[2] * dRGB[2];
            const float ddz = dz_[0] * dRGB[0] + dz_[1] * dRGB[1] + dz_[2] * dRGB[2];
            const float sum2 = ox * ox + oy * oy + oz * oz;
            const float inv = 1.0f / sqrtf(sum2 * sum2 * sum2);
            dm[0] += ((+sum2 - ox * ox) * ddx - oy * ox * ddy - oz * ox * ddz) * inv;
            dm[1] += (-ox * oy * ddx + (sum2 - oy * oy) * ddy - oz * oy * ddz) * inv;
            dm[2] += (-ox * oz * ddx - oy * oz * ddy + (sum2 - oz * oz) * ddz) * inv;
        }
        dL_dmean3D[3 * i] = dm[0]; dL_dmean3D[3 * i + 1] = dm[1]; dL_dmean3D[3 * i + 2] = dm[2];
        if (scales) { /* backward.cu:278-341 */
            const float r = rotations[4 * i], x = rotations[4 * i + 1], y = rotations[4 * i + 2], z = rotations[4 * i + 3];
            const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                                   {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                                   {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
            const float s[3] = {scale_modifier * scales[3 * i], scale_modifier * scales[3 * i + 1],
                                scale_modifier * scales[3 * i + 2]};
            float Mm[3][3], dM[3][3], dMt[3][3];
            for (int cc = 0; cc < 3; ++cc)
                for (int rr = 0; rr < 3; ++rr) Mm[cc][rr] = s[rr] * R[cc][rr];
            const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                                    {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                                    {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
            for (int cc = 0; cc < 3; ++cc)
                for (int rr = 0; rr < 3; ++rr)
                    dM[cc][rr] = 2.0f * Mm[0][rr] * dS[cc][0] + 2.0f * Mm[1][rr] * dS[cc][1] + 2.0f * Mm[2][rr] * dS[cc][2];
            for (int cc = 0; cc < 3; ++cc)
                for (int rr = 0; rr < 3; ++rr) dMt[cc][rr] = dM[rr][cc];
            for (int j = 0; j < 3; ++j) {
                dL_dscale[3 * i + j] = R[0][j] * dMt[j][0] + R[1][j] * dMt[j][1] + R[2][j] * dMt[j][2];
                for (int rr = 0; rr < 3; ++rr) dMt[j][rr] *= s[j];
            }
            dL_drot[4 * i] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
            dL_drot[4 * i + 1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) +
                                 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
            dL_drot[4 * i + 2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) +
                                 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
            dL_drot[4 * i + 3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) +
                                 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
        }
    }
}

/* ---- simple-knn distCUDA2 (reference submodules/simple-knn/simple_knn.cu:127-183): for every point the mean of the
 * three smallest squared distances to the OTHER points, each distance rounded like the reference build rounds
 * `d.x*d.x + d.y*d.y + d.z*d.z` (d = other - point; SASS: FMUL on d.y, FFMA with d.x, FFMA with d.z -- nvcc contracts
 * a*b + c*d to fma(a, b, c*d)), summed smallest first, divided by 3.
 * Brute force O(P^2): the reference's Morton/box search is exact, so its result does not depend on the search. */
void oracle_knn_mean_dist2(int P, const float* pts, float* out) {
    for (int i = 0; i < P; ++i) {
        float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
        const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        for (int j = 0; j < P; ++j) {
            if (j == i) continue;
            const float dx = pts[3 * j] - x, dy = pts[3 * j + 1] - y, dz = pts[3 * j + 2] - z;
            float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
            for (int k = 0; k < 3; ++k)
                if (best[k] > d) { const float t = best[k]; best[k] = d; d = t; }
        }
        out[i] = ((best[0] + best[1]) + best[2]) / 3.0f;
    }
}

int oracle_version(void) { return 1; }
