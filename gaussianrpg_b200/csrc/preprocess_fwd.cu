// preprocess_fwd.cu -- per-Gaussian forward stage: near cull, EWA projection, conic, screen
// radius, tile rectangle, SH -> RGB, plus the conservative alpha>=1/255 footprint used by the
// blend kernels for warp-level culling.
//
// Behavioural spec: reference preprocessCUDA (cuda_rasterizer/forward.cu:155-256) with
// in_frustum (auxiliary.h:139-164), computeCov3D (forward.cu:118-152), computeCov2D
// (forward.cu:74-113), computeColorFromSH (forward.cu:20-71), ndc2Pix/getRect
// (auxiliary.h:41-56).  Rounding sequence: see DESIGN.md "arithmetic contract".
#include <cstdio>
#include <cstdlib>
#include "grpg_common.cuh"

namespace grpg {

struct ProjOut {
    float depth;      // view-space z
    float px, py;     // pixel centre
    float a, b, c;    // cov2D incl. +0.3 blur
    float conx, cony, conz;
    int radius;
    uint32_t xmin, xmax, ymin, ymax;
    float cov3d[6];
    bool visible;
};

__device__ __forceinline__ void cov3d_from_scale_rot(const float sx_in, const float sy_in, const float sz_in, float mod,
                                                     float r, float x, float y, float z, float* cov3D) {
    // forward.cu:118-152 (quaternion used un-normalised, :127)
    const float sx = fmul(sx_in, mod), sy = fmul(sy_in, mod), sz = fmul(sz_in, mod);
    const float xz = fmul(x, z), rx = fmul(r, x), rz = fmul(r, z), yy = fmul(y, y), zz = fmul(z, z);
    const float xz_p_ry = ffma(r, y, xz), xz_m_ry = ffma(-r, y, xz);
    const float yz_m_rx = ffma(y, z, -rx), yz_p_rx = ffma(y, z, rx);
    const float xy_m_rz = ffma(x, y, -rz), xy_p_rz = ffma(x, y, rz);
    const float xx_yy = ffma(x, x, yy), yy_zz = fadd(yy, zz), xx_zz = ffma(x, x, zz);
    // R[c][r] (column c, row r) of the glm matrix built at forward.cu:134-138
    const float R00 = fadd(-fadd(yy_zz, yy_zz), 1.0f), R01 = fadd(xy_m_rz, xy_m_rz), R02 = fadd(xz_p_ry, xz_p_ry);
    const float R10 = fadd(xy_p_rz, xy_p_rz), R11 = fadd(-fadd(xx_zz, xx_zz), 1.0f), R12 = fadd(yz_m_rx, yz_m_rx);
    const float R20 = fadd(xz_m_ry, xz_m_ry), R21 = fadd(yz_p_rx, yz_p_rx), R22 = fadd(-fadd(xx_yy, xx_yy), 1.0f);
    // M = S * R  ->  M[c][r] = s_r * R[c][r]
    const float M00 = fmul(sx, R00), M01 = fmul(sy, R01), M02 = fmul(sz, R02);
    const float M10 = fmul(sx, R10), M11 = fmul(sy, R11), M12 = fmul(sz, R12);
    const float M20 = fmul(sx, R20), M21 = fmul(sy, R21), M22 = fmul(sz, R22);
    // Sigma = M^T M : Sigma[c][r] = sum_k M[r][k] * M[c][k]
    cov3D[0] = dot3(M00, M00, M01, M01, M02, M02);
    cov3D[1] = dot3(M10, M00, M11, M01, M12, M02);
    cov3D[2] = dot3(M20, M00, M21, M01, M22, M02);
    cov3D[3] = dot3(M10, M10, M11, M11, M12, M12);
    cov3D[4] = dot3(M20, M10, M21, M11, M22, M12);
    cov3D[5] = dot3(M20, M20, M21, M21, M22, M22);
}

// Quaternion of Gaussian idx: one 16-byte load when the array is 16-byte aligned, four scalar loads for a contiguous
// view whose storage offset is not a multiple of four floats (the reference has no alignment requirement).
__device__ __forceinline__ float4 load_quat(const float* __restrict__ rotations, size_t idx) {
    if ((reinterpret_cast<uintptr_t>(rotations) & 15u) == 0) return reinterpret_cast<const float4*>(rotations)[idx];
    const float* q = rotations + 4 * idx;
    return make_float4(q[0], q[1], q[2], q[3]);
}

// Everything that determines visibility, keys and the 2D footprint of one Gaussian.
__device__ __forceinline__ void project_gaussian(float x, float y, float z, const float* __restrict__ v,
                                                 const float* __restrict__ m, const float* cov3D, int W, int H,
                                                 float tan_fovx, float tan_fovy, float focal_x, float focal_y,
                                                 uint32_t grid_x, uint32_t grid_y, ProjOut& o) {
    o.visible = false;
    o.radius = 0;
    const float tz = xform_row(v, 2, x, y, z);
    o.depth = tz;
    if (tz <= 0.2f) return;  // auxiliary.h:154 (NaN passes, as in the reference)

    const float hx = xform_row(m, 0, x, y, z);
    const float hy = xform_row(m, 1, x, y, z);
    const float hw = xform_row(m, 3, x, y, z);
    const float p_w = frcp(fadd(hw, 0.0000001f));
    const float ndc_x = fmul(hx, p_w), ndc_y = fmul(hy, p_w);

    // computeCov2D, forward.cu:74-113
    const float tx = xform_row(v, 0, x, y, z);
    const float ty = xform_row(v, 1, x, y, z);
    const float limx = fmul(tan_fovx, 1.3f), limy = fmul(tan_fovy, 1.3f);
    const float cx = fminf(fmaxf(fdiv(tx, tz), -limx), limx);
    const float cy = fminf(fmaxf(fdiv(ty, tz), -limy), limy);
    const float tz2 = fmul(tz, tz);
    const float J00 = fdiv(focal_x, tz);
    const float J02 = fdiv(fmul(fmul(tz, -cx), focal_x), tz2);
    const float J11 = fdiv(focal_y, tz);
    const float J12 = fdiv(fmul(fmul(tz, -cy), focal_y), tz2);
    // T = W * J, rows r of W^T: W[0][r]=v[4r], W[1][r]=v[4r+1], W[2][r]=v[4r+2]
    float T0[3], T1[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        T0[r] = ffma(v[4 * r + 2], J02, fmul(v[4 * r], J00));
        T1[r] = ffma(v[4 * r + 2], J12, fmul(v[4 * r + 1], J11));
    }
    // X = T^T * Vrk^T ; cov = X * T
    const float B0[3] = {cov3D[0], cov3D[1], cov3D[2]};
    const float B1[3] = {cov3D[1], cov3D[3], cov3D[4]};
    const float B2[3] = {cov3D[2], cov3D[4], cov3D[5]};
    const float X00 = dot3(T0[0], B0[0], T0[1], B0[1], T0[2], B0[2]);
    const float X10 = dot3(T0[0], B1[0], T0[1], B1[1], T0[2], B1[2]);
    const float X20 = dot3(T0[0], B2[0], T0[1], B2[1], T0[2], B2[2]);
    const float X01 = dot3(T1[0], B0[0], T1[1], B0[1], T1[2], B0[2]);
    const float X11 = dot3(T1[0], B1[0], T1[1], B1[1], T1[2], B1[2]);
    const float X21 = dot3(T1[0], B2[0], T1[1], B2[1], T1[2], B2[2]);
    const float a = fadd(dot3(X00, T0[0], X10, T0[1], X20, T0[2]), 0.3f);
    const float b = dot3(X01, T0[0], X11, T0[1], X21, T0[2]);
    const float c = fadd(dot3(X01, T1[0], X11, T1[1], X21, T1[2]), 0.3f);
    o.a = a; o.b = b; o.c = c;

    const float det = ffma(a, c, -fmul(b, b));
    if (det == 0.0f) return;  // forward.cu:219-221
    const float det_inv = frcp(det);
    o.conx = fmul(c, det_inv);
    o.cony = fmul(b, -det_inv);
    o.conz = fmul(a, det_inv);

    const float mid = fmul(fadd(a, c), 0.5f);
    const float s = fsqrt(fmaxf(ffma(mid, mid, -det), 0.1f));
    const float lam = fmaxf(fadd(mid, s), fadd(mid, -s));
    const int radius = __float2int_ru(fmul(fsqrt(lam), 3.0f));

    // ndc2Pix is evaluated in double in the reference (auxiliary.h:41-44)
    const float px = (float)__dmul_rn(__fma_rn(__dadd_rn((double)ndc_x, 1.0), (double)W, -1.0), 0.5);
    const float py = (float)__dmul_rn(__fma_rn(__dadd_rn((double)ndc_y, 1.0), (double)H, -1.0), 0.5);
    o.px = px; o.py = py;

    // getRect, auxiliary.h:46-56
    const float rf = (float)radius;
    const int ix0 = __float2int_rz(fmul(fadd(px, -rf), 0.0625f));
    const int iy0 = __float2int_rz(fmul(fadd(py, -rf), 0.0625f));
    const int ix1 = __float2int_rz(fmul(fadd(fadd(fadd(px, rf), 16.0f), -1.0f), 0.0625f));
    const int iy1 = __float2int_rz(fmul(fadd(fadd(fadd(py, rf), 16.0f), -1.0f), 0.0625f));
    o.xmin = min(grid_x, (uint32_t)max(0, ix0));
    o.ymin = min(grid_y, (uint32_t)max(0, iy0));
    o.xmax = min(grid_x, (uint32_t)max(0, ix1));
    o.ymax = min(grid_y, (uint32_t)max(0, iy1));
    if ((o.xmax - o.xmin) * (o.ymax - o.ymin) == 0) return;  // forward.cu:236
    o.radius = radius;
    o.visible = true;
}

// SH -> RGB, forward.cu:20-71.  Returns the clamp mask in bits 0..2.
__device__ __forceinline__ uint32_t sh_to_rgb(int deg, const float* __restrict__ sh /* [M][3] of this Gaussian */,
                                              float px, float py, float pz, const float* __restrict__ campos,
                                              float* rgb) {
    const float dx = fadd(-campos[0], px), dy = fadd(-campos[1], py), dz = fadd(-campos[2], pz);
    const float len = fsqrt(ffma(dz, dz, ffma(dx, dx, fmul(dy, dy))));
    const float x = fdiv(dx, len), y = fdiv(dy, len), z = fdiv(dz, len);
    float res[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) res[ch] = fmul(sh[ch], 0.28209479177387814f);
    if (deg > 0) {
        const float c1y = fmul(y, 0.4886025119029199f), c1z = fmul(z, 0.4886025119029199f),
                    c1x = fmul(x, 0.4886025119029199f);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float r = ffma(-c1y, sh[3 + ch], res[ch]);
            r = ffma(c1z, sh[6 + ch], r);
            res[ch] = ffma(-c1x, sh[9 + ch], r);
        }
        if (deg > 1) {
            const float xx = fmul(x, x), yy = fmul(y, y), zz = fmul(z, z);
            const float xy = fmul(y, x), yz = fmul(z, y), xz = fmul(z, x);
            const float k4 = fmul(xy, 1.0925484305920792f);
            const float k5 = fmul(yz, -1.0925484305920792f);
            const float zz2 = fadd(zz, zz);
            const float k6 = fmul(fadd(-yy, fadd(-xx, zz2)), 0.31539156525252005f);
            const float k7 = fmul(xz, -1.0925484305920792f);
            const float xx_yy = fadd(xx, -yy);
            const float k8 = fmul(xx_yy, 0.5462742152960396f);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float r = ffma(k4, sh[12 + ch], res[ch]);
                r = ffma(k5, sh[15 + ch], r);
                r = ffma(k6, sh[18 + ch], r);
                r = ffma(k7, sh[21 + ch], r);
                res[ch] = ffma(k8, sh[24 + ch], r);
            }
            if (deg > 2) {
                const float k9 = fmul(fmul(y, -0.5900435899266435f), ffma(xx, 3.0f, -yy));
                const float k10 = fmul(fmul(xy, 2.890611442640554f), z);
                const float f4 = fadd(-yy, ffma(zz, 4.0f, -xx));  // 4zz - xx - yy
                const float k11 = fmul(fmul(y, -0.4570457994644658f), f4);
                const float k12 = fmul(fmul(z, 0.3731763325901154f), ffma(yy, -3.0f, ffma(xx, -3.0f, zz2)));
                const float k13 = fmul(f4, fmul(x, -0.4570457994644658f));
                const float k14 = fmul(xx_yy, fmul(z, 1.445305721320277f));
                const float k15 = fmul(fmul(x, -0.5900435899266435f), ffma(yy, -3.0f, xx));
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    float r = ffma(k9, sh[27 + ch], res[ch]);
                    r = ffma(k10, sh[30 + ch], r);
                    r = ffma(k11, sh[33 + ch], r);
                    r = ffma(k12, sh[36 + ch], r);
                    r = ffma(k13, sh[39 + ch], r);
                    r = ffma(k14, sh[42 + ch], r);
                    res[ch] = ffma(k15, sh[45 + ch], r);
                }
            }
        }
    }
    uint32_t clamp_mask = 0;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float r = fadd(res[ch], 0.5f);
        if (r < 0.0f) clamp_mask |= 1u << ch;
        rgb[ch] = r < 0.0f ? 0.0f : r;
    }
    return clamp_mask;
}

// Conservative half extents of {pixels that can reach alpha >= 1/255} for the *stored*
// conic/opacity (the quantities the blend evaluates).  Used only to skip work whose result
// the reference would discard (forward.cu:418-426), never to change a result.
__device__ __forceinline__ void alpha_footprint(float A, float B, float C, float opacity, int radius, float& hx,
                                                float& hy, float& thr) {
    const float inf = __int_as_float(0x7f800000);
    hx = inf; hy = inf; thr = -inf;
    if (!(opacity >= 0.0f) || !isfinite(A) || !isfinite(B) || !isfinite(C)) return;  // keep exact behaviour for odd inputs
    if (opacity < 1.0f / 255.0f) { hx = -inf; hy = -inf; thr = inf; return; }  // alpha = o*G <= o < 1/255 always
    const double dA = A, dB = B, dC = C;
    // Worst-case error of the float evaluation of `power` (three products, two sums) anywhere in a tile the
    // Gaussian is binned to (|d| <= radius + 16): the blend may only skip a pixel when even that error cannot lift
    // alpha over 1/255.
    const double dmax = (double)radius + 17.0;
    const double slack_q = 8.0 * 1.1920929e-7 * (fabs(dA) + fabs(dC) + 2.0 * fabs(dB)) * dmax * dmax + 1e-4;
    // `power < thr` must imply min(0.99, o*expf(power)) < 1/255 (expf is accurate to 2 ulp); thr also serves as
    // the geometric bound q = A dx^2 + 2B dx dy + C dy^2 > -2 thr of the phase-1 ellipse test, hence the slack
    const double lthr = log(1.0 / (255.0 * (double)opacity)) - 1e-4 - 0.5 * slack_q;
    thr = __double2float_rd(lthr);
    const double D = dA * dC - dB * dB;
    if (!(D > 0.0) || !(dA > 0.0) || !(dC > 0.0)) return;  // not an ellipse: no bounding box
    // alpha >= 1/255  <=>  q <= 2 ln(255 o) =: tau
    double tau = (-2.0 * lthr) * (1.0 + 1e-5);
    if (tau <= 0.0) { hx = -inf; hy = -inf; return; }
    const double ex = sqrt(tau * dC / D) + 1e-2;
    const double ey = sqrt(tau * dA / D) + 1e-2;
    hx = __double2float_ru(ex);
    hy = __double2float_ru(ey);
}

template <bool TILE_MASK>  // per-tile culling of small rectangles (binning mode 0); false compiles it out
__global__ void __launch_bounds__(256) preprocess_fwd_kernel(
    int P, int D, int M, const float* __restrict__ means3D, const float* __restrict__ scales, float scale_modifier,
    const float* __restrict__ rotations, const float* __restrict__ opacities, const float* __restrict__ shs,
    const float* __restrict__ cov3D_precomp, const float* __restrict__ colors_precomp,
    const float* __restrict__ viewmatrix, const float* __restrict__ projmatrix, const float* __restrict__ cam_pos,
    int W, int H, float tan_fovx, float tan_fovy, float focal_x, float focal_y, uint32_t grid_x, uint32_t grid_y,
    int prefiltered, int row_stride, int row_phase, int forward_only, int* __restrict__ radii, Rec* __restrict__ rec, uint32_t* __restrict__ depth_key,
    uint2* __restrict__ rect, uint32_t* __restrict__ tiles_touched, float* __restrict__ cov3D_out,
    uint8_t* __restrict__ clamped, int reference_binning, unsigned long long* __restrict__ ref_instances,
    uint32_t* __restrict__ tile_mask) {
    __shared__ float s_cam[35];
    __shared__ uint32_t s_ref_count;  // instances the reference bins for this CTA's Gaussians (its num_rendered)
    if (threadIdx.x == 0) s_ref_count = 0;
    // Per-CTA slices of the AoS attribute arrays are contiguous in global memory (256 x 12 B means, 256 x 12 B
    // scales, 256 x 48 B SH at M = 4), so they are staged with TMA bulk copies (one elected thread, one mbarrier)
    // and then read conflict-free from shared memory -- instead of 12-byte-strided per-thread loads.
    __shared__ __align__(16) float s_means[256 * 3];
    __shared__ __align__(16) float s_scales[256 * 3];
    __shared__ __align__(16) float s_sh[256 * 12];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmask[256];
    const int cta_first = blockIdx.x * blockDim.x;
    auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    // the ragged last CTA and 16-byte-misaligned views use plain loads
    const bool full_cta = cta_first + 256 <= P && aligned16(means3D);
    const bool stage_scales = full_cta && scales != nullptr && cov3D_precomp == nullptr && aligned16(scales);
    const bool stage_sh = full_cta && shs != nullptr && colors_precomp == nullptr && M == 4 && aligned16(shs);
    if (full_cta) {
        if (threadIdx.x == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t bytes = 256 * 12;
            if (stage_scales) bytes += 256 * 12;
            if (stage_sh) bytes += 256 * 48;
            mbar_expect_tx(&s_bar, bytes);
            tma_bulk_g2s(s_means, means3D + 3 * (size_t)cta_first, 256 * 12, &s_bar);
            if (stage_scales) tma_bulk_g2s(s_scales, scales + 3 * (size_t)cta_first, 256 * 12, &s_bar);
            if (stage_sh) tma_bulk_g2s(s_sh, shs + 12 * (size_t)cta_first, 256 * 48, &s_bar);
        }
    }
    if (threadIdx.x < 16) s_cam[threadIdx.x] = viewmatrix[threadIdx.x];
    else if (threadIdx.x < 32) s_cam[threadIdx.x] = projmatrix[threadIdx.x - 16];
    else if (threadIdx.x < 35) s_cam[threadIdx.x] = cam_pos[threadIdx.x - 32];
    __syncthreads();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (full_cta) mbar_wait(&s_bar, 0);
    if (idx >= P) return;
    const float* v = s_cam;
    const float* m = s_cam + 16;

    float x, y, z;
    if (full_cta) { x = s_means[3 * threadIdx.x]; y = s_means[3 * threadIdx.x + 1]; z = s_means[3 * threadIdx.x + 2]; }
    else { x = means3D[3 * idx]; y = means3D[3 * idx + 1]; z = means3D[3 * idx + 2]; }
    float cov3D[6];
    ProjOut o;
    // cheap early cull before touching the other attribute arrays
    const float tz = xform_row(v, 2, x, y, z);
    bool vis = !(tz <= 0.2f);
    if (!vis && prefiltered) {
        printf("Point is filtered although prefiltered is set. This shouldn't happen!");
        __trap();
    }
    if (vis) {
        if (cov3D_precomp != nullptr) {
#pragma unroll
            for (int k = 0; k < 6; ++k) cov3D[k] = cov3D_precomp[6 * (size_t)idx + k];
        } else {
            const float4 q = load_quat(rotations, (size_t)idx);
            const float* sc = stage_scales ? s_scales + 3 * threadIdx.x : scales + 3 * (size_t)idx;
            cov3d_from_scale_rot(sc[0], sc[1], sc[2], scale_modifier, q.x, q.y, q.z, q.w, cov3D);
            if (!forward_only) {
#pragma unroll
                for (int k = 0; k < 6; ++k) cov3D_out[6 * (size_t)idx + k] = cov3D[k];
            }
        }
        project_gaussian(x, y, z, v, m, cov3D, W, H, tan_fovx, tan_fovy, focal_x, focal_y, grid_x, grid_y, o);
        vis = o.visible;
    }
    Rec r;
    uint32_t ref_count = 0;
    uint32_t n_tiles = 0, mask = 0xFFFFFFFFu, cull_cnt = 0, cull_w = 1, cull_x0 = 0, cull_y0 = 0;
    if (!vis) {
        radii[idx] = 0;
        tiles_touched[idx] = 0;
        depth_key[idx] = 0xFFFFFFFFu;
        rect[idx] = make_uint2(0u, 0u);
        r.a = r.b = r.c = make_float4(0.f, 0.f, 0.f, 0.f);  // never read: culled Gaussians emit no instance
    } else {
        float rgb[3];
        uint32_t cmask = 0;
        if (colors_precomp != nullptr) {
            rgb[0] = colors_precomp[3 * idx]; rgb[1] = colors_precomp[3 * idx + 1]; rgb[2] = colors_precomp[3 * idx + 2];
        } else {
            cmask = sh_to_rgb(D, stage_sh ? s_sh + 12 * threadIdx.x : shs + (size_t)idx * M * 3, x, y, z, s_cam + 32, rgb);
            if (!forward_only) clamped[idx] = (uint8_t)cmask;
        }
        const float opacity = opacities[idx];
        // rows of [y0, y1) owned by this band (row % stride == phase), expressed as local row indices
        auto band_clip = [&](uint32_t y0, uint32_t y1, uint32_t& l0, uint32_t& l1) {
            l0 = y0; l1 = y1;
            if (row_stride > 1) {
                const uint32_t st = (uint32_t)row_stride, ph = (uint32_t)row_phase;
                const uint32_t first = y0 + ((ph + st - y0 % st) % st);
                if (first < y1) { l0 = (first - ph) / st; l1 = l0 + (y1 - 1 - first) / st + 1; }
                else { l0 = 0; l1 = 0; }
            }
        };
        uint32_t ly0, ly1;
        band_clip(o.ymin, o.ymax, ly0, ly1);
        ref_count = (o.xmax - o.xmin) * (ly1 - ly0);  // what the reference bins: getRect of the 3-sigma radius
        // A Gaussian whose rectangle has no row in this rank's band emits no instance here, so its record is never
        // staged: the alpha footprint (double-precision log and square roots, the most expensive part of this kernel)
        // is skipped for it -- four in five visible Gaussians on a rank of an 8-way sharded frame.
        float hx = __int_as_float(0x7f800000), hy = hx, thr = -hx;
        if (ly1 > ly0) alpha_footprint(o.conx, o.cony, o.conz, opacity, o.radius, hx, hy, thr);
        r.a = make_float4(o.px, o.py, pack_extents(hx, hy), thr);
        r.b = make_float4(o.conx, o.cony, o.conz, opacity);
        r.c = make_float4(rgb[0], rgb[1], rgb[2], o.depth);
        radii[idx] = o.radius;
        uint32_t bx0 = o.xmin, bx1 = o.xmax;
        if (reference_binning != 1 && !(hx == __int_as_float(0x7f800000))) {
            // Clip the rectangle to the tiles that hold a pixel with |dx| <= hx and |dy| <= hy: outside, alpha is
            // provably < 1/255 and the reference's blend skips the pair (forward.cu:418-426), so the instance is
            // never binned.  Images, gradients and radii are unchanged; the instance lists become the reference's
            // lists minus those dead entries.
            uint32_t by0 = o.ymin, by1 = o.ymax;
            if (hx < 0.0f || hy < 0.0f) {
                bx1 = bx0; by1 = by0;
            } else {
                const double xl = floor(((double)o.px - (double)hx) * 0.0625), xh = floor(((double)o.px + (double)hx) * 0.0625) + 1.0;
                const double yl = floor(((double)o.py - (double)hy) * 0.0625), yh = floor(((double)o.py + (double)hy) * 0.0625) + 1.0;
                bx0 = (uint32_t)fmin(fmax(xl, (double)o.xmin), (double)o.xmax);
                bx1 = (uint32_t)fmax(fmin(xh, (double)o.xmax), (double)bx0);
                by0 = (uint32_t)fmin(fmax(yl, (double)o.ymin), (double)o.ymax);
                by1 = (uint32_t)fmax(fmin(yh, (double)o.ymax), (double)by0);
            }
            band_clip(by0, by1, ly0, ly1);
        }
        n_tiles = (bx1 - bx0) * (ly1 - ly0);
        cull_w = bx1 - bx0; cull_x0 = bx0; cull_y0 = ly0;
        if (TILE_MASK && reference_binning == 0 && n_tiles > 0u && n_tiles <= 32u) {
            // Exact per-tile culling for small rectangles: a tile is binned only if it can hold a pixel with
            // alpha >= 1/255 -- the same conservative ellipse-vs-rectangle test the blend kernels run per 8x8 block
            // (a block is a subset of its tile, so a tile that fails is failed by all of its blocks: the blend would
            // discard the instance).  The surviving tiles are a bit mask over the rectangle (row-major).
            if (full_cta) {
                cull_cnt = n_tiles;  // tested below, one (Gaussian, tile) pair per lane
            } else {
                const uint32_t st = row_stride > 1 ? (uint32_t)row_stride : 1u, ph = row_stride > 1 ? (uint32_t)row_phase : 0u;
                mask = 0u;
                for (uint32_t t = 0; t < n_tiles; ++t) {
                    const uint32_t ty_l = ly0 + t / cull_w, tx = bx0 + t % cull_w;
                    const float x_lo = (float)(tx * GRPG_TILE), y_lo = (float)((ty_l * st + ph) * GRPG_TILE);
                    if (footprint_hits_exact(r.a, r.b, x_lo, x_lo + (GRPG_TILE - 1), y_lo, y_lo + (GRPG_TILE - 1))) mask |= 1u << t;
                }
                if (mask != 0xFFFFFFFFu) n_tiles = (uint32_t)__popc(mask);
                // (a 32-tile rectangle whose every tile survives keeps the all-ones mask, which also means "no mask")
            }
        }
        depth_key[idx] = __float_as_uint(o.depth);
        rect[idx] = make_uint2(bx0 | (bx1 << 16), ly0 | (ly1 << 16));
    }
    if (full_cta) {  // records go to the slots the bulk store below reads (and the per-tile test reads back)
        float4* slot = reinterpret_cast<float4*>(s_sh) + 3 * threadIdx.x;
        slot[0] = r.a; slot[1] = r.b; slot[2] = r.c;
    }
    if (TILE_MASK && full_cta) {
        // Warp-cooperative form of the per-tile test (all 32 lanes of a full CTA are here): the (Gaussian, tile) pairs
        // of the warp's small rectangles are numbered by a prefix sum and tested 32 at a time, one pair per lane,
        // instead of every lane looping over its own rectangle (measured: the per-lane loop doubled this kernel,
        // 0.104 -> 0.212 ms on the 2 M scene).  Records are read back from the slots they are stored from anyway.
        const int lane = threadIdx.x & 31;
        uint32_t incl = cull_cnt;
#pragma unroll
        for (int o_ = 1; o_ < 32; o_ <<= 1) {
            const uint32_t t_ = __shfl_up_sync(0xffffffffu, incl, o_);
            if (lane >= o_) incl += t_;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total) {
            uint32_t* wmask = s_tmask + (threadIdx.x & ~31u);
            wmask[lane] = 0u;
            __syncwarp();
            const uint32_t excl = incl - cull_cnt;
            const uint32_t st = row_stride > 1 ? (uint32_t)row_stride : 1u, ph = row_stride > 1 ? (uint32_t)row_phase : 0u;
            for (uint32_t base = 0; base < total; base += 32) {
                const uint32_t pr = base + lane;
                int owner = 0;  // number of lanes whose inclusive count is <= pr = the lane that owns pair pr
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1) {
                    const uint32_t probe = __shfl_sync(0xffffffffu, incl, (owner + step - 1) & 31);
                    if (probe <= pr) owner += step;
                }
                owner &= 31;
                const uint32_t t = pr - __shfl_sync(0xffffffffu, excl, owner);
                const uint32_t ow = __shfl_sync(0xffffffffu, cull_w, owner);
                const uint32_t ox0 = __shfl_sync(0xffffffffu, cull_x0, owner);
                const uint32_t oy0 = __shfl_sync(0xffffffffu, cull_y0, owner);
                if (pr < total) {
                    // t < 32 and ow <= 32: (t + 0.5) / ow is at least 1/64 away from every integer, float is exact enough
                    const uint32_t ty = (uint32_t)(((float)t + 0.5f) * __frcp_rn((float)ow));
                    const uint32_t tx = t - ty * ow;
                    const float4* os = reinterpret_cast<const float4*>(s_sh) + 3 * ((threadIdx.x & ~31u) + owner);
                    const float x_lo = (float)((ox0 + tx) * GRPG_TILE), y_lo = (float)(((oy0 + ty) * st + ph) * GRPG_TILE);
                    if (footprint_hits_exact(os[0], os[1], x_lo, x_lo + (GRPG_TILE - 1), y_lo, y_lo + (GRPG_TILE - 1)))
                        atomicOr(&wmask[owner], 1u << t);
                }
            }
            __syncwarp();
            if (cull_cnt) {
                mask = wmask[lane];
                // (a 32-tile rectangle whose every tile survives keeps the all-ones mask, which also means "no mask")
                if (mask != 0xFFFFFFFFu) n_tiles = (uint32_t)__popc(mask);
            }
        }
    }
    if (vis) {
        if (TILE_MASK) tile_mask[idx] = mask;
        tiles_touched[idx] = n_tiles;
    }
    // The 256 records of a full CTA are one contiguous 12 KB range: each thread drops its record into its own
    // (already consumed) 48-byte SH slot and one thread writes the range back with a TMA bulk store.
    if (full_cta) {
        ref_count = __reduce_add_sync(0xffffffffu, ref_count);
        if ((threadIdx.x & 31) == 0 && ref_count) atomicAdd(&s_ref_count, ref_count);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            if (s_ref_count) atomicAdd(ref_instances, (unsigned long long)s_ref_count);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(rec + cta_first),
                         "r"(smem_u32(s_sh)), "r"(256u * 48u)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else {
        if (vis) rec[idx] = r;
        // ragged / misaligned CTAs (threads past P have exited): warp-aggregated instead of CTA-aggregated
        const unsigned active = __activemask();
        ref_count = __reduce_add_sync(active, ref_count);
        if ((threadIdx.x & 31) == (unsigned)(__ffs(active) - 1) && ref_count)
            atomicAdd(ref_instances, (unsigned long long)ref_count);
    }
}

// checkFrustum, rasterizer_impl.cu:54-66
__global__ void __launch_bounds__(256) mark_visible_kernel(int P, const float* __restrict__ means3D,
                                                           const float* __restrict__ viewmatrix,
                                                           uint8_t* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float tz = xform_row(viewmatrix, 2, means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
    present[idx] = !(tz <= 0.2f);
}

// filter_preprocessCUDA, forward.cu:259-334
__global__ void __launch_bounds__(256) visible_filter_kernel(
    int P, const float* __restrict__ means3D, const float* __restrict__ scales, float scale_modifier,
    const float* __restrict__ rotations, const float* __restrict__ cov3D_precomp,
    const float* __restrict__ viewmatrix, const float* __restrict__ projmatrix, int W, int H, float tan_fovx,
    float tan_fovy, float focal_x, float focal_y, uint32_t grid_x, uint32_t grid_y, int* __restrict__ radii,
    float* __restrict__ means2D) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float x = means3D[3 * idx], y = means3D[3 * idx + 1], z = means3D[3 * idx + 2];
    radii[idx] = 0;
    const float tz = xform_row(viewmatrix, 2, x, y, z);
    if (tz <= 0.2f) return;
    float cov3D[6];
    if (cov3D_precomp != nullptr) {
#pragma unroll
        for (int k = 0; k < 6; ++k) cov3D[k] = cov3D_precomp[6 * (size_t)idx + k];
    } else {
        const float4 q = load_quat(rotations, (size_t)idx);
        cov3d_from_scale_rot(scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2], scale_modifier, q.x, q.y, q.z,
                             q.w, cov3D);
    }
    ProjOut o;
    project_gaussian(x, y, z, viewmatrix, projmatrix, cov3D, W, H, tan_fovx, tan_fovy, focal_x, focal_y, grid_x,
                     grid_y, o);
    if (!o.visible) return;
    radii[idx] = o.radius;
    means2D[2 * idx] = o.px;
    means2D[2 * idx + 1] = o.py;
}

// GRPG_EXACT_TILE_CULL=1: per-tile masks for small rectangles (see launch_preprocess_fwd); read once per process
bool exact_tile_cull() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GRPG_EXACT_TILE_CULL");
        v = (e && atoi(e) != 0) ? 1 : 0;
    }
    return v != 0;
}

void launch_preprocess_fwd(const grpg_forward_args* a, float focal_x, float focal_y, uint32_t grid_x, uint32_t grid_y,
                           Rec* rec, uint32_t* depth_key, uint2* rect, uint32_t* tiles_touched, float* cov3d,
                           uint8_t* clamped, uint32_t* tile_mask, unsigned long long* ref_instances, cudaStream_t stream) {
    const int P = a->P;
    // binning mode of the kernel: 1 = the reference's rectangles, 2 = rectangles clipped to the alpha footprint (default),
    // 0 = clipped rectangles + per-tile mask (GRPG_EXACT_TILE_CULL=1).  Measured on the 2 M scene:
    // the mask removes 13 % of the binned instances and 0.037 ms from the sort and the blends, and costs 0.045 ms here
    // plus 0.017 ms in the emission -- a net loss, so it stays an A/B switch.
    const int binning_mode = a->reference_binning ? 1 : (exact_tile_cull() ? 0 : 2);
    ProfScope ps("preprocess_fwd", stream);
#define GRPG_PRE_ARGS                                                                                                     \
        P, a->D, a->M, a->means3D, a->scales, a->scale_modifier, a->rotations, a->opacities, a->shs, a->cov3D_precomp,   \
        a->colors_precomp, a->viewmatrix, a->projmatrix, a->cam_pos, a->width, a->height, a->tan_fovx, a->tan_fovy,      \
        focal_x, focal_y, grid_x, grid_y, a->prefiltered, a->tile_row_stride, a->tile_row_phase, a->forward_only,        \
        a->radii, rec, depth_key, rect, tiles_touched, cov3d, clamped, binning_mode, ref_instances, tile_mask
    if (binning_mode == 0) preprocess_fwd_kernel<true><<<(P + 255) / 256, 256, 0, stream>>>(GRPG_PRE_ARGS);
    else preprocess_fwd_kernel<false><<<(P + 255) / 256, 256, 0, stream>>>(GRPG_PRE_ARGS);
#undef GRPG_PRE_ARGS
}

void launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t stream) {
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, viewmatrix, present);
}

void launch_visible_filter(int P, int W, int H, const float* means3D, const float* scales, float scale_modifier,
                           const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                           const float* projmatrix, float tan_fovx, float tan_fovy, int* radii, float* means2D,
                           cudaStream_t stream) {
    const float focal_y = H / (2.0f * tan_fovy), focal_x = W / (2.0f * tan_fovx);
    const uint32_t gx = (W + 15) / 16, gy = (H + 15) / 16;
    visible_filter_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, scales, scale_modifier, rotations,
                                                               cov3D_precomp, viewmatrix, projmatrix, W, H, tan_fovx,
                                                               tan_fovy, focal_x, focal_y, gx, gy, radii, means2D);
}

}  // namespace grpg
