// preprocess_bwd.cu -- per-Gaussian backward: 2D gradients -> 3D parameter gradients.
//
// Fuses the reference's two kernels computeCov2DCUDA (cuda_rasterizer/backward.cu:144-274) and
// preprocessCUDA backward (backward.cu:346-412, with computeColorFromSH :20-139 and
// computeCov3D :278-341) into one pass that reads the 48-byte gradient record produced by the
// blend backward once and writes every output row exactly once (zeros for culled Gaussians),
// so the caller needs no zero-initialised gradient tensors (reference: 11 torch::zeros,
// rasterize_points.cu:166-176).
//
// Conventions kept from the reference: 1/(denom^2 + 1e-7) (:203); clamped t.x/t.y gradients are
// zeroed and t.x,t.y treated as independent of t.z (:175-176,262-264); SH gradient zeroed per
// clamped channel (:31-34); quaternion gradient is for the raw, un-normalised quaternion (:340).
#include "grpg_common.cuh"

namespace grpg {

struct f3 { float x, y, z; };

// Peer-mapped record buffers of all ranks (full size, global row index): the record all-reduce of the multi-GPU
// backward is done here, in the load path of its only consumer.
struct PeerGrad {
    const float* p[8];
    int n;
};

// Inputs are full-size arrays indexed by the global Gaussian id p_begin + i; grad_rec and every output hold
// `P` rows for the slice [p_begin, p_begin + P) (row i).
template <int MINB>
__global__ void __launch_bounds__(256, MINB) preprocess_bwd_kernel(
    int P, int p_begin, int D, int M, const float* __restrict__ means3D_full, const int* __restrict__ radii_full,
    const float* __restrict__ shs_full, const uint8_t* __restrict__ clamped_full, const float* __restrict__ scales_full,
    const float* __restrict__ rotations_full, float scale_modifier, const float* __restrict__ cov3Ds_full,
    const float* __restrict__ view, const float* __restrict__ proj, float h_x, float h_y, float tan_fovx,
    float tan_fovy, const float* __restrict__ campos, const float* __restrict__ grad_rec /*[P][12]*/,
    PeerGrad peers, float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic_out, float* __restrict__ dL_dopacity,
    float* __restrict__ dL_dcolor, float* __restrict__ dL_ddepth, float* __restrict__ dL_dmean3D,
    float* __restrict__ dL_dcov3D, float* __restrict__ dL_dsh, float* __restrict__ dL_dscale,
    float* __restrict__ dL_drot) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    // shift the input views to the slice so that everything below indexes with the local row `idx`
    const size_t g0_ = (size_t)p_begin;
    const float* means3D = means3D_full + 3 * g0_;
    const int* radii = radii_full + g0_;
    const float* shs = shs_full ? shs_full + g0_ * M * 3 : nullptr;
    const uint8_t* clamped = clamped_full + g0_;
    const float* scales = scales_full ? scales_full + 3 * g0_ : nullptr;
    const float* rotations = rotations_full ? rotations_full + 4 * g0_ : nullptr;
    const float* cov3Ds = cov3Ds_full + 6 * g0_;
    const bool vis = radii[idx] > 0;

    float4 g0 = make_float4(0, 0, 0, 0), g1 = g0, g2 = g0;
    if (vis && peers.n == 0) {
        const float4* gp = reinterpret_cast<const float4*>(grad_rec) + 3 * (size_t)idx;
        g0 = gp[0]; g1 = gp[1]; g2 = gp[2];
    } else if (vis) {
        // fixed summation order (rank 0, 1, ...) so that every rank obtains bit-identical sums
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (r >= peers.n) break;
            const float4* gp = reinterpret_cast<const float4*>(peers.p[r]) + 3 * (g0_ + (size_t)idx);
            const float4 a0 = __ldcg(gp), a1 = __ldcg(gp + 1), a2 = __ldcg(gp + 2);
            g0.x += a0.x; g0.y += a0.y; g0.z += a0.z; g0.w += a0.w;
            g1.x += a1.x; g1.y += a1.y; g1.z += a1.z; g1.w += a1.w;
            g2.x += a2.x; g2.y += a2.y; g2.z += a2.z; g2.w += a2.w;
        }
    }
    // g0 = (dmean2D.x, dmean2D.y, |.|, dconic.x) g1 = (dconic.y, dconic.w, dopacity, dcolor.r) g2 = (dcolor.g, dcolor.b, ddepth, -)
    const float dm2x = g0.x, dm2y = g0.y, dm2abs = g0.z;
    const float dcon_x = g0.w, dcon_y = g1.x, dcon_w = g1.y;
    const float dopac = g1.z;
    const float dcol[3] = {g1.w, g2.x, g2.y};
    const float ddep = g2.z;

    // outputs nobody asked for (NULL) are not written: dL_dconic / dL_ddepth are internal to the chain rule, dL_dcolor
    // only matters with precomputed colours, dL_dcov3D with precomputed covariances (56 of 168 B per Gaussian)
    if (dL_dmean2D) { dL_dmean2D[3 * idx] = dm2x; dL_dmean2D[3 * idx + 1] = dm2y; dL_dmean2D[3 * idx + 2] = dm2abs; }
    if (dL_dconic_out) reinterpret_cast<float4*>(dL_dconic_out)[idx] = make_float4(dcon_x, dcon_y, 0.f, dcon_w);
    dL_dopacity[idx] = dopac;
    if (dL_dcolor) { dL_dcolor[3 * idx] = dcol[0]; dL_dcolor[3 * idx + 1] = dcol[1]; dL_dcolor[3 * idx + 2] = dcol[2]; }
    if (dL_ddepth) dL_ddepth[idx] = ddep;

    float dmean[3] = {0.f, 0.f, 0.f};
    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dscale[3] = {0.f, 0.f, 0.f};
    float drot[4] = {0.f, 0.f, 0.f, 0.f};

    if (vis) {
        const float mx = means3D[3 * idx], my = means3D[3 * idx + 1], mz = means3D[3 * idx + 2];
        const float* cov3D = cov3Ds + 6 * (size_t)idx;

        // ---- conic -> cov2D -> cov3D / mean (backward.cu:144-274) ----
        float tx = view[0] * mx + view[4] * my + view[8] * mz + view[12];
        float ty = view[1] * mx + view[5] * my + view[9] * mz + view[13];
        const float tz = view[2] * mx + view[6] * my + view[10] * mz + view[14];
        const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
        const float txtz = tx / tz, tytz = ty / tz;
        tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
        ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
        const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;

        const float J00 = h_x / tz, J02 = -(h_x * tx) / (tz * tz);
        const float J11 = h_y / tz, J12 = -(h_y * ty) / (tz * tz);
        // T[c][r] as in the forward: T0r = W0r*J00 + W2r*J02 ; T1r = W1r*J11 + W2r*J12
        float T0[3], T1[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            T0[r] = view[4 * r] * J00 + view[4 * r + 2] * J02;
            T1[r] = view[4 * r + 1] * J11 + view[4 * r + 2] * J12;
        }
        const float V[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
        float TV0[3], TV1[3];  // (T0 . Vrk[k]), (T1 . Vrk[k])
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            TV0[k] = T0[0] * V[k][0] + T0[1] * V[k][1] + T0[2] * V[k][2];
            TV1[k] = T1[0] * V[k][0] + T1[1] * V[k][1] + T1[2] * V[k][2];
        }
        const float a = TV0[0] * T0[0] + TV0[1] * T0[1] + TV0[2] * T0[2] + 0.3f;
        const float b = TV0[0] * T1[0] + TV0[1] * T1[1] + TV0[2] * T1[2];
        const float c = TV1[0] * T1[0] + TV1[1] * T1[1] + TV1[2] * T1[2] + 0.3f;

        const float denom = a * c - b * b;
        float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        if (denom2inv != 0.f) {
            dL_da = denom2inv * (-c * c * dcon_x + 2 * b * c * dcon_y + (denom - a * c) * dcon_w);
            dL_dc = denom2inv * (-a * a * dcon_w + 2 * a * b * dcon_y + (denom - a * c) * dcon_x);
            dL_db = denom2inv * 2 * (b * c * dcon_x - (denom + 2 * b * b) * dcon_y + a * b * dcon_w);
            dcov[0] = T0[0] * T0[0] * dL_da + T0[0] * T1[0] * dL_db + T1[0] * T1[0] * dL_dc;
            dcov[3] = T0[1] * T0[1] * dL_da + T0[1] * T1[1] * dL_db + T1[1] * T1[1] * dL_dc;
            dcov[5] = T0[2] * T0[2] * dL_da + T0[2] * T1[2] * dL_db + T1[2] * T1[2] * dL_dc;
            dcov[1] = 2 * T0[0] * T0[1] * dL_da + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_db + 2 * T1[0] * T1[1] * dL_dc;
            dcov[2] = 2 * T0[0] * T0[2] * dL_da + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_db + 2 * T1[0] * T1[2] * dL_dc;
            dcov[4] = 2 * T0[2] * T0[1] * dL_da + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_db + 2 * T1[1] * T1[2] * dL_dc;
        }
        // dL/dT (upper 2x3)
        const float dT00 = 2 * TV0[0] * dL_da + TV1[0] * dL_db;
        const float dT01 = 2 * TV0[1] * dL_da + TV1[1] * dL_db;
        const float dT02 = 2 * TV0[2] * dL_da + TV1[2] * dL_db;
        const float dT10 = 2 * TV1[0] * dL_dc + TV0[0] * dL_db;
        const float dT11 = 2 * TV1[1] * dL_dc + TV0[1] * dL_db;
        const float dT12 = 2 * TV1[2] * dL_dc + TV0[2] * dL_db;
        // T = W * J : dL/dJ
        const float dJ00 = view[0] * dT00 + view[4] * dT01 + view[8] * dT02;
        const float dJ02 = view[2] * dT00 + view[6] * dT01 + view[10] * dT02;
        const float dJ11 = view[1] * dT10 + view[5] * dT11 + view[9] * dT12;
        const float dJ12 = view[2] * dT10 + view[6] * dT11 + view[10] * dT12;
        const float tzi = 1.f / tz, tz2 = tzi * tzi, tz3 = tz2 * tzi;
        const float dtx = x_grad_mul * -h_x * tz2 * dJ02;
        const float dty = y_grad_mul * -h_y * tz2 * dJ12;
        const float dtz = -h_x * tz2 * dJ00 - h_y * tz2 * dJ11 + (2 * h_x * tx) * tz3 * dJ02 + (2 * h_y * ty) * tz3 * dJ12;
        dmean[0] = view[0] * dtx + view[1] * dty + view[2] * dtz;
        dmean[1] = view[4] * dtx + view[5] * dty + view[6] * dtz;
        dmean[2] = view[8] * dtx + view[9] * dty + view[10] * dtz;

        // ---- 2D mean -> 3D mean through the projection (backward.cu:371-389) ----
        const float hw_ = proj[3] * mx + proj[7] * my + proj[11] * mz + proj[15];
        const float m_w = 1.0f / (hw_ + 0.0000001f);
        const float mul1 = (proj[0] * mx + proj[4] * my + proj[8] * mz + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * mx + proj[5] * my + proj[9] * mz + proj[13]) * m_w * m_w;
        dmean[0] += (proj[0] * m_w - proj[3] * mul1) * dm2x + (proj[1] * m_w - proj[3] * mul2) * dm2y;
        dmean[1] += (proj[4] * m_w - proj[7] * mul1) * dm2x + (proj[5] * m_w - proj[7] * mul2) * dm2y;
        dmean[2] += (proj[8] * m_w - proj[11] * mul1) * dm2x + (proj[9] * m_w - proj[11] * mul2) * dm2y;

        // ---- depth -> 3D mean (backward.cu:391-403) ----
        const float mul3 = view[2] * mx + view[6] * my + view[10] * mz + view[14];
        dmean[0] += (view[2] - view[3] * mul3) * ddep;
        dmean[1] += (view[6] - view[7] * mul3) * ddep;
        dmean[2] += (view[10] - view[11] * mul3) * ddep;

        // ---- SH (backward.cu:20-139) ----
        if (shs != nullptr) {
            const float* sh = shs + (size_t)idx * M * 3;
            float* dsh = dL_dsh + (size_t)idx * M * 3;
            const uint32_t cm = clamped[idx];
            const float dRGB[3] = {(cm & 1u) ? 0.f : dcol[0], (cm & 2u) ? 0.f : dcol[1], (cm & 4u) ? 0.f : dcol[2]};
            const float ox = mx - campos[0], oy = my - campos[1], oz = mz - campos[2];
            const float inv_len = 1.0f / sqrtf(ox * ox + oy * oy + oz * oz);
            const float x = ox * inv_len, y = oy * inv_len, z = oz * inv_len;
            float dRdx[3] = {0, 0, 0}, dRdy[3] = {0, 0, 0}, dRdz[3] = {0, 0, 0};
            const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
            const float C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                                 0.5462742152960396f};
            const float C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                                 -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};
            float basis[16];
            int nb = 1;
            basis[0] = C0;
            if (D > 0) {
                basis[1] = -C1 * y; basis[2] = C1 * z; basis[3] = -C1 * x;
                nb = 4;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    dRdx[ch] = -C1 * sh[9 + ch];
                    dRdy[ch] = -C1 * sh[3 + ch];
                    dRdz[ch] = C1 * sh[6 + ch];
                }
                if (D > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    basis[4] = C2[0] * xy; basis[5] = C2[1] * yz; basis[6] = C2[2] * (2.f * zz - xx - yy);
                    basis[7] = C2[3] * xz; basis[8] = C2[4] * (xx - yy);
                    nb = 9;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        dRdx[ch] += C2[0] * y * sh[12 + ch] + C2[2] * 2.f * -x * sh[18 + ch] + C2[3] * z * sh[21 + ch] +
                                    C2[4] * 2.f * x * sh[24 + ch];
                        dRdy[ch] += C2[0] * x * sh[12 + ch] + C2[1] * z * sh[15 + ch] + C2[2] * 2.f * -y * sh[18 + ch] +
                                    C2[4] * 2.f * -y * sh[24 + ch];
                        dRdz[ch] += C2[1] * y * sh[15 + ch] + C2[2] * 2.f * 2.f * z * sh[18 + ch] + C2[3] * x * sh[21 + ch];
                    }
                    if (D > 2) {
                        basis[9] = C3[0] * y * (3.f * xx - yy);
                        basis[10] = C3[1] * xy * z;
                        basis[11] = C3[2] * y * (4.f * zz - xx - yy);
                        basis[12] = C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                        basis[13] = C3[4] * x * (4.f * zz - xx - yy);
                        basis[14] = C3[5] * z * (xx - yy);
                        basis[15] = C3[6] * x * (xx - 3.f * yy);
                        nb = 16;
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            dRdx[ch] += C3[0] * sh[27 + ch] * 3.f * 2.f * xy + C3[1] * sh[30 + ch] * yz +
                                        C3[2] * sh[33 + ch] * -2.f * xy + C3[3] * sh[36 + ch] * -3.f * 2.f * xz +
                                        C3[4] * sh[39 + ch] * (-3.f * xx + 4.f * zz - yy) + C3[5] * sh[42 + ch] * 2.f * xz +
                                        C3[6] * sh[45 + ch] * 3.f * (xx - yy);
                            dRdy[ch] += C3[0] * sh[27 + ch] * 3.f * (xx - yy) + C3[1] * sh[30 + ch] * xz +
                                        C3[2] * sh[33 + ch] * (-3.f * yy + 4.f * zz - xx) +
                                        C3[3] * sh[36 + ch] * -3.f * 2.f * yz + C3[4] * sh[39 + ch] * -2.f * xy +
                                        C3[5] * sh[42 + ch] * -2.f * yz + C3[6] * sh[45 + ch] * -3.f * 2.f * xy;
                            dRdz[ch] += C3[1] * sh[30 + ch] * xy + C3[2] * sh[33 + ch] * 4.f * 2.f * yz +
                                        C3[3] * sh[36 + ch] * 3.f * (2.f * zz - xx - yy) +
                                        C3[4] * sh[39 + ch] * 4.f * 2.f * xz + C3[5] * sh[42 + ch] * (xx - yy);
                        }
                    }
                }
            }
            for (int k = 0; k < M; ++k) {
                const float bk = k < nb ? basis[k < 16 ? k : 15] : 0.f;
                dsh[3 * k] = bk * dRGB[0]; dsh[3 * k + 1] = bk * dRGB[1]; dsh[3 * k + 2] = bk * dRGB[2];
            }
            const float ddx = dRdx[0] * dRGB[0] + dRdx[1] * dRGB[1] + dRdx[2] * dRGB[2];
            const float ddy = dRdy[0] * dRGB[0] + dRdy[1] * dRGB[1] + dRdy[2] * dRGB[2];
            const float ddz = dRdz[0] * dRGB[0] + dRdz[1] * dRGB[1] + dRdz[2] * dRGB[2];
            // dnormvdv, auxiliary.h:106-117
            const float sum2 = ox * ox + oy * oy + oz * oz;
            const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            dmean[0] += ((+sum2 - ox * ox) * ddx - oy * ox * ddy - oz * ox * ddz) * invsum32;
            dmean[1] += (-ox * oy * ddx + (sum2 - oy * oy) * ddy - oz * oy * ddz) * invsum32;
            dmean[2] += (-ox * oz * ddx - oy * oz * ddy + (sum2 - oz * oz) * ddz) * invsum32;
        }

        // ---- cov3D -> scale / rotation (backward.cu:278-341) ----
        if (scales != nullptr) {
            float4 q;  // 16-byte load when aligned, scalar loads for an offset contiguous view
            if ((reinterpret_cast<uintptr_t>(rotations) & 15u) == 0) q = reinterpret_cast<const float4*>(rotations)[idx];
            else q = make_float4(rotations[4 * (size_t)idx], rotations[4 * (size_t)idx + 1], rotations[4 * (size_t)idx + 2], rotations[4 * (size_t)idx + 3]);
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            // Rm[c][r] as the glm matrix in the forward
            const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                                   {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                                   {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
            const float s[3] = {scale_modifier * scales[3 * idx], scale_modifier * scales[3 * idx + 1],
                                scale_modifier * scales[3 * idx + 2]};
            float Mm[3][3];  // M = S*R : M[c][r] = s_r R[c][r]
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) Mm[c][rr] = s[rr] * R[c][rr];
            // dL_dSigma (glm column-major, symmetric)
            const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                                    {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                                    {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
            // dL_dM = 2 * M * dL_dSigma  (glm product: (A*B)[c][r] = sum_k A[k][r] * B[c][k])
            float dM[3][3];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int rr = 0; rr < 3; ++rr)
                    dM[c][rr] = 2.0f * (Mm[0][rr] * dS[c][0] + Mm[1][rr] * dS[c][1] + Mm[2][rr] * dS[c][2]);
            // Rt = transpose(R): Rt[c][r] = R[r][c];  dL_dMt[c][r] = dM[r][c]
            // dL_dscale_j = dot(Rt[j], dL_dMt[j]) = sum_r R[r][j] * dM[r][j]
            float dMt[3][3];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) dMt[c][rr] = dM[rr][c];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                dscale[j] = R[0][j] * dMt[j][0] + R[1][j] * dMt[j][1] + R[2][j] * dMt[j][2];
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) dMt[j][rr] *= s[j];
            }
            drot[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
            drot[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) -
                      4 * x * (dMt[2][2] + dMt[1][1]);
            drot[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) -
                      4 * y * (dMt[2][2] + dMt[0][0]);
            drot[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) -
                      4 * z * (dMt[1][1] + dMt[0][0]);
        }
    } else if (shs != nullptr) {
        float* dsh = dL_dsh + (size_t)idx * M * 3;
        for (int k = 0; k < 3 * M; ++k) dsh[k] = 0.f;
    }

    dL_dmean3D[3 * idx] = dmean[0]; dL_dmean3D[3 * idx + 1] = dmean[1]; dL_dmean3D[3 * idx + 2] = dmean[2];
    if (dL_dcov3D) {
#pragma unroll
        for (int k = 0; k < 6; ++k) dL_dcov3D[6 * (size_t)idx + k] = dcov[k];
    }
    if (dL_dscale) { dL_dscale[3 * idx] = dscale[0]; dL_dscale[3 * idx + 1] = dscale[1]; dL_dscale[3 * idx + 2] = dscale[2]; }
    if (dL_drot) reinterpret_cast<float4*>(dL_drot)[idx] = make_float4(drot[0], drot[1], drot[2], drot[3]);
}

void launch_preprocess_bwd(const grpg_backward_args* a, const float* cov3D, const uint8_t* clamped,
                           const float* grad_rec, cudaStream_t stream) {
    const int P = a->p_count;
    if (P <= 0) return;
    const float focal_y = a->height / (2.0f * a->tan_fovy);
    const float focal_x = a->width / (2.0f * a->tan_fovx);
    PeerGrad peers;
    peers.n = a->n_peer_grad > 0 ? (a->n_peer_grad < 8 ? a->n_peer_grad : 8) : 0;
    for (int r = 0; r < 8; ++r) peers.p[r] = r < peers.n ? a->peer_grad_ws[r] : nullptr;
    ProfScope ps("preprocess_bwd", stream);
#define GRPG_PBWD(MB)                                                                                                    \
    preprocess_bwd_kernel<MB><<<(P + 255) / 256, 256, 0, stream>>>(                                                      \
        P, a->p_begin, a->D, a->M, a->means3D, a->radii, a->shs, clamped, a->scales, a->rotations, a->scale_modifier, cov3D, \
        a->viewmatrix, a->projmatrix, focal_x, focal_y, a->tan_fovx, a->tan_fovy, a->cam_pos, grad_rec, peers, a->dL_dmean2D, \
        a->dL_dconic, a->dL_dopacity, a->dL_dcolor, a->dL_ddepth, a->dL_dmean3D, a->dL_dcov3D, a->dL_dsh, a->dL_dscale,     \
        a->dL_drot)
    GRPG_PBWD(4);  // 64 registers, 4 CTAs/SM: measured 0.129 ms (1: 90 regs 0.186, 3: 79 regs 0.145, 5: 48 regs + spills 0.135)
#undef GRPG_PBWD
}

}  // namespace grpg
