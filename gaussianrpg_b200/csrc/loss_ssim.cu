// loss_ssim.cu -- fused L1 + SSIM loss: value and gradient in one pass over the two images.
//
// Behavioural spec: lib/utils/loss_utils.py:21-37 (l1_loss), :81-124 (ssim/_ssim) and their autograd.
// One CTA owns a 32x32 tile of one image plane.  Everything between the two global reads (img1, img2) and
// the global write (gradient) lives in shared memory:
//
//   A  stage img1/img2 on the tile + 10-pixel halo (zero outside the image = conv2d's padding; masked
//      pixels read as 0 in BOTH images, loss_utils.py:96-98)
//   B  horizontal 11-tap pass of the five window statistics  x, y, x^2, y^2, xy     (52 rows x 42 columns)
//   C  vertical pass -> mu1, mu2, E11, E22, E12 on the tile + 5-pixel halo, then per pixel the SSIM value and
//      the partials of SSIM with respect to (mu1, E11, E12) -- the three statistics img1 enters; statistics
//      outside the image do not exist in the reference, their partials are 0
//   D  horizontal pass of the three partial maps, E vertical pass, then
//      d(sum ssim)/d img1(p) = conv(D_mu1)(p) + 2 img1(p) conv(D_E11)(p) + img2(p) conv(D_E12)(p)
//
// The second halo (10 instead of 5 pixels) trades 1.6x of the arithmetic for never writing the 3 partial maps
// to HBM and reading them back: the kernel's HBM traffic is 2 reads + 1 write of the image.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "grpg_common.cuh"
#include "../../include/grpg_loss.h"

namespace grpg {

constexpr int LT = 32;            // tile edge
constexpr int LR = 5;             // window radius
constexpr int LIN = LT + 4 * LR;  // 52: staged input edge
constexpr int LST = LT + 2 * LR;  // 42: statistics edge
constexpr int LOSS_THREADS = 256;

// gaussian(11, 1.5) exactly as the reference's float32 tensor (loss_utils.py:81-83)
__constant__ float c_win[11] = {0x1.0d956cp-10f, 0x1.f1fe02p-8f, 0x1.26eb18p-5f, 0x1.bff0fep-4f, 0x1.b43c3ep-3f,
                                0x1.106560p-2f,  0x1.b43c3ep-3f, 0x1.bff0fep-4f, 0x1.26eb18p-5f, 0x1.f1fe02p-8f,
                                0x1.0d956cp-10f};

struct LossSmem {
    float x[LIN][LIN];
    float y[LIN][LIN];
    union {
        float h[5][LIN][LST];   // B: horizontally filtered statistics
        float hd[3][LST][LT];   // D: horizontally filtered partials (h is dead by then)
    };
    float d[3][LST][LST];       // C: partials of SSIM
    float red[2][LOSS_THREADS / 32];
};

__global__ void __launch_bounds__(LOSS_THREADS) l1_ssim_kernel(
    int H, int W, int planes_per_mask, const float* __restrict__ img1, const float* __restrict__ img2,
    const uint8_t* __restrict__ mask, float coef_l1, float coef_ssim, double* __restrict__ sums,
    float* __restrict__ grad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LossSmem& s = *reinterpret_cast<LossSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int plane = blockIdx.z;
    const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
    const size_t hw = (size_t)H * W;
    const float* p1 = img1 + plane * hw;
    const float* p2 = img2 + plane * hw;
    const uint8_t* pm = mask ? mask + (size_t)(plane / planes_per_mask) * hw : nullptr;
    float w[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) w[k] = c_win[k];

    // A: stage inputs
    for (int i = tid; i < LIN * LIN; i += LOSS_THREADS) {
        const int r = i / LIN, c = i - r * LIN;
        const int gy = y0 - 2 * LR + r, gx = x0 - 2 * LR + c;
        float a = 0.f, b = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const size_t o = (size_t)gy * W + gx;
            if (!pm || pm[o]) { a = __ldg(p1 + o); b = __ldg(p2 + o); }
        }
        s.x[r][c] = a; s.y[r][c] = b;
    }
    __syncthreads();

    // B: horizontal pass of the five statistics
    for (int i = tid; i < LIN * LST; i += LOSS_THREADS) {
        const int r = i / LST, c = i - r * LST;
        float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float a = s.x[r][c + k], b = s.y[r][c + k];
            const float wa = w[k] * a, wb = w[k] * b;
            sx += wa; sy += wb; sxx = fmaf(wa, a, sxx); syy = fmaf(wb, b, syy); sxy = fmaf(wa, b, sxy);
        }
        s.h[0][r][c] = sx; s.h[1][r][c] = sy; s.h[2][r][c] = sxx; s.h[3][r][c] = syy; s.h[4][r][c] = sxy;
    }
    __syncthreads();

    // C: vertical pass, SSIM value and partials
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    float ssim_acc = 0.f;
    for (int i = tid; i < LST * LST; i += LOSS_THREADS) {
        const int r = i / LST, c = i - r * LST;
        const int gy = y0 - LR + r, gx = x0 - LR + c;
        float d_mu1 = 0.f, d_e11 = 0.f, d_e12 = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
            for (int k = 0; k < 11; ++k) {
                mu1 = fmaf(w[k], s.h[0][r + k][c], mu1); mu2 = fmaf(w[k], s.h[1][r + k][c], mu2);
                e11 = fmaf(w[k], s.h[2][r + k][c], e11); e22 = fmaf(w[k], s.h[3][r + k][c], e22);
                e12 = fmaf(w[k], s.h[4][r + k][c], e12);
            }
            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
            const float s1 = e11 - mu1_sq, s2 = e22 - mu2_sq, s12 = e12 - mu12;
            const float A1 = 2.f * mu12 + C1, A2 = 2.f * s12 + C2;
            const float B1 = mu1_sq + mu2_sq + C1, B2 = s1 + s2 + C2;
            const float inv = 1.f / (B1 * B2);
            const float S = (A1 * A2) * inv;  // loss_utils.py:119
            d_mu1 = 2.f * mu2 * (A2 - A1) * inv - S * (2.f * mu1 / B1 - 2.f * mu1 / B2);
            d_e11 = -S / B2;
            d_e12 = 2.f * A1 * inv;
            // this CTA's own pixels contribute to the sum once
            if (r >= LR && r < LR + LT && c >= LR && c < LR + LT) ssim_acc += S;
        }
        s.d[0][r][c] = d_mu1; s.d[1][r][c] = d_e11; s.d[2][r][c] = d_e12;
    }
    __syncthreads();

    if (grad != nullptr) {
        // D: horizontal pass of the partials (aliases h)
        for (int i = tid; i < LST * LT; i += LOSS_THREADS) {
            const int r = i / LT, c = i - r * LT;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int k = 0; k < 11; ++k) {
                a0 = fmaf(w[k], s.d[0][r][c + k], a0); a1 = fmaf(w[k], s.d[1][r][c + k], a1);
                a2 = fmaf(w[k], s.d[2][r][c + k], a2);
            }
            s.hd[0][r][c] = a0; s.hd[1][r][c] = a1; s.hd[2][r][c] = a2;
        }
        __syncthreads();
    }

    // E: vertical pass, gradient, L1
    float l1_acc = 0.f;
    for (int i = tid; i < LT * LT; i += LOSS_THREADS) {
        const int r = i / LT, c = i - r * LT;
        const int gy = y0 + r, gx = x0 + c;
        if (gy >= H || gx >= W) continue;
        const size_t o = (size_t)gy * W + gx;
        const bool valid = !pm || pm[o];
        const float a = s.x[r + 2 * LR][c + 2 * LR], b = s.y[r + 2 * LR][c + 2 * LR];  // already 0 when masked
        const float diff = a - b;
        if (valid) l1_acc += fabsf(diff);
        if (grad != nullptr) {
            float g = 0.f;
            if (valid) {
                float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
                for (int k = 0; k < 11; ++k) {
                    g0 = fmaf(w[k], s.hd[0][r + k][c], g0); g1 = fmaf(w[k], s.hd[1][r + k][c], g1);
                    g2 = fmaf(w[k], s.hd[2][r + k][c], g2);
                }
                const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);  // torch.abs backward: sign(0) = 0
                g = coef_l1 * sgn + coef_ssim * (g0 + 2.f * a * g1 + b * g2);
            }
            grad[plane * hw + o] = g;
        }
    }

    // block reduction -> one double atomic per sum and CTA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l1_acc += __shfl_xor_sync(0xffffffffu, l1_acc, o);
        ssim_acc += __shfl_xor_sync(0xffffffffu, ssim_acc, o);
    }
    if ((tid & 31) == 0) { s.red[0][tid >> 5] = l1_acc; s.red[1][tid >> 5] = ssim_acc; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int k = 0; k < LOSS_THREADS / 32; ++k) { a += (double)s.red[0][k]; b += (double)s.red[1][k]; }
        atomicAdd(sums + 2 * plane, a);
        atomicAdd(sums + 2 * plane + 1, b);
    }
}

}  // namespace grpg

extern "C" int grpg_loss_fail(const char* msg);  // api.cu: records the message for grpg_last_error()

extern "C" int grpg_l1_ssim(const grpg_l1_ssim_args* a) {
    using namespace grpg;
    if (!a) return grpg_loss_fail("grpg_l1_ssim: null arguments");
    if (a->planes < 0 || a->height < 0 || a->width < 0) return grpg_loss_fail("grpg_l1_ssim: bad sizes");
    if (a->planes == 0 || a->height == 0 || a->width == 0) return 0;
    if (!a->img1 || !a->img2 || !a->sums) return grpg_loss_fail("grpg_l1_ssim: missing pointer");
    if (a->mask && (a->planes_per_mask <= 0 || a->planes % a->planes_per_mask != 0))
        return grpg_loss_fail("grpg_l1_ssim: planes must be a multiple of planes_per_mask");
    if (a->planes > 65535) return grpg_loss_fail("grpg_l1_ssim: more than 65535 planes");
    cudaStream_t stream = (cudaStream_t)a->stream;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(l1_ssim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LossSmem));
        configured = true;
    }
    cudaMemsetAsync(a->sums, 0, (size_t)a->planes * 2 * sizeof(double), stream);
    const dim3 grid((a->width + LT - 1) / LT, (a->height + LT - 1) / LT, a->planes);
    ProfScope ps("l1_ssim", stream);
    l1_ssim_kernel<<<grid, LOSS_THREADS, sizeof(LossSmem), stream>>>(
        a->height, a->width, a->planes_per_mask > 0 ? a->planes_per_mask : 1, a->img1, a->img2, a->mask, a->coef_l1,
        a->coef_ssim, a->sums, a->grad);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return grpg_loss_fail(cudaGetErrorString(e));
    return 0;
}
