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

    // A: stage inputs.  All of a thread's global loads are issued before the first shared store, so the
    // ~11 x 2 loads per thread are in flight together (the SM holds only two of these CTAs).
    constexpr int A_ITERS = (LIN * LIN + LOSS_THREADS - 1) / LOSS_THREADS;
    float va[A_ITERS], vb[A_ITERS];
#pragma unroll
    for (int it = 0; it < A_ITERS; ++it) {
        const int i = tid + it * LOSS_THREADS;
        const int r = i / LIN, c = i - r * LIN;
        const int gy = y0 - 2 * LR + r, gx = x0 - 2 * LR + c;
        va[it] = 0.f; vb[it] = 0.f;
        if (i < LIN * LIN && gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const size_t o = (size_t)gy * W + gx;
            if (!pm || pm[o]) { va[it] = __ldg(p1 + o); vb[it] = __ldg(p2 + o); }
        }
    }
#pragma unroll
    for (int it = 0; it < A_ITERS; ++it) {
        const int i = tid + it * LOSS_THREADS;
        if (i < LIN * LIN) { (&s.x[0][0])[i] = va[it]; (&s.y[0][0])[i] = vb[it]; }
    }
    __syncthreads();

    // B: horizontal pass of the five statistics.  One work item = 7 adjacent outputs of one row: 17 staged
    // pixels are read once (instead of 7 x 11) and their squares/products are formed once per pixel.
    constexpr int BS = 7, B_STRIPS = LST / BS;  // 42 = 6 x 7
    for (int item = tid; item < LIN * B_STRIPS; item += LOSS_THREADS) {
        const int r = item / B_STRIPS, c0 = (item - r * B_STRIPS) * BS;
        float sx[BS], sy[BS], sxx[BS], syy[BS], sxy[BS];
#pragma unroll
        for (int o = 0; o < BS; ++o) { sx[o] = 0.f; sy[o] = 0.f; sxx[o] = 0.f; syy[o] = 0.f; sxy[o] = 0.f; }
#pragma unroll
        for (int j = 0; j < BS + 10; ++j) {
            const float a = s.x[r][c0 + j], b = s.y[r][c0 + j];
            const float aa = a * a, bb = b * b, ab = a * b;
#pragma unroll
            for (int o = 0; o < BS; ++o) {
                const int k = j - o;
                if (k >= 0 && k < 11) {
                    sx[o] = fmaf(w[k], a, sx[o]); sy[o] = fmaf(w[k], b, sy[o]);
                    sxx[o] = fmaf(w[k], aa, sxx[o]); syy[o] = fmaf(w[k], bb, syy[o]); sxy[o] = fmaf(w[k], ab, sxy[o]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < BS; ++o) {
            s.h[0][r][c0 + o] = sx[o]; s.h[1][r][c0 + o] = sy[o]; s.h[2][r][c0 + o] = sxx[o];
            s.h[3][r][c0 + o] = syy[o]; s.h[4][r][c0 + o] = sxy[o];
        }
    }
    __syncthreads();

    // C: vertical pass, SSIM value and partials.  One work item = 7 vertically adjacent outputs of one column
    // (lanes walk along the row: conflict-free): 17 x 5 loads feed 7 x 5 x 11 FMAs.
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    float ssim_acc = 0.f;
    constexpr int CS = 7, C_STRIPS = LST / CS;
    for (int item = tid; item < LST * C_STRIPS; item += LOSS_THREADS) {
        const int strip = item / LST, c = item - strip * LST, r0 = strip * CS;
        float acc[5][CS];
#pragma unroll
        for (int q = 0; q < 5; ++q)
#pragma unroll
            for (int o = 0; o < CS; ++o) acc[q][o] = 0.f;
#pragma unroll
        for (int j = 0; j < CS + 10; ++j) {
            float v[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) v[q] = s.h[q][r0 + j][c];
#pragma unroll
            for (int o = 0; o < CS; ++o) {
                const int k = j - o;
                if (k >= 0 && k < 11) {
#pragma unroll
                    for (int q = 0; q < 5; ++q) acc[q][o] = fmaf(w[k], v[q], acc[q][o]);
                }
            }
        }
        const int gx = x0 - LR + c;
#pragma unroll
        for (int o = 0; o < CS; ++o) {
            const int r = r0 + o, gy = y0 - LR + r;
            float d_mu1 = 0.f, d_e11 = 0.f, d_e12 = 0.f;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                const float mu1 = acc[0][o], mu2 = acc[1][o], e11 = acc[2][o], e22 = acc[3][o], e12 = acc[4][o];
                const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
                const float s1 = e11 - mu1_sq, s2 = e22 - mu2_sq, s12 = e12 - mu12;
                const float A1 = 2.f * mu12 + C1, A2 = 2.f * s12 + C2;
                const float B1 = mu1_sq + mu2_sq + C1, B2 = s1 + s2 + C2;
                const float iB1 = 1.f / B1, iB2 = 1.f / B2;
                const float inv = iB1 * iB2;
                const float S = (A1 * A2) * inv;  // loss_utils.py:119
                d_mu1 = 2.f * mu2 * (A2 - A1) * inv - S * (2.f * mu1) * (iB1 - iB2);
                d_e11 = -S * iB2;
                d_e12 = 2.f * A1 * inv;
                // this CTA's own pixels contribute to the sum once
                if (r >= LR && r < LR + LT && c >= LR && c < LR + LT) ssim_acc += S;
            }
            s.d[0][r][c] = d_mu1; s.d[1][r][c] = d_e11; s.d[2][r][c] = d_e12;
        }
    }
    __syncthreads();

    if (grad != nullptr) {
        // D: horizontal pass of the partials (aliases h); one work item = 8 adjacent outputs of one row
        constexpr int DS = 8, D_STRIPS = LT / DS;
        for (int item = tid; item < LST * D_STRIPS; item += LOSS_THREADS) {
            const int r = item / D_STRIPS, c0 = (item - r * D_STRIPS) * DS;
            float acc[3][DS];
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int o = 0; o < DS; ++o) acc[q][o] = 0.f;
#pragma unroll
            for (int j = 0; j < DS + 10; ++j) {
                const float v0 = s.d[0][r][c0 + j], v1 = s.d[1][r][c0 + j], v2 = s.d[2][r][c0 + j];
#pragma unroll
                for (int o = 0; o < DS; ++o) {
                    const int k = j - o;
                    if (k >= 0 && k < 11) {
                        acc[0][o] = fmaf(w[k], v0, acc[0][o]); acc[1][o] = fmaf(w[k], v1, acc[1][o]);
                        acc[2][o] = fmaf(w[k], v2, acc[2][o]);
                    }
                }
            }
#pragma unroll
            for (int o = 0; o < DS; ++o) {
                s.hd[0][r][c0 + o] = acc[0][o]; s.hd[1][r][c0 + o] = acc[1][o]; s.hd[2][r][c0 + o] = acc[2][o];
            }
        }
        __syncthreads();
    }

    // E: vertical pass, gradient, L1; one work item = 4 vertically adjacent pixels of one column (256 items)
    float l1_acc = 0.f;
    constexpr int ES = 4;
    {
        const int c = tid & (LT - 1), r0 = (tid / LT) * ES;
        const int gx = x0 + c;
        float acc[3][ES];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int o = 0; o < ES; ++o) acc[q][o] = 0.f;
        if (grad != nullptr) {
#pragma unroll
            for (int j = 0; j < ES + 10; ++j) {
                const float v0 = s.hd[0][r0 + j][c], v1 = s.hd[1][r0 + j][c], v2 = s.hd[2][r0 + j][c];
#pragma unroll
                for (int o = 0; o < ES; ++o) {
                    const int k = j - o;
                    if (k >= 0 && k < 11) {
                        acc[0][o] = fmaf(w[k], v0, acc[0][o]); acc[1][o] = fmaf(w[k], v1, acc[1][o]);
                        acc[2][o] = fmaf(w[k], v2, acc[2][o]);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 0; o < ES; ++o) {
            const int r = r0 + o, gy = y0 + r;
            if (gy >= H || gx >= W) continue;
            const size_t off = (size_t)gy * W + gx;
            const bool valid = !pm || pm[off];
            const float a = s.x[r + 2 * LR][c + 2 * LR], b = s.y[r + 2 * LR][c + 2 * LR];  // already 0 when masked
            const float diff = a - b;
            if (valid) l1_acc += fabsf(diff);
            if (grad != nullptr) {
                float g = 0.f;
                if (valid) {
                    const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);  // torch.abs backward: sign(0) = 0
                    g = coef_l1 * sgn + coef_ssim * (acc[0][o] + 2.f * a * acc[1][o] + b * acc[2][o]);
                }
                grad[plane * hw + off] = g;
            }
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
