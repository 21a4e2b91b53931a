// image_epilogue.cu -- sky composite + clamp + float->uint8 + CHW->HWC, optionally straight into pinned host memory.
// Behavioural spec: street_gaussian_renderer.py:336-346 and simulator.py:313-314 (see include/grpg_image.h).
#include <cstdint>
#include <cuda_runtime.h>

#include "grpg_common.cuh"
#include "../../include/grpg_image.h"

extern "C" int grpg_loss_fail(const char* msg);

namespace grpg {

// the reference's float32 operation sequence: t = sky * (1 - acc); r = rgb + t; clamp; (r * 255) truncated
__device__ __forceinline__ float composite(float c, float sky, float acc, bool has_sky) {
    float r = c;
    if (has_sky) r = fadd(c, fmul(sky, fadd(1.0f, -acc)));
    return fminf(fmaxf(r, 0.0f), 1.0f);  // NaN -> 0 (the reference's uint8 cast of NaN is undefined)
}
__device__ __forceinline__ uint32_t to_u8(float clamped) { return (uint32_t)__float2int_rz(fmul(clamped, 255.0f)); }

// One thread = 4 horizontally adjacent pixels: 3 (+4) coalesced 16-byte loads, one 12-byte interleaved store.
__global__ void __launch_bounds__(256) compose_rgb8_vec4_kernel(int H, int W, const float* __restrict__ rgb,
                                                                const float* __restrict__ acc,
                                                                const float* __restrict__ sky, uint8_t* out8,
                                                                float* __restrict__ outf) {
    const size_t hw = (size_t)H * W;
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 pixels
    if (q * 4 >= hw) return;
    const bool has_sky = sky != nullptr;
    float4 c[3], s[3], a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        c[ch] = __ldcs(reinterpret_cast<const float4*>(rgb + ch * hw) + q);
        s[ch] = has_sky ? __ldcs(reinterpret_cast<const float4*>(sky + ch * hw) + q) : a;
    }
    if (has_sky) a = __ldcs(reinterpret_cast<const float4*>(acc) + q);
    float v[4][3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        v[0][ch] = composite(c[ch].x, s[ch].x, a.x, has_sky); v[1][ch] = composite(c[ch].y, s[ch].y, a.y, has_sky);
        v[2][ch] = composite(c[ch].z, s[ch].z, a.z, has_sky); v[3][ch] = composite(c[ch].w, s[ch].w, a.w, has_sky);
        if (outf) reinterpret_cast<float4*>(outf + ch * hw)[q] = make_float4(v[0][ch], v[1][ch], v[2][ch], v[3][ch]);
    }
    if (out8) {
        uint32_t b[12];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) b[3 * p + ch] = to_u8(v[p][ch]);
        uint32_t w0 = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
        uint32_t w1 = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
        uint32_t w2 = b[8] | (b[9] << 8) | (b[10] << 16) | (b[11] << 24);
        uint32_t* o = reinterpret_cast<uint32_t*>(out8 + q * 12);
        o[0] = w0; o[1] = w1; o[2] = w2;
    }
}

// generic path: one thread per pixel (H*W not a multiple of 4, or misaligned views)
__global__ void __launch_bounds__(256) compose_rgb8_scalar_kernel(int H, int W, const float* __restrict__ rgb,
                                                                  const float* __restrict__ acc,
                                                                  const float* __restrict__ sky, uint8_t* out8,
                                                                  float* __restrict__ outf) {
    const size_t hw = (size_t)H * W;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= hw) return;
    const bool has_sky = sky != nullptr;
    const float a = has_sky ? acc[p] : 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float v = composite(rgb[ch * hw + p], has_sky ? sky[ch * hw + p] : 0.f, a, has_sky);
        if (outf) outf[ch * hw + p] = v;
        if (out8) out8[p * 3 + ch] = (uint8_t)to_u8(v);
    }
}

}  // namespace grpg

extern "C" int grpg_compose_rgb8(const grpg_rgb8_args* a) {
    using namespace grpg;
    if (!a) return grpg_loss_fail("grpg_compose_rgb8: null arguments");
    if (a->height < 0 || a->width < 0) return grpg_loss_fail("grpg_compose_rgb8: bad sizes");
    const size_t hw = (size_t)a->height * a->width;
    if (hw == 0) return 0;
    if (!a->rgb) return grpg_loss_fail("grpg_compose_rgb8: missing rgb");
    if (a->sky && !a->acc) return grpg_loss_fail("grpg_compose_rgb8: sky needs acc");
    if (!a->out_rgb8 && !a->out_rgb) return 0;
    cudaStream_t stream = (cudaStream_t)a->stream;
    uint8_t* out8 = a->out_rgb8;
    if (out8) {  // pinned host memory: use the device alias of the mapping
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, out8) != cudaSuccess) { cudaGetLastError(); return grpg_loss_fail("grpg_compose_rgb8: out_rgb8 is not CUDA-accessible memory"); }
        if (attr.type == cudaMemoryTypeHost) out8 = (uint8_t*)attr.devicePointer;
        else if (attr.type == cudaMemoryTypeUnregistered)
            return grpg_loss_fail("grpg_compose_rgb8: out_rgb8 must be device memory or pinned host memory");
    }
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    const bool vec = hw % 4 == 0 && al16(a->rgb) && (!a->sky || (al16(a->sky) && al16(a->acc))) &&
                     (!a->out_rgb || al16(a->out_rgb)) && (!out8 || (reinterpret_cast<uintptr_t>(out8) & 3u) == 0);
    ProfScope ps("compose_rgb8", stream);
    if (vec) {
        const size_t groups = hw / 4;
        compose_rgb8_vec4_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, stream>>>(a->height, a->width, a->rgb, a->acc,
                                                                                     a->sky, out8, a->out_rgb);
    } else {
        compose_rgb8_scalar_kernel<<<(unsigned)((hw + 255) / 256), 256, 0, stream>>>(a->height, a->width, a->rgb, a->acc,
                                                                                   a->sky, out8, a->out_rgb);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return grpg_loss_fail(cudaGetErrorString(e));
    return 0;
}
