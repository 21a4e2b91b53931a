// sky_cubemap.cu -- sky cube-map lookup, its backward, and the lookup fused into the post-render epilogue
// (include/grpg_sky.h; behavioural spec: lib/models/sky_cubemap.py:77-124 of the reference, nvdiffrast cube mapping).
#include <cstdint>
#include <cuda_runtime.h>

#include "grpg_common.cuh"
#include "../../include/grpg_sky.h"

extern "C" int grpg_loss_fail(const char* msg);

namespace grpg {

// Face and face coordinates (u, v in [0, 1]) of a direction: major axis, OpenGL face order +x,-x,+y,-y,+z,-z.
__device__ __forceinline__ int cube_face(float x, float y, float z, float& u, float& v) {
    const float ax = fabsf(x), ay = fabsf(y), az = fabsf(z);
    int idx;
    float c;
    if (az > fmaxf(ax, ay)) { idx = 4; c = z; }
    else if (ay > ax) { idx = 2; c = y; y = z; }
    else { idx = 0; c = x; x = z; }
    if (c < 0.f) idx += 1;
    const float m = 0.5f / fabsf(c);
    const float m0 = (idx == 0 || idx == 5) ? -m : m;
    const float m1 = (idx != 2) ? -m : m;
    u = x * m0 + 0.5f;
    v = y * m1 + 0.5f;
    if (!isfinite(u) || !isfinite(v)) return -1;
    u = fminf(fmaxf(u, 0.f), 1.f);
    v = fminf(fmaxf(v, 0.f), 1.f);
    return idx;
}

// Inverse of cube_face on the (extended) face plane: the point with face coordinates a = 2u - 1, b = 2v - 1.
__device__ __forceinline__ void face_point(int idx, float a, float b, float& x, float& y, float& z) {
    switch (idx) {
        case 0: x = 1.f; z = -a; y = -b; break;
        case 1: x = -1.f; z = a; y = -b; break;
        case 2: y = 1.f; x = a; z = b; break;
        case 3: y = -1.f; x = a; z = -b; break;
        case 4: z = 1.f; x = a; y = -b; break;
        default: z = -1.f; x = -a; y = -b; break;
    }
}

// Linear index (in texels) of texel (i, j) of `face`, where i or j may lie one step off the face: the tap is then the
// texel of the adjacent face across that edge (the cube unfolded around the edge).  -1 at a cube corner.
__device__ __forceinline__ int cube_texel(int face, int i, int j, int res) {
    const bool oi = i < 0 || i >= res, oj = j < 0 || j >= res;
    if (!oi && !oj) return (face * res + j) * res + i;
    if (oi && oj) return -1;
    const float inv = 1.0f / (float)res;
    const float a = (2.f * i + 1.f) * inv - 1.f, b = (2.f * j + 1.f) * inv - 1.f;  // texel centre, |.| = 1 + 1/res off the face
    float q[3];
    face_point(face, a, b, q[0], q[1], q[2]);
    // fold around the edge (the cube unfolded): the coordinate that left [-1, 1] becomes the new major axis, the old
    // major axis moves in from the edge by the overshoot (one texel centre = 1/res); the third coordinate, which runs
    // along the shared edge, is kept -- so the tap is the texel at the same position along the edge
    const int A = face >> 1;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        if (ax != A && fabsf(q[ax]) > 1.f) q[ax] = copysignf(1.f, q[ax]);
    }
#pragma unroll
    for (int ax = 0; ax < 3; ++ax)
        if (ax == A) q[ax] *= (1.f - inv);
    float u, v;
    const int f2 = cube_face(q[0], q[1], q[2], u, v);
    if (f2 < 0) return -1;
    const int i2 = min(res - 1, (int)(u * res)), j2 = min(res - 1, (int)(v * res));
    return (f2 * res + j2) * res + i2;
}

struct Taps {
    int idx[4];
    float w[4];
};

// the four bilinear taps of a direction; weights of missing taps are 0 and the rest renormalised
__device__ __forceinline__ bool cube_taps(float dx, float dy, float dz, int res, Taps& t) {
    float u, v;
    const int face = cube_face(dx, dy, dz, u, v);
    if (face < 0) return false;
    const float fu = u * res - 0.5f, fv = v * res - 0.5f;
    const float i0f = floorf(fu), j0f = floorf(fv);
    const int i0 = (int)i0f, j0 = (int)j0f;
    const float wu = fu - i0f, wv = fv - j0f;
    t.idx[0] = cube_texel(face, i0, j0, res);         t.w[0] = (1.f - wu) * (1.f - wv);
    t.idx[1] = cube_texel(face, i0 + 1, j0, res);     t.w[1] = wu * (1.f - wv);
    t.idx[2] = cube_texel(face, i0, j0 + 1, res);     t.w[2] = (1.f - wu) * wv;
    t.idx[3] = cube_texel(face, i0 + 1, j0 + 1, res); t.w[3] = wu * wv;
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (t.idx[k] < 0) t.w[k] = 0.f; sum += t.w[k]; }
    if (sum <= 0.f) return false;
    if (sum != 1.f) {
        const float r = 1.f / sum;
#pragma unroll
        for (int k = 0; k < 4; ++k) t.w[k] *= r;
    }
    return true;
}

struct SkyParams {
    int H, W, res;
    const float* cubemap;
    const float* M;
    const float* jitter;
    const uint8_t* mask;
    const float* acc;
    float fill;
};

__device__ __forceinline__ bool sky_pixel_taps(const SkyParams& p, const float* sM, int px, int py, Taps& t) {
    const size_t pid = (size_t)py * p.W + px;
    if (p.mask != nullptr) { if (p.mask[pid] == 0) return false; }
    else if (p.acc != nullptr) { if (!((1.0f - p.acc[pid]) > 1e-3f)) return false; }
    const float jx = p.jitter ? p.jitter[pid] : 0.5f, jy = p.jitter ? p.jitter[(size_t)p.H * p.W + pid] : 0.5f;
    const float x = (float)px + jx, y = (float)py + jy;
    float dx = sM[0] * x + sM[1] * y + sM[2], dy = sM[3] * x + sM[4] * y + sM[5], dz = sM[6] * x + sM[7] * y + sM[8];
    const float inv = rsqrtf(dx * dx + dy * dy + dz * dz);  // rays_d / |rays_d| (graphics_utils.py:205)
    return cube_taps(dx * inv, dy * inv, dz * inv, p.res, t);
}

__global__ void __launch_bounds__(256) sky_forward_kernel(SkyParams p, float* __restrict__ sky) {
    __shared__ float sM[9];
    if (threadIdx.x < 9) sM[threadIdx.x] = p.M[threadIdx.x];
    __syncthreads();
    const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (px >= p.W || py >= p.H) return;
    const size_t hw = (size_t)p.H * p.W, pid = (size_t)py * p.W + px;
    Taps t;
    float c[3] = {p.fill, p.fill, p.fill};
    if (sky_pixel_taps(p, sM, px, py, t)) {
        c[0] = c[1] = c[2] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (t.w[k] == 0.f) continue;
            const float* tex = p.cubemap + 3 * (size_t)t.idx[k];
            c[0] += t.w[k] * __ldg(tex); c[1] += t.w[k] * __ldg(tex + 1); c[2] += t.w[k] * __ldg(tex + 2);
        }
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) sky[ch * hw + pid] = fminf(fmaxf(c[ch], 0.f), 1.f);
}

__global__ void __launch_bounds__(256) sky_backward_kernel(SkyParams p, const float* __restrict__ sky,
                                                           const float* __restrict__ dL_dsky, float* __restrict__ d_cube) {
    __shared__ float sM[9];
    if (threadIdx.x < 9) sM[threadIdx.x] = p.M[threadIdx.x];
    __syncthreads();
    const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (px >= p.W || py >= p.H) return;
    const size_t hw = (size_t)p.H * p.W, pid = (size_t)py * p.W + px;
    Taps t;
    if (!sky_pixel_taps(p, sM, px, py, t)) return;
    float g[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        // the forward clamped: like torch.clamp, the gradient passes where the unclamped value lay in [0, 1]; the
        // stored value equals the bound both when it was on it and when it was cut, so the value is recomputed
        g[ch] = dL_dsky[ch * hw + pid];
    }
    float c[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (t.w[k] == 0.f) continue;
        const float* tex = p.cubemap + 3 * (size_t)t.idx[k];
        c[0] += t.w[k] * __ldg(tex); c[1] += t.w[k] * __ldg(tex + 1); c[2] += t.w[k] * __ldg(tex + 2);
    }
    (void)sky;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
        if (!(c[ch] >= 0.f && c[ch] <= 1.f)) g[ch] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (t.w[k] == 0.f) continue;
        float* d = d_cube + 3 * (size_t)t.idx[k];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
            if (g[ch] != 0.f) atomicAdd(d + ch, t.w[k] * g[ch]);
    }
}

// lookup + rgb + sky (1 - acc) + clamp + x255 + uint8 HWC (+ the float image): image_epilogue.cu's arithmetic
__global__ void __launch_bounds__(256) sky_compose_rgb8_kernel(SkyParams p, const float* __restrict__ rgb,
                                                               uint8_t* out8, float* __restrict__ outf) {
    __shared__ float sM[9];
    if (threadIdx.x < 9) sM[threadIdx.x] = p.M[threadIdx.x];
    __syncthreads();
    const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (px >= p.W || py >= p.H) return;
    const size_t hw = (size_t)p.H * p.W, pid = (size_t)py * p.W + px;
    Taps t;
    float c[3] = {p.fill, p.fill, p.fill};
    if (sky_pixel_taps(p, sM, px, py, t)) {
        c[0] = c[1] = c[2] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (t.w[k] == 0.f) continue;
            const float* tex = p.cubemap + 3 * (size_t)t.idx[k];
            c[0] += t.w[k] * __ldg(tex); c[1] += t.w[k] * __ldg(tex + 1); c[2] += t.w[k] * __ldg(tex + 2);
        }
    }
    const float one_m_acc = fadd(1.0f, -p.acc[pid]);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float s = fminf(fmaxf(c[ch], 0.f), 1.f);                                 // sky_color.clamp(0, 1)
        const float r = fminf(fmaxf(fadd(rgb[ch * hw + pid], fmul(s, one_m_acc)), 0.0f), 1.0f);
        if (outf) outf[ch * hw + pid] = r;
        if (out8) out8[pid * 3 + ch] = (uint8_t)__float2int_rz(fmul(r, 255.0f));
    }
}

static int sky_params(const grpg_sky_args* a, SkyParams& p, const char* who) {
    if (!a) return grpg_loss_fail("grpg_sky: null arguments");
    if (a->height < 0 || a->width < 0 || a->resolution <= 0) return grpg_loss_fail("grpg_sky: bad sizes");
    if (!a->cubemap || !a->ray_matrix) return grpg_loss_fail("grpg_sky: missing cube map or ray matrix");
    (void)who;
    p.H = a->height; p.W = a->width; p.res = a->resolution;
    p.cubemap = a->cubemap; p.M = a->ray_matrix; p.jitter = a->jitter; p.mask = a->mask; p.acc = a->acc; p.fill = a->fill;
    return 0;
}

static int sky_check(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return grpg_loss_fail(cudaGetErrorString(e));
    (void)what;
    return 0;
}

}  // namespace grpg

using namespace grpg;

extern "C" int grpg_sky_forward(const grpg_sky_args* a) {
    SkyParams p;
    if (int rc = sky_params(a, p, "forward")) return rc;
    if ((size_t)p.H * p.W == 0) return 0;
    if (!a->sky) return grpg_loss_fail("grpg_sky_forward: missing output");
    cudaStream_t stream = (cudaStream_t)a->stream;
    ProfScope ps("sky_forward", stream);
    sky_forward_kernel<<<dim3((p.W + 31) / 32, (p.H + 7) / 8), 256, 0, stream>>>(p, a->sky);
    return sky_check("grpg_sky_forward");
}

extern "C" int grpg_sky_backward(const grpg_sky_args* a, const float* dL_dsky, float* d_cubemap) {
    SkyParams p;
    if (int rc = sky_params(a, p, "backward")) return rc;
    if ((size_t)p.H * p.W == 0) return 0;
    if (!dL_dsky || !d_cubemap) return grpg_loss_fail("grpg_sky_backward: missing gradient buffers");
    cudaStream_t stream = (cudaStream_t)a->stream;
    ProfScope ps("sky_backward", stream);
    sky_backward_kernel<<<dim3((p.W + 31) / 32, (p.H + 7) / 8), 256, 0, stream>>>(p, a->sky, dL_dsky, d_cubemap);
    return sky_check("grpg_sky_backward");
}

extern "C" int grpg_sky_compose_rgb8(const grpg_sky_args* a, const float* rgb, uint8_t* out_rgb8, float* out_rgb) {
    SkyParams p;
    if (int rc = sky_params(a, p, "compose")) return rc;
    if ((size_t)p.H * p.W == 0) return 0;
    if (!rgb || !a->acc) return grpg_loss_fail("grpg_sky_compose_rgb8: rgb and acc are required");
    if (!out_rgb8 && !out_rgb) return 0;
    uint8_t* out8 = out_rgb8;
    if (out8) {  // pinned host memory: use the device alias of the mapping (see image_epilogue.cu)
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, out8) != cudaSuccess) { cudaGetLastError(); return grpg_loss_fail("grpg_sky_compose_rgb8: out_rgb8 is not CUDA-accessible memory"); }
        if (attr.type == cudaMemoryTypeHost) out8 = (uint8_t*)attr.devicePointer;
        else if (attr.type == cudaMemoryTypeUnregistered)
            return grpg_loss_fail("grpg_sky_compose_rgb8: out_rgb8 must be device memory or pinned host memory");
    }
    cudaStream_t stream = (cudaStream_t)a->stream;
    ProfScope ps("sky_compose_rgb8", stream);
    sky_compose_rgb8_kernel<<<dim3((p.W + 31) / 32, (p.H + 7) / 8), 256, 0, stream>>>(p, rgb, out8, out_rgb);
    return sky_check("grpg_sky_compose_rgb8");
}
