// optim.cu -- fused multi-tensor Adam step and densification statistics (include/grpg_optim.h).
//
// Behavioural spec: torch/optim/adam.py `_multi_tensor_adam` (non-capturable, amsgrad = False, weight_decay = 0), which
// is what the reference's per-sub-model `torch.optim.Adam(l, lr=0.0, eps=1e-15)` runs (gaussian_model.py:286-318), and
// StreetGaussianModel.set_max_radii2D / add_densification_stats (street_gaussian_model.py:555-578).
//
// B200 design.  Both ops are single-pass streaming: Adam reads 16 B and writes 12 B per parameter element (1.3 GB per
// iteration at 2 M Gaussians), so one launch walks ALL tensors of ALL sub-models: a CTA owns a 4096-element chunk of one
// tensor (looked up from a descriptor table by binary search), every thread keeps four independent 16-byte
// loads per array in flight, nothing is re-read.  torch's foreach path makes seven passes over the same bytes per
// optimiser and the reference has nine optimisers.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "grpg_common.cuh"
#include "../../include/grpg_optim.h"

extern "C" int grpg_loss_fail(const char* msg);

namespace grpg {

constexpr int ADAM_THREADS = 256;
constexpr int ADAM_VEC = 4;                                  // float4 groups per thread
constexpr int ADAM_CHUNK = ADAM_THREADS * ADAM_VEC * 4;      // 4096 elements per CTA
constexpr int ADAM_MAX_TENSORS = 1024;

struct AdamDev {
    float* p; const float* g; float* m; float* v;
    long long numel;
    int chunk_begin, vec_ok;  // vec_ok: all four pointers 16-byte aligned
    float w1, beta2, w2, step_size, bc2_sqrt, eps;
};
static_assert(sizeof(AdamDev) % 8 == 0, "table rows are read as 8-byte words");

// Row of a descriptor table that owns CTA `b`: the largest k with chunk_begin[k] <= b.  The column is monotone
// (empty tensors repeat their successor's value, so the largest k is never an empty one), hence a binary search;
// every thread runs the same ~10 broadcast loads.  (A __syncthreads_count over per-thread partial counts would
// under-count once a thread holds two matching rows, i.e. for tables longer than the block.)
template <typename Row>
__device__ __forceinline__ int owner_row(const Row* __restrict__ tab, int n, int b) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tab[mid].chunk_begin <= b) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamDev& d) {
    m = __fmaf_rn(d.w1, g - m, m);                  // lerp_(g, 1 - beta1): m + w (g - m)
    v = v * d.beta2;                                // mul_(beta2)
    v = __fmaf_rn(d.w2 * g, g, v);                  // addcmul_(g, g, value = 1 - beta2)
    const float denom = sqrtf(v) / d.bc2_sqrt + d.eps;
    p = __fmaf_rn(d.step_size, m / denom, p);       // addcdiv_(m, denom, value = step_size)
}

__global__ void __launch_bounds__(ADAM_THREADS) adam_step_kernel(const AdamDev* __restrict__ tab, int n) {
    __shared__ AdamDev d;
    const int k = owner_row(tab, n, (int)blockIdx.x);
    {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(tab + k);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(&d);
        for (int i = threadIdx.x; i < (int)(sizeof(AdamDev) / 8); i += ADAM_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    const long long e0 = (long long)((int)blockIdx.x - d.chunk_begin) * ADAM_CHUNK;
    const long long left = d.numel - e0;
    if (left <= 0) return;
    const int cnt = left < ADAM_CHUNK ? (int)left : ADAM_CHUNK;
    float* p = d.p + e0; const float* g = d.g + e0; float* m = d.m + e0; float* v = d.v + e0;
    const int tid = threadIdx.x;
    if (d.vec_ok) {  // chunk starts are multiples of 4096 elements: alignment of the chunk = alignment of the tensor
        const int G = cnt >> 2;
        float4 P[ADAM_VEC], Gd[ADAM_VEC], Mv[ADAM_VEC], V[ADAM_VEC];
#pragma unroll
        for (int u = 0; u < ADAM_VEC; ++u) {
            const int i = tid + u * ADAM_THREADS;
            if (i < G) {
                P[u] = reinterpret_cast<const float4*>(p)[i];
                Gd[u] = __ldcs(reinterpret_cast<const float4*>(g) + i);
                Mv[u] = reinterpret_cast<const float4*>(m)[i];
                V[u] = reinterpret_cast<const float4*>(v)[i];
            }
        }
#pragma unroll
        for (int u = 0; u < ADAM_VEC; ++u) {
            const int i = tid + u * ADAM_THREADS;
            if (i < G) {
                adam_update(P[u].x, Gd[u].x, Mv[u].x, V[u].x, d);
                adam_update(P[u].y, Gd[u].y, Mv[u].y, V[u].y, d);
                adam_update(P[u].z, Gd[u].z, Mv[u].z, V[u].z, d);
                adam_update(P[u].w, Gd[u].w, Mv[u].w, V[u].w, d);
                reinterpret_cast<float4*>(p)[i] = P[u];
                reinterpret_cast<float4*>(m)[i] = Mv[u];
                reinterpret_cast<float4*>(v)[i] = V[u];
            }
        }
        const int e = 4 * G + tid;  // < 4 trailing elements of the tensor
        if (e < cnt) {
            float pp = p[e], mm = m[e], vv = v[e];
            adam_update(pp, g[e], mm, vv, d);
            p[e] = pp; m[e] = mm; v[e] = vv;
        }
    } else {
        for (int e = tid; e < cnt; e += ADAM_THREADS) {
            float pp = p[e], mm = m[e], vv = v[e];
            adam_update(pp, g[e], mm, vv, d);
            p[e] = pp; m[e] = mm; v[e] = vv;
        }
    }
}

// ---- densification statistics ----------------------------------------------------------------------------------
constexpr int STATS_THREADS = 256;
constexpr int STATS_MAX_SUB = 1024;
struct StatsDev {
    float* max_radii; float* accum; float* denom;
    long long offset;
    int n, chunk_begin;
};
static_assert(sizeof(StatsDev) % 8 == 0, "");

__global__ void __launch_bounds__(STATS_THREADS) densify_stats_kernel(const StatsDev* __restrict__ tab, int n_sub,
                                                                      const int* __restrict__ radii,
                                                                      const float* __restrict__ vgrad,
                                                                      const unsigned char* __restrict__ filter,
                                                                      int what) {
    __shared__ StatsDev d;
    const int k = owner_row(tab, n_sub, (int)blockIdx.x);
    if (threadIdx.x == 0) d = tab[k];
    __syncthreads();
    const int j = ((int)blockIdx.x - d.chunk_begin) * STATS_THREADS + threadIdx.x;  // row inside the sub-model
    if (j >= d.n) return;
    const size_t i = (size_t)d.offset + j;                                           // row of the composed arrays
    // visibility_filter: the caller's boolean mask when one is passed, else radii > 0 (what the reference renderer
    // builds, street_gaussian_renderer.py:268)
    const int r = (filter == nullptr || (what & GRPG_STATS_MAX_RADII)) ? __ldcs(radii + i) : 0;
    if (filter != nullptr ? filter[i] == 0 : r <= 0) return;
    if (what & GRPG_STATS_MAX_RADII)
        d.max_radii[j] = fmaxf(d.max_radii[j], (float)r);   // street_gaussian_model.py:564-565 (radii.float())
    if (what & GRPG_STATS_GRADIENTS) {
        const float gx = __ldcs(vgrad + 3 * i), gy = __ldcs(vgrad + 3 * i + 1), gz = __ldcs(vgrad + 3 * i + 2);
        d.accum[2 * j] += sqrtf(gx * gx + gy * gy);         // :576  torch.norm(grad[:, :2], dim=-1)
        d.accum[2 * j + 1] += fabsf(gz);                    // :577  torch.norm(grad[:, 2:], dim=-1)
        d.denom[j] += 1.0f;                                 // :578
    }
}

static int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        static thread_local char buf[256];
        snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
        return grpg_loss_fail(buf);
    }
    return 0;
}

}  // namespace grpg

using namespace grpg;

extern "C" size_t grpg_adam_workspace_bytes(int n) { return align_up((size_t)(n > 0 ? n : 1) * sizeof(AdamDev), 256); }

extern "C" int grpg_adam_step(const grpg_adam_tensor* tensors, int n, void* workspace, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return 0;
    if (!tensors || n < 0 || n > ADAM_MAX_TENSORS) return grpg_loss_fail("grpg_adam_step: 1..1024 tensors per call");
    if (!workspace) return grpg_loss_fail("grpg_adam_step: null workspace");
    std::vector<AdamDev> tab(n);
    long long chunks = 0;
    for (int k = 0; k < n; ++k) {
        const grpg_adam_tensor& t = tensors[k];
        if (t.numel < 0 || (t.numel > 0 && (!t.param || !t.grad || !t.exp_avg || !t.exp_avg_sq)))
            return grpg_loss_fail("grpg_adam_step: null tensor pointer");
        AdamDev& d = tab[k];
        std::memset(&d, 0, sizeof(d));
        d.p = t.param; d.g = t.grad; d.m = t.exp_avg; d.v = t.exp_avg_sq; d.numel = t.numel;
        d.chunk_begin = (int)chunks;
        d.vec_ok = ((((uintptr_t)t.param | (uintptr_t)t.grad | (uintptr_t)t.exp_avg | (uintptr_t)t.exp_avg_sq) & 15) == 0) ? 1 : 0;
        d.w1 = t.one_minus_beta1; d.beta2 = t.beta2; d.w2 = t.one_minus_beta2; d.step_size = t.step_size;
        d.bc2_sqrt = t.bias_correction2_sqrt; d.eps = t.eps;
        chunks += (t.numel + ADAM_CHUNK - 1) / ADAM_CHUNK;
        if (chunks > 0x7fffffffLL) return grpg_loss_fail("grpg_adam_step: too many elements");
    }
    if (chunks == 0) return 0;
    if (cudaMemcpyAsync(workspace, tab.data(), tab.size() * sizeof(AdamDev), cudaMemcpyHostToDevice, stream) != cudaSuccess)
        return check_launch("grpg_adam_step: table upload");
    ProfScope ps("adam_step", stream);
    adam_step_kernel<<<(unsigned)chunks, ADAM_THREADS, 0, stream>>>(reinterpret_cast<const AdamDev*>(workspace), n);
    return check_launch("grpg_adam_step");
}

extern "C" size_t grpg_stats_workspace_bytes(int n_sub) {
    return align_up((size_t)(n_sub > 0 ? n_sub : 1) * sizeof(StatsDev), 256);
}

extern "C" int grpg_densify_stats(const grpg_stats_submodel* subs, int n_sub, const int* radii, const float* viewspace_grad,
                                  void* workspace, void* stream_) {
    return grpg_densify_stats_ex(subs, n_sub, radii, viewspace_grad, nullptr, GRPG_STATS_MAX_RADII | GRPG_STATS_GRADIENTS,
                                 workspace, stream_);
}

extern "C" int grpg_densify_stats_ex(const grpg_stats_submodel* subs, int n_sub, const int* radii,
                                     const float* viewspace_grad, const unsigned char* visibility_filter, int what,
                                     void* workspace, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_sub == 0) return 0;
    if ((what & ~(GRPG_STATS_MAX_RADII | GRPG_STATS_GRADIENTS)) || what == 0) return grpg_loss_fail("grpg_densify_stats_ex: bad `what` mask");
    if (!subs || n_sub < 0 || n_sub > STATS_MAX_SUB) return grpg_loss_fail("grpg_densify_stats: 1..1024 sub-models per call");
    std::vector<StatsDev> tab(n_sub);
    long long offset = 0, chunks = 0;
    for (int k = 0; k < n_sub; ++k) {
        const grpg_stats_submodel& s = subs[k];
        if (s.n < 0 || (s.n > 0 && (!s.max_radii2D || !s.xyz_gradient_accum || !s.denom)))
            return grpg_loss_fail("grpg_densify_stats: null statistics pointer");
        StatsDev& d = tab[k];
        std::memset(&d, 0, sizeof(d));
        d.max_radii = s.max_radii2D; d.accum = s.xyz_gradient_accum; d.denom = s.denom;
        d.offset = offset; d.n = s.n; d.chunk_begin = (int)chunks;
        offset += s.n;
        chunks += (s.n + STATS_THREADS - 1) / STATS_THREADS;
    }
    if (chunks == 0) return 0;
    if (!workspace || ((what & GRPG_STATS_GRADIENTS) && !viewspace_grad) ||
        (((what & GRPG_STATS_MAX_RADII) || !visibility_filter) && !radii))
        return grpg_loss_fail("grpg_densify_stats: null input or workspace");
    if (cudaMemcpyAsync(workspace, tab.data(), tab.size() * sizeof(StatsDev), cudaMemcpyHostToDevice, stream) != cudaSuccess)
        return check_launch("grpg_densify_stats: table upload");
    ProfScope ps("densify_stats", stream);
    densify_stats_kernel<<<(unsigned)chunks, STATS_THREADS, 0, stream>>>(reinterpret_cast<const StatsDev*>(workspace), n_sub,
                                                                        radii, viewspace_grad, visibility_filter, what);
    return check_launch("grpg_densify_stats");
}
