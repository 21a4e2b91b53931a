// simple_knn.cu -- mean squared distance to the three nearest neighbours of every point (include/grpg_knn.h).
// Behavioural spec: reference submodules/simple-knn/simple_knn.cu:140-219 (exact 3-NN; result independent of the search
// structure).  Search structure here: Morton order + three levels of axis-aligned boxes over that order.
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "radix_sort.cuh"
#include "../../include/grpg_knn.h"

extern "C" int grpg_loss_fail(const char* msg);

namespace grpg {

constexpr int KNN_LEAF = 128;   // points per leaf box
constexpr int KNN_L1 = 8;       // leaves per level-1 box  (1024 points, the reference's BOX_SIZE)
constexpr int KNN_L2 = 32;      // level-1 boxes per level-2 box (32768 points)

struct Box { float lo[3], hi[3]; };

// ---- bounding box of the cloud: per-CTA reduction + float atomics on the ordered-int image of the floats ---------
__device__ __forceinline__ int ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float unord(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void __launch_bounds__(256) knn_bounds_kernel(int P, const float* __restrict__ pts, int* __restrict__ bounds /*[6]*/) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float v = pts[3 * (size_t)i + a]; lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(bounds + a, ord(lo[a])); atomicMax(bounds + 3 + a, ord(hi[a])); }
    }
}

__device__ __forceinline__ uint32_t spread10(uint32_t x) {  // simple_knn.cu:45-52
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}

__global__ void __launch_bounds__(256) knn_morton_kernel(int P, const float* __restrict__ pts, const int* __restrict__ bounds,
                                                         uint32_t* __restrict__ codes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    uint32_t c = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {  // simple_knn.cu:54-61 (an empty extent gives NaN -> 0, a single cell on that axis)
        const float lo = unord(bounds[a]), hi = unord(bounds[3 + a]);
        const float t = ((pts[3 * (size_t)i + a] - lo) / (hi - lo)) * 1023.0f;
        const uint32_t q = t >= 0.0f ? min(1023u, (uint32_t)t) : 0u;
        c |= spread10(q) << a;
    }
    codes[i] = c;
}

// points gathered into Morton order as float4 (x, y, z, original index) + the leaf boxes
__global__ void __launch_bounds__(KNN_LEAF) knn_gather_leaf_kernel(int P, const float* __restrict__ pts,
                                                                   const uint32_t* __restrict__ order,
                                                                   float4* __restrict__ sorted, Box* __restrict__ leaf) {
    const int i = blockIdx.x * KNN_LEAF + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < P) {
        const uint32_t src = order[i];
        const float x = pts[3 * (size_t)src], y = pts[3 * (size_t)src + 1], z = pts[3 * (size_t)src + 2];
        sorted[i] = make_float4(x, y, z, __uint_as_float(src));
        lo[0] = hi[0] = x; lo[1] = hi[1] = y; lo[2] = hi[2] = z;
    }
    __shared__ float s[KNN_LEAF / 32][6];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) { s[threadIdx.x >> 5][a] = lo[a]; s[threadIdx.x >> 5][3 + a] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float l = FLT_MAX, h = -FLT_MAX;
        for (int w = 0; w < KNN_LEAF / 32; ++w) { l = fminf(l, s[w][threadIdx.x]); h = fmaxf(h, s[w][3 + threadIdx.x]); }
        leaf[blockIdx.x].lo[threadIdx.x] = l;
        leaf[blockIdx.x].hi[threadIdx.x] = h;
    }
}

// parent boxes: union of `fan` consecutive children
__global__ void __launch_bounds__(256) knn_merge_boxes_kernel(int n_child, int fan, const Box* __restrict__ child,
                                                              Box* __restrict__ parent) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p * fan >= n_child) return;
    Box b;
#pragma unroll
    for (int a = 0; a < 3; ++a) { b.lo[a] = FLT_MAX; b.hi[a] = -FLT_MAX; }
    for (int c = p * fan; c < min(n_child, (p + 1) * fan); ++c) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { b.lo[a] = fminf(b.lo[a], child[c].lo[a]); b.hi[a] = fmaxf(b.hi[a], child[c].hi[a]); }
    }
    parent[p] = b;
}

// squared distance from a point to a box (0 inside), simple_knn.cu:115-125
__device__ __forceinline__ float box_dist2(const Box& b, float x, float y, float z) {
    float dx = 0.f, dy = 0.f, dz = 0.f;
    if (x < b.lo[0] || x > b.hi[0]) dx = fminf(fabsf(x - b.lo[0]), fabsf(x - b.hi[0]));
    if (y < b.lo[1] || y > b.hi[1]) dy = fminf(fabsf(y - b.lo[1]), fabsf(y - b.hi[1]));
    if (z < b.lo[2] || z > b.hi[2]) dz = fminf(fabsf(z - b.lo[2]), fabsf(z - b.hi[2]));
    return dx * dx + dy * dy + dz * dz;
}

// insert into the ascending triple (simple_knn.cu:127-141); the distance is the reference build's contraction
__device__ __forceinline__ void consider(float x, float y, float z, const float4 q, float (&best)[3]) {
    const float dx = __fadd_rn(q.x, -x), dy = __fadd_rn(q.y, -y), dz = __fadd_rn(q.z, -z);
    float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));  // a*b + c*d contracts to fma(a, b, c*d)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (best[j] > d) { const float t = best[j]; best[j] = d; d = t; }
    }
}

__global__ void __launch_bounds__(128) knn_search_kernel(int P, const float4* __restrict__ sorted, const Box* __restrict__ leaf,
                                                         const Box* __restrict__ l1, const Box* __restrict__ l2, int n_leaf,
                                                         int n_l1, int n_l2, float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float4 me = sorted[idx];
    const float x = me.x, y = me.y, z = me.z;
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    // rejection radius from the six neighbours in Morton order (simple_knn.cu:151-160)
    for (int i = max(0, idx - 3); i <= min(P - 1, idx + 3); ++i)
        if (i != idx) consider(x, y, z, sorted[i], best);
    const float reject = best[2];
    best[0] = best[1] = best[2] = FLT_MAX;
    // every box whose distance does not exceed the rejection radius (nor the current third-best) is searched; a
    // parent's distance never exceeds its children's, so pruning a parent prunes nothing the flat loop would visit
    for (int c2 = 0; c2 < n_l2; ++c2) {
        const float d2 = box_dist2(l2[c2], x, y, z);
        if (d2 > reject || d2 > best[2]) continue;
        for (int c1 = c2 * KNN_L2; c1 < min(n_l1, (c2 + 1) * KNN_L2); ++c1) {
            const float d1 = box_dist2(l1[c1], x, y, z);
            if (d1 > reject || d1 > best[2]) continue;
            for (int lf = c1 * KNN_L1; lf < min(n_leaf, (c1 + 1) * KNN_L1); ++lf) {
                const float d0 = box_dist2(leaf[lf], x, y, z);
                if (d0 > reject || d0 > best[2]) continue;
                const int lo = lf * KNN_LEAF, hi = min(P, (lf + 1) * KNN_LEAF);
                for (int i = lo; i < hi; ++i)
                    if (i != idx) consider(x, y, z, sorted[i], best);
            }
        }
    }
    out[__float_as_uint(me.w)] = __fdiv_rn(__fadd_rn(__fadd_rn(best[0], best[1]), best[2]), 3.0f);  // :183
}

static size_t knn_layout(int P, size_t* o_codes, size_t* o_order, size_t* o_codes_b, size_t* o_order_b, size_t* o_aux,
                         size_t* o_sorted, size_t* o_leaf, size_t* o_l1, size_t* o_l2, size_t* o_bounds) {
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += align_up(bytes, 256); return at; };
    const size_t p = (size_t)(P > 0 ? P : 1);
    const size_t n_leaf = (p + KNN_LEAF - 1) / KNN_LEAF, n_l1 = (n_leaf + KNN_L1 - 1) / KNN_L1, n_l2 = (n_l1 + KNN_L2 - 1) / KNN_L2;
    *o_codes = take(p * 4); *o_order = take(p * 4); *o_codes_b = take(p * 4); *o_order_b = take(p * 4);
    *o_aux = take((size_t)SORT_MAX_PASSES * (256 + 64) * 4 + (size_t)SORT_MAX_PASSES * sort_num_tiles((long long)p) * 256 * 4);
    *o_sorted = take(p * 16);
    *o_leaf = take(n_leaf * sizeof(Box)); *o_l1 = take(n_l1 * sizeof(Box)); *o_l2 = take(n_l2 * sizeof(Box));
    *o_bounds = take(64);
    return o;
}

}  // namespace grpg

using namespace grpg;

extern "C" size_t grpg_knn_workspace_bytes(int P) {
    size_t a, b, c, d, e, f, g, h, i, j;
    return knn_layout(P, &a, &b, &c, &d, &e, &f, &g, &h, &i, &j);
}

extern "C" int grpg_knn_mean_dist2(int P, const float* points, float* mean_dist2, void* workspace, void* stream_) {
    if (P == 0) return 0;
    if (P < 0 || !points || !mean_dist2 || !workspace) return grpg_loss_fail("grpg_knn_mean_dist2: bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    size_t o_codes, o_order, o_codes_b, o_order_b, o_aux, o_sorted, o_leaf, o_l1, o_l2, o_bounds;
    knn_layout(P, &o_codes, &o_order, &o_codes_b, &o_order_b, &o_aux, &o_sorted, &o_leaf, &o_l1, &o_l2, &o_bounds);
    char* w = (char*)workspace;
    uint32_t *codes = (uint32_t*)(w + o_codes), *order = (uint32_t*)(w + o_order);
    uint32_t *codes_b = (uint32_t*)(w + o_codes_b), *order_b = (uint32_t*)(w + o_order_b);
    int* bounds = (int*)(w + o_bounds);
    const int init[6] = {INT32_MAX, INT32_MAX, INT32_MAX, INT32_MIN, INT32_MIN, INT32_MIN};
    // (a 24-byte constant: cudaMemcpyAsync from pageable memory copies it before returning)
    cudaMemcpyAsync(bounds, init, sizeof init, cudaMemcpyHostToDevice, stream);
    {
        ProfScope ps("knn_bounds", stream);
        knn_bounds_kernel<<<min(148 * 8, (P + 255) / 256), 256, 0, stream>>>(P, points, bounds);
    }
    {
        ProfScope ps("knn_morton", stream);
        knn_morton_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, points, bounds, codes);
    }
    const bool in_a = onesweep_sort_pairs(codes, order, codes_b, order_b, P, 30, (uint32_t*)(w + o_aux), 148, stream,
                                          "knn_sort_hist", "knn_sort_scan", "knn_sort_pass", /*iota_values=*/true);
    const uint32_t* ord_sorted = in_a ? order : order_b;
    const int n_leaf = (P + KNN_LEAF - 1) / KNN_LEAF, n_l1 = (n_leaf + KNN_L1 - 1) / KNN_L1, n_l2 = (n_l1 + KNN_L2 - 1) / KNN_L2;
    float4* sorted = (float4*)(w + o_sorted);
    Box *leaf = (Box*)(w + o_leaf), *l1 = (Box*)(w + o_l1), *l2 = (Box*)(w + o_l2);
    {
        ProfScope ps("knn_boxes", stream);
        knn_gather_leaf_kernel<<<n_leaf, KNN_LEAF, 0, stream>>>(P, points, ord_sorted, sorted, leaf);
        knn_merge_boxes_kernel<<<(n_l1 + 255) / 256, 256, 0, stream>>>(n_leaf, KNN_L1, leaf, l1);
        knn_merge_boxes_kernel<<<(n_l2 + 255) / 256, 256, 0, stream>>>(n_l1, KNN_L2, l1, l2);
    }
    ProfScope ps("knn_search", stream);
    knn_search_kernel<<<(P + 127) / 128, 128, 0, stream>>>(P, sorted, leaf, l1, l2, n_leaf, n_l1, n_l2, mean_dist2);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return grpg_loss_fail(cudaGetErrorString(e));
    return 0;
}
