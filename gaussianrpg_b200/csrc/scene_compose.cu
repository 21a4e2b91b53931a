// scene_compose.cu -- fused scene-graph compose: sub-model parameters -> the rasterizer's [P, .] inputs, and back.
//
// Behavioural spec (see include/grpg_compose.h): StreetGaussianModel.get_xyz / get_rotation / get_scaling /
// get_opacity / get_features (lib/models/street_gaussian_model.py:295-384,438-453), the sub-model activations
// (gaussian_model.py:214-251), the actors' Fourier dc (gaussian_model_actor.py:73-82), quaternion_raw_multiply and
// quaternion_to_matrix (lib/utils/general_utils.py:220-238,125-146).
//
// B200 design.  The op is pure streaming (184 B read+written per background Gaussian at M = 4), so the only things
// that matter are coalescing and the number of passes.  One CTA owns a chunk of 1024 Gaussians of ONE sub-model and
// walks each attribute as a FLAT float range (scaling, opacity, SH rows, background xyz are elementwise or pure row
// re-packing), so every warp access is a contiguous 128-byte line whatever the row length (3, 9, 3M floats);
// quaternions move as one 16-byte vector per row.  Everything the reference materialises in between (clones,
// normalised copies, expanded per-Gaussian pose rows, rotation matrices, the concatenations) never exists.  The
// backward has the same shape; the per-actor pose gradient (dR, dt, dq: 16 floats) is reduced per CTA and added
// with 16 atomics, a one-block kernel finishes the chain through quaternion_to_matrix.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "grpg_common.cuh"
#include "../../include/grpg_compose.h"

extern "C" int grpg_loss_fail(const char* msg);

namespace grpg {

constexpr int CMP_CHUNK = 256;  // rows of ONE sub-model per CTA (one row per thread in the quaternion stage)
constexpr int CMP_THREADS = 256;
constexpr float CMP_EPS = 1e-12f;  // torch.nn.functional.normalize default

struct SubDev {
    const float *xyz, *scaling, *rotation, *opacity, *dc, *rest;
    const unsigned char* flip;
    const float *obj_rot, *obj_trans;  // device: 4 and 3 floats (actors)
    float *d_xyz, *d_scaling, *d_rotation, *d_opacity, *d_dc, *d_rest;
    long long offset;  // first output row of this sub-model
    int n, is_actor, F, chunk_begin;
    int aligned16, pad_;  // every parameter (and gradient) pointer is 16-byte aligned: bulk staging allowed
    float idft[GRPG_COMPOSE_MAX_FOURIER];
};
static_assert(sizeof(SubDev) % 8 == 0, "descriptor table is copied as 8-byte words");

struct Quat { float w, x, y, z; };

__device__ __forceinline__ Quat qmul(const Quat a, const Quat b) {  // general_utils.py:232-238
    Quat o;
    o.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    o.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    o.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    o.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    return o;
}
// product with the flip quaternion (0, 0, 1, 0) = matrix_to_quaternion(diag(-1, 1, -1)) on the left
__device__ __forceinline__ Quat qflip(const Quat b) { return Quat{-b.y, b.z, b.w, -b.x}; }
// cotangent of b for o = (0,0,1,0) * b
__device__ __forceinline__ Quat qflip_bwd(const Quat g) { return Quat{g.y, -g.z, -g.w, g.x}; }

__device__ __forceinline__ Quat qnormalize(const Quat v) {  // F.normalize: v / max(|v|, eps)
    const float n = fmaxf(sqrtf(v.w * v.w + v.x * v.x + v.y * v.y + v.z * v.z), CMP_EPS);
    return Quat{v.w / n, v.x / n, v.y / n, v.z / n};
}
__device__ __forceinline__ Quat qnormalize_bwd(const Quat v, const Quat g) {
    const float nrm = sqrtf(v.w * v.w + v.x * v.x + v.y * v.y + v.z * v.z);
    if (!(nrm > CMP_EPS)) return Quat{g.w / CMP_EPS, g.x / CMP_EPS, g.y / CMP_EPS, g.z / CMP_EPS};
    const float inv = 1.0f / nrm;
    const Quat y{v.w * inv, v.x * inv, v.y * inv, v.z * inv};
    const float d = y.w * g.w + y.x * g.x + y.y * g.y + y.z * g.z;
    return Quat{(g.w - y.w * d) * inv, (g.x - y.x * d) * inv, (g.y - y.y * d) * inv, (g.z - y.z * d) * inv};
}

struct Pose {  // per-CTA copy of the actor pose
    float q[4];  // obj_rot as given (the rotation product uses it un-normalised)
    float R[9];  // quaternion_to_matrix(obj_rot) (general_utils.py:125-146), row-major
    float t[3];
};

// Finds the sub-model that owns this CTA's chunk (one parallel look at the chunk_begin column: sub-model k owns
// chunks [chunk_begin[k], chunk_begin[k+1]), empty sub-models own none), copies its descriptor into shared memory and
// derives the pose constants.
__device__ __forceinline__ int load_desc(SubDev* dst, Pose* pose, const SubDev* __restrict__ subs, int n_sub) {
    int c = 0;
    for (int k = threadIdx.x; k < n_sub; k += blockDim.x) c += subs[k].chunk_begin <= (int)blockIdx.x ? 1 : 0;
    const int k = __syncthreads_count(c) - 1;  // n_sub <= blockDim.x: at most one candidate per thread
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(subs + k);
    unsigned long long* d = reinterpret_cast<unsigned long long*>(dst);
    for (int i = threadIdx.x; i < (int)(sizeof(SubDev) / 8); i += blockDim.x) d[i] = src[i];
    __syncthreads();
    if (threadIdx.x == 0 && dst->is_actor) {
        const float q0 = dst->obj_rot[0], q1 = dst->obj_rot[1], q2 = dst->obj_rot[2], q3 = dst->obj_rot[3];
        const float norm = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
        const float r = q0 / norm, x = q1 / norm, y = q2 / norm, z = q3 / norm;
        pose->q[0] = q0; pose->q[1] = q1; pose->q[2] = q2; pose->q[3] = q3;
        pose->R[0] = 1.f - 2.f * (y * y + z * z); pose->R[1] = 2.f * (x * y - r * z); pose->R[2] = 2.f * (x * z + r * y);
        pose->R[3] = 2.f * (x * y + r * z); pose->R[4] = 1.f - 2.f * (x * x + z * z); pose->R[5] = 2.f * (y * z - r * x);
        pose->R[6] = 2.f * (x * z - r * y); pose->R[7] = 2.f * (y * z + r * x); pose->R[8] = 1.f - 2.f * (x * x + y * y);
        pose->t[0] = dst->obj_trans[0]; pose->t[1] = dst->obj_trans[1]; pose->t[2] = dst->obj_trans[2];
    }
    __syncthreads();
    return k;
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

// Flat element loop with U independent loads in flight per thread before the first store: the kernels are pure
// streaming, so bytes in flight per SM (not instruction count) set the bandwidth.
constexpr int CMP_U = 8;
template <typename Load, typename Store>
__device__ __forceinline__ void flat_loop(int n, Load ld, Store st) {
    for (int j0 = threadIdx.x; j0 < n; j0 += CMP_U * CMP_THREADS) {
        decltype(ld(0)) v[CMP_U];
#pragma unroll
        for (int u = 0; u < CMP_U; ++u) {
            const int j = j0 + u * CMP_THREADS;
            if (j < n) v[u] = ld(j);
        }
#pragma unroll
        for (int u = 0; u < CMP_U; ++u) {
            const int j = j0 + u * CMP_THREADS;
            if (j < n) st(j, v[u]);
        }
    }
}


// ---- staged fast path ------------------------------------------------------------------------------------------
// When the chunk's rows start at a multiple of 4 on both sides (sub-model offset % 4 == 0 -- always true for the
// background, 92 % of a street scene) every input range of the chunk is a 16-byte-aligned contiguous span: one
// thread hands them to the TMA engine (cp.async.bulk + mbarrier), so the whole chunk (23-70 KB) is in flight at once
// without costing registers or issue slots, and the outputs leave as aligned 16-byte stores computed from shared
// memory.  Other chunks (actors at odd offsets, misaligned views) take the flat-loop path above.
struct Stager {
    uint64_t* bar;
    uint32_t bytes;
    __device__ __forceinline__ void add(float* dst, const float* src, int floats) {
        if (floats > 0) { tma_bulk_g2s(dst, src, (uint32_t)floats * 4u, bar); }
    }
};
// rows [rows_a, cnt) (< 4 rows: a bulk copy needs multiples of 16 bytes) are fetched with plain loads
__device__ __forceinline__ void stage_tail(float* dst, const float* src, int rows_a, int cnt, int row_floats) {
    const int n = (cnt - rows_a) * row_floats;
    for (int j = threadIdx.x; j < n; j += CMP_THREADS) dst[rows_a * row_floats + j] = src[rows_a * row_floats + j];
}
// n_el floats to a 16-byte-aligned destination: aligned float4 groups + a scalar tail (< 4 elements)
template <typename F4, typename F1>
__device__ __forceinline__ void emit_vec4(float* __restrict__ out, int n_el, F4 group, F1 elem) {
    const int G = n_el >> 2;
    for (int g = threadIdx.x; g < G; g += CMP_THREADS) reinterpret_cast<float4*>(out)[g] = group(g);
    const int e = 4 * G + threadIdx.x;
    if (e < n_el) out[e] = elem(e);
}
__device__ __forceinline__ float4 ld4(const float* s, int g) { return reinterpret_cast<const float4*>(s)[g]; }

__global__ void __launch_bounds__(CMP_THREADS) compose_fwd_kernel(const SubDev* __restrict__ subs, int n_sub, int M,
                                                                  float* __restrict__ xyz, float* __restrict__ rot,
                                                                  float* __restrict__ scl, float* __restrict__ opa,
                                                                  float* __restrict__ feat, int allow_staged) {
    extern __shared__ __align__(16) float s_dyn[];
    __shared__ SubDev sd;
    __shared__ Pose pose;
    __shared__ uint64_t s_bar;
    load_desc(&sd, &pose, subs, n_sub);
    const int r0 = ((int)blockIdx.x - sd.chunk_begin) * CMP_CHUNK;
    const int cnt = min(CMP_CHUNK, sd.n - r0);
    if (cnt <= 0) return;
    const size_t ob = (size_t)sd.offset + r0;  // first output row of this chunk
    const size_t ib = (size_t)r0;
    const bool actor = sd.is_actor != 0;
    const unsigned char* fl = (actor && sd.flip) ? sd.flip + ib : nullptr;

    if (allow_staged && sd.aligned16 && (sd.offset & 3) == 0) {
        const int tid = threadIdx.x;
        const int rowlen = 3 * M, restlen = rowlen - 3, F = sd.F;
        float* s_xyz = s_dyn;
        float* s_scl = s_xyz + CMP_CHUNK * 3;
        float* s_rot = s_scl + CMP_CHUNK * 3;
        float* s_opa = s_rot + CMP_CHUNK * 4;
        float* s_dc = s_opa + CMP_CHUNK;
        float* s_rest = s_dc + CMP_CHUNK * 3 * F;
        const int ra = cnt & ~3;  // rows that move as bulk copies
        if (tid == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&s_bar, (uint32_t)ra * 4u * (uint32_t)(3 + 3 + 4 + 1 + 3 * F + restlen));
            Stager st{&s_bar, 0};
            st.add(s_xyz, sd.xyz + ib * 3, ra * 3);
            st.add(s_scl, sd.scaling + ib * 3, ra * 3);
            st.add(s_rot, sd.rotation + ib * 4, ra * 4);
            st.add(s_opa, sd.opacity + ib, ra);
            st.add(s_dc, sd.dc + ib * 3 * F, ra * 3 * F);
            st.add(s_rest, sd.rest + ib * restlen, ra * restlen);
        }
        stage_tail(s_xyz, sd.xyz + ib * 3, ra, cnt, 3);
        stage_tail(s_scl, sd.scaling + ib * 3, ra, cnt, 3);
        stage_tail(s_rot, sd.rotation + ib * 4, ra, cnt, 4);
        stage_tail(s_opa, sd.opacity + ib, ra, cnt, 1);
        stage_tail(s_dc, sd.dc + ib * 3 * F, ra, cnt, 3 * F);
        if (restlen > 0) stage_tail(s_rest, sd.rest + ib * restlen, ra, cnt, restlen);
        if (ra > 0) mbar_wait(&s_bar, 0);
        __syncthreads();

        if (tid < cnt) {  // rotations: one row per thread
            const float4 v = ld4(s_rot, tid);
            Quat q = qnormalize(Quat{v.x, v.y, v.z, v.w});
            if (actor) {
                if (fl && fl[tid]) q = qflip(q);
                q = qnormalize(qmul(Quat{pose.q[0], pose.q[1], pose.q[2], pose.q[3]}, q));
            }
            reinterpret_cast<float4*>(rot)[ob + tid] = make_float4(q.w, q.x, q.y, q.z);
        }
        emit_vec4(scl + ob * 3, cnt * 3,
                  [&](int g) { const float4 v = ld4(s_scl, g); return make_float4(expf(v.x), expf(v.y), expf(v.z), expf(v.w)); },
                  [&](int e) { return expf(s_scl[e]); });
        emit_vec4(opa + ob, cnt,
                  [&](int g) { const float4 v = ld4(s_opa, g); return make_float4(sigmoidf(v.x), sigmoidf(v.y), sigmoidf(v.z), sigmoidf(v.w)); },
                  [&](int e) { return sigmoidf(s_opa[e]); });
        const bool plain_dc = F == 1 && !actor;
        auto feat_elem = [&](int row, int c) {
            if (c >= 3) return s_rest[row * restlen + (c - 3)];
            if (plain_dc) return s_dc[row * 3 + c];
            float v = 0.f;
            for (int f = 0; f < F; ++f) v += s_dc[(row * F + f) * 3 + c] * sd.idft[f];
            return v;
        };
        emit_vec4(feat + ob * rowlen, cnt * rowlen,
                  [&](int g) {
                      int row = (4 * g) / rowlen, c = 4 * g - row * rowlen;
                      float v[4];
#pragma unroll
                      for (int u = 0; u < 4; ++u) {
                          v[u] = feat_elem(row, c);
                          if (++c == rowlen) { c = 0; ++row; }
                      }
                      return make_float4(v[0], v[1], v[2], v[3]);
                  },
                  [&](int e) { const int row = e / rowlen; return feat_elem(row, e - row * rowlen); });
        if (!actor) {
            emit_vec4(xyz + ob * 3, cnt * 3, [&](int g) { return ld4(s_xyz, g); }, [&](int e) { return s_xyz[e]; });
        } else {
            auto xyz_elem = [&](int e) {
                const int row = e / 3, c = e - row * 3;
                const float x0 = s_xyz[row * 3], x2 = s_xyz[row * 3 + 2];
                float x1 = s_xyz[row * 3 + 1];
                if (fl && fl[row]) x1 = -x1;  // flip_axis = 1
                return pose.R[c * 3] * x0 + pose.R[c * 3 + 1] * x1 + pose.R[c * 3 + 2] * x2 + pose.t[c];
            };
            emit_vec4(xyz + ob * 3, cnt * 3,
                      [&](int g) { return make_float4(xyz_elem(4 * g), xyz_elem(4 * g + 1), xyz_elem(4 * g + 2), xyz_elem(4 * g + 3)); },
                      xyz_elem);
        }
        return;
    }

    {  // rotations: one 16-byte vector per row
        const float4* s = reinterpret_cast<const float4*>(sd.rotation) + ib;
        float4* o = reinterpret_cast<float4*>(rot) + ob;
        const Quat qo{pose.q[0], pose.q[1], pose.q[2], pose.q[3]};
        flat_loop(cnt, [&](int row) { return __ldcs(s + row); },
                  [&](int row, float4 v) {
                      Quat q = qnormalize(Quat{v.x, v.y, v.z, v.w});
                      if (actor) {
                          if (fl && fl[row]) q = qflip(q);
                          q = qnormalize(qmul(qo, q));
                      }
                      o[row] = make_float4(q.w, q.x, q.y, q.z);
                  });
    }
    {  // scaling: exp, flat
        const float* s = sd.scaling + ib * 3;
        float* o = scl + ob * 3;
        flat_loop(cnt * 3, [&](int j) { return __ldcs(s + j); }, [&](int j, float v) { o[j] = expf(v); });
    }
    {  // opacity: sigmoid, flat
        const float* s = sd.opacity + ib;
        float* o = opa + ob;
        flat_loop(cnt, [&](int j) { return __ldcs(s + j); }, [&](int j, float v) { o[j] = sigmoidf(v); });
    }
    {  // SH rows: [dc | rest] re-packed; the actors' dc is the inverse DFT over their fourier_dim rows
        const int rowlen = 3 * M, restlen = rowlen - 3, F = sd.F;
        const float* dc = sd.dc + ib * 3 * F;
        const float* rest = sd.rest + ib * restlen;
        float* o = feat + ob * rowlen;
        const bool plain_dc = F == 1 && !actor;
        flat_loop(cnt * rowlen,
                  [&](int j) {
                      const int row = j / rowlen, c = j - row * rowlen;
                      if (c >= 3) return __ldcs(rest + (size_t)row * restlen + (c - 3));
                      if (plain_dc) return __ldcs(dc + row * 3 + c);
                      float v = 0.f;
                      for (int f = 0; f < F; ++f) v += __ldcs(dc + ((size_t)row * F + f) * 3 + c) * sd.idft[f];
                      return v;
                  },
                  [&](int j, float v) { o[j] = v; });
    }
    {  // means
        const float* s = sd.xyz + ib * 3;
        float* o = xyz + ob * 3;
        if (!actor) {
            flat_loop(cnt * 3, [&](int j) { return __ldcs(s + j); }, [&](int j, float v) { o[j] = v; });
        } else {
            flat_loop(cnt * 3,
                      [&](int j) {
                          const int row = j / 3, c = j - row * 3;
                          const float x0 = s[row * 3], x2 = s[row * 3 + 2];
                          float x1 = s[row * 3 + 1];
                          if (fl && fl[row]) x1 = -x1;  // flip_axis = 1
                          return pose.R[c * 3] * x0 + pose.R[c * 3 + 1] * x1 + pose.R[c * 3 + 2] * x2 + pose.t[c];
                      },
                      [&](int j, float v) { o[j] = v; });
        }
    }
}

__global__ void __launch_bounds__(CMP_THREADS) compose_bwd_kernel(
    const SubDev* __restrict__ subs, int n_sub, int M, const float* __restrict__ g_xyz, const float* __restrict__ g_rot,
    const float* __restrict__ g_scl, const float* __restrict__ g_opa, const float* __restrict__ g_feat,
    float* __restrict__ pose_acc /*[n_sub][16]*/, int allow_staged) {
    extern __shared__ __align__(16) float s_dyn[];
    __shared__ SubDev sd;
    __shared__ Pose pose;
    __shared__ float s_red[CMP_THREADS / 32][16];
    __shared__ uint64_t s_bar;
    const int k_sub = load_desc(&sd, &pose, subs, n_sub);
    const int tid = threadIdx.x;
    const int r0 = ((int)blockIdx.x - sd.chunk_begin) * CMP_CHUNK;
    const int cnt = min(CMP_CHUNK, sd.n - r0);
    if (cnt <= 0) return;
    const size_t ob = (size_t)sd.offset + r0;
    const size_t ib = (size_t)r0;
    const bool actor = sd.is_actor != 0;
    const unsigned char* fl = (actor && sd.flip) ? sd.flip + ib : nullptr;
    const bool staged = allow_staged && sd.aligned16 && (sd.offset & 3) == 0;

    // shared-memory views of the staged path (parameters, then cotangents)
    const int rowlen_ = 3 * M;
    float* s_scl = s_dyn;
    float* s_opa = s_scl + CMP_CHUNK * 3;
    float* s_rot = s_opa + CMP_CHUNK;
    float* s_xyz = s_rot + CMP_CHUNK * 4;
    float* s_gx = s_xyz + CMP_CHUNK * 3;
    float* s_gr = s_gx + CMP_CHUNK * 3;
    float* s_gs = s_gr + CMP_CHUNK * 4;
    float* s_go = s_gs + CMP_CHUNK * 3;
    float* s_gf = s_go + CMP_CHUNK;
    if (staged) {
        const int ra = cnt & ~3;
        if (tid == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&s_bar, (uint32_t)ra * 4u * (uint32_t)(3 + 1 + 4 + (actor ? 3 : 0) + 3 + 4 + 3 + 1 + rowlen_));
            Stager st{&s_bar, 0};
            st.add(s_scl, sd.scaling + ib * 3, ra * 3);
            st.add(s_opa, sd.opacity + ib, ra);
            st.add(s_rot, sd.rotation + ib * 4, ra * 4);
            if (actor) st.add(s_xyz, sd.xyz + ib * 3, ra * 3);
            st.add(s_gx, g_xyz + ob * 3, ra * 3);
            st.add(s_gr, g_rot + ob * 4, ra * 4);
            st.add(s_gs, g_scl + ob * 3, ra * 3);
            st.add(s_go, g_opa + ob, ra);
            st.add(s_gf, g_feat + ob * rowlen_, ra * rowlen_);
        }
        stage_tail(s_scl, sd.scaling + ib * 3, ra, cnt, 3);
        stage_tail(s_opa, sd.opacity + ib, ra, cnt, 1);
        stage_tail(s_rot, sd.rotation + ib * 4, ra, cnt, 4);
        if (actor) stage_tail(s_xyz, sd.xyz + ib * 3, ra, cnt, 3);
        stage_tail(s_gx, g_xyz + ob * 3, ra, cnt, 3);
        stage_tail(s_gr, g_rot + ob * 4, ra, cnt, 4);
        stage_tail(s_gs, g_scl + ob * 3, ra, cnt, 3);
        stage_tail(s_go, g_opa + ob, ra, cnt, 1);
        stage_tail(s_gf, g_feat + ob * rowlen_, ra, cnt, rowlen_);
        if (ra > 0) mbar_wait(&s_bar, 0);
        __syncthreads();

        emit_vec4(sd.d_scaling + ib * 3, cnt * 3,
                  [&](int g) {
                      const float4 v = ld4(s_scl, g), c = ld4(s_gs, g);
                      return make_float4(c.x * expf(v.x), c.y * expf(v.y), c.z * expf(v.z), c.w * expf(v.w));
                  },
                  [&](int e) { return s_gs[e] * expf(s_scl[e]); });
        auto dsig = [](float x, float g) { const float sg = sigmoidf(x); return g * (sg * (1.0f - sg)); };
        emit_vec4(sd.d_opacity + ib, cnt,
                  [&](int g) {
                      const float4 v = ld4(s_opa, g), c = ld4(s_go, g);
                      return make_float4(dsig(v.x, c.x), dsig(v.y, c.y), dsig(v.z, c.z), dsig(v.w, c.w));
                  },
                  [&](int e) { return dsig(s_opa[e], s_go[e]); });
        const int restlen = rowlen_ - 3, F = sd.F, dclen = 3 * F;
        if (restlen > 0) {
            auto rest_elem = [&](int row, int c) { return s_gf[row * rowlen_ + 3 + c]; };
            emit_vec4(sd.d_rest + ib * restlen, cnt * restlen,
                      [&](int g) {
                          int row = (4 * g) / restlen, c = 4 * g - row * restlen;
                          float v[4];
#pragma unroll
                          for (int u = 0; u < 4; ++u) {
                              v[u] = rest_elem(row, c);
                              if (++c == restlen) { c = 0; ++row; }
                          }
                          return make_float4(v[0], v[1], v[2], v[3]);
                      },
                      [&](int e) { const int row = e / restlen; return rest_elem(row, e - row * restlen); });
        }
        const bool plain_dc = F == 1 && !actor;
        auto dc_elem = [&](int e) {
            const int row = e / dclen, rem = e - row * dclen, f = rem / 3, c = rem - f * 3;
            const float gv = s_gf[row * rowlen_ + c];
            return plain_dc ? gv : gv * sd.idft[f];
        };
        emit_vec4(sd.d_dc + ib * dclen, cnt * dclen,
                  [&](int g) { return make_float4(dc_elem(4 * g), dc_elem(4 * g + 1), dc_elem(4 * g + 2), dc_elem(4 * g + 3)); },
                  dc_elem);
        if (!actor) {
            emit_vec4(sd.d_xyz + ib * 3, cnt * 3, [&](int g) { return ld4(s_gx, g); }, [&](int e) { return s_gx[e]; });
            if (tid < cnt) {
                const float4 v = ld4(s_rot, tid), gv = ld4(s_gr, tid);
                const Quat d = qnormalize_bwd(Quat{v.x, v.y, v.z, v.w}, Quat{gv.x, gv.y, gv.z, gv.w});
                reinterpret_cast<float4*>(sd.d_rotation)[ib + tid] = make_float4(d.w, d.x, d.y, d.z);
            }
            return;
        }
        auto dxyz_elem = [&](int e) {
            const int row = e / 3, c = e - row * 3;  // R^T g
            float d = pose.R[c] * s_gx[row * 3] + pose.R[3 + c] * s_gx[row * 3 + 1] + pose.R[6 + c] * s_gx[row * 3 + 2];
            if (c == 1 && fl && fl[row]) d = -d;
            return d;
        };
        emit_vec4(sd.d_xyz + ib * 3, cnt * 3,
                  [&](int g) { return make_float4(dxyz_elem(4 * g), dxyz_elem(4 * g + 1), dxyz_elem(4 * g + 2), dxyz_elem(4 * g + 3)); },
                  dxyz_elem);
        // the quaternion rows and the pose partials continue below, reading the staged copies
    }

    if (!staged) {
    {  // d log-scale = g * exp(s)
        const float* s = sd.scaling + ib * 3;
        const float* g = g_scl + ob * 3;
        float* o = sd.d_scaling + ib * 3;
        flat_loop(cnt * 3, [&](int j) { return make_float2(__ldcs(g + j), __ldcs(s + j)); },
                  [&](int j, float2 v) { o[j] = v.x * expf(v.y); });
    }
    {  // d logit = g * s (1 - s)
        const float* s = sd.opacity + ib;
        const float* g = g_opa + ob;
        float* o = sd.d_opacity + ib;
        flat_loop(cnt, [&](int j) { return make_float2(__ldcs(g + j), __ldcs(s + j)); },
                  [&](int j, float2 v) {
                      const float sg = sigmoidf(v.y);
                      o[j] = v.x * (sg * (1.0f - sg));
                  });
    }
    {  // SH
        const int rowlen = 3 * M, restlen = rowlen - 3, F = sd.F;
        const float* g = g_feat + ob * rowlen;
        float* drest = sd.d_rest + ib * restlen;
        flat_loop(cnt * restlen,
                  [&](int j) {
                      const int row = j / restlen, c = j - row * restlen;
                      return __ldcs(g + (size_t)row * rowlen + 3 + c);
                  },
                  [&](int j, float v) { drest[j] = v; });
        float* ddc = sd.d_dc + ib * 3 * F;
        const int dclen = 3 * F;
        const bool plain_dc = F == 1 && !actor;
        flat_loop(cnt * dclen,
                  [&](int j) {
                      const int row = j / dclen, rem = j - row * dclen, f = rem / 3, c = rem - f * 3;
                      const float gv = g[(size_t)row * rowlen + c];
                      return plain_dc ? gv : gv * sd.idft[f];
                  },
                  [&](int j, float v) { ddc[j] = v; });
    }
    {  // means
        const float* g = g_xyz + ob * 3;
        float* o = sd.d_xyz + ib * 3;
        if (!actor) {
            flat_loop(cnt * 3, [&](int j) { return __ldcs(g + j); }, [&](int j, float v) { o[j] = v; });
        } else {
            flat_loop(cnt * 3,
                      [&](int j) {
                          const int row = j / 3, c = j - row * 3;  // R^T g
                          float d = pose.R[c] * g[row * 3] + pose.R[3 + c] * g[row * 3 + 1] + pose.R[6 + c] * g[row * 3 + 2];
                          if (c == 1 && fl && fl[row]) d = -d;
                          return d;
                      },
                      [&](int j, float v) { o[j] = v; });
        }
    }
    }  // !staged
    // rotations (+ the pose partials of actors)
    const float4* s = staged ? reinterpret_cast<const float4*>(s_rot) : reinterpret_cast<const float4*>(sd.rotation) + ib;
    const float4* g = staged ? reinterpret_cast<const float4*>(s_gr) : reinterpret_cast<const float4*>(g_rot) + ob;
    float4* o = reinterpret_cast<float4*>(sd.d_rotation) + ib;
    struct RotIn { float4 v, g; };
    if (!actor) {
        flat_loop(cnt, [&](int row) { return RotIn{__ldcs(s + row), __ldcs(g + row)}; },
                  [&](int row, RotIn in) {
                      const Quat d = qnormalize_bwd(Quat{in.v.x, in.v.y, in.v.z, in.v.w}, Quat{in.g.x, in.g.y, in.g.z, in.g.w});
                      o[row] = make_float4(d.w, d.x, d.y, d.z);
                  });
        return;
    }
    const Quat qo{pose.q[0], pose.q[1], pose.q[2], pose.q[3]};
    const float* xs = staged ? s_xyz : sd.xyz + ib * 3;
    const float* gx = staged ? s_gx : g_xyz + ob * 3;
    float acc[16];  // dR (9, row-major), dt (3), dq from the rotation product (4)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (int row = tid; row < cnt; row += CMP_THREADS) {
        const float4 v = s[row], gv = g[row];
        const bool flip = fl && fl[row];
        const Quat raw{v.x, v.y, v.z, v.w};
        const Quat ql0 = qnormalize(raw);
        const Quat ql = flip ? qflip(ql0) : ql0;
        const Quat gq = qnormalize_bwd(qmul(qo, ql), Quat{gv.x, gv.y, gv.z, gv.w});
        // cotangents of a (= obj_rot) and b (= ql) of the raw product
        acc[12] += ql.w * gq.w + ql.x * gq.x + ql.y * gq.y + ql.z * gq.z;
        acc[13] += -ql.x * gq.w + ql.w * gq.x - ql.z * gq.y + ql.y * gq.z;
        acc[14] += -ql.y * gq.w + ql.z * gq.x + ql.w * gq.y - ql.x * gq.z;
        acc[15] += -ql.z * gq.w - ql.y * gq.x + ql.x * gq.y + ql.w * gq.z;
        Quat db;
        db.w = qo.w * gq.w + qo.x * gq.x + qo.y * gq.y + qo.z * gq.z;
        db.x = -qo.x * gq.w + qo.w * gq.x + qo.z * gq.y - qo.y * gq.z;
        db.y = -qo.y * gq.w - qo.z * gq.x + qo.w * gq.y + qo.x * gq.z;
        db.z = -qo.z * gq.w + qo.y * gq.x - qo.x * gq.y + qo.w * gq.z;
        if (flip) db = qflip_bwd(db);
        const Quat d = qnormalize_bwd(raw, db);
        o[row] = make_float4(d.w, d.x, d.y, d.z);
        // X = R x_l + t
        const float x0 = xs[row * 3], x2 = xs[row * 3 + 2];
        const float x1 = flip ? -xs[row * 3 + 1] : xs[row * 3 + 1];
        const float g0 = gx[row * 3], g1 = gx[row * 3 + 1], g2 = gx[row * 3 + 2];
        acc[0] += g0 * x0; acc[1] += g0 * x1; acc[2] += g0 * x2;
        acc[3] += g1 * x0; acc[4] += g1 * x1; acc[5] += g1 * x2;
        acc[6] += g2 * x0; acc[7] += g2 * x1; acc[8] += g2 * x2;
        acc[9] += g0; acc[10] += g1; acc[11] += g2;
    }
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float v = acc[i];
#pragma unroll
        for (int ofs = 16; ofs >= 1; ofs >>= 1) v += __shfl_xor_sync(0xffffffffu, v, ofs);
        if (lane == 0) s_red[warp][i] = v;
    }
    __syncthreads();
    if (tid < 16) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < CMP_THREADS / 32; ++w) v += s_red[w][tid];
        atomicAdd(pose_acc + (size_t)k_sub * 16 + tid, v);
    }
}

// chain the accumulated dL/dR through quaternion_to_matrix (which normalises its argument) and add the cotangent that
// reached obj_rot through the rotation product
__global__ void compose_pose_finish_kernel(const SubDev* __restrict__ subs, int n_sub, const float* __restrict__ pose_acc,
                                           float* __restrict__ d_rots, float* __restrict__ d_trans) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_sub) return;
    float dq[4] = {0.f, 0.f, 0.f, 0.f}, dt[3] = {0.f, 0.f, 0.f};
    if (subs[k].is_actor && subs[k].n > 0) {
        const float* G = pose_acc + (size_t)k * 16;
        const float* q = subs[k].obj_rot;
        const float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        const float inv = 1.0f / nrm;
        const float r = q[0] * inv, x = q[1] * inv, y = q[2] * inv, z = q[3] * inv;
        float d[4];
        d[0] = 2.f * (-z * G[1] + y * G[2] + z * G[3] - x * G[5] - y * G[6] + x * G[7]);
        d[1] = 2.f * (y * G[1] + z * G[2] + y * G[3] - 2.f * x * G[4] - r * G[5] + z * G[6] + r * G[7] - 2.f * x * G[8]);
        d[2] = 2.f * (-2.f * y * G[0] + x * G[1] + r * G[2] + x * G[3] + z * G[5] - r * G[6] + z * G[7] - 2.f * y * G[8]);
        d[3] = 2.f * (-2.f * z * G[0] - r * G[1] + x * G[2] + r * G[3] - 2.f * z * G[4] + y * G[5] + x * G[6] + y * G[7]);
        const float dot = r * d[0] + x * d[1] + y * d[2] + z * d[3];
        dq[0] = (d[0] - r * dot) * inv + G[12];
        dq[1] = (d[1] - x * dot) * inv + G[13];
        dq[2] = (d[2] - y * dot) * inv + G[14];
        dq[3] = (d[3] - z * dot) * inv + G[15];
        dt[0] = G[9]; dt[1] = G[10]; dt[2] = G[11];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) d_rots[k * 4 + i] = dq[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) d_trans[k * 3 + i] = dt[i];
}

static size_t desc_bytes(int n_sub) { return align_up((size_t)(n_sub > 0 ? n_sub : 1) * sizeof(SubDev), 256); }

// Validates the host descriptors and builds the device table; returns the number of CTAs (chunks), or -1.
static int build_table(const grpg_compose_submodel* subs, int n_sub, int M, bool backward, std::vector<SubDev>& tab) {
    if (!subs || n_sub <= 0) { grpg_loss_fail("grpg_compose: no sub-models"); return -1; }
    if (n_sub > CMP_THREADS) { grpg_loss_fail("grpg_compose: more than 256 sub-models"); return -1; }
    if (M < 1) { grpg_loss_fail("grpg_compose: M (SH coefficients per Gaussian) must be >= 1"); return -1; }
    tab.resize(n_sub);
    long long offset = 0, chunks = 0;
    for (int k = 0; k < n_sub; ++k) {
        const grpg_compose_submodel& s = subs[k];
        SubDev& d = tab[k];
        std::memset(&d, 0, sizeof(d));
        if (s.n < 0) { grpg_loss_fail("grpg_compose: negative sub-model size"); return -1; }
        if (s.fourier_dim < 1 || s.fourier_dim > GRPG_COMPOSE_MAX_FOURIER) {
            grpg_loss_fail("grpg_compose: fourier_dim out of range [1, GRPG_COMPOSE_MAX_FOURIER]"); return -1;
        }
        if (s.n > 0) {
            if (!s.xyz || !s.scaling || !s.rotation || !s.opacity || !s.features_dc || (M > 1 && !s.features_rest)) {
                grpg_loss_fail("grpg_compose: null parameter pointer"); return -1;
            }
            if (((uintptr_t)s.rotation & 15) || (backward && ((uintptr_t)s.d_rotation & 15))) {
                grpg_loss_fail("grpg_compose: rotation pointers must be 16-byte aligned"); return -1;
            }
            if (backward && (!s.d_xyz || !s.d_scaling || !s.d_rotation || !s.d_opacity || !s.d_features_dc ||
                             (M > 1 && !s.d_features_rest))) {
                grpg_loss_fail("grpg_compose: null gradient pointer"); return -1;
            }
        }
        d.xyz = s.xyz; d.scaling = s.scaling; d.rotation = s.rotation; d.opacity = s.opacity;
        d.dc = s.features_dc; d.rest = s.features_rest; d.flip = s.flip_mask;
        d.d_xyz = s.d_xyz; d.d_scaling = s.d_scaling; d.d_rotation = s.d_rotation; d.d_opacity = s.d_opacity;
        d.d_dc = s.d_features_dc; d.d_rest = s.d_features_rest;
        d.offset = offset; d.n = s.n; d.is_actor = s.is_actor ? 1 : 0; d.F = s.fourier_dim; d.chunk_begin = (int)chunks;
        {
            uintptr_t bits = (uintptr_t)s.xyz | (uintptr_t)s.scaling | (uintptr_t)s.rotation | (uintptr_t)s.opacity |
                             (uintptr_t)s.features_dc | (uintptr_t)s.features_rest;
            if (backward)
                bits |= (uintptr_t)s.d_xyz | (uintptr_t)s.d_scaling | (uintptr_t)s.d_rotation | (uintptr_t)s.d_opacity |
                        (uintptr_t)s.d_features_dc | (uintptr_t)s.d_features_rest;
            d.aligned16 = (bits & 15) == 0 ? 1 : 0;
        }
        for (int i = 0; i < GRPG_COMPOSE_MAX_FOURIER; ++i) d.idft[i] = i < s.fourier_dim ? s.idft[i] : 0.f;
        if (d.is_actor) {
            if (!s.obj_rot || !s.obj_trans) { grpg_loss_fail("grpg_compose: actor without a pose"); return -1; }
            d.obj_rot = s.obj_rot; d.obj_trans = s.obj_trans;
        } else {
            d.idft[0] = 1.0f;
        }
        offset += s.n;
        chunks += (s.n + CMP_CHUNK - 1) / CMP_CHUNK;
        if (offset > 0x7fffffffLL / 64) { grpg_loss_fail("grpg_compose: too many Gaussians"); return -1; }
    }
    return (int)chunks;
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

extern "C" size_t grpg_compose_workspace_bytes(int n_sub) {
    return desc_bytes(n_sub) + align_up((size_t)(n_sub > 0 ? n_sub : 1) * 16 * sizeof(float), 256);
}

extern "C" int grpg_compose_forward(const grpg_compose_submodel* subs, int n_sub, int M, void* workspace, float* xyz,
                                    float* rotation, float* scaling, float* opacity, float* features, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    std::vector<SubDev> tab;
    const int chunks = build_table(subs, n_sub, M, false, tab);
    if (chunks < 0) return 1;
    if (chunks == 0) return 0;
    if (!workspace || !xyz || !rotation || !scaling || !opacity || !features) return grpg_loss_fail("grpg_compose_forward: null output or workspace");
    if ((uintptr_t)rotation & 15) return grpg_loss_fail("grpg_compose_forward: rotation output must be 16-byte aligned");
    if (cudaMemcpyAsync(workspace, tab.data(), tab.size() * sizeof(SubDev), cudaMemcpyHostToDevice, stream) != cudaSuccess)
        return check_launch("grpg_compose_forward: descriptor upload");
    // staged path: shared memory for one chunk of the widest sub-model; outputs must be 16-byte aligned
    int maxF = 1;
    for (int k = 0; k < n_sub; ++k) maxF = subs[k].fourier_dim > maxF ? subs[k].fourier_dim : maxF;
    size_t smem = (size_t)CMP_CHUNK * 4 * (3 + 3 + 4 + 1 + 3 * maxF + 3 * (M - 1));
    int staged = (((uintptr_t)xyz | (uintptr_t)rotation | (uintptr_t)scaling | (uintptr_t)opacity | (uintptr_t)features) & 15) == 0;
    if (smem > 160 * 1024) staged = 0;
    if (!staged) smem = 0;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(compose_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return check_launch("grpg_compose_forward: shared memory opt-in");
    ProfScope ps("compose_fwd", stream);
    compose_fwd_kernel<<<chunks, CMP_THREADS, smem, stream>>>(reinterpret_cast<const SubDev*>(workspace), n_sub, M, xyz,
                                                              rotation, scaling, opacity, features, staged);
    return check_launch("grpg_compose_forward");
}

extern "C" int grpg_compose_backward(const grpg_compose_submodel* subs, int n_sub, int M, void* workspace,
                                     const float* g_xyz, const float* g_rotation, const float* g_scaling,
                                     const float* g_opacity, const float* g_features, float* d_obj_rots,
                                     float* d_obj_trans, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    std::vector<SubDev> tab;
    const int chunks = build_table(subs, n_sub, M, true, tab);
    if (chunks < 0) return 1;
    if (!workspace || !d_obj_rots || !d_obj_trans) return grpg_loss_fail("grpg_compose_backward: null output or workspace");
    if (chunks > 0 && (!g_xyz || !g_rotation || !g_scaling || !g_opacity || !g_features))
        return grpg_loss_fail("grpg_compose_backward: null cotangent");
    if ((uintptr_t)g_rotation & 15) return grpg_loss_fail("grpg_compose_backward: rotation cotangent must be 16-byte aligned");
    float* pose_acc = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + desc_bytes(n_sub));
    if (cudaMemcpyAsync(workspace, tab.data(), tab.size() * sizeof(SubDev), cudaMemcpyHostToDevice, stream) != cudaSuccess)
        return check_launch("grpg_compose_backward: descriptor upload");
    cudaMemsetAsync(pose_acc, 0, (size_t)n_sub * 16 * sizeof(float), stream);
    ProfScope ps("compose_bwd", stream);
    size_t smem = (size_t)CMP_CHUNK * 4 * (22 + 3 * M);
    int staged = (((uintptr_t)g_xyz | (uintptr_t)g_rotation | (uintptr_t)g_scaling | (uintptr_t)g_opacity | (uintptr_t)g_features) & 15) == 0;
    if (smem > 160 * 1024) staged = 0;
    if (!staged) smem = 0;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(compose_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return check_launch("grpg_compose_backward: shared memory opt-in");
    if (chunks > 0)
        compose_bwd_kernel<<<chunks, CMP_THREADS, smem, stream>>>(reinterpret_cast<const SubDev*>(workspace), n_sub, M, g_xyz,
                                                                  g_rotation, g_scaling, g_opacity, g_features, pose_acc, staged);
    compose_pose_finish_kernel<<<(n_sub + 63) / 64, 64, 0, stream>>>(reinterpret_cast<const SubDev*>(workspace), n_sub,
                                                                      pose_acc, d_obj_rots, d_obj_trans);
    return check_launch("grpg_compose_backward");
}
