// blend_bwd.cu -- per-tile back-to-front replay producing per-Gaussian 2D gradients.
//
// Behavioural spec: reference renderCUDA backward (cuda_rasterizer/backward.cu:415-641):
// T starts at T_final = 1 - alpha_out (:468), instances behind a pixel's last contributor are
// skipped (:530-532), the 0.99 clamp is straight-through (:543,618), dL_dmean2D is in NDC units
// with a third |gx|+|gy| channel (:625-628), dL_dconic.z is unused (:633-635).
//
// B200 design.  The reference issues 11+S global float atomics per (pixel, Gaussian) hit.
// Here a warp owns an 8x4 pixel block, culls instances with the same conservative footprint
// as the forward, reduces the per-lane partials with shuffles and issues ONE vector of
// atomics per (warp, Gaussian) -- lanes 0..10 each add one component of the 48-byte gradient
// record, so the 11 adds leave the SM as a single coalesced RED request.
#include "grpg_common.cuh"
#include <stdlib.h>

namespace grpg {

constexpr int BWD_BATCH = 256;
constexpr int GREC = 12;  // floats per Gaussian in the gradient record
#define GRPG_BWD_PPL_DEFAULT 3  // 3 = the packed two-pixel kernel; measured on the 2 M scene: 1 -> 1.21 ms, 2 -> 1.08 ms, 4 -> 1.20 ms

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sums 12 per-lane values over the warp with 13 shuffles instead of 60: at every butterfly step a lane hands
// half of its remaining slots to its partner and keeps the other half (12 -> 6 -> 3(+1 pad) -> 2 -> 1), and
// the xor-1 step completes the one slot the lane is left with.  On return the lane holds the warp total of
// slot 6*b4 + 3*b3 + (2*b2 + b1) (b_k = bit k of the lane id); lanes with 2*b2 + b1 == 3 hold the pad.
__device__ __forceinline__ float warp_reduce12(const float (&v)[12], int lane) {
    float a[6], b[3];
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float send = h4 ? v[i] : v[i + 6];
        const float keep = h4 ? v[i + 6] : v[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float send = h3 ? a[i] : a[i + 3];
        const float keep = h3 ? a[i + 3] : a[i];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const float c0 = (h2 ? b[2] : b[0]) + __shfl_xor_sync(0xffffffffu, h2 ? b[0] : b[2], 4);
    const float c1 = (h2 ? 0.f : b[1]) + __shfl_xor_sync(0xffffffffu, h2 ? b[1] : 0.f, 4);
    float r = (h1 ? c1 : c0) + __shfl_xor_sync(0xffffffffu, h1 ? c0 : c1, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

template <int SB>
__global__ void __launch_bounds__(256) blend_bwd_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, const Rec* __restrict__ rec,
    const float* __restrict__ semantics, int S, int W, int H, const float* __restrict__ bg_color,
    const float* __restrict__ alphas, const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
    const float* __restrict__ dL_dpixel_depths, const float* __restrict__ dL_dalphas,
    const float* __restrict__ dL_dpixel_semantics, float* __restrict__ grad_rec /*[P][12]*/,
    float* __restrict__ dL_dsemantics /*[P][S]*/, int HL, int row_stride, int row_phase, int grads_full) {
    // staged records (48-byte stride) and ids, double buffered and filled one batch ahead with cp.async
    // (see blend_fwd.cu)
    __shared__ __align__(16) float4 s_rec2[2][BWD_BATCH * 3];
    __shared__ uint32_t s_id2[2][BWD_BATCH];
    __shared__ int s_maxlast[8];
    __shared__ uint16_t s_q[8][BWD_BATCH];  // per-warp queue of surviving staged slots

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tiles_x = (W + GRPG_TILE - 1) / GRPG_TILE;
    const uint32_t tile = blockIdx.y * tiles_x + blockIdx.x;
    const int bx0 = blockIdx.x * GRPG_TILE + (warp & 1) * 8;
    const int by0 = (blockIdx.y * row_stride + row_phase) * GRPG_TILE + (warp >> 1) * 4;
    const int pix_x = bx0 + (lane & 7), pix_y = by0 + (lane >> 3);
    const int loc_y = blockIdx.y * GRPG_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = pix_x < W && pix_y < H;
    const float pxf = (float)pix_x, pyf = (float)pix_y;
    const float bx_lo = (float)bx0, bx_hi = (float)(bx0 + 7), by_lo = (float)by0, by_hi = (float)(by0 + 3);
    const size_t pid = (size_t)loc_y * W + pix_x;
    // pixel gradients: band-compact like `alphas`, or full frames indexed by the true row
    const size_t hw = grads_full ? (size_t)H * W : (size_t)HL * W;
    const size_t gid_px = grads_full ? (size_t)pix_y * W + pix_x : pid;

    const uint2 range = ranges[tile];
    const int n_inst = (int)(range.y - range.x);

    const float T_final = inside ? 1.0f - alphas[pid] : 0.0f;
    float T = T_final;
    const int last_contributor = inside ? (int)n_contrib[pid] : 0;

    float dpix0 = 0.f, dpix1 = 0.f, dpix2 = 0.f, dpix_depth = 0.f, dpix_alpha = 0.f;
    float dsem[SB > 0 ? SB : 1];
    if (inside) {
        dpix0 = dL_dpixels[gid_px]; dpix1 = dL_dpixels[hw + gid_px]; dpix2 = dL_dpixels[2 * hw + gid_px];
        dpix_depth = dL_dpixel_depths[gid_px];
        dpix_alpha = dL_dalphas[gid_px];
    }
#pragma unroll
    for (int i = 0; i < (SB > 0 ? SB : 1); ++i)
        dsem[i] = (SB > 0 && inside && i < S) ? dL_dpixel_semantics[(size_t)i * hw + gid_px] : 0.f;
    const float bg_dot_dpixel = bg_color[0] * dpix0 + bg_color[1] * dpix1 + bg_color[2] * dpix2;

    // the reference's per-channel "last contributor" recurrences (backward.cu:560-604) collapse into one scalar: see
    // the wide kernel below (A <- a_k s_k + (1 - a_k) A with s_k = sum_u u_k g_u, dL/da = s_k - A)
    float accA = 0.f;

    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    // record component this lane owns after warp_reduce12 (11 = nothing)
    const int red_sub = ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    const int red_slot = red_sub < 3 ? ((lane >> 4) & 1) * 6 + ((lane >> 3) & 1) * 3 + red_sub : 11;

    // nothing behind the deepest last-contributor of the tile can receive gradient
    int wmax = __reduce_max_sync(0xffffffffu, last_contributor);
    if (lane == 0) s_maxlast[warp] = wmax;
    __syncthreads();
    int tile_last = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tile_last = max(tile_last, s_maxlast[w]);
    tile_last = min(tile_last, n_inst);

    // slot t of the batch that starts at `top` holds 0-based position top-1-t: ascending slot = back-to-front
    auto stage = [&](int buf, int top, uint32_t id) {
        if (top - 1 - tid >= 0) {
            const float4* r = reinterpret_cast<const float4*>(rec + id);
            float4* d = &s_rec2[buf][3 * tid];
            cp_async16(d, r); cp_async16(d + 1, r + 1); cp_async16(d + 2, r + 2);
            s_id2[buf][tid] = id;
        }
        cp_async_commit();
    };
    auto fetch_id = [&](int top) { return top - 1 - tid >= 0 ? point_list[range.x + top - 1 - tid] : 0u; };
    uint32_t id_next = fetch_id(tile_last);
    stage(0, tile_last, id_next);
    id_next = fetch_id(tile_last - BWD_BATCH);

    for (int top = tile_last, it = 0; top > 0; top -= BWD_BATCH, ++it) {
        const int cnt = min(BWD_BATCH, top);
        cp_async_wait_all();
        __syncthreads();  // batch `it` has landed for everyone, and every warp has left batch it-1
        const float4* s_rec = s_rec2[it & 1];
        const uint32_t* s_id = s_id2[it & 1];
        stage((it + 1) & 1, top - BWD_BATCH, id_next);
        id_next = fetch_id(top - 2 * BWD_BATCH);
        if (wmax <= top - cnt) continue;  // every pixel of this warp ended before this batch

        // phase 1: footprint test of the whole batch, survivors compacted into a per-warp byte queue
        uint16_t* q = s_q[warp];
        const char* rec_base = reinterpret_cast<const char*>(s_rec);
        int n_q = 0;
        for (int g0 = 0; g0 < cnt; g0 += 32) {
            const int j = g0 + lane;
            // slots whose position lies behind every pixel's last contributor cannot receive gradient
            const bool hit = j < cnt && (top - 1 - j) < wmax && footprint_hits_exact(s_rec[3 * j], s_rec[3 * j + 1], bx_lo, bx_hi, by_lo, by_hi);
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            if (hit) q[n_q + __popc(m & ((1u << lane) - 1u))] = (uint16_t)j;
            n_q += __popc(m);
        }
        __syncwarp();
        // phase 2: back-to-front replay of the survivors
        {
            for (int qi = 0; qi < n_q; ++qi) {
                const int k = (int)q[qi];
                const float4* rk = reinterpret_cast<const float4*>(rec_base + (uint32_t)k * 48u);
                const int pos = top - 1 - k;  // 0-based position in the tile's range
                const float4 a = rk[0];
                const float4 b = rk[1];
                const uint32_t gid = s_id[k];
                const float dx = a.x - pxf, dy = a.y - pyf;
                const float power = ffma(ffma(dx, fmul(dx, b.x), fmul(dy, fmul(dy, b.z))), -0.5f, -fmul(dy, fmul(dx, b.y)));
                const bool maybe = (pos < last_contributor) && !(power > 0.0f) && !(power < a.w);
                if (!__any_sync(0xffffffffu, maybe)) continue;  // exact-ellipse vote, see blend_fwd.cu
                const float G = expf(power);
                const float alpha = fminf(0.99f, b.w * G);
                const bool active = maybe && (alpha >= 1.0f / 255.0f);
                if (!__any_sync(0xffffffffu, active)) continue;

                float g_mx = 0.f, g_my = 0.f, g_mabs = 0.f, g_cx = 0.f, g_cy = 0.f, g_cw = 0.f, g_op = 0.f;
                float g_c0 = 0.f, g_c1 = 0.f, g_c2 = 0.f, g_d = 0.f;
                float g_sem[SB > 0 ? SB : 1];
#pragma unroll
                for (int i = 0; i < (SB > 0 ? SB : 1); ++i) g_sem[i] = 0.f;
                // all loads happen before the divergent section
                const float4 c = rk[2];
                float sv[SB > 0 ? SB : 1];
                if (SB > 0) {
                    const float* sp = semantics + (size_t)gid * S;
#pragma unroll
                    for (int i = 0; i < SB; ++i) sv[i] = i < S ? __ldg(sp + i) : 0.f;
                }
                if (active) {
                    // one reciprocal serves T/(1-a) and T_final/(1-a); 1-a lies in [0.01, 1], so the single-instruction
                    // approximation (1 ulp) needs no range fix-up -- its error is far inside the gradient bar
                    float inv_1ma;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_1ma) : "f"(1.f - alpha));
                    T = T * inv_1ma;
                    const float w_at = alpha * T;
                    float s_k = __fmaf_rn(c.z, dpix2, __fmaf_rn(c.y, dpix1, __fmaf_rn(c.x, dpix0, dpix_alpha)));
                    g_c0 = w_at * dpix0; g_c1 = w_at * dpix1; g_c2 = w_at * dpix2;
                    if (SB > 0) {
#pragma unroll
                        for (int i = 0; i < SB; ++i) {
                            if (i < S) {
                                s_k = __fmaf_rn(sv[i], dsem[i], s_k);
                                g_sem[i] = w_at * dsem[i];
                            }
                        }
                    }
                    s_k = __fmaf_rn(c.w, dpix_depth, s_k); g_d = w_at * dpix_depth;
                    float dL_dopa = (s_k - accA) * T;
                    accA = __fmaf_rn(alpha, s_k, (1.f - alpha) * accA);
                    dL_dopa += (-T_final * inv_1ma) * bg_dot_dpixel;

                    const float dL_dG = b.w * dL_dopa;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * b.x - gdy * b.y;
                    const float dG_ddely = -gdy * b.z - gdx * b.y;
                    g_mx = dL_dG * dG_ddelx * ddelx_dx;
                    g_my = dL_dG * dG_ddely * ddely_dy;
                    g_mabs = fabsf(g_mx) + fabsf(g_my);
                    g_cx = -0.5f * gdx * dx * dL_dG;
                    g_cy = -0.5f * gdx * dy * dL_dG;
                    g_cw = -0.5f * gdy * dy * dL_dG;
                    g_op = G * dL_dopa;
                }
                // warp reduction (12-slot butterfly), then one vector of atomics per (warp, Gaussian):
                // even lanes own one component each of the 48-byte gradient record
                const float vals[12] = {g_mx, g_my, g_mabs, g_cx, g_cy, g_cw, g_op, g_c0, g_c1, g_c2, g_d, 0.f};
                const float mine = warp_reduce12(vals, lane);
                if (!(lane & 1) && red_slot < 11) atomicAdd(grad_rec + (size_t)gid * GREC + red_slot, mine);
                if (SB > 0) {
#pragma unroll
                    for (int i = 0; i < SB; ++i) {
                        if (i < S) {
                            const float v = warp_sum(g_sem[i]);
                            if (lane == 0) atomicAdd(dL_dsemantics + (size_t)gid * S + i, v);
                        }
                    }
                }
            }
        }
    }
    cp_async_wait_all();  // nothing may be in flight into shared memory when the CTA retires
}


// ---- wide variant (no semantic channels): PPL pixels per lane ------------------------------------------------------
// A warp owns an 8 x (4*PPL) pixel block and lane l the pixels (l&7, (l>>3) + 4p), p < PPL.  The 12-slot butterfly
// and the RED vector -- 50 of the ~185 warp instructions of a (warp, Gaussian) pair in the one-pixel layout -- and
// the queue/record loads are then paid once per 32*PPL pixels; each lane adds its PPL partials in registers first.
// Per-pixel arithmetic is the same sequence as in blend_bwd_kernel.
template <int PPL, int MINB>
__global__ void __launch_bounds__(256 / PPL, MINB) blend_bwd_wide_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, const Rec* __restrict__ rec, int W, int H,
    const float* __restrict__ bg_color, const float* __restrict__ alphas, const uint32_t* __restrict__ n_contrib,
    const float* __restrict__ dL_dpixels, const float* __restrict__ dL_dpixel_depths,
    const float* __restrict__ dL_dalphas, float* __restrict__ grad_rec /*[P][12]*/, int HL, int row_stride,
    int row_phase, int grads_full) {
    constexpr int NT = 256 / PPL, NW = 8 / PPL, RPT = BWD_BATCH / NT;
    __shared__ __align__(16) float4 s_rec2[2][BWD_BATCH * 3];
    __shared__ uint32_t s_id2[2][BWD_BATCH];
    __shared__ int s_maxlast[NW];
    __shared__ uint16_t s_q[NW][BWD_BATCH];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tiles_x = (W + GRPG_TILE - 1) / GRPG_TILE;
    const uint32_t tile = blockIdx.y * tiles_x + blockIdx.x;
    const int bx0 = blockIdx.x * GRPG_TILE + (warp & 1) * 8;
    const int wy0 = (warp >> 1) * 4 * PPL;
    const int by0 = (blockIdx.y * row_stride + row_phase) * GRPG_TILE + wy0;
    const int pix_x = bx0 + (lane & 7);
    const float pxf = (float)pix_x;
    const float bx_lo = (float)bx0, bx_hi = (float)(bx0 + 7), by_lo = (float)by0, by_hi = (float)(by0 + 4 * PPL - 1);
    const size_t hw = grads_full ? (size_t)H * W : (size_t)HL * W;

    const uint2 range = ranges[tile];
    const int n_inst = (int)(range.y - range.x);
    const float bg0 = bg_color[0], bg1 = bg_color[1], bg2 = bg_color[2];

    // Per-pixel state: 9 registers.  The reference carries, per channel u (3 colours, depth, alpha with u_k = 1), the
    // recurrence acc_u <- a_last u_last + (1 - a_last) acc_u and forms dL/da = sum_u (u_k - acc_u) g_u
    // (backward.cu:560-604).  Only the g-weighted sum of the acc_u is ever used, and it obeys the same recurrence:
    // with s_k = sum_u u_k g_u,  A <- a_k s_k + (1 - a_k) A  and  dL/da = s_k - A.  One scalar replaces five (and the
    // five "last" values), 7 instead of 20 instructions per pixel and Gaussian; rounding is of the same order as the
    // reference's own (each acc_u carries the rounding of its recurrence too).
    float pyf[PPL], T[PPL], tfb[PPL] /* T_final * (bg . dL_dpixel) */, dpix0[PPL], dpix1[PPL], dpix2[PPL], dpix_depth[PPL],
        dpix_alpha[PPL];
    float accA[PPL];
    int last_contributor[PPL];
    int lmax = 0;
#pragma unroll
    for (int p = 0; p < PPL; ++p) {
        const int pix_y = by0 + (lane >> 3) + 4 * p;
        const int loc_y = blockIdx.y * GRPG_TILE + wy0 + (lane >> 3) + 4 * p;
        const bool inside = pix_x < W && pix_y < H;
        const size_t pid = (size_t)loc_y * W + pix_x;
        const size_t gpx = grads_full ? (size_t)pix_y * W + pix_x : pid;
        pyf[p] = (float)pix_y;
        T[p] = inside ? 1.0f - alphas[pid] : 0.0f;  // T_final (backward.cu:468)
        last_contributor[p] = inside ? (int)n_contrib[pid] : 0;
        lmax = max(lmax, last_contributor[p]);
        dpix0[p] = inside ? dL_dpixels[gpx] : 0.f;
        dpix1[p] = inside ? dL_dpixels[hw + gpx] : 0.f;
        dpix2[p] = inside ? dL_dpixels[2 * hw + gpx] : 0.f;
        dpix_depth[p] = inside ? dL_dpixel_depths[gpx] : 0.f;
        dpix_alpha[p] = inside ? dL_dalphas[gpx] : 0.f;
        tfb[p] = T[p] * (bg0 * dpix0[p] + bg1 * dpix1[p] + bg2 * dpix2[p]);
        accA[p] = 0.f;
    }
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    const int red_sub = ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    const int red_slot = red_sub < 3 ? ((lane >> 4) & 1) * 6 + ((lane >> 3) & 1) * 3 + red_sub : 11;
    // factor applied to the warp total of the slot this lane ends up owning (see the per-pixel section)
    const float slot_scale = red_slot == 0 ? -ddelx_dx : red_slot == 1 ? -ddely_dy : (red_slot >= 3 && red_slot <= 5) ? -0.5f : 1.0f;

    const int wmax = __reduce_max_sync(0xffffffffu, lmax);
    if (lane == 0) s_maxlast[warp] = wmax;
    __syncthreads();
    int tile_last = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) tile_last = max(tile_last, s_maxlast[w]);
    tile_last = min(tile_last, n_inst);

    auto stage = [&](int buf, int top, const uint32_t (&id)[RPT]) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int t = tid + r * NT;
            if (top - 1 - t >= 0) {
                const float4* src = reinterpret_cast<const float4*>(rec + id[r]);
                float4* d = &s_rec2[buf][3 * t];
                cp_async16(d, src); cp_async16(d + 1, src + 1); cp_async16(d + 2, src + 2);
                s_id2[buf][t] = id[r];
            }
        }
        cp_async_commit();
    };
    auto fetch_id = [&](int top, uint32_t (&id)[RPT]) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int t = tid + r * NT;
            id[r] = top - 1 - t >= 0 ? point_list[range.x + top - 1 - t] : 0u;
        }
    };
    uint32_t id_next[RPT];
    fetch_id(tile_last, id_next);
    stage(0, tile_last, id_next);
    fetch_id(tile_last - BWD_BATCH, id_next);

    for (int top = tile_last, it = 0; top > 0; top -= BWD_BATCH, ++it) {
        const int cnt = min(BWD_BATCH, top);
        cp_async_wait_all();
        __syncthreads();
        const float4* s_rec = s_rec2[it & 1];
        const uint32_t* s_id = s_id2[it & 1];
        stage((it + 1) & 1, top - BWD_BATCH, id_next);
        fetch_id(top - 2 * BWD_BATCH, id_next);
        if (wmax <= top - cnt) continue;

        uint16_t* q = s_q[warp];
        const char* rec_base = reinterpret_cast<const char*>(s_rec);
        int n_q = 0;
        for (int g0 = 0; g0 < cnt; g0 += 32) {
            const int j = g0 + lane;
            const bool hit = j < cnt && (top - 1 - j) < wmax &&
                             footprint_hits_exact(s_rec[3 * j], s_rec[3 * j + 1], bx_lo, bx_hi, by_lo, by_hi);
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            if (hit) q[n_q + __popc(m & ((1u << lane) - 1u))] = (uint16_t)j;
            n_q += __popc(m);
        }
        __syncwarp();
        for (int qi = 0; qi < n_q; ++qi) {
            const int k = (int)q[qi];
            const float4* rk = reinterpret_cast<const float4*>(rec_base + (uint32_t)k * 48u);
            const int pos = top - 1 - k;
            const float4 a = rk[0];
            const float4 b = rk[1];
            const uint32_t gid = s_id[k];
            const float dx = a.x - pxf;
            const float dxA = fmul(dx, b.x), dxB = fmul(dx, b.y);
            float dy[PPL], power[PPL];
            bool maybe[PPL], any_maybe = false;
#pragma unroll
            for (int p = 0; p < PPL; ++p) {
                dy[p] = a.y - pyf[p];
                power[p] = ffma(ffma(dx, dxA, fmul(dy[p], fmul(dy[p], b.z))), -0.5f, -fmul(dy[p], dxB));
                maybe[p] = (pos < last_contributor[p]) && !(power[p] > 0.0f) && !(power[p] < a.w);
                any_maybe |= maybe[p];
            }
            if (!__any_sync(0xffffffffu, any_maybe)) continue;
            float G[PPL], alpha[PPL];
            bool active[PPL], any_active = false;
#pragma unroll
            for (int p = 0; p < PPL; ++p) {
                G[p] = expf(power[p]);
                alpha[p] = fminf(0.99f, b.w * G[p]);
                active[p] = maybe[p] && (alpha[p] >= 1.0f / 255.0f);
                any_active |= active[p];
            }
            if (!__any_sync(0xffffffffu, any_active)) continue;

            const float4 c = rk[2];
            float vals[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) vals[i] = 0.f;
#pragma unroll
            for (int p = 0; p < PPL; ++p) {
                if (active[p]) {
                    float inv_1ma;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_1ma) : "f"(1.f - alpha[p]));
                    T[p] = T[p] * inv_1ma;
                    const float w_at = alpha[p] * T[p];
                    const float s_k = __fmaf_rn(c.w, dpix_depth[p], __fmaf_rn(c.z, dpix2[p], __fmaf_rn(c.y, dpix1[p],
                                                __fmaf_rn(c.x, dpix0[p], dpix_alpha[p]))));
                    float dL_dopa = (s_k - accA[p]) * T[p];
                    dL_dopa += -(tfb[p] * inv_1ma);
                    accA[p] = __fmaf_rn(alpha[p], s_k, (1.f - alpha[p]) * accA[p]);  // for the next (nearer) Gaussian

                    // With u = dL/dG * G * dx and v = dL/dG * G * dy the mean gradients are -(uA + vB) 0.5W and
                    // -(vC + uB) 0.5H and the conic gradients -0.5 (u dx, u dy, v dy): the per-Gaussian constants
                    // (0.5W, 0.5H, -0.5, the signs) are applied once to the warp sums (slot_scale), not per pixel.
                    const float uG = (b.w * dL_dopa) * G[p];
                    const float u = uG * dx, v = uG * dy[p];
                    const float px_ = __fmaf_rn(u, b.x, v * b.y), py_ = __fmaf_rn(v, b.z, u * b.y);
                    vals[0] += px_;
                    vals[1] += py_;
                    vals[2] = __fmaf_rn(fabsf(px_), ddelx_dx, __fmaf_rn(fabsf(py_), ddely_dy, vals[2]));
                    vals[3] = __fmaf_rn(u, dx, vals[3]);
                    vals[4] = __fmaf_rn(u, dy[p], vals[4]);
                    vals[5] = __fmaf_rn(v, dy[p], vals[5]);
                    vals[6] = __fmaf_rn(G[p], dL_dopa, vals[6]);
                    vals[7] += w_at * dpix0[p];
                    vals[8] += w_at * dpix1[p];
                    vals[9] += w_at * dpix2[p];
                    vals[10] += w_at * dpix_depth[p];
                }
            }
            const float mine = warp_reduce12(vals, lane) * slot_scale;
            if (!(lane & 1) && red_slot < 11) atomicAdd(grad_rec + (size_t)gid * GREC + red_slot, mine);
        }
    }
    cp_async_wait_all();
}

// Same butterfly as warp_reduce12 with the additions of each level issued as packed FADD2 (13 -> 8 add instructions).
__device__ __forceinline__ float warp_reduce12_packed(const float (&v)[12], int lane) {
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
    float k[6], r[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float send = h4 ? v[i] : v[i + 6];
        k[i] = h4 ? v[i + 6] : v[i];
        r[i] = __shfl_xor_sync(0xffffffffu, send, 16);
    }
    const float2 a01 = fadd2(f2(k[0], k[1]), f2(r[0], r[1]));
    const float2 a23 = fadd2(f2(k[2], k[3]), f2(r[2], r[3]));
    const float2 a45 = fadd2(f2(k[4], k[5]), f2(r[4], r[5]));
    const float a[6] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y};
    float kb[3], rb[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float send = h3 ? a[i] : a[i + 3];
        kb[i] = h3 ? a[i + 3] : a[i];
        rb[i] = __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const float2 b01 = fadd2(f2(kb[0], kb[1]), f2(rb[0], rb[1]));
    const float b2 = kb[2] + rb[2];
    const float s0 = __shfl_xor_sync(0xffffffffu, h2 ? b01.x : b2, 4);
    const float s1 = __shfl_xor_sync(0xffffffffu, h2 ? b01.y : 0.f, 4);
    const float2 c = fadd2(f2(h2 ? b2 : b01.x, h2 ? 0.f : b01.y), f2(s0, s1));
    float q = (h1 ? c.y : c.x) + __shfl_xor_sync(0xffffffffu, h1 ? c.x : c.y, 2);
    q += __shfl_xor_sync(0xffffffffu, q, 1);
    return q;
}

// ---- packed variant (no semantic channels): two vertically adjacent pixels per lane, FADD2 / FMUL2 / FFMA2 ------------
// Pixel layout and masking as in blend_fwd_packed_kernel: lane l owns (l & 7, 2 (l >> 3)) and the pixel below it, and the
// whole per-pixel chain -- `power`, expf, alpha, T / (1 - alpha), the collapsed recurrence, the 2D-mean / conic /
// opacity / colour / depth partials -- is issued once per lane on float2 operands.  A pixel that does not take part
// (behind its last contributor, power > 0, alpha < 1/255) runs with alpha = 0 and G = 0: rcp(1 - 0) = 1 leaves T, the
// recurrence keeps A (fma(0, s, 1 * A) = A) and every partial it adds is an exact zero.
template <int MINB, int NW = 4, int BATCH = BWD_BATCH>
__global__ void __launch_bounds__(32 * NW, MINB) blend_bwd_packed_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, const Rec* __restrict__ rec, int W, int H,
    const float* __restrict__ bg_color, const float* __restrict__ alphas, const uint32_t* __restrict__ n_contrib,
    const float* __restrict__ dL_dpixels, const float* __restrict__ dL_dpixel_depths,
    const float* __restrict__ dL_dalphas, float* __restrict__ grad_rec /*[P][12]*/, int HL, int row_stride,
    int row_phase, int grads_full) {
    constexpr int NT = 32 * NW, RPT = BATCH / NT;
    static_assert(NW == 4 || NW == 1, "a CTA is a tile (4 warps) or one 8x8 block (1 warp)");
    __shared__ __align__(16) float4 s_rec2[2][BATCH * 3];
    __shared__ uint32_t s_id2[2][BATCH];
    __shared__ int s_maxlast[NW];
    __shared__ uint16_t s_q[NW][BATCH];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tiles_x = (W + GRPG_TILE - 1) / GRPG_TILE;
    const int sub = NW == 4 ? warp : (int)(blockIdx.x & 3u);  // NW == 1: one single-warp CTA per 8x8 block, see blend_fwd.cu
    const uint32_t tile_x = NW == 4 ? blockIdx.x : (blockIdx.x >> 2);
    const uint32_t tile = blockIdx.y * tiles_x + tile_x;
    const int bx0 = (int)tile_x * GRPG_TILE + (sub & 1) * 8;
    const int wy0 = (sub >> 1) * 8;
    const int by0 = (blockIdx.y * row_stride + row_phase) * GRPG_TILE + wy0;
    const int pix_x = bx0 + (lane & 7);
    const int row0 = by0 + 2 * (lane >> 3);
    const float pxf = (float)pix_x;
    const float2 npy = f2(-(float)row0, -(float)(row0 + 1));
    const float bx_lo = (float)bx0, bx_hi = (float)(bx0 + 7), by_lo = (float)by0, by_hi = (float)(by0 + 7);
    const size_t hw = grads_full ? (size_t)H * W : (size_t)HL * W;

    const uint2 range = ranges[tile];
    const int n_inst = (int)(range.y - range.x);
    const float bg0 = bg_color[0], bg1 = bg_color[1], bg2 = bg_color[2];

    float sT[2], sL[2], sd0[2], sd1[2], sd2[2], sdd[2], sda[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int pix_y = row0 + p;
        const int loc_y = blockIdx.y * GRPG_TILE + wy0 + 2 * (lane >> 3) + p;
        const bool inside = pix_x < W && pix_y < H;
        const size_t pid = (size_t)loc_y * W + pix_x;
        const size_t gpx = grads_full ? (size_t)pix_y * W + pix_x : pid;
        sT[p] = inside ? 1.0f - alphas[pid] : 0.0f;  // T_final (backward.cu:468)
        sL[p] = inside ? __uint_as_float(n_contrib[pid]) : 0.0f;
        sd0[p] = inside ? dL_dpixels[gpx] : 0.f;
        sd1[p] = inside ? dL_dpixels[hw + gpx] : 0.f;
        sd2[p] = inside ? dL_dpixels[2 * hw + gpx] : 0.f;
        sdd[p] = inside ? dL_dpixel_depths[gpx] : 0.f;
        sda[p] = inside ? dL_dalphas[gpx] : 0.f;
    }
    float2 T = f2(sT[0], sT[1]);
    const float2 dp0 = f2(sd0[0], sd0[1]), dp1 = f2(sd1[0], sd1[1]), dp2 = f2(sd2[0], sd2[1]);
    const float2 dpd = f2(sdd[0], sdd[1]), dpa = f2(sda[0], sda[1]);
    // -(T_final * (bg . dL_dpixel)): the background term of dL/dalpha, multiplied by 1 / (1 - alpha) per Gaussian
    const float2 ntfb = f2(-(sT[0] * (bg0 * sd0[0] + bg1 * sd1[0] + bg2 * sd2[0])),
                           -(sT[1] * (bg0 * sd0[1] + bg1 * sd1[1] + bg2 * sd2[1])));
    const int last0 = (int)__float_as_uint(sL[0]), last1 = (int)__float_as_uint(sL[1]);
    float2 accA = f2(0.f);

    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    const int red_sub = ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    const int red_slot = red_sub < 3 ? ((lane >> 4) & 1) * 6 + ((lane >> 3) & 1) * 3 + red_sub : 12;  // 12 = pad lane
    // Slots 2 and 11 carry sum |px_| and sum |py_|: both go to record component 2 (the |gx| + |gy| channel), scaled by
    // 0.5 W and 0.5 H here -- once per (warp, Gaussian) instead of two packed multiplies per lane.
    const float slot_scale = red_slot == 0 ? -ddelx_dx : red_slot == 1 ? -ddely_dy : red_slot == 2 ? ddelx_dx :
                             red_slot == 11 ? ddely_dy : (red_slot >= 3 && red_slot <= 5) ? -0.5f : 1.0f;
    const int red_comp = red_slot == 11 ? 2 : red_slot;

    const int wmax = __reduce_max_sync(0xffffffffu, max(last0, last1));
    if (lane == 0) s_maxlast[warp] = wmax;
    __syncthreads();
    int tile_last = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) tile_last = max(tile_last, s_maxlast[w]);
    tile_last = min(tile_last, n_inst);

    auto stage = [&](int buf, int top, const uint32_t (&id)[RPT]) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int t = tid + r * NT;
            if (top - 1 - t >= 0) {
                const float4* src = reinterpret_cast<const float4*>(rec + id[r]);
                float4* d = &s_rec2[buf][3 * t];
                cp_async16(d, src); cp_async16(d + 1, src + 1); cp_async16(d + 2, src + 2);
                s_id2[buf][t] = id[r];
            }
        }
        cp_async_commit();
    };
    auto fetch_id = [&](int top, uint32_t (&id)[RPT]) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int t = tid + r * NT;
            id[r] = top - 1 - t >= 0 ? point_list[range.x + top - 1 - t] : 0u;
        }
    };
    uint32_t id_next[RPT];
    fetch_id(tile_last, id_next);
    stage(0, tile_last, id_next);
    fetch_id(tile_last - BATCH, id_next);

    for (int top = tile_last, it = 0; top > 0; top -= BATCH, ++it) {
        const int cnt = min(BATCH, top);
        cp_async_wait_all();
        __syncthreads();
        const float4* s_rec = s_rec2[it & 1];
        const uint32_t* s_id = s_id2[it & 1];
        stage((it + 1) & 1, top - BATCH, id_next);
        fetch_id(top - 2 * BATCH, id_next);
        if (wmax <= top - cnt) continue;

        uint16_t* q = s_q[warp];
        const char* rec_base = reinterpret_cast<const char*>(s_rec);
        int n_q = 0;
        for (int g0 = 0; g0 < cnt; g0 += 32) {
            const int j = g0 + lane;
            const bool hit = j < cnt && (top - 1 - j) < wmax &&
                             footprint_hits_exact(s_rec[3 * j], s_rec[3 * j + 1], bx_lo, bx_hi, by_lo, by_hi);
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            if (hit) q[n_q + __popc(m & ((1u << lane) - 1u))] = (uint16_t)j;
            n_q += __popc(m);
        }
        __syncwarp();
        for (int qi = 0; qi < n_q; ++qi) {
            const int k = (int)q[qi];
            const float4* rk = reinterpret_cast<const float4*>(rec_base + (uint32_t)k * 48u);
            const int pos = top - 1 - k;
            const float4 a = rk[0];
            const float4 b = rk[1];
            const uint32_t gid = s_id[k];
            const float dx = fadd(-pxf, a.x);
            const float2 dy = fadd2(f2(a.y), npy);
            const float2 dxAB = fmul2(f2(dx), f2(b.x, b.y));
            const float2 tc = fmul2(dy, fmul2(dy, f2(b.z)));
            const float2 tb = fmul2(dy, f2(dxAB.y));
            const float2 power = ffma2(ffma2(f2(dx), f2(dxAB.x), tc), f2(-0.5f), f2(-tb.x, -tb.y));
            const bool mb0 = (pos < last0) && !(power.x > 0.0f) && !(power.x < a.w);
            const bool mb1 = (pos < last1) && !(power.y > 0.0f) && !(power.y < a.w);
            if (!__any_sync(0xffffffffu, mb0 || mb1)) continue;
            const float2 G = expf2_exact(power);
            float2 al = fmul2(f2(b.w), G);
            al = f2(fminf(0.99f, al.x), fminf(0.99f, al.y));
            const bool ac0 = mb0 && (al.x >= 1.0f / 255.0f), ac1 = mb1 && (al.y >= 1.0f / 255.0f);
            if (!__any_sync(0xffffffffu, ac0 || ac1)) continue;

            const float4 c = rk[2];
            const float2 ae = f2(ac0 ? al.x : 0.0f, ac1 ? al.y : 0.0f);
            const float2 Ge = f2(ac0 ? G.x : 0.0f, ac1 ? G.y : 0.0f);
            const float2 om = fadd2(f2(-ae.x, -ae.y), f2(1.0f));
            // one reciprocal serves T/(1-a) and T_final/(1-a); 1-a lies in [0.01, 1]: the single-instruction
            // approximation needs no range fix-up, and rcp(1) = 1 exactly keeps a masked pixel's T untouched
            float2 inv;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.x) : "f"(om.x));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.y) : "f"(om.y));
            T = fmul2(T, inv);
            const float2 w_at = fmul2(ae, T);
            float2 s_k = ffma2(f2(c.x), dp0, dpa);
            s_k = ffma2(f2(c.y), dp1, s_k);
            s_k = ffma2(f2(c.z), dp2, s_k);
            s_k = ffma2(f2(c.w), dpd, s_k);
            float2 dLo = fmul2(fadd2(s_k, f2(-accA.x, -accA.y)), T);
            dLo = ffma2(ntfb, inv, dLo);
            accA = ffma2(ae, s_k, fmul2(om, accA));  // for the next (nearer) Gaussian

            // With u = dL/dG * G * dx and v = dL/dG * G * dy the mean gradients are -(uA + vB) 0.5W and
            // -(vC + uB) 0.5H and the conic gradients -0.5 (u dx, u dy, v dy): the per-Gaussian constants
            // (0.5W, 0.5H, -0.5, the signs) are applied once to the warp sums (slot_scale), not per pixel.
            const float2 uG = fmul2(fmul2(f2(b.w), dLo), Ge);
            const float2 u = fmul2(uG, f2(dx)), v = fmul2(uG, dy);
            const float2 px_ = ffma2(u, f2(b.x), fmul2(v, f2(b.y)));
            const float2 py_ = ffma2(v, f2(b.z), fmul2(u, f2(b.y)));
            const float2 udy = fmul2(u, dy), vdy = fmul2(v, dy), gd = fmul2(Ge, dLo);
            const float2 w0 = fmul2(w_at, dp0), w1 = fmul2(w_at, dp1), w2 = fmul2(w_at, dp2), wd = fmul2(w_at, dpd);
            float vals[12];
            vals[0] = px_.x + px_.y;
            vals[1] = py_.x + py_.y;
            vals[2] = fabsf(px_.x) + fabsf(px_.y);
            vals[3] = (u.x + u.y) * dx;
            vals[4] = udy.x + udy.y;
            vals[5] = vdy.x + vdy.y;
            vals[6] = gd.x + gd.y;
            vals[7] = w0.x + w0.y;
            vals[8] = w1.x + w1.y;
            vals[9] = w2.x + w2.y;
            vals[10] = wd.x + wd.y;
            vals[11] = fabsf(py_.x) + fabsf(py_.y);
            const float mine = warp_reduce12_packed(vals, lane) * slot_scale;
            if (!(lane & 1) && red_slot < 12) atomicAdd(grad_rec + (size_t)gid * GREC + red_comp, mine);
        }
    }
    cp_async_wait_all();
}

// ---- batched ("pipelined") variant of the packed kernel -------------------------------------------------------------
// Same arithmetic per pixel.  The packed kernel above runs every queue entry as one serial chain (loads -> power ->
// vote -> expf -> vote -> recurrence -> 5-level butterfly -> RED), several hundred cycles that only other warps can
// hide; a tile-row band of a sharded frame lasts as long as its heaviest tile, whose warps end up alone on their SM
// (DESIGN section 8).  Here NB queue entries are taken at a time: the state-independent part (loads, power, expf,
// alpha: "stage P") is issued for all NB as straight-line code, then the recurrences and the butterflies of the NB
// entries follow in one basic block, so the butterfly of one entry overlaps the chain of the next.  Entries whose
// pixels all turn out inactive are not skipped (they add exact zeros and issue no RED).
template <int MINB, int NB, int NW = 4, int BATCH = BWD_BATCH>
__global__ void __launch_bounds__(32 * NW, MINB) blend_bwd_pipe_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, const Rec* __restrict__ rec, int W, int H,
    const float* __restrict__ bg_color, const float* __restrict__ alphas, const uint32_t* __restrict__ n_contrib,
    const float* __restrict__ dL_dpixels, const float* __restrict__ dL_dpixel_depths,
    const float* __restrict__ dL_dalphas, float* __restrict__ grad_rec /*[P][12]*/, int HL, int row_stride,
    int row_phase, int grads_full) {
    constexpr int NT = 32 * NW, RPT = BATCH / NT;
    __shared__ __align__(16) float4 s_rec2[2][BATCH * 3];
    __shared__ uint32_t s_id2[2][BATCH];
    __shared__ int s_maxlast[NW];
    __shared__ uint16_t s_q[NW][BATCH];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tiles_x = (W + GRPG_TILE - 1) / GRPG_TILE;
    const int sub = NW == 4 ? warp : (int)(blockIdx.x & 3u);  // NW == 1: one single-warp CTA per 8x8 block, see blend_fwd.cu
    const uint32_t tile_x = NW == 4 ? blockIdx.x : (blockIdx.x >> 2);
    const uint32_t tile = blockIdx.y * tiles_x + tile_x;
    const int bx0 = (int)tile_x * GRPG_TILE + (sub & 1) * 8;
    const int wy0 = (sub >> 1) * 8;
    const int by0 = (blockIdx.y * row_stride + row_phase) * GRPG_TILE + wy0;
    const int pix_x = bx0 + (lane & 7);
    const int row0 = by0 + 2 * (lane >> 3);
    const float pxf = (float)pix_x;
    const float2 npy = f2(-(float)row0, -(float)(row0 + 1));
    const float bx_lo = (float)bx0, bx_hi = (float)(bx0 + 7), by_lo = (float)by0, by_hi = (float)(by0 + 7);
    const size_t hw = grads_full ? (size_t)H * W : (size_t)HL * W;

    const uint2 range = ranges[tile];
    const int n_inst = (int)(range.y - range.x);
    const float bg0 = bg_color[0], bg1 = bg_color[1], bg2 = bg_color[2];

    float sT[2], sL[2], sd0[2], sd1[2], sd2[2], sdd[2], sda[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int pix_y = row0 + p;
        const int loc_y = blockIdx.y * GRPG_TILE + wy0 + 2 * (lane >> 3) + p;
        const bool inside = pix_x < W && pix_y < H;
        const size_t pid = (size_t)loc_y * W + pix_x;
        const size_t gpx = grads_full ? (size_t)pix_y * W + pix_x : pid;
        sT[p] = inside ? 1.0f - alphas[pid] : 0.0f;  // T_final (backward.cu:468)
        sL[p] = inside ? __uint_as_float(n_contrib[pid]) : 0.0f;
        sd0[p] = inside ? dL_dpixels[gpx] : 0.f;
        sd1[p] = inside ? dL_dpixels[hw + gpx] : 0.f;
        sd2[p] = inside ? dL_dpixels[2 * hw + gpx] : 0.f;
        sdd[p] = inside ? dL_dpixel_depths[gpx] : 0.f;
        sda[p] = inside ? dL_dalphas[gpx] : 0.f;
    }
    float2 T = f2(sT[0], sT[1]);
    const float2 dp0 = f2(sd0[0], sd0[1]), dp1 = f2(sd1[0], sd1[1]), dp2 = f2(sd2[0], sd2[1]);
    const float2 dpd = f2(sdd[0], sdd[1]), dpa = f2(sda[0], sda[1]);
    const float2 ntfb = f2(-(sT[0] * (bg0 * sd0[0] + bg1 * sd1[0] + bg2 * sd2[0])),
                           -(sT[1] * (bg0 * sd0[1] + bg1 * sd1[1] + bg2 * sd2[1])));
    const int last0 = (int)__float_as_uint(sL[0]), last1 = (int)__float_as_uint(sL[1]);
    float2 accA = f2(0.f);

    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    const int red_sub = ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    const int red_slot = red_sub < 3 ? ((lane >> 4) & 1) * 6 + ((lane >> 3) & 1) * 3 + red_sub : 12;  // 12 = pad lane
    const float slot_scale = red_slot == 0 ? -ddelx_dx : red_slot == 1 ? -ddely_dy : red_slot == 2 ? ddelx_dx :
                             red_slot == 11 ? ddely_dy : (red_slot >= 3 && red_slot <= 5) ? -0.5f : 1.0f;
    const int red_comp = red_slot == 11 ? 2 : red_slot;
    const bool red_lane = !(lane & 1) && red_slot < 12;

    const int wmax = __reduce_max_sync(0xffffffffu, max(last0, last1));
    if (lane == 0) s_maxlast[warp] = wmax;
    __syncthreads();
    int tile_last = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) tile_last = max(tile_last, s_maxlast[w]);
    tile_last = min(tile_last, n_inst);

    auto stage = [&](int buf, int top, const uint32_t (&id)[RPT]) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int t = tid + r * NT;
            if (top - 1 - t >= 0) {
                const float4* src = reinterpret_cast<const float4*>(rec + id[r]);
                float4* d = &s_rec2[buf][3 * t];
                cp_async16(d, src); cp_async16(d + 1, src + 1); cp_async16(d + 2, src + 2);
                s_id2[buf][t] = id[r];
            }
        }
        cp_async_commit();
    };
    auto fetch_id = [&](int top, uint32_t (&id)[RPT]) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int t = tid + r * NT;
            id[r] = top - 1 - t >= 0 ? point_list[range.x + top - 1 - t] : 0u;
        }
    };
    uint32_t id_next[RPT];
    fetch_id(tile_last, id_next);
    stage(0, tile_last, id_next);
    fetch_id(tile_last - BATCH, id_next);

    for (int top = tile_last, it = 0; top > 0; top -= BATCH, ++it) {
        const int cnt = min(BATCH, top);
        cp_async_wait_all();
        __syncthreads();
        const float4* s_rec = s_rec2[it & 1];
        const uint32_t* s_id = s_id2[it & 1];
        stage((it + 1) & 1, top - BATCH, id_next);
        fetch_id(top - 2 * BATCH, id_next);
        if (wmax <= top - cnt) continue;

        uint16_t* q = s_q[warp];
        const char* rec_base = reinterpret_cast<const char*>(s_rec);
        int n_q = 0;
        for (int g0 = 0; g0 < cnt; g0 += 32) {
            const int j = g0 + lane;
            const bool hit = j < cnt && (top - 1 - j) < wmax &&
                             footprint_hits_exact(s_rec[3 * j], s_rec[3 * j + 1], bx_lo, bx_hi, by_lo, by_hi);
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            if (hit) q[n_q + __popc(m & ((1u << lane) - 1u))] = (uint16_t)j;
            n_q += __popc(m);
        }
        __syncwarp();
        for (int i0 = 0; i0 < n_q; i0 += NB) {
            // stage P (state-independent).  A slot past the end of the queue replays the last entry, inactive.
            float2 ae[NB], Ge[NB], dy[NB];
            float dx[NB];
            float4 b[NB], c[NB];
            uint32_t gid[NB];
            bool any_active[NB];
#pragma unroll
            for (int e = 0; e < NB; ++e) {
                const bool valid = i0 + e < n_q;
                const int k = (int)q[valid ? i0 + e : n_q - 1];
                const float4* rk = reinterpret_cast<const float4*>(rec_base + (uint32_t)k * 48u);
                const int pos = top - 1 - k;
                const float4 a = rk[0];
                b[e] = rk[1];
                c[e] = rk[2];
                gid[e] = s_id[k];
                dx[e] = fadd(-pxf, a.x);
                dy[e] = fadd2(f2(a.y), npy);
                const float2 dxAB = fmul2(f2(dx[e]), f2(b[e].x, b[e].y));
                const float2 tc = fmul2(dy[e], fmul2(dy[e], f2(b[e].z)));
                const float2 tb = fmul2(dy[e], f2(dxAB.y));
                const float2 power = ffma2(ffma2(f2(dx[e]), f2(dxAB.x), tc), f2(-0.5f), f2(-tb.x, -tb.y));
                const bool mb0 = valid && (pos < last0) && !(power.x > 0.0f) && !(power.x < a.w);
                const bool mb1 = valid && (pos < last1) && !(power.y > 0.0f) && !(power.y < a.w);
                const float2 G = expf2_exact(power);
                float2 al = fmul2(f2(b[e].w), G);
                al = f2(fminf(0.99f, al.x), fminf(0.99f, al.y));
                const bool ac0 = mb0 && (al.x >= 1.0f / 255.0f), ac1 = mb1 && (al.y >= 1.0f / 255.0f);
                ae[e] = f2(ac0 ? al.x : 0.0f, ac1 ? al.y : 0.0f);
                Ge[e] = f2(ac0 ? G.x : 0.0f, ac1 ? G.y : 0.0f);
                any_active[e] = __any_sync(0xffffffffu, ac0 || ac1);
            }
            // stage B: recurrence + reduction, entry by entry (back to front), one basic block: the REDs (a branch each)
            // follow after the last butterfly
            float mine[NB];
#pragma unroll
            for (int e = 0; e < NB; ++e) {
                const float2 om = fadd2(f2(-ae[e].x, -ae[e].y), f2(1.0f));
                float2 inv;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.x) : "f"(om.x));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.y) : "f"(om.y));
                T = fmul2(T, inv);
                const float2 w_at = fmul2(ae[e], T);
                float2 s_k = ffma2(f2(c[e].x), dp0, dpa);
                s_k = ffma2(f2(c[e].y), dp1, s_k);
                s_k = ffma2(f2(c[e].z), dp2, s_k);
                s_k = ffma2(f2(c[e].w), dpd, s_k);
                float2 dLo = fmul2(fadd2(s_k, f2(-accA.x, -accA.y)), T);
                dLo = ffma2(ntfb, inv, dLo);
                accA = ffma2(ae[e], s_k, fmul2(om, accA));  // for the next (nearer) Gaussian

                const float2 uG = fmul2(fmul2(f2(b[e].w), dLo), Ge[e]);
                const float2 u = fmul2(uG, f2(dx[e])), v = fmul2(uG, dy[e]);
                const float2 px_ = ffma2(u, f2(b[e].x), fmul2(v, f2(b[e].y)));
                const float2 py_ = ffma2(v, f2(b[e].z), fmul2(u, f2(b[e].y)));
                const float2 udy = fmul2(u, dy[e]), vdy = fmul2(v, dy[e]), gd = fmul2(Ge[e], dLo);
                const float2 w0 = fmul2(w_at, dp0), w1 = fmul2(w_at, dp1), w2 = fmul2(w_at, dp2), wd = fmul2(w_at, dpd);
                float vals[12];
                vals[0] = px_.x + px_.y;
                vals[1] = py_.x + py_.y;
                vals[2] = fabsf(px_.x) + fabsf(px_.y);
                vals[3] = (u.x + u.y) * dx[e];
                vals[4] = udy.x + udy.y;
                vals[5] = vdy.x + vdy.y;
                vals[6] = gd.x + gd.y;
                vals[7] = w0.x + w0.y;
                vals[8] = w1.x + w1.y;
                vals[9] = w2.x + w2.y;
                vals[10] = wd.x + wd.y;
                vals[11] = fabsf(py_.x) + fabsf(py_.y);
                mine[e] = warp_reduce12_packed(vals, lane) * slot_scale;
            }
#pragma unroll
            for (int e = 0; e < NB; ++e)
                if (red_lane && any_active[e]) atomicAdd(grad_rec + (size_t)gid[e] * GREC + red_comp, mine[e]);
        }
    }
    cp_async_wait_all();
}

// pixels per lane of the S == 0 backward blend (1 = blend_bwd_kernel<0>); GRPG_BWD_PPL overrides for A/B runs
static int bwd_pixels_per_lane() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GRPG_BWD_PPL");
        v = e ? atoi(e) : GRPG_BWD_PPL_DEFAULT;
        if (v != 1 && v != 2 && v != 3) v = GRPG_BWD_PPL_DEFAULT;
    }
    return v;
}
// GRPG_BWD_PIPE: 0 = blend_bwd_packed_kernel always, 1 = the batched kernel always, 2 = the batched kernel for
// tile-row bands only (sharded frames); GRPG_BWD_PIPE_CFG = "<min blocks><entries per batch>" (42, 52, 62, 44)
#define GRPG_BWD_PIPE_DEFAULT 2  // bands: 0.238 -> 0.207 ms (band of 8), 0.313 -> 0.292 (of 4), 0.439 -> 0.426 (of 2); whole frame: no gain
#define GRPG_BWD_PIPE_CFG_DEFAULT 52
int blend_split_mode();  // blend_fwd.cu
static int bwd_pipe_mode() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GRPG_BWD_PIPE");
        v = e ? atoi(e) : GRPG_BWD_PIPE_DEFAULT;
    }
    return v;
}
static int bwd_pipe_cfg() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GRPG_BWD_PIPE_CFG");
        v = e ? atoi(e) : GRPG_BWD_PIPE_CFG_DEFAULT;
    }
    return v;
}
static int bwd_min_blocks() {  // A/B knob of the packed kernel's occupancy target (GRPG_BWD_MINB = 6 | 7 | 8)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GRPG_BWD_MINB");
        v = e ? atoi(e) : 7;
    }
    return v;
}

void launch_blend_bwd(const grpg_backward_args* a, const uint2* ranges, const uint32_t* point_list, const Rec* rec,
                      const uint32_t* n_contrib, float* grad_rec, cudaStream_t stream) {
    const int stride = a->tile_row_stride > 1 ? a->tile_row_stride : 1, phase = a->tile_row_stride > 1 ? a->tile_row_phase : 0;
    const int HL = band_height(a->height, stride, phase);
    const dim3 grid((a->width + GRPG_TILE - 1) / GRPG_TILE, band_rows(a->height, stride, phase), 1);
    if (grid.y == 0) return;
    const int S = a->S;
    const int gfull = (stride > 1 && a->pixel_grads_full_frame) ? 1 : 0;
    ProfScope ps("blend_bwd", stream);
#define GRPG_BWD_LAUNCH(SBV)                                                                                      \
    blend_bwd_kernel<SBV><<<grid, 256, 0, stream>>>(ranges, point_list, rec, a->semantics, S, a->width, a->height,  \
                                                     a->background, a->alphas, n_contrib, a->dL_dpix, a->dL_dpix_depth, \
                                                     a->dL_dalphas, a->dL_dpix_semantic, grad_rec, a->dL_dsemantic, HL, stride, phase, gfull)
    const int ppl = S == 0 ? bwd_pixels_per_lane() : 1;
#define GRPG_BWD_WIDE(PPLV, MINBV)                                                                                         \
    blend_bwd_wide_kernel<PPLV, MINBV><<<grid, 256 / PPLV, 0, stream>>>(ranges, point_list, rec, a->width, a->height,      \
                                                                 a->background, a->alphas, n_contrib, a->dL_dpix,    \
                                                                 a->dL_dpix_depth, a->dL_dalphas, grad_rec, HL, stride, phase, gfull)
#define GRPG_BWD_PACKED(MINBV)                                                                                   \
    blend_bwd_packed_kernel<MINBV><<<grid, 128, 0, stream>>>(ranges, point_list, rec, a->width, a->height, a->background, \
                                                             a->alphas, n_contrib, a->dL_dpix, a->dL_dpix_depth,      \
                                                             a->dL_dalphas, grad_rec, HL, stride, phase, gfull)
    if (ppl == 3 && (bwd_pipe_mode() == 1 || (bwd_pipe_mode() == 2 && stride > 1))) {
#define GRPG_BWD_PIPE(MINBV, NBV)                                                                                         \
    blend_bwd_pipe_kernel<MINBV, NBV><<<grid, 128, 0, stream>>>(ranges, point_list, rec, a->width, a->height, a->background, \
                                                                a->alphas, n_contrib, a->dL_dpix, a->dL_dpix_depth,        \
                                                                a->dL_dalphas, grad_rec, HL, stride, phase, gfull)
        if (blend_split_mode() == 1 || (blend_split_mode() == 2 && stride > 1)) {
            blend_bwd_pipe_kernel<24, 2, 1, 64><<<dim3(grid.x * 4, grid.y, 1), 32, 0, stream>>>(
                ranges, point_list, rec, a->width, a->height, a->background, a->alphas, n_contrib, a->dL_dpix,
                a->dL_dpix_depth, a->dL_dalphas, grad_rec, HL, stride, phase, gfull);
        } else
        switch (bwd_pipe_cfg()) {
            case 42: GRPG_BWD_PIPE(4, 2); break;
            case 62: GRPG_BWD_PIPE(6, 2); break;
            case 44: GRPG_BWD_PIPE(4, 4); break;
            default: GRPG_BWD_PIPE(5, 2); break;
        }
#undef GRPG_BWD_PIPE
    } else if (ppl == 3 && (blend_split_mode() == 1 || (blend_split_mode() == 2 && stride > 1))) {
        // one single-warp CTA per 8x8 block (see blend_fwd.cu)
        blend_bwd_packed_kernel<28, 1, 64><<<dim3(grid.x * 4, grid.y, 1), 32, 0, stream>>>(
            ranges, point_list, rec, a->width, a->height, a->background, a->alphas, n_contrib, a->dL_dpix, a->dL_dpix_depth,
            a->dL_dalphas, grad_rec, HL, stride, phase, gfull);
    } else if (ppl == 3) {
        const int mb = bwd_min_blocks();
        if (mb == 6) GRPG_BWD_PACKED(6);
        else if (mb == 8) GRPG_BWD_PACKED(8);
        else GRPG_BWD_PACKED(7);
    } else if (ppl == 2) GRPG_BWD_WIDE(2, 7);  // 72 registers, 7 CTAs of 4 warps per SM (measured 6: 0.911 ms, 7: 0.912, 8: 0.931)
    else if (S == 0) GRPG_BWD_LAUNCH(0);
    else if (S <= 4) GRPG_BWD_LAUNCH(4);
    else if (S <= 8) GRPG_BWD_LAUNCH(8);
    else if (S <= 16) GRPG_BWD_LAUNCH(16);
    else GRPG_BWD_LAUNCH(32);
#undef GRPG_BWD_LAUNCH
#undef GRPG_BWD_WIDE
#undef GRPG_BWD_PACKED
}

}  // namespace grpg
