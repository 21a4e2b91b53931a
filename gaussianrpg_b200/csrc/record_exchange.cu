// record_exchange.cu -- sparse exchange of the per-Gaussian 2D gradient records between the ranks of a tile-row sharded
// frame over NVLink peer memory (include/grpg_b200.h, grpg_exchange_*).  No reference counterpart.
#include <cstdint>
#include "grpg_common.cuh"

extern "C" int grpg_loss_fail(const char* msg);

namespace grpg {

constexpr int EX_HEADER = 256;      // bytes: [0..7] u32 counts received per source rank, [8..15] u32 send counters (local use)
constexpr int EX_ENTRY = 12;        // floats per entry: Gaussian id (bits) + the 11 record components = 48 bytes

static inline size_t ex_slice(int P, int world) { return ((size_t)P + world - 1) / world; }

struct ExPtrs { char* inbox[8]; };

// One thread per Gaussian.  In-band Gaussians (tiles_touched != 0) of a foreign slice are appended to the owner's inbox
// segment reserved for this rank.  Owners are monotone in the id, so a warp addresses one or two owners: per owner the
// warp allocates a run of consecutive slots with one atomic on a LOCAL counter, assembles the run in shared memory and
// copies it out with fully coalesced 16-byte PEER stores (a contiguous 48 x n byte range of the remote inbox) -- strided
// 16-byte pieces per lane cost three times the NVLink packets (measured: 135 -> see profiles/ for the coalesced form).
__global__ void __launch_bounds__(256) exchange_pack_kernel(int P, int world, int rank, uint32_t slice,
                                                            const uint32_t* __restrict__ tiles_touched,
                                                            const float4* __restrict__ grad_rec4, ExPtrs ptrs) {
    __shared__ __align__(16) float4 s_run[8][32 * 3];
    __shared__ uint32_t s_wcnt[8][2];  // senders per warp for the (at most) two owners a CTA of 256 consecutive ids spans
    __shared__ uint32_t s_wbase[8][2];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* send_count = reinterpret_cast<uint32_t*>(ptrs.inbox[rank]) + 8;
    const bool in_range = i < (uint32_t)P;
    const int owner = in_range ? (int)(i / slice) : -1;
    bool send = in_range && owner != rank && tiles_touched[i] != 0u;
    // Slots are allocated per CTA: consecutive CTAs address the same owner, so one atomic per WARP on the owner's
    // counter puts all resident warps on one L2 address (62 k same-address atomics at P = 2 M); per CTA it is 8 k.  A
    // CTA of 256 consecutive ids spans at most two owners when slice >= 256; smaller slices keep the per-warp atomics.
    // (Measured on 2 GPUs: 0.078 ms either way -- there the kernel is bound by the 29 MB it stores over NVLink; the
    // 8-GPU case, 14 MB per rank, was not re-measured.)
    const uint32_t c0 = blockIdx.x * blockDim.x;
    const int o0 = (int)(c0 / slice);
    const bool cta_alloc = slice >= 256u;
    if (cta_alloc) {
        const uint32_t m0 = __ballot_sync(0xffffffffu, send && owner == o0);
        const uint32_t m1 = __ballot_sync(0xffffffffu, send && owner == o0 + 1);
        if (lane == 0) { s_wcnt[warp][0] = (uint32_t)__popc(m0); s_wcnt[warp][1] = (uint32_t)__popc(m1); }
        __syncthreads();
        if (threadIdx.x < 2) {
            const int j = threadIdx.x;
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += s_wcnt[w][j];
            uint32_t base = tot ? atomicAdd(send_count + o0 + j, tot) : 0u;  // (tot != 0 implies a valid foreign owner)
#pragma unroll
            for (int w = 0; w < 8; ++w) { s_wbase[w][j] = base; base += s_wcnt[w][j]; }
        }
        __syncthreads();
    }
    uint32_t todo = __ballot_sync(0xffffffffu, send);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const int lead_owner = __shfl_sync(0xffffffffu, owner, leader);
        const bool mine = send && owner == lead_owner;
        const uint32_t peers = __ballot_sync(0xffffffffu, mine);
        const int n = __popc(peers);
        uint32_t base = 0;
        if (cta_alloc) {
            base = s_wbase[warp][lead_owner - o0];
        } else {
            if (lane == leader) base = atomicAdd(send_count + lead_owner, (uint32_t)n);
            base = __shfl_sync(0xffffffffu, base, leader);
        }
        if (mine) {
            const int k = __popc(peers & ((1u << lane) - 1u));
            const float4 a = grad_rec4[3 * (size_t)i], b = grad_rec4[3 * (size_t)i + 1], c = grad_rec4[3 * (size_t)i + 2];
            s_run[warp][3 * k] = make_float4(__uint_as_float(i), a.x, a.y, a.z);
            s_run[warp][3 * k + 1] = make_float4(a.w, b.x, b.y, b.z);
            s_run[warp][3 * k + 2] = make_float4(b.w, c.x, c.y, c.z);
            send = false;
        }
        __syncwarp();
        float4* dst = reinterpret_cast<float4*>(ptrs.inbox[lead_owner] + EX_HEADER + ((size_t)rank * slice + base) * (EX_ENTRY * 4));
        for (int t = lane; t < 3 * n; t += 32) dst[t] = s_run[warp][t];
        __syncwarp();
        todo &= ~peers;
    }
}

// publishes how many entries this rank appended to every owner's inbox (remote 4-byte stores into the headers)
__global__ void exchange_publish_kernel(int world, int rank, ExPtrs ptrs) {
    const int q = threadIdx.x;
    if (q >= world || q == rank) return;
    const uint32_t* send_count = reinterpret_cast<const uint32_t*>(ptrs.inbox[rank]) + 8;
    reinterpret_cast<uint32_t*>(ptrs.inbox[q])[rank] = send_count[q];
}

// one half-warp per inbox entry: its lanes 1..11 add the record's components (one coalesced RED per entry)
__global__ void __launch_bounds__(256) exchange_accumulate_kernel(int world, int rank, uint32_t slice, const char* __restrict__ inbox,
                                                                  float* __restrict__ grad_rec) {
    const int lane = threadIdx.x & 31, h = lane & 15, half = lane >> 4;
    const uint32_t hw = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 2 + half, n_hw = ((gridDim.x * blockDim.x) >> 5) * 2;
    const uint32_t* counts = reinterpret_cast<const uint32_t*>(inbox);
    for (int q = 0; q < world; ++q) {
        if (q == rank) continue;
        const uint32_t n = min(counts[q], slice);
        const float* seg = reinterpret_cast<const float*>(inbox + EX_HEADER + (size_t)q * slice * (EX_ENTRY * 4));
        // whole warps iterate together (the shuffle below needs every lane), two entries per step
        for (uint32_t e0 = hw - half; e0 < n; e0 += n_hw) {
            const uint32_t e = e0 + half;
            const float v = (e < n && h < 12) ? seg[(size_t)e * EX_ENTRY + h] : 0.f;
            const uint32_t id = __float_as_uint(__shfl_sync(0xffffffffu, v, half * 16));
            if (e < n && h >= 1 && h < 12) atomicAdd(grad_rec + (size_t)id * 12 + (h - 1), v);
        }
    }
}

}  // namespace grpg

using namespace grpg;

extern "C" size_t grpg_exchange_inbox_bytes(int P, int world) {
    if (P < 0 || world < 1 || world > 8) return 0;
    return EX_HEADER + ex_slice(P, world) * (size_t)world * (EX_ENTRY * 4);
}

static int ex_check(const grpg_exchange_args* a, ExPtrs& p) {
    if (!a || a->P < 0 || a->world < 1 || a->world > 8 || a->rank < 0 || a->rank >= a->world)
        return grpg_loss_fail("grpg_exchange: bad arguments");
    if (!a->grad_rec || !a->geom_ws) return grpg_loss_fail("grpg_exchange: missing buffers");
    for (int q = 0; q < a->world; ++q) {
        if (!a->inbox[q]) return grpg_loss_fail("grpg_exchange: missing inbox pointer");
        p.inbox[q] = (char*)a->inbox[q];
    }
    return 0;
}

extern "C" int grpg_exchange_pack(const grpg_exchange_args* a) {
    ExPtrs p{};
    if (int rc = ex_check(a, p)) return rc;
    if (a->P == 0 || a->world == 1) return 0;
    cudaStream_t stream = (cudaStream_t)a->stream;
    grpg_geom_layout L;
    grpg_get_geometry_layout(a->P, &L);
    const uint32_t* tiles_touched = (const uint32_t*)((const char*)a->geom_ws + L.tiles_touched);
    ProfScope ps("exchange_pack", stream);
    cudaMemsetAsync(p.inbox[a->rank] + 32, 0, 32, stream);  // this rank's send counters
    exchange_pack_kernel<<<(a->P + 255) / 256, 256, 0, stream>>>(a->P, a->world, a->rank, (uint32_t)ex_slice(a->P, a->world),
                                                                 tiles_touched, reinterpret_cast<const float4*>(a->grad_rec), p);
    exchange_publish_kernel<<<1, 32, 0, stream>>>(a->world, a->rank, p);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return grpg_loss_fail(cudaGetErrorString(e));
    return 0;
}

extern "C" int grpg_exchange_accumulate(const grpg_exchange_args* a) {
    ExPtrs p{};
    if (int rc = ex_check(a, p)) return rc;
    if (a->P == 0 || a->world == 1) return 0;
    cudaStream_t stream = (cudaStream_t)a->stream;
    ProfScope ps("exchange_accumulate", stream);
    exchange_accumulate_kernel<<<148 * 4, 256, 0, stream>>>(a->world, a->rank, (uint32_t)ex_slice(a->P, a->world),
                                                            p.inbox[a->rank], a->grad_rec);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return grpg_loss_fail(cudaGetErrorString(e));
    return 0;
}
