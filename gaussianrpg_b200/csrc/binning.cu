// binning.cu -- ordering of Gaussian/tile instances.
//
// The reference emits one 64-bit key ((tile << 32) | depth bits) per Gaussian/tile overlap
// (duplicateWithKeys, rasterizer_impl.cu:70-111) and runs a stable LSD radix sort over the
// low 32+bit bits of R such pairs (rasterizer_impl.cu:303-311): 6 digit passes over 12 B * R.
//
// B200 redesign (same resulting order, ~5x fewer bytes): the depth ordering is a property of
// the Gaussian, not of the instance.  So
//   1. sort the P Gaussians once by (depth bits, id)          -- 4 passes over 8 B * P
//   2. prefix-sum tile counts in that order                    -- replaces InclusiveSum (:280)
//   3. emit instances in depth order, balanced over warps      -- replaces duplicateWithKeys
//   4. stable-sort instances by tile id only (14 bits at 1920x1280 -> 2 passes over 8 B * R)
// A stable sort by tile of a depth-ordered sequence is exactly the (tile, depth, id) order of
// the reference; ties on equal depth keep ascending Gaussian id in both.
#include "radix_sort.cuh"

namespace grpg {

// ---- 2. exclusive scan of tiles_touched in sorted order (decoupled look-back) ------------
constexpr int SCAN_IPT = 8;
constexpr int SCAN_TILE = 256 * SCAN_IPT;
constexpr uint32_t EMIT_CHUNK = 1024;                      // instances emitted per warp
constexpr uint32_t EMIT_MAX_CHUNKS = (1u << 31) / EMIT_CHUNK + 2;  // R < 2^31 (the reference's int num_rendered) is enforced by the API
constexpr unsigned long long SC_FLAG_AGG = 1ull << 62;
constexpr unsigned long long SC_FLAG_PREFIX = 2ull << 62;
constexpr unsigned long long SC_FLAG_MASK = 3ull << 62;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(256) scan_tiles_kernel(const uint32_t* __restrict__ sorted_idx,
                                                         const uint32_t* __restrict__ tiles_touched, uint32_t P_cap,
                                                         uint32_t* __restrict__ offsets,
                                                         unsigned long long* __restrict__ status /*[tiles+1]*/,
                                                         uint32_t* __restrict__ tile_counter,
                                                         unsigned long long* __restrict__ total_out,
                                                         const unsigned long long* __restrict__ ref_instances,
                                                         uint32_t* __restrict__ chunk_start,
                                                         unsigned long long capacity) {
    __shared__ uint32_t s_scan[8];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_excl;
    // total_out = the forward's device-side counters: [0] instances to bin, [1] the reference's num_rendered,
    // [2] overflow flag of the static-capacity mode, [3] Gaussians in the depth order (written by the depth sort's
    // compacting first pass; the grid of this kernel covers all P_cap, tiles past the count leave at once)
    const uint32_t P = (uint32_t)min((unsigned long long)P_cap, total_out[3]);
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    if (P == 0) {  // nothing visible: the first ticket publishes the (zero) totals
        if (tile == 0 && threadIdx.x == 0) { total_out[0] = 0; total_out[1] = *ref_instances; total_out[2] = 0; }
        return;
    }
    if ((unsigned long long)tile * SCAN_TILE >= P) return;
    const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_IPT;
    uint32_t v[SCAN_IPT];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i) {
        uint32_t j = base + i;
        // the last depth-sort pass left tiles_touched[sorted_idx[j]] in offsets[j]; it is overwritten below with
        // the exclusive prefix (each CTA reads its own range before writing it)
        v[i] = j < P ? offsets[j] : 0u;
        sum += v[i];
    }
    uint32_t block_total;
    uint32_t texcl = block_exclusive_scan_256(sum, s_scan, &block_total);
    // warp-parallel decoupled look-back: all ~1000 tiles of a 2 M-Gaussian scan are resident at once, so a
    // one-predecessor-per-step walk would be a serial chain; warp 0 inspects 32 predecessors per step instead
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        if (lane == 0)
            st_volatile_u64(status + tile, (unsigned long long)block_total | (tile == 0 ? SC_FLAG_PREFIX : SC_FLAG_AGG));
        unsigned long long excl = 0;
        int t = (int)tile - 1;
        while (t >= 0) {
            const int idx = t - lane;
            unsigned long long sv = idx >= 0 ? ld_volatile_u64(status + idx) : SC_FLAG_PREFIX;  // before tile 0: prefix 0
            while (__any_sync(0xffffffffu, (sv & SC_FLAG_MASK) == 0)) {
                if ((sv & SC_FLAG_MASK) == 0) sv = ld_volatile_u64(status + idx);
            }
            const uint32_t pm = __ballot_sync(0xffffffffu, (sv & SC_FLAG_MASK) == SC_FLAG_PREFIX);
            const int first = pm ? __ffs(pm) - 1 : 32;  // nearest predecessor that already holds a full prefix
            uint32_t v = lane <= first ? (uint32_t)(sv & ~SC_FLAG_MASK) : 0u;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            excl += v;
            if (pm) break;
            t -= 32;
        }
        if (lane == 0) {
            if (tile > 0) st_volatile_u64(status + tile, (excl + block_total) | SC_FLAG_PREFIX);
            s_excl = excl;
            if ((tile + 1) * (unsigned long long)SCAN_TILE >= P) {
                total_out[0] = excl + block_total;  // instances to bin
                total_out[1] = *ref_instances;       // instances the reference would bin (its num_rendered)
                total_out[2] = (capacity != 0 && excl + block_total > capacity) ? 1ull : 0ull;
            }
        }
    }
    __syncthreads();
    uint32_t run = (uint32_t)s_excl + texcl;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i) {
        uint32_t j = base + i;
        if (j < P) {
            offsets[j] = run;
            // every multiple of EMIT_CHUNK inside [run, run + v) starts in sorted position j
            for (uint32_t kb = (run + EMIT_CHUNK - 1) / EMIT_CHUNK; (unsigned long long)kb * EMIT_CHUNK < (unsigned long long)run + v[i]; ++kb)
                if (kb < EMIT_MAX_CHUNKS) chunk_start[kb] = j;
        }
        run += v[i];
    }
}

// position of the (n+1)-th set bit of `mask` (n < popc(mask)): five popc steps, no data-dependent loop
__device__ __forceinline__ uint32_t nth_set_bit(uint32_t mask, uint32_t n) {
    uint32_t pos = 0;
    uint32_t c = (uint32_t)__popc(mask & 0xFFFFu);
    if (n >= c) { pos = 16; n -= c; }
    c = (uint32_t)__popc((mask >> pos) & 0xFFu);
    if (n >= c) { pos += 8; n -= c; }
    c = (uint32_t)__popc((mask >> pos) & 0xFu);
    if (n >= c) { pos += 4; n -= c; }
    c = (uint32_t)__popc((mask >> pos) & 0x3u);
    if (n >= c) { pos += 2; n -= c; }
    if (n >= ((mask >> pos) & 1u)) pos += 1;
    return pos;
}

// ---- 3. instance emission ------------------------------------------------------------------
// Output-balanced: warp k writes instances [k*EMIT_CHUNK, (k+1)*EMIT_CHUNK) no matter how they are
// distributed over Gaussians (a splat covering the whole screen is shared by many warps; the
// reference gives all of its tiles to one thread, rasterizer_impl.cu:98-109).  The scan kernel
// recorded the depth-sorted position that owns each chunk boundary, so no search is needed.
template <bool TILE_MASK>  // rectangles of <= 32 tiles carry a mask of the tiles to bin (GRPG_EXACT_TILE_CULL=1)
__global__ void __launch_bounds__(256) emit_instances_kernel(const uint32_t* __restrict__ sorted_idx,
                                                             const uint32_t* __restrict__ offsets,
                                                             const uint2* __restrict__ rect,
                                                             const uint32_t* __restrict__ tile_mask,
                                                             const uint32_t* __restrict__ chunk_start, uint32_t P,
                                                             uint32_t R, const unsigned long long* __restrict__ counts,
                                                             uint32_t grid_x,
                                                             uint32_t* __restrict__ inst_tile,
                                                             uint32_t* __restrict__ inst_gauss,
                                                             uint32_t* __restrict__ tile_hist /*[256][sort tiles]*/,
                                                             uint32_t sort_tiles, uint32_t digit_mask) {
    // A CTA writes 8 x 1024 consecutive instances = exactly two 4096-key tiles of the tile-id sort, so the
    // digit histograms of that sort's first pass are accumulated here instead of re-reading the keys.
    __shared__ uint32_t s_hist[2][256];
    s_hist[0][threadIdx.x] = 0;
    s_hist[1][threadIdx.x] = 0;
    __syncthreads();
    if (counts != nullptr) {  // static-capacity mode: the counts live on the device, R and P are capacities
        R = (uint32_t)min((unsigned long long)R, counts[0]);
        P = (uint32_t)min((unsigned long long)P, counts[3]);
    }
    const int lane = threadIdx.x & 31;
    const int half = (threadIdx.x >> 5) >> 2;  // warps 0-3 -> first sort tile, 4-7 -> second
    const uint32_t chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long lo64 = (unsigned long long)chunk * EMIT_CHUNK;
    uint32_t cur = lo64 < R ? (uint32_t)lo64 : R;
    const uint32_t hi = lo64 < R ? min(R, cur + EMIT_CHUNK) : R;
    uint32_t pos = lo64 < R ? chunk_start[chunk] : P;
    // one depth-sorted Gaussian per lane: id, first instance, rectangle (a dependent gather chain: sorted_idx -> rect)
    struct Item { uint32_t g, off, x0, y0, w, end, mask; };
    auto load_item = [&](uint32_t at) {
        Item it{0u, 0xFFFFFFFFu, 0u, 0u, 1u, 0u, 0xFFFFFFFFu};
        const uint32_t p = at + lane;
        if (at < P && p < P) {
            it.g = sorted_idx[p];
            it.off = offsets[p];
            const uint2 r = rect[it.g];
            it.x0 = r.x & 0xffffu; it.y0 = r.y & 0xffffu;
            it.w = (r.x >> 16) - it.x0;
            const uint32_t area = it.w * ((r.y >> 16) - it.y0);
            // rectangles of <= 32 tiles carry a mask of the tiles that can hold a visible pixel: only those are binned
            if (TILE_MASK) it.mask = area <= 32u ? tile_mask[it.g] : 0xFFFFFFFFu;
            it.end = it.off + ((TILE_MASK && it.mask != 0xFFFFFFFFu) ? (uint32_t)__popc(it.mask) : area);
            if (it.w == 0) it.w = 1;
        }
        return it;
    };
    // The next group of 32 Gaussians is requested before the current one is expanded: with few CTAs per SM (the band
    // of a sharded frame emits an eighth of the instances) the kernel is a chain of gather latencies otherwise
    // (measured 0.062 ms for 1.2 M instances against 0.072 ms for 8.6 M).
    Item nxt = load_item(pos);
    while (cur < hi && pos < P) {
        const Item itm = nxt;
        nxt = load_item(pos + 32);
        const uint32_t g = itm.g, off = itm.off, x0 = itm.x0, y0 = itm.y0, w = itm.w, end = itm.end, mask = itm.mask;
        const uint32_t group_end = __reduce_max_sync(0xffffffffu, end);
        const uint32_t stop = min(hi, group_end);
        for (uint32_t k0 = cur; k0 < stop; k0 += 32) {
            const uint32_t k = k0 + lane;
            int owner = 0;  // last lane whose offset <= k (offsets are non-decreasing across lanes)
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const uint32_t probe = __shfl_sync(0xffffffffu, off, (owner + step) & 31);
                if (probe <= k) owner += step;
            }
            const uint32_t o_off = __shfl_sync(0xffffffffu, off, owner);
            const uint32_t o_w = __shfl_sync(0xffffffffu, w, owner);
            const uint32_t o_x0 = __shfl_sync(0xffffffffu, x0, owner);
            const uint32_t o_y0 = __shfl_sync(0xffffffffu, y0, owner);
            const uint32_t o_g = __shfl_sync(0xffffffffu, g, owner);
            const uint32_t o_mask = TILE_MASK ? __shfl_sync(0xffffffffu, mask, owner) : 0xFFFFFFFFu;
            if (k < stop) {
                uint32_t m = k - o_off;
                if (TILE_MASK && o_mask != 0xFFFFFFFFu) m = nth_set_bit(o_mask, m);  // the (m+1)-th binned tile of the rectangle
                // m / o_w through a float estimate and one correction step (exact below 2^22; a 32-bit integer
                // division costs ~20 instructions per instance)
                uint32_t ty;
                if (m < (1u << 22)) {
                    ty = (uint32_t)((float)m * __frcp_rn((float)o_w));
                    const int rem = (int)(m - ty * o_w);
                    if (rem < 0) --ty;
                    else if (rem >= (int)o_w) ++ty;
                } else {
                    ty = m / o_w;
                }
                const uint32_t tx = m - ty * o_w;
                const uint32_t tile_id = (o_y0 + ty) * grid_x + (o_x0 + tx);
                inst_tile[k] = tile_id;
                inst_gauss[k] = o_g;
                atomicAdd(&s_hist[half][tile_id & digit_mask], 1u);
            }
        }
        if (stop > cur) cur = stop;
        pos += 32;
    }
    __syncthreads();
    const uint32_t t0 = blockIdx.x * 2;
    if (t0 < sort_tiles) tile_hist[(size_t)threadIdx.x * sort_tiles + t0] = s_hist[0][threadIdx.x];
    if (t0 + 1 < sort_tiles) tile_hist[(size_t)threadIdx.x * sort_tiles + t0 + 1] = s_hist[1][threadIdx.x];
}

// ---- 5. tile ranges (identifyTileRanges, rasterizer_impl.cu:116-138) -----------------------
__global__ void __launch_bounds__(256) tile_ranges_kernel(const uint32_t* __restrict__ tile_keys, uint32_t R,
                                                          uint2* __restrict__ ranges,
                                                          const unsigned long long* __restrict__ counts) {
    if (counts != nullptr) R = (uint32_t)min((unsigned long long)R, counts[0]);
    // 4 keys per thread (one 16-byte load) + the key before them
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= R) return;
    uint32_t k[4];
    int n = 4;
    if (i0 + 4 <= R) {
        const uint4 v = *reinterpret_cast<const uint4*>(tile_keys + i0);
        k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
    } else {
        n = (int)(R - i0);
        for (int j = 0; j < n; ++j) k[j] = tile_keys[i0 + j];
    }
    uint32_t prev = i0 == 0 ? 0xFFFFFFFFu : tile_keys[i0 - 1];
    for (int j = 0; j < n; ++j) {
        const uint32_t i = i0 + j, cur = k[j];
        if (i == 0) ranges[cur].x = 0;
        else if (cur != prev) { ranges[prev].y = i; ranges[cur].x = i; }
        if (i == R - 1) ranges[cur].y = R;
        prev = cur;
    }
}

__global__ void __launch_bounds__(256) iota_kernel(uint32_t* __restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

__global__ void __launch_bounds__(256) reference_keys_kernel(const uint32_t* __restrict__ point_list,
                                                             const uint32_t* __restrict__ tile_keys,
                                                             const Rec* __restrict__ rec, uint32_t R,
                                                             unsigned long long* __restrict__ keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    keys[i] = ((unsigned long long)tile_keys[i] << 32) | __float_as_uint(rec[point_list[i]].c.w);
}

// ---- host orchestration ----------------------------------------------------------------------
// scratch layout inside geom_ws.scratch: [keys_b P][vals_b P][sort aux][scan status][counters]
size_t geom_scratch_bytes(int P) {
    size_t b = 0;
    b += align_up((size_t)P * 4, 256) * 2;
    b += align_up((size_t)SORT_MAX_PASSES * (256 + 64) * 4 + (size_t)SORT_MAX_PASSES * sort_num_tiles(P) * 256 * 4, 256);
    b += align_up(((size_t)(P + SCAN_TILE - 1) / SCAN_TILE + 2) * 8, 256);
    b += 256;  // scan tile counter + total
    b += align_up((size_t)EMIT_MAX_CHUNKS * 4, 256);  // chunk_start
    return b;
}
size_t binning_scratch_bytes(long long R) {
    size_t b = 0;
    b += align_up((size_t)R * 4, 256) * 2;
    b += align_up((size_t)SORT_MAX_PASSES * (256 + 64) * 4 + (size_t)SORT_MAX_PASSES * sort_num_tiles(R) * 256 * 4, 256);
    return b;
}

// offset of the [scan status | counters] block inside the geometry scratch
static size_t geom_status_offset(int P) {
    return align_up((size_t)P * 4, 256) * 2 +
           align_up((size_t)SORT_MAX_PASSES * (256 + 64) * 4 + (size_t)SORT_MAX_PASSES * sort_num_tiles(P) * 256 * 4, 256);
}
// Zeroes the scan status words and the counters (before the projection kernel, which accumulates the reference's
// instance count into the returned counter).
unsigned long long* prepare_geometry_scratch(int P, void* scratch, cudaStream_t stream) {
    char* p = (char*)scratch + geom_status_offset(P);
    const size_t scan_tiles = ((size_t)P + SCAN_TILE - 1) / SCAN_TILE;
    const size_t status_bytes = align_up((scan_tiles + 2) * 8, 256);
    cudaMemsetAsync(p, 0, status_bytes + 256, stream);
    return (unsigned long long*)(p + status_bytes + 64);  // counters block: [0] scan tile ticket, [64] ref instances
}

// Steps 1-2.  depth_key is consumed (left sorted); sorted_idx receives the permutation of the Gaussians that emit at
// least one instance (tiles_touched != 0) -- the sort's first digit pass drops the others for free, so the remaining
// passes, the scan and the emission work on the visible (and, with tile-row bands, in-band) Gaussians only.
// `counts` = the forward's device-side counters (see scan_tiles_kernel).
void run_depth_order_and_scan(int P, uint32_t* depth_key, uint32_t* sorted_idx, const uint32_t* tiles_touched,
                              uint32_t* offsets, void* scratch, unsigned long long* counts, unsigned long long capacity,
                              int num_sms, cudaStream_t stream) {
    char* p = (char*)scratch;
    uint32_t* keys_b = (uint32_t*)p; p += align_up((size_t)P * 4, 256);
    uint32_t* vals_b = (uint32_t*)p; p += align_up((size_t)P * 4, 256);
    uint32_t* aux = (uint32_t*)p;
    p += align_up((size_t)SORT_MAX_PASSES * (256 + 64) * 4 + (size_t)SORT_MAX_PASSES * sort_num_tiles(P) * 256 * 4, 256);
    unsigned long long* status = (unsigned long long*)p;
    size_t scan_tiles = ((size_t)P + SCAN_TILE - 1) / SCAN_TILE;
    p += align_up((scan_tiles + 2) * 8, 256);
    uint32_t* scan_counter = (uint32_t*)p;
    p += 256;
    uint32_t* chunk_start = (uint32_t*)p;

    // values start as the identity permutation: generated inside the first digit pass, no iota kernel
    bool in_a = onesweep_sort_pairs(depth_key, sorted_idx, keys_b, vals_b, P, 32, aux, num_sms, stream,
                                    "depth_sort_hist", "depth_sort_scan", "depth_sort_pass", /*iota_values=*/true,
                                    /*first_hist_ready=*/false, /*gather_src=*/tiles_touched, /*gather_dst=*/offsets,
                                    /*n_dev=*/nullptr, /*keep_src=*/tiles_touched, /*n_kept_out=*/counts + 3);
    if (!in_a) {  // 32 bits -> 4 passes -> always lands back in the a-buffers; kept for safety
        cudaMemcpyAsync(depth_key, keys_b, (size_t)P * 4, cudaMemcpyDeviceToDevice, stream);
        cudaMemcpyAsync(sorted_idx, vals_b, (size_t)P * 4, cudaMemcpyDeviceToDevice, stream);
    }
    // status words and counters were zeroed by prepare_geometry_scratch
    ProfScope ps("scan_tiles", stream);
    scan_tiles_kernel<<<(unsigned)scan_tiles, 256, 0, stream>>>(
        sorted_idx, tiles_touched, (uint32_t)P, offsets, status, scan_counter, counts,
        (const unsigned long long*)((const char*)scan_counter + 64), chunk_start, capacity);
}

// Steps 3-5.  Final order lands in (tile_keys, point_list).  `R` is the instance count when the host knows it
// (counts_are_exact) or the capacity of the binning workspace in static-capacity mode, where every kernel reads the
// count from `counts` on the device and the grids cover the capacity.
void run_instance_binning(int P, long long R, uint32_t grid_x, uint32_t num_tiles, const uint32_t* sorted_idx,
                          const uint32_t* offsets, const uint2* rect, const uint32_t* tile_mask, const void* geom_scratch,
                          uint32_t* tile_keys, uint32_t* point_list, void* scratch, uint2* ranges,
                          const unsigned long long* counts, int static_capacity, int num_sms, cudaStream_t stream) {
    cudaMemsetAsync(ranges, 0, (size_t)num_tiles * sizeof(uint2), stream);
    if (R <= 0) return;
    char* p = (char*)scratch;
    uint32_t* keys_b = (uint32_t*)p; p += align_up((size_t)R * 4, 256);
    uint32_t* vals_b = (uint32_t*)p; p += align_up((size_t)R * 4, 256);
    uint32_t* aux = (uint32_t*)p;
    int bits = 1;
    while (bits < 32 && (1ull << bits) < (unsigned long long)num_tiles) ++bits;
    SortPlan plan = make_sort_plan(bits);
    // choose the start buffer so that the last pass writes (tile_keys, point_list)
    const bool start_in_final = (plan.passes % 2) == 0;
    uint32_t* k0 = start_in_final ? tile_keys : keys_b;
    uint32_t* v0 = start_in_final ? point_list : vals_b;
    uint32_t* k1 = start_in_final ? keys_b : tile_keys;
    uint32_t* v1 = start_in_final ? vals_b : point_list;
    {
        ProfScope ps("emit_instances", stream);
        // chunk_start lives at the end of the geometry scratch (see run_depth_order_and_scan)
        const char* gs = (const char*)geom_scratch;
        gs += align_up((size_t)P * 4, 256) * 2;
        gs += align_up((size_t)SORT_MAX_PASSES * (256 + 64) * 4 + (size_t)SORT_MAX_PASSES * sort_num_tiles(P) * 256 * 4, 256);
        gs += align_up((((size_t)P + SCAN_TILE - 1) / SCAN_TILE + 2) * 8, 256) + 256;
        const uint32_t* chunk_start = (const uint32_t*)gs;
        static_assert(EMIT_CHUNK * 8 == 2 * SORT_TILE, "an emit CTA must cover exactly two sort tiles");
        const uint32_t sort_tiles = (uint32_t)sort_num_tiles(R);
        if (tile_mask != nullptr)
            emit_instances_kernel<true><<<(sort_tiles + 1) / 2, 256, 0, stream>>>(
                sorted_idx, offsets, rect, tile_mask, chunk_start, (uint32_t)P, (uint32_t)R, counts, grid_x, k0, v0,
                sort_first_pass_hist(aux), sort_tiles, (1u << plan.bits[0]) - 1u);
        else
            emit_instances_kernel<false><<<(sort_tiles + 1) / 2, 256, 0, stream>>>(
                sorted_idx, offsets, rect, nullptr, chunk_start, (uint32_t)P, (uint32_t)R, counts, grid_x, k0, v0,
                sort_first_pass_hist(aux), sort_tiles, (1u << plan.bits[0]) - 1u);
    }
    onesweep_sort_pairs(k0, v0, k1, v1, R, bits, aux, num_sms, stream, "tile_sort_hist", "tile_sort_scan", "tile_sort_pass",
                        /*iota_values=*/false, /*first_hist_ready=*/true, nullptr, nullptr,
                        /*n_dev=*/static_capacity ? counts : nullptr);
    ProfScope ps("tile_ranges", stream);
    tile_ranges_kernel<<<(unsigned)((R + 1023) / 1024), 256, 0, stream>>>(tile_keys, (uint32_t)R, ranges,
                                                                          static_capacity ? counts : nullptr);
}

void launch_reference_keys(long long R, const uint32_t* point_list, const uint32_t* tile_keys, const Rec* rec,
                           unsigned long long* keys, cudaStream_t stream) {
    if (R <= 0) return;
    reference_keys_kernel<<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(point_list, tile_keys, rec, (uint32_t)R, keys);
}

}  // namespace grpg
