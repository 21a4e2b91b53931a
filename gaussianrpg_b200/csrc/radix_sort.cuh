// radix_sort.cuh -- hand-written onesweep LSD radix sort of (uint32 key, uint32 value) pairs.
//
// Replaces cub::DeviceRadixSort::SortPairs as used at rasterizer_impl.cu:306-311 of the
// reference.  One kernel per digit pass; each CTA ranks a 4096-key tile with warp-level
// match/ballot multi-split (stable), publishes its per-digit counts and resolves its global
// offsets by decoupled look-back over the preceding tiles (single read + single write of the
// data per pass).  A separate histogram kernel reads the keys once for all passes.
//
// HBM traffic per pass: 8 B read + 8 B written per pair (+ 4 B per pair once for the
// histogram); look-back state is 1 KB per tile and lives in L2.
#pragma once
#include "grpg_common.cuh"

namespace grpg {

constexpr uint32_t LB_FLAG_AGG = 1u << 30;
constexpr uint32_t LB_FLAG_PREFIX = 2u << 30;
constexpr uint32_t LB_FLAG_MASK = 3u << 30;
constexpr uint32_t LB_VALUE_MASK = ~LB_FLAG_MASK;

struct SortPlan {
    int passes;
    int bits[SORT_MAX_PASSES];
    int shift[SORT_MAX_PASSES];
};

static inline SortPlan make_sort_plan(int total_bits) {
    SortPlan p{};
    if (total_bits < 1) total_bits = 1;
    p.passes = (total_bits + 7) / 8;
    int per = (total_bits + p.passes - 1) / p.passes;
    int s = 0;
    for (int i = 0; i < p.passes; ++i) {
        int b = per;
        if (s + b > total_bits) b = total_bits - s;
        p.bits[i] = b;
        p.shift[i] = s;
        s += b;
    }
    return p;
}

// ---- histogram of every pass' digit in one read ---------------------------------------
static __global__ void __launch_bounds__(256) sort_histogram_kernel(const uint32_t* __restrict__ keys, uint32_t n,
                                                             SortPlan plan, uint32_t* __restrict__ hist /*[passes][256]*/) {
    __shared__ uint32_t s_hist[SORT_MAX_PASSES][256];
    for (int i = threadIdx.x; i < SORT_MAX_PASSES * 256; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t stride = gridDim.x * blockDim.x * 4;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) * 4; base < n; base += stride) {
        uint32_t k[4];
        int cnt = 4;
        if (base + 4 <= n) {
            uint4 v = *reinterpret_cast<const uint4*>(keys + base);
            k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
        } else {
            cnt = n - base;
            for (int j = 0; j < cnt; ++j) k[j] = keys[base + j];
        }
        for (int j = 0; j < cnt; ++j) {
#pragma unroll
            for (int p = 0; p < SORT_MAX_PASSES; ++p)
                if (p < plan.passes) atomicAdd(&s_hist[p][(k[j] >> plan.shift[p]) & ((1u << plan.bits[p]) - 1)], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.passes * 256; i += blockDim.x) {
        uint32_t c = (&s_hist[0][0])[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// exclusive scan of one value per thread across a 256-thread block
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* s_warp /*[8]*/, uint32_t* total = nullptr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        uint32_t c = s_warp[w];
        if (w < warp) wbase += c;
        tot += c;
    }
    if (total) *total = tot;
    __syncthreads();
    return wbase + inc - v;
}

// ---- one onesweep digit pass -----------------------------------------------------------
template <bool COMPACT, int NBAL = 8>  // COMPACT: pass 0 of a compacting sort (keep_src given); false removes that code
                                        // entirely.  NBAL: ballots per key = digit bits covered (7 for the 14-bit tile sort)
__global__ void __launch_bounds__(SORT_THREADS, 4) onesweep_pass_kernel(
    const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int shift, int bits,
    const uint32_t* __restrict__ hist /*[256] for this pass*/, uint32_t* __restrict__ tile_counter,
    uint32_t* __restrict__ lookback /*[tiles][256]*/, int iota_values, int precomputed_offsets,
    const uint32_t* __restrict__ gather_src, uint32_t* __restrict__ gather_dst,
    const unsigned long long* __restrict__ n_dev /* device-side element count (clamped to n) or NULL */,
    const uint32_t* __restrict__ keep_src /* pass 0 only: element i takes part iff keep_src[i] != 0, or NULL */,
    unsigned long long* __restrict__ n_out /* receives the number of elements written (the kept ones) or NULL */) {
    __shared__ uint32_t s_warp_hist[8][256];
    __shared__ uint32_t s_local_start[256];
    __shared__ uint32_t s_bin_base[256];
    __shared__ uint32_t s_keys[SORT_TILE];
    __shared__ uint32_t s_vals[SORT_TILE];
    __shared__ uint32_t s_scan[8];
    __shared__ uint32_t s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t mask = (1u << bits) - 1u;
    if (n_dev != nullptr) {  // the count lives on the device (no host sync): the grid covers the capacity `n`
        const unsigned long long nd = *n_dev;
        n = nd < n ? (uint32_t)nd : n;
        if (blockIdx.x * (uint32_t)SORT_TILE >= n && !(blockIdx.x == 0 && n_out != nullptr)) return;
    }
    // look-back mode needs tile ids in scheduling order; with precomputed offsets the block index will do
    if (tid == 0) s_tile = precomputed_offsets ? blockIdx.x : atomicAdd(tile_counter, 1u);
#pragma unroll
    for (int w = 0; w < 8; ++w) s_warp_hist[w][tid] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tile_base = tile * SORT_TILE;
    const uint32_t warp_base = tile_base + warp * (32 * SORT_IPT);

    uint32_t key[SORT_IPT], val[SORT_IPT];
    uint16_t rank[SORT_IPT];
    uint32_t keep_bits = 0xFFFFu;  // bit i: item i of this thread takes part (pass-0 compaction drops the others)
#pragma unroll
    for (int i = 0; i < SORT_IPT; ++i) {
        uint32_t idx = warp_base + i * 32 + lane;
        key[i] = idx < n ? keys_in[idx] : 0xFFFFFFFFu;
        if (COMPACT && !(idx < n && keep_src[idx] != 0u)) keep_bits &= ~(1u << i);
    }
    // values are requested now so that their latency hides behind the ranking
#pragma unroll
    for (int i = 0; i < SORT_IPT; ++i) {
        uint32_t idx = warp_base + i * 32 + lane;
        val[i] = idx < n ? (iota_values ? idx : vals_in[idx]) : 0u;
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
    // Stable multi-split ranking.  `match.any` runs on the slow ADU pipe on sm_100 (measured: 57 % ADU
    // utilisation at 18 % issue utilisation), so the peer mask is built from one ballot per digit bit instead;
    // the group leader bumps the warp's digit counter with one shared-memory atomic (program order keeps
    // item i before item i+1) and broadcasts the old value.
#pragma unroll
    for (int i = 0; i < SORT_IPT; ++i) {
        const uint32_t d = (key[i] >> shift) & mask;
        const bool kept = !COMPACT || ((keep_bits >> i) & 1u);
        uint32_t peers = COMPACT ? __ballot_sync(0xffffffffu, kept) : 0xffffffffu;
#pragma unroll
        for (int b = 0; b < NBAL; ++b) {
            const bool bit = (d >> b) & 1u;  // (a run-time `b < bits` guard was measured slower than fixed ballots)
            const uint32_t bal = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? bal : ~bal;
        }
        // a dropped lane keeps a peer mask without itself in it (possibly empty): it neither leads nor ranks
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (kept && lane == leader) pre = atomicAdd(&s_warp_hist[warp][d], (uint32_t)__popc(peers));
        pre = __shfl_sync(0xffffffffu, pre, (COMPACT && leader < 0) ? 0 : leader);
        rank[i] = (uint16_t)(pre + __popc(peers & lt_mask));
    }
    __syncthreads();

    // per-digit: exclusive offsets of each warp inside the digit, tile total
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        uint32_t c = s_warp_hist[w][tid];
        s_warp_hist[w][tid] = total;
        total += c;
    }
    // publish the tile aggregate as early as possible
    uint32_t* my_lb = lookback + (size_t)tile * 256 + tid;
    if (!precomputed_offsets) st_volatile_u32(my_lb, total | (tile == 0 ? LB_FLAG_PREFIX : LB_FLAG_AGG));

    uint32_t tile_kept = 0;  // items of this tile that take part (== its size unless pass-0 compaction drops some)
    uint32_t local_start = block_exclusive_scan_256(total, s_scan, &tile_kept);
    uint32_t ghist = hist[tid];
    uint32_t gtotal = 0;
    uint32_t gexcl = block_exclusive_scan_256(ghist, s_scan, &gtotal);
    if (n_out != nullptr && blockIdx.x == 0 && tid == 0) *n_out = gtotal;  // elements this pass writes
    if (n_dev != nullptr && tile_base >= n) return;  // block 0 of an empty input only reports the count

    // decoupled look-back over preceding tiles for digit `tid`
    uint32_t excl = 0;
    if (precomputed_offsets) {
        // exclusive count of this digit in all earlier tiles (radix_tile_scan_kernel, digit-major layout)
        excl = lookback[(size_t)tid * gridDim.x + tile];
    } else if (tile > 0) {
        // windowed look-back: fetch up to 4 predecessor states per round trip instead of one
        int t = (int)tile - 1;
        bool found = false;
        while (!found) {
            uint32_t v[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) v[w] = (t - w >= 0) ? ld_volatile_u32(lookback + (size_t)(t - w) * 256 + tid) : LB_FLAG_PREFIX;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                if (found) break;
                uint32_t x = v[w];
                while ((x & LB_FLAG_MASK) == 0) x = ld_volatile_u32(lookback + (size_t)(t - w) * 256 + tid);
                excl += x & LB_VALUE_MASK;
                if ((x & LB_FLAG_MASK) == LB_FLAG_PREFIX) found = true;
            }
            t -= 4;
        }
        st_volatile_u32(my_lb, (excl + total) | LB_FLAG_PREFIX);
    }
    s_local_start[tid] = local_start;
    s_bin_base[tid] = gexcl + excl - local_start;
    __syncthreads();

    // scatter keys and values into tile-sorted order in shared memory
#pragma unroll
    for (int i = 0; i < SORT_IPT; ++i) {
        if (COMPACT && !((keep_bits >> i) & 1u)) continue;
        uint32_t d = (key[i] >> shift) & mask;
        uint32_t p = s_local_start[d] + s_warp_hist[warp][d] + rank[i];
        s_keys[p] = key[i];
        s_vals[p] = val[i];
    }
    __syncthreads();
    // coalesced-by-run write-out: slot p of the tile goes to s_bin_base[digit] + p
    uint32_t valid = (n - tile_base) < (uint32_t)SORT_TILE ? (n - tile_base) : (uint32_t)SORT_TILE;
    if (COMPACT) valid = tile_kept;
#pragma unroll
    for (int k = 0; k < SORT_IPT; ++k) {
        uint32_t p = k * SORT_THREADS + tid;
        if (p < valid) {
            uint32_t kk = s_keys[p];
            uint32_t o = s_bin_base[(kk >> shift) & mask] + p;
            keys_out[o] = kk;
            const uint32_t vv = s_vals[p];
            vals_out[o] = vv;
            // optional payload gathered through the sorted value (the scan of tile counts then reads it in order)
            if (gather_src != nullptr) gather_dst[o] = gather_src[vv];
        }
    }
}

// ---- split pass (no look-back): per-tile digit histograms, a scan over tiles per digit, then the scatter ------
// For the sizes of this path (2 M Gaussians -> 488 tiles, ~15 M instances -> ~3.6 k tiles) every tile of a
// onesweep pass is resident at once and the decoupled look-back degenerates into a serial prefix chain
// (measured: 30 us per pass at 2 M keys, 22 % of stall samples on the look-back load).  Reading the keys a second
// time (4 B per pair) to precompute the offsets is cheaper.
static __global__ void __launch_bounds__(SORT_THREADS) radix_tile_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n,
                                                                      int shift, int bits,
                                                                      uint32_t* __restrict__ tile_hist /*[256][tiles]*/,
                                                                      const unsigned long long* __restrict__ n_dev,
                                                                      const uint32_t* __restrict__ keep_src) {
    __shared__ uint32_t s_hist[256];
    const int tid = threadIdx.x;
    s_hist[tid] = 0;
    __syncthreads();
    if (n_dev != nullptr) {  // device-side count: tiles past it publish zero histograms
        const unsigned long long nd = *n_dev;
        n = nd < n ? (uint32_t)nd : n;
    }
    const uint32_t mask = (1u << bits) - 1u;
    const uint32_t base = blockIdx.x * SORT_TILE;
#pragma unroll
    for (int i = 0; i < SORT_IPT / 4; ++i) {
        const uint32_t idx = base + (i * SORT_THREADS + tid) * 4;
        if (idx + 4 <= n) {
            const uint4 v = *reinterpret_cast<const uint4*>(keys + idx);
            uint4 kp = make_uint4(1u, 1u, 1u, 1u);
            if (keep_src != nullptr) kp = *reinterpret_cast<const uint4*>(keep_src + idx);
            if (kp.x) atomicAdd(&s_hist[(v.x >> shift) & mask], 1u);
            if (kp.y) atomicAdd(&s_hist[(v.y >> shift) & mask], 1u);
            if (kp.z) atomicAdd(&s_hist[(v.z >> shift) & mask], 1u);
            if (kp.w) atomicAdd(&s_hist[(v.w >> shift) & mask], 1u);
        } else {
            for (uint32_t j = idx; j < n && j < idx + 4; ++j)
                if (keep_src == nullptr || keep_src[j] != 0u) atomicAdd(&s_hist[(keys[j] >> shift) & mask], 1u);
        }
    }
    __syncthreads();
    tile_hist[(size_t)tid * gridDim.x + blockIdx.x] = s_hist[tid];  // digit-major: the scan reads contiguously
}

// one CTA per digit: exclusive scan of that digit's contiguous row of tile counts (in place) + the digit total
static __global__ void __launch_bounds__(256) radix_tile_scan_kernel(uint32_t* __restrict__ tile_hist /*[256][tiles]*/,
                                                              uint32_t tiles, uint32_t* __restrict__ digit_totals) {
    __shared__ uint32_t s_scan[8];
    uint32_t* row = tile_hist + (size_t)blockIdx.x * tiles;
    uint32_t running = 0;
    constexpr uint32_t PER = 16;  // values per thread per chunk
    for (uint32_t t0 = 0; t0 < tiles; t0 += 256 * PER) {
        const uint32_t base = t0 + threadIdx.x * PER;
        uint32_t v[PER], sum = 0;
#pragma unroll
        for (uint32_t i = 0; i < PER; ++i) { v[i] = base + i < tiles ? row[base + i] : 0u; sum += v[i]; }
        uint32_t chunk_total;
        uint32_t ex = running + block_exclusive_scan_256(sum, s_scan, &chunk_total);
#pragma unroll
        for (uint32_t i = 0; i < PER; ++i) { if (base + i < tiles) row[base + i] = ex; ex += v[i]; }
        running += chunk_total;
    }
    if (threadIdx.x == 0) digit_totals[blockIdx.x] = running;
}

static inline uint32_t* sort_first_pass_hist(uint32_t* aux) { return aux + SORT_MAX_PASSES * (256 + 64); }

// Sorts n pairs on bits [0, total_bits).  Input in (keys_a, vals_a); (keys_b, vals_b) is the
// ping-pong partner.  Returns true when the sorted result ends in the *_a buffers, false for *_b.
// `aux` must hold SORT_MAX_PASSES*(256+64) uint32 + SORT_MAX_PASSES*tiles*256 uint32 and is
// cleared here.
static inline bool onesweep_sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                                       long long n, int total_bits, uint32_t* aux, int num_sms, cudaStream_t stream,
                                       const char* hist_name = "sort_hist", const char* scan_name = "sort_scan",
                                       const char* pass_name = "sort_pass",
                                       bool iota_values = false, bool first_hist_ready = false,
                                       const uint32_t* gather_src = nullptr, uint32_t* gather_dst = nullptr,
                                       const unsigned long long* n_dev = nullptr, const uint32_t* keep_src = nullptr,
                                       unsigned long long* n_kept_out = nullptr) {
    // n_dev: the element count lives on the device (n is then the capacity the grids cover).  keep_src / n_kept_out:
    // pass 0 drops every element i with keep_src[i] == 0 (stable compaction for free: later passes and the caller
    // work on *n_kept_out elements, which later passes read from the device).
    if (n <= 0) return true;
    SortPlan plan = make_sort_plan(total_bits);
    size_t tiles = sort_num_tiles(n);
    uint32_t* hist = aux;                                  // [4][256]
    uint32_t* counters = aux + SORT_MAX_PASSES * 256;      // [4] (padded to 64 words)
    uint32_t* lookback = aux + SORT_MAX_PASSES * (256 + 64);
    (void)num_sms;
    bool in_a = true;
    for (int p = 0; p < plan.passes; ++p) {
        const uint32_t* ki = in_a ? keys_a : keys_b;
        const uint32_t* vi = in_a ? vals_a : vals_b;
        uint32_t* ko = in_a ? keys_b : keys_a;
        uint32_t* vo = in_a ? vals_b : vals_a;
        uint32_t* tile_offsets = lookback + (size_t)p * tiles * 256;
        if (!(first_hist_ready && p == 0)) {  // else the producer of the keys already filled pass 0's tile histograms
            ProfScope ps(hist_name, stream);
            radix_tile_hist_kernel<<<(unsigned)tiles, SORT_THREADS, 0, stream>>>(
                ki, (uint32_t)n, plan.shift[p], plan.bits[p], tile_offsets, (keep_src && p > 0) ? n_kept_out : n_dev,
                p == 0 ? keep_src : nullptr);
        }
        {
            ProfScope ps(scan_name, stream);
            radix_tile_scan_kernel<<<256, 256, 0, stream>>>(tile_offsets, (uint32_t)tiles, hist + p * 256);
        }
        ProfScope ps(pass_name, stream);
        // (a counting-rank scatter -- private 16-bit counters per thread and digit, one scan over them -- was
        //  measured for the 7-bit tile passes: 17 % fewer instructions but 104 vs 82 us per pass, bound by the
        //  dependent shared-memory chains at 3 CTAs/SM; the ballot ranking stays)
        const bool last = p == plan.passes - 1;
        if (keep_src && p == 0)
            onesweep_pass_kernel<true><<<(unsigned)tiles, SORT_THREADS, 0, stream>>>(
                ki, vi, ko, vo, (uint32_t)n, plan.shift[p], plan.bits[p], hist + p * 256, counters + p, tile_offsets,
                iota_values ? 1 : 0, /*precomputed_offsets=*/1, last ? gather_src : nullptr, last ? gather_dst : nullptr,
                n_dev, keep_src, n_kept_out);
        else if (plan.bits[p] <= 7)
            onesweep_pass_kernel<false, 7><<<(unsigned)tiles, SORT_THREADS, 0, stream>>>(
                ki, vi, ko, vo, (uint32_t)n, plan.shift[p], plan.bits[p], hist + p * 256, counters + p, tile_offsets,
                (iota_values && p == 0) ? 1 : 0, /*precomputed_offsets=*/1, last ? gather_src : nullptr,
                last ? gather_dst : nullptr, (keep_src && p > 0) ? n_kept_out : n_dev, nullptr, nullptr);
        else
            onesweep_pass_kernel<false><<<(unsigned)tiles, SORT_THREADS, 0, stream>>>(
                ki, vi, ko, vo, (uint32_t)n, plan.shift[p], plan.bits[p], hist + p * 256, counters + p, tile_offsets,
                (iota_values && p == 0) ? 1 : 0, /*precomputed_offsets=*/1, last ? gather_src : nullptr,
                last ? gather_dst : nullptr, (keep_src && p > 0) ? n_kept_out : n_dev, nullptr, nullptr);
        in_a = !in_a;
    }
    return in_a;
}

}  // namespace grpg
