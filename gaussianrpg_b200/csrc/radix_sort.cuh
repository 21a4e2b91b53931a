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
__global__ void __launch_bounds__(256) sort_histogram_kernel(const uint32_t* __restrict__ keys, uint32_t n,
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
__global__ void __launch_bounds__(SORT_THREADS, 4) onesweep_pass_kernel(
    const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int shift, int bits,
    const uint32_t* __restrict__ hist /*[256] for this pass*/, uint32_t* __restrict__ tile_counter,
    uint32_t* __restrict__ lookback /*[tiles][256]*/, int iota_values, int precomputed_offsets) {
    __shared__ uint32_t s_warp_hist[8][256];
    __shared__ uint32_t s_local_start[256];
    __shared__ uint32_t s_bin_base[256];
    __shared__ uint32_t s_keys[SORT_TILE];
    __shared__ uint32_t s_vals[SORT_TILE];
    __shared__ uint32_t s_scan[8];
    __shared__ uint32_t s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t mask = (1u << bits) - 1u;
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
#pragma unroll
    for (int i = 0; i < SORT_IPT; ++i) {
        uint32_t idx = warp_base + i * 32 + lane;
        key[i] = idx < n ? keys_in[idx] : 0xFFFFFFFFu;
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
        uint32_t peers = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const bool bit = (d >> b) & 1u;  // (a `b < bits` guard was measured slower than the 8 fixed ballots)
            const uint32_t bal = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? bal : ~bal;
        }
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (lane == leader) pre = atomicAdd(&s_warp_hist[warp][d], (uint32_t)__popc(peers));
        pre = __shfl_sync(0xffffffffu, pre, leader);
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

    uint32_t local_start = block_exclusive_scan_256(total, s_scan);
    uint32_t ghist = hist[tid];
    uint32_t gexcl = block_exclusive_scan_256(ghist, s_scan);

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
        uint32_t d = (key[i] >> shift) & mask;
        uint32_t p = s_local_start[d] + s_warp_hist[warp][d] + rank[i];
        s_keys[p] = key[i];
        s_vals[p] = val[i];
    }
    __syncthreads();
    // coalesced-by-run write-out: slot p of the tile goes to s_bin_base[digit] + p
    const uint32_t valid = (n - tile_base) < (uint32_t)SORT_TILE ? (n - tile_base) : (uint32_t)SORT_TILE;
#pragma unroll
    for (int k = 0; k < SORT_IPT; ++k) {
        uint32_t p = k * SORT_THREADS + tid;
        if (p < valid) {
            uint32_t kk = s_keys[p];
            uint32_t o = s_bin_base[(kk >> shift) & mask] + p;
            keys_out[o] = kk;
            vals_out[o] = s_vals[p];
        }
    }
}

// ---- split pass (no look-back): per-tile digit histograms, a scan over tiles per digit, then the scatter ------
// For the sizes of this path (2 M Gaussians -> 488 tiles, ~15 M instances -> ~3.6 k tiles) every tile of a
// onesweep pass is resident at once and the decoupled look-back degenerates into a serial prefix chain
// (measured: 30 us per pass at 2 M keys, 22 % of stall samples on the look-back load).  Reading the keys a second
// time (4 B per pair) to precompute the offsets is cheaper.
__global__ void __launch_bounds__(SORT_THREADS) radix_tile_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n,
                                                                      int shift, int bits,
                                                                      uint32_t* __restrict__ tile_hist /*[256][tiles]*/) {
    __shared__ uint32_t s_hist[256];
    const int tid = threadIdx.x;
    s_hist[tid] = 0;
    __syncthreads();
    const uint32_t mask = (1u << bits) - 1u;
    const uint32_t base = blockIdx.x * SORT_TILE;
#pragma unroll
    for (int i = 0; i < SORT_IPT / 4; ++i) {
        const uint32_t idx = base + (i * SORT_THREADS + tid) * 4;
        if (idx + 4 <= n) {
            const uint4 v = *reinterpret_cast<const uint4*>(keys + idx);
            atomicAdd(&s_hist[(v.x >> shift) & mask], 1u);
            atomicAdd(&s_hist[(v.y >> shift) & mask], 1u);
            atomicAdd(&s_hist[(v.z >> shift) & mask], 1u);
            atomicAdd(&s_hist[(v.w >> shift) & mask], 1u);
        } else {
            for (uint32_t j = idx; j < n && j < idx + 4; ++j) atomicAdd(&s_hist[(keys[j] >> shift) & mask], 1u);
        }
    }
    __syncthreads();
    tile_hist[(size_t)tid * gridDim.x + blockIdx.x] = s_hist[tid];  // digit-major: the scan reads contiguously
}

// one CTA per digit: exclusive scan of that digit's contiguous row of tile counts (in place) + the digit total
__global__ void __launch_bounds__(256) radix_tile_scan_kernel(uint32_t* __restrict__ tile_hist /*[256][tiles]*/,
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

// ---- scatter pass with counting ranks (digits of at most 7 bits) ---------------------------------------------------
// The ballot ranking above costs ~4 warp instructions per key (8 ballots + bookkeeping, 64 % issue utilisation at
// 2.9 TB/s).  For narrow digits the classic counting rank is cheaper: every thread owns 16 CONSECUTIVE keys, bumps
// a private 16-bit counter per digit, one exclusive scan over the counters in (digit-major, thread) order turns them
// into tile-local rank bases, and each key's rank is base + its arrival number inside the thread.  Stable by
// construction.  Counters: 128 digits x 256 threads x 2 B = 64 KB (padded against bank conflicts), reused afterwards
// as the staging buffer for the coalesced write-out.
constexpr int CS_DIG = 128;
constexpr int CS_CNT_WORDS = CS_DIG * SORT_THREADS / 2;           // 16384 words of 2 x u16
constexpr int CS_PAD_WORDS = (CS_CNT_WORDS / 64) * 4;             // 4 pad words per 64: rows stay 16-byte aligned
constexpr size_t CS_SMEM_BYTES = (size_t)(CS_CNT_WORDS + CS_PAD_WORDS) * 4 + (CS_DIG + 16) * 4;

__device__ __forceinline__ uint32_t cs_entry(uint32_t d, uint32_t t) {  // u16 index of counter (digit d, thread t)
    const uint32_t e = d * SORT_THREADS + t;
    return e + ((e >> 7) << 3);  // every 128 entries (64 words) are followed by 8 pad entries (4 words)
}

__global__ void __launch_bounds__(SORT_THREADS, 3) radix_scatter_count_kernel(
    const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, uint32_t n, int shift, int bits, const uint32_t* __restrict__ digit_totals /*[256]*/,
    const uint32_t* __restrict__ tile_offsets /*[256][tiles]*/) {
    extern __shared__ __align__(16) unsigned char cs_smem[];
    uint16_t* cnt = reinterpret_cast<uint16_t*>(cs_smem);
    uint32_t* cntw = reinterpret_cast<uint32_t*>(cs_smem);
    uint32_t* s_base = cntw + CS_CNT_WORDS + CS_PAD_WORDS;  // [128]
    uint32_t* s_scan = s_base + CS_DIG;                     // [8]
    const uint32_t t = threadIdx.x, tile = blockIdx.x, tile_base = tile * SORT_TILE;
    const uint32_t mask = (1u << bits) - 1u;

    // zero the counters (16-byte stores)
    {
        uint4* z = reinterpret_cast<uint4*>(cs_smem);
        constexpr int N16 = (CS_CNT_WORDS + CS_PAD_WORDS) / 4;
        for (int i = t; i < N16; i += SORT_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    // blocked loads: thread t owns the 16 consecutive pairs starting at tile_base + 16 t
    uint32_t key[SORT_IPT], val[SORT_IPT];
    const uint32_t base = tile_base + t * SORT_IPT;
    if (base + SORT_IPT <= n) {
#pragma unroll
        for (int q = 0; q < SORT_IPT / 4; ++q) {
            const uint4 k4 = *reinterpret_cast<const uint4*>(keys_in + base + 4 * q);
            const uint4 v4 = *reinterpret_cast<const uint4*>(vals_in + base + 4 * q);
            key[4 * q] = k4.x; key[4 * q + 1] = k4.y; key[4 * q + 2] = k4.z; key[4 * q + 3] = k4.w;
            val[4 * q] = v4.x; val[4 * q + 1] = v4.y; val[4 * q + 2] = v4.z; val[4 * q + 3] = v4.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < SORT_IPT; ++i) {
            key[i] = base + i < n ? keys_in[base + i] : 0xFFFFFFFFu;  // pads rank last inside the last digit
            val[i] = base + i < n ? vals_in[base + i] : 0u;
        }
    }
    __syncthreads();

    // phase 1: private counters; remember each key's arrival number among equal digits of this thread (4 bits)
    unsigned long long arrival = 0ull;
#pragma unroll
    for (int i = 0; i < SORT_IPT; ++i) {
        const uint32_t d = (key[i] >> shift) & mask;
        uint16_t* c = cnt + cs_entry(d, t);
        const uint32_t old = *c;
        *c = (uint16_t)(old + 1);
        arrival |= (unsigned long long)old << (4 * i);
    }
    __syncthreads();

    // phase 2: exclusive scan of all counters in (digit-major, thread) order; thread t rakes entries [128t, 128t+128)
    {
        uint4* row = reinterpret_cast<uint4*>(cntw + t * 68);  // 64 data words + 4 pad words per row
        uint32_t sum = 0;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const uint4 x = row[q];
            sum += (x.x & 0xffffu) + (x.x >> 16) + (x.y & 0xffffu) + (x.y >> 16) + (x.z & 0xffffu) + (x.z >> 16) +
                   (x.w & 0xffffu) + (x.w >> 16);
        }
        uint32_t run = block_exclusive_scan_256(sum, s_scan);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            uint4 x = row[q];
            uint32_t lo, hi;
            lo = x.x & 0xffffu; hi = x.x >> 16; x.x = run | ((run + lo) << 16); run += lo + hi;
            lo = x.y & 0xffffu; hi = x.y >> 16; x.y = run | ((run + lo) << 16); run += lo + hi;
            lo = x.z & 0xffffu; hi = x.z >> 16; x.z = run | ((run + lo) << 16); run += lo + hi;
            lo = x.w & 0xffffu; hi = x.w >> 16; x.w = run | ((run + lo) << 16); run += lo + hi;
            row[q] = x;
        }
    }
    __syncthreads();

    // phase 3: ranks, and the per-digit output base (global position of the digit run minus its tile-local start)
    uint16_t rank[SORT_IPT];
#pragma unroll
    for (int i = 0; i < SORT_IPT; ++i) {
        const uint32_t d = (key[i] >> shift) & mask;
        rank[i] = (uint16_t)(cnt[cs_entry(d, t)] + (uint32_t)((arrival >> (4 * i)) & 15ull));
    }
    const uint32_t gexcl = block_exclusive_scan_256(digit_totals[t], s_scan);
    if (t < CS_DIG) {
        const uint32_t local_start = cnt[cs_entry(t, 0)];
        s_base[t] = gexcl + tile_offsets[(size_t)t * gridDim.x + tile] - local_start;
    }
    __syncthreads();  // every rank base has been read: the counter area becomes the staging buffer

    uint32_t* s_keys = cntw;
    uint32_t* s_vals = cntw + SORT_TILE;
#pragma unroll
    for (int i = 0; i < SORT_IPT; ++i) {
        s_keys[rank[i]] = key[i];
        s_vals[rank[i]] = val[i];
    }
    __syncthreads();
    const uint32_t valid = (n - tile_base) < (uint32_t)SORT_TILE ? (n - tile_base) : (uint32_t)SORT_TILE;
#pragma unroll
    for (int k = 0; k < SORT_IPT; ++k) {
        const uint32_t p = k * SORT_THREADS + t;
        if (p < valid) {
            const uint32_t kk = s_keys[p];
            const uint32_t o = s_base[(kk >> shift) & mask] + p;
            keys_out[o] = kk;
            vals_out[o] = s_vals[p];
        }
    }
}

static inline uint32_t* sort_first_pass_hist(uint32_t* aux) { return aux + SORT_MAX_PASSES * (256 + 64); }

// Sorts n pairs on bits [0, total_bits).  Input in (keys_a, vals_a); (keys_b, vals_b) is the
// ping-pong partner.  Returns true when the sorted result ends in the *_a buffers, false for *_b.
// `aux` must hold SORT_MAX_PASSES*(256+64) uint32 + SORT_MAX_PASSES*tiles*256 uint32 and is
// cleared here.
static inline bool onesweep_sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                                       long long n, int total_bits, uint32_t* aux, int num_sms, cudaStream_t stream,
                                       const char* hist_name = "sort_hist", const char* pass_name = "sort_pass",
                                       bool iota_values = false, bool first_hist_ready = false) {
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
        {
            ProfScope ps(hist_name, stream);
            if (!(first_hist_ready && p == 0))  // the producer of the keys already filled pass 0's tile histograms
                radix_tile_hist_kernel<<<(unsigned)tiles, SORT_THREADS, 0, stream>>>(ki, (uint32_t)n, plan.shift[p],
                                                                                    plan.bits[p], tile_offsets);
            radix_tile_scan_kernel<<<256, 256, 0, stream>>>(tile_offsets, (uint32_t)tiles, hist + p * 256);
        }
        ProfScope ps(pass_name, stream);
        if (plan.bits[p] <= 7 && !(iota_values && p == 0)) {
            static bool smem_opt_in = false;
            if (!smem_opt_in) {
                cudaFuncSetAttribute(radix_scatter_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)CS_SMEM_BYTES);
                smem_opt_in = true;
            }
            radix_scatter_count_kernel<<<(unsigned)tiles, SORT_THREADS, CS_SMEM_BYTES, stream>>>(
                ki, vi, ko, vo, (uint32_t)n, plan.shift[p], plan.bits[p], hist + p * 256, tile_offsets);
        } else {
            onesweep_pass_kernel<<<(unsigned)tiles, SORT_THREADS, 0, stream>>>(
                ki, vi, ko, vo, (uint32_t)n, plan.shift[p], plan.bits[p], hist + p * 256, counters + p, tile_offsets,
                (iota_values && p == 0) ? 1 : 0, /*precomputed_offsets=*/1);
        }
        in_a = !in_a;
    }
    return in_a;
}

}  // namespace grpg
