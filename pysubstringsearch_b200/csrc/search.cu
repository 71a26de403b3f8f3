// search.cu — batched substring search kernels for sm_100a (see search.cuh).
#include "search.cuh"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>

namespace pss {

namespace {

// ------------------------------------------------------------------------------------
// bounds: one warp per (query, chunk)
// ------------------------------------------------------------------------------------
constexpr int BD_THREADS = 256;

// Sign of (suffix starting at s) versus pattern P[0..m): -1 suffix < P, 0 P is a prefix of
// the suffix (lib.rs:220 starts_with), +1 suffix > P.  Unsigned bytes; a suffix that ends
// inside the pattern is smaller (slice cmp, lib.rs:224).  The warp compares 32 bytes per
// step: lane l owns byte w + l of the window.
__device__ __forceinline__ int cmp_suffix(const uint8_t *__restrict__ text, uint32_t n, uint32_t s,
                                          const uint8_t *__restrict__ P, uint32_t m, uint32_t pc0,
                                          uint32_t lane) {
    const uint32_t avail = n - s;
    for (uint32_t w = 0; w < m; w += 32) {
        const uint32_t b   = w + lane;
        const bool in_pat  = b < m;
        const uint32_t pc  = (w == 0) ? pc0 : (in_pat ? (uint32_t)__ldg(P + b) : 0u);
        const bool in_txt  = in_pat && b < avail;
        const uint32_t tc  = in_txt ? (uint32_t)__ldg(text + s + b) : 0u;
        const bool neq     = in_pat && (!in_txt || tc != pc);
        const uint32_t msk = __ballot_sync(0xffffffffu, neq);
        if (msk) {
            const int f        = __ffs(msk) - 1;
            const uint32_t enc = (in_txt ? 0u : 0x10000u) | (tc << 8) | pc;
            const uint32_t e   = __shfl_sync(0xffffffffu, enc, f);
            if (e & 0x10000u) return -1;                       // suffix ended first
            return ((e >> 8) & 0xFFu) < (e & 0xFFu) ? -1 : 1;
        }
    }
    return 0;
}

// Range of SA slots whose suffixes start with P, found by the whole warp: lower bound, then
// upper bound starting from it (the reference runs the same two searches, lib.rs:212-252).
__device__ __forceinline__ void warp_bounds(const uint8_t *__restrict__ text, const int32_t *__restrict__ sa, uint32_t n,
                                            const uint32_t *__restrict__ bucket,
                                            const uint8_t *__restrict__ P, uint32_t m, uint32_t lane,
                                            uint32_t *lb_out, uint32_t *cnt_out) {
    const uint32_t pc0 = lane < m ? (uint32_t)__ldg(P + lane) : 0u;
    // every match lies inside the bucket of the pattern's first two bytes
    uint32_t lo = 0, hi_all = n;
    if (bucket != nullptr && m >= 2) {
        const uint32_t key = (__shfl_sync(0xffffffffu, pc0, 0) << 8) | __shfl_sync(0xffffffffu, pc0, 1);
        lo     = min(__ldg(bucket + key), n);                 // clamped: a corrupt index must not send probes out of range
        hi_all = min(max(__ldg(bucket + key + 1), lo), n);
    }
    // A probe is two dependent DRAM round trips: SA[mid], then the text it points to.  The SA
    // slots of BOTH possible next probes are therefore requested while the text of this one is
    // still on its way (uniform addresses: one sector each), so that a level costs one round
    // trip instead of two — the kernel is bound by that latency (ncu: one load in flight per
    // warp, 1 TB/s of random sectors), not by the extra sector.
    // UPPER = false: smallest slot whose suffix is >= P (as a prefix comparison);
    // UPPER = true : smallest slot whose suffix is > P and does not start with it.
    auto bisect = [&](uint32_t lo_, uint32_t hi_, const bool upper) -> uint32_t {
        if (lo_ >= hi_) return lo_;
        uint32_t mid = lo_ + ((hi_ - lo_) >> 1);
        uint32_t s   = (uint32_t)__ldg(sa + mid);
        while (true) {
            const uint32_t ml = lo_ + ((mid - lo_) >> 1), mr = (mid + 1) + ((hi_ - (mid + 1)) >> 1);
            const bool vl = lo_ < mid, vr = mid + 1 < hi_;
            const uint32_t sl = vl ? (uint32_t)__ldg(sa + ml) : 0u;
            const uint32_t sr = vr ? (uint32_t)__ldg(sa + mr) : 0u;
            const int c = cmp_suffix(text, n, s, P, m, pc0, lane);
            if (upper ? c <= 0 : c < 0) {
                lo_ = mid + 1;
                if (!vr) return lo_;
                mid = mr; s = sr;
            } else {
                hi_ = mid;
                if (!vl) return lo_;
                mid = ml; s = sl;
            }
        }
    };
    const uint32_t lb = bisect(lo, hi_all, false);
    const uint32_t ub = bisect(lb, hi_all, true);    // starts from the lower bound (lib.rs:235)
    *lb_out  = lb;
    *cnt_out = ub - lb;
}

__global__ void __launch_bounds__(BD_THREADS)
bounds_kernel(const DeviceChunk *__restrict__ chunks, int nc, const uint8_t *__restrict__ patterns,
              const int64_t *__restrict__ pat_off, uint32_t npairs, uint32_t *__restrict__ lb_out,
              uint32_t *__restrict__ cnt_out) {
    const uint32_t lane = lane_id();
    const uint32_t pair = (uint32_t)(((uint64_t)blockIdx.x * BD_THREADS + threadIdx.x) >> 5);   // up to 2^31 pairs x 32 lanes
    if (pair >= npairs) return;
    const uint32_t q = pair / (uint32_t)nc, c = pair % (uint32_t)nc;
    uint32_t lb, cnt;
    // a malformed offsets array (device callers) must not turn into a wild pattern length
    const int64_t o0 = pat_off[q], o1 = pat_off[q + 1];
    const uint32_t m = o1 > o0 ? (uint32_t)min(o1 - o0, (int64_t)0x7FFFFFFF) : 0u;
    warp_bounds(chunks[c].text, chunks[c].sa, chunks[c].n, chunks[c].bucket, patterns + o0, m, lane, &lb, &cnt);
    if (lane == 0) {
        lb_out[pair]  = lb;
        cnt_out[pair] = cnt;
    }
}

// Several (query, chunk) pairs per warp: G lanes own one pair and compare G text bytes per
// step, the 32 / G pairs of a warp advance in lockstep (every loop condition is a warp vote,
// every ballot / shuffle runs with the full mask).  Patterns are short (config 2: 4-32 bytes),
// so a whole warp per pair left most lanes idle AND left the kernel with one dependent load
// chain per warp; this form keeps 32 / G chains (plus their look-ahead SA loads) in flight per
// warp at the same occupancy.  Same searches, same results as warp_bounds.
template <int G, bool LOOKAHEAD>
__global__ void __launch_bounds__(BD_THREADS)
bounds_group_kernel(const DeviceChunk *__restrict__ chunks, int nc, const uint8_t *__restrict__ patterns,
                    const int64_t *__restrict__ pat_off, uint32_t npairs, uint32_t *__restrict__ lb_out,
                    uint32_t *__restrict__ cnt_out) {
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr uint32_t GM   = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
    const uint32_t lane   = lane_id();
    const uint32_t sub    = lane % G;              // byte of the window this lane owns
    const uint32_t gbase  = lane - sub;            // first lane of the group
    const uint32_t pair   = (uint32_t)(((uint64_t)blockIdx.x * BD_THREADS + threadIdx.x) / G);   // up to 2^31 pairs x G lanes
    const bool     valid  = pair < npairs;
    const uint32_t q = valid ? pair / (uint32_t)nc : 0u, c = valid ? pair % (uint32_t)nc : 0u;
    const uint8_t *__restrict__ text = chunks[c].text;
    const int32_t *__restrict__ sa   = chunks[c].sa;
    const uint32_t n = chunks[c].n;
    const uint32_t *__restrict__ bucket = chunks[c].bucket;
    int64_t o0 = 0, o1 = 0;
    if (valid) { o0 = pat_off[q]; o1 = pat_off[q + 1]; }
    const uint8_t *__restrict__ P = patterns + o0;
    const uint32_t m   = o1 > o0 ? (uint32_t)min(o1 - o0, (int64_t)0x7FFFFFFF) : 0u;
    const uint32_t pc0 = sub < m ? (uint32_t)__ldg(P + sub) : 0u;

    uint32_t lo = 0, hi_all = n;
    if (valid && bucket != nullptr && m >= 2) {
        const uint32_t key = ((uint32_t)__ldg(P) << 8) | (uint32_t)__ldg(P + 1);
        lo     = min(__ldg(bucket + key), n);
        hi_all = min(max(__ldg(bucket + key + 1), lo), n);
    }
    uint32_t hi = hi_all, lb = 0, ub = 0;
    int phase = valid ? 0 : 2;                     // 0 lower bound, 1 upper bound, 2 done
    bool have = false;                             // (mid, s) is the next probe of [lo, hi)
    uint32_t mid = 0, s = 0;
    if (phase == 0 && lo < hi) {
        mid = lo + ((hi - lo) >> 1);
        s = (uint32_t)__ldg(sa + mid);
        have = true;
    }
    while (true) {
        if (phase == 0 && !have) {                 // lower bound found: the upper search starts from it (lib.rs:235)
            lb = lo; hi = hi_all; phase = 1;
            if (lo < hi) {
                mid = lo + ((hi - lo) >> 1);
                s = (uint32_t)__ldg(sa + mid);
                have = true;
            }
        }
        if (phase == 1 && !have) { ub = lo; phase = 2; }
        const bool probing = phase < 2;
        if (!__any_sync(FULL, probing)) break;
        // SA slots of both possible next probes, requested before this probe's text arrives
        const uint32_t ml = lo + ((mid - lo) >> 1), mr = (mid + 1) + ((hi - (mid + 1)) >> 1);
        const bool vl = probing && lo < mid, vr = probing && mid + 1 < hi;
        uint32_t sl = 0, sr = 0;
        if (LOOKAHEAD) {
            sl = vl ? (uint32_t)__ldg(sa + ml) : 0u;
            sr = vr ? (uint32_t)__ldg(sa + mr) : 0u;
        }
        // suffix at s versus P, G bytes per step (cmp_suffix's rules)
        const uint32_t avail = n - s;
        int res = 2;                               // undecided
        for (uint32_t w = 0;; w += G) {
            const bool need = probing && res == 2 && w < m;
            if (!__any_sync(FULL, need)) break;
            const uint32_t b   = w + sub;
            const bool in_pat  = need && b < m;
            const uint32_t pc  = (w == 0) ? pc0 : (in_pat ? (uint32_t)__ldg(P + b) : 0u);
            const bool in_txt  = in_pat && b < avail;
            const uint32_t tc  = in_txt ? (uint32_t)__ldg(text + s + b) : 0u;
            const bool neq     = in_pat && (!in_txt || tc != pc);
            const uint32_t msk = (__ballot_sync(FULL, neq) >> gbase) & GM;
            const uint32_t enc = (in_txt ? 0u : 0x10000u) | (tc << 8) | pc;
            const uint32_t e   = __shfl_sync(FULL, enc, gbase + (msk ? __ffs(msk) - 1 : 0));
            if (need && msk) res = (e & 0x10000u) ? -1 : (((e >> 8) & 0xFFu) < (e & 0xFFu) ? -1 : 1);
        }
        if (res == 2) res = 0;                     // P is a prefix of the suffix
        if (probing) {
            if (phase ? res <= 0 : res < 0) { lo = mid + 1; have = vr; mid = mr; s = sr; }
            else                            { hi = mid;     have = vl; mid = ml; s = sl; }
            if (!LOOKAHEAD && have) s = (uint32_t)__ldg(sa + mid);
        }
    }
    if (valid && sub == 0) {
        lb_out[pair]  = lb;
        cnt_out[pair] = ub - lb;
    }
}

constexpr int SCAN_THREADS = 1024;   // single-CTA scan over the compaction's tile sums (tile_scan_kernel)

// Dedup tiers by the number of matching suffixes h of a (query, chunk) pair:
//   light   h <= LIGHT_MAX    one warp, all-pairs compare in shared memory    (pair_dedup_kernel)
//   medium  h <= MEDIUM_MAX   one CTA, hash set of entry starts in shared memory (pair_dedup_hash_kernel:
//                             32 KB tables for h <= 2048, a persistent second launch with 128 KB tables above)
//   heavy   above             global stable sort of (pair, entry start)        (sort_async + mark_kernel)
constexpr uint32_t LIGHT_MAX  = 256;
constexpr uint32_t MEDIUM_MAX = 8192;

// ------------------------------------------------------------------------------------
// Scans over the per-pair arrays, PS_TILE pairs per CTA in two launches each: per-CTA partials,
// then every CTA folds the partials of the CTAs before (after) it and scans its own tile.
// (One 1024-thread CTA walking the whole array — a thread per contiguous range, one sector per
// load — took 0.25 + 0.14 ms of a 150 000-pair batch whose other kernels add up to 1.1 ms.)
// ------------------------------------------------------------------------------------
constexpr int PS_THREADS = 256;
constexpr int PS_IPT     = 8;
constexpr int PS_TILE    = PS_THREADS * PS_IPT;

struct Tri {
    unsigned long long h, v;   // hits, hits of heavy pairs
    uint32_t m;                // medium pairs
};
__device__ __forceinline__ Tri tri_add(const Tri &a, const Tri &b) { return Tri{a.h + b.h, a.v + b.v, a.m + b.m}; }
__device__ __forceinline__ Tri tri_of(uint32_t c) {
    return Tri{c, c > MEDIUM_MAX ? c : 0ull, (c > LIGHT_MAX && c <= MEDIUM_MAX) ? 1u : 0u};
}
// Sum over the block, returned to every thread; *excl = the sum over the threads before this one.
__device__ __forceinline__ Tri tri_block_scan(const Tri &mine, Tri *s_warp, Tri *excl) {
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    Tri incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Tri y{__shfl_up_sync(0xffffffffu, incl.h, o), __shfl_up_sync(0xffffffffu, incl.v, o),
              __shfl_up_sync(0xffffffffu, incl.m, o)};
        if (lane >= (uint32_t)o) incl = tri_add(incl, y);
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    Tri pre{0, 0, 0}, tot{0, 0, 0};
#pragma unroll
    for (int w = 0; w < PS_THREADS / 32; ++w) {
        const Tri t = s_warp[w];
        if ((uint32_t)w < warp) pre = tri_add(pre, t);
        tot = tri_add(tot, t);
    }
    __syncthreads();
    *excl = Tri{pre.h + incl.h - mine.h, pre.v + incl.v - mine.v, pre.m + incl.m - mine.m};
    return tot;
}

// partials[b] = (hits, heavy hits, medium pairs) of the pairs of tile b
__global__ void __launch_bounds__(PS_THREADS)
hit_partials_kernel(const uint32_t *__restrict__ cnt, uint32_t npairs, Tri *__restrict__ partials) {
    __shared__ Tri s_warp[PS_THREADS / 32];
    const uint32_t base = blockIdx.x * PS_TILE + threadIdx.x * PS_IPT;
    Tri mine{0, 0, 0};
#pragma unroll
    for (int e = 0; e < PS_IPT; ++e)
        if (base + e < npairs) mine = tri_add(mine, tri_of(__ldg(cnt + base + e)));
    Tri excl;
    const Tri tot = tri_block_scan(mine, s_warp, &excl);
    if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

// hit_off[p] = sum of cnt[0..p) (u32, wraps only when the batch is oversized, which the
// u64 total reveals); hit_off[npairs] = total.  heavy_off[p] = the same scan over the heavy
// pairs only (where pair p's hits go in the sort's input); med_list = the medium pairs, in order.
// totals[0] = all hits, totals[1] = hits of heavy pairs, totals[2] = number of medium pairs.
__global__ void __launch_bounds__(PS_THREADS)
hit_offsets_kernel(const uint32_t *__restrict__ cnt, uint32_t npairs, const Tri *__restrict__ partials,
                   uint32_t *__restrict__ hit_off, uint32_t *__restrict__ heavy_off, uint32_t *__restrict__ med_list,
                   unsigned long long *__restrict__ totals) {
    __shared__ Tri s_warp[PS_THREADS / 32];
    // everything before this tile
    Tri before{0, 0, 0};
    for (uint32_t j = threadIdx.x; j < blockIdx.x; j += PS_THREADS) before = tri_add(before, partials[j]);
    Tri unused;
    before = tri_block_scan(before, s_warp, &unused);
    // this tile
    const uint32_t base = blockIdx.x * PS_TILE + threadIdx.x * PS_IPT;
    uint32_t c[PS_IPT];
    Tri mine{0, 0, 0};
#pragma unroll
    for (int e = 0; e < PS_IPT; ++e) {
        c[e] = base + e < npairs ? __ldg(cnt + base + e) : 0u;
        mine = tri_add(mine, tri_of(c[e]));
    }
    Tri run;
    const Tri tot = tri_block_scan(mine, s_warp, &run);
    run = tri_add(run, before);
#pragma unroll
    for (int e = 0; e < PS_IPT; ++e) {
        const uint32_t p = base + e;
        if (p >= npairs) break;
        hit_off[p]   = (uint32_t)run.h;
        heavy_off[p] = (uint32_t)run.v;
        if (c[e] > LIGHT_MAX && c[e] <= MEDIUM_MAX) med_list[run.m] = p;
        run = tri_add(run, tri_of(c[e]));
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        const Tri all = tri_add(before, tot);
        hit_off[npairs] = (uint32_t)all.h;
        totals[0]       = all.h;
        totals[1]       = all.v;
        totals[2]       = all.m;
    }
}

// pair_first[p] = output offset of pair p's first entry, 0xFFFFFFFF for pairs without hits
// (written by compact_kernel).  entry_off[p] = base + entries before pair p = the first
// offset of the next pair that has hits (suffix minimum), or the sub-batch total.
__device__ __forceinline__ uint32_t block_min_256(uint32_t v, uint32_t *s_warp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane_id() == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t r = 0xFFFFFFFFu;
#pragma unroll
    for (int w = 0; w < PS_THREADS / 32; ++w) r = min(r, s_warp[w]);
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(PS_THREADS)
pair_first_min_kernel(const uint32_t *__restrict__ pair_first, uint32_t np, uint32_t *__restrict__ part_min) {
    __shared__ uint32_t s_warp[PS_THREADS / 32];
    const uint32_t base = blockIdx.x * PS_TILE + threadIdx.x * PS_IPT;
    uint32_t mn = 0xFFFFFFFFu;
#pragma unroll
    for (int e = 0; e < PS_IPT; ++e)
        if (base + e < np) mn = min(mn, __ldg(pair_first + base + e));
    mn = block_min_256(mn, s_warp);
    if (threadIdx.x == 0) part_min[blockIdx.x] = mn;
}

__global__ void __launch_bounds__(PS_THREADS)
entry_offsets_kernel(const uint32_t *__restrict__ pair_first, uint32_t np, const uint32_t *__restrict__ part_min,
                     const uint32_t *__restrict__ kept_total, uint32_t base, uint32_t *__restrict__ entry_off) {
    __shared__ uint32_t s_warp[PS_THREADS / 32];
    const uint32_t total = *kept_total;
    // the first entry of any later tile (or the total)
    uint32_t after = total;
    for (uint32_t j = blockIdx.x + 1 + threadIdx.x; j < gridDim.x; j += PS_THREADS) after = min(after, part_min[j]);
    after = block_min_256(after, s_warp);
    const uint32_t at = blockIdx.x * PS_TILE + threadIdx.x * PS_IPT;
    uint32_t v[PS_IPT];
#pragma unroll
    for (int e = 0; e < PS_IPT; ++e) v[e] = at + e < np ? __ldg(pair_first + at + e) : 0xFFFFFFFFu;
#pragma unroll
    for (int e = PS_IPT - 2; e >= 0; --e) v[e] = min(v[e], v[e + 1]);      // suffix minimum inside the thread
    // minimum over the threads after this one: inclusive suffix scan in the warp, then the later warps
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t incl = v[0];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + (uint32_t)o < 32) incl = min(incl, y);
    }
    if (lane == 0) s_warp[warp] = incl;
    uint32_t later = __shfl_down_sync(0xffffffffu, incl, 1);
    if (lane == 31) later = 0xFFFFFFFFu;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < PS_THREADS / 32; ++w)
        if ((uint32_t)w > warp) later = min(later, s_warp[w]);
    later = min(later, after);
#pragma unroll
    for (int e = 0; e < PS_IPT; ++e)
        if (at + e < np) entry_off[at + e] = base + min(v[e], later);
    if (blockIdx.x == 0 && threadIdx.x == 0) entry_off[np] = base + total;
}

__global__ void __launch_bounds__(256)
query_offsets_kernel(const uint32_t *__restrict__ entry_off, uint32_t nq, uint32_t nc, int64_t *__restrict__ query_off) {
    const uint32_t q = blockIdx.x * 256 + threadIdx.x;
    if (q <= nq) query_off[q] = (int64_t)entry_off[(size_t)q * nc];
}

// ------------------------------------------------------------------------------------
// newline side index: sorted offsets of every '\n' of a chunk
// ------------------------------------------------------------------------------------
constexpr int NL_THREADS = 256;
constexpr int NL_BYTES_PER_THREAD = 16;
constexpr int NL_TILE = NL_THREADS * NL_BYTES_PER_THREAD;   // 4096 text bytes per CTA

__device__ __forceinline__ uint32_t block_excl_sum_256(uint32_t v, uint32_t *s_warp, uint32_t *total) {
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t pre = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        uint32_t t = s_warp[w];
        if ((uint32_t)w < warp) pre += t;
        tot += t;
    }
    *total = tot;
    __syncthreads();
    return pre + incl - v;
}

// bit j of the result is set when byte j of the thread's 16-byte slice is '\n' (the text
// buffer has 16 readable zero bytes past n, and its base is 256-byte aligned)
__device__ __forceinline__ uint32_t newline_mask16(const uint8_t *__restrict__ text, uint32_t n, uint32_t at) {
    if (at >= n) return 0;
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(text + at));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t eq = __vcmpeq4(w[k], 0x0A0A0A0Au);   // 0xFF per equal byte
        m |= ((eq & 1u) | ((eq >> 7) & 2u) | ((eq >> 14) & 4u) | ((eq >> 21) & 8u)) << (4 * k);
    }
    const uint32_t left = n - at;
    if (left < 16) m &= (1u << left) - 1u;
    return m;
}

__global__ void __launch_bounds__(NL_THREADS)
newline_count_kernel(const uint8_t *__restrict__ text, uint32_t n, uint32_t *__restrict__ tile_sum) {
    __shared__ uint32_t s_warp[8];
    const uint32_t at = blockIdx.x * NL_TILE + threadIdx.x * NL_BYTES_PER_THREAD;
    uint32_t total;
    (void)block_excl_sum_256(__popc(newline_mask16(text, n, at)), s_warp, &total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(NL_THREADS)
newline_fill_kernel(const uint8_t *__restrict__ text, uint32_t n, const uint32_t *__restrict__ tile_prefix,
                    uint32_t *__restrict__ nl) {
    __shared__ uint32_t s_warp[8];
    const uint32_t at = blockIdx.x * NL_TILE + threadIdx.x * NL_BYTES_PER_THREAD;
    uint32_t m = newline_mask16(text, n, at);
    uint32_t total;
    uint32_t o = tile_prefix[blockIdx.x] + block_excl_sum_256(__popc(m), s_warp, &total);
    while (m) {
        const int b = __ffs(m) - 1;
        nl[o++] = at + b;
        m &= m - 1;
    }
}

// Line directory: dir[j] = number of '\n' at text positions < j * LINE_BLOCK (j = 0 .. ceil(n / LINE_BLOCK)),
// i.e. where the newlines of text block j start inside the sorted offset list.  One binary search
// per block at open; extraction then needs no search over the whole list (entry_bounds).
__global__ void __launch_bounds__(256)
line_directory_kernel(const uint32_t *__restrict__ nl, uint32_t n_lines, uint32_t n_dir, uint32_t *__restrict__ dir) {
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n_dir) return;
    const uint64_t at = (uint64_t)j * LINE_BLOCK;
    uint32_t lo = 0, hi = n_lines;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if ((uint64_t)__ldg(nl + mid) < at) lo = mid + 1;
        else hi = mid;
    }
    dir[j] = lo;
}

// Line records: one 32-byte sector per LINE_REC_BLOCK (128) text bytes that answers "which entry
// contains position p" for every p of the block with a single load:
//   x = start of the entry that is open at the block's first byte (1 + last '\n' before the block, 0 if none)
//   y = first '\n' at or after the block's end, n - 1 if there is none (lib.rs:266-269: None -> len - 1)
//   z, w, x', y' = 128-bit map of the '\n' inside the block
// Built from the sorted newline offsets, one thread per block.
__global__ void __launch_bounds__(256)
line_records_kernel(const uint32_t *__restrict__ nl, uint32_t n_lines, uint32_t n, uint32_t n_blocks,
                    uint4 *__restrict__ rec) {
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n_blocks) return;
    const uint64_t at = (uint64_t)j * LINE_REC_BLOCK, end = at + LINE_REC_BLOCK;
    uint32_t lo = 0, hi = n_lines;
    while (lo < hi) {                                   // first newline at or after the block's start
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if ((uint64_t)__ldg(nl + mid) < at) lo = mid + 1;
        else hi = mid;
    }
    uint32_t bm[4] = {0, 0, 0, 0};
    const uint32_t open_at = lo > 0 ? __ldg(nl + lo - 1) + 1 : 0u;
    uint32_t k = lo;
    for (; k < n_lines; ++k) {
        const uint64_t x = __ldg(nl + k);
        if (x >= end) break;
        const uint32_t o = (uint32_t)(x - at);
#pragma unroll
        for (int w = 0; w < 4; ++w)
            if ((o >> 5) == (uint32_t)w) bm[w] |= 1u << (o & 31);
    }
    const uint32_t next = k < n_lines ? __ldg(nl + k) : n - 1;
    rec[2 * (size_t)j]     = make_uint4(open_at, next, bm[0], bm[1]);
    rec[2 * (size_t)j + 1] = make_uint4(bm[2], bm[3], 0u, 0u);
}

// ------------------------------------------------------------------------------------
// 2-byte prefix table: bucket[a << 8 | b] = first SA slot whose suffix starts with a, b.
// Boundaries are where the prefix changes along the suffix array; the thread at a boundary
// fills every table entry between the two prefixes.  A one-byte suffix "c" (the text's last
// position) takes the key (c, 0): it sorts directly before every "c\0..." suffix, so buckets stay
// contiguous and ordered (a bucket may hold it without matching — the searches skip it).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prefix_key(const uint8_t *__restrict__ text, uint32_t n, uint32_t pos) {
    const uint32_t a = __ldg(text + pos);
    const uint32_t b = pos + 1 < n ? (uint32_t)__ldg(text + pos + 1) : 0u;
    return (a << 8) | b;
}

__global__ void __launch_bounds__(256)
prefix_bucket_kernel(const uint8_t *__restrict__ text, const int32_t *__restrict__ sa, uint32_t n,
                     uint32_t *__restrict__ bucket) {
    const uint32_t k = blockIdx.x * 256 + threadIdx.x;
    if (k >= n) return;
    const uint32_t cur = prefix_key(text, n, (uint32_t)__ldg(sa + k));
    if (k == 0) {
        for (uint32_t j = 0; j <= cur; ++j) bucket[j] = 0;
    } else {
        const uint32_t prev = prefix_key(text, n, (uint32_t)__ldg(sa + k - 1));
        for (uint32_t j = prev + 1; j <= cur; ++j) bucket[j] = k;       // empty when prev == cur
    }
    if (k == n - 1)
        for (uint32_t j = cur + 1; j <= 65536u; ++j) bucket[j] = n;
}

// ------------------------------------------------------------------------------------
// extract: one thread per matching suffix
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t find_pair(const uint32_t *__restrict__ hit_off, uint32_t npairs, uint32_t f) {
    // largest p with hit_off[p] <= f  (hit_off has npairs + 1 entries, hit_off[npairs] = H > f)
    uint32_t lo = 0, hi = npairs;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(hit_off + mid) <= f) lo = mid;
        else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint32_t ld_text_word(const uint8_t *text, uint32_t i) {
    return __ldg(reinterpret_cast<const uint32_t *>(text + i));
}

constexpr uint32_t SCAN_LIMIT = 64;   // bytes scanned each way before the newline index decides

// Entry around text position pos: *b = 1 + last '\n' strictly before pos (none → 0,
// lib.rs:270-273), *e = first '\n' at or after pos (none → n - 1, lib.rs:266-269).
//
// With the line directory (every Reader builds it at open) no text is read at all: the block
// of pos gives the first newline offset of that block inside the sorted list (one random
// sector), and the nine offsets nl[lo - 1 .. lo + 8) loaded together (one or two more
// sectors, all in flight at once) hold both neighbours of pos unless the 256-byte block has
// more than eight newlines before pos.  Two dependent DRAM round trips per matching suffix,
// whatever the line length; the text scan it replaces walked 4 bytes per dependent load, crossed
// two or three sectors per 45-byte line and fell back to a 24-probe search over the whole list
// for every line end further than 64 bytes away (a quarter of the hits each way).
__device__ __forceinline__ void entry_bounds_dir(const DeviceChunk &ch, uint32_t pos, uint32_t *b_out, uint32_t *e_out) {
    const uint32_t *__restrict__ nl = ch.nl;
    const uint32_t L = ch.n_lines;
    if (L == 0) {                                            // a chunk without any '\n'
        *b_out = 0;
        *e_out = ch.n - 1;
        return;
    }
    const uint32_t blk = pos / LINE_BLOCK;
    const uint32_t lo = __ldg(ch.dir + blk);
    // unconditional loads at clamped indices: all nine are in flight together (a load guarded by
    // `idx < L` compiles to nine dependent branch-load-compare steps)
    uint32_t x[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) x[j] = __ldg(nl + min(lo + (uint32_t)j - 1u, L - 1u));
    uint32_t e = 0xFFFFFFFFu, b = 0;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const uint32_t idx = lo + (uint32_t)j - 1u;          // lo == 0, j == 0 wraps to 0xFFFFFFFF: out of range
        const uint32_t v   = idx < L ? x[j] : 0xFFFFFFFFu;
        b = v < pos ? v + 1 : b;                             // ascending: the last one below pos stands
        e = v < pos ? e : min(e, v);
    }
    if (e == 0xFFFFFFFFu) {
        if (lo + 8 >= L) {
            e = ch.n - 1;                                    // no '\n' at or after pos (lib.rs:266-269: None -> len - 1)
        } else {
            // more than eight newlines of this block lie before pos: search the rest of the block
            uint32_t l = lo + 8, h = __ldg(ch.dir + blk + 1);
            while (l < h) {
                const uint32_t mid = l + ((h - l) >> 1);
                if (__ldg(nl + mid) < pos) l = mid + 1;
                else h = mid;
            }
            e = l < L ? __ldg(nl + l) : ch.n - 1;
            b = __ldg(nl + l - 1) + 1;                       // l >= lo + 8 >= 1
        }
    }
    *b_out = b;
    *e_out = e;
}

// With the line records (the default) ONE 32-byte sector decides: the entry's end is the first
// set bit at or after pos in the block's newline map, else the stored next newline; its start
// follows the last set bit before pos, else it is the stored start of the entry open at the
// block's first byte.
__device__ __forceinline__ void entry_bounds_rec(const DeviceChunk &ch, uint32_t pos, uint32_t *b_out, uint32_t *e_out) {
    const uint32_t blk = pos / LINE_REC_BLOCK, off = pos % LINE_REC_BLOCK;
    const uint4 r0 = __ldg(ch.rec + 2 * (size_t)blk), r1 = __ldg(ch.rec + 2 * (size_t)blk + 1);
    const unsigned long long lo64 = ((unsigned long long)r0.w << 32) | r0.z, hi64 = ((unsigned long long)r1.y << 32) | r1.x;
    const uint32_t base = blk * LINE_REC_BLOCK;
    // at or after pos
    unsigned long long f_lo = off < 64 ? (lo64 & (~0ull << off)) : 0ull;
    unsigned long long f_hi = off < 64 ? hi64 : (hi64 & (~0ull << (off - 64)));
    uint32_t e = r0.y;
    if (f_lo) e = base + (uint32_t)__ffsll((long long)f_lo) - 1u;
    else if (f_hi) e = base + 64u + (uint32_t)__ffsll((long long)f_hi) - 1u;
    // strictly before pos
    unsigned long long b_lo = off < 64 ? (off ? (lo64 & (~0ull >> (64 - off))) : 0ull) : lo64;
    unsigned long long b_hi = off > 64 ? (hi64 & (~0ull >> (128 - off))) : 0ull;
    uint32_t b = r0.x;
    if (b_hi) b = base + 64u + (63u - (uint32_t)__clzll((long long)b_hi)) + 1u;
    else if (b_lo) b = base + (63u - (uint32_t)__clzll((long long)b_lo)) + 1u;
    *b_out = b;
    *e_out = e;
}

__device__ __forceinline__ void entry_bounds(const DeviceChunk &ch, uint32_t pos, uint32_t *b_out, uint32_t *e_out) {
    if (ch.rec != nullptr) {
        entry_bounds_rec(ch, pos, b_out, e_out);
        return;
    }
    if (ch.dir != nullptr) {
        entry_bounds_dir(ch, pos, b_out, e_out);
        return;
    }
    const uint8_t *__restrict__ text = ch.text;
    const uint32_t n = ch.n;
    const bool bounded = ch.nl != nullptr;
    // forward
    bool     have_e = false;
    uint32_t e = 0;
    {
        uint32_t i  = pos & ~3u;
        uint32_t eq = __vcmpeq4(ld_text_word(text, i), 0x0A0A0A0Au) & (0xFFFFFFFFu << (8 * (pos & 3u)));
        const uint32_t stop = bounded ? min(n, (pos & ~3u) + SCAN_LIMIT) : n;
        while (eq == 0) {
            i += 4;
            if (i >= stop) break;
            eq = __vcmpeq4(ld_text_word(text, i), 0x0A0A0A0Au);
        }
        if (eq) {
            e = i + ((__ffs(eq) - 1) >> 3);
            e = e < n ? e : n - 1;
            have_e = true;
        } else if (i >= n) {
            e = n - 1;
            have_e = true;
        }
    }
    // backward
    bool     have_b = false;
    uint32_t b = 0;
    if (pos == 0) {
        have_b = true;
    } else {
        const uint32_t q = pos - 1;
        uint32_t i  = q & ~3u;
        uint32_t eq = __vcmpeq4(ld_text_word(text, i), 0x0A0A0A0Au) & (0xFFFFFFFFu >> (8 * (3u - (q & 3u))));
        const uint32_t stop = (bounded && i > SCAN_LIMIT) ? i - SCAN_LIMIT : 0u;
        bool at_zero = false;
        while (eq == 0) {
            if (i == 0) { at_zero = true; break; }
            if (i <= stop) break;
            i -= 4;
            eq = __vcmpeq4(ld_text_word(text, i), 0x0A0A0A0Au);
        }
        if (eq) {
            b = i + ((31 - __clz(eq)) >> 3) + 1;
            have_b = true;
        } else if (at_zero) {
            have_b = true;
        }
    }
    if (!(have_e && have_b)) {
        // long line: j = number of newlines before pos (first index with nl[j] >= pos)
        uint32_t lo = 0, hi = ch.n_lines;
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (__ldg(ch.nl + mid) < pos) lo = mid + 1;
            else hi = mid;
        }
        if (!have_e) e = lo < ch.n_lines ? __ldg(ch.nl + lo) : n - 1;
        if (!have_b) b = lo > 0 ? __ldg(ch.nl + lo - 1) + 1 : 0u;
    }
    *b_out = b;
    *e_out = e;
}

__global__ void __launch_bounds__(256)
extract_kernel(const DeviceChunk *__restrict__ chunks, int nc, uint32_t pair_base, uint32_t npairs,
               const uint32_t *__restrict__ hit_off, const uint32_t *__restrict__ heavy_off,
               const uint32_t *__restrict__ lb, const uint32_t *__restrict__ cnt, uint32_t nhits, int sbits,
               uint32_t *__restrict__ line_start, uint32_t *__restrict__ line_end,
               uint64_t *__restrict__ heavy_keys, uint32_t *__restrict__ heavy_vals) {
    const uint32_t f = blockIdx.x * 256 + threadIdx.x;
    const uint32_t f0 = f & ~31u;               // the warp's first hit
    if (f0 >= nhits) return;
    // one binary search per warp (uniform loads), then every lane walks forward from the warp's
    // pair: 32 consecutive hits span few pairs; long runs of empty pairs fall back to a search
    uint32_t p = find_pair(hit_off, npairs, f0);
    if (f >= nhits) return;
    int steps = 0;
    while (__ldg(hit_off + p + 1) <= f && steps < 16) { ++p; ++steps; }
    if (__ldg(hit_off + p + 1) <= f) p = find_pair(hit_off, npairs, f);
    const uint32_t pair = pair_base + p;
    const DeviceChunk ch = chunks[pair % (uint32_t)nc];
    const uint32_t k    = f - __ldg(hit_off + p);
    const uint32_t pos  = (uint32_t)__ldg(ch.sa + __ldg(lb + pair) + k);
    uint32_t b, e;
    entry_bounds(ch, pos, &b, &e);
    line_start[f] = b;
    line_end[f]   = e;
    if (__ldg(cnt + pair) > MEDIUM_MAX) {     // heavy pair: its hits also go to the sort's input
        const uint32_t j = __ldg(heavy_off + p) + k;
        heavy_keys[j] = ((uint64_t)p << sbits) | b;
        heavy_vals[j] = f;
    }
}

// ------------------------------------------------------------------------------------
// dedup of light pairs: one warp per (query, chunk) pair with at most LIGHT_MAX matching
// suffixes.  A hit is kept iff no earlier hit of the pair (SA order) has the same entry
// start (lib.rs:262,274: the first hit in SA order stands for the entry).  The pair's entry
// starts sit in shared memory; every lane checks its hits against all earlier ones —
// O(h^2 / 32) compares per lane and no global sort for the bulk of a batch.
// ------------------------------------------------------------------------------------
constexpr int PD_WARPS = 8;

__global__ void __launch_bounds__(PD_WARPS * 32)
pair_dedup_kernel(const uint32_t *__restrict__ hit_off, const uint32_t *__restrict__ cnt, uint32_t pair_base,
                  uint32_t npairs, const uint32_t *__restrict__ line_start, uint32_t *__restrict__ flag) {
    __shared__ uint32_t s_start[PD_WARPS][LIGHT_MAX];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t p = blockIdx.x * PD_WARPS + warp;
    if (p >= npairs) return;
    const uint32_t h = __ldg(cnt + pair_base + p);
    if (h == 0 || h > LIGHT_MAX) return;
    const uint32_t off = __ldg(hit_off + p);
    uint32_t *st = s_start[warp];
    for (uint32_t i = lane; i < h; i += 32) st[i] = line_start[off + i];
    __syncwarp();
    for (uint32_t i = lane; i < h; i += 32) {
        const uint32_t mine = st[i];
        bool first = true;
        for (uint32_t g = 0; g < i; ++g)
            if (st[g] == mine) { first = false; break; }
        flag[off + i] = first ? (0x80000000u | mine) : 0u;
    }
}

// dedup of medium pairs: one CTA per pair, a hash set of the pair's entry starts in shared
// memory (open addressing, at most half full) keeps the smallest hit index per entry start;
// a hit is kept iff it is that smallest index.  O(h) per pair instead of a sort.
constexpr int PH_THREADS = 512;
constexpr uint32_t PH_SMALL_MAX = 2048;            // pairs up to here use the 32 KB table (several CTAs per SM)

__device__ __forceinline__ void hash_dedup_pair(uint32_t *keys, uint32_t *best, uint32_t h, uint32_t off,
                                                const uint32_t *__restrict__ line_start, uint32_t *__restrict__ flag) {
    uint32_t bits = 9;                                              // table of at least 512 slots, >= 2h
    while ((1u << bits) < 2 * h) ++bits;
    const uint32_t size = 1u << bits, mask = size - 1u;
    for (uint32_t i = threadIdx.x; i < size; i += PH_THREADS) { keys[i] = 0xFFFFFFFFu; best[i] = 0xFFFFFFFFu; }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < h; i += PH_THREADS) {
        const uint32_t st = line_start[off + i];
        uint32_t slot = (st * 2654435761u) >> (32 - bits);
        while (true) {
            const uint32_t prev = atomicCAS(&keys[slot], 0xFFFFFFFFu, st);
            if (prev == 0xFFFFFFFFu || prev == st) { atomicMin(&best[slot], i); break; }
            slot = (slot + 1) & mask;
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < h; i += PH_THREADS) {
        const uint32_t st = line_start[off + i];
        uint32_t slot = (st * 2654435761u) >> (32 - bits);
        while (keys[slot] != st) slot = (slot + 1) & mask;
        flag[off + i] = best[slot] == i ? (0x80000000u | st) : 0u;
    }
    __syncthreads();
}

// One CTA per medium pair.  Pairs above PH_SMALL_MAX are only queued (big_list) for the second,
// persistent launch, which owns a table four times as large.
__global__ void __launch_bounds__(PH_THREADS)
pair_dedup_hash_kernel(const uint32_t *__restrict__ med_list, const uint32_t *__restrict__ hit_off,
                       const uint32_t *__restrict__ cnt, uint32_t pair_base, const uint32_t *__restrict__ line_start,
                       uint32_t *__restrict__ flag, uint32_t *__restrict__ big_list, uint32_t *__restrict__ big_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t p = med_list[blockIdx.x];
    const uint32_t h = __ldg(cnt + pair_base + p);
    if (h > PH_SMALL_MAX) {
        if (threadIdx.x == 0) big_list[atomicAdd(big_count, 1u)] = p;
        return;
    }
    uint32_t *keys = reinterpret_cast<uint32_t *>(smem_raw);      // entry start, 0xFFFFFFFF = empty
    hash_dedup_pair(keys, keys + 2 * PH_SMALL_MAX, h, __ldg(hit_off + p), line_start, flag);
}

__global__ void __launch_bounds__(PH_THREADS)
pair_dedup_hash_big_kernel(const uint32_t *__restrict__ big_list, const uint32_t *__restrict__ big_count,
                           uint32_t *__restrict__ cursor, const uint32_t *__restrict__ hit_off,
                           const uint32_t *__restrict__ cnt, uint32_t pair_base, const uint32_t *__restrict__ line_start,
                           uint32_t *__restrict__ flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_next;
    uint32_t *keys = reinterpret_cast<uint32_t *>(smem_raw);
    const uint32_t total = *big_count;
    while (true) {
        if (threadIdx.x == 0) s_next = atomicAdd(cursor, 1u);
        __syncthreads();
        const uint32_t j = s_next;
        __syncthreads();
        if (j >= total) return;
        const uint32_t p = big_list[j];
        hash_dedup_pair(keys, keys + 2 * MEDIUM_MAX, __ldg(cnt + pair_base + p), __ldg(hit_off + p), line_start, flag);
    }
}

// ------------------------------------------------------------------------------------
// dedup: after the stable sort by (pair, entry start) the head of every run is the entry's
// first hit in SA order; flag it at its original hit index.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mark_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t nhits, int sbits,
            uint32_t *__restrict__ flag) {
    const uint32_t k = blockIdx.x * 256 + threadIdx.x;
    if (k >= nhits) return;
    const uint64_t cur = keys[k];
    if (k == 0 || keys[k - 1] != cur)
        flag[vals[k]] = 0x80000000u | (uint32_t)(cur & ((1ull << sbits) - 1ull));
}

constexpr int CP_THREADS = 256;
constexpr int CP_IPT     = 8;
constexpr int CP_TILE    = CP_THREADS * CP_IPT;

__global__ void __launch_bounds__(CP_THREADS)
flag_reduce_kernel(const uint32_t *__restrict__ flag, uint32_t nhits, uint32_t *__restrict__ tile_sum) {
    __shared__ uint32_t s_warp[CP_THREADS / 32];
    const uint32_t base = blockIdx.x * CP_TILE + threadIdx.x * CP_IPT;
    uint32_t c = 0;
#pragma unroll
    for (int e = 0; e < CP_IPT; ++e)
        if (base + e < nhits) c += flag[base + e] >> 31;
    uint32_t total;
    (void)block_excl_sum_256(c, s_warp, &total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
tile_scan_kernel(uint32_t *__restrict__ tile_sum, uint32_t tiles, uint32_t *__restrict__ total_out) {
    __shared__ uint32_t s_part[SCAN_THREADS];
    const uint32_t per = (tiles + SCAN_THREADS - 1) / SCAN_THREADS;
    const uint32_t lo  = min(tiles, threadIdx.x * per), hi = min(tiles, lo + per);
    uint32_t sum = 0;
    for (uint32_t t = lo; t < hi; ++t) sum += tile_sum[t];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele over 1024 partials
    for (int o = 1; o < SCAN_THREADS; o <<= 1) {
        uint32_t y = threadIdx.x >= (uint32_t)o ? s_part[threadIdx.x - o] : 0u;
        __syncthreads();
        s_part[threadIdx.x] += y;
        __syncthreads();
    }
    uint32_t run = s_part[threadIdx.x] - sum;
    for (uint32_t t = lo; t < hi; ++t) {
        uint32_t v  = tile_sum[t];
        tile_sum[t] = run;
        run += v;
    }
    if (threadIdx.x == SCAN_THREADS - 1) *total_out = s_part[SCAN_THREADS - 1];
}

__global__ void __launch_bounds__(CP_THREADS)
compact_kernel(const uint32_t *__restrict__ flag, const uint32_t *__restrict__ line_end,
               const uint32_t *__restrict__ tile_prefix, const uint32_t *__restrict__ hit_off, uint32_t hit_base,
               const DeviceChunk *__restrict__ chunks, int nc, uint32_t pair_base, uint32_t npairs, uint32_t nhits,
               uint32_t *__restrict__ pair_first, int32_t *__restrict__ out_chunk,
               uint32_t *__restrict__ out_start, uint32_t *__restrict__ out_end) {
    __shared__ uint32_t s_warp[CP_THREADS / 32];
    // the tile's kept tuples, staged so that they leave the SM as one contiguous run per array
    // (a thread's own tuples are 8 apart from its neighbour's: one sector per 4-byte store otherwise)
    __shared__ uint32_t s_start[CP_TILE], s_end[CP_TILE];
    __shared__ int32_t  s_chunk[CP_TILE];
    const uint32_t base = blockIdx.x * CP_TILE + threadIdx.x * CP_IPT;
    uint32_t fl[CP_IPT], le[CP_IPT];
    uint32_t c = 0;
    if (base + CP_IPT <= nhits) {
        // 8 consecutive words, 32-byte aligned: two 16-byte loads per array
        const uint4 f0 = __ldg(reinterpret_cast<const uint4 *>(flag + base)), f1 = __ldg(reinterpret_cast<const uint4 *>(flag + base) + 1);
        const uint4 e0 = __ldg(reinterpret_cast<const uint4 *>(line_end + base)), e1 = __ldg(reinterpret_cast<const uint4 *>(line_end + base) + 1);
        fl[0] = f0.x; fl[1] = f0.y; fl[2] = f0.z; fl[3] = f0.w; fl[4] = f1.x; fl[5] = f1.y; fl[6] = f1.z; fl[7] = f1.w;
        le[0] = e0.x; le[1] = e0.y; le[2] = e0.z; le[3] = e0.w; le[4] = e1.x; le[5] = e1.y; le[6] = e1.z; le[7] = e1.w;
    } else {
#pragma unroll
        for (int e = 0; e < CP_IPT; ++e) {
            fl[e] = (base + e < nhits) ? flag[base + e] : 0u;
            le[e] = (base + e < nhits) ? line_end[base + e] : 0u;
        }
    }
#pragma unroll
    for (int e = 0; e < CP_IPT; ++e) c += fl[e] >> 31;
    uint32_t total;
    const uint32_t local = block_excl_sum_256(c, s_warp, &total);
    const uint32_t tile_o = tile_prefix[blockIdx.x];
    if (base < nhits) {
        uint32_t o = local;
        uint32_t p = find_pair(hit_off, npairs, base + hit_base);
        int32_t  gid = chunks[(pair_base + p) % (uint32_t)nc].global_id;
#pragma unroll
        for (int e = 0; e < CP_IPT; ++e) {
            const uint32_t f = base + e;
            if (f >= nhits) break;
            if (f + hit_base >= __ldg(hit_off + p + 1)) {
                do { ++p; } while (f + hit_base >= __ldg(hit_off + p + 1));       // skips pairs without hits
                gid = chunks[(pair_base + p) % (uint32_t)nc].global_id;
            }
            if (f + hit_base == __ldg(hit_off + p)) pair_first[p] = tile_o + o;   // first hit of pair p: its output offset
            if (fl[e] >> 31) {
                s_chunk[o] = gid;
                s_start[o] = fl[e] & 0x7FFFFFFFu;
                s_end[o]   = le[e];
                ++o;
            }
        }
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < total; t += CP_THREADS) {
        out_chunk[tile_o + t] = s_chunk[t];
        out_start[tile_o + t] = s_start[t];
        out_end[tile_o + t]   = s_end[t];
    }
}

// Deferred compaction (the distributed search): first only the output offset of every pair's
// first entry (pair_first_kernel → entry_offsets_kernel give the per-pair entry offsets without
// writing a tuple), later compact_place_kernel writes entry i of pair p to pair_dst[p] + i of
// arrays that may sit in another GPU's memory.
__global__ void __launch_bounds__(CP_THREADS)
pair_first_kernel(const uint32_t *__restrict__ flag, const uint32_t *__restrict__ tile_prefix,
                  const uint32_t *__restrict__ hit_off, uint32_t npairs, uint32_t nhits, uint32_t *__restrict__ pair_first) {
    __shared__ uint32_t s_warp[CP_THREADS / 32];
    const uint32_t base = blockIdx.x * CP_TILE + threadIdx.x * CP_IPT;
    uint32_t fl[CP_IPT];
    uint32_t c = 0;
#pragma unroll
    for (int e = 0; e < CP_IPT; ++e) {
        fl[e] = (base + e < nhits) ? flag[base + e] : 0u;
        c += fl[e] >> 31;
    }
    uint32_t total;
    uint32_t o = block_excl_sum_256(c, s_warp, &total) + tile_prefix[blockIdx.x];
    if (base >= nhits) return;
    uint32_t p = find_pair(hit_off, npairs, base);
#pragma unroll
    for (int e = 0; e < CP_IPT; ++e) {
        const uint32_t f = base + e;
        if (f >= nhits) break;
        while (f >= __ldg(hit_off + p + 1)) ++p;
        if (f == __ldg(hit_off + p)) pair_first[p] = o;
        o += fl[e] >> 31;
    }
}

__global__ void __launch_bounds__(CP_THREADS)
compact_place_kernel(const uint32_t *__restrict__ flag, const uint32_t *__restrict__ line_end,
                     const uint32_t *__restrict__ tile_prefix, const uint32_t *__restrict__ hit_off,
                     const uint32_t *__restrict__ entry_off, const uint32_t *__restrict__ pair_dst,
                     const DeviceChunk *__restrict__ chunks, int nc, uint32_t npairs, uint32_t nhits,
                     int32_t *out_chunk, uint32_t *out_start, uint32_t *out_end) {
    __shared__ uint32_t s_warp[CP_THREADS / 32];
    __shared__ uint32_t s_start[CP_TILE], s_end[CP_TILE], s_dst[CP_TILE];
    __shared__ int32_t  s_chunk[CP_TILE];
    const uint32_t base = blockIdx.x * CP_TILE + threadIdx.x * CP_IPT;
    uint32_t fl[CP_IPT];
    uint32_t c = 0;
#pragma unroll
    for (int e = 0; e < CP_IPT; ++e) {
        fl[e] = (base + e < nhits) ? flag[base + e] : 0u;
        c += fl[e] >> 31;
    }
    uint32_t total;
    const uint32_t tile_o = tile_prefix[blockIdx.x];
    uint32_t l = block_excl_sum_256(c, s_warp, &total);     // index among the tile's kept entries
    if (base < nhits && c) {
        uint32_t p = find_pair(hit_off, npairs, base);
#pragma unroll
        for (int e = 0; e < CP_IPT; ++e) {
            const uint32_t f = base + e;
            if (f >= nhits) break;
            if (fl[e] >> 31) {
                while (f >= __ldg(hit_off + p + 1)) ++p;
                s_start[l] = fl[e] & 0x7FFFFFFFu;
                s_end[l]   = line_end[f];
                s_dst[l]   = __ldg(pair_dst + p) + (tile_o + l - __ldg(entry_off + p));
                s_chunk[l] = chunks[p % (uint32_t)nc].global_id;
                ++l;
            }
        }
    }
    __syncthreads();
    // consecutive threads write consecutive entries: runs of one pair are contiguous at the
    // destination, so the stores leave as full 32-byte sectors (NVLink-friendly)
    for (uint32_t t = threadIdx.x; t < total; t += CP_THREADS) {
        const uint32_t d = s_dst[t];
        if (out_chunk) out_chunk[d] = s_chunk[t];
        out_start[d] = s_start[t];
        out_end[d]   = s_end[t];
    }
}

// ------------------------------------------------------------------------------------
// Small-batch path: ONE launch answers a handful of (query, chunk) pairs end to end.
// CTA p finds pair p's SA range with two concurrent 512-ary searches (lower bound in warps
// 0-15, upper bound in warps 16-31: 4 rounds of independent probes at n = 2^29 instead of
// 2 x 29 dependent ones); the last CTA to finish extracts the entries of all pairs, dedups
// them with a bitonic sort in shared memory and writes header + tuples to `out`, which is
// mapped pinned host memory: the host sees the result without a copy or a stream sync.
// status = 1 when the pairs have more than SMALL_CAP matching suffixes (general path).
// ------------------------------------------------------------------------------------
constexpr int SMALL_THREADS = 1024;
constexpr int SMALL_CAP     = 8192;   // matching suffixes handled in shared memory

struct SmallHeader {
    uint32_t seq;         // written last: the host polls it
    uint32_t status;      // 0 = answered, 1 = too many hits (use the general path)
    uint32_t n_hits;
    uint32_t n_entries;
    uint32_t entry_off[SMALL_MAX_PAIRS + 1];
    uint32_t pad[3];
    long long query_off[SMALL_MAX_QUERIES + 1];
};
// Result block: header, then chunk / start / end arrays of SMALL_CAP each.
constexpr size_t SMALL_OUT_BYTES = sizeof(SmallHeader) + 3 * (size_t)SMALL_CAP * sizeof(uint32_t);

struct SmallSmem {
    uint64_t key[SMALL_CAP];       // (pair << 43) | (entry start << 13) | hit index
    uint32_t end[SMALL_CAP];       // entry end, by hit index
    uint32_t start[SMALL_CAP];     // bit 31 = first hit of its entry, low bits = entry start
    uint32_t lb[SMALL_MAX_PAIRS], cnt[SMALL_MAX_PAIRS], off[SMALL_MAX_PAIRS + 1], pair_entries[SMALL_MAX_PAIRS];
    uint32_t warp_sum[SMALL_THREADS / 32];
    uint32_t total;
    uint32_t count[2];
    uint32_t is_last;
    uint8_t  pat[SMALL_PAT_BYTES];
};

// One thread compares the suffix at s with P: -1 / 0 (P is a prefix) / +1, as cmp_suffix.
__device__ __forceinline__ int cmp_suffix_thread(const uint8_t *__restrict__ text, uint32_t n, uint32_t s,
                                                 const uint8_t *P, uint32_t m) {
    const uint32_t avail = n - s;
    for (uint32_t b = 0; b < m; ++b) {
        if (b >= avail) return -1;
        const uint32_t tc = __ldg(text + s + b), pc = P[b];
        if (tc != pc) return tc < pc ? -1 : 1;
    }
    return 0;
}

__global__ void __launch_bounds__(SMALL_THREADS, 1)
small_search_kernel(const DeviceChunk *__restrict__ chunks, int nc, const __grid_constant__ SmallPatterns pats,
                    uint32_t npairs, uint32_t seq, uint32_t *__restrict__ lb_out, uint32_t *__restrict__ cnt_out,
                    uint32_t *__restrict__ ticket, unsigned char *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmallSmem &s = *reinterpret_cast<SmallSmem *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;

    // ---- phase 1: this CTA's pair -------------------------------------------------------
    {
        const uint32_t pair = blockIdx.x;
        const uint32_t q = pair / (uint32_t)nc, c = pair % (uint32_t)nc;
        const uint32_t m = pats.off[q + 1] - pats.off[q];
        for (uint32_t i = tid; i < m; i += SMALL_THREADS) s.pat[i] = pats.bytes[pats.off[q] + i];
        if (tid < 2) s.count[tid] = 0;
        __syncthreads();
        const uint8_t *__restrict__ text = chunks[c].text;
        const int32_t *__restrict__ sa   = chunks[c].sa;
        const uint32_t n = chunks[c].n;
        constexpr uint32_t T = SMALL_THREADS / 2;
        const uint32_t which = tid / T;        // 0: lower bound (first slot with cmp >= 0), 1: upper (first with cmp > 0)
        const uint32_t t     = tid % T;
        uint32_t lo = 0, hi = n;               // the answer lies in [lo, hi]
        if (chunks[c].bucket != nullptr && m >= 2) {          // inside the bucket of the first two pattern bytes
            const uint32_t key = ((uint32_t)s.pat[0] << 8) | s.pat[1];
            lo = min(__ldg(chunks[c].bucket + key), n);
            hi = min(max(__ldg(chunks[c].bucket + key + 1), lo), n);
        }
        while (true) {                         // both halves iterate in lockstep (block barriers)
            const uint32_t R = hi - lo;
            bool below = false;                // predicate "boundary is past my pivot"
            uint32_t step = 1, pivots = 0;
            if (R > 0) {
                step   = (R + T - 1) / T;
                pivots = (R + step - 1) / step;            // <= T
                if (t < pivots) {
                    const int cmp = cmp_suffix_thread(text, n, (uint32_t)__ldg(sa + lo + t * step), s.pat, m);
                    below = which == 0 ? (cmp < 0) : (cmp <= 0);
                }
            }
            const uint32_t votes = __popc(__ballot_sync(0xffffffffu, below));
            if (lane == 0 && votes) atomicAdd(&s.count[which], votes);
            __syncthreads();
            const uint32_t c_true = s.count[which];
            __syncthreads();
            if (tid < 2) s.count[tid] = 0;
            if (R > 0) {
                // pivots 0..c_true-1 are below the boundary, pivot c_true (if any) is not:
                // the range shrinks to fewer than `step` slots; step == 1 ends the search
                if (c_true == 0) {
                    hi = lo;                   // pivot 0 is the slot lo itself
                } else {
                    if (c_true < pivots) hi = lo + c_true * step;
                    lo = lo + (c_true - 1) * step + 1;
                }
            }
            // uniform exit: both searches must have converged (also orders the reset of the
            // counters before the next round's votes)
            if (__syncthreads_and(hi == lo ? 1 : 0)) break;
        }
        if (t == 0) s.lb[which] = lo;          // reuse s.lb[0..1] as scratch for the two boundaries
        __syncthreads();
        if (tid == 0) {
            const uint32_t lbv = s.lb[0], ubv = s.lb[1];
            lb_out[pair]  = lbv;
            cnt_out[pair] = ubv > lbv ? ubv - lbv : 0u;
            __threadfence();
            const uint32_t prev = atomicAdd(ticket, 1u);
            s.is_last = (prev == npairs - 1) ? 1u : 0u;
        }
        __syncthreads();
        if (!s.is_last) return;
        __threadfence();
    }

    // ---- phase 2 (last CTA): extraction, dedup, compaction for all pairs ----------------------
    SmallHeader *hdr  = reinterpret_cast<SmallHeader *>(out);
    int32_t  *o_chunk = reinterpret_cast<int32_t *>(out + sizeof(SmallHeader));
    uint32_t *o_start = reinterpret_cast<uint32_t *>(o_chunk + SMALL_CAP);
    uint32_t *o_end   = o_start + SMALL_CAP;
    if (tid < npairs) {
        s.lb[tid]  = ld_volatile_u32(lb_out + tid);
        s.cnt[tid] = ld_volatile_u32(cnt_out + tid);
    }
    if (tid < SMALL_MAX_PAIRS) s.pair_entries[tid] = 0;
    __syncthreads();
    if (tid == 0) {
        *ticket = 0;                            // ready for the next launch
        uint64_t run = 0;
        for (uint32_t p = 0; p < npairs; ++p) {
            s.off[p] = (uint32_t)min(run, (uint64_t)0xFFFFFFFFu);
            run += s.cnt[p];
        }
        s.off[npairs] = (uint32_t)min(run, (uint64_t)0xFFFFFFFFu);
        s.total = (uint32_t)min(run, (uint64_t)0xFFFFFFFFu);
    }
    __syncthreads();
    const uint32_t H = s.total;
    if (H > SMALL_CAP) {
        if (tid == 0) {
            hdr->status = 1; hdr->n_hits = H; hdr->n_entries = 0;
            __threadfence_system();
            *reinterpret_cast<volatile uint32_t *>(&hdr->seq) = seq;
        }
        return;
    }
    uint32_t P2 = 1;
    while (P2 < H) P2 <<= 1;

    for (uint32_t f = tid; f < P2; f += SMALL_THREADS) {
        uint64_t key = ~0ull;
        if (f < H) {
            uint32_t p = 0;
            while (s.off[p + 1] <= f) ++p;
            const DeviceChunk ch = chunks[p % (uint32_t)nc];
            const uint32_t pos   = (uint32_t)__ldg(ch.sa + s.lb[p] + (f - s.off[p]));
            uint32_t b, e;
            entry_bounds(ch, pos, &b, &e);
            s.end[f]   = e;
            s.start[f] = 0;
            key = ((uint64_t)p << 43) | ((uint64_t)b << 13) | f;
        }
        s.key[f] = key;
    }
    __syncthreads();

    constexpr uint32_t ALL_PAIRS_MAX = 768;
    if (H <= ALL_PAIRS_MAX) {
        // few hits (the common single query): an entry's first hit is the one with no earlier hit
        // (smaller hit index) of the same (pair, entry start) — H^2 / 1024 key compares per thread
        // in shared memory instead of ~log^2(H) block-wide barriers of the sort below
        for (uint32_t f = tid; f < H; f += SMALL_THREADS) {
            const uint64_t mine = s.key[f] >> 13;
            bool first = true;
            for (uint32_t g = 0; g < f; ++g)
                if ((s.key[g] >> 13) == mine) { first = false; break; }
            if (first) s.start[f] = 0x80000000u | (uint32_t)(mine & 0x3FFFFFFFu);
        }
        __syncthreads();
    } else {
    // bitonic sort by (pair, entry start, hit index); the head of every (pair, entry start)
    // run is the entry's first hit in SA order
    for (uint32_t k = 2; k <= P2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = tid; i < P2; i += SMALL_THREADS) {
                const uint32_t x = i ^ j;
                if (x > i) {
                    const uint64_t a = s.key[i], b = s.key[x];
                    if ((a > b) == ((i & k) == 0)) { s.key[i] = b; s.key[x] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t k = tid; k < H; k += SMALL_THREADS) {
        const uint64_t cur = s.key[k];
        if (k == 0 || (s.key[k - 1] >> 13) != (cur >> 13))
            s.start[(uint32_t)cur & (SMALL_CAP - 1)] = 0x80000000u | (uint32_t)((cur >> 13) & 0x3FFFFFFFu);
    }
    __syncthreads();
    }

    // compaction in hit order = (query, chunk, SA order).  The kept tuples are first packed in
    // shared memory (the sort keys are dead by now) and then written by consecutive threads:
    // `out` is host memory behind PCIe, where one 128-byte store per warp instead of 32
    // scattered 4-byte ones is the difference between ~15 and ~500 bus transactions per array.
    constexpr int PER = SMALL_CAP / SMALL_THREADS;
    uint32_t *st_start = reinterpret_cast<uint32_t *>(s.key);
    uint32_t *st_end   = st_start + SMALL_CAP;
    const uint32_t base = tid * PER;
    uint32_t kept = 0;
    uint32_t vs[PER], ve[PER];
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        vs[e] = (base + e < H) ? s.start[base + e] : 0u;
        ve[e] = (base + e < H) ? s.end[base + e] : 0u;
        kept += vs[e] >> 31;
    }
    uint32_t incl = kept;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += y;
    }
    if (lane == 31) s.warp_sum[warp] = incl;
    __syncthreads();                            // also: every thread has read its sort keys / flags
    uint32_t pre = 0, tot = 0;
    for (uint32_t w = 0; w < SMALL_THREADS / 32; ++w) {
        const uint32_t t = s.warp_sum[w];
        if (w < warp) pre += t;
        tot += t;
    }
    uint32_t o = pre + incl - kept;
    if (base < H) {
        uint32_t p = 0;
        while (s.off[p + 1] <= base) ++p;
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const uint32_t f = base + e;
            if (f >= H) break;
            while (s.off[p + 1] <= f) ++p;
            if (vs[e] >> 31) {
                st_start[o] = vs[e] & 0x7FFFFFFFu;
                st_end[o]   = ve[e];
                atomicAdd(&s.pair_entries[p], 1u);
                ++o;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {                             // entries before every pair (reuses s.off: the hit offsets are dead)
        uint32_t run = 0;
        for (uint32_t p = 0; p < npairs; ++p) {
            const uint32_t c = s.pair_entries[p];
            s.off[p] = run;
            run += c;
        }
        s.off[npairs] = run;
    }
    __syncthreads();
    for (uint32_t t = tid; t < tot; t += SMALL_THREADS) {
        uint32_t p = 0;
        while (s.off[p + 1] <= t) ++p;          // <= 64 pairs
        o_chunk[t] = chunks[p % (uint32_t)nc].global_id;
        o_start[t] = st_start[t];
        o_end[t]   = st_end[t];
    }
    if (tid <= npairs) hdr->entry_off[tid] = s.off[tid];
    if (tid <= npairs / (uint32_t)nc) hdr->query_off[tid] = s.off[tid * (uint32_t)nc];
    if (tid == 0) { hdr->status = 0; hdr->n_hits = H; hdr->n_entries = tot; }
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();                 // tuples + header fields are visible to the host first
        *reinterpret_cast<volatile uint32_t *>(&hdr->seq) = seq;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------
int Searcher::init(int device) {
    if (device_ >= 0) return PSS_OK;
    if (device < 0) device = default_device();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(PSS_ERR_CUDA, "no CUDA device available (libpss_b200 has no CPU fallback)");
    if (device >= ndev) return fail(PSS_ERR_ARG, "device index out of range");
    PSS_CUDA_TRY(cudaSetDevice(device));
    device_ = device;
    PSS_CUDA_TRY(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    for (auto &e : ev_) PSS_CUDA_TRY(cudaEventCreate(&e));
    PSS_CUDA_TRY(cudaMalloc(&d_scalar_, 16 * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMemset(d_scalar_, 0, 16 * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMallocHost(&h_scalar_, 16 * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaHostAlloc(&h_small_out_, SMALL_OUT_BYTES, cudaHostAllocMapped));
    std::memset(h_small_out_, 0, SMALL_OUT_BYTES);
    PSS_CUDA_TRY(cudaFuncSetAttribute(small_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(SmallSmem)));
    PSS_CUDA_TRY(cudaFuncSetAttribute(pair_dedup_hash_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(4 * MEDIUM_MAX * sizeof(uint32_t))));
    small_path_ = true;
    if (const char *e = std::getenv("PSS_SMALL_PATH")) small_path_ = std::atoi(e) != 0;
    if (const char *e = std::getenv("PSS_BOUNDS_GROUP")) bounds_group_ = std::atoi(e);
    PSS_TRY(sorter_.init(device_));
    PSS_TRY(ensure_pairs(SMALL_MAX_PAIRS, SMALL_MAX_QUERIES));
    return PSS_OK;
}

void Searcher::release() {
    if (device_ < 0) return;
    cudaSetDevice(device_);
    cudaFree(d_chunks_);
    cudaFree(d_lb_); cudaFree(d_cnt_); cudaFree(d_hit_off_); cudaFree(d_pair_first_);
    cudaFree(d_entry_off_); cudaFree(d_query_off_); cudaFree(d_heavy_off_); cudaFree(d_start_); cudaFree(d_med_list_); cudaFree(d_big_list_);
    cudaFree(d_scan_part_);
    d_scan_part_ = nullptr;
    if (h_cnt_) cudaFreeHost(h_cnt_);
    if (h_hit_off_) cudaFreeHost(h_hit_off_);
    if (h_heavy_off_) cudaFreeHost(h_heavy_off_);
    if (h_med_list_) cudaFreeHost(h_med_list_);
    cudaFree(d_keys_); cudaFree(d_keys_alt_); cudaFree(d_vals_); cudaFree(d_vals_alt_);
    cudaFree(d_end_); cudaFree(d_flag_); cudaFree(d_tile_sum_); cudaFree(d_scalar_);
    cudaFree(d_out_chunk_); cudaFree(d_out_start_); cudaFree(d_out_end_);
    if (h_small_out_) cudaFreeHost(h_small_out_);
    if (h_scalar_) cudaFreeHost(h_scalar_);
    for (auto &e : ev_)
        if (e) cudaEventDestroy(e);
    if (stream_) cudaStreamDestroy(stream_);
    sorter_.release();
    d_chunks_ = nullptr;
    d_lb_ = d_cnt_ = d_hit_off_ = d_pair_first_ = d_entry_off_ = d_heavy_off_ = d_start_ = d_med_list_ = d_big_list_ = nullptr;
    d_query_off_ = nullptr;
    h_cnt_ = h_hit_off_ = h_heavy_off_ = h_med_list_ = nullptr;
    d_keys_ = d_keys_alt_ = nullptr;
    d_vals_ = d_vals_alt_ = d_end_ = d_flag_ = d_tile_sum_ = d_scalar_ = h_scalar_ = nullptr;
    d_out_chunk_ = nullptr;
    d_out_start_ = d_out_end_ = nullptr;
    h_small_out_ = nullptr;
    for (auto &e : ev_) e = nullptr;
    stream_ = nullptr;
    pair_cap_ = query_cap_ = hit_cap_ = heavy_cap_ = out_cap_ = 0;
    chunks_.clear();
    device_ = -1;
}

int Searcher::set_chunks(const std::vector<DeviceChunk> &chunks) {
    PSS_CUDA_TRY(cudaSetDevice(device_));
    cudaFree(d_chunks_);
    d_chunks_ = nullptr;
    chunks_   = chunks;
    if (!chunks_.empty()) {
        PSS_CUDA_TRY(cudaMalloc(&d_chunks_, chunks_.size() * sizeof(DeviceChunk)));
        PSS_CUDA_TRY(cudaMemcpy(d_chunks_, chunks_.data(), chunks_.size() * sizeof(DeviceChunk), cudaMemcpyHostToDevice));
    }
    return PSS_OK;
}

int Searcher::build_newline_index(const uint8_t *d_text, uint32_t n, uint32_t **d_nl, uint32_t *n_lines) {
    *d_nl = nullptr;
    *n_lines = 0;
    if (n == 0) return PSS_OK;
    PSS_CUDA_TRY(cudaSetDevice(device_));
    if (reinterpret_cast<uintptr_t>(d_text) & 15) return fail(PSS_ERR_ARG, "chunk text must be 16-byte aligned on the device");
    const uint32_t tiles = (uint32_t)div_up(n, NL_TILE);
    uint32_t *d_tiles = nullptr;
    PSS_CUDA_TRY(cudaMalloc(&d_tiles, (size_t)tiles * sizeof(uint32_t)));
    struct Free { uint32_t *p; ~Free() { cudaFree(p); } } guard{d_tiles};
    newline_count_kernel<<<tiles, NL_THREADS, 0, stream_>>>(d_text, n, d_tiles);
    PSS_LAUNCH_CHECK();
    tile_scan_kernel<<<1, SCAN_THREADS, 0, stream_>>>(d_tiles, tiles, d_scalar_ + 4);
    PSS_LAUNCH_CHECK();
    PSS_CUDA_TRY(cudaMemcpyAsync(h_scalar_ + 4, d_scalar_ + 4, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream_));
    PSS_CUDA_TRY(cudaStreamSynchronize(stream_));
    const uint32_t L = h_scalar_[4];
    uint32_t *nl = nullptr;
    PSS_CUDA_TRY(cudaMalloc(&nl, (size_t)std::max<uint32_t>(L, 1) * sizeof(uint32_t)));
    newline_fill_kernel<<<tiles, NL_THREADS, 0, stream_>>>(d_text, n, d_tiles, nl);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    if (e != cudaSuccess) {
        cudaFree(nl);
        return fail(PSS_ERR_CUDA, std::string("newline index: ") + cudaGetErrorString(e));
    }
    *d_nl = nl;
    *n_lines = L;
    return PSS_OK;
}

int Searcher::build_line_directory(const uint32_t *d_nl, uint32_t n_lines, uint32_t n, uint32_t **d_dir) {
    *d_dir = nullptr;
    if (n == 0 || !d_nl) return PSS_OK;
    PSS_CUDA_TRY(cudaSetDevice(device_));
    const uint32_t n_dir = (uint32_t)div_up(n, LINE_BLOCK) + 1;   // dir[ceil(n / LINE_BLOCK)] = n_lines
    uint32_t *dir = nullptr;
    PSS_CUDA_TRY(cudaMalloc(&dir, (size_t)n_dir * sizeof(uint32_t)));
    line_directory_kernel<<<(unsigned)div_up(n_dir, 256), 256, 0, stream_>>>(d_nl, n_lines, n_dir, dir);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    if (e != cudaSuccess) {
        cudaFree(dir);
        return fail(PSS_ERR_CUDA, std::string("line directory: ") + cudaGetErrorString(e));
    }
    *d_dir = dir;
    return PSS_OK;
}

int Searcher::build_line_records(const uint32_t *d_nl, uint32_t n_lines, uint32_t n, uint4 **d_rec) {
    *d_rec = nullptr;
    if (n == 0 || !d_nl) return PSS_OK;
    PSS_CUDA_TRY(cudaSetDevice(device_));
    const uint32_t n_blocks = (uint32_t)div_up(n, LINE_REC_BLOCK);
    uint4 *rec = nullptr;
    PSS_CUDA_TRY(cudaMalloc(&rec, (size_t)n_blocks * 2 * sizeof(uint4)));
    line_records_kernel<<<(unsigned)div_up(n_blocks, 256), 256, 0, stream_>>>(d_nl, n_lines, n, n_blocks, rec);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    if (e != cudaSuccess) {
        cudaFree(rec);
        return fail(PSS_ERR_CUDA, std::string("line records: ") + cudaGetErrorString(e));
    }
    *d_rec = rec;
    return PSS_OK;
}

int Searcher::build_prefix_buckets(const uint8_t *d_text, const int32_t *d_sa, uint32_t n, uint32_t **d_bucket) {
    *d_bucket = nullptr;
    if (n == 0) return PSS_OK;
    PSS_CUDA_TRY(cudaSetDevice(device_));
    uint32_t *b = nullptr;
    PSS_CUDA_TRY(cudaMalloc(&b, 65537 * sizeof(uint32_t)));
    cudaMemsetAsync(b, 0, 65537 * sizeof(uint32_t), stream_);
    prefix_bucket_kernel<<<(unsigned)div_up(n, 256), 256, 0, stream_>>>(d_text, d_sa, n, b);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    if (e != cudaSuccess) {
        cudaFree(b);
        return fail(PSS_ERR_CUDA, std::string("prefix buckets: ") + cudaGetErrorString(e));
    }
    *d_bucket = b;
    return PSS_OK;
}

int Searcher::ensure_pairs(int64_t npairs, int64_t nq) {
    if (npairs > pair_cap_) {
        cudaFree(d_lb_); cudaFree(d_cnt_); cudaFree(d_hit_off_); cudaFree(d_pair_first_); cudaFree(d_entry_off_);
        cudaFree(d_heavy_off_); cudaFree(d_med_list_); cudaFree(d_big_list_); cudaFree(d_scan_part_);
        d_scan_part_ = nullptr;
        if (h_cnt_) cudaFreeHost(h_cnt_);
        if (h_hit_off_) cudaFreeHost(h_hit_off_);
        if (h_heavy_off_) cudaFreeHost(h_heavy_off_);
        if (h_med_list_) cudaFreeHost(h_med_list_);
        d_lb_ = d_cnt_ = d_hit_off_ = d_pair_first_ = d_entry_off_ = d_heavy_off_ = d_med_list_ = d_big_list_ = nullptr;
        h_cnt_ = h_hit_off_ = h_heavy_off_ = h_med_list_ = nullptr;
        pair_cap_ = 0;
        int64_t cap = std::max<int64_t>(npairs + npairs / 4, 1024);
        PSS_CUDA_TRY(cudaMalloc(&d_lb_, cap * sizeof(uint32_t)));
        PSS_CUDA_TRY(cudaMalloc(&d_cnt_, cap * sizeof(uint32_t)));
        PSS_CUDA_TRY(cudaMalloc(&d_hit_off_, (cap + 1) * sizeof(uint32_t)));
        PSS_CUDA_TRY(cudaMalloc(&d_pair_first_, cap * sizeof(uint32_t)));
        PSS_CUDA_TRY(cudaMalloc(&d_entry_off_, (cap + 1) * sizeof(uint32_t)));
        PSS_CUDA_TRY(cudaMalloc(&d_heavy_off_, (cap + 1) * sizeof(uint32_t)));
        PSS_CUDA_TRY(cudaMalloc(&d_med_list_, (cap + 1) * sizeof(uint32_t)));
        PSS_CUDA_TRY(cudaMalloc(&d_big_list_, (cap + 1) * sizeof(uint32_t)));
        PSS_CUDA_TRY(cudaMalloc(&d_scan_part_, (size_t)(div_up(cap, PS_TILE) + 1) * sizeof(Tri)));
        pair_cap_ = cap;
    }
    if (nq > query_cap_) {
        cudaFree(d_query_off_);
        d_query_off_ = nullptr;
        query_cap_ = 0;
        int64_t cap = std::max<int64_t>(nq + nq / 4, 1024);
        PSS_CUDA_TRY(cudaMalloc(&d_query_off_, (cap + 1) * sizeof(int64_t)));
        query_cap_ = cap;
    }
    return PSS_OK;
}

int Searcher::ensure_hits(int64_t nhits) {
    if (nhits <= hit_cap_) return PSS_OK;
    cudaFree(d_start_); cudaFree(d_end_); cudaFree(d_flag_); cudaFree(d_tile_sum_);
    d_start_ = d_end_ = d_flag_ = d_tile_sum_ = nullptr;
    hit_cap_ = 0;
    int64_t cap = std::max<int64_t>(nhits + nhits / 4, 1 << 16);
    if (cap >= (1ll << 30)) cap = (1ll << 30) - 1;
    PSS_CUDA_TRY(cudaMalloc(&d_start_, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_end_, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_flag_, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_tile_sum_, (size_t)div_up(cap, CP_TILE) * sizeof(uint32_t)));
    hit_cap_ = cap;
    return PSS_OK;
}

// Sort buffers for the hits of heavy pairs (more than LIGHT_MAX matching suffixes).
int Searcher::ensure_heavy(int64_t n) {
    if (n <= heavy_cap_) return PSS_OK;
    cudaFree(d_keys_); cudaFree(d_keys_alt_); cudaFree(d_vals_); cudaFree(d_vals_alt_);
    d_keys_ = d_keys_alt_ = nullptr;
    d_vals_ = d_vals_alt_ = nullptr;
    heavy_cap_ = 0;
    int64_t cap = std::max<int64_t>(n + n / 4, 1 << 16);
    if (cap >= (1ll << 30)) cap = (1ll << 30) - 1;
    PSS_CUDA_TRY(cudaMalloc(&d_keys_, cap * sizeof(uint64_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_keys_alt_, cap * sizeof(uint64_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_vals_, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_vals_alt_, cap * sizeof(uint32_t)));
    PSS_TRY(sorter_.ensure(cap));
    heavy_cap_ = cap;
    return PSS_OK;
}

// Output arrays for `entries` tuples; what is already there is kept (sub-batches append).
// Room for `entries` result tuples; the first `keep` (the entries earlier sub-batches of this
// batch have produced) survive a regrowth, nothing else is copied.
int Searcher::ensure_out(int64_t entries, int64_t keep, cudaStream_t s) {
    if (entries <= out_cap_) return PSS_OK;
    int64_t cap = std::max<int64_t>(entries + entries / 4, 1 << 16);
    int32_t  *nc = nullptr;
    uint32_t *ns = nullptr, *ne = nullptr;
    PSS_CUDA_TRY(cudaMalloc(&nc, cap * sizeof(int32_t)));
    PSS_CUDA_TRY(cudaMalloc(&ns, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&ne, cap * sizeof(uint32_t)));
    keep = std::min(keep, out_cap_);
    if (keep > 0) {
        PSS_CUDA_TRY(cudaMemcpyAsync(nc, d_out_chunk_, keep * 4, cudaMemcpyDeviceToDevice, s));
        PSS_CUDA_TRY(cudaMemcpyAsync(ns, d_out_start_, keep * 4, cudaMemcpyDeviceToDevice, s));
        PSS_CUDA_TRY(cudaMemcpyAsync(ne, d_out_end_, keep * 4, cudaMemcpyDeviceToDevice, s));
    }
    if (out_cap_) PSS_CUDA_TRY(cudaStreamSynchronize(s));   // nothing queued may still use the old arrays
    cudaFree(d_out_chunk_); cudaFree(d_out_start_); cudaFree(d_out_end_);
    d_out_chunk_ = nc; d_out_start_ = ns; d_out_end_ = ne;
    out_cap_ = cap;
    return PSS_OK;
}

int Searcher::search_small(const uint8_t *h_patterns, const int64_t *h_offsets, int32_t nq, SearchOutput *out,
                           SearchTimes *times, bool *handled) {
    *handled = false;
    if (device_ < 0) return fail(PSS_ERR_ARG, "searcher not initialised");
    const int nc = (int)chunks_.size();
    const int64_t npairs = (int64_t)nq * nc;
    if (!small_path_ || nq <= 0 || nq > SMALL_MAX_QUERIES || npairs <= 0 || npairs > SMALL_MAX_PAIRS) return PSS_OK;
    if (h_offsets[nq] > SMALL_PAT_BYTES) return PSS_OK;
    SmallPatterns pats;
    pats.nq = (uint32_t)nq;
    for (int32_t q = 0; q <= nq; ++q) pats.off[q] = (uint32_t)h_offsets[q];
    if (h_offsets[nq]) std::memcpy(pats.bytes, h_patterns, (size_t)h_offsets[nq]);
    PSS_CUDA_TRY(cudaSetDevice(device_));
    const uint32_t seq = ++small_seq_ ? small_seq_ : ++small_seq_;   // never 0
    unsigned char *d_out = nullptr;
    PSS_CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void **>(&d_out), h_small_out_, 0));
    const auto t0 = std::chrono::steady_clock::now();
    small_search_kernel<<<(unsigned)npairs, SMALL_THREADS, sizeof(SmallSmem), stream_>>>(
        d_chunks_, nc, pats, (uint32_t)npairs, seq, d_lb_, d_cnt_, d_scalar_ + 3, d_out);
    PSS_LAUNCH_CHECK();
    // The kernel writes its result into mapped pinned memory and the sequence number last:
    // poll it instead of paying a stream synchronisation.  cudaStreamQuery now and then
    // catches a failed launch (which would never write the word).
    volatile SmallHeader *hdr = reinterpret_cast<volatile SmallHeader *>(h_small_out_);
    uint32_t spins = 0;
    while (hdr->seq != seq) {
        if ((++spins & 0xFFFu) == 0) {
            cudaError_t q = cudaStreamQuery(stream_);
            if (q != cudaErrorNotReady) {
                if (q != cudaSuccess) return fail(PSS_ERR_CUDA, std::string("small search kernel: ") + cudaGetErrorString(q));
                if (hdr->seq != seq) PSS_CUDA_TRY(cudaStreamSynchronize(stream_));
                if (hdr->seq != seq) return fail(PSS_ERR_CUDA, "small search kernel finished without publishing a result");
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (times) {
        *times = SearchTimes();
        times->ms_bounds = times->ms_total =
            std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    if (hdr->status != 0) return PSS_OK;    // too many matching suffixes: not handled (lb/cnt are recomputed)
    const SmallHeader *h = reinterpret_cast<const SmallHeader *>(h_small_out_);
    out->on_host     = true;
    out->n_entries   = h->n_entries;
    out->n_hits      = h->n_hits;
    out->d_entry_off = h->entry_off;
    out->d_query_off = reinterpret_cast<const int64_t *>(h->query_off);
    out->d_chunk     = reinterpret_cast<const int32_t *>(h_small_out_ + sizeof(SmallHeader));
    out->d_start     = reinterpret_cast<const uint32_t *>(out->d_chunk + SMALL_CAP);
    out->d_end       = out->d_start + SMALL_CAP;
    *handled = true;
    return PSS_OK;
}

int Searcher::search(const uint8_t *d_patterns, const int64_t *d_offsets, int32_t nq, cudaStream_t stream,
                     SearchOutput *out, SearchTimes *times, bool defer_compact) {
    deferred_.valid = false;
    if (device_ < 0) return fail(PSS_ERR_ARG, "searcher not initialised");
    if (nq < 0 || !out) return fail(PSS_ERR_ARG, "bad search arguments");
    *out = SearchOutput();
    if (times) *times = SearchTimes();
    const int nc = (int)chunks_.size();
    const int64_t npairs64 = (int64_t)nq * nc;
    if (npairs64 >= (1ll << 31) - 1) return fail(PSS_ERR_ARG, "too many (query, chunk) pairs in one batch");
    PSS_CUDA_TRY(cudaSetDevice(device_));
    cudaStream_t s = stream ? stream : stream_;
    const uint32_t npairs = (uint32_t)npairs64;
    PSS_TRY(ensure_pairs(std::max<int64_t>(npairs, 1), std::max<int64_t>(nq, 1)));
    out->d_entry_off = d_entry_off_;
    out->d_query_off = d_query_off_;
    if (npairs == 0) {
        PSS_CUDA_TRY(cudaMemsetAsync(d_entry_off_, 0, sizeof(uint32_t), s));
        PSS_CUDA_TRY(cudaMemsetAsync(d_query_off_, 0, ((size_t)nq + 1) * sizeof(int64_t), s));
        PSS_CUDA_TRY(cudaStreamSynchronize(s));
        return PSS_OK;
    }

    uint32_t max_n = 1;
    for (const auto &c : chunks_) max_n = std::max(max_n, c.n);
    const int sbits = std::max(1, bit_width_u64((uint64_t)max_n - 1));

    // ---- bounds + hit offsets, one host read: the number of matching suffixes ----------------
    PSS_CUDA_TRY(cudaEventRecord(ev_[0], s));
    // Lanes per (query, chunk) pair, measured on a 2^29-byte chunk (profiles/r02_search_bounds_ab.txt):
    // up to about two warps' worth of pairs per warp slot the kernel is one dependent chain per
    // pair — a whole warp per pair with the SA look-ahead is fastest (10 k pairs: 0.071 ms vs
    // 0.09-0.12 ms grouped); far above that it is bound by random-sector throughput, where the
    // look-ahead's extra sector per level costs more than it hides and more pairs per warp win
    // (150 k pairs: 0.715 ms warp-per-pair without look-ahead, 0.504 with, 0.375 / 0.347 ms with
    // 8 / 4 lanes per pair and no look-ahead).
    int group = bounds_group_;
    if (group == 0) group = npairs < 16384u ? 32 : (npairs < 65536u ? -8 : -4);
    if (group == 8)
        bounds_group_kernel<8, true><<<(unsigned)div_up((int64_t)npairs * 8, BD_THREADS), BD_THREADS, 0, s>>>(
            d_chunks_, nc, d_patterns, d_offsets, npairs, d_lb_, d_cnt_);
    else if (group == -8)
        bounds_group_kernel<8, false><<<(unsigned)div_up((int64_t)npairs * 8, BD_THREADS), BD_THREADS, 0, s>>>(
            d_chunks_, nc, d_patterns, d_offsets, npairs, d_lb_, d_cnt_);
    else if (group == 4)
        bounds_group_kernel<4, true><<<(unsigned)div_up((int64_t)npairs * 4, BD_THREADS), BD_THREADS, 0, s>>>(
            d_chunks_, nc, d_patterns, d_offsets, npairs, d_lb_, d_cnt_);
    else if (group == -4)
        bounds_group_kernel<4, false><<<(unsigned)div_up((int64_t)npairs * 4, BD_THREADS), BD_THREADS, 0, s>>>(
            d_chunks_, nc, d_patterns, d_offsets, npairs, d_lb_, d_cnt_);
    else if (group == -32)
        bounds_group_kernel<32, false><<<(unsigned)div_up((int64_t)npairs * 32, BD_THREADS), BD_THREADS, 0, s>>>(
            d_chunks_, nc, d_patterns, d_offsets, npairs, d_lb_, d_cnt_);
    else
        bounds_kernel<<<(unsigned)div_up((int64_t)npairs * 32, BD_THREADS), BD_THREADS, 0, s>>>(
            d_chunks_, nc, d_patterns, d_offsets, npairs, d_lb_, d_cnt_);
    PSS_LAUNCH_CHECK();
    PSS_CUDA_TRY(cudaEventRecord(ev_[1], s));
    const unsigned ps_tiles = (unsigned)div_up(npairs, PS_TILE);
    hit_partials_kernel<<<ps_tiles, PS_THREADS, 0, s>>>(d_cnt_, npairs, reinterpret_cast<Tri *>(d_scan_part_));
    PSS_LAUNCH_CHECK();
    hit_offsets_kernel<<<ps_tiles, PS_THREADS, 0, s>>>(d_cnt_, npairs, reinterpret_cast<const Tri *>(d_scan_part_), d_hit_off_,
                                                      d_heavy_off_, d_med_list_,
                                                      reinterpret_cast<unsigned long long *>(d_scalar_ + 10));
    PSS_LAUNCH_CHECK();
    PSS_CUDA_TRY(cudaMemcpyAsync(h_scalar_ + 10, d_scalar_ + 10, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    PSS_CUDA_TRY(cudaStreamSynchronize(s));
    const uint64_t total_hits  = (uint64_t)h_scalar_[10] | ((uint64_t)h_scalar_[11] << 32);
    const uint64_t total_heavy = (uint64_t)h_scalar_[12] | ((uint64_t)h_scalar_[13] << 32);
    const uint32_t total_medium = h_scalar_[14];
    out->n_hits = (int64_t)total_hits;

    // ---- sub-batches of pairs whose hits fit the workspace (normally: one) --------------------
    constexpr int64_t HIT_BUDGET = 1ll << 27;
    struct Sub { uint32_t a, np; uint32_t nh; uint32_t n_heavy, n_medium; bool device_offsets; };
    std::vector<Sub> subs;
    if (total_hits <= (uint64_t)HIT_BUDGET) {
        if (total_hits) subs.push_back({0u, npairs, (uint32_t)total_hits, (uint32_t)total_heavy, total_medium, true});
    } else {
        // oversized batch: the per-pair counts come to the host once and are cut there
        if (!h_cnt_) {
            PSS_CUDA_TRY(cudaMallocHost(&h_cnt_, (size_t)pair_cap_ * sizeof(uint32_t)));
            PSS_CUDA_TRY(cudaMallocHost(&h_hit_off_, ((size_t)pair_cap_ + 1) * sizeof(uint32_t)));
            PSS_CUDA_TRY(cudaMallocHost(&h_heavy_off_, ((size_t)pair_cap_ + 1) * sizeof(uint32_t)));
            PSS_CUDA_TRY(cudaMallocHost(&h_med_list_, ((size_t)pair_cap_ + 1) * sizeof(uint32_t)));
        }
        PSS_CUDA_TRY(cudaMemcpyAsync(h_cnt_, d_cnt_, (size_t)npairs * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        PSS_CUDA_TRY(cudaStreamSynchronize(s));
        uint32_t a = 0;
        while (a < npairs) {
            int64_t H = 0;
            uint32_t b = a;
            while (b < npairs && (H == 0 || H + h_cnt_[b] <= HIT_BUDGET)) H += h_cnt_[b++];
            if (H >= (1ll << 30)) return fail(PSS_ERR_ARG, "a single (query, chunk) pair has >= 2^30 hits");
            if (H) subs.push_back({a, b - a, (uint32_t)H, 0u, 0u, false});
            a = b;
        }
    }

    const bool defer = defer_compact && subs.size() == 1 && subs[0].device_offsets;
    uint32_t entries = 0;          // entries produced so far (all sub-batches)
    uint32_t next_pair = 0;        // entry_off is filled up to here
    float ms_extract = 0.f, ms_dedup = 0.f;
    for (const Sub &sb : subs) {
        PSS_TRY(ensure_hits(sb.nh));
        if ((int64_t)entries + sb.nh >= (1ll << 32)) return fail(PSS_ERR_ARG, "more than 2^32 entries in one batch");
        if (!defer) PSS_TRY(ensure_out((int64_t)entries + sb.nh, entries, s));   // entries <= matching suffixes: no count needed up front
        const uint32_t *hit_off = d_hit_off_ + sb.a, *heavy_off = d_heavy_off_ + sb.a;
        const uint32_t hit_base = 0;                         // offsets are relative to the sub-batch
        uint32_t n_heavy = sb.n_heavy, n_medium = sb.n_medium;
        const uint32_t *med_list = d_med_list_ + sb.a;
        if (sb.device_offsets) {
            med_list = d_med_list_;
        } else {
            uint32_t run = 0, hrun = 0, mrun = 0;
            for (uint32_t p = 0; p < sb.np; ++p) {
                const uint32_t c = h_cnt_[sb.a + p];
                h_hit_off_[p] = run;
                h_heavy_off_[p] = hrun;
                run += c;
                if (c > MEDIUM_MAX) hrun += c;
                else if (c > LIGHT_MAX) h_med_list_[mrun++] = p;
            }
            h_hit_off_[sb.np] = run;
            n_heavy = hrun;
            n_medium = mrun;
            PSS_CUDA_TRY(cudaMemcpyAsync(d_hit_off_ + sb.a, h_hit_off_, ((size_t)sb.np + 1) * sizeof(uint32_t),
                                         cudaMemcpyHostToDevice, s));
            PSS_CUDA_TRY(cudaMemcpyAsync(d_heavy_off_ + sb.a, h_heavy_off_, (size_t)sb.np * sizeof(uint32_t),
                                         cudaMemcpyHostToDevice, s));
            if (mrun)
                PSS_CUDA_TRY(cudaMemcpyAsync(d_med_list_ + sb.a, h_med_list_, (size_t)mrun * sizeof(uint32_t),
                                             cudaMemcpyHostToDevice, s));
        }
        PSS_TRY(ensure_heavy(n_heavy));
        // pairs skipped between sub-batches (no hits) get their entry offset here
        if (sb.a > next_pair) {
            std::vector<uint32_t> fill(sb.a - next_pair, entries);
            PSS_CUDA_TRY(cudaMemcpyAsync(d_entry_off_ + next_pair, fill.data(), fill.size() * sizeof(uint32_t),
                                         cudaMemcpyHostToDevice, s));
            PSS_CUDA_TRY(cudaStreamSynchronize(s));
        }
        const uint32_t nh = sb.nh;
        const uint32_t tiles = (uint32_t)div_up(nh, CP_TILE);
        PSS_CUDA_TRY(cudaMemsetAsync(d_flag_, 0, (size_t)nh * sizeof(uint32_t), s));
        PSS_CUDA_TRY(cudaEventRecord(ev_[2], s));
        extract_kernel<<<(unsigned)div_up(nh, 256), 256, 0, s>>>(d_chunks_, nc, sb.a, sb.np, hit_off, heavy_off, d_lb_, d_cnt_,
                                                                nh, sbits, d_start_, d_end_, d_keys_, d_vals_);
        PSS_LAUNCH_CHECK();
        PSS_CUDA_TRY(cudaEventRecord(ev_[3], s));

        // dedup: light pairs by one warp each; heavy pairs through the stable sort of
        // (pair, entry start) with the hit index as value
        pair_dedup_kernel<<<(unsigned)div_up(sb.np, PD_WARPS), PD_WARPS * 32, 0, s>>>(hit_off, d_cnt_, sb.a, sb.np, d_start_, d_flag_);
        PSS_LAUNCH_CHECK();
        if (n_medium) {
            PSS_CUDA_TRY(cudaMemsetAsync(d_scalar_ + 5, 0, 2 * sizeof(uint32_t), s));   // big-pair count, cursor
            pair_dedup_hash_kernel<<<n_medium, PH_THREADS, 4 * PH_SMALL_MAX * sizeof(uint32_t), s>>>(
                med_list, hit_off, d_cnt_, sb.a, d_start_, d_flag_, d_big_list_, d_scalar_ + 5);
            PSS_LAUNCH_CHECK();
            pair_dedup_hash_big_kernel<<<(unsigned)std::min<uint32_t>(n_medium, (uint32_t)sorter_.num_sms()), PH_THREADS,
                                         4 * MEDIUM_MAX * sizeof(uint32_t), s>>>(d_big_list_, d_scalar_ + 5, d_scalar_ + 6, hit_off,
                                                                                 d_cnt_, sb.a, d_start_, d_flag_);
            PSS_LAUNCH_CHECK();
        }
        if (n_heavy) {
            bool in_alt = false;
            const int end_bit = sbits + std::max(1, bit_width_u64((uint64_t)sb.np - 1));
            PSS_TRY(sorter_.sort_async(d_keys_, d_keys_alt_, d_vals_, d_vals_alt_, n_heavy, 0, end_bit, /*iota=*/false, s, &in_alt));
            const uint64_t *k_sorted = in_alt ? d_keys_alt_ : d_keys_;
            const uint32_t *v_sorted = in_alt ? d_vals_alt_ : d_vals_;
            mark_kernel<<<(unsigned)div_up(n_heavy, 256), 256, 0, s>>>(k_sorted, v_sorted, n_heavy, sbits, d_flag_);
            PSS_LAUNCH_CHECK();
        }
        flag_reduce_kernel<<<tiles, CP_THREADS, 0, s>>>(d_flag_, nh, d_tile_sum_);
        PSS_LAUNCH_CHECK();
        tile_scan_kernel<<<1, SCAN_THREADS, 0, s>>>(d_tile_sum_, tiles, d_scalar_ + 2);
        PSS_LAUNCH_CHECK();
        PSS_CUDA_TRY(cudaMemsetAsync(d_pair_first_, 0xFF, (size_t)sb.np * sizeof(uint32_t), s));
        if (defer) {
            pair_first_kernel<<<tiles, CP_THREADS, 0, s>>>(d_flag_, d_tile_sum_, hit_off, sb.np, nh, d_pair_first_);
            deferred_.valid = true;
            deferred_.npairs = sb.np; deferred_.nhits = nh; deferred_.tiles = tiles; deferred_.nc = nc;
        } else {
            compact_kernel<<<tiles, CP_THREADS, 0, s>>>(d_flag_, d_end_, d_tile_sum_, hit_off, hit_base, d_chunks_, nc, sb.a,
                                                        sb.np, nh, d_pair_first_, d_out_chunk_ + entries,
                                                        d_out_start_ + entries, d_out_end_ + entries);
        }
        PSS_LAUNCH_CHECK();
        {
            const unsigned pt = (unsigned)div_up(sb.np, PS_TILE);
            uint32_t *part_min = reinterpret_cast<uint32_t *>(d_scan_part_);
            pair_first_min_kernel<<<pt, PS_THREADS, 0, s>>>(d_pair_first_, sb.np, part_min);
            PSS_LAUNCH_CHECK();
            entry_offsets_kernel<<<pt, PS_THREADS, 0, s>>>(d_pair_first_, sb.np, part_min, d_scalar_ + 2, entries,
                                                           d_entry_off_ + sb.a);
            PSS_LAUNCH_CHECK();
        }
        PSS_CUDA_TRY(cudaEventRecord(ev_[4], s));
        PSS_CUDA_TRY(cudaMemcpyAsync(h_scalar_ + 2, d_scalar_ + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        PSS_CUDA_TRY(cudaMemcpyAsync(h_scalar_ + 8, sorter_.d_error_flag(), sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        PSS_CUDA_TRY(cudaStreamSynchronize(s));
        if (h_scalar_[8]) return sorter_.poll_error(s);
        entries += h_scalar_[2];
        next_pair = sb.a + sb.np;
        if (times) {
            float ms = 0.f;
            PSS_CUDA_TRY(cudaEventElapsedTime(&ms, ev_[2], ev_[3]));
            ms_extract += ms;
            PSS_CUDA_TRY(cudaEventElapsedTime(&ms, ev_[3], ev_[4]));
            ms_dedup += ms;
        }
    }
    // pairs after the last sub-batch with hits (or all pairs when nothing matched)
    if (next_pair <= npairs) {
        std::vector<uint32_t> fill((size_t)npairs - next_pair + 1, entries);
        if (next_pair == 0 || next_pair < npairs) {
            // entry_offsets_kernel already wrote entry_off[next_pair] = entries when a sub-batch ran
            PSS_CUDA_TRY(cudaMemcpyAsync(d_entry_off_ + next_pair, fill.data(), fill.size() * sizeof(uint32_t),
                                         cudaMemcpyHostToDevice, s));
            PSS_CUDA_TRY(cudaStreamSynchronize(s));   // `fill` is pageable: complete before it goes out of scope
        }
    }
    query_offsets_kernel<<<(unsigned)div_up((int64_t)nq + 1, 256), 256, 0, s>>>(d_entry_off_, (uint32_t)nq, (uint32_t)nc,
                                                                               d_query_off_);
    PSS_LAUNCH_CHECK();
    out->n_entries = entries;
    out->deferred  = defer;
    out->d_chunk   = defer ? nullptr : d_out_chunk_;
    out->d_start   = defer ? nullptr : d_out_start_;
    out->d_end     = defer ? nullptr : d_out_end_;
    if (times) {
        float ms = 0.f;
        PSS_CUDA_TRY(cudaEventElapsedTime(&ms, ev_[0], ev_[1]));
        times->ms_bounds  = ms;
        times->ms_extract = ms_extract;
        times->ms_dedup   = ms_dedup;
    }
    return PSS_OK;
}

int Searcher::compact_deferred(const uint32_t *d_pair_dst, int32_t *d_chunk, uint32_t *d_start, uint32_t *d_end,
                               cudaStream_t stream) {
    if (!deferred_.valid) return PSS_OK;      // nothing matched (or nothing was deferred): no tuples to write
    if (!d_pair_dst || !d_start || !d_end) return fail(PSS_ERR_ARG, "compact_deferred: null destination");
    cudaStream_t s = stream ? stream : stream_;
    compact_place_kernel<<<deferred_.tiles, CP_THREADS, 0, s>>>(d_flag_, d_end_, d_tile_sum_, d_hit_off_, d_entry_off_,
                                                               d_pair_dst, d_chunks_, deferred_.nc, deferred_.npairs,
                                                               deferred_.nhits, d_chunk, d_start, d_end);
    PSS_LAUNCH_CHECK();
    deferred_.valid = false;
    return PSS_OK;
}

}  // namespace pss
