// sa_build.cu — prefix-doubling suffix-array builder for sm_100a (see sa_build.cuh).
#include "sa_build.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

namespace pss {

namespace {

// d_small_ layout (32-bit words)
constexpr int SM_PRESENCE = 0;    // [0..8)   256-bit set of byte values present in the text
constexpr int SM_SCALARS  = 8;    // [8..16)  [8] = active suffixes after the current re-rank, [9] = tile ticket,
                                  //          [10] = look-back watchdog flag, [11] = active-set append cursor,
                                  //          [12] = the sorter's watchdog flag (copied in for the read-back)
constexpr int SM_LUT      = 16;   // [16..144) 256 x u16: byte value → dense code (1..sigma)
constexpr int SM_WORDS    = 144;

// ------------------------------------------------------------------------------------
// Alphabet: which byte values occur (a 256-bit presence set is all the packing needs).
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
presence_kernel(const uint8_t *__restrict__ text, uint32_t n, uint32_t *__restrict__ presence) {
    __shared__ uint32_t s_seen[256];
    s_seen[threadIdx.x] = 0;
    __syncthreads();
    // 16-byte vector body between the first and last 16-byte boundary; ragged ends by
    // one thread.
    const uintptr_t addr = reinterpret_cast<uintptr_t>(text);
    uint32_t head = (uint32_t)((16 - (addr & 15)) & 15);
    if (head > n) head = n;
    const uint32_t nvec = (n - head) / 16;
    const uint4 *vec    = reinterpret_cast<const uint4 *>(text + head);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
        uint4 v = ld_stream_u128(vec + i);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            s_seen[w[q] & 0xFF]         = 1;
            s_seen[(w[q] >> 8) & 0xFF]  = 1;
            s_seen[(w[q] >> 16) & 0xFF] = 1;
            s_seen[w[q] >> 24]          = 1;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (uint32_t i = 0; i < head; ++i) s_seen[text[i]] = 1;
        for (uint32_t i = head + nvec * 16; i < n; ++i) s_seen[text[i]] = 1;
    }
    __syncthreads();
    if (s_seen[threadIdx.x]) atomicOr(&presence[threadIdx.x >> 5], 1u << (threadIdx.x & 31));
}

// ------------------------------------------------------------------------------------
// Round 0 keys: key(i) = codes of T[i .. i+m) packed big-endian, b bits each, 0 past end.
// ------------------------------------------------------------------------------------
constexpr int KG_THREADS = 256;
constexpr int KG_IPT     = 8;
constexpr int KG_TILE    = KG_THREADS * KG_IPT;

// Persistent: each CTA walks tiles grid-stride and also accumulates the radix histograms of
// the keys it produces (so the sort needs no separate histogram pass over them).
__global__ void __launch_bounds__(KG_THREADS)
keygen_kernel(const uint8_t *__restrict__ text, uint32_t n, const uint16_t *__restrict__ lut, int b, int m,
              uint64_t *__restrict__ keys, HistLayout hl, uint32_t *__restrict__ g_hist) {
    __shared__ uint16_t s_lut[256];
    __shared__ uint16_t s_code[KG_TILE + 64];
    __shared__ uint64_t s_key[KG_TILE + KG_TILE / KG_IPT];  // one pad word per thread row
    __shared__ uint32_t s_hist[MAX_PASSES * RADIX];
    const uint32_t tid = threadIdx.x;
    s_lut[tid] = lut[tid];
    for (int i = tid; i < hl.npass * RADIX; i += KG_THREADS) s_hist[i] = 0;
    __syncthreads();
    const uint32_t tiles = (n + KG_TILE - 1) / KG_TILE;
    for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const uint32_t base = t * KG_TILE;
    for (uint32_t i = tid; i < (uint32_t)(KG_TILE + m); i += KG_THREADS) {
        uint32_t pos = base + i;
        s_code[i] = pos < n ? s_lut[text[pos]] : (uint16_t)0;
    }
    __syncthreads();
    const uint64_t mask = (m * b >= 64) ? ~0ull : ((1ull << (m * b)) - 1ull);
    const uint32_t r0 = tid * KG_IPT;
    uint64_t w = 0;
    for (int j = 0; j < m; ++j) w = (w << b) | s_code[r0 + j];
#pragma unroll
    for (int j = 0; j < KG_IPT; ++j) {
        s_key[tid * (KG_IPT + 1) + j] = w;
        w = ((w << b) & mask) | s_code[r0 + j + m];
    }
    __syncthreads();
    const bool tile_full = base + KG_TILE <= n;
#pragma unroll
    for (int j = 0; j < KG_IPT; ++j) {
        uint32_t i = j * KG_THREADS + tid;
        const bool ok = base + i < n;
        const uint64_t kv = s_key[(i / KG_IPT) * (KG_IPT + 1) + (i % KG_IPT)];
        if (ok) keys[base + i] = kv;
        hist_accumulate(s_hist, kv, ok, tile_full, hl);
    }
    __syncthreads();   // s_code / s_key are reused by the next tile
    }
    hist_flush(s_hist, g_hist, hl.npass, KG_THREADS);
}

// ------------------------------------------------------------------------------------
// Round r >= 1 keys: (group rank << rbits) | rank of suffix i + h (0 past the end).
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_kernel(const uint32_t *__restrict__ idx, const uint32_t *__restrict__ grp,
              const uint32_t *__restrict__ isa, uint32_t n, uint32_t h, int rbits, uint32_t n_active,
              uint64_t *__restrict__ keys, HistLayout hl, uint32_t *__restrict__ g_hist, int l2_hints) {
    constexpr int U = 4;
    __shared__ uint32_t s_hist[MAX_PASSES * RADIX];
    for (int i = threadIdx.x; i < hl.npass * RADIX; i += 256) s_hist[i] = 0;
    __syncthreads();
    const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first();
    const uint32_t blocks = (n_active + 256 * U - 1) / (256 * U);
    for (uint32_t blk = blockIdx.x; blk < blocks; blk += gridDim.x) {
        const uint32_t base = (blk * 256 * U) + threadIdx.x;
        uint32_t i[U], g[U], r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t k = base + u * 256;
            i[u] = k < n_active ? (l2_hints ? ld_stream_u32_hint(idx + k, stream) : ld_stream_u32(idx + k)) : 0u;
            g[u] = k < n_active ? (l2_hints ? ld_stream_u32_hint(grp + k, stream) : ld_stream_u32(grp + k)) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t k = base + u * 256;
            uint64_t j = (uint64_t)i[u] + h;
            // the active set is bucketed by index window: these reads stay inside a slice of ISA
            // that fits in L2, provided L2 keeps it (evict_last) while the records stream by
            r[u] = (k < n_active && j < n) ? (l2_hints ? ld_nc_u32_hint(isa + j, keep) : __ldg(isa + j)) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t k = base + u * 256;
            const uint64_t kv = ((uint64_t)g[u] << rbits) | r[u];
            if (k < n_active) {
                if (l2_hints) st_u64_hint(keys + k, kv, stream);
                else keys[k] = kv;
            }
            // the warp's 32 keys are consecutive: full iff its last one is in range
            hist_accumulate(s_hist, kv, k < n_active, (k | 31u) < n_active, hl);
        }
    }
    __syncthreads();
    hist_flush(s_hist, g_hist, hl.npass, 256);
}

// ------------------------------------------------------------------------------------
// Rank scatter ISA[index] = rank from (index << 32 | rank) pairs that were partitioned by
// the top bits of the index: concurrently running CTAs then write into one window of ISA
// that fits in L2, so every 32-byte sector reaches DRAM once, fully written, instead of
// being read-modified-written once per 4-byte rank.
// ------------------------------------------------------------------------------------
// The same pass also extracts the next round's active set — pairs whose bit 31 is set are
// suffixes that still share a group; their (index, group rank = rank - 1) are appended to
// out_idx / out_grp.  Taking them from here rather than from the sorted order leaves the
// active set bucketed by index window as well, so the next round's ISA[i + h] gather reads
// stay inside an L2-resident window too.  (The order of the active set is irrelevant: the
// next sort is a full radix sort.)  Space is claimed with one atomicAdd per CTA.
__global__ void __launch_bounds__(256)
isa_scatter_kernel(const uint64_t *__restrict__ pairs, uint32_t n_pairs, uint32_t *__restrict__ isa,
                   uint32_t *__restrict__ out_idx, uint32_t *__restrict__ out_grp, uint32_t *__restrict__ counter) {
    constexpr int U = 4;
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base;
    __shared__ uint32_t s_idx[256 * U], s_grp[256 * U];   // the CTA's kept records, staged for coalesced appends
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    // blocked: a thread owns U consecutive pairs, so that its kept records are written together
    const uint32_t base = (blockIdx.x * 256 + threadIdx.x) * U;
    uint64_t v[U];
    uint32_t kept = 0;
    if (base + U <= n_pairs) {   // 4 pairs = 32 aligned bytes: two 16-byte loads
        const uint4 *vp = reinterpret_cast<const uint4 *>(pairs + base);
        const uint4 a = ld_stream_u128(vp), b = ld_stream_u128(vp + 1);
        v[0] = ((uint64_t)a.y << 32) | a.x; v[1] = ((uint64_t)a.w << 32) | a.z;
        v[2] = ((uint64_t)b.y << 32) | b.x; v[3] = ((uint64_t)b.w << 32) | b.z;
    } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t k = base + u;
            v[u] = k < n_pairs ? ld_stream_u64(pairs + k) : 0ull;
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) kept += (uint32_t)(v[u] >> 31) & 1u;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint32_t k = base + u;
        if (k < n_pairs) isa[(uint32_t)(v[u] >> 32)] = (uint32_t)v[u] & 0x7FFFFFFFu;
    }
    uint32_t incl = kept;
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
        const uint32_t t = s_warp[w];
        if ((uint32_t)w < warp) pre += t;
        tot += t;
    }
    if (threadIdx.x == 0) s_base = tot ? atomicAdd(counter, tot) : 0u;
    uint32_t o = pre + incl - kept;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if ((v[u] >> 31) & 1u) {
            s_idx[o] = (uint32_t)(v[u] >> 32);
            s_grp[o] = ((uint32_t)v[u] & 0x7FFFFFFFu) - 1u;
            ++o;
        }
    }
    __syncthreads();
    // the CTA's run [s_base, s_base + tot) is written by consecutive threads: full sectors
    // instead of one 32-byte sector per 4-byte record
    const uint32_t gb = s_base;
    for (uint32_t t = threadIdx.x; t < tot; t += 256) {
        out_idx[gb + t] = s_idx[t];
        out_grp[gb + t] = s_grp[t];
    }
}

// ------------------------------------------------------------------------------------
// Segmented re-rank after a sort.  For sorted position k (old group rank g = key >> gs):
//   A(k) = last k' <= k that starts an old group,  B(k) = last k' <= k that starts a new
//   group (new group = run of equal full keys).  SA position p = g + (k - A), new group
//   rank = g + (B - A).  A suffix alone in its new group is final: SA[p] = i.  The rest
//   are compacted (order preserved) into the next round's active set.
// One pass over tiles of 2048 records: per-tile aggregates are chained by decoupled look-back.
// ------------------------------------------------------------------------------------
constexpr int RR_THREADS = 256;
constexpr int RR_IPT     = 8;
constexpr int RR_TILE    = RR_THREADS * RR_IPT;

struct Tup {
    uint32_t a, b, s;  // a/b: (position + 1) of the last old/new head seen, 0 = none; s: kept count
};
__device__ __forceinline__ Tup tup_comb(const Tup &x, const Tup &y) {
    Tup r;
    r.a = max(x.a, y.a);
    r.b = max(x.b, y.b);
    r.s = x.s + y.s;
    return r;
}
__device__ __forceinline__ Tup tup_shfl_up(const Tup &x, int o) {
    Tup r;
    r.a = __shfl_up_sync(0xffffffffu, x.a, o);
    r.b = __shfl_up_sync(0xffffffffu, x.b, o);
    r.s = __shfl_up_sync(0xffffffffu, x.s, o);
    return r;
}

// Exclusive scan of one Tup per thread across the block; returns the exclusive prefix and
// writes the block total to *total (valid in every thread).  s_warp: one Tup per warp.
template <int THREADS>
__device__ __forceinline__ Tup block_excl_scan(Tup v, Tup *s_warp, Tup *total) {
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    Tup incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Tup y = tup_shfl_up(incl, o);
        if (lane >= (uint32_t)o) incl = tup_comb(y, incl);
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    Tup pre = {0, 0, 0};
    Tup tot = {0, 0, 0};
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) {
        Tup t = s_warp[w];
        if ((uint32_t)w < warp) pre = tup_comb(pre, t);
        tot = tup_comb(tot, t);
    }
    Tup up = tup_shfl_up(incl, 1);
    if (lane > 0) pre = tup_comb(pre, up);
    *total = tot;
    __syncthreads();
    return pre;
}

// Sorted keys of a tile (+ one halo record each side) in shared memory, padded so that
// a thread's 8 consecutive records do not collide on banks.  Local index u = k - base + 1.
struct RerankTile {
    uint64_t k[RR_TILE + 2 + (RR_TILE + 2) / 8 + 1];
    __device__ __forceinline__ static uint32_t slot(uint32_t u) { return u + (u >> 3); }
};

__device__ __forceinline__ void rerank_load(RerankTile &t, const uint64_t *__restrict__ keys, uint32_t base,
                                            uint32_t n_active) {
    // all of a thread's loads are issued before the first store to shared memory: one DRAM round
    // trip per tile instead of nine dependent ones (ncu: a third of the kernel's stall samples sat
    // on the store that waited for each load in turn)
    constexpr int PER = (RR_TILE + 2 + RR_THREADS - 1) / RR_THREADS;   // 9: the last round holds the 2 halo words
    uint64_t v[PER];
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        const uint32_t u = e * RR_THREADS + threadIdx.x;
        const int64_t k  = (int64_t)base + u - 1;
        v[e] = 0;
        if (u < RR_TILE + 2 && k >= 0 && k < (int64_t)n_active) v[e] = ld_stream_u64(keys + k);
    }
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        const uint32_t u = e * RR_THREADS + threadIdx.x;
        if (u < RR_TILE + 2) t.k[RerankTile::slot(u)] = v[e];
    }
    __syncthreads();
}

// Flags of sorted record k (u = local index + 1): bit0 old head, bit1 new head, bit2 kept.
__device__ __forceinline__ uint32_t rerank_flags(const RerankTile &t, uint32_t u, uint32_t k, uint32_t n_active,
                                                 int gs, bool first) {
    uint64_t prev = t.k[RerankTile::slot(u - 1)];
    uint64_t cur  = t.k[RerankTile::slot(u)];
    uint64_t next = t.k[RerankTile::slot(u + 1)];
    bool hn  = (k == 0) || (cur != prev);
    bool ho  = (k == 0) || (!first && ((cur >> gs) != (prev >> gs)));
    bool nhn = (k + 1 == n_active) || (next != cur);
    bool keep = !(hn && nhn);
    return (ho ? 1u : 0u) | (hn ? 2u : 0u) | (keep ? 4u : 0u);
}

// Look-back state of the single-pass re-rank: two self-validating 64-bit words per tile,
//   w0 = flag(2) | a(31) | b(31),  w1 = flag(2) | s(32);  flag 1 = the tile's own aggregate,
//   2 = inclusive prefix over tiles 0..t, 3 = abort.  A reader accepts a pair only when
//   both flags agree, so no fence is needed between the two stores.
__device__ __forceinline__ void rr_publish(unsigned long long *state, uint32_t tile, uint32_t flag, const Tup &v) {
    const unsigned long long w0 = ((unsigned long long)flag << 62) | ((unsigned long long)v.a << 31) | v.b;
    const unsigned long long w1 = ((unsigned long long)flag << 62) | v.s;
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(state + 2 * (size_t)tile), "l"(w0) : "memory");
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(state + 2 * (size_t)tile + 1), "l"(w1) : "memory");
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(RR_THREADS, 5)
rerank_apply_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t n_active,
                    int gs, int first, unsigned long long *tile_state, uint32_t *scalars,
                    uint32_t *__restrict__ isa, uint64_t *__restrict__ pairs_out, int32_t *__restrict__ sa,
                    uint32_t *__restrict__ out_idx, uint32_t *__restrict__ out_grp) {
    __shared__ RerankTile tile;
    __shared__ Tup s_warp[RR_THREADS / 32];
    __shared__ Tup s_excl;
    __shared__ uint32_t s_tile, s_abort;
    // single pass: tiles are ticketed in launch order and chained by decoupled look-back
    if (threadIdx.x == 0) {
        s_tile  = atomicAdd(&scalars[1], 1u);
        s_abort = 0;
    }
    __syncthreads();
    const uint32_t tile_id = s_tile;
    const uint32_t base    = tile_id * RR_TILE;

    static_assert(RR_IPT == 8, "the vector loads/stores below move 8 records per thread");
    uint32_t flags[RR_IPT];
    uint32_t idx[RR_IPT];
    const bool full_tile = base + RR_TILE <= n_active;
    if (full_tile) {
        // a thread's 8 suffix indices are 32 contiguous, 32-byte aligned bytes: two 16-byte loads
        // instead of eight 4-byte ones that would each pull a whole sector for 4 useful bytes.
        // Issued BEFORE the key loads: both streams share one DRAM round trip.
        const uint4 *vp = reinterpret_cast<const uint4 *>(vals + base + threadIdx.x * RR_IPT);
        const uint4 a = ld_stream_u128(vp), b = ld_stream_u128(vp + 1);
        idx[0] = a.x; idx[1] = a.y; idx[2] = a.z; idx[3] = a.w;
        idx[4] = b.x; idx[5] = b.y; idx[6] = b.z; idx[7] = b.w;
    }
    rerank_load(tile, keys, base, n_active);
    Tup agg = {0, 0, 0};
#pragma unroll
    for (int e = 0; e < RR_IPT; ++e) {
        uint32_t loc = threadIdx.x * RR_IPT + e;
        uint32_t k   = base + loc;
        flags[e] = 0;
        if (!full_tile) idx[e] = 0;
        if (k < n_active) {
            uint32_t f = rerank_flags(tile, loc + 1, k, n_active, gs, first != 0);
            flags[e]   = f | 8u;  // bit3: record exists
            if (!full_tile) idx[e] = ld_stream_u32(vals + k);
            if (f & 1u) agg.a = k + 1;
            if (f & 2u) agg.b = k + 1;
            agg.s += (f >> 2) & 1u;
        }
    }
    Tup total;
    Tup run = block_excl_scan<RR_THREADS>(agg, s_warp, &total);

    // ---- decoupled look-back over the preceding tiles' aggregates (warp 0, 32 tiles a step) ----
    if (threadIdx.x < 32) {
        const uint32_t lane = threadIdx.x;
        if (lane == 0) rr_publish(tile_state, tile_id, tile_id == 0 ? 2u : 1u, total);
        Tup excl = {0, 0, 0};
        bool aborted = false;
        if (tile_id > 0) {
            // Four 32-tile windows of predecessor states are fetched per L2 round trip.  With P tiles
            // in flight, a look-back that covers W tiles per round trip tau sustains W / tau tiles per
            // second once the walk is the critical path (a tile finds its nearest inclusive prefix
            // about tau * P / W behind): 32 per step capped the kernel at ~2 TB/s (ncu: 13.8 us per
            // 2048-record tile, stalls on barrier + long scoreboard).
            constexpr int LBW = 4;
            int64_t basep  = (int64_t)tile_id - 1;
            uint32_t spins = 0;
            uint64_t t0 = 0;
            bool done = false;
            while (!done && !aborted) {
                unsigned long long w0[LBW], w1[LBW];
#pragma unroll
                for (int j = 0; j < LBW; ++j) {
                    const int64_t q = basep - 32 * j - lane;
                    w0[j] = 2ull << 62; w1[j] = 2ull << 62;   // before tile 0: inclusive identity
                    if (q >= 0) {
                        w0[j] = ld_volatile_u64(tile_state + 2 * (size_t)q);
                        w1[j] = ld_volatile_u64(tile_state + 2 * (size_t)q + 1);
                    }
                }
                int consumed = 0;
#pragma unroll
                for (int j = 0; j < LBW; ++j) {
                    if (done || aborted || consumed < j) continue;        // an earlier window has to be polled again
                    const uint32_t f0 = (uint32_t)(w0[j] >> 62), f1 = (uint32_t)(w1[j] >> 62);
                    const bool ready  = f0 != 0 && f0 == f1;
                    const uint32_t m_incl  = __ballot_sync(0xffffffffu, ready && f0 == 2u);
                    const uint32_t m_abort = __ballot_sync(0xffffffffu, f0 == 3u || f1 == 3u);
                    const int k_incl       = m_incl ? __ffs(m_incl) - 1 : 32;
                    const uint32_t need    = k_incl >= 31 ? 0xffffffffu : ((2u << k_incl) - 1u);   // lanes 0..k_incl
                    const uint32_t m_wait  = __ballot_sync(0xffffffffu, !ready) & need;
                    if (m_abort & need) { aborted = true; continue; }
                    if (m_wait) continue;                                  // consumed stays at j: re-poll from here
                    Tup v = {0, 0, 0};
                    if ((int)lane <= k_incl) {
                        v.a = (uint32_t)((w0[j] >> 31) & 0x7FFFFFFFu);
                        v.b = (uint32_t)(w0[j] & 0x7FFFFFFFu);
                        v.s = (uint32_t)w1[j];
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        Tup y;
                        y.a = __shfl_xor_sync(0xffffffffu, v.a, o);
                        y.b = __shfl_xor_sync(0xffffffffu, v.b, o);
                        y.s = __shfl_xor_sync(0xffffffffu, v.s, o);
                        v = tup_comb(v, y);
                    }
                    excl = tup_comb(excl, v);
                    consumed = j + 1;
                    if (k_incl < 32) done = true;
                }
                basep -= 32 * consumed;
                if (!done && !aborted && consumed == 0) {
                    // wall-clock watchdog (warp-uniform: every lane evaluates the same values)
                    if ((++spins & 0x3FFFu) == 0) {
                        const uint64_t now = __shfl_sync(0xffffffffu, global_timer_ns(), 0);
                        if (t0 == 0) t0 = now;
                        else if (now - t0 > 30ull * 1000ull * 1000ull * 1000ull) aborted = true;
                    }
                }
            }
            if (lane == 0) {
                if (aborted) {
                    Tup z = {0, 0, 0};
                    rr_publish(tile_state, tile_id, 3u, z);
                    atomicExch(&scalars[2], 1u);
                    s_abort = 1;
                } else {
                    rr_publish(tile_state, tile_id, 2u, tup_comb(excl, total));
                }
            }
        }
        if (lane == 0) {
            s_excl = excl;
            if (base + RR_TILE >= n_active) scalars[0] = excl.s + total.s;   // last tile: next active count
        }
    }
    __syncthreads();
    if (s_abort) return;
    run = tup_comb(s_excl, run);

    uint64_t pr[RR_IPT];
#pragma unroll
    for (int e = 0; e < RR_IPT; ++e) {
        pr[e] = 0;
        if (!(flags[e] & 8u)) continue;
        uint32_t loc = threadIdx.x * RR_IPT + e;
        uint32_t k   = base + loc;
        uint32_t f   = flags[e];
        if (f & 1u) run.a = k + 1;
        if (f & 2u) run.b = k + 1;
        const uint32_t A = run.a - 1, B = run.b - 1;
        const uint32_t g = first ? 0u : (uint32_t)(tile.k[RerankTile::slot(loc + 1)] >> gs);
        const uint32_t p  = g + (k - A);   // final SA slot of this record inside its old group
        const uint32_t ng = g + (B - A);   // rank of its new group = SA slot of the group's head
        // New rank of the suffix (1-based; 0 is "past the end"): the group's rank, or the
        // suffix's final SA slot if it is alone in its group.
        const uint32_t rank1 = ((f & 4u) ? ng : p) + 1;
        if (!(f & 4u)) sa[p] = (int32_t)idx[e];
        if (pairs_out) {
            // large rounds: the scatter into ISA and the extraction of the next active set
            // happen later, partitioned by index window; bit 31 marks "still active"
            pr[e] = ((uint64_t)idx[e] << 32) | rank1 | ((f & 4u) ? 0x80000000u : 0u);
            if (!full_tile) pairs_out[k] = pr[e];
        } else {
            if (f & 4u) {
                const uint32_t c = run.s++;
                out_idx[c] = idx[e];
                out_grp[c] = ng;
            }
            if (first || rank1 != g + 1) isa[idx[e]] = rank1;
        }
    }
    if (pairs_out && full_tile) {
        ulonglong2 *pp = reinterpret_cast<ulonglong2 *>(pairs_out + base + threadIdx.x * RR_IPT);
#pragma unroll
        for (int j = 0; j < RR_IPT / 2; ++j) pp[j] = make_ulonglong2(pr[2 * j], pr[2 * j + 1]);
    }
}

}  // namespace

// ------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------
int SaBuilder::init(int device, int64_t max_n) {
    if (device_ >= 0) return PSS_OK;
    if (device < 0) device = default_device();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(PSS_ERR_CUDA, "no CUDA device available (libpss_b200 has no CPU fallback)");
    if (device >= ndev) return fail(PSS_ERR_ARG, "device index out of range");
    PSS_CUDA_TRY(cudaSetDevice(device));
    device_ = device;
    // The rank gather / scatter of prefix doubling touches 4 bytes per random sector: ask
    // L2 not to pull in neighbouring sectors on those misses (a hint; ignored if unsupported).
    if (const char *e = std::getenv("PSS_L2_FETCH")) {
        int g = std::atoi(e);
        if (g == 32 || g == 64 || g == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g);
        cudaGetLastError();
    }
    if (const char *e = std::getenv("PSS_L2_HINTS")) l2_hints_ = std::atoi(e);
    if (const char *e = std::getenv("PSS_GATHER_CTAS")) gather_ctas_ = std::max(1, std::atoi(e));
    PSS_CUDA_TRY(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    PSS_CUDA_TRY(cudaEventCreate(&ev_begin_));
    PSS_CUDA_TRY(cudaEventCreate(&ev_end_));
    PSS_CUDA_TRY(cudaMalloc(&d_small_, SM_WORDS * sizeof(uint32_t)));
    PSS_TRY(alloc_mapped_words(&h_small_, SM_WORDS));
    PSS_TRY(sorter_.init(device_));
    if (max_n > 0) PSS_TRY(ensure(max_n));
    return PSS_OK;
}

int SaBuilder::ensure(int64_t n) {
    if (n <= cap_) return PSS_OK;
    PSS_CUDA_TRY(cudaSetDevice(device_));
    cudaFree(keys_a_); cudaFree(keys_b_); cudaFree(vals_a_); cudaFree(vals_b_);
    cudaFree(grp_); cudaFree(isa_); cudaFree(tile_aggr_);
    keys_a_ = keys_b_ = nullptr; vals_a_ = vals_b_ = grp_ = isa_ = tile_aggr_ = nullptr;
    cap_ = 0;
    int64_t cap = std::max<int64_t>(n, 1 << 16);
    PSS_CUDA_TRY(cudaMalloc(&keys_a_, cap * sizeof(uint64_t)));
    PSS_CUDA_TRY(cudaMalloc(&keys_b_, cap * sizeof(uint64_t)));
    PSS_CUDA_TRY(cudaMalloc(&vals_a_, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&vals_b_, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&grp_, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&isa_, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&tile_aggr_, (size_t)div_up(cap, RR_TILE) * 2 * sizeof(unsigned long long)));
    PSS_TRY(sorter_.ensure(cap));
    cap_ = cap;
    return PSS_OK;
}

int SaBuilder::ensure_io(int64_t n) {
    if (n <= io_cap_) return PSS_OK;
    PSS_CUDA_TRY(cudaSetDevice(device_));
    cudaFree(d_text_); cudaFree(d_sa_);
    d_text_ = nullptr; d_sa_ = nullptr; io_cap_ = 0;
    int64_t cap = std::max<int64_t>(n, 1 << 16);
    PSS_CUDA_TRY(cudaMalloc(&d_text_, cap + 64));
    PSS_CUDA_TRY(cudaMalloc(&d_sa_, cap * sizeof(int32_t)));
    io_cap_ = cap;
    return PSS_OK;
}

void SaBuilder::release_workspace() {
    if (device_ < 0) return;
    cudaSetDevice(device_);
    cudaFree(keys_a_); cudaFree(keys_b_); cudaFree(vals_a_); cudaFree(vals_b_);
    cudaFree(grp_); cudaFree(isa_); cudaFree(tile_aggr_);
    cudaFree(d_text_); cudaFree(d_sa_);
    keys_a_ = keys_b_ = nullptr; vals_a_ = vals_b_ = grp_ = isa_ = tile_aggr_ = nullptr;
    d_text_ = nullptr; d_sa_ = nullptr;
    cap_ = io_cap_ = 0;
    stager_.release();
    sorter_.release_workspace();
}

void SaBuilder::release() {
    if (device_ < 0) return;
    cudaSetDevice(device_);
    cudaFree(keys_a_); cudaFree(keys_b_); cudaFree(vals_a_); cudaFree(vals_b_);
    cudaFree(grp_); cudaFree(isa_); cudaFree(tile_aggr_); cudaFree(d_small_);
    cudaFree(d_text_); cudaFree(d_sa_);
    if (h_small_) cudaFreeHost(h_small_);
    stager_.release();
    if (ev_begin_) cudaEventDestroy(ev_begin_);
    if (ev_end_) cudaEventDestroy(ev_end_);
    if (stream_) cudaStreamDestroy(stream_);
    sorter_.release();
    keys_a_ = keys_b_ = nullptr; vals_a_ = vals_b_ = grp_ = isa_ = tile_aggr_ = d_small_ = nullptr;
    h_small_ = nullptr; d_text_ = nullptr; d_sa_ = nullptr; stream_ = nullptr;
    ev_begin_ = ev_end_ = nullptr;
    cap_ = io_cap_ = 0;
    device_ = -1;
}

int SaBuilder::build_device(const uint8_t *d_text, int32_t n, int32_t *d_sa, cudaStream_t stream) {
    if (device_ < 0) return fail(PSS_ERR_ARG, "builder not initialised");
    if (n < 0 || (n > 0 && (!d_text || !d_sa))) return fail(PSS_ERR_ARG, "bad build arguments");
    if ((int64_t)n >= (1ll << 30)) return fail(PSS_ERR_ARG, "n must be < 2^30 (container stores 4n in a u32)");
    std::memset(&stats_, 0, sizeof(stats_));
    pass_stats_.clear();
    stats_.n = n;
    if (n == 0) return PSS_OK;
    PSS_CUDA_TRY(cudaSetDevice(device_));
    cudaStream_t s = stream ? stream : stream_;
    const long long launches0 = g_kernel_launches.load();
    if (n == 1) {
        PSS_CUDA_TRY(cudaMemsetAsync(d_sa, 0, sizeof(int32_t), s));
        PSS_CUDA_TRY(cudaStreamSynchronize(s));
        return PSS_OK;
    }
    PSS_TRY(ensure(n));
    const uint32_t un = (uint32_t)n;

    PSS_CUDA_TRY(cudaEventRecord(ev_begin_, s));

    // ---- alphabet → dense codes -------------------------------------------------------
    PSS_CUDA_TRY(cudaMemsetAsync(d_small_, 0, SM_WORDS * sizeof(uint32_t), s));
    {
        int grid = (int)std::min<int64_t>(div_up(n, 256 * 64), (int64_t)sm_count(device_) * 8);
        presence_kernel<<<std::max(grid, 1), 256, 0, s>>>(d_text, un, d_small_ + SM_PRESENCE);
        PSS_LAUNCH_CHECK();
    }
    PSS_TRY(copy_words(h_small_, d_small_, 8, s));
    PSS_CUDA_TRY(cudaStreamSynchronize(s));
    int sigma = 0;
    uint16_t *lut = reinterpret_cast<uint16_t *>(h_small_ + SM_LUT);
    for (int c = 0; c < 256; ++c) {
        bool present = (h_small_[SM_PRESENCE + (c >> 5)] >> (c & 31)) & 1u;
        lut[c] = present ? (uint16_t)(++sigma) : (uint16_t)0;
    }
    const int b = bit_width_u64((uint64_t)sigma);  // codes 1..sigma, 0 = past the end
    int m = 64 / b;
    if (const char *e = std::getenv("PSS_H0")) {
        int v = std::atoi(e);
        if (v >= 1 && v <= 64 / b) m = v;
    }
    if ((int64_t)m > (int64_t)n) m = n;  // no point packing past the text
    stats_.sigma = sigma;
    stats_.bits_per_symbol = b;
    stats_.h0 = m;
    PSS_TRY(copy_words(d_small_ + SM_LUT, h_small_ + SM_LUT, 128, s));   // device reads the mapped host words

    // ---- round 0: packed-prefix keys, sort, rank ---------------------------------------
    PSS_TRY(sorter_.hist_reset(s));
    {
        const int grid = (int)std::min<int64_t>(div_up(n, KG_TILE), (int64_t)sorter_.num_sms() * 4);
        keygen_kernel<<<grid, KG_THREADS, 0, s>>>(d_text, un, reinterpret_cast<const uint16_t *>(d_small_ + SM_LUT), b, m,
                                                  keys_a_, hist_layout(0, m * b), sorter_.d_hist());
        PSS_LAUNCH_CHECK();
    }

    SortProfile prof;
    prof.timed = profiling_;
    auto record_passes = [&](int round, uint32_t n_rec) {
        stats_.n_passes += prof.n_passes;
        for (int p = 0; p < prof.n_passes; ++p) {
            stats_.records_sorted += n_rec;
            stats_.sort_ms += prof.ms[p];
            if (profiling_ && (int)pass_stats_.size() < PSS_MAX_PASS_STATS) {
                pss_pass_stat ps = {};
                ps.round = round; ps.pass = p; ps.shift = prof.shift[p]; ps.reserved = prof.spread[p];
                ps.n_records = n_rec; ps.ms = prof.ms[p];
                pass_stats_.push_back(ps);
            }
        }
    };

    const int rbits = bit_width_u64((uint64_t)un);       // ranks 0..n  (0 = past the end)
    const int gbits = bit_width_u64((uint64_t)un - 1);   // group ranks 0..n-1
    bool in_alt = false;
    stats_.active_per_round[0] = n;
    PSS_TRY(sorter_.sort(keys_a_, keys_b_, vals_a_, vals_b_, un, 0, m * b, /*iota=*/true, s, &in_alt, &prof,
                         /*hist_done=*/true, /*defer_check=*/true));

    uint32_t partition_min = 1u << 22;
    if (const char *e = std::getenv("PSS_PARTITION_MIN")) partition_min = (uint32_t)std::strtoul(e, nullptr, 10);
    uint32_t  n_active = un;
    uint32_t *v_sorted = in_alt ? vals_b_ : vals_a_;
    uint32_t *v_free   = in_alt ? vals_a_ : vals_b_;
    uint64_t *k_sorted = in_alt ? keys_b_ : keys_a_;
    auto rerank = [&](bool first) -> int {
        const uint32_t tiles = (uint32_t)div_up(n_active, RR_TILE);
        PSS_CUDA_TRY(cudaMemsetAsync(tile_aggr_, 0, (size_t)tiles * 2 * sizeof(unsigned long long), s));
        PSS_CUDA_TRY(cudaMemsetAsync(d_small_ + SM_SCALARS, 0, 8 * sizeof(uint32_t), s));
        // Above the threshold the rank scatter is partitioned (pairs → one keys-only onesweep
        // pass on the index's top 8 bits → windowed scatter); below it the direct scatter
        // is cheaper than the extra launches.
        const bool partitioned = n_active >= partition_min;
        uint64_t *k_other = (k_sorted == keys_a_) ? keys_b_ : keys_a_;
        rerank_apply_kernel<<<tiles, RR_THREADS, 0, s>>>(k_sorted, v_sorted, n_active, rbits, first ? 1 : 0,
                                                         reinterpret_cast<unsigned long long *>(tile_aggr_),
                                                         d_small_ + SM_SCALARS, isa_, partitioned ? k_other : nullptr,
                                                         d_sa, v_free, grp_);
        PSS_LAUNCH_CHECK();
        if (partitioned) {
            bool alt = false;
            const int ibits = bit_width_u64((uint64_t)un - 1);
            const int shift = 32 + std::max(0, ibits - RADIX_BITS);
            PSS_TRY(sorter_.partition(k_other, k_sorted, n_active, shift, (uint32_t)(RADIX - 1), s, &alt));
            isa_scatter_kernel<<<(unsigned)div_up(n_active, 256 * 4), 256, 0, s>>>(
                alt ? k_sorted : k_other, n_active, isa_, v_free, grp_, d_small_ + SM_SCALARS + 3);
            PSS_LAUNCH_CHECK();
        }
        // ONE read-back per round: active count, this kernel's watchdog and the sorter's (the
        // sort before and the partition pass above deferred their checks to here)
        PSS_TRY(copy_words(d_small_ + SM_SCALARS + 4, sorter_.d_error_flag(), 1, s));
        PSS_TRY(copy_words(h_small_ + SM_SCALARS, d_small_ + SM_SCALARS, 5, s));
        PSS_CUDA_TRY(cudaStreamSynchronize(s));
        if (h_small_[SM_SCALARS + 2]) return fail(PSS_ERR_CUDA, "re-rank: look-back watchdog fired");
        PSS_TRY(sorter_.check_error_word(h_small_[SM_SCALARS + 4], s));
        n_active = h_small_[SM_SCALARS];
        return PSS_OK;
    };
    PSS_TRY(rerank(true));
    PSS_TRY(sorter_.collect_times(&prof));
    record_passes(0, un);

    // ---- doubling rounds ------------------------------------------------------------------
    uint64_t h = (uint64_t)m;
    int round = 0;
    while (n_active > 0) {
        ++round;
        if (round >= 63 || h >= (uint64_t)n * 2) return fail(PSS_ERR_CUDA, "prefix doubling failed to converge");
        stats_.active_per_round[round] = n_active;
        uint32_t *v_in = v_free;  // compacted active suffix indices written by the last re-rank
        uint32_t *v_alt = (v_in == vals_a_) ? vals_b_ : vals_a_;
        PSS_TRY(sorter_.hist_reset(s));
        {
            const int grid = (int)std::min<int64_t>(div_up(n_active, 256 * 4), (int64_t)sorter_.num_sms() * gather_ctas_);
            gather_kernel<<<grid, 256, 0, s>>>(v_in, grp_, isa_, un, (uint32_t)std::min<uint64_t>(h, un), rbits, n_active,
                                               keys_a_, hist_layout(0, rbits + gbits), sorter_.d_hist(), l2_hints_);
            PSS_LAUNCH_CHECK();
        }
        PSS_TRY(sorter_.sort(keys_a_, keys_b_, v_in, v_alt, n_active, 0, rbits + gbits, /*iota=*/false, s, &in_alt,
                             &prof, /*hist_done=*/true, /*defer_check=*/true));
        const uint32_t n_sorted = n_active;
        k_sorted = in_alt ? keys_b_ : keys_a_;
        v_sorted = in_alt ? v_alt : v_in;
        v_free   = in_alt ? v_in : v_alt;
        PSS_TRY(rerank(false));
        PSS_TRY(sorter_.collect_times(&prof));
        record_passes(round, n_sorted);
        h *= 2;
    }
    stats_.rounds = round;

    PSS_CUDA_TRY(cudaEventRecord(ev_end_, s));
    PSS_CUDA_TRY(cudaEventSynchronize(ev_end_));
    PSS_CUDA_TRY(cudaEventElapsedTime(&stats_.total_ms, ev_begin_, ev_end_));
    stats_.n_pass_stats = (int32_t)pass_stats_.size();
    stats_.n_kernel_launches = (int32_t)(g_kernel_launches.load() - launches0);
    return PSS_OK;
}

// ---- host <-> device copies for callers with ordinary (pageable) memory (HostStager) --------------
namespace {

constexpr size_t STAGE_SLICE = 16u << 20;
constexpr size_t STAGE_MIN   = 8u << 20;   // below this a plain cudaMemcpy is as good


void parallel_memcpy(uint8_t *dst, const uint8_t *src, size_t len, int threads) {
    if (threads <= 1 || len < (1u << 20)) {
        std::memcpy(dst, src, len);
        return;
    }
    const size_t part = (len / threads + 4095) & ~(size_t)4095;
    std::vector<std::thread> pool;
    size_t off = 0;
    for (int t = 0; t < threads - 1 && off + part < len; ++t, off += part)
        pool.emplace_back([=] { std::memcpy(dst + off, src + off, part); });
    std::memcpy(dst + off, src + off, len - off);
    for (auto &th : pool) th.join();
}

int copy_threads() {
    unsigned hc = std::thread::hardware_concurrency();
    int t = hc ? (int)std::min<unsigned>(8, std::max<unsigned>(1, hc / 2)) : 4;
    if (const char *e = std::getenv("PSS_COPY_THREADS")) t = std::max(1, std::atoi(e));
    return t;
}

}  // namespace

bool host_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

void HostStager::release() {
    for (int i = 0; i < 2; ++i) {
        if (stage_[i]) cudaFreeHost(stage_[i]);
        if (ev_[i]) cudaEventDestroy(ev_[i]);
        stage_[i]   = nullptr;
        ev_[i]      = nullptr;
        pending_[i] = false;
    }
}

int HostStager::copy(void *dst, const void *src, size_t bytes, bool to_device, cudaStream_t stream) {
    if (bytes == 0) return PSS_OK;
    const void *host = to_device ? src : dst;
    if (bytes < STAGE_MIN || host_is_pinned(host)) {
        PSS_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, stream));
        // pageable + small: the runtime has staged the bytes on return; pinned: the caller's
        // buffer must stay untouched until the stream reaches this point
        if (!to_device) PSS_CUDA_TRY(cudaStreamSynchronize(stream));
        return PSS_OK;
    }
    if (!stage_[0]) {
        for (int i = 0; i < 2; ++i) {
            PSS_CUDA_TRY(cudaMallocHost(&stage_[i], STAGE_SLICE));
            PSS_CUDA_TRY(cudaEventCreateWithFlags(&ev_[i], cudaEventDisableTiming));
        }
    }
    const int threads = copy_threads();
    uint8_t *d = static_cast<uint8_t *>(dst);
    const uint8_t *s = static_cast<const uint8_t *>(src);
    const size_t nsl = (bytes + STAGE_SLICE - 1) / STAGE_SLICE;
    auto len_of = [&](size_t i) { return std::min(STAGE_SLICE, bytes - i * STAGE_SLICE); };
    auto wait_free = [&](int b) -> int {     // the DMA that last used bounce buffer b has finished
        if (pending_[b]) {
            PSS_CUDA_TRY(cudaEventSynchronize(ev_[b]));
            pending_[b] = false;
        }
        return PSS_OK;
    };
    if (to_device) {
        for (size_t i = 0; i < nsl; ++i) {
            const int b = (int)(i & 1);
            PSS_TRY(wait_free(b));
            parallel_memcpy(static_cast<uint8_t *>(stage_[b]), s + i * STAGE_SLICE, len_of(i), threads);
            PSS_CUDA_TRY(cudaMemcpyAsync(d + i * STAGE_SLICE, stage_[b], len_of(i), cudaMemcpyHostToDevice, stream));
            PSS_CUDA_TRY(cudaEventRecord(ev_[b], stream));
            pending_[b] = true;
        }
        return PSS_OK;   // completion is ordered on `stream`
    }
    for (size_t i = 0; i <= nsl; ++i) {
        if (i < nsl) {
            const int b = (int)(i & 1);
            PSS_TRY(wait_free(b));
            PSS_CUDA_TRY(cudaMemcpyAsync(stage_[b], s + i * STAGE_SLICE, len_of(i), cudaMemcpyDeviceToHost, stream));
            PSS_CUDA_TRY(cudaEventRecord(ev_[b], stream));
            pending_[b] = true;
        }
        if (i >= 1) {   // while slice i is in flight, move slice i-1 out of its bounce buffer
            const int pb = (int)((i - 1) & 1);
            PSS_TRY(wait_free(pb));
            parallel_memcpy(d + (i - 1) * STAGE_SLICE, static_cast<const uint8_t *>(stage_[pb]), len_of(i - 1), threads);
        }
    }
    return PSS_OK;
}

int SaBuilder::build_host(const uint8_t *h_text, int32_t n, int32_t *h_sa) {
    if (n < 0 || (n > 0 && (!h_text || !h_sa))) return fail(PSS_ERR_ARG, "bad build arguments");
    if (n == 0) return PSS_OK;
    PSS_CUDA_TRY(cudaSetDevice(device_));
    PSS_TRY(ensure_io(n));
    PSS_TRY(stager_.copy(d_text_, h_text, (size_t)n, /*to_device=*/true, stream_));
    PSS_TRY(build_device(d_text_, n, d_sa_, stream_));
    PSS_TRY(stager_.copy(h_sa, d_sa_, (size_t)n * sizeof(int32_t), /*to_device=*/false, stream_));
    return PSS_OK;
}

}  // namespace pss
