// search.cu — batched substring search kernels for sm_100a (see search.cuh).
#include "search.cuh"

#include <algorithm>
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
                                            const uint8_t *__restrict__ P, uint32_t m, uint32_t lane,
                                            uint32_t *lb_out, uint32_t *cnt_out) {
    const uint32_t pc0 = lane < m ? (uint32_t)__ldg(P + lane) : 0u;
    // smallest slot whose suffix is >= P (as a prefix comparison)
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (cmp_suffix(text, n, (uint32_t)__ldg(sa + mid), P, m, pc0, lane) < 0) lo = mid + 1;
        else hi = mid;
    }
    const uint32_t lb = lo;
    // smallest slot past lb whose suffix is > P and does not start with it
    hi = n;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (cmp_suffix(text, n, (uint32_t)__ldg(sa + mid), P, m, pc0, lane) <= 0) lo = mid + 1;
        else hi = mid;
    }
    *lb_out  = lb;
    *cnt_out = lo - lb;
}

__global__ void __launch_bounds__(BD_THREADS)
bounds_kernel(const DeviceChunk *__restrict__ chunks, int nc, const uint8_t *__restrict__ patterns,
              const int64_t *__restrict__ pat_off, uint32_t npairs, uint32_t *__restrict__ lb_out,
              uint32_t *__restrict__ cnt_out) {
    const uint32_t lane = lane_id();
    const uint32_t pair = (blockIdx.x * BD_THREADS + threadIdx.x) >> 5;
    if (pair >= npairs) return;
    const uint32_t q = pair / (uint32_t)nc, c = pair % (uint32_t)nc;
    uint32_t lb, cnt;
    warp_bounds(chunks[c].text, chunks[c].sa, chunks[c].n, patterns + pat_off[q],
                (uint32_t)(pat_off[q + 1] - pat_off[q]), lane, &lb, &cnt);
    if (lane == 0) {
        lb_out[pair]  = lb;
        cnt_out[pair] = cnt;
    }
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

// first '\n' at or after pos (lib.rs:266-269; none → n - 1)
__device__ __forceinline__ uint32_t next_newline(const uint8_t *__restrict__ text, uint32_t n, uint32_t pos) {
    uint32_t i  = pos & ~3u;
    uint32_t eq = __vcmpeq4(ld_text_word(text, i), 0x0A0A0A0Au) & (0xFFFFFFFFu << (8 * (pos & 3u)));
    while (eq == 0) {
        i += 4;
        if (i >= n) return n - 1;
        eq = __vcmpeq4(ld_text_word(text, i), 0x0A0A0A0Au);
    }
    uint32_t e = i + ((__ffs(eq) - 1) >> 3);
    return e < n ? e : n - 1;
}

// 1 + last '\n' strictly before pos (lib.rs:270-273; none → 0)
__device__ __forceinline__ uint32_t line_begin(const uint8_t *__restrict__ text, uint32_t pos) {
    if (pos == 0) return 0;
    const uint32_t q = pos - 1;
    uint32_t i  = q & ~3u;
    uint32_t eq = __vcmpeq4(ld_text_word(text, i), 0x0A0A0A0Au) & (0xFFFFFFFFu >> (8 * (3u - (q & 3u))));
    while (eq == 0) {
        if (i == 0) return 0;
        i -= 4;
        eq = __vcmpeq4(ld_text_word(text, i), 0x0A0A0A0Au);
    }
    return i + ((31 - __clz(eq)) >> 3) + 1;
}

__global__ void __launch_bounds__(256)
extract_kernel(const DeviceChunk *__restrict__ chunks, int nc, uint32_t pair_base, uint32_t npairs,
               const uint32_t *__restrict__ hit_off, const uint32_t *__restrict__ lb, uint32_t nhits, int sbits,
               uint64_t *__restrict__ keys, uint32_t *__restrict__ line_end) {
    const uint32_t f = blockIdx.x * 256 + threadIdx.x;
    if (f >= nhits) return;
    const uint32_t p    = find_pair(hit_off, npairs, f);
    const uint32_t pair = pair_base + p;
    const uint32_t c    = pair % (uint32_t)nc;
    const uint8_t *text = chunks[c].text;
    const uint32_t n    = chunks[c].n;
    const uint32_t pos  = (uint32_t)__ldg(chunks[c].sa + __ldg(lb + pair) + (f - __ldg(hit_off + p)));
    const uint32_t e    = next_newline(text, n, pos);
    const uint32_t b    = line_begin(text, pos);
    keys[f]     = ((uint64_t)p << sbits) | b;
    line_end[f] = e;
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

__device__ __forceinline__ uint32_t block_excl_sum(uint32_t v, uint32_t *s_warp, uint32_t *total) {
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
    for (int w = 0; w < CP_THREADS / 32; ++w) {
        uint32_t t = s_warp[w];
        if ((uint32_t)w < warp) pre += t;
        tot += t;
    }
    *total = tot;
    __syncthreads();
    return pre + incl - v;
}

__global__ void __launch_bounds__(CP_THREADS)
flag_reduce_kernel(const uint32_t *__restrict__ flag, uint32_t nhits, uint32_t *__restrict__ tile_sum) {
    __shared__ uint32_t s_warp[CP_THREADS / 32];
    const uint32_t base = blockIdx.x * CP_TILE + threadIdx.x * CP_IPT;
    uint32_t c = 0;
#pragma unroll
    for (int e = 0; e < CP_IPT; ++e)
        if (base + e < nhits) c += flag[base + e] >> 31;
    uint32_t total;
    (void)block_excl_sum(c, s_warp, &total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
tile_scan_kernel(uint32_t *__restrict__ tile_sum, uint32_t tiles, uint32_t *__restrict__ total_out) {
    __shared__ uint32_t s_part[1024];
    const uint32_t per = (tiles + 1023) / 1024;
    const uint32_t lo  = min(tiles, threadIdx.x * per), hi = min(tiles, lo + per);
    uint32_t sum = 0;
    for (uint32_t t = lo; t < hi; ++t) sum += tile_sum[t];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele over 1024 partials
    for (int o = 1; o < 1024; o <<= 1) {
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
    if (threadIdx.x == 1023) *total_out = s_part[1023];
}

__global__ void __launch_bounds__(CP_THREADS)
compact_kernel(const uint32_t *__restrict__ flag, const uint32_t *__restrict__ line_end,
               const uint32_t *__restrict__ tile_prefix, const uint32_t *__restrict__ hit_off,
               const DeviceChunk *__restrict__ chunks, int nc, uint32_t pair_base, uint32_t npairs, uint32_t nhits,
               uint32_t *__restrict__ pair_first, int32_t *__restrict__ out_query, int32_t *__restrict__ out_chunk,
               uint32_t *__restrict__ out_start, uint32_t *__restrict__ out_end) {
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
    uint32_t o = block_excl_sum(c, s_warp, &total) + tile_prefix[blockIdx.x];
    if (base >= nhits) return;
    uint32_t p = find_pair(hit_off, npairs, base);
#pragma unroll
    for (int e = 0; e < CP_IPT; ++e) {
        const uint32_t f = base + e;
        if (f >= nhits) break;
        while (f >= __ldg(hit_off + p + 1)) ++p;           // skips pairs without hits
        if (f == __ldg(hit_off + p)) pair_first[p] = o;    // first hit of pair p: its output offset
        if (fl[e] >> 31) {
            const uint32_t pair = pair_base + p;
            if (out_query) out_query[o] = (int32_t)(pair / (uint32_t)nc);
            if (out_chunk) out_chunk[o] = chunks[pair % (uint32_t)nc].global_id;
            out_start[o] = fl[e] & 0x7FFFFFFFu;
            out_end[o]   = line_end[f];
            ++o;
        }
    }
}


// ------------------------------------------------------------------------------------
// Small-batch path: one CTA answers a handful of (query, chunk) pairs end to end — bounds,
// extraction, dedup (bitonic sort in shared memory) and compaction — so that a single
// `Reader.search` costs one kernel launch and one device→host copy instead of ~15 launches.
// Falls back (status = 1) when the pairs have more than SMALL_CAP matching suffixes.
// ------------------------------------------------------------------------------------
constexpr int SMALL_THREADS   = 1024;
constexpr int SMALL_CAP       = 8192;   // matching suffixes handled in shared memory
constexpr int SMALL_MAX_PAIRS = 64;

struct SmallHeader {
    uint32_t status;      // 0 = answered, 1 = too many hits (use the general path)
    uint32_t n_hits;
    uint32_t n_entries;
    uint32_t reserved;
    uint32_t pair_entries[SMALL_MAX_PAIRS];
};
// Packed result buffer: header, then query / chunk / start / end arrays of SMALL_CAP each.
constexpr size_t SMALL_OUT_BYTES = sizeof(SmallHeader) + 4 * (size_t)SMALL_CAP * sizeof(uint32_t);

struct SmallSmem {
    uint64_t key[SMALL_CAP];       // (pair << 43) | (entry start << 13) | hit index
    uint32_t end[SMALL_CAP];       // entry end, by hit index
    uint32_t start[SMALL_CAP];     // bit 31 = first hit of its entry, low bits = entry start
    uint32_t lb[SMALL_MAX_PAIRS], cnt[SMALL_MAX_PAIRS], off[SMALL_MAX_PAIRS + 1], pair_entries[SMALL_MAX_PAIRS];
    uint32_t warp_sum[SMALL_THREADS / 32];
    uint32_t total;
};

__global__ void __launch_bounds__(SMALL_THREADS, 1)
small_search_kernel(const DeviceChunk *__restrict__ chunks, int nc, const uint8_t *__restrict__ patterns,
                    const int64_t *__restrict__ pat_off, uint32_t npairs, uint32_t *__restrict__ lb_out,
                    uint32_t *__restrict__ cnt_out, unsigned char *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmallSmem &s = *reinterpret_cast<SmallSmem *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    SmallHeader *hdr   = reinterpret_cast<SmallHeader *>(out);
    int32_t  *o_query  = reinterpret_cast<int32_t *>(out + sizeof(SmallHeader));
    int32_t  *o_chunk  = o_query + SMALL_CAP;
    uint32_t *o_start  = reinterpret_cast<uint32_t *>(o_chunk + SMALL_CAP);
    uint32_t *o_end    = o_start + SMALL_CAP;

    // ---- bounds: one warp per pair (same search as bounds_kernel) -------------------------------
    for (uint32_t pair = warp; pair < npairs; pair += SMALL_THREADS / 32) {
        const uint32_t q = pair / (uint32_t)nc, c = pair % (uint32_t)nc;
        uint32_t lb, cnt;
        warp_bounds(chunks[c].text, chunks[c].sa, chunks[c].n, patterns + pat_off[q],
                    (uint32_t)(pat_off[q + 1] - pat_off[q]), lane, &lb, &cnt);
        if (lane == 0) {
            s.lb[pair] = lb;
            s.cnt[pair] = cnt;
            lb_out[pair] = lb;
            cnt_out[pair] = cnt;
        }
    }
    if (tid < SMALL_MAX_PAIRS) s.pair_entries[tid] = 0;
    __syncthreads();
    if (tid == 0) {
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
        if (tid == 0) { hdr->status = 1; hdr->n_hits = H; hdr->n_entries = 0; }
        return;
    }
    uint32_t P2 = 1;
    while (P2 < H) P2 <<= 1;

    // ---- extract: entry boundaries of every matching suffix ---------------------------------
    for (uint32_t f = tid; f < P2; f += SMALL_THREADS) {
        uint64_t key = ~0ull;
        if (f < H) {
            uint32_t p = 0;
            while (s.off[p + 1] <= f) ++p;
            const uint32_t c    = p % (uint32_t)nc;
            const uint8_t *text = chunks[c].text;
            const uint32_t n    = chunks[c].n;
            const uint32_t pos  = (uint32_t)__ldg(chunks[c].sa + s.lb[p] + (f - s.off[p]));
            s.end[f]   = next_newline(text, n, pos);
            s.start[f] = 0;
            key = ((uint64_t)p << 43) | ((uint64_t)line_begin(text, pos) << 13) | f;
        }
        s.key[f] = key;
    }
    __syncthreads();

    // ---- dedup: bitonic sort by (pair, entry start, hit index); the head of every
    //      (pair, entry start) run is the entry's first hit in SA order -------------------------
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

    // ---- compaction in hit order = (query, chunk, SA order) -------------------------------------
    constexpr int PER = SMALL_CAP / SMALL_THREADS;
    const uint32_t base = tid * PER;
    uint32_t kept = 0;
#pragma unroll
    for (int e = 0; e < PER; ++e)
        if (base + e < H) kept += s.start[base + e] >> 31;
    uint32_t incl = kept;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += y;
    }
    if (lane == 31) s.warp_sum[warp] = incl;
    __syncthreads();
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
            const uint32_t v = s.start[f];
            if (v >> 31) {
                o_query[o] = (int32_t)(p / (uint32_t)nc);
                o_chunk[o] = chunks[p % (uint32_t)nc].global_id;
                o_start[o] = v & 0x7FFFFFFFu;
                o_end[o]   = s.end[f];
                atomicAdd(&s.pair_entries[p], 1u);
                ++o;
            }
        }
    }
    __syncthreads();
    if (tid < SMALL_MAX_PAIRS) hdr->pair_entries[tid] = s.pair_entries[tid];
    if (tid == 0) { hdr->status = 0; hdr->n_hits = H; hdr->n_entries = tot; }
}

}  // namespace

// ------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------
int SearchSink::deliver_host(int64_t count, const int32_t *q, const int32_t *c, const uint32_t *s, const uint32_t *e,
                             cudaStream_t st) {
    int32_t *dq = nullptr, *dc = nullptr;
    uint32_t *ds = nullptr, *de = nullptr;
    PSS_TRY(reserve(count, &dq, &dc, &ds, &de));
    if (count) {
        if (dq) PSS_CUDA_TRY(cudaMemcpyAsync(dq, q, count * 4, cudaMemcpyHostToDevice, st));
        if (dc) PSS_CUDA_TRY(cudaMemcpyAsync(dc, c, count * 4, cudaMemcpyHostToDevice, st));
        PSS_CUDA_TRY(cudaMemcpyAsync(ds, s, count * 4, cudaMemcpyHostToDevice, st));
        PSS_CUDA_TRY(cudaMemcpyAsync(de, e, count * 4, cudaMemcpyHostToDevice, st));
    }
    PSS_TRY(commit(count, st));
    PSS_CUDA_TRY(cudaStreamSynchronize(st));
    return PSS_OK;
}

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
    PSS_CUDA_TRY(cudaMallocHost(&h_scalar_, 16 * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_small_out_, SMALL_OUT_BYTES));
    PSS_CUDA_TRY(cudaMemset(d_small_out_, 0, SMALL_OUT_BYTES));   // the whole block is copied back every call
    PSS_CUDA_TRY(cudaMallocHost(&h_small_out_, SMALL_OUT_BYTES));
    PSS_CUDA_TRY(cudaFuncSetAttribute(small_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(SmallSmem)));
    small_path_ = true;
    if (const char *e = std::getenv("PSS_SMALL_PATH")) small_path_ = std::atoi(e) != 0;
    PSS_TRY(sorter_.init(device_));
    return PSS_OK;
}

void Searcher::release() {
    if (device_ < 0) return;
    cudaSetDevice(device_);
    cudaFree(d_chunks_);
    cudaFree(d_lb_); cudaFree(d_cnt_); cudaFree(d_hit_off_); cudaFree(d_pair_first_);
    if (h_lb_) cudaFreeHost(h_lb_);
    if (h_cnt_) cudaFreeHost(h_cnt_);
    if (h_hit_off_) cudaFreeHost(h_hit_off_);
    if (h_pair_first_) cudaFreeHost(h_pair_first_);
    cudaFree(d_keys_); cudaFree(d_keys_alt_); cudaFree(d_vals_); cudaFree(d_vals_alt_);
    cudaFree(d_end_); cudaFree(d_flag_); cudaFree(d_tile_sum_); cudaFree(d_scalar_);
    cudaFree(d_small_out_);
    if (h_small_out_) cudaFreeHost(h_small_out_);
    d_small_out_ = h_small_out_ = nullptr;
    if (h_scalar_) cudaFreeHost(h_scalar_);
    for (auto &e : ev_)
        if (e) cudaEventDestroy(e);
    if (stream_) cudaStreamDestroy(stream_);
    sorter_.release();
    d_chunks_ = nullptr;
    d_lb_ = d_cnt_ = d_hit_off_ = d_pair_first_ = nullptr;
    h_lb_ = h_cnt_ = h_hit_off_ = h_pair_first_ = nullptr;
    d_keys_ = d_keys_alt_ = nullptr;
    d_vals_ = d_vals_alt_ = d_end_ = d_flag_ = d_tile_sum_ = d_scalar_ = h_scalar_ = nullptr;
    for (auto &e : ev_) e = nullptr;
    stream_ = nullptr;
    pair_cap_ = hit_cap_ = 0;
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

int Searcher::ensure_pairs(int64_t npairs) {
    if (npairs <= pair_cap_) return PSS_OK;
    cudaFree(d_lb_); cudaFree(d_cnt_); cudaFree(d_hit_off_); cudaFree(d_pair_first_);
    if (h_lb_) cudaFreeHost(h_lb_);
    if (h_cnt_) cudaFreeHost(h_cnt_);
    if (h_hit_off_) cudaFreeHost(h_hit_off_);
    if (h_pair_first_) cudaFreeHost(h_pair_first_);
    d_lb_ = d_cnt_ = d_hit_off_ = d_pair_first_ = nullptr;
    h_lb_ = h_cnt_ = h_hit_off_ = h_pair_first_ = nullptr;
    pair_cap_ = 0;
    int64_t cap = std::max<int64_t>(npairs, 1024);
    PSS_CUDA_TRY(cudaMalloc(&d_lb_, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_cnt_, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_hit_off_, (cap + 1) * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_pair_first_, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMallocHost(&h_lb_, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMallocHost(&h_cnt_, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMallocHost(&h_hit_off_, (cap + 1) * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMallocHost(&h_pair_first_, cap * sizeof(uint32_t)));
    pair_cap_ = cap;
    return PSS_OK;
}

int Searcher::ensure_hits(int64_t nhits) {
    if (nhits <= hit_cap_) return PSS_OK;
    cudaFree(d_keys_); cudaFree(d_keys_alt_); cudaFree(d_vals_); cudaFree(d_vals_alt_);
    cudaFree(d_end_); cudaFree(d_flag_); cudaFree(d_tile_sum_);
    d_keys_ = d_keys_alt_ = nullptr;
    d_vals_ = d_vals_alt_ = d_end_ = d_flag_ = d_tile_sum_ = nullptr;
    hit_cap_ = 0;
    int64_t cap = std::max<int64_t>(nhits + nhits / 4, 1 << 16);
    if (cap >= (1ll << 30)) cap = (1ll << 30) - 1;
    PSS_CUDA_TRY(cudaMalloc(&d_keys_, cap * sizeof(uint64_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_keys_alt_, cap * sizeof(uint64_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_vals_, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_vals_alt_, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_end_, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_flag_, cap * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_tile_sum_, (size_t)div_up(cap, CP_TILE) * sizeof(uint32_t)));
    PSS_TRY(sorter_.ensure(cap));
    hit_cap_ = cap;
    return PSS_OK;
}

int Searcher::search(const uint8_t *d_patterns, const int64_t *d_offsets, int32_t nq, cudaStream_t stream,
                     SearchSink *sink, int64_t *per_pair_count, int64_t *n_hits, SearchTimes *times) {
    if (device_ < 0) return fail(PSS_ERR_ARG, "searcher not initialised");
    if (nq < 0 || !sink) return fail(PSS_ERR_ARG, "bad search arguments");
    if (n_hits) *n_hits = 0;
    if (times) *times = SearchTimes();
    const int nc = (int)chunks_.size();
    const int64_t npairs64 = (int64_t)nq * nc;
    if (npairs64 == 0) return PSS_OK;
    if (npairs64 >= (1ll << 31)) return fail(PSS_ERR_ARG, "too many (query, chunk) pairs in one batch");
    PSS_CUDA_TRY(cudaSetDevice(device_));
    cudaStream_t s = stream ? stream : stream_;
    const uint32_t npairs = (uint32_t)npairs64;
    PSS_TRY(ensure_pairs(npairs));

    uint32_t max_n = 1;
    for (const auto &c : chunks_) max_n = std::max(max_n, c.n);
    const int sbits = std::max(1, bit_width_u64((uint64_t)max_n - 1));

    // ---- small batches: one fused kernel, one device→host copy -------------------------------
    bool have_bounds = false;
    if (small_path_ && npairs <= (uint32_t)SMALL_MAX_PAIRS) {
        PSS_CUDA_TRY(cudaEventRecord(ev_[0], s));
        small_search_kernel<<<1, SMALL_THREADS, sizeof(SmallSmem), s>>>(d_chunks_, nc, d_patterns, d_offsets, npairs, d_lb_,
                                                                       d_cnt_, d_small_out_);
        PSS_LAUNCH_CHECK();
        PSS_CUDA_TRY(cudaEventRecord(ev_[1], s));
        PSS_CUDA_TRY(cudaMemcpyAsync(h_small_out_, d_small_out_, SMALL_OUT_BYTES, cudaMemcpyDeviceToHost, s));
        PSS_CUDA_TRY(cudaStreamSynchronize(s));
        const SmallHeader *hdr = reinterpret_cast<const SmallHeader *>(h_small_out_);
        if (hdr->status == 0) {
            const int32_t *hq  = reinterpret_cast<const int32_t *>(h_small_out_ + sizeof(SmallHeader));
            const int32_t *hc  = hq + SMALL_CAP;
            const uint32_t *hs = reinterpret_cast<const uint32_t *>(hc + SMALL_CAP);
            const uint32_t *he = hs + SMALL_CAP;
            if (n_hits) *n_hits = hdr->n_hits;
            if (per_pair_count)
                for (uint32_t p = 0; p < npairs; ++p) per_pair_count[p] = hdr->pair_entries[p];
            if (times) {
                float ms = 0.f;
                PSS_CUDA_TRY(cudaEventElapsedTime(&ms, ev_[0], ev_[1]));
                times->ms_bounds = ms;   // the fused kernel: bounds + extract + dedup
            }
            return sink->deliver_host(hdr->n_entries, hq, hc, hs, he, s);
        }
        have_bounds = true;   // too many hits: lb/cnt are already in d_lb_/d_cnt_
    }

    // ---- bounds -----------------------------------------------------------------------
    PSS_CUDA_TRY(cudaEventRecord(ev_[0], s));
    if (!have_bounds) {
    bounds_kernel<<<(unsigned)div_up((int64_t)npairs * 32, BD_THREADS), BD_THREADS, 0, s>>>(
        d_chunks_, nc, d_patterns, d_offsets, npairs, d_lb_, d_cnt_);
    PSS_LAUNCH_CHECK();
    }
    PSS_CUDA_TRY(cudaEventRecord(ev_[1], s));
    PSS_CUDA_TRY(cudaMemcpyAsync(h_cnt_, d_cnt_, npairs * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    PSS_CUDA_TRY(cudaStreamSynchronize(s));
    {
        float ms = 0.f;
        PSS_CUDA_TRY(cudaEventElapsedTime(&ms, ev_[0], ev_[1]));
        if (times) times->ms_bounds += ms;
    }
    if (per_pair_count) std::memset(per_pair_count, 0, sizeof(int64_t) * npairs);

    // ---- sub-batches of pairs whose hits fit the workspace ----------------------------------
    constexpr int64_t HIT_BUDGET = 1ll << 27;
    uint32_t a = 0;
    while (a < npairs) {
        int64_t H = 0;
        uint32_t b = a;
        while (b < npairs && (H == 0 || H + h_cnt_[b] <= HIT_BUDGET)) {
            H += h_cnt_[b];
            ++b;
        }
        if (H >= (1ll << 30)) return fail(PSS_ERR_ARG, "a single (query, chunk) pair has >= 2^30 hits");
        const uint32_t np = b - a;
        if (H == 0) { a = b; continue; }
        if (n_hits) *n_hits += H;
        uint32_t run = 0;
        for (uint32_t p = 0; p < np; ++p) {
            h_hit_off_[p] = run;
            run += h_cnt_[a + p];
        }
        h_hit_off_[np] = run;
        const uint32_t nh = (uint32_t)H;
        PSS_TRY(ensure_hits(H));
        PSS_CUDA_TRY(cudaMemcpyAsync(d_hit_off_, h_hit_off_, (np + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s));

        PSS_CUDA_TRY(cudaEventRecord(ev_[2], s));
        extract_kernel<<<(unsigned)div_up(nh, 256), 256, 0, s>>>(d_chunks_, nc, a, np, d_hit_off_, d_lb_, nh, sbits,
                                                                d_keys_, d_end_);
        PSS_LAUNCH_CHECK();
        PSS_CUDA_TRY(cudaEventRecord(ev_[3], s));

        bool in_alt = false;
        const int end_bit = sbits + std::max(1, bit_width_u64((uint64_t)np - 1));
        PSS_TRY(sorter_.sort(d_keys_, d_keys_alt_, d_vals_, d_vals_alt_, nh, 0, end_bit, /*iota=*/true, s, &in_alt, nullptr));
        const uint64_t *k_sorted = in_alt ? d_keys_alt_ : d_keys_;
        const uint32_t *v_sorted = in_alt ? d_vals_alt_ : d_vals_;
        const uint32_t tiles = (uint32_t)div_up(nh, CP_TILE);
        PSS_CUDA_TRY(cudaMemsetAsync(d_flag_, 0, (size_t)nh * sizeof(uint32_t), s));
        mark_kernel<<<(unsigned)div_up(nh, 256), 256, 0, s>>>(k_sorted, v_sorted, nh, sbits, d_flag_);
        PSS_LAUNCH_CHECK();
        flag_reduce_kernel<<<tiles, CP_THREADS, 0, s>>>(d_flag_, nh, d_tile_sum_);
        PSS_LAUNCH_CHECK();
        tile_scan_kernel<<<1, 1024, 0, s>>>(d_tile_sum_, tiles, d_scalar_);
        PSS_LAUNCH_CHECK();
        PSS_CUDA_TRY(cudaMemcpyAsync(h_scalar_, d_scalar_, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        PSS_CUDA_TRY(cudaStreamSynchronize(s));
        const int64_t kept = h_scalar_[0];

        int32_t *o_q = nullptr, *o_c = nullptr;
        uint32_t *o_s = nullptr, *o_e = nullptr;
        PSS_TRY(sink->reserve(kept, &o_q, &o_c, &o_s, &o_e));
        if (!o_s || !o_e) return fail(PSS_ERR_ARG, "search sink returned null output buffers");
        if (per_pair_count)   // pairs without hits never get an entry: keep the copied-back array defined
            PSS_CUDA_TRY(cudaMemsetAsync(d_pair_first_, 0, (size_t)np * sizeof(uint32_t), s));
        compact_kernel<<<tiles, CP_THREADS, 0, s>>>(d_flag_, d_end_, d_tile_sum_, d_hit_off_, d_chunks_, nc, a, np, nh,
                                                    d_pair_first_, o_q, o_c, o_s, o_e);
        PSS_LAUNCH_CHECK();
        PSS_CUDA_TRY(cudaEventRecord(ev_[4], s));
        PSS_TRY(sink->commit(kept, s));
        if (per_pair_count) {
            PSS_CUDA_TRY(cudaMemcpyAsync(h_pair_first_, d_pair_first_, np * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            PSS_CUDA_TRY(cudaStreamSynchronize(s));
            // pairs without hits have no pair_first entry: walk backwards from the total
            int64_t next = kept;
            for (int64_t p = (int64_t)np - 1; p >= 0; --p) {
                if (h_cnt_[a + p] == 0) continue;
                per_pair_count[a + p] = next - (int64_t)h_pair_first_[p];
                next = h_pair_first_[p];
            }
        } else {
            PSS_CUDA_TRY(cudaStreamSynchronize(s));
        }
        if (times) {
            float ms = 0.f;
            PSS_CUDA_TRY(cudaEventElapsedTime(&ms, ev_[2], ev_[3]));
            times->ms_extract += ms;
            PSS_CUDA_TRY(cudaEventElapsedTime(&ms, ev_[3], ev_[4]));
            times->ms_dedup += ms;
        }
        a = b;
    }
    return PSS_OK;
}

}  // namespace pss
