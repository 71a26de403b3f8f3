// radix_sort.cu — onesweep LSD radix sort kernels for sm_100a (see radix_sort.cuh).
#include "radix_sort.cuh"

#include <algorithm>
#include <cstring>

namespace pss {

namespace {

constexpr uint32_t FLAG_SHIFT = 30;
constexpr uint32_t VALUE_MASK = (1u << FLAG_SHIFT) - 1u;
constexpr uint32_t FLAG_EMPTY = 0u;
constexpr uint32_t FLAG_LOCAL = 1u;  // value = this tile's count
constexpr uint32_t FLAG_INCL  = 2u;  // value = inclusive prefix over tiles 0..t
constexpr uint32_t FLAG_ABORT = 3u;  // a predecessor gave up (watchdog)
// Look-back watchdog: a wall-clock limit (not a poll count), so that a predecessor tile
// slowed down by time-slicing, MPS or a debugger does not turn into a spurious failure.
constexpr uint64_t WATCHDOG_NS = 30ull * 1000ull * 1000ull * 1000ull;

constexpr int CTRL_TICKET  = 0;   // [0..8)
constexpr int CTRL_TRIVIAL = 16;  // [16..24)
constexpr int CTRL_MAXBIN  = 24;  // [24..32) digit spread (expected distinct digits per warp x1000)
constexpr int CTRL_RESET   = 32;  // words [0..32) are cleared by every sort
constexpr int CTRL_ERROR   = 32;  // sticky look-back watchdog flag (cleared when it is reported)
constexpr int CTRL_WORDS   = 40;

// ------------------------------------------------------------------------------------
// Upfront histogram: one read of the keys, all digits at once.
// ------------------------------------------------------------------------------------
constexpr int HIST_THREADS = 512;
constexpr int HIST_UNROLL  = 4;

__global__ void __launch_bounds__(HIST_THREADS)
radix_hist_kernel(const uint64_t *__restrict__ keys, uint32_t n, int begin_bit, int npass,
                  uint32_t last_mask, uint32_t *__restrict__ g_hist) {
    __shared__ uint32_t s_hist[MAX_PASSES * RADIX];
    for (int i = threadIdx.x; i < npass * RADIX; i += HIST_THREADS) s_hist[i] = 0;
    __syncthreads();

    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    constexpr uint32_t WARP_SPAN = 32 * HIST_UNROLL;
    const uint32_t warps_total = gridDim.x * (HIST_THREADS / 32);
    const uint32_t warp_global = blockIdx.x * (HIST_THREADS / 32) + warp;

    HistLayout h;
    h.begin_bit = begin_bit; h.npass = npass; h.last_mask = last_mask;
    for (uint64_t base = (uint64_t)warp_global * WARP_SPAN; base < n;
         base += (uint64_t)warps_total * WARP_SPAN) {
        uint64_t k[HIST_UNROLL];
        bool     ok[HIST_UNROLL];
#pragma unroll
        for (int u = 0; u < HIST_UNROLL; ++u) {
            uint64_t i = base + u * 32 + lane;
            ok[u] = i < n;
            k[u]  = ok[u] ? ld_stream_u64(keys + i) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < HIST_UNROLL; ++u)
            hist_accumulate(s_hist, k[u], ok[u], (base + u * 32 + 32) <= n, h);
    }
    __syncthreads();
    hist_flush(s_hist, g_hist, npass, HIST_THREADS);
}

// Exclusive scan of each digit's histogram → bin_base; flag digits with one non-empty bin.
__global__ void __launch_bounds__(RADIX)
radix_scan_kernel(const uint32_t *__restrict__ g_hist, uint32_t *__restrict__ bin_base,
                  uint32_t *__restrict__ ctrl, uint32_t n) {
    __shared__ uint32_t s_warp[RADIX / 32];
    __shared__ uint32_t s_trivial, s_max;
    const int p = blockIdx.x;
    const uint32_t t = threadIdx.x, lane = lane_id(), warp = t >> 5;
    if (t == 0) { s_trivial = 0; s_max = 0; }
    __syncthreads();
    uint32_t c = g_hist[p * RADIX + t];
    if (c == n) s_trivial = 1;
    // Expected number of distinct digit values among the 32 records a warp ranks at once,
    // x1000: sum over bins of 1 - (1 - p_bin)^32.  MATCH.ANY's latency grows with it.
    const float pb = (float)c / (float)n;
    atomicAdd(&s_max, (uint32_t)(1000.f * (1.f - __powf(1.f - pb, 32.f)) + 0.5f));
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t add = 0;
    for (uint32_t w = 0; w < warp; ++w) add += s_warp[w];
    bin_base[p * RADIX + t] = add + incl - c;
    if (t == 0) {
        ctrl[CTRL_TRIVIAL + p] = s_trivial;
        ctrl[CTRL_MAXBIN + p]  = s_max;   // digit spread: picks the ranking variant
    }
}

// ------------------------------------------------------------------------------------
// One digit pass.
// ------------------------------------------------------------------------------------
template <int THREADS, int IPT>
struct PassSmem {
    static constexpr int TILE = THREADS * IPT;
    uint64_t keys[TILE];
    uint32_t vals[TILE];
    uint32_t warp_hist[THREADS / 32][RADIX];
    uint32_t bin_start[RADIX];   // exclusive scan over digits of the tile's counts
    uint32_t glob_off[RADIX];    // global offset of the digit's run minus bin_start (mod 2^32)
    uint32_t scan_warp[RADIX / 32];
    uint32_t tile;
    uint32_t abort;
};

// tile_state[tile][digit]: bits 31..30 = FLAG_*, bits 29..0 = count (LOCAL) or inclusive
// prefix over tiles 0..tile (INCL).  One word carries flag and value, so no fence is needed.
// MODE: 0 = (key, value) records, 1 = values are 0..n-1 (generated, not loaded),
//       2 = keys only (used to partition (index, rank) pairs before the rank scatter).
// BALLOT: how the lanes holding the same digit find each other.  false = the hardware
// MATCH.ANY instruction, whose latency grows with the number of distinct digits in the warp
// (fast for skewed digits such as packed text); true = 8 independent ballots, one per
// digit bit, AND-ed together — a fixed cost that wins on uniformly random digits (ranks).
template <int THREADS, int IPT, int MIN_BLOCKS, int MODE, bool BALLOT>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
onesweep_pass_kernel(const uint64_t *__restrict__ keys_in, uint64_t *__restrict__ keys_out,
                     const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ vals_out,
                     uint32_t n, int shift, uint32_t digit_mask,
                     const uint32_t *__restrict__ bin_base, uint32_t *tile_state,
                     uint32_t *ctrl, int pass_slot) {
    static_assert(THREADS >= RADIX && THREADS % 32 == 0, "one thread per digit is needed");
    using Smem = PassSmem<THREADS, IPT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &s = *reinterpret_cast<Smem *>(smem_raw);

    constexpr int TILE       = THREADS * IPT;
    constexpr int WARPS      = THREADS / 32;
    constexpr int WARP_ITEMS = 32 * IPT;
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;

    // Tiles are handed out in launch order so that every predecessor of a running tile
    // is itself running or finished: the look-back below can never wait on a tile that
    // has not been scheduled.
    if (tid == 0) {
        s.tile  = atomicAdd(&ctrl[CTRL_TICKET + pass_slot], 1u);
        s.abort = 0;
    }
#pragma unroll
    for (int i = 0; i < RADIX / 32; ++i) s.warp_hist[warp][i * 32 + lane] = 0;
    __syncthreads();
    const uint32_t tile      = s.tile;
    const uint32_t tile_base = tile * TILE;
    const uint32_t warp_base = tile_base + warp * WARP_ITEMS;
    const uint32_t valid     = min((uint32_t)TILE, n - tile_base);

    // ---- load (warp-striped: item j of lane l sits at warp_base + j*32 + l) ----------
    constexpr bool IOTA_VALS = (MODE == 1);
    constexpr bool HAS_VALS  = (MODE != 2);
    uint64_t key[IPT];
    uint32_t val[HAS_VALS ? IPT : 1];
    if (valid == TILE) {
#pragma unroll
        for (int j = 0; j < IPT; ++j) key[j] = ld_stream_u64(keys_in + warp_base + j * 32 + lane);
        if (HAS_VALS) {
#pragma unroll
            for (int j = 0; j < IPT; ++j)
                val[j] = IOTA_VALS ? (warp_base + j * 32 + lane) : ld_stream_u32(vals_in + warp_base + j * 32 + lane);
        }
    } else {
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            uint32_t i = warp_base + j * 32 + lane;
            bool ok = i < n;
            key[j] = ok ? ld_stream_u64(keys_in + i) : ~0ull;
            if (HAS_VALS) val[j] = ok ? (IOTA_VALS ? i : ld_stream_u32(vals_in + i)) : 0u;
        }
    }

    // ---- early counts: warp-private digit histogram, so the tile's counts can be
    //      published (and the look-back started) before the expensive ranking -----------
    uint32_t *wh = s.warp_hist[warp];
#pragma unroll
    for (int j = 0; j < IPT; ++j) atomicAdd(&wh[(uint32_t)(key[j] >> shift) & digit_mask], 1u);
    __syncthreads();

    // ---- per digit (threads 0..255): offsets of each warp inside the digit's bin, publish
    //      the tile count, put the first look-back loads in flight, scan over digits -------
    constexpr int LB = 4;
    uint32_t count = 0;
    uint32_t lbv[LB];
    if (tid < RADIX) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            uint32_t t        = s.warp_hist[w][tid];
            s.warp_hist[w][tid] = run;
            run += t;
        }
        count = run;
        st_volatile_u32(tile_state + (size_t)tile * RADIX + tid,
                        ((tile == 0 ? FLAG_INCL : FLAG_LOCAL) << FLAG_SHIFT) | count);
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            const int64_t q = (int64_t)tile - 1 - j;
            lbv[j] = q >= 0 ? ld_volatile_u32(tile_state + (size_t)q * RADIX + tid) : (FLAG_INCL << FLAG_SHIFT);
        }
        uint32_t incl = count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += y;
        }
        if (lane == 31) s.scan_warp[warp] = incl;
        asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 digit warps only
        uint32_t add = 0;
        for (uint32_t w = 0; w < warp; ++w) add += s.scan_warp[w];
        s.bin_start[tid] = incl - count + add;
    }
    __syncthreads();

    // ---- decoupled look-back state machine (digit threads).  lb_consume() folds the LB
    //      predecessor states loaded last into `excl` (in order, stopping at an inclusive
    //      prefix or at a predecessor that has not published yet); lb_issue() puts the next
    //      LB loads in flight.  The steps are interleaved with the ranking loop below, so the
    //      L2 round trips of the walk overlap with the ranking instead of following it. -----
    uint32_t excl     = 0;
    int64_t  lb_p     = (int64_t)tile - 1;
    bool     lb_done  = (tile == 0) || (tid >= RADIX);
    bool     lb_abort = false;
    bool     lb_progress = false;
    auto lb_consume = [&]() {
        int consumed = 0;
        bool stopped = false;
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            const uint32_t f = lbv[j] >> FLAG_SHIFT;
            if (lb_done || stopped) continue;
            if (f == FLAG_EMPTY) {
                stopped = true;                    // predecessor not published yet: poll it again
            } else if (f == FLAG_ABORT) {
                lb_abort = lb_done = true;
            } else {
                excl += lbv[j] & VALUE_MASK;
                ++consumed;
                if (f == FLAG_INCL) lb_done = true;
            }
        }
        lb_p -= consumed;
        lb_progress = consumed > 0;
    };
    auto lb_issue = [&]() {
        if (lb_done) return;
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            const int64_t q = lb_p - j;
            lbv[j] = q >= 0 ? ld_volatile_u32(tile_state + (size_t)q * RADIX + tid) : (FLAG_INCL << FLAG_SHIFT);
        }
    };

    // ---- rank and stage: match_any groups equal digits inside the warp; the group's lowest
    //      lane advances the warp's running offset; records go to shared memory in digit order
    const uint32_t lt = lanemask_lt();
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        const uint32_t d     = (uint32_t)(key[j] >> shift) & digit_mask;
        const uint32_t bs    = s.bin_start[d];
        uint32_t peers;
        if (BALLOT) {
            peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < RADIX_BITS; ++b) {
                const bool bit    = (d >> b) & 1u;
                const uint32_t vm = __ballot_sync(0xffffffffu, bit);
                peers &= bit ? vm : ~vm;
            }
        } else {
            peers = __match_any_sync(0xffffffffu, d);
        }
        const int leader     = __ffs(peers) - 1;
        uint32_t old         = 0;
        if ((int)lane == leader) {
            old   = wh[d];
            wh[d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        const uint32_t pos = bs + old + __popc(peers & lt);
        s.keys[pos] = key[j];
        if (HAS_VALS) s.vals[pos] = val[j];
        __syncwarp();
        if ((j & 3) == 3 && j + 1 < IPT && !lb_done) {   // warp-divergent only in the digit warps' walk
            lb_consume();
            lb_issue();
        }
    }

    // ---- finish the walk (blocking), publish the inclusive prefix ---------------------------
    if (tid < RADIX) {
        if (tile > 0) {
            uint32_t spins = 0;
            uint64_t t0 = 0;
            while (true) {
                lb_consume();
                if (lb_done) break;
                if (!lb_progress && (++spins & 0x3FFFu) == 0) {
                    const uint64_t now = global_timer_ns();
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > WATCHDOG_NS) { lb_abort = true; break; }
                }
                lb_issue();
            }
            const bool aborted = lb_abort;
            if (aborted) {
                st_volatile_u32(tile_state + (size_t)tile * RADIX + tid, FLAG_ABORT << FLAG_SHIFT);
                s.abort = 1;
                atomicExch(&ctrl[CTRL_ERROR], 1u);
            } else {
                st_volatile_u32(tile_state + (size_t)tile * RADIX + tid,
                                (FLAG_INCL << FLAG_SHIFT) | (excl + count));
            }
        }
        s.glob_off[tid] = bin_base[tid] + excl - s.bin_start[tid];
    }
    __syncthreads();
    if (s.abort) return;

    // ---- coalesced per-digit runs to global memory ------------------------------------
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        uint32_t i = j * THREADS + tid;
        if (i < valid) {
            uint64_t k   = s.keys[i];
            uint32_t d   = (uint32_t)(k >> shift) & digit_mask;
            uint32_t pos = s.glob_off[d] + i;
            if (pos < n) {  // always true; keeps a corrupted run (watchdog abort upstream) in bounds
                keys_out[pos] = k;
                if (HAS_VALS) vals_out[pos] = s.vals[i];
            }
        }
    }
}

// Tile geometries compiled in; PSS_PASS_CFG selects one (default 0).
struct PassConfig {
    int threads, ipt, smem;
    const void *fn[3][2];   // [MODE][BALLOT]
};
#define PSS_PASS_CONFIG(T, I, B)                                                                    \
    {T, I, (int)sizeof(PassSmem<T, I>),                                                              \
     {{(const void *)onesweep_pass_kernel<T, I, B, 0, false>, (const void *)onesweep_pass_kernel<T, I, B, 0, true>},  \
      {(const void *)onesweep_pass_kernel<T, I, B, 1, false>, (const void *)onesweep_pass_kernel<T, I, B, 1, true>},  \
      {(const void *)onesweep_pass_kernel<T, I, B, 2, false>, (const void *)onesweep_pass_kernel<T, I, B, 2, true>}}}
const PassConfig kPassConfigs[] = {
    PSS_PASS_CONFIG(256, 18, 3),   // 0 (default): 4608-record tiles, 3 CTAs/SM — best on B200 (3.75 TB/s avg)
    PSS_PASS_CONFIG(256, 16, 3),   // 1: 4096-record tiles, 3 CTAs/SM (3.45-3.55 TB/s)
    PSS_PASS_CONFIG(512, 8, 2),    // 2: 4096-record tiles, 2 CTAs/SM
    PSS_PASS_CONFIG(384, 16, 2),   // 3: 6144-record tiles, 2 CTAs/SM
    PSS_PASS_CONFIG(256, 24, 2),   // 4: 6144-record tiles, 2 CTAs/SM
    PSS_PASS_CONFIG(256, 14, 4),   // 5: 3584-record tiles, 4 CTAs/SM at 64 registers (3.09 TB/s: smaller tiles lose more
    PSS_PASS_CONFIG(256, 12, 4),   // 6: 3072-record tiles, 4 CTAs/SM (3.22 TB/s)       than the occupancy gains)
};
constexpr int kNumPassConfigs = (int)(sizeof(kPassConfigs) / sizeof(kPassConfigs[0]));

__global__ void iota_kernel(uint32_t *v, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

__global__ void copy_words_kernel(uint32_t *dst, const uint32_t *src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}

}  // namespace

int copy_words(uint32_t *dst, const uint32_t *src, int n, cudaStream_t stream) {
    copy_words_kernel<<<1, 128, 0, stream>>>(dst, src, n);
    PSS_LAUNCH_CHECK();
    return PSS_OK;
}

int alloc_mapped_words(uint32_t **p, int n) {
    PSS_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(p), (size_t)n * sizeof(uint32_t), cudaHostAllocMapped));
    std::memset(*p, 0, (size_t)n * sizeof(uint32_t));
    return PSS_OK;
}

// ------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------
int RadixSorter::init(int device) {
    if (device_ >= 0) return PSS_OK;
    device_ = device;
    PSS_CUDA_TRY(cudaSetDevice(device_));
    num_sms_ = sm_count(device_);
    PSS_CUDA_TRY(cudaMalloc(&d_hist_, MAX_PASSES * RADIX * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_bin_base_, MAX_PASSES * RADIX * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMalloc(&d_ctrl_, CTRL_WORDS * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMemset(d_ctrl_, 0, CTRL_WORDS * sizeof(uint32_t)));   // the error flag is sticky across sort_async calls
    PSS_TRY(alloc_mapped_words(&h_ctrl_, CTRL_WORDS));
    for (auto &e : ev_) PSS_CUDA_TRY(cudaEventCreate(&e));
    ev_ready_ = true;
    cfg_ = 0;
    if (const char *e = std::getenv("PSS_PASS_CFG")) {
        int v = std::atoi(e);
        if (v >= 0 && v < kNumPassConfigs) cfg_ = v;
    }
    const PassConfig &pc = kPassConfigs[cfg_];
    tile_items_ = pc.threads * pc.ipt;
    for (int m = 0; m < 3; ++m)
        for (int b = 0; b < 2; ++b)
            PSS_CUDA_TRY(cudaFuncSetAttribute(pc.fn[m][b], cudaFuncAttributeMaxDynamicSharedMemorySize, pc.smem));
    ballot_mode_ = 2;   // 0 = always MATCH.ANY, 1 = always ballots, 2 = per pass from the histogram
    if (const char *e = std::getenv("PSS_BALLOT")) ballot_mode_ = std::atoi(e);
    if (const char *e = std::getenv("PSS_SPREAD_THRESHOLD")) spread_threshold_ = (uint32_t)std::atoi(e);
    return PSS_OK;
}

int RadixSorter::ensure(int64_t n) {
    int64_t tiles = div_up(n, tile_items_);
    if (tiles < 1) tiles = 1;
    if (tiles <= tile_capacity_) return PSS_OK;
    PSS_CUDA_TRY(cudaSetDevice(device_));
    if (d_tile_state_) cudaFree(d_tile_state_);
    d_tile_state_  = nullptr;
    tile_capacity_ = 0;
    PSS_CUDA_TRY(cudaMalloc(&d_tile_state_, (size_t)tiles * RADIX * sizeof(uint32_t)));
    tile_capacity_ = tiles;
    return PSS_OK;
}

void RadixSorter::release_workspace() {
    if (device_ < 0) return;
    cudaSetDevice(device_);
    cudaFree(d_tile_state_);
    d_tile_state_  = nullptr;
    tile_capacity_ = 0;
}

void RadixSorter::release() {
    if (device_ < 0) return;
    cudaSetDevice(device_);
    cudaFree(d_hist_);
    cudaFree(d_bin_base_);
    cudaFree(d_tile_state_);
    cudaFree(d_ctrl_);
    if (h_ctrl_) cudaFreeHost(h_ctrl_);
    if (ev_ready_)
        for (auto &e : ev_) cudaEventDestroy(e);
    d_hist_ = d_bin_base_ = d_tile_state_ = d_ctrl_ = h_ctrl_ = nullptr;
    ev_ready_      = false;
    tile_capacity_ = 0;
    device_        = -1;
}

// One keys-only pass: stable partition of `keys` by the 8-bit digit at `shift` (masked).
int RadixSorter::partition(uint64_t *keys, uint64_t *keys_alt, uint32_t n, int shift, uint32_t mask,
                           cudaStream_t stream, bool *in_alt) {
    *in_alt = false;
    if (n == 0) return PSS_OK;
    if (n > VALUE_MASK) return fail(PSS_ERR_ARG, "radix partition: n must be < 2^30");
    PSS_TRY(ensure(n));
    const uint32_t tiles = (uint32_t)div_up(n, tile_items_);
    const PassConfig &pc = kPassConfigs[cfg_];
    PSS_CUDA_TRY(cudaMemsetAsync(d_hist_, 0, RADIX * sizeof(uint32_t), stream));
    PSS_CUDA_TRY(cudaMemsetAsync(d_ctrl_, 0, CTRL_RESET * sizeof(uint32_t), stream));
    PSS_CUDA_TRY(cudaMemsetAsync(d_tile_state_, 0, (size_t)tiles * RADIX * sizeof(uint32_t), stream));
    int64_t want = div_up(n, (int64_t)HIST_THREADS * HIST_UNROLL);
    int grid     = (int)std::min<int64_t>(want, (int64_t)num_sms_ * 4);
    radix_hist_kernel<<<grid, HIST_THREADS, 0, stream>>>(keys, n, shift, 1, mask, d_hist_);
    PSS_LAUNCH_CHECK();
    radix_scan_kernel<<<1, RADIX, 0, stream>>>(d_hist_, d_bin_base_, d_ctrl_, n);
    PSS_LAUNCH_CHECK();
    // A constant digit would make the pass a plain copy; it is not worth a host round trip
    // to find out, so the pass always runs.
    const uint32_t *vals_arg = nullptr, *base_arg = d_bin_base_;
    uint32_t *vout = nullptr;
    uint32_t n_arg = n, mask_arg = mask;
    int shift_arg = shift, slot_arg = 0;
    void *args[] = {&keys, &keys_alt, &vals_arg, &vout, &n_arg, &shift_arg, &mask_arg, &base_arg,
                    &d_tile_state_, &d_ctrl_, &slot_arg};
    // index windows are near-uniform: the ballot ranking is the right one
    PSS_CUDA_TRY(cudaLaunchKernel(pc.fn[2][ballot_mode_ == 0 ? 0 : 1], dim3(tiles), dim3(pc.threads), args,
                                  (size_t)pc.smem, stream));
    count_launch();
    *in_alt = true;
    return PSS_OK;
}

// Reads back the look-back watchdog flag (synchronises the stream) and clears it.
int RadixSorter::poll_error(cudaStream_t stream) {
    PSS_TRY(copy_words(h_ctrl_, d_ctrl_, CTRL_WORDS, stream));
    PSS_CUDA_TRY(cudaStreamSynchronize(stream));
    return check_error_word(h_ctrl_[CTRL_ERROR], stream);
}

const uint32_t *RadixSorter::d_error_flag() const { return d_ctrl_ + CTRL_ERROR; }

// `word`: the error flag as read back by the caller (together with its own scalars).
int RadixSorter::check_error_word(uint32_t word, cudaStream_t stream) {
    if (!word) return PSS_OK;
    PSS_CUDA_TRY(cudaMemsetAsync(d_ctrl_ + CTRL_ERROR, 0, sizeof(uint32_t), stream));
    return fail(PSS_ERR_CUDA, "radix sort: look-back watchdog fired");
}

// Per-pass CUDA-event durations of the last timed sort (its events must have completed).
int RadixSorter::collect_times(SortProfile *prof) {
    if (!prof || !prof->timed) return PSS_OK;
    for (int e = 0; e < prof->n_passes; ++e) PSS_CUDA_TRY(cudaEventElapsedTime(&prof->ms[e], ev_[2 * e], ev_[2 * e + 1]));
    PSS_CUDA_TRY(cudaEventElapsedTime(&prof->hist_ms, ev_[2 * MAX_PASSES], ev_[2 * MAX_PASSES + 1]));
    return PSS_OK;
}

int RadixSorter::sort_async(uint64_t *keys, uint64_t *keys_alt, uint32_t *vals, uint32_t *vals_alt, uint32_t n,
                            int begin_bit, int end_bit, bool iota_vals, cudaStream_t stream, bool *in_alt) {
    *in_alt = false;
    if (begin_bit < 0 || end_bit > 64 || end_bit < begin_bit) return fail(PSS_ERR_ARG, "radix sort: bad bit range");
    if (n > VALUE_MASK) return fail(PSS_ERR_ARG, "radix sort: n must be < 2^30");
    const int npass = (end_bit - begin_bit + RADIX_BITS - 1) / RADIX_BITS;
    if (npass > MAX_PASSES) return fail(PSS_ERR_ARG, "radix sort: more than 8 digits");
    if (n == 0) return PSS_OK;
    if (npass == 0) {
        if (iota_vals) {
            iota_kernel<<<(unsigned)div_up(n, 256), 256, 0, stream>>>(vals, n);
            PSS_LAUNCH_CHECK();
        }
        return PSS_OK;
    }
    PSS_TRY(ensure(n));
    const int last_bits      = (end_bit - begin_bit) - (npass - 1) * RADIX_BITS;
    const uint32_t last_mask = (1u << last_bits) - 1u;
    const uint32_t tiles     = (uint32_t)div_up(n, tile_items_);
    const PassConfig &pc     = kPassConfigs[cfg_];
    PSS_CUDA_TRY(cudaMemsetAsync(d_hist_, 0, MAX_PASSES * RADIX * sizeof(uint32_t), stream));
    PSS_CUDA_TRY(cudaMemsetAsync(d_ctrl_, 0, CTRL_RESET * sizeof(uint32_t), stream));   // not the sticky error flag
    {
        int64_t want = div_up(n, (int64_t)HIST_THREADS * HIST_UNROLL);
        int grid     = (int)std::min<int64_t>(want, (int64_t)num_sms_ * 4);
        radix_hist_kernel<<<grid, HIST_THREADS, 0, stream>>>(keys, n, begin_bit, npass, last_mask, d_hist_);
        PSS_LAUNCH_CHECK();
        radix_scan_kernel<<<npass, RADIX, 0, stream>>>(d_hist_, d_bin_base_, d_ctrl_, n);
        PSS_LAUNCH_CHECK();
    }
    uint64_t *kin = keys, *kout = keys_alt;
    uint32_t *vin = vals, *vout = vals_alt;
    bool iota = iota_vals;
    for (int p = 0; p < npass; ++p) {
        const int shift     = begin_bit + p * RADIX_BITS;
        const uint32_t mask = (p == npass - 1) ? last_mask : (uint32_t)(RADIX - 1);
        PSS_CUDA_TRY(cudaMemsetAsync(d_tile_state_, 0, (size_t)tiles * RADIX * sizeof(uint32_t), stream));
        const uint32_t *vals_arg = iota ? nullptr : vin;
        const uint32_t *base_arg = d_bin_base_ + p * RADIX;
        uint32_t n_arg = n, mask_arg = mask;
        int shift_arg = shift, slot_arg = p;
        void *args[] = {&kin, &kout, &vals_arg, &vout, &n_arg, &shift_arg, &mask_arg, &base_arg,
                        &d_tile_state_, &d_ctrl_, &slot_arg};
        PSS_CUDA_TRY(cudaLaunchKernel(pc.fn[iota ? 1 : 0][ballot_mode_ == 0 ? 0 : 1], dim3(tiles), dim3(pc.threads), args,
                                      (size_t)pc.smem, stream));
        count_launch();
        iota = false;
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
    *in_alt = (npass & 1) != 0;
    return PSS_OK;
}

int RadixSorter::hist_reset(cudaStream_t stream) {
    PSS_CUDA_TRY(cudaMemsetAsync(d_hist_, 0, MAX_PASSES * RADIX * sizeof(uint32_t), stream));
    return PSS_OK;
}

int RadixSorter::sort(uint64_t *keys, uint64_t *keys_alt, uint32_t *vals, uint32_t *vals_alt,
                      uint32_t n, int begin_bit, int end_bit, bool iota_vals, cudaStream_t stream,
                      bool *in_alt, SortProfile *prof, bool hist_done, bool defer_check) {
    *in_alt = false;
    const bool timed = prof && prof->timed;
    if (prof) { prof->n_passes = 0; prof->hist_ms = 0.f; }
    if (begin_bit < 0 || end_bit > 64 || end_bit < begin_bit) return fail(PSS_ERR_ARG, "radix sort: bad bit range");
    if (n > VALUE_MASK) return fail(PSS_ERR_ARG, "radix sort: n must be < 2^30");
    const int npass = (end_bit - begin_bit + RADIX_BITS - 1) / RADIX_BITS;
    if (npass > MAX_PASSES) return fail(PSS_ERR_ARG, "radix sort: more than 8 digits");
    if (n == 0 || npass == 0) return PSS_OK;
    PSS_TRY(ensure(n));

    const int last_bits      = (end_bit - begin_bit) - (npass - 1) * RADIX_BITS;
    const uint32_t last_mask = (1u << last_bits) - 1u;
    const uint32_t tiles     = (uint32_t)div_up(n, tile_items_);
    const PassConfig &pc     = kPassConfigs[cfg_];

    if (!hist_done) PSS_CUDA_TRY(cudaMemsetAsync(d_hist_, 0, MAX_PASSES * RADIX * sizeof(uint32_t), stream));
    PSS_CUDA_TRY(cudaMemsetAsync(d_ctrl_, 0, CTRL_RESET * sizeof(uint32_t), stream));
    if (timed) PSS_CUDA_TRY(cudaEventRecord(ev_[2 * MAX_PASSES], stream));
    {
        if (!hist_done) {
            int64_t want = div_up(n, (int64_t)HIST_THREADS * HIST_UNROLL);
            int grid     = (int)std::min<int64_t>(want, (int64_t)num_sms_ * 4);
            radix_hist_kernel<<<grid, HIST_THREADS, 0, stream>>>(keys, n, begin_bit, npass, last_mask, d_hist_);
            PSS_LAUNCH_CHECK();
        }
        radix_scan_kernel<<<npass, RADIX, 0, stream>>>(d_hist_, d_bin_base_, d_ctrl_, n);
        PSS_LAUNCH_CHECK();
    }
    if (timed) PSS_CUDA_TRY(cudaEventRecord(ev_[2 * MAX_PASSES + 1], stream));
    PSS_TRY(copy_words(h_ctrl_, d_ctrl_, CTRL_WORDS, stream));
    PSS_CUDA_TRY(cudaStreamSynchronize(stream));

    uint32_t max_bin[MAX_PASSES];
    bool     trivial[MAX_PASSES];
    for (int p = 0; p < MAX_PASSES; ++p) {
        max_bin[p] = h_ctrl_[CTRL_MAXBIN + p];
        trivial[p] = h_ctrl_[CTRL_TRIVIAL + p] != 0;
    }
    uint64_t *kin = keys, *kout = keys_alt;
    uint32_t *vin = vals, *vout = vals_alt;
    bool iota = iota_vals;  // first executed pass generates 0..n-1 instead of reading vin
    int executed = 0;
    for (int p = 0; p < npass; ++p) {
        if (trivial[p]) continue;  // every key has the same digit: order unchanged
        const int shift     = begin_bit + p * RADIX_BITS;
        const uint32_t mask = (p == npass - 1) ? last_mask : (uint32_t)(RADIX - 1);
        // the event pair brackets the pass AND the reset of its look-back words (tiles x 1 KiB)
        if (timed) PSS_CUDA_TRY(cudaEventRecord(ev_[2 * executed], stream));
        PSS_CUDA_TRY(cudaMemsetAsync(d_tile_state_, 0, (size_t)tiles * RADIX * sizeof(uint32_t), stream));
        {
            const uint32_t *vals_arg = iota ? nullptr : vin;
            const uint32_t *base_arg = d_bin_base_ + p * RADIX;
            uint32_t n_arg = n, mask_arg = mask;
            int shift_arg = shift, slot_arg = p;
            void *args[] = {&kin, &kout, &vals_arg, &vout, &n_arg, &shift_arg, &mask_arg, &base_arg,
                            &d_tile_state_, &d_ctrl_, &slot_arg};
            // many distinct digits per warp → ballots (threshold calibrated on B200)
            const bool ballot = ballot_mode_ == 1 || (ballot_mode_ == 2 && max_bin[p] > spread_threshold_);
            PSS_CUDA_TRY(cudaLaunchKernel(pc.fn[iota ? 1 : 0][ballot ? 1 : 0], dim3(tiles), dim3(pc.threads), args,
                                          (size_t)pc.smem, stream));
        }
        PSS_LAUNCH_CHECK();
        if (timed) PSS_CUDA_TRY(cudaEventRecord(ev_[2 * executed + 1], stream));
        if (prof) { prof->shift[executed] = shift; prof->spread[executed] = (int)max_bin[p]; }
        iota = false;
        std::swap(kin, kout);
        std::swap(vin, vout);
        ++executed;
    }
    *in_alt = (executed & 1) != 0;
    if (iota) {  // no pass ran: the (already ordered) values still have to exist
        iota_kernel<<<(unsigned)div_up(n, 256), 256, 0, stream>>>(vals, n);
        PSS_LAUNCH_CHECK();
    }

    if (prof) prof->n_passes = executed;
    // Watchdog flag (look-back gave up): never expected; surfaces as an error, not a hang.
    // With defer_check the caller reads d_error_flag() at its own next synchronisation
    // (check_error_word) and fetches the pass timings afterwards (collect_times): the sort
    // then costs ONE host round trip (the skip mask above).
    if (defer_check) return PSS_OK;
    PSS_TRY(copy_words(h_ctrl_, d_ctrl_, CTRL_WORDS, stream));
    PSS_CUDA_TRY(cudaStreamSynchronize(stream));
    PSS_TRY(check_error_word(h_ctrl_[CTRL_ERROR], stream));
    PSS_TRY(collect_times(prof));
    return PSS_OK;
}

}  // namespace pss
