// radix_sort.cu — onesweep LSD radix sort kernels for sm_100a (see radix_sort.cuh).
#include "radix_sort.cuh"

#include <algorithm>

namespace pss {

namespace {

constexpr uint32_t FLAG_SHIFT = 30;
constexpr uint32_t VALUE_MASK = (1u << FLAG_SHIFT) - 1u;
constexpr uint32_t FLAG_EMPTY = 0u;
constexpr uint32_t FLAG_LOCAL = 1u;  // value = this tile's count
constexpr uint32_t FLAG_INCL  = 2u;  // value = inclusive prefix over tiles 0..t
constexpr uint32_t FLAG_ABORT = 3u;  // a predecessor gave up (watchdog)
constexpr uint32_t SPIN_LIMIT = 1u << 24;

constexpr int CTRL_TICKET  = 0;   // [0..8)
constexpr int CTRL_ERROR   = 8;
constexpr int CTRL_TRIVIAL = 16;  // [16..24)
constexpr int CTRL_WORDS   = 32;

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
        for (int u = 0; u < HIST_UNROLL; ++u) {
            const bool full = (base + u * 32 + 32) <= n;  // warp-uniform
            for (int p = 0; p < npass; ++p) {
                uint32_t d = (uint32_t)(k[u] >> (begin_bit + p * RADIX_BITS)) &
                             (p == npass - 1 ? last_mask : (uint32_t)(RADIX - 1));
                if (full) {
                    // Constant digits (high bits of small ranks, padded alphabets) would
                    // serialise 32 same-address atomics; aggregate them instead.
                    uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
                    if (__all_sync(0xffffffffu, d == d0)) {
                        if (lane == 0) atomicAdd(&s_hist[p * RADIX + d0], 32u);
                    } else {
                        atomicAdd(&s_hist[p * RADIX + d], 1u);
                    }
                } else if (ok[u]) {
                    atomicAdd(&s_hist[p * RADIX + d], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * RADIX; i += HIST_THREADS) {
        uint32_t c = s_hist[i];
        if (c) atomicAdd(&g_hist[i], c);
    }
}

// Exclusive scan of each digit's histogram → bin_base; flag digits with one non-empty bin.
__global__ void __launch_bounds__(RADIX)
radix_scan_kernel(const uint32_t *__restrict__ g_hist, uint32_t *__restrict__ bin_base,
                  uint32_t *__restrict__ ctrl, uint32_t n) {
    __shared__ uint32_t s_warp[RADIX / 32];
    __shared__ uint32_t s_trivial;
    const int p = blockIdx.x;
    const uint32_t t = threadIdx.x, lane = lane_id(), warp = t >> 5;
    if (t == 0) s_trivial = 0;
    __syncthreads();
    uint32_t c = g_hist[p * RADIX + t];
    if (c == n) s_trivial = 1;
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
    if (t == 0) ctrl[CTRL_TRIVIAL + p] = s_trivial;
}

// ------------------------------------------------------------------------------------
// One digit pass.
// ------------------------------------------------------------------------------------
struct PassSmem {
    uint64_t keys[PASS_TILE];
    uint32_t vals[PASS_TILE];
    uint32_t warp_hist[PASS_THREADS / 32][RADIX];
    uint32_t bin_start[RADIX];   // exclusive scan over digits of the tile's counts
    uint32_t glob_off[RADIX];    // global offset of the digit's run minus bin_start (mod 2^32)
    uint32_t scan_warp[RADIX / 32];
    uint32_t tile;
    uint32_t abort;
};

template <bool IOTA_VALS>
__global__ void __launch_bounds__(PASS_THREADS, 2)
onesweep_pass_kernel(const uint64_t *__restrict__ keys_in, uint64_t *__restrict__ keys_out,
                     const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ vals_out,
                     uint32_t n, int shift, uint32_t digit_mask,
                     const uint32_t *__restrict__ bin_base, uint32_t *tile_state,
                     uint32_t *ctrl, int pass_slot) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PassSmem &s = *reinterpret_cast<PassSmem *>(smem_raw);

    constexpr int WARPS      = PASS_THREADS / 32;
    constexpr int WARP_ITEMS = 32 * PASS_IPT;
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
    const uint32_t tile_base = tile * PASS_TILE;
    const uint32_t warp_base = tile_base + warp * WARP_ITEMS;
    const uint32_t valid     = min((uint32_t)PASS_TILE, n - tile_base);

    // ---- load (warp-striped: item j of lane l sits at warp_base + j*32 + l) ----------
    uint64_t key[PASS_IPT];
    uint32_t val[PASS_IPT];
    if (valid == PASS_TILE) {
#pragma unroll
        for (int j = 0; j < PASS_IPT; ++j) key[j] = ld_stream_u64(keys_in + warp_base + j * 32 + lane);
#pragma unroll
        for (int j = 0; j < PASS_IPT; ++j)
            val[j] = IOTA_VALS ? (warp_base + j * 32 + lane) : ld_stream_u32(vals_in + warp_base + j * 32 + lane);
    } else {
#pragma unroll
        for (int j = 0; j < PASS_IPT; ++j) {
            uint32_t i = warp_base + j * 32 + lane;
            bool ok = i < n;
            key[j] = ok ? ld_stream_u64(keys_in + i) : ~0ull;
            val[j] = ok ? (IOTA_VALS ? i : ld_stream_u32(vals_in + i)) : 0u;
        }
    }

    // ---- rank inside the warp: match_any groups equal digits, the group's lowest lane
    //      bumps the warp's private counter, everyone derives a stable rank ------------
    uint16_t rank[PASS_IPT];
    uint32_t *wh = s.warp_hist[warp];
    const uint32_t lt = lanemask_lt();
#pragma unroll
    for (int j = 0; j < PASS_IPT; ++j) {
        uint32_t d     = (uint32_t)(key[j] >> shift) & digit_mask;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        int leader     = __ffs(peers) - 1;
        uint32_t old   = 0;
        if ((int)lane == leader) {
            old   = wh[d];
            wh[d] = old + __popc(peers);
        }
        old     = __shfl_sync(0xffffffffu, old, leader);
        rank[j] = (uint16_t)(old + __popc(peers & lt));
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit: prefix over warps, publish the tile count, scan over digits -------
    uint32_t count = 0;
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
        uint32_t incl = count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += y;
        }
        if (lane == 31) s.scan_warp[warp] = incl;
        s.bin_start[tid] = incl - count;  // warp-local exclusive; fixed up below
    }
    __syncthreads();
    if (tid < RADIX) {
        uint32_t add = 0;
        for (uint32_t w = 0; w < warp; ++w) add += s.scan_warp[w];
        uint32_t bstart  = s.bin_start[tid] + add;
        s.bin_start[tid] = bstart;

        // ---- decoupled look-back: one thread per digit walks the preceding tiles -----
        uint32_t excl = 0;
        bool aborted  = false;
        if (tile > 0) {
            int64_t p = (int64_t)tile - 1;
            while (true) {
                const uint32_t *slot = tile_state + (size_t)p * RADIX + tid;
                uint32_t v, spins = 0;
                do {
                    v = ld_volatile_u32(slot);
                } while ((v >> FLAG_SHIFT) == FLAG_EMPTY && ++spins < SPIN_LIMIT);
                uint32_t f = v >> FLAG_SHIFT;
                if (f == FLAG_EMPTY || f == FLAG_ABORT) { aborted = true; break; }
                excl += v & VALUE_MASK;
                if (f == FLAG_INCL) break;
                --p;
            }
            if (aborted) {
                st_volatile_u32(tile_state + (size_t)tile * RADIX + tid, FLAG_ABORT << FLAG_SHIFT);
                s.abort = 1;
                atomicExch(&ctrl[CTRL_ERROR], 1u);
            } else {
                st_volatile_u32(tile_state + (size_t)tile * RADIX + tid,
                                (FLAG_INCL << FLAG_SHIFT) | (excl + count));
            }
        }
        s.glob_off[tid] = bin_base[tid] + excl - bstart;
    }
    __syncthreads();
    if (s.abort) return;

    // ---- stage the tile in shared memory in digit order ------------------------------
#pragma unroll
    for (int j = 0; j < PASS_IPT; ++j) {
        uint32_t d   = (uint32_t)(key[j] >> shift) & digit_mask;
        uint32_t pos = s.bin_start[d] + wh[d] + rank[j];
        s.keys[pos]  = key[j];
        s.vals[pos]  = val[j];
    }
    __syncthreads();

    // ---- coalesced per-digit runs to global memory ------------------------------------
#pragma unroll
    for (int j = 0; j < PASS_IPT; ++j) {
        uint32_t i = j * PASS_THREADS + tid;
        if (i < valid) {
            uint64_t k   = s.keys[i];
            uint32_t d   = (uint32_t)(k >> shift) & digit_mask;
            uint32_t pos = s.glob_off[d] + i;
            if (pos < n) {  // always true; keeps a corrupted run (watchdog abort upstream) in bounds
                keys_out[pos] = k;
                vals_out[pos] = s.vals[i];
            }
        }
    }
}

__global__ void iota_kernel(uint32_t *v, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

}  // namespace

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
    PSS_CUDA_TRY(cudaMallocHost(&h_ctrl_, CTRL_WORDS * sizeof(uint32_t)));
    for (auto &e : ev_) PSS_CUDA_TRY(cudaEventCreate(&e));
    ev_ready_ = true;
    PSS_CUDA_TRY(cudaFuncSetAttribute(onesweep_pass_kernel<true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem)));
    PSS_CUDA_TRY(cudaFuncSetAttribute(onesweep_pass_kernel<false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem)));
    return PSS_OK;
}

int RadixSorter::ensure(int64_t n) {
    int64_t tiles = div_up(n, PASS_TILE);
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

int RadixSorter::sort(uint64_t *keys, uint64_t *keys_alt, uint32_t *vals, uint32_t *vals_alt,
                      uint32_t n, int begin_bit, int end_bit, bool iota_vals, cudaStream_t stream,
                      bool *in_alt, SortProfile *prof) {
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
    const uint32_t tiles     = (uint32_t)div_up(n, PASS_TILE);

    PSS_CUDA_TRY(cudaMemsetAsync(d_hist_, 0, MAX_PASSES * RADIX * sizeof(uint32_t), stream));
    PSS_CUDA_TRY(cudaMemsetAsync(d_ctrl_, 0, CTRL_WORDS * sizeof(uint32_t), stream));
    if (timed) PSS_CUDA_TRY(cudaEventRecord(ev_[2 * MAX_PASSES], stream));
    {
        int64_t want = div_up(n, (int64_t)HIST_THREADS * HIST_UNROLL);
        int grid     = (int)std::min<int64_t>(want, (int64_t)num_sms_ * 4);
        radix_hist_kernel<<<grid, HIST_THREADS, 0, stream>>>(keys, n, begin_bit, npass, last_mask, d_hist_);
        PSS_LAUNCH_CHECK();
        radix_scan_kernel<<<npass, RADIX, 0, stream>>>(d_hist_, d_bin_base_, d_ctrl_, n);
        PSS_LAUNCH_CHECK();
    }
    if (timed) PSS_CUDA_TRY(cudaEventRecord(ev_[2 * MAX_PASSES + 1], stream));
    PSS_CUDA_TRY(cudaMemcpyAsync(h_ctrl_, d_ctrl_, CTRL_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    PSS_CUDA_TRY(cudaStreamSynchronize(stream));

    uint64_t *kin = keys, *kout = keys_alt;
    uint32_t *vin = vals, *vout = vals_alt;
    bool iota = iota_vals;  // first executed pass generates 0..n-1 instead of reading vin
    int executed = 0;
    for (int p = 0; p < npass; ++p) {
        if (h_ctrl_[CTRL_TRIVIAL + p]) continue;  // every key has the same digit: order unchanged
        const int shift     = begin_bit + p * RADIX_BITS;
        const uint32_t mask = (p == npass - 1) ? last_mask : (uint32_t)(RADIX - 1);
        PSS_CUDA_TRY(cudaMemsetAsync(d_tile_state_, 0, (size_t)tiles * RADIX * sizeof(uint32_t), stream));
        if (timed) PSS_CUDA_TRY(cudaEventRecord(ev_[2 * executed], stream));
        if (iota) {
            onesweep_pass_kernel<true><<<tiles, PASS_THREADS, sizeof(PassSmem), stream>>>(
                kin, kout, nullptr, vout, n, shift, mask, d_bin_base_ + p * RADIX, d_tile_state_, d_ctrl_, p);
        } else {
            onesweep_pass_kernel<false><<<tiles, PASS_THREADS, sizeof(PassSmem), stream>>>(
                kin, kout, vin, vout, n, shift, mask, d_bin_base_ + p * RADIX, d_tile_state_, d_ctrl_, p);
        }
        PSS_LAUNCH_CHECK();
        if (timed) PSS_CUDA_TRY(cudaEventRecord(ev_[2 * executed + 1], stream));
        if (prof) prof->shift[executed] = shift;
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

    // Watchdog flag (look-back gave up): never expected; surfaces as an error, not a hang.
    PSS_CUDA_TRY(cudaMemcpyAsync(h_ctrl_, d_ctrl_, CTRL_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    PSS_CUDA_TRY(cudaStreamSynchronize(stream));
    if (h_ctrl_[CTRL_ERROR]) return fail(PSS_ERR_CUDA, "radix sort: look-back watchdog fired");
    if (prof) prof->n_passes = executed;
    if (timed) {
        for (int e = 0; e < executed; ++e) PSS_CUDA_TRY(cudaEventElapsedTime(&prof->ms[e], ev_[2 * e], ev_[2 * e + 1]));
        PSS_CUDA_TRY(cudaEventElapsedTime(&prof->hist_ms, ev_[2 * MAX_PASSES], ev_[2 * MAX_PASSES + 1]));
    }
    return PSS_OK;
}

}  // namespace pss
