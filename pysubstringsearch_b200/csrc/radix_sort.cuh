// radix_sort.cuh — hand-written onesweep LSD radix sort over (key u64, value u32) records.
//
// One upfront histogram kernel computes the 256-bin histograms of every 8-bit digit in
// [begin_bit, end_bit); each digit then costs ONE pass over the data: a tile (CTA) ranks
// its records inside each warp (ballots or MATCH.ANY), publishes its per-digit counts, resolves its
// global offsets by decoupled look-back over the preceding tiles, stages the records in
// shared memory in digit order and writes them out in coalesced per-digit runs.
// Digits whose histogram has a single non-empty bin are skipped on the host.
//
// This replaces (as the inner engine of a prefix-doubling builder) the serial induced
// sorting scans of the reference's libsais (src/libsais/libsais.c:2105, :2936, :4565,
// :5194); it is not a translation of them.
#pragma once

#include "common.cuh"

namespace pss {

constexpr int RADIX_BITS = 8;
constexpr int RADIX      = 1 << RADIX_BITS;
constexpr int MAX_PASSES = 8;


// Digit layout of a sort over key bits [begin_bit, end_bit): what a kernel that produces
// the keys needs in order to accumulate the upfront histograms itself (hist_accumulate).
struct HistLayout {
    int      begin_bit = 0, npass = 0;
    uint32_t last_mask = 0;
};
inline HistLayout hist_layout(int begin_bit, int end_bit) {
    HistLayout h;
    h.begin_bit = begin_bit;
    h.npass     = (end_bit - begin_bit + RADIX_BITS - 1) / RADIX_BITS;
    const int last_bits = (end_bit - begin_bit) - (h.npass - 1) * RADIX_BITS;
    h.last_mask = (1u << last_bits) - 1u;
    return h;
}

#ifdef __CUDACC__
// Adds one key per lane to the block's shared-memory histograms s_hist[npass][RADIX].
// `warp_full` (warp-uniform) says all 32 lanes hold a valid key: only then can a digit that
// is identical across the warp be added with one atomic instead of 32 same-address ones.
__device__ __forceinline__ void hist_accumulate(uint32_t *s_hist, uint64_t key, bool valid, bool warp_full,
                                                const HistLayout &h) {
    const uint32_t lane = threadIdx.x & 31u;
    for (int p = 0; p < h.npass; ++p) {
        const uint32_t d = (uint32_t)(key >> (h.begin_bit + p * RADIX_BITS)) &
                           (p == h.npass - 1 ? h.last_mask : (uint32_t)(RADIX - 1));
        if (warp_full) {
            const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
            if (__all_sync(0xffffffffu, d == d0)) {
                if (lane == 0) atomicAdd(&s_hist[p * RADIX + d0], 32u);
            } else {
                atomicAdd(&s_hist[p * RADIX + d], 1u);
            }
        } else if (valid) {
            atomicAdd(&s_hist[p * RADIX + d], 1u);
        }
    }
}
__device__ __forceinline__ void hist_flush(const uint32_t *s_hist, uint32_t *g_hist, int npass, int threads) {
    for (int i = threadIdx.x; i < npass * RADIX; i += threads) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&g_hist[i], c);
    }
}
#endif

struct SortProfile {
    bool  timed    = false;  // in: record CUDA events around the histogram and every pass
    int   n_passes = 0;      // out: passes executed (constant digits are skipped)
    int   shift[MAX_PASSES] = {};
    int   spread[MAX_PASSES] = {};  // expected distinct digits per warp x1000 (chooses the ranking variant)
    float ms[MAX_PASSES]    = {};   // valid when timed
    float hist_ms = 0.f;            // valid when timed
};

class RadixSorter {
public:
    RadixSorter() = default;
    ~RadixSorter() { release(); }
    RadixSorter(const RadixSorter &) = delete;
    RadixSorter &operator=(const RadixSorter &) = delete;

    int  init(int device);
    int  ensure(int64_t n);  // workspace for up to n records
    void release();
    void release_workspace();   // frees the tile-state array only (regrows on demand)

    // Sorts records by key bits [begin_bit, end_bit), stable.  iota_vals means the input
    // values are 0..n-1 and `vals` is only scratch (they are generated on the fly in the
    // first executed pass).  *in_alt tells where the result is.  prof (optional) receives
    // the executed pass list and, if prof->timed, per-pass CUDA-event durations.
    // Synchronises the stream before returning.
    int sort(uint64_t *keys, uint64_t *keys_alt, uint32_t *vals, uint32_t *vals_alt,
             uint32_t n, int begin_bit, int end_bit, bool iota_vals, cudaStream_t stream,
             bool *in_alt, SortProfile *prof, bool hist_done = false, bool defer_check = false);
    int check_error_word(uint32_t word, cudaStream_t stream);
    int collect_times(SortProfile *prof);

    // For key-producing kernels that fill the histograms themselves: zero them (hist_reset),
    // accumulate into d_hist() with hist_accumulate/hist_flush, then sort(..., hist_done = true).
    int       hist_reset(cudaStream_t stream);
    uint32_t *d_hist() const { return d_hist_; }
    int       num_sms() const { return num_sms_; }

    // Stable partition of keys by one 8-bit digit ((key >> shift) & mask), keys only; the
    // result is in keys_alt (*in_alt = true).  Does not synchronise; the look-back
    // watchdog flag of this pass is read by poll_error().
    int partition(uint64_t *keys, uint64_t *keys_alt, uint32_t n, int shift, uint32_t mask,
                  cudaStream_t stream, bool *in_alt);

    // Same sort with no host round trip: every digit's pass runs (a constant digit costs a
    // stable copy instead of being skipped) and lanes are matched with ballots.  Nothing is
    // synchronised; *in_alt is known up front (odd number of digits).  The look-back
    // watchdog flag is sticky until poll_error() / error_flag() is consulted.
    int sort_async(uint64_t *keys, uint64_t *keys_alt, uint32_t *vals, uint32_t *vals_alt,
                   uint32_t n, int begin_bit, int end_bit, bool iota_vals, cudaStream_t stream, bool *in_alt);

    int poll_error(cudaStream_t stream);
    const uint32_t *d_error_flag() const;   // device word, non-zero once a watchdog fired

    int device() const { return device_; }

private:
    int       device_        = -1;
    int       num_sms_       = 0;
    int       cfg_           = 0;     // index into the compiled tile geometries (PSS_PASS_CFG)
    int       tile_items_    = 4608;  // records per tile of the selected geometry
    int       ballot_mode_   = 2;     // ranking variant: 0 MATCH.ANY, 1 ballots, 2 chosen per pass (PSS_BALLOT)
    uint32_t  spread_threshold_ = 8000;   // MATCH.ANY only when a warp sees at most ~8 distinct digits
    int64_t   tile_capacity_ = 0;
    uint32_t *d_hist_        = nullptr;  // [MAX_PASSES][RADIX]
    uint32_t *d_bin_base_    = nullptr;  // [MAX_PASSES][RADIX]
    uint32_t *d_tile_state_  = nullptr;  // [tile_capacity][RADIX]
    uint32_t *d_ctrl_        = nullptr;  // [0..8) tickets, [8] error flag, [16..24) trivial flags
    uint32_t *h_ctrl_        = nullptr;  // pinned mirror of d_ctrl_
    cudaEvent_t ev_[2 * MAX_PASSES + 2] = {};
    bool      ev_ready_      = false;
};

}  // namespace pss
