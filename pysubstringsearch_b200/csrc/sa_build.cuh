// sa_build.cuh — GPU prefix-doubling suffix-array builder (BUILD hot path).
//
// Replaces the reference's per-chunk `libsais(T, SA, n, 0, NULL)` call
// (src/lib.rs:24-40 → src/libsais/libsais.c:6597) with a B200-native design:
//
//   round 0   pack the first h0 symbols of every suffix into one u64 key (dense alphabet
//             codes, 0 = past the end, so a proper prefix sorts first without a sentinel),
//             onesweep-sort (key, index), derive group ranks
//   round r   for the suffixes still sharing a group: key = (group rank, rank of suffix
//             i + h); onesweep-sort; single-pass segmented re-rank; settled suffixes are
//             written to SA and leave the active set (Larsson–Sadakane style filtering);
//             ranks are scattered into ISA window by window after one partition pass;
//             h doubles
//
// The suffix array of a text is unique, so the result is byte-identical to libsais'.
#pragma once

#include <vector>

#include "common.cuh"
#include "radix_sort.cuh"

namespace pss {

class SaBuilder {
public:
    SaBuilder() = default;
    ~SaBuilder() { release(); }
    SaBuilder(const SaBuilder &) = delete;
    SaBuilder &operator=(const SaBuilder &) = delete;

    int  init(int device, int64_t max_n);
    void release();

    int build_device(const uint8_t *d_text, int32_t n, int32_t *d_sa, cudaStream_t stream);
    int build_host(const uint8_t *h_text, int32_t n, int32_t *h_sa);

    void set_profiling(bool on) { profiling_ = on; }
    const pss_build_stats &stats() const { return stats_; }
    const std::vector<pss_pass_stat> &pass_stats() const { return pass_stats_; }
    int device() const { return device_; }
    cudaStream_t stream() const { return stream_; }

    // Device staging buffers for host-side callers (Writer): text and SA of the chunk.
    int ensure_io(int64_t n);
    uint8_t *d_text_io() const { return d_text_; }
    int32_t *d_sa_io() const { return d_sa_; }

private:
    int ensure(int64_t n);
    int staged_copy(void *dst, const void *src, size_t bytes, bool to_device);

    int          device_   = -1;
    cudaStream_t stream_   = nullptr;
    RadixSorter  sorter_;
    int64_t      cap_      = 0;   // records the workspace can hold
    int64_t      io_cap_   = 0;
    uint64_t    *keys_a_   = nullptr, *keys_b_ = nullptr;
    uint32_t    *vals_a_   = nullptr, *vals_b_ = nullptr;
    uint32_t    *grp_      = nullptr;
    uint32_t    *isa_      = nullptr;
    uint32_t    *tile_aggr_ = nullptr;   // re-rank look-back state: two u64 words per 2048-record tile
    uint32_t    *d_small_  = nullptr;    // presence[8] | scalars[8] | code LUT (256 x u16)
    uint32_t    *h_small_  = nullptr;    // pinned mirror
    uint8_t     *d_text_   = nullptr;
    int32_t     *d_sa_     = nullptr;
    cudaEvent_t  ev_begin_ = nullptr, ev_end_ = nullptr;
    void        *stage_[2]    = {nullptr, nullptr};   // pinned bounce slices for pageable callers
    cudaEvent_t  stage_ev_[2] = {nullptr, nullptr};
    bool         profiling_ = false;
    pss_build_stats stats_ = {};
    std::vector<pss_pass_stat> pass_stats_;
};

}  // namespace pss
