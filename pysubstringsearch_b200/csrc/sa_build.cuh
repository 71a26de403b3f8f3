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

// Host <-> device copies for callers with ordinary (pageable) memory.  A Rust/C host calls
// pss_libsais with plain heap buffers; cudaMemcpy from/to pageable memory runs at a fraction
// of the PCIe rate (one driver thread stages through a small pinned buffer), and the 4n-byte
// suffix array is four times the text.  The stager bounces through two pinned slices itself
// and moves the slices with several CPU threads while the DMA of the next one is in flight.
// Pinned callers' buffers are copied directly.
bool host_is_pinned(const void *p);   // page-locked (cudaMallocHost / cudaHostRegister) or managed memory

class HostStager {
public:
    HostStager() = default;
    ~HostStager() { release(); }
    HostStager(const HostStager &) = delete;
    HostStager &operator=(const HostStager &) = delete;
    // to_device: returns once the host buffer has been consumed; completion on the device is
    // ordered on `stream`.  from device: complete (host buffer filled) on return.
    int  copy(void *dst, const void *src, size_t bytes, bool to_device, cudaStream_t stream);
    void release();

private:
    void       *stage_[2]   = {nullptr, nullptr};
    cudaEvent_t ev_[2]      = {nullptr, nullptr};
    bool        pending_[2] = {false, false};
};

class SaBuilder {
public:
    SaBuilder() = default;
    ~SaBuilder() { release(); }
    SaBuilder(const SaBuilder &) = delete;
    SaBuilder &operator=(const SaBuilder &) = delete;

    int  init(int device, int64_t max_n);
    void release();
    void release_workspace();   // frees the sort workspace and I/O buffers (they regrow on demand)

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
    HostStager   stager_;                              // pinned bounce slices for pageable callers
    bool         profiling_ = false;
    // CTAs per SM of the rank gather (PSS_GATHER_CTAS).  More records in flight than ~600 k and
    // the 128-byte fills of the random ISA reads evict the index window from L2 before it is
    // reused: at 8 CTAs/SM the round-1 gather of a 2^29-byte chunk read 38.7 GB from DRAM in
    // 7.35 ms, at 4 it reads 12.3 GB in 4.60 ms, at 2 the ideal 5.4 GB but latency-bound (6.65 ms).
    int          gather_ctas_ = 4;
    int          l2_hints_ = 1;           // PSS_L2_HINTS=0 disables the evict_last / evict_first cache hints
    pss_build_stats stats_ = {};
    std::vector<pss_pass_stat> pass_stats_;
};

}  // namespace pss
