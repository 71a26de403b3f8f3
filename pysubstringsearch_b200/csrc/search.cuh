// search.cuh — batched GPU substring search over per-chunk suffix arrays (SEARCH hot path).
//
// Replaces the per-query, per-chunk closure of the reference's Reader::search
// (src/lib.rs:208-281) and the Python loop of search_multiple (__init__.py:61-73) with
// one batched pipeline over all (query, chunk) pairs:
//
//   bounds   one warp per (query, chunk): cooperative lower/upper bound over the SA,
//            pattern bytes in registers, text compared in coalesced 32-byte windows
//            (lib.rs:212-252 computes the same range with two sequential binary searches)
//   extract  one thread per matching suffix: entry start = 1 + previous '\n', entry end =
//            next '\n' (lib.rs:266-273), found with 4-byte SIMD compares
//   dedup    stable onesweep sort of (pair id, entry start) → the first record of every
//            run is the entry's first hit in SA order (lib.rs:262,274 uses a hash set);
//            survivors are compacted back in SA order, which is the reference's order
#pragma once

#include <vector>

#include "common.cuh"
#include "radix_sort.cuh"

namespace pss {

struct DeviceChunk {
    const uint8_t *text;   // device, zero-padded to a multiple of 16 bytes past n
    const int32_t *sa;     // device
    uint32_t       n;
    int32_t        global_id;
};

struct SearchTimes {
    float ms_bounds = 0.f, ms_extract = 0.f, ms_dedup = 0.f, ms_total = 0.f;
};

// Receives the entries of one sub-batch, already in final order, while they are still
// on the device.
struct SearchSink {
    virtual ~SearchSink() {}
    // Called before the compaction kernel of a sub-batch: `count` entries are about to be
    // produced; return the device pointers they must be written to (any may be nullptr).
    virtual int reserve(int64_t count, int32_t **d_query, int32_t **d_chunk, uint32_t **d_start,
                        uint32_t **d_end) = 0;
    // Called after the compaction kernel has been enqueued on `stream`.
    virtual int commit(int64_t count, cudaStream_t stream) = 0;
    // Small-batch path: the entries are already in (pinned) host memory.  The default
    // forwards them through reserve()/commit(); host sinks override it with a plain copy.
    virtual int deliver_host(int64_t count, const int32_t *query, const int32_t *chunk, const uint32_t *start,
                             const uint32_t *end, cudaStream_t stream);
};

class Searcher {
public:
    Searcher() = default;
    ~Searcher() { release(); }
    Searcher(const Searcher &) = delete;
    Searcher &operator=(const Searcher &) = delete;

    int  init(int device);
    void release();
    int  set_chunks(const std::vector<DeviceChunk> &chunks);
    int  num_chunks() const { return (int)chunks_.size(); }
    cudaStream_t stream() const { return stream_; }
    int  device() const { return device_; }

    // d_patterns / d_offsets: device.  Entries are delivered to `sink` sub-batch by
    // sub-batch in (query, chunk, SA order) order.  per_pair_count (host, nq * num_chunks,
    // may be nullptr) receives the entries produced by each (query, chunk) pair.
    int search(const uint8_t *d_patterns, const int64_t *d_offsets, int32_t nq, cudaStream_t stream,
               SearchSink *sink, int64_t *per_pair_count, int64_t *n_hits, SearchTimes *times);

private:
    int ensure_pairs(int64_t npairs);
    int ensure_hits(int64_t nhits);

    int          device_ = -1;
    cudaStream_t stream_ = nullptr;
    RadixSorter  sorter_;
    std::vector<DeviceChunk> chunks_;
    DeviceChunk *d_chunks_ = nullptr;

    int64_t   pair_cap_ = 0;
    uint32_t *d_lb_ = nullptr, *d_cnt_ = nullptr, *d_hit_off_ = nullptr, *d_pair_first_ = nullptr;
    uint32_t *h_lb_ = nullptr, *h_cnt_ = nullptr, *h_hit_off_ = nullptr, *h_pair_first_ = nullptr;  // pinned

    int64_t   hit_cap_ = 0;
    uint64_t *d_keys_ = nullptr, *d_keys_alt_ = nullptr;
    uint32_t *d_vals_ = nullptr, *d_vals_alt_ = nullptr;
    uint32_t *d_end_ = nullptr, *d_flag_ = nullptr, *d_tile_sum_ = nullptr;
    uint32_t *d_scalar_ = nullptr, *h_scalar_ = nullptr;
    unsigned char *d_small_out_ = nullptr, *h_small_out_ = nullptr;   // small-batch path result block
    bool      small_path_ = true;
    cudaEvent_t ev_[8] = {};
};

}  // namespace pss
