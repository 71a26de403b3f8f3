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
//            next '\n' (lib.rs:266-273): a short SIMD scan around the hit, and beyond
//            SCAN_LIMIT bytes one binary search in the chunk's newline side index
//            (sorted '\n' offsets, derived on the device when the chunk is loaded), so the
//            cost per hit is bounded whatever the line length
//   dedup    the first hit in SA order stands for an entry (lib.rs:262,274 uses a hash set).
//            Three tiers by the pair's number of matching suffixes: up to 256 — one warp,
//            all-pairs compare in shared memory; up to 8192 — one CTA, a hash set of entry
//            starts in shared memory that keeps the smallest hit index; above — a stable
//            onesweep sort of (pair id, entry start) whose run heads are the first hits.
//            Survivors are compacted back in SA order, which is the reference's order
//
// Counts, offsets and the per-(query, chunk) entry offsets stay on the device; the host
// reads two scalars per batch (matching suffixes, entries).  A batch of a few pairs
// (a single Reader.search) is answered by ONE launch: every pair's bounds by a CTA-wide
// 512-ary search (4 dependent probes instead of 58), the last CTA to finish extracts,
// dedups and writes the tuples straight into mapped pinned host memory.
#pragma once

#include <vector>

#include "common.cuh"
#include "radix_sort.cuh"

namespace pss {

struct DeviceChunk {
    const uint8_t  *text;      // device, 16 readable zero bytes past n
    const int32_t  *sa;        // device
    const uint32_t *nl;        // device: sorted offsets of every '\n' (may be null: unbounded scans)
    const uint32_t *bucket;    // device: [65537] first SA slot of every 2-byte prefix (may be null: full-range search)
    const uint32_t *dir;       // device: [ceil(n / LINE_BLOCK) + 1] newlines before every text block (may be null: text scans)
    const uint4    *rec;       // device: [ceil(n / LINE_REC_BLOCK)][2] line records, one sector per text block (may be null: dir / scans)
    uint32_t        n;
    uint32_t        n_lines;   // entries of nl
    int32_t         global_id;
    int32_t         reserved;
};

// Text bytes per entry of the line directory: 256 → 1/64 of the text's size, ~6 newlines of a
// 45-byte-line corpus per block, so the nine offsets extraction loads almost always decide.
constexpr uint32_t LINE_BLOCK = 256;
// Text bytes per line record (a 32-byte sector: entry start before the block, next newline after
// it, 128-bit newline map): a quarter of the text's size, and extraction is one load per hit.
constexpr uint32_t LINE_REC_BLOCK = 128;

struct SearchTimes {
    float ms_bounds = 0.f, ms_extract = 0.f, ms_dedup = 0.f, ms_total = 0.f;
};

// Result of one batch, resident on the searcher's device and owned by the searcher
// (valid until its next search).  Entries are in (query, local chunk, SA order of first
// hit) order; pair p = query * num_chunks + local chunk.
struct SearchOutput {
    int64_t         n_entries = 0, n_hits = 0;
    const uint32_t *d_entry_off = nullptr;   // [npairs + 1] entries before pair p
    const int32_t  *d_chunk = nullptr;       // [n_entries] global chunk id
    const uint32_t *d_start = nullptr;       // [n_entries] entry start (line_tail, lib.rs:270-273)
    const uint32_t *d_end = nullptr;         // [n_entries] offset of the terminating '\n' (line_head, lib.rs:266-269)
    const int64_t  *d_query_off = nullptr;   // [nq + 1] entries before query q
    bool            on_host = false;         // small path: the five arrays are (mapped pinned) HOST pointers
    bool            deferred = false;        // compaction not run yet (search(..., defer_compact)): d_chunk/start/end are null
};

// Patterns of a small batch, passed to the fused kernel by value (no H2D copy).
constexpr int SMALL_MAX_PAIRS   = 64;
constexpr int SMALL_MAX_QUERIES = 64;
constexpr int SMALL_PAT_BYTES   = 1024;
struct SmallPatterns {
    uint32_t nq;
    uint32_t off[SMALL_MAX_QUERIES + 1];
    uint8_t  bytes[SMALL_PAT_BYTES];
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
    const std::vector<DeviceChunk> &chunks() const { return chunks_; }
    cudaStream_t stream() const { return stream_; }
    int  device() const { return device_; }

    // Builds the newline side index of a chunk already resident on this device: *d_nl
    // (cudaMalloc'ed here, owned by the caller) and *n_lines.
    int build_newline_index(const uint8_t *d_text, uint32_t n, uint32_t **d_nl, uint32_t *n_lines);
    // Builds the line directory over a newline index: (*d_dir)[j] = newlines before text byte
    // j * LINE_BLOCK (cudaMalloc'ed here, owned by the caller).
    int build_line_directory(const uint32_t *d_nl, uint32_t n_lines, uint32_t n, uint32_t **d_dir);
    // Builds the line records over a newline index (cudaMalloc'ed here, owned by the caller).
    int build_line_records(const uint32_t *d_nl, uint32_t n_lines, uint32_t n, uint4 **d_rec);

    // Builds the 2-byte prefix table of a resident chunk: (*d_bucket)[a << 8 | b] = first SA slot
    // whose suffix starts with bytes a, b (entry 65536 = n).  A pattern of two or more bytes is
    // then searched inside its bucket only: ~10 of the 29 probe levels of a 2^29-byte chunk
    // disappear from both binary searches.  cudaMalloc'ed here, owned by the caller.
    int build_prefix_buckets(const uint8_t *d_text, const int32_t *d_sa, uint32_t n, uint32_t **d_bucket);

    // d_patterns / d_offsets: device.  Synchronises `stream` before returning (the counts
    // in *out are host values).
    //
    // defer_compact: stop before the compaction — the per-pair entry offsets and the entry
    // count are final, the tuples are not written yet; compact_deferred() then writes them
    // wherever the caller wants them (another GPU's memory included).  Ignored (normal
    // compaction, out->deferred = false) when the batch needs more than one sub-batch.
    int search(const uint8_t *d_patterns, const int64_t *d_offsets, int32_t nq, cudaStream_t stream,
               SearchOutput *out, SearchTimes *times, bool defer_compact = false);

    // Compaction of the last deferred search: entry i of local pair p (i counted inside the
    // pair, SA order) is written to index d_pair_dst[p] + i of the three output arrays, which
    // may live in peer memory (kept entries are staged in shared memory and leave the SM in
    // contiguous runs).  d_chunk may be null.  Asynchronous on `stream`.
    int compact_deferred(const uint32_t *d_pair_dst, int32_t *d_chunk, uint32_t *d_start, uint32_t *d_end,
                         cudaStream_t stream);

    // Single-launch path for host-side callers: patterns (host) are passed by value.
    // *handled = false when the batch does not qualify (too many pairs / pattern bytes /
    // matching suffixes); nothing has been produced then and search() must be used.
    int search_small(const uint8_t *h_patterns, const int64_t *h_offsets, int32_t nq, SearchOutput *out,
                     SearchTimes *times, bool *handled);

private:
    int ensure_pairs(int64_t npairs, int64_t nq);
    int ensure_hits(int64_t nhits);
    int ensure_heavy(int64_t n);
    int ensure_out(int64_t entries, int64_t keep, cudaStream_t s);

    int          device_ = -1;
    cudaStream_t stream_ = nullptr;
    RadixSorter  sorter_;
    std::vector<DeviceChunk> chunks_;
    DeviceChunk *d_chunks_ = nullptr;

    int64_t   pair_cap_ = 0, query_cap_ = 0;
    uint32_t *d_lb_ = nullptr, *d_cnt_ = nullptr, *d_hit_off_ = nullptr, *d_pair_first_ = nullptr;
    uint32_t *d_entry_off_ = nullptr, *d_heavy_off_ = nullptr, *d_med_list_ = nullptr, *d_big_list_ = nullptr;
    void     *d_scan_part_ = nullptr;   // per-tile partials of the per-pair scans (hit offsets, entry offsets)
    int64_t  *d_query_off_ = nullptr;
    // pinned; only the oversized-batch path uses them
    uint32_t *h_cnt_ = nullptr, *h_hit_off_ = nullptr, *h_heavy_off_ = nullptr, *h_med_list_ = nullptr;

    int64_t   hit_cap_ = 0;                              // per matching suffix: entry start / end, keep flag
    uint32_t *d_start_ = nullptr, *d_end_ = nullptr, *d_flag_ = nullptr, *d_tile_sum_ = nullptr;
    int64_t   heavy_cap_ = 0;                            // sort buffers for the hits of heavy pairs only
    uint64_t *d_keys_ = nullptr, *d_keys_alt_ = nullptr;
    uint32_t *d_vals_ = nullptr, *d_vals_alt_ = nullptr;

    int64_t   out_cap_ = 0;
    int32_t  *d_out_chunk_ = nullptr;
    uint32_t *d_out_start_ = nullptr, *d_out_end_ = nullptr;

    // scalars: [2] entries of the current sub-batch, [3] small-path ticket, [4] newline count,
    // [5] pairs queued for the large hash tables, [6] their cursor,
    // [10..11] total matching suffixes (u64), [12..13] those of heavy pairs (u64), [14..15] medium pairs (u64)
    uint32_t *d_scalar_ = nullptr, *h_scalar_ = nullptr;
    unsigned char *h_small_out_ = nullptr;   // mapped pinned result block of the small path
    uint32_t  small_seq_ = 0;
    bool      small_path_ = true;
    int       bounds_group_ = 0;     // lanes per (query, chunk) pair of the batched bounds kernel (PSS_BOUNDS_GROUP): 0 = by
                                     // batch size, 32 / 8 / 4 with the SA look-ahead, -32 / -8 / -4 without
    cudaEvent_t ev_[8] = {};
    struct Deferred {
        bool     valid = false;
        uint32_t npairs = 0, nhits = 0, tiles = 0;
        int      nc = 0;
    } deferred_;
};

}  // namespace pss
