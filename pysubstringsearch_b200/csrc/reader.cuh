// reader.cuh — internals of pss_reader shared by host_index.cu (open / single-process search)
// and dist.cu (the NCCL exchange of the one-process-per-GPU search).
#pragma once

#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "search.cuh"

namespace pss {

// Pinned bounce buffer (grow-only).
struct Pinned {
    void  *p   = nullptr;
    size_t cap = 0;
    ~Pinned() { if (p) cudaFreeHost(p); }
    int ensure(size_t n) {
        if (n <= cap) return PSS_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        PSS_CUDA_TRY(cudaMallocHost(&p, n));
        cap = n;
        return PSS_OK;
    }
};

// Device buffer (grow-only, contents not preserved).
struct DeviceBuf {
    void  *p   = nullptr;
    size_t cap = 0;
    ~DeviceBuf() { cudaFree(p); }
    int ensure(size_t n) {
        if (n <= cap) return PSS_OK;
        cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = std::max<size_t>(n + n / 4, 1 << 16);
        PSS_CUDA_TRY(cudaMalloc(&p, want));
        cap = want;
        return PSS_OK;
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

// Pool of pinned host blocks that back pss_result arrays: the device→host copy of a batch
// lands directly in the memory the caller reads, and the blocks are recycled across batches
// (pinning a fresh block per batch costs more than the whole search).
struct PinnedBlock {
    void  *p   = nullptr;
    size_t cap = 0;
};
struct PinnedPool {
    std::mutex mu;
    std::vector<PinnedBlock> free_;
    ~PinnedPool() {
        for (auto &b : free_) cudaFreeHost(b.p);
    }
    int acquire(size_t bytes, PinnedBlock *out) {
        {
            std::lock_guard<std::mutex> lock(mu);
            size_t best = free_.size();
            for (size_t i = 0; i < free_.size(); ++i)
                if (free_[i].cap >= bytes && (best == free_.size() || free_[i].cap < free_[best].cap)) best = i;
            if (best != free_.size()) {
                *out = free_[best];
                free_.erase(free_.begin() + best);
                return PSS_OK;
            }
        }
        size_t cap = 1 << 16;
        while (cap < bytes) cap *= 2;
        PinnedBlock b;
        PSS_CUDA_TRY(cudaMallocHost(&b.p, cap));
        b.cap = cap;
        *out  = b;
        return PSS_OK;
    }
    void release(PinnedBlock b) {
        if (!b.p) return;
        std::lock_guard<std::mutex> lock(mu);
        if (free_.size() < 4) free_.push_back(b);
        else cudaFreeHost(b.p);
    }
};

// One batch of results in ONE pinned block:
//   [query_offsets i64 x (nq+1)] [chunk i32 x cap] [start u32 x cap] [end u32 x cap]
struct ResultOwner {
    pss_result pub;                      // first member: pss_result* == ResultOwner*
    std::shared_ptr<PinnedPool> pool;
    PinnedBlock blk;
    int64_t nq = 0, cap = 0;
    ~ResultOwner() { if (pool) pool->release(blk); }
    int64_t  *query_off() const { return static_cast<int64_t *>(blk.p); }
    int32_t  *chunk() const { return reinterpret_cast<int32_t *>(query_off() + nq + 1); }
    uint32_t *start() const { return reinterpret_cast<uint32_t *>(chunk() + cap); }
    uint32_t *end() const { return start() + cap; }
    int alloc(int64_t n_queries, int64_t entries) {
        nq = n_queries;
        const size_t bytes = ((size_t)nq + 1) * 8 + (size_t)std::max<int64_t>(entries, 1) * 12;
        PSS_TRY(pool->acquire(bytes, &blk));
        cap = (int64_t)((blk.cap - ((size_t)nq + 1) * 8) / 12);
        return PSS_OK;
    }
    void publish(int64_t entries) {
        pub.n_queries     = (int32_t)nq;
        pub.n_entries     = entries;
        pub.query_offsets = query_off();
        pub.chunk_id      = entries ? chunk() : nullptr;
        pub.line_start    = entries ? start() : nullptr;
        pub.line_end      = entries ? end() : nullptr;
    }
};

struct ChunkHost {
    uint64_t file_text_off = 0, file_sa_off = 0;
    uint32_t n = 0;
    uint64_t sa_bytes = 0;
    bool     owned = false;
    bool     borrowed = false;             // device pointers belong to the caller (device-chunk reader)
    std::vector<uint8_t> text;             // host copy (result materialisation), owned chunks only
    const uint8_t *h_text = nullptr;       // text.data(), or the caller's host copy (may be null)
    uint8_t  *d_text = nullptr;
    int32_t  *d_sa   = nullptr;
    uint32_t *d_nl   = nullptr;            // newline side index (always ours)
    uint32_t *d_bucket = nullptr;          // 2-byte prefix table over the SA (always ours)
    uint32_t *d_dir  = nullptr;            // line directory over d_nl (always ours; PSS_LINE_DIR=1 only)
    uint4    *d_rec  = nullptr;            // line records (always ours)
    uint32_t  n_lines = 0;
};

}  // namespace pss

struct pss_reader {
    std::string path;
    std::vector<pss::ChunkHost> chunks;    // every chunk of the container (or of the whole sharded index)
    pss::Searcher searcher;
    int shard_rank = 0, shard_count = 1;
    // pattern staging
    pss::Pinned   h_pat;
    uint8_t *d_pat = nullptr;              // [offsets i64 x (nq+1)] [pattern bytes]: one broadcast / one H2D
    size_t   d_pat_cap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    std::shared_ptr<pss::PinnedPool> pool = std::make_shared<pss::PinnedPool>();
    // rank-0 buffers of the distributed search (dist.cu)
    pss::DeviceBuf dist_recv, dist_final, dist_desc, dist_qoff;
    // Multi-GPU front (two or more devices listed): this object then owns no GPU state
    // itself; subs[g] is an ordinary sharded reader on device g holding the chunks k with
    // k % G == g, and a batch is answered by all of them concurrently.
    std::vector<pss_reader *> subs;

    ~pss_reader();
    int ensure_patterns(size_t bytes);     // d_pat / h_pat capacity
};
