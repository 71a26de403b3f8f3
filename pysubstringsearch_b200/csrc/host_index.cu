// host_index.cu — C++ host side of the index container: Writer (chunk accumulator +
// serialiser) and Reader (container parser, GPU upload, batched search), exported through
// the C ABI of include/pss.h.  Mirrors the reference's Rust host (src/lib.rs:42-288); the
// suffix array and every search step run on the GPU — there is no CPU path here.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <string>
#include <vector>

#include "common.cuh"
#include "sa_build.cuh"
#include "search.cuh"

using namespace pss;

namespace {

int io_fail(const std::string &what, const char *path) {
    const int e = errno;
    return fail(e == ENOENT ? PSS_ERR_NOTFOUND : PSS_ERR_IO,
                what + " '" + (path ? path : "") + "': " + std::strerror(e));
}

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

}  // namespace

// ======================================================================================
// Writer
// ======================================================================================
struct pss_writer {
    FILE *file = nullptr;
    std::vector<uint8_t> text;   // the chunk being accumulated (size() == bytes buffered)
    size_t capacity = 0;         // logical Vec<u8> capacity the flush rule compares against
    std::unique_ptr<SaBuilder> builder;

    // In-order background serialiser: dump_data() hands a finished chunk (text + suffix
    // array) to this thread and returns, so ingestion of chunk k+1 overlaps the multi-GB
    // file write of chunk k.  Two pinned SA buffers ping-pong; records are written strictly
    // in chunk order by the single thread, so the container is byte-identical to the
    // reference's synchronous writer (lib.rs:112-119) and is valid after every completed record.
    struct Job {
        std::vector<uint8_t> text;
        std::vector<int32_t> sa_small;   // suffix array of a small chunk (owned by the job)
        int stage = -1;                  // or: index of the pinned buffer holding it
    };
    Pinned sa_stage[2];
    bool   stage_busy[2] = {false, false};
    std::deque<Job> queue;
    std::mutex mu;
    std::condition_variable cv;
    std::thread io_thread;
    bool io_started = false, io_stop = false, io_active = false;
    int  io_status = PSS_OK;
    std::string io_error;

    // Rust's Vec::reserve on the logical capacity (RawVec::grow_amortized): the reference
    // flushes on `len + entry + 1 > capacity()` (lib.rs:75, :96), and an entry that exactly
    // fills the buffer, or an over-long line from a file, grows it for good.
    void reserve_logical(size_t additional) {
        const size_t len = text.size();
        if (capacity - len >= additional) return;
        const size_t need = len + additional;
        capacity = std::max<size_t>(std::max(capacity * 2, need), 8);
    }
    void push_entry(const uint8_t *p, size_t n) {
        reserve_logical(n);
        text.insert(text.end(), p, p + n);
        reserve_logical(1);
        text.push_back('\n');
    }

    void io_main() {
        std::unique_lock<std::mutex> lock(mu);
        while (true) {
            cv.wait(lock, [&] { return io_stop || !queue.empty(); });
            if (queue.empty()) break;   // io_stop and nothing left
            Job job = std::move(queue.front());
            queue.pop_front();
            io_active = true;
            lock.unlock();
            const size_t n = job.text.size();
            const int32_t *sa = job.stage >= 0 ? static_cast<const int32_t *>(sa_stage[job.stage].p) : job.sa_small.data();
            const uint32_t n32 = (uint32_t)n, sab = (uint32_t)(n * 4);
            bool ok = io_status == PSS_OK;
            int err = 0;
            if (ok) {
                ok = std::fwrite(&n32, 4, 1, file) == 1 && std::fwrite(job.text.data(), 1, n, file) == n &&
                     std::fwrite(&sab, 4, 1, file) == 1 && std::fwrite(sa, 4, n, file) == n;
                err = errno;
            }
            lock.lock();
            if (!ok && io_status == PSS_OK) {
                io_status = PSS_ERR_IO;
                io_error  = std::string("write to index file: ") + std::strerror(err);
            }
            if (job.stage >= 0) stage_busy[job.stage] = false;
            io_active = false;
            cv.notify_all();
        }
    }
    // Status of the background writes so far (call with `mu` NOT held).
    int check_io() {
        std::lock_guard<std::mutex> lock(mu);
        if (io_status != PSS_OK) return fail(io_status, io_error);
        return PSS_OK;
    }
    // Wait until everything queued has reached the FILE buffer.
    int drain() {
        std::unique_lock<std::mutex> lock(mu);
        cv.wait(lock, [&] { return queue.empty() && !io_active; });
        if (io_status != PSS_OK) return fail(io_status, io_error);
        return PSS_OK;
    }
    void stop_io() {
        if (!io_started) return;
        {
            std::lock_guard<std::mutex> lock(mu);
            io_stop = true;
        }
        cv.notify_all();
        io_thread.join();
        io_started = false;
    }
};

static int writer_dump(pss_writer *w) {
    const size_t n = w->text.size();
    if (n == 0) return PSS_OK;
    if (n >= (1ull << 30)) return fail(PSS_ERR_ARG, "chunk of 2^30 bytes or more: the container's u32 length fields would wrap");
    PSS_TRY(w->check_io());
    if (!w->builder) {
        w->builder.reset(new (std::nothrow) SaBuilder());
        if (!w->builder) return fail(PSS_ERR_NOMEM, "out of host memory");
        int rc = w->builder->init(-1, 0);
        if (rc != PSS_OK) { w->builder.reset(); return rc; }
    }
    pss_writer::Job job;
    // SA lands in pinned memory for large chunks (full-rate D2H), in a plain vector otherwise.
    int32_t *sa = nullptr;
    if (n >= (1u << 20)) {
        std::unique_lock<std::mutex> lock(w->mu);
        w->cv.wait(lock, [&] { return !w->stage_busy[0] || !w->stage_busy[1]; });
        job.stage = w->stage_busy[0] ? 1 : 0;
        w->stage_busy[job.stage] = true;
        lock.unlock();
        int rc = w->sa_stage[job.stage].ensure(n * sizeof(int32_t));
        if (rc != PSS_OK) {
            std::lock_guard<std::mutex> relock(w->mu);
            w->stage_busy[job.stage] = false;
            return rc;
        }
        sa = static_cast<int32_t *>(w->sa_stage[job.stage].p);
    } else {
        job.sa_small.resize(n);
        sa = job.sa_small.data();
    }
    int rc = w->builder->build_host(w->text.data(), (int32_t)n, sa);
    if (rc != PSS_OK) {
        if (job.stage >= 0) {
            std::lock_guard<std::mutex> relock(w->mu);
            w->stage_busy[job.stage] = false;
            w->cv.notify_all();
        }
        return rc;
    }
    job.text.swap(w->text);   // the accumulator starts the next chunk empty
    if (!w->io_started) {
        try {
            w->io_thread = std::thread([w] { w->io_main(); });
        } catch (...) {
            w->text.swap(job.text);
            if (job.stage >= 0) w->stage_busy[job.stage] = false;
            return fail(PSS_ERR_NOMEM, "cannot start the index writer thread");
        }
        w->io_started = true;
    }
    {
        std::lock_guard<std::mutex> lock(w->mu);
        w->queue.push_back(std::move(job));
    }
    w->cv.notify_all();
    return PSS_OK;
}

static int writer_finalize(pss_writer *w) {
    if (!w->text.empty()) PSS_TRY(writer_dump(w));
    PSS_TRY(w->drain());
    if (std::fflush(w->file) != 0) return io_fail("flush index file", nullptr);
    return PSS_OK;
}

extern "C" {

int32_t pss_writer_open(const char *index_file_path, int64_t max_chunk_len, pss_writer **out) {
    if (!index_file_path || !out) return fail(PSS_ERR_ARG, "null argument");
    *out = nullptr;
    std::unique_ptr<pss_writer> w(new (std::nothrow) pss_writer());
    if (!w) return fail(PSS_ERR_NOMEM, "out of host memory");
    w->file = std::fopen(index_file_path, "wb");
    if (!w->file) return io_fail("create", index_file_path);
    w->capacity = max_chunk_len < 0 ? (size_t)512 * 1024 * 1024 : (size_t)max_chunk_len;
    *out = w.release();
    return PSS_OK;
}

int32_t pss_writer_add_entry(pss_writer *w, const uint8_t *text, size_t len) {
    if (!w || (!text && len)) return fail(PSS_ERR_ARG, "null argument");
    if (len > w->capacity) return fail(PSS_ERR_TOOBIG, "entry is too big");
    if (w->text.size() + len + 1 > w->capacity) PSS_TRY(writer_dump(w));
    w->push_entry(text, len);
    return PSS_OK;
}

int32_t pss_writer_add_entries_from_file_lines(pss_writer *w, const char *input_file_path) {
    if (!w || !input_file_path) return fail(PSS_ERR_ARG, "null argument");
    FILE *in = std::fopen(input_file_path, "rb");
    if (!in) return io_fail("open", input_file_path);
    // Line splitting as bstr's for_byte_line does it (lib.rs:73): records end at '\n'; the
    // terminator and one '\r' directly before it are dropped; a final unterminated record
    // counts and keeps a trailing '\r' (bstr trims "\r" only as part of "\r\n").
    std::vector<uint8_t> block(1 << 20), line;
    int rc = PSS_OK;
    auto emit = [&](const uint8_t *p, size_t n, bool terminated) -> int {
        if (terminated && n && p[n - 1] == '\r') --n;
        if (w->text.size() + n + 1 > w->capacity) PSS_TRY(writer_dump(w));
        w->push_entry(p, n);
        return PSS_OK;
    };
    size_t got;
    while (rc == PSS_OK && (got = std::fread(block.data(), 1, block.size(), in)) > 0) {
        size_t from = 0;
        while (rc == PSS_OK) {
            const uint8_t *nl = static_cast<const uint8_t *>(std::memchr(block.data() + from, '\n', got - from));
            if (!nl) break;
            const size_t upto = (size_t)(nl - block.data());
            if (line.empty()) {
                rc = emit(block.data() + from, upto - from, true);
            } else {
                line.insert(line.end(), block.data() + from, block.data() + upto);
                rc = emit(line.data(), line.size(), true);
                line.clear();
            }
            from = upto + 1;
        }
        line.insert(line.end(), block.data() + from, block.data() + got);
    }
    if (rc == PSS_OK && std::ferror(in)) rc = io_fail("read", input_file_path);
    if (rc == PSS_OK && !line.empty()) rc = emit(line.data(), line.size(), false);
    std::fclose(in);
    return rc;
}

int32_t pss_writer_dump_data(pss_writer *w) {
    if (!w) return fail(PSS_ERR_ARG, "null writer");
    return writer_dump(w);
}

int32_t pss_writer_finalize(pss_writer *w) {
    if (!w) return fail(PSS_ERR_ARG, "null writer");
    return writer_finalize(w);
}

int32_t pss_writer_close(pss_writer *w) {
    if (!w) return PSS_OK;
    int rc = writer_finalize(w);
    w->stop_io();
    if (std::fclose(w->file) != 0 && rc == PSS_OK) rc = io_fail("close index file", nullptr);
    delete w;
    return rc;
}

}  // extern "C"

// ======================================================================================
// Reader
// ======================================================================================
namespace {

struct ChunkHost {
    uint64_t file_text_off = 0, file_sa_off = 0;
    uint32_t n = 0;
    uint64_t sa_bytes = 0;
    bool     owned = false;
    std::vector<uint8_t> text;  // host copy (result materialisation), owned chunks only
    uint8_t *d_text = nullptr;
    int32_t *d_sa   = nullptr;
};

// One batch of results in a pinned block: [chunk i32 x cap][start u32 x cap][end u32 x cap].
struct ResultOwner {
    pss_result pub;
    std::vector<int64_t> query_offsets;
    std::shared_ptr<PinnedPool> pool;
    PinnedBlock blk;
    int64_t cap = 0, used = 0;
    ~ResultOwner() { if (pool) pool->release(blk); }
    int32_t  *chunk() const { return static_cast<int32_t *>(blk.p); }
    uint32_t *start() const { return reinterpret_cast<uint32_t *>(chunk() + cap); }
    uint32_t *end() const { return start() + cap; }
    int ensure(int64_t need) {
        if (need <= cap) return PSS_OK;
        int64_t ncap = std::max<int64_t>(need + need / 2, 1 << 12);
        PinnedBlock nb;
        PSS_TRY(pool->acquire((size_t)ncap * 12, &nb));
        ncap = (int64_t)(nb.cap / 12);
        if (used) {
            int32_t *nc = static_cast<int32_t *>(nb.p);
            std::memcpy(nc, chunk(), (size_t)used * 4);
            std::memcpy(nc + ncap, start(), (size_t)used * 4);
            std::memcpy(nc + 2 * ncap, end(), (size_t)used * 4);
        }
        pool->release(blk);
        blk = nb;
        cap = ncap;
        return PSS_OK;
    }
};

// Sends the entries of every sub-batch to the result being built.  Lives in the reader:
// its device staging buffers are allocated once and only grow.
struct HostSink : SearchSink {
    ResultOwner *res = nullptr;
    int32_t  *d_chunk = nullptr;
    uint32_t *d_start = nullptr, *d_end = nullptr;
    int64_t   cap = 0;
    ~HostSink() override { cudaFree(d_chunk); cudaFree(d_start); cudaFree(d_end); }
    int reserve(int64_t count, int32_t **q, int32_t **c, uint32_t **s, uint32_t **e) override {
        if (count > cap) {
            cudaFree(d_chunk); cudaFree(d_start); cudaFree(d_end);
            d_chunk = nullptr; d_start = d_end = nullptr; cap = 0;
            int64_t nc = std::max<int64_t>(count + count / 2, 1 << 16);
            PSS_CUDA_TRY(cudaMalloc(&d_chunk, nc * sizeof(int32_t)));
            PSS_CUDA_TRY(cudaMalloc(&d_start, nc * sizeof(uint32_t)));
            PSS_CUDA_TRY(cudaMalloc(&d_end, nc * sizeof(uint32_t)));
            cap = nc;
        }
        *q = nullptr; *c = d_chunk; *s = d_start; *e = d_end;
        return PSS_OK;
    }
    int commit(int64_t count, cudaStream_t st) override {
        if (count == 0) return PSS_OK;
        PSS_TRY(res->ensure(res->used + count));
        const int64_t at = res->used;
        PSS_CUDA_TRY(cudaMemcpyAsync(res->chunk() + at, d_chunk, count * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        PSS_CUDA_TRY(cudaMemcpyAsync(res->start() + at, d_start, count * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        PSS_CUDA_TRY(cudaMemcpyAsync(res->end() + at, d_end, count * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        res->used += count;
        return PSS_OK;
    }
    int deliver_host(int64_t count, const int32_t *, const int32_t *c, const uint32_t *s, const uint32_t *e,
                     cudaStream_t) override {
        if (count == 0) return PSS_OK;
        PSS_TRY(res->ensure(res->used + count));
        std::memcpy(res->chunk() + res->used, c, (size_t)count * 4);
        std::memcpy(res->start() + res->used, s, (size_t)count * 4);
        std::memcpy(res->end() + res->used, e, (size_t)count * 4);
        res->used += count;
        return PSS_OK;
    }
};

// Writes straight into caller-provided device buffers (one-process-per-GPU + NCCL gather).
struct DeviceSink : SearchSink {
    int32_t *q, *c;
    uint32_t *s, *e;
    int64_t cap, used = 0, wanted = 0;
    bool overflow = false;
    int32_t *spill_q = nullptr, *spill_c = nullptr;
    uint32_t *spill_s = nullptr, *spill_e = nullptr;
    int64_t spill_cap = 0;
    DeviceSink(int32_t *q_, int32_t *c_, uint32_t *s_, uint32_t *e_, int64_t cap_) : q(q_), c(c_), s(s_), e(e_), cap(cap_) {}
    ~DeviceSink() override { cudaFree(spill_q); cudaFree(spill_c); cudaFree(spill_s); cudaFree(spill_e); }
    int reserve(int64_t count, int32_t **oq, int32_t **oc, uint32_t **os, uint32_t **oe) override {
        wanted += count;
        if (!overflow && used + count <= cap) {
            *oq = q ? q + used : nullptr; *oc = c ? c + used : nullptr; *os = s + used; *oe = e + used;
            return PSS_OK;
        }
        // past capacity: keep counting (so the caller learns the size it needs) but send
        // the data to a scratch area
        overflow = true;
        if (count > spill_cap) {
            cudaFree(spill_q); cudaFree(spill_c); cudaFree(spill_s); cudaFree(spill_e);
            spill_q = spill_c = nullptr; spill_s = spill_e = nullptr; spill_cap = 0;
            PSS_CUDA_TRY(cudaMalloc(&spill_q, std::max<int64_t>(count, 1) * 4));
            PSS_CUDA_TRY(cudaMalloc(&spill_c, std::max<int64_t>(count, 1) * 4));
            PSS_CUDA_TRY(cudaMalloc(&spill_s, std::max<int64_t>(count, 1) * 4));
            PSS_CUDA_TRY(cudaMalloc(&spill_e, std::max<int64_t>(count, 1) * 4));
            spill_cap = count;
        }
        *oq = spill_q; *oc = spill_c; *os = spill_s; *oe = spill_e;
        return PSS_OK;
    }
    int commit(int64_t count, cudaStream_t) override {
        if (!overflow) used += count;
        return PSS_OK;
    }
};

}  // namespace

struct pss_reader {
    std::string path;
    std::vector<ChunkHost> chunks;
    Searcher searcher;
    int shard_rank = 0, shard_count = 1;
    // pattern staging
    Pinned   h_pat;
    uint8_t *d_pat = nullptr;
    int64_t *d_off = nullptr;
    size_t   d_pat_cap = 0, d_off_cap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    HostSink sink;
    std::shared_ptr<PinnedPool> pool = std::make_shared<PinnedPool>();
    std::vector<int64_t> per_pair;
    // Multi-GPU front (PSS_DEVICES lists two or more devices): this object then owns no GPU
    // state itself; subs[g] is an ordinary sharded reader on device g holding the chunks
    // k with k % G == g, and a batch is answered by all of them concurrently.
    std::vector<pss_reader *> subs;

    ~pss_reader() {
        for (pss_reader *sub : subs) delete sub;
        if (searcher.device() >= 0) cudaSetDevice(searcher.device());
        for (auto &c : chunks) { cudaFree(c.d_text); cudaFree(c.d_sa); }
        cudaFree(d_pat); cudaFree(d_off);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    }
};

static int read_exact(int fd, void *dst, size_t n, uint64_t off) {
    uint8_t *p = static_cast<uint8_t *>(dst);
    while (n) {
        ssize_t g = ::pread(fd, p, n, (off_t)off);
        if (g < 0) { if (errno == EINTR) continue; return -1; }
        if (g == 0) { errno = 0; return -2; }
        p += g; off += (uint64_t)g; n -= (size_t)g;
    }
    return 0;
}

static int reader_open(const char *path, int shard_rank, int shard_count, int device, pss_reader **out) {
    if (!path || !out) return fail(PSS_ERR_ARG, "null argument");
    *out = nullptr;
    if (shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count) return fail(PSS_ERR_ARG, "bad shard spec");
    int fd = ::open(path, O_RDONLY);
    if (fd < 0) return io_fail("open", path);
    struct FdGuard { int fd; ~FdGuard() { ::close(fd); } } guard{fd};
    struct stat st;
    if (::fstat(fd, &st) != 0) return io_fail("stat", path);
    const uint64_t file_len = (uint64_t)st.st_size;

    std::unique_ptr<pss_reader> r(new (std::nothrow) pss_reader());
    if (!r) return fail(PSS_ERR_NOMEM, "out of host memory");
    r->path = path;
    r->shard_rank = shard_rank;
    r->shard_count = shard_count;

    // Walk the container (lib.rs:174-196): [u32 n][text n][u32 sa_bytes][sa]...
    uint64_t pos = 0;
    while (pos < file_len) {
        uint32_t n = 0, sab = 0;
        if (read_exact(fd, &n, 4, pos) != 0) return fail(PSS_ERR_FORMAT, "truncated index file (chunk header)");
        if (pos + 4 + (uint64_t)n + 4 > file_len) return fail(PSS_ERR_FORMAT, "truncated index file (chunk text)");
        if (read_exact(fd, &sab, 4, pos + 4 + n) != 0) return fail(PSS_ERR_FORMAT, "truncated index file (sa header)");
        ChunkHost c;
        c.n = n;
        c.file_text_off = pos + 4;
        c.file_sa_off   = pos + 8 + n;
        c.sa_bytes      = sab;
        if (c.file_sa_off + c.sa_bytes > file_len) return fail(PSS_ERR_FORMAT, "truncated index file (suffix array)");
        if (c.sa_bytes != (uint64_t)n * 4) return fail(PSS_ERR_FORMAT, "suffix array length does not match text length");
        c.owned = ((int)(r->chunks.size() % (size_t)shard_count) == shard_rank);
        r->chunks.push_back(std::move(c));
        pos += 8 + (uint64_t)n + (uint64_t)sab;
    }

    // GPU upload of the owned chunks: text (+16 zero bytes so 4-byte text reads never leave
    // the allocation) and SA, streamed through a double pinned bounce buffer.
    PSS_TRY(r->searcher.init(device));
    PSS_CUDA_TRY(cudaSetDevice(r->searcher.device()));
    cudaStream_t s = r->searcher.stream();
    PSS_CUDA_TRY(cudaEventCreate(&r->ev0));
    PSS_CUDA_TRY(cudaEventCreate(&r->ev1));
    constexpr size_t SLICE = 32u << 20;
    Pinned bounce[2];
    cudaEvent_t done[2] = {nullptr, nullptr};
    struct EvGuard { cudaEvent_t *e; ~EvGuard() { for (int i = 0; i < 2; ++i) if (e[i]) cudaEventDestroy(e[i]); } } evg{done};
    for (int i = 0; i < 2; ++i) {
        PSS_TRY(bounce[i].ensure(SLICE));
        PSS_CUDA_TRY(cudaEventCreate(&done[i]));
    }
    int which = 0;
    std::vector<DeviceChunk> dchunks;
    for (size_t k = 0; k < r->chunks.size(); ++k) {
        ChunkHost &c = r->chunks[k];
        if (!c.owned || c.n == 0) continue;
        c.text.resize(c.n);
        if (read_exact(fd, c.text.data(), c.n, c.file_text_off) != 0) return io_fail("read", path);
        PSS_CUDA_TRY(cudaMalloc(&c.d_text, (size_t)c.n + 16));
        PSS_CUDA_TRY(cudaMemsetAsync(c.d_text + c.n, 0, 16, s));
        PSS_CUDA_TRY(cudaMemcpyAsync(c.d_text, c.text.data(), c.n, cudaMemcpyHostToDevice, s));
        PSS_CUDA_TRY(cudaMalloc(&c.d_sa, (size_t)c.n * sizeof(int32_t)));
        for (uint64_t off = 0; off < c.sa_bytes; off += SLICE) {
            const size_t len = (size_t)std::min<uint64_t>(SLICE, c.sa_bytes - off);
            PSS_CUDA_TRY(cudaEventSynchronize(done[which]));  // bounce buffer free again
            if (read_exact(fd, bounce[which].p, len, c.file_sa_off + off) != 0) return io_fail("read", path);
            PSS_CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<uint8_t *>(c.d_sa) + off, bounce[which].p, len,
                                         cudaMemcpyHostToDevice, s));
            PSS_CUDA_TRY(cudaEventRecord(done[which], s));
            which ^= 1;
        }
        DeviceChunk dc;
        dc.text = c.d_text; dc.sa = c.d_sa; dc.n = c.n; dc.global_id = (int32_t)k;
        dchunks.push_back(dc);
    }
    PSS_CUDA_TRY(cudaStreamSynchronize(s));
    PSS_TRY(r->searcher.set_chunks(dchunks));
    *out = r.release();
    return PSS_OK;
}

extern "C" {

int32_t pss_reader_open(const char *index_file_path, pss_reader **out) {
    // PSS_DEVICES = "all" or "0,1,2,..." spreads the chunks of ONE Reader over several GPUs
    // of the box (chunk k -> listed device k % G), inside this process.
    std::vector<int> devices;
    if (const char *e = std::getenv("PSS_DEVICES")) {
        const int ndev = pss_device_count();
        if (std::strcmp(e, "all") == 0) {
            for (int d = 0; d < ndev; ++d) devices.push_back(d);
        } else {
            for (const char *q = e; *q;) {
                char *endp = nullptr;
                long v = std::strtol(q, &endp, 10);
                if (endp == q) break;
                if (v < 0 || v >= ndev) return fail(PSS_ERR_ARG, "PSS_DEVICES names a device that does not exist");
                devices.push_back((int)v);
                q = (*endp == ',') ? endp + 1 : endp;
            }
        }
    }
    if (devices.size() < 2) return reader_open(index_file_path, 0, 1, devices.empty() ? -1 : devices[0], out);
    if (!out) return fail(PSS_ERR_ARG, "null argument");
    *out = nullptr;
    std::unique_ptr<pss_reader> front(new (std::nothrow) pss_reader());
    if (!front) return fail(PSS_ERR_NOMEM, "out of host memory");
    front->path = index_file_path ? index_file_path : "";
    const int G = (int)devices.size();
    for (int g = 0; g < G; ++g) {
        pss_reader *sub = nullptr;
        PSS_TRY(reader_open(index_file_path, g, G, devices[g], &sub));   // ~pss_reader frees the earlier ones
        front->subs.push_back(sub);
    }
    *out = front.release();
    return PSS_OK;
}

int32_t pss_reader_open_sharded(const char *index_file_path, int32_t shard_rank, int32_t shard_count,
                                pss_reader **out) {
    return reader_open(index_file_path, shard_rank, shard_count, -1, out);
}

int32_t pss_reader_close(pss_reader *r) {
    delete r;
    return PSS_OK;
}

int32_t pss_reader_num_chunks(const pss_reader *r) {
    if (!r) return 0;
    return (int32_t)(r->subs.empty() ? r->chunks.size() : r->subs[0]->chunks.size());
}
int32_t pss_reader_num_local_chunks(const pss_reader *r) {
    if (!r) return 0;
    if (r->subs.empty()) return r->searcher.num_chunks();
    int32_t total = 0;
    for (const pss_reader *sub : r->subs) total += sub->searcher.num_chunks();
    return total;
}

int32_t pss_reader_chunk_text(const pss_reader *r, int32_t chunk, const uint8_t **text, int64_t *len) {
    if (r && !r->subs.empty() && chunk >= 0)
        return pss_reader_chunk_text(r->subs[(size_t)chunk % r->subs.size()], chunk, text, len);
    if (!r || !text || !len || chunk < 0 || chunk >= (int32_t)r->chunks.size()) return fail(PSS_ERR_ARG, "bad chunk index");
    const ChunkHost &c = r->chunks[chunk];
    *text = c.owned ? c.text.data() : nullptr;
    *len  = c.owned ? (int64_t)c.n : 0;
    return PSS_OK;
}

// Multi-GPU front: every sub-reader answers the batch for its own chunks on its own GPU (one
// host thread each); the per-device results, each ordered by (query, chunk), are merged into
// the single-process order: query, then ascending global chunk id.
static int32_t search_batch_multi(pss_reader *r, const uint8_t *patterns, const int64_t *offsets, int32_t nq,
                                  pss_result **out) {
    const size_t G = r->subs.size();
    std::vector<pss_result *> part(G, nullptr);
    std::vector<int> rcs(G, PSS_OK);
    std::vector<std::string> errs(G);
    {
        std::vector<std::thread> workers;
        for (size_t g = 0; g < G; ++g)
            workers.emplace_back([&, g] {
                rcs[g] = pss_reader_search_batch(r->subs[g], patterns, offsets, nq, &part[g]);
                if (rcs[g] != PSS_OK) errs[g] = pss_last_error();
            });
        for (auto &w : workers) w.join();
    }
    struct Cleanup {
        std::vector<pss_result *> &p;
        ~Cleanup() { for (pss_result *x : p) pss_result_free(x); }
    } cleanup{part};
    for (size_t g = 0; g < G; ++g)
        if (rcs[g] != PSS_OK) return fail(rcs[g], errs[g]);

    std::unique_ptr<ResultOwner> res(new (std::nothrow) ResultOwner());
    if (!res) return fail(PSS_ERR_NOMEM, "out of host memory");
    std::memset(&res->pub, 0, sizeof(res->pub));
    res->query_offsets.assign((size_t)nq + 1, 0);
    res->pool = r->pool;
    int64_t total = 0;
    for (size_t g = 0; g < G; ++g) total += part[g]->n_entries;
    PSS_TRY(res->ensure(std::max<int64_t>(total, 1)));
    std::vector<int64_t> cur(G);
    int64_t at = 0;
    for (int32_t q = 0; q < nq; ++q) {
        for (size_t g = 0; g < G; ++g) cur[g] = part[g]->query_offsets[q];
        while (true) {
            // the sub-reader whose next entry of this query has the smallest chunk id goes next
            int best = -1;
            int32_t best_chunk = 0;
            for (size_t g = 0; g < G; ++g)
                if (cur[g] < part[g]->query_offsets[q + 1] && (best < 0 || part[g]->chunk_id[cur[g]] < best_chunk)) {
                    best = (int)g;
                    best_chunk = part[g]->chunk_id[cur[g]];
                }
            if (best < 0) break;
            const pss_result *p = part[best];
            int64_t end = cur[best];
            while (end < p->query_offsets[q + 1] && p->chunk_id[end] == best_chunk) ++end;
            const int64_t cnt = end - cur[best];
            std::memcpy(res->chunk() + at, p->chunk_id + cur[best], (size_t)cnt * 4);
            std::memcpy(res->start() + at, p->line_start + cur[best], (size_t)cnt * 4);
            std::memcpy(res->end() + at, p->line_end + cur[best], (size_t)cnt * 4);
            at += cnt;
            cur[best] = end;
        }
        res->query_offsets[q + 1] = at;
    }
    res->used = at;
    for (size_t g = 0; g < G; ++g) {
        res->pub.n_hits += part[g]->n_hits;
        res->pub.ms_bounds  = std::max(res->pub.ms_bounds, part[g]->ms_bounds);
        res->pub.ms_extract = std::max(res->pub.ms_extract, part[g]->ms_extract);
        res->pub.ms_dedup   = std::max(res->pub.ms_dedup, part[g]->ms_dedup);
        res->pub.ms_total   = std::max(res->pub.ms_total, part[g]->ms_total);
    }
    res->pub.n_queries     = nq;
    res->pub.n_entries     = res->used;
    res->pub.query_offsets = res->query_offsets.data();
    res->pub.chunk_id      = res->used ? res->chunk() : nullptr;
    res->pub.line_start    = res->used ? res->start() : nullptr;
    res->pub.line_end      = res->used ? res->end() : nullptr;
    *out = &res.release()->pub;
    return PSS_OK;
}

int32_t pss_reader_search_batch(pss_reader *r, const uint8_t *patterns, const int64_t *offsets, int32_t nq,
                                pss_result **out) {
    if (!r || !out || nq < 0 || (nq > 0 && !offsets)) return fail(PSS_ERR_ARG, "bad search arguments");
    *out = nullptr;
    if (!r->subs.empty()) return search_batch_multi(r, patterns, offsets, nq, out);
    std::unique_ptr<ResultOwner> res(new (std::nothrow) ResultOwner());
    if (!res) return fail(PSS_ERR_NOMEM, "out of host memory");
    std::memset(&res->pub, 0, sizeof(res->pub));
    res->query_offsets.assign((size_t)nq + 1, 0);
    res->pool = r->pool;
    const int nc = r->searcher.num_chunks();
    if (nq > 0 && nc > 0) {
        const int64_t total = offsets[nq];
        if (total < 0 || (total > 0 && !patterns)) return fail(PSS_ERR_ARG, "bad pattern buffer");
        PSS_CUDA_TRY(cudaSetDevice(r->searcher.device()));
        cudaStream_t s = r->searcher.stream();
        // stage patterns + offsets in pinned memory, one H2D each
        const size_t off_bytes = ((size_t)nq + 1) * sizeof(int64_t);
        PSS_TRY(r->h_pat.ensure(off_bytes + (size_t)total + 64));
        std::memcpy(r->h_pat.p, offsets, off_bytes);
        if (total) std::memcpy(static_cast<uint8_t *>(r->h_pat.p) + off_bytes, patterns, (size_t)total);
        if ((size_t)total + 64 > r->d_pat_cap) {
            cudaFree(r->d_pat); r->d_pat = nullptr; r->d_pat_cap = 0;
            size_t cap = std::max<size_t>((size_t)total + 64, 1 << 16);
            PSS_CUDA_TRY(cudaMalloc(&r->d_pat, cap));
            r->d_pat_cap = cap;
        }
        if ((size_t)nq + 1 > r->d_off_cap) {
            cudaFree(r->d_off); r->d_off = nullptr; r->d_off_cap = 0;
            size_t cap = std::max<size_t>((size_t)nq + 1, 1 << 12);
            PSS_CUDA_TRY(cudaMalloc(&r->d_off, cap * sizeof(int64_t)));
            r->d_off_cap = cap;
        }
        PSS_CUDA_TRY(cudaEventRecord(r->ev0, s));
        PSS_CUDA_TRY(cudaMemcpyAsync(r->d_off, r->h_pat.p, off_bytes, cudaMemcpyHostToDevice, s));
        if (total)
            PSS_CUDA_TRY(cudaMemcpyAsync(r->d_pat, static_cast<uint8_t *>(r->h_pat.p) + off_bytes, (size_t)total,
                                         cudaMemcpyHostToDevice, s));
        r->sink.res = res.get();
        std::vector<int64_t> &per_pair = r->per_pair;
        per_pair.assign((size_t)nq * nc, 0);
        SearchTimes times;
        int64_t n_hits = 0;
        int rc = r->searcher.search(r->d_pat, r->d_off, nq, s, &r->sink, per_pair.data(), &n_hits, &times);
        r->sink.res = nullptr;
        if (rc != PSS_OK) return rc;
        PSS_CUDA_TRY(cudaEventRecord(r->ev1, s));
        PSS_CUDA_TRY(cudaEventSynchronize(r->ev1));
        PSS_CUDA_TRY(cudaEventElapsedTime(&times.ms_total, r->ev0, r->ev1));
        for (int32_t q = 0; q < nq; ++q) {
            int64_t cnt = 0;
            for (int c = 0; c < nc; ++c) cnt += per_pair[(size_t)q * nc + c];
            res->query_offsets[q + 1] = res->query_offsets[q] + cnt;
        }
        if (res->used != res->query_offsets[nq])
            return fail(PSS_ERR_CUDA, "internal error: per-query counts do not add up to the entry count");
        res->pub.n_hits     = n_hits;
        res->pub.ms_bounds  = times.ms_bounds;
        res->pub.ms_extract = times.ms_extract;
        res->pub.ms_dedup   = times.ms_dedup;
        res->pub.ms_total   = times.ms_total;
    }
    res->pub.n_queries     = nq;
    res->pub.n_entries     = res->used;
    res->pub.query_offsets = res->query_offsets.data();
    res->pub.chunk_id      = res->used ? res->chunk() : nullptr;
    res->pub.line_start    = res->used ? res->start() : nullptr;
    res->pub.line_end      = res->used ? res->end() : nullptr;
    *out = &res.release()->pub;
    return PSS_OK;
}

int32_t pss_reader_search_batch_device(pss_reader *r, const uint8_t *d_patterns, const int64_t *d_offsets,
                                       int32_t nq, int64_t total_pattern_bytes, int32_t *d_query_id,
                                       int32_t *d_chunk_id, uint32_t *d_line_start, uint32_t *d_line_end,
                                       int64_t capacity, int64_t *n_entries, int64_t *n_hits, void *stream) {
    (void)total_pattern_bytes;
    if (!r || nq < 0 || !n_entries || capacity < 0 || (capacity > 0 && (!d_line_start || !d_line_end)))
        return fail(PSS_ERR_ARG, "bad search arguments");
    *n_entries = 0;
    if (n_hits) *n_hits = 0;
    if (!r->subs.empty())
        return fail(PSS_ERR_ARG, "device-resident search needs a single-device reader (unset PSS_DEVICES)");
    if (nq == 0 || r->searcher.num_chunks() == 0) return PSS_OK;
    DeviceSink sink(d_query_id, d_chunk_id, d_line_start, d_line_end, capacity);
    PSS_TRY(r->searcher.search(d_patterns, d_offsets, nq, static_cast<cudaStream_t>(stream), &sink, nullptr, n_hits,
                               nullptr));
    *n_entries = sink.wanted;
    if (sink.overflow) return fail(PSS_ERR_NOMEM, "result buffers too small; *n_entries holds the required capacity");
    return PSS_OK;
}

void pss_result_free(pss_result *res) {
    if (!res) return;
    // pub is the first member of ResultOwner
    delete reinterpret_cast<ResultOwner *>(res);
}

}  // extern "C"
