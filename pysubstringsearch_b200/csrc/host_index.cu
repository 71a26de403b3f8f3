// host_index.cu — C++ host side of the index container: Writer (chunk accumulator +
// pipelined serialiser) and Reader (container parser, GPU upload, batched search), exported
// through the C ABI of include/pss.h.  Mirrors the reference's Rust host (src/lib.rs:42-288);
// the suffix array and every search step run on the GPU — there is no CPU path here.
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

#include "build_engine.cuh"
#include "common.cuh"
#include "reader.cuh"
#include "search.cuh"

using namespace pss;

namespace {

int io_fail(const std::string &what, const char *path) {
    const int e = errno;
    return fail(e == ENOENT ? PSS_ERR_NOTFOUND : PSS_ERR_IO,
                what + " '" + (path ? path : "") + "': " + std::strerror(e));
}

// "all" | "0,1,2": the device list of env PSS_DEVICES (empty when unset).
int devices_from_env(std::vector<int> *devices) {
    devices->clear();
    const char *e = std::getenv("PSS_DEVICES");
    if (!e || !*e) return PSS_OK;
    const int ndev = pss_device_count();
    if (std::strcmp(e, "all") == 0) {
        for (int d = 0; d < ndev; ++d) devices->push_back(d);
        return PSS_OK;
    }
    for (const char *q = e; *q;) {
        char *endp = nullptr;
        long v = std::strtol(q, &endp, 10);
        if (endp == q) break;
        if (v < 0 || v >= ndev) return fail(PSS_ERR_ARG, "PSS_DEVICES names a device that does not exist");
        devices->push_back((int)v);
        q = (*endp == ',') ? endp + 1 : endp;
    }
    return PSS_OK;
}

}  // namespace

// ======================================================================================
// Writer
// ======================================================================================
// dump_data() hands the buffered chunk to the build engine of the next GPU in the chunk →
// GPU map (pss_sa_build_begin) and returns; ingestion of chunk k+1 goes on while chunk k
// is built.  One serialiser thread takes the chunks strictly in order: it waits for the
// chunk's suffix array (D2H into a pinned buffer; the engine's second slot lets the next
// build run meanwhile) and writes the record (u32 n, text, u32 4n, SA — lib.rs:112-119).
// The container is therefore byte-identical to the reference's synchronous writer and is
// valid after every completed record.
struct pss_writer {
    FILE *file = nullptr;
    std::vector<uint8_t> text;   // the chunk being accumulated (size() == bytes buffered)
    size_t capacity = 0;         // logical Vec<u8> capacity the flush rule compares against
    std::vector<int> devices;    // chunk k → devices[k % size]; empty = the default device
    size_t n_dumped = 0;

    struct Job {
        std::vector<uint8_t> text;
        BuildEngine *engine = nullptr;
        BuildEngine::Job *handle = nullptr;
    };
    std::deque<Job> queue;
    std::vector<std::vector<uint8_t>> spare;   // recycled text buffers (no regrowth per chunk)
    size_t max_queue = 2;
    Pinned sa_stage;
    std::vector<int32_t> sa_small;
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
            const bool healthy = io_status == PSS_OK;
            lock.unlock();
            cv.notify_all();            // a queue slot is free: ingestion may hand over the next chunk

            const size_t n = job.text.size();
            int32_t *sa = nullptr;
            if (n * sizeof(int32_t) >= (1u << 20) && sa_stage.ensure(n * sizeof(int32_t)) == PSS_OK) {
                sa = static_cast<int32_t *>(sa_stage.p);    // full-rate D2H
            } else {
                sa_small.resize(n);
                sa = sa_small.data();
            }
            int rc = job.engine->wait(job.handle, sa);      // always: frees the engine's slot
            std::string err = rc == PSS_OK ? std::string() : std::string(pss_last_error());
            if (rc == PSS_OK && healthy) {
                const uint32_t n32 = (uint32_t)n, sab = (uint32_t)(n * 4);
                const bool ok = std::fwrite(&n32, 4, 1, file) == 1 && std::fwrite(job.text.data(), 1, n, file) == n &&
                                std::fwrite(&sab, 4, 1, file) == 1 && std::fwrite(sa, 4, n, file) == n;
                if (!ok) {
                    rc  = PSS_ERR_IO;
                    err = std::string("write to index file: ") + std::strerror(errno);
                }
            }
            job.text.clear();
            lock.lock();
            if (rc != PSS_OK && io_status == PSS_OK) {
                io_status = rc;
                io_error  = err;
            }
            if (spare.size() < 2) spare.push_back(std::move(job.text));
            io_active = false;
            cv.notify_all();
        }
    }
    // Status of the background work so far (call with `mu` NOT held).
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
    const int device = w->devices.empty() ? -1 : w->devices[w->n_dumped % w->devices.size()];
    BuildEngine *engine = nullptr;
    PSS_TRY(BuildEngine::get(device, &engine));      // no GPU: fails here, loudly, chunk stays buffered
    if (!w->io_started) {
        try {
            w->io_thread = std::thread([w] { w->io_main(); });
        } catch (...) {
            return fail(PSS_ERR_NOMEM, "cannot start the index writer thread");
        }
        w->io_started = true;
    }
    pss_writer::Job job;
    job.engine = engine;
    std::vector<uint8_t> next;
    {
        // back-pressure: at most one chunk per GPU in flight plus one waiting to be written
        std::unique_lock<std::mutex> lock(w->mu);
        w->cv.wait(lock, [&] { return w->queue.size() + (w->io_active ? 1 : 0) < w->max_queue; });
        if (!w->spare.empty()) {
            next = std::move(w->spare.back());
            w->spare.pop_back();
        }
    }
    next.clear();
    next.reserve(std::min<size_t>(w->text.size(), w->capacity));   // the next chunk will be about as large
    job.text.swap(w->text);
    w->text.swap(next);
    int rc = engine->begin(job.text.data(), (int32_t)n, &job.handle);
    if (rc != PSS_OK) {
        w->text.swap(job.text);
        return rc;
    }
    ++w->n_dumped;
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

int32_t pss_writer_open_devices(const char *index_file_path, int64_t max_chunk_len, const int32_t *devices,
                                int32_t ndev, pss_writer **out) {
    if (!index_file_path || !out || ndev < 0 || (ndev > 0 && !devices)) return fail(PSS_ERR_ARG, "null argument");
    *out = nullptr;
    std::unique_ptr<pss_writer> w(new (std::nothrow) pss_writer());
    if (!w) return fail(PSS_ERR_NOMEM, "out of host memory");
    if (ndev > 0) {
        const int have = pss_device_count();
        for (int32_t i = 0; i < ndev; ++i) {
            if (devices[i] < 0 || (have > 0 && devices[i] >= have)) return fail(PSS_ERR_ARG, "device index out of range");
            w->devices.push_back(devices[i]);
        }
    } else {
        PSS_TRY(devices_from_env(&w->devices));
    }
    w->max_queue = std::max<size_t>(w->devices.size(), 1) + 1;
    w->file = std::fopen(index_file_path, "wb");
    if (!w->file) return io_fail("create", index_file_path);
    w->capacity = max_chunk_len < 0 ? (size_t)512 * 1024 * 1024 : (size_t)max_chunk_len;
    // Address space only (untouched pages cost nothing): a buffer that instead grows by doubling
    // copies the chunk once more and faults twice the pages — 0.77 s vs 0.30 s per 512 MiB here.
    try {
        w->text.reserve(std::min<size_t>(w->capacity, (size_t)1 << 30));
    } catch (const std::bad_alloc &) {
    }
    *out = w.release();
    return PSS_OK;
}

int32_t pss_writer_open(const char *index_file_path, int64_t max_chunk_len, pss_writer **out) {
    return pss_writer_open_devices(index_file_path, max_chunk_len, nullptr, 0, out);
}

int32_t pss_writer_add_entry(pss_writer *w, const uint8_t *text, size_t len) {
    if (!w || (!text && len)) return fail(PSS_ERR_ARG, "null argument");
    if (len > w->capacity) return fail(PSS_ERR_TOOBIG, "entry is too big");
    if (w->text.size() + len + 1 > w->capacity) PSS_TRY(writer_dump(w));
    w->push_entry(text, len);
    return PSS_OK;
}

/* 1 when adding an entry of `len` bytes would flush the buffered chunk first (the binding
 * releases the GIL around such calls only). */
int32_t pss_writer_would_flush(const pss_writer *w, size_t len) {
    return (w && w->text.size() + len + 1 > w->capacity && !w->text.empty()) ? 1 : 0;
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
        const uint8_t *b = block.data();
        size_t from = 0;
        // Bulk path.  A run of complete records without any '\r' is already in the chunk's own
        // format (entry, '\n', entry, '\n', ...), and appending its records one at a time
        // would neither flush nor grow the buffer as long as the whole run fits the room left
        // (the flush test `len + entry + 1 > capacity` is monotone in len): such a run is
        // appended with ONE copy.  Everything else — the record that continues the previous
        // block, the record that triggers the flush, "\r\n" records — takes the per-record
        // path below, so the chunk boundaries and bytes are those of the per-record loop.
        const bool no_cr = std::memchr(b, '\r', got) == nullptr;
        while (rc == PSS_OK) {
            if (no_cr && line.empty() && w->text.size() < w->capacity) {
                const size_t room = std::min(w->capacity - w->text.size(), got - from);
                const uint8_t *last = room ? static_cast<const uint8_t *>(memrchr(b + from, '\n', room)) : nullptr;
                if (last) {
                    w->text.insert(w->text.end(), b + from, last + 1);
                    from = (size_t)(last + 1 - b);
                }
            }
            const uint8_t *nl = static_cast<const uint8_t *>(std::memchr(b + from, '\n', got - from));
            if (!nl) break;
            const size_t upto = (size_t)(nl - b);
            if (line.empty()) {
                rc = emit(b + from, upto - from, true);
            } else {
                line.insert(line.end(), b + from, b + upto);
                rc = emit(line.data(), line.size(), true);
                line.clear();
            }
            from = upto + 1;
        }
        line.insert(line.end(), b + from, b + got);
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
    std::string err = rc == PSS_OK ? std::string() : std::string(pss_last_error());
    // whatever is still queued must be waited for (the engines hold pointers into the jobs)
    w->drain();
    w->stop_io();
    if (std::fclose(w->file) != 0 && rc == PSS_OK) {
        rc  = io_fail("close index file", nullptr);
        err = pss_last_error();
    }
    delete w;
    if (rc != PSS_OK) return fail(rc, err);
    return PSS_OK;
}

}  // extern "C"

// ======================================================================================
// Reader
// ======================================================================================
pss_reader::~pss_reader() {
    for (pss_reader *sub : subs) delete sub;
    if (searcher.device() >= 0) cudaSetDevice(searcher.device());
    for (auto &c : chunks) {
        if (!c.borrowed) { cudaFree(c.d_text); cudaFree(c.d_sa); }
        cudaFree(c.d_nl);
        cudaFree(c.d_bucket);
        cudaFree(c.d_dir);
        cudaFree(c.d_rec);
    }
    cudaFree(d_pat);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (ev2) cudaEventDestroy(ev2);
}

int pss_reader::ensure_patterns(size_t bytes) {
    PSS_TRY(h_pat.ensure(bytes + 64));
    if (bytes + 64 > d_pat_cap) {
        cudaFree(d_pat);
        d_pat = nullptr; d_pat_cap = 0;
        size_t cap = std::max<size_t>(bytes + 64 + bytes / 4, 1 << 16);
        PSS_CUDA_TRY(cudaMalloc(&d_pat, cap));
        d_pat_cap = cap;
    }
    return PSS_OK;
}

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

static int reader_init_gpu(pss_reader *r, int device) {
    PSS_TRY(r->searcher.init(device));
    PSS_CUDA_TRY(cudaSetDevice(r->searcher.device()));
    PSS_CUDA_TRY(cudaEventCreate(&r->ev0));
    PSS_CUDA_TRY(cudaEventCreate(&r->ev1));
    PSS_CUDA_TRY(cudaEventCreate(&r->ev2));
    return PSS_OK;
}

// Newline side indexes + the searcher's chunk table, for the chunks that are on the device.
static int reader_finish_open(pss_reader *r) {
    std::vector<DeviceChunk> dchunks;
    for (size_t k = 0; k < r->chunks.size(); ++k) {
        ChunkHost &c = r->chunks[k];
        if (!c.owned || c.n == 0) continue;
        PSS_TRY(r->searcher.build_newline_index(c.d_text, c.n, &c.d_nl, &c.n_lines));
        // what extraction consults (A/B measurements): 2 (default) line records, 1 line directory, 0 text scans
        const char *use_dir = std::getenv("PSS_LINE_DIR");
        const int line_mode = use_dir ? std::atoi(use_dir) : 2;
        if (line_mode == 1) PSS_TRY(r->searcher.build_line_directory(c.d_nl, c.n_lines, c.n, &c.d_dir));
        else if (line_mode != 0) PSS_TRY(r->searcher.build_line_records(c.d_nl, c.n_lines, c.n, &c.d_rec));
        PSS_TRY(r->searcher.build_prefix_buckets(c.d_text, c.d_sa, c.n, &c.d_bucket));
        DeviceChunk dc = {};
        dc.text = c.d_text; dc.sa = c.d_sa; dc.nl = c.d_nl; dc.bucket = c.d_bucket; dc.dir = c.d_dir; dc.rec = c.d_rec; dc.n = c.n; dc.n_lines = c.n_lines;
        dc.global_id = (int32_t)k;
        dchunks.push_back(dc);
    }
    return r->searcher.set_chunks(dchunks);
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

    // GPU upload of the owned chunks: text (+16 zero bytes so vector text reads never leave
    // the allocation) and SA, streamed through a double pinned bounce buffer.
    PSS_TRY(reader_init_gpu(r.get(), device));
    cudaStream_t s = r->searcher.stream();
    constexpr size_t SLICE = 32u << 20;
    Pinned bounce[2];
    cudaEvent_t done[2] = {nullptr, nullptr};
    struct EvGuard { cudaEvent_t *e; ~EvGuard() { for (int i = 0; i < 2; ++i) if (e[i]) cudaEventDestroy(e[i]); } } evg{done};
    for (int i = 0; i < 2; ++i) {
        PSS_TRY(bounce[i].ensure(SLICE));
        PSS_CUDA_TRY(cudaEventCreate(&done[i]));
    }
    int which = 0;
    for (size_t k = 0; k < r->chunks.size(); ++k) {
        ChunkHost &c = r->chunks[k];
        if (!c.owned || c.n == 0) continue;
        c.text.resize(c.n);
        c.h_text = c.text.data();
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
    }
    PSS_CUDA_TRY(cudaStreamSynchronize(s));
    PSS_TRY(reader_finish_open(r.get()));
    *out = r.release();
    return PSS_OK;
}

static int reader_open_multi(const char *path, const std::vector<int> &devices, pss_reader **out) {
    if (!out) return fail(PSS_ERR_ARG, "null argument");
    *out = nullptr;
    if (devices.size() < 2) return reader_open(path, 0, 1, devices.empty() ? -1 : devices[0], out);
    std::unique_ptr<pss_reader> front(new (std::nothrow) pss_reader());
    if (!front) return fail(PSS_ERR_NOMEM, "out of host memory");
    front->path = path ? path : "";
    const int G = (int)devices.size();
    for (int g = 0; g < G; ++g) {
        pss_reader *sub = nullptr;
        PSS_TRY(reader_open(path, g, G, devices[g], &sub));   // ~pss_reader frees the earlier ones
        front->subs.push_back(sub);
    }
    *out = front.release();
    return PSS_OK;
}

// offsets[0] == 0, non-decreasing (ADVICE: a bad array would become a huge pattern length on the device)
static int validate_offsets(const int64_t *offsets, int32_t nq) {
    if (nq == 0) return PSS_OK;
    if (offsets[0] != 0) return fail(PSS_ERR_ARG, "pattern offsets must start at 0");
    for (int32_t q = 0; q < nq; ++q)
        if (offsets[q + 1] < offsets[q]) return fail(PSS_ERR_ARG, "pattern offsets must be non-decreasing");
    return PSS_OK;
}

extern "C" {

int32_t pss_reader_open(const char *index_file_path, pss_reader **out) {
    // PSS_DEVICES = "all" or "0,1,2,..." spreads the chunks of ONE Reader over several GPUs
    // of the box (chunk k -> listed device k % G), inside this process.
    std::vector<int> devices;
    PSS_TRY(devices_from_env(&devices));
    return reader_open_multi(index_file_path, devices, out);
}

int32_t pss_reader_open_devices(const char *index_file_path, const int32_t *devices, int32_t ndev, pss_reader **out) {
    if (ndev < 0 || (ndev > 0 && !devices)) return fail(PSS_ERR_ARG, "bad device list");
    std::vector<int> list;
    const int have = pss_device_count();
    for (int32_t i = 0; i < ndev; ++i) {
        if (devices[i] < 0 || (have > 0 && devices[i] >= have)) return fail(PSS_ERR_ARG, "device index out of range");
        list.push_back(devices[i]);
    }
    return reader_open_multi(index_file_path, list, out);
}

int32_t pss_reader_open_sharded(const char *index_file_path, int32_t shard_rank, int32_t shard_count,
                                pss_reader **out) {
    return reader_open(index_file_path, shard_rank, shard_count, -1, out);
}

int32_t pss_reader_open_device_chunks(const pss_device_chunk *chunks, int32_t n_local, int32_t n_chunks_total,
                                      int32_t device, pss_reader **out) {
    if (!out || n_local < 0 || n_chunks_total < n_local || (n_local > 0 && !chunks)) return fail(PSS_ERR_ARG, "bad chunk list");
    *out = nullptr;
    std::unique_ptr<pss_reader> r(new (std::nothrow) pss_reader());
    if (!r) return fail(PSS_ERR_NOMEM, "out of host memory");
    r->path = "<device chunks>";
    r->chunks.resize((size_t)n_chunks_total);
    int32_t prev = -1;
    for (int32_t i = 0; i < n_local; ++i) {
        const pss_device_chunk &in = chunks[i];
        if (in.global_id <= prev || in.global_id >= n_chunks_total) return fail(PSS_ERR_ARG, "chunk ids must be ascending and < n_chunks_total");
        if (in.n >= (1u << 30) || (in.n > 0 && (!in.d_text || !in.d_sa))) return fail(PSS_ERR_ARG, "bad device chunk");
        prev = in.global_id;
        ChunkHost &c = r->chunks[(size_t)in.global_id];
        c.n = in.n;
        c.sa_bytes = (uint64_t)in.n * 4;
        c.owned = true;
        c.borrowed = true;
        c.h_text = in.h_text;
        c.d_text = const_cast<uint8_t *>(in.d_text);
        c.d_sa   = const_cast<int32_t *>(in.d_sa);
    }
    PSS_TRY(reader_init_gpu(r.get(), device));
    PSS_TRY(reader_finish_open(r.get()));
    *out = r.release();
    return PSS_OK;
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
    *text = c.owned ? c.h_text : nullptr;
    *len  = (c.owned && c.h_text) ? (int64_t)c.n : 0;
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
    res->pool = r->pool;
    int64_t total = 0;
    for (size_t g = 0; g < G; ++g) total += part[g]->n_entries;
    PSS_TRY(res->alloc(nq, total));
    std::vector<int64_t> cur(G);
    int64_t at = 0;
    res->query_off()[0] = 0;
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
        res->query_off()[q + 1] = at;
    }
    for (size_t g = 0; g < G; ++g) {
        res->pub.n_hits += part[g]->n_hits;
        res->pub.ms_bounds  = std::max(res->pub.ms_bounds, part[g]->ms_bounds);
        res->pub.ms_extract = std::max(res->pub.ms_extract, part[g]->ms_extract);
        res->pub.ms_dedup   = std::max(res->pub.ms_dedup, part[g]->ms_dedup);
        res->pub.ms_total   = std::max(res->pub.ms_total, part[g]->ms_total);
    }
    res->pub.n_ranks = (int32_t)G;
    res->publish(at);
    *out = &res.release()->pub;
    return PSS_OK;
}

int32_t pss_reader_search_batch(pss_reader *r, const uint8_t *patterns, const int64_t *offsets, int32_t nq,
                                pss_result **out) {
    if (!r || !out || nq < 0 || (nq > 0 && !offsets)) return fail(PSS_ERR_ARG, "bad search arguments");
    *out = nullptr;
    PSS_TRY(validate_offsets(offsets, nq));
    const int64_t total = nq > 0 ? offsets[nq] : 0;
    if (total > 0 && !patterns) return fail(PSS_ERR_ARG, "bad pattern buffer");
    if (!r->subs.empty()) return search_batch_multi(r, patterns, offsets, nq, out);
    std::unique_ptr<ResultOwner> res(new (std::nothrow) ResultOwner());
    if (!res) return fail(PSS_ERR_NOMEM, "out of host memory");
    std::memset(&res->pub, 0, sizeof(res->pub));
    res->pool = r->pool;
    res->pub.n_ranks = 1;
    const int nc = r->searcher.num_chunks();
    if (nq == 0 || nc == 0) {
        PSS_TRY(res->alloc(nq, 0));
        std::memset(res->query_off(), 0, ((size_t)nq + 1) * 8);
        res->publish(0);
        *out = &res.release()->pub;
        return PSS_OK;
    }
    DeviceGuard dev_guard;
    PSS_CUDA_TRY(cudaSetDevice(r->searcher.device()));
    SearchOutput so;
    SearchTimes times;
    // ---- a handful of pairs (Reader.search): one launch, result read from mapped memory ----
    bool handled = false;
    PSS_TRY(r->searcher.search_small(patterns, offsets, nq, &so, &times, &handled));
    if (handled) {
        PSS_TRY(res->alloc(nq, so.n_entries));
        std::memcpy(res->query_off(), so.d_query_off, ((size_t)nq + 1) * 8);
        if (so.n_entries) {
            std::memcpy(res->chunk(), so.d_chunk, (size_t)so.n_entries * 4);
            std::memcpy(res->start(), so.d_start, (size_t)so.n_entries * 4);
            std::memcpy(res->end(), so.d_end, (size_t)so.n_entries * 4);
        }
    } else {
        cudaStream_t s = r->searcher.stream();
        // stage offsets + patterns in pinned memory, ONE H2D
        const size_t off_bytes = ((size_t)nq + 1) * sizeof(int64_t);
        PSS_TRY(r->ensure_patterns(off_bytes + (size_t)total));
        std::memcpy(r->h_pat.p, offsets, off_bytes);
        if (total) std::memcpy(static_cast<uint8_t *>(r->h_pat.p) + off_bytes, patterns, (size_t)total);
        PSS_CUDA_TRY(cudaEventRecord(r->ev0, s));
        PSS_CUDA_TRY(cudaMemcpyAsync(r->d_pat, r->h_pat.p, off_bytes + (size_t)total, cudaMemcpyHostToDevice, s));
        PSS_TRY(r->searcher.search(r->d_pat + off_bytes, reinterpret_cast<const int64_t *>(r->d_pat), nq, s, &so, &times));
        PSS_TRY(res->alloc(nq, so.n_entries));
        PSS_CUDA_TRY(cudaMemcpyAsync(res->query_off(), so.d_query_off, ((size_t)nq + 1) * 8, cudaMemcpyDeviceToHost, s));
        if (so.n_entries) {
            const size_t b = (size_t)so.n_entries * 4;
            PSS_CUDA_TRY(cudaMemcpyAsync(res->chunk(), so.d_chunk, b, cudaMemcpyDeviceToHost, s));
            PSS_CUDA_TRY(cudaMemcpyAsync(res->start(), so.d_start, b, cudaMemcpyDeviceToHost, s));
            PSS_CUDA_TRY(cudaMemcpyAsync(res->end(), so.d_end, b, cudaMemcpyDeviceToHost, s));
        }
        PSS_CUDA_TRY(cudaEventRecord(r->ev1, s));
        PSS_CUDA_TRY(cudaEventSynchronize(r->ev1));
        PSS_CUDA_TRY(cudaEventElapsedTime(&times.ms_total, r->ev0, r->ev1));
    }
    if (res->query_off()[nq] != so.n_entries)
        return fail(PSS_ERR_CUDA, "internal error: per-query counts do not add up to the entry count");
    res->pub.n_hits     = so.n_hits;
    res->pub.ms_bounds  = times.ms_bounds;
    res->pub.ms_extract = times.ms_extract;
    res->pub.ms_dedup   = times.ms_dedup;
    res->pub.ms_total   = times.ms_total;
    res->publish(so.n_entries);
    *out = &res.release()->pub;
    return PSS_OK;
}

int32_t pss_reader_search_batch_device(pss_reader *r, const uint8_t *d_patterns, const int64_t *d_offsets,
                                       int32_t nq, int64_t total_pattern_bytes, pss_device_result *out, void *stream) {
    (void)total_pattern_bytes;
    if (!r || nq < 0 || !out) return fail(PSS_ERR_ARG, "bad search arguments");
    std::memset(out, 0, sizeof(*out));
    if (!r->subs.empty())
        return fail(PSS_ERR_ARG, "device-resident search needs a single-device reader (no device list / PSS_DEVICES)");
    DeviceGuard dev_guard;
    SearchOutput so;
    SearchTimes times;
    PSS_TRY(r->searcher.search(d_patterns, d_offsets, nq, static_cast<cudaStream_t>(stream), &so, &times));
    out->n_queries       = nq;
    out->n_chunks        = r->searcher.num_chunks();
    out->n_entries       = so.n_entries;
    out->n_hits          = so.n_hits;
    out->d_query_offsets = so.d_query_off;
    out->d_entry_offsets = so.d_entry_off;
    out->d_chunk_id      = so.d_chunk;
    out->d_line_start    = so.d_start;
    out->d_line_end      = so.d_end;
    out->ms_bounds       = times.ms_bounds;
    out->ms_extract      = times.ms_extract;
    out->ms_dedup        = times.ms_dedup;
    return PSS_OK;
}

void pss_result_free(pss_result *res) {
    if (!res) return;
    // pub is the first member of ResultOwner
    delete reinterpret_cast<ResultOwner *>(res);
}

}  // extern "C"
