// build_engine.cu — asynchronous per-GPU build engines (see build_engine.cuh).
#include "build_engine.cuh"

#include <chrono>
#include <cstdlib>
#include <map>
#include <new>

namespace pss {

namespace {
std::mutex g_engines_mu;
std::map<int, BuildEngine *> g_engines;   // never destroyed at exit: the CUDA context may be gone by then
}  // namespace

int BuildEngine::get(int device, BuildEngine **out) {
    *out = nullptr;
    if (device < 0) device = default_device();
    std::lock_guard<std::mutex> lock(g_engines_mu);
    auto it = g_engines.find(device);
    if (it != g_engines.end()) {
        *out = it->second;
        return PSS_OK;
    }
    BuildEngine *e = new (std::nothrow) BuildEngine();
    if (!e) return fail(PSS_ERR_NOMEM, "out of host memory");
    int rc = e->init(device);
    if (rc != PSS_OK) {
        delete e;
        return rc;
    }
    g_engines[device] = e;
    *out = e;
    return PSS_OK;
}

void BuildEngine::release_idle() {
    std::lock_guard<std::mutex> lock(g_engines_mu);
    for (auto &kv : g_engines) {
        BuildEngine *e = kv.second;
        std::lock_guard<std::mutex> l2(e->mu_);
        if (e->in_flight_ == 0) e->free_device_memory();
    }
}

int BuildEngine::init(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(PSS_ERR_CUDA, "no CUDA device available (libpss_b200 has no CPU fallback)");
    if (device >= ndev) return fail(PSS_ERR_ARG, "device index out of range");
    DeviceGuard guard;
    PSS_TRY(builder_.init(device, 0));
    device_ = device;
    PSS_CUDA_TRY(cudaSetDevice(device_));
    PSS_CUDA_TRY(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    PSS_CUDA_TRY(cudaStreamCreateWithFlags(&h2d_stream_, cudaStreamNonBlocking));
    for (Slot &s : slots_) PSS_CUDA_TRY(cudaEventCreateWithFlags(&s.uploaded, cudaEventDisableTiming));
    if (const char *e = std::getenv("PSS_ENGINE_TRACE")) trace_ = std::atoi(e) != 0;
    return PSS_OK;
}

// Called with mu_ held and nothing in flight.
void BuildEngine::free_device_memory() {
    DeviceGuard guard;
    cudaSetDevice(device_);
    for (Slot &s : slots_) {
        cudaFree(s.d_text);
        cudaFree(s.d_sa);
        s.d_text = nullptr; s.d_sa = nullptr; s.cap = 0;
    }
    h2d_stager_.release();
    d2h_stager_.release();
    builder_.release_workspace();
}

int BuildEngine::begin(const uint8_t *h_text, int32_t n, Job **out) {
    if (!out) return fail(PSS_ERR_ARG, "null out pointer");
    *out = nullptr;
    if (n < 0 || (n > 0 && !h_text)) return fail(PSS_ERR_ARG, "bad build arguments");
    if ((int64_t)n >= (1ll << 30)) return fail(PSS_ERR_ARG, "n must be < 2^30 (container stores 4n in a u32)");
    Job *job = new (std::nothrow) Job();
    if (!job) return fail(PSS_ERR_NOMEM, "out of host memory");
    job->h_text = h_text;
    job->n      = n;
    job->engine = this;
    {
        std::lock_guard<std::mutex> lock(mu_);
        if (!started_) {
            try {
                thread_ = std::thread([this] { worker(); });
            } catch (...) {
                delete job;
                return fail(PSS_ERR_NOMEM, "cannot start the build worker thread");
            }
            started_ = true;
        }
        queue_.push_back(job);
        ++in_flight_;
    }
    cv_.notify_all();
    *out = job;
    return PSS_OK;
}

int BuildEngine::free_slot() const {
    for (int i = 0; i < NSLOTS; ++i)
        if (!slots_[i].busy) return i;
    return -1;
}

// Makes slot `si` large enough for the job and issues the upload of its text on the H2D
// stream; the build waits for slot.uploaded.  Pinned text: one asynchronous DMA.  Pageable
// text: bounced through the stager by this thread (returns when the host buffer is consumed).
int BuildEngine::stage(Job *job, int si) {
    Slot &slot = slots_[si];
    const int64_t n = job->n;
    if (n > slot.cap) {
        cudaFree(slot.d_text);
        cudaFree(slot.d_sa);
        slot.d_text = nullptr; slot.d_sa = nullptr; slot.cap = 0;
        const int64_t cap = std::max<int64_t>(n, 1 << 16);
        cudaError_t e = cudaMalloc(&slot.d_text, (size_t)cap + 64);
        if (e == cudaSuccess) e = cudaMalloc(&slot.d_sa, (size_t)cap * sizeof(int32_t));
        if (e != cudaSuccess) {
            cudaFree(slot.d_text);
            slot.d_text = nullptr;
            return fail(e == cudaErrorMemoryAllocation ? PSS_ERR_NOMEM : PSS_ERR_CUDA,
                        std::string("device buffers for the chunk: ") + cudaGetErrorString(e));
        }
        slot.cap = cap;
    }
    if (n > 0) PSS_TRY(h2d_stager_.copy(slot.d_text, job->h_text, (size_t)n, /*to_device=*/true, h2d_stream_));
    PSS_CUDA_TRY(cudaEventRecord(slot.uploaded, h2d_stream_));
    return PSS_OK;
}

void BuildEngine::worker() {
    cudaSetDevice(device_);
    std::unique_lock<std::mutex> lock(mu_);
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    while (true) {
        cv_.wait(lock, [&] { return stop_ || !queue_.empty(); });
        if (queue_.empty()) break;
        Job *job = queue_.front();
        int rc = PSS_OK;
        std::string err;
        if (!job->staged) {
            // a free (text, SA) slot: freed by wait() of an earlier job
            cv_.wait(lock, [&] { return stop_ || free_slot() >= 0; });
            if (free_slot() < 0) break;
            job->slot = free_slot();
            slots_[job->slot].busy = true;
            lock.unlock();
            rc = stage(job, job->slot);
            if (rc != PSS_OK) err = pss_last_error();
            lock.lock();
            job->staged = true;
        }
        queue_.pop_front();
        job->state = 1;
        // the next chunk's text starts arriving now, while this one is built (pinned sources
        // only: bouncing a pageable text would hold this thread for tens of milliseconds)
        Job *next = queue_.empty() ? nullptr : queue_.front();
        int next_slot = -1;
        if (next && !next->staged && free_slot() >= 0 && host_is_pinned(next->h_text)) {
            next_slot = free_slot();
            slots_[next_slot].busy = true;
            next->slot = next_slot;
        }
        lock.unlock();

        const double t0 = now_ms();
        if (next_slot >= 0) {
            int rc2 = stage(next, next_slot);      // asynchronous; a failure is reported by that job
            lock.lock();
            next->staged = true;
            if (rc2 != PSS_OK) { next->rc = rc2; next->err = pss_last_error(); }
            lock.unlock();
        }
        if (rc == PSS_OK && job->rc != PSS_OK) { rc = job->rc; err = job->err; }   // its own prefetch failed earlier
        Slot &slot = slots_[job->slot];
        if (rc == PSS_OK && job->n > 0) {
            cudaError_t e = cudaStreamWaitEvent(builder_.stream(), slot.uploaded, 0);
            if (e != cudaSuccess) {
                rc = fail(PSS_ERR_CUDA, std::string("cudaStreamWaitEvent: ") + cudaGetErrorString(e));
            } else {
                rc = builder_.build_device(slot.d_text, job->n, slot.d_sa, builder_.stream());
            }
            if (rc != PSS_OK) err = pss_last_error();
        }
        if (trace_)
            fprintf(stderr, "[pss engine %d] job n=%d slot=%d: staged+built in %.1f ms (device build %.1f ms)%s\n", device_,
                    job->n, job->slot, now_ms() - t0, builder_.stats().total_ms, next_slot >= 0 ? " +prefetch" : "");

        lock.lock();
        job->rc       = rc;
        job->err      = err;
        job->build_ms = builder_.stats().total_ms;
        job->state    = 2;
        cv_.notify_all();
    }
}

int BuildEngine::wait(Job *job, int32_t *h_sa) {
    if (!job) return fail(PSS_ERR_ARG, "null build handle");
    {
        std::unique_lock<std::mutex> lock(mu_);
        cv_.wait(lock, [&] { return job->state == 2; });
    }
    int rc = job->rc;
    std::string err = job->err;
    if (rc == PSS_OK && job->n > 0) {
        if (!h_sa) {
            rc  = PSS_ERR_ARG;
            err = "null suffix array pointer";
        } else {
            DeviceGuard guard;
            std::lock_guard<std::mutex> d2h(d2h_mu_);
            cudaSetDevice(device_);
            const auto t0 = std::chrono::steady_clock::now();
            rc = d2h_stager_.copy(h_sa, slots_[job->slot].d_sa, (size_t)job->n * sizeof(int32_t), /*to_device=*/false,
                                  copy_stream_);
            if (rc != PSS_OK) err = pss_last_error();
            if (trace_)
                fprintf(stderr, "[pss engine %d] job n=%d slot=%d: suffix array copied out in %.1f ms\n", device_, job->n,
                        job->slot, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        }
    }
    {
        std::lock_guard<std::mutex> lock(mu_);
        if (job->slot >= 0) slots_[job->slot].busy = false;
        --in_flight_;
    }
    cv_.notify_all();
    delete job;
    if (rc != PSS_OK) return fail(rc, err);
    return PSS_OK;
}

}  // namespace pss
