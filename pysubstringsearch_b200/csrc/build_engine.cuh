// build_engine.cuh — asynchronous suffix-array builds, one engine per GPU.
//
// The seam SURVEY §8(b) asks for next to the synchronous `libsais` drop-in:
//     pss_sa_build_begin(device, T, n) -> handle      returns at once
//     pss_sa_build_wait(handle, SA)                   blocks, then copies the SA out
// so that a Writer can keep ingesting chunk k+1 while chunk k is being built
// (src/lib.rs:67-124 blocks on libsais inside dump_data) and can spread chunks over GPUs.
//
// Per device: one SaBuilder (the 32 B/byte sort workspace), one worker thread that runs the
// builds in request order, and THREE (text, SA) slots in HBM: while chunk k is being built,
// the text of chunk k+1 is already arriving (separate H2D stream, issued before build k
// starts) and the suffix array of chunk k-1 is leaving (D2H by whoever calls wait, on a third
// stream).  Engines are process-wide and cached: pss_libsais, the Writers and the async
// C ABI all share them; pss_release_cached() frees the idle ones.
#pragma once

#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>

#include "common.cuh"
#include "sa_build.cuh"

namespace pss {

class BuildEngine {
public:
    struct Job {
        const uint8_t *h_text = nullptr;
        int32_t        n = 0;
        int            slot = -1;     // assigned when the text upload is issued
        bool           staged = false;
        int            state = 0;     // 0 queued, 1 building, 2 built (or failed: rc != 0)
        int            rc = PSS_OK;
        std::string    err;
        float          build_ms = 0.f;   // CUDA-event time of the device build
        BuildEngine   *engine = nullptr;
    };

    // The engine of `device` (-1 = the default device), created on first use.
    static int  get(int device, BuildEngine **out);
    // Frees the workspaces of all engines with nothing in flight.
    static void release_idle();

    // Queues a build of h_text[0..n); the text must stay valid and unchanged until wait()
    // returns.  Never blocks.
    int begin(const uint8_t *h_text, int32_t n, Job **out);
    // Blocks until the build is done, copies the suffix array to h_sa[0..n) (pinned memory
    // is written by DMA directly, pageable memory through the stager) and frees the job.
    // With three slots per device, at most three builds per device can be past their H2D at
    // any time: wait for them in begin order.
    int wait(Job *job, int32_t *h_sa);

    int device() const { return device_; }

private:
    BuildEngine() = default;
    int  init(int device);
    void worker();
    void free_device_memory();

    struct Slot {
        uint8_t    *d_text = nullptr;
        int32_t    *d_sa = nullptr;
        int64_t     cap = 0;
        bool        busy = false;
        cudaEvent_t uploaded = nullptr;    // the text of the slot's job has arrived
    };
    static constexpr int NSLOTS = 3;

    int  free_slot() const;                         // call with mu_ held; -1 if none
    int  stage(Job *job, int si);                   // slot buffers + text upload (mu_ NOT held)

    int          device_ = -1;
    SaBuilder    builder_;
    Slot         slots_[NSLOTS];
    cudaStream_t h2d_stream_ = nullptr;    // text uploads (ahead of the build that needs them)
    cudaStream_t copy_stream_ = nullptr;   // D2H of finished suffix arrays
    bool         trace_ = false;           // PSS_ENGINE_TRACE=1: per-job timings on stderr
    HostStager   h2d_stager_, d2h_stager_;
    std::mutex   d2h_mu_;                  // one wait() copies out at a time (d2h_stager_ is shared)

    std::mutex              mu_;
    std::condition_variable cv_;
    std::deque<Job *>       queue_;
    int                     in_flight_ = 0;   // begun and not yet waited
    bool                    stop_ = false, started_ = false;
    std::thread             thread_;
};

// RAII: restores the calling thread's current device (API entry points must not leave it changed).
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace pss
