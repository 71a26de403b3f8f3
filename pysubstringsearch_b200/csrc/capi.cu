// capi.cu — extern "C" surface of libpss_b200.so for the BUILD path (declared in
// include/pss.h), plus the process-wide error / device plumbing.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "build_engine.cuh"
#include "common.cuh"
#include "radix_sort.cuh"
#include "sa_build.cuh"

namespace pss {

static thread_local std::string t_last_error;
std::atomic<long long> g_kernel_launches{0};
static std::atomic<int> g_device{-1};

void set_error(const std::string &msg) { t_last_error = msg; }
int fail(int code, const std::string &msg) {
    t_last_error = msg;
    return code;
}

int default_device() {
    int d = g_device.load();
    if (d >= 0) return d;
    const char *e = std::getenv("PSS_DEVICE");
    if (e && *e) return std::atoi(e);
    // One process per GPU launched by torchrun: CUDA_VISIBLE_DEVICES is normally NOT
    // narrowed, so the local rank picks the device.
    e = std::getenv("LOCAL_RANK");
    if (e && *e) {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) return std::atoi(e) % ndev;
    }
    return 0;
}

int sm_count(int device) {
    static int cached[64] = {0};
    if (device >= 0 && device < 64 && cached[device]) return cached[device];
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
    if (device >= 0 && device < 64) cached[device] = v;
    return v;
}

}  // namespace pss

using namespace pss;

struct pss_sa_builder {
    SaBuilder impl;
};

extern "C" {

const char *pss_last_error(void) { return t_last_error.c_str(); }

const char *pss_version(void) { return "pss_b200 0.2 (sm_100a)"; }

int32_t pss_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int32_t pss_set_device(int32_t device) {
    int n = pss_device_count();
    if (device < 0 || device >= n) return fail(PSS_ERR_ARG, "device index out of range");
    g_device.store(device);
    return PSS_OK;
}

int32_t pss_get_device(void) { return default_device(); }

int64_t pss_kernel_launch_count(void) { return g_kernel_launches.load(); }

int32_t pss_libsais(const uint8_t *T, int32_t *SA, int32_t n, int32_t fs, int32_t *freq) {
    // Argument contract of libsais (src/libsais/libsais.c:6599-6607).
    if (T == nullptr || SA == nullptr || n < 0 || fs < 0) return fail(PSS_ERR_ARG, "libsais: bad arguments");
    if (freq) {
        std::memset(freq, 0, 256 * sizeof(int32_t));
        for (int32_t i = 0; i < n; ++i) freq[T[i]]++;
    }
    if (n == 0) return PSS_OK;
    // the synchronous seam is the asynchronous one waited for at once (same cached engine)
    BuildEngine *engine = nullptr;
    PSS_TRY(BuildEngine::get(-1, &engine));
    BuildEngine::Job *job = nullptr;
    PSS_TRY(engine->begin(T, n, &job));
    return engine->wait(job, SA);
}

int32_t pss_sa_build_begin(int32_t device, const uint8_t *T, int32_t n, pss_sa_build **out) {
    if (!out) return fail(PSS_ERR_ARG, "null out pointer");
    *out = nullptr;
    if (n < 0 || (n > 0 && !T)) return fail(PSS_ERR_ARG, "bad build arguments");
    BuildEngine *engine = nullptr;
    PSS_TRY(BuildEngine::get(device, &engine));
    BuildEngine::Job *job = nullptr;
    PSS_TRY(engine->begin(T, n, &job));
    *out = reinterpret_cast<pss_sa_build *>(job);
    return PSS_OK;
}

int32_t pss_sa_build_wait(pss_sa_build *h, int32_t *SA) {
    if (!h) return fail(PSS_ERR_ARG, "null build handle");
    BuildEngine::Job *job = reinterpret_cast<BuildEngine::Job *>(h);
    return job->engine->wait(job, SA);
}

int32_t pss_memcpy_d2h(void *dst, const void *d_src, size_t bytes) {
    if (bytes && (!dst || !d_src)) return fail(PSS_ERR_ARG, "null pointer");
    if (bytes) PSS_CUDA_TRY(cudaMemcpy(dst, d_src, bytes, cudaMemcpyDeviceToHost));
    return PSS_OK;
}

int32_t pss_release_cached(void) {
    BuildEngine::release_idle();
    return PSS_OK;
}

int32_t pss_sa_builder_create(int32_t device, int64_t max_n, pss_sa_builder **out) {
    if (!out) return fail(PSS_ERR_ARG, "null out pointer");
    *out = nullptr;
    pss_sa_builder *b = new (std::nothrow) pss_sa_builder();
    if (!b) return fail(PSS_ERR_NOMEM, "out of host memory");
    int rc = b->impl.init(device, max_n);
    if (rc != PSS_OK) {
        delete b;
        return rc;
    }
    *out = b;
    return PSS_OK;
}

void pss_sa_builder_destroy(pss_sa_builder *b) { delete b; }

int32_t pss_sa_builder_set_profiling(pss_sa_builder *b, int32_t on) {
    if (!b) return fail(PSS_ERR_ARG, "null builder");
    b->impl.set_profiling(on != 0);
    return PSS_OK;
}

int32_t pss_sa_builder_build_device(pss_sa_builder *b, const uint8_t *d_text, int32_t n, int32_t *d_sa,
                                    void *stream) {
    if (!b) return fail(PSS_ERR_ARG, "null builder");
    return b->impl.build_device(d_text, n, d_sa, static_cast<cudaStream_t>(stream));
}

int32_t pss_sa_builder_build_host(pss_sa_builder *b, const uint8_t *h_text, int32_t n, int32_t *h_sa) {
    if (!b) return fail(PSS_ERR_ARG, "null builder");
    return b->impl.build_host(h_text, n, h_sa);
}

int32_t pss_sa_builder_stats(pss_sa_builder *b, pss_build_stats *stats, pss_pass_stat *pass_stats) {
    if (!b || !stats) return fail(PSS_ERR_ARG, "null argument");
    *stats = b->impl.stats();
    if (pass_stats) {
        const auto &v = b->impl.pass_stats();
        for (size_t i = 0; i < v.size() && i < PSS_MAX_PASS_STATS; ++i) pass_stats[i] = v[i];
    }
    return PSS_OK;
}

int32_t pss_radix_sort_pairs(uint64_t *d_keys, uint64_t *d_keys_alt, uint32_t *d_vals, uint32_t *d_vals_alt,
                             int64_t n, int32_t begin_bit, int32_t end_bit, int32_t *result_in_alt,
                             float *pass_ms, int32_t *n_passes, void *stream) {
    if (n < 0 || !d_keys || !d_keys_alt || !d_vals_alt || !result_in_alt)
        return fail(PSS_ERR_ARG, "radix sort: null argument");
    if (n >= (1ll << 30)) return fail(PSS_ERR_ARG, "radix sort: n must be < 2^30");
    static std::mutex mu;
    static RadixSorter *sorter = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    PSS_CUDA_TRY(cudaGetDevice(&dev));
    if (sorter && sorter->device() != dev) {
        delete sorter;
        sorter = nullptr;
    }
    if (!sorter) {
        sorter = new (std::nothrow) RadixSorter();
        if (!sorter) return fail(PSS_ERR_NOMEM, "out of host memory");
        int rc = sorter->init(dev);
        if (rc != PSS_OK) {
            delete sorter;
            sorter = nullptr;
            return rc;
        }
    }
    SortProfile prof;
    prof.timed = (pass_ms != nullptr);
    bool in_alt = false;
    // vals == NULL means "values are 0..n-1"; the sorter then needs a second value buffer
    // for its ping-pong, which the caller does not have: allocate a scratch one.
    uint32_t *scratch = nullptr;
    bool iota = (d_vals == nullptr);
    if (iota) {
        PSS_CUDA_TRY(cudaMalloc(&scratch, std::max<int64_t>(n, 1) * sizeof(uint32_t)));
        d_vals = scratch;
    }
    int rc = sorter->sort(d_keys, d_keys_alt, d_vals, d_vals_alt, (uint32_t)n, begin_bit, end_bit, iota,
                          static_cast<cudaStream_t>(stream), &in_alt, &prof);
    if (rc == PSS_OK && iota && !in_alt && n > 0) {
        // result values live in the scratch buffer: hand them back through vals_alt
        cudaError_t e = cudaMemcpyAsync(d_vals_alt, scratch, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                                        static_cast<cudaStream_t>(stream));
        if (e == cudaSuccess) e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
        if (e != cudaSuccess) rc = fail(PSS_ERR_CUDA, cudaGetErrorString(e));
    }
    if (scratch) cudaFree(scratch);
    if (rc != PSS_OK) return rc;
    *result_in_alt = in_alt ? 1 : 0;
    if (n_passes) *n_passes = prof.n_passes;
    if (pass_ms)
        for (int p = 0; p < prof.n_passes; ++p) pass_ms[p] = prof.ms[p];
    return PSS_OK;
}

}  // extern "C"
