// common.cuh — shared helpers for libpss_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <atomic>

#include "../../include/pss.h"

namespace pss {

// ---- error plumbing -----------------------------------------------------------------
void set_error(const std::string &msg);
int  fail(int code, const std::string &msg);

extern std::atomic<long long> g_kernel_launches;
inline void count_launch(int k = 1) { g_kernel_launches.fetch_add(k, std::memory_order_relaxed); }

#define PSS_CUDA_TRY(expr)                                                                 \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            return ::pss::fail(_e == cudaErrorMemoryAllocation ? PSS_ERR_NOMEM : PSS_ERR_CUDA, \
                               std::string(#expr) + ": " + cudaGetErrorString(_e));        \
        }                                                                                  \
    } while (0)

#define PSS_TRY(expr)                                                                      \
    do {                                                                                   \
        int _rc = (expr);                                                                  \
        if (_rc != PSS_OK) return _rc;                                                     \
    } while (0)

// Check the launch of the kernel just enqueued.
#define PSS_LAUNCH_CHECK()                                                                 \
    do {                                                                                   \
        ::pss::count_launch();                                                             \
        PSS_CUDA_TRY(cudaGetLastError());                                                  \
    } while (0)

int default_device();
int sm_count(int device);

// Copies a few 32-bit words between device memory and MAPPED pinned host memory with a tiny
// kernel instead of cudaMemcpyAsync: no copy engine is involved, so the scalar read-backs of
// a build do not queue behind the multi-gigabyte D2H / H2D of the neighbouring chunk (measured:
// a build that overlapped a 2 GiB suffix-array copy took 137 ms instead of 96 ms because its
// first 32-byte read-back waited for that copy on the shared D2H engine).
int copy_words(uint32_t *dst, const uint32_t *src, int n, cudaStream_t stream);
int alloc_mapped_words(uint32_t **p, int n);   // cudaHostAlloc(..., cudaHostAllocMapped), zeroed

// ---- device helpers -----------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Streaming (read-once) loads: keep them out of L1 so the staging tiles own it.
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_stream_u128(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

// L2 eviction-priority hints (createpolicy + ld/st .L2::cache_hint).  The rank gather / scatter
// of prefix doubling touches 4 bytes of a window of ISA that fits in L2 while gigabytes of
// records stream past: without hints the streams (and the 128-byte fills of the random
// misses themselves) turn L2 over faster than the window is reused (ncu: every ISA line was
// fetched from DRAM ~18 times).  Window data is marked evict_last, streams evict_first.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint32_t ld_nc_u32_hint(const uint32_t *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream_u32_hint(const uint32_t *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint4 ld_stream_u128_hint(const uint4 *p, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_u32_hint(uint32_t *p, uint32_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_u64_hint(uint64_t *p, uint64_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}

static inline int bit_width_u64(uint64_t v) {
    int b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

static inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace pss
