// dist.cu — the one exchange step of the one-process-per-GPU search, inside the library:
// chunk k of the index lives on rank k % world; the query batch is broadcast from rank 0,
// every rank searches its own chunks, and rank 0 receives exactly the tuples found
// (gather-v over NCCL send/recv: no padding, no host round trip per rank) and places
// them into the single-process order (query, ascending chunk id, SA order) with one kernel.
//
// Reference: the rayon fan-out over chunks and the Mutex<Vec>::extend merge of
// Reader::search (src/lib.rs:205-207, 280-284) — there in one address space, here across
// the GPUs of an NVSwitch box.
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the library is bound at run time (see NcclApi)

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>

#include "build_engine.cuh"   // DeviceGuard
#include "common.cuh"
#include "reader.cuh"
#include "search.cuh"

using namespace pss;

// NCCL is bound with dlopen when the first communicator is created, not at link time: a
// process that also hosts PyTorch must end up with ONE libnccl.so.2, and PyTorch's bundled
// copy is newer than the system one (a link-time dependency on the system copy makes
// `import torch` fail after this library has been loaded).  Order: the copy already in the
// process, then $PSS_NCCL_LIB, then the system libnccl.so.2.
namespace {
struct NcclApi {
    decltype(&ncclGetUniqueId)    GetUniqueId = nullptr;
    decltype(&ncclCommInitRank)   CommInitRank = nullptr;
    decltype(&ncclCommDestroy)    CommDestroy = nullptr;
    decltype(&ncclBroadcast)      Broadcast = nullptr;
    decltype(&ncclAllGather)      AllGather = nullptr;
    decltype(&ncclAllReduce)      AllReduce = nullptr;
    decltype(&ncclSend)           Send = nullptr;
    decltype(&ncclRecv)           Recv = nullptr;
    decltype(&ncclGroupStart)     GroupStart = nullptr;
    decltype(&ncclGroupEnd)       GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
    std::string error;
};

const NcclApi &nccl() {
    static NcclApi api = [] {
        NcclApi a;
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) {
            const char *e = std::getenv("PSS_NCCL_LIB");
            if (e && *e) h = dlopen(e, RTLD_NOW | RTLD_LOCAL);
        }
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) {
            a.error = std::string("cannot load libnccl.so.2: ") + dlerror();
            return a;
        }
#define PSS_BIND(name)                                                              \
        a.name = reinterpret_cast<decltype(a.name)>(dlsym(h, "nccl" #name));        \
        if (!a.name) { a.error = "libnccl.so.2 lacks nccl" #name; return a; }
        PSS_BIND(GetUniqueId) PSS_BIND(CommInitRank) PSS_BIND(CommDestroy) PSS_BIND(Broadcast) PSS_BIND(Send)
        PSS_BIND(AllGather) PSS_BIND(AllReduce)
        PSS_BIND(Recv) PSS_BIND(GroupStart) PSS_BIND(GroupEnd) PSS_BIND(GetErrorString)
#undef PSS_BIND
        a.ok = true;
        return a;
    }();
    return api;
}
}  // namespace

#define PSS_NCCL_TRY(expr)                                                                     \
    do {                                                                                       \
        ncclResult_t _r = (expr);                                                              \
        if (_r != ncclSuccess)                                                                 \
            return ::pss::fail(PSS_ERR_CUDA, std::string(#expr) + ": " + nccl().GetErrorString(_r)); \
    } while (0)

static bool dist_trace() {
    static const bool on = [] { const char *e = std::getenv("PSS_DIST_TRACE"); return e && std::atoi(e) != 0; }();
    return on;
}
static double wall_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct pss_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    void *h_pinned = nullptr;     // descriptors + counts staging (pinned)
    size_t h_cap = 0;
    // Fused exchange: rank 0's result arrays [chunk | start | end], `final_cap` entries each, are
    // one cudaMalloc block exported with CUDA IPC and mapped by every other rank, whose
    // compaction kernels store their tuples straight into it over NVLink.
    int       fused = -1;             // -1 not tried yet, 0 unavailable (NCCL gather-v is used), 1 in use
    uint32_t *final_base = nullptr;   // rank 0: the block; other ranks: its mapping
    int64_t   final_cap = 0;
    uint32_t *d_small = nullptr;      // 64 words: handle broadcast, counts, barrier word
    uint32_t *h_small = nullptr;      // mapped pinned mirror
    pss::DeviceBuf all_e, f_me, desc_dev, qoff, pairs_all;
};

namespace {

// What rank 0 knows about rank r's part of the batch.
struct RankDesc {
    const uint32_t *E;        // [npairs + 1] entries of rank r before its pair p
    const uint32_t *start;    // [count]
    const uint32_t *end;      // [count]
    uint32_t       *F;        // [npairs] final offset of the first entry of pair p
    uint32_t        nc;       // chunks owned by rank r
    uint32_t        npairs;   // nq * nc
    uint32_t        count;    // entries found by rank r
    uint32_t        pad;
};

// chunks owned by rank r (of G) with global id < k, under the map chunk k -> rank k % G
__device__ __forceinline__ uint32_t chunks_before(uint32_t r, uint32_t k, uint32_t G) {
    return k > r ? (k - r - 1) / G + 1 : 0u;
}

__global__ void __launch_bounds__(32)
dist_counts_kernel(RankDesc *__restrict__ desc, int G, uint32_t *__restrict__ counts) {
    const int r = threadIdx.x;
    if (r < G) {
        const uint32_t c = desc[r].E[desc[r].npairs];
        desc[r].count = c;
        counts[r]     = c;
    }
}

// F_r[p]: position in the merged result of the first entry of rank r's pair p = entries of
// every rank in pairs that precede (query q, chunk k) in (query, chunk) order.
__global__ void __launch_bounds__(256)
dist_pair_final_kernel(const RankDesc *__restrict__ desc, int G) {
    const RankDesc d = desc[blockIdx.y];
    const uint32_t p = blockIdx.x * 256 + threadIdx.x;
    if (p >= d.npairs) return;
    const uint32_t q = p / d.nc, j = p % d.nc;
    const uint32_t k = j * (uint32_t)G + blockIdx.y;
    uint32_t sum = 0;
    for (int r2 = 0; r2 < G; ++r2) {
        const RankDesc o = desc[r2];
        if (o.nc) sum += __ldg(o.E + (size_t)q * o.nc + chunks_before((uint32_t)r2, k, (uint32_t)G));
    }
    d.F[p] = sum;
}

// Same as dist_pair_final_kernel for ONE rank's pairs (every rank computes where its own
// entries go once all entry-offset arrays have been all-gathered).
__global__ void __launch_bounds__(256)
dist_pair_final_one_kernel(const RankDesc *__restrict__ desc, int G, int me) {
    const RankDesc d = desc[me];
    const uint32_t p = blockIdx.x * 256 + threadIdx.x;
    if (p >= d.npairs) return;
    const uint32_t q = p / d.nc, j = p % d.nc;
    const uint32_t k = j * (uint32_t)G + (uint32_t)me;
    uint32_t sum = 0;
    for (int r2 = 0; r2 < G; ++r2) {
        const RankDesc o = desc[r2];
        if (o.nc) sum += __ldg(o.E + (size_t)q * o.nc + chunks_before((uint32_t)r2, k, (uint32_t)G));
    }
    d.F[p] = sum;
}

// P[q * n_total + k] = entries of the merged result before (query q, chunk k) — the same sum as
// F, laid out in (query, chunk) order: the merged result's per-pair entry offsets.
__global__ void __launch_bounds__(256)
dist_global_pairs_kernel(const RankDesc *__restrict__ desc, int G, uint32_t nq, uint32_t n_total, uint32_t *__restrict__ P) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i > nq * n_total) return;
    const uint32_t q = i / n_total, k = i % n_total;      // i == nq * n_total: q = nq, k = 0 → the total
    uint32_t sum = 0;
    for (int r2 = 0; r2 < G; ++r2) {
        const RankDesc o = desc[r2];
        if (o.nc) sum += __ldg(o.E + (size_t)q * o.nc + chunks_before((uint32_t)r2, k, (uint32_t)G));
    }
    P[i] = sum;
}

// largest p in [0, npairs) with E[p] <= i (pairs without entries repeat the offset and are skipped)
__device__ __forceinline__ uint32_t pair_of_entry(const uint32_t *__restrict__ E, uint32_t npairs, uint32_t i) {
    uint32_t lo = 0, hi = npairs;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(E + mid) <= i) lo = mid;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
dist_place_kernel(const RankDesc *__restrict__ desc, int G, int rank_base, int32_t *out_chunk,
                  uint32_t *out_start, uint32_t *out_end) {
    const RankDesc d = desc[rank_base + blockIdx.y];
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    const uint32_t i0 = i & ~31u;                 // the warp's first entry
    if (i0 >= d.count) return;
    // One binary search per warp (every lane issues the same loads: one transaction each),
    // then each lane walks forward from the warp's pair: a warp's 32 consecutive entries span
    // few pairs.  Long runs of empty pairs fall back to the lane's own search.
    uint32_t p = pair_of_entry(d.E, d.npairs, i0);
    if (i >= d.count) return;
    int steps = 0;
    while (p + 1 < d.npairs && __ldg(d.E + p + 1) <= i && steps < 16) { ++p; ++steps; }
    if (p + 1 < d.npairs && __ldg(d.E + p + 1) <= i) p = pair_of_entry(d.E, d.npairs, i);
    const uint32_t dst = d.F[p] + (i - __ldg(d.E + p));
    out_chunk[dst] = (int32_t)((p % d.nc) * (uint32_t)G + (uint32_t)rank_base + blockIdx.y);
    out_start[dst] = d.start[i];
    out_end[dst]   = d.end[i];
}

// chunk id of every merged entry from the global pair offsets (rank 0 fills this array itself
// while the other ranks' tuples arrive: a third less NVLink traffic).  One warp per
// (query, chunk) pair: its entries are one contiguous run.
__global__ void __launch_bounds__(256)
dist_fill_chunk_kernel(const uint32_t *__restrict__ P, uint32_t npairs, uint32_t n_total, int32_t *__restrict__ out_chunk) {
    const uint32_t p = (blockIdx.x * 256 + threadIdx.x) >> 5;
    if (p >= npairs) return;
    const uint32_t a = __ldg(P + p), b = __ldg(P + p + 1);
    const int32_t chunk = (int32_t)(p % n_total);
    for (uint32_t i = a + (threadIdx.x & 31u); i < b; i += 32) out_chunk[i] = chunk;
}

__global__ void __launch_bounds__(256)
dist_query_off_kernel(const RankDesc *__restrict__ desc, int G, uint32_t nq, int64_t *__restrict__ qoff) {
    const uint32_t q = blockIdx.x * 256 + threadIdx.x;
    if (q > nq) return;
    int64_t sum = 0;
    for (int r = 0; r < G; ++r) sum += (int64_t)__ldg(desc[r].E + (size_t)q * desc[r].nc);
    qoff[q] = sum;
}

struct DistOut {
    int64_t n_entries = 0, n_hits = 0;
    const int64_t  *d_qoff = nullptr;
    const uint32_t *d_entry_off = nullptr;
    const int32_t  *d_chunk = nullptr;
    const uint32_t *d_start = nullptr, *d_end = nullptr;
    int32_t n_chunks = 0;
    SearchTimes times;
    float ms_exchange = 0.f;
};

uint32_t owned_chunks(uint32_t rank, uint32_t world, uint32_t n_total) {
    return rank < n_total ? (n_total - rank + world - 1) / world : 0u;
}

int dist_search_fused(pss_reader *r, pss_comm *c, int32_t nq, int64_t total, DistOut *out);
int fused_setup(pss_comm *c, cudaStream_t s);

// The whole collective.  On entry the batch is in r->d_pat on rank 0 ([offsets | bytes],
// enqueued on the reader's stream); on exit rank 0 holds the merged result on the device.
int dist_search_core(pss_reader *r, pss_comm *c, int32_t nq, int64_t total, DistOut *out) {
    const int G = c->world, me = c->rank;
    cudaStream_t s = r->searcher.stream();
    const size_t off_bytes = ((size_t)nq + 1) * sizeof(int64_t);
    const uint32_t n_total = (uint32_t)r->chunks.size();
    const int nc_me = r->searcher.num_chunks();
    if ((uint32_t)nc_me != owned_chunks((uint32_t)me, (uint32_t)G, n_total))
        return fail(PSS_ERR_ARG, "distributed search: this reader does not own chunks k with k % world == rank "
                                 "(open it with pss_reader_open_sharded(path, rank, world))");
    for (int j = 0; j < nc_me; ++j)
        if (r->searcher.chunks()[j].global_id != j * G + me)
            return fail(PSS_ERR_ARG, "distributed search: chunk map is not chunk k -> rank k % world");

    if (G > 1) {
        if (c->fused < 0) PSS_TRY(fused_setup(c, s));      // collective, first batch only
        if (c->fused == 1) return dist_search_fused(r, c, nq, total, out);
    }
    const double tw0 = wall_ms();
    PSS_CUDA_TRY(cudaEventRecord(r->ev0, s));
    // ---- 1. the batch reaches every GPU -------------------------------------------------
    if (G > 1) PSS_NCCL_TRY(nccl().Broadcast(r->d_pat, r->d_pat, off_bytes + (size_t)total, ncclUint8, 0, c->comm, s));

    // ---- 2. local search ---------------------------------------------------------------------
    SearchOutput so;
    PSS_TRY(r->searcher.search(r->d_pat + off_bytes, reinterpret_cast<const int64_t *>(r->d_pat), nq, s, &so, &out->times));
    out->n_hits = so.n_hits;
    const uint32_t npairs_me = (uint32_t)nq * (uint32_t)nc_me;
    if (G == 1) {
        out->n_entries   = so.n_entries;
        out->d_qoff      = so.d_query_off;
        out->d_entry_off = so.d_entry_off;
        out->d_chunk     = so.d_chunk;
        out->d_start     = so.d_start;
        out->d_end       = so.d_end;
        out->n_chunks    = nc_me;
        return PSS_OK;
    }
    PSS_CUDA_TRY(cudaEventRecord(r->ev1, s));
    const double tw1 = wall_ms();

    // ---- 3. gather-v to rank 0 ----------------------------------------------------------------
    if (me != 0) {
        PSS_NCCL_TRY(nccl().Send(so.d_entry_off, (size_t)npairs_me + 1, ncclUint32, 0, c->comm, s));
        if (so.n_entries) {
            PSS_NCCL_TRY(nccl().GroupStart());
            PSS_NCCL_TRY(nccl().Send(so.d_start, (size_t)so.n_entries, ncclUint32, 0, c->comm, s));
            PSS_NCCL_TRY(nccl().Send(so.d_end, (size_t)so.n_entries, ncclUint32, 0, c->comm, s));
            PSS_NCCL_TRY(nccl().GroupEnd());
        }
        PSS_CUDA_TRY(cudaStreamSynchronize(s));
        return PSS_OK;
    }

    // rank 0: per-rank entry-offset arrays (sizes are known: nq * chunks owned + 1)
    std::vector<uint32_t> nc_of(G), np_of(G);
    size_t e_words = 0, f_words = 0;
    for (int q = 0; q < G; ++q) {
        nc_of[q] = owned_chunks((uint32_t)q, (uint32_t)G, n_total);
        np_of[q] = (uint32_t)nq * nc_of[q];
        if (q != 0) e_words += (size_t)np_of[q] + 1;
        f_words += np_of[q];
    }
    const size_t desc_bytes = ((size_t)G * sizeof(RankDesc) + 255) & ~(size_t)255;
    const size_t cnt_bytes  = 256;
    PSS_TRY(r->dist_desc.ensure(desc_bytes + cnt_bytes + (e_words + f_words) * sizeof(uint32_t)));
    RankDesc *d_desc   = r->dist_desc.as<RankDesc>();
    uint32_t *d_counts = reinterpret_cast<uint32_t *>(r->dist_desc.as<uint8_t>() + desc_bytes);
    uint32_t *d_E      = reinterpret_cast<uint32_t *>(r->dist_desc.as<uint8_t>() + desc_bytes + cnt_bytes);
    uint32_t *d_F      = d_E + e_words;
    const size_t need_h = (size_t)G * sizeof(RankDesc) + (size_t)G * sizeof(uint32_t);
    if (need_h > c->h_cap) {
        if (c->h_pinned) cudaFreeHost(c->h_pinned);
        c->h_pinned = nullptr; c->h_cap = 0;
        PSS_CUDA_TRY(cudaMallocHost(&c->h_pinned, need_h));
        c->h_cap = need_h;
    }
    RankDesc *h_desc   = static_cast<RankDesc *>(c->h_pinned);
    uint32_t *h_counts = reinterpret_cast<uint32_t *>(h_desc + G);
    {
        uint32_t *e_at = d_E, *f_at = d_F;
        for (int q = 0; q < G; ++q) {
            RankDesc &d = h_desc[q];
            d = RankDesc();
            d.nc = nc_of[q];
            d.npairs = np_of[q];
            d.F = f_at;
            f_at += np_of[q];
            if (q == 0) {
                d.E = so.d_entry_off; d.start = so.d_start; d.end = so.d_end;
            } else {
                d.E = e_at;
                e_at += (size_t)np_of[q] + 1;
            }
        }
    }
    PSS_NCCL_TRY(nccl().GroupStart());
    for (int q = 1; q < G; ++q)
        PSS_NCCL_TRY(nccl().Recv(const_cast<uint32_t *>(h_desc[q].E), (size_t)np_of[q] + 1, ncclUint32, q, c->comm, s));
    PSS_NCCL_TRY(nccl().GroupEnd());
    PSS_CUDA_TRY(cudaMemcpyAsync(d_desc, h_desc, (size_t)G * sizeof(RankDesc), cudaMemcpyHostToDevice, s));
    dist_counts_kernel<<<1, 32, 0, s>>>(d_desc, G, d_counts);
    PSS_LAUNCH_CHECK();
    PSS_CUDA_TRY(cudaMemcpyAsync(h_counts, d_counts, (size_t)G * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    PSS_CUDA_TRY(cudaStreamSynchronize(s));     // the one host read of the exchange: how much each rank sends
    const double tw2 = wall_ms();

    int64_t total_entries = 0, recv_entries = 0;
    uint32_t max_count = 0;
    for (int q = 0; q < G; ++q) {
        total_entries += h_counts[q];
        if (q) recv_entries += h_counts[q];
        max_count = std::max(max_count, h_counts[q]);
    }
    if (h_counts[0] != (uint32_t)so.n_entries) return fail(PSS_ERR_CUDA, "distributed search: inconsistent local entry count");
    if (total_entries >= (1ll << 32)) return fail(PSS_ERR_ARG, "more than 2^32 entries in one batch");
    PSS_TRY(r->dist_recv.ensure((size_t)std::max<int64_t>(recv_entries, 1) * 2 * sizeof(uint32_t)));
    PSS_TRY(r->dist_final.ensure((size_t)std::max<int64_t>(total_entries, 1) * 3 * sizeof(uint32_t)));
    PSS_TRY(r->dist_qoff.ensure(off_bytes));
    {
        uint32_t *at = r->dist_recv.as<uint32_t>();
        PSS_NCCL_TRY(nccl().GroupStart());
        for (int q = 1; q < G; ++q) {
            const uint32_t cnt = h_counts[q];
            h_desc[q].start = at;
            h_desc[q].end   = at + cnt;
            h_desc[q].count = cnt;
            at += 2 * (size_t)cnt;
            if (cnt) {
                PSS_NCCL_TRY(nccl().Recv(const_cast<uint32_t *>(h_desc[q].start), cnt, ncclUint32, q, c->comm, s));
                PSS_NCCL_TRY(nccl().Recv(const_cast<uint32_t *>(h_desc[q].end), cnt, ncclUint32, q, c->comm, s));
            }
        }
        PSS_NCCL_TRY(nccl().GroupEnd());
        h_desc[0].count = h_counts[0];
    }
    PSS_CUDA_TRY(cudaMemcpyAsync(d_desc, h_desc, (size_t)G * sizeof(RankDesc), cudaMemcpyHostToDevice, s));
    if (dist_trace()) PSS_CUDA_TRY(cudaStreamSynchronize(s));
    const double tw3 = wall_ms();

    // ---- 4. placement into (query, chunk, SA order) ------------------------------------------
    int32_t  *f_chunk = r->dist_final.as<int32_t>();
    uint32_t *f_start = reinterpret_cast<uint32_t *>(f_chunk) + total_entries;
    uint32_t *f_end   = f_start + total_entries;
    uint32_t max_pairs = 0;
    for (int q = 0; q < G; ++q) max_pairs = std::max(max_pairs, np_of[q]);
    if (max_pairs) {
        dist_pair_final_kernel<<<dim3((unsigned)div_up(max_pairs, 256), (unsigned)G), 256, 0, s>>>(d_desc, G);
        PSS_LAUNCH_CHECK();
    }
    if (max_count) {
        dist_place_kernel<<<dim3((unsigned)div_up(max_count, 256), (unsigned)G), 256, 0, s>>>(d_desc, G, 0, f_chunk, f_start, f_end);
        PSS_LAUNCH_CHECK();
    }
    dist_query_off_kernel<<<(unsigned)div_up((int64_t)nq + 1, 256), 256, 0, s>>>(d_desc, G, (uint32_t)nq,
                                                                                r->dist_qoff.as<int64_t>());
    PSS_LAUNCH_CHECK();
    PSS_CUDA_TRY(cudaEventRecord(r->ev2, s));
    PSS_CUDA_TRY(cudaStreamSynchronize(s));
    PSS_CUDA_TRY(cudaEventElapsedTime(&out->ms_exchange, r->ev1, r->ev2));
    if (dist_trace())
        fprintf(stderr, "[pss dist] rank 0: bcast+local search %.3f ms | entry offsets in + counts %.3f ms | %lld tuples in "
                        "(%.1f MB) %.3f ms | placement %.3f ms | total %.3f ms\n",
                tw1 - tw0, tw2 - tw1, (long long)recv_entries, recv_entries * 8e-6, tw3 - tw2, wall_ms() - tw3, wall_ms() - tw0);
    out->n_entries   = total_entries;
    out->d_qoff      = r->dist_qoff.as<int64_t>();
    out->d_entry_off = nullptr;          // per-rank arrays only; the merged result is indexed by query
    out->d_chunk     = f_chunk;
    out->d_start     = f_start;
    out->d_end       = f_end;
    out->n_chunks    = (int32_t)n_total;
    return PSS_OK;
}

// ---- fused exchange: compaction kernels store into rank 0's arrays over peer memory --------------

// One NCCL all-reduce of a word as a stream-ordered barrier: when it completes on rank 0,
// every rank's earlier work on its stream (its compaction kernel, whose peer stores are
// flushed at kernel end) is done.
int stream_barrier(pss_comm *c, cudaStream_t s) {
    PSS_NCCL_TRY(nccl().AllReduce(c->d_small + 32, c->d_small + 33, 1, ncclUint32, ncclSum, c->comm, s));
    return PSS_OK;
}

// min over ranks of a host value (collective, synchronises the stream)
int all_min(pss_comm *c, cudaStream_t s, uint32_t mine, uint32_t *out) {
    c->h_small[40] = mine;
    PSS_TRY(copy_words(c->d_small + 40, c->h_small + 40, 1, s));
    PSS_NCCL_TRY(nccl().AllReduce(c->d_small + 40, c->d_small + 41, 1, ncclUint32, ncclMin, c->comm, s));
    PSS_TRY(copy_words(c->h_small + 41, c->d_small + 41, 1, s));
    PSS_CUDA_TRY(cudaStreamSynchronize(s));
    *out = c->h_small[41];
    return PSS_OK;
}

// (Re)allocates rank 0's result block for `entries` tuples and maps it on every rank.
// Collective; returns PSS_OK with *ok = false when some rank cannot map peer memory.
int fused_grow(pss_comm *c, cudaStream_t s, int64_t entries, bool *ok) {
    *ok = false;
    const int64_t cap = std::max<int64_t>(entries + entries / 2, 1 << 20);
    // 1. every mapping of the old block goes away before the block itself
    if (c->rank != 0 && c->final_base) {
        cudaIpcCloseMemHandle(c->final_base);
        c->final_base = nullptr;
    }
    uint32_t all_ok = 0;
    PSS_TRY(all_min(c, s, 1u, &all_ok));
    cudaIpcMemHandle_t handle;
    static_assert(sizeof(handle) == 64, "cudaIpcMemHandle_t is 64 bytes");
    uint32_t good = 1;
    if (c->rank == 0) {
        cudaFree(c->final_base);
        c->final_base = nullptr;
        c->final_cap = 0;
        if (cudaMalloc(&c->final_base, (size_t)cap * 3 * sizeof(uint32_t)) != cudaSuccess ||
            cudaIpcGetMemHandle(&handle, c->final_base) != cudaSuccess) {
            cudaGetLastError();
            good = 0;
            std::memset(&handle, 0, sizeof(handle));
        }
        std::memcpy(c->h_small, &handle, sizeof(handle));
        PSS_TRY(copy_words(c->d_small, c->h_small, 16, s));
    }
    // 2. the handle reaches every rank
    PSS_NCCL_TRY(nccl().Broadcast(c->d_small, c->d_small, 64, ncclUint8, 0, c->comm, s));
    if (c->rank != 0) {
        PSS_TRY(copy_words(c->h_small, c->d_small, 16, s));
        PSS_CUDA_TRY(cudaStreamSynchronize(s));
        std::memcpy(&handle, c->h_small, sizeof(handle));
        void *mapped = nullptr;
        if (cudaIpcOpenMemHandle(&mapped, handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            good = 0;
        } else {
            c->final_base = static_cast<uint32_t *>(mapped);
        }
    }
    PSS_TRY(all_min(c, s, good, &all_ok));
    if (!all_ok) {
        if (c->rank != 0 && c->final_base) cudaIpcCloseMemHandle(c->final_base);
        if (c->rank == 0) cudaFree(c->final_base);
        c->final_base = nullptr;
        c->final_cap = 0;
        return PSS_OK;
    }
    c->final_cap = cap;
    *ok = true;
    return PSS_OK;
}

int fused_setup(pss_comm *c, cudaStream_t s) {
    c->fused = 0;
    if (c->world > 16) return PSS_OK;           // the count words of the staging block hold 16 ranks: one node
    if (const char *e = std::getenv("PSS_DIST_FUSED"))
        if (std::atoi(e) == 0) return PSS_OK;
    PSS_CUDA_TRY(cudaMalloc(&c->d_small, 64 * sizeof(uint32_t)));
    PSS_CUDA_TRY(cudaMemset(c->d_small, 0, 64 * sizeof(uint32_t)));
    PSS_TRY(alloc_mapped_words(&c->h_small, 64));
    bool ok = false;
    PSS_TRY(fused_grow(c, s, 1 << 20, &ok));
    c->fused = ok ? 1 : 0;
    if (dist_trace())
        fprintf(stderr, "[pss dist] rank %d: fused peer-memory exchange %s\n", c->rank, ok ? "enabled" : "unavailable (NCCL gather-v)");
    return PSS_OK;
}

// The whole collective, fused variant.  Same contract as dist_search_core.
int dist_search_fused(pss_reader *r, pss_comm *c, int32_t nq, int64_t total, DistOut *out) {
    const int G = c->world, me = c->rank;
    cudaStream_t s = r->searcher.stream();
    const size_t off_bytes = ((size_t)nq + 1) * sizeof(int64_t);
    const uint32_t n_total = (uint32_t)r->chunks.size();
    const int nc_me = r->searcher.num_chunks();
    const double tw0 = wall_ms();
    PSS_CUDA_TRY(cudaEventRecord(r->ev0, s));
    // ---- 1. the batch reaches every GPU; 2. local search, compaction deferred ---------------------
    PSS_NCCL_TRY(nccl().Broadcast(r->d_pat, r->d_pat, off_bytes + (size_t)total, ncclUint8, 0, c->comm, s));
    SearchOutput so;
    PSS_TRY(r->searcher.search(r->d_pat + off_bytes, reinterpret_cast<const int64_t *>(r->d_pat), nq, s, &so, &out->times,
                               /*defer_compact=*/true));
    out->n_hits = so.n_hits;
    PSS_CUDA_TRY(cudaEventRecord(r->ev1, s));
    const double tw1 = wall_ms();

    // ---- 3. all-gather of the per-pair entry offsets (fixed size: nq * most chunks per rank + 1) ---
    std::vector<uint32_t> nc_of(G), np_of(G);
    uint32_t max_np = 0;
    for (int q = 0; q < G; ++q) {
        nc_of[q] = owned_chunks((uint32_t)q, (uint32_t)G, n_total);
        np_of[q] = (uint32_t)nq * nc_of[q];
        max_np = std::max(max_np, np_of[q]);
    }
    const size_t words = (size_t)max_np + 1;
    PSS_TRY(c->all_e.ensure((size_t)(G + 1) * words * sizeof(uint32_t)));
    uint32_t *e_send = c->all_e.as<uint32_t>(), *e_all = e_send + words;
    PSS_CUDA_TRY(cudaMemcpyAsync(e_send, so.d_entry_off, ((size_t)np_of[me] + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    PSS_NCCL_TRY(nccl().AllGather(e_send, e_all, words, ncclUint32, c->comm, s));
    PSS_TRY(c->f_me.ensure((size_t)std::max<uint32_t>(np_of[me], 1) * sizeof(uint32_t)));
    PSS_TRY(c->desc_dev.ensure((size_t)G * sizeof(RankDesc) + 256));
    const size_t need_h = (size_t)G * sizeof(RankDesc);
    if (need_h > c->h_cap) {
        if (c->h_pinned) cudaFreeHost(c->h_pinned);
        c->h_pinned = nullptr; c->h_cap = 0;
        PSS_CUDA_TRY(cudaMallocHost(&c->h_pinned, need_h));
        c->h_cap = need_h;
    }
    RankDesc *h_desc = static_cast<RankDesc *>(c->h_pinned), *d_desc = c->desc_dev.as<RankDesc>();
    for (int q = 0; q < G; ++q) {
        h_desc[q] = RankDesc();
        h_desc[q].E = e_all + (size_t)q * words;
        h_desc[q].nc = nc_of[q];
        h_desc[q].npairs = np_of[q];
        h_desc[q].F = q == me ? c->f_me.as<uint32_t>() : nullptr;
    }
    PSS_CUDA_TRY(cudaMemcpyAsync(d_desc, h_desc, need_h, cudaMemcpyHostToDevice, s));
    dist_counts_kernel<<<1, 32, 0, s>>>(d_desc, G, c->d_small + 44);
    PSS_LAUNCH_CHECK();
    PSS_TRY(copy_words(c->h_small + 44, c->d_small + 44, G, s));
    PSS_CUDA_TRY(cudaStreamSynchronize(s));      // the one host read of the exchange: entries per rank
    int64_t total_entries = 0;
    for (int q = 0; q < G; ++q) total_entries += c->h_small[44 + q];
    if (c->h_small[44 + me] != (uint32_t)so.n_entries) return fail(PSS_ERR_CUDA, "distributed search: inconsistent local entry count");
    if (total_entries >= (1ll << 32)) return fail(PSS_ERR_ARG, "more than 2^32 entries in one batch");
    const double tw2 = wall_ms();
    if (total_entries > c->final_cap) {          // every rank sees the same total: the grow step is collective
        bool ok = false;
        PSS_TRY(fused_grow(c, s, total_entries, &ok));
        if (!ok) return fail(PSS_ERR_CUDA, "distributed search: cannot re-map rank 0's result block on every rank");
    }

    // ---- 4. where this rank's pairs go, then the compaction writes them there (peer stores) ------
    int32_t  *f_chunk = reinterpret_cast<int32_t *>(c->final_base);
    uint32_t *f_start = c->final_base + c->final_cap;
    uint32_t *f_end   = c->final_base + 2 * c->final_cap;
    if (np_of[me]) {
        dist_pair_final_one_kernel<<<(unsigned)div_up(np_of[me], 256), 256, 0, s>>>(d_desc, G, me);
        PSS_LAUNCH_CHECK();
    }
    if (so.deferred) {
        // chunk ids are not sent: rank 0 derives them from the global pair offsets (below)
        PSS_TRY(r->searcher.compact_deferred(c->f_me.as<uint32_t>(), nullptr, f_start, f_end, s));
    } else if (so.n_entries) {
        // oversized batch (several sub-batches were compacted locally): place from the local arrays
        h_desc[me].E = so.d_entry_off; h_desc[me].start = so.d_start; h_desc[me].end = so.d_end;
        h_desc[me].count = (uint32_t)so.n_entries;
        PSS_CUDA_TRY(cudaMemcpyAsync(d_desc + me, h_desc + me, sizeof(RankDesc), cudaMemcpyHostToDevice, s));
        dist_place_kernel<<<dim3((unsigned)div_up(so.n_entries, 256), 1), 256, 0, s>>>(d_desc, G, me, f_chunk, f_start, f_end);
        PSS_LAUNCH_CHECK();
    }
    if (me == 0 && total_entries) {
        const uint32_t np_all = (uint32_t)nq * n_total;
        PSS_TRY(c->pairs_all.ensure(((size_t)np_all + 1) * sizeof(uint32_t)));
        dist_global_pairs_kernel<<<(unsigned)div_up((int64_t)np_all + 1, 256), 256, 0, s>>>(d_desc, G, (uint32_t)nq, n_total,
                                                                                           c->pairs_all.as<uint32_t>());
        PSS_LAUNCH_CHECK();
        dist_fill_chunk_kernel<<<(unsigned)div_up((int64_t)np_all * 32, 256), 256, 0, s>>>(c->pairs_all.as<uint32_t>(), np_all,
                                                                                          n_total, f_chunk);
        PSS_LAUNCH_CHECK();
    }
    // ---- 5. everyone's stores have landed ------------------------------------------------------
    PSS_TRY(stream_barrier(c, s));
    if (me == 0) {
        PSS_TRY(c->qoff.ensure(off_bytes));
        dist_query_off_kernel<<<(unsigned)div_up((int64_t)nq + 1, 256), 256, 0, s>>>(d_desc, G, (uint32_t)nq, c->qoff.as<int64_t>());
        PSS_LAUNCH_CHECK();
    }
    PSS_CUDA_TRY(cudaEventRecord(r->ev2, s));
    PSS_CUDA_TRY(cudaStreamSynchronize(s));
    if (me != 0) return PSS_OK;
    PSS_CUDA_TRY(cudaEventElapsedTime(&out->ms_exchange, r->ev1, r->ev2));
    if (dist_trace())
        fprintf(stderr, "[pss dist] rank 0 (fused): bcast+local search %.3f ms | entry offsets all-gathered + counts %.3f ms | "
                        "%lld entries placed by peer stores + barrier %.3f ms | total %.3f ms\n",
                tw1 - tw0, tw2 - tw1, (long long)total_entries, wall_ms() - tw2, wall_ms() - tw0);
    out->n_entries   = total_entries;
    out->d_qoff      = c->qoff.as<int64_t>();
    out->d_entry_off = total_entries ? c->pairs_all.as<uint32_t>() : nullptr;   // (query, chunk) order over ALL chunks
    out->d_chunk     = f_chunk;
    out->d_start     = f_start;
    out->d_end       = f_end;
    out->n_chunks    = (int32_t)n_total;
    return PSS_OK;
}

int check_dist_args(pss_reader *r, pss_comm *c, int32_t nq) {
    if (!r || !c || nq < 0) return fail(PSS_ERR_ARG, "bad search arguments");
    if (!r->subs.empty()) return fail(PSS_ERR_ARG, "distributed search needs a single-device reader per rank");
    if (c->world > 1 && !c->comm) return fail(PSS_ERR_ARG, "communicator not initialised");
    if (r->searcher.device() != c->device) return fail(PSS_ERR_ARG, "reader and communicator are on different GPUs");
    return PSS_OK;
}

}  // namespace

extern "C" {

int32_t pss_comm_unique_id(uint8_t id[PSS_COMM_ID_BYTES]) {
    if (!id) return fail(PSS_ERR_ARG, "null id");
    static_assert(sizeof(ncclUniqueId) == PSS_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    if (!nccl().ok) return fail(PSS_ERR_CUDA, nccl().error);
    ncclUniqueId u;
    PSS_NCCL_TRY(nccl().GetUniqueId(&u));
    std::memcpy(id, &u, PSS_COMM_ID_BYTES);
    return PSS_OK;
}

int32_t pss_comm_create(const uint8_t id[PSS_COMM_ID_BYTES], int32_t rank, int32_t world, pss_comm **out) {
    if (!out || world < 1 || rank < 0 || rank >= world || (world > 1 && !id)) return fail(PSS_ERR_ARG, "bad communicator arguments");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(PSS_ERR_CUDA, "no CUDA device available (libpss_b200 has no CPU fallback)");
    std::unique_ptr<pss_comm> c(new (std::nothrow) pss_comm());
    if (!c) return fail(PSS_ERR_NOMEM, "out of host memory");
    c->rank = rank;
    c->world = world;
    c->device = default_device();
    if (c->device >= ndev) return fail(PSS_ERR_ARG, "device index out of range");
    if (world > 1) {
        if (!nccl().ok) return fail(PSS_ERR_CUDA, nccl().error);
        DeviceGuard guard;
        PSS_CUDA_TRY(cudaSetDevice(c->device));
        ncclUniqueId u;
        std::memcpy(&u, id, PSS_COMM_ID_BYTES);
        PSS_NCCL_TRY(nccl().CommInitRank(&c->comm, world, u, rank));
    }
    *out = c.release();
    return PSS_OK;
}

int32_t pss_comm_destroy(pss_comm *c) {
    if (!c) return PSS_OK;
    if (c->comm) {
        DeviceGuard guard;
        cudaSetDevice(c->device);
        nccl().CommDestroy(c->comm);
    }
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    {
        DeviceGuard guard;
        cudaSetDevice(c->device);
        if (c->final_base) {
            if (c->rank == 0) cudaFree(c->final_base);
            else cudaIpcCloseMemHandle(c->final_base);
        }
        cudaFree(c->d_small);
        if (c->h_small) cudaFreeHost(c->h_small);
    }
    delete c;
    return PSS_OK;
}

int32_t pss_reader_search_batch_dist(pss_reader *r, pss_comm *c, const uint8_t *patterns, const int64_t *offsets,
                                     int32_t nq, pss_result **out) {
    if (!out || (nq > 0 && !offsets)) return fail(PSS_ERR_ARG, "bad search arguments");
    *out = nullptr;
    PSS_TRY(check_dist_args(r, c, nq));
    if (nq > 0) {
        if (offsets[0] != 0) return fail(PSS_ERR_ARG, "pattern offsets must start at 0");
        for (int32_t q = 0; q < nq; ++q)
            if (offsets[q + 1] < offsets[q]) return fail(PSS_ERR_ARG, "pattern offsets must be non-decreasing");
    }
    const int64_t total = nq > 0 ? offsets[nq] : 0;
    if (c->rank == 0 && total > 0 && !patterns) return fail(PSS_ERR_ARG, "bad pattern buffer");
    std::unique_ptr<ResultOwner> res(new (std::nothrow) ResultOwner());
    if (!res) return fail(PSS_ERR_NOMEM, "out of host memory");
    std::memset(&res->pub, 0, sizeof(res->pub));
    res->pool = r->pool;
    res->pub.n_ranks = c->world;
    DeviceGuard guard;
    PSS_CUDA_TRY(cudaSetDevice(r->searcher.device()));
    cudaStream_t s = r->searcher.stream();
    const size_t off_bytes = ((size_t)nq + 1) * sizeof(int64_t);
    PSS_TRY(r->ensure_patterns(off_bytes + (size_t)total));
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    PSS_CUDA_TRY(cudaEventCreate(&ev_begin));
    PSS_CUDA_TRY(cudaEventCreate(&ev_end));
    struct EvFree { cudaEvent_t a, b; ~EvFree() { cudaEventDestroy(a); cudaEventDestroy(b); } } evfree{ev_begin, ev_end};
    PSS_CUDA_TRY(cudaEventRecord(ev_begin, s));
    if (c->rank == 0) {
        if (nq > 0) std::memcpy(r->h_pat.p, offsets, off_bytes);
        else std::memset(r->h_pat.p, 0, off_bytes);
        if (total) std::memcpy(static_cast<uint8_t *>(r->h_pat.p) + off_bytes, patterns, (size_t)total);
        PSS_CUDA_TRY(cudaMemcpyAsync(r->d_pat, r->h_pat.p, off_bytes + (size_t)total, cudaMemcpyHostToDevice, s));
    }
    DistOut d;
    const double th0 = wall_ms();
    PSS_TRY(dist_search_core(r, c, nq, total, &d));
    const double th1 = wall_ms();
    if (c->rank == 0) {
        PSS_TRY(res->alloc(nq, d.n_entries));
        const double th2 = wall_ms();
        PSS_CUDA_TRY(cudaMemcpyAsync(res->query_off(), d.d_qoff, off_bytes, cudaMemcpyDeviceToHost, s));
        if (d.n_entries) {
            const size_t b = (size_t)d.n_entries * 4;
            PSS_CUDA_TRY(cudaMemcpyAsync(res->chunk(), d.d_chunk, b, cudaMemcpyDeviceToHost, s));
            PSS_CUDA_TRY(cudaMemcpyAsync(res->start(), d.d_start, b, cudaMemcpyDeviceToHost, s));
            PSS_CUDA_TRY(cudaMemcpyAsync(res->end(), d.d_end, b, cudaMemcpyDeviceToHost, s));
        }
        PSS_CUDA_TRY(cudaEventRecord(ev_end, s));
        PSS_CUDA_TRY(cudaEventSynchronize(ev_end));
        PSS_CUDA_TRY(cudaEventElapsedTime(&res->pub.ms_total, ev_begin, ev_end));
        if (dist_trace())
            fprintf(stderr, "[pss dist] rank 0 host call: core %.3f ms | result block %.3f ms | D2H of %lld entries %.3f ms\n",
                    th1 - th0, th2 - th1, (long long)d.n_entries, wall_ms() - th2);
        if (res->query_off()[nq] != d.n_entries)
            return fail(PSS_ERR_CUDA, "internal error: per-query counts do not add up to the entry count");
        res->publish(d.n_entries);
    } else {
        PSS_TRY(res->alloc(nq, 0));
        std::memset(res->query_off(), 0, off_bytes);
        res->publish(0);
    }
    res->pub.n_hits      = d.n_hits;
    res->pub.ms_bounds   = d.times.ms_bounds;
    res->pub.ms_extract  = d.times.ms_extract;
    res->pub.ms_dedup    = d.times.ms_dedup;
    res->pub.ms_exchange = d.ms_exchange;
    *out = &res.release()->pub;
    return PSS_OK;
}

int32_t pss_reader_search_batch_dist_device(pss_reader *r, pss_comm *c, const uint8_t *d_patterns,
                                            const int64_t *d_offsets, int32_t nq, int64_t total_pattern_bytes,
                                            pss_device_result *out) {
    if (!out || total_pattern_bytes < 0) return fail(PSS_ERR_ARG, "bad search arguments");
    std::memset(out, 0, sizeof(*out));
    PSS_TRY(check_dist_args(r, c, nq));
    if (c->rank == 0 && nq > 0 && (!d_offsets || (total_pattern_bytes > 0 && !d_patterns)))
        return fail(PSS_ERR_ARG, "bad pattern buffer");
    DeviceGuard guard;
    PSS_CUDA_TRY(cudaSetDevice(r->searcher.device()));
    cudaStream_t s = r->searcher.stream();
    const size_t off_bytes = ((size_t)nq + 1) * sizeof(int64_t);
    PSS_TRY(r->ensure_patterns(off_bytes + (size_t)total_pattern_bytes));
    if (c->rank == 0) {
        if (nq > 0) PSS_CUDA_TRY(cudaMemcpyAsync(r->d_pat, d_offsets, off_bytes, cudaMemcpyDeviceToDevice, s));
        else PSS_CUDA_TRY(cudaMemsetAsync(r->d_pat, 0, off_bytes, s));
        if (total_pattern_bytes)
            PSS_CUDA_TRY(cudaMemcpyAsync(r->d_pat + off_bytes, d_patterns, (size_t)total_pattern_bytes,
                                         cudaMemcpyDeviceToDevice, s));
    }
    DistOut d;
    PSS_TRY(dist_search_core(r, c, nq, total_pattern_bytes, &d));
    out->n_queries = nq;
    out->n_hits    = d.n_hits;
    out->ms_bounds   = d.times.ms_bounds;
    out->ms_extract  = d.times.ms_extract;
    out->ms_dedup    = d.times.ms_dedup;
    out->ms_exchange = d.ms_exchange;
    if (c->rank == 0) {
        out->n_chunks        = d.n_chunks;
        out->n_entries       = d.n_entries;
        out->d_query_offsets = d.d_qoff;
        out->d_entry_offsets = d.d_entry_off;
        out->d_chunk_id      = d.d_chunk;
        out->d_line_start    = d.d_start;
        out->d_line_end      = d.d_end;
    }
    return PSS_OK;
}

}  // extern "C"
