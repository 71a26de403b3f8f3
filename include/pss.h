/*
 * pss.h — C ABI of libpss_b200.so, the B200 (sm_100a) implementation of the two
 * hot paths of Intsights/PySubstringSearch:
 *
 *   BUILD   suffix-array construction per chunk      (reference: src/lib.rs:24-40 →
 *           src/libsais/libsais.c:6597 `libsais`)
 *   SEARCH  Reader.search / search_multiple           (reference: src/lib.rs:201-287,
 *           pysubstringsearch/__init__.py:61-73)
 *
 * Plain C types only (pointers, sizes, int status codes): a Rust/pyo3 host
 * (src/lib.rs:14-22), cgo, JNI or ctypes can bind it directly.  No exceptions and no
 * aborts cross this boundary; every function that can fail returns a status code and
 * leaves a human-readable message in pss_last_error() (thread-local).
 *
 * Every compute entry point runs on the GPU.  There is no CPU fallback: with no usable
 * CUDA device the calls fail with PSS_ERR_CUDA.
 */
#ifndef PSS_H_
#define PSS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------- */
#define PSS_OK             0
#define PSS_ERR_ARG      (-1)  /* same value libsais returns for bad arguments (libsais.c:6599-6602) */
#define PSS_ERR_NOMEM    (-2)  /* same value libsais returns for allocation failure (libsais.c:6505-6507) */
#define PSS_ERR_CUDA     (-3)  /* CUDA runtime / kernel failure, or no device */
#define PSS_ERR_IO       (-4)  /* file I/O error other than "not found" (reference: io::Error → OSError) */
#define PSS_ERR_NOTFOUND (-5)  /* file does not exist (reference: FileNotFoundError, tests:48-56) */
#define PSS_ERR_TOOBIG   (-6)  /* "entry is too big" (reference: src/lib.rs:92-94 → ValueError) */
#define PSS_ERR_FORMAT   (-7)  /* truncated / malformed index container */

/* Message describing the last failure on the calling thread ("" if none). */
const char *pss_last_error(void);

/* Library version string, e.g. "pss_b200 0.2 (sm_100a)". */
const char *pss_version(void);

/* Number of CUDA devices visible to the library (0 if none / driver missing). */
int32_t pss_device_count(void);

/* Select the device used by subsequently created builders/writers/readers on this
 * thread's process (default: env PSS_DEVICE, else LOCAL_RANK, else 0). */
int32_t pss_set_device(int32_t device);
int32_t pss_get_device(void);

/* ===================================================================================== */
/* BUILD                                                                                 */
/* ===================================================================================== */

/*
 * Drop-in for `libsais` exactly as the reference declares and calls it
 * (src/lib.rs:14-22 declaration; src/lib.rs:29-37 call with fs = 0, freq = NULL;
 * src/libsais/libsais.h:56-65 contract).
 *
 *   T   [0..n)      input bytes (host memory, caller-owned)
 *   SA  [0..n+fs)   output suffix array (host memory, caller-owned); SA[0..n) is written
 *   n               text length, 0 <= n < 2^30 here (the container's u32 sa_bytes = 4n
 *                   must not wrap, src/lib.rs:116)
 *   fs              extra space after SA[n); accepted (>= 0) and ignored
 *   freq [0..256)   optional output symbol frequency table (may be NULL)
 *
 * Returns 0, or -1 for bad arguments, -2 for (device) allocation failure — the libsais
 * codes — or PSS_ERR_CUDA.  Synchronous; does H2D(T) → prefix-doubling build on the
 * GPU → D2H(SA).  Thread-safe (calls are serialised on an internal cached workspace).
 * Output is byte-identical to libsais: suffixes in unsigned-byte lexicographic order,
 * a proper prefix sorting first.
 */
int32_t pss_libsais(const uint8_t *T, int32_t *SA, int32_t n, int32_t fs, int32_t *freq);

/*
 * Asynchronous form of the same seam: lets the caller overlap the build of
 * chunk k with the ingestion of chunk k+1 (the reference blocks inside dump_data,
 * src/lib.rs:105-124 → :115) and spread chunks over the GPUs of the box.
 *
 *   pss_sa_build_begin(device, T, n, &h)   queues the build on `device` (-1 = default) and
 *                                          returns at once; T must stay valid and unchanged
 *                                          until pss_sa_build_wait(h, ..) returns
 *   pss_sa_build_wait(h, SA)               blocks until the build is done, copies SA[0..n)
 *                                          to host memory (pinned or pageable), frees h
 *
 * Each device runs its builds in begin order on one worker thread and keeps three
 * (text, SA) slots in HBM: while chunk k is built, the text of k+1 is uploaded and the
 * device→host copy done by wait(k-1) runs.  Any number of builds may be queued; wait for
 * the handles of one device in begin order.  Same return codes as pss_libsais.
 */
typedef struct pss_sa_build pss_sa_build;
int32_t pss_sa_build_begin(int32_t device, const uint8_t *T, int32_t n, pss_sa_build **out);
int32_t pss_sa_build_wait(pss_sa_build *h, int32_t *SA);

/* pss_libsais, the async builds and the Writers share one cached build engine per GPU
 * (≈ 32 bytes of HBM per text byte of the largest chunk seen, plus two text/SA slots).
 * This frees the device memory of every engine with nothing in flight; it regrows on
 * demand. */
int32_t pss_release_cached(void);

/* Reusable builder: owns the device workspace (≈ 32 bytes per text byte of capacity). */
typedef struct pss_sa_builder pss_sa_builder;

/* Per-radix-pass record (filled when profiling is enabled). */
typedef struct pss_pass_stat {
    int32_t  round;        /* 0 = initial packed-prefix sort, r >= 1 = doubling round r */
    int32_t  pass;         /* digit index inside the round's sort */
    int32_t  shift;        /* bit offset of the 8-bit digit */
    int32_t  reserved;     /* digit spread: expected distinct digits per warp x1000 */
    int64_t  n_records;    /* records entering the pass (N_active) */
    float    ms;           /* CUDA-event duration of the pass kernel */
    float    reserved2;
} pss_pass_stat;

typedef struct pss_build_stats {
    int32_t  n;                 /* text length of the last build */
    int32_t  sigma;             /* distinct byte values present */
    int32_t  bits_per_symbol;   /* code width used for the packed initial key */
    int32_t  h0;                /* symbols packed in the initial key (initial prefix depth) */
    int32_t  rounds;            /* doubling rounds executed after the initial sort */
    int32_t  n_passes;          /* radix passes executed (all rounds) */
    int32_t  n_pass_stats;      /* valid entries in pass_stats (<= PSS_MAX_PASS_STATS) */
    int32_t  n_kernel_launches; /* kernels launched by the last build */
    int64_t  active_per_round[64]; /* active suffixes entering round r (index 0 = n) */
    float    total_ms;          /* CUDA-event duration of the whole device build */
    float    sort_ms;           /* sum of pass durations (profiling only) */
    int64_t  records_sorted;    /* sum over passes of n_records */
} pss_build_stats;

#define PSS_MAX_PASS_STATS 512

/* Create a builder on `device` (-1 = current default) able to build texts up to
 * `max_n` bytes (grown on demand if a larger text arrives). */
int32_t pss_sa_builder_create(int32_t device, int64_t max_n, pss_sa_builder **out);
void    pss_sa_builder_destroy(pss_sa_builder *b);

/* Enable (1) / disable (0) per-pass CUDA-event profiling (default off; when on, events
 * are recorded around every radix pass on the build stream). */
int32_t pss_sa_builder_set_profiling(pss_sa_builder *b, int32_t on);

/*
 * Device-resident build: d_text[0..n) and d_sa[0..n) are DEVICE pointers on the
 * builder's device (e.g. torch tensor data_ptr()).  `stream` is a cudaStream_t (NULL =
 * the builder's own stream).  Returns after the build has completed on the device.
 */
int32_t pss_sa_builder_build_device(pss_sa_builder *b, const uint8_t *d_text, int32_t n,
                                    int32_t *d_sa, void *stream);

/* Host-buffer build through the same builder: H2D, build, D2H (what pss_libsais does). */
int32_t pss_sa_builder_build_host(pss_sa_builder *b, const uint8_t *h_text, int32_t n,
                                  int32_t *h_sa);

/* Statistics of the last build. pass_stats may be NULL; otherwise room for
 * PSS_MAX_PASS_STATS entries. */
int32_t pss_sa_builder_stats(pss_sa_builder *b, pss_build_stats *stats,
                             pss_pass_stat *pass_stats);

/*
 * The onesweep LSD radix sort the builder is made of, exposed for parity tests and for
 * the roofline measurement: sorts n (key u64, value u32) records by key bits
 * [begin_bit, end_bit), stable.  All pointers are DEVICE pointers; keys_alt/vals_alt are
 * scratch of the same size.  On return *result_in_alt is 1 if the sorted data sits in
 * the alt buffers, 0 if in the primary ones.  vals may be NULL (treated as 0..n-1).
 * pass_ms (host, may be NULL, room for 8 floats) receives per-pass kernel durations;
 * *n_passes the number of passes actually run (constant digits are skipped).
 */
int32_t pss_radix_sort_pairs(uint64_t *d_keys, uint64_t *d_keys_alt,
                             uint32_t *d_vals, uint32_t *d_vals_alt,
                             int64_t n, int32_t begin_bit, int32_t end_bit,
                             int32_t *result_in_alt, float *pass_ms, int32_t *n_passes,
                             void *stream);

/* ===================================================================================== */
/* Writer  (reference: src/lib.rs:42-144)                                                */
/* ===================================================================================== */

typedef struct pss_writer pss_writer;

/* File::create(index_file_path) + buffer of capacity max_chunk_len
 * (src/lib.rs:50-65).  max_chunk_len < 0 means "None" → 512 MiB default (:57). */
int32_t pss_writer_open(const char *index_file_path, int64_t max_chunk_len, pss_writer **out);

/* Same, with an explicit chunk → GPU map: chunk k of this writer is built on
 * devices[k % ndev] (ndev >= 1).  pss_writer_open uses env PSS_DEVICES ("all" or
 * "0,1,..") if set, else the default device.  Chunks are built concurrently (one engine
 * per listed GPU) while ingestion continues; records still reach the file strictly in
 * chunk order, so the container is byte-identical to the single-GPU / reference one. */
int32_t pss_writer_open_devices(const char *index_file_path, int64_t max_chunk_len,
                                const int32_t *devices, int32_t ndev, pss_writer **out);

/* src/lib.rs:88-103.  `text` is the entry's UTF-8 bytes (no terminator needed).
 * PSS_ERR_TOOBIG if len > capacity. */
int32_t pss_writer_add_entry(pss_writer *w, const uint8_t *text, size_t len);

/* 1 when adding an entry of `len` bytes would hand the buffered chunk to the builder first
 * (bindings use it to release their interpreter lock only around such calls). */
int32_t pss_writer_would_flush(const pss_writer *w, size_t len);

/* src/lib.rs:67-86 (bstr for_byte_line: split at '\n', strip "\n" or "\r\n"). */
int32_t pss_writer_add_entries_from_file_lines(pss_writer *w, const char *input_file_path);

/* src/lib.rs:105-124: serialise the buffered chunk (u32le n, text, u32le 4n, i32le SA[n])
 * with the SA built on the GPU; no-op on an empty buffer.  The build and the file write
 * proceed in the background (records reach the file in chunk order); a failure of either
 * is reported by the next dump / finalize / close. */
int32_t pss_writer_dump_data(pss_writer *w);

/* src/lib.rs:126-135: dump the last partial chunk and flush. */
int32_t pss_writer_finalize(pss_writer *w);

/* Drop (src/lib.rs:138-144): finalize, close the file, free. Returns finalize's status. */
int32_t pss_writer_close(pss_writer *w);

/* ===================================================================================== */
/* Reader  (reference: src/lib.rs:146-288)                                               */
/* ===================================================================================== */

typedef struct pss_reader pss_reader;

/*
 * One batch of results, in the reference's order: query order (the search_multiple
 * concatenation, __init__.py:61-73); inside a query, chunk order (ascending chunk id —
 * the reference's cross-chunk order is a rayon race, lib.rs:207,280); inside a chunk,
 * suffix-array order of each entry's first matching suffix, deduplicated by entry start
 * offset (lib.rs:262-278).  Arrays are owned by the result and freed by pss_result_free.
 */
typedef struct pss_result {
    int32_t         n_queries;
    int32_t         reserved;
    int64_t         n_entries;      /* total returned entries over all queries */
    const int64_t  *query_offsets;  /* [n_queries + 1] : entries of query q are [off[q], off[q+1]) */
    const int32_t  *chunk_id;       /* [n_entries] global chunk index in the container */
    const uint32_t *line_start;     /* [n_entries] entry start offset inside the chunk text (line_tail, lib.rs:270-273) */
    const uint32_t *line_end;       /* [n_entries] offset of the terminating '\n' (line_head, lib.rs:266-269) */
    int64_t         n_hits;         /* total matching suffixes before dedup (diagnostic) */
    float           ms_bounds;      /* device time: lower/upper-bound kernel */
    float           ms_extract;     /* device time: hit expansion + newline extraction */
    float           ms_dedup;       /* device time: sort-based dedup + compaction */
    float           ms_total;       /* device time of the whole batch, incl. H2D/D2H */
    float           ms_exchange;    /* multi-GPU: NCCL broadcast + gather + placement on rank 0 (0 otherwise) */
    int32_t         n_ranks;        /* ranks that contributed (1 for a single-process reader) */
} pss_result;

/* Reader::new (src/lib.rs:162-199): parse the container, keep every chunk's text in host
 * memory, upload text + SA of every owned chunk to the GPU.
 * PSS_ERR_NOTFOUND if the file is missing. */
int32_t pss_reader_open(const char *index_file_path, pss_reader **out);

/* Same, but this process owns only chunks k with k % shard_count == shard_rank
 * (chunk → GPU map for one-process-per-GPU runs).  Text of foreign chunks is not kept. */
int32_t pss_reader_open_sharded(const char *index_file_path, int32_t shard_rank,
                                int32_t shard_count, pss_reader **out);

/* Same as pss_reader_open with an explicit chunk → GPU map inside this process: chunk k
 * lives on devices[k % ndev]; a batch is answered by all listed GPUs concurrently and
 * merged into the single-process order.  (pss_reader_open reads the same list from env
 * PSS_DEVICES = "all" | "0,1,..".) */
int32_t pss_reader_open_devices(const char *index_file_path, const int32_t *devices, int32_t ndev,
                                pss_reader **out);

/* A reader over chunks that already sit in this process's GPU memory (e.g. suffix arrays
 * just built with pss_sa_builder_build_device): no file, no copy — the pointers are
 * borrowed and must outlive the reader.  d_text must be 16-byte aligned with 16 readable
 * bytes past n; h_text (host copy of the text, may be NULL) is what
 * pss_reader_chunk_text returns.  global_id is the chunk's index in the whole index
 * (ascending within one reader). */
typedef struct pss_device_chunk {
    const uint8_t *d_text;
    const int32_t *d_sa;
    const uint8_t *h_text;
    uint32_t       n;
    int32_t        global_id;
} pss_device_chunk;
int32_t pss_reader_open_device_chunks(const pss_device_chunk *chunks, int32_t n_local,
                                      int32_t n_chunks_total, int32_t device, pss_reader **out);

int32_t pss_reader_close(pss_reader *r);

/* Chunks in the container / owned by this reader. */
int32_t pss_reader_num_chunks(const pss_reader *r);
int32_t pss_reader_num_local_chunks(const pss_reader *r);

/* Host text of global chunk `chunk` (NULL/0 if not owned). */
int32_t pss_reader_chunk_text(const pss_reader *r, int32_t chunk, const uint8_t **text,
                              int64_t *len);

/*
 * Batched search (Reader.search = batch of one; search_multiple = one call).
 * patterns: concatenated pattern bytes (HOST); offsets[nq+1] (HOST, offsets[0] = 0,
 * non-decreasing — checked, PSS_ERR_ARG otherwise): pattern q is
 * patterns[offsets[q] .. offsets[q+1]).  Empty patterns are legal (match
 * every entry).  On success *out is a new result (free with pss_result_free).
 * The handle is not thread-safe (mirrors `&mut self`, lib.rs:202).
 */
int32_t pss_reader_search_batch(pss_reader *r, const uint8_t *patterns, const int64_t *offsets,
                                int32_t nq, pss_result **out);

/*
 * Result of a device-resident search: DEVICE pointers owned by the reader, valid until
 * its next search.  Entries are ordered as in pss_result; pair p = query * n_chunks +
 * chunk position (position among the chunks this result covers, ascending chunk id).
 */
typedef struct pss_device_result {
    int32_t         n_queries;
    int32_t         n_chunks;        /* chunks covered: local chunks, or all of them after a gather */
    int64_t         n_entries;
    int64_t         n_hits;
    const int64_t  *d_query_offsets; /* [n_queries + 1] */
    const uint32_t *d_entry_offsets; /* [n_queries * n_chunks + 1] entries before pair p */
    const int32_t  *d_chunk_id;      /* [n_entries] */
    const uint32_t *d_line_start;    /* [n_entries] */
    const uint32_t *d_line_end;      /* [n_entries] */
    float           ms_bounds, ms_extract, ms_dedup, ms_exchange;
} pss_device_result;

/*
 * Same search with DEVICE-resident inputs and outputs: d_patterns / d_offsets are device
 * pointers on the reader's GPU (total_pattern_bytes = offsets[nq], known to the caller);
 * the tuples stay in HBM.  `stream` is a cudaStream_t (NULL = the reader's own stream);
 * the call returns after the batch has completed on the device.  Single-device readers only.
 */
int32_t pss_reader_search_batch_device(pss_reader *r, const uint8_t *d_patterns,
                                       const int64_t *d_offsets, int32_t nq,
                                       int64_t total_pattern_bytes, pss_device_result *out,
                                       void *stream);

/* ===================================================================================== */
/* One process per GPU: sharded index, hits gathered to rank 0 over NCCL                  */
/* (reference: rayon fan-out over chunks + Mutex<Vec>::extend, src/lib.rs:205-207,280-284) */
/* ===================================================================================== */

/*
 * Communicator of the search exchange step.  Rank 0 creates a 128-byte id
 * (pss_comm_unique_id) and hands it to the other ranks by any means (the launcher's store,
 * a file, torch.distributed.broadcast); every rank then calls pss_comm_create — a
 * collective — with its rank and the world size, on the GPU selected by
 * pss_set_device / PSS_DEVICE / LOCAL_RANK.  Built on NCCL (NVLink / NVSwitch).
 */
typedef struct pss_comm pss_comm;
#define PSS_COMM_ID_BYTES 128
int32_t pss_comm_unique_id(uint8_t id[PSS_COMM_ID_BYTES]);
int32_t pss_comm_create(const uint8_t id[PSS_COMM_ID_BYTES], int32_t rank, int32_t world, pss_comm **out);
int32_t pss_comm_destroy(pss_comm *c);

/*
 * Collective batched search over an index sharded chunk k → rank k % world (readers opened
 * with pss_reader_open_sharded(path, rank, world) or pss_reader_open_device_chunks).
 * Every rank calls it with the same nq and the same offsets[] (host); pattern bytes are
 * taken from rank 0 and broadcast (other ranks may pass patterns = NULL).  Each rank
 * searches its own chunks; rank 0 receives exactly the tuples found (no padding), places
 * them in the single-process order (query, ascending chunk id, SA order) and returns them
 * in *out; on other ranks *out is an empty result.  One exchange step per batch:
 * broadcast(patterns) → local search → all-gather(per-pair entry offsets) → every rank's
 * compaction kernel stores its tuples at their final positions in rank 0's result arrays
 * over NVLink peer memory (CUDA IPC; where peers cannot be mapped, or with
 * PSS_DIST_FUSED=0 on every rank: NCCL gather-v of exact counts + a placement kernel).
 * As with any collective, a rank that fails (or does not call) leaves the others waiting:
 * treat an error on one rank as fatal for the job.
 */
int32_t pss_reader_search_batch_dist(pss_reader *r, pss_comm *c, const uint8_t *patterns,
                                     const int64_t *offsets, int32_t nq, pss_result **out);

/* Same exchange with the batch already in rank 0's HBM (d_patterns, d_offsets valid on
 * rank 0 only; nq and total_pattern_bytes identical on every rank) and the merged tuples
 * left in rank 0's HBM (*out; n_entries = 0 elsewhere). */
int32_t pss_reader_search_batch_dist_device(pss_reader *r, pss_comm *c, const uint8_t *d_patterns,
                                            const int64_t *d_offsets, int32_t nq,
                                            int64_t total_pattern_bytes, pss_device_result *out);

void pss_result_free(pss_result *res);

/* Copies `bytes` from device memory (e.g. a pss_device_result array) to host memory, for
 * callers that do not link a CUDA runtime themselves.  Synchronous. */
int32_t pss_memcpy_d2h(void *dst, const void *d_src, size_t bytes);

/* Number of kernels this library has launched in the calling process (for bench.py's
 * gpu_launches claim). */
int64_t pss_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PSS_H_ */
