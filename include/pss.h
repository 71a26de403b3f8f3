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

/* Library version string, e.g. "pss_b200 0.1 (sm_100a)". */
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

/* src/lib.rs:88-103.  `text` is the entry's UTF-8 bytes (no terminator needed).
 * PSS_ERR_TOOBIG if len > capacity. */
int32_t pss_writer_add_entry(pss_writer *w, const uint8_t *text, size_t len);

/* src/lib.rs:67-86 (bstr for_byte_line: split at '\n', strip "\n" or "\r\n"). */
int32_t pss_writer_add_entries_from_file_lines(pss_writer *w, const char *input_file_path);

/* src/lib.rs:105-124: serialise the buffered chunk (u32le n, text, u32le 4n, i32le SA[n])
 * with the SA built on the GPU; no-op on an empty buffer. */
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
} pss_result;

/* Reader::new (src/lib.rs:162-199): parse the container, keep every chunk's text in host
 * memory, upload text + SA of every owned chunk to the GPU.
 * PSS_ERR_NOTFOUND if the file is missing. */
int32_t pss_reader_open(const char *index_file_path, pss_reader **out);

/* Same, but this process owns only chunks k with k % shard_count == shard_rank
 * (chunk → GPU map for one-process-per-GPU runs).  Text of foreign chunks is not kept. */
int32_t pss_reader_open_sharded(const char *index_file_path, int32_t shard_rank,
                                int32_t shard_count, pss_reader **out);

int32_t pss_reader_close(pss_reader *r);

/* Chunks in the container / owned by this reader. */
int32_t pss_reader_num_chunks(const pss_reader *r);
int32_t pss_reader_num_local_chunks(const pss_reader *r);

/* Host text of global chunk `chunk` (NULL/0 if not owned). */
int32_t pss_reader_chunk_text(const pss_reader *r, int32_t chunk, const uint8_t **text,
                              int64_t *len);

/*
 * Batched search (Reader.search = batch of one; search_multiple = one call).
 * patterns: concatenated pattern bytes (HOST); offsets[nq+1] (HOST, offsets[0] = 0):
 * pattern q is patterns[offsets[q] .. offsets[q+1]).  Empty patterns are legal (match
 * every entry).  On success *out is a new result (free with pss_result_free).
 * The handle is not thread-safe (mirrors `&mut self`, lib.rs:202).
 */
int32_t pss_reader_search_batch(pss_reader *r, const uint8_t *patterns, const int64_t *offsets,
                                int32_t nq, pss_result **out);

/*
 * Same search with DEVICE-resident inputs and outputs, for one-process-per-GPU runs that
 * gather hits with NCCL: d_patterns / d_offsets are device pointers; the result tuples
 * are left on the device in caller-provided buffers of `capacity` entries each
 * (d_query_id, d_chunk_id, d_line_start, d_line_end).  *n_entries receives the number
 * of entries produced; if it exceeds `capacity` nothing past capacity is written and
 * PSS_ERR_NOMEM is returned (call again with larger buffers).
 */
int32_t pss_reader_search_batch_device(pss_reader *r, const uint8_t *d_patterns,
                                       const int64_t *d_offsets, int32_t nq,
                                       int64_t total_pattern_bytes,
                                       int32_t *d_query_id, int32_t *d_chunk_id,
                                       uint32_t *d_line_start, uint32_t *d_line_end,
                                       int64_t capacity, int64_t *n_entries,
                                       int64_t *n_hits, void *stream);

void pss_result_free(pss_result *res);

/* Number of kernels this library has launched in the calling process (for bench.py's
 * gpu_launches claim). */
int64_t pss_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PSS_H_ */
