/*
 * pss_oracle.c — ORACLE (test infrastructure only; never linked into the product).
 *
 * CPU restatement of the reference's host logic around the two hot paths, following
 * /root/reference/src/lib.rs and pysubstringsearch/__init__.py line by line in meaning:
 *
 *   Writer::new                      lib.rs:50-65    (default max_chunk_len 512 MiB, :57)
 *   Writer::add_entries_from_file_lines  lib.rs:67-86 (bstr for_byte_line: split at '\n',
 *                                        strip the '\n' and one preceding '\r')
 *   Writer::add_entry                lib.rs:88-103   ("entry is too big" check :92-94,
 *                                        flush rule :96-98)
 *   Writer::dump_data                lib.rs:105-124  (u32le n | text | u32le 4n | i32le SA[n])
 *   Writer::finalize                 lib.rs:126-135
 *   Reader::new                      lib.rs:162-199  (walk the container)
 *   Reader::search                   lib.rs:201-287  (lower bound :212-233, upper bound
 *                                        :235-252, SA range read :257-260, newline-delimited
 *                                        extraction + dedup by entry start :262-278)
 *   Reader.search_multiple           __init__.py:61-73 (concatenation in query order)
 *
 * The suffix array comes from a pluggable function with libsais' signature: by default
 * oracle_libsais (sais_port.c); tests may plug in the reference's own compiled libsais
 * (oracle/_ref/libsais_ref.so) with oracle_set_sa_function().
 *
 * Vec<u8> capacity growth (Rust RawVec::grow_amortized: max(2*cap, needed), min 8) is
 * restated because the flush rule compares against `capacity()` (lib.rs:75, :92, :96).
 *
 * Cross-chunk result order: the reference appends per-chunk results in rayon completion
 * order (lib.rs:207, :280), i.e. nondeterministic; this oracle uses ascending chunk order.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

typedef int32_t (*sa_fn_t)(const uint8_t *, int32_t *, int32_t, int32_t, int32_t *);
int32_t oracle_libsais(const uint8_t *T, int32_t *SA, int32_t n, int32_t fs, int32_t *freq);

static sa_fn_t g_sa_fn = oracle_libsais;
void oracle_set_sa_function(void *fn) { g_sa_fn = fn ? (sa_fn_t)fn : oracle_libsais; }

#define ORC_OK 0
#define ORC_ERR_IO (-4)
#define ORC_ERR_NOTFOUND (-5)
#define ORC_ERR_TOOBIG (-6)
#define ORC_ERR_NOMEM (-2)
#define ORC_ERR_FORMAT (-7)

/* ------------------------------------------------------------------------------------ */
/* Writer                                                                                */
/* ------------------------------------------------------------------------------------ */
typedef struct oracle_writer {
    FILE *f;
    uint8_t *buf;
    size_t len;
    size_t cap;      /* logical Vec capacity (what lib.rs compares against) */
    size_t alloc;    /* bytes actually malloc'ed (lazy: the 512 MiB default is not touched) */
} oracle_writer;

static int w_reserve_real(oracle_writer *w, size_t need) {
    if (need <= w->alloc) return ORC_OK;
    size_t na = w->alloc ? w->alloc : 4096;
    while (na < need) na *= 2;
    uint8_t *nb = (uint8_t *)realloc(w->buf, na);
    if (!nb) return ORC_ERR_NOMEM;
    w->buf = nb;
    w->alloc = na;
    return ORC_OK;
}

/* Vec::reserve(additional) semantics on the logical capacity. */
static void w_grow_logical(oracle_writer *w, size_t additional) {
    if (w->cap - w->len >= additional) return;
    size_t required = w->len + additional;
    size_t nc = w->cap * 2 > required ? w->cap * 2 : required;
    if (nc < 8) nc = 8;
    w->cap = nc;
}

static int w_append(oracle_writer *w, const uint8_t *p, size_t n) {
    w_grow_logical(w, n);                       /* extend_from_slice */
    int rc = w_reserve_real(w, w->len + n + 1);
    if (rc) return rc;
    if (n) memcpy(w->buf + w->len, p, n);
    w->len += n;
    w_grow_logical(w, 1);                       /* push(b'\n') */
    w->buf[w->len++] = '\n';
    return ORC_OK;
}

oracle_writer *oracle_writer_open(const char *path, long long max_chunk_len, int *status) {
    oracle_writer *w = (oracle_writer *)calloc(1, sizeof(*w));
    if (!w) { if (status) *status = ORC_ERR_NOMEM; return NULL; }
    w->f = fopen(path, "wb");                   /* File::create, lib.rs:55 */
    if (!w->f) {
        if (status) *status = (errno == ENOENT) ? ORC_ERR_NOTFOUND : ORC_ERR_IO;
        free(w);
        return NULL;
    }
    w->cap = max_chunk_len < 0 ? (size_t)512 * 1024 * 1024 : (size_t)max_chunk_len;  /* lib.rs:57 */
    if (status) *status = ORC_OK;
    return w;
}

int oracle_writer_dump_data(oracle_writer *w) {
    if (w->len == 0) return ORC_OK;             /* lib.rs:108-110 */
    uint32_t n32 = (uint32_t)w->len;
    if (fwrite(&n32, 4, 1, w->f) != 1) return ORC_ERR_IO;           /* :112 (little-endian host) */
    if (fwrite(w->buf, 1, w->len, w->f) != w->len) return ORC_ERR_IO; /* :113 */
    int32_t *sa = (int32_t *)malloc(sizeof(int32_t) * w->len);      /* :27 */
    if (!sa) return ORC_ERR_NOMEM;
    g_sa_fn(w->buf, sa, (int32_t)w->len, 0, NULL);                  /* :29-37, rc ignored like the reference */
    uint32_t sab = (uint32_t)(w->len * 4);                          /* :116 (wraps like the reference) */
    int ok = fwrite(&sab, 4, 1, w->f) == 1 && fwrite(sa, 4, w->len, w->f) == w->len;  /* :117-119 */
    free(sa);
    if (!ok) return ORC_ERR_IO;
    w->len = 0;                                 /* :121 */
    return ORC_OK;
}

int oracle_writer_add_entry(oracle_writer *w, const uint8_t *text, size_t len) {
    if (len > w->cap) return ORC_ERR_TOOBIG;    /* :92-94 */
    if (w->len + len + 1 > w->cap) {            /* :96-98 */
        int rc = oracle_writer_dump_data(w);
        if (rc) return rc;
    }
    return w_append(w, text, len);              /* :99-100 */
}

int oracle_writer_add_entries_from_file_lines(oracle_writer *w, const char *path) {
    FILE *in = fopen(path, "rb");               /* :71 */
    if (!in) return errno == ENOENT ? ORC_ERR_NOTFOUND : ORC_ERR_IO;
    char *line = NULL;
    size_t lcap = 0;
    ssize_t got;
    int rc = ORC_OK;
    while ((got = getline(&line, &lcap, in)) >= 0) {   /* bstr for_byte_line, :73 */
        size_t n = (size_t)got;
        if (n && line[n - 1] == '\n') {
            --n;
            if (n && line[n - 1] == '\r') --n;
        }
        if (w->len + n + 1 > w->cap) {          /* :75-77 — note: no "too big" check here */
            rc = oracle_writer_dump_data(w);
            if (rc) break;
        }
        rc = w_append(w, (const uint8_t *)line, n);    /* :78-79 */
        if (rc) break;
    }
    free(line);
    fclose(in);
    return rc;
}

int oracle_writer_finalize(oracle_writer *w) {
    if (w->len) {                               /* :129-131 */
        int rc = oracle_writer_dump_data(w);
        if (rc) return rc;
    }
    return fflush(w->f) == 0 ? ORC_OK : ORC_ERR_IO;    /* :132 */
}

int oracle_writer_close(oracle_writer *w) {     /* Drop, :138-144 */
    if (!w) return ORC_OK;
    int rc = oracle_writer_finalize(w);
    fclose(w->f);
    free(w->buf);
    free(w);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* Reader                                                                                */
/* ------------------------------------------------------------------------------------ */
typedef struct oracle_chunk {
    uint8_t *data;
    size_t   len;
    size_t   sa_start;   /* suffixes_file_start, lib.rs:180 */
    size_t   sa_end;     /* suffixes_file_end,   lib.rs:181 */
} oracle_chunk;

typedef struct oracle_reader {
    int fd;
    size_t n_chunks;
    oracle_chunk *chunks;
} oracle_reader;

typedef struct oracle_hits {
    size_t   count, cap;
    int32_t *chunk;
    uint32_t *start;   /* line_tail */
    uint32_t *end;     /* line_head */
    uint64_t  n_matches;   /* matching suffixes before dedup */
    uint64_t  n_probes;    /* SA probes issued */
} oracle_hits;

void oracle_reader_close(oracle_reader *r) {
    if (!r) return;
    for (size_t i = 0; i < r->n_chunks; ++i) free(r->chunks[i].data);
    free(r->chunks);
    if (r->fd >= 0) close(r->fd);
    free(r);
}

static int read_full(int fd, void *dst, size_t n, size_t off) {
    uint8_t *p = (uint8_t *)dst;
    while (n) {
        ssize_t g = pread(fd, p, n, (off_t)off);
        if (g <= 0) return -1;
        p += g; off += (size_t)g; n -= (size_t)g;
    }
    return 0;
}

oracle_reader *oracle_reader_open(const char *path, int *status) {
    int fd = open(path, O_RDONLY);              /* lib.rs:166 */
    if (fd < 0) { if (status) *status = errno == ENOENT ? ORC_ERR_NOTFOUND : ORC_ERR_IO; return NULL; }
    struct stat st;
    if (fstat(fd, &st)) { close(fd); if (status) *status = ORC_ERR_IO; return NULL; }
    oracle_reader *r = (oracle_reader *)calloc(1, sizeof(*r));
    r->fd = fd;
    size_t file_len = (size_t)st.st_size, pos = 0, ccap = 0;
    while (pos < file_len) {                    /* :174 */
        uint32_t dlen, slen;
        if (read_full(fd, &dlen, 4, pos)) goto bad;                 /* :175 */
        uint8_t *data = (uint8_t *)malloc(dlen ? dlen : 1);
        if (!data) goto bad;
        if (read_full(fd, data, dlen, pos + 4)) { free(data); goto bad; }   /* :176-177 */
        if (read_full(fd, &slen, 4, pos + 4 + dlen)) { free(data); goto bad; } /* :179 */
        if (r->n_chunks == ccap) {
            ccap = ccap ? ccap * 2 : 4;
            r->chunks = (oracle_chunk *)realloc(r->chunks, ccap * sizeof(oracle_chunk));
        }
        oracle_chunk *c = &r->chunks[r->n_chunks++];
        c->data = data;
        c->len = dlen;
        c->sa_start = pos + 8 + dlen;           /* :180 */
        c->sa_end = c->sa_start + slen;         /* :181 */
        pos += 8 + (size_t)dlen + (size_t)slen; /* :184 */
    }
    if (status) *status = ORC_OK;
    return r;
bad:
    oracle_reader_close(r);
    if (status) *status = ORC_ERR_FORMAT;
    return NULL;
}

size_t oracle_reader_num_chunks(const oracle_reader *r) { return r->n_chunks; }
const uint8_t *oracle_reader_chunk_text(const oracle_reader *r, size_t c, size_t *len) {
    *len = r->chunks[c].len;
    return r->chunks[c].data;
}
/* Copies SA[0..n) of chunk c (for tests). */
int oracle_reader_chunk_sa(const oracle_reader *r, size_t c, int32_t *out) {
    const oracle_chunk *ch = &r->chunks[c];
    return read_full(r->fd, out, ch->sa_end - ch->sa_start, ch->sa_start) ? ORC_ERR_IO : ORC_OK;
}

/* One SA probe the way the reference does it: BufReader::seek drops its buffer, then
 * read_i32 refills it with one read of up to 8 KiB (lib.rs:216-217). */
static int32_t probe(const oracle_reader *r, const oracle_chunk *c, size_t off, oracle_hits *h) {
    uint8_t buf[8192];
    size_t want = sizeof(buf);
    struct stat st;
    (void)st;
    ssize_t g = pread(r->fd, buf, want, (off_t)off);
    int32_t v = 0;
    if (g >= 4) memcpy(&v, buf, 4);
    (void)c;
    h->n_probes++;
    return v;
}

/* Ordering of `sub` against the suffix (lib.rs:220-228): 0 = suffix starts with sub,
 * <0 = sub < suffix, >0 = sub > suffix (slice cmp: memcmp then length). */
static int cmp_sub(const uint8_t *sub, size_t m, const uint8_t *suf, size_t sl) {
    size_t k = m < sl ? m : sl;
    int c = k ? memcmp(sub, suf, k) : 0;
    if (c) return c;
    if (sl >= m) return 0;   /* starts_with */
    return 1;                /* suffix is a proper prefix of sub → sub is greater */
}

/* open-addressing set of entry starts (AHashSet<usize>, lib.rs:262) */
typedef struct { uint64_t *slot; size_t cap, used; } u64set;
static int set_insert(u64set *s, uint64_t key) {
    if ((s->used + 1) * 2 > s->cap) {
        size_t nc = s->cap ? s->cap * 2 : 64;
        uint64_t *ns = (uint64_t *)malloc(nc * sizeof(uint64_t));
        memset(ns, 0xFF, nc * sizeof(uint64_t));
        for (size_t i = 0; i < s->cap; ++i)
            if (s->slot[i] != UINT64_MAX) {
                size_t p = (size_t)((s->slot[i] * 0x9E3779B97F4A7C15ull) >> 17) & (nc - 1);
                while (ns[p] != UINT64_MAX) p = (p + 1) & (nc - 1);
                ns[p] = s->slot[i];
            }
        free(s->slot);
        s->slot = ns;
        s->cap = nc;
    }
    size_t p = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 17) & (s->cap - 1);
    while (s->slot[p] != UINT64_MAX) {
        if (s->slot[p] == key) return 0;
        p = (p + 1) & (s->cap - 1);
    }
    s->slot[p] = key;
    s->used++;
    return 1;
}

static int hits_push(oracle_hits *h, int32_t chunk, uint32_t s, uint32_t e) {
    if (h->count == h->cap) {
        size_t nc = h->cap ? h->cap * 2 : 256;
        h->chunk = (int32_t *)realloc(h->chunk, nc * sizeof(int32_t));
        h->start = (uint32_t *)realloc(h->start, nc * sizeof(uint32_t));
        h->end = (uint32_t *)realloc(h->end, nc * sizeof(uint32_t));
        if (!h->chunk || !h->start || !h->end) return ORC_ERR_NOMEM;
        h->cap = nc;
    }
    h->chunk[h->count] = chunk;
    h->start[h->count] = s;
    h->end[h->count] = e;
    h->count++;
    return ORC_OK;
}

static int search_chunk(const oracle_reader *r, size_t ci, const uint8_t *sub, size_t m, oracle_hits *h) {
    const oracle_chunk *c = &r->chunks[ci];
    if (c->sa_end < c->sa_start + 4) return ORC_OK;
    int have_start = 0, have_end = 0;
    size_t start = 0, end = 0;
    /* byte offsets into the file, signed so that `mid - 4` below the range is representable
     * (the reference's usize arithmetic never underflows because sa_start >= 8) */
    int64_t left = (int64_t)c->sa_start, right = (int64_t)c->sa_end - 4;      /* :212-213 */
    while (left <= right) {                                                  /* :214 */
        int64_t mid = left + ((right - left) / 4 / 2 * 4);                   /* :215 */
        int32_t di = probe(r, c, (size_t)mid, h);                            /* :216-217 */
        int o = cmp_sub(sub, m, c->data + di, c->len - (size_t)di);          /* :219-228 */
        if (o == 0) { have_start = 1; start = (size_t)mid; right = mid - 4; }
        else if (o < 0) right = mid - 4;
        else left = mid + 4;
    }
    if (!have_start) return ORC_OK;                                          /* :231-233 */
    right = (int64_t)c->sa_end - 4;                                          /* :235 (left is kept) */
    while (left <= right) {                                                  /* :236 */
        int64_t mid = left + ((right - left) / 4 / 2 * 4);
        int32_t di = probe(r, c, (size_t)mid, h);
        int o = cmp_sub(sub, m, c->data + di, c->len - (size_t)di);
        if (o == 0) { have_end = 1; end = (size_t)mid; left = mid + 4; }     /* :242-244 */
        else if (o < 0) right = mid - 4;
        else left = mid + 4;
    }
    if (!have_end) return ORC_ERR_FORMAT;  /* reference would panic on unwrap(); unreachable for a valid SA */

    size_t nbytes = end - start + 4;                                         /* :257 */
    int32_t *suf = (int32_t *)malloc(nbytes);
    if (!suf) return ORC_ERR_NOMEM;
    if (read_full(r->fd, suf, nbytes, start)) { free(suf); return ORC_ERR_IO; }  /* :259-260 */
    u64set seen = {0};
    int rc = ORC_OK;
    for (size_t k = 0; k < nbytes / 4 && rc == ORC_OK; ++k) {                /* :264 */
        size_t di = (size_t)suf[k];
        const uint8_t *nl = (const uint8_t *)memchr(c->data + di, '\n', c->len - di);
        size_t head = nl ? (size_t)(nl - c->data) : c->len - 1;             /* :266-269 */
        const uint8_t *pl = di ? (const uint8_t *)memrchr(c->data, '\n', di) : NULL;
        size_t tail = pl ? (size_t)(pl - c->data) + 1 : 0;                  /* :270-273 */
        h->n_matches++;
        if (set_insert(&seen, (uint64_t)tail))                              /* :274 */
            rc = hits_push(h, (int32_t)ci, (uint32_t)tail, (uint32_t)head); /* :275-276 */
    }
    free(seen.slot);
    free(suf);
    return rc;
}

oracle_hits *oracle_hits_new(void) { return (oracle_hits *)calloc(1, sizeof(oracle_hits)); }
void oracle_hits_free(oracle_hits *h) {
    if (!h) return;
    free(h->chunk); free(h->start); free(h->end); free(h);
}
void oracle_hits_clear(oracle_hits *h) { h->count = 0; h->n_matches = 0; h->n_probes = 0; }
size_t oracle_hits_count(const oracle_hits *h) { return h->count; }
const int32_t *oracle_hits_chunk(const oracle_hits *h) { return h->chunk; }
const uint32_t *oracle_hits_start(const oracle_hits *h) { return h->start; }
const uint32_t *oracle_hits_end(const oracle_hits *h) { return h->end; }
uint64_t oracle_hits_matches(const oracle_hits *h) { return h->n_matches; }
uint64_t oracle_hits_probes(const oracle_hits *h) { return h->n_probes; }

/* Reader::search: appends this query's entries to `h` (so search_multiple is a loop of
 * calls on the same `h`, __init__.py:65-71). */
int oracle_reader_search(const oracle_reader *r, const uint8_t *sub, size_t m, oracle_hits *h) {
    for (size_t ci = 0; ci < r->n_chunks; ++ci) {                            /* :207, chunk order */
        int rc = search_chunk(r, ci, sub, m, h);
        if (rc) return rc;
    }
    return ORC_OK;
}

/* search_multiple over packed patterns; per_query_count[q] receives the number of entries
 * query q contributed.  Returns total wall seconds spent (for the CPU baseline) via *secs. */
int oracle_reader_search_multiple(const oracle_reader *r, const uint8_t *patterns, const int64_t *offsets,
                                  int32_t nq, oracle_hits *h, int64_t *per_query_count) {
    for (int32_t q = 0; q < nq; ++q) {
        size_t before = h->count;
        int rc = oracle_reader_search(r, patterns + offsets[q], (size_t)(offsets[q + 1] - offsets[q]), h);
        if (rc) return rc;
        if (per_query_count) per_query_count[q] = (int64_t)(h->count - before);
    }
    return ORC_OK;
}
