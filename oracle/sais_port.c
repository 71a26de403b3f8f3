/*
 * sais_port.c — ORACLE (test infrastructure only; never linked into the product).
 *
 * CPU restatement of what the reference's BUILD path computes:
 *   src/lib.rs:24-40  construct_suffix_array → libsais(T, SA, n, 0, NULL)
 *   src/libsais/libsais.c:6597-6610 (libsais) → :6500 (libsais_main) → :6458 (libsais_main_8u)
 *
 * libsais is an engineered SA-IS (suffix array by induced sorting, Nong/Zhang/Chan 2009):
 * classify suffixes S/L, bucket the LMS suffixes (libsais.c:692-736, :1537-1561), induce the
 * order of LMS substrings (two scans, :2105/:2936), name them (:3853-3881), recurse on the
 * reduced string if names collide (:6481), then induce the final order (:4565/:5194).
 * This file restates that algorithm in its plain textbook form — one recursive function
 * over int32 symbols with an explicit smallest sentinel — not libsais' code.
 *
 * Output contract (what parity is checked on): SA[0..n) is the permutation of 0..n-1 that
 * orders the suffixes of T by unsigned-byte lexicographic order, a proper prefix first
 * (equivalently: a virtual sentinel smaller than every byte terminates T).  The suffix
 * array of a text is unique, so any correct construction is bit-identical to libsais'.
 *
 * Pinning: tests/test_oracle.py checks this port against (i) brute-force suffix sorting,
 * (ii) the reference's own libsais.c compiled unmodified into oracle/_ref/ (when present),
 * (iii) the committed golden vectors in tests/golden/ that were generated from (ii).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TGET(i) ((types[(i) >> 3] >> ((i) & 7)) & 1)
#define TSET(i, b)                                                         \
    do {                                                                   \
        if (b) types[(i) >> 3] |= (uint8_t)(1u << ((i) & 7));              \
        else   types[(i) >> 3] &= (uint8_t)~(1u << ((i) & 7));             \
    } while (0)
#define IS_LMS(i) ((i) > 0 && TGET(i) && !TGET((i) - 1))

static void bucket_bounds(const int32_t *s, int32_t *bkt, int32_t n, int32_t K, int ends) {
    int32_t i, sum = 0;
    for (i = 0; i < K; ++i) bkt[i] = 0;
    for (i = 0; i < n; ++i) bkt[s[i]]++;
    for (i = 0; i < K; ++i) {
        sum += bkt[i];
        bkt[i] = ends ? sum : sum - bkt[i];
    }
}

/* left-to-right scan: place L-type predecessors at bucket heads */
static void induce_l(const uint8_t *types, int32_t *SA, const int32_t *s, int32_t *bkt, int32_t n, int32_t K) {
    int32_t i, j;
    bucket_bounds(s, bkt, n, K, 0);
    for (i = 0; i < n; ++i) {
        j = SA[i] - 1;
        if (SA[i] > 0 && !TGET(j)) SA[bkt[s[j]]++] = j;
    }
}

/* right-to-left scan: place S-type predecessors at bucket tails */
static void induce_s(const uint8_t *types, int32_t *SA, const int32_t *s, int32_t *bkt, int32_t n, int32_t K) {
    int32_t i, j;
    bucket_bounds(s, bkt, n, K, 1);
    for (i = n - 1; i >= 0; --i) {
        j = SA[i] - 1;
        if (SA[i] > 0 && TGET(j)) SA[--bkt[s[j]]] = j;
    }
}

/* s[0..n) over alphabet [0,K), s[n-1] == 0 is the unique smallest symbol. */
static int sais_rec(const int32_t *s, int32_t *SA, int32_t n, int32_t K) {
    int32_t i, j, n1, names, prev;
    uint8_t *types = (uint8_t *)calloc((size_t)n / 8 + 1, 1);
    int32_t *bkt = (int32_t *)malloc(sizeof(int32_t) * (size_t)K);
    if (!types || !bkt) { free(types); free(bkt); return -2; }

    if (n == 1) { SA[0] = 0; free(types); free(bkt); return 0; }
    TSET(n - 1, 1);
    TSET(n - 2, 0);
    for (i = n - 3; i >= 0; --i)
        TSET(i, (s[i] < s[i + 1] || (s[i] == s[i + 1] && TGET(i + 1))) ? 1 : 0);

    /* stage 1: sort the LMS substrings by induced sorting */
    bucket_bounds(s, bkt, n, K, 1);
    for (i = 0; i < n; ++i) SA[i] = -1;
    for (i = 1; i < n; ++i)
        if (IS_LMS(i)) SA[--bkt[s[i]]] = i;
    induce_l(types, SA, s, bkt, n, K);
    induce_s(types, SA, s, bkt, n, K);

    /* compact the sorted LMS positions, then name the substrings */
    n1 = 0;
    for (i = 0; i < n; ++i)
        if (SA[i] >= 0 && IS_LMS(SA[i])) SA[n1++] = SA[i];
    for (i = n1; i < n; ++i) SA[i] = -1;
    names = 0;
    prev = -1;
    for (i = 0; i < n1; ++i) {
        int32_t pos = SA[i], d;
        int diff = 0;
        for (d = 0; d < n; ++d) {
            if (prev == -1 || s[pos + d] != s[prev + d] || TGET(pos + d) != TGET(prev + d)) { diff = 1; break; }
            if (d > 0 && (IS_LMS(pos + d) || IS_LMS(prev + d))) break;
        }
        if (diff) { ++names; prev = pos; }
        SA[n1 + pos / 2] = names - 1;
    }
    for (i = n - 1, j = n - 1; i >= n1; --i)
        if (SA[i] >= 0) SA[j--] = SA[i];

    /* stage 2: suffix array of the reduced string (recursion only if names collide) */
    {
        int32_t *SA1 = SA, *s1 = SA + n - n1;
        if (names < n1) {
            int rc = sais_rec(s1, SA1, n1, names);
            if (rc) { free(types); free(bkt); return rc; }
        } else {
            for (i = 0; i < n1; ++i) SA1[s1[i]] = i;
        }

        /* stage 3: seed the LMS suffixes in their final relative order and induce the rest */
        bucket_bounds(s, bkt, n, K, 1);
        for (i = 1, j = 0; i < n; ++i)
            if (IS_LMS(i)) s1[j++] = i;
        for (i = 0; i < n1; ++i) SA1[i] = s1[SA1[i]];
        for (i = n1; i < n; ++i) SA[i] = -1;
        for (i = n1 - 1; i >= 0; --i) {
            j = SA[i];
            SA[i] = -1;
            SA[--bkt[s[j]]] = j;
        }
    }
    induce_l(types, SA, s, bkt, n, K);
    induce_s(types, SA, s, bkt, n, K);
    free(types);
    free(bkt);
    return 0;
}

/*
 * Same signature and return codes as the reference's libsais (src/libsais/libsais.h:56-65;
 * libsais.c:6599-6609): 0 ok, -1 bad arguments, -2 allocation failure.
 */
int32_t oracle_libsais(const uint8_t *T, int32_t *SA, int32_t n, int32_t fs, int32_t *freq) {
    int32_t i, rc;
    int32_t *s, *sa;
    if (T == NULL || SA == NULL || n < 0 || fs < 0) return -1;
    if (freq) {
        memset(freq, 0, 256 * sizeof(int32_t));
        for (i = 0; i < n; ++i) freq[T[i]]++;
    }
    if (n == 0) return 0;
    if (n == 1) { SA[0] = 0; return 0; }
    s = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 1));
    sa = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 1));
    if (!s || !sa) { free(s); free(sa); return -2; }
    for (i = 0; i < n; ++i) s[i] = (int32_t)T[i] + 1;
    s[n] = 0; /* the virtual sentinel: makes a proper prefix sort first */
    rc = sais_rec(s, sa, n + 1, 257);
    if (rc == 0) memcpy(SA, sa + 1, sizeof(int32_t) * (size_t)n); /* sa[0] == n is the sentinel suffix */
    free(s);
    free(sa);
    return rc;
}
