"""Deterministic synthetic corpora for the benchmark configs of BASELINE.json / SURVEY §8(d).

Shared by bench.py and the tests; numpy only (vectorised: the 500 MB config builds in
seconds, not minutes).
"""
import numpy as np

_LETTERS = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz", dtype=np.uint8)


def zipf_words_text(n, seed=20240501, vocab=65536, s=1.1, newline_p=1.0 / 6.0, block=1 << 22):
    """n bytes of newline-delimited lowercase pseudo-words (config 1): vocabulary of
    `vocab` words with lengths U[3,10], word ids Zipf(p ∝ rank^-s), each word followed by
    '\\n' with probability newline_p, else ' '.  The last byte is forced to '\\n'."""
    rng = np.random.default_rng(seed)
    wlen = rng.integers(3, 11, size=vocab).astype(np.int64)
    width = 11  # longest word (10) + its separator slot
    vmat = _LETTERS[rng.integers(0, 26, size=(vocab, width))]
    p = 1.0 / np.arange(1, vocab + 1, dtype=np.float64) ** s
    cdf = np.cumsum(p / p.sum())
    col = np.arange(width, dtype=np.int64)[None, :]
    out = np.empty(n, dtype=np.uint8)
    filled = 0
    while filled < n:
        ids = np.searchsorted(cdf, rng.random(block), side="right")
        np.minimum(ids, vocab - 1, out=ids)
        seps = np.where(rng.random(block) < newline_p, 10, 32).astype(np.uint8)
        lens = wlen[ids]
        rows = vmat[ids]                                  # block x 11 letters
        rows[np.arange(block), lens] = seps               # separator right after the word
        chunk = rows[col <= lens[:, None]]                # row-major flatten of word+sep
        take = min(len(chunk), n - filled)
        out[filled:filled + take] = chunk[:take]
        filled += take
    text = out[:n]
    text[n - 1] = 10
    return text


def plant(text, word, count, seed):
    """Overwrite `count` distinct lines' interiors with `word` (config 1 plants 'google'
    in 5 943 lines and 'text_two' in 159, README.md:49,51).  Returns the line count hit."""
    rng = np.random.default_rng(seed)
    w = np.frombuffer(word.encode(), dtype=np.uint8)
    nl = np.flatnonzero(text == 10)
    starts = np.concatenate(([0], nl[:-1] + 1))
    lens = nl - starts
    ok = np.flatnonzero(lens >= len(w))
    pick = rng.choice(ok, size=min(count, len(ok)), replace=False)
    for li in pick:
        room = int(lens[li]) - len(w)
        o = int(starts[li]) + (int(rng.integers(0, room + 1)) if room > 0 else 0)
        text[o:o + len(w)] = w
    return len(pick)


def config1_text(n=500_000_000, seed=20240501):
    t = zipf_words_text(n, seed=seed)
    plant(t, "google", 5943, seed + 1)
    plant(t, "text_two", 159, seed + 2)
    return t


def _gram_codes(text):
    lut = np.full(256, 31, dtype=np.int32)
    lut[97:123] = np.arange(26)
    lut[32], lut[10], lut[95] = 26, 27, 28
    return lut


def fourgram_counts(text, block=1 << 26):
    """Occurrences of every 4-gram (5-bit symbol codes → 2^20 bins; foreign bytes share a
    code, which only over-counts)."""
    lut = _gram_codes(text)
    counts = np.zeros(1 << 20, dtype=np.int64)
    n = len(text)
    for lo in range(0, n - 3, block):
        c = lut[text[lo:min(n, lo + block + 3)]]
        g = (c[:-3] << 15) | (c[1:-2] << 10) | (c[2:-1] << 5) | c[3:]
        counts += np.bincount(g, minlength=1 << 20)
    return counts


def config2_queries(text, nq=10_000, seed=7, hit_frac=0.9, max_count=5000):
    """nq patterns, length U[4,32]: hit_frac cut from the text at uniform random offsets
    (may cross '\\n'), the rest random lowercase; shuffled.  Returns a list of bytes.

    Cut patterns are rejection-sampled to be SELECTIVE: a candidate is kept only if its
    rarest 4-gram occurs at most `max_count` times in the text (an upper bound on its own
    hit count; 5 000 mirrors the reference README's 159- and 5 943-result queries).
    Without this, uniformly cut 4-grams of Zipf text match millions of lines each and a
    10 k batch would return billions of strings — the high-hit regime is measured on its
    own as config 5, not smeared over config 2."""
    rng = np.random.default_rng(seed)
    n_hit = int(nq * hit_frac)
    counts = fourgram_counts(text)
    lut = _gram_codes(text)
    pats = []
    while len(pats) < n_hit:
        m = 4 * n_hit
        offs = rng.integers(0, len(text) - 36, size=m)
        lens = rng.integers(4, 33, size=m)
        best = np.full(m, np.iinfo(np.int64).max, dtype=np.int64)
        for j in range(29):
            c = [lut[text[offs + j + k]] for k in range(4)]
            g = (c[0] << 15) | (c[1] << 10) | (c[2] << 5) | c[3]
            cnt = counts[g]
            best = np.where(j <= lens - 4, np.minimum(best, cnt), best)
        for k in np.flatnonzero(best <= max_count):
            pats.append(bytes(text[offs[k]:offs[k] + lens[k]]))
            if len(pats) == n_hit:
                break
    lens = rng.integers(4, 33, size=nq - n_hit)
    for k in range(nq - n_hit):
        pats.append(bytes(_LETTERS[rng.integers(0, 26, size=lens[k])]))
    order = rng.permutation(nq)
    return [pats[i] for i in order]


def acgt_text(n, seed=4, base_len=1 << 20, mut_every=1 << 16, nl_every=1000):
    """Config 4a: random ACGT base block tiled, one point mutation per 64 KiB, '\\n' every
    1000 symbols — long repeats, the worst case for prefix-doubling depth."""
    rng = np.random.default_rng(seed)
    sym = np.frombuffer(b"ACGT", dtype=np.uint8)
    base = sym[rng.integers(0, 4, size=min(base_len, n))]
    t = np.tile(base, n // len(base) + 1)[:n].copy()
    mpos = np.arange(mut_every // 2, n, mut_every)
    t[mpos] = sym[rng.integers(0, 4, size=len(mpos))]
    t[nl_every - 1::nl_every] = 10
    t[n - 1] = 10
    return t


def pack_patterns(pats):
    """bytes list → (uint8 blob, int64 offsets[nq+1]) as the C ABI takes them."""
    offs = np.zeros(len(pats) + 1, dtype=np.int64)
    if pats:
        np.cumsum([len(p) for p in pats], out=offs[1:])
    blob = np.frombuffer(b"".join(pats) + b"\0", dtype=np.uint8).copy()
    return blob, offs
