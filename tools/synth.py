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


# --------------------------------------------------------------------------------------
# Config 3 (BASELINE configs[2]): 15 chunks x 2^29 bytes, ONE shared vocabulary, generated with
# torch ops from a counter-based hash, so that the same bytes come out of a GPU (bench ranks:
# 0.3 s per chunk) and of the CPU (reference arm).  Every chunk is an independent word
# sequence (seed 1000 + k); its first m bytes are the same whatever length is generated, so the
# query set can be cut from short prefixes of all chunks by any process.
# --------------------------------------------------------------------------------------
CONFIG3_CHUNKS = 15
CONFIG3_CHUNK_BYTES = 1 << 29
CONFIG3_PREFIX = 32 << 20

_M64 = (1 << 64) - 1


def _i64(c):
    """unsigned 64-bit constant → the int64 with the same bits"""
    c &= _M64
    return c - (1 << 64) if c >= (1 << 63) else c


def _lsr(x, k):
    """logical shift right of an int64 tensor"""
    return (x >> k) & ((1 << (64 - k)) - 1)


def _splitmix64(x):
    """splitmix64 finaliser on an int64 torch tensor (arithmetic wraps mod 2^64 on CPU and CUDA alike)"""
    z = x + _i64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


_VOCAB_CACHE = {}


def _config3_vocab(device, vocab=65536, s=1.1):
    import torch
    key = (str(device), vocab, s)
    if key not in _VOCAB_CACHE:
        rng = np.random.default_rng(20240501)
        wlen = rng.integers(3, 11, size=vocab).astype(np.int64)
        vmat = _LETTERS[rng.integers(0, 26, size=(vocab, 11))]
        p = 1.0 / np.arange(1, vocab + 1, dtype=np.float64) ** s
        cdf = np.cumsum(p / p.sum())
        _VOCAB_CACHE[key] = (torch.from_numpy(wlen).to(device), torch.from_numpy(vmat.copy()).to(device),
                             torch.from_numpy(cdf).to(device))
    return _VOCAB_CACHE[key]


def config3_chunk_torch(k, n=CONFIG3_CHUNK_BYTES, device="cpu", pad=16, words_per_block=1 << 22, force_newline=True):
    """uint8 torch tensor of n + pad bytes on `device`: chunk k of the config-3 corpus (Zipf words
    over the shared 65 536-word vocabulary, '\\n' after a word w.p. 1/6 else ' '), zero padding
    after byte n.  With force_newline the last byte is '\\n' (a chunk always ends an entry)."""
    import torch
    wlen, vmat, cdf = _config3_vocab(device)
    vocab = wlen.numel()
    out = torch.zeros(n + pad, dtype=torch.uint8, device=device)
    seed = _i64((1000 + k) * 0x9E3779B97F4A7C15)
    filled, counter = 0, 0
    cols = torch.arange(11, device=device, dtype=torch.int64)
    while filled < n:
        c = torch.arange(counter, counter + words_per_block, device=device, dtype=torch.int64)
        h1 = _splitmix64(seed ^ (c * 2))
        h2 = _splitmix64(seed ^ (c * 2 + 1))
        u = _lsr(h1, 11).to(torch.float64) * (1.0 / (1 << 53))
        ids = torch.searchsorted(cdf, u, right=True).clamp_(max=vocab - 1)
        sep = torch.where(_lsr(h2, 40) < (1 << 24) // 6, 10, 32).to(torch.uint8)
        lens = wlen[ids]
        pos = torch.cumsum(lens + 1, 0)
        total = int(pos[-1])
        pos = pos - (lens + 1) + filled
        rows = vmat[ids]                                              # W x 11 letters
        rows = torch.where(cols[None, :] == lens[:, None], sep[:, None], rows)
        at = pos[:, None] + cols[None, :]
        keep = (cols[None, :] <= lens[:, None]) & (at < n)
        out[at[keep]] = rows[keep]
        filled += total
        counter += words_per_block
    if force_newline:
        out[n - 1] = 10
    return out


def config3_chunk(k, n=CONFIG3_CHUNK_BYTES, device="cpu"):
    """numpy uint8[n]: chunk k; chunk 0 carries config 1's planted lines ('google' in 5 943 lines,
    'text_two' in 159 — README.md:49,51), the others are plain."""
    t = config3_chunk_torch(k, n, device=device)[:n].cpu().numpy()
    if k == 0:
        plant(t, "google", 5943, 20240502)
        plant(t, "text_two", 159, 20240503)
    return t


def fourgram_counts_torch(t):
    """fourgram_counts on a uint8 torch tensor (any device) → int64 tensor [2^20]"""
    import torch
    lut = torch.from_numpy(_gram_codes(None).astype(np.int64)).to(t.device)
    c = lut[t.long()]
    g = (c[:-3] << 15) | (c[1:-2] << 10) | (c[2:-1] << 5) | c[3:]
    return torch.bincount(g, minlength=1 << 20), lut


def config3_queries(prefixes, nq=10_000, seed=7, hit_frac=0.9, max_count_per_chunk=5000,
                    chunk_bytes=CONFIG3_CHUNK_BYTES):
    """config-2-style batch for the multi-chunk index: hit_frac of the nq patterns (length U[4,32])
    are cut at uniform random offsets from the prefixes of ALL chunks in turn (query j from
    chunk j % len(prefixes); may cross '\\n'), the rest are random lowercase; shuffled.

    Cut patterns are rejection-sampled to be SELECTIVE: a candidate is kept only if its rarest
    4-gram, counted on the prefix it was cut from and scaled to the chunk size, occurs at most
    `max_count_per_chunk` times per chunk (5 000 mirrors the reference README's 159- and
    5 943-result queries on its 500 MB file).  Uniformly cut 4-grams of Zipf text match millions
    of lines each; 10 000 of them would return billions of strings — that regime is config 5.

    prefixes: list of uint8 torch tensors (any device), one per chunk.  Returns a list of bytes."""
    import torch
    rng = np.random.default_rng(seed)
    n_hit = int(nq * hit_frac)
    nchunks = len(prefixes)
    per = [len(range(c, n_hit, nchunks)) for c in range(nchunks)]
    cut = []
    for c, t in enumerate(prefixes):
        m = t.numel()
        counts, lut = fourgram_counts_torch(t)
        thresh = max(1, int(max_count_per_chunk * (m / float(chunk_bytes))))
        got = []
        while len(got) < per[c]:
            k = 4 * per[c]
            offs = torch.from_numpy(rng.integers(0, m - 40, size=k)).to(t.device)
            lens = torch.from_numpy(rng.integers(4, 33, size=k)).to(t.device)
            best = torch.full((k,), 1 << 62, dtype=torch.int64, device=t.device)
            codes = [lut[t[offs + j].long()] for j in range(32)]
            for j in range(29):
                g = (codes[j] << 15) | (codes[j + 1] << 10) | (codes[j + 2] << 5) | codes[j + 3]
                best = torch.where(j <= lens - 4, torch.minimum(best, counts[g]), best)
            ok = torch.nonzero(best <= thresh).flatten().cpu().numpy()
            offs_h, lens_h = offs.cpu().numpy(), lens.cpu().numpy()
            win = t[(offs[:, None] + torch.arange(32, device=t.device)[None, :])].cpu().numpy()
            for i in ok:
                got.append(bytes(win[i, :lens_h[i]]))
                if len(got) == per[c]:
                    break
        cut.append(got)
    pats = [cut[j % nchunks][j // nchunks] for j in range(n_hit)]
    lens = rng.integers(4, 33, size=nq - n_hit)
    for k in range(nq - n_hit):
        pats.append(bytes(_LETTERS[rng.integers(0, 26, size=lens[k])]))
    order = rng.permutation(nq)
    return [pats[i] for i in order]
