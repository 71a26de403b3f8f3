"""Executable specification (numpy, CPU) of the GPU builder's algorithm — the same key packing,
active-set filtering and re-rank arithmetic as csrc/sa_build.cu, without any CUDA — checked
against the oracle.  It documents WHY the kernels are right; the kernels themselves are
checked bit-for-bit in test_gpu_parity.py."""
import numpy as np


def model_suffix_array(text, key_bits=64):
    t = np.frombuffer(bytes(text), dtype=np.uint8)
    n = len(t)
    if n == 0:
        return np.zeros(0, np.int32)
    if n == 1:
        return np.zeros(1, np.int32)
    # alphabet -> dense codes 1..sigma, 0 = past the end (presence_kernel + LUT)
    present = np.unique(t)
    lut = np.zeros(256, dtype=np.uint64)
    lut[present] = np.arange(1, len(present) + 1, dtype=np.uint64)
    b = int(len(present)).bit_length()
    m = min(key_bits // b, n)
    code = np.concatenate([lut[t], np.zeros(m, dtype=np.uint64)])
    # round 0 key: m codes packed big-endian (keygen_kernel)
    key = np.zeros(n, dtype=np.uint64)
    for j in range(m):
        key = (key << np.uint64(b)) | code[j:j + n]
    sa = np.full(n, -1, dtype=np.int64)
    isa = np.zeros(n, dtype=np.int64)          # 1-based rank, 0 = past the end
    idx = np.arange(n, dtype=np.int64)
    grp = np.zeros(n, dtype=np.int64)
    first, h = True, m
    while len(idx):
        if not first:                          # gather_kernel: (group rank, rank of suffix i + h)
            j = idx + h
            r2 = np.where(j < n, isa[np.minimum(j, n - 1)], 0)
            key = (grp.astype(np.uint64) << np.uint64(32)) | r2.astype(np.uint64)
        order = np.argsort(key, kind="stable")  # the onesweep sort
        key, idx = key[order], idx[order]
        k = np.arange(len(idx))
        g = np.zeros(len(idx), dtype=np.int64) if first else (key >> np.uint64(32)).astype(np.int64)
        hn = np.ones(len(idx), dtype=bool)
        hn[1:] = key[1:] != key[:-1]                       # new-group heads
        ho = np.zeros(len(idx), dtype=bool)
        ho[0] = True
        if not first:
            ho[1:] = g[1:] != g[:-1]                       # old-group heads
        A = np.maximum.accumulate(np.where(ho, k, 0))      # the 3-component scan of rerank_apply
        B = np.maximum.accumulate(np.where(hn, k, 0))
        p = g + (k - A)                                    # final SA slot inside the old group
        ng = g + (B - A)                                   # rank of the new group
        nhn = np.ones(len(idx), dtype=bool)
        nhn[:-1] = hn[1:]
        single = hn & nhn
        sa[p[single]] = idx[single]                        # settled suffixes leave the active set
        isa[idx] = np.where(single, p, ng) + 1
        idx, grp = idx[~single], ng[~single]
        if not first:
            h *= 2                                         # round 0 sorted m symbols: round 1 looks at i + m
        first = False
        assert h <= 4 * n + 64, "prefix doubling failed to converge"
    return sa.astype(np.int32)


def test_model_matches_oracle(oracle):
    rng = np.random.default_rng(3)
    cases = [b"ab", b"banana", b"mississippi\n", b"aaaaaaaaaaaaaaaa", b"x\x00x", b"\n\n\n", b"abab" * 40 + b"\n"]
    for _ in range(60):
        n = int(rng.integers(2, 300))
        cases.append(bytes(rng.choice(np.array([0, 10, 97, 98, 255], dtype=np.uint8), size=n)))
    for _ in range(10):
        cases.append(bytes(rng.integers(0, 256, size=int(rng.integers(2, 500)), dtype=np.uint8)))
    for t in cases:
        assert model_suffix_array(t).tolist() == oracle.suffix_array_port(t).tolist(), t[:40]


def test_model_with_narrow_keys_needs_more_rounds_but_same_answer(oracle):
    """The packed-prefix width only changes the number of doubling rounds, never the result
    (the PSS_H0 knob)."""
    t = (b"the quick brown fox jumps over the lazy dog\n" * 7) + b"the quick brown cat\n"
    want = oracle.suffix_array_port(t).tolist()
    for bits in (64, 16, 5):
        assert model_suffix_array(t, key_bits=bits).tolist() == want


# ---- extraction through the line directory (csrc/search.cu: line_directory_kernel, entry_bounds_dir) ----
LINE_BLOCK = 256


def model_entry_bounds_dir(nl, dirv, n, pos):
    """Statement-for-statement model of entry_bounds_dir: nine offsets nl[lo-1 .. lo+8) decide,
    else a search inside the rest of the block."""
    L = len(nl)
    if L == 0:
        return 0, n - 1
    blk = pos // LINE_BLOCK
    lo = int(dirv[blk])
    e, b = 0xFFFFFFFF, 0
    for j in range(9):
        idx = lo + j - 1
        v = int(nl[idx]) if 0 <= idx < L else 0xFFFFFFFF
        if v < pos:
            b = v + 1
        else:
            e = min(e, v)
    if e == 0xFFFFFFFF:
        if lo + 8 >= L:
            e = n - 1
        else:
            lo2, hi = lo + 8, int(dirv[blk + 1])
            while lo2 < hi:
                mid = lo2 + ((hi - lo2) >> 1)
                if nl[mid] < pos:
                    lo2 = mid + 1
                else:
                    hi = mid
            e = int(nl[lo2]) if lo2 < L else n - 1
            b = int(nl[lo2 - 1]) + 1
    return b, e


def reference_entry_bounds(t, pos):
    """lib.rs:266-273: first '\\n' at or after pos (none: len - 1); 1 + last '\\n' before pos (none: 0)."""
    n = len(t)
    e = t.find(b"\n", pos)
    e = n - 1 if e < 0 else e
    b = t.rfind(b"\n", 0, pos) + 1
    return b, e


def test_line_directory_model_matches_reference_scans():
    rng = np.random.default_rng(11)
    texts = [b"\n" * 700, b"a" * 1000, b"ab\n", b"x", b"\n", b"a" * 255 + b"\n" + b"b" * 300 + b"\n\n\n" + b"c" * 513]
    for p_nl in (0.5, 0.1, 0.02, 0.001):
        for n in (1, 255, 256, 257, 2049, 5000):
            a = rng.integers(97, 100, size=n, dtype=np.uint8)
            a[rng.random(n) < p_nl] = 10
            texts.append(a.tobytes())
            a2 = a.copy()
            a2[-1] = 10                       # Writer-made chunks end in '\n'
            texts.append(a2.tobytes())
    for t in texts:
        n = len(t)
        arr = np.frombuffer(t, dtype=np.uint8)
        nl = np.flatnonzero(arr == 10).astype(np.int64)
        n_dir = (n + LINE_BLOCK - 1) // LINE_BLOCK + 1
        dirv = np.searchsorted(nl, np.arange(n_dir, dtype=np.int64) * LINE_BLOCK, side="left")  # line_directory_kernel
        assert dirv[-1] == len(nl)
        for pos in (range(n) if n <= 600 else rng.integers(0, n, size=600)):
            assert model_entry_bounds_dir(nl, dirv, n, int(pos)) == reference_entry_bounds(t, int(pos)), (n, int(pos))


# ---- extraction through the line records (csrc/search.cu: line_records_kernel, entry_bounds_rec) ----
LINE_REC_BLOCK = 128
M64 = (1 << 64) - 1


def model_line_records(nl, n):
    """line_records_kernel: per 128-byte block (start of the entry open at its first byte, first
    '\\n' at or after its end else n - 1, 128-bit newline map as two 64-bit halves)."""
    recs = []
    for j in range((n + LINE_REC_BLOCK - 1) // LINE_REC_BLOCK):
        at, end = j * LINE_REC_BLOCK, (j + 1) * LINE_REC_BLOCK
        lo = int(np.searchsorted(nl, at, side="left"))
        open_at = int(nl[lo - 1]) + 1 if lo > 0 else 0
        k, bm = lo, 0
        while k < len(nl) and nl[k] < end:
            bm |= 1 << int(nl[k] - at)
            k += 1
        nxt = int(nl[k]) if k < len(nl) else n - 1
        recs.append((open_at, nxt, bm & M64, bm >> 64))
    return recs


def model_entry_bounds_rec(recs, pos):
    blk, off = divmod(pos, LINE_REC_BLOCK)
    open_at, nxt, lo64, hi64 = recs[blk]
    base = blk * LINE_REC_BLOCK
    f_lo = (lo64 & ((M64 << off) & M64)) if off < 64 else 0
    f_hi = hi64 if off < 64 else (hi64 & ((M64 << (off - 64)) & M64))
    e = nxt
    if f_lo:
        e = base + ((f_lo & -f_lo).bit_length() - 1)
    elif f_hi:
        e = base + 64 + ((f_hi & -f_hi).bit_length() - 1)
    b_lo = ((lo64 & (M64 >> (64 - off))) if off else 0) if off < 64 else lo64
    b_hi = (hi64 & (M64 >> (128 - off))) if off > 64 else 0
    b = open_at
    if b_hi:
        b = base + 64 + b_hi.bit_length()
    elif b_lo:
        b = base + b_lo.bit_length()
    return b, e


def test_line_records_model_matches_reference_scans():
    rng = np.random.default_rng(12)
    texts = [b"\n" * 300, b"a" * 700, b"ab\n", b"x", b"\n", b"a" * 127 + b"\n" + b"b" * 200 + b"\n\n\n" + b"c" * 257,
             b"a" * 63 + b"\n" + b"a" * 64 + b"\n" + b"a" * 63 + b"\n\n" + b"q" * 62 + b"\n"]
    for p_nl in (0.5, 0.1, 0.02, 0.002):
        for n in (1, 63, 64, 65, 127, 128, 129, 1000, 3000):
            a = rng.integers(97, 100, size=n, dtype=np.uint8)
            a[rng.random(n) < p_nl] = 10
            texts.append(a.tobytes())
            a2 = a.copy()
            a2[-1] = 10
            texts.append(a2.tobytes())
    for t in texts:
        n = len(t)
        nl = np.flatnonzero(np.frombuffer(t, dtype=np.uint8) == 10).astype(np.int64)
        recs = model_line_records(nl, n)
        for pos in (range(n) if n <= 700 else rng.integers(0, n, size=700)):
            assert model_entry_bounds_rec(recs, int(pos)) == reference_entry_bounds(t, int(pos)), (n, int(pos))


# ---- Writer.add_entries_from_file_lines: bulk append of whole runs of lines (csrc/host_index.cu) ----
def model_ingest_chunks(data, capacity, block=1 << 20):
    """Statement-for-statement model of pss_writer_add_entries_from_file_lines + finalize: returns
    the chunk texts in order.  A run of complete records without '\\r' that fits the room left is
    appended with one copy; everything else goes record by record through the reference's flush
    rule `len + entry + 1 > capacity` (lib.rs:75) with Rust's Vec growth of the logical capacity."""
    chunks, text, line = [], bytearray(), bytearray()
    cap = [capacity]

    def dump():
        if text:
            chunks.append(bytes(text))
            text.clear()

    def reserve_logical(additional):
        if cap[0] - len(text) >= additional:
            return
        cap[0] = max(max(cap[0] * 2, len(text) + additional), 8)

    def emit(rec, terminated):
        if terminated and rec.endswith(b"\r"):
            rec = rec[:-1]
        if len(text) + len(rec) + 1 > cap[0]:
            dump()
        reserve_logical(len(rec))
        text.extend(rec)
        reserve_logical(1)
        text.extend(b"\n")

    for at in range(0, len(data), block):
        b = data[at:at + block]
        frm, no_cr = 0, b"\r" not in b
        while True:
            if no_cr and not line and len(text) < cap[0]:
                room = min(cap[0] - len(text), len(b) - frm)
                last = b.rfind(b"\n", frm, frm + room)
                if last >= 0:
                    text.extend(b[frm:last + 1])
                    frm = last + 1
            nl = b.find(b"\n", frm)
            if nl < 0:
                break
            if not line:
                emit(b[frm:nl], True)
            else:
                line.extend(b[frm:nl])
                emit(bytes(line), True)
                line.clear()
            frm = nl + 1
        line.extend(b[frm:])
    if line:
        emit(bytes(line), False)
    dump()
    return chunks


def _container_chunk_texts(path):
    raw, out, pos = open(path, "rb").read(), [], 0
    while pos < len(raw):
        n = int.from_bytes(raw[pos:pos + 4], "little")
        out.append(raw[pos + 4:pos + 4 + n])
        pos += 8 + 5 * n
    return out


def test_bulk_ingestion_model_matches_oracle_writer(tmp_path):
    """The bulk path must cut chunks exactly where the reference's per-line loop does: compared
    with the oracle Writer (restatement of lib.rs:67-86) on LF, CRLF and mixed files, chunk sizes
    below / around / above the read block, over-long lines, and a missing final newline."""
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    words = [bytes(rng.integers(97, 123, size=int(rng.integers(1, 9)), dtype=np.uint8)) for _ in range(200)]

    def lines(k, sep):
        return sep.join(b" ".join(words[int(j)] for j in rng.integers(0, 200, size=int(rng.integers(0, 12)))) for _ in range(k))

    files = [lines(3000, b"\n") + b"\n", lines(3000, b"\n"), lines(2000, b"\r\n") + b"\r\n",
             lines(1500, b"\n") + b"\r\n" + lines(1500, b"\n") + b"\n", b"", b"\n", b"x", b"\n\n\n",
             lines(500, b"\n") + b"\n" + b"q" * 5000 + b"\n" + lines(500, b"\n") + b"\n", b"a\rb\n" + lines(100, b"\n")]
    src, idx = str(tmp_path / "in.txt"), str(tmp_path / "o.idx")
    for data in files:
        for capacity, block in ((64, 256), (1000, 256), (1000, 4096), (4096, 1000), (100_000, 1 << 14), (1 << 20, 1 << 20)):
            open(src, "wb").write(data)
            w = O.Writer(idx, capacity)
            w.add_entries_from_file_lines(src)
            w.close()
            assert model_ingest_chunks(data, capacity, block) == _container_chunk_texts(idx), (len(data), capacity, block)
