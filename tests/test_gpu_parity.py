"""GPU parity tests (B200): the CUDA path, called through the C ABI (tests/capi.py) and the
Python API, against the CPU oracle (oracle/) and the committed golden fixtures.
Bit-exact everywhere: suffix arrays, index files, ordered result tuples."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from tools import synth

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------
# radix sort
# ------------------------------------------------------------------------------------
def _radix_case(pss, n, begin, end, iota, seed, kind="random"):
    import torch
    rng = np.random.default_rng(seed)
    if kind == "random":
        keys = rng.integers(0, 2**63, size=n, dtype=np.uint64) * 2 + rng.integers(0, 2, size=n, dtype=np.uint64)
    elif kind == "few":
        keys = rng.integers(0, 3, size=n, dtype=np.uint64) << np.uint64(begin)
    else:
        keys = np.full(n, 0x0123456789ABCDEF, dtype=np.uint64)
    vals = np.arange(n, dtype=np.uint32) if iota else rng.integers(0, 2**32, size=n, dtype=np.uint32)
    dk = torch.from_numpy(keys.view(np.int64)).cuda()
    dka = torch.empty_like(dk)
    dv = torch.from_numpy(vals.view(np.int32)).cuda()
    dva = torch.empty_like(dv)
    torch.cuda.synchronize()
    in_alt, npass = C.c_int32(0), C.c_int32(0)
    pss.check(pss.lib.pss_radix_sort_pairs(dk.data_ptr(), dka.data_ptr(), 0 if iota else dv.data_ptr(), dva.data_ptr(),
                                           n, begin, end, C.byref(in_alt), None, C.byref(npass), None))
    torch.cuda.synchronize()
    gk = (dka if in_alt.value else dk).cpu().numpy().view(np.uint64)
    gv = (dva if (in_alt.value or iota) else dv).cpu().numpy().view(np.uint32)
    width = end - begin
    mask = np.uint64((1 << width) - 1) if width < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    order = np.argsort((keys >> np.uint64(begin)) & mask, kind="stable")
    assert np.array_equal(gk, keys[order])
    assert np.array_equal(gv, vals[order])


@pytest.mark.parametrize("n", [1, 2, 31, 33, 4095, 4096, 4097, 100_000, (1 << 21) + 12345])
def test_radix_sort_sizes(pss, n):
    _radix_case(pss, n, 0, 64, False, n)


def test_radix_sort_bit_ranges_and_skipped_digits(pss):
    _radix_case(pss, 1 << 20, 0, 64, True, 1)
    _radix_case(pss, 1 << 20, 5, 37, False, 2)
    _radix_case(pss, 1 << 20, 0, 59, True, 3)
    _radix_case(pss, 1 << 20, 8, 24, False, 4, kind="few")
    _radix_case(pss, 1 << 20, 0, 64, True, 5, kind="const")     # every digit constant: zero passes
    _radix_case(pss, 1 << 20, 0, 64, False, 6, kind="const")


# ------------------------------------------------------------------------------------
# suffix array
# ------------------------------------------------------------------------------------
def _text(kind, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "words":
        return synth.zipf_words_text(n, seed=seed, vocab=4096, block=1 << 16)
    if kind == "bin":
        return rng.integers(0, 256, size=n, dtype=np.uint8)
    if kind == "tiny":
        return rng.choice(np.array([0, 10, 97, 98, 255], dtype=np.uint8), size=n)
    if kind == "acgt":
        return synth.acgt_text(n, seed=seed, base_len=max(64, n // 9), mut_every=max(16, n // 7))
    if kind == "same":
        return np.full(n, 97, dtype=np.uint8)
    if kind == "period":
        t = np.tile(np.frombuffer(b"ACGT", dtype=np.uint8), n // 4 + 1)[:n].copy()
        t[-1] = 10
        return t
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["tiny", "words", "bin", "acgt", "same", "period"])
def test_sa_matches_oracle(pss, oracle, kind):
    for n in [1, 2, 3, 5, 17, 100, 1000, 4096, 4097, 65536, 300_000]:
        if kind in ("same", "period") and n > 65536:
            continue
        t = _text(kind, n, n + 7)
        assert np.array_equal(pss.libsais(t), oracle.suffix_array_port(t)), (kind, n)


def test_sa_many_small_random(pss, oracle):
    rng = np.random.default_rng(99)
    for _ in range(150):
        n = int(rng.integers(2, 400))
        t = rng.choice(np.array([0, 10, 97, 98, 255], dtype=np.uint8), size=n)
        assert np.array_equal(pss.libsais(t), oracle.suffix_array_bruteforce(t))


def test_sa_golden_vectors(pss, vectors):
    for v in vectors["sa"]:
        t = bytes.fromhex(v["text"])
        assert pss.libsais(t).tolist() == v["sa"]


def test_sa_larger_vs_reference_libsais(pss, oracle):
    """16 MiB config-1-like text and a 16 MiB repeat-heavy text against the reference's own
    compiled libsais (oracle/_ref), else the port."""
    ref = oracle.suffix_array_reference if oracle.reference_libsais_available() else oracle.suffix_array_port
    for t in (synth.zipf_words_text(1 << 24, seed=5), synth.acgt_text(1 << 24, seed=6)):
        assert np.array_equal(pss.libsais(t), ref(t))


def test_libsais_shim_on_gpu(pss, oracle):
    """libsais_pss_shim.so: the symbol `libsais` itself (what src/lib.rs:14-22 declares), forwarding
    to pss_libsais."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "pysubstringsearch_b200", "libsais_pss_shim.so")
    if not os.path.exists(path):
        pytest.skip("shim not built (make -C pysubstringsearch_b200/csrc shim)")
    shim = C.CDLL(path)
    shim.libsais.restype = C.c_int32
    shim.libsais.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    t = synth.zipf_words_text(200_000, seed=77, vocab=1024, block=1 << 14)
    sa = np.empty(len(t), dtype=np.int32)
    assert shim.libsais(t.ctypes.data, sa.ctypes.data, len(t), 0, None) == 0
    assert np.array_equal(sa, oracle.suffix_array_port(t))


def test_libsais_freq_table(pss):
    t = np.frombuffer(b"abracadabra\n", dtype=np.uint8)
    sa = np.empty(len(t), dtype=np.int32)
    freq = np.zeros(256, dtype=np.int32)
    assert pss.lib.pss_libsais(t.ctypes.data, sa.ctypes.data, len(t), 0, freq.ctypes.data) == 0
    assert freq[ord("a")] == 5 and freq[10] == 1 and freq.sum() == len(t)


# ------------------------------------------------------------------------------------
# Writer: index file byte-identical to the reference-equivalent writer
# ------------------------------------------------------------------------------------
def _write(cls, path, entries, mcl):
    w = cls(path, mcl)
    for e in entries:
        w.add_entry(e)
    w.finalize()
    w.close()


def test_writer_golden_containers(pss, vectors):
    with tempfile.TemporaryDirectory() as d:
        for case in vectors["containers"]:
            p = os.path.join(d, "g.idx")
            _write(pss.Writer, p, case["entries"], case["max_chunk_len"])
            assert open(p, "rb").read().hex() == case["container_hex"]


def test_writer_file_lines_golden(pss, vectors):
    with tempfile.TemporaryDirectory() as d:
        for case in vectors["file_lines"]:
            src, p = os.path.join(d, "in.txt"), os.path.join(d, "g.idx")
            open(src, "wb").write(bytes.fromhex(case["raw_hex"]))
            w = pss.Writer(p)
            assert w.add_entries_from_file_lines(src) == 0
            assert w.finalize() == 0
            w.close()
            assert open(p, "rb").read().hex() == case["container_hex"]


def test_writer_multichunk_vs_oracle(pss, oracle):
    text = synth.zipf_words_text(3_000_000, seed=21, vocab=4096, block=1 << 16)
    entries = bytes(text).split(b"\n")[:-1]
    with tempfile.TemporaryDirectory() as d:
        for mcl in (None, 1 << 20, 300_000, 4096):
            a, b = os.path.join(d, "gpu.idx"), os.path.join(d, "cpu.idx")
            _write(pss.Writer, a, entries, mcl)
            _write(oracle.Writer, b, entries, mcl)
            assert open(a, "rb").read() == open(b, "rb").read(), mcl
        # file ingestion, same splitting rule, CRLF mixed in
        src = os.path.join(d, "in.txt")
        open(src, "wb").write(b"\r\n".join(entries[:5000]) + b"\n" + b"\n".join(entries[5000:20000]))
        a, b = os.path.join(d, "gpu2.idx"), os.path.join(d, "cpu2.idx")
        w = pss.Writer(a, 1 << 18)
        assert w.add_entries_from_file_lines(src) == 0
        w.close()
        w = oracle.Writer(b, 1 << 18)
        w.add_entries_from_file_lines(src)
        w.close()
        assert open(a, "rb").read() == open(b, "rb").read()
        # LF-only files take the bulk-append path (whole runs of lines per copy): chunk boundaries
        # must still fall where the per-line flush rule puts them, for chunk sizes far below and
        # above the 1 MiB read block, with and without a final newline
        for tail, mcl in ((b"\n", 4096), (b"", 300_000), (b"\n", 1_500_000), (b"", None)):
            open(src, "wb").write(b"\n".join(entries[:30000]) + tail)
            w = pss.Writer(a, mcl)
            assert w.add_entries_from_file_lines(src) == 0
            w.close()
            w = oracle.Writer(b, mcl)
            w.add_entries_from_file_lines(src)
            w.close()
            assert open(a, "rb").read() == open(b, "rb").read(), (tail, mcl)


def test_writer_overlong_file_line_grows_capacity(pss, oracle):
    """add_entries_from_file_lines has no "too big" check (lib.rs:73-79): a line longer than
    max_chunk_len grows the buffer for good, which moves every later chunk boundary."""
    rng = np.random.default_rng(5)
    lines = [bytes(rng.integers(97, 123, size=int(k), dtype=np.uint8)) for k in
             [10, 20, 3000, 15, 15, 900, 40, 40, 40, 7000, 5, 5, 5, 5, 5, 5, 5, 5]]
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "in.txt")
        open(src, "wb").write(b"\n".join(lines) + b"\n")
        for mcl in (64, 1000, 8):
            a, b = os.path.join(d, "a.idx"), os.path.join(d, "b.idx")
            w = pss.Writer(a, mcl)
            assert w.add_entries_from_file_lines(src) == 0
            assert w.add_entry("tail") in (0, -6)
            w.close()
            ow = oracle.Writer(b, mcl)
            ow.add_entries_from_file_lines(src)
            try:
                ow.add_entry("tail")
            except ValueError:
                pass
            ow.close()
            assert open(a, "rb").read() == open(b, "rb").read(), mcl


def test_writer_capacity_quirks(pss, oracle):
    """An entry that exactly fills the chunk doubles the capacity for good (Vec growth)."""
    with tempfile.TemporaryDirectory() as d:
        a, b = os.path.join(d, "a.idx"), os.path.join(d, "b.idx")
        entries = ["1234", "x", "yy", "zzzz", "12345678", "q"]
        for cls, p in ((pss.Writer, a), (oracle.Writer, b)):
            w = cls(p, 4)
            for e in entries:
                try:
                    rc = w.add_entry(e)
                    assert rc in (0, None)
                except ValueError:
                    pass
            w.close()
        assert open(a, "rb").read() == open(b, "rb").read()


# ------------------------------------------------------------------------------------
# Reader: ordered result tuples identical to the oracle
# ------------------------------------------------------------------------------------
def _compare_searches(pss_reader, oracle_reader, patterns):
    qo, ch, st, en, stats = pss_reader.search_batch(patterns)
    counts, och, ost, oen = oracle_reader.search_multiple_tuples(patterns)
    assert np.array_equal(np.diff(qo), counts)
    assert np.array_equal(ch, och)
    assert np.array_equal(st, ost)
    assert np.array_equal(en, oen)
    return stats


def test_search_golden_containers(pss, vectors):
    with tempfile.TemporaryDirectory() as d:
        for case in vectors["containers"]:
            p = os.path.join(d, "g.idx")
            open(p, "wb").write(bytes.fromhex(case["container_hex"]))
            r = pss.Reader(p)
            for s in case["searches"]:
                qo, ch, st, en, _ = r.search_batch([s["pattern"]])
                assert (ch.tolist(), st.tolist(), en.tolist()) == (s["chunk"], s["start"], s["end"]), s["pattern"]
            r.close()


def test_search_random_patterns_vs_oracle(pss, oracle):
    text = synth.zipf_words_text(4_000_000, seed=31, vocab=4096, block=1 << 16)
    synth.plant(text, "google", 300, 1)
    entries = bytes(text).split(b"\n")[:-1]
    with tempfile.TemporaryDirectory() as d:
        for mcl in (None, 1 << 20):
            p = os.path.join(d, "s.idx")
            _write(oracle.Writer, p, entries, mcl)      # reference-format file, read by the GPU reader
            r, o = pss.Reader(p), oracle.Reader(p)
            assert r.num_chunks == o.num_chunks
            pats = synth.config2_queries(text, nq=600, seed=3)
            pats += [b"", b"\n", b"e ", b"google", b"zzzzzz", b"a", b" ", b"\n\n", b"sojqmxwtxw",
                     bytes(text[1000:1100]), bytes(text[5000:5070]), b"q" * 300]
            stats = _compare_searches(r, o, pats)
            assert stats["n_hits"] > len(text)          # the empty pattern alone matches every suffix
            # one at a time == batched
            for pat in pats[::50]:
                _compare_searches(r, o, [pat])
            r.close()
            o.close()


@pytest.mark.parametrize("env", [{"PSS_BOUNDS_GROUP": "-4"}, {"PSS_BOUNDS_GROUP": "4"}, {"PSS_BOUNDS_GROUP": "-8"},
                                 {"PSS_BOUNDS_GROUP": "8"}, {"PSS_BOUNDS_GROUP": "-32"}, {"PSS_BOUNDS_GROUP": "32"},
                                 {"PSS_LINE_DIR": "0"}, {"PSS_LINE_DIR": "1"}])
def test_search_kernel_variants_vs_oracle(pss, oracle, env, monkeypatch):
    """The batched bounds kernel picks its geometry by batch size (a warp per pair below 16 384
    pairs, 8 or 4 lanes per pair above) and extraction reads the line records (PSS_LINE_DIR: 1 = the
    line directory, 0 = text scans); every
    variant is forced here on the same patterns — empty, '\\n', longer than any window, absent,
    crossing entries — and compared with the oracle's ordered tuples."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)               # read when the Reader is opened
    monkeypatch.setenv("PSS_SMALL_PATH", "0")  # one-query batches through the batched kernels as well
    text = synth.zipf_words_text(3_000_000, seed=32, vocab=2048, block=1 << 16)
    entries = bytes(text).split(b"\n")[:-1]
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "v.idx")
        _write(oracle.Writer, p, entries, 1 << 20)       # 3 chunks
        r, o = pss.Reader(p), oracle.Reader(p)
        pats = synth.config2_queries(text, nq=400, seed=5)
        pats += [b"", b"\n", b"e ", b"zzzzzz", b"a", b" ", b"\n\n", bytes(text[1000:1100]), bytes(text[5000:5033]),
                 bytes(text[7000:7004]), bytes(text[7000:7005]), bytes(text[len(text) - 9:]), b"q" * 300]
        _compare_searches(r, o, pats)
        for pat in pats[::40]:
            _compare_searches(r, o, [pat])
        r.close()
        o.close()


def test_search_binary_text_and_long_lines(pss, oracle):
    rng = np.random.default_rng(8)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "b.idx")
        w = oracle.Writer(p, 1 << 16)
        # few newlines → very long entries; 2-symbol alphabet → huge hit ranges
        for _ in range(40):
            w.add_entry(bytes(rng.choice(np.frombuffer(b"ab", dtype=np.uint8), size=int(rng.integers(1, 9000)))))
        w.finalize()
        w.close()
        r, o = pss.Reader(p), oracle.Reader(p)
        pats = [b"a", b"b", b"ab", b"ba", b"aaaa", b"abababab", b"", b"b\na", b"a" * 40, b"c"]
        _compare_searches(r, o, pats)
        r.close()
        o.close()


def test_single_query_paths_small_and_overflow(pss, oracle):
    """A one-query batch takes the fused single-launch path up to 8192 matching suffixes and
    falls back to the general pipeline above that; both must agree with the oracle."""
    text = synth.zipf_words_text(2_000_000, seed=61, vocab=1024, block=1 << 16)
    entries = bytes(text).split(b"\n")[:-1]
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "s.idx")
        _write(oracle.Writer, p, entries, 1 << 20)          # 2 chunks → 2 pairs per query
        r, o = pss.Reader(p), oracle.Reader(p)
        for pat in [b"zzzzzzzz", bytes(text[777:800]), bytes(text[5000:5006]), b"e", b" ", b"", b"\n", b"ab"]:
            stats = _compare_searches(r, o, [pat])
        # a few queries at once still fit the small path (<= 64 pairs)
        _compare_searches(r, o, [bytes(text[k:k + 9]) for k in range(100, 3000, 100)])
        r.close()
        o.close()


def test_empty_index_and_no_hits(pss, oracle):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "e.idx")
        w = pss.Writer(p)
        assert w.finalize() == 0
        w.close()
        assert os.path.getsize(p) == 0
        r = pss.Reader(p)
        qo, ch, st, en, _ = r.search_batch(["a", ""])
        assert qo.tolist() == [0, 0, 0] and len(ch) == 0
        qo, ch, st, en, _ = r.search_batch([])
        assert qo.tolist() == [0]
        r.close()


def test_sharded_readers_cover_the_index(pss, oracle):
    text = synth.zipf_words_text(1_500_000, seed=41, vocab=2048, block=1 << 16)
    entries = bytes(text).split(b"\n")[:-1]
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "m.idx")
        _write(pss.Writer, p, entries, 1 << 18)
        full = pss.Reader(p)
        pats = synth.config2_queries(text, nq=100, seed=5) + [b"", b"e "]
        qo, ch, st, en, _ = full.search_batch(pats)
        parts = [pss.Reader(p, shard=(k, 3)) for k in range(3)]
        rows = []
        for k, part in enumerate(parts):
            pqo, pch, pst, pen, _ = part.search_batch(pats)
            assert all(c % 3 == k for c in pch)
            q = np.repeat(np.arange(len(pats)), np.diff(pqo))
            rows.append(np.stack([q, pch, np.arange(len(pch)), pst, pen], axis=1))
        merged = np.concatenate(rows)
        # (query, chunk) order; inside a pair the shard's own order is kept (stable)
        merged = merged[np.lexsort((merged[:, 2], merged[:, 1], merged[:, 0]))]
        assert np.array_equal(merged[:, 1], ch) and np.array_equal(merged[:, 3], st) and np.array_equal(merged[:, 4], en)
        for part in parts:
            part.close()
        full.close()


def test_multi_device_reader_front(pss, oracle, monkeypatch):
    """PSS_DEVICES spreads the chunks of ONE Reader over several GPUs inside the process
    (chunk k -> listed device k % G) and merges the per-device results back into the
    single-process order.  Listing device 0 three times exercises the whole front (threads,
    merge, text lookup) on a one-GPU box; with more GPUs visible "all" is tried as well."""
    import pysubstringsearch
    text = synth.zipf_words_text(1_500_000, seed=71, vocab=2048, block=1 << 16)
    entries = bytes(text).split(b"\n")[:-1]
    pats = synth.config2_queries(text, nq=150, seed=8) + [b"", b"e ", b"zzzz", b"\n"]
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "m.idx")
        _write(oracle.Writer, p, entries, 1 << 18)          # 6 chunks
        o = oracle.Reader(p)
        specs = ["0,0,0", "0,0"] + (["all"] if pss.lib.pss_device_count() >= 2 else [])
        for spec in specs:
            monkeypatch.setenv("PSS_DEVICES", spec)
            r = pss.Reader(p)
            assert r.num_chunks == o.num_chunks
            _compare_searches(r, o, pats)
            _compare_searches(r, o, [pats[3]])
            r.close()
            pr = pysubstringsearch.Reader(index_file_path=p)
            assert pr.search_multiple(substrings=[q.decode() for q in pats[:40]]) == o.search_multiple(pats[:40])
            del pr
        monkeypatch.delenv("PSS_DEVICES")
        o.close()


def _device_result_arrays(pss, dr):
    return (pss.from_device(dr.d_query_offsets, dr.n_queries + 1, np.int64),
            pss.from_device(dr.d_chunk_id, dr.n_entries, np.int32),
            pss.from_device(dr.d_line_start, dr.n_entries, np.uint32),
            pss.from_device(dr.d_line_end, dr.n_entries, np.uint32))


def test_device_resident_search_matches_host_api(pss):
    import torch
    text = synth.zipf_words_text(1_000_000, seed=51, vocab=2048, block=1 << 16)
    entries = bytes(text).split(b"\n")[:-1]
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "m.idx")
        _write(pss.Writer, p, entries, 1 << 19)
        r = pss.Reader(p)
        pats = synth.config2_queries(text, nq=200, seed=6)
        qo, ch, st, en, _ = r.search_batch(pats)
        blob, offs = synth.pack_patterns(pats)
        d_blob = torch.from_numpy(blob).cuda()
        d_offs = torch.from_numpy(offs).cuda()
        torch.cuda.synchronize()
        dr = pss.DeviceResult()
        rc = pss.lib.pss_reader_search_batch_device(r.h, d_blob.data_ptr(), d_offs.data_ptr(), len(pats), int(offs[-1]),
                                                    C.byref(dr), None)
        assert rc == 0, pss.err()
        assert dr.n_entries == len(ch) and dr.n_queries == len(pats) and dr.n_chunks == r.num_chunks
        gqo, gch, gst, gen = _device_result_arrays(pss, dr)
        assert np.array_equal(gqo, qo) and np.array_equal(gch, ch) and np.array_equal(gst, st) and np.array_equal(gen, en)
        # per-(query, chunk) entry offsets: the exclusive scan of the per-pair entry counts
        npairs = len(pats) * r.num_chunks
        eo = pss.from_device(dr.d_entry_offsets, npairs + 1, np.uint32).astype(np.int64)
        pair = np.repeat(np.arange(len(pats)), np.diff(qo)) * r.num_chunks + ch
        want = np.concatenate(([0], np.cumsum(np.bincount(pair, minlength=npairs))))
        assert np.array_equal(eo, want)
        r.close()


def test_bad_pattern_offsets_are_rejected(pss):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "o.idx")
        _write(pss.Writer, p, [b"alpha", b"beta"], None)
        r = pss.Reader(p)
        blob = np.frombuffer(b"alphabeta", dtype=np.uint8).copy()
        res = C.c_void_p()
        for bad in ([1, 5, 9], [0, 6, 5], [0, -1, 4]):
            offs = np.array(bad, dtype=np.int64)
            assert pss.lib.pss_reader_search_batch(r.h, blob.ctypes.data, offs.ctypes.data, 2, C.byref(res)) == -1
        r.close()


def test_async_build_seam_pipelined(pss, oracle):
    """pss_sa_build_begin/_wait: several builds queued at once (two device slots), pageable and
    pinned destinations, every suffix array identical to the oracle's."""
    import torch
    texts = [synth.zipf_words_text(n, seed=100 + i, vocab=4096, block=1 << 16)
             for i, n in enumerate([3_000_000, 1, 10_000_000, 70_000, 9_000_000, 2])]
    handles = []
    for t in texts[:2]:
        h = C.c_void_p()
        pss.check(pss.lib.pss_sa_build_begin(-1, t.ctypes.data, len(t), C.byref(h)))
        handles.append(h)
    out = []
    for k, t in enumerate(texts):
        if k + 2 < len(texts):                      # keep two builds in flight ahead of the one waited for
            h = C.c_void_p()
            pss.check(pss.lib.pss_sa_build_begin(-1, texts[k + 2].ctypes.data, len(texts[k + 2]), C.byref(h)))
            handles.append(h)
        if k % 2 == 0:
            sa = np.empty(len(t), dtype=np.int32)
            pss.check(pss.lib.pss_sa_build_wait(handles[k], sa.ctypes.data))
        else:
            pinned = torch.empty(len(t), dtype=torch.int32).pin_memory()
            pss.check(pss.lib.pss_sa_build_wait(handles[k], pinned.data_ptr()))
            sa = pinned.numpy().copy()
        out.append(sa)
    for t, sa in zip(texts, out):
        assert np.array_equal(sa, oracle.suffix_array_port(t))
    assert pss.lib.pss_sa_build_wait(None, None) == -1
    assert pss.lib.pss_release_cached() == 0
    assert np.array_equal(pss.libsais(texts[3]), out[3])          # the workspace regrows after a release


def test_writer_explicit_device_list(pss, oracle):
    """Chunk k → devices[k % G]: builds run on per-device engines, records reach the file in chunk order."""
    text = synth.zipf_words_text(6_000_000, seed=23, vocab=4096, block=1 << 16)
    entries = bytes(text).split(b"\n")[:-1]
    ndev = pss.lib.pss_device_count()
    lists = [[0], [0, 0, 0]] + ([list(range(ndev))] if ndev >= 2 else [])
    with tempfile.TemporaryDirectory() as d:
        b = os.path.join(d, "cpu.idx")
        _write(oracle.Writer, b, entries, 1 << 20)
        want = open(b, "rb").read()
        for devs in lists:
            a = os.path.join(d, "gpu.idx")
            w = pss.Writer(a, 1 << 20, devices=devs)
            for e in entries:
                assert w.add_entry(e) == 0
            assert w.finalize() == 0
            w.close()
            assert open(a, "rb").read() == want, devs
            r = pss.Reader(a, devices=devs)
            o = oracle.Reader(b)
            _compare_searches(r, o, [b"e ", b"", bytes(text[500:520])])
            r.close()
            o.close()


def test_newline_side_index_long_lines(pss, oracle):
    """Entries far longer than the scan window: extraction goes through the newline side index
    (one binary search), results identical to the oracle; a 64 MiB single-line chunk stays
    bounded (the unindexed scan would walk up to 64 MiB per hit)."""
    rng = np.random.default_rng(12)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "l.idx")
        w = oracle.Writer(p, 1 << 22)
        lens = [0, 1, 63, 64, 65, 127, 128, 129, 200, 5000, 70000, 3, 0, 0, 900_000, 64, 1]
        for ln in lens * 2:
            w.add_entry(bytes(rng.choice(np.frombuffer(b"abc ", dtype=np.uint8), size=ln)))
        w.finalize()
        w.close()
        r, o = pss.Reader(p), oracle.Reader(p)
        _compare_searches(r, o, [b"abc", b"a", b"", b"\n", b"cab a", b" ", b"\n\n", b"bbbbbbbb", b"c\na"])
        for pat in (b"abca", b"zz", b"\n"):
            _compare_searches(r, o, [pat])
        r.close()
        o.close()
    # one 64 MiB line (2-symbol text: every 2-gram has ~16 M matching suffixes, all in the same entry)
    n = 64 << 20
    text = rng.integers(97, 99, size=n, dtype=np.uint8)
    text[-1] = 10
    sa = pss.libsais(text)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "one.idx")
        with open(p, "wb") as f:
            f.write(np.uint32(n).tobytes()); f.write(memoryview(text))
            f.write(np.uint32(4 * n).tobytes()); f.write(memoryview(sa))
        r = pss.Reader(p)
        qo, ch, st, en, stats = r.search_batch([b"ab", b"ba"])
        assert qo.tolist() == [0, 1, 2] and st.tolist() == [0, 0] and en.tolist() == [n - 1, n - 1]
        assert stats["n_hits"] > n // 3
        assert stats["ms_extract"] < 50.0, stats       # 33 M hits x one bounded lookup each
        qo, ch, st, en, stats = r.search_batch([bytes(text[1000:1040])])
        assert st.tolist() == [0] and en.tolist() == [n - 1]
        r.close()


def test_reader_over_device_resident_chunks(pss, oracle):
    """pss_reader_open_device_chunks: chunks built and kept in HBM (no file), searched through the
    same pipeline; with world = 1 the distributed call is the plain search."""
    import torch
    texts = [synth.zipf_words_text(1_500_000 + 1000 * k, seed=300 + k, vocab=2048, block=1 << 16) for k in range(3)]
    d_text, d_sa, chunks = [], [], []
    builder = C.c_void_p()
    pss.check(pss.lib.pss_sa_builder_create(-1, 0, C.byref(builder)))
    for k, t in enumerate(texts):
        dt = torch.zeros(len(t) + 16, dtype=torch.uint8, device="cuda")
        dt[:len(t)] = torch.from_numpy(t).cuda()
        ds = torch.empty(len(t), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        pss.check(pss.lib.pss_sa_builder_build_device(builder, dt.data_ptr(), len(t), ds.data_ptr(), None))
        d_text.append(dt); d_sa.append(ds)
        chunks.append(pss.DeviceChunk(dt.data_ptr(), ds.data_ptr(), t.ctypes.data, len(t), k))
    pss.lib.pss_sa_builder_destroy(builder)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "ref.idx")
        with open(p, "wb") as f:
            for t, ds in zip(texts, d_sa):
                f.write(np.uint32(len(t)).tobytes()); f.write(memoryview(t))
                f.write(np.uint32(4 * len(t)).tobytes()); f.write(memoryview(ds.cpu().numpy()))
        o = oracle.Reader(p)
        for k, t in enumerate(texts):
            assert np.array_equal(o.chunk_sa(k), oracle.suffix_array_port(t))
        r = pss.Reader(device_chunks=chunks, n_chunks_total=3)
        pats = synth.config2_queries(texts[1], nq=300, seed=9) + [b"", b"e ", b"\n"]
        _compare_searches(r, o, pats)
        comm = pss.Comm(0, 1)
        qo, ch, st, en, stats = r.search_batch_dist(comm, pats)
        counts, och, ost, oen = o.search_multiple_tuples(pats)
        assert np.array_equal(np.diff(qo), counts) and np.array_equal(ch, och)
        assert np.array_equal(st, ost) and np.array_equal(en, oen)
        comm.close()
        r.close()
        o.close()


def test_two_gpu_distributed_search_matches_single_process():
    """One process per GPU (torchrun, 2 ranks): index sharded chunk k → rank k % 2, batch broadcast
    from rank 0, hits gathered with the in-library NCCL gather-v; rank 0's merged result must equal
    the single-process Reader's and the oracle's (tools/dist_check.py asserts it)."""
    import subprocess
    import sys
    from pysubstringsearch_b200 import capi
    if capi.lib.pss_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tests.conftest import ROOT
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "dist_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "dist_check ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


# ------------------------------------------------------------------------------------
# Python API
# ------------------------------------------------------------------------------------
def test_python_api_reference_kats(kats):
    """The reference's own test vectors through the drop-in import name."""
    import pysubstringsearch
    for k in kats:
        if k["method"] == "open_missing":
            with pytest.raises(FileNotFoundError):
                pysubstringsearch.Reader(index_file_path=k["query"])
            continue
        with tempfile.TemporaryDirectory() as d:
            p = f"{d}/output.idx"
            w = pysubstringsearch.Writer(index_file_path=p)
            for e in k["entries"]:
                w.add_entry(text=e)
            w.finalize()
            r = pysubstringsearch.Reader(index_file_path=p)
            got = r.search(substring=k["query"]) if k["method"] == "search" else r.search_multiple(substrings=k["query"])
            assert sorted(got) == sorted(k["expected"]), k


def test_python_api_ordered_vs_oracle(oracle, vectors):
    import pysubstringsearch
    with tempfile.TemporaryDirectory() as d:
        for case in vectors["containers"]:
            p = os.path.join(d, "g.idx")
            w = pysubstringsearch.Writer(index_file_path=p, max_chunk_len=case["max_chunk_len"])
            for e in case["entries"]:
                w.add_entry(text=e)
            w.finalize()
            del w
            r = pysubstringsearch.Reader(index_file_path=p)
            for s in case["searches"]:
                assert r.search(substring=s["pattern"]) == s["strings"]
            assert r.search_multiple(substrings=[s["pattern"] for s in case["searches"]]) == case["search_multiple"]


def test_python_writer_drop_flushes(oracle):
    import pysubstringsearch
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "drop.idx")
        w = pysubstringsearch.Writer(index_file_path=p)
        w.add_entry(text="never finalized")
        del w                                   # Drop → finalize (lib.rs:138-144)
        assert oracle.Reader(p).search("final") == ["never finalized"]


# ------------------------------------------------------------------------------------
# full size: BASELINE configs at their stated sizes
# ------------------------------------------------------------------------------------
def _assert_suffix_array_properties(text, sa):
    """Pins a suffix array completely without an oracle: SA is a permutation of 0..n-1 (every
    index exactly once), and for neighbours a = SA[k-1], b = SA[k]: T[a] < T[b], or T[a] == T[b]
    and rank(a+1) < rank(b+1) (past-the-end = -1)."""
    n = len(text)
    assert sa.dtype == np.int32 and len(sa) == n
    assert int(sa.min()) == 0 and int(sa.max()) == n - 1
    seen = np.zeros(n, dtype=bool)
    seen[sa] = True
    assert seen.all(), "SA is not a permutation"
    del seen
    isa = np.empty(n, dtype=np.int32)
    isa[sa] = np.arange(n, dtype=np.int32)
    step = 50_000_000                       # bounded temporaries
    for lo in range(1, n, step):
        hi = min(n, lo + step)
        a, b = sa[lo - 1:hi - 1], sa[lo:hi]
        ta, tb = text[a], text[b]
        ra = np.where(a + 1 < n, isa[np.minimum(a + 1, n - 1)], -1)
        rb = np.where(b + 1 < n, isa[np.minimum(b + 1, n - 1)], -1)
        assert bool(np.all((ta < tb) | ((ta == tb) & (ra < rb)))), "suffixes out of order in [%d, %d)" % (lo, hi)


def _write_one_chunk(path, text, sa):
    with open(path, "wb") as f:
        f.write(np.uint32(len(text)).tobytes()); f.write(memoryview(text))
        f.write(np.uint32(4 * len(text)).tobytes()); f.write(memoryview(sa))


def test_full_size_config1_2_5(pss, oracle):
    """BASELINE configs[0], [1], [4] on the 500 000 000-byte chunk.  The GPU-built suffix array is
    pinned by its properties; then, on the GPU-written index file, the ORDERED result tuples of
    the reference restatement (oracle.Reader, which probes the file as lib.rs does) must equal
    the GPU reader's for: the config-1 queries, the full 10 000-query config-2 batch, and the
    config-5 query (the most frequent bigram of the text: millions of matching suffixes)."""
    n = 500_000_000
    text = synth.config1_text(n)
    sa = pss.libsais(text)
    _assert_suffix_array_properties(text, sa)
    pats = synth.config2_queries(text, nq=10_000, seed=7)
    pairs = text[:-1].astype(np.uint16) << 8 | text[1:]
    top = int(np.bincount(pairs, minlength=1 << 16).argmax())
    bigram = bytes([top >> 8, top & 255])
    del pairs
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        p = os.path.join(d, "full.idx")
        _write_one_chunk(p, text, sa)
        del sa
        r, o = pss.Reader(p), oracle.Reader(p)
        for pat in (b"google", b"text_two", b"qqqqqqqqqq"):          # config 1
            _compare_searches(r, o, [pat])
        qo, ch, st, en, _ = r.search_batch([b"google", b"text_two"])
        assert np.diff(qo).tolist() == [5943, 159]
        stats = _compare_searches(r, o, pats)                          # config 2: ordered, full batch
        assert stats["n_hits"] > 100_000
        stats = _compare_searches(r, o, [bigram])                      # config 5
        assert stats["n_hits"] > 5_000_000
        r.close()
        o.close()


def test_full_size_config3_chunk_and_writer(pss, oracle):
    """BASELINE configs[2]: (a) one full 2^29-byte chunk of the config-3 corpus: the GPU suffix
    array equals the reference's own compiled libsais byte for byte; (b) a 3-chunk index written
    by the GPU Writer from a text file is byte-identical to the oracle Writer's (with the
    reference libsais), and the sharded readers over it agree with the oracle."""
    ref = oracle.suffix_array_reference if oracle.reference_libsais_available() else oracle.suffix_array_port
    text = synth.config3_chunk(1, synth.CONFIG3_CHUNK_BYTES, device="cuda")
    assert len(text) == 1 << 29 and text[-1] == 10
    sa = pss.libsais(text)
    assert np.array_equal(sa, ref(text)), "2^29-byte chunk: suffix array differs from libsais"
    del sa
    m = 48 << 20
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        src, a, b = os.path.join(d, "in.txt"), os.path.join(d, "gpu.idx"), os.path.join(d, "cpu.idx")
        with open(src, "wb") as f:
            for k in range(3):
                f.write(memoryview(synth.config3_chunk(k, m, device="cuda")))
        w = pss.Writer(a, m)
        assert w.add_entries_from_file_lines(src) == 0
        assert w.close() == 0
        oracle.use_reference_libsais(oracle.reference_libsais_available())
        try:
            ow = oracle.Writer(b, m)
            ow.add_entries_from_file_lines(src)
            ow.close()
        finally:
            oracle.use_reference_libsais(False)
        assert os.path.getsize(a) == 3 * (8 + 5 * m)
        assert subprocess_cmp(a, b), "3-chunk index differs from the oracle Writer's"
        r, o = pss.Reader(a), oracle.Reader(b)
        assert r.num_chunks == 3
        _compare_searches(r, o, [b"google", b"text_two", b"sojq", b"e "] + [bytes(text[k:k + 11]) for k in range(7, 4000, 97)])
        r.close()
        o.close()


def subprocess_cmp(a, b):
    import subprocess
    return subprocess.run(["cmp", "-s", a, b]).returncode == 0


def test_full_size_config4_low_entropy(pss):
    """BASELINE configs[3]: a 2^29-byte ACGT chunk with 64 KiB repeats (the worst case for the
    doubling depth): permutation + neighbour-order properties of the GPU suffix array."""
    n = 1 << 29
    text = synth.acgt_text(n)
    sa = pss.libsais(text)
    _assert_suffix_array_properties(text, sa)


