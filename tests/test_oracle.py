"""CPU tests: pin the oracle (oracle/) against golden vectors, brute force and — where the
prebuilt oracle/_ref exists — the reference's own compiled libsais.c."""
import os
import tempfile

import numpy as np
import pytest


def test_port_matches_bruteforce(oracle):
    rng = np.random.default_rng(1)
    for _ in range(200):
        n = int(rng.integers(1, 120))
        t = rng.choice(np.array([0, 10, 97, 98, 255], dtype=np.uint8), size=n)
        assert np.array_equal(oracle.suffix_array_port(t), oracle.suffix_array_bruteforce(t))


def test_port_matches_golden_reference_sa(oracle, vectors):
    """SA vectors computed by the reference's libsais.c (committed fixture)."""
    for v in vectors["sa"]:
        t = bytes.fromhex(v["text"])
        assert oracle.suffix_array_port(t).tolist() == v["sa"]


def test_port_matches_compiled_reference(oracle):
    if not oracle.reference_libsais_available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(2)
    for kind in range(4):
        n = 200_000
        if kind == 0:
            t = rng.integers(0, 256, size=n, dtype=np.uint8)
        elif kind == 1:
            t = rng.integers(97, 101, size=n, dtype=np.uint8)
        elif kind == 2:
            t = np.tile(np.frombuffer(b"ACGT", dtype=np.uint8), n // 4)
        else:
            t = np.full(n, 97, dtype=np.uint8)
            t[::997] = 10
        assert np.array_equal(oracle.suffix_array_port(t), oracle.suffix_array_reference(t))


def test_libsais_contract(oracle):
    assert oracle.suffix_array_port(b"").tolist() == []
    assert oracle.suffix_array_port(b"z").tolist() == [0]
    # proper prefix first, no sentinel byte: "x" < "x\0"
    assert oracle.suffix_array_port(b"x\0x").tolist() == [1, 2, 0]


def test_known_answer_container(oracle):
    """entries ['ab'] → 03000000 61620a 0c000000 02000000 00000000 01000000."""
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "k.idx")
        w = oracle.Writer(p)
        w.add_entry("ab")
        w.finalize()
        w.close()
        assert open(p, "rb").read().hex() == "0300000061620a0c000000020000000000000001000000"


def test_oracle_reproduces_golden_containers(oracle, vectors):
    """The port-backed oracle writer/reader reproduces fixtures made with the reference libsais."""
    oracle.use_reference_libsais(False)
    for case in vectors["containers"]:
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "g.idx")
            w = oracle.Writer(p, case["max_chunk_len"])
            for e in case["entries"]:
                w.add_entry(e)
            w.finalize()
            w.close()
            assert open(p, "rb").read().hex() == case["container_hex"]
            r = oracle.Reader(p)
            for s in case["searches"]:
                ch, st, en = r.search_tuples(s["pattern"])
                assert (ch.tolist(), st.tolist(), en.tolist()) == (s["chunk"], s["start"], s["end"])
                assert r.search(s["pattern"]) == s["strings"]
            assert r.search_multiple([s["pattern"] for s in case["searches"]]) == case["search_multiple"]
            r.close()


def test_oracle_file_lines(oracle, vectors):
    for case in vectors["file_lines"]:
        with tempfile.TemporaryDirectory() as d:
            src = os.path.join(d, "in.txt")
            open(src, "wb").write(bytes.fromhex(case["raw_hex"]))
            p = os.path.join(d, "g.idx")
            w = oracle.Writer(p)
            w.add_entries_from_file_lines(src)
            w.finalize()
            w.close()
            assert open(p, "rb").read().hex() == case["container_hex"]


def test_oracle_passes_reference_kats(oracle, kats):
    """The reference's own test vectors (tests/test_pysubstringsearch.py), order-insensitive
    exactly as the reference asserts them (assertCountEqual)."""
    for k in kats:
        if k["method"] == "open_missing":
            with pytest.raises(FileNotFoundError):
                oracle.Reader(k["query"])
            continue
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "k.idx")
            w = oracle.Writer(p)
            for e in k["entries"]:
                w.add_entry(e)
            w.finalize()
            w.close()
            r = oracle.Reader(p)
            got = r.search(k["query"]) if k["method"] == "search" else r.search_multiple(k["query"])
            assert sorted(got) == sorted(k["expected"]), k
            r.close()


def test_oracle_semantics(oracle):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "s.idx")
        w = oracle.Writer(p)
        for e in ["one", "ten", "tenten", "dup", "dup", ""]:
            w.add_entry(e)
        w.finalize()
        w.close()
        r = oracle.Reader(p)
        assert r.search("ten") == ["tenten", "ten"]        # SA order of the first matching suffix
        assert r.search("dup") == ["dup", "dup"]            # dedup key is the entry offset, not its text
        assert sorted(r.search("")) == sorted(["one", "ten", "tenten", "dup", "dup", ""])
        assert r.search("n\nt") == ["ten"]                   # a match may start in one entry and cross '\n'
        assert r.search("zzz") == []
        # capacity quirks (lib.rs:92-98): too-big check, and an entry that exactly fills the chunk
        w = oracle.Writer(p, 4)
        with pytest.raises(ValueError):
            w.add_entry("12345")
        w.add_entry("1234")
        w.add_entry("x")
        w.close()
        r2 = oracle.Reader(p)
        assert [r2.chunk_text(c) for c in range(r2.num_chunks)] == [b"1234\nx\n"]  # capacity doubled to 8
        # empty index
        w = oracle.Writer(p)
        w.finalize()
        w.close()
        assert os.path.getsize(p) == 0
        assert oracle.Reader(p).search("a") == []


def test_port_property_based(oracle):
    """Hypothesis: for arbitrary short byte strings over a tiny alphabet (many ties, NUL and
    0xFF included) the port equals the definition-level suffix sort."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.binary(min_size=0, max_size=64).map(lambda b: bytes(b"\x00\na\xff"[c & 3] for c in b)))
    def check(t):
        assert oracle.suffix_array_port(t).tolist() == oracle.suffix_array_bruteforce(t).tolist()

    check()


def test_oracle_search_equals_naive_scan(oracle):
    """Reader.search of the oracle against a definition-level scan: every entry that contains
    the pattern (as bytes, matches may start anywhere in the entry and run past its end into
    the following text of the chunk), each once."""
    rng = np.random.default_rng(11)
    entries = ["".join(rng.choice(list("ab "), size=int(rng.integers(0, 12)))) for _ in range(300)]
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "n.idx")
        w = oracle.Writer(p, 256)
        for e in entries:
            w.add_entry(e)
        w.close()
        r = oracle.Reader(p)
        chunks = [r.chunk_text(c) for c in range(r.num_chunks)]
        for pat in ["a", "ab", "b a", "", "aa\n", "zz", " "]:
            want = []
            for text in chunks:
                starts = [0] + [i + 1 for i, ch in enumerate(text[:-1]) if ch == 10]
                for s in starts:
                    e = text.index(b"\n", s)
                    # a match belongs to the entry in which it STARTS (positions s..e inclusive)
                    if any(text.startswith(pat.encode(), k) for k in range(s, e + 1)):
                        want.append(text[s:e].decode())
            assert sorted(r.search(pat)) == sorted(want), pat
        r.close()
