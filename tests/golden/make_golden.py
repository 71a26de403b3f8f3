#!/usr/bin/env python
"""Generates the committed golden fixtures in tests/golden/ (run in the authoring container,
where /root/reference exists; the fixtures then travel to the GPU box).

1. reference_kats.json — the known-answer vectors the reference's OWN test file holds
   (/root/reference/tests/test_pysubstringsearch.py:48-294).  They are extracted by
   running that file, unmodified, against a recording `pysubstringsearch` module backed
   by the oracle (oracle/pss_oracle.c + the reference's compiled libsais): every
   Writer.add_entry / Reader.search / search_multiple call and every assertCountEqual
   expectation is captured.  The run itself must pass — that pins the oracle.
2. sa_vectors.json — suffix arrays of small adversarial texts computed by the reference's
   libsais.c (oracle/_ref), whole index containers (hex) written by the oracle Writer with
   that libsais, and ORDERED search results for a set of patterns (incl. empty pattern,
   patterns containing '\\n', duplicates, multi-chunk containers).
"""
import importlib.util
import json
import os
import sys
import tempfile
import types
import unittest

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF_TESTS = "/root/reference/tests/test_pysubstringsearch.py"


def extract_reference_kats():
    O.use_reference_libsais(True)
    records = []
    state = {}

    class RecWriter(O.Writer):
        def __init__(self, index_file_path, max_chunk_len=None):
            super().__init__(index_file_path, max_chunk_len)
            state[index_file_path] = []
            self._path = index_file_path

        def add_entry(self, text):
            state[self._path].append(text)
            super().add_entry(text)

        def finalize(self):
            super().finalize()
            self.close()

    class RecReader(O.Reader):
        def __init__(self, index_file_path):
            super().__init__(index_file_path)
            self._entries = list(state.get(index_file_path, []))

        def search(self, substring):
            res = super().search(substring)
            self._last = dict(entries=self._entries, method="search", query=substring)
            RecReader.last = self._last
            return res

        def search_multiple(self, substrings):
            res = super().search_multiple(substrings)
            RecReader.last = dict(entries=self._entries, method="search_multiple", query=list(substrings))
            return res

    shim = types.ModuleType("pysubstringsearch")
    shim.Writer = RecWriter
    shim.Reader = RecReader
    sys.modules["pysubstringsearch"] = shim
    spec = importlib.util.spec_from_file_location("ref_tests", REF_TESTS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    case_cls = mod.PySubstringSearchTestCase
    orig = unittest.TestCase.assertCountEqual

    def recording_assert(self, first, second, msg=None):
        orig(self, first, second, msg)
        rec = dict(RecReader.last)
        rec["expected"] = list(second)
        rec["test"] = self._testMethodName
        records.append(rec)

    case_cls.assertCountEqual = recording_assert
    suite = unittest.defaultTestLoader.loadTestsFromTestCase(case_cls)
    result = unittest.TextTestRunner(verbosity=0).run(suite)
    del sys.modules["pysubstringsearch"]
    assert result.wasSuccessful(), "the oracle fails the reference's own tests"
    assert result.testsRun == 7
    records.append(dict(test="test_file_not_found", method="open_missing", entries=[], query="missing_index_file_path",
                        expected="FileNotFoundError"))
    return records


def sa_vectors():
    O.use_reference_libsais(True)
    rng = np.random.default_rng(12345)
    texts = [b"a", b"ab", b"ba", b"aa", b"ab\n", b"banana", b"mississippi\n", b"\n\n\n", b"\x00\x00a\x00",
             b"\xff\xfe\xff\xff\n", b"abcabcabcabc\n", b"aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa",
             b"x\nx\0\nx\n", "في البداية\nكان\n".encode()]
    for n in (7, 33, 64, 65, 200, 513):
        texts.append(bytes(rng.choice(np.array([0, 10, 97, 98, 255], dtype=np.uint8), size=n)))
        texts.append(bytes(rng.integers(0, 256, size=n, dtype=np.uint8)))
        texts.append(bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)))
    sa = [dict(text=t.hex(), sa=O.suffix_array_reference(t).tolist()) for t in texts]

    containers = []
    words = ["one", "two", "three", "four", "five", "six", "seven", "eight", "nine", "ten", "tenten"]
    cases = [
        (["ab"], None, ["a", "b", "ab", "", "c"]),
        (words, None, ["ten", "f", "our", "aaa", "onet", "e", "", "n\nt", "\n", "ee"]),
        (words, 16, ["ten", "e", "", "t", "seven", "x"]),
        (["dup", "dup", "", "dup", "x\ty", "", "dupdup"], None, ["dup", "", "up", "pd", "\n\n"]),
        (["alpha", "beta", "gamma", "delta", "epsilon", "dup", "dup", ""], 16, ["a", "dup", "", "lta"]),
        (["héllo wörld", "wörld", "日本語のテキスト", "テキスト", "في البداية"], None, ["ö", "テキスト", "wörld", "ي", "語"]),
        (["a" * 40 + "b", "a" * 39 + "c" + "a" * 40, "a" * 100], None, ["a" * 33, "a" * 40, "a" * 41, "a" * 39 + "c", "a"]),
    ]
    for entries, mcl, patterns in cases:
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "g.idx")
            w = O.Writer(p, mcl)
            for e in entries:
                w.add_entry(e)
            w.finalize()
            w.close()
            blob = open(p, "rb").read()
            r = O.Reader(p)
            searches = []
            for pat in patterns:
                ch, st, en = r.search_tuples(pat)
                searches.append(dict(pattern=pat, chunk=ch.tolist(), start=st.tolist(), end=en.tolist(),
                                     strings=r.search(pat)))
            multi = r.search_multiple(patterns)
            r.close()
            containers.append(dict(entries=entries, max_chunk_len=mcl, container_hex=blob.hex(), searches=searches,
                                   search_multiple=multi))
    # file-lines ingestion (bstr for_byte_line semantics)
    lines_cases = []
    for raw in [b"one\ntwo\r\nthree", b"\n\nx\n", b"", b"no newline", b"crlf\r\n\r\nend\r", b"a\rb\n"]:
        with tempfile.TemporaryDirectory() as d:
            src = os.path.join(d, "in.txt")
            open(src, "wb").write(raw)
            p = os.path.join(d, "g.idx")
            w = O.Writer(p, None)
            w.add_entries_from_file_lines(src)
            w.finalize()
            w.close()
            lines_cases.append(dict(raw_hex=raw.hex(), container_hex=open(p, "rb").read().hex()))
    return dict(sa=sa, containers=containers, file_lines=lines_cases)


def main():
    assert os.path.exists(REF_TESTS), "run this where /root/reference exists"
    assert O.reference_libsais_available(), "run `make -C oracle ref` first"
    kats = extract_reference_kats()
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump(kats, f, ensure_ascii=False, indent=1)
    vec = sa_vectors()
    with open(os.path.join(HERE, "sa_vectors.json"), "w") as f:
        json.dump(vec, f, ensure_ascii=False, indent=0)
    print("wrote %d reference KATs, %d SA vectors, %d containers, %d file-line cases"
          % (len(kats), len(vec["sa"]), len(vec["containers"]), len(vec["file_lines"])))


if __name__ == "__main__":
    main()
