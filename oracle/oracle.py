"""oracle.py — ctypes front-end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product (pysubstringsearch_b200) never does.

`Writer` / `Reader` mirror the reference's Python surface
(/root/reference/pysubstringsearch/__init__.py:6-73) on top of oracle/pss_oracle.c, which
restates /root/reference/src/lib.rs.  The suffix array comes from oracle/sais_port.c, or —
`use_reference_libsais(True)` — from the reference's own libsais.c compiled unmodified
into oracle/_ref/libsais_ref.so (built by `make -C oracle ref` where /root/reference
exists; the prebuilt .so travels to the GPU box).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "_build", "libpss_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libsais_ref.so")


def build(quiet=True):
    """Compile the C restatement (and the reference libsais when its source is present)."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-C", _HERE, "port"], stdout=out)
    subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=out)


def _load():
    if not os.path.exists(_PORT_SO):
        build()
    lib = C.CDLL(_PORT_SO)
    vp, sz, i32, i64 = C.c_void_p, C.c_size_t, C.c_int32, C.c_int64
    lib.oracle_libsais.restype = i32
    lib.oracle_libsais.argtypes = [vp, vp, i32, i32, vp]
    lib.oracle_set_sa_function.argtypes = [vp]
    lib.oracle_writer_open.restype = vp
    lib.oracle_writer_open.argtypes = [C.c_char_p, C.c_longlong, C.POINTER(C.c_int)]
    for name in ("dump_data", "finalize", "close"):
        f = getattr(lib, "oracle_writer_" + name)
        f.restype = C.c_int
        f.argtypes = [vp]
    lib.oracle_writer_add_entry.restype = C.c_int
    lib.oracle_writer_add_entry.argtypes = [vp, C.c_char_p, sz]
    lib.oracle_writer_add_entries_from_file_lines.restype = C.c_int
    lib.oracle_writer_add_entries_from_file_lines.argtypes = [vp, C.c_char_p]
    lib.oracle_reader_open.restype = vp
    lib.oracle_reader_open.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    lib.oracle_reader_close.argtypes = [vp]
    lib.oracle_reader_num_chunks.restype = sz
    lib.oracle_reader_num_chunks.argtypes = [vp]
    lib.oracle_reader_chunk_text.restype = vp
    lib.oracle_reader_chunk_text.argtypes = [vp, sz, C.POINTER(sz)]
    lib.oracle_reader_chunk_sa.restype = C.c_int
    lib.oracle_reader_chunk_sa.argtypes = [vp, sz, vp]
    lib.oracle_hits_new.restype = vp
    lib.oracle_hits_free.argtypes = [vp]
    lib.oracle_hits_clear.argtypes = [vp]
    lib.oracle_hits_count.restype = sz
    lib.oracle_hits_count.argtypes = [vp]
    for name in ("chunk", "start", "end"):
        f = getattr(lib, "oracle_hits_" + name)
        f.restype = vp
        f.argtypes = [vp]
    lib.oracle_hits_matches.restype = C.c_uint64
    lib.oracle_hits_matches.argtypes = [vp]
    lib.oracle_hits_probes.restype = C.c_uint64
    lib.oracle_hits_probes.argtypes = [vp]
    lib.oracle_reader_search.restype = C.c_int
    lib.oracle_reader_search.argtypes = [vp, C.c_char_p, sz, vp]
    lib.oracle_reader_search_multiple.restype = C.c_int
    lib.oracle_reader_search_multiple.argtypes = [vp, vp, vp, i32, vp, vp]
    return lib


_lib = _load()
_ref = None


def reference_libsais_available():
    return os.path.exists(_REF_SO)


def _ref_lib():
    global _ref
    if _ref is None:
        _ref = C.CDLL(_REF_SO)
        _ref.libsais.restype = C.c_int32
        _ref.libsais.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    return _ref


def use_reference_libsais(on=True):
    """Make the oracle Writer build its suffix arrays with the reference's compiled libsais."""
    if on:
        _lib.oracle_set_sa_function(C.cast(_ref_lib().libsais, C.c_void_p))
    else:
        _lib.oracle_set_sa_function(None)


def _as_u8(text):
    if isinstance(text, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(text), dtype=np.uint8)
    return np.ascontiguousarray(text, dtype=np.uint8)


def suffix_array_port(text):
    """SA by the C restatement (oracle/sais_port.c)."""
    t = _as_u8(text)
    sa = np.empty(len(t), dtype=np.int32)
    rc = _lib.oracle_libsais(t.ctypes.data, sa.ctypes.data, len(t), 0, None)
    if rc != 0:
        raise RuntimeError("oracle_libsais rc=%d" % rc)
    return sa


def suffix_array_reference(text):
    """SA by the reference's own libsais.c (oracle/_ref)."""
    t = _as_u8(text)
    sa = np.empty(len(t), dtype=np.int32)
    rc = _ref_lib().libsais(t.ctypes.data, sa.ctypes.data, len(t), 0, None)
    if rc != 0:
        raise RuntimeError("libsais rc=%d" % rc)
    return sa


def suffix_array_bruteforce(text):
    """O(n^2 log n) definition-level SA for tiny inputs: sort suffixes as byte strings."""
    b = bytes(_as_u8(text))
    return np.array(sorted(range(len(b)), key=lambda i: b[i:]), dtype=np.int32)


def _raise(rc, what):
    if rc == -5:
        raise FileNotFoundError(what)
    if rc == -6:
        raise ValueError("entry is too big")
    if rc != 0:
        raise OSError("%s failed (rc=%d)" % (what, rc))


class Writer:
    def __init__(self, index_file_path, max_chunk_len=None):
        st = C.c_int(0)
        self._w = _lib.oracle_writer_open(os.fsencode(index_file_path),
                                          -1 if max_chunk_len is None else int(max_chunk_len), C.byref(st))
        if not self._w:
            _raise(st.value, index_file_path)

    def add_entries_from_file_lines(self, input_file_path):
        _raise(_lib.oracle_writer_add_entries_from_file_lines(self._w, os.fsencode(input_file_path)),
               input_file_path)

    def add_entry(self, text):
        b = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        _raise(_lib.oracle_writer_add_entry(self._w, b, len(b)), "add_entry")

    def dump_data(self):
        _raise(_lib.oracle_writer_dump_data(self._w), "dump_data")

    def finalize(self):
        _raise(_lib.oracle_writer_finalize(self._w), "finalize")

    def close(self):
        if self._w:
            w, self._w = self._w, None
            _raise(_lib.oracle_writer_close(w), "close")

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Reader:
    def __init__(self, index_file_path):
        st = C.c_int(0)
        self._r = _lib.oracle_reader_open(os.fsencode(index_file_path), C.byref(st))
        if not self._r:
            _raise(st.value, index_file_path)
        self._h = _lib.oracle_hits_new()
        self._texts = None

    def close(self):
        if getattr(self, "_r", None):
            _lib.oracle_hits_free(self._h)
            _lib.oracle_reader_close(self._r)
            self._r = None

    def __del__(self):
        self.close()

    @property
    def num_chunks(self):
        return _lib.oracle_reader_num_chunks(self._r)

    def chunk_text(self, c):
        n = C.c_size_t(0)
        p = _lib.oracle_reader_chunk_text(self._r, c, C.byref(n))
        return C.string_at(p, n.value)

    def chunk_sa(self, c):
        n = len(self.chunk_text(c))
        sa = np.empty(n, dtype=np.int32)
        _raise(_lib.oracle_reader_chunk_sa(self._r, c, sa.ctypes.data), "chunk_sa")
        return sa

    def _tuples(self):
        n = _lib.oracle_hits_count(self._h)
        if n == 0:
            z = np.zeros(0, dtype=np.uint32)
            return np.zeros(0, dtype=np.int32), z, z
        ch = np.ctypeslib.as_array(C.cast(_lib.oracle_hits_chunk(self._h), C.POINTER(C.c_int32)), (n,)).copy()
        st = np.ctypeslib.as_array(C.cast(_lib.oracle_hits_start(self._h), C.POINTER(C.c_uint32)), (n,)).copy()
        en = np.ctypeslib.as_array(C.cast(_lib.oracle_hits_end(self._h), C.POINTER(C.c_uint32)), (n,)).copy()
        return ch, st, en

    def search_tuples(self, substring):
        """(chunk, line_start, line_end) arrays in the reference's order."""
        b = substring.encode("utf-8") if isinstance(substring, str) else bytes(substring)
        _lib.oracle_hits_clear(self._h)
        _raise(_lib.oracle_reader_search(self._r, b, len(b), self._h), "search")
        return self._tuples()

    def search_multiple_tuples(self, substrings):
        pats = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in substrings]
        offs = np.zeros(len(pats) + 1, dtype=np.int64)
        np.cumsum([len(p) for p in pats], out=offs[1:])
        blob = np.frombuffer(b"".join(pats) + b"\0", dtype=np.uint8)
        counts = np.zeros(len(pats), dtype=np.int64)
        _lib.oracle_hits_clear(self._h)
        _raise(_lib.oracle_reader_search_multiple(self._r, blob.ctypes.data, offs.ctypes.data, len(pats),
                                                  self._h, counts.ctypes.data), "search_multiple")
        ch, st, en = self._tuples()
        return counts, ch, st, en

    def last_stats(self):
        return dict(matches=int(_lib.oracle_hits_matches(self._h)), probes=int(_lib.oracle_hits_probes(self._h)))

    def _strings(self, ch, st, en):
        if self._texts is None:
            self._texts = [self.chunk_text(c) for c in range(self.num_chunks)]
        return [self._texts[c][s:e].decode("utf-8", "surrogateescape") for c, s, e in zip(ch, st, en)]

    def search(self, substring):
        return self._strings(*self.search_tuples(substring))

    def search_multiple(self, substrings):
        _, ch, st, en = self.search_multiple_tuples(substrings)
        return self._strings(ch, st, en)
