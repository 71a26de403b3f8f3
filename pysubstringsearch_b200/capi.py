"""ctypes binding of include/pss.h (libpss_b200.so): the plain C ABI, for callers that want
result tuples / device-resident entry points instead of the Writer/Reader classes.  Used by
bench.py and the tests; mirrored by the stubs in INTEGRATION.md.  No compute happens here."""
import ctypes as C
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpss_b200.so")
lib = C.CDLL(LIB_PATH)

vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
lib.pss_last_error.restype = C.c_char_p
lib.pss_version.restype = C.c_char_p
lib.pss_kernel_launch_count.restype = i64
lib.pss_libsais.restype = i32
lib.pss_libsais.argtypes = [vp, vp, i32, i32, vp]
lib.pss_sa_builder_create.argtypes = [i32, i64, C.POINTER(vp)]
lib.pss_sa_builder_destroy.argtypes = [vp]
lib.pss_sa_builder_destroy.restype = None
lib.pss_sa_builder_set_profiling.argtypes = [vp, i32]
lib.pss_sa_builder_build_device.argtypes = [vp, vp, i32, vp, vp]
lib.pss_sa_builder_build_host.argtypes = [vp, vp, i32, vp]
lib.pss_sa_builder_stats.argtypes = [vp, vp, vp]
lib.pss_radix_sort_pairs.argtypes = [vp, vp, vp, vp, i64, i32, i32, C.POINTER(i32), vp, C.POINTER(i32), vp]
lib.pss_writer_open.argtypes = [C.c_char_p, i64, C.POINTER(vp)]
lib.pss_writer_add_entry.argtypes = [vp, C.c_char_p, sz]
lib.pss_writer_add_entries_from_file_lines.argtypes = [vp, C.c_char_p]
lib.pss_writer_dump_data.argtypes = [vp]
lib.pss_writer_finalize.argtypes = [vp]
lib.pss_writer_close.argtypes = [vp]
lib.pss_reader_open.argtypes = [C.c_char_p, C.POINTER(vp)]
lib.pss_reader_open_sharded.argtypes = [C.c_char_p, i32, i32, C.POINTER(vp)]
lib.pss_reader_close.argtypes = [vp]
lib.pss_reader_num_chunks.argtypes = [vp]
lib.pss_reader_num_local_chunks.argtypes = [vp]
lib.pss_reader_chunk_text.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(i64)]
lib.pss_reader_search_batch.argtypes = [vp, vp, vp, i32, C.POINTER(vp)]
lib.pss_reader_search_batch_device.argtypes = [vp, vp, vp, i32, i64, vp, vp]
lib.pss_sa_build_begin.argtypes = [i32, vp, i32, C.POINTER(vp)]
lib.pss_sa_build_wait.argtypes = [vp, vp]
lib.pss_release_cached.argtypes = []
lib.pss_memcpy_d2h.argtypes = [vp, vp, sz]
lib.pss_writer_open_devices.argtypes = [C.c_char_p, i64, vp, i32, C.POINTER(vp)]
lib.pss_writer_would_flush.argtypes = [vp, sz]
lib.pss_reader_open_devices.argtypes = [C.c_char_p, vp, i32, C.POINTER(vp)]
lib.pss_reader_open_device_chunks.argtypes = [vp, i32, i32, i32, C.POINTER(vp)]
lib.pss_comm_unique_id.argtypes = [vp]
lib.pss_comm_create.argtypes = [vp, i32, i32, C.POINTER(vp)]
lib.pss_comm_destroy.argtypes = [vp]
lib.pss_reader_search_batch_dist.argtypes = [vp, vp, vp, vp, i32, C.POINTER(vp)]
lib.pss_reader_search_batch_dist_device.argtypes = [vp, vp, vp, vp, i32, i64, vp]
COMM_ID_BYTES = 128
lib.pss_result_free.argtypes = [vp]
lib.pss_result_free.restype = None


class PassStat(C.Structure):
    _fields_ = [("round", i32), ("pass_", i32), ("shift", i32), ("reserved", i32), ("n_records", i64),
                ("ms", C.c_float), ("reserved2", C.c_float)]


class BuildStats(C.Structure):
    _fields_ = [("n", i32), ("sigma", i32), ("bits_per_symbol", i32), ("h0", i32), ("rounds", i32),
                ("n_passes", i32), ("n_pass_stats", i32), ("n_kernel_launches", i32),
                ("active_per_round", i64 * 64), ("total_ms", C.c_float), ("sort_ms", C.c_float),
                ("records_sorted", i64)]


class Result(C.Structure):
    _fields_ = [("n_queries", i32), ("reserved", i32), ("n_entries", i64), ("query_offsets", C.POINTER(i64)),
                ("chunk_id", C.POINTER(i32)), ("line_start", C.POINTER(C.c_uint32)),
                ("line_end", C.POINTER(C.c_uint32)), ("n_hits", i64), ("ms_bounds", C.c_float),
                ("ms_extract", C.c_float), ("ms_dedup", C.c_float), ("ms_total", C.c_float),
                ("ms_exchange", C.c_float), ("n_ranks", i32)]


class DeviceChunk(C.Structure):
    _fields_ = [("d_text", vp), ("d_sa", vp), ("h_text", vp), ("n", C.c_uint32), ("global_id", i32)]


class DeviceResult(C.Structure):
    _fields_ = [("n_queries", i32), ("n_chunks", i32), ("n_entries", i64), ("n_hits", i64),
                ("d_query_offsets", vp), ("d_entry_offsets", vp), ("d_chunk_id", vp), ("d_line_start", vp),
                ("d_line_end", vp), ("ms_bounds", C.c_float), ("ms_extract", C.c_float), ("ms_dedup", C.c_float),
                ("ms_exchange", C.c_float)]


def err():
    return lib.pss_last_error().decode()


def check(rc):
    if rc != 0:
        raise RuntimeError("pss rc=%d: %s" % (rc, err()))


def libsais(text):
    t = np.ascontiguousarray(np.frombuffer(bytes(text), dtype=np.uint8) if not isinstance(text, np.ndarray) else text)
    sa = np.empty(len(t), dtype=np.int32)
    check(lib.pss_libsais(t.ctypes.data, sa.ctypes.data, len(t), 0, None))
    return sa


def from_device(ptr, n, dtype):
    """n elements of `dtype` at device pointer `ptr` → numpy array"""
    out = np.empty(n, dtype=dtype)
    if n:
        check(lib.pss_memcpy_d2h(out.ctypes.data, ptr, out.nbytes))
    return out


def result_arrays(res):
    """pss_result* (c_void_p) → (query_offsets, chunk, start, end, stats) numpy copies; frees the result."""
    r = C.cast(res, C.POINTER(Result)).contents
    n, nq = r.n_entries, r.n_queries
    qo = np.ctypeslib.as_array(r.query_offsets, (nq + 1,)).copy()
    if n:
        ch = np.ctypeslib.as_array(r.chunk_id, (n,)).copy()
        st = np.ctypeslib.as_array(r.line_start, (n,)).copy()
        en = np.ctypeslib.as_array(r.line_end, (n,)).copy()
    else:
        ch = np.zeros(0, np.int32)
        st = en = np.zeros(0, np.uint32)
    stats = dict(n_hits=r.n_hits, ms_bounds=r.ms_bounds, ms_extract=r.ms_extract, ms_dedup=r.ms_dedup,
                 ms_total=r.ms_total, ms_exchange=r.ms_exchange, n_ranks=r.n_ranks)
    lib.pss_result_free(res)
    return qo, ch, st, en, stats


def pack(patterns):
    pats = [p.encode() if isinstance(p, str) else bytes(p) for p in patterns]
    offs = np.zeros(len(pats) + 1, dtype=np.int64)
    if pats:
        np.cumsum([len(p) for p in pats], out=offs[1:])
    blob = np.frombuffer(b"".join(pats) + b"\0", dtype=np.uint8).copy()
    return blob, offs


class Comm:
    """Communicator of the distributed search (NCCL inside libpss_b200.so).  `exchange(id_bytes)`
    must hand rank 0's id to every rank (e.g. a torch.distributed broadcast)."""

    def __init__(self, rank, world, exchange=None):
        self.h = vp()
        ident = (C.c_uint8 * COMM_ID_BYTES)()
        if world > 1:
            if rank == 0:
                check(lib.pss_comm_unique_id(ident))
            raw = exchange(bytes(ident))
            ident = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(raw)
        check(lib.pss_comm_create(ident, rank, world, C.byref(self.h)))
        self.rank, self.world = rank, world

    def close(self):
        if self.h:
            lib.pss_comm_destroy(self.h)
            self.h = None


class Writer:
    def __init__(self, path, max_chunk_len=None, devices=None):
        self.h = vp()
        mcl = -1 if max_chunk_len is None else max_chunk_len
        if devices is None:
            check(lib.pss_writer_open(os.fsencode(path), mcl, C.byref(self.h)))
        else:
            arr = (i32 * len(devices))(*devices)
            check(lib.pss_writer_open_devices(os.fsencode(path), mcl, arr, len(devices), C.byref(self.h)))

    def add_entry(self, text):
        b = text.encode() if isinstance(text, str) else bytes(text)
        return lib.pss_writer_add_entry(self.h, b, len(b))

    def add_entries_from_file_lines(self, path):
        return lib.pss_writer_add_entries_from_file_lines(self.h, os.fsencode(path))

    def dump_data(self):
        return lib.pss_writer_dump_data(self.h)

    def finalize(self):
        return lib.pss_writer_finalize(self.h)

    def close(self):
        rc = 0
        if self.h:
            rc = lib.pss_writer_close(self.h)
            self.h = None
        return rc


class Reader:
    def __init__(self, path=None, shard=None, devices=None, device_chunks=None, n_chunks_total=None, device=-1):
        self.h = vp()
        if device_chunks is not None:
            arr = (DeviceChunk * len(device_chunks))(*device_chunks)
            self._keep = arr
            check(lib.pss_reader_open_device_chunks(arr, len(device_chunks), n_chunks_total, device, C.byref(self.h)))
        elif devices is not None:
            arr = (i32 * len(devices))(*devices)
            check(lib.pss_reader_open_devices(os.fsencode(path), arr, len(devices), C.byref(self.h)))
        elif shard is None:
            check(lib.pss_reader_open(os.fsencode(path), C.byref(self.h)))
        else:
            check(lib.pss_reader_open_sharded(os.fsencode(path), shard[0], shard[1], C.byref(self.h)))

    def close(self):
        if self.h:
            lib.pss_reader_close(self.h)
            self.h = None

    @property
    def num_chunks(self):
        return lib.pss_reader_num_chunks(self.h)

    def search_batch(self, patterns):
        """→ (query_offsets, chunk, start, end, stats dict); ordered as the reference orders."""
        blob, offs = pack(patterns)
        res = vp()
        check(lib.pss_reader_search_batch(self.h, blob.ctypes.data, offs.ctypes.data, len(offs) - 1, C.byref(res)))
        return result_arrays(res)

    def search_batch_dist(self, comm, patterns):
        """Collective (every rank calls it with the same patterns); rank 0 gets the merged result."""
        blob, offs = pack(patterns)
        res = vp()
        check(lib.pss_reader_search_batch_dist(self.h, comm.h, blob.ctypes.data, offs.ctypes.data, len(offs) - 1,
                                               C.byref(res)))
        return result_arrays(res)
