"""CPU tests of the C-ABI library and the Python boundary: it loads, exports what
include/pss.h declares, and its host-side logic (argument checks, error mapping, container
validation) behaves — no compute call needs a GPU here."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

from tests.conftest import HAVE_GPU, ROOT


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "pss.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pss_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(pss):
    syms = declared_symbols()
    assert len(syms) >= 42
    for s in syms:
        assert hasattr(pss.lib, s), "libpss_b200.so does not export %s" % s


def test_version_and_device_count(pss):
    assert b"sm_100a" in pss.lib.pss_version()
    assert pss.lib.pss_device_count() >= 0


def test_libsais_argument_contract(pss):
    # libsais.c:6599-6602: NULL / negative arguments → -1, before any device work
    assert pss.lib.pss_libsais(None, None, 5, 0, None) == -1
    buf = (C.c_uint8 * 4)()
    sa = (C.c_int32 * 4)()
    assert pss.lib.pss_libsais(buf, sa, -1, 0, None) == -1
    assert pss.lib.pss_libsais(buf, sa, 4, -1, None) == -1
    assert pss.lib.pss_libsais(buf, sa, 0, 0, None) == 0      # n == 0 is a no-op


def test_reader_missing_file(pss):
    h = C.c_void_p()
    rc = pss.lib.pss_reader_open(b"/nonexistent/dir/missing.idx", C.byref(h))
    assert rc == -5 and "No such file" in pss.err()


def test_reader_rejects_truncated_container(pss):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "bad.idx").encode()
        h = C.c_void_p()
        good = bytes.fromhex("0300000061620a0c000000020000000000000001000000")
        for cut in (2, 6, 9, 15, len(good) - 1):
            open(p, "wb").write(good[:cut])
            assert pss.lib.pss_reader_open(p, C.byref(h)) == -7, cut
        open(p, "wb").write(good[:7] + bytes.fromhex("08000000") + good[11:])   # sa_bytes != 4n
        assert pss.lib.pss_reader_open(p, C.byref(h)) == -7


def test_writer_host_logic(pss):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "w.idx")
        w = pss.Writer(p, max_chunk_len=4)
        assert w.add_entry("12345") == -6 and "too big" in pss.err()
        assert w.add_entry("123") == 0
        if not HAVE_GPU:
            # the flush needs the GPU builder: must fail loudly, never fall back to a CPU path
            assert w.add_entry("x") == -3
            assert "no CUDA device" in pss.err() or "CUDA" in pss.err()
            assert w.finalize() == -3
        h = C.c_void_p()
        assert pss.lib.pss_writer_open(os.path.join(d, "no/such/dir/x.idx").encode(), -1, C.byref(h)) == -5
        assert w.add_entries_from_file_lines(os.path.join(d, "missing.txt")) == -5


def test_new_entry_points_check_their_arguments(pss):
    """Async build seam, device lists, communicator: argument errors surface as status codes
    before any device work (no GPU needed)."""
    h = C.c_void_p()
    assert pss.lib.pss_sa_build_begin(-1, None, 5, C.byref(h)) == -1
    assert pss.lib.pss_sa_build_begin(-1, None, -1, C.byref(h)) == -1
    assert pss.lib.pss_sa_build_wait(None, None) == -1
    assert pss.lib.pss_release_cached() == 0
    assert pss.lib.pss_reader_open_devices(b"/nonexistent/x.idx", None, 2, C.byref(h)) == -1
    assert pss.lib.pss_reader_open_device_chunks(None, 1, 1, -1, C.byref(h)) == -1
    assert pss.lib.pss_comm_create(None, 3, 2, C.byref(h)) == -1
    assert pss.lib.pss_comm_create(None, 0, 0, C.byref(h)) == -1
    if not HAVE_GPU:
        buf = (C.c_uint8 * 4)(97, 98, 99, 10)
        assert pss.lib.pss_sa_build_begin(-1, buf, 4, C.byref(h)) == -3 and "no CUDA device" in pss.err()
        assert pss.lib.pss_comm_create(None, 0, 1, C.byref(h)) == -3


@pytest.mark.filterwarnings("ignore::pytest.PytestUnraisableExceptionWarning")   # a Writer with buffered entries is dropped without a GPU
def test_python_module_surface():
    import pysubstringsearch
    import pysubstringsearch_b200 as m
    assert pysubstringsearch.Writer is m.Writer and pysubstringsearch.Reader is m.Reader
    for cls, names in ((m.Writer, ["add_entries_from_file_lines", "add_entry", "dump_data", "finalize"]),
                       (m.Reader, ["search", "search_multiple"])):
        for n in names:
            assert callable(getattr(cls, n))
    with pytest.raises(FileNotFoundError):
        m.Reader(index_file_path="missing_index_file_path")
    with pytest.raises(TypeError):
        m.Reader(index_file_path=b"bytes-not-str")
    with tempfile.TemporaryDirectory() as d:
        w = m.Writer(index_file_path=os.path.join(d, "a.idx"), max_chunk_len=3)
        with pytest.raises(ValueError, match="entry is too big"):
            w.add_entry(text="abcd")
        with pytest.raises(TypeError):
            w.add_entry(text=b"ab")
        with pytest.raises(FileNotFoundError):
            w.add_entries_from_file_lines(input_file_path=os.path.join(d, "nope.txt"))
        # add_entry uses the vectorcall convention with a fast path for add_entry(s) / add_entry(text=s);
        # every other call shape must still fail like the generic parser.  (The entries stay buffered:
        # nothing is flushed here, and the failing close at teardown — no GPU — is reported, not raised.)
        big = m.Writer(index_file_path=os.path.join(d, "b.idx"), max_chunk_len=1 << 20)
        big.writer.add_entry("positional")
        big.writer.add_entry(text="keyword")
        big.add_entry("façade, non-ASCII ✓")
        for bad in (lambda: big.writer.add_entry(), lambda: big.writer.add_entry(b"bytes"), lambda: big.writer.add_entry(txt="x"),
                    lambda: big.writer.add_entry("a", "b"), lambda: big.writer.add_entry("a", text="b"),
                    lambda: big.writer.add_entry(text="a", other=1)):
            with pytest.raises(TypeError):
                bad()


def test_product_never_imports_oracle():
    """The product path must not reference oracle/ (a CPU fallback would void parity)."""
    pkg = os.path.join(ROOT, "pysubstringsearch_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"(import|from|include|dlopen|CDLL)[^\n]*oracle", src), \
                    "%s pulls in the oracle" % f
    out = subprocess.run([sys.executable, "-c",
                          "import sys, pysubstringsearch_b200; print(any(m.startswith('oracle') for m in sys.modules))"],
                         cwd=ROOT, capture_output=True, text=True)
    assert out.stdout.strip() == "False", out.stderr


def test_ctypes_structs_match_the_header(pss):
    """ABI drift guard: sizeof/offsetof of the public structs as gcc sees include/pss.h must
    equal the ctypes mirrors in pysubstringsearch_b200/capi.py (and INTEGRATION.md's repr(C))."""
    import ctypes as C
    src = r'''
#include <stddef.h>
#include <stdio.h>
#include "pss.h"
#define P(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void) {
    printf("pss_result %zu\n", sizeof(pss_result));
    P(pss_result, n_queries); P(pss_result, n_entries); P(pss_result, query_offsets); P(pss_result, chunk_id);
    P(pss_result, line_start); P(pss_result, line_end); P(pss_result, n_hits); P(pss_result, ms_bounds); P(pss_result, ms_total);
    printf("pss_pass_stat %zu\n", sizeof(pss_pass_stat));
    P(pss_pass_stat, shift); P(pss_pass_stat, n_records); P(pss_pass_stat, ms);
    printf("pss_build_stats %zu\n", sizeof(pss_build_stats));
    P(pss_build_stats, n_kernel_launches); P(pss_build_stats, active_per_round); P(pss_build_stats, total_ms);
    P(pss_build_stats, records_sorted);
    P(pss_result, ms_exchange); P(pss_result, n_ranks);
    printf("pss_device_chunk %zu\n", sizeof(pss_device_chunk));
    P(pss_device_chunk, d_sa); P(pss_device_chunk, h_text); P(pss_device_chunk, n); P(pss_device_chunk, global_id);
    printf("pss_device_result %zu\n", sizeof(pss_device_result));
    P(pss_device_result, n_chunks); P(pss_device_result, n_entries); P(pss_device_result, d_query_offsets);
    P(pss_device_result, d_entry_offsets); P(pss_device_result, d_line_end); P(pss_device_result, ms_exchange);
    return 0;
}
'''
    with tempfile.TemporaryDirectory() as d:
        c, exe = os.path.join(d, "abi.c"), os.path.join(d, "abi")
        open(c, "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        got = dict(line.rsplit(" ", 1) for line in subprocess.check_output([exe], text=True).splitlines())
    mirrors = {"pss_result": pss.Result, "pss_pass_stat": pss.PassStat, "pss_build_stats": pss.BuildStats,
               "pss_device_chunk": pss.DeviceChunk, "pss_device_result": pss.DeviceResult}
    for key, val in got.items():
        if "." in key:
            struct, field = key.split(".")
            field = {"pass": "pass_"}.get(field, field)
            assert getattr(mirrors[struct], field).offset == int(val), key
        else:
            assert C.sizeof(mirrors[key]) == int(val), key


def test_header_is_plain_c_and_static_archive_links(tmp_path):
    """include/pss.h must stay bindable from C / Rust (plain C99, no C++), and the static archive a
    build.rs would link (`make static`) must resolve with nothing but the
    static CUDA runtime: a C program is compiled against the header with -pedantic, linked against
    libpss_b200.a and run (no GPU needed for pss_version / a failing pss_reader_open)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "pysubstringsearch_b200", "csrc")
    subprocess.check_call(["make", "-C", csrc, "static"], stdout=subprocess.DEVNULL)
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc")), "lib64")
    src = tmp_path / "smoke.c"
    src.write_text(
        '#include <stdio.h>\n#include "pss.h"\n'
        'int main(void) {\n'
        '    pss_reader *r = 0;\n'
        '    int rc = pss_reader_open("/nonexistent/file.idx", &r);\n'
        '    printf("%s|%d|%s\\n", pss_version(), rc, pss_last_error());\n'
        '    return 0;\n}\n')
    obj, exe = str(tmp_path / "smoke.o"), str(tmp_path / "smoke")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                           "-c", str(src), "-o", obj])
    subprocess.check_call(["g++", obj, "-L", os.path.join(root, "pysubstringsearch_b200"), "-l:libpss_b200.a",
                           "-L", cuda_lib, "-lcudart_static", "-lpthread", "-ldl", "-lrt", "-o", exe])
    out = subprocess.check_output([exe]).decode().strip().split("|")
    assert "sm_100a" in out[0] and int(out[1]) != 0 and "No such file" in out[2]
    needed = subprocess.check_output(["ldd", exe]).decode()
    assert "libcudart" not in needed and "libpss" not in needed      # everything CUDA-side is inside the binary


def test_libsais_named_shim_forwards_the_contract():
    """The optional shim exports the reference's own foreign symbol `libsais` (lib.rs:14-22): same
    argument contract as pss_libsais, checked here without a GPU; with one it must build the same
    suffix array (tests/test_gpu_parity.py::test_libsais_shim_on_gpu)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "pysubstringsearch_b200", "csrc"), "shim"], stdout=subprocess.DEVNULL)
    shim = C.CDLL(os.path.join(root, "pysubstringsearch_b200", "libsais_pss_shim.so"))
    shim.libsais.restype = C.c_int32
    shim.libsais.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    buf = (C.c_uint8 * 4)()
    sa = (C.c_int32 * 4)()
    assert shim.libsais(None, None, 5, 0, None) == -1
    assert shim.libsais(buf, sa, -1, 0, None) == -1
    assert shim.libsais(buf, sa, 4, -1, None) == -1
    assert shim.libsais(buf, sa, 0, 0, None) == 0
