"""pysubstringsearch_b200 — B200 (sm_100a) implementation of PySubstringSearch's two hot
paths behind the reference's own Python surface.

Same classes, method names, keyword names and exception types as the reference façade
(/root/reference/pysubstringsearch/__init__.py:6-73, stubs pysubstringsearch.pyi:1-44):

    Writer(index_file_path, max_chunk_len=None)
        .add_entries_from_file_lines(input_file_path) / .add_entry(text) / .dump_data() / .finalize()
    Reader(index_file_path)
        .search(substring) -> list[str] / .search_multiple(substrings) -> list[str]

The native module `pysubstringsearch_b200.pysubstringsearch` (csrc/pymodule.cpp) sits on
the C ABI of include/pss.h (libpss_b200.so): suffix arrays are built by a GPU
prefix-doubling builder, searches run as batched CUDA kernels.  There is no CPU fallback:
importing works anywhere the extension is built, but building an index or opening a
Reader needs a CUDA device and fails loudly without one.

`search_multiple` is ONE native batched call (the reference loops in Python,
__init__.py:61-73); it returns the same concatenation in query order.
"""
import typing

try:
    from . import pysubstringsearch
except ImportError as exc:  # pragma: no cover - build problem, never a silent fallback
    raise ImportError(
        "pysubstringsearch_b200: the native extension is not built "
        "(run `python -c 'import __graft_entry__ as g; g.build()'` or "
        "`make -C pysubstringsearch_b200/csrc`): %s" % (exc,)
    ) from exc

__all__ = ["Writer", "Reader"]
__version__ = "0.1.0"


class Writer:
    """Accumulates newline-terminated entries into chunks of at most `max_chunk_len` bytes
    (default 512 MiB) and writes each chunk with its suffix array to `index_file_path`."""

    def __init__(self, index_file_path: str, max_chunk_len: typing.Optional[int] = None) -> None:
        self.writer = pysubstringsearch.Writer(index_file_path=index_file_path, max_chunk_len=max_chunk_len)

    def add_entries_from_file_lines(self, input_file_path: str) -> None:
        self.writer.add_entries_from_file_lines(input_file_path=input_file_path)

    def add_entry(self, text: str) -> None:
        self.writer.add_entry(text=text)

    def dump_data(self) -> None:
        self.writer.dump_data()

    def finalize(self) -> None:
        self.writer.finalize()


class Reader:
    """Loads every chunk of an index file onto the GPU and answers substring queries."""

    def __init__(self, index_file_path: str) -> None:
        self.reader = pysubstringsearch.Reader(index_file_path=index_file_path)

    def search(self, substring: str) -> typing.List[str]:
        return self.reader.search(substring=substring)

    def search_multiple(self, substrings: typing.List[str]) -> typing.List[str]:
        return self.reader.search_multiple(substrings=list(substrings))
