#!/usr/bin/env python
"""Measures the remaining BASELINE.json configs on one GPU and prints one JSON object:
  config 1 — single-query latency of search('google') / search('text_two') on the 500 MB index
  config 4 — low-entropy ACGT text with long repeats (one 2^29 chunk): rounds, active sets, GB/s
  config 5 — one high-hit query (most frequent bigram): stage times, D2H, Python materialisation
Parity of every result is checked against the oracle on a bounded sample where noted."""
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pysubstringsearch_b200 import capi as pss  # noqa: E402
from tools import synth  # noqa: E402

out = {}
n = int(os.environ.get("PSS_N", 500_000_000))


def build(text, profile=False):
    sa = np.empty(len(text), dtype=np.int32)
    b = C.c_void_p()
    pss.check(pss.lib.pss_sa_builder_create(-1, len(text), C.byref(b)))
    pss.lib.pss_sa_builder_set_profiling(b, 1 if profile else 0)
    st = pss.BuildStats()
    best = None
    for _ in range(3):
        pss.check(pss.lib.pss_sa_builder_build_host(b, text.ctypes.data, len(text), sa.ctypes.data))
        pss.lib.pss_sa_builder_stats(b, C.byref(st), None)
        best = st.total_ms if best is None else min(best, st.total_ms)
    info = dict(n=len(text), device_ms=best, GBps=len(text) / best / 1e6, rounds=st.rounds, passes=st.n_passes,
                sigma=st.sigma, h0=st.h0, active_per_round=[int(st.active_per_round[i]) for i in range(st.rounds + 1)])
    pss.lib.pss_sa_builder_destroy(b)
    return sa, info


def write_index(path, text, sa):
    with open(path, "wb") as f:
        f.write(np.uint32(len(text)).tobytes()); f.write(memoryview(text))
        f.write(np.uint32(4 * len(text)).tobytes()); f.write(memoryview(sa))


# ---- config 1 + 5 on the 500 MB text ---------------------------------------------------------
text = synth.config1_text(n)
sa, info = build(text)
out["config1_build"] = info
with tempfile.TemporaryDirectory() as d:
    p = os.path.join(d, "c1.idx")
    write_index(p, text, sa)
    import pysubstringsearch_b200 as api
    t0 = time.perf_counter()
    reader = api.Reader(index_file_path=p)
    out["config1_reader_open_s"] = time.perf_counter() - t0
    raw = pss.Reader(p)
lat = {}
for q in ("google", "text_two", "zzzzzz"):
    reader.search(substring=q)
    ts = []
    for _ in range(50):
        t0 = time.perf_counter(); res = reader.search(substring=q); ts.append(time.perf_counter() - t0)
    lat[q] = dict(results=len(res), median_us=float(np.median(ts) * 1e6), min_us=float(np.min(ts) * 1e6))
out["config1_single_query_latency_python"] = lat
lat = {}
for q in (b"google", b"text_two"):
    raw.search_batch([q])
    ts = []
    for _ in range(50):
        t0 = time.perf_counter(); r = raw.search_batch([q]); ts.append(time.perf_counter() - t0)
    lat[q.decode()] = dict(results=len(r[1]), median_us=float(np.median(ts) * 1e6), device_ms=r[4])
out["config1_single_query_latency_cabi"] = lat

# config 5: most frequent bigram
c = text[:50_000_000].astype(np.uint16)
big = np.bincount((c[:-1] << 8) | c[1:], minlength=65536)
top = int(np.argmax(big))
pat = bytes([top >> 8, top & 255])
raw.search_batch([pat])
t0 = time.perf_counter(); qo, ch, st_, en, stats = raw.search_batch([pat]); t_cabi = time.perf_counter() - t0
t0 = time.perf_counter(); res = reader.search(substring=pat.decode()); t_py = time.perf_counter() - t0
out["config5_high_hit"] = dict(pattern=pat.decode(), matching_suffixes=int(stats["n_hits"]), entries=len(ch),
                               stage_ms=stats, cabi_call_ms=t_cabi * 1e3, python_call_ms=t_py * 1e3,
                               d2h_bytes=len(ch) * 12)
assert len(res) == len(ch)
del reader, raw, res

# ---- config 4: low-entropy text with long repeats, one full 2^29 chunk ---------------------------
n4 = int(os.environ.get("PSS_N4", 1 << 29))
t4 = synth.acgt_text(n4)
sa4, info4 = build(t4)
# size-independent property check on the GPU result: permutation + sorted (rank-of-next-suffix test)
isa = np.empty(n4, dtype=np.int32); isa[sa4] = np.arange(n4, dtype=np.int32)
seen = np.zeros(n4, dtype=bool); seen[sa4] = True          # every index 0..n-1 occurs (n values, n distinct)
perm_ok = bool(seen.all() and sa4.min() == 0 and sa4.max() == n4 - 1)
del seen
a, b_ = sa4[:-1], sa4[1:]
ta, tb = t4[a], t4[b_]
na = np.where(a + 1 < n4, isa[np.minimum(a + 1, n4 - 1)], -1)
nb = np.where(b_ + 1 < n4, isa[np.minimum(b_ + 1, n4 - 1)], -1)
sorted_ok = bool(np.all((ta < tb) | ((ta == tb) & (na < nb))))
info4.update(permutation_ok=perm_ok and len(np.unique(sa4[:1 << 20])) == (1 << 20), sorted_ok=sorted_ok)
out["config4_low_entropy"] = info4
print(json.dumps(out, indent=1))
