#!/usr/bin/env python
"""Small driver for ncu: builds one suffix array (and optionally runs a query batch) through the
C ABI.  usage: profile_build.py [n_bytes] [kind=words|acgt|c3] [search=0|1] [repeat]"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pysubstringsearch_b200 import capi as pss  # noqa: E402
from tools import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
kind = sys.argv[2] if len(sys.argv) > 2 else "words"
do_search = len(sys.argv) > 3 and sys.argv[3] == "1"
repeat = int(sys.argv[4]) if len(sys.argv) > 4 else 1
if kind == "c3":      # chunk 1 of the config-3 corpus, generated on the GPU (fast)
    text = synth.config3_chunk(1, n, device="cuda")
else:
    text = synth.config1_text(n) if kind == "words" else synth.acgt_text(n)
sa = np.empty(n, dtype=np.int32)
b = C.c_void_p()
pss.check(pss.lib.pss_sa_builder_create(-1, n, C.byref(b)))
pss.lib.pss_sa_builder_set_profiling(b, 1)
st = pss.BuildStats()
ps = (pss.PassStat * 512)()
for _ in range(repeat):
    pss.check(pss.lib.pss_sa_builder_build_host(b, text.ctypes.data, n, sa.ctypes.data))
    pss.lib.pss_sa_builder_stats(b, C.byref(st), ps)
    print("build n=%d total_ms=%.3f sort_ms=%.3f rounds=%d passes=%d launches=%d active=%s" % (
        n, st.total_ms, st.sort_ms, st.rounds, st.n_passes, st.n_kernel_launches,
        [int(st.active_per_round[i]) for i in range(st.rounds + 1)]))
    tot_b = sum(24.0 * ps[i].n_records for i in range(st.n_pass_stats))
    print("pass GB/s avg %.1f" % (tot_b / (st.sort_ms * 1e-3) / 1e9))
    if os.environ.get("PSS_PRINT_PASSES"):
        for i in range(st.n_pass_stats):
            q = ps[i]
            print("  round %d shift %2d spread %5.1f n=%10d %.3f ms %7.1f GB/s" % (q.round, q.shift, q.reserved / 1000.0, q.n_records, q.ms, 24.0 * q.n_records / (q.ms * 1e-3) / 1e9))
if do_search:
    pats = synth.config2_queries(text, nq=10000, seed=7)
    rep = int(os.environ.get("PSS_PROFILE_QUERY_REPEAT", "1"))   # 15: as many (query, chunk) pairs as config 3 has
    pats = pats * rep if rep > 1 else pats + [b"sojq", b"google", b"e "]
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "p.idx")
        with open(p, "wb") as f:
            f.write(np.uint32(n).tobytes()); f.write(memoryview(text)); f.write(np.uint32(4 * n).tobytes()); f.write(memoryview(sa))
        r = pss.Reader(p)
    for _ in range(repeat):
        qo, ch, st_, en, stats = r.search_batch(pats)
        print("search entries=%d %s" % (len(ch), stats))
    r.close()
