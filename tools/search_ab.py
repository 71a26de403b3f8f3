#!/usr/bin/env python
"""A/B of the extraction step on one config-3 chunk: PSS_LINE_DIR=0 (text scans) vs 1 (line
directory).  Builds the chunk's suffix array once, opens the index twice and runs the same
10 000-query batch (plus one high-hit bigram) through the C ABI, printing the stage times and
checking that both readers return identical tuples.  usage: search_ab.py [n_bytes] [repeat]"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pysubstringsearch_b200 import capi as pss  # noqa: E402
from tools import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 29
repeat = int(sys.argv[2]) if len(sys.argv) > 2 else 5
text = synth.config3_chunk(1, n, device="cuda")
sa = pss.libsais(text)
pats = synth.config2_queries(text, nq=10000, seed=7)
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
    p = os.path.join(d, "ab.idx")
    with open(p, "wb") as f:
        f.write(np.uint32(n).tobytes()); f.write(memoryview(text)); f.write(np.uint32(4 * n).tobytes()); f.write(memoryview(sa))
    results = {}
    for mode in ("0", "1"):
        os.environ["PSS_LINE_DIR"] = mode
        r = pss.Reader(p)
        for label, batch in (("10k batch", pats), ("bigram", [b"e "]), ("google", [b"google"])):
            best = None
            for _ in range(repeat):
                qo, ch, st, en, stats = r.search_batch(batch)
                if best is None or stats["ms_total"] < best["ms_total"]:
                    best = stats
            results[(mode, label)] = (qo, ch, st, en)
            print("PSS_LINE_DIR=%s %-9s entries=%d hits=%d bounds %.3f extract %.3f dedup %.3f total %.3f ms" % (
                mode, label, len(ch), best["n_hits"], best["ms_bounds"], best["ms_extract"], best["ms_dedup"], best["ms_total"]))
        r.close()
    for label in ("10k batch", "bigram", "google"):
        a, b = results[("0", label)], results[("1", label)]
        same = all(np.array_equal(x, y) for x, y in zip(a, b))
        print("identical tuples (%s): %s" % (label, same))
        assert same
