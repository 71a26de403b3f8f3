#!/usr/bin/env python
"""A/B of the search kernels' variants on one config-3 chunk: PSS_LINE_DIR=0 (extraction scans
the text) vs 1 (line directory) vs 2 (line records); with PSS_AB_BOUNDS=1 instead the bounds
kernel's geometries (PSS_BOUNDS_GROUP = lanes per pair, negative = no SA look-ahead).  Builds the chunk's suffix array once, opens the index once per variant and runs the same
10 000-query batch (plus one high-hit bigram) through the C ABI, printing the stage times and
checking that every variant returns identical tuples.  usage: search_ab.py [n_bytes] [repeat]"""
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
    modes = [("scan", {"PSS_LINE_DIR": "0"}), ("directory", {"PSS_LINE_DIR": "1"}), ("records", {"PSS_LINE_DIR": "2"})]
    if os.environ.get("PSS_AB_BOUNDS"):
        modes = [("g%s" % g, {"PSS_BOUNDS_GROUP": g}) for g in ("32", "-32", "8", "-8", "4", "-4")]
    for mode, env in modes:
        os.environ.update(env)
        r = pss.Reader(p)
        for label, batch in (("10k batch", pats), ("150k batch", pats * 15), ("bigram", [b"e "]), ("google", [b"google"])):
            best = None
            for _ in range(repeat):
                qo, ch, st, en, stats = r.search_batch(batch)
                if best is None or stats["ms_total"] < best["ms_total"]:
                    best = stats
            results[(mode, label)] = (qo, ch, st, en)
            print("%-22s %-10s entries=%d hits=%d bounds %.3f extract %.3f dedup %.3f total %.3f ms" % (
                mode, label, len(ch), best["n_hits"], best["ms_bounds"], best["ms_extract"], best["ms_dedup"], best["ms_total"]))
        r.close()
    for label in ("10k batch", "150k batch", "bigram", "google"):
        for mode, _ in modes[1:]:
            a, b = results[(modes[0][0], label)], results[(mode, label)]
            same = all(np.array_equal(x, y) for x, y in zip(a, b))
            print("identical tuples (%s, %s vs %s): %s" % (label, mode, modes[0][0], same))
            assert same
