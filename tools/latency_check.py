import os, sys, time, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pysubstringsearch_b200 import capi as pss
import pysubstringsearch_b200 as api
from tools import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64 << 20
text = synth.config1_text(n)
sa = pss.libsais(text)
with tempfile.TemporaryDirectory() as d:
    p = os.path.join(d, "c.idx")
    with open(p, "wb") as f:
        f.write(np.uint32(n).tobytes()); f.write(memoryview(text)); f.write(np.uint32(4 * n).tobytes()); f.write(memoryview(sa))
    reader = api.Reader(index_file_path=p)
    raw = pss.Reader(p)
for q in ("google", "text_two", "zzzzzz", "sojq"):
    reader.search(substring=q)
    ts = []
    for _ in range(100):
        t0 = time.perf_counter(); res = reader.search(substring=q); ts.append(time.perf_counter() - t0)
    r = raw.search_batch([q])
    ts2 = []
    for _ in range(100):
        t0 = time.perf_counter(); r = raw.search_batch([q]); ts2.append(time.perf_counter() - t0)
    print("%-10s results=%7d python median %.1f us  min %.1f us | c-abi(ctypes) median %.1f us  kernel %.3f ms" % (
        q, len(res), np.median(ts) * 1e6, np.min(ts) * 1e6, np.median(ts2) * 1e6, r[4]["ms_bounds"]))
