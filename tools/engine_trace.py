"""Times the asynchronous build seam on a few full-size chunks (PSS_ENGINE_TRACE=1 prints the
engine's own per-job timings): pinned host text → pinned host suffix array, all builds queued."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pysubstringsearch_b200 import capi as pss  # noqa: E402
from tools import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 29
k = int(sys.argv[2]) if len(sys.argv) > 2 else 5
texts = []
for i in range(k):
    t = synth.config3_chunk_torch(i, n, device="cuda")[:n]
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.copy_(t)
    texts.append(h)
torch.cuda.synchronize()
sa = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(2)]
for rep in range(3):
    t0 = time.perf_counter()
    hs = []
    for h in texts:
        x = C.c_void_p()
        pss.check(pss.lib.pss_sa_build_begin(0, h.data_ptr(), n, C.byref(x)))
        hs.append(x)
    marks = []
    for i, x in enumerate(hs):
        pss.check(pss.lib.pss_sa_build_wait(x, sa[i % 2].data_ptr()))
        marks.append(time.perf_counter() - t0)
    print("rep %d: %d chunks of %d bytes: waits returned at %s ms → %.1f ms per chunk, %.2f GB/s" % (
        rep, k, n, [round(m * 1e3) for m in marks], marks[-1] * 1e3 / k, k * n / marks[-1] / 1e9), flush=True)
# raw copy rates for reference
d = torch.empty(n, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
for name, fn, nbytes in (("D2H 2 GiB", lambda: sa[0].copy_(d, non_blocking=True), 4 * n),
                         ("H2D 0.5 GiB", lambda: d.view(torch.uint8)[:n].copy_(texts[0], non_blocking=True), n)):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%s: %.1f ms = %.1f GB/s" % (name, dt * 1e3, nbytes / dt / 1e9))
