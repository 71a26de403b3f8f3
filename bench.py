#!/usr/bin/env python
"""bench.py — headline benchmark of the two hot paths (BASELINE.json metric: index build GB/s
and search_multiple queries/s), one JSON line on stdout from rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE configs[0]+[1] — a 500 000 000-byte synthetic
newline-delimited ASCII chunk (tools/synth.config1_text) whose suffix array is built, then a
search_multiple batch of 10 000 substrings (len 4-32) over that index.  One "step" = one
BUILD of the chunk + one SEARCH batch.  With N > 1 every rank owns one such chunk (its own
seed): chunks shard with no data-path collective for BUILD ("weak" scaling); for SEARCH the
query batch is broadcast and the per-chunk hits are gathered to rank 0 over NCCL.

  value      BUILD, device-resident: text already in HBM, SA left in HBM
  e2e        BUILD through the C-ABI host call (pss_sa_builder_build_host) with pinned host
             buffers: H2D of the text and D2H of the suffix array inside the timed region
  search     same pair of numbers for search_multiple (queries/s), plus the Python boundary
  roofline   the onesweep radix pass: 24 B per record per pass / CUDA-event pass time
  cpu_baseline  the reference's own libsais.c (oracle/_ref) on a bounded sample, 1 thread —
             exactly how the reference runs it (libsais.c:6609 threads = 1)

--impl reference times the reference CPU path (libsais + the lib.rs search restatement).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402

N_TEXT = 500_000_000
N_QUERIES = 10_000
CPU_SAMPLE = int(os.environ.get("PSS_BENCH_CPU_SAMPLE", 64 << 20))   # bytes of the chunk the CPU arm processes per step


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE line (the JSON).  Libraries (NCCL prints its version banner to
# stdout) must not pollute it: fd 1 is pointed at stderr for the whole run and the JSON line
# is written to the saved descriptor at the end.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        in_window = [x for x in self.lines if t0 is None or (t0 - 0.05 <= x[0] <= t1 + 0.05)]
        # a timed region shorter than the sampling period can fall between two samples: then
        # report the samples taken around it (the sampler starts 0.3 s before the region)
        for t, line in (in_window or self.lines):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[0]))
                smax.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------
# reference arm / CPU baseline (the only place bench.py executes oracle/)
# --------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, sample_bytes=CPU_SAMPLE, nq=N_QUERIES):
    from oracle import oracle as O
    kind = "reference" if O.reference_libsais_available() else "port"
    sa_fn = O.suffix_array_reference if kind == "reference" else O.suffix_array_port
    O.use_reference_libsais(kind == "reference")
    text = synth.config1_text(sample_bytes, seed=20240501)
    pats = synth.config2_queries(text, nq=nq, seed=7)
    build_s, search_s = [], []
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "sample.idx")
        reader = None
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            sa = sa_fn(text)                       # libsais(T, SA, n, 0, NULL): lib.rs:29-37
            t1 = time.perf_counter()
            if reader is None:                     # container written once (reference layout)
                with open(path, "wb") as f:
                    f.write(np.uint32(len(text)).tobytes())
                    f.write(text.tobytes())
                    f.write(np.uint32(len(text) * 4).tobytes())
                    f.write(sa.tobytes())
                reader = O.Reader(path)
            t2 = time.perf_counter()
            counts, ch, st, en = reader.search_multiple_tuples(pats)   # lib.rs:201-287 per query
            t3 = time.perf_counter()
            if it >= warmup:
                build_s.append(t1 - t0)
                search_s.append(t3 - t2)
            log("[reference] step %d: libsais %.2fs, search_multiple %.3fs (%d entries)" % (it, t1 - t0, t3 - t2, len(ch)))
    b, s = float(np.mean(build_s)), float(np.mean(search_s))
    sample = "first %d bytes of the config-1 text; %d queries over that 1-chunk index" % (sample_bytes, nq)
    return {
        "kind": kind, "cores": 1, "sample": sample,
        "build_GBps": sample_bytes / b / 1e9, "build_s": b,
        "search_qps": nq / s, "search_s": s,
    }


def workload_config(n, nq, world, n_chunks=0):
    n_chunks = n_chunks or world
    if n_chunks == world:
        what = "BASELINE configs[0]+[1]: %d-byte newline-delimited ASCII chunk per GPU (tools/synth.config1_text), " \
               "SA build + search_multiple of %d substrings (len 4-32) over it" % (n, nq)
    else:
        what = "BASELINE configs[2] shape: %d chunks x %d bytes (tools/synth.config1_text, seed per chunk) sharded " \
               "chunk k -> rank k %% %d, full build + search_multiple of %d substrings" % (n_chunks, n, world, nq)
    return {
        "workload": what, "chunk_bytes": n, "chunks": n_chunks, "queries": nq,
        "l2_policy": "inputs larger than L2 (text %d MB, SA %d MB, sort buffers %d GB)" % (n >> 20, (4 * n) >> 20, (24 * n) >> 30),
        "parallelism": "chunk k -> GPU k %% %d (%d chunk(s) on the busiest GPU)" % (world, -(-n_chunks // world)),
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = cpu_reference_run(args.steps, args.warmup, nq=args.queries)
    line = {
        "impl": "reference", "metric": "index_build_GBps", "value": r["build_GBps"], "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": (r["build_s"] + r["search_s"]) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32", "data": "synthetic",
        "config": dict(workload_config(args.size, args.queries, max(args.gpus, 1), args.chunks),
                       reference_sample="each step = the reference CPU path on the first %d bytes of that chunk" % CPU_SAMPLE),
        "cpu_baseline": {"value": r["build_GBps"], "unit": "GB/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["build_GBps"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "search": {"metric": "search_multiple_qps", "value": r["search_qps"], "unit": "queries/s",
                   "e2e": {"value": r["search_qps"], "unit": "queries/s"}},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pysubstringsearch_b200 import capi as pss   # ctypes binding of include/pss.h
    from pysubstringsearch_b200 import distributed as D

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a collective that cannot complete must fail in minutes, not hold N GPUs for the default 10
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=180))
    torch.cuda.set_device(local_rank)
    pss.check(pss.lib.pss_set_device(local_rank))
    dev = torch.device("cuda", local_rank)
    lib = pss.lib
    n = args.size
    K, W = args.steps, max(args.warmup, 0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs: chunk k of the index belongs to rank k % world ---------------------------------
    n_chunks = args.chunks if args.chunks > 0 else world
    assert n_chunks >= world, "--chunks must be at least the number of ranks"
    chunk_ids = [k for k in range(n_chunks) if D.chunk_owner(k, world) == rank]
    t_gen = time.perf_counter()
    texts = [synth.config1_text(n, seed=20240501 + 1000 * k) for k in chunk_ids]
    text = texts[0]
    pats = synth.config2_queries(text, nq=args.queries, seed=7)     # same seed: rank 0's are used
    blob, offs = synth.pack_patterns(pats)
    log("[rank %d] generated %d x %d bytes + %d queries in %.1fs" % (rank, len(texts), n, len(pats), time.perf_counter() - t_gen))
    h_texts = [torch.from_numpy(t).pin_memory() for t in texts]
    h_sa = torch.empty(n, dtype=torch.int32).pin_memory()
    d_texts = [h.to(dev) for h in h_texts]
    d_sa = torch.empty(n, dtype=torch.int32, device=dev)
    builder = C.c_void_p()
    pss.check(lib.pss_sa_builder_create(local_rank, n, C.byref(builder)))
    pss.check(lib.pss_sa_builder_set_profiling(builder, 1))

    # ---- index for the SEARCH half: this rank's chunks in the reference container, once, untimed ----
    tmpdir = tempfile.TemporaryDirectory()
    path = os.path.join(tmpdir.name, "bench_rank%d.idx" % rank)
    t0 = time.perf_counter()
    with open(path, "wb") as f:                          # container layout of lib.rs:112-119
        for t, h in zip(texts, h_texts):
            pss.check(lib.pss_sa_builder_build_host(builder, h.data_ptr(), n, h_sa.data_ptr()))
            f.write(np.uint32(n).tobytes())
            f.write(memoryview(t))
            f.write(np.uint32(n * 4).tobytes())
            f.write(memoryview(h_sa.numpy()))
    t1 = time.perf_counter()
    reader = pss.Reader(path)
    log("[rank %d] index written (%.1fs) and opened on the GPU (%.1fs)" % (rank, t1 - t0, time.perf_counter() - t1))
    if rank != 0 or args.skip_python:
        tmpdir.cleanup()
    d_blob = torch.from_numpy(blob).to(dev)
    d_offs = torch.from_numpy(offs).to(dev)
    if world > 1:
        # rank 0's batch is THE query batch: agree on its size once, so that the per-step
        # broadcasts below have identical shapes on every rank
        d_blob, d_offs = D.broadcast_queries(d_blob if rank == 0 else None, d_offs if rank == 0 else None, dev, src=0)
        offs = d_offs.cpu().numpy()
        blob = d_blob.cpu().numpy()
        pats = [bytes(blob[offs[i]:offs[i + 1]]) for i in range(len(offs) - 1)]
    cap = 1 << 22
    outs = [torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(4)]

    def search_device():
        """Device-resident SEARCH: patterns in HBM → result tuples in HBM (then, N > 1, gathered to rank 0)."""
        nonlocal cap, outs
        n_entries, n_hits = C.c_int64(0), C.c_int64(0)
        if world > 1:
            dist.broadcast(d_blob, 0)
            dist.broadcast(d_offs, 0)
            torch.cuda.synchronize()
        while True:
            rc = lib.pss_reader_search_batch_device(reader.h, d_blob.data_ptr(), d_offs.data_ptr(), len(pats), int(offs[-1]),
                                                    outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(),
                                                    outs[3].data_ptr(), cap, C.byref(n_entries), C.byref(n_hits), None)
            if rc == -2 and n_entries.value > cap:
                cap = int(n_entries.value * 1.25)
                outs = [torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(4)]
                continue
            pss.check(rc)
            break
        k = n_entries.value
        if world > 1:
            # the one exchange step of the path: per-chunk hit tuples → rank 0 over NCCL
            parts = D.gather_hits(outs[0][:k], outs[1][:k], outs[2][:k], outs[3][:k], dst=0)
            torch.cuda.synchronize()
            if parts is not None:
                k = sum(int(p.shape[1]) for p in parts)
        return k, n_hits.value

    stats = pss.BuildStats()
    pstats = (pss.PassStat * 512)()

    def build_device(acc=None):
        """Device-resident BUILD of every chunk this rank owns (text in HBM → SA in HBM)."""
        for d_text in d_texts:
            pss.check(lib.pss_sa_builder_build_device(builder, d_text.data_ptr(), n, d_sa.data_ptr(), None))
            if acc is not None:
                lib.pss_sa_builder_stats(builder, C.byref(stats), pstats)
                acc["dev_ms"] += stats.total_ms
                for i in range(stats.n_pass_stats):
                    acc["pass_bytes"] += 24.0 * pstats[i].n_records
                    acc["pass_ms"] += pstats[i].ms
                    acc["launches"] += 1

    # ---- device-resident timed region -------------------------------------------------------
    idx = torch.cuda.current_device()
    for _ in range(W):
        build_device()
        search_device()
    # rank 0 samples its own GPU's clocks; one nvidia-smi poller per rank would contend for the
    # driver lock and add milliseconds to every small kernel launch / sync of the other ranks
    sampler = ClockSampler(idx)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = lib.pss_kernel_launch_count()
    barrier()
    t_begin = time.perf_counter()
    build_t = search_t = 0.0
    acc = {"dev_ms": 0.0, "pass_bytes": 0.0, "pass_ms": 0.0, "launches": 0}
    entries = hits = 0
    for _ in range(K):
        a = time.perf_counter()
        build_device(acc)
        b = time.perf_counter()
        # ranks may own different numbers of chunks (15 over 8): align them here, so that the
        # wait for the busiest builder is not booked as search time by the idle ranks
        if world > 1:
            barrier()
        b2 = time.perf_counter()
        entries, hits = search_device()
        c = time.perf_counter()
        build_t += b - a
        search_t += c - b2
    dev_build_ms, pass_bytes, pass_ms, n_pass_launch = acc["dev_ms"], acc["pass_bytes"], acc["pass_ms"], acc["launches"]
    barrier()
    t_end = time.perf_counter()
    launches = lib.pss_kernel_launch_count() - launches0
    clocks = sampler.stop(t_begin, t_end)
    step_s = max_over_ranks((t_end - t_begin) / K)
    build_s = max_over_ranks(build_t / K)
    search_s = max_over_ranks(search_t / K)
    dev_build_s = max_over_ranks(dev_build_ms / K / 1e3)
    rounds, passes = stats.rounds, stats.n_passes
    active = [int(stats.active_per_round[i]) for i in range(stats.rounds + 1)]

    # ---- end-to-end through the C-ABI host calls (pinned host buffers) ------------------------
    def e2e_build():
        for h in h_texts:
            pss.check(lib.pss_sa_builder_build_host(builder, h.data_ptr(), n, h_sa.data_ptr()))

    def e2e_search():
        """The C-ABI host call itself: host patterns in, host result tuples out (then freed)."""
        res = C.c_void_p()
        pss.check(lib.pss_reader_search_batch(reader.h, blob.ctypes.data, offs.ctypes.data, len(pats), C.byref(res)))
        r = C.cast(res, C.POINTER(pss.Result)).contents
        out = (int(r.n_entries), dict(ms_bounds=r.ms_bounds, ms_extract=r.ms_extract, ms_dedup=r.ms_dedup, ms_total=r.ms_total))
        lib.pss_result_free(res)
        return out

    for _ in range(min(W, 2)):
        e2e_build()
        e2e_search()
    barrier()
    eb = es = 0.0
    for _ in range(K):
        a = time.perf_counter()
        e2e_build()
        b = time.perf_counter()
        n_e2e_entries, sstats = e2e_search()
        c = time.perf_counter()
        eb += b - a
        es += c - b
    barrier()
    e2e_build_s = max_over_ranks(eb / K)
    e2e_search_s = max_over_ranks(es / K)
    # the literal drop-in symbol with ordinary (pageable) host memory, as lib.rs:29-37 calls it
    sa_pageable = np.empty(n, dtype=np.int32)
    pss.check(lib.pss_libsais(text.ctypes.data, sa_pageable.ctypes.data, n, 0, None))
    t0 = time.perf_counter()
    pss.check(lib.pss_libsais(text.ctypes.data, sa_pageable.ctypes.data, n, 0, None))
    libsais_pageable_s = max_over_ranks(time.perf_counter() - t0)
    pss.check(lib.pss_sa_builder_build_host(builder, h_texts[0].data_ptr(), n, h_sa.data_ptr()))
    sa_matches = bool(np.array_equal(sa_pageable, h_sa.numpy()))
    del sa_pageable
    h2d_search = int(blob.nbytes + offs.nbytes)
    d2h_search = int(n_e2e_entries * 12 + len(pats) * 4)

    # ---- Python boundary (list[str]) — one measurement, rank 0 ---------------------------------
    py_qps = None
    if rank == 0 and not args.skip_python:
        import pysubstringsearch_b200
        str_pats = [p.decode("ascii") for p in pats]
        py_reader = pysubstringsearch_b200.Reader(index_file_path=path)
        py_reader.search_multiple(substrings=str_pats[:100])
        t0 = time.perf_counter()
        res = py_reader.search_multiple(substrings=str_pats)
        py_qps = len(str_pats) / (time.perf_counter() - t0)
        assert len(res) == n_e2e_entries
        del res, py_reader
        tmpdir.cleanup()

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        r = cpu_reference_run(1, 0)
        cpu = r

    reader.close()
    lib.pss_sa_builder_destroy(builder)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    achieved = pass_bytes / (pass_ms * 1e-3) / 1e9 if pass_ms > 0 else 0.0
    # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel
    # (profiles/r01_ncu_pass_kernel_final.txt): 4.699 GB for 4.606 GB algorithmic → 1.02x
    NCU_TRAFFIC_RATIO = 4.699 / 4.606
    avg_alg_bytes = pass_bytes / n_pass_launch if n_pass_launch else 0.0
    total_bytes = n * n_chunks
    line = {
        "metric": "index_build_GBps", "value": total_bytes / build_s / 1e9, "unit": "GB/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": step_s * 1e3,
        "higher_is_better": True, "scaling": "weak" if n_chunks == world else "strong", "vs_baseline": None,
        "dtype": "u8 text / u32 ranks / u64 keys", "data": "synthetic",
        "config": workload_config(n, len(pats), world, n_chunks),
        "build": {"device_event_ms": dev_build_s * 1e3, "host_call_ms": build_s * 1e3, "rounds": rounds,
                  "radix_passes": passes, "active_per_round": active, "h0_symbols": stats.h0,
                  "bits_per_symbol": stats.bits_per_symbol},
        "e2e": {"value": total_bytes / e2e_build_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": n, "d2h_bytes_per_step": 4 * n,
                "ms_per_step": e2e_build_s * 1e3,
                "pss_libsais_pageable": {"value": n * world / libsais_pageable_s / 1e9, "unit": "GB/s",
                                         "ms": libsais_pageable_s * 1e3, "same_sa_as_pinned_path": sa_matches}},
        "search": {
            "metric": "search_multiple_qps", "value": len(pats) / search_s, "unit": "queries/s",
            "ms_per_batch": search_s * 1e3, "entries": int(entries), "matching_suffixes": int(hits),
            "e2e": {"value": len(pats) / e2e_search_s, "unit": "queries/s", "ms_per_batch": e2e_search_s * 1e3,
                    "h2d_bytes_per_step": h2d_search, "d2h_bytes_per_step": d2h_search},
            "stage_ms": {"bounds": sstats["ms_bounds"], "extract": sstats["ms_extract"], "dedup": sstats["ms_dedup"],
                         "total_device": sstats["ms_total"]},
            "python_boundary_qps": py_qps,
        },
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                     "traffic": avg_alg_bytes * NCU_TRAFFIC_RATIO / 1e9 if n_pass_launch else None,
                     "traffic_unit": "GB per launch (ncu dram read+write = 1.02x algorithmic, profiles/r01_ncu_pass_kernel_final.txt)",
                     "algorithmic_GB_per_launch": avg_alg_bytes / 1e9,
                     "kernel": "onesweep_pass_kernel", "peak_source": peak_kind,
                     "algorithmic_bytes": "24 B per record per pass (8 B key + 4 B value, read once + written once)",
                     "launches_timed": n_pass_launch, "avg_launch_ms": pass_ms / n_pass_launch if n_pass_launch else None,
                     "share_of_build": (pass_ms / K) / (dev_build_s * 1e3) if dev_build_s else None},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if cpu:
        line["cpu_baseline"] = {"value": cpu["build_GBps"], "unit": "GB/s", "cores": cpu["cores"], "kind": cpu["kind"],
                                "sample": cpu["sample"], "search_qps": cpu["search_qps"]}
    emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=N_TEXT, help="chunk bytes per GPU (default: the 500 MB config)")
    ap.add_argument("--queries", type=int, default=N_QUERIES)
    ap.add_argument("--chunks", type=int, default=0,
                    help="total chunks of the index, sharded chunk k -> rank k %% N (default: one per GPU = weak scaling; "
                         "15 with --size 536870912 is BASELINE configs[2])")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-python", action="store_true")
    args = ap.parse_args()
    capture_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
