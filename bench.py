#!/usr/bin/env python
"""bench.py — headline benchmark of the two hot paths (BASELINE.json metric: index build GB/s
and search_multiple queries/s at 1/2/4/8 B200), one JSON line on stdout from rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload at EVERY N (config.workload): BASELINE configs[2] — the 7.5 GiB multi-chunk index:
15 chunks x 2^29 bytes of synthetic newline-delimited ASCII over ONE shared vocabulary
(tools/synth.config3_chunk_torch), chunk k → GPU k % N ("strong" scaling: the total work is
fixed), plus a config-2-style search_multiple batch of 10 000 substrings (len 4-32) cut from
all 15 chunks.  One "step" = the suffix arrays of all 15 chunks are built, then the batch is
searched over the whole index.  It fits one B200 (37.5 GiB of index), so N = 1 runs the same
workload and the per-N values are directly comparable.

  value      BUILD, device-resident: texts already in HBM, suffix arrays left in HBM;
             GB/s = 15 x 2^29 bytes / time of the slowest rank
  e2e        BUILD through the C ABI's host seam (pss_sa_build_begin / pss_sa_build_wait):
             pinned host text in, pinned host suffix array out, H2D + D2H inside the timed
             region, device→host copy of chunk k overlapped with the build of chunk k+1
  search     search_multiple, queries/s: `value` device-resident (batch in rank 0's HBM →
             merged tuples in rank 0's HBM, through broadcast → local search → NCCL gather-v →
             placement, all inside libpss_b200.so), `e2e` host patterns → host tuples on rank 0
  roofline   the onesweep radix pass: 24 B per record per pass / CUDA-event pass time
  cpu_baseline / --impl reference
             the reference's own libsais.c (oracle/_ref, 1 thread — libsais.c:6609) on ONE
             full-size chunk of the same corpus (chunks are built serially by the reference,
             so its GB/s on one chunk is its GB/s on the index), and the lib.rs search
             restatement answering the SAME 10 000 queries over that chunk's index
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

N_QUERIES = 10_000


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
# workload
# --------------------------------------------------------------------------------------
def workload_config(n, nq, world, n_chunks):
    return {
        "workload": "BASELINE configs[2]: %d chunks x %d bytes of newline-delimited ASCII (tools/synth.config3_chunk_torch: "
                    "Zipf words over one shared 65 536-word vocabulary, seed per chunk; chunk 0 carries config 1's planted "
                    "'google' x 5943 / 'text_two' x 159 lines), full build + search_multiple of %d substrings (len 4-32; 90%% "
                    "cut from all %d chunks in turn and rejection-sampled to be selective: rarest 4-gram <= 5000 occurrences "
                    "per chunk, 10%% random lowercase)" % (n_chunks, n, nq, n_chunks),
        "chunk_bytes": n, "chunks": n_chunks, "queries": nq, "index_bytes": n_chunks * (8 + 5 * n),
        "l2_policy": "inputs larger than L2 (per chunk: text %d MB, SA %d MB, sort buffers %d GB)" % (n >> 20, (4 * n) >> 20, (24 * n) >> 30),
        "parallelism": "chunk k -> GPU k %% %d (%d chunk(s) on the busiest GPU, ideal speed-up %.3fx)" % (
            world, -(-n_chunks // world), n_chunks / float(-(-n_chunks // world))),
    }


def make_queries(n, n_chunks, nq, device):
    """The query batch: cut from the first 32 MiB of every chunk (any process can regenerate it)."""
    m = min(synth.CONFIG3_PREFIX, n)
    prefixes = [synth.config3_chunk_torch(k, m, device=device, force_newline=False)[:m] for k in range(n_chunks)]
    return synth.config3_queries(prefixes, nq=nq, seed=7, chunk_bytes=n)


def write_container(path, text, sa):
    """One chunk in the reference container layout (lib.rs:112-119)."""
    with open(path, "wb") as f:
        f.write(np.uint32(len(text)).tobytes())
        f.write(memoryview(text))
        f.write(np.uint32(len(text) * 4).tobytes())
        f.write(memoryview(sa))


def median_latency_us(fn, reps=200):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts) * 1e6)


# --------------------------------------------------------------------------------------
# reference arm / CPU baseline (the only place bench.py executes oracle/)
# --------------------------------------------------------------------------------------
def cpu_reference_run(text0, pats, search_steps, search_warmup, path=None, check=None):
    """The reference's CPU path on ONE full-size chunk (chunk 0): libsais timed once, then the
    lib.rs search restatement over that chunk's index for the full query batch.
    `check` (optional): dict with ordered tuples of another implementation to compare with."""
    from oracle import oracle as O
    kind = "reference" if O.reference_libsais_available() else "port"
    sa_fn = O.suffix_array_reference if kind == "reference" else O.suffix_array_port
    n = len(text0)
    t0 = time.perf_counter()
    sa = sa_fn(text0)                          # libsais(T, SA, n, 0, NULL): lib.rs:29-37, one thread
    build_s = time.perf_counter() - t0
    log("[reference] libsais on %d bytes: %.1fs (%.4f GB/s, 1 thread)" % (n, build_s, n / build_s / 1e9))
    out = {"kind": kind, "cores": 1, "build_s": build_s, "build_GBps": n / build_s / 1e9}
    tmp = None
    if path is None:
        tmp = tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        path = os.path.join(tmp.name, "chunk0.idx")
        write_container(path, text0, sa)
    reader = O.Reader(path)
    times = []
    for it in range(search_warmup + search_steps):
        t0 = time.perf_counter()
        counts, ch, st, en = reader.search_multiple_tuples(pats)      # lib.rs:201-287 per query, sequential
        if it >= search_warmup:
            times.append(time.perf_counter() - t0)
    out.update(search_s=float(np.mean(times)), search_qps=len(pats) / float(np.mean(times)), entries=int(len(ch)))
    log("[reference] search_multiple of %d queries over the 1-chunk index: %.3fs (%d entries)" % (len(pats), out["search_s"], len(ch)))
    out["single_query_us"] = {q: median_latency_us(lambda q=q: reader.search_tuples(q), 100)
                              for q in (b"google", b"text_two", b"zzzzzz")}
    out["single_query_results"] = {q.decode(): int(len(reader.search_tuples(q)[0])) for q in (b"google", b"text_two")}
    out["single_query_us"] = {k.decode(): v for k, v in out["single_query_us"].items()}
    if check is not None:
        out["sa_identical_to_gpu"] = bool(np.array_equal(sa, check["sa"]))
        out["search_identical_to_gpu"] = bool(np.array_equal(counts, check["counts"]) and np.array_equal(st, check["start"])
                                              and np.array_equal(en, check["end"]))
    out["sample"] = ("chunk 0 of the %s at full size (%d bytes): libsais timed once on it; the reference builds chunks "
                     "serially on one thread, so this is its GB/s on the whole index.  Search: the same %d queries over "
                     "that chunk's index (1 of the chunks; the reference searches chunks in parallel with rayon)"
                     % ("config-3 corpus", n, len(pats)))
    reader.close()
    if tmp:
        tmp.cleanup()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    gen_dev = "cuda" if torch.cuda.is_available() else "cpu"     # input synthesis only; the timed path is CPU code
    n = args.size
    t0 = time.perf_counter()
    text0 = synth.config3_chunk(0, n, device=gen_dev)
    pats = make_queries(n, args.chunks, args.queries, gen_dev)
    log("[reference] chunk 0 + queries generated on %s in %.1fs" % (gen_dev, time.perf_counter() - t0))
    r = cpu_reference_run(text0, pats, search_steps=max(1, min(args.steps, 5)), search_warmup=1)
    cfg = workload_config(n, len(pats), max(args.gpus, 1), args.chunks)
    cfg["reference_steps"] = ("libsais on one full-size chunk is timed ONCE per run (%.0f s per build; K timed full builds "
                              "would not fit a few minutes); search_multiple is timed over %d steps" % (r["build_s"], max(1, min(args.steps, 5))))
    line = {
        "impl": "reference", "metric": "index_build_GBps", "value": r["build_GBps"], "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": (r["build_s"] * args.chunks + r["search_s"]) * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8/i32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": r["build_GBps"], "unit": "GB/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["build_GBps"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "search": {"metric": "search_multiple_qps", "value": r["search_qps"], "unit": "queries/s", "entries_chunk0": r["entries"],
                   "e2e": {"value": r["search_qps"], "unit": "queries/s"},
                   "single_query_us": r["single_query_us"], "single_query_results": r["single_query_results"]},
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

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=300))
    torch.cuda.set_device(local_rank)
    pss.check(pss.lib.pss_set_device(local_rank))
    dev = torch.device("cuda", local_rank)
    lib = pss.lib
    n, n_chunks = args.size, args.chunks
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

    def exchange_id(raw):
        t = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())

    comm = pss.Comm(rank, world, exchange_id if world > 1 else None)

    # ---- inputs: chunk k of the index belongs to rank k % world ---------------------------------
    own = [k for k in range(n_chunks) if k % world == rank]
    t_gen = time.perf_counter()
    d_text, h_text = {}, {}
    for k in own:
        t = synth.config3_chunk_torch(k, n, device=dev)            # n + 16 bytes, zero padded
        if k == 0:                                                  # config 1's planted lines
            host = t[:n].cpu().numpy()
            synth.plant(host, "google", 5943, 20240502)
            synth.plant(host, "text_two", 159, 20240503)
            t[:n] = torch.from_numpy(host).to(dev)
        d_text[k] = t
        h = torch.empty(n, dtype=torch.uint8).pin_memory()
        h.copy_(t[:n])
        h_text[k] = h
    pats = make_queries(n, n_chunks, args.queries, dev)             # identical on every rank
    blob, offs = synth.pack_patterns(pats)
    torch.cuda.synchronize()
    log("[rank %d] generated %d x %d bytes + %d queries in %.1fs" % (rank, len(own), n, len(pats), time.perf_counter() - t_gen))
    d_sa = {k: torch.empty(n, dtype=torch.int32, device=dev) for k in own}
    h_sa = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(2)]
    builder = C.c_void_p()
    pss.check(lib.pss_sa_builder_create(local_rank, n, C.byref(builder)))
    pss.check(lib.pss_sa_builder_set_profiling(builder, 1))

    stats = pss.BuildStats()
    pstats = (pss.PassStat * 512)()

    def build_device(acc=None):
        """Device-resident BUILD of every chunk this rank owns (text in HBM → SA in HBM)."""
        for k in own:
            pss.check(lib.pss_sa_builder_build_device(builder, d_text[k].data_ptr(), n, d_sa[k].data_ptr(), None))
            if acc is not None:
                lib.pss_sa_builder_stats(builder, C.byref(stats), pstats)
                acc["dev_ms"] += stats.total_ms
                for i in range(stats.n_pass_stats):
                    acc["pass_bytes"] += 24.0 * pstats[i].n_records
                    acc["pass_ms"] += pstats[i].ms
                    acc["launches"] += 1

    build_device()
    # ---- the index for SEARCH: this rank's chunks, resident in HBM (no file: 37.5 GiB) -------------
    chunks = [pss.DeviceChunk(d_text[k].data_ptr(), d_sa[k].data_ptr(), h_text[k].data_ptr(), n, k) for k in own]
    reader = pss.Reader(device_chunks=chunks, n_chunks_total=n_chunks, device=local_rank)
    d_blob = torch.from_numpy(blob).to(dev) if rank == 0 else None
    d_offs = torch.from_numpy(offs).to(dev) if rank == 0 else None
    dres = pss.DeviceResult()

    def search_device():
        """Device-resident SEARCH: batch in rank 0's HBM → merged tuples in rank 0's HBM."""
        pss.check(lib.pss_reader_search_batch_dist_device(
            reader.h, comm.h, d_blob.data_ptr() if rank == 0 else None, d_offs.data_ptr() if rank == 0 else None,
            len(pats), int(offs[-1]), C.byref(dres)))
        return int(dres.n_entries), int(dres.n_hits), float(dres.ms_exchange)

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
    exch_ms = 0.0
    for _ in range(K):
        a = time.perf_counter()
        build_device(acc)
        b = time.perf_counter()
        # ranks own different numbers of chunks (15 over 8): align them here, so that the
        # wait for the busiest builder is not booked as search time by the idle ranks
        if world > 1:
            barrier()
        b2 = time.perf_counter()
        entries, hits_local, ex = search_device()
        c = time.perf_counter()
        build_t += b - a
        search_t += c - b2
        exch_ms += ex
    barrier()
    t_end = time.perf_counter()
    launches = lib.pss_kernel_launch_count() - launches0
    clocks = sampler.stop(t_begin, t_end)
    step_s = max_over_ranks((t_end - t_begin) / K)
    build_s = max_over_ranks(build_t / K)
    search_s = max_over_ranks(search_t / K)
    dev_build_s = max_over_ranks(acc["dev_ms"] / K / 1e3)
    pass_bytes, pass_ms, n_pass_launch = acc["pass_bytes"], acc["pass_ms"], acc["launches"]
    rounds, passes = stats.rounds, stats.n_passes
    active = [int(stats.active_per_round[i]) for i in range(stats.rounds + 1)]
    dev_stage = {"bounds": float(dres.ms_bounds), "extract": float(dres.ms_extract), "dedup": float(dres.ms_dedup),
                 "exchange": exch_ms / K}

    # ---- the same search with a 10x larger batch (the 10 000 queries ten times over): per-batch
    #      fixed costs (launches, the collectives' latency, scalar read-backs) stop hiding how the
    #      per-chunk work itself scales over the ranks ---------------------------------------------
    big = None
    if args.big_batch > 1:
        nbig = len(pats) * args.big_batch
        big_total = int(offs[-1]) * args.big_batch
        if rank == 0:
            lens = np.diff(offs)
            big_offs = np.zeros(nbig + 1, dtype=np.int64)
            np.cumsum(np.tile(lens, args.big_batch), out=big_offs[1:])
            d_big_blob = torch.from_numpy(np.tile(blob[:int(offs[-1])], args.big_batch)).to(dev)
            d_big_offs = torch.from_numpy(big_offs).to(dev)
        bres = pss.DeviceResult()

        def search_big():
            pss.check(lib.pss_reader_search_batch_dist_device(
                reader.h, comm.h, d_big_blob.data_ptr() if rank == 0 else None, d_big_offs.data_ptr() if rank == 0 else None,
                nbig, big_total, C.byref(bres)))
            return int(bres.n_entries)

        for _ in range(2):
            search_big()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            big_entries = search_big()
        barrier()
        big_s = max_over_ranks((time.perf_counter() - t0) / 3)
        big = {"queries": nbig, "value": nbig / big_s, "unit": "queries/s", "ms_per_batch": big_s * 1e3,
               "entries": big_entries, "what": "the %d queries x %d in one batch, device-resident, same exchange" % (len(pats), args.big_batch)}
        if rank == 0:
            del d_big_blob, d_big_offs

    # ---- end-to-end through the C-ABI host seams (pinned host buffers) ------------------------
    def e2e_build():
        """Host text → host suffix array for every owned chunk: all builds are queued at once
        (pss_sa_build_begin), the waits copy SA k out while chunk k+1 is being built."""
        handles = []
        for k in own:
            h = C.c_void_p()
            pss.check(lib.pss_sa_build_begin(local_rank, h_text[k].data_ptr(), n, C.byref(h)))
            handles.append(h)
        for i, h in enumerate(handles):
            pss.check(lib.pss_sa_build_wait(h, h_sa[i % 2].data_ptr()))

    def e2e_search():
        """The collective C-ABI host call: host patterns in (rank 0), host result tuples out on rank 0."""
        res = C.c_void_p()
        pss.check(lib.pss_reader_search_batch_dist(reader.h, comm.h, blob.ctypes.data, offs.ctypes.data, len(pats), C.byref(res)))
        r = C.cast(res, C.POINTER(pss.Result)).contents
        out = (int(r.n_entries), dict(ms_bounds=r.ms_bounds, ms_extract=r.ms_extract, ms_dedup=r.ms_dedup,
                                      ms_exchange=r.ms_exchange, ms_total=r.ms_total))
        lib.pss_result_free(res)
        return out

    Ke = max(1, min(K, args.e2e_steps))
    e2e_build()
    e2e_search()
    barrier()
    eb = es = 0.0
    for _ in range(Ke):
        a = time.perf_counter()
        e2e_build()
        b = time.perf_counter()
        if world > 1:
            barrier()
        b2 = time.perf_counter()
        n_e2e_entries, sstats = e2e_search()
        c = time.perf_counter()
        eb += b - a
        es += c - b2
        if world > 1:
            barrier()     # rank 0 still copies the merged result out: the other ranks' next build must not share its PCIe/host path
    barrier()
    e2e_build_s = max_over_ranks(eb / Ke)
    e2e_search_s = max_over_ranks(es / Ke)
    sa_last = own[-1]
    e2e_sa_matches = bool(torch.equal(h_sa[(len(own) - 1) % 2], d_sa[sa_last].cpu()))
    h2d_search = int(blob.nbytes + offs.nbytes)
    d2h_search = int(n_e2e_entries * 12 + (len(pats) + 1) * 8)
    pss.check(lib.pss_release_cached())          # the async engine's workspace is not needed any more

    # ---- rank 0 extras: Writer end to end, Python boundary, single-query latency, CPU baseline ----
    extras = {}
    if rank == 0:
        shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
        tmpdir = tempfile.TemporaryDirectory(dir=shm)
        where = "tmpfs (/dev/shm)" if shm else "the default temp directory"
        try:
            if not args.skip_writer:
                # Writer.add_entries_from_file_lines → finalize on the first two owned chunks' text
                src, dst = os.path.join(tmpdir.name, "input.txt"), os.path.join(tmpdir.name, "writer.idx")
                ks = own[:2]
                with open(src, "wb") as f:
                    for k in ks:
                        f.write(memoryview(h_text[k].numpy()))
                t0 = time.perf_counter()
                w = pss.Writer(dst, None, devices=[local_rank])
                pss.check(w.add_entries_from_file_lines(src))
                pss.check(w.finalize())
                pss.check(w.close())
                wall = time.perf_counter() - t0
                size = os.path.getsize(dst)
                sa0 = np.fromfile(dst, dtype=np.int32, count=n, offset=8 + n)
                extras["writer_e2e"] = {
                    "what": "Writer(index).add_entries_from_file_lines(%d-byte text file) → finalize → close; build of chunk k+1 "
                            "overlaps D2H + file write of chunk k; files on %s" % (len(ks) * n, where),
                    "input_bytes": len(ks) * n, "index_bytes": int(size), "seconds": wall,
                    "text_GBps": len(ks) * n / wall / 1e9, "index_write_GBps": size / wall / 1e9,
                    "chunks_written": len(ks), "sa_identical_to_device_build": bool(np.array_equal(sa0, d_sa[ks[0]].cpu().numpy())),
                }
                os.unlink(src)
                os.unlink(dst)
                del sa0
            path0 = os.path.join(tmpdir.name, "chunk0.idx")
            sa0 = d_sa[0].cpu().numpy()
            text0 = h_text[0].numpy()
            if not (args.skip_python and args.skip_cpu):
                write_container(path0, text0, sa0)
            one = pss.Reader(path0) if not (args.skip_python and args.skip_cpu) else None
            if not args.skip_python:
                import pysubstringsearch_b200
                str_pats = [p.decode("ascii") for p in pats]
                py_reader = pysubstringsearch_b200.Reader(index_file_path=path0)
                py_reader.search_multiple(substrings=str_pats[:100])
                t0 = time.perf_counter()
                res = py_reader.search_multiple(substrings=str_pats)
                py_s = time.perf_counter() - t0
                extras["python_boundary"] = {
                    "what": "pysubstringsearch.Reader over chunk 0's index file: search_multiple(10 000 str) → list[str]",
                    "qps": len(str_pats) / py_s, "strings_returned": len(res),
                    "single_query_us": {q: median_latency_us(lambda q=q: py_reader.search(substring=q))
                                        for q in ("google", "text_two", "zzzzzz")},
                    "single_query_results": {q: len(py_reader.search(substring=q)) for q in ("google", "text_two")},
                    "single_query_reference_readme_us": {"google": 497.0, "text_two": 14.9},
                }
                extras["cabi_single_query_us"] = {q: median_latency_us(lambda q=q: one.search_batch([q]))
                                                  for q in ("google", "text_two", "zzzzzz")}
                del res, py_reader
            if world == 1 and not args.skip_cpu:
                qo, ch, st, en, _ = one.search_batch(pats)
                check = {"sa": sa0, "counts": np.diff(qo), "start": st, "end": en}
                extras["cpu"] = cpu_reference_run(text0, pats, search_steps=1, search_warmup=0, path=path0, check=check)
            if one is not None:
                one.close()
        finally:
            tmpdir.cleanup()

    reader.close()
    comm.close()
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
    avg_alg_bytes = pass_bytes / n_pass_launch if n_pass_launch else 0.0
    total_bytes = n * n_chunks
    npairs_busiest = len(pats) * (-(-n_chunks // world))
    probe_bytes = 2 * int(np.ceil(np.log2(max(n, 2)))) * 64        # 2 searches x log2(n) probes x (SA sector + text sector)
    line = {
        "metric": "index_build_GBps", "value": total_bytes / build_s / 1e9, "unit": "GB/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": step_s * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8 text / u32 ranks / u64 keys", "data": "synthetic",
        "config": workload_config(n, len(pats), world, n_chunks),
        "build": {"device_event_ms_per_step": dev_build_s * 1e3, "host_call_ms_per_step": build_s * 1e3,
                  "per_chunk_ms": dev_build_s * 1e3 / max(1, -(-n_chunks // world)),
                  "per_chunk_GBps": n / (dev_build_s / max(1, -(-n_chunks // world))) / 1e9 if dev_build_s else None,
                  "rounds": rounds, "radix_passes": passes, "active_per_round": active, "h0_symbols": stats.h0,
                  "bits_per_symbol": stats.bits_per_symbol},
        "e2e": {"value": total_bytes / e2e_build_s / 1e9, "unit": "GB/s",
                "h2d_bytes_per_step": n * len(own), "d2h_bytes_per_step": 4 * n * len(own),
                "ms_per_step": e2e_build_s * 1e3, "steps": Ke, "same_sa_as_device_build": e2e_sa_matches,
                "api": "pss_sa_build_begin / pss_sa_build_wait, pinned host text and suffix array (bytes are rank 0's)"},
        "search": {
            "metric": "search_multiple_qps", "value": len(pats) / search_s, "unit": "queries/s",
            "ms_per_batch": search_s * 1e3, "entries": int(entries),
            "e2e": {"value": len(pats) / e2e_search_s, "unit": "queries/s", "ms_per_batch": e2e_search_s * 1e3,
                    "h2d_bytes_per_step": h2d_search, "d2h_bytes_per_step": d2h_search, "entries": int(n_e2e_entries),
                    "steps": Ke, "api": "pss_reader_search_batch_dist (host patterns → host tuples on rank 0)"},
            "stage_ms_rank0": dev_stage,
            "big_batch": big,
            "e2e_stage_ms_rank0": sstats,
            "roofline": {"bound": "hbm-random", "unit": "GB/s",
                         "algorithmic_bytes": "2 x ceil(log2 n) probes x (one 32 B SA sector + one 32 B text sector) per (query, chunk)",
                         "achieved": npairs_busiest * probe_bytes / (dev_stage["bounds"] * 1e-3) / 1e9 if dev_stage["bounds"] else None,
                         "kernel": "bounds_group_kernel (4 or 8 lanes per pair from 16 384 pairs per rank on) / bounds_kernel (a warp per pair)",
                         "peak": peak,
                         "note": "random 32-byte sectors: a fraction of the streaming peak is the ceiling; "
                                 "profiles/r02_search_bounds_ab.txt has the geometry sweep",
                         "ncu": "profiles/r02_ncu_search_kernels.txt (sectors per request, DRAM bytes)"},
        },
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                     "traffic": None,
                     "traffic_note": "not measured in this run: dram__bytes of one `ncu --set full` launch is in profiles/ "
                                     "(r02_ncu_pass_kernel.txt; r01: 1.02x algorithmic)",
                     "algorithmic_GB_per_launch": avg_alg_bytes / 1e9,
                     "kernel": "onesweep_pass_kernel", "peak_source": peak_kind,
                     "algorithmic_bytes": "24 B per record per pass (8 B key + 4 B value, read once + written once); "
                                          "the event pair includes the reset of the pass's look-back words",
                     "launches_timed": n_pass_launch, "avg_launch_ms": pass_ms / n_pass_launch if n_pass_launch else None,
                     "share_of_build": (pass_ms / K) / (acc["dev_ms"] / K) if acc["dev_ms"] else None},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    line.update({k: v for k, v in extras.items() if k != "cpu"})
    cpu = extras.get("cpu")
    if cpu:
        line["cpu_baseline"] = {"value": cpu["build_GBps"], "unit": "GB/s", "cores": cpu["cores"], "kind": cpu["kind"],
                                "sample": cpu["sample"], "build_s_one_chunk": cpu["build_s"], "search_qps": cpu["search_qps"],
                                "search_entries_chunk0": cpu["entries"], "single_query_us": cpu["single_query_us"],
                                "sa_identical_to_gpu": cpu.get("sa_identical_to_gpu"),
                                "search_identical_to_gpu": cpu.get("search_identical_to_gpu")}
    emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=synth.CONFIG3_CHUNK_BYTES, help="bytes per chunk (default 2^29)")
    ap.add_argument("--chunks", type=int, default=synth.CONFIG3_CHUNKS, help="chunks of the index (default 15), chunk k -> rank k %% N")
    ap.add_argument("--queries", type=int, default=N_QUERIES)
    ap.add_argument("--e2e-steps", type=int, default=5, help="timed steps of the host-to-host legs (<= --steps)")
    ap.add_argument("--big-batch", type=int, default=10, help="extra search measurement with the query list repeated this many times in one batch (<= 1: off)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-python", action="store_true")
    ap.add_argument("--skip-writer", action="store_true")
    args = ap.parse_args()
    capture_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
