"""CPU tests of the measurement plumbing: the reference arm of bench.py (which needs no GPU)
prints exactly one JSON line with the contract's keys, and the synthetic generators are
deterministic."""
import json
import os
import subprocess
import sys

import numpy as np

from tests.conftest import ROOT
from tools import synth


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--queries", "300", "--size", str(1 << 21), "--chunks", "3"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "index_build_GBps" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    assert "workload" in d["config"] and d["search"]["unit"] == "queries/s" and d["value"] > 0
    assert d["scaling"] == "strong" and d["config"]["chunks"] == 3 and "selective" in d["config"]["workload"]
    assert set(d["search"]["single_query_us"]) == {"google", "text_two", "zzzzzz"}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_synthetic_generators_are_deterministic():
    a, b = synth.config1_text(300_000), synth.config1_text(300_000)
    assert np.array_equal(a, b) and a[-1] == 10 and len(a) == 300_000
    assert set(np.unique(a).tolist()) <= set(range(97, 123)) | {10, 32, 95}
    qa, qb = synth.config2_queries(a, nq=200, seed=7), synth.config2_queries(b, nq=200, seed=7)
    assert qa == qb and len(qa) == 200 and all(4 <= len(q) <= 32 for q in qa)
    blob, offs = synth.pack_patterns(qa)
    assert offs[0] == 0 and offs[-1] == sum(len(q) for q in qa)
    assert bytes(blob[offs[3]:offs[4]]) == qa[3]
    t = synth.acgt_text(100_000)
    assert set(np.unique(t).tolist()) <= {10, 65, 67, 71, 84} and t[-1] == 10


def test_config3_corpus_is_reproducible_from_prefixes():
    """Chunks of the config-3 corpus share one vocabulary, differ by seed, and a short prefix of a
    chunk equals the head of the full chunk — what lets every process cut the same query batch."""
    a = synth.config3_chunk_torch(2, 300_000)
    b = synth.config3_chunk_torch(2, 100_000, force_newline=False)
    assert a.numel() == 300_016 and int(a[299_999]) == 10 and not a[300_000:].any()
    assert bool((a[:100_000] == b[:100_000]).all())
    c = synth.config3_chunk_torch(3, 100_000, force_newline=False)
    assert not bool((c[:100_000] == b[:100_000]).all())
    words = lambda t: set(bytes(t[:100_000].numpy()).replace(b"\n", b" ").split()[1:-1])
    assert len(words(b) & words(c)) > 100                       # shared vocabulary
    pre = [synth.config3_chunk_torch(k, 1 << 20, force_newline=False)[:1 << 20] for k in range(3)]
    q1 = synth.config3_queries(pre, nq=60, chunk_bytes=1 << 20)
    q2 = synth.config3_queries(pre, nq=60, chunk_bytes=1 << 20)
    assert q1 == q2 and len(q1) == 60 and all(4 <= len(q) <= 32 for q in q1)
    hits = [sum(bytes(p.numpy()).count(q) > 0 for p in pre) for q in q1]
    assert sum(h > 0 for h in hits) >= 50                       # 90 % are cut from the text
    t0 = synth.config3_chunk(0, 400_000)
    assert bytes(t0).count(b"text_two") == 159 or bytes(t0).count(b"text_two") > 100


def test_selective_queries_have_bounded_hits(oracle):
    """config-2 queries are rejection-sampled so that their rarest 4-gram occurs <= max_count
    times, which bounds their own hit count."""
    text = synth.config1_text(2_000_000)
    pats = synth.config2_queries(text, nq=100, seed=3, max_count=500)
    raw = bytes(text)
    for p in pats[:40]:
        assert raw.count(p) <= 500
