"""N > 1 host logic on CPU: two gloo ranks each answer for the chunks they own (the oracle
plays the per-rank searcher here), hits are gathered to rank 0 and merged; the result must
equal the single-process search."""
import os
import subprocess
import sys
import tempfile

import numpy as np

from tests.conftest import ROOT

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["PSS_ROOT"])
from oracle import oracle as O
from pysubstringsearch_b200 import distributed as D
from tools import synth

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
path = os.environ["PSS_INDEX"]
reader = O.Reader(path)
pats = [b"e ", b"ab", b"", b"zzzz", b"\n", b"qu", b"the"]
blob, offs = synth.pack_patterns(pats)
tb, to = D.broadcast_queries(torch.from_numpy(blob) if rank == 0 else None,
                             torch.from_numpy(offs) if rank == 0 else None, torch.device("cpu"))
blob_r, offs_r = tb.numpy(), to.numpy()
q_l, c_l, s_l, e_l = [], [], [], []
for qi in range(len(offs_r) - 1):
    pat = bytes(blob_r[offs_r[qi]:offs_r[qi + 1]])
    ch, st, en = reader.search_tuples(pat)
    mine = np.array([D.chunk_owner(int(c), world) == rank for c in ch], dtype=bool)
    q_l.append(np.full(int(mine.sum()), qi, dtype=np.int32)); c_l.append(ch[mine]); s_l.append(st[mine]); e_l.append(en[mine])
cat = lambda xs, dt: torch.from_numpy(np.concatenate(xs).astype(dt).view(np.int32))
parts = D.gather_hits(cat(q_l, np.int32), cat(c_l, np.int32), cat(s_l, np.uint32), cat(e_l, np.uint32))
if rank == 0:
    merged = D.merge_hits(parts)
    counts, ch, st, en = reader.search_multiple_tuples(pats)
    q = np.repeat(np.arange(len(pats)), counts)
    ok = (np.array_equal(merged[0], q) and np.array_equal(merged[1], ch)
          and np.array_equal(merged[2].view(np.uint32), st) and np.array_equal(merged[3].view(np.uint32), en))
    print("MERGE_OK" if ok else "MERGE_BAD", merged.shape[1])
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_gloo_gather_matches_single_process(oracle):
    from tools import synth
    text = synth.zipf_words_text(200_000, seed=9, vocab=512, block=1 << 14)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "m.idx")
        w = oracle.Writer(path, 1 << 15)            # ~7 chunks
        for e in bytes(text).split(b"\n")[:-1]:
            w.add_entry(e)
        w.close()
        script = os.path.join(d, "worker.py")
        open(script, "w").write(WORKER)
        env = dict(os.environ, PSS_ROOT=ROOT, PSS_INDEX=path, MASTER_ADDR="127.0.0.1")
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                              "--master-addr", "127.0.0.1", "--master-port", "29533", script],
                             env=env, capture_output=True, text=True, timeout=300)
        assert "MERGE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_chunk_owner_round_robin():
    from pysubstringsearch_b200 import distributed as D
    assert [D.chunk_owner(k, 4) for k in range(8)] == [0, 1, 2, 3, 0, 1, 2, 3]
    owned = [[k for k in range(15) if D.chunk_owner(k, 8) == r] for r in range(8)]
    assert sorted(sum(owned, [])) == list(range(15)) and max(len(o) for o in owned) == 2   # ceil(15/8)
