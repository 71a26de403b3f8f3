"""N > 1 host logic on CPU.  The exchange step itself runs inside libpss_b200.so over NCCL and
is tested on GPUs (tests/test_gpu_parity.py::test_two_gpu_distributed_search_matches_single_process);
here two gloo ranks exercise the same protocol with the oracle as the per-rank searcher: each
rank produces what a GPU rank produces (entry offsets per (query, local chunk) pair + tuples in
pair order), rank 0 places them with the numpy restatement of the library's placement
arithmetic, and the result must equal the single-process search."""
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

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
reader = O.Reader(os.environ["PSS_INDEX"])
n_total = reader.num_chunks
pats = [[b"e ", b"ab", b"", b"zzzz", b"\n", b"qu", b"the"]] if rank == 0 else [None]
dist.broadcast_object_list(pats, 0)                      # step 1: the batch reaches every rank
pats = pats[0]
mine = D.owned_chunks(rank, world, n_total)
eo, st_l, en_l = [0], [], []
for pat in pats:                                          # step 2: local search, pair order
    ch, st, en = reader.search_tuples(pat)
    for k in mine:
        sel = ch == k
        st_l.append(st[sel]); en_l.append(en[sel])
        eo.append(eo[-1] + int(sel.sum()))
part = (np.array(eo, dtype=np.uint32), np.concatenate(st_l) if st_l else np.zeros(0, np.uint32),
        np.concatenate(en_l) if en_l else np.zeros(0, np.uint32))
parts = [None] * world
dist.gather_object(part, parts if rank == 0 else None, 0)   # step 3: gather-v to rank 0
if rank == 0:
    qoff, ch, st, en = D.merge_reference(parts, len(pats), n_total)   # step 4: placement
    counts, och, ost, oen = reader.search_multiple_tuples(pats)
    ok = (np.array_equal(np.diff(qoff), counts) and np.array_equal(ch, och)
          and np.array_equal(st, ost) and np.array_equal(en, oen))
    print("MERGE_OK" if ok else "MERGE_BAD", len(ch))
dist.barrier()
dist.destroy_process_group()
'''


def _run(world, port, oracle):
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
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
                              "--master-addr", "127.0.0.1", "--master-port", str(port), script],
                             env=env, capture_output=True, text=True, timeout=300)
        assert "MERGE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_two_rank_gloo_exchange_matches_single_process(oracle):
    _run(2, 29533, oracle)


def test_three_rank_gloo_uneven_shards(oracle):
    _run(3, 29534, oracle)          # 7 chunks over 3 ranks: 3 + 2 + 2


def test_chunk_map_and_placement_arithmetic():
    from pysubstringsearch_b200 import distributed as D
    assert [D.chunk_owner(k, 4) for k in range(8)] == [0, 1, 2, 3, 0, 1, 2, 3]
    owned = [D.owned_chunks(r, 8, 15) for r in range(8)]
    assert sorted(sum(owned, [])) == list(range(15)) and max(len(o) for o in owned) == 2   # ceil(15/8)
    for world in (1, 2, 3, 8, 20):
        for k in range(15):
            for r in range(world):
                assert D.chunks_before(r, k, world) == sum(1 for c in D.owned_chunks(r, world, 15) if c < k)
    # random per-pair counts: the placed offsets are the exclusive scan in (query, chunk) order
    rng = np.random.default_rng(1)
    nq, n_total, world = 5, 7, 3
    counts = rng.integers(0, 4, size=(nq, n_total))
    eo = []
    for r in range(world):
        c = counts[:, D.owned_chunks(r, world, n_total)].reshape(-1)
        eo.append(np.concatenate(([0], np.cumsum(c))).astype(np.uint32))
    final, qoff = D.place_reference(eo, nq, n_total)
    flat = np.concatenate(([0], np.cumsum(counts.reshape(-1))))
    for r in range(world):
        own = D.owned_chunks(r, world, n_total)
        for p in range(nq * len(own)):
            q, j = divmod(p, len(own))
            assert final[r][p] == flat[q * n_total + own[j]]
    assert np.array_equal(qoff, flat[::n_total])
