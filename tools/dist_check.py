"""Distributed search check, run under torchrun with N ranks (N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py

Rank 0 writes a multi-chunk index; every rank opens its shard (chunk k → rank k % N) and all
ranks call the collective pss_reader_search_batch_dist; rank 0 compares the merged result with
the single-process Reader's and with the CPU oracle's (ordered tuples, bit-exact)."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from pysubstringsearch_b200 import capi as pss
    from tools import synth

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pss.check(pss.lib.pss_set_device(local))

    def exchange(raw):
        t = torch.tensor(list(raw), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())

    comm = pss.Comm(rank, world, exchange)
    text = synth.zipf_words_text(5_000_000, seed=77, vocab=4096, block=1 << 16)
    entries = bytes(text).split(b"\n")[:-1]
    path_t = [None]
    tmp = None
    if rank == 0:
        tmp = tempfile.TemporaryDirectory()
        path_t[0] = os.path.join(tmp.name, "dist.idx")
        w = pss.Writer(path_t[0], 700_000)        # 8 chunks: uneven over 3 ranks, 1+ per rank up to 8
        for e in entries:
            assert w.add_entry(e) == 0
        assert w.finalize() == 0
        w.close()
    dist.broadcast_object_list(path_t, 0)
    path = path_t[0]
    r = pss.Reader(path, shard=(rank, world))
    batches = [synth.config2_queries(text, nq=500, seed=4) + [b"", b"e ", b"\n", b"zzzzzzzz"],
               [b"google"], [], [b"qqqqqqqqqqqq", b"xxxxxxxxxxxxxxx"], synth.config2_queries(text, nq=37, seed=5)]
    full = o = None
    if rank == 0:
        from oracle import oracle as O
        full = pss.Reader(path)
        o = O.Reader(path)
    for pats in batches * 2:
        qo, ch, st, en, stats = r.search_batch_dist(comm, pats)
        if rank == 0:
            assert stats["n_ranks"] == world
            fqo, fch, fst, fen, _ = full.search_batch(pats)
            assert np.array_equal(qo, fqo) and np.array_equal(ch, fch) and np.array_equal(st, fst) and np.array_equal(en, fen)
            if pats:
                counts, och, ost, oen = o.search_multiple_tuples(pats)
                assert np.array_equal(np.diff(qo), counts) and np.array_equal(ch, och)
                assert np.array_equal(st, ost) and np.array_equal(en, oen)
        else:
            assert len(ch) == 0
    r.close()
    comm.close()
    dist.barrier()
    if rank == 0:
        full.close()
        print("dist_check ok: %d ranks, %d chunks, %d batches" % (world, full and len(batches) * 2 or 0, len(batches) * 2))
        tmp.cleanup()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
