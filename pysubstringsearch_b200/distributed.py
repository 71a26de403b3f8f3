"""One-process-per-GPU plumbing for sharded indexes.

Chunks are the shard unit (the reference already treats them as the parallel
unit at search time, lib.rs:207): chunk k is owned by rank k % world.  BUILD needs no
collective.  SEARCH has one exchange step per batch, and it lives INSIDE libpss_b200.so
(csrc/dist.cu, NCCL over NVLink): broadcast of the batch from rank 0 → local search →
gather-v of exactly the tuples found → placement into the single-process order.  This module
only bootstraps the library's communicator from an existing torch.distributed process group
(`init_comm`) and restates the placement arithmetic in numpy (`place_reference`) so that the
host-side logic can be tested on CPU with gloo ranks.
"""
import numpy as np


def chunk_owner(chunk, world):
    return chunk % world


def owned_chunks(rank, world, n_total):
    """Global ids of the chunks rank `rank` holds, ascending (local chunk j ↔ global j * world + rank)."""
    return list(range(rank, n_total, world))


def init_comm():
    """pss Comm over the current torch.distributed group: rank 0's NCCL id reaches the other
    ranks through a torch broadcast (works on the nccl and the gloo backend)."""
    import torch
    import torch.distributed as dist
    from . import capi
    rank, world = dist.get_rank(), dist.get_world_size()
    on_gpu = dist.get_backend() == "nccl"

    def exchange(raw):
        t = torch.tensor(list(raw), dtype=torch.uint8, device="cuda" if on_gpu else "cpu")
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())

    return capi.Comm(rank, world, exchange if world > 1 else None)


def chunks_before(rank, k, world):
    """Chunks owned by `rank` whose global id is below k (csrc/dist.cu: chunks_before)."""
    return (k - rank - 1) // world + 1 if k > rank else 0


def place_reference(entry_offsets, nq, n_total):
    """numpy restatement of the placement step of csrc/dist.cu.

    entry_offsets[r]: uint array [nq * nc_r + 1] — entries of rank r before its pair
    p = q * nc_r + j (j = local chunk).  Returns (final, query_offsets): final[r][p] is the
    position in the merged result of the first entry of rank r's pair p, i.e. the number of
    entries of every rank in pairs that precede (query q, chunk k = j * world + r) in
    (query, chunk) order; query_offsets[q] = entries before query q."""
    world = len(entry_offsets)
    nc = [len(owned_chunks(r, world, n_total)) for r in range(world)]
    final = []
    for r in range(world):
        f = np.zeros(nq * nc[r], dtype=np.int64)
        for p in range(nq * nc[r]):
            q, j = divmod(p, nc[r])
            k = j * world + r
            f[p] = sum(int(entry_offsets[r2][q * nc[r2] + chunks_before(r2, k, world)]) for r2 in range(world) if nc[r2])
        final.append(f)
    qoff = np.array([sum(int(entry_offsets[r][q * nc[r]]) for r in range(world)) for q in range(nq + 1)], dtype=np.int64)
    return final, qoff


def merge_reference(parts, nq, n_total):
    """parts[r] = (entry_offsets, start, end) of rank r → (query_offsets, chunk, start, end) in
    the single-process order, placed exactly as the library's kernel places them."""
    world = len(parts)
    final, qoff = place_reference([p[0] for p in parts], nq, n_total)
    total = int(qoff[-1])
    chunk = np.zeros(total, dtype=np.int32)
    start = np.zeros(total, dtype=np.uint32)
    end = np.zeros(total, dtype=np.uint32)
    for r, (eo, st, en) in enumerate(parts):
        nc = len(owned_chunks(r, world, n_total))
        for p in range(nq * nc):
            a, b = int(eo[p]), int(eo[p + 1])
            if b > a:
                d = int(final[r][p])
                chunk[d:d + b - a] = (p % nc) * world + r
                start[d:d + b - a] = st[a:b]
                end[d:d + b - a] = en[a:b]
    return qoff, chunk, start, end
