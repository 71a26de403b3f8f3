"""One-process-per-GPU plumbing for sharded indexes (torch.distributed; NCCL on GPUs, gloo
on CPU for tests).

Chunks are the shard unit (the reference already treats them as the parallel
unit at search time, lib.rs:207): chunk k is owned by rank k % world.  BUILD needs no
collective.  SEARCH has one exchange step per batch: the packed query batch is broadcast
from rank 0, every rank searches its own chunks, and the per-chunk hit tuples are gathered
to rank 0, which orders them by (query, chunk) — the order a single-process Reader returns.
"""
import numpy as np
import torch
import torch.distributed as dist


def chunk_owner(chunk, world):
    return chunk % world


def broadcast_queries(blob, offsets, device, src=0):
    """blob: uint8 tensor, offsets: int64 tensor (valid on `src`; other ranks may pass None).
    Returns (blob, offsets) tensors on `device` on every rank."""
    rank = dist.get_rank()
    meta = torch.zeros(2, dtype=torch.int64, device=device)
    if rank == src:
        meta[0], meta[1] = blob.numel(), offsets.numel()
    dist.broadcast(meta, src)
    nb, no = int(meta[0]), int(meta[1])
    if rank == src:
        blob, offsets = blob.to(device).contiguous(), offsets.to(device).contiguous()
    else:
        blob = torch.empty(nb, dtype=torch.uint8, device=device)
        offsets = torch.empty(no, dtype=torch.int64, device=device)
    dist.broadcast(blob, src)
    dist.broadcast(offsets, src)
    return blob, offsets


_BUFFERS = {}


def _buffer(key, shape, dtype, device):
    buf = _BUFFERS.get(key)
    if buf is None or buf.shape != torch.Size(shape) or buf.device != device:
        buf = torch.empty(shape, dtype=dtype, device=device)
        _BUFFERS[key] = buf
    return buf


def gather_hits(query, chunk, start, end, dst=0):
    """Variable-length gather of hit tuples to `dst`.  Inputs: 1-D int32 tensors of equal
    length on this rank's device (start/end carry uint32 bit patterns).  Returns on `dst` a
    list with one (4, k_r) int32 tensor per rank (views into a reused buffer: consume them
    before the next call), elsewhere None.

    Two collectives per batch: an all-gather of the per-rank counts (so every rank agrees
    on the padded width) and one gather of the padded (4, kmax) blocks.  The staging
    buffers persist across calls and only grow."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = query.device
    n = query.numel()
    counts_t = _buffer(("counts", world), (world,), torch.int64, dev)
    mine = _buffer(("mine",), (1,), torch.int64, dev)
    mine.fill_(n)
    dist.all_gather_into_tensor(counts_t, mine)
    counts = counts_t.tolist()
    kmax = max(max(counts), 1)
    cap = 1 << (kmax - 1).bit_length()          # power-of-two widths → few reallocations
    send = _buffer(("send",), (4, cap), torch.int32, dev)
    send[0, :n], send[1, :n], send[2, :n], send[3, :n] = query, chunk, start, end
    if rank == dst:
        recv = _buffer(("recv",), (world, 4, cap), torch.int32, dev)
        dist.gather(send, list(recv.unbind(0)), dst=dst)
        return [recv[r, :, :c] for r, c in enumerate(counts)]
    dist.gather(send, None, dst=dst)
    return None


def merge_hits(parts):
    """Per-rank (4, k) tensors → one (4, K) int32 numpy array ordered by (query, chunk),
    keeping each rank's own order inside a (query, chunk) pair (SA order of first hit)."""
    cols = [p.cpu().numpy() for p in parts if p.shape[1]]
    if not cols:
        return np.zeros((4, 0), dtype=np.int32)
    allc = np.concatenate(cols, axis=1)
    seq = np.arange(allc.shape[1])
    order = np.lexsort((seq, allc[1], allc[0]))
    return allc[:, order]
