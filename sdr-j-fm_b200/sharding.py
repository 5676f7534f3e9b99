"""Multi-GPU host logic: streams are the only thing that shards.

One process per GPU; rank r owns a contiguous block of streams; the data path has no
collective.  The single collective is the one-time broadcast of the table blob (taps and
LUTs) that rank 0 designs — `broadcast_tables` works with any torch.distributed backend
(NCCL on the GPU box, gloo in the CPU tests)."""
import numpy as np


def stream_range(n_total, world, rank):
    """contiguous block of streams of `rank` (first `n_total % world` ranks get one more)."""
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_tables(blob, dist, device=None, src=0):
    """blob: uint8 ndarray (rank `src`: the designed tables; others: anything of the same size or
    None).  Returns the blob of rank `src` on every rank."""
    import torch
    n = torch.tensor([0 if blob is None else int(blob.size)], dtype=torch.int64, device=device)
    dist.broadcast(n, src=src)
    if dist.get_rank() == src:
        t = torch.from_numpy(np.ascontiguousarray(blob, dtype=np.uint8)).to(device)
    else:
        t = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    dist.broadcast(t, src=src)
    return t.cpu().numpy()
