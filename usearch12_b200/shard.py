"""Query sharding across ranks and the final hit gather.

usearch_global shards naturally: queries are independent (search.cpp:63-86), the index is
replicated on every GPU, and the only exchange is the gather of the fixed-width hit records onto
rank 0.  torch.distributed is the plumbing (NCCL on GPUs, gloo in the CPU tests); the records are
opaque bytes here.
"""
import numpy as np

from .capi import HIT_DTYPE


def shard_range(n_items, rank, world):
    """Contiguous [lo, hi) slice of n_items owned by rank (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_hits(local_hits, query_base, rank, world, device=None):
    """Gathers every rank's hit records on rank 0.

    local_hits: numpy structured array (HIT_DTYPE) or a uint8 torch tensor holding packed records
    (already on `device` for the NCCL path); query_base: global index of this rank's first query.
    Returns on rank 0 one HIT_DTYPE array with global query indexes, ordered by rank then local
    order (i.e. by global query); None on the other ranks."""
    import torch
    import torch.distributed as dist
    if isinstance(local_hits, np.ndarray):
        buf = torch.from_numpy(local_hits.view(np.uint8).reshape(-1).copy())
        if device is not None:
            buf = buf.to(device)
    else:
        buf = local_hits
    meta = torch.tensor([buf.numel(), query_base], dtype=torch.int64, device=buf.device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    sizes = [int(m[0]) for m in metas]
    cap = max(sizes + [HIT_DTYPE.itemsize])
    padded = torch.zeros(cap, dtype=torch.uint8, device=buf.device)
    padded[:buf.numel()] = buf
    outs = [torch.zeros_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, outs, dst=0)
    if rank != 0:
        return None
    parts = []
    for r in range(world):
        a = outs[r][:sizes[r]].cpu().numpy().view(HIT_DTYPE).copy()
        a["query"] += np.uint32(int(metas[r][1]))
        parts.append(a)
    return np.concatenate(parts) if parts else np.zeros(0, HIT_DTYPE)
