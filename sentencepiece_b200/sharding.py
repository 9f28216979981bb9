"""Multi-GPU sharding of a sentence batch.

Sentences are independent, so the path shards with no data-path collective: contiguous
sentence ranges balanced by INPUT BYTES, one range per rank, model tables replicated.
The only exchange is the optional gather of the packed id buffers to one rank
(variable length: counts first, then padded buffers), done with torch.distributed --
NCCL over NVLink on GPUs, gloo in the CPU tests.  Rank-major order == file order.
"""
import numpy as np


def shard_ranges(offsets, world):
    """offsets: uint64[n+1] -> list of (lo, hi) sentence ranges, one per rank, contiguous,
    covering [0, n), balanced by bytes."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    base = int(offsets[0])
    total = int(offsets[n]) - base
    cuts = [0]
    for r in range(1, world):
        target = base + (total * r) // world
        k = int(np.searchsorted(offsets, np.uint64(target), side="left"))
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_ids(ids, id_offsets, dst=0, group=None):
    """ids: int32 tensor [m], id_offsets: int64 tensor [k+1] (this rank's shard, local offsets).
    Returns (ids, id_offsets) of the whole batch on rank `dst` (None elsewhere)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = ids.device
    meta = torch.tensor([ids.numel(), id_offsets.numel() - 1], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    max_ids = max(int(m[0]) for m in metas)
    max_sent = max(int(m[1]) for m in metas)
    send_ids = torch.zeros(max(max_ids, 1), dtype=torch.int32, device=dev)
    send_ids[: ids.numel()] = ids
    counts = (id_offsets[1:] - id_offsets[:-1]).to(torch.int64)
    send_cnt = torch.zeros(max(max_sent, 1), dtype=torch.int64, device=dev)
    send_cnt[: counts.numel()] = counts
    recv_ids = [torch.empty_like(send_ids) for _ in range(world)] if rank == dst else None
    recv_cnt = [torch.empty_like(send_cnt) for _ in range(world)] if rank == dst else None
    dist.gather(send_ids, recv_ids, dst=dst, group=group)
    dist.gather(send_cnt, recv_cnt, dst=dst, group=group)
    if rank != dst:
        return None
    all_ids = torch.cat([recv_ids[r][: int(metas[r][0])] for r in range(world)])
    all_cnt = torch.cat([recv_cnt[r][: int(metas[r][1])] for r in range(world)])
    offs = torch.zeros(all_cnt.numel() + 1, dtype=torch.int64, device=dev)
    offs[1:] = torch.cumsum(all_cnt, 0)
    return all_ids, offs
