"""LociPartitioning across the GPUs of one box.

The path shards naturally: a locus's result depends only on the reads overlapping it
(DistributedUtil.scala:537-545; T/DistributedUtilSuite.scala:208-220 asserts 800 tasks == 1 task).  So:
  1. `partition_loci_uniformly(world, loci)` gives every rank contiguous contig ranges (DistributedUtil.scala:83-108);
  2. each rank keeps the reads that overlap its ranges — a read crossing a boundary is duplicated, like the reference's
     read -> task expansion (DistributedUtil.scala:585-597);
  3. every rank runs the same kernels on its shard: no data-path collective;
  4. the per-rank record buffers (variable length) are gathered to rank 0 — the path's only exchange.
`torch.distributed` is the plumbing (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .reads import ReadBatch


def ranges_of_rank(partitions: Sequence[tuple], rank: int) -> List[tuple]:
    return [p for p in partitions if p[3] == rank]


def shard_reads(batch: ReadBatch, ranges: Sequence[tuple]) -> ReadBatch:
    """Reads of `batch` that overlap any of `ranges` (contig, start, end[, task]); order preserved."""
    keep = np.zeros(len(batch), dtype=bool)
    end = batch.end()
    for r in ranges:
        keep |= (batch.contig == r[0]) & (batch.start < r[2]) & (end > r[1])
    return batch.select(keep)


def gather_records(records: np.ndarray, pool: bytes, dst: int = 0, group=None, device=None):
    """Gathers one structured record array + its allele byte pool per rank to `dst`.
    Returns (records, pool) concatenated in rank order with ref/alt offsets re-based (None on other ranks)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = device or torch.device("cpu")
    payloads = [records.view(np.uint8).reshape(-1), np.frombuffer(pool, dtype=np.uint8)]
    sizes = torch.tensor([p.size for p in payloads], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    caps = torch.stack(all_sizes).max(dim=0).values.tolist()
    gathered = []
    for k, p in enumerate(payloads):
        cap = max(int(caps[k]), 1)
        buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
        if p.size:
            buf[:p.size] = torch.from_numpy(p.copy()).to(dev)
        out = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == dst else None
        dist.gather(buf, out, dst=dst, group=group)
        gathered.append(out)
    if rank != dst:
        return None, None
    recs, pools, base = [], [], 0
    for r in range(world):
        n_rec_bytes, n_pool = int(all_sizes[r][0]), int(all_sizes[r][1])
        a = gathered[0][r][:n_rec_bytes].cpu().numpy().view(records.dtype).copy()
        if base + n_pool >= 2 ** 32:
            raise OverflowError("merged allele pools exceed the 32-bit offsets of the records: gather fewer ranks at a time")
        if len(a):
            a["ref_off"] += base
            a["alt_off"] += base
        recs.append(a)
        pools.append(gathered[1][r][:n_pool].cpu().numpy().tobytes())
        base += n_pool
    return np.concatenate(recs) if recs else records[:0], b"".join(pools)
