"""The N > 1 path on real GPUs (-m gpu; skipped on a one-GPU box): loci partitioned over ranks, every rank generating and
packing the reads of its own shard on its device, records and depth histograms gathered to rank 0 by the library's NCCL
exchange (guac_result_gather / guac_comm_reduce_depth_histogram) — against one run over everything."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(sys.argv[0])) + "/..")
import torch, torch.distributed as dist
from guacamole_b200 import abi, callers, synth
from guacamole_b200.distributed import ranges_of_rank
from guacamole_b200.loci import partition_loci_uniformly

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo")
contigs = [("1", 400000), ("2", 250000), ("3", 1000)]
loci = [(0, 0, 399999), (1, 0, 249999), (2, 0, 999)]
ctx = callers.Context(rank)
ids = [callers.Comm.unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
comm = callers.Comm(ctx, ids[0], rank, world)
mine = ranges_of_rank(partition_loci_uniformly(world, loci), rank)   # contiguous ranges; contig 1 is cut mid-contig
shard = synth.generate_device(ctx, contigs, depth=30, seed=77, windows=synth.shard_windows(mine))
reads = ctx.pack_device(shard.c, [c[0] for c in contigs])
for sort in (1, 0):
    ctx.set_option(abi.OPT_SORT_RECORDS, sort)
    local = callers.germline_threshold(ctx, reads, mine)
    merged = comm.gather(local, 0)
    callers.depth_histogram(ctx, reads, mine, fetch=False)
    hist = comm.reduce_depth_histogram(0)
    if rank == 0:
        whole = synth.generate_device(ctx, contigs, depth=30, seed=77)
        all_reads = ctx.pack_device(whole.c, [c[0] for c in contigs])
        want = callers.germline_threshold(ctx, all_reads, loci)
        key = lambda g: (g["contig"], g["start"], g["ref"], g["alt"], g["gt"], g["tie"])
        got, exp = [key(g) for g in merged.genotypes()], [key(g) for g in want.genotypes()]
        assert len(got) == len(exp) > 1500, (len(got), len(exp))
        assert (got == exp) if sort else (sorted(got) == sorted(exp))
        assert merged.stats["loci_visited"] == want.stats["loci_visited"]
        assert np.array_equal(hist, callers.depth_histogram(ctx, all_reads, loci))
        assert len(merged.compact()[1]) == len(want.compact()[1]) > 20   # general (indel) records crossed with their allele bytes
        all_reads.free(); whole.free()
    else:
        assert len(merged) == 0 and hist is None
dist.barrier()
print("rank", rank, "ok", flush=True)
'''


def test_two_rank_gather(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = os.path.join(HERE, "_multi_worker.py")
    with open(script, "w") as f:
        f.write(WORKER)
    try:
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                              "--master-port", "29571", script], capture_output=True, text=True, timeout=600)
    finally:
        os.remove(script)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2
