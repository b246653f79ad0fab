"""The N > 1 path on CPU: world_size-2 gloo.  Each rank shards the reads by its loci ranges, runs the per-rank caller
(here the oracle stands in for the CUDA engine, which needs a GPU), rank 0 gathers the records; the result must equal
the single-task run."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle_binding as orc
    from guacamole_b200 import synth
    from guacamole_b200.callers import THRESHOLD_DTYPE
    from guacamole_b200.distributed import gather_records, ranges_of_rank, shard_reads
    from guacamole_b200.loci import partition_loci_uniformly
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    contigs = [("1", 40000), ("2", 25000)]
    batch = synth.generate(contigs, depth=25, seed=17).to_read_batch()
    loci = [(0, 0, 39999), (1, 0, 24999)]
    parts = partition_loci_uniformly(world, loci)
    mine = ranges_of_rank(parts, rank)
    shard = shard_reads(batch, mine)
    assert 0 < len(shard) < len(batch)
    res = orc.germline_threshold(shard, [(c, s, e) for (c, s, e, _) in mine], orc.threshold_params(8))
    recs = res.threshold()
    arr = np.zeros(len(recs), THRESHOLD_DTYPE)
    pool = bytearray()
    for i, r in enumerate(recs):
        arr[i]["start"], arr[i]["contig"], arr[i]["sample"] = r["start"], r["contig"], r["sample"]
        arr[i]["ref_off"], arr[i]["ref_len"] = len(pool), len(r["ref"])
        pool += r["ref"].encode()
        arr[i]["alt_off"], arr[i]["alt_len"] = len(pool), len(r["alt"])
        pool += r["alt"].encode()
        arr[i]["gt"] = r["gt"]
    allrec, allpool = gather_records(arr, bytes(pool), dst=0)
    if rank == 0:
        got = [(int(r["contig"]), int(r["start"]), allpool[r["ref_off"]:r["ref_off"] + r["ref_len"]].decode(),
                allpool[r["alt_off"]:r["alt_off"] + r["alt_len"]].decode(), (int(r["gt"][0]), int(r["gt"][1]))) for r in allrec]
        full = orc.germline_threshold(batch, loci, orc.threshold_params(8)).threshold()
        want = [(r["contig"], r["start"], r["ref"], r["alt"], r["gt"]) for r in full]
        with open(out_path, "w") as fh:
            fh.write("ok" if got == want and len(got) > 50 else f"mismatch {len(got)} {len(want)}")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather(tmp_path):
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_shard_reads_duplicates_boundary_reads():
    sys.path.insert(0, ROOT)
    from guacamole_b200.distributed import shard_reads
    from guacamole_b200.reads import ReadBatch, make_read
    b = ReadBatch.from_records([make_read("ACGTACGT", "8M", "8", 0), make_read("ACGTACGT", "8M", "8", 6), make_read("ACGTACGT", "8M", "8", 20)])
    left, right = shard_reads(b, [(0, 0, 10, 0)]), shard_reads(b, [(0, 10, 40, 1)])
    assert list(left.start) == [0, 6] and list(right.start) == [6, 20]   # the read over the cut is on both sides
