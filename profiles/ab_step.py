"""A/B driver: device-generated chr20-shaped read sets, resident germline-threshold and somatic-standard steps timed with the
context's CUDA-event stopwatch; prints the record count and a checksum so that variants (environment toggles, options) can be
compared on one box.  usage: python profiles/ab_step.py [germline|somatic|both] [contig_length] [calls]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guacamole_b200 import abi, callers, synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "both"
length = int(sys.argv[2]) if len(sys.argv) > 2 else 63_025_520
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ctx = callers.Context(0)
ranges = [(0, 0, length - 1)]
contigs = [("20", length)]


def checksum(res, somatic):
    r = res.records
    if somatic:
        return int(r["start"].astype(np.int64).sum() + r["tumor"]["allele_read_depth"].astype(np.int64).sum() * 7
                   + np.round(r["tumor"]["mean_base_quality"] * 1000).astype(np.int64).sum())
    return int(r["start"].astype(np.int64).sum() + r["gt"].astype(np.int64).sum() * 7)


def timed(fn, somatic):
    for _ in range(3):
        res = fn()
    ctx.timer_start()
    t0 = time.perf_counter()
    for _ in range(calls):
        res = fn()
    ms = ctx.timer_stop() / calls
    wall = (time.perf_counter() - t0) * 1e3 / calls
    print("somatic" if somatic else "germline", "ms/call", round(ms, 4), "wall", round(wall, 4), "records", len(res), "checksum", checksum(res, somatic),
          {k: round(float(res.stats[k]), 4) for k in ("tile_kernel_ms", "exact_kernel_ms", "kernel_launches", "exact_loci")}, flush=True)


if what in ("germline", "both"):
    ctx.set_option(abi.OPT_PACK_QUALITIES, 0)
    dv = synth.generate_device(ctx, contigs, depth=30, seed=20261020, sample=0, with_qualities=False)
    reads = ctx.pack_synth(dv)
    print("germline pack_kernel_ms", round(reads.pack_kernel_ms, 3), "expand", round(reads.expand_kernel_ms, 3), flush=True)
    timed(lambda: callers.germline_threshold(ctx, reads, ranges, threshold=8), False)
    reads.free()
    ctx.set_option(abi.OPT_TRIM_CACHE, 1)
if what in ("somatic", "both"):
    ctx.set_option(abi.OPT_PACK_QUALITIES, 1)
    packed = []
    for s, d in ((1, 60), (0, 30)):
        dv = synth.generate_device(ctx, contigs, depth=d, seed=20261020, sample=s, with_qualities=True)
        packed.append(ctx.pack_synth(dv))
    timed(lambda: callers.somatic_standard(ctx, packed[0], packed[1], ranges, odds_threshold=20), True)
