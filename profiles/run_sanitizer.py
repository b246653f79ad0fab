"""Small driver for compute-sanitizer (memcheck / racecheck): both callers over slices of the benchmark shapes, in both pack
modes, plus a deep slice that takes the wide (16-bit) counter tile.
usage: compute-sanitizer --tool racecheck python profiles/run_sanitizer.py [loci]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guacamole_b200 import abi, callers, synth  # noqa: E402

length = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = callers.Context(0)
for lists in (1, 0):
    ctx.set_option(abi.OPT_DIFFERENCE_LISTS, lists)
    sb = synth.generate([("20", length)], depth=30, seed=7)
    reads = ctx.pack_c(sb.c, ["20"])
    res = callers.germline_threshold(ctx, reads, [(0, 0, length - 1)], threshold=8)
    cnt = callers.pileup_counts(ctx, reads, [(0, 0, min(length, 50_000))])
    std = callers.germline_standard(ctx, reads, [(0, 0, min(length, 200_000))])
    print("germline", "streams" if lists else "read walk", len(res), len(cnt), len(std), res.stats["exact_loci"])
    cb = callers.CompactBatch(sb.c)
    r2 = ctx.pack_v2(cb, ["20"])
    print("compact batch", len(callers.germline_threshold(ctx, r2, [(0, 0, length - 1)], threshold=8)))
    r2.free()
    reads.free()
    n = max(60_000, length // 4)
    ht, hm = synth.generate([("20", n)], depth=60, seed=7, sample=1), synth.generate([("20", n)], depth=30, seed=7, sample=0)
    t, m = ctx.pack_c(ht.c, ["20"]), ctx.pack_c(hm.c, ["20"])
    som = callers.somatic_standard(ctx, t, m, [(0, 0, n - 1)], odds_threshold=20)
    print("somatic", "rows" if lists else "read walk", len(som), som.stats["exact_loci"])
    t.free()
    m.free()
    hd = synth.generate([("amp", 3000)], depth=700, seed=9)
    deep = ctx.pack_c(hd.c, ["amp"])
    res = callers.germline_threshold(ctx, deep, [(0, 0, 2999)], threshold=8)
    print("deep", len(res))
    deep.free()
ctx.close()
