import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from guacamole_b200 import abi, callers, synth
length = 4_000_000
sb = synth.generate([("20", length)], depth=30, seed=20261020)
ctx = callers.Context(0)
ctx.set_option(abi.OPT_PACK_QUALITIES, 0)
out = {}
for mode in (1, 0):
    ctx.set_option(abi.OPT_DIFFERENCE_LISTS, mode)
    reads = ctx.pack_c(sb.c, ["20"])
    res = callers.germline_threshold(ctx, reads, [(0, 0, length - 1)], threshold=8)
    print(mode, len(res), res.stats["exact_loci"], res.stats["tile_kernel_ms"], res.stats["exact_kernel_ms"])
    c, g, s = res.compact()
    out[mode] = (np.sort(c.copy()), g.copy(), res.bytes)
    if mode == 1:
        counts = callers.pileup_counts(ctx, reads, [(0, 0, length - 1)]).records.copy()
    reads.free()
g1, g0 = out[1][1], out[0][1]
extra = np.setdiff1d(g1["start"], g0["start"])
print("extra general loci", len(extra))
idx = {int(l): i for i, l in enumerate(counts["locus"])}
for l in extra[:25]:
    r = counts[idx[int(l)]]
    print(int(l), l % 1024, "depth", r["depth"], "other", r["other_count"], "bases", r["base_count"], chr(r["reference_base"]))
