"""Small driver for ncu: packs a synthetic tumor/normal pair (60x/30x) and runs somatic-standard a few times.
usage: python profiles/run_somatic.py [contig_length] [calls]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guacamole_b200 import abi, callers, synth  # noqa: E402

length = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = callers.Context(0)
ctx.set_option(abi.OPT_SORT_RECORDS, int(os.environ.get("GUAC_SORT", "0")))
t = synth.generate([("20", length)], depth=60, seed=20261021, sample=1)
n = synth.generate([("20", length)], depth=30, seed=20261021, sample=0)
rt, rn = ctx.pack_c(t.c, ["20"]), ctx.pack_c(n.c, ["20"])
for _ in range(calls):
    res = callers.somatic_standard(ctx, rt, rn, [(0, 0, length - 1)], odds_threshold=20)
print(len(res), res.stats)
