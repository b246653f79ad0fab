"""Small driver for ncu: packs a synthetic chr20-shape slice and runs germline-threshold a few times.
usage: python profiles/run_germline.py [contig_length] [calls]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guacamole_b200 import abi, callers, synth  # noqa: E402

length = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sb = synth.generate([("20", length)], depth=30, seed=20261020)
ctx = callers.Context(0)
ctx.set_option(abi.OPT_PACK_QUALITIES, 0)
ctx.set_option(abi.OPT_SORT_RECORDS, int(os.environ.get("GUAC_SORT", "1")))
ctx.set_option(abi.OPT_SEGMENTS, int(os.environ.get("GUAC_SEGMENTS", "1")))
reads = ctx.pack_c(sb.c, ["20"])
import time
for _ in range(3):
    res = callers.germline_threshold(ctx, reads, [(0, 0, length - 1)], threshold=8)
ctx.timer_start()
t0 = time.perf_counter()
for _ in range(calls):
    res = callers.germline_threshold(ctx, reads, [(0, 0, length - 1)], threshold=8)
ms = ctx.timer_stop() / calls
print(len(res), "ms/call", round(ms, 4), "wall", round((time.perf_counter() - t0) * 1e3 / calls, 4),
      {k: res.stats[k] for k in ("tile_kernel_ms", "exact_kernel_ms", "kernel_launches", "exact_loci")})
