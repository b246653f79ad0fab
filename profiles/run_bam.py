"""Decode rate of guac_bam_load (host threads, no GPU): writes a synthetic 30x BAM of `loci` loci, loads it a few times.
usage: python profiles/run_bam.py [loci] [threads]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guacamole_b200 import callers, synth  # noqa: E402
from guacamole_b200.reads import write_bam  # noqa: E402

loci = int(sys.argv[1]) if len(sys.argv) > 1 else 1_500_000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
batch = synth.generate([("20", loci)], depth=30, seed=11).to_read_batch()
path = os.path.join(tempfile.mkdtemp(), "synth.bam")
t0 = time.perf_counter()
write_bam(batch, path, level=1)
print(f"wrote {len(batch):,} reads, {os.path.getsize(path) / 1e6:.1f} MB in {time.perf_counter() - t0:.1f} s (Python writer)")
for with_q in (True, False):
    best = None
    for _ in range(3):
        cb = callers.CompactBatch.from_bam(path, non_duplicate=True, passed_qc=True, has_md_tag=True, with_qualities=with_q, n_threads=threads)
        st = cb.decode_stats
        best = st if best is None or st["decode_ms"] < best["decode_ms"] else best
        cb.free()
    s = best["decode_ms"] * 1e-3
    print(f"qualities={with_q}: {best['reads']:,} reads in {best['decode_ms']:.1f} ms = {best['reads'] / s / 1e6:.2f} M reads/s, "
          f"{best['file_bytes'] / s / 1e9:.2f} GB/s of file, {best['inflated_bytes'] / s / 1e9:.2f} GB/s inflated ({threads or os.cpu_count()} threads)")
os.remove(path)
