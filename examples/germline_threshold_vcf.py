"""The reference's CPU-runnable case end to end (BASELINE.json configs[0]):
    guacamole germline-threshold --reads chrM.sorted.bam --out calls.vcf
on the B200 engine.  usage: python examples/germline_threshold_vcf.py reads.{bam,sam,npz} out.vcf [threshold]
(needs a GPU; input filters of GermlineThresholdCaller.scala:61-63: mapped, non-duplicate, has an MD tag)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guacamole_b200 import callers, loci, reads, vcf  # noqa: E402


def main(path, out_path, threshold=8):
    if path.endswith(".bam"):
        batch = reads.load_bam(path)
    elif path.endswith(".sam"):
        batch = reads.load_sam(path)
    else:
        batch = reads.ReadBatch.load_npz(path)
    batch = batch.filtered(non_duplicate=True, has_md=True).sorted()
    ctx = callers.Context(0)
    packed = ctx.pack(batch)
    ranges = loci.parse_loci("all", batch.contig_names, batch.contig_lengths)  # drops the last base of every contig
    result = callers.germline_threshold(ctx, packed, ranges, threshold=threshold)
    n = vcf.write_vcf(out_path, result.genotypes(), batch.contig_names, batch.sample_names, batch.contig_lengths)
    print(f"{len(batch)} reads, {result.stats['loci_visited']} non-empty loci, {len(result)} genotypes, {n} VCF lines -> {out_path}")
    packed.free()
    ctx.close()


if __name__ == "__main__":
    if len(sys.argv) < 3:
        sys.exit(__doc__)
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 8)
