"""Converts the reference's test resources (SAM/BAM read files) into compact columnar fixtures (.npz) that travel
with the repo (the GPU box has no /root/reference).  Run in the build container:

    python tests/golden/make_fixtures.py

Only read DATA is converted (no reference source code); every mapped read is kept in file order with all flag
bits, so each test applies the same Read.InputFilters the reference suite applies.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from guacamole_b200.reads import load_reads  # noqa: E402

RES = "/root/reference/src/test/resources"
FILES = [
    "chrM.sorted.bam", "gatk_mini_bundle_extract.bam", "different_start_reads.sam", "same_start_reads.sam",
    "same_start_reads_snv_tumor.sam", "testrna.sam", "rna_chr17_41244936.sam", "mdtagissue.sam",
    "tumor.chr20.tough.sam", "normal.chr20.tough.sam", "tumor.chr20.simplefp.sam", "normal.chr20.simplefp.sam",
    "synthetic.challenge.set1.normal.v2.withMDTags.chr2.syn1fp.sam",
    "synthetic.challenge.set1.tumor.v2.withMDTags.chr2.syn1fp.sam",
    "synthetic.challenge.set1.normal.v2.withMDTags.chr2.complexvar.sam",
    "synthetic.challenge.set1.tumor.v2.withMDTags.chr2.complexvar.sam",
]

if __name__ == "__main__":
    for f in FILES:
        b = load_reads(os.path.join(RES, f))
        out = os.path.join(HERE, f.rsplit(".", 1)[0] + ".npz")
        b.save_npz(out)
        print(f, len(b), "reads ->", os.path.getsize(out), "bytes")
