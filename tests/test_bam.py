"""guac_bam_load: the native BAM -> compact batch front end (host threads) against the Python reader and the
batches the files were written from.  Replaces Read.fromSAMRecord / Read.InputFilters (reads/Read.scala:88-136, 217-291)."""
import os

import numpy as np
import pytest

from conftest import load_golden
from guacamole_b200 import abi, callers
from guacamole_b200.reads import ReadBatch, load_bam, make_read, write_bam

COLUMNS = ("contig", "start", "cigar_off", "cigar", "seq_off", "seq", "qual", "mapq", "flags", "md_off", "md")


def assert_same_reads(got: ReadBatch, want: ReadBatch):
    assert len(got) == len(want)
    assert list(got.contig_names) == list(want.contig_names)
    for col in COLUMNS:
        a, b = getattr(got, col), getattr(want, col)
        assert np.array_equal(np.asarray(a).astype(np.int64), np.asarray(b).astype(np.int64)), col


def test_bam_round_trip_chrm(tmp_path):
    want = load_golden("chrM.sorted")
    path = str(tmp_path / "chrM.bam")
    write_bam(want, path, block_bytes=3001)  # (small, odd members: records and their fields straddle member boundaries)
    cb = callers.CompactBatch.from_bam(path, n_threads=5)
    st = cb.decode_stats
    assert st["reads"] == len(want) == st["records_in_file"] and st["inflated_bytes"] > st["file_bytes"] > 0
    assert_same_reads(cb.to_read_batch(), want)
    assert_same_reads(load_bam(path), want)       # the Python reader agrees on the same file
    cb.free()
    # Read.InputFilters on the flag bits
    for kw in (dict(non_duplicate=True), dict(passed_qc=True, has_md_tag=True), dict(non_duplicate=True, passed_qc=True, has_md_tag=True, is_paired=True)):
        cb = callers.CompactBatch.from_bam(path, **kw)
        f = want.filtered(non_duplicate=kw.get("non_duplicate", False), passed_qc=kw.get("passed_qc", False), has_md=kw.get("has_md_tag", False),
                          is_paired=kw.get("is_paired", False))
        assert_same_reads(cb.to_read_batch(), f)
        cb.free()
    cb = callers.CompactBatch.from_bam(path, with_qualities=False)
    assert not cb.c.qual and cb.h2d_bytes < st["inflated_bytes"]
    cb.free()


def test_bam_unsorted_odd_lengths_and_samples(tmp_path):
    recs = [make_read("ACGTN", "5M", "5", 7, chr="b", sample="t"), make_read("ACG", "3M", "3", 2, chr="b", sample="t"),
            make_read("TTGCA=A", "2S5M", "5", 4, chr="a", sample="t", is_positive_strand=False, alignment_quality=3),
            make_read("ACGTACGTA", "4M1D5M", "4^C5", 1, chr="c", sample="n"), make_read("GG", "2M", None, 9, chr="a", sample="t")]
    b = ReadBatch.from_records(recs, contig_names=["a", "b", "c"])
    path = str(tmp_path / "mixed.bam")
    write_bam(b, path, block_bytes=97)
    with pytest.raises(callers.GuacError) as e:   # two samples and none named
        callers.CompactBatch.from_bam(path)
    assert e.value.code == abi.ERR_INVALID_ARGUMENT and "several samples" in str(e.value)
    cb = callers.CompactBatch.from_bam(path, sample="t")
    got = cb.to_read_batch()
    want = b.select(b.sample == b.sample_names.index("t")).sorted()
    assert cb.sample_names == ["t"] and list(got.start) == [4, 9, 2, 7] and cb.c.read_length == 0
    assert_same_reads(got, want)
    assert [cb.c.contig_read_off[i] for i in range(4)] == [0, 2, 4, 4]
    cb.free()
    cb = callers.CompactBatch.from_bam(path, sample="n", has_md_tag=True)
    assert cb.c.n_reads == 1 and cb.c.read_length == 9 and cb.to_read_batch().seq.tobytes() == b"ACGTACGTA"
    cb.free()
    with pytest.raises(callers.GuacError):
        callers.CompactBatch.from_bam(str(tmp_path / "missing.bam"))
    bad = tmp_path / "bad.bam"
    bad.write_bytes(open(path, "rb").read()[:-40])
    with pytest.raises(callers.GuacError):
        callers.CompactBatch.from_bam(str(bad))


@pytest.mark.skipif(not os.path.exists("/root/reference/src/test/resources/gatk_mini_bundle_extract.bam"), reason="reference resources absent")
def test_bam_reference_resources():
    """The reference's own BAM test resources (this container only): native loader == Python reader, record for record."""
    for name in ("chrM.sorted.bam", "gatk_mini_bundle_extract.bam"):
        path = "/root/reference/src/test/resources/" + name
        want = load_bam(path)
        for sample in sorted(set(want.sample_names)):
            cb = callers.CompactBatch.from_bam(path, sample=sample)
            w = want.select(want.sample == want.sample_names.index(sample))
            order = np.lexsort((np.arange(len(w)), w.start, w.contig))
            assert_same_reads(cb.to_read_batch(), w.select(order))
            cb.free()


@pytest.mark.gpu
def test_bam_to_records_on_gpu(tmp_path):
    """BASELINE.json configs[0] from the file: BAM -> guac_bam_load -> guac_reads_pack_v2 -> germline-threshold = 138 records."""
    want = load_golden("chrM.sorted")
    path = str(tmp_path / "chrM.bam")
    write_bam(want, path)
    ctx = callers.Context(0)
    cb = callers.CompactBatch.from_bam(path, non_duplicate=True, passed_qc=True, has_md_tag=True, with_qualities=False)
    ctx.set_option(abi.OPT_PACK_QUALITIES, 0)
    reads = ctx.pack_v2(cb, cb.contig_names, cb.sample_names)
    got = callers.germline_threshold(ctx, reads, [(cb.contig_names.index("chrM"), 0, 16571)], threshold=8)
    ref = ctx.pack(want.filtered(non_duplicate=True, passed_qc=True, has_md=True).sorted())
    exp = callers.germline_threshold(ctx, ref, [(want.contig_names.index("chrM"), 0, 16571)], threshold=8)
    assert len(got) == 138 and got.genotypes() == exp.genotypes()
    reads.free(); ref.free(); cb.free(); ctx.close()
