"""Parity tests proper (-m gpu): the CUDA engine, called through the C ABI of include/guac.h, against the CPU oracle on
the same inputs.  Integer outputs (depths, allele counts, called loci, genotypes, allele strings) must be bit-exact."""
import numpy as np
import pytest

import oracle_binding as orc
from conftest import load_golden
from guacamole_b200 import abi
from guacamole_b200.reads import ReadBatch, make_read

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from guacamole_b200.callers import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module", params=[1, 0], ids=["difference-lists", "cigar-walk"], autouse=True)
def difference_lists(request, ctx):
    """Every parity test runs twice: with reads packed as differences against the reference track (the default) and with
    the pileup kernels walking bit planes / CIGARs for every read (the path reads with many differences take)."""
    from guacamole_b200 import abi
    ctx.set_option(abi.OPT_DIFFERENCE_LISTS, request.param)
    yield request.param
    ctx.set_option(abi.OPT_DIFFERENCE_LISTS, 1)


def gpu_threshold(ctx, batch, ranges, **kw):
    from guacamole_b200 import callers
    reads = ctx.pack(batch)
    res = callers.germline_threshold(ctx, reads, ranges, **kw)
    reads.free()
    return res


def gpu_counts(ctx, batch, ranges, skip_empty=True, reference=None):
    from guacamole_b200 import callers
    reads = ctx.pack(batch, reference)
    res = callers.pileup_counts(ctx, reads, ranges, skip_empty)
    reads.free()
    return res


def assert_threshold_equal(ctx, batch, ranges, threshold=8, emit_ref=False, emit_no_call=False, skip_empty=True, oracle_ranges=None):
    want = orc.germline_threshold(batch, oracle_ranges or ranges, orc.threshold_params(threshold, emit_ref, emit_no_call, skip_empty))
    got = gpu_threshold(ctx, batch, ranges, threshold=threshold, emit_ref=emit_ref, emit_no_call=emit_no_call, skip_empty=skip_empty)
    w, g = want.threshold(), got.genotypes()
    assert len(g) == len(w), (len(g), len(w))
    for a, b in zip(g, w):
        assert a == b, (a, b)
    assert got.stats["loci_visited"] == want.stats["loci_visited"]
    assert got.stats["tie_loci"] == want.stats["tie_loci"]
    return got


def assert_counts_equal(ctx, batch, ranges, skip_empty=True, reference=None):
    want = orc.pileup_counts(batch, ranges, skip_empty, reference=reference).counts()
    got = gpu_counts(ctx, batch, ranges, skip_empty, reference).records
    assert len(got) == len(want)
    for f in ("locus", "contig", "depth", "positive_depth", "reference_depth", "base_count", "other_count", "reference_base"):
        assert np.array_equal(got[f], want[f]), f
    return got


UNIT_SETS = {
    "long_insert": [make_read("TCGATCGA", "8M", "8", 1), make_read("TCGATCGA", "8M", "8", 1), make_read("TCGACCCTCGA", "4M3I4M", "8", 1)],
    "deletion": [make_read("AATTGAATTG", "5M1D5M", "5^C5", 0), make_read("AATTGCAATTG", "11M", "11", 0, is_positive_strand=False)],
    "contig_start_insertion": [make_read("AAAAAACGT", "5I4M", "4", 0), make_read("ACGT", "4M", "4", 0)],
    "mixed_indels": [make_read("TCATCTCAAAAGAGATCGA", "2M2D1M2I2M4I2M2D6M", "2^GA5^TC6", 10)] * 3 + [make_read("TCGAATCGATCGATCGA", "17M", "17", 10)] * 2,
    "het_snv": [make_read("TCGATCGA", "8M", "8", 1), make_read("TCGATCGA", "8M", "8", 1), make_read("GCGATCGA", "8M", "0T7", 1)],
    "clips_and_eq": [make_read("GGACGTACGTACGTACGCC", "2S5M4=1X5=2S", "9G5", 10), make_read("ACGTACGTAC", "2H10M3H", "10", 12)],
    "n_base_in_read": [make_read("ACNTACGT", "8M", "8", 3), make_read("ACGTACGT", "8M", "8", 3), make_read("ACGTACGT", "8M", "2N5", 3)],
    "rna_skip": [make_read("CCCCAGCCTAGG", "7M5000N5M", "12", 100), make_read("CCCCAGC", "7M", "7", 100)],
    "x_before_i": [make_read("ACGTTTACGT", "3M1X2I4M", "3C4", 5), make_read("ACGAACGT", "8M", "8", 5)],
    "deletion_then_mismatch": [make_read("ACGTACGT", "4M2D4M", "4^TT0C3", 20), make_read("ACGTTTCCGT", "10M", "10", 20)],
}


@pytest.mark.parametrize("name", sorted(UNIT_SETS))
def test_unit_sets(ctx, name):
    b = ReadBatch.from_records(UNIT_SETS[name]).sorted()
    ranges = [(0, 0, 6000)]
    assert_counts_equal(ctx, b, ranges)
    assert_counts_equal(ctx, b, [(0, 0, 40)], skip_empty=False)
    for thr in (0, 8, 34, 50):
        assert_threshold_equal(ctx, b, ranges, threshold=thr)
    assert_threshold_equal(ctx, b, ranges, threshold=8, emit_ref=True, emit_no_call=True)
    assert_threshold_equal(ctx, b, ranges, threshold=60, emit_ref=False, emit_no_call=True)


def test_germline_suite_vectors(ctx):  # GermlineThresholdCallerSuite.scala:72-85 through the engine
    hom = ReadBatch.from_records([make_read("TCGATCGA", "8M", "8", 1), make_read("GCGATCGA", "8M", "0T7", 1), make_read("GCGATCGA", "8M", "0T7", 1)])
    g = gpu_threshold(ctx, hom, [(0, 0, 100)], threshold=50).genotypes()
    assert [(x["start"], x["ref"], x["alt"], x["gt"]) for x in g] == [(1, "T", "G", (abi.GT_ALT, abi.GT_ALT))]


def chrm():
    return load_golden("chrM.sorted").filtered(non_duplicate=True, has_md=True).sorted()


def test_chrm_identical_vcf(ctx, tmp_path):
    """BASELINE.json configs[0] end to end: the VCF written from the engine's records equals, line for line, the one written
    from the oracle's (identical field sets after canonical sort)."""
    from guacamole_b200 import callers, loci, vcf
    b = chrm()
    ranges = loci.parse_loci("all", b.contig_names, b.contig_lengths)
    want = orc.germline_threshold(b, ranges).threshold()
    reads = ctx.pack(b)
    got = callers.germline_threshold(ctx, reads, ranges).genotypes()
    reads.free()
    vcf.write_vcf(str(tmp_path / "gpu.vcf"), got, b.contig_names, b.sample_names, b.contig_lengths)
    vcf.write_vcf(str(tmp_path / "oracle.vcf"), want, b.contig_names, b.sample_names, b.contig_lengths)
    g, w = open(tmp_path / "gpu.vcf").read(), open(tmp_path / "oracle.vcf").read()
    assert g == w and g.count("\n") > 138


def test_chrm_config1(ctx):  # BASELINE.json configs[0]: germline-threshold on chrM.sorted.bam, loci "all", threshold 8
    b = chrm()
    got = assert_threshold_equal(ctx, b, [(0, 0, 16570)])
    assert len(got) == 138 and got.stats["loci_visited"] == 15904
    assert_counts_equal(ctx, b, [(0, 0, 16570)])
    assert_threshold_equal(ctx, b, [(0, 0, 16570)], emit_ref=True, emit_no_call=True)
    # LociPartitioning: 8 contiguous tasks give the same records as one task (the oracle runs single-task: its own
    # result at the order-sensitive locus 13854 depends on the task cut)
    parts = orc.partition_loci_uniformly(8, [(0, 0, 16570)])
    assert_threshold_equal(ctx, b, parts, oracle_ranges=[(0, 0, 16570)])


def test_gatk_bundle_indel_heavy(ctx):
    g = load_golden("gatk_mini_bundle_extract").filtered(has_md=True, non_duplicate=True).sorted()
    c = g.contig_names.index("20")
    ranges = [(c, 9999000, 10271000)]
    assert_counts_equal(ctx, g, ranges)
    assert_threshold_equal(ctx, g, ranges, threshold=8)
    assert_threshold_equal(ctx, g, ranges, threshold=0)
    assert_threshold_equal(ctx, g, [(c, 10006000, 10010000)], emit_ref=True, emit_no_call=True)


def test_somatic_fixture_counts(ctx):
    for name, contig in (("tumor.chr20.tough", "20"), ("synthetic.challenge.set1.tumor.v2.withMDTags.chr2.complexvar", "2")):
        b = load_golden(name).filtered(non_duplicate=True, passed_qc=True, has_md=True).sorted()
        c = b.contig_names.index(contig)
        ranges = [(c, 0, int(b.end().max()) + 10)]
        assert_counts_equal(ctx, b, ranges)
        assert_threshold_equal(ctx, b, ranges)


def test_synthetic_shape(ctx):
    from guacamole_b200 import synth
    b = synth.generate([("20", 300000)], depth=30, seed=11).to_read_batch()
    ranges = [(0, 0, 299999)]
    assert_counts_equal(ctx, b, ranges)
    got = assert_threshold_equal(ctx, b, ranges)
    assert len(got) > 100
    assert_threshold_equal(ctx, b, [(0, 1000, 5000), (0, 70000, 70001), (0, 100000, 200000)])


def test_synthetic_multi_contig_and_deep(ctx):
    from guacamole_b200 import synth
    b = synth.generate([("1", 60000), ("2", 30000), ("3", 5000)], depth=40, seed=5).to_read_batch()
    ranges = [(0, 0, 59999), (1, 0, 29999), (2, 0, 4999)]
    assert_counts_equal(ctx, b, ranges)
    assert_threshold_equal(ctx, b, ranges)
    deep = synth.generate([("amp", 3000)], depth=3000, seed=9).to_read_batch()   # > 255 reads deep: 12-plane counters
    assert_counts_equal(ctx, deep, [(0, 0, 2999)])
    assert_threshold_equal(ctx, deep, [(0, 0, 2999)])


def test_fasta_reference(ctx):
    b = ReadBatch.from_records([make_read("TCGATCGA", "8M", "8", 1), make_read("TCGATCGA", "8M", "8", 1), make_read("GCGATCGA", "8M", "0T7", 1)])
    ref = [b"NTCGATCGANNNN"]
    assert_counts_equal(ctx, b, [(0, 0, 12)], reference=ref)
    assert_counts_equal(ctx, b, [(0, 0, 12)], skip_empty=False, reference=ref)


def test_errors(ctx):
    from guacamole_b200._lib import GuacError
    unsorted = ReadBatch.from_records([make_read("TCGATCGA", "8M", "8", 5), make_read("TCGATCGA", "8M", "8", 2)])
    with pytest.raises(GuacError) as e:
        ctx.pack(unsorted)
    assert e.value.code == abi.ERR_UNSORTED_READS
    no_md = ReadBatch.from_records([make_read("TCGATCGA", "8M", None, 5)])
    with pytest.raises(GuacError) as e:
        ctx.pack(no_md)
    assert e.value.code == abi.ERR_MISSING_MD
    bad_md = ReadBatch.from_records([make_read("AATTGAATTG", "5M1D5M", "10", 0)])
    with pytest.raises(GuacError) as e:
        ctx.pack(bad_md)
    assert e.value.code == abi.ERR_MISSING_MD
    bad_cigar = ReadBatch.from_records([make_read("AATTG", "7M", "7", 0)])
    with pytest.raises(GuacError) as e:
        ctx.pack(bad_cigar)
    assert e.value.code == abi.ERR_INVALID_CIGAR
    contig_order = ReadBatch.from_records([make_read("ACGT", "4M", "4", 1, "a"), make_read("ACGT", "4M", "4", 1, "b"), make_read("ACGT", "4M", "4", 9, "a")])
    with pytest.raises(GuacError) as e:
        ctx.pack(contig_order)
    assert e.value.code == abi.ERR_CONTIG_ORDER


def test_empty_inputs(ctx):
    from guacamole_b200 import callers
    empty = ReadBatch.from_records([], contig_names=["chr1"], contig_lengths=[1000])
    reads = ctx.pack(empty)
    assert len(callers.germline_threshold(ctx, reads, [(0, 0, 1000)])) == 0
    assert len(callers.pileup_counts(ctx, reads, [(0, 0, 10)], skip_empty=False)) == 10
    assert len(callers.germline_threshold(ctx, reads, [])) == 0


def test_sharded_ranks_equal_single_run(ctx):
    """LociPartitioning across ranks (emulated one after the other on one GPU): shard reads by each rank's ranges,
    run the engine per shard, concatenate — identical to the unsharded run (T/DistributedUtilSuite.scala:208-220)."""
    from guacamole_b200 import callers, synth
    from guacamole_b200.distributed import ranges_of_rank, shard_reads
    from guacamole_b200.loci import partition_loci_uniformly
    b = synth.generate([("1", 50000), ("2", 30000)], depth=30, seed=23).to_read_batch()
    loci = [(0, 0, 49999), (1, 0, 29999)]
    full = gpu_threshold(ctx, b, loci).genotypes()
    parts = partition_loci_uniformly(4, loci)
    merged = []
    for rank in range(4):
        mine = ranges_of_rank(parts, rank)
        merged += gpu_threshold(ctx, shard_reads(b, mine), mine).genotypes()
    assert sorted(merged, key=lambda r: (r["contig"], r["start"], r["ref"], r["alt"])) == full and len(full) > 100


def test_more_input_errors(ctx):
    from guacamole_b200._lib import GuacError
    import numpy as np

    def code_of(batch, **kw):
        with pytest.raises(GuacError) as e:
            ctx.pack(batch, **kw)
        return e.value.code

    assert code_of(ReadBatch.from_records([make_read("ACGTACGT", "4M2P4M", "8", 3)])) == abi.ERR_INVALID_CIGAR            # P operator
    past = ReadBatch.from_records([make_read("ACGTACGT", "8M", "8", 95, "c")], contig_names=["c"], contig_lengths=[100])
    assert code_of(past) == abi.ERR_INVALID_ARGUMENT                                                                        # read ends past its contig
    badq = ReadBatch.from_records([make_read("ACGT", "4M", "4", 3, quality_scores=[30, 200, 30, 30])])
    assert code_of(badq) == abi.ERR_BAD_QUALITY                                                                             # quality byte > 127
    two = ReadBatch.from_records([make_read("ACGT", "4M", "4", 3, sample="a"), make_read("ACGT", "4M", "4", 4, sample="b")])
    assert code_of(two) == abi.ERR_UNSUPPORTED                                                                              # one sample per read set
    ok = ctx.pack(ReadBatch.from_records([make_read("ACGT", "4M", "4", 3)]))
    from guacamole_b200 import callers
    with pytest.raises(GuacError) as e:
        callers.germline_threshold(ctx, ok, [(5, 0, 10)])
    assert e.value.code == abi.ERR_INVALID_ARGUMENT                                                                         # contig out of range
    ok.free()


def test_depth_overflow_retry_and_long_reads(ctx):
    from guacamole_b200 import synth
    # ~300x: deeper than the 8-bit counter fields but under 2048 reads per granule -> the engine detects the overflow and
    # reruns with 16-bit fields
    b = synth.generate([("c", 6000)], depth=300, seed=41).to_read_batch()
    assert_counts_equal(ctx, b, [(0, 0, 5999)])
    assert_threshold_equal(ctx, b, [(0, 0, 5999)])
    # 700 bp reads span more than the seven plane pairs the fast path preloads (rolling-window path)
    long_reads = synth.generate([("c", 60000)], depth=20, read_length=700, seed=43).to_read_batch()
    assert_counts_equal(ctx, long_reads, [(0, 0, 59999)])
    assert_threshold_equal(ctx, long_reads, [(0, 0, 59999)])
    # reads much shorter than a word
    short_reads = synth.generate([("c", 20000)], depth=15, read_length=25, seed=45, frac_clip=0.0).to_read_batch()
    assert_counts_equal(ctx, short_reads, [(0, 0, 19999)])
    assert_threshold_equal(ctx, short_reads, [(0, 3, 19990)], threshold=0)


def test_no_skip_empty_and_fasta_threshold(ctx):
    b = ReadBatch.from_records([make_read("TCGATCGA", "8M", "8", 1), make_read("TCGATCGA", "8M", "8", 1), make_read("GCGATCGA", "8M", "0T7", 1),
                                make_read("ACGTACGT", "8M", "8", 2000)])
    assert_threshold_equal(ctx, b, [(0, 0, 3000)], threshold=0, skip_empty=False)
    assert_threshold_equal(ctx, b, [(0, 0, 3000)], threshold=0, skip_empty=False, emit_ref=True, emit_no_call=True)
    # with a FASTA reference the pileup's reference base is the FASTA base, whatever the MD tags say
    ref = [b"N" + b"ACGATCGA" + b"N" * 3000]
    from guacamole_b200 import callers
    want = orc.germline_threshold(b, [(0, 0, 20)], orc.threshold_params(0), reference=ref).threshold()
    reads = ctx.pack(b, ref)
    got = callers.germline_threshold(ctx, reads, [(0, 0, 20)], threshold=0).genotypes()
    reads.free()
    assert got == want and any(g["start"] == 1 and g["ref"] == "A" for g in got)


def test_allele_counts_and_variant_loci(ctx):  # VariantSupport / VAFHistogram closures
    from guacamole_b200 import callers, synth
    g = load_golden("gatk_mini_bundle_extract").filtered(has_md=True, non_duplicate=True).sorted()
    c = g.contig_names.index("20")
    syn = synth.generate([("s", 30000)], depth=25, seed=31, sample=0).to_read_batch()
    for batch, ranges in ((g, [(c, 10006800, 10006850), (c, 10008900, 10008960), (c, 9999990, 10000010), (c, 5, 7)]),
                          (syn, [(0, 0, 3000), (0, 12000, 12500)])):
        want = orc.allele_counts(batch, ranges).allele_counts()
        reads = ctx.pack(batch)
        got = callers.allele_counts(ctx, reads, ranges).genotypes()
        key = lambda x: (x["contig"], x["start"], x["ref"], x["alt"], x["count"])
        assert sorted(key(x) for x in got) == sorted(key(x) for x in want)
        assert len(got) > 50
        # VAFHistogram.variantLociFromReads over the per-locus counts
        wc = orc.pileup_counts(batch, ranges).counts()
        gc = callers.pileup_counts(ctx, reads, ranges).records
        reads.free()
        wl, gl = callers.variant_loci(wc, 2, 5), callers.variant_loci(gc, 2, 5)
        assert gl.tolist() == wl.tolist() and len(gl) > 0
        assert callers.generate_vaf_histogram(gl["variant_allele_frequency"], 20) == \
            callers.generate_vaf_histogram(wl["variant_allele_frequency"], 20)


def test_chunked_pack_above_a_million_reads(ctx):
    """guac_reads_pack copies the bases of >= 1,000,000 reads in eight chunks and runs k_pack_bases / k_md_track per chunk
    as they land: parity in windows around chunk boundaries (reads n/8, n/4, n/2, 3n/4) and at both ends."""
    from guacamole_b200 import synth
    L = 5_300_000
    b = synth.generate([("20", L)], depth=30, seed=4242, sample=0).to_read_batch()
    assert len(b) >= 1_000_000
    ranges = [(0, 0, 30_000)]
    for k in (1, 2, 4, 6):
        s = int(b.start[len(b) * k // 8])
        ranges.append((0, max(0, s - 20_000), s + 20_000))
    ranges.append((0, L - 30_000, L - 1))
    assert_threshold_equal(ctx, b, ranges)
    assert_counts_equal(ctx, b, ranges)
    # the by-locus stores finished chunk by chunk underneath the copies (GUAC_OPT_PACK_OVERLAP, the default) and in one launch
    # after the last chunk: the same records over every locus, from the wide and from the compact batch
    from guacamole_b200 import abi, callers
    whole = [(0, 0, L - 1)]
    got = {}
    for overlap in (1, 0):
        ctx.set_option(abi.OPT_PACK_OVERLAP, overlap)
        try:
            for compact in (False, True):
                host = callers.CompactBatch(b, pinned=False, fixed_length=False) if compact else None
                reads = ctx.pack_v2(host, b.contig_names) if compact else ctx.pack(b)
                res = callers.germline_threshold(ctx, reads, whole, threshold=8)
                got[(overlap, compact)] = res.genotypes()
                del res
                reads.free()
                if host is not None:
                    host.free()
        finally:
            ctx.set_option(abi.OPT_PACK_OVERLAP, 1)
    assert len(got[(1, False)]) > 1000
    for k, v in got.items():
        assert v == got[(1, False)], k


def test_by_sample_split(ctx):
    """A batch holding two samples: the reference calls each sample on its own elements (pileup.bySample); the host side
    splits, packs and calls per sample and merges in canonical order — the oracle does the grouping natively."""
    from guacamole_b200 import callers, synth
    from guacamole_b200.reads import concat
    a = synth.generate([("s", 20000)], depth=20, seed=71, sample=0).to_read_batch()   # same seed = same reference,
    b = synth.generate([("s", 20000)], depth=12, seed=71, sample=1).to_read_batch()   # other reads (+ tumor-only variants)
    b.sample[:] = 1
    both = concat([a, b])
    both.sample_names = ["first", "second"]
    both = both.sorted()
    assert set(np.unique(both.sample).tolist()) == {0, 1}
    want = orc.germline_threshold(both, [(0, 0, 19999)]).threshold()
    got = callers.germline_threshold_by_sample(ctx, both, [(0, 0, 19999)], threshold=8)
    key = lambda x: (x["contig"], x["start"], x["sample"], x["ref"], x["alt"], tuple(x["gt"]))
    assert [key(x) for x in got] == [key(x) for x in want]
    assert len({x["sample"] for x in got}) == 2


def test_partition_loci_by_approximate_depth(ctx):  # DistributedUtilSuite.scala:76-93 + the oracle on uneven depth
    from guacamole_b200 import synth
    from guacamole_b200.loci import partition_loci_by_approximate_depth
    golden = ReadBatch.from_records([make_read("A" * n, f"{n}M", f"{n}", s, chr="chr1") for s, n in ((5, 1), (6, 1), (7, 1), (8, 1))])
    reads = ctx.pack(golden)
    got = partition_loci_by_approximate_depth(ctx, 2, [(0, 0, 100)], 100, reads)
    reads.free()
    assert got == [(0, 0, 7, 0), (0, 7, 100, 1)]
    # uneven depth: a deep amplicon inside a shallow contig, two read sets, gaps in the loci
    shallow = synth.generate([("1", 60000), ("2", 30000)], depth=6, seed=11).to_read_batch()
    deep = synth.generate([("1", 60000), ("2", 30000)], depth=200, seed=12, window=(0, 20000, 24000)).to_read_batch()
    loci = [(0, 0, 25000), (0, 30000, 59999), (1, 500, 29999)]
    ra, rb = ctx.pack(shallow), ctx.pack(deep)
    for tasks, acc in ((4, 250), (8, 13), (3, 1), (16, 1000)):
        want = orc.partition_loci_by_approximate_depth(tasks, loci, acc, shallow, deep)
        assert partition_loci_by_approximate_depth(ctx, tasks, loci, acc, ra, rb) == want
    ra.free()
    rb.free()


def test_device_generator_equals_host_generator(ctx):
    """The two builds of the synthetic generator produce the same bytes, a shard holds what the whole genome would, and a
    batch packed from device memory gives the records of the same batch packed from the host."""
    from guacamole_b200 import callers, synth
    contigs = [("1", 300000), ("2", 120000), ("3", 900)]
    for sample, depth in ((0, 30), (1, 45)):
        host = synth.generate(contigs, depth=depth, seed=97, sample=sample)
        dev = synth.generate_device(ctx, contigs, depth=depth, seed=97, sample=sample)
        back = dev.download()
        hb, db = host.to_read_batch(), back.to_read_batch()
        assert len(hb.start) == len(db.start) > 50000
        for col in ("contig", "start", "cigar_off", "cigar", "seq_off", "seq", "qual", "mapq", "flags", "md_off", "md"):
            assert np.array_equal(getattr(hb, col), getattr(db, col)), col
        ranges = [(0, 0, 299999), (1, 0, 119999)]
        a = ctx.pack(hb)
        b = ctx.pack_device(dev.c, [c[0] for c in contigs])
        ra = callers.germline_threshold(ctx, a, ranges).genotypes()
        rb = callers.germline_threshold(ctx, b, ranges).genotypes()
        for g in rb:
            g["sample"] = 0  # (to_read_batch interns the sample as index 0; the generated column carries the sample id)
        assert ra == rb and len(ra) > 300
        a.free()
        b.free()
        back.free()
        dev.free()
    # a shard generated alone: the reads STARTING in its windows, byte for byte
    whole = synth.generate(contigs, depth=30, seed=97).to_read_batch()
    windows = synth.shard_windows([(0, 250000, 300000, 1), (1, 0, 40000, 1)])
    shard = synth.generate_device(ctx, contigs, depth=30, seed=97, windows=windows)
    sb = shard.download().to_read_batch()
    keep = np.zeros(len(whole.start), bool)
    for c, s, e in windows:
        keep |= (whole.contig == c) & (whole.start >= s) & (whole.start < e)
    sel = whole.select(keep)
    assert len(sb.start) == len(sel.start) > 10000
    for col in ("contig", "start", "cigar", "seq", "qual", "mapq", "flags", "md"):
        assert np.array_equal(getattr(sb, col), getattr(sel, col)), col
    shard.free()


def test_depth_histogram(ctx, difference_lists):
    """guac_depth_histogram (the histogram the shards reduce to rank 0) against the depths of guac_pileup_counts."""
    if not difference_lists:
        pytest.skip("the histogram reads the difference streams")
    from guacamole_b200 import callers, synth
    for contigs, depth, ranges in (([("1", 70000), ("2", 9000)], 30, [(0, 100, 65000), (1, 0, 9000)]),
                                   ([("amp", 4000)], 900, [(0, 0, 4100)])):
        b = synth.generate(contigs, depth=depth, seed=3).to_read_batch()
        reads = ctx.pack(b)
        counts = callers.pileup_counts(ctx, reads, ranges, skip_empty=False).records
        want = np.bincount(np.minimum(counts["depth"], 255), minlength=256).astype(np.uint64)
        got = callers.depth_histogram(ctx, reads, ranges)
        reads.free()
        assert int(got.sum()) == sum(r[2] - r[1] for r in ranges)
        assert np.array_equal(got, want)


def test_overlapping_ranges_are_refused(ctx):
    """LociSet merges overlapping ranges; the C ABI asks the caller to (include/guac.h): overlapping input is refused."""
    from guacamole_b200 import callers
    from guacamole_b200._lib import GuacError
    b = ReadBatch.from_records(UNIT_SETS["het_snv"]).sorted()
    reads = ctx.pack(b)
    for fn in (lambda r: callers.germline_threshold(ctx, reads, r), lambda r: callers.pileup_counts(ctx, reads, r),
               lambda r: callers.allele_counts(ctx, reads, r)):
        with pytest.raises(GuacError) as e:
            fn([(0, 0, 6), (0, 4, 9)])
        assert e.value.code == abi.ERR_INVALID_ARGUMENT
        assert len(fn([(0, 4, 9), (0, 0, 4)])) > 0   # unordered but disjoint is fine
    reads.free()


def _same_results(ctx, batch, ranges, reference=None):
    """The compact batch (guac_reads_pack_v2) packs into the same store as the wide one: identical results."""
    from guacamole_b200 import callers
    wide = ctx.pack(batch, reference)
    cb = callers.CompactBatch(batch, fixed_length=True)
    compact = ctx.pack_v2(cb, batch.contig_names, batch.sample_names, reference=reference)
    assert compact.n_reads == wide.n_reads
    assert compact.h2d_bytes < wide.h2d_bytes or len(batch) == 0
    n = 0
    for thr, emit in ((8, False), (0, True)):
        a = callers.germline_threshold(ctx, wide, ranges, threshold=thr, emit_ref=emit, emit_no_call=emit)
        b = callers.germline_threshold(ctx, compact, ranges, threshold=thr, emit_ref=emit, emit_no_call=emit)
        # (general records pool their allele strings in the order the warps finish: compare the decoded records)
        assert a.genotypes() == b.genotypes() and a.stats["loci_visited"] == b.stats["loci_visited"]
        n += len(a)
    a, b = callers.pileup_counts(ctx, wide, ranges), callers.pileup_counts(ctx, compact, ranges)
    assert np.array_equal(a.records, b.records)
    a, b = callers.germline_standard(ctx, wide, ranges), callers.germline_standard(ctx, compact, ranges)
    assert repr(a.genotypes()) == repr(b.genotypes())  # (repr: NaN fields compare equal; allele strings are pooled in the order the warps finish: compare the decoded records)
    wide.free()
    compact.free()
    cb.free()
    return n


@pytest.mark.parametrize("name", sorted(UNIT_SETS))
def test_compact_batch_unit_sets(ctx, name):
    b = ReadBatch.from_records(UNIT_SETS[name]).sorted()
    assert _same_results(ctx, b, [(0, 0, 6000)]) > 0


def test_compact_batch_real_and_synthetic(ctx):
    from guacamole_b200 import callers, synth
    from guacamole_b200._lib import GuacError
    chrm = load_golden("chrM.sorted").filtered(non_duplicate=True, passed_qc=True, has_md=True).sorted()
    assert _same_results(ctx, chrm, [(chrm.contig_names.index("chrM"), 0, 16571)]) >= 138
    # several contigs (one of them empty), reads of unequal length (seq_off stays), above the chunked-copy limit
    contigs = [("1", 2_700_000), ("2", 900), ("4", 2_600_000)]
    b = synth.generate(contigs, depth=30, seed=515, sample=0).to_read_batch()
    assert len(b) >= 1_000_000
    ranges = [(0, 0, 40_000), (0, 2_650_000, 2_700_000), (1, 0, 900), (2, 0, 50_000), (2, 2_560_000, 2_600_000)]
    for k in (1, 2, 3):
        i = len(b) * k // 4
        s, c = int(b.start[i]), int(b.contig[i])
        ranges.append((c, s + 60_000, s + 90_000) if k == 2 and c == 0 else (c, max(50_000, s - 15_000), s + 15_000))
    ranges = [r for r in ranges if r[2] > r[1]]
    assert _same_results(ctx, b, ranges) > 300
    reference = [bytes(np.random.default_rng(3).choice(np.frombuffer(b"ACGT", np.uint8), n).tobytes()) for _, n in contigs[:2]]
    small = synth.generate(contigs[1:2] + contigs[1:2], depth=20, seed=5, sample=0).to_read_batch()
    assert _same_results(ctx, small, [(0, 0, 900), (1, 0, 900)], reference=[reference[1], reference[1]]) > 0
    # what the compact form cannot hold is refused by the converter, not mangled
    lower = ReadBatch.from_records([make_read("acgtACGT", "8M", "8", 1)])
    with pytest.raises(GuacError) as e:
        callers.CompactBatch(lower)
    assert e.value.code == abi.ERR_INVALID_ARGUMENT
    two = ReadBatch.from_records([make_read("ACGT", "4M", "4", 1, sample="a"), make_read("ACGT", "4M", "4", 2, sample="b")])
    with pytest.raises(GuacError) as e:
        callers.CompactBatch(two)
    assert e.value.code == abi.ERR_INVALID_ARGUMENT
    empty = callers.CompactBatch(ReadBatch.from_records([], contig_names=["1"]))
    r = ctx.pack_v2(empty, ["1"], ["s"])
    assert r.n_reads == 0 and len(callers.germline_threshold(ctx, r, [(0, 0, 100)])) == 0
    # a contig without reads between two that have some
    gap = ReadBatch.from_records([make_read("TCGATCGA", "8M", "8", 1, chr="a"), make_read("GCGATCGA", "8M", "0T7", 1, chr="a"),
                                  make_read("GCGATCGA", "8M", "0T7", 3, chr="c")], contig_names=["a", "b", "c"])
    assert _same_results(ctx, gap, [(0, 0, 50), (1, 0, 50), (2, 0, 50)]) > 0


@pytest.mark.gpu
def test_call_segments_and_repeated_calls(ctx):
    """GUAC_OPT_SEGMENTS cuts a germline call into up to four segments of tiles whose exact kernels share the general-record
    buffers and draw their loci from per-segment ticket counters (which reset themselves): the same records as one segment,
    equal to the oracle's, call after call on one context (indel-rich reads so that the exact kernel has work in every segment)."""
    from guacamole_b200 import abi, callers, synth
    L = 1_000_000
    b = synth.generate([("20", L)], depth=30, seed=977, sample=0, frac_ins=0.03, frac_del=0.03).to_read_batch()
    ranges = [(0, 0, L - 1)]
    want = orc.germline_threshold(b, ranges, orc.threshold_params(8)).threshold()
    reads = ctx.pack(b)
    try:
        for seg in (1, 2, 3, 4, 1):
            ctx.set_option(abi.OPT_SEGMENTS, seg)
            for _ in range(2):
                got = callers.germline_threshold(ctx, reads, ranges, threshold=8)
                assert got.stats["exact_loci"] > 50
                assert got.genotypes() == want, seg
    finally:
        ctx.set_option(abi.OPT_SEGMENTS, 1)
        reads.free()
