"""Pins the oracle's callers, likelihoods, windowing and partitioning to the reference's own known-answer tests:
GermlineThresholdCallerSuite, LikelihoodSuite (eps 1e-12), SomaticStandardCallerSuite (47 call/no-call decisions on
real SAM slices + 8 indel allele strings), AlleleEvidenceSuite, VariantSupportSuite (distinct alleles),
DistributedUtilSuite and SlidingWindowSuite (all under src/test/scala/org/hammerlab/guacamole/)."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_binding as orc
from conftest import load_golden
from guacamole_b200 import abi
from guacamole_b200.reads import ReadBatch, make_read


def batch(*reads):
    return ReadBatch.from_records(list(reads))


REF3 = [make_read("TCGATCGA", "8M", "8", 1)] * 3
HET = [make_read("TCGATCGA", "8M", "8", 1), make_read("TCGATCGA", "8M", "8", 1), make_read("GCGATCGA", "8M", "0T7", 1)]
HOM = [make_read("TCGATCGA", "8M", "8", 1), make_read("GCGATCGA", "8M", "0T7", 1), make_read("GCGATCGA", "8M", "0T7", 1)]
REF, ALT, OTHER, NOCALL = abi.GT_REF, abi.GT_ALT, abi.GT_OTHER_ALT, abi.GT_NO_CALL


# ---- GermlineThresholdCallerSuite.scala:30-113 ----------------------------------------------------------------------
def test_threshold_no_variants():
    g = orc.threshold_at(batch(*REF3), 0, 1, orc.threshold_params(0, True, True))
    assert [x["gt"] for x in g] == [(REF, REF)] and g[0]["ref"] == "T" and g[0]["alt"] == "<ALT>"


@pytest.mark.parametrize("thr,expected", [(0, (REF, ALT)), (30, (REF, ALT)), (50, (REF, REF))])
def test_threshold_het(thr, expected):
    g = orc.threshold_at(batch(*HET), 0, 1, orc.threshold_params(thr, True, True))
    assert [x["gt"] for x in g] == [expected]


def test_threshold_hom_alt():
    g = orc.threshold_at(batch(*HOM), 0, 1, orc.threshold_params(50, False, True))
    assert len(g) == 1 and g[0]["gt"] == (ALT, ALT) and (g[0]["start"], g[0]["ref"], g[0]["alt"]) == (1, "T", "G")
    b = batch(*[make_read("TGGATCGA", "8M", "1C6", 1)] * 3)
    g = orc.threshold_at(b, 0, 2, orc.threshold_params(50, False, True))
    assert len(g) == 1 and g[0]["gt"] == (ALT, ALT) and (g[0]["start"], g[0]["ref"], g[0]["alt"]) == (2, "C", "G")


def test_threshold_heterozygous_deletion_regression():  # issue 302, :104-113
    b = load_golden("synthetic.challenge.set1.normal.v2.withMDTags.chr2.syn1fp").filtered(non_duplicate=True, passed_qc=True)
    contig = b.contig_names.index("2")
    assert orc.threshold_at(b, contig, 16050070, orc.threshold_params(8, False, True)) == []


# ---- LikelihoodSuite.scala (eps = 1e-12) ---------------------------------------------------------------------------
E30, E40 = 10 ** -3.0, 10 ** -4.0


def ref_read(q):
    return make_read("C", "1M", "1", 1, "chr1", [q])


def alt_read(q):
    return make_read("A", "1M", "0C0", 1, "chr1", [q])


def lk(reads, **kw):
    return {(g["a1"][1], g["a2"][1]): g["value"] for g in orc.likelihoods_at(batch(*reads), 0, 1, **kw)}


def test_likelihood_all_ref():
    got = lk([ref_read(30), ref_read(40), ref_read(30)])
    assert set(got) == {("C", "C")}
    assert abs(got[("C", "C")] - (1 - E30) * (1 - E40) * (1 - E30)) < 1e-12


def test_likelihood_mix():
    got = lk([ref_read(30), ref_read(40), alt_read(30)])
    assert set(got) == {("C", "C"), ("A", "C"), ("A", "A")}
    assert abs(got[("C", "C")] - (1 - E30) * (1 - E40) * E30) < 1e-12
    assert abs(got[("A", "C")] - 1 / 8.0) < 1e-12
    assert abs(got[("A", "A")] - E30 * E40 * (1 - E30)) < 1e-12
    got = lk([ref_read(30), ref_read(40), alt_read(30)], log_space=True)
    assert abs(got[("C", "C")] - (math.log(1 - E30) + math.log(1 - E40) + math.log(E30))) < 1e-12
    assert abs(got[("A", "C")] - math.log(1 / 8.0)) < 1e-12
    assert abs(got[("A", "A")] - (math.log(E30) + math.log(E40) + math.log(1 - E30))) < 1e-12


def test_likelihood_all_alt():
    got = lk([alt_read(30), alt_read(40), alt_read(30)])
    assert set(got) == {("A", "A")}
    assert abs(got[("A", "A")] - (1 - E30) * (1 - E40) * (1 - E30)) < 1e-12
    got = lk([alt_read(30), alt_read(40), alt_read(30)], log_space=True)
    assert abs(got[("A", "A")] - (2 * math.log(1 - E30) + math.log(1 - E40))) < 1e-12


def test_phred_utils():
    assert orc.lib().orc_phred_to_success_probability(30) == 1.0 - 10 ** -3.0
    assert orc.lib().orc_success_probability_to_phred(C_double(0.999)) == 30


def C_double(x):
    import ctypes
    return ctypes.c_double(x)


# ---- SomaticStandardCallerSuite.scala ----------------------------------------------------------------------------------
SOM = dict(odds=120, min_mapq=1, filter_multi_allelic=False)
FILTER = dict(min_tumor_read_depth=8, max_tumor_read_depth=200, min_normal_read_depth=4,
              min_tumor_alternate_read_depth=3, min_log_odds=120, min_vaf=5, min_likelihood=70)


def found_variant(tumor, normal, contig, locus):
    recs = orc.somatic_at(tumor, normal, contig, locus, orc.somatic_params(**SOM))
    return any(orc.somatic_genotype_filter(r["_raw"], **FILTER) for r in recs)


def tn(tumor_name, normal_name):
    t = load_golden(tumor_name).filtered(non_duplicate=True, passed_qc=True)
    n = load_golden(normal_name).filtered(non_duplicate=True, passed_qc=True)
    return t, n


POSITIVE_TOUGH = [42999694, 25031215, 44061033, 45175149, 755754, 1843813, 3555766, 3868620, 9896926, 14017900,
                  17054263, 35951019, 50472935, 51858471, 58201903, 7087895, 19772181, 30430960, 32150541, 42186626,
                  44973412, 46814443, 52311925, 53774355, 57280858, 62262870]


def test_somatic_positive_tough():  # :82-88
    t, n = tn("tumor.chr20.tough", "normal.chr20.tough")
    c = t.contig_names.index("20")
    assert [found_variant(t, n, c, p) for p in POSITIVE_TOUGH] == [True] * len(POSITIVE_TOUGH)


def test_somatic_negative_syn1():  # :90-97
    t, n = tn("synthetic.challenge.set1.tumor.v2.withMDTags.chr2.syn1fp", "synthetic.challenge.set1.normal.v2.withMDTags.chr2.syn1fp")
    c = t.contig_names.index("2")
    neg = [216094721, 3529313, 8789794, 104043280, 104175801, 126651101, 241901237, 57270796, 120757852]
    assert [found_variant(t, n, c, p) for p in neg] == [False] * len(neg)


def test_somatic_complexvar():  # :99-108
    t, n = tn("synthetic.challenge.set1.tumor.v2.withMDTags.chr2.complexvar", "synthetic.challenge.set1.normal.v2.withMDTags.chr2.complexvar")
    c = t.contig_names.index("2")
    neg = [148487667, 134307261, 90376213, 3638733, 109347468]
    assert [found_variant(t, n, c, p) for p in neg] == [False] * len(neg)
    assert [found_variant(t, n, c, p) for p in (82949713, 130919744)] == [True, True]


def test_somatic_difficult_negative():  # :110-115   (one read at 20:26211835 has no MD tag; the path tolerates it)
    t, n = tn("tumor.chr20.simplefp", "normal.chr20.simplefp")
    c = t.contig_names.index("20")
    neg = [26211835, 29652479, 54495768, 13046318, 25939088]
    assert [found_variant(t, n, c, p) for p in neg] == [False] * len(neg)


NORMAL8 = [make_read("TCGATCGA", "8M", "8", 0)] * 3


def som_alleles(tumor, normal, locus):
    return [(r["ref"], r["alt"]) for r in orc.somatic_at(batch(*tumor), batch(*normal), 0, locus, orc.somatic_params(odds=2))]


def test_somatic_indels():  # :117-262
    assert som_alleles([make_read("TCGGTCGA", "8M", "3G4", 0)] * 3, NORMAL8, 2) == []
    assert som_alleles([make_read("TCGTCGA", "3M1D4M", "3^A4", 0)] * 3, NORMAL8, 2) == [("GA", "G")]
    assert som_alleles([make_read("TCGAAAAGCT", "5M6D5M", "5^GCTTCG5", 0)] * 3,
                       [make_read("TCGAAGCTTCGAAGCT", "16M", "16", 0)] * 3, 4) == [("AGCTTCG", "A")]
    assert som_alleles([make_read("TCGAGTCGA", "4M1I4M", "8", 0)] * 3, NORMAL8, 3) == [("A", "AG")]
    assert som_alleles([make_read("TCGAGGTCTCGA", "4M4I4M", "8", 0)] * 3, NORMAL8, 3) == [("A", "AGGTC")]
    normal = [make_read("TCGAATCGATCGATCGA", "17M", "17", 10)] * 3
    tumor = [make_read("TCATCTCAAAAGAGATCGA", "2M2D1M2I2M4I2M2D6M", "2^GA5^TC6", 10)] * 3
    assert som_alleles(tumor, normal, 11) == [("CGA", "C")]
    assert som_alleles(tumor, normal, 14) == [("A", "ATC")]
    assert som_alleles(tumor, normal, 16) == [("C", "CAAAA")]
    assert som_alleles(tumor, normal, 18) == [("ATC", "A")]


# ---- AlleleEvidenceSuite.scala:22-60 -------------------------------------------------------------------------------------
def test_allele_evidence():
    reads = [make_read("TCGATCGA", "8M", "1A6", 1, alignment_quality=30),
             make_read("TCGATCGA", "8M", "1A6", 1, alignment_quality=30),
             make_read("TCGACCCTCGA", "4M3I4M", "1A6", 1, alignment_quality=60)]
    ev = orc.allele_evidence_at(batch(*reads), 0, 2, "A", "C", 0.5)
    assert ev["mean_mapping_quality"] == 40.0 and ev["median_mapping_quality"] == 30 and ev["median_mismatches_per_read"] == 1
    reads = [make_read("TAGATCGA", "8M", "8", 1, alignment_quality=30),
             make_read("TCGATCGA", "8M", "1A6", 1, alignment_quality=60),
             make_read("TAGACCCTCGA", "4M3I4M", "8", 1, alignment_quality=60)]
    ev = orc.allele_evidence_at(batch(*reads), 0, 2, "A", "C", 0.5)
    assert ev["mean_mapping_quality"] == 60.0 and ev["median_mapping_quality"] == 60 and ev["median_mismatches_per_read"] == 1
    reads = [make_read("TAGATCGA", "8M", "8", 1, alignment_quality=30),
             make_read("TAGATCGA", "8M", "8", 1, alignment_quality=60),
             make_read("TAGACCCTCGA", "4M3I4M", "8", 1, alignment_quality=60)]
    ev = orc.allele_evidence_at(batch(*reads), 0, 2, "A", "C", 0.5)
    assert all(math.isnan(ev[k]) for k in ("mean_mapping_quality", "median_mapping_quality", "median_mismatches_per_read"))


# ---- VariantSupportSuite.scala:55-108: number of distinct alleles per locus on a real BAM ----------------------------------
def distinct_alleles(b, contig, locus, ref=None):
    return len({(e["ref"], e["seq"]) for e in orc.pileup_at(b, contig, locus, ref).elements()})


def test_variant_support_distinct_alleles():
    g = load_golden("gatk_mini_bundle_extract")
    c = g.contig_names.index("20")
    allr = g.filtered(has_md=True).sorted()
    nodup = g.filtered(has_md=True, non_duplicate=True).sorted()
    assert [distinct_alleles(allr, c, p) for p in (10008951, 10006822, 10009053)] == [2, 2, 1]
    assert [distinct_alleles(allr, c, p, "N") for p in (1, 9999996, 10007174, 10260442)] == [0, 1, 2, 1]
    assert [distinct_alleles(nodup, c, p, "N") for p in (9999996, 10006822, 10008920, 10009053)] == [1, 2, 3, 1]


# ---- SlidingWindowSuite.scala ---------------------------------------------------------------------------------------------
def test_sliding_window_counts():  # :79-123  01222333210
    b = batch(make_read("TCGATCGA", "8M", "8", 1), make_read("CGATCGAT", "8M", "8", 2), make_read("TCG", "3M", "3", 5))
    loci, ca, _ = orc.visited_loci(b, None, [(0, 0, 11)], skip_empty=False)
    assert loci == list(range(11)) and ca == [0, 1, 2, 2, 2, 3, 3, 3, 2, 1, 0]


def test_advance_multiple_windows():  # :232-283
    r1 = batch(make_read("TCGATCGA", "8M", "8", 2), make_read("CGATCGAT", "8M", "8", 3), make_read("TCG", "3M", "3", 6))
    r2 = batch(make_read("TCGATCGA", "8M", "8", 5), make_read("CGATCGAT", "8M", "8", 80), make_read("TCG", "3M", "3", 100))
    loci, ca, cb = orc.visited_loci(r1, r2, [(0, 0, 3), (0, 60, 101)], skip_empty=True)
    assert loci == [2, 80, 81, 82, 83, 84, 85, 86, 87, 100]
    assert ca[0] > 0 and cb[0] == 0 and ca[1] == 0 and cb[1] > 0


def test_unsorted_reads_rejected():  # :54-58 "Regions must be sorted by start locus"
    b = batch(make_read("TCGATCGA", "8M", "8", 5), make_read("TCGATCGA", "8M", "8", 2))
    with pytest.raises(orc.OracleError) as e:
        orc.visited_loci(b, None, [(0, 0, 20)])
    assert e.value.code == abi.ERR_UNSORTED_READS


# ---- DistributedUtilSuite.scala -------------------------------------------------------------------------------------------
def fmt(parts, name="chrM"):
    return ",".join(f"{name}:{s}-{e}={t}" for (_, s, e, t) in parts)


def test_partition_loci_by_approximate_depth():  # DistributedUtilSuite.scala:76-93
    from guacamole_b200.reads import ReadBatch, make_read
    reads = ReadBatch.from_records([make_read("A" * n, f"{n}M", f"{n}", s, chr="chr1") for s, n in ((5, 1), (6, 1), (7, 1), (8, 1))])
    got = orc.partition_loci_by_approximate_depth(2, [(0, 0, 100)], 100, reads)
    assert fmt(got, "chr1") == "chr1:0-7=0,chr1:7-100=1"
    # every locus is assigned exactly once, tasks ascend along the loci (result.count == lociUsed.count, :249)
    from guacamole_b200 import synth
    b = synth.generate([("1", 30000), ("2", 20000)], depth=12, seed=5).to_read_batch()
    for tasks, acc in ((3, 7), (8, 250), (5, 1)):
        parts = orc.partition_loci_by_approximate_depth(tasks, [(0, 0, 29999), (1, 100, 19999)], acc, b)
        assert sum(e - s for _, s, e, _ in parts) == 29999 + 19899
        assert [t for *_, t in parts] == sorted(t for *_, t in parts) and parts[-1][3] < tasks
        ends = {}
        for c, s, e, _ in parts:
            assert s == ends.get(c, 0 if c == 0 else 100)
            ends[c] = e


def test_partition_loci_uniformly():  # :46-63
    assert fmt(orc.partition_loci_uniformly(4, [(0, 0, 16571)])) == "chrM:0-4143=0,chrM:4143-8286=1,chrM:8286-12428=2,chrM:12428-16571=3"
    assert fmt(orc.partition_loci_uniformly(3, [(0, 0, 10)])) == "chrM:0-3=0,chrM:3-7=1,chrM:7-10=2"
    assert fmt(orc.partition_loci_uniformly(4, [(0, 0, 3)])) == "chrM:0-1=0,chrM:1-2=1,chrM:2-3=2"
    assert orc.partition_loci_uniformly(4, [(0, 10, 10)]) == []
    p = orc.partition_loci_uniformly(100, [(0, 1000, 1100)])
    assert p == [(0, 1000 + i, 1001 + i, i) for i in range(100)]
    p = orc.partition_loci_uniformly(2, [(0, 0, 100), (1, 0, 100)])
    assert sum(e - s for (_, s, e, t) in p if t == 0) == 100 and sum(e - s for (_, s, e, t) in p if t == 1) == 100


def parts(n, loci):
    return orc.partition_loci_uniformly(n, loci)


def test_pileup_flatmap_skip_empty():  # :141-155  loci 1..8 with 5 tasks over 4 contigs
    b = ReadBatch.from_records([make_read("TCGATCGA", "8M", "8", 1)] * 3, contig_names=["chr0", "chr1", "chr2"])
    ranges = parts(5, [(0, 5, 10), (1, 0, 100), (2, 0, 1000), (2, 5000, 6000)])
    res = orc.pileup_counts(b, ranges, skip_empty=True)
    assert list(res.counts()["locus"]) == [1, 2, 3, 4, 5, 6, 7, 8]
    assert res.stats["loci_visited"] == 8


def test_pileup_flatmap_two_rdds_skip_empty():  # :157-179
    names = ["chr0", "chr1", "chr2"]
    r1 = ReadBatch.from_records([make_read("TCGATCGA", "8M", "8", 1)] * 3 + [make_read("GGGGGGGG", "8M", "8", 100)] * 3, contig_names=names)
    r2 = ReadBatch.from_records([make_read("AAAAAAAA", "8M", "8", 1), make_read("CCCCCCCC", "8M", "8", 1),
                                 make_read("TTTTTTTT", "8M", "8", 1), make_read("XXX", "3M", "8", 99)], contig_names=names)
    loci, _, _ = orc.visited_loci(r1, r2, [(1, 1, 500)], skip_empty=True)
    assert loci == [1, 2, 3, 4, 5, 6, 7, 8, 99, 100, 101, 102, 103, 104, 105, 106, 107]


def test_pileup_flatmap_no_skip_and_many_tasks():  # :102-139, 208-220: 800-way result == 1-way result
    b = batch(*[make_read("TCGATCGA", "8M", "8", 1)] * 4)
    res = orc.pileup_counts(b, parts(1, [(0, 1, 9)]), skip_empty=False)
    c = res.counts()
    assert len(c) == 8 and c["locus"][0] == 1 and chr(c["reference_base"][0]) == "T" and all(c["reference_depth"] == 4)
    one = orc.pileup_counts(b, parts(1, [(0, 0, 500)]), skip_empty=True).counts()
    many = orc.pileup_counts(b, parts(800, [(0, 0, 500)]), skip_empty=True, n_threads=4).counts()
    assert np.array_equal(one, many)


def test_threshold_through_pileup_flatmap():  # :320-374
    p = orc.threshold_params(0, False, False, skip_empty=False)
    ranges = parts(3, [(0, 1, 100)])
    assert orc.germline_threshold(batch(*[make_read("TCGATCGA", "8M", "8", 1)] * 3), ranges, p).threshold() == []
    het = batch(make_read("TCGATCGA", "8M", "8", 1), make_read("TCGGTCGA", "8M", "3A4", 1), make_read("TCGGTCGA", "8M", "3A4", 1))
    g = orc.germline_threshold(het, ranges, p).threshold()
    assert len(g) == 1 and (g[0]["start"], g[0]["ref"], g[0]["alt"], g[0]["gt"]) == (4, "A", "G", (REF, ALT))
    hom = batch(*[make_read("CCGATCGA", "8M", "0T7", 1)] * 3)
    g = orc.germline_threshold(hom, ranges, p).threshold()
    assert len(g) == 1 and (g[0]["start"], g[0]["ref"], g[0]["alt"], g[0]["gt"]) == (1, "T", "C", (ALT, ALT))


def test_window_fold_depths():  # :376-416
    b = batch(make_read("TCGATCGGC", "8M", "8", 0), make_read("CCCCCCCC", "8M", "8", 1),
              make_read("TCGATCGA", "8M", "8", 4), make_read("GGGGGGG", "7M", "7", 9))
    ranges = parts(5, [(0, 0, 20)])
    c = orc.pileup_counts(b, ranges, skip_empty=False).counts()
    sums = [int(c["depth"][(c["locus"] >= 4 * i) & (c["locus"] < 4 * i + 4)].sum()) for i in range(5)]
    assert sums == [7, 12, 8, 4, 0] and len(c) == 20


# ---- GermlineStandard.Caller.callVariantsAtLocus (commands/GermlineStandardCaller.scala:90-124) ---------------------------------
# The reference has no suite for this caller (parity unpinned end to end): these cases tie the oracle's restatement to the
# pieces that ARE pinned — Likelihood (LikelihoodSuite), AlleleEvidence (AlleleEvidenceSuite) — and to the caller's text.
def _std(reads, **kw):
    b = ReadBatch.from_records(reads).sorted()
    return b, orc.germline_standard(b, [(0, 0, 64)], **kw).called()


def test_germline_standard_het_and_hom():
    ref8 = make_read("TCGATCGA", "8M", "8", 0)
    alt8 = make_read("TCGGTCGA", "8M", "3A4", 0)
    b, got = _std([ref8] * 3 + [alt8] * 3)
    assert [(g["start"], g["ref"], g["alt"]) for g in got] == [(3, "A", "G")]      # het: one non-reference allele
    lk = orc.likelihoods_at(b, 0, 3, include_alignment=False, log_space=True, normalize=True)
    best = max(x["value"] for x in lk)
    ev = got[0]["evidence"]
    assert abs(ev["likelihood"] - math.exp(best)) < 1e-15
    assert (ev["read_depth"], ev["allele_read_depth"], ev["forward_depth"], ev["allele_forward_depth"]) == (6, 3, 6, 3)
    assert got[0]["phred"] == orc.lib().orc_success_probability_to_phred(C.c_double(ev["likelihood"] - 1e-10))
    # homozygous alternate: Genotype.getNonReferenceAlleles keeps both copies -> two equal records
    b, got = _std([alt8] * 3)
    assert [(g["start"], g["ref"], g["alt"]) for g in got] == [(3, "A", "G"), (3, "A", "G")]
    assert got[0] == got[1] and got[0]["evidence"]["likelihood"] == 1.0
    # all reference: nothing
    assert _std([ref8] * 4)[1] == []


def test_germline_standard_indels_and_filter():
    dele = make_read("TCGTCGA", "3M1D4M", "3^A4", 0)
    _, got = _std([dele] * 3)
    assert [(g["start"], g["ref"], g["alt"]) for g in got] == [(2, "GA", "G")] * 2 + [(3, "A", "")] * 2
    ins = make_read("TCGAGTCGA", "4M1I4M", "8", 0)
    _, got = _std([ins] * 3)
    assert [(g["start"], g["ref"], g["alt"]) for g in got] == [(3, "A", "AG")] * 2
    # QualityAlignedReadsFilter drops low-mapq reads from the likelihoods; the evidence still counts them (unfiltered pileup)
    ref8 = make_read("TCGATCGA", "8M", "8", 0)
    alt_lo = make_read("TCGGTCGA", "8M", "3A4", 0, alignment_quality=5)
    alt_hi = make_read("TCGGTCGA", "8M", "3A4", 0, alignment_quality=50)
    _, got = _std([ref8] * 2 + [alt_lo] * 4 + [alt_hi] * 2, min_mapq=10)
    assert [(g["start"], g["ref"], g["alt"]) for g in got] == [(3, "A", "G")]      # het over the 4 kept reads
    assert got[0]["evidence"]["read_depth"] == 8 and got[0]["evidence"]["allele_read_depth"] == 6
    _, got = _std([ref8] * 2 + [alt_lo] * 4 + [alt_hi] * 2, min_mapq=0)
    assert len(got) == 1 and got[0]["evidence"]["allele_read_depth"] == 6          # still het (2 good reference reads)
    _, got = _std([alt_lo] * 3, min_mapq=10)
    assert got == []                                                                # every element filtered: Seq.empty


# ---- VariantSupport.pileupToAlleleCounts / VAFHistogram (commands/VariantSupport.scala:110-118, VAFHistogram.scala:31-38, 185-194)
def test_variant_support_allele_counts():
    """VariantSupportSuite.scala:55-108.  The suite's `computedAlleleCounts.size` exhausts the iterator, so only the NUMBER
    of alleles per locus is really asserted there (test_variant_support_distinct_alleles); the per-allele counts it lists
    are checked here where they agree with that (10008920, 10007174)."""
    g = load_golden("gatk_mini_bundle_extract")
    c = g.contig_names.index("20")
    allr = g.filtered(has_md=True).sorted()
    nodup = g.filtered(has_md=True, non_duplicate=True).sorted()
    by_alt = lambda b, p: {x["alt"]: x["count"] for x in orc.allele_counts(b, [(c, p, p + 1)]).allele_counts()}
    assert by_alt(nodup, 10008920) == {"C": 2, "CA": 1, "CAA": 1}
    assert by_alt(allr, 10007174) == {"T": 5, "C": 3}
    assert by_alt(allr, 1) == {}
    r = orc.allele_counts(nodup, [(c, 10006822, 10006823)]).allele_counts()
    assert [(x["ref"], x["alt"], x["count"]) for x in r] == [("C", "", 3), ("C", "C", 1)]   # mid-deletion + match
    # a range: rows sorted by (locus, ref, alt), counts add up to the depth
    r = orc.allele_counts(nodup, [(c, 10008900, 10008960)])
    rows = r.allele_counts()
    depth = {int(x["locus"]): int(x["depth"]) for x in orc.pileup_counts(nodup, [(c, 10008900, 10008960)]).counts()}
    total = {}
    for x in rows:
        total[x["start"]] = total.get(x["start"], 0) + x["count"]
    assert total == depth
    assert rows == sorted(rows, key=lambda x: (x["start"], x["ref"], x["alt"]))


def test_vaf_histogram_suite():  # VAFHistogramSuite.scala:8-42
    from guacamole_b200.callers import generate_vaf_histogram, variant_loci, COUNTS_DTYPE
    vafs = [0.25, 0.35, 0.4, 0.5, 0.55]
    assert generate_vaf_histogram(vafs, 10) == {20: 1, 30: 1, 40: 1, 50: 2}
    assert generate_vaf_histogram(vafs, 20) == {25: 1, 35: 1, 40: 1, 50: 1, 55: 1}
    assert generate_vaf_histogram(vafs, 100) == {25: 1, 35: 1, 40: 1, 50: 1, 55: 1}
    with pytest.raises(ValueError):
        generate_vaf_histogram(vafs, 0)
    # VariantLocus.apply + the two filters of variantLociFromReads on hand-made count rows
    rows = np.zeros(4, COUNTS_DTYPE)
    rows["locus"] = [10, 11, 12, 13]
    rows["depth"] = [10, 10, 4, 20]
    rows["reference_depth"] = [10, 7, 1, 19]
    got = variant_loci(rows)
    assert got["locus"].tolist() == [11, 12, 13]
    assert np.allclose(got["variant_allele_frequency"], [0.3, 0.75, 0.05])
    assert variant_loci(rows, min_read_depth=5)["locus"].tolist() == [11, 13]
    assert variant_loci(rows, min_variant_allele_frequency=30)["locus"].tolist() == [11, 12]   # 0.3f >= 0.30 as doubles


def test_vcf_writer_on_chrm(tmp_path):
    """"identical VCF" = identical field sets after canonical sort.  The writer turns records into sorted lines
    (POS = start + 1, GT from the GenotypeAllele pair); here on the oracle's 138 chrM records (config 1)."""
    from guacamole_b200 import vcf
    b = load_golden("chrM.sorted").filtered(non_duplicate=True, has_md=True).sorted()
    recs = orc.germline_threshold(b, [(0, 0, 16570)]).threshold()
    lines = vcf.vcf_lines(recs, b.contig_names, b.sample_names)
    assert len(recs) == 138 and len(lines) == 138
    gts = [ln.split("\t")[-1] for ln in lines]
    assert gts.count("0/1") == 120 and gts.count("1/1") == 18
    first = lines[0].split("\t")
    assert first[0] == b.contig_names[0] and int(first[1]) == recs[0]["start"] + 1 and first[3] == recs[0]["ref"]
    assert [int(ln.split("\t")[1]) for ln in lines] == sorted(int(ln.split("\t")[1]) for ln in lines)
    n = vcf.write_vcf(str(tmp_path / "chrM.vcf"), recs, b.contig_names, b.sample_names, b.contig_lengths)
    text = open(tmp_path / "chrM.vcf").read().splitlines()
    assert n == 138 and text[0] == "##fileformat=VCFv4.2" and text[-1] == lines[-1]
    # called alleles carry GQ / DP / AD; an empty alternate is spelled <DEL>
    dele = ReadBatch.from_records([make_read("TCGTCGA", "3M1D4M", "3^A4", 0)] * 3).sorted()
    cl = vcf.vcf_lines([dict(g, gt=(0, 1)) for g in orc.germline_standard(dele, [(0, 0, 64)]).called()], dele.contig_names, dele.sample_names)
    assert [ln.split("\t")[1:5] for ln in cl] == [["3", ".", "GA", "G"], ["4", ".", "A", "*"]]
    assert cl[0].split("\t")[8] == "GT:GQ:DP:AD" and cl[0].split("\t")[9] == "0/1:100:3:0,3"


def test_vcf_lines_for_somatic_records():
    """calledSomaticAlleleToADAMGenotype (AlleleConversions.scala:47-62): GT Ref/Alt, GQ = phredScaledSomaticLikelihood,
    DP / AD from the tumor evidence."""
    from guacamole_b200 import vcf
    tumor = ReadBatch.from_records([make_read("TCGGTCGA", "8M", "3A4", 0)] * 3).sorted()
    normal = ReadBatch.from_records([make_read("TCGATCGA", "8M", "8", 0)] * 3).sorted()
    recs = orc.somatic_standard(tumor, normal, [(0, 0, 64)], orc.somatic_params(odds=2)).somatic()
    assert [(r["start"], r["ref"], r["alt"]) for r in recs] == [(3, "A", "G")]
    lines = vcf.vcf_lines([dict(r, gt=(abi.GT_REF, abi.GT_ALT)) for r in recs], tumor.contig_names, tumor.sample_names)
    f = lines[0].split("\t")
    assert f[1:5] == ["4", ".", "A", "G"] and f[8] == "GT:GQ:DP:AD"
    gt, gq, dp, ad = f[9].split(":")
    assert gt == "0/1" and int(gq) == recs[0]["phred"] and dp == "3" and ad == "0,3"
