"""Parity of the somatic-standard kernels (K_somatic, K_somatic_exact, K_evidence) with the oracle, through the C ABI.
Integers and allele strings bit-exact; fp64 likelihoods / odds / means within 1e-9 relative
(summation order differs from colt's last-to-first walk); quantities obtained by cancellation next to 1
(log of odds ~ 1, 1 - total) get an absolute floor of 1e-12, the reference's own assertAlmostEqual epsilon."""
import math

import numpy as np
import pytest

import oracle_binding as orc
from conftest import load_golden
from guacamole_b200.reads import ReadBatch, make_read

pytestmark = pytest.mark.gpu
REL = 1e-9
ABS = 1e-12   # TestUtil.assertAlmostEqual's epsilon (src/test/.../util/TestUtil.scala:204-206)


@pytest.fixture(scope="module")
def ctx():
    from guacamole_b200.callers import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module", params=[1, 0], ids=["rows", "read-walk"], autouse=True)
def row_store(request, ctx):
    """Every somatic parity test runs twice: over the per-word row stores built at pack time (the default), and with the
    likelihood kernel walking every word's candidate reads itself (stores packed without rows)."""
    from guacamole_b200 import abi
    ctx.set_option(abi.OPT_DIFFERENCE_LISTS, request.param)
    yield request.param
    ctx.set_option(abi.OPT_DIFFERENCE_LISTS, 1)


def close(a, b):
    if isinstance(a, float) or isinstance(b, float):
        if math.isnan(a) or math.isnan(b):
            return math.isnan(a) and math.isnan(b)
        if math.isinf(a) or math.isinf(b):
            return a == b
        return abs(a - b) <= REL * max(abs(a), abs(b)) + ABS
    return a == b


def drop_tied_loci(g, w):
    """Loci where two tumor genotypes are exactly as likely (say het(ref, A) and het(ref, G) from one A and one G element of the
    same quality): the reference's maxBy then follows the rounding noise of its per-element sums (terms log(s + (1 - s)),
    zero up to an ulp), which no other summation order reproduces.  Such a locus is recognised by both sides
    emitting a record there with the same tumor likelihood but another alternate; it is taken out of the comparison."""
    gw = {(x["contig"], x["start"]): x for x in w}
    tied = set()
    for x in g:
        y = gw.get((x["contig"], x["start"]))
        if y is not None and (x["ref"], x["alt"]) != (y["ref"], y["alt"]) and close(x["tumor"]["likelihood"], y["tumor"]["likelihood"]):
            tied.add((x["contig"], x["start"]))
    keep = lambda rs: [x for x in rs if (x["contig"], x["start"]) not in tied]
    return keep(g), keep(w), len(tied)


def assert_somatic_equal(ctx, tumor, normal, ranges, max_tied=0, **kw):
    from guacamole_b200 import callers
    want = orc.somatic_standard(tumor, normal, ranges, orc.somatic_params(**kw))
    t, n = ctx.pack(tumor), ctx.pack(normal)
    got = callers.somatic_standard(ctx, t, n, ranges, odds_threshold=kw.get("odds", 20), min_alignment_quality=kw.get("min_mapq", 1),
                                   filter_multi_allelic=kw.get("filter_multi_allelic", False),
                                   max_read_depth=kw.get("max_read_depth", 2 ** 31 - 1))
    t.free()
    n.free()
    w, g = want.somatic(), got.genotypes()
    if max_tied:
        g, w, n_tied = drop_tied_loci(g, w)
        assert n_tied <= max_tied, n_tied
    assert [(x["contig"], x["start"], x["ref"], x["alt"]) for x in g] == [(x["contig"], x["start"], x["ref"], x["alt"]) for x in w]
    for a, b in zip(g, w):
        assert a["phred"] == b["phred"], (a, b)
        assert close(a["somatic_log_odds"], b["somatic_log_odds"]), (a, b)
        for side in ("tumor", "normal"):
            for k, v in b[side].items():
                assert close(a[side][k], v), (side, k, a, b)
    assert got.stats["loci_visited"] == want.stats["loci_visited"]
    return got


NORMAL8 = [make_read("TCGATCGA", "8M", "8", 0)] * 3


def pair(tumor, normal):
    return ReadBatch.from_records(tumor).sorted(), ReadBatch.from_records(normal).sorted()


def test_suite_indels(ctx):  # SomaticStandardCallerSuite.scala:117-262 through the engine
    cases = [
        ([make_read("TCGGTCGA", "8M", "3G4", 0)] * 3, NORMAL8),
        ([make_read("TCGTCGA", "3M1D4M", "3^A4", 0)] * 3, NORMAL8),
        ([make_read("TCGAAAAGCT", "5M6D5M", "5^GCTTCG5", 0)] * 3, [make_read("TCGAAGCTTCGAAGCT", "16M", "16", 0)] * 3),
        ([make_read("TCGAGTCGA", "4M1I4M", "8", 0)] * 3, NORMAL8),
        ([make_read("TCGAGGTCTCGA", "4M4I4M", "8", 0)] * 3, NORMAL8),
        ([make_read("TCATCTCAAAAGAGATCGA", "2M2D1M2I2M4I2M2D6M", "2^GA5^TC6", 10)] * 3, [make_read("TCGAATCGATCGATCGA", "17M", "17", 10)] * 3),
    ]
    for tumor, normal in cases:
        t, n = pair(tumor, normal)
        got = assert_somatic_equal(ctx, t, n, [(0, 0, 64)], odds=2)
        assert_somatic_equal(ctx, t, n, [(0, 0, 64)], odds=2, filter_multi_allelic=True)
    t, n = pair(*cases[1])
    g = assert_somatic_equal(ctx, t, n, [(0, 0, 64)], odds=2).genotypes()
    assert [(x["start"], x["ref"], x["alt"]) for x in g] == [(2, "GA", "G")]


def tn(tumor_name, normal_name):
    t = load_golden(tumor_name).filtered(non_duplicate=True, passed_qc=True, has_md=True).sorted()
    n = load_golden(normal_name).filtered(non_duplicate=True, passed_qc=True, has_md=True).sorted()
    return t, n


FIXTURES = [("tumor.chr20.tough", "normal.chr20.tough", "20"),
            ("synthetic.challenge.set1.tumor.v2.withMDTags.chr2.syn1fp", "synthetic.challenge.set1.normal.v2.withMDTags.chr2.syn1fp", "2"),
            ("synthetic.challenge.set1.tumor.v2.withMDTags.chr2.complexvar", "synthetic.challenge.set1.normal.v2.withMDTags.chr2.complexvar", "2"),
            ("tumor.chr20.simplefp", "normal.chr20.simplefp", "20")]


@pytest.mark.parametrize("tumor_name,normal_name,contig", FIXTURES)
def test_real_fixtures(ctx, tumor_name, normal_name, contig):
    t, n = tn(tumor_name, normal_name)
    c = t.contig_names.index(contig)
    hi = int(max(t.end().max(), n.end().max())) + 10
    got = assert_somatic_equal(ctx, t, n, [(c, 0, hi)], odds=20)
    assert_somatic_equal(ctx, t, n, [(c, 0, hi)], odds=120)
    assert_somatic_equal(ctx, t, n, [(c, 0, hi)], odds=20, min_mapq=30, max_read_depth=60)
    if "tough" in tumor_name:
        assert len(got) > 20


def test_synthetic_pair(ctx):  # BASELINE.json configs[2] shape, small
    from guacamole_b200 import synth
    contigs = [("20", 200000)]
    tumor = synth.generate(contigs, depth=60, seed=21, sample=1).to_read_batch()
    normal = synth.generate(contigs, depth=30, seed=21, sample=0).to_read_batch()
    got = assert_somatic_equal(ctx, tumor, normal, [(0, 0, 199999)], odds=20)
    assert len(got) >= 1
    assert_somatic_equal(ctx, tumor, normal, [(0, 5000, 90000), (0, 120000, 120777)], odds=20, filter_multi_allelic=True)


def test_deep_underflow_flow(ctx):  # the reference's naive normalisation underflows at ~1,100x; Inf / NaN must flow alike
    from guacamole_b200 import synth
    contigs = [("amp", 2000)]
    tumor = synth.generate(contigs, depth=2500, seed=33, sample=1).to_read_batch()
    normal = synth.generate(contigs, depth=2500, seed=33, sample=0).to_read_batch()
    assert_somatic_equal(ctx, tumor, normal, [(0, 0, 1999)], odds=20)


@pytest.mark.parametrize("seed", [5, 6])
def test_quality_extremes(ctx, seed):
    """Base qualities 0..3 (success probability < 1/2: log((1-s) + (1-s)) > 0, log(s + s) = -inf at 0), qualities > 63 (the
    reads leave the one-byte-per-element path), mapq 0 / 255 (table rows of -inf): the early outs of the tumor half (hom-ref
    lead test, bounds over indel elements) must agree with the literal enumeration the oracle does, NaN / Inf included."""
    from guacamole_b200 import synth
    contigs = [("q", 30000)]
    rng = np.random.default_rng(seed)
    tumor = synth.generate(contigs, depth=40, seed=100 + seed, sample=1).to_read_batch()
    normal = synth.generate(contigs, depth=25, seed=100 + seed, sample=0).to_read_batch()
    for b, frac in ((tumor, 0.35), (normal, 0.25)):
        q = b.qual.copy()
        hit = rng.random(q.shape[0]) < frac
        q[hit] = rng.choice(np.array([0, 1, 2, 3, 5, 9, 17, 33, 62, 63], dtype=np.uint8), size=int(hit.sum()))
        # whole reads of wide qualities (a few), so that the general path sees enough elements
        wide = rng.random(len(b)) < 0.04
        for i in np.nonzero(wide)[0]:
            q[int(b.seq_off[i]):int(b.seq_off[i + 1])] = rng.choice(np.array([64, 70, 93, 127], dtype=np.uint8))
        b.qual = q
        m = b.mapq.copy()
        hit = rng.random(m.shape[0]) < 0.3
        m[hit] = rng.choice(np.array([0, 1, 2, 7, 29, 30, 254, 255], dtype=np.uint8), size=int(hit.sum()))
        b.mapq = m
    got = assert_somatic_equal(ctx, tumor, normal, [(0, 0, 29999)], max_tied=30, odds=1, min_mapq=0)
    assert len(got) > 5000
    assert_somatic_equal(ctx, tumor, normal, [(0, 0, 29999)], max_tied=30, odds=20, min_mapq=1)
    assert_somatic_equal(ctx, tumor, normal, [(0, 0, 29999)], max_tied=30, odds=20, min_mapq=30, max_read_depth=45)


def test_amplicon_shape_config5(ctx):
    """BASELINE.json configs[4] at test size: 10,000x tumor and normal over an amplicon.  Every likelihood underflows (SURVEY
    H4), the per-locus element counts pass 8 bits by far, and a call's supporting elements run into the thousands (the
    AlleleEvidence medians are taken over all of them)."""
    from guacamole_b200 import synth
    contigs = [("amp", 520)]
    tumor = synth.generate(contigs, depth=10000, seed=55, sample=1).to_read_batch()
    normal = synth.generate(contigs, depth=10000, seed=56, sample=0).to_read_batch()
    assert_somatic_equal(ctx, tumor, normal, [(0, 0, 519)], odds=20)
    # a call whose supporting elements run into the thousands: medians / means over all 1,500 of them
    alt = [make_read("TCGGTCGA", "8M", "3A4", 0, quality_scores=[10 + (i * 7) % 50] * 8, alignment_quality=20 + (i * 11) % 41)
           for i in range(1500)]
    ref = [make_read("TCGATCGA", "8M", "8", 0, quality_scores=[12 + (i * 5) % 45] * 8, alignment_quality=25 + (i * 3) % 36)
           for i in range(1500)]
    t, n = pair(alt, ref)
    got = assert_somatic_equal(ctx, t, n, [(0, 0, 16)], odds=2).genotypes()
    assert [(x["start"], x["ref"], x["alt"], x["tumor"]["allele_read_depth"]) for x in got] == [(3, "A", "G", 1500)]
    # the same reads through germline-threshold (16-bit counter fields) and per-locus counts
    from guacamole_b200 import callers
    for b in (tumor,):
        want = orc.germline_threshold(b, [(0, 0, 519)]).threshold()
        reads = ctx.pack(b)
        g = callers.germline_threshold(ctx, reads, [(0, 0, 519)]).genotypes()
        wc = orc.pileup_counts(b, [(0, 0, 519)]).counts()
        gc = callers.pileup_counts(ctx, reads, [(0, 0, 519)]).records
        reads.free()
        key = lambda x: (x["contig"], x["start"], x["ref"], x["alt"], x["gt"])
        assert [key(x) for x in g] == [key(x) for x in want]
        assert gc["depth"].tolist() == wc["depth"].tolist() and gc["reference_depth"].tolist() == wc["reference_depth"].tolist()
        assert int(gc["depth"].max()) > 9000


# ---- SomaticStandardCallerSuite.scala:82-115 through the engine, genotype filters applied on the device -----------------------
SUITE_FILTERS = dict(min_tumor_read_depth=8, max_tumor_read_depth=200, min_normal_read_depth=4, min_tumor_alternate_read_depth=3,
                     min_lod=120, min_vaf=5, min_likelihood=70, seq_overload=True)
SUITE_LOCI = [
    ("tumor.chr20.tough", "normal.chr20.tough", "20",
     [42999694, 25031215, 44061033, 45175149, 755754, 1843813, 3555766, 3868620, 9896926, 14017900, 17054263, 35951019, 50472935,
      51858471, 58201903, 7087895, 19772181, 30430960, 32150541, 42186626, 44973412, 46814443, 52311925, 53774355, 57280858, 62262870], []),
    ("synthetic.challenge.set1.tumor.v2.withMDTags.chr2.syn1fp", "synthetic.challenge.set1.normal.v2.withMDTags.chr2.syn1fp", "2",
     [], [216094721, 3529313, 8789794, 104043280, 104175801, 126651101, 241901237, 57270796, 120757852]),
    ("synthetic.challenge.set1.tumor.v2.withMDTags.chr2.complexvar", "synthetic.challenge.set1.normal.v2.withMDTags.chr2.complexvar", "2",
     [82949713, 130919744], [148487667, 134307261, 90376213, 3638733, 109347468]),
    ("tumor.chr20.simplefp", "normal.chr20.simplefp", "20", [], [26211835, 29652479, 54495768, 13046318, 25939088]),
]


@pytest.mark.parametrize("tumor_name,normal_name,contig,positive,negative", SUITE_LOCI)
def test_suite_decisions_with_device_filters(ctx, tumor_name, normal_name, contig, positive, negative):
    """The reference's 28 "variant found" and 19 "no variant" loci: findPotentialVariantAtLocus(odds 120) followed by
    SomaticGenotypeFilter(Seq(...), 8, 200, 4, 3, 120, 5, 70), the filters running in the engine's epilogue
    (guac_somatic_standard_filtered); the surviving records equal the host-side filter over the unfiltered call."""
    from guacamole_b200 import callers
    t, n = tn(tumor_name, normal_name)
    c = t.contig_names.index(contig)
    loci = sorted(positive + negative)
    ranges = [(c, p, p + 1) for p in loci]
    rt, rn = ctx.pack(t), ctx.pack(n)
    kept = callers.somatic_standard(ctx, rt, rn, ranges, odds_threshold=120, min_alignment_quality=1, filters=SUITE_FILTERS)
    everything = callers.somatic_standard(ctx, rt, rn, ranges, odds_threshold=120, min_alignment_quality=1)
    rt.free()
    rn.free()
    found = {int(r["start"]) for r in kept.records}
    assert found == set(positive), (sorted(found), sorted(positive))
    mask = callers.somatic_genotype_filter(everything.records, **SUITE_FILTERS)
    assert [g["start"] for g in kept.genotypes()] == [g["start"] for g, m in zip(everything.genotypes(), mask) if m]
    assert len(kept) <= len(everything)


def test_compact_batch_somatic(ctx):
    """guac_reads_pack_v2 (4-bit bases, 32-bit columns) gives the store of guac_reads_pack: byte-identical somatic records."""
    from guacamole_b200 import callers
    t, n = tn(*FIXTURES[0][:2])
    c = t.contig_names.index(FIXTURES[0][2])
    hi = int(max(t.end().max(), n.end().max())) + 10
    wide = [ctx.pack(t), ctx.pack(n)]
    cbs = [callers.CompactBatch(t, fixed_length=False), callers.CompactBatch(n)]
    compact = [ctx.pack_v2(cbs[0], t.contig_names, t.sample_names), ctx.pack_v2(cbs[1], n.contig_names, n.sample_names)]
    a = callers.somatic_standard(ctx, wide[0], wide[1], [(c, 0, hi)], odds_threshold=20)
    b = callers.somatic_standard(ctx, compact[0], compact[1], [(c, 0, hi)], odds_threshold=20)
    assert len(a) > 20 and repr(a.genotypes()) == repr(b.genotypes())  # (repr: NaN evidence fields compare equal)
    for r in wide + compact:
        r.free()
