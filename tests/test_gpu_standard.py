"""Parity of the germline-standard kernels (K_standard, K_standard_exact, K_standard_evidence) with the oracle, through the
C ABI.  Integers and allele strings bit-exact; fp64 likelihoods / means within 1e-9 relative (+ 1e-12 absolute,
TestUtil.assertAlmostEqual's epsilon): the per-allele sums are taken in another order than colt's last-to-first walk."""
import math

import numpy as np
import pytest

import oracle_binding as orc
from conftest import load_golden
from guacamole_b200.reads import ReadBatch, make_read

pytestmark = pytest.mark.gpu
REL, ABS = 1e-9, 1e-12


@pytest.fixture(scope="module")
def ctx():
    from guacamole_b200.callers import Context
    c = Context(0)
    yield c
    c.close()


def close(a, b):
    if isinstance(a, float) or isinstance(b, float):
        if math.isnan(a) or math.isnan(b):
            return math.isnan(a) and math.isnan(b)
        if math.isinf(a) or math.isinf(b):
            return a == b
        return abs(a - b) <= REL * max(abs(a), abs(b)) + ABS
    return a == b


def assert_standard_equal(ctx, batch, ranges, min_mapq=1, max_tied=0):
    from guacamole_b200 import callers
    want = orc.germline_standard(batch, ranges, min_mapq=min_mapq)
    reads = ctx.pack(batch)
    got = callers.germline_standard(ctx, reads, ranges, min_alignment_quality=min_mapq)
    reads.free()
    w, g = want.called(), got.genotypes()
    if max_tied:  # exact likelihood ties between genotypes follow rounding noise
        gw = {}
        for x in w:
            gw.setdefault((x["contig"], x["start"]), []).append(x)
        tied = set()
        for x in g:
            ys = gw.get((x["contig"], x["start"]))
            if ys and (x["ref"], x["alt"]) not in [(y["ref"], y["alt"]) for y in ys] and \
                    close(x["evidence"]["likelihood"], ys[0]["evidence"]["likelihood"]):
                tied.add((x["contig"], x["start"]))
        assert len(tied) <= max_tied, len(tied)
        g = [x for x in g if (x["contig"], x["start"]) not in tied]
        w = [x for x in w if (x["contig"], x["start"]) not in tied]
    key = lambda x: (x["contig"], x["start"], x["ref"], x["alt"])
    assert [key(x) for x in g] == [key(x) for x in w]
    for a, b in zip(g, w):
        assert a["phred"] == b["phred"], (a, b)
        for k, v in b["evidence"].items():
            assert close(a["evidence"][k], v), (k, a, b)
    assert got.stats["loci_visited"] == want.stats["loci_visited"]
    return got


def test_unit_cases(ctx):
    ref8 = make_read("TCGATCGA", "8M", "8", 0)
    alt8 = make_read("TCGGTCGA", "8M", "3A4", 0)
    cases = [
        [ref8] * 3 + [alt8] * 3,                                         # het SNV
        [alt8] * 3,                                                      # hom alt: two equal records
        [ref8] * 4,                                                      # nothing
        [make_read("TCGTCGA", "3M1D4M", "3^A4", 0)] * 3,                 # deletion: anchor + mid-deletion loci
        [make_read("TCGAGTCGA", "4M1I4M", "8", 0)] * 3 + [ref8] * 2,     # insertion, het
        [make_read("TCATCTCAAAAGAGATCGA", "2M2D1M2I2M4I2M2D6M", "2^GA5^TC6", 10)] * 3,
        [ref8] * 2 + [make_read("TCGGTCGA", "8M", "3A4", 0, alignment_quality=5)] * 4 + [alt8] * 2,
        [make_read("TCGNTCGA", "8M", "3A4", 0)] * 2 + [ref8],           # non-ACGT allele drops out of the enumeration
    ]
    for reads in cases:
        b = ReadBatch.from_records(reads).sorted()
        for mq in (0, 1, 10):
            assert_standard_equal(ctx, b, [(0, 0, 64)], min_mapq=mq)
    got = assert_standard_equal(ctx, ReadBatch.from_records(cases[1]).sorted(), [(0, 0, 64)]).genotypes()
    assert [(x["start"], x["ref"], x["alt"]) for x in got] == [(3, "A", "G")] * 2


@pytest.mark.parametrize("name,contig", [("tumor.chr20.tough", "20"), ("normal.chr20.tough", "20"),
                                         ("synthetic.challenge.set1.tumor.v2.withMDTags.chr2.complexvar", "2")])
def test_real_fixtures(ctx, name, contig):
    b = load_golden(name).filtered(non_duplicate=True, has_md=True).sorted()
    c = b.contig_names.index(contig)
    hi = int(b.end().max()) + 10
    got = assert_standard_equal(ctx, b, [(c, 0, hi)], min_mapq=1, max_tied=3)
    assert len(got) > 5
    assert_standard_equal(ctx, b, [(c, 0, hi)], min_mapq=30, max_tied=3)


def test_chrm(ctx):  # the reference's own CPU-runnable input (BASELINE.json configs[0]) through the third caller
    b = load_golden("chrM.sorted").filtered(non_duplicate=True, has_md=True).sorted()
    got = assert_standard_equal(ctx, b, [(0, 0, 16570)], min_mapq=1, max_tied=3)
    assert len(got) > 20


def test_synthetic_and_extremes(ctx):
    from guacamole_b200 import synth
    contigs = [("20", 150000)]
    b = synth.generate(contigs, depth=30, seed=77, sample=0).to_read_batch()
    got = assert_standard_equal(ctx, b, [(0, 0, 149999)], min_mapq=1)
    assert len(got) > 50
    assert_standard_equal(ctx, b, [(0, 1000, 60000), (0, 90000, 90500)], min_mapq=20)
    # base qualities 0..3 and > 63, mapq 0 / 255 (see test_gpu_somatic.test_quality_extremes)
    rng = np.random.default_rng(9)
    q = b.qual.copy()
    hit = rng.random(q.shape[0]) < 0.3
    q[hit] = rng.choice(np.array([0, 1, 2, 3, 5, 9, 17, 33, 62, 63], dtype=np.uint8), size=int(hit.sum()))
    for i in np.nonzero(rng.random(len(b)) < 0.04)[0]:
        q[int(b.seq_off[i]):int(b.seq_off[i + 1])] = rng.choice(np.array([64, 70, 93, 127], dtype=np.uint8))
    b.qual = q
    m = b.mapq.copy()
    hit = rng.random(m.shape[0]) < 0.3
    m[hit] = rng.choice(np.array([0, 1, 2, 7, 29, 30, 254, 255], dtype=np.uint8), size=int(hit.sum()))
    b.mapq = m
    # (ten discrete quality values over 150,000 loci: two single-read alternates of equal quality are common -> exact ties)
    assert_standard_equal(ctx, b, [(0, 0, 149999)], min_mapq=0, max_tied=300)
    assert_standard_equal(ctx, b, [(0, 0, 149999)], min_mapq=30, max_tied=300)
