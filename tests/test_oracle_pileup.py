"""Pins the oracle's CIGAR walk / alignment classification / quality rules to the known-answer vectors of the
reference's PileupSuite (src/test/scala/org/hammerlab/guacamole/pileup/PileupSuite.scala; line numbers cited per
test) and MDTagUtilsSuite."""
import pytest

import oracle_binding as orc
from conftest import load_golden
from guacamole_b200.reads import ReadBatch, make_read


def batch(*reads):
    return ReadBatch.from_records(list(reads))


def elems(b, locus, ref=None, contig=0):
    return orc.pileup_at(b, contig, locus, ref).elements()


LONG_INSERT = [make_read("TCGATCGA", "8M", "8", 1), make_read("TCGATCGA", "8M", "8", 1),
               make_read("TCGACCCTCGA", "4M3I4M", "8", 1)]
Q = [10, 15, 20, 25, 10, 15, 20, 25]
LONG_INSERT_Q = [make_read("TCGATCGA", "8M", "8", 1, "chr1", Q), make_read("TCGATCGA", "8M", "8", 1, "chr1", Q),
                 make_read("TCGACCCTCGA", "4M3I4M", "8", 1, "chr1", [10, 15, 20, 25, 5, 5, 5, 10, 15, 20, 25])]


def test_long_insert_reads():  # PileupSuite.scala:52-71
    b = batch(*LONG_INSERT)
    assert elems(b, 0) == []
    first = elems(b, 1)
    assert all(e["kind"] == "Match" and e["quality"] == 31 for e in first)
    ins = elems(b, 4)
    assert [e["kind"] for e in ins] == ["Match", "Match", "Insertion"]
    assert all(e["quality"] == 31 for e in ins)
    assert ins[0]["seq"] == "A" and ins[2]["seq"] == "ACCC" and ins[2]["ref"] == "A"


def test_insert_qualities():  # :73-89, :91-107, :118-133
    b = batch(*LONG_INSERT_Q)
    ins = elems(b, 4)
    assert [e["quality"] for e in ins] == [25, 25, 5]
    past = elems(b, 5)
    assert all(e["kind"] == "Match" and e["quality"] == 10 for e in past)
    last = elems(b, 8)
    assert all(e["kind"] == "Match" and e["seq"] == "A" and e["quality"] == 25 for e in last)
    at7 = elems(batch(*LONG_INSERT), 7)
    assert all(e["kind"] == "Match" and e["seq"] == "G" for e in at7)


def test_same_start_reads():  # :135-144, 222-244
    b = load_golden("same_start_reads")
    assert len(elems(b, 0)) == 10
    for i in range(1, 60):
        assert len(elems(b, i, "N")) == 10
    dels = [e for e in elems(b, 9, "A") if e["kind"] == "Deletion"]
    assert len(dels) == 5 and all(e["ref"] == "AAAAAAAAAAA" for e in dels)
    for i in range(10, 20):
        assert sum(e["kind"] == "MidDeletion" for e in elems(b, i, "N")) == 5
    for i in range(60, 70):
        assert len(elems(b, i, "N")) == 5


def test_element_creation():  # :146-174
    b = batch(make_read("AATTG", "5M", "5", 0))
    for i in range(3):
        e = elems(b, i)[0]
        assert e["kind"] == "Match" and e["index_within"] == i
    b = batch(make_read("AAATTT", "3M3M", "6", 0))
    e = elems(b, 3)[0]
    assert e["kind"] == "Match" and e["index_within"] == 0
    e = elems(b, 4)[0]
    assert e["kind"] == "Match" and e["index_within"] == 1


def test_contig_start_insertion():  # :176-180
    b = batch(make_read("AAAAAACGT", "5I4M", "4", 0))
    e = elems(b, 0)[0]
    assert e["kind"] == "Insertion" and e["seq"] == "AAAAAA" and e["quality"] == 31 and e["ref"] == "A"
    # subsequent loci walk off the insertion onto the match
    e = elems(b, 1)[0]
    assert e["kind"] == "Match" and e["seq"] == "C" and e["read_position"] == 6


def test_deletion_elements():  # :196-220
    b = batch(make_read("AATTGAATTG", "5M1D5M", "5^C5", 0))
    e = elems(b, 0)[0]
    assert e["kind"] == "Match" and e["index_within"] == 0
    e = elems(b, 4)[0]
    assert e["kind"] == "Deletion" and e["ref"] == "GC" and e["seq"] == "G" and e["index_within"] == 4
    e = elems(b, 5)[0]
    assert e["kind"] == "MidDeletion" and e["index_within"] == 0 and e["seq"] == "" and e["ref"] == "C"
    e = elems(b, 6)[0]
    assert e["kind"] == "Match" and e["index_within"] == 0
    e = elems(b, 9)[0]
    assert e["kind"] == "Match" and e["index_within"] == 3


def test_different_start_reads():  # :246-344  (29M10D31M, 10M10I10D40M, 5M4=1X5=)
    b = load_golden("different_start_reads")
    r1 = b.select([0])
    for bad in (0, 4, 75):
        assert elems(r1, bad) == []          # no overlap: the reference asserts; Pileup(...) filters
    assert elems(r1, 5)[0]["seq"] == "A"
    assert len(elems(r1, 74)) == 1
    e = elems(r1, 5 + 28)[0]
    assert e["kind"] == "Deletion" and e["ref"] == "AGGGGGGGGGG"
    assert elems(r1, 5 + 29)[0]["seq"] == "" and elems(r1, 5 + 38)[0]["seq"] == ""
    assert elems(r1, 5 + 39)[0]["seq"] == "A"
    r3 = b.select([2])
    assert [elems(r3, x)[0]["seq"] for x in (15, 16, 17, 18)] == ["A", "T", "C", "G"]
    r4 = b.select([3])
    for i in range(2):
        assert [elems(r4, 20 + i * 4 + k)[0]["seq"][0] for k in range(4)] == ["A", "C", "G", "T"]
    e = elems(r4, 29)[0]
    assert e["kind"] == "Insertion" and e["seq"] == "CGTACGTACGT"
    r5 = b.select([4])
    got = {x: elems(r5, x)[0]["seq"] for x in (10, 14, 18, 19, 20, 21, 22, 24)}
    assert got == {10: "A", 14: "A", 18: "A", 19: "C", 20: "G", 21: "T", 22: "A", 24: "G"}


@pytest.mark.parametrize("idx", [5, 6])
def test_clipped_reads(idx):  # :346-378  (the suite's comments say 4=1N4=4S / 4=1N4=4H; the fixture holds 4=1D4=4S / 4=1D4=4H)
    b = load_golden("different_start_reads").select([idx])
    got = [elems(b, x)[0]["seq"] for x in (40, 41, 42, 43, 44, 45, 48)]
    assert got == ["A", "C", "G", "T", "", "A", "T"]
    assert elems(b, 44)[0]["kind"] == "MidDeletion"
    assert elems(b, 43)[0]["kind"] == "Deletion" and elems(b, 43)[0]["ref"] == "TG"
    assert elems(b, 49) == []


def test_rna_read():  # :380-402
    b = batch(make_read("CCCCAGCCTAGGCCTTCGACACTGGGGGGCTGAGGGAAGGGGCACCTGCC", "7M191084N43M", "9T24T7G7", 229538779))
    assert elems(b, 229538780)[0]["seq"] == "C"
    assert elems(b, 229538781)[0]["seq"] == "C"
    assert elems(b, 229539779)[0]["seq"] == ""
    assert elems(b, 229729912)[0]["seq"] == "C"


def test_rna_pileup_depths():  # :404-415
    b = load_golden("testrna")
    assert len(elems(b, 229580594)) == 94
    assert len(elems(b, 229580706, "A")) == 4
    assert len(elems(b, 229580707, "N")) == 1


def test_mid_deletion_alleles():  # :417-432
    b = batch(*[make_read("TCGAAAAGCT", "5M6D5M", "5^GCTTCG5", 0)] * 3)
    es = elems(b, 4)
    assert {(e["ref"], e["seq"]) for e in es} == {("AGCTTCG", "A")}
    es = elems(b, 5)
    assert {(e["ref"], e["seq"]) for e in es} == {("G", "")}


# ---- MDTagUtilsSuite (src/test/scala/org/hammerlab/guacamole/reads/MDTagUtilsSuite.scala) ----------------------------
MD_CASES = [
    # (sequence, cigar, md, start, expected per-read reference, expected mismatch count)
    ("GATA", "3M6D1M", "3^GATTCG1", 1, "GATGATTCGA", 0),                  # :29-34
    ("TCGATCGA", "8M", "1A6", 1, "TAGATCGA", 1),                            # :236-240
    ("GCTACTCGAA", "10M", "1A9", 5, "GATACTCGAA", 1),                       # :50-64
    ("GCTACTCAAA", "10M", "1A5G2", 5, "GATACTCGAA", 2),                     # :65-79
    ("GAGGGTACTCGAA", "2M3I8M", "10", 5, "GATACTCGAA", 0),                  # :95-109
    ("GCGGGTACTCGAA", "2M3I8M", "1A5G2", 5, "GATACTCGAA", 2),               # :110-124
    ("ACTCGAATTA", "10M", "7CG1", 8, "ACTCGAACGA", 2),                      # :121
    ("GAGAA", "2M5D3M", "2^TACTC3", 5, "GATACTCGAA", 0),                    # :125-139
    ("ACTCGA", "5M4D1M", "5^AACG1", 8, "ACTCGAACGA", 0),                    # :140-154
    ("AATTGAATTG", "5M1D5M", "5^C5", 0, "AATTGCAATTG", 0),
    ("acgtACGT", "8M", "2t5", 0, "acTtACGT", 1),                            # MdTag upper-cases the tag only
    ("AAAAA", "5M", "0", 0, "AAAAA", 0),
    ("ACGT", "2M2M", "1C0T1", 0, "ACTT", 2),
]


@pytest.mark.parametrize("seq,cigar,md,start,expected,nm", MD_CASES)
def test_md_reference(seq, cigar, md, start, expected, nm):
    b = batch(make_read(seq, cigar, md, start))
    ref, got_nm = orc.md_reference(b, 0)
    assert ref == expected and got_nm == nm


def test_md_reference_rna():  # :215-233  N-skipping RNA read
    seq = "CCCCAGCCTAGGCCTTCGACACTGGGGGGCTGAGGGAAGGGGCACCTGCC"
    b = batch(make_read(seq, "7M191084N43M", "9T24T7G7", 229538779))
    ref, nm = orc.md_reference(b, 0)
    assert len(ref) == 7 + 191084 + 43 and nm == 3
    assert ref[:7] == "CCCCAGC" and set(ref[7:-43]) == {"N"}
    assert ref[-43:] == "CTTGGCCTTCGACACTGGGGGGCTGAGTGAAGGGGGACCTGCC"


def test_md_errors():
    b = batch(make_read("AATTGAATTG", "5M1D5M", "10", 0))  # deletion missing from MD
    with pytest.raises(orc.OracleError):
        orc.md_reference(b, 0)
    b = batch(make_read("AATTG", "5M", None, 0))          # ReferenceWithoutMDTagException
    with pytest.raises(orc.OracleError) as e:
        orc.md_reference(b, 0)
    assert e.value.code == 5
