"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/guac.h declares,
ctypes mirrors match the header, host-only entry points work, and the engine refuses to run without a GPU."""
import ctypes as C
import os
import re

import pytest

from guacamole_b200 import abi
from guacamole_b200._lib import EXPORTED, GuacError, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(guac_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = lib()
    declared = declared_functions("guac.h")
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(EXPORTED) == declared


def test_synth_library_exports():
    """guac_synth.h: the host build lives in libguac_synth.so, the device build in libguac_b200.so."""
    from guacamole_b200.synth import _load
    host, device = _load(), lib()
    declared = declared_functions("guac_synth.h")
    assert len(declared) >= 10
    for name in declared:
        on_device = "device" in name or "host_batch" in name or name == "guac_reads_pack_synth"
        assert hasattr(device if on_device else host, name), name


def test_abi_version_and_status_strings():
    L = lib()
    assert L.guac_abi_version() == abi.GUAC_ABI_VERSION
    for code, name in abi.STATUS_NAMES.items():
        assert L.guac_status_string(code).decode() == name


def test_struct_sizes_match_header():
    assert C.sizeof(abi.ThresholdRecordC) == 32
    assert C.sizeof(abi.AlleleEvidenceC) == 64
    assert C.sizeof(abi.SomaticRecordC) == 40 + 2 * 64
    assert C.sizeof(abi.LocusCountsC) == 48
    assert C.sizeof(abi.LocusRangeC) == 24
    assert C.sizeof(abi.StatsC) == 120


def test_partition_loci_uniformly_goldens():  # DistributedUtilSuite.scala:46-63 through the product's host entry point
    from guacamole_b200.loci import partition_loci_uniformly
    fmt = lambda parts: ",".join(f"chrM:{s}-{e}={t}" for (_, s, e, t) in parts)
    assert fmt(partition_loci_uniformly(4, [(0, 0, 16571)])) == "chrM:0-4143=0,chrM:4143-8286=1,chrM:8286-12428=2,chrM:12428-16571=3"
    assert fmt(partition_loci_uniformly(3, [(0, 0, 10)])) == "chrM:0-3=0,chrM:3-7=1,chrM:7-10=2"
    assert fmt(partition_loci_uniformly(4, [(0, 0, 3)])) == "chrM:0-1=0,chrM:1-2=1,chrM:2-3=2"
    assert partition_loci_uniformly(100, [(0, 1000, 1100)]) == [(0, 1000 + i, 1001 + i, i) for i in range(100)]


def test_partition_matches_oracle():
    import oracle_binding as orc
    from guacamole_b200.loci import partition_loci_uniformly
    from guacamole_b200.synth import GRCH37
    loci = [(i, 0, ln) for i, (_, ln) in enumerate(GRCH37)]
    for tasks in (1, 2, 4, 8, 2000):
        assert partition_loci_uniformly(tasks, loci) == orc.partition_loci_uniformly(tasks, loci)


def test_parse_loci():
    from guacamole_b200.loci import parse_loci
    assert parse_loci("all", ["chrM"], [16571]) == [(0, 0, 16570)]          # LociSet.scala:205-207 drops the last base
    assert parse_loci("chr1:5-10, chr1:8-20,chr2", ["chr1", "chr2"], [100, 50]) == [(0, 5, 20), (1, 0, 50)]
    with pytest.raises(ValueError):
        parse_loci("chr9:1-2", ["chr1"], [10])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from guacamole_b200.callers import Context
    with pytest.raises(GuacError) as e:
        Context(0)
    assert e.value.code == abi.ERR_NO_DEVICE


def test_synth_is_deterministic_and_consistent():
    import numpy as np
    import oracle_binding as orc
    from guacamole_b200 import synth
    a = synth.generate([("20", 50000)], depth=20, seed=3, n_threads=1).to_read_batch()
    b = synth.generate([("20", 50000)], depth=20, seed=3, n_threads=4).to_read_batch()
    assert np.array_equal(a.seq, b.seq) and np.array_equal(a.md, b.md) and np.array_equal(a.cigar, b.cigar)
    assert np.all(np.diff(a.start) >= 0)
    # MD tags and CIGARs are mutually consistent: the oracle rebuilds every read's reference without error and
    # overlapping reads agree on it (no order-sensitive loci by construction)
    c = orc.pileup_counts(a, [(0, 0, 49999)]).counts()
    assert len(c) > 40000 and c["depth"].max() < 80


def test_somatic_genotype_filter_matches_oracle():
    """guac_somatic_genotype_filter (host code of the engine, no GPU needed) against the oracle's restatement of
    SomaticGenotypeFilter.apply(Seq, ...) on the records the oracle calls on the reference's tumor/normal fixture."""
    import numpy as np
    import oracle_binding as orc
    from conftest import load_golden
    from guacamole_b200.callers import SOMATIC_DTYPE, somatic_genotype_filter
    t = load_golden("tumor.chr20.tough").filtered(non_duplicate=True, passed_qc=True, has_md=True).sorted()
    n = load_golden("normal.chr20.tough").filtered(non_duplicate=True, passed_qc=True, has_md=True).sorted()
    c = t.contig_names.index("20")
    recs = orc.somatic_standard(t, n, [(c, 0, 63025520)], orc.somatic_params(odds=20)).somatic()
    assert len(recs) > 30
    arr = np.zeros(len(recs), SOMATIC_DTYPE)
    for i, r in enumerate(recs):
        arr[i] = np.frombuffer(bytes(r["_raw"]), dtype=SOMATIC_DTYPE)[0]
    kw = dict(min_tumor_read_depth=8, max_tumor_read_depth=200, min_normal_read_depth=4, min_tumor_alternate_read_depth=3,
              min_vaf=5, min_likelihood=70)
    got = somatic_genotype_filter(arr, seq_overload=True, **kw)
    want = [orc.somatic_genotype_filter(r["_raw"], 8, 200, 4, 3, 120, 5, 70) for r in recs]
    assert list(got) == want and 0 < sum(want) < len(want)
    # the RDD overload adds LOD / mean-MQ / median-mismatch predicates: each can only remove records
    strict = somatic_genotype_filter(arr, min_lod=1, min_average_mapping_quality=40, max_median_mismatches=3, **kw)
    assert np.all(~strict | got) and strict.sum() <= got.sum()
    # all defaults: what is left are the predicates that bite at 0 (log odds > 0, VAF > 0, phred >= 0, NaN means fail)
    default = (arr["somatic_log_odds"] > 0) & (arr["tumor"]["allele_read_depth"] > 0) & (arr["phred_scaled_somatic_likelihood"] >= 0) & \
              (arr["tumor"]["mean_mapping_quality"] >= 0) & (arr["normal"]["mean_mapping_quality"] >= 0) & \
              (arr["tumor"]["median_mismatches_per_read"] <= 2 ** 31 - 1)
    assert np.array_equal(somatic_genotype_filter(arr), default)


def test_plain_c_consumer_builds_and_runs(tmp_path):
    """include/guac.h is the boundary a JNI / cgo / Panama shim compiles against: a plain-C program (tests/c/abi_driver.c)
    must build with gcc -std=c99 -Wall -Werror against it, link the product library and pass its own checks."""
    import shutil
    import subprocess
    import sys
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    from guacamole_b200._lib import LIB_PATH as so
    lib()  # (raises with build instructions if the library is missing)
    exe = str(tmp_path / "abi_driver")
    cmd = [gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_driver.c"),
           "-o", exe, so, "-Wl,-rpath," + os.path.dirname(so)]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = ""   # the driver's last check wants a process without a device
    r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_driver ok" in r.stdout


def test_compact_batch_conversion():
    """guac_read_batch_compact (host only): the columns at BAM width hold what the wide batch held."""
    from guacamole_b200 import callers
    from guacamole_b200.reads import ReadBatch, make_read
    recs = [make_read("TCGATCGAN", "9M", "9", 1, chr="a"), make_read("=CMGRSVTWYHKDBNAA", "17M", "17", 4, chr="a", alignment_quality=7),
            make_read("ACG", "3M", "3", 2, chr="c", is_positive_strand=False)]
    b = ReadBatch.from_records(recs, contig_names=["a", "b", "c"])
    cb = callers.CompactBatch(b, fixed_length=True)
    v = cb.c
    assert v.n_reads == 3 and v.n_contigs == 3 and v.read_length == 0  # (unequal lengths: the offsets stay)
    assert [v.contig_read_off[i] for i in range(4)] == [0, 2, 2, 3]
    assert [v.start[i] for i in range(3)] == [int(x) for x in b.start]
    assert [v.seq_off[i] for i in range(4)] == [0, 9, 26, 29]
    letters = "=ACMGRSVTWYHKDBN"
    bases = "".join(letters[(v.seq4[g >> 1] >> (0 if g & 1 else 4)) & 15] for g in range(29))
    assert bases == "TCGATCGAN" + "=CMGRSVTWYHKDBNAA" + "ACG"
    assert [v.mapq[i] for i in range(3)] == [30, 7, 30] and [v.flags[i] for i in range(3)] == [int(x) for x in b.flags]
    assert [v.qual[i] for i in range(29)] == [int(x) for x in b.qual]
    assert cb.h2d_bytes == 3 * (4 + 4 + 4 + 4 + 1 + 1) + 3 * 4 + 15 + 29 + int(b.md_off[-1])
    cb.free()
    same = ReadBatch.from_records([make_read("ACGT", "4M", "4", 1), make_read("ACGA", "4M", "3T0", 2)])
    cb = callers.CompactBatch(same, fixed_length=True)
    assert cb.c.read_length == 4 and not cb.c.seq_off
    cb.free()
    with pytest.raises(callers.GuacError):
        callers.CompactBatch(ReadBatch.from_records([make_read("ACgT", "4M", "4", 1)]))


def test_python_constants_mirror_the_header():
    """Every integer constant of include/guac.h that guacamole_b200/abi.py mirrors (status codes, read flags, context options,
    the ABI version) carries the header's value: the two are edited by hand."""
    import re
    text = open(os.path.join(ROOT, "include", "guac.h")).read()
    header = {}
    for name, value in re.findall(r"^#define\s+GUAC_([A-Z0-9_]+)\s+(0x[0-9A-Fa-f]+|\d+)u?\b", text, flags=re.M):
        header[name] = int(value, 0)
    for name, value in re.findall(r"^\s*GUAC_(OK|ERR_[A-Z_]+)\s*=\s*(\d+)", text, flags=re.M):
        header[name] = int(value)
    checked = 0
    for name, value in header.items():
        mirror = getattr(abi, name, None)
        if mirror is None and name == "ABI_VERSION":
            mirror = abi.GUAC_ABI_VERSION
        if mirror is None:
            continue
        assert mirror == value, (name, mirror, value)
        checked += 1
    assert checked >= 25, checked
    for name in ("OPT_SORT_RECORDS", "OPT_PACK_QUALITIES", "OPT_HOST_THREADS", "OPT_DIFFERENCE_LISTS", "OPT_SEGMENTS", "OPT_TRIM_CACHE",
                 "OPT_PACK_OVERLAP"):
        assert name in header and getattr(abi, name) == header[name], name
