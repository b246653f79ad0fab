"""The C ABI's hot path driven from C on a GPU (-m gpu): tests/c/gpu_driver.c — no Python, no torch in that process — packs
the reads of the reference's own CPU-runnable input (chrM.sorted.bam, BASELINE.json configs[0]), runs germline-threshold over
loci "all" and prints the records; they must be the oracle's 138 (120 het + 18 hom-alt)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle_binding as orc
from conftest import load_golden

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dump_columns(batch, path):
    n = len(batch.start)
    with open(path, "wb") as f:
        np.asarray([n, len(batch.contig_names), int(batch.cigar_off[-1]), int(batch.seq_off[-1]), int(batch.md_off[-1]), 0], np.uint64).tofile(f)
        np.asarray(batch.contig_lengths, np.int64).tofile(f)
        batch.contig.astype(np.int32).tofile(f)
        batch.start.astype(np.int64).tofile(f)
        batch.cigar_off.astype(np.uint64).tofile(f)
        batch.cigar.astype(np.uint32).tofile(f)
        batch.seq_off.astype(np.uint64).tofile(f)
        batch.seq.astype(np.uint8).tofile(f)
        batch.qual.astype(np.uint8).tofile(f)
        batch.mapq.astype(np.uint8).tofile(f)
        batch.flags.astype(np.uint8).tofile(f)
        batch.md_off.astype(np.uint64).tofile(f)
        batch.md.astype(np.uint8).tofile(f)


def test_c_driver_chrm(tmp_path):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    from guacamole_b200 import loci
    from guacamole_b200._lib import LIB_PATH as so
    b = load_golden("chrM.sorted").filtered(non_duplicate=True, has_md=True).sorted()
    cols = str(tmp_path / "chrM.bin")
    dump_columns(b, cols)
    exe = str(tmp_path / "gpu_driver")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "gpu_driver.c"),
                    "-o", exe, so, "-Wl,-rpath," + os.path.dirname(so)], check=True, capture_output=True, text=True)
    r = subprocess.run([exe, cols], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "gpu_driver ok" in r.stdout
    ranges = loci.parse_loci("all", b.contig_names, b.contig_lengths)
    want = orc.germline_threshold(b, ranges).threshold()
    lines = [ln.split() for ln in r.stdout.splitlines()]
    got = [(int(x[1]), int(x[2]), x[3], x[4], (int(x[5]), int(x[6]))) for x in lines if x and x[0] == "R"]
    assert got == [(g["contig"], g["start"], g["ref"], g["alt"], g["gt"]) for g in want]
    assert len(got) == 138 and sum(1 for g in got if g[4] == (0, 1)) == 120 and sum(1 for g in got if g[4] == (1, 1)) == 18
    compact = [(int(x[1]), int(x[2]), x[3], x[4], (int(x[5]), int(x[6]))) for x in lines if x and x[0] == "C"]
    head = next(x for x in lines if x and x[0] == "records")
    assert len(compact) == int(head[3]) and len(got) == int(head[1]) == int(head[3]) + int(head[5])
    assert set(compact) <= set(got)          # (the exact kernel's records are the rest: variable-length or non-ACGT alleles, ...)
