"""guac_inflate.h — the BAM front end's own raw-deflate decoder (CPU only): byte-for-byte against zlib over streams of every
block type, level and strategy, refusal of corrupt / truncated / mis-sized streams without a write outside the output, and the
loader returning the same reads with and without it."""
import os
import subprocess

import pytest

from conftest import load_golden
from guacamole_b200 import callers
from guacamole_b200.reads import write_bam

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_inflate_matches_zlib_on_4000_streams(tmp_path):
    exe = str(tmp_path / "inflate_check")
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "c", "inflate_check.cpp"), "-lz"]
    subprocess.check_call(cmd)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    status, streams, fast = out.stdout.split()
    # every valid stream is decoded by the fast path (none needs the zlib fallback), 16 sizes x 5 contents x 10 levels x 5 strategies
    assert status == "ok" and int(streams) == 4000 and int(fast) == 4000


@pytest.mark.parametrize("level,block_bytes", [(1, 65280), (6, 3001), (9, 257), (0, 40000)])
def test_bam_loader_same_reads_with_and_without_own_inflate(tmp_path, monkeypatch, level, block_bytes):
    """BGZF members of several sizes and compression levels (level 0 = stored blocks): the loader's columns are identical
    whether its members go through guac_inflate or through zlib (GUAC_BAM_ZLIB_ONLY)."""
    want = load_golden("chrM.sorted")
    path = str(tmp_path / "x.bam")
    write_bam(want, path, level=level, block_bytes=block_bytes)
    monkeypatch.delenv("GUAC_BAM_ZLIB_ONLY", raising=False)
    a = callers.CompactBatch.from_bam(path, n_threads=3)
    monkeypatch.setenv("GUAC_BAM_ZLIB_ONLY", "1")
    b = callers.CompactBatch.from_bam(path, n_threads=3)
    ra, rb = a.to_read_batch(), b.to_read_batch()
    assert len(ra) == len(rb) == len(want)
    for col in ("start", "cigar", "cigar_off", "seq", "seq_off", "qual", "md", "md_off", "mapq", "flags", "contig"):
        va, vb = getattr(ra, col), getattr(rb, col)
        assert (va == vb).all(), col
    a.free()
    b.free()


def test_bam_loader_on_corrupt_members(tmp_path, monkeypatch):
    """Bytes flipped inside BGZF members: the loader fails with an error or — where the damage is harmless to the deflate
    stream — returns what it returns through zlib; it never crashes and the two paths never disagree."""
    import numpy as np
    from guacamole_b200._lib import GuacError
    want = load_golden("chrM.sorted")
    path = str(tmp_path / "x.bam")
    write_bam(want, path, level=6, block_bytes=8192)
    raw = bytearray(open(path, "rb").read())
    rng = np.random.default_rng(5)
    outcomes = set()
    for trial in range(12):
        bad = bytearray(raw)
        for _ in range(3):
            at = int(rng.integers(400, len(bad) - 64))  # (behind the header's members, in front of the EOF marker)
            bad[at] ^= 1 << int(rng.integers(0, 8))
        p2 = str(tmp_path / f"bad{trial}.bam")
        open(p2, "wb").write(bytes(bad))
        res = []
        for zlib_only in (False, True):
            if zlib_only:
                monkeypatch.setenv("GUAC_BAM_ZLIB_ONLY", "1")
            else:
                monkeypatch.delenv("GUAC_BAM_ZLIB_ONLY", raising=False)
            try:
                cb = callers.CompactBatch.from_bam(p2, n_threads=2)
                rb = cb.to_read_batch()
                res.append(("ok", len(rb), bytes(np.asarray(rb.seq)[:2000]), bytes(np.asarray(rb.start).tobytes()[:2000])))
                cb.free()
            except (GuacError, ValueError, RuntimeError) as e:
                res.append(("error",))
        assert res[0][0] == res[1][0], (trial, res[0][0], res[1][0])
        if res[0][0] == "ok":
            assert res[0] == res[1], trial
        outcomes.add(res[0][0])
    assert "error" in outcomes  # (most flips break a member)
