"""bench.py's record fingerprint (the `e2e_check` of every bench line): independent of record order, sensitive to any change of a
locus, an allele byte or a genotype."""
import importlib.util
import os
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _result(n=50, seed=3):
    rng = np.random.default_rng(seed)
    dt = np.dtype([("contig", "i4"), ("start", "i8"), ("ref_off", "u4"), ("ref_len", "u2"), ("alt_off", "u4"), ("alt_len", "u2"), ("gt", "u1", (2,))])
    rec = np.zeros(n, dt)
    rec["contig"] = rng.integers(0, 3, n)
    rec["start"] = rng.integers(0, 10_000, n)
    rec["ref_off"] = rng.integers(0, 200, n)
    rec["alt_off"] = rng.integers(0, 200, n)
    rec["ref_len"] = rng.integers(1, 4, n)
    rec["alt_len"] = rng.integers(0, 4, n)
    rec["gt"] = rng.integers(0, 3, (n, 2))
    pool = rng.integers(65, 90, 256).astype(np.uint8)
    return types.SimpleNamespace(records=rec, bytes=pool)


def test_record_digest_is_order_independent_and_sensitive():
    b = _bench()
    r = _result()
    d0 = b.record_digest(r, False)
    shuffled = types.SimpleNamespace(records=r.records[np.random.default_rng(9).permutation(len(r.records))], bytes=r.bytes)
    assert b.record_digest(shuffled, False) == d0
    for field, delta in (("start", 1), ("contig", 1), ("ref_len", 1)):
        x = _result()
        x.records[field][7] += delta
        assert b.record_digest(x, False) != d0, field
    x = _result()
    x.records["gt"][11, 1] ^= 1
    assert b.record_digest(x, False) != d0
    x = _result()
    x.bytes[int(x.records["ref_off"][5])] ^= 1   # the first reference byte of record 5
    assert b.record_digest(x, False) != d0
    assert b.record_digest(types.SimpleNamespace(records=r.records[:0], bytes=r.bytes), False) == "empty"
