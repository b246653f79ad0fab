"""Synthetic reads of the BASELINE.json shapes (binding of libguac_synth.so, see include/guac_synth.h)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import abi
from .reads import ReadBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libguac_synth.so")

# GRCh37 primary contigs (the lengths the reference's own suite lists, DistributedUtilSuite.scala:72)
GRCH37 = [("1", 249250621), ("2", 243199373), ("3", 198022430), ("4", 191154276), ("5", 180915260), ("6", 171115067),
          ("7", 159138663), ("8", 146364022), ("9", 141213431), ("10", 135534747), ("11", 135006516), ("12", 133851895),
          ("13", 115169878), ("14", 107349540), ("15", 102531392), ("16", 90354753), ("17", 81195210), ("18", 78077248),
          ("19", 59128983), ("20", 63025520), ("21", 48129895), ("22", 51304566), ("X", 155270560), ("Y", 59373566),
          ("MT", 16569)]
CHR20_LENGTH = 63025520


class SynthParamsC(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_contigs", C.c_uint32), ("read_length", C.c_int32),
                ("contig_length", C.POINTER(C.c_int64)), ("n_reads", C.c_uint64), ("sample", C.c_int32),
                ("window_contig", C.c_int32), ("window_start", C.c_int64), ("window_end", C.c_int64),
                ("frac_clip", C.c_double), ("frac_ins", C.c_double), ("frac_del", C.c_double),
                ("frac_both", C.c_double), ("n_threads", C.c_int32), ("pad_", C.c_int32)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(f"{_LIB_PATH} is missing: run __graft_entry__.build()")
        _lib = C.CDLL(_LIB_PATH)
        _lib.guac_synth_batch_view.restype = C.POINTER(abi.ReadBatchC)
        _lib.guac_synth_batch_view.argtypes = [C.c_void_p]
        _lib.guac_synth_batch_free.argtypes = [C.c_void_p]
        _lib.guac_synth_generate.argtypes = [C.POINTER(SynthParamsC), C.POINTER(C.c_void_p)]
    return _lib


class SynthBatch:
    """Owns a generated guac_read_batch; `.c` is the C view, `.to_read_batch()` copies it into numpy columns."""

    def __init__(self, handle, contig_names, sample_name):
        self._h = handle
        self.contig_names = list(contig_names)
        self.sample_name = sample_name
        self.c = _load().guac_synth_batch_view(handle).contents

    @property
    def n_reads(self):
        return int(self.c.n_reads)

    def _arr(self, ptr, n, dtype):
        if n == 0:
            return np.zeros(0, dtype)
        return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype)

    def to_read_batch(self) -> ReadBatch:
        c, n = self.c, int(self.c.n_reads)
        cigar_off = self._arr(c.cigar_off, n + 1, np.uint64).copy()
        seq_off = self._arr(c.seq_off, n + 1, np.uint64).copy()
        md_off = self._arr(c.md_off, n + 1, np.uint64).copy()
        md_addr = C.c_void_p.from_address(C.addressof(c) + abi.ReadBatchC.md.offset).value  # raw pointer, not bytes
        md = np.frombuffer(C.string_at(md_addr, int(md_off[-1])), dtype=np.uint8).copy() if n else np.zeros(0, np.uint8)
        return ReadBatch(
            self.contig_names, self._arr(c.contig_length, int(c.n_contigs), np.int64).copy(), [self.sample_name],
            self._arr(c.contig, n, np.int32).copy(), self._arr(c.start, n, np.int64).copy(), cigar_off,
            self._arr(c.cigar, int(cigar_off[-1]), np.uint32).copy(), seq_off,
            self._arr(c.seq, int(seq_off[-1]), np.uint8).copy(), self._arr(c.qual, int(seq_off[-1]), np.uint8).copy(),
            self._arr(c.mapq, n, np.uint8).copy(), self._arr(c.flags, n, np.uint8).copy(),
            np.zeros(n, np.int32), md_off, md)

    def free(self):
        if self._h:
            _load().guac_synth_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def generate(contigs: Sequence[tuple], depth: float, read_length: int = 150, seed: int = 20261018, sample: int = 0,
             window: Optional[tuple] = None, n_threads: int = 0, n_reads: Optional[int] = None,
             frac_clip=0.20, frac_ins=0.009, frac_del=0.009, frac_both=0.002) -> SynthBatch:
    """contigs: [(name, length)]; window: (contig_index, start, end) restricts the reads (and their count) to a slice."""
    L = _load()
    names = [c[0] for c in contigs]
    lengths = np.asarray([c[1] for c in contigs], dtype=np.int64)
    if n_reads is None:
        loci = (window[2] - window[1]) if window else int(lengths.sum())
        n_reads = int(depth * loci / read_length)
    p = SynthParamsC()
    p.seed, p.n_contigs, p.read_length = seed, len(contigs), read_length
    p.contig_length = lengths.ctypes.data_as(C.POINTER(C.c_int64))
    p.n_reads, p.sample = n_reads, sample
    if window:
        p.window_contig, p.window_start, p.window_end = window
    p.frac_clip, p.frac_ins, p.frac_del, p.frac_both = frac_clip, frac_ins, frac_del, frac_both
    p.n_threads = n_threads
    h = C.c_void_p()
    rc = L.guac_synth_generate(C.byref(p), C.byref(h))
    if rc != 0:
        raise RuntimeError(f"guac_synth_generate failed: {abi.STATUS_NAMES.get(rc, rc)}")
    return SynthBatch(h, names, "tumor" if sample == 1 else "normal")
