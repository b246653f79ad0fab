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
                ("contig_length", C.POINTER(C.c_int64)), ("reads_per_locus", C.c_double), ("sample", C.c_int32),
                ("n_windows", C.c_uint32), ("windows", C.POINTER(abi.LocusRangeC)),
                ("frac_clip", C.c_double), ("frac_ins", C.c_double), ("frac_del", C.c_double),
                ("frac_both", C.c_double), ("n_threads", C.c_int32), ("with_qualities", C.c_int32)]


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


def make_params(contigs: Sequence[tuple], depth: float, read_length: int = 150, seed: int = 20261018, sample: int = 0,
                windows: Optional[Sequence[tuple]] = None, n_threads: int = 0, with_qualities: bool = True,
                frac_clip=0.20, frac_ins=0.009, frac_del=0.009, frac_both=0.002):
    """guac_synth_params + the arrays it points at (keep the second value alive while the struct is in use)."""
    lengths = np.asarray([c[1] for c in contigs], dtype=np.int64)
    p = SynthParamsC()
    p.seed, p.n_contigs, p.read_length = seed, len(contigs), read_length
    p.contig_length = lengths.ctypes.data_as(C.POINTER(C.c_int64))
    p.reads_per_locus, p.sample = float(depth) / float(read_length), sample
    keep = [lengths]
    if windows:
        arr = (abi.LocusRangeC * len(windows))()
        for i, w in enumerate(windows):
            arr[i].contig, arr[i].start, arr[i].end = int(w[0]), int(w[1]), int(w[2])
        p.n_windows, p.windows = len(windows), arr
        keep.append(arr)
    p.frac_clip, p.frac_ins, p.frac_del, p.frac_both = frac_clip, frac_ins, frac_del, frac_both
    p.n_threads, p.with_qualities = n_threads, int(with_qualities)
    return p, keep


def generate(contigs: Sequence[tuple], depth: float, read_length: int = 150, seed: int = 20261018, sample: int = 0,
             window: Optional[tuple] = None, n_threads: int = 0, windows: Optional[Sequence[tuple]] = None,
             frac_clip=0.20, frac_ins=0.009, frac_del=0.009, frac_both=0.002) -> SynthBatch:
    """contigs: [(name, length)]; window = (contig_index, start, end) / windows = [...] restrict the reads to those STARTING
    there.  The reads starting at a locus do not depend on the windows: a slice holds what the whole genome would."""
    L = _load()
    names = [c[0] for c in contigs]
    if window is not None:
        windows = [window]
    p, keep = make_params(contigs, depth, read_length, seed, sample, windows, n_threads, True, frac_clip, frac_ins, frac_del, frac_both)
    h = C.c_void_p()
    rc = L.guac_synth_generate(C.byref(p), C.byref(h))
    del keep
    if rc != 0:
        raise RuntimeError(f"guac_synth_generate failed: {abi.STATUS_NAMES.get(rc, rc)}")
    return SynthBatch(h, names, "tumor" if sample == 1 else "normal")


class HostBatch(SynthBatch):
    """Host copy of a device-generated batch (guac_synth_device_batch_download); page-locked when asked for."""

    def __init__(self, handle, contig_names, sample_name):
        from ._lib import lib
        self._h = handle
        self.contig_names = list(contig_names)
        self.sample_name = sample_name
        self.c = lib().guac_synth_host_batch_view(handle).contents

    def free(self):
        if self._h:
            from ._lib import lib
            lib().guac_synth_host_batch_free(self._h)
            self._h = None


class DeviceBatch:
    """A batch generated straight into device memory (guac_synth_generate_device): `.c` is a guac_read_batch whose column
    pointers are DEVICE pointers — hand it to Context.pack_device."""

    def __init__(self, ctx, handle, contig_names, sample_name):
        from ._lib import lib
        self.ctx, self._h = ctx, handle
        self.contig_names = list(contig_names)
        self.sample_name = sample_name
        self.c = lib().guac_synth_device_batch_view(handle).contents
        self.kernel_ms = float(lib().guac_synth_device_batch_ms(handle))
        ops, md, bases = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib().guac_synth_device_batch_totals(handle, C.byref(ops), C.byref(md), C.byref(bases))
        self.n_cigar_ops, self.n_md_bytes, self.n_bases = int(ops.value), int(md.value), int(bases.value)

    @property
    def n_reads(self):
        return int(self.c.n_reads)

    def download(self, pinned: bool = False) -> HostBatch:
        from ._lib import lib
        h = C.c_void_p()
        self.ctx._check(lib().guac_synth_device_batch_download(self.ctx._h, self._h, int(pinned), C.byref(h)))
        return HostBatch(h, self.contig_names, self.sample_name)

    def free(self):
        if self._h:
            from ._lib import lib
            lib().guac_synth_device_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def generate_device(ctx, contigs: Sequence[tuple], depth: float, read_length: int = 150, seed: int = 20261018, sample: int = 0,
                    windows: Optional[Sequence[tuple]] = None, with_qualities: bool = True,
                    frac_clip=0.20, frac_ins=0.009, frac_del=0.009, frac_both=0.002) -> DeviceBatch:
    """generate() on the device of `ctx` (a callers.Context): the same reads, byte for byte, left in HBM."""
    from ._lib import lib
    p, keep = make_params(contigs, depth, read_length, seed, sample, windows, 0, with_qualities, frac_clip, frac_ins, frac_del, frac_both)
    h = C.c_void_p()
    ctx._check(lib().guac_synth_generate_device(ctx._h, C.byref(p), C.byref(h)))
    del keep
    return DeviceBatch(ctx, h, [c[0] for c in contigs], "tumor" if sample == 1 else "normal")


def shard_windows(ranges: Sequence[tuple], read_length: int = 150) -> list:
    """Start windows of the reads that can overlap the loci `ranges` (contig, start, end[, task]) of one shard: a read
    starting up to read_length + 40 before a range reaches into it (reads crossing a shard boundary are generated by both
    neighbours, like the reference duplicates them across tasks, DistributedUtil.scala:585-597).  Merged per contig."""
    out = []
    for r in sorted((int(r[0]), int(r[1]), int(r[2])) for r in ranges):
        s, e = max(0, r[1] - (read_length + 40)), r[2]
        if out and out[-1][0] == r[0] and s <= out[-1][2]:
            out[-1] = (r[0], out[-1][1], max(out[-1][2], e))
        else:
            out.append((r[0], s, e))
    return out
