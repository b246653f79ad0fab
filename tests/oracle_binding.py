"""ctypes binding of the CPU oracle (oracle/libguac_oracle.so).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs — never by guacamole_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

from guacamole_b200 import abi
from guacamole_b200.reads import ReadBatch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB_PATH = os.path.join(_ROOT, "oracle", "libguac_oracle.so")


class ElementC(C.Structure):
    _fields_ = [("read_index", C.c_int64), ("kind", C.c_int32), ("quality_score", C.c_int32),
                ("read_position", C.c_int32), ("cigar_element_index", C.c_int32),
                ("index_within_cigar_element", C.c_int32), ("ref_off", C.c_uint32), ("ref_len", C.c_uint32),
                ("seq_off", C.c_uint32), ("seq_len", C.c_uint32), ("is_positive_strand", C.c_uint8),
                ("pad_", C.c_uint8 * 3)]


class GenotypeLikelihoodC(C.Structure):
    _fields_ = [("a1_ref_off", C.c_uint32), ("a1_ref_len", C.c_uint32), ("a1_alt_off", C.c_uint32),
                ("a1_alt_len", C.c_uint32), ("a2_ref_off", C.c_uint32), ("a2_ref_len", C.c_uint32),
                ("a2_alt_off", C.c_uint32), ("a2_alt_len", C.c_uint32), ("value", C.c_double)]


KINDS = ["Match", "Mismatch", "Insertion", "Deletion", "MidDeletion", "Clipped"]


def build_oracle() -> str:
    src = os.path.join(_ROOT, "oracle", "guac_oracle.cpp")
    if (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle())
        _lib.orc_last_error.restype = C.c_char_p
        _lib.orc_result_n.restype = C.c_size_t
        _lib.orc_result_threshold_records.restype = C.POINTER(abi.ThresholdRecordC)
        _lib.orc_result_somatic_records.restype = C.POINTER(abi.SomaticRecordC)
        _lib.orc_result_called_alleles.restype = C.POINTER(abi.CalledAlleleC)
        _lib.orc_result_allele_counts.restype = C.POINTER(abi.AlleleCountC)
        _lib.orc_result_counts.restype = C.POINTER(abi.LocusCountsC)
        _lib.orc_result_elements.restype = C.POINTER(ElementC)
        _lib.orc_result_likelihoods.restype = C.POINTER(GenotypeLikelihoodC)
        _lib.orc_result_bytes.restype = C.POINTER(C.c_uint8)
        _lib.orc_result_stats.restype = C.POINTER(abi.StatsC)
        _lib.orc_result_reference_base.restype = C.c_uint8
        _lib.orc_phred_to_success_probability.restype = C.c_double
        for f in ("orc_result_n", "orc_result_threshold_records", "orc_result_somatic_records", "orc_result_called_alleles",
                  "orc_result_allele_counts", "orc_result_counts",
                  "orc_result_elements", "orc_result_likelihoods", "orc_result_stats", "orc_result_free",
                  "orc_result_reference_base"):
            getattr(_lib, f).argtypes = [C.c_void_p]
        _lib.orc_result_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
    return _lib


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(f"{abi.STATUS_NAMES.get(code, code)}: {msg}")
        self.code = code


def _check(rc):
    if rc != 0:
        raise OracleError(rc, lib().orc_last_error().decode())


def ranges_array(ranges: Sequence[tuple]):
    """[(contig_index, start, end[, task])] -> LocusRangeC array"""
    arr = (abi.LocusRangeC * max(1, len(ranges)))()
    for i, r in enumerate(ranges):
        arr[i].contig, arr[i].start, arr[i].end = r[0], r[1], r[2]
        arr[i].task = r[3] if len(r) > 3 else 0
    return arr


class Result:
    def __init__(self, handle):
        self.h = handle
        L = lib()
        nb = C.c_size_t()
        p = L.orc_result_bytes(self.h, C.byref(nb))
        self.bytes = bytes(bytearray(p[i] for i in range(nb.value))) if nb.value else b""
        self.stats = abi.struct_to_dict(L.orc_result_stats(self.h).contents)
        self.reference_base = chr(L.orc_result_reference_base(self.h))

    def _s(self, off, ln):
        return self.bytes[off:off + ln].decode("latin1")

    def threshold(self) -> List[dict]:
        L = lib()
        n = L.orc_result_n(self.h)
        p = L.orc_result_threshold_records(self.h)
        out = []
        for i in range(n):
            r = p[i]
            out.append(dict(contig=r.contig, start=r.start, sample=r.sample, ref=self._s(r.ref_off, r.ref_len),
                            alt=self._s(r.alt_off, r.alt_len), gt=(r.gt[0], r.gt[1]), tie=r.tie))
        return out

    def somatic(self) -> List[dict]:
        L = lib()
        n = L.orc_result_n(self.h)
        p = L.orc_result_somatic_records(self.h)
        out = []
        for i in range(n):
            r = p[i]
            d = dict(contig=r.contig, start=r.start, sample=r.sample, ref=self._s(r.ref_off, r.ref_len),
                     alt=self._s(r.alt_off, r.alt_len), phred=r.phred_scaled_somatic_likelihood,
                     somatic_log_odds=r.somatic_log_odds, tumor=abi.struct_to_dict(r.tumor),
                     normal=abi.struct_to_dict(r.normal))
            d["_raw"] = abi.SomaticRecordC.from_buffer_copy(r)
            out.append(d)
        return out

    def allele_counts(self) -> List[dict]:
        L = lib()
        n = L.orc_result_n(self.h)
        p = L.orc_result_allele_counts(self.h)
        return [dict(contig=p[i].contig, start=p[i].start, sample=p[i].sample, ref=self._s(p[i].ref_off, p[i].ref_len),
                     alt=self._s(p[i].alt_off, p[i].alt_len), count=p[i].count) for i in range(n)]

    def called(self) -> List[dict]:
        L = lib()
        n = L.orc_result_n(self.h)
        p = L.orc_result_called_alleles(self.h)
        out = []
        for i in range(n):
            r = p[i]
            out.append(dict(contig=r.contig, start=r.start, sample=r.sample, ref=self._s(r.ref_off, r.ref_len),
                            alt=self._s(r.alt_off, r.alt_len), phred=r.phred_scaled_likelihood,
                            evidence=abi.struct_to_dict(r.evidence)))
        return out

    def counts(self) -> np.ndarray:
        L = lib()
        n = L.orc_result_n(self.h)
        p = L.orc_result_counts(self.h)
        dt = np.dtype([("locus", "<i8"), ("contig", "<i4"), ("depth", "<i4"), ("positive_depth", "<i4"),
                       ("reference_depth", "<i4"), ("base_count", "<i4", 4), ("other_count", "<i4"),
                       ("reference_base", "u1"), ("pad_", "u1", 3)])
        assert dt.itemsize == C.sizeof(abi.LocusCountsC)
        if n == 0:
            return np.zeros(0, dt)
        buf = C.string_at(p, n * dt.itemsize)
        return np.frombuffer(buf, dtype=dt).copy()

    def elements(self) -> List[dict]:
        L = lib()
        n = L.orc_result_n(self.h)
        p = L.orc_result_elements(self.h)
        out = []
        for i in range(n):
            e = p[i]
            out.append(dict(read_index=e.read_index, kind=KINDS[e.kind], quality=e.quality_score,
                            read_position=e.read_position, cigar_element_index=e.cigar_element_index,
                            index_within=e.index_within_cigar_element, ref=self._s(e.ref_off, e.ref_len),
                            seq=self._s(e.seq_off, e.seq_len), positive=bool(e.is_positive_strand)))
        return out

    def likelihoods(self) -> List[dict]:
        L = lib()
        n = L.orc_result_n(self.h)
        p = L.orc_result_likelihoods(self.h)
        out = []
        for i in range(n):
            g = p[i]
            out.append(dict(a1=(self._s(g.a1_ref_off, g.a1_ref_len), self._s(g.a1_alt_off, g.a1_alt_len)),
                            a2=(self._s(g.a2_ref_off, g.a2_ref_len), self._s(g.a2_alt_off, g.a2_alt_len)),
                            value=g.value))
        return out

    def __del__(self):
        try:
            lib().orc_result_free(self.h)
        except Exception:
            pass


def threshold_params(threshold=8, emit_ref=False, emit_no_call=False, skip_empty=True):
    return abi.ThresholdParamsC(threshold, int(emit_ref), int(emit_no_call), int(skip_empty))


def somatic_params(odds=20, min_mapq=1, filter_multi_allelic=False, max_read_depth=2**31 - 1, skip_empty=True):
    return abi.SomaticParamsC(odds, min_mapq, int(filter_multi_allelic), max_read_depth, int(skip_empty))


def reference_c(ref_bases: Optional[Sequence[bytes]]):
    if ref_bases is None:
        return None, None
    offs = np.zeros(len(ref_bases) + 1, np.uint64)
    offs[1:] = np.cumsum([len(b) for b in ref_bases])
    data = np.frombuffer(b"".join(ref_bases) or b"\0", dtype=np.uint8).copy()
    r = abi.ReferenceC(len(ref_bases), offs.ctypes.data_as(C.POINTER(C.c_uint64)),
                       data.ctypes.data_as(C.POINTER(C.c_uint8)))
    return r, (offs, data)


def germline_threshold(batch: ReadBatch, ranges, params=None, n_threads=1, reference=None) -> Result:
    params = params or threshold_params()
    h = C.c_void_p()
    b = batch.to_c()
    ref, keep = reference_c(reference)
    arr = ranges_array(ranges)
    _check(lib().orc_germline_threshold(C.byref(b), C.byref(ref) if ref else None, arr, C.c_size_t(len(ranges)),
                                        C.byref(params), n_threads, C.byref(h)))
    return Result(h)


def somatic_standard(tumor: ReadBatch, normal: ReadBatch, ranges, params=None, n_threads=1, reference=None) -> Result:
    params = params or somatic_params()
    h = C.c_void_p()
    bt, bn = tumor.to_c(), normal.to_c()
    ref, keep = reference_c(reference)
    arr = ranges_array(ranges)
    _check(lib().orc_somatic_standard(C.byref(bt), C.byref(bn), C.byref(ref) if ref else None, arr,
                                      C.c_size_t(len(ranges)), C.byref(params), n_threads, C.byref(h)))
    return Result(h)


def allele_counts(batch: ReadBatch, ranges, n_threads=1, reference=None) -> Result:
    h = C.c_void_p()
    b = batch.to_c()
    ref, keep = reference_c(reference)
    arr = ranges_array(ranges)
    _check(lib().orc_allele_counts(C.byref(b), C.byref(ref) if ref else None, arr, C.c_size_t(len(ranges)), n_threads,
                                   C.byref(h)))
    return Result(h)


def germline_standard(batch: ReadBatch, ranges, min_mapq=1, skip_empty=True, n_threads=1, reference=None) -> Result:
    params = abi.StandardParamsC(min_mapq, int(skip_empty))
    h = C.c_void_p()
    b = batch.to_c()
    ref, keep = reference_c(reference)
    arr = ranges_array(ranges)
    _check(lib().orc_germline_standard(C.byref(b), C.byref(ref) if ref else None, arr, C.c_size_t(len(ranges)),
                                       C.byref(params), n_threads, C.byref(h)))
    return Result(h)


def pileup_counts(batch: ReadBatch, ranges, skip_empty=True, n_threads=1, reference=None) -> Result:
    h = C.c_void_p()
    b = batch.to_c()
    ref, keep = reference_c(reference)
    arr = ranges_array(ranges)
    _check(lib().orc_pileup_counts(C.byref(b), C.byref(ref) if ref else None, arr, C.c_size_t(len(ranges)),
                                   int(skip_empty), n_threads, C.byref(h)))
    return Result(h)


def pileup_at(batch: ReadBatch, contig: int, locus: int, reference_base: Optional[str] = None) -> Result:
    h = C.c_void_p()
    b = batch.to_c()
    _check(lib().orc_pileup_at(C.byref(b), contig, C.c_int64(locus), ord(reference_base) if reference_base else -1,
                               C.byref(h)))
    return Result(h)


def md_reference(batch: ReadBatch, i: int):
    b = batch.to_c()
    buf = (C.c_uint8 * 1_000_000)()
    n = C.c_size_t()
    nm = C.c_int()
    _check(lib().orc_md_reference(C.byref(b), C.c_uint64(i), buf, C.c_size_t(len(buf)), C.byref(n), C.byref(nm)))
    return bytes(buf[:n.value]).decode("latin1"), nm.value


def likelihoods_at(batch, contig, locus, include_alignment=False, log_space=False, normalize=False) -> List[dict]:
    h = C.c_void_p()
    b = batch.to_c()
    _check(lib().orc_likelihoods_at(C.byref(b), contig, C.c_int64(locus), int(include_alignment), int(log_space),
                                    int(normalize), C.byref(h)))
    return Result(h).likelihoods()


def somatic_at(tumor, normal, contig, locus, params=None) -> List[dict]:
    params = params or somatic_params()
    h = C.c_void_p()
    bt, bn = tumor.to_c(), normal.to_c()
    _check(lib().orc_somatic_at(C.byref(bt), C.byref(bn), contig, C.c_int64(locus), C.byref(params), C.byref(h)))
    return Result(h).somatic()


def threshold_at(batch, contig, locus, params=None) -> List[dict]:
    params = params or threshold_params()
    h = C.c_void_p()
    b = batch.to_c()
    _check(lib().orc_threshold_at(C.byref(b), contig, C.c_int64(locus), C.byref(params), C.byref(h)))
    return Result(h).threshold()


def allele_evidence_at(batch, contig, locus, ref: str, alt: str, likelihood: float) -> dict:
    b = batch.to_c()
    ev = abi.AlleleEvidenceC()
    r, a = ref.encode(), alt.encode()
    _check(lib().orc_allele_evidence_at(C.byref(b), contig, C.c_int64(locus), r, C.c_size_t(len(r)), a,
                                        C.c_size_t(len(a)), C.c_double(likelihood), C.byref(ev)))
    return abi.struct_to_dict(ev)


def visited_loci(a: ReadBatch, b: Optional[ReadBatch], ranges, skip_empty=True, half_window=0, max_out=100000):
    ba = a.to_c()
    bb = b.to_c() if b is not None else None
    loci = (C.c_int64 * max_out)()
    ca = (C.c_int32 * max_out)()
    cb = (C.c_int32 * max_out)()
    n = C.c_size_t()
    arr = ranges_array(ranges)
    _check(lib().orc_visited_loci(C.byref(ba), C.byref(bb) if bb is not None else None, arr, C.c_size_t(len(ranges)),
                                  int(skip_empty), C.c_int64(half_window), loci, ca, cb, C.c_size_t(max_out),
                                  C.byref(n)))
    k = min(n.value, max_out)
    return list(loci[:k]), list(ca[:k]), list(cb[:k])


def partition_loci_uniformly(tasks: int, loci) -> List[tuple]:
    arr = ranges_array(loci)
    out = (abi.LocusRangeC * 100000)()
    n = C.c_size_t()
    _check(lib().orc_partition_loci_uniformly(C.c_int64(tasks), arr, C.c_size_t(len(loci)), out,
                                              C.c_size_t(100000), C.byref(n)))
    return [(out[i].contig, out[i].start, out[i].end, out[i].task) for i in range(n.value)]


def partition_loci_by_approximate_depth(tasks: int, loci, accuracy: int, *batches) -> List[tuple]:
    """DistributedUtil.partitionLociByApproximateDepth over the reads of `batches` (ReadBatch objects)."""
    arr = ranges_array(loci)
    cs = [b.to_c() for b in batches]
    ptrs = (C.c_void_p * len(cs))(*[C.cast(C.pointer(c), C.c_void_p) for c in cs])
    out = (abi.LocusRangeC * 400000)()
    n = C.c_size_t()
    _check(lib().orc_partition_loci_by_approximate_depth(C.c_int64(tasks), arr, C.c_size_t(len(loci)), C.c_int64(accuracy), ptrs,
                                                         C.c_size_t(len(cs)), out, C.c_size_t(400000), C.byref(n)))
    return [(out[i].contig, out[i].start, out[i].end, out[i].task) for i in range(n.value)]


def somatic_genotype_filter(raw_record, min_tumor_read_depth, max_tumor_read_depth, min_normal_read_depth,
                            min_tumor_alternate_read_depth, min_log_odds, min_vaf, min_likelihood) -> bool:
    return bool(lib().orc_somatic_genotype_filter(C.byref(raw_record), min_tumor_read_depth, max_tumor_read_depth,
                                                  min_normal_read_depth, min_tumor_alternate_read_depth, min_log_odds,
                                                  min_vaf, min_likelihood))
