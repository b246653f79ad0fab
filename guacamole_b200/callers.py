"""Host-side mirror of the reference's caller interface for the pileup-and-call path.

    reference (Scala)                                                       here
    ---------------------------------------------------------------------  -----------------------------------------
    DistributedUtil.pileupFlatMap(reads, lociPartitions, skipEmpty,          germline_threshold(ctx, reads, loci, ...)
        GermlineThreshold.Caller.callVariantsAtLocus(_, threshold, ...))      (commands/GermlineThresholdCaller.scala:73-81)
    DistributedUtil.pileupFlatMapTwoRDDs(tumor, normal, lociPartitions,      somatic_standard(ctx, tumor, normal, loci, ...)
        skipEmpty, SomaticStandard.Caller.findPotentialVariantAtLocus(...))   (commands/SomaticStandardCaller.scala:103-119)
    pileupFlatMap(reads, loci, skipEmpty, p => (depth, ...))                 pileup_counts(ctx, reads, loci, skip_empty)

Everything goes through the C ABI of include/guac.h (libguac_b200.so); nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import abi
from ._lib import GuacError, lib
from .loci import ranges_to_c
from .reads import ReadBatch

COUNTS_DTYPE = np.dtype([("locus", "<i8"), ("contig", "<i4"), ("depth", "<i4"), ("positive_depth", "<i4"),
                         ("reference_depth", "<i4"), ("base_count", "<i4", 4), ("other_count", "<i4"),
                         ("reference_base", "u1"), ("pad_", "u1", 3)])
THRESHOLD_DTYPE = np.dtype([("start", "<i8"), ("contig", "<i4"), ("sample", "<i4"), ("ref_off", "<u4"),
                            ("alt_off", "<u4"), ("ref_len", "<u2"), ("alt_len", "<u2"), ("gt", "u1", 2), ("tie", "u1"),
                            ("pad_", "u1")])
_EVIDENCE = [("likelihood", "<f8"), ("mean_mapping_quality", "<f8"), ("median_mapping_quality", "<f8"),
             ("mean_base_quality", "<f8"), ("median_base_quality", "<f8"), ("median_mismatches_per_read", "<f8"),
             ("read_depth", "<i4"), ("allele_read_depth", "<i4"), ("forward_depth", "<i4"),
             ("allele_forward_depth", "<i4")]
SOMATIC_DTYPE = np.dtype([("start", "<i8"), ("contig", "<i4"), ("sample", "<i4"), ("ref_off", "<u4"),
                          ("alt_off", "<u4"), ("ref_len", "<u2"), ("alt_len", "<u2"),
                          ("phred_scaled_somatic_likelihood", "<i4"), ("somatic_log_odds", "<f8"),
                          ("tumor", _EVIDENCE), ("normal", _EVIDENCE)])
CALLED_DTYPE = np.dtype([("start", "<i8"), ("contig", "<i4"), ("sample", "<i4"), ("ref_off", "<u4"),
                         ("alt_off", "<u4"), ("ref_len", "<u2"), ("alt_len", "<u2"),
                         ("phred_scaled_likelihood", "<i4"), ("evidence", _EVIDENCE)])
assert CALLED_DTYPE.itemsize == C.sizeof(abi.CalledAlleleC)
ALLELE_COUNT_DTYPE = np.dtype([("start", "<i8"), ("contig", "<i4"), ("sample", "<i4"), ("ref_off", "<u4"),
                               ("alt_off", "<u4"), ("ref_len", "<u2"), ("alt_len", "<u2"), ("count", "<i4")])
assert ALLELE_COUNT_DTYPE.itemsize == C.sizeof(abi.AlleleCountC)
assert COUNTS_DTYPE.itemsize == C.sizeof(abi.LocusCountsC)
assert THRESHOLD_DTYPE.itemsize == C.sizeof(abi.ThresholdRecordC)
assert SOMATIC_DTYPE.itemsize == C.sizeof(abi.SomaticRecordC)


class Context:
    """guac_ctx: one CUDA device + stream.  Single-threaded; use one per thread / rank."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        rc = lib().guac_ctx_create(device, C.byref(self._h))
        if rc != 0:
            raise GuacError(rc, "guac_ctx_create failed (no CUDA device of compute capability 10.x?)")

    def _check(self, rc):
        if rc != 0:
            raise GuacError(rc, lib().guac_last_error(self._h).decode())

    def close(self):
        if self._h:
            lib().guac_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, option: int, value: int) -> None:
        self._check(lib().guac_ctx_set_option(self._h, option, value))

    def timer_start(self) -> None:
        self._check(lib().guac_ctx_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        self._check(lib().guac_ctx_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def pack_c(self, batch_c, contig_names=None, sample_names=None) -> "PackedReads":
        """Pack straight from a guac_read_batch (e.g. the synthetic generator's buffers) without a numpy copy."""
        return PackedReads(self, None, None, batch_c=batch_c, contig_names=contig_names, sample_names=sample_names)

    def pack_device(self, batch_c, contig_names=None, sample_names=None) -> "PackedReads":
        """guac_reads_pack_device: the batch's columns already live in this context's device memory."""
        return PackedReads(self, None, None, batch_c=batch_c, contig_names=contig_names, sample_names=sample_names, on_device=True)

    def pack_v2(self, compact: "CompactBatch", contig_names=None, sample_names=None, reference=None) -> "PackedReads":
        """guac_reads_pack_v2: a compact host batch (4-bit bases, 32-bit columns), widened on the device."""
        return PackedReads(self, None, reference, batch_c=compact.c, contig_names=contig_names, sample_names=sample_names, v2=True)

    def pack_synth(self, device_batch) -> "PackedReads":
        """guac_reads_pack_synth: packs a synth.DeviceBatch, taking its large columns over instead of copying them."""
        return PackedReads(self, None, None, synth_batch=device_batch, contig_names=device_batch.contig_names,
                           sample_names=[device_batch.sample_name])

    def pack(self, batch: ReadBatch, reference: Optional[Sequence[bytes]] = None) -> "PackedReads":
        return PackedReads(self, batch, reference)


class CompactBatch:
    """guac_host_batch_v2: a read batch converted to the compact columns (guac_read_batch_compact).  `source` is a ReadBatch or
    a guac_read_batch (ctypes); it is only read during the conversion."""

    def __init__(self, source, pinned: bool = False, fixed_length: bool = True):
        self._h = C.c_void_p()
        if source is None:  # (from_bam fills the handle in)
            return
        self.contig_names = list(getattr(source, "contig_names", []) or [])
        self.sample_names = list(getattr(source, "sample_names", []) or ["default"])
        b = source.to_c() if hasattr(source, "to_c") else source
        rc = lib().guac_read_batch_compact(C.byref(b), int(pinned), int(fixed_length), C.byref(self._h))
        if rc != abi.OK:
            raise GuacError(rc, lib().guac_status_string(rc).decode())
        self.c = lib().guac_host_batch_v2_view(self._h).contents

    @classmethod
    def from_bam(cls, path: str, non_duplicate=False, passed_qc=False, has_md_tag=False, is_paired=False, with_qualities=True,
                 pinned=False, sample: Optional[str] = None, n_threads: int = 0) -> "CompactBatch":
        """guac_bam_load: a BAM file decoded on host threads straight into the compact columns (no GPU needed)."""
        self = cls(None)
        opt = abi.BamOptionsC(n_threads, int(non_duplicate), int(passed_qc), int(has_md_tag), int(is_paired), int(with_qualities),
                              int(pinned), 0, sample.encode() if sample is not None else None)
        rc = lib().guac_bam_load(path.encode(), C.byref(opt), C.byref(self._h))
        if rc != abi.OK:
            raise GuacError(rc, lib().guac_bam_last_error().decode())
        self.c = lib().guac_host_batch_v2_view(self._h).contents
        self.contig_names = [lib().guac_host_batch_v2_contig_name(self._h, i).decode() for i in range(self.c.n_contigs)]
        self.sample_names = [lib().guac_host_batch_v2_sample_name(self._h).decode()]
        return self

    @property
    def decode_stats(self) -> dict:
        st = (C.c_uint64 * 4)()
        ms = lib().guac_host_batch_v2_decode_stats(self._h, st)
        return {"file_bytes": int(st[0]), "inflated_bytes": int(st[1]), "records_in_file": int(st[2]), "reads": int(st[3]), "decode_ms": float(ms)}

    def to_read_batch(self) -> ReadBatch:
        """The wide ReadBatch holding the same reads (tests; examples)."""
        v, n = self.c, int(self.c.n_reads)
        arr = lambda p, k, dt: np.ctypeslib.as_array(p, shape=(max(int(k), 1),))[:int(k)].astype(dt) if k else np.zeros(0, dt)
        cigar_off, md_off = arr(v.cigar_off, n + 1, np.uint64), arr(v.md_off, n + 1, np.uint64)
        seq_off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(v.read_length)) if v.read_length else arr(v.seq_off, n + 1, np.uint64)
        if n == 0:
            cigar_off = md_off = seq_off = np.zeros(1, np.uint64)
        n_bases = int(seq_off[-1])
        packed = arr(v.seq4, (n_bases + 1) // 2, np.uint8)
        nib = np.empty(2 * len(packed), np.uint8)
        nib[0::2], nib[1::2] = packed >> 4, packed & 15
        seq = np.frombuffer(b"=ACMGRSVTWYHKDBN", np.uint8)[nib[:n_bases]]
        off = [int(v.contig_read_off[i]) for i in range(v.n_contigs + 1)] if n else [0] * (v.n_contigs + 1)
        contig = np.zeros(n, np.int32)
        for ci in range(v.n_contigs):
            contig[off[ci]:off[ci + 1]] = ci
        lengths = np.asarray([int(v.contig_length[i]) for i in range(v.n_contigs)], np.int64) if v.contig_length else None
        md = np.frombuffer(C.string_at(v.md, int(md_off[-1])), np.uint8).copy() if int(md_off[-1]) else np.zeros(0, np.uint8)
        return ReadBatch(contig_names=list(self.contig_names), sample_names=list(self.sample_names), contig=contig,
                         start=arr(v.start, n, np.int64), cigar_off=cigar_off, cigar=arr(v.cigar, int(cigar_off[-1]), np.uint32),
                         seq_off=seq_off, seq=seq.copy(), qual=arr(v.qual, n_bases, np.uint8) if v.qual else np.zeros(n_bases, np.uint8),
                         mapq=arr(v.mapq, n, np.uint8), flags=arr(v.flags, n, np.uint8), sample=np.zeros(n, np.int32), md_off=md_off, md=md,
                         contig_lengths=lengths)

    @property
    def h2d_bytes(self) -> int:
        return int(lib().guac_host_batch_v2_bytes(self._h))

    def free(self):
        if self._h:
            lib().guac_host_batch_v2_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PackedReads:
    """guac_reads: one sample's start-sorted reads packed into the device SoA (guac_reads_pack)."""

    def __init__(self, ctx: Context, batch: Optional[ReadBatch], reference: Optional[Sequence[bytes]] = None,
                 batch_c=None, contig_names=None, sample_names=None, on_device=False, synth_batch=None, v2=False):
        self.ctx = ctx
        self.contig_names = list(batch.contig_names if batch is not None else (contig_names or []))
        self.sample_names = list(batch.sample_names if batch is not None else (sample_names or ["default"]))
        self._h = C.c_void_p()
        if synth_batch is not None:
            ctx._check(lib().guac_reads_pack_synth(ctx._h, synth_batch._h, None, C.byref(self._h)))
            return
        b = batch.to_c() if batch is not None else batch_c
        ref = None
        if reference is not None:
            offs = np.zeros(len(reference) + 1, np.uint64)
            offs[1:] = np.cumsum([len(x) for x in reference])
            data = np.frombuffer(b"".join(reference) or b"\0", dtype=np.uint8).copy()
            ref = abi.ReferenceC(len(reference), offs.ctypes.data_as(C.POINTER(C.c_uint64)),
                                 data.ctypes.data_as(C.POINTER(C.c_uint8)))
        pack = lib().guac_reads_pack_v2 if v2 else lib().guac_reads_pack_device if on_device else lib().guac_reads_pack
        ctx._check(pack(ctx._h, C.byref(b), C.byref(ref) if ref is not None else None, C.byref(self._h)))

    @property
    def n_reads(self) -> int:
        return int(lib().guac_reads_count(self._h))

    @property
    def device_bytes(self) -> int:
        return int(lib().guac_reads_device_bytes(self._h))

    @property
    def order_sensitive_loci(self) -> int:
        return int(lib().guac_reads_order_sensitive_loci(self._h))

    @property
    def expand_kernel_ms(self) -> float:
        return float(lib().guac_reads_expand_kernel_ms(self._h))

    @property
    def h2d_bytes(self) -> int:
        return int(lib().guac_reads_h2d_bytes(self._h))

    @property
    def pack_kernel_ms(self) -> float:
        return float(lib().guac_reads_pack_kernel_ms(self._h))

    def free(self):
        if self._h:
            lib().guac_reads_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _ResultHandle:
    """Owns one guac_result; freed when the last view onto its buffers goes away."""

    def __init__(self, h):
        self.h = h

    def release(self):
        if self.h:
            h, self.h = self.h, None
            lib().guac_result_free(h)

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Result:
    """guac_result: numpy views straight onto the library-owned (page-locked) record buffer and allele byte pool — no
    copy.  The buffers live as long as this object (or any array derived from `records`) does."""

    def __init__(self, handle, kind: str):
        L = lib()
        self.kind = kind
        self._h = _ResultHandle(handle)
        self._n = int(L.guac_result_n(handle))
        self._records = None
        self._stats = None
        self._pool = None

    @property
    def stats(self) -> dict:
        if self._stats is None:
            self._stats = abi.struct_to_dict(lib().guac_result_stats(self._h.h).contents)
        return self._stats

    @property
    def _bytes(self):
        if self._pool is None:
            nb = C.c_size_t()
            bp = lib().guac_result_bytes(self._h.h, C.byref(nb))
            if nb.value:
                rawb = (C.c_uint8 * nb.value).from_address(C.cast(bp, C.c_void_p).value)
                rawb._owner = self._h
                self._pool = memoryview(rawb).cast("B")
            else:
                self._pool = memoryview(b"")
        return self._pool

    def _view(self, p, n, dt):
        if not n:
            return np.zeros(0, dt)
        raw = (C.c_uint8 * (n * dt.itemsize)).from_address(C.cast(p, C.c_void_p).value)
        raw._owner = self._h  # the view keeps the library buffer alive (no reference cycle through self)
        return np.frombuffer(raw, dtype=np.uint8).view(dt)

    @property
    def records(self) -> np.ndarray:
        """The records as a structured array over the library's buffer.  For germline-threshold results the library builds
        the guac_threshold_record view on this first access (the records crossed the bus in compact form, see compact())."""
        if self._records is None:
            L, h = lib(), self._h.h
            if self.kind == "threshold":
                p, dt = L.guac_result_threshold_records(h), THRESHOLD_DTYPE
            elif self.kind == "somatic":
                p, dt = L.guac_result_somatic_records(h), SOMATIC_DTYPE
            elif self.kind == "called":
                p, dt = L.guac_result_called_alleles(h), CALLED_DTYPE
            elif self.kind == "allele_counts":
                p, dt = L.guac_result_allele_counts(h), ALLELE_COUNT_DTYPE
            else:
                p, dt = L.guac_result_counts(h), COUNTS_DTYPE
            self._records = self._view(p, self._n, dt)
        return self._records

    def compact(self):
        """guac_result_compact_records: (compact u64 records, general guac_threshold_record array, sample)."""
        L = lib()
        pc, pg, ng, sm = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_int32()
        n = L.guac_result_compact_records(self._h.h, C.byref(pc), C.byref(pg), C.byref(ng), C.byref(sm))
        return (self._view(pc, int(n), np.dtype("<u8")), self._view(pg, int(ng.value), THRESHOLD_DTYPE), int(sm.value))

    @property
    def bytes(self) -> bytes:
        return bytes(self._bytes)

    def free(self):
        """Detach from the library buffers (copying what is still referenced) and release them now."""
        self._records = self.records.copy()
        self._pool = memoryview(bytes(self._bytes))
        _ = self.stats
        self._h.release()

    def __len__(self):
        return self._n

    def _s(self, off, ln):
        return bytes(self._bytes[int(off):int(off) + int(ln)]).decode("latin1")

    def genotypes(self) -> List[dict]:
        """Records as the fields of the bdg-formats Genotype the reference builds
        (GermlineThresholdCaller.scala:106-117, AlleleConversions.scala:47-62)."""
        out = []
        for r in self.records:
            d = dict(contig=int(r["contig"]), start=int(r["start"]), sample=int(r["sample"]),
                     ref=self._s(r["ref_off"], r["ref_len"]), alt=self._s(r["alt_off"], r["alt_len"]))
            if self.kind == "threshold":
                d["gt"] = (int(r["gt"][0]), int(r["gt"][1]))
                d["tie"] = int(r["tie"])
            elif self.kind == "allele_counts":  # VariantSupport.AlleleCount (commands/VariantSupport.scala:36-41)
                d["count"] = int(r["count"])
            elif self.kind == "called":  # AlleleConversions.calledAlleleToADAMGenotype (AlleleConversions.scala:30-45)
                d["gt"] = (abi.GT_REF, abi.GT_ALT)
                d["phred"] = int(r["phred_scaled_likelihood"])
                d["evidence"] = {k: (float(r["evidence"][k]) if k[0] == "m" or k == "likelihood" else int(r["evidence"][k]))
                                 for k, _ in _EVIDENCE}
            else:
                d["gt"] = (abi.GT_REF, abi.GT_ALT)
                d["phred"] = int(r["phred_scaled_somatic_likelihood"])
                d["somatic_log_odds"] = float(r["somatic_log_odds"])
                for side in ("tumor", "normal"):
                    d[side] = {k: (float(r[side][k]) if k[0] == "m" or k == "likelihood" else int(r[side][k]))
                               for k, _ in _EVIDENCE}
            out.append(d)
        return out


def _ranges(loci):
    return ranges_to_c(loci), len(loci)


def germline_threshold(ctx: Context, reads: PackedReads, loci_partitions, threshold: int = 8, emit_ref: bool = False,
                       emit_no_call: bool = False, skip_empty: bool = True) -> Result:
    """pileupFlatMap(reads, lociPartitions, skipEmpty, callVariantsAtLocus(_, threshold, emitRef, emitNoCall))."""
    arr, n = _ranges(loci_partitions)
    prm = abi.ThresholdParamsC(threshold, int(emit_ref), int(emit_no_call), int(skip_empty))
    h = C.c_void_p()
    ctx._check(lib().guac_germline_threshold(ctx._h, reads._h, arr, n, C.byref(prm), C.byref(h)))
    return Result(h, "threshold")


def somatic_standard(ctx: Context, tumor: PackedReads, normal: PackedReads, loci_partitions, odds_threshold: int = 20,
                     min_alignment_quality: int = 1, filter_multi_allelic: bool = False,
                     max_read_depth: int = 2 ** 31 - 1, skip_empty: bool = True, filters: Optional[dict] = None) -> Result:
    """pileupFlatMapTwoRDDs(tumor, normal, lociPartitions, skipEmpty, findPotentialVariantAtLocus(...)).  `filters`: the
    keyword arguments of somatic_genotype_filter — the post-call genotype filters (SomaticStandardCaller.scala:125-151) then
    run on the device, and only the records that pass come back (guac_somatic_standard_filtered)."""
    arr, n = _ranges(loci_partitions)
    prm = abi.SomaticParamsC(odds_threshold, min_alignment_quality, int(filter_multi_allelic), max_read_depth,
                             int(skip_empty))
    h = C.c_void_p()
    if filters is None:
        ctx._check(lib().guac_somatic_standard(ctx._h, tumor._h, normal._h, arr, n, C.byref(prm), C.byref(h)))
    else:
        fp = _filter_params(**filters)
        ctx._check(lib().guac_somatic_standard_filtered(ctx._h, tumor._h, normal._h, arr, n, C.byref(prm), C.byref(fp), C.byref(h)))
    return Result(h, "somatic")


def _filter_params(min_tumor_read_depth=0, max_tumor_read_depth=2 ** 31 - 1, min_normal_read_depth=0,
                   min_tumor_alternate_read_depth=0, min_lod=0, min_likelihood=0, min_vaf=0,
                   min_average_mapping_quality=0, min_average_base_quality=0, max_median_mismatches=2 ** 31 - 1,
                   seq_overload=False):
    return abi.SomaticFilterParamsC(min_tumor_read_depth, max_tumor_read_depth, min_normal_read_depth,
                                    min_tumor_alternate_read_depth, min_lod, min_likelihood, min_vaf,
                                    min_average_mapping_quality, min_average_base_quality, max_median_mismatches,
                                    int(seq_overload), 0)


def germline_threshold_by_sample(ctx: Context, batch: ReadBatch, loci_partitions, reference: Optional[Sequence[bytes]] = None,
                                 **params) -> List[dict]:
    """The reference's callers group a pileup `bySample` (pileup/Pileup.scala:49-53; GermlineThresholdCaller.scala:97) and
    call every sample on its own elements; a packed read set holds ONE sample (its reference track is derived from its
    own reads), so a batch of several samples is split here, packed and called sample by sample, and the records merged
    in canonical (contig, start, sample, ref, alt) order.  Returns Result.genotypes()-style dicts."""
    out = []
    for s in np.unique(batch.sample):
        sub = batch.select(np.nonzero(batch.sample == s)[0])
        reads = ctx.pack(sub, reference)
        try:
            out.extend(germline_threshold(ctx, reads, loci_partitions, **params).genotypes())
        finally:
            reads.free()
    out.sort(key=lambda g: (g["contig"], g["start"], g["sample"], g["ref"], g["alt"]))
    return out


def germline_standard(ctx: Context, reads: PackedReads, loci_partitions, min_alignment_quality: int = 1,
                      skip_empty: bool = True) -> Result:
    """pileupFlatMap(reads, lociPartitions, skipEmpty, GermlineStandard.callVariantsAtLocus(_, minAlignmentQuality))
    (commands/GermlineStandardCaller.scala:66-70): one CalledAllele per non-reference allele of the most likely genotype."""
    arr, n = _ranges(loci_partitions)
    prm = abi.StandardParamsC(min_alignment_quality, int(skip_empty))
    h = C.c_void_p()
    ctx._check(lib().guac_germline_standard(ctx._h, reads._h, arr, n, C.byref(prm), C.byref(h)))
    return Result(h, "called")


def allele_counts(ctx: Context, reads: PackedReads, loci_partitions) -> Result:
    """pileupFlatMap(reads, lociPartitions, true, VariantSupport.Caller.pileupToAlleleCounts)
    (commands/VariantSupport.scala:93-98, 110-118): every distinct allele of every non-empty pileup with its read count."""
    arr, n = _ranges(loci_partitions)
    h = C.c_void_p()
    ctx._check(lib().guac_allele_counts(ctx._h, reads._h, arr, n, C.byref(h)))
    return Result(h, "allele_counts")


def variant_loci(counts: np.ndarray, min_read_depth: int = 0, min_variant_allele_frequency: int = 0) -> np.ndarray:
    """VAFHistogram.variantLociFromReads' closure (commands/VAFHistogram.scala:31-38, 215-223) over the rows of
    pileup_counts(): loci with referenceDepth != depth, depth >= minReadDepth and variantAlleleFrequency (a Float:
    (depth - referenceDepth).toFloat / depth) >= minVariantAlleleFrequency / 100.0.  Returns (contig, locus, vaf) rows."""
    depth = counts["depth"].astype(np.int64)
    refd = counts["reference_depth"].astype(np.int64)
    with np.errstate(divide="ignore", invalid="ignore"):
        vaf = (depth - refd).astype(np.float32) / depth.astype(np.float32)   # Float / Int -> Float
    keep = (refd != depth) & (depth >= min_read_depth) & (vaf.astype(np.float64) >= min_variant_allele_frequency / 100.0)
    out = np.zeros(int(keep.sum()), dtype=[("contig", "<i4"), ("locus", "<i8"), ("variant_allele_frequency", "<f4")])
    out["contig"], out["locus"], out["variant_allele_frequency"] = counts["contig"][keep], counts["locus"][keep], vaf[keep]
    return out


def generate_vaf_histogram(variant_allele_frequencies: Sequence[float], bins: int) -> dict:
    """VAFHistogram.generateVAFHistogram (commands/VAFHistogram.scala:185-194): bin = percent - percent % (100 / bins)
    with percent = (vaf * 100).toInt in Float arithmetic; returns {bin: number of loci}."""
    if not (1 <= bins <= 100):
        raise ValueError("Bins should be between 1 and 100")
    v = np.asarray(variant_allele_frequencies, dtype=np.float32)
    percent = (v * np.float32(100)).astype(np.int64)   # Float * Int -> Float, .toInt truncates
    width = 100 // bins
    keys = percent - percent % width
    uniq, cnt = np.unique(keys, return_counts=True)
    return {int(k): int(c) for k, c in zip(uniq, cnt)}


def pileup_counts(ctx: Context, reads: PackedReads, loci_partitions, skip_empty: bool = True) -> Result:
    """pileupFlatMap(reads, lociPartitions, skipEmpty, p => depth / positiveDepth / referenceDepth / base counts)."""
    arr, n = _ranges(loci_partitions)
    h = C.c_void_p()
    ctx._check(lib().guac_pileup_counts(ctx._h, reads._h, arr, n, int(skip_empty), C.byref(h)))
    return Result(h, "counts")


def somatic_genotype_filter(records: np.ndarray, min_tumor_read_depth=0, max_tumor_read_depth=2 ** 31 - 1, min_normal_read_depth=0,
                            min_tumor_alternate_read_depth=0, min_lod=0, min_likelihood=0, min_vaf=0,
                            min_average_mapping_quality=0, min_average_base_quality=0, max_median_mismatches=2 ** 31 - 1,
                            seq_overload=False) -> np.ndarray:
    """SomaticGenotypeFilter over an array of guac_somatic_record (SOMATIC_DTYPE): boolean keep mask
    (filters/SomaticGenotypeFilter.scala:282-335; called after findPotentialVariantAtLocus, SomaticStandardCaller.scala:125-151)."""
    recs = np.ascontiguousarray(records)
    assert recs.dtype == SOMATIC_DTYPE
    prm = abi.SomaticFilterParamsC(min_tumor_read_depth, max_tumor_read_depth, min_normal_read_depth,
                                   min_tumor_alternate_read_depth, min_lod, min_likelihood, min_vaf,
                                   min_average_mapping_quality, min_average_base_quality, max_median_mismatches,
                                   int(seq_overload), 0)
    keep = np.zeros(len(recs), np.uint8)
    if len(recs):
        lib().guac_somatic_genotype_filter(recs.ctypes.data_as(C.c_void_p), len(recs), C.byref(prm), keep.ctypes.data_as(C.c_void_p))
    return keep.astype(bool)


DEPTH_BINS = 256


def depth_histogram(ctx: Context, reads: PackedReads, loci_partitions, fetch: bool = True) -> Optional[np.ndarray]:
    """guac_depth_histogram: hist[d] = requested loci covered by exactly d reads (last bin: deeper).  fetch = False leaves it
    on the device for Comm.reduce_depth_histogram."""
    arr, n = _ranges(loci_partitions)
    hist = np.zeros(DEPTH_BINS, np.uint64) if fetch else None
    ctx._check(lib().guac_depth_histogram(ctx._h, reads._h, arr, n,
                                          hist.ctypes.data_as(C.POINTER(C.c_uint64)) if fetch else None))
    return hist


class Comm:
    """guac_comm: NCCL communicator of the record gather, one per rank over that rank's Context.  `comm_id()` on one rank, the
    bytes handed to every rank by the host side (here: torch.distributed / any broadcast), then Comm(ctx, id, rank, world)."""

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = lib().guac_comm_unique_id(buf)
        if rc != 0:
            raise GuacError(rc, "ncclGetUniqueId failed")
        return bytes(buf)

    def __init__(self, ctx: Context, comm_id: bytes, rank: int, world: int):
        self.ctx, self.rank, self.world = ctx, rank, world
        self._h = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(comm_id)
        ctx._check(lib().guac_comm_create(ctx._h, buf, rank, world, C.byref(self._h)))

    def gather(self, result: Result, root: int = 0) -> Result:
        """guac_result_gather (collective): all ranks' germline-threshold records on `root`, in rank order."""
        h = C.c_void_p()
        self.ctx._check(lib().guac_result_gather(self._h, result._h.h, root, C.byref(h)))
        return Result(h, "threshold")

    def reduce_depth_histogram(self, root: int = 0) -> Optional[np.ndarray]:
        hist = np.zeros(DEPTH_BINS, np.uint64)
        self.ctx._check(lib().guac_comm_reduce_depth_histogram(self._h, root, hist.ctypes.data_as(C.POINTER(C.c_uint64))))
        return hist if self.rank == root else None

    def close(self):
        if self._h:
            lib().guac_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
