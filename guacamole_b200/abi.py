"""ctypes mirror of include/guac.h — the plain-C data contract of the pileup-and-call engine.

Structure layouts follow include/guac.h field by field; tests/test_abi.py checks sizes and offsets against the
compiled library (guac_abi_sizeof).  No torch types cross this boundary.
"""
import ctypes as C

GUAC_ABI_VERSION = 2

# guac_status
OK = 0
ERR_INVALID_ARGUMENT = 1
ERR_UNSORTED_READS = 2
ERR_CONTIG_ORDER = 3
ERR_INVALID_CIGAR = 4
ERR_MISSING_MD = 5
ERR_MULTIPLE_REFERENCE_BASES = 6
ERR_BAD_QUALITY = 7
ERR_CUDA = 8
ERR_OOM = 9
ERR_NO_DEVICE = 10
ERR_UNSUPPORTED = 11

STATUS_NAMES = {
    0: "GUAC_OK", 1: "GUAC_ERR_INVALID_ARGUMENT", 2: "GUAC_ERR_UNSORTED_READS", 3: "GUAC_ERR_CONTIG_ORDER",
    4: "GUAC_ERR_INVALID_CIGAR", 5: "GUAC_ERR_MISSING_MD", 6: "GUAC_ERR_MULTIPLE_REFERENCE_BASES",
    7: "GUAC_ERR_BAD_QUALITY", 8: "GUAC_ERR_CUDA", 9: "GUAC_ERR_OOM", 10: "GUAC_ERR_NO_DEVICE",
    11: "GUAC_ERR_UNSUPPORTED",
}

READ_POSITIVE_STRAND = 0x01
READ_DUPLICATE = 0x02
READ_FAILED_QC = 0x04
READ_HAS_MD = 0x08
READ_PAIRED = 0x10

CIGAR_OPS = "MIDNSHP=X"

GT_REF, GT_ALT, GT_OTHER_ALT, GT_NO_CALL = 0, 1, 2, 3
GT_NAMES = {0: "Ref", 1: "Alt", 2: "OtherAlt", 3: "NoCall"}


class ReadBatchC(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint64),
        ("n_contigs", C.c_uint32),
        ("contig_length", C.POINTER(C.c_int64)),
        ("contig", C.POINTER(C.c_int32)),
        ("start", C.POINTER(C.c_int64)),
        ("cigar_off", C.POINTER(C.c_uint64)),
        ("cigar", C.POINTER(C.c_uint32)),
        ("seq_off", C.POINTER(C.c_uint64)),
        ("seq", C.POINTER(C.c_uint8)),
        ("qual", C.POINTER(C.c_uint8)),
        ("mapq", C.POINTER(C.c_uint8)),
        ("flags", C.POINTER(C.c_uint8)),
        ("sample", C.POINTER(C.c_int32)),
        ("md_off", C.POINTER(C.c_uint64)),
        ("md", C.c_char_p),
    ]


class ReadBatchV2C(C.Structure):
    """guac_read_batch_v2: the columns at their BAM width (4-bit bases, 32-bit starts and offsets)."""
    _fields_ = [
        ("n_reads", C.c_uint64),
        ("n_contigs", C.c_uint32),
        ("read_length", C.c_uint32),
        ("contig_length", C.POINTER(C.c_int64)),
        ("contig_read_off", C.POINTER(C.c_uint64)),
        ("start", C.POINTER(C.c_int32)),
        ("cigar_off", C.POINTER(C.c_uint32)),
        ("cigar", C.POINTER(C.c_uint32)),
        ("seq_off", C.POINTER(C.c_uint32)),
        ("seq4", C.POINTER(C.c_uint8)),
        ("qual", C.POINTER(C.c_uint8)),
        ("mapq", C.POINTER(C.c_uint8)),
        ("flags", C.POINTER(C.c_uint8)),
        ("md_off", C.POINTER(C.c_uint32)),
        ("md", C.c_char_p),
        ("sample", C.c_int32),
        ("reserved", C.c_int32),
    ]


class BamOptionsC(C.Structure):
    _fields_ = [("n_threads", C.c_int32), ("non_duplicate", C.c_int32), ("passed_qc", C.c_int32), ("has_md_tag", C.c_int32),
                ("is_paired", C.c_int32), ("with_qualities", C.c_int32), ("pinned", C.c_int32), ("reserved", C.c_int32),
                ("sample", C.c_char_p)]


class ReferenceC(C.Structure):
    _fields_ = [
        ("n_contigs", C.c_uint32),
        ("base_off", C.POINTER(C.c_uint64)),
        ("bases", C.POINTER(C.c_uint8)),
    ]


class LocusRangeC(C.Structure):
    _fields_ = [("contig", C.c_int32), ("task", C.c_int32), ("start", C.c_int64), ("end", C.c_int64)]


class ThresholdParamsC(C.Structure):
    _fields_ = [("threshold_percent", C.c_int32), ("emit_ref", C.c_int32), ("emit_no_call", C.c_int32),
                ("skip_empty", C.c_int32)]


class SomaticParamsC(C.Structure):
    _fields_ = [("odds_threshold", C.c_int32), ("min_alignment_quality", C.c_int32),
                ("filter_multi_allelic", C.c_int32), ("max_read_depth", C.c_int32), ("skip_empty", C.c_int32)]


class StandardParamsC(C.Structure):
    _fields_ = [("min_alignment_quality", C.c_int32), ("skip_empty", C.c_int32)]


class ThresholdRecordC(C.Structure):
    _fields_ = [("start", C.c_int64), ("contig", C.c_int32), ("sample", C.c_int32), ("ref_off", C.c_uint32),
                ("alt_off", C.c_uint32), ("ref_len", C.c_uint16), ("alt_len", C.c_uint16), ("gt", C.c_uint8 * 2),
                ("tie", C.c_uint8), ("pad_", C.c_uint8)]


class AlleleEvidenceC(C.Structure):
    _fields_ = [("likelihood", C.c_double), ("mean_mapping_quality", C.c_double),
                ("median_mapping_quality", C.c_double), ("mean_base_quality", C.c_double),
                ("median_base_quality", C.c_double), ("median_mismatches_per_read", C.c_double),
                ("read_depth", C.c_int32), ("allele_read_depth", C.c_int32), ("forward_depth", C.c_int32),
                ("allele_forward_depth", C.c_int32)]


class SomaticRecordC(C.Structure):
    _fields_ = [("start", C.c_int64), ("contig", C.c_int32), ("sample", C.c_int32), ("ref_off", C.c_uint32),
                ("alt_off", C.c_uint32), ("ref_len", C.c_uint16), ("alt_len", C.c_uint16),
                ("phred_scaled_somatic_likelihood", C.c_int32), ("somatic_log_odds", C.c_double),
                ("tumor", AlleleEvidenceC), ("normal", AlleleEvidenceC)]


class AlleleCountC(C.Structure):
    _fields_ = [("start", C.c_int64), ("contig", C.c_int32), ("sample", C.c_int32), ("ref_off", C.c_uint32),
                ("alt_off", C.c_uint32), ("ref_len", C.c_uint16), ("alt_len", C.c_uint16), ("count", C.c_int32)]


class CalledAlleleC(C.Structure):
    _fields_ = [("start", C.c_int64), ("contig", C.c_int32), ("sample", C.c_int32), ("ref_off", C.c_uint32),
                ("alt_off", C.c_uint32), ("ref_len", C.c_uint16), ("alt_len", C.c_uint16),
                ("phred_scaled_likelihood", C.c_int32), ("evidence", AlleleEvidenceC)]


class SomaticFilterParamsC(C.Structure):
    _fields_ = [("min_tumor_read_depth", C.c_int32), ("max_tumor_read_depth", C.c_int32),
                ("min_normal_read_depth", C.c_int32), ("min_tumor_alternate_read_depth", C.c_int32),
                ("min_lod", C.c_int32), ("min_likelihood", C.c_int32), ("min_vaf", C.c_int32),
                ("min_average_mapping_quality", C.c_int32), ("min_average_base_quality", C.c_int32),
                ("max_median_mismatches", C.c_int32), ("seq_overload", C.c_int32), ("pad_", C.c_int32)]


class LocusCountsC(C.Structure):
    _fields_ = [("locus", C.c_int64), ("contig", C.c_int32), ("depth", C.c_int32), ("positive_depth", C.c_int32),
                ("reference_depth", C.c_int32), ("base_count", C.c_int32 * 4), ("other_count", C.c_int32),
                ("reference_base", C.c_uint8), ("pad_", C.c_uint8 * 3)]


class StatsC(C.Structure):
    _fields_ = [("reads_total", C.c_uint64), ("reads_relevant", C.c_uint64), ("reads_expanded", C.c_uint64),
                ("loci_requested", C.c_uint64), ("loci_visited", C.c_uint64), ("records", C.c_uint64),
                ("tie_loci", C.c_uint64), ("order_sensitive_loci", C.c_uint64), ("kernel_ms", C.c_double),
                ("kernel_launches", C.c_uint64), ("tile_kernel_ms", C.c_double), ("exact_kernel_ms", C.c_double),
                ("exact_loci", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]

OPT_SORT_RECORDS = 1
OPT_PACK_QUALITIES = 2
OPT_HOST_THREADS = 3
OPT_DIFFERENCE_LISTS = 4
OPT_SEGMENTS = 5
OPT_TRIM_CACHE = 6
OPT_PACK_OVERLAP = 7


def struct_to_dict(s):
    out = {}
    for name, _ in s._fields_:
        v = getattr(s, name)
        if isinstance(v, C.Structure):
            v = struct_to_dict(v)
        elif isinstance(v, C.Array):
            v = list(v)
        out[name] = v
    return out
