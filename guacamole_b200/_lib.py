"""Loads libguac_b200.so (the CUDA engine behind include/guac.h) and declares its C ABI for ctypes.

There is no CPU fallback: if the library is missing the import of any caller fails loudly, and on a machine without
a B200 `guac_ctx_create` returns GUAC_ERR_NO_DEVICE."""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GUAC_B200_LIBRARY") or os.path.join(_HERE, "libguac_b200.so")  # (override: A/B of two builds on one box)

# every symbol include/guac.h declares (tests/test_abi.py checks the library exports them all)
EXPORTED = [
    "guac_abi_version", "guac_ctx_create", "guac_ctx_destroy", "guac_last_error", "guac_status_string",
    "guac_ctx_set_option", "guac_ctx_timer_start", "guac_ctx_timer_stop", "guac_host_register", "guac_host_unregister",
    "guac_reads_pack", "guac_reads_pack_device", "guac_reads_pack_v2", "guac_read_batch_compact", "guac_host_batch_v2_view",
    "guac_host_batch_v2_bytes", "guac_host_batch_v2_free", "guac_bam_load", "guac_bam_last_error", "guac_host_batch_v2_contig_name",
    "guac_host_batch_v2_sample_name", "guac_host_batch_v2_decode_stats", "guac_reads_free", "guac_reads_count", "guac_reads_device_bytes",
    "guac_reads_order_sensitive_loci", "guac_reads_h2d_bytes", "guac_reads_pack_kernel_ms", "guac_reads_expand_kernel_ms",
    "guac_germline_threshold", "guac_somatic_standard", "guac_somatic_standard_filtered", "guac_germline_standard", "guac_pileup_counts",
    "guac_allele_counts", "guac_result_allele_counts",
    "guac_result_n", "guac_result_threshold_records", "guac_result_compact_records", "guac_result_somatic_records", "guac_result_counts",
    "guac_result_called_alleles",
    "guac_result_bytes", "guac_result_stats", "guac_result_free", "guac_partition_loci_uniformly",
    "guac_partition_loci_by_approximate_depth", "guac_depth_histogram", "guac_comm_unique_id", "guac_comm_create",
    "guac_comm_destroy", "guac_result_gather", "guac_comm_reduce_depth_histogram",
    "guac_somatic_genotype_filter",
]

_lib = None


class GuacError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{abi.STATUS_NAMES.get(code, code)}: {msg}")
        self.code = code


def _preload_nccl():
    """libguac_b200.so links libnccl.so.2 (the record gather).  In a process that also imports torch, the NCCL torch was built
    against must be the one the loader binds for that SONAME: load torch's bundled copy first when there is one (a host
    without torch, e.g. the JVM shim, simply gets the system library)."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            p = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(p):
                C.CDLL(p, mode=C.RTLD_GLOBAL)
    except Exception:
        pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(guacamole_b200 has no CPU fallback)")
    _preload_nccl()
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.guac_abi_version.restype = C.c_int
    L.guac_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.guac_ctx_destroy.argtypes = [vp]
    L.guac_ctx_destroy.restype = None
    L.guac_last_error.argtypes = [vp]
    L.guac_last_error.restype = C.c_char_p
    L.guac_ctx_set_option.argtypes = [vp, C.c_int, C.c_int64]
    L.guac_ctx_timer_start.argtypes = [vp]
    L.guac_ctx_timer_stop.argtypes = [vp, C.POINTER(C.c_double)]
    L.guac_host_register.argtypes = [vp, C.c_size_t]
    L.guac_host_unregister.argtypes = [vp]
    L.guac_status_string.argtypes = [C.c_int]
    L.guac_status_string.restype = C.c_char_p
    L.guac_reads_pack.argtypes = [vp, C.POINTER(abi.ReadBatchC), C.POINTER(abi.ReferenceC), C.POINTER(vp)]
    L.guac_reads_free.argtypes = [vp]
    L.guac_reads_pack_device.argtypes = [vp, C.POINTER(abi.ReadBatchC), C.POINTER(abi.ReferenceC), C.POINTER(vp)]
    L.guac_reads_pack_v2.argtypes = [vp, C.POINTER(abi.ReadBatchV2C), C.POINTER(abi.ReferenceC), C.POINTER(vp)]
    L.guac_read_batch_compact.argtypes = [C.POINTER(abi.ReadBatchC), C.c_int, C.c_int, C.POINTER(vp)]
    L.guac_host_batch_v2_view.argtypes = [vp]
    L.guac_host_batch_v2_view.restype = C.POINTER(abi.ReadBatchV2C)
    L.guac_host_batch_v2_bytes.argtypes = [vp]
    L.guac_host_batch_v2_bytes.restype = C.c_uint64
    L.guac_bam_load.argtypes = [C.c_char_p, C.POINTER(abi.BamOptionsC), C.POINTER(vp)]
    L.guac_bam_last_error.restype = C.c_char_p
    L.guac_host_batch_v2_contig_name.argtypes = [vp, C.c_uint32]
    L.guac_host_batch_v2_contig_name.restype = C.c_char_p
    L.guac_host_batch_v2_sample_name.argtypes = [vp]
    L.guac_host_batch_v2_sample_name.restype = C.c_char_p
    L.guac_host_batch_v2_decode_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.guac_host_batch_v2_decode_stats.restype = C.c_double
    L.guac_host_batch_v2_free.argtypes = [vp]
    L.guac_host_batch_v2_free.restype = None
    L.guac_reads_free.restype = None
    for f in ("guac_reads_count", "guac_reads_device_bytes", "guac_reads_order_sensitive_loci", "guac_reads_h2d_bytes"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = C.c_uint64
    L.guac_reads_pack_kernel_ms.argtypes = [vp]
    L.guac_reads_pack_kernel_ms.restype = C.c_double
    L.guac_reads_expand_kernel_ms.argtypes = [vp]
    L.guac_reads_expand_kernel_ms.restype = C.c_double
    L.guac_germline_threshold.argtypes = [vp, vp, C.POINTER(abi.LocusRangeC), C.c_size_t,
                                          C.POINTER(abi.ThresholdParamsC), C.POINTER(vp)]
    L.guac_somatic_standard.argtypes = [vp, vp, vp, C.POINTER(abi.LocusRangeC), C.c_size_t,
                                        C.POINTER(abi.SomaticParamsC), C.POINTER(vp)]
    L.guac_somatic_standard_filtered.argtypes = [vp, vp, vp, C.POINTER(abi.LocusRangeC), C.c_size_t,
                                                 C.POINTER(abi.SomaticParamsC), C.POINTER(abi.SomaticFilterParamsC), C.POINTER(vp)]
    L.guac_germline_standard.argtypes = [vp, vp, C.POINTER(abi.LocusRangeC), C.c_size_t,
                                         C.POINTER(abi.StandardParamsC), C.POINTER(vp)]
    L.guac_allele_counts.argtypes = [vp, vp, C.POINTER(abi.LocusRangeC), C.c_size_t, C.POINTER(vp)]
    L.guac_result_allele_counts.argtypes = [vp]
    L.guac_result_allele_counts.restype = C.POINTER(abi.AlleleCountC)
    L.guac_result_called_alleles.argtypes = [vp]
    L.guac_result_called_alleles.restype = C.POINTER(abi.CalledAlleleC)
    L.guac_pileup_counts.argtypes = [vp, vp, C.POINTER(abi.LocusRangeC), C.c_size_t, C.c_int, C.POINTER(vp)]
    L.guac_result_n.argtypes = [vp]
    L.guac_result_n.restype = C.c_size_t
    L.guac_result_threshold_records.argtypes = [vp]
    L.guac_result_threshold_records.restype = C.POINTER(abi.ThresholdRecordC)
    L.guac_result_compact_records.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(C.c_int32)]
    L.guac_result_compact_records.restype = C.c_size_t
    L.guac_result_somatic_records.argtypes = [vp]
    L.guac_result_somatic_records.restype = C.POINTER(abi.SomaticRecordC)
    L.guac_result_counts.argtypes = [vp]
    L.guac_result_counts.restype = C.POINTER(abi.LocusCountsC)
    L.guac_result_bytes.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.guac_result_bytes.restype = C.POINTER(C.c_uint8)
    L.guac_result_stats.argtypes = [vp]
    L.guac_result_stats.restype = C.POINTER(abi.StatsC)
    L.guac_result_free.argtypes = [vp]
    L.guac_result_free.restype = None
    L.guac_partition_loci_uniformly.argtypes = [C.c_int64, C.POINTER(abi.LocusRangeC), C.c_size_t,
                                                C.POINTER(abi.LocusRangeC), C.c_size_t, C.POINTER(C.c_size_t)]
    L.guac_partition_loci_by_approximate_depth.argtypes = [vp, C.c_int64, C.POINTER(abi.LocusRangeC), C.c_size_t, C.c_int64,
                                                           C.POINTER(vp), C.c_size_t, C.POINTER(abi.LocusRangeC), C.c_size_t,
                                                           C.POINTER(C.c_size_t)]
    L.guac_somatic_genotype_filter.argtypes = [vp, C.c_size_t, C.POINTER(abi.SomaticFilterParamsC), vp]
    L.guac_somatic_genotype_filter.restype = C.c_size_t
    L.guac_depth_histogram.argtypes = [vp, vp, C.POINTER(abi.LocusRangeC), C.c_size_t, C.POINTER(C.c_uint64)]
    L.guac_comm_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    L.guac_comm_create.argtypes = [vp, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.POINTER(vp)]
    L.guac_comm_destroy.argtypes = [vp]
    L.guac_comm_destroy.restype = None
    L.guac_result_gather.argtypes = [vp, vp, C.c_int, C.POINTER(vp)]
    L.guac_comm_reduce_depth_histogram.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64)]
    # device build of the synthetic generator (include/guac_synth.h)
    L.guac_synth_generate_device.argtypes = [vp, vp, C.POINTER(vp)]
    L.guac_reads_pack_synth.argtypes = [vp, vp, C.POINTER(abi.ReferenceC), C.POINTER(vp)]
    L.guac_synth_device_batch_view.argtypes = [vp]
    L.guac_synth_device_batch_view.restype = C.POINTER(abi.ReadBatchC)
    L.guac_synth_device_batch_ms.argtypes = [vp]
    L.guac_synth_device_batch_ms.restype = C.c_double
    L.guac_synth_device_batch_totals.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.guac_synth_device_batch_totals.restype = None
    L.guac_synth_device_batch_free.argtypes = [vp]
    L.guac_synth_device_batch_free.restype = None
    L.guac_synth_device_batch_download.argtypes = [vp, vp, C.c_int, C.POINTER(vp)]
    L.guac_synth_host_batch_view.argtypes = [vp]
    L.guac_synth_host_batch_view.restype = C.POINTER(abi.ReadBatchC)
    L.guac_synth_host_batch_free.argtypes = [vp]
    L.guac_synth_host_batch_free.restype = None
    _lib = L
    return L
