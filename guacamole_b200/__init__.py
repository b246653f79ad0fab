"""guacamole_b200 — a B200-native (sm_100a) pileup-and-call engine behind Guacamole's caller interface.

Only the hot path lives here: packed read store, bit-sliced pileup kernels, the germline-threshold and
somatic-standard callers, loci partitioning.  See DESIGN.md."""
from . import abi  # noqa: F401
from .reads import ReadBatch, ReadRecord, make_read, load_reads  # noqa: F401

__all__ = ["abi", "ReadBatch", "ReadRecord", "make_read", "load_reads"]
