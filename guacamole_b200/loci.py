"""LociSet / LociMap as the hot path needs them: flat (contig, start, end[, task]) ranges.

Mirrors LociSet.parse / Builder.result (LociSet.scala:166-217: "all" leaves out the LAST base of every contig,
:205-207) and DistributedUtil.partitionLociUniformly (DistributedUtil.scala:83-108, done in C by
guac_partition_loci_uniformly)."""
from __future__ import annotations

import ctypes as C
import re
from typing import Dict, List, Optional, Sequence, Tuple

from . import abi

Range = Tuple[int, int, int]            # (contig index, start, end)
TaskRange = Tuple[int, int, int, int]   # (contig index, start, end, task)

_CONTIG_AND_LOCI = re.compile(r"^([\w.]+):(\d+)-(\d+)$")
_CONTIG_ONLY = re.compile(r"^([\w.]+)$")


def parse_loci(expr: str, contig_names: Sequence[str], contig_lengths: Optional[Sequence[int]] = None) -> List[Range]:
    """LociSet.parse(expr).result(contigLengths) as sorted, merged ranges over contig INDICES."""
    names = list(contig_names)
    out: Dict[int, List[Tuple[int, int]]] = {}
    if expr == "all":
        if contig_lengths is None:
            raise ValueError("'all' needs contig lengths")
        for i, ln in enumerate(contig_lengths):
            if ln - 1 > 0:
                out.setdefault(i, []).append((0, int(ln) - 1))
    else:
        for part in re.sub(r"\s", "", expr).split(","):
            if part == "":
                continue
            m = _CONTIG_AND_LOCI.match(part)
            if m:
                name, s, e = m.group(1), int(m.group(2)), int(m.group(3))
            else:
                m = _CONTIG_ONLY.match(part)
                if not m:
                    raise ValueError("Couldn't parse loci range: %s" % part)
                name, s, e = m.group(1), 0, None
            if name not in names:
                raise ValueError("No such contig: %s" % name)
            i = names.index(name)
            if e is None:
                if contig_lengths is None:
                    raise ValueError("open range needs contig lengths")
                e = int(contig_lengths[i])
            if contig_lengths is not None and e > contig_lengths[i]:
                raise ValueError("Invalid range %d-%d for contig '%s' which has length %d" % (s, e, name, contig_lengths[i]))
            if e > s:
                out.setdefault(i, []).append((s, e))
    ranges: List[Range] = []
    for i in sorted(out):
        merged: List[List[int]] = []
        for s, e in sorted(out[i]):
            if merged and s <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], e)
            else:
                merged.append([s, e])
        ranges.extend((i, s, e) for s, e in merged)
    return ranges


def ranges_to_c(ranges: Sequence[tuple]):
    arr = (abi.LocusRangeC * max(1, len(ranges)))()
    for i, r in enumerate(ranges):
        arr[i].contig, arr[i].start, arr[i].end = int(r[0]), int(r[1]), int(r[2])
        arr[i].task = int(r[3]) if len(r) > 3 else 0
    return arr


def partition_loci_uniformly(tasks: int, loci: Sequence[Range]) -> List[TaskRange]:
    """DistributedUtil.partitionLociUniformly: LociMap[Long] as (contig, start, end, task) ranges."""
    from ._lib import GuacError, lib
    L = lib()
    arr = ranges_to_c(loci)
    n = C.c_size_t()
    cap = len(loci) + int(tasks) + 8
    out = (abi.LocusRangeC * cap)()
    rc = L.guac_partition_loci_uniformly(tasks, arr, len(loci), out, cap, C.byref(n))
    if rc != 0:
        raise GuacError(rc, "partitionLociUniformly failed")
    return [(out[i].contig, out[i].start, out[i].end, out[i].task) for i in range(n.value)]


def partition_loci_by_approximate_depth(ctx, tasks: int, loci: Sequence[Range], accuracy: int, *read_sets) -> List[TaskRange]:
    """DistributedUtil.partitionLociByApproximateDepth (DistributedUtil.scala:162-251): loci assigned to tasks so that every
    task sees about the same number of the reads packed in `read_sets` (PackedReads).  `loci` in LociSet order."""
    from ._lib import lib
    L = lib()
    arr = ranges_to_c(loci)
    handles = (C.c_void_p * len(read_sets))(*[r._h for r in read_sets])
    total = sum(r[2] - r[1] for r in loci)
    cap = len(loci) + 2 * int(min(int(tasks) * int(accuracy), total)) + 8
    out = (abi.LocusRangeC * cap)()
    n = C.c_size_t()
    ctx._check(L.guac_partition_loci_by_approximate_depth(ctx._h, tasks, arr, len(loci), accuracy, handles, len(read_sets),
                                                          out, cap, C.byref(n)))
    return [(out[i].contig, out[i].start, out[i].end, out[i].task) for i in range(n.value)]
