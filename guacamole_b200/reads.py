"""Columnar read batches: the host-side image of the reference's RDD[MappedRead].

`ReadBatch` holds the fields of MappedRead (reads/MappedRead.scala:35-48) as flat numpy columns in exactly the
layout `guac_read_batch` (include/guac.h) expects, so handing it to the C ABI is zero-copy.  The loaders mirror
the conventions of Read.fromSAMRecord (reads/Read.scala:217-291): 0-based start = POS-1, numeric qualities,
sample from the read group else "default", and the callers' input filters (Read.InputFilters :88-136).
"""
from __future__ import annotations

import ctypes as C
import gzip
import re
import struct
from dataclasses import dataclass, field
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import abi

_CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=X])")
_BAM_SEQ = "=ACMGRSVTWYHKDBN"


def parse_cigar(cigar: str) -> List[int]:
    """'4M3I4M' -> BAM-encoded ops (len << 4 | op)."""
    if cigar in ("*", ""):
        return []
    ops = []
    pos = 0
    for m in _CIGAR_RE.finditer(cigar):
        if m.start() != pos:
            raise ValueError(f"bad CIGAR {cigar!r}")
        pos = m.end()
        ops.append((int(m.group(1)) << 4) | abi.CIGAR_OPS.index(m.group(2)))
    if pos != len(cigar):
        raise ValueError(f"bad CIGAR {cigar!r}")
    return ops


def cigar_to_string(ops: Iterable[int]) -> str:
    return "".join(f"{int(o) >> 4}{abi.CIGAR_OPS[int(o) & 0xF]}" for o in ops)


@dataclass
class ReadRecord:
    """One MappedRead.  Defaults follow TestUtil.makeRead (src/test/.../util/TestUtil.scala:65-89)."""
    sequence: str
    cigar: str
    md: Optional[str]
    start: int = 1
    contig: str = "chr1"
    quals: Optional[Sequence[int]] = None   # numeric phred; default '@' = 31
    mapq: int = 30
    is_positive_strand: bool = True
    is_duplicate: bool = False
    failed_qc: bool = False
    is_paired: bool = False
    sample: str = "default"


def make_read(sequence, cigar, mdtag, start=1, chr="chr1", quality_scores=None, alignment_quality=30, **kw) -> ReadRecord:
    """TestUtil.makeRead with the reference's argument order."""
    return ReadRecord(sequence, cigar, mdtag, start, chr, quality_scores, alignment_quality, **kw)


@dataclass
class ReadBatch:
    contig_names: List[str]
    contig_lengths: Optional[np.ndarray]
    sample_names: List[str]
    contig: np.ndarray       # int32 [n]
    start: np.ndarray        # int64 [n]
    cigar_off: np.ndarray    # uint64 [n+1]
    cigar: np.ndarray        # uint32
    seq_off: np.ndarray      # uint64 [n+1]
    seq: np.ndarray          # uint8
    qual: np.ndarray         # uint8
    mapq: np.ndarray         # uint8 [n]
    flags: np.ndarray        # uint8 [n]
    sample: np.ndarray       # int32 [n]
    md_off: np.ndarray       # uint64 [n+1]
    md: np.ndarray           # uint8
    _keep: list = field(default_factory=list, repr=False)

    def __len__(self) -> int:
        return int(self.start.shape[0])

    @property
    def n_reads(self) -> int:
        return len(self)

    # ---- construction ------------------------------------------------------------------------------------------
    @staticmethod
    def from_records(records: Sequence[ReadRecord], contig_names: Optional[List[str]] = None,
                     contig_lengths: Optional[Sequence[int]] = None) -> "ReadBatch":
        names = list(contig_names) if contig_names is not None else []
        samples: List[str] = []
        contig, start, mapq, flags, sample = [], [], [], [], []
        cigar_off, seq_off, md_off = [0], [0], [0]
        cigar: List[int] = []
        seq = bytearray()
        qual = bytearray()
        md = bytearray()
        for r in records:
            if r.contig not in names:
                if contig_names is not None:
                    raise ValueError(f"unknown contig {r.contig}")
                names.append(r.contig)
            if r.sample not in samples:
                samples.append(r.sample)
            contig.append(names.index(r.contig))
            start.append(r.start)
            cigar.extend(parse_cigar(r.cigar))
            cigar_off.append(len(cigar))
            s = r.sequence.encode() if isinstance(r.sequence, str) else bytes(r.sequence)
            seq += s
            q = bytes([31] * len(s)) if r.quals is None else bytes(int(x) & 0xFF for x in r.quals)
            if len(q) != len(s):
                raise ValueError("Base qualities have length %d but sequence has length %d" % (len(q), len(s)))
            qual += q
            seq_off.append(len(seq))
            f = 0
            if r.is_positive_strand:
                f |= abi.READ_POSITIVE_STRAND
            if r.is_duplicate:
                f |= abi.READ_DUPLICATE
            if r.failed_qc:
                f |= abi.READ_FAILED_QC
            if r.is_paired:
                f |= abi.READ_PAIRED
            if r.md is not None:
                f |= abi.READ_HAS_MD
                md += r.md.encode()
            md_off.append(len(md))
            flags.append(f)
            mapq.append(r.mapq)
            sample.append(samples.index(r.sample))
        return ReadBatch(
            contig_names=names,
            contig_lengths=None if contig_lengths is None else np.asarray(contig_lengths, dtype=np.int64),
            sample_names=samples or ["default"],
            contig=np.asarray(contig, dtype=np.int32), start=np.asarray(start, dtype=np.int64),
            cigar_off=np.asarray(cigar_off, dtype=np.uint64), cigar=np.asarray(cigar, dtype=np.uint32),
            seq_off=np.asarray(seq_off, dtype=np.uint64), seq=np.frombuffer(bytes(seq), dtype=np.uint8).copy(),
            qual=np.frombuffer(bytes(qual), dtype=np.uint8).copy(), mapq=np.asarray(mapq, dtype=np.uint8),
            flags=np.asarray(flags, dtype=np.uint8), sample=np.asarray(sample, dtype=np.int32),
            md_off=np.asarray(md_off, dtype=np.uint64), md=np.frombuffer(bytes(md), dtype=np.uint8).copy())

    # ---- views ---------------------------------------------------------------------------------------------------
    def ref_length(self) -> np.ndarray:
        """Cigar.getPaddedReferenceLength without P (MappedRead.end, reads/MappedRead.scala:87)."""
        op = self.cigar & 0xF
        ln = (self.cigar >> 4).astype(np.int64)
        consumes = np.isin(op, [0, 2, 3, 7, 8])
        contrib = np.where(consumes, ln, 0)
        csum = np.concatenate([[0], np.cumsum(contrib)])
        return (csum[self.cigar_off[1:].astype(np.int64)] - csum[self.cigar_off[:-1].astype(np.int64)]).astype(np.int64)

    def end(self) -> np.ndarray:
        return self.start + self.ref_length()

    def select(self, index: np.ndarray) -> "ReadBatch":
        """Rows `index` (bool mask or integer indices), order preserved as given."""
        idx = np.asarray(index)
        if idx.dtype == bool:
            idx = np.nonzero(idx)[0]
        idx = idx.astype(np.int64)

        def gather(off, data):
            lo = off[:-1].astype(np.int64)[idx]
            hi = off[1:].astype(np.int64)[idx]
            ln = hi - lo
            new_off = np.concatenate([[0], np.cumsum(ln)]).astype(np.uint64)
            if len(idx) == 0 or ln.sum() == 0:
                return new_off, data[:0].copy()
            pos = np.repeat(lo - new_off[:-1].astype(np.int64), ln) + np.arange(int(ln.sum()), dtype=np.int64)
            return new_off, data[pos]

        cigar_off, cigar = gather(self.cigar_off, self.cigar)
        seq_off, seq = gather(self.seq_off, self.seq)
        _, qual = gather(self.seq_off, self.qual)
        md_off, md = gather(self.md_off, self.md)
        return ReadBatch(self.contig_names, self.contig_lengths, self.sample_names, self.contig[idx].copy(),
                         self.start[idx].copy(), cigar_off, cigar, seq_off, seq, qual, self.mapq[idx].copy(),
                         self.flags[idx].copy(), self.sample[idx].copy(), md_off, md)

    def sorted(self) -> "ReadBatch":
        """Stable sort by (contig, start) — the order windowTaskFlatMapMultipleRDDs establishes
        (DistributedUtil.scala:518-529, 621-626)."""
        order = np.lexsort((self.start, self.contig))
        return self.select(order)

    def filtered(self, non_duplicate=False, passed_qc=False, has_md=False, is_paired=False) -> "ReadBatch":
        """Read.InputFilters (reads/Read.scala:88-136) on the flag bits."""
        keep = np.ones(len(self), dtype=bool)
        if non_duplicate:
            keep &= (self.flags & abi.READ_DUPLICATE) == 0
        if passed_qc:
            keep &= (self.flags & abi.READ_FAILED_QC) == 0
        if has_md:
            keep &= (self.flags & abi.READ_HAS_MD) != 0
        if is_paired:
            keep &= (self.flags & abi.READ_PAIRED) != 0
        return self.select(keep)

    def record(self, i: int) -> ReadRecord:
        co, ce = int(self.cigar_off[i]), int(self.cigar_off[i + 1])
        so, se = int(self.seq_off[i]), int(self.seq_off[i + 1])
        mo, me = int(self.md_off[i]), int(self.md_off[i + 1])
        f = int(self.flags[i])
        return ReadRecord(
            sequence=self.seq[so:se].tobytes().decode("latin1"), cigar=cigar_to_string(self.cigar[co:ce]),
            md=self.md[mo:me].tobytes().decode() if f & abi.READ_HAS_MD else None, start=int(self.start[i]),
            contig=self.contig_names[int(self.contig[i])], quals=list(self.qual[so:se]), mapq=int(self.mapq[i]),
            is_positive_strand=bool(f & abi.READ_POSITIVE_STRAND), is_duplicate=bool(f & abi.READ_DUPLICATE),
            failed_qc=bool(f & abi.READ_FAILED_QC), is_paired=bool(f & abi.READ_PAIRED),
            sample=self.sample_names[int(self.sample[i])])

    # ---- C view -----------------------------------------------------------------------------------------------------
    def to_c(self) -> abi.ReadBatchC:
        """A guac_read_batch pointing at this batch's arrays (keep `self` alive while it is in use).  The arrays the last call
        pinned are released first: repeated packs of one batch do not pile references up."""
        self._keep.clear()

        def ptr(a, t):
            a = np.ascontiguousarray(a)
            self._keep.append(a)
            return a.ctypes.data_as(C.POINTER(t))

        b = abi.ReadBatchC()
        b.n_reads = len(self)
        b.n_contigs = len(self.contig_names)
        b.contig_length = ptr(self.contig_lengths, C.c_int64) if self.contig_lengths is not None else None
        b.contig = ptr(self.contig, C.c_int32)
        b.start = ptr(self.start, C.c_int64)
        b.cigar_off = ptr(self.cigar_off, C.c_uint64)
        b.cigar = ptr(self.cigar if len(self.cigar) else np.zeros(1, np.uint32), C.c_uint32)
        b.seq_off = ptr(self.seq_off, C.c_uint64)
        b.seq = ptr(self.seq if len(self.seq) else np.zeros(1, np.uint8), C.c_uint8)
        b.qual = ptr(self.qual if len(self.qual) else np.zeros(1, np.uint8), C.c_uint8)
        b.mapq = ptr(self.mapq if len(self.mapq) else np.zeros(1, np.uint8), C.c_uint8)
        b.flags = ptr(self.flags if len(self.flags) else np.zeros(1, np.uint8), C.c_uint8)
        b.sample = ptr(self.sample if len(self.sample) else np.zeros(1, np.int32), C.c_int32)
        b.md_off = ptr(self.md_off, C.c_uint64)
        md = np.ascontiguousarray(self.md if len(self.md) else np.zeros(1, np.uint8))
        self._keep.append(md)
        b.md = C.cast(md.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_char_p)
        return b

    # ---- (de)serialisation for committed fixtures -----------------------------------------------------------------------
    def save_npz(self, path: str) -> None:
        np.savez_compressed(
            path, contig_names=np.array(self.contig_names), sample_names=np.array(self.sample_names),
            contig_lengths=self.contig_lengths if self.contig_lengths is not None else np.zeros(0, np.int64),
            contig=self.contig, start=self.start, cigar_off=self.cigar_off, cigar=self.cigar, seq_off=self.seq_off,
            seq=self.seq, qual=self.qual, mapq=self.mapq, flags=self.flags, sample=self.sample, md_off=self.md_off,
            md=self.md)

    @staticmethod
    def load_npz(path: str) -> "ReadBatch":
        z = np.load(path, allow_pickle=False)
        cl = z["contig_lengths"]
        return ReadBatch([str(x) for x in z["contig_names"]], cl if len(cl) else None,
                         [str(x) for x in z["sample_names"]], z["contig"], z["start"], z["cigar_off"], z["cigar"],
                         z["seq_off"], z["seq"], z["qual"], z["mapq"], z["flags"], z["sample"], z["md_off"], z["md"])


def concat(batches: Sequence[ReadBatch]) -> ReadBatch:
    first = batches[0]

    def cat_off(offs):
        out = [np.zeros(1, np.uint64)]
        base = 0
        for o in offs:
            out.append(o[1:] + np.uint64(base))
            base += int(o[-1])
        return np.concatenate(out).astype(np.uint64)

    return ReadBatch(first.contig_names, first.contig_lengths, first.sample_names,
                     np.concatenate([b.contig for b in batches]), np.concatenate([b.start for b in batches]),
                     cat_off([b.cigar_off for b in batches]), np.concatenate([b.cigar for b in batches]),
                     cat_off([b.seq_off for b in batches]), np.concatenate([b.seq for b in batches]),
                     np.concatenate([b.qual for b in batches]), np.concatenate([b.mapq for b in batches]),
                     np.concatenate([b.flags for b in batches]), np.concatenate([b.sample for b in batches]),
                     cat_off([b.md_off for b in batches]), np.concatenate([b.md for b in batches]))


# ---- SAM / BAM front end (mapped reads only) ---------------------------------------------------------------------------
def _flags_from_sam(flag: int, has_md: bool) -> int:
    f = 0
    if not flag & 0x10:
        f |= abi.READ_POSITIVE_STRAND
    if flag & 0x400:
        f |= abi.READ_DUPLICATE
    if flag & 0x200:
        f |= abi.READ_FAILED_QC
    if flag & 0x1:
        f |= abi.READ_PAIRED
    if has_md:
        f |= abi.READ_HAS_MD
    return f


def _header_info(text: str):
    names, lengths, rg_sample = [], [], {}
    for line in text.splitlines():
        if line.startswith("@SQ"):
            d = dict(x.split(":", 1) for x in line.split("\t")[1:] if ":" in x)
            names.append(d["SN"])
            lengths.append(int(d["LN"]))
        elif line.startswith("@RG"):
            d = dict(x.split(":", 1) for x in line.split("\t")[1:] if ":" in x)
            if "ID" in d and "SM" in d:
                rg_sample[d["ID"]] = d["SM"]
    return names, lengths, rg_sample


def _records_to_batch(recs, names, lengths) -> ReadBatch:
    return ReadBatch.from_records(recs, contig_names=names, contig_lengths=lengths)


def load_sam(path: str) -> ReadBatch:
    """Mapped reads of a SAM text file, in file order (Read.fromSAMRecord, reads/Read.scala:217-291)."""
    header, lines = [], []
    with open(path) as fh:
        for line in fh:
            (header if line.startswith("@") else lines).append(line.rstrip("\n"))
    names, lengths, rg_sample = _header_info("\n".join(header))
    recs = []
    for line in lines:
        if not line:
            continue
        t = line.split("\t")
        flag, rname, pos, mapq, cigar, seq, qual = int(t[1]), t[2], int(t[3]), int(t[4]), t[5], t[9], t[10]
        if flag & 0x4 or rname == "*" or pos < 1 or cigar == "*":
            continue  # unmapped
        md, rg = None, None
        for tag in t[11:]:
            if tag.startswith("MD:Z:"):
                md = tag[5:]
            elif tag.startswith("RG:Z:"):
                rg = tag[5:]
        quals = [0] * len(seq) if qual == "*" else [ord(c) - 33 for c in qual]
        if rname not in names:
            names.append(rname)
            lengths.append(0)
        recs.append(ReadRecord(seq, cigar, md, pos - 1, rname, quals, mapq,
                               is_positive_strand=not flag & 0x10, is_duplicate=bool(flag & 0x400),
                               failed_qc=bool(flag & 0x200), is_paired=bool(flag & 0x1),
                               sample=rg_sample.get(rg, "default")))
    return _records_to_batch(recs, names, lengths)


def load_bam(path: str) -> ReadBatch:
    """Mapped reads of a BAM file, in file order.  BGZF is a concatenation of gzip members."""
    data = gzip.open(path, "rb").read()
    if data[:4] != b"BAM\x01":
        raise ValueError("not a BAM file")
    p = 4
    (l_text,) = struct.unpack_from("<i", data, p)
    p += 4
    text = data[p:p + l_text].split(b"\0")[0].decode()
    p += l_text
    (n_ref,) = struct.unpack_from("<i", data, p)
    p += 4
    names, lengths = [], []
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", data, p)
        p += 4
        names.append(data[p:p + l_name - 1].decode())
        p += l_name
        (l_ref,) = struct.unpack_from("<i", data, p)
        p += 4
        lengths.append(l_ref)
    _, _, rg_sample = _header_info(text)
    recs = []
    n = len(data)
    while p < n:
        (block_size,) = struct.unpack_from("<i", data, p)
        p += 4
        end = p + block_size
        ref_id, pos, l_read_name, mapq, _bin, n_cigar, flag, l_seq, _nref, _npos, _tlen = struct.unpack_from(
            "<iiBBHHHiiii", data, p)
        q = p + 32 + l_read_name
        cigar = struct.unpack_from("<%dI" % n_cigar, data, q)
        q += 4 * n_cigar
        packed = data[q:q + (l_seq + 1) // 2]
        q += (l_seq + 1) // 2
        seq = "".join(_BAM_SEQ[(packed[i >> 1] >> (4 if i % 2 == 0 else 0)) & 0xF] for i in range(l_seq))
        quals = list(data[q:q + l_seq])
        q += l_seq
        if quals and quals[0] == 0xFF:
            quals = [0] * l_seq
        md, rg = None, None
        while q < end:
            tag = data[q:q + 2]
            typ = chr(data[q + 2])
            q += 3
            if typ == "Z" or typ == "H":
                e = data.index(b"\0", q)
                val = data[q:e].decode()
                q = e + 1
                if tag == b"MD":
                    md = val
                elif tag == b"RG":
                    rg = val
            elif typ in "AcC":
                q += 1
            elif typ in "sS":
                q += 2
            elif typ in "iIf":
                q += 4
            elif typ == "B":
                sub = chr(data[q])
                (cnt,) = struct.unpack_from("<i", data, q + 1)
                q += 5 + cnt * {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[sub]
            else:
                raise ValueError("bad BAM tag type " + typ)
        p = end
        if flag & 0x4 or ref_id < 0 or pos < 0 or n_cigar == 0:
            continue
        recs.append(ReadRecord(seq, cigar_to_string(cigar), md, pos, names[ref_id], quals, mapq,
                               is_positive_strand=not flag & 0x10, is_duplicate=bool(flag & 0x400),
                               failed_qc=bool(flag & 0x200), is_paired=bool(flag & 0x1),
                               sample=rg_sample.get(rg, "default")))
    return _records_to_batch(recs, names, lengths)


def load_reads(path: str) -> ReadBatch:
    return load_bam(path) if path.endswith(".bam") else load_sam(path)


def write_bam(batch: ReadBatch, path: str, read_groups: Optional[dict] = None, block_bytes: int = 0xFF00, level: int = 6) -> None:
    """Write the batch as a BAM file (BGZF members of `block_bytes` uncompressed bytes, SAM spec 4.1 / 4.2): the test and
    bench input of guac_bam_load.  Samples other than "default" get one read group each (ID = SM = sample name)."""
    import zlib
    names = list(batch.contig_names)
    lengths = [int(x) for x in batch.contig_lengths] if batch.contig_lengths is not None else [0] * len(names)
    ends = batch.end() if len(batch) else np.zeros(0, np.int64)
    for c in range(len(names)):
        if lengths[c] <= 0:
            sel = batch.contig == c
            lengths[c] = int(ends[sel].max()) + 1 if sel.any() else 1
    text = "@HD\tVN:1.4\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in zip(names, lengths))
    text += "".join(f"@RG\tID:{s}\tSM:{s}\n" for s in batch.sample_names if s != "default")
    out = bytearray(b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(names)))
    for n, l in zip(names, lengths):
        out += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l)
    code = np.full(256, 15, np.uint8)
    for k, ch in enumerate(b"=ACMGRSVTWYHKDBN"):
        code[ch] = k
    for i in range(len(batch)):
        co, ce = int(batch.cigar_off[i]), int(batch.cigar_off[i + 1])
        so, se = int(batch.seq_off[i]), int(batch.seq_off[i + 1])
        mo, me = int(batch.md_off[i]), int(batch.md_off[i + 1])
        f = int(batch.flags[i])
        flag = (0 if f & abi.READ_POSITIVE_STRAND else 0x10) | (0x400 if f & abi.READ_DUPLICATE else 0) | \
               (0x200 if f & abi.READ_FAILED_QC else 0) | (0x1 if f & abi.READ_PAIRED else 0)
        name = b"r%d\0" % i
        nib = code[batch.seq[so:se]]
        if len(nib) & 1:
            nib = np.append(nib, np.uint8(0))
        packed = ((nib[0::2] << 4) | nib[1::2]).astype(np.uint8).tobytes()
        aux = b""
        if f & abi.READ_HAS_MD:
            aux += b"MDZ" + batch.md[mo:me].tobytes() + b"\0"
        sample = batch.sample_names[int(batch.sample[i])] if len(batch.sample_names) else "default"
        if sample != "default":
            aux += b"RGZ" + sample.encode() + b"\0"
        aux += b"NMi" + struct.pack("<i", 0)
        body = struct.pack("<iiBBHHHiiii", int(batch.contig[i]), int(batch.start[i]), len(name), int(batch.mapq[i]), 4680, ce - co, flag,
                           se - so, -1, -1, 0) + name + batch.cigar[co:ce].astype("<u4").tobytes() + packed + \
            batch.qual[so:se].tobytes() + aux
        out += struct.pack("<i", len(body)) + body

    def member(chunk: bytes) -> bytes:
        z = zlib.compressobj(level, zlib.DEFLATED, -15)
        data = z.compress(chunk) + z.flush()
        head = b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(data) + 25)
        return head + data + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))

    with open(path, "wb") as fh:
        for a in range(0, len(out), block_bytes):
            fh.write(member(bytes(out[a:a + block_bytes])))
        fh.write(member(b""))
