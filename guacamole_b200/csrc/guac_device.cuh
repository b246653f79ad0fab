// guac_device.cuh — device-side data layout and helpers of the B200 pileup-and-call engine (sm_100a only).
//
// Layout in HBM (one guac_reads object = one sample's start-sorted reads):
//   rec[n+1]      16 B/read   {start, end, pair_off, info}   contig-relative 0-based [start, end)   (MappedRead.start/end)
//   pairs[]       8 B/32 bases  bit-plane pairs (lo, hi) of the 2-bit read bases, read coordinates, 32 bases per pair
//   xmask[]       4 B/32 bases  1 = base is not A/C/G/T (read only for reads flagged HAS_EXC)
//   gs_hdr[G]     16 B/granule  {offset of the granule's difference stream, its length, depth entering the granule (all reads,
//                             positive strand)}: the reads of a granule as their DIFFERENCES against the reference track
//   gs_diffs[]    2 B/difference  (locus offset in the granule, class) as the counter slot it lands in, contiguous per granule
//                             (reference-based encoding, like CRAM): class 1..3 = read base code ^ reference base code,
//                             0 = element that is not a plain base
//   gs_dd/gs_dp[] 1 B/locus   reads starting (low nibble) / ending (high nibble) at the locus, all reads / positive strand
//                             (wide stores: 4 B/locus, two 16-bit fields): the depth is their running sum
//   cig_off/cigar BAM-encoded run-length CIGAR ops (read only for reads that are not SIMPLE)
//   seq/qual      raw bytes (the exact per-locus paths read them)
//   qc[]          1 B/base   quality (6 bits) | base code << 6: the one byte per pileup element the likelihood kernel loads
//   md_off/md     upper-cased MD strings (deleted bases for the exact per-locus path)
//   trk_lo/hi/std reference track per contig as three bit-planes, 32 loci per word (MD- or FASTA-derived)
//   gran_first/last per 1024-loci granule: the range of read indices that can overlap it
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/guac.h"

namespace guac {

constexpr int kGranuleShift = 10;               // 1024 loci per granule
constexpr int kGranuleLoci = 1 << kGranuleShift;
constexpr int kTileWords = 128;                 // somatic kernel: one CTA per 128 words = 4096 loci
constexpr int kTileLoci = kTileWords * 32;
constexpr int kChunkReads = 512;                // reads staged per chunk
constexpr int kChunkPairs = 4096;               // plane pairs staged per chunk (32 KB)
constexpr int kMaxReadLen = 65535;

// ReadRec.info bits
constexpr uint32_t kInfoLeadMask = 0xFFFFu;     // SIMPLE reads: read bases skipped before the aligned segment
constexpr uint32_t kInfoSimple = 1u << 16;      // one M/=/X segment plus S/H clips only
constexpr uint32_t kInfoHasExc = 1u << 17;      // read holds a non-ACGT base
constexpr uint32_t kInfoPositive = 1u << 18;    // isPositiveStrand
constexpr uint32_t kInfoEmpty = 1u << 19;       // consumes no reference (overlaps nothing)
constexpr uint32_t kInfoWideQ = 1u << 20;       // read holds a base quality > 63 (its qc bytes are not usable)
constexpr int kInfoMapqShift = 24;

// header of one granule's difference stream (built at pack time by k_expand, guac_tile.cuh)
struct __align__(16) GranHdr {
  uint32_t df_off;     // first entry of the granule in gs_diffs, in units of 8 entries (16 bytes)
  uint32_t n_df;       // entries
  uint32_t depth_in;   // reads that cover the granule's first locus and started before it
  uint32_t pos_in;     // ... of which on the positive strand
};

struct __align__(16) ReadRec {
  int32_t start;
  int32_t end;
  uint32_t pair_off;
  uint32_t info;
};

struct ContigInfo {
  uint64_t read_begin, read_end;  // global read index range
  uint32_t word_off;              // into trk_* (words)
  uint32_t gran_off;              // into gran_*
  int32_t length;                 // loci covered by the track
  int32_t n_words;
  int32_t n_grans;
  int32_t pad_;
};

struct DevReads {
  uint64_t n;
  int32_t max_ref_span;       // longest reference span of a read (how far back of a locus a start-sorted search must look)
  int32_t pad_;
  const ReadRec* rec;
  const uint32_t* cig_off;
  const uint32_t* cigar;
  const uint2* pairs;
  const uint32_t* xmask;
  const GranHdr* gs_hdr;      // per granule (all contigs, ContigInfo.gran_off): nullptr when the store was packed without streams
  const uint16_t* gs_diffs;   // CntLayout::slot_code(locus offset, class) entries, 16-byte aligned per granule
  const uint8_t* gs_dd;       // per locus of every granule: starts | ends << 4 (narrow) or u32 starts | ends << 16 (wide)
  const uint8_t* gs_dp;       // same, positive-strand reads only
  const uint32_t* gs_imp;     // per 32 loci of every granule: loci holding an element that is neither a plain base nor a
                              //   mid-deletion element carrying the track's base (only those need the exact per-locus walk)
  int32_t gs_wide;            // 1: 4 bytes per locus in gs_dd / gs_dp
  int32_t pad2_;
  // likelihood callers: per 32-locus word (all contigs, ContigInfo.word_off) the reads overlapping it as rows (guac_rows.cuh)
  const uint4* q_hdr;         // {first column pair, columns, first row group, rows | highest rank << 24}; nullptr: no such store
  const uint16_t* q_depth;    // per locus: plain elements (all mapping qualities)
  const uint32_t* q_cols;     // per (block of 8 columns, lane): eight 16-bit elements (class | quality << 4 | mapq rank << 10)
  const uint4* q_groups;      // general rows, per group of 4: their headers (mapq | type << 8)
  const uint32_t* q_rows;     // ... and per group and lane: the 4 rows' bytes for the lane's locus
  const uint64_t* seq_off;
  const uint8_t* seq;
  const uint8_t* qual;
  const uint8_t* qc;          // per base: quality (6 bits) | base code << 6; valid for reads without HAS_EXC / WIDE_Q
  const uint32_t* md_off;
  const char* md;
  const uint16_t* nm;
  const int32_t* del_start;   // per read: reference position of its first deleted base (-1: no deletion), its offset in the
  const uint32_t* del_md;     //   read's MD string, and the number of contiguous deleted bases: the exact per-locus path finds
  const uint16_t* del_len;    //   a deleted base without walking the tag (reads with several deletions walk for the others)
  const ContigInfo* contigs;
  const uint32_t* trk_lo;
  const uint32_t* trk_hi;
  const uint32_t* trk_std;
  const uint8_t* fasta;       // optional raw reference bytes (FASTA mode), else nullptr
  const uint64_t* fasta_off;
  const uint32_t* gran_first;
  const uint32_t* gran_last;
};

struct DevError {
  int code;
  int pad;
  unsigned long long where;
};

__device__ __forceinline__ void report_error(DevError* e, int code, unsigned long long where) {
  if (atomicCAS(&e->code, 0, code) == 0) e->where = where;
}

__device__ __forceinline__ bool op_consumes_read(uint32_t op) {
  return op == GUAC_CIGAR_M || op == GUAC_CIGAR_I || op == GUAC_CIGAR_S || op == GUAC_CIGAR_EQ || op == GUAC_CIGAR_X;
}
__device__ __forceinline__ bool op_consumes_ref(uint32_t op) {
  return op == GUAC_CIGAR_M || op == GUAC_CIGAR_D || op == GUAC_CIGAR_N || op == GUAC_CIGAR_EQ || op == GUAC_CIGAR_X;
}
__device__ __forceinline__ bool op_is_match_like(uint32_t op) {
  return op == GUAC_CIGAR_M || op == GUAC_CIGAR_EQ || op == GUAC_CIGAR_X;
}
__device__ __forceinline__ bool is_std_base(uint8_t b) { return b == 'A' || b == 'C' || b == 'G' || b == 'T'; }
// A=0 C=1 G=2 T=3 (lo = bit 0, hi = bit 1)
__device__ __forceinline__ uint32_t base_code(uint8_t b) { return b == 'C' ? 1u : b == 'G' ? 2u : b == 'T' ? 3u : 0u; }
__device__ __forceinline__ uint8_t code_base(uint32_t c) { return (uint8_t)("ACGT"[c & 3]); }

// mask of bits b in [lo_b, hi_b) of a 32-bit word, 0 <= lo_b, hi_b <= 32
__device__ __forceinline__ uint32_t bit_range(int lo_b, int hi_b) {
  lo_b = max(lo_b, 0);
  hi_b = min(hi_b, 32);
  if (hi_b <= lo_b) return 0u;
  uint32_t hi_m = hi_b >= 32 ? 0xFFFFFFFFu : ((1u << hi_b) - 1u);
  return hi_m & (0xFFFFFFFFu << lo_b);
}

// 32 bits of a read's bit-plane starting at (possibly negative, > -32) base index q0.  `get(j)` returns word j (j >= 0).
template <typename Get>
__device__ __forceinline__ uint32_t plane_window(Get get, int q0) {
  int j = q0 >> 5;  // arithmetic shift: floor
  int sh = q0 & 31;
  uint32_t a = j >= 0 ? get(j) : 0u;
  uint32_t b = get(j + 1);
  return __funnelshift_r(a, b, sh);
}

// Reads are sorted by start, and a granule's candidate range [first, last) can hold several hundred of them.  One
// warp-cooperative probe round (32 evenly spaced records) narrows it to the reads that can overlap loci [lo, hi): those with
// start < hi and start > lo - max_ref_span.  Conservative: it only drops whole probe intervals.  Warp-uniform result.
__device__ __forceinline__ void narrow_candidates(const DevReads& R, uint32_t& first, uint32_t& last, int lo, int hi) {
  const uint32_t n = last - first;
  if (first == 0xFFFFFFFFu || last <= first || n <= 96u) return;
  const int lane = threadIdx.x & 31;
  const uint32_t step = (n + 31u) >> 5;
  const uint32_t idx = first + (uint32_t)lane * step;
  const int s = idx < last ? R.rec[idx].start : 0x7FFFFFFF;
  const uint32_t before = __ballot_sync(0xFFFFFFFFu, (long long)s + R.max_ref_span <= (long long)lo);  // a prefix of the lanes
  const uint32_t after = __ballot_sync(0xFFFFFFFFu, s >= hi);                                         // a suffix of the lanes
  const int k = __popc(before);
  if (after) last = min(last, first + (uint32_t)(__ffs(after) - 1) * step);
  if (k > 0) first = first + (uint32_t)(k - 1) * step;
}

// ---- MD walk shared by the track builder, its verifier and the exact per-locus path ------------------------------
// Visits the alignment of one read in reference order, driven by the CIGAR with the MD tag consumed alongside
// (ADAM MdTag semantics, restated in oracle/guac_oracle.cpp Read::parse_md).  The visitor gets
//   run(ref_pos, read_pos, k)        k aligned bases where MD says "match"  (reference base == read base)
//   mismatch(ref_pos, read_pos, ch)  MD mismatch: reference base ch
//   deleted(ref_pos, ch, md_pos)     reference base ch inside a D op (md_pos = its offset in the read's MD string)
//   skipped(ref_pos, len)            N op
// and returns false from any callback to stop early.  Returns 0 or a guac_status.
template <typename V>
__device__ int md_walk(const DevReads& R, uint64_t r, V& v) {
  const ReadRec rec = R.rec[r];
  const uint32_t c0 = R.cig_off[r], c1 = R.cig_off[r + 1];
  const char* md = R.md + R.md_off[r];
  const int md_len = (int)(R.md_off[r + 1] - R.md_off[r]);
  int pos = 0;
  int64_t pending = 0;
  int ref_pos = rec.start, read_pos = 0;
  bool trivial = (md_len == 1 && md[0] == '0') || md_len == 0;  // MdTag: null or "0" -> empty tag
  if (!trivial && !(md[0] >= '0' && md[0] <= '9')) return GUAC_ERR_MISSING_MD;
  for (uint32_t c = c0; c < c1; ++c) {
    const uint32_t op = R.cigar[c] & 0xF;
    int len = (int)(R.cigar[c] >> 4);
    if (op_is_match_like(op)) {
      int remaining = len;
      if (trivial) {  // no mismatches known: everything matches
        if (!v.run(ref_pos, read_pos, remaining)) return 0;
        ref_pos += remaining;
        read_pos += remaining;
        continue;
      }
      while (remaining > 0) {
        if (pending > 0) {
          int k = (int)(pending < remaining ? pending : remaining);
          if (!v.run(ref_pos, read_pos, k)) return 0;
          ref_pos += k;
          read_pos += k;
          remaining -= k;
          pending -= k;
        } else if (pos >= md_len) {
          return GUAC_ERR_MISSING_MD;
        } else {
          char ch = md[pos];
          if (ch >= '0' && ch <= '9') {
            int64_t n = 0;
            while (pos < md_len && md[pos] >= '0' && md[pos] <= '9') n = n * 10 + (md[pos++] - '0');
            pending = n;
          } else if (ch == '^') {
            return GUAC_ERR_MISSING_MD;
          } else {
            ++pos;
            if (!v.mismatch(ref_pos, read_pos, (uint8_t)ch)) return 0;
            ++ref_pos;
            ++read_pos;
            --remaining;
          }
        }
      }
    } else if (op == GUAC_CIGAR_D) {
      int remaining = len;
      if (trivial) return GUAC_ERR_MISSING_MD;
      while (remaining > 0) {
        if (pending > 0) return GUAC_ERR_MISSING_MD;  // "found matching bases in deletion"
        if (pos >= md_len) return GUAC_ERR_MISSING_MD;
        char ch = md[pos];
        if (ch >= '0' && ch <= '9') {
          int64_t n = 0;
          while (pos < md_len && md[pos] >= '0' && md[pos] <= '9') n = n * 10 + (md[pos++] - '0');
          pending = n;
        } else if (ch == '^') {
          ++pos;
        } else {
          ++pos;
          if (!v.deleted(ref_pos, (uint8_t)ch, pos - 1)) return 0;
          ++ref_pos;
          --remaining;
        }
      }
    } else if (op == GUAC_CIGAR_N) {
      if (!v.skipped(ref_pos, len)) return 0;
      ref_pos += len;
    } else if (op == GUAC_CIGAR_I || op == GUAC_CIGAR_S) {
      read_pos += len;
    }
  }
  return 0;
}

}  // namespace guac
