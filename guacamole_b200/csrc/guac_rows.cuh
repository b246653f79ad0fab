// guac_rows.cuh — the likelihood callers' view of a read set: per 32-locus word, the overlapping reads as ROWS.
//
// Likelihood.likelihoodsOfAllPossibleGenotypesFromPileup (likelihood/Likelihood.scala:99-113, 149-201) needs, per locus, every
// overlapping element's base and quality (and its read's mapping quality): 1 byte per (read, locus).  Walking the reads from
// the loci (k_somatic's gather_sample) costs ~60 instructions and two dependent loads per (read, word).  K_expand_rows does
// that walk ONCE, at pack time — the CIGAR expansion of PileupElement.advanceToLocus / alignment
// (pileup/PileupElement.scala:68-248) — and leaves, per word, one row per overlapping read in read order:
//   lean rows      (one M/=/X run between clips, A/C/G/T bases, qualities < 64 — 98 % of the reads): 32 bytes, byte l =
//                  quality | base code << 6 of the read's base at locus l of the word;
//   general rows   (reads with insertions / deletions / skips / non-ACGT bases / wide qualities), two rows each: the element
//                  class per locus (0xF8 | base code = plain base, 0xFE = an element that is not a plain base, 0xFD = an
//                  element the exact kernel must report an error for, 0xFF = no element) and its qualityScore.
// Rows are stored four to a group: per group one uint4 of row headers (mapq | type << 8 | first lane << 10 | lanes << 15, the
// same for every locus: a row's table row is warp-uniform) and per lane one 32-bit word holding the lane's four bytes, so the
// likelihood kernel streams 128 coalesced bytes per four reads and does no per-read address arithmetic at all.
#pragma once

#include "guac_pileup.cuh"

namespace guac {

constexpr uint32_t kRowLean = 0u, kRowGeneral = 1u, kRowQuality = 2u;
constexpr uint32_t kElemNone = 0xFFu, kElemOther = 0xFEu, kElemHard = 0xFDu, kElemPlain = 0xF8u;

// row header: mapq 7..0 | type 9..8 | first lane 14..10 | lanes 20..15 | rank of the mapq among those present, from the top, 28..21 (the row
// of the likelihood kernel's shared-memory table); 0 = padding (a lean row of no lanes)
__device__ __forceinline__ uint32_t row_header(uint32_t mapq, uint32_t type, int lo, int len, uint32_t rank) {
  return mapq | (type << 8) | ((uint32_t)lo << 10) | ((uint32_t)len << 15) | (rank << 21);
}

struct RowsArgs {
  DevReads R;
  uint2* hdr_w;
  uint4* groups_w;
  uint32_t* rows_w;
  unsigned long long cap_groups;
  uint32_t w_begin, w_end;     // global word indices this launch covers
  uint32_t n_contigs;
  uint32_t pad_;
  unsigned long long* counters;  // [4] groups reserved
  uint32_t mapq_mask[8];         // mapping qualities present in the read set (k_header): a row's table row is its mapq's rank
};

constexpr int kRowsWarps = 8;

__global__ void __launch_bounds__(kRowsWarps * 32) k_expand_rows(RowsArgs A) {
  const DevReads& R = A.R;
  const int lane = threadIdx.x & 31;
  const uint32_t W = A.w_begin + blockIdx.x * kRowsWarps + (threadIdx.x >> 5);
  if (W >= A.w_end) return;  // whole warp
  uint32_t c = 0;
  {
    uint32_t lo = 0, hi = A.n_contigs - 1;
    while (lo < hi) {
      const uint32_t mid = (lo + hi + 1) >> 1;
      if (R.contigs[mid].word_off <= W) lo = mid; else hi = mid - 1;
    }
    c = lo;
  }
  const ContigInfo ci = R.contigs[c];
  const int span_lo = (int)(W - ci.word_off) << 5, x = span_lo + lane;
  uint32_t first = 0xFFFFFFFFu, last = 0;
  if (span_lo < ci.length) {
    const int g = span_lo >> kGranuleShift;
    first = R.gran_first[ci.gran_off + g];
    last = R.gran_last[ci.gran_off + g];
  }
  if (first == 0xFFFFFFFFu) {
    if (lane == 0) A.hdr_w[W] = make_uint2(0u, 0u);
    return;
  }
  narrow_candidates(R, first, last, span_lo, span_lo + 32);
  constexpr uint32_t kLeanMask = kInfoSimple | kInfoHasExc | kInfoWideQ;

  uint32_t n_rows = 0;
  unsigned long long goff = 0;
  bool fits = false;
  uint32_t acc = 0, hacc[4] = {0u, 0u, 0u, 0u};
  for (int pass = 0; pass < 2; ++pass) {
    uint32_t row = 0;
    auto append = [&](uint32_t header, uint32_t byte) {  // warp-uniform call; `byte` per lane
      if (pass == 1 && fits) {
        const uint32_t k = row & 3u;
        acc |= (byte & 0xFFu) << (8 * k);
        hacc[0] = k == 0 ? header : hacc[0];
        hacc[1] = k == 1 ? header : hacc[1];
        hacc[2] = k == 2 ? header : hacc[2];
        hacc[3] = k == 3 ? header : hacc[3];
        if (k == 3u) {
          const size_t grp = (size_t)goff + (row >> 2);
          A.rows_w[grp * 32 + lane] = acc;
          if (lane == 0) A.groups_w[grp] = make_uint4(hacc[0], hacc[1], hacc[2], hacc[3]);
          acc = 0;
          hacc[0] = hacc[1] = hacc[2] = hacc[3] = 0u;
        }
      }
      ++row;
    };
    for (uint32_t base = first; base < last; base += 32) {
      const uint32_t mine = base + lane;
      ReadRec my{0, 0, 0, 0};
      if (mine < last) my = R.rec[mine];
      const bool overlaps = mine < last && my.start < span_lo + 32 && my.end > span_lo && my.end > my.start;
      const bool lean_mine = overlaps && (my.info & kLeanMask) == kInfoSimple;
      uint32_t ov = __ballot_sync(0xFFFFFFFFu, overlaps);
      const uint32_t ov_lean = __ballot_sync(0xFFFFFFFFu, lean_mine);
      unsigned long long my_qa = 0;
      if (pass == 1 && lean_mine)
        my_qa = (unsigned long long)(uintptr_t)R.qc + R.seq_off[mine] + (unsigned long long)(my.info & kInfoLeadMask) - (unsigned long long)(long long)my.start;
      while (ov) {  // warp-uniform, reads in index order
        const int j = __ffs(ov) - 1;
        ov &= ov - 1;
        const bool lean = (ov_lean >> j) & 1u;
        if (pass == 0) {
          if (!lean && (row & 1u)) ++row;  // a general read's two rows share a group: pad to an even row
          row += lean ? 1u : 2u;
          continue;
        }
        const int start = __shfl_sync(0xFFFFFFFFu, my.start, j), end = __shfl_sync(0xFFFFFFFFu, my.end, j);
        const uint32_t info = __shfl_sync(0xFFFFFFFFu, my.info, j);
        const uint32_t mapq = info >> kInfoMapqShift;
        uint32_t rank = 0;  // mapping qualities present ABOVE this one (high ones are the common ones: they get the first rows)
#pragma unroll
        for (int wd = 0; wd < 8; ++wd) {
          const uint32_t m = A.mapq_mask[wd];
          rank += (uint32_t)wd > (mapq >> 5) ? __popc(m) : ((uint32_t)wd == (mapq >> 5) ? __popc(m & ~((2u << (mapq & 31u)) - 1u)) : 0u);
        }
        if (lean) {
          const unsigned long long qa = ((unsigned long long)__shfl_sync(0xFFFFFFFFu, (uint32_t)(my_qa >> 32), j) << 32) |
                                        __shfl_sync(0xFFFFFFFFu, (uint32_t)my_qa, j);
          const int lo = max(start, span_lo) - span_lo, hi = min(end, span_lo + 32) - span_lo;
          uint32_t b = 0;
          if (lane >= lo && lane < hi) b = (uint32_t)__ldg(reinterpret_cast<const uint8_t*>((uintptr_t)(qa + (unsigned long long)(long long)x)));
          append(row_header(mapq, kRowLean, lo, hi - lo, rank), b);
        } else {
          if (row & 1u) append(0u, 0u);
          uint32_t b = kElemNone, q = 0;
          if (x >= start && x < end) {
            Elem e;
            const int rc = classify(R, (uint64_t)(base + j), x, (uint8_t)'N', e);
            if (rc || e.kind == kNone) b = kElemHard;
            else if ((e.kind == kMatch || e.kind == kMismatch) && is_std_base(e.base)) b = kElemPlain | base_code(e.base);
            else b = kElemOther;
            q = (uint32_t)e.qual & 0xFFu;
          }
          append(row_header(mapq, kRowGeneral, 0, 32, rank), b);
          append(row_header(mapq, kRowQuality, 0, 32, rank), q);
        }
      }
    }
    if (pass == 0) {
      n_rows = row;
      const uint32_t groups = (n_rows + 3u) >> 2;
      if (lane == 0) goff = atomicAdd(&A.counters[4], (unsigned long long)groups);
      goff = __shfl_sync(0xFFFFFFFFu, goff, 0);
      fits = goff + groups <= A.cap_groups;  // (else the host grows the buffers and packs the rows again)
      if (lane == 0) A.hdr_w[W] = make_uint2((uint32_t)goff, n_rows);
      if (!fits || n_rows == 0) return;
    } else if (row & 3u) {  // the last, partial group
      const size_t grp = (size_t)goff + (row >> 2);
      A.rows_w[grp * 32 + lane] = acc;
      if (lane == 0) A.groups_w[grp] = make_uint4(hacc[0], hacc[1], hacc[2], hacc[3]);
    }
  }
}

}  // namespace guac
