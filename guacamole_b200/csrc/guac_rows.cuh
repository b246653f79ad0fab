// guac_rows.cuh — the likelihood callers' view of a read set: per 32-locus word, the pileup TRANSPOSED.
//
// Likelihood.likelihoodsOfAllPossibleGenotypesFromPileup (likelihood/Likelihood.scala:99-113, 149-201) needs, per locus, every
// overlapping element's base and quality and its read's mapping quality.  Walking the reads from the loci (k_somatic's
// gather_sample) costs ~60 instructions and two dependent loads per (read, word).  K_expand_rows does that walk ONCE, at pack
// time — the CIGAR expansion of PileupElement.advanceToLocus / alignment (pileup/PileupElement.scala:68-248) — and leaves,
// per word:
//   columns   the plain elements (A/C/G/T bases of reads that are one M/=/X run between clips, qualities < 64: 98 % of all
//             elements) of every locus, in read order, 16 bits each: (base code ^ reference code) | quality << 4 | rank of
//             the read's mapping quality among those present (from the top) << 10 — bits 15..4 are the element's byte offset
//             in the likelihood table (64 rows of 64 double2).  Column k of the word holds the k-th
//             element of each of its 32 loci; loci with fewer elements are filled with a sentinel that reads a zero row of
//             the likelihood table.  Two columns per 32-bit word per locus: the likelihood kernel streams 128 coalesced bytes
//             per two columns and does nothing per element but one table look-up and two additions.  Elements that carry
//             the reference base fill a locus' column from the FRONT, the mismatching ones (a percent) from the BACK, the
//             sentinels sit in between: only the word's last block(s) of columns — how many is in the header — can hold an
//             element whose class is not 0, every other block is summed without looking at the class bits;
//   depth     the number of plain elements per locus (u16);
//   rows      the other reads (insertions / deletions / skips / non-ACGT bases / wide qualities), two rows of 32 bytes each:
//             the element class per locus (0xF8 | base code = plain base, 0xFE = an element that is not a plain base, 0xFD =
//             an element the exact kernel must report an error for, 0xFF = no element) and its qualityScore; four rows to a
//             group with one uint4 of row headers (mapq | type << 8).
#pragma once

#include "guac_pileup.cuh"

namespace guac {

constexpr uint32_t kRowLean = 0u, kRowGeneral = 1u, kRowQuality = 2u;
constexpr uint32_t kElemNone = 0xFFu, kElemOther = 0xFEu, kElemHard = 0xFDu, kElemPlain = 0xF8u;

constexpr uint32_t kRankZero = 63u;        // rank of the sentinel element: the all-zero row of the likelihood table
constexpr uint32_t kMaxRank = 62u;         // reads whose mapping quality ranks at or beyond this take the general rows
constexpr uint32_t kSentinel = kRankZero << 10;
constexpr uint32_t kSentinel2 = kSentinel | (kSentinel << 16);
constexpr uint32_t kRareAll = 255u;        // header value: "any block may hold a mismatching element"

// halfword j (0..7) of a block of eight column elements held in four registers (no dynamic register index)
__device__ __forceinline__ void set_half(uint32_t (&p)[4], uint32_t j, uint32_t elem) {
  const uint32_t sh = 16u * (j & 1u), keep = ~(0xFFFFu << sh), v = elem << sh;
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = (j >> 1) == (uint32_t)i ? ((p[i] & keep) | v) : p[i];
}

struct RowsArgs {
  DevReads R;
  uint4* hdr_w;                // per word: {first column block, columns | trailing blocks with mismatches << 24, first row group,
                               //            rows | highest rank << 24}
  uint16_t* depth_w;           // per locus: plain elements
  uint32_t* cols_w;            // per (block of eight columns, lane): 16 bytes
  uint4* groups_w;
  uint32_t* rows_w;
  unsigned long long cap_pairs, cap_groups;
  uint32_t w_begin, w_end;     // global word indices this launch covers
  uint32_t n_contigs;
  uint32_t pad_;
  unsigned long long* counters;  // [4] column blocks reserved, [5] row groups reserved
  uint32_t mapq_mask[8];         // mapping qualities present in the read set (k_header)
  DevError* err;
};

constexpr int kRowsWarps = 8;

__device__ __forceinline__ uint32_t mapq_rank(const uint32_t (&mask)[8], uint32_t mapq) {  // mapping qualities present above this one
  uint32_t rank = 0;
#pragma unroll
  for (int wd = 0; wd < 8; ++wd) {
    const uint32_t m = mask[wd];
    rank += (uint32_t)wd > (mapq >> 5) ? __popc(m) : ((uint32_t)wd == (mapq >> 5) ? __popc(m & ~((2u << (mapq & 31u)) - 1u)) : 0u);
  }
  return rank;
}

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
    if (lane == 0) A.hdr_w[W] = make_uint4(0u, 0u, 0u, 0u);
    A.depth_w[(size_t)W * 32 + lane] = 0;
    return;
  }
  narrow_candidates(R, first, last, span_lo, span_lo + 32);
  constexpr uint32_t kLeanMask = kInfoSimple | kInfoHasExc | kInfoWideQ;
  const uint32_t rcode = ((R.trk_lo[W] >> lane) & 1u) | (((R.trk_hi[W] >> lane) & 1u) << 1);  // this locus' reference code
  const bool std_ref = (R.trk_std[W] >> lane) & 1u;  // (no reference class without a standard reference base: all elements in front)

  uint32_t depth = 0, n_cols = 0, max_rank = 0, n_rows = 0, slots = 0;
  unsigned long long coff = 0, goff = 0;
  uint32_t acc = 0, hacc[4] = {0u, 0u, 0u, 0u};
  for (int pass = 0; pass < 2; ++pass) {
    uint32_t row = 0, kf = 0, kb = 0;
    uint32_t pf[4] = {kSentinel2, kSentinel2, kSentinel2, kSentinel2}, pb[4] = {kSentinel2, kSentinel2, kSentinel2, kSentinel2};
    uint4* const blocks_w = reinterpret_cast<uint4*>(A.cols_w) + (size_t)coff * 32 + lane;
    // this lane's next column element: reference-class elements from the front, the others from the back; a block of eight
    // columns (16 bytes per locus) leaves when it is full
    auto put = [&](uint32_t elem) {
      if ((elem & 3u) == 0u || !std_ref) {
        set_half(pf, kf & 7u, elem);
        if ((kf & 7u) == 7u) {
          blocks_w[(size_t)(kf >> 3) * 32] = make_uint4(pf[0], pf[1], pf[2], pf[3]);
          pf[0] = pf[1] = pf[2] = pf[3] = kSentinel2;
        }
        ++kf;
      } else {
        const uint32_t pos = slots - 1u - kb;
        set_half(pb, pos & 7u, elem);
        if ((pos & 7u) == 0u) {
          blocks_w[(size_t)(pos >> 3) * 32] = make_uint4(pb[0], pb[1], pb[2], pb[3]);
          pb[0] = pb[1] = pb[2] = pb[3] = kSentinel2;
        }
        ++kb;
      }
    };
    auto append = [&](uint32_t header, uint32_t byte) {  // warp-uniform call; `byte` per lane
      const uint32_t j = row & 3u;
      acc |= (byte & 0xFFu) << (8 * j);
      hacc[0] = j == 0 ? header : hacc[0];
      hacc[1] = j == 1 ? header : hacc[1];
      hacc[2] = j == 2 ? header : hacc[2];
      hacc[3] = j == 3 ? header : hacc[3];
      if (j == 3u) {
        const size_t grp = (size_t)goff + (row >> 2);
        A.rows_w[grp * 32 + lane] = acc;
        if (lane == 0) A.groups_w[grp] = make_uint4(hacc[0], hacc[1], hacc[2], hacc[3]);
        acc = 0;
        hacc[0] = hacc[1] = hacc[2] = hacc[3] = 0u;
      }
      ++row;
    };
    for (uint32_t base = first; base < last; base += 32) {
      const uint32_t mine = base + lane;
      ReadRec my{0, 0, 0, 0};
      if (mine < last) my = R.rec[mine];
      const bool overlaps = mine < last && my.start < span_lo + 32 && my.end > span_lo && my.end > my.start;
      const uint32_t my_rank = overlaps ? mapq_rank(A.mapq_mask, my.info >> kInfoMapqShift) : 0u;
      const bool lean_mine = overlaps && (my.info & kLeanMask) == kInfoSimple && my_rank < kMaxRank;
      uint32_t ov = __ballot_sync(0xFFFFFFFFu, overlaps);
      const uint32_t ov_lean = __ballot_sync(0xFFFFFFFFu, lean_mine);
      unsigned long long my_qa = 0;
      if (pass == 1 && lean_mine)
        my_qa = (unsigned long long)(uintptr_t)R.qc + R.seq_off[mine] + (unsigned long long)(my.info & kInfoLeadMask) - (unsigned long long)(long long)my.start;
      while (ov) {  // warp-uniform, reads in index order
        const int j = __ffs(ov) - 1;
        ov &= ov - 1;
        const bool lean = (ov_lean >> j) & 1u;
        const int start = __shfl_sync(0xFFFFFFFFu, my.start, j), end = __shfl_sync(0xFFFFFFFFu, my.end, j);
        const uint32_t rank = __shfl_sync(0xFFFFFFFFu, my_rank, j);
        const bool covered = x >= start && x < end;
        if (pass == 0) {
          if (lean) {
            depth += covered ? 1u : 0u;
            max_rank = max(max_rank, rank);
          } else {
            row += 2u;
          }
          continue;
        }
        if (lean) {
          const unsigned long long qa = ((unsigned long long)__shfl_sync(0xFFFFFFFFu, (uint32_t)(my_qa >> 32), j) << 32) |
                                        __shfl_sync(0xFFFFFFFFu, (uint32_t)my_qa, j);
          if (covered) {
            const uint32_t b = (uint32_t)__ldg(reinterpret_cast<const uint8_t*>((uintptr_t)(qa + (unsigned long long)(long long)x)));
            put(((b & 63u) << 4) | (((b >> 6) ^ rcode) & 3u) | (rank << 10));
          }
        } else {
          const uint32_t info = __shfl_sync(0xFFFFFFFFu, my.info, j);
          const uint32_t mapq = info >> kInfoMapqShift;
          uint32_t b = kElemNone, q = 0;
          if (covered) {
            Elem e;
            const int rc = classify(R, (uint64_t)(base + j), x, (uint8_t)'N', e);
            if (rc || e.kind == kNone) b = kElemHard;
            else if ((e.kind == kMatch || e.kind == kMismatch) && is_std_base(e.base)) b = kElemPlain | base_code(e.base);
            else b = kElemOther;
            q = (uint32_t)e.qual & 0xFFu;
          }
          append(mapq | (kRowGeneral << 8), b);
          append(mapq | (kRowQuality << 8), q);
        }
      }
    }
    if (pass == 0) {
      n_rows = row;
      n_cols = __reduce_max_sync(0xFFFFFFFFu, depth);
      max_rank = __reduce_max_sync(0xFFFFFFFFu, max_rank);
      const uint32_t pairs = (n_cols + 7u) >> 3, groups = (n_rows + 3u) >> 2;  // (blocks of eight columns)
      slots = pairs << 3;
      if (n_cols > 0xFFFFFFu) { report_error(A.err, GUAC_ERR_UNSUPPORTED, (unsigned long long)W); return; }
      if (lane == 0) {
        coff = atomicAdd(&A.counters[4], (unsigned long long)pairs);
        goff = atomicAdd(&A.counters[5], (unsigned long long)groups);
      }
      coff = __shfl_sync(0xFFFFFFFFu, coff, 0);
      goff = __shfl_sync(0xFFFFFFFFu, goff, 0);
      const bool fits = coff + pairs <= A.cap_pairs && goff + groups <= A.cap_groups;  // (else the host grows the buffers and repeats)
      if (lane == 0) A.hdr_w[W] = make_uint4((uint32_t)coff, n_cols | (kRareAll << 24), (uint32_t)goff, n_rows | (max_rank << 24));
      A.depth_w[(size_t)W * 32 + lane] = (uint16_t)min(depth, 0xFFFFu);
      if (!fits || (n_cols == 0 && n_rows == 0)) return;
    } else {
      // the free positions [kf, slots - kb) of this lane's column hold sentinels: the two partial blocks (one, where front and
      // back meet inside a block) and the whole blocks in between
      const uint32_t free_lo = kf, free_hi = slots - kb;
      const bool part_f = (free_lo & 7u) != 0u, part_b = (free_hi & 7u) != 0u;
      if (part_f && part_b && (free_lo >> 3) == (free_hi >> 3)) {
        const uint32_t nf = free_lo & 7u;  // halfwords below nf come from the front block, the others from the back block
        uint32_t m[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t lo = (uint32_t)(2 * i) < nf ? pf[i] : pb[i], hi = (uint32_t)(2 * i + 1) < nf ? pf[i] : pb[i];
          m[i] = (lo & 0xFFFFu) | (hi & 0xFFFF0000u);
        }
        blocks_w[(size_t)(free_lo >> 3) * 32] = make_uint4(m[0], m[1], m[2], m[3]);
      } else {
        if (part_f) blocks_w[(size_t)(free_lo >> 3) * 32] = make_uint4(pf[0], pf[1], pf[2], pf[3]);
        if (part_b) blocks_w[(size_t)(free_hi >> 3) * 32] = make_uint4(pb[0], pb[1], pb[2], pb[3]);
      }
      for (uint32_t b = (free_lo + 7u) >> 3; b < (free_hi >> 3); ++b) blocks_w[(size_t)b * 32] = make_uint4(kSentinel2, kSentinel2, kSentinel2, kSentinel2);
      // trailing blocks that hold an element of a mismatch class on some lane
      const uint32_t max_back = __reduce_max_sync(0xFFFFFFFFu, kb);
      const uint32_t rare_blocks = max_back ? (slots >> 3) - ((slots - max_back) >> 3) : 0u;
      if (lane == 0) A.hdr_w[W] = make_uint4((uint32_t)coff, n_cols | (min(rare_blocks, kRareAll) << 24), (uint32_t)goff, n_rows | (max_rank << 24));
      if (row & 3u) {  // the last, partial group of rows
        const size_t grp = (size_t)goff + (row >> 2);
        A.rows_w[grp * 32 + lane] = acc;
        if (lane == 0) A.groups_w[grp] = make_uint4(hacc[0], hacc[1], hacc[2], hacc[3]);
      }
    }
  }
}

}  // namespace guac
