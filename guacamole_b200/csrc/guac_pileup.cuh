// guac_pileup.cuh — the pileup hot path: CIGAR expansion + per-locus allele / depth / strand counting + fused callers.
//
// K_tile   (k_pileup_tile): one CTA per tile of 4096 loci, one THREAD per 32-loci word.  The tile's reads (a contiguous
//           index range of the start-sorted store, found through the granule index) are staged chunk-wise into shared
//           memory; every thread aligns each overlapping read's 2-bit base planes onto its word with funnel shifts
//           (CIGAR expansion, bit-parallel over 32 loci) and adds them into bit-sliced ("vertical") counters with
//           carry-save adders — no atomics, no per-(read, locus) work.  The epilogue evaluates the caller per locus:
//             GermlineThreshold.Caller.callVariantsAtLocus      commands/GermlineThresholdCaller.scala:90-179
//             Pileup.depth/positiveDepth/referenceDepth           pileup/Pileup.scala:76-91
//           Replaces SlidingWindow.setCurrentLocus (windowing/SlidingWindow.scala:83-110), Pileup.atGreaterLocus
//           (pileup/Pileup.scala:103-132) and PileupElement.advanceToLocus/alignment (pileup/PileupElement.scala:68-248).
// K_exact  (k_exact_loci): one thread per locus that the bit-sliced path cannot decide exactly (insertions, deletions,
//           clipped/N-skipped elements, non-ACGT bases, non-standard reference base): a literal per-element walk.
#pragma once

#include "guac_device.cuh"

namespace guac {

// ---- device output ------------------------------------------------------------------------------------------------
constexpr uint32_t kPoolAltOff = 0;    // "<ALT>" lives at pool[0..5)
constexpr uint32_t kPoolByteOff = 8;   // byte value v lives at pool[8 + v]
constexpr uint32_t kPoolDynOff = 264;  // dynamically allocated allele strings start here

struct SlowLocus {
  int32_t contig;
  int32_t locus;
};

struct DevOut {
  guac_threshold_record* trec;
  guac_locus_counts* crec;
  uint32_t cap_rec;
  uint8_t* pool;
  uint32_t cap_pool;
  SlowLocus* slow;
  uint32_t cap_slow;
  // counters: [0] records, [1] pool bytes, [2] slow loci, [3] visited loci, [4] tie loci, [5] counter overflow
  unsigned long long* counters;
  DevError* err;
};

struct TileDesc {
  int32_t contig;
  int32_t word0;       // first word of the tile (contig-relative)
  int32_t locus_begin; // requested loci of this tile: [locus_begin, locus_end)
  int32_t locus_end;
};

struct CallParams {
  int32_t mode;             // 0 germline threshold, 1 per-locus counts
  int32_t threshold_percent;
  int32_t emit_ref;
  int32_t emit_no_call;
  int32_t skip_empty;
  int32_t sample;
};

// ---- bit-sliced counters --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a ^ b)); }

template <int W>
struct VCounter {
  uint32_t p[W];  // p[k] holds bit k of the 32 per-locus counts
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < W; ++k) p[k] = 0;
  }
  // add four one-bit-per-locus words; returns the carry out of the top plane
  __device__ __forceinline__ uint32_t add4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t t0 = maj3(p[0], a, b);
    uint32_t s0 = p[0] ^ a ^ b;
    uint32_t t1 = maj3(s0, c, d);
    p[0] = s0 ^ c ^ d;
    uint32_t f = maj3(p[1], t0, t1);
    p[1] = p[1] ^ t0 ^ t1;
#pragma unroll
    for (int k = 2; k < W; ++k) {
      uint32_t cy = p[k] & f;
      p[k] ^= f;
      f = cy;
    }
    return f;
  }
  __device__ __forceinline__ uint32_t any() const {
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < W; ++k) o |= p[k];
    return o;
  }
  __device__ __forceinline__ int at(int b) const {
    int v = 0;
#pragma unroll
    for (int k = 0; k < W; ++k) v |= (int)((p[k] >> b) & 1u) << k;
    return v;
  }
};

// ---- one read aligned onto one word ------------------------------------------------------------------------------------
struct Aligned {
  uint32_t plain;  // Match/Mismatch elements whose read base is A/C/G/T
  uint32_t lo, hi; // their 2-bit codes
  uint32_t other;  // every other element of this read in the word (insertion / deletion anchors, mid-deletion, N-skip, non-ACGT)
};

// SIMPLE reads: a single aligned segment [start, end) <-> read bases [lead, lead + end - start)
__device__ __forceinline__ Aligned align_simple(const ReadRec& rec, const uint2* P, int wbase) {
  Aligned a;
  int q0 = (int)(rec.info & kInfoLeadMask) + (wbase - rec.start);
  uint32_t valid = bit_range(rec.start - wbase, rec.end - wbase);
  a.lo = plane_window([&](int j) { return P[j].x; }, q0) & valid;
  a.hi = plane_window([&](int j) { return P[j].y; }, q0) & valid;
  a.plain = valid;
  a.other = 0;
  return a;
}

// reads with a non-ACGT base: take the exception mask from HBM (rare)
__device__ __noinline__ void apply_exceptions(Aligned& a, const DevReads& R, uint32_t pair_off, int q0_of_bit0, uint32_t valid) {
  const uint32_t* X = R.xmask + pair_off;
  uint32_t x = plane_window([&](int j) { return X[j]; }, q0_of_bit0) & valid & a.plain;
  a.plain &= ~x;
  a.lo &= ~x;
  a.hi &= ~x;
  a.other |= x;
}

// general reads: walk the run-length CIGAR; each M/=/X run is one funnel-shifted window
__device__ __noinline__ Aligned align_cigar(const DevReads& R, uint64_t r, const ReadRec& rec, const uint2* P, int wbase) {
  Aligned a{0, 0, 0, 0};
  const uint32_t c0 = R.cig_off[r], c1 = R.cig_off[r + 1];
  int ref_pos = rec.start, read_pos = 0;
  uint32_t prev_op = 0xFFu;
  for (uint32_t c = c0; c < c1 && ref_pos <= wbase + 32; ++c) {
    const uint32_t op = R.cigar[c] & 0xF;
    const int len = (int)(R.cigar[c] >> 4);
    if (op_is_match_like(op)) {
      if (ref_pos + len > wbase) {
        uint32_t valid = bit_range(ref_pos - wbase, ref_pos + len - wbase);
        int q0 = read_pos + (wbase - ref_pos);
        a.lo |= plane_window([&](int j) { return P[j].x; }, q0) & valid;
        a.hi |= plane_window([&](int j) { return P[j].y; }, q0) & valid;
        a.plain |= valid;
        if (rec.info & kInfoHasExc) apply_exceptions(a, R, rec.pair_off, q0, valid);
      }
      ref_pos += len;
      read_pos += len;
    } else if (op == GUAC_CIGAR_I) {
      // (M|=, I): the last base of the preceding run carries the insertion; an I at reference position 0 of the
      // contig is the contig-start insertion.  Either way that locus is not a plain element.
      int anchor = (ref_pos == 0) ? 0 : ref_pos - 1;
      if ((prev_op == GUAC_CIGAR_M || prev_op == GUAC_CIGAR_EQ || ref_pos == 0) && anchor >= wbase && anchor < wbase + 32)
        a.other |= 1u << (anchor - wbase);
      read_pos += len;
    } else if (op == GUAC_CIGAR_D || op == GUAC_CIGAR_N) {
      if (op == GUAC_CIGAR_D && ref_pos - 1 >= wbase && ref_pos - 1 < wbase + 32 && ref_pos > rec.start)
        a.other |= 1u << (ref_pos - 1 - wbase);  // deletion anchor (or an invalid predecessor: decided exactly later)
      if (ref_pos + len > wbase) a.other |= bit_range(ref_pos - wbase, ref_pos + len - wbase);
      ref_pos += len;
    } else if (op == GUAC_CIGAR_S) {
      read_pos += len;
    }
    prev_op = op;
  }
  // (the loop runs while ref_pos <= wbase + 32 so that an I / D starting right after the word still marks its anchor)
  a.plain &= ~a.other;
  a.lo &= a.plain;
  a.hi &= a.plain;
  return a;
}

// ---- K_tile ---------------------------------------------------------------------------------------------------------------
struct __align__(16) TileSmem {
  ReadRec rec[kChunkReads + 4];
  uint2 pairs[kChunkPairs + 8];
  int32_t pmax[kChunkReads];
  int32_t warp_max[8];
  uint32_t first, last;
  uint32_t cut;
};

template <int W, int MODE>
__global__ void __launch_bounds__(kTileWords) k_pileup_tile(DevReads R, const TileDesc* __restrict__ tiles, CallParams prm, DevOut out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem& S = *reinterpret_cast<TileSmem*>(smem_raw);
  const TileDesc td = tiles[blockIdx.x];
  const ContigInfo ci = R.contigs[td.contig];
  const int tid = threadIdx.x;
  const int w = td.word0 + tid;          // this thread's word (contig-relative)
  const int wbase = w << 5;

  // candidate reads of the tile through the granule index
  if (tid == 0) {
    S.first = 0xFFFFFFFFu;
    S.last = 0;
  }
  __syncthreads();
  {
    int g0 = (td.word0 << 5) >> kGranuleShift;
    int g1 = min(((td.word0 + kTileWords) << 5) - 1, ci.length - 1) >> kGranuleShift;
    int g = g0 + tid;
    if (g <= g1 && g < ci.n_grans) {
      uint32_t f = R.gran_first[ci.gran_off + g], l = R.gran_last[ci.gran_off + g];
      if (f != 0xFFFFFFFFu) {
        atomicMin(&S.first, f);
        atomicMax(&S.last, l);
      }
    }
  }
  __syncthreads();
  const uint32_t first = S.first, last = S.last;

  VCounter<W> cV, cL, cH, cHL;   // plain elements, lo bit, hi bit, both
  VCounter<(MODE == 1 ? W : 3)> cO;  // other elements (narrow + sticky overflow in caller mode)
  VCounter<(MODE == 1 ? W : 1)> cP;  // positive-strand elements (counts mode only)
  cV.clear(); cL.clear(); cH.clear(); cHL.clear(); cO.clear(); cP.clear();
  uint32_t ovf = 0, o_sat = 0;

  for (uint32_t c0 = first; c0 < last && first != 0xFFFFFFFFu;) {
    // ---- stage a chunk of records, cut it so that its plane pairs fit, then stage the pairs
    const uint32_t cn_max = min((uint32_t)kChunkReads, last - c0);
    __syncthreads();
    for (uint32_t i = tid; i <= cn_max; i += kTileWords) S.rec[i] = R.rec[c0 + i];
    if (tid == 0) S.cut = cn_max;
    __syncthreads();
    const uint32_t pbase = S.rec[0].pair_off;
    for (uint32_t i = tid + 1; i <= cn_max; i += kTileWords)
      if (S.rec[i].pair_off - pbase > (uint32_t)kChunkPairs) atomicMin(&S.cut, i - 1);
    __syncthreads();
    const uint32_t cn = S.cut;
    const uint32_t np = S.rec[cn].pair_off - pbase;
    for (uint32_t i = tid; i < np + 2; i += kTileWords) S.pairs[i] = R.pairs[pbase + i];
    // prefix maximum of `end` over the chunk: reads before the first index with pmax > wbase cannot reach this word
    {
      int m = -0x7FFFFFFF;
      const uint32_t per = (cn + kTileWords - 1) / kTileWords;
      const uint32_t b0 = min(cn, tid * per), b1 = min(cn, b0 + per);
      for (uint32_t i = b0; i < b1; ++i) m = max(m, S.rec[i].end);
      int incl = m;
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((tid & 31) >= o) incl = max(incl, t);
      }
      if ((tid & 31) == 31) S.warp_max[tid >> 5] = incl;
      __syncthreads();
      int before = -0x7FFFFFFF;
      for (int k = 0; k < (tid >> 5); ++k) before = max(before, S.warp_max[k]);
      int excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
      if ((tid & 31) == 0) excl = -0x7FFFFFFF;
      int run = max(before, excl);
      for (uint32_t i = b0; i < b1; ++i) {
        run = max(run, S.rec[i].end);
        S.pmax[i] = run;
      }
    }
    __syncthreads();

    // ---- this thread's reads in the chunk: [lb, ub)
    uint32_t lb, ub;
    {
      uint32_t lo_i = 0, hi_i = cn;  // first i with rec[i].start >= wbase + 32
      while (lo_i < hi_i) {
        uint32_t mid = (lo_i + hi_i) >> 1;
        if (S.rec[mid].start >= wbase + 32) hi_i = mid; else lo_i = mid + 1;
      }
      ub = lo_i;
      lo_i = 0; hi_i = ub;           // first i with pmax[i] > wbase
      while (lo_i < hi_i) {
        uint32_t mid = (lo_i + hi_i) >> 1;
        if (S.pmax[mid] > wbase) hi_i = mid; else lo_i = mid + 1;
      }
      lb = lo_i;
    }
    for (uint32_t i = lb; i < ub; i += 4) {
      Aligned a[4];
      uint32_t pos[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = Aligned{0, 0, 0, 0};
        pos[u] = 0;
        const uint32_t idx = i + u;
        if (idx < ub) {
          const ReadRec rec = S.rec[idx];
          if (rec.end > wbase) {
            const uint2* P = S.pairs + (rec.pair_off - pbase);
            if (rec.info & kInfoSimple) {
              a[u] = align_simple(rec, P, wbase);
              if (rec.info & kInfoHasExc)
                apply_exceptions(a[u], R, rec.pair_off, (int)(rec.info & kInfoLeadMask) + (wbase - rec.start), a[u].plain);
            } else {
              a[u] = align_cigar(R, c0 + idx, rec, P, wbase);
            }
            if (MODE == 1 && (rec.info & kInfoPositive)) pos[u] = a[u].plain | a[u].other;
          }
        }
      }
      ovf |= cV.add4(a[0].plain, a[1].plain, a[2].plain, a[3].plain);
      cL.add4(a[0].lo, a[1].lo, a[2].lo, a[3].lo);
      cH.add4(a[0].hi, a[1].hi, a[2].hi, a[3].hi);
      cHL.add4(a[0].lo & a[0].hi, a[1].lo & a[1].hi, a[2].lo & a[2].hi, a[3].lo & a[3].hi);
      uint32_t oc = cO.add4(a[0].other, a[1].other, a[2].other, a[3].other);
      if (MODE == 1) { ovf |= oc; ovf |= cP.add4(pos[0], pos[1], pos[2], pos[3]); } else o_sat |= oc;
    }
    c0 += cn;
  }
  if (ovf) atomicAdd(&out.counters[5], 1ull);

  // ---- epilogue: the caller, per locus of this word ----------------------------------------------------------------------
  if (w >= ci.n_words) return;
  const uint32_t in_range = bit_range(td.locus_begin - wbase, td.locus_end - wbase);
  if (!in_range) return;
  const uint32_t rlo = R.trk_lo[ci.word_off + w], rhi = R.trk_hi[ci.word_off + w], rstd = R.trk_std[ci.word_off + w];
  const uint32_t anyO = cO.any() | o_sat;
  const uint32_t covered = cV.any() | anyO;
  uint32_t mm = 0;
#pragma unroll
  for (int k = 0; k < W; ++k) mm |= (cL.p[k] ^ (rlo & cV.p[k])) | (cH.p[k] ^ (rhi & cV.p[k]));
  uint32_t visit = (prm.skip_empty ? covered : 0xFFFFFFFFu) & in_range;
  if (MODE == 0 && !prm.skip_empty) visit &= covered;  // callVariantsAtLocus returns nothing on an empty pileup
  atomicAdd(&out.counters[3], (unsigned long long)__popc(visit));
  uint32_t todo = visit;
  if (MODE == 0 && !prm.emit_ref && !prm.emit_no_call) todo &= (mm | anyO | ~rstd);
  while (todo) {
    const int b = __ffs(todo) - 1;
    todo &= todo - 1;
    const int locus = wbase + b;
    const int v = cV.at(b), l = cL.at(b), h = cH.at(b), hl = cHL.at(b), o = cO.at(b);
    const bool o_over = (o_sat >> b) & 1u;
    const bool std_ref = (rstd >> b) & 1u;
    int cnt[4] = {v - l - h + hl, l - hl, h - hl, hl};
    const int rcode = (int)(((rlo >> b) & 1u) | (((rhi >> b) & 1u) << 1));
    const uint8_t rbase = code_base(rcode);
    if (MODE == 1) {
      if (!std_ref && (v + o) > 0) {
        uint32_t s = (uint32_t)atomicAdd(&out.counters[2], 1ull);
        if (s < out.cap_slow) out.slow[s] = SlowLocus{td.contig, locus};
        continue;
      }
      uint32_t s = (uint32_t)atomicAdd(&out.counters[0], 1ull);
      if (s < out.cap_rec) {
        guac_locus_counts c;
        c.locus = locus;
        c.contig = td.contig;
        c.depth = v + o;
        c.positive_depth = cP.at(b);
        c.reference_depth = std_ref ? cnt[rcode] : 0;
        c.base_count[0] = cnt[0]; c.base_count[1] = cnt[1]; c.base_count[2] = cnt[2]; c.base_count[3] = cnt[3];
        c.other_count = o;
        c.reference_base = std_ref ? rbase : (uint8_t)'N';
        c.pad_[0] = c.pad_[1] = c.pad_[2] = 0;
        out.crec[s] = c;
      }
      continue;
    }
    // ---- GermlineThreshold.Caller.callVariantsAtLocus on the SNV alleles -------------------------------------------------
    const int total = v + o;
    // any allele made of "other" elements has count <= o: if o cannot pass the threshold the SNV counts decide alone
    const bool exact = std_ref && !o_over && ((long long)o * 100 / total <= prm.threshold_percent);
    if (!exact) {
      uint32_t s = (uint32_t)atomicAdd(&out.counters[2], 1ull);
      if (s < out.cap_slow) out.slow[s] = SlowLocus{td.contig, locus};
      continue;
    }
    int n = 0, sc[4], sb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (cnt[k] > 0 && (long long)cnt[k] * 100 / total > prm.threshold_percent) {
        int j = n++;  // stable insertion sort: count descending, allele order (= base code order) ascending on ties
        while (j > 0 && sc[j - 1] < cnt[k]) { sc[j] = sc[j - 1]; sb[j] = sb[j - 1]; --j; }
        sc[j] = cnt[k];
        sb[j] = k;
      }
    const uint8_t tie = (n >= 3 && sc[1] == sc[2]) ? 1 : 0;
    uint8_t e_alt[2], e_g0[2], e_g1[2];
    bool e_sym[2];
    int ne = 0;
    auto emit = [&](int code, bool sym, uint8_t g0, uint8_t g1) { e_alt[ne] = code_base(code); e_sym[ne] = sym; e_g0[ne] = g0; e_g1[ne] = g1; ++ne; };
    if (n == 0) {
      if (prm.emit_no_call) emit(0, true, GUAC_GT_NO_CALL, GUAC_GT_NO_CALL);
    } else if (n == 1) {
      if (sb[0] == rcode) { if (prm.emit_ref) emit(0, true, GUAC_GT_REF, GUAC_GT_REF); }
      else emit(sb[0], false, GUAC_GT_ALT, GUAC_GT_ALT);
    } else {
      const bool v1 = sb[0] != rcode, v2 = sb[1] != rcode;
      if (v1 != v2) emit(v1 ? sb[0] : sb[1], false, GUAC_GT_REF, GUAC_GT_ALT);
      else { emit(sb[0], false, GUAC_GT_ALT, GUAC_GT_OTHER_ALT); emit(sb[1], false, GUAC_GT_ALT, GUAC_GT_OTHER_ALT); }
    }
    if (tie) atomicAdd(&out.counters[4], 1ull);
    for (int k = 0; k < ne; ++k) {
      uint32_t s = (uint32_t)atomicAdd(&out.counters[0], 1ull);
      if (s < out.cap_rec) {
        guac_threshold_record r;
        r.start = locus;
        r.contig = td.contig;
        r.sample = prm.sample;
        r.ref_off = kPoolByteOff + rbase;
        r.ref_len = 1;
        r.alt_off = e_sym[k] ? kPoolAltOff : kPoolByteOff + e_alt[k];
        r.alt_len = e_sym[k] ? 5 : 1;
        r.gt[0] = e_g0[k];
        r.gt[1] = e_g1[k];
        r.tie = tie;
        r.pad_ = 0;
        out.trec[s] = r;
      }
    }
  }
}

// ---- the exact per-element walk ------------------------------------------------------------------------------------------
enum ElemKind : int { kMatch = 0, kMismatch = 1, kInsertion = 2, kDeletion = 3, kMidDeletion = 4, kClipped = 5, kNone = -1 };

struct Elem {
  int kind;
  int qual;        // PileupElement.qualityScore
  uint8_t base;    // Match/Mismatch: read base; MidDeletion: MD base
  int len;         // Insertion: anchor + inserted bases; Deletion: deleted bases
  uint64_t ptr;    // Insertion: offset into seq; Deletion: offset into md of the first deleted base
};

// offset into R.md of the deleted base at reference position `pos` of read r (inside a D op), or -1
__device__ long md_deleted_offset(const DevReads& R, uint64_t r, int pos) {
  // walk CIGAR and MD together; deleted bases of one D op are contiguous in the tag
  const ReadRec rec = R.rec[r];
  const uint32_t c0 = R.cig_off[r], c1 = R.cig_off[r + 1];
  const char* md = R.md + R.md_off[r];
  const int md_len = (int)(R.md_off[r + 1] - R.md_off[r]);
  int p = 0;
  long pending = 0;
  int ref_pos = rec.start;
  for (uint32_t c = c0; c < c1; ++c) {
    const uint32_t op = R.cigar[c] & 0xF;
    int len = (int)(R.cigar[c] >> 4);
    if (op_is_match_like(op)) {
      int remaining = len;
      while (remaining > 0) {
        if (pending > 0) { long k = pending < remaining ? pending : remaining; remaining -= (int)k; pending -= k; ref_pos += (int)k; }
        else if (p >= md_len) return -1;
        else if (md[p] >= '0' && md[p] <= '9') { long n = 0; while (p < md_len && md[p] >= '0' && md[p] <= '9') n = n * 10 + (md[p++] - '0'); pending = n; }
        else if (md[p] == '^') return -1;
        else { ++p; --remaining; ++ref_pos; }
      }
    } else if (op == GUAC_CIGAR_D) {
      int remaining = len;
      while (remaining > 0) {
        if (pending > 0 || p >= md_len) return -1;
        if (md[p] >= '0' && md[p] <= '9') { long n = 0; while (p < md_len && md[p] >= '0' && md[p] <= '9') n = n * 10 + (md[p++] - '0'); pending = n; }
        else if (md[p] == '^') ++p;
        else { if (ref_pos == pos) return (long)(R.md_off[r] + p); ++p; --remaining; ++ref_pos; }
      }
    } else if (op == GUAC_CIGAR_N) {
      ref_pos += len;
    }
    if (ref_pos > pos) return -1;
  }
  return -1;
}

// PileupElement(read, locus, referenceBase) + alignment + qualityScore  (pileup/PileupElement.scala:68-171, 220-274)
__device__ int classify(const DevReads& R, uint64_t r, int locus, uint8_t ref_base, Elem& e) {
  const ReadRec rec = R.rec[r];
  const uint32_t c0 = R.cig_off[r], c1 = R.cig_off[r + 1];
  const uint8_t* seq = R.seq + R.seq_off[r];
  const uint8_t* qual = R.qual ? R.qual + R.seq_off[r] : nullptr;  // absent when packed without qualities
  const int read_len = (int)(R.seq_off[r + 1] - R.seq_off[r]);
  const int mapq = (int)(rec.info >> kInfoMapqShift);
  int ref_pos = rec.start, read_pos = 0;
  e.kind = kNone;
  for (uint32_t c = c0; c < c1; ++c) {
    const uint32_t op = R.cigar[c] & 0xF;
    const int len = (int)(R.cigar[c] >> 4);
    const int ref_len = op_consumes_ref(op) ? len : 0;
    const bool here = ref_pos <= locus && locus < ref_pos + ref_len;
    const bool stay_on_insertion = !here && locus == 0 && op == GUAC_CIGAR_I;  // insertion at the start of a contig
    if (!here && !stay_on_insertion) {
      if (op_consumes_read(op)) read_pos += len;
      ref_pos += ref_len;
      continue;
    }
    const int idx = here ? locus - ref_pos : 0;
    const int rp = read_pos + ((here && op_consumes_read(op)) ? idx : 0);
    const bool is_final = idx == len - 1;
    const bool has_next = c + 1 < c1;
    const uint32_t next_op = is_final ? (has_next ? (R.cigar[c + 1] & 0xF) : 0xFFu) : op;
    const int next_len = has_next ? (int)(R.cigar[c + 1] >> 4) : 0;
    auto insertion = [&](int ins_len) {
      int from = min(max(rp, 0), read_len), until = min(rp + ins_len + 1, read_len);
      if (until <= from) return (int)GUAC_ERR_INVALID_CIGAR;
      e.kind = kInsertion;
      e.len = until - from;
      e.ptr = R.seq_off[r] + from;
      int q = 255;
      for (int k = from; k < until; ++k) q = min(q, qual ? (int)(int8_t)qual[k] : 0);
      e.qual = q;
      e.base = seq[from];
      return 0;
    };
    if ((op == GUAC_CIGAR_M || op == GUAC_CIGAR_EQ) && next_op == GUAC_CIGAR_I) return insertion(next_len);
    if (op == GUAC_CIGAR_I && next_op != 0xFFu && ref_pos == 0) return insertion(len);
    if (op == GUAC_CIGAR_I) return GUAC_ERR_INVALID_CIGAR;
    if (op_is_match_like(op) && next_op == GUAC_CIGAR_D) {
      long off = md_deleted_offset(R, r, locus + 1);
      if (off < 0 || rp < 0 || rp >= read_len) return GUAC_ERR_MISSING_MD;
      // all next_len deleted bases must be present in the tag
      if (md_deleted_offset(R, r, locus + next_len) != off + next_len - 1) return GUAC_ERR_MISSING_MD;
      e.kind = kDeletion;
      e.len = next_len;
      e.ptr = (uint64_t)off;
      e.qual = qual ? (int)(int8_t)qual[rp] : 0;
      e.base = ref_base;
      return 0;
    }
    if (op == GUAC_CIGAR_D) {
      long off = md_deleted_offset(R, r, locus);
      if (off < 0) return GUAC_ERR_MISSING_MD;
      e.kind = kMidDeletion;
      e.base = (uint8_t)R.md[off];
      e.qual = mapq;
      e.len = 0;
      return 0;
    }
    if (next_op == GUAC_CIGAR_D) return GUAC_ERR_INVALID_CIGAR;  // deletion preceded by a non-match operator
    if (op_is_match_like(op)) {
      if (rp < 0 || rp >= read_len) return GUAC_ERR_INVALID_CIGAR;
      e.base = seq[rp];
      e.qual = qual ? (int)(int8_t)qual[rp] : 0;
      e.kind = (e.base == ref_base) ? kMatch : kMismatch;
      e.len = 1;
      return 0;
    }
    e.kind = kClipped;  // N (S and H have no reference length)
    e.qual = mapq;
    e.len = 0;
    return 0;
  }
  return 0;
}

// ---- allele table of one locus ---------------------------------------------------------------------------------------------
constexpr int kMaxAlleles = 48;

struct AlleleEntry {
  int kind;      // 0 SNV (Match/Mismatch), 2 insertion, 3 deletion, 4 mid-deletion, 5 clipped
  int len;
  uint64_t ptr;
  uint8_t base;
  int count;
};

struct AlleleView {
  const DevReads& R;
  uint8_t ref_base;
  __device__ int ref_len(const AlleleEntry& a) const { return a.kind == 3 ? 1 + a.len : (a.kind == 5 ? 0 : 1); }
  __device__ int alt_len(const AlleleEntry& a) const { return a.kind == 2 ? a.len : (a.kind == 4 || a.kind == 5 ? 0 : 1); }
  __device__ uint8_t ref_at(const AlleleEntry& a, int i) const {
    switch (a.kind) {
      case 0: return ref_base;
      case 2: return R.seq[a.ptr];
      case 3: return i == 0 ? ref_base : (uint8_t)R.md[a.ptr + i - 1];
      default: return a.base;  // mid-deletion
    }
  }
  __device__ uint8_t alt_at(const AlleleEntry& a, int i) const {
    switch (a.kind) {
      case 0: return a.base;
      case 2: return R.seq[a.ptr + i];
      default: return ref_base;  // deletion
    }
  }
  __device__ bool is_variant(const AlleleEntry& a) const { return a.kind == 0 ? a.base != ref_base : a.kind != 5; }
  __device__ bool alt_empty(const AlleleEntry& a) const { return a.kind == 4 || a.kind == 5; }
  __device__ bool same(const AlleleEntry& a, const Elem& e) const {
    int ek = (e.kind == kMatch || e.kind == kMismatch) ? 0 : e.kind;
    if (a.kind != ek) return false;
    if (ek == 0 || ek == 4) return a.base == e.base;
    if (ek == 5) return true;
    if (a.len != e.len) return false;
    if (ek == 2) { for (int i = 0; i < a.len; ++i) if (R.seq[a.ptr + i] != R.seq[e.ptr + i]) return false; return true; }
    for (int i = 0; i < a.len; ++i) if (R.md[a.ptr + i] != R.md[e.ptr + i]) return false;
    return true;
  }
  // Allele.compare: java String.compareTo on ref, then alt (bytes widened by Byte.toChar)
  __device__ static int jchar(uint8_t b) { return (int)(uint16_t)(int16_t)(int8_t)b; }
  __device__ int compare(const AlleleEntry& a, const AlleleEntry& b) const {
    int la = ref_len(a), lb = ref_len(b), n = min(la, lb);
    for (int i = 0; i < n; ++i) { int d = jchar(ref_at(a, i)) - jchar(ref_at(b, i)); if (d) return d; }
    if (la != lb) return la - lb;
    la = alt_len(a); lb = alt_len(b); n = min(la, lb);
    for (int i = 0; i < n; ++i) { int d = jchar(alt_at(a, i)) - jchar(alt_at(b, i)); if (d) return d; }
    return la - lb;
  }
};

__device__ uint8_t reference_base_of(const DevReads& R, const ContigInfo& ci, int contig, int locus, bool* is_std) {
  const uint32_t w = ci.word_off + (uint32_t)(locus >> 5);
  const int b = locus & 31;
  *is_std = (R.trk_std[w] >> b) & 1u;
  if (*is_std) return code_base(((R.trk_lo[w] >> b) & 1u) | (((R.trk_hi[w] >> b) & 1u) << 1));
  if (R.fasta) {
    uint64_t o = R.fasta_off[contig] + (uint64_t)locus;
    if (o < R.fasta_off[contig + 1]) return R.fasta[o];
  }
  return 'N';  // Pileup.referenceBaseAtLocus: no read offers a standard base
}

__device__ uint32_t pool_alloc(DevOut& out, uint32_t n) {
  uint32_t o = (uint32_t)atomicAdd(&out.counters[1], (unsigned long long)n);
  return kPoolDynOff + o;
}

// ---- K_exact: thread per locus ------------------------------------------------------------------------------------------------
__device__ void exact_locus(const DevReads& R, int contig, int locus, const CallParams& prm, DevOut& out);

// grid-stride over the loci K_tile deferred; their number is read from the device counter (no host round trip)
__global__ void __launch_bounds__(64) k_exact_loci(DevReads R, const SlowLocus* __restrict__ loci, CallParams prm, DevOut out) {
  const uint32_t n_loci = (uint32_t)min(out.counters[2], (unsigned long long)out.cap_slow);
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_loci; t += gridDim.x * blockDim.x)
    exact_locus(R, loci[t].contig, loci[t].locus, prm, out);
}

__device__ void exact_locus(const DevReads& R, const int contig, const int locus, const CallParams& prm, DevOut& out) {
  const ContigInfo ci = R.contigs[contig];
  bool std_ref;
  const uint8_t ref_base = reference_base_of(R, ci, contig, locus, &std_ref);
  AlleleView av{R, ref_base};
  AlleleEntry tab[kMaxAlleles];
  int na = 0, total = 0, pos_depth = 0, ref_depth = 0, other = 0;
  int bc[4] = {0, 0, 0, 0};
  const int g = locus >> kGranuleShift;
  const uint32_t first = R.gran_first[ci.gran_off + g], last = R.gran_last[ci.gran_off + g];
  for (uint32_t r = first; r < last && first != 0xFFFFFFFFu; ++r) {
    const ReadRec rec = R.rec[r];
    if (rec.start > locus || rec.end <= locus) continue;
    Elem e;
    int rc = classify(R, r, locus, ref_base, e);
    if (rc || e.kind == kNone) {
      report_error(out.err, rc ? rc : GUAC_ERR_INVALID_CIGAR, r);
      return;
    }
    ++total;
    if (rec.info & kInfoPositive) ++pos_depth;
    if (e.kind == kMatch) ++ref_depth;
    if ((e.kind == kMatch || e.kind == kMismatch) && is_std_base(e.base)) ++bc[base_code(e.base)]; else ++other;
    if (prm.mode == 1) continue;
    int k = 0;
    for (; k < na; ++k)
      if (av.same(tab[k], e)) { ++tab[k].count; break; }
    if (k == na) {
      if (na == kMaxAlleles) { report_error(out.err, GUAC_ERR_UNSUPPORTED, ((unsigned long long)contig << 32) | (uint32_t)locus); return; }
      tab[na].kind = (e.kind == kMatch || e.kind == kMismatch) ? 0 : e.kind;
      tab[na].len = e.len;
      tab[na].ptr = e.ptr;
      tab[na].base = e.base;
      tab[na].count = 1;
      ++na;
    }
  }
  if (total == 0) return;
  if (prm.mode == 1) {
    uint32_t s = (uint32_t)atomicAdd(&out.counters[0], 1ull);
    if (s < out.cap_rec) {
      guac_locus_counts c;
      c.locus = locus; c.contig = contig; c.depth = total; c.positive_depth = pos_depth; c.reference_depth = ref_depth;
      c.base_count[0] = bc[0]; c.base_count[1] = bc[1]; c.base_count[2] = bc[2]; c.base_count[3] = bc[3];
      c.other_count = other; c.reference_base = ref_base; c.pad_[0] = c.pad_[1] = c.pad_[2] = 0;
      out.crec[s] = c;
    }
    return;
  }
  // counts.toList.filter(count * 100 / total > threshold).sortBy(-count)  — canonical pre-order: Allele.compare
  int idx[kMaxAlleles], n = 0;
  for (int k = 0; k < na; ++k)
    if ((long long)tab[k].count * 100 / total > prm.threshold_percent) {
      int j = n++;
      while (j > 0 && (tab[idx[j - 1]].count < tab[k].count ||
                       (tab[idx[j - 1]].count == tab[k].count && av.compare(tab[idx[j - 1]], tab[k]) > 0))) {
        idx[j] = idx[j - 1];
        --j;
      }
      idx[j] = k;
    }
  const uint8_t tie = (n >= 3 && tab[idx[1]].count == tab[idx[2]].count) ? 1 : 0;
  if (tie) atomicAdd(&out.counters[4], 1ull);
  auto emit = [&](const AlleleEntry* a, bool sym_ref, uint8_t sym_ref_base, uint8_t g0, uint8_t g1) {
    uint32_t s = (uint32_t)atomicAdd(&out.counters[0], 1ull);
    guac_threshold_record rcd;
    rcd.start = locus; rcd.contig = contig; rcd.sample = prm.sample; rcd.gt[0] = g0; rcd.gt[1] = g1; rcd.tie = tie; rcd.pad_ = 0;
    if (sym_ref) {
      rcd.ref_off = kPoolByteOff + sym_ref_base; rcd.ref_len = 1; rcd.alt_off = kPoolAltOff; rcd.alt_len = 5;
    } else {
      int rl = av.ref_len(*a), al = av.alt_len(*a);
      uint32_t o = pool_alloc(out, (uint32_t)(rl + al));
      if ((unsigned long long)o + rl + al <= out.cap_pool) {
        for (int i = 0; i < rl; ++i) out.pool[o + i] = av.ref_at(*a, i);
        for (int i = 0; i < al; ++i) out.pool[o + rl + i] = av.alt_at(*a, i);
      }
      rcd.ref_off = o; rcd.ref_len = (uint16_t)rl; rcd.alt_off = o + rl; rcd.alt_len = (uint16_t)al;
    }
    if (s < out.cap_rec) out.trec[s] = rcd;
  };
  if (n == 0) {
    if (prm.emit_no_call) emit(nullptr, true, ref_base, GUAC_GT_NO_CALL, GUAC_GT_NO_CALL);
  } else if (n == 1 && !av.is_variant(tab[idx[0]])) {
    if (prm.emit_ref) emit(nullptr, true, ref_base, GUAC_GT_REF, GUAC_GT_REF);
  } else if (n == 1) {
    emit(&tab[idx[0]], false, 0, GUAC_GT_ALT, GUAC_GT_ALT);
  } else {
    const AlleleEntry& a1 = tab[idx[0]];
    const AlleleEntry& a2 = tab[idx[1]];
    const bool v1 = av.is_variant(a1), v2 = av.is_variant(a2);
    if ((!v1 || !v2) && (av.alt_empty(a1) != av.alt_empty(a2))) {
      // heterozygous deletion: nothing
    } else if (v1 != v2) {
      emit(v1 ? &a1 : &a2, false, 0, GUAC_GT_REF, GUAC_GT_ALT);
    } else if (v1 && v2) {
      emit(&a1, false, 0, GUAC_GT_ALT, GUAC_GT_OTHER_ALT);
      emit(&a2, false, 0, GUAC_GT_ALT, GUAC_GT_OTHER_ALT);
    } else {
      // two non-variant alleles: only possible with differing reference bases
      const bool n1 = av.ref_len(a1) == 1 && av.ref_at(a1, 0) == 'N', n2 = av.ref_len(a2) == 1 && av.ref_at(a2, 0) == 'N';
      if (n1 || n2) {
        const AlleleEntry& p = n1 ? a2 : a1;
        emit(nullptr, true, av.ref_len(p) ? av.ref_at(p, 0) : (uint8_t)'N', GUAC_GT_REF, GUAC_GT_REF);
      } else {
        report_error(out.err, GUAC_ERR_MULTIPLE_REFERENCE_BASES, ((unsigned long long)contig << 32) | (uint32_t)locus);
      }
    }
  }
}

}  // namespace guac
