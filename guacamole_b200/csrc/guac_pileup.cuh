// guac_pileup.cuh — the pileup hot path: CIGAR expansion + per-locus allele / depth / strand counting + fused callers.
//
// K_tile   (k_pileup_tile): one WARP per granule of 1024 loci (4 warps per CTA, no block barrier).  Phase 1 is
//           READ-centric: each lane takes one read of the granule's candidate range (a contiguous index range of the
//           start-sorted store, found through the granule index) and replays the read's differences against the reference
//           track (mm[], built once at pack time by k_mismatch_lists: the CIGAR expansion, bit-parallel over 32 loci):
//           the per-locus counter tile is touched only where the read DIFFERS from the reference (sparse shared-memory
//           atomics: one per mismatch / insertion / deletion / clipped element) plus two atomics per read for the depth
//           difference array.  Reads with more differences than mm[] holds align their 2-bit base planes onto the
//           reference planes word by word with funnel shifts + XOR, walking their CIGAR in the kernel.
//           Phase 2 scans the difference array into depth (+ strand depth).  Phase 3 is LOCUS-centric: the caller.
//             GermlineThreshold.Caller.callVariantsAtLocus      commands/GermlineThresholdCaller.scala:90-179
//             Pileup.depth/positiveDepth/referenceDepth           pileup/Pileup.scala:76-91
//           Replaces SlidingWindow.setCurrentLocus (windowing/SlidingWindow.scala:83-110), Pileup.atGreaterLocus
//           (pileup/Pileup.scala:103-132) and PileupElement.advanceToLocus/alignment (pileup/PileupElement.scala:68-248).
// K_exact  (k_exact_loci): one WARP per locus that the counter tile cannot decide exactly (insertions, deletions,
//           clipped/N-skipped elements, non-ACGT bases, non-standard reference base): a literal per-element walk.
#pragma once

#include <type_traits>

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
  guac_threshold_record* trec;   // general records (exact kernel); per-locus rows in counts mode
  guac_locus_counts* crec;
  uint32_t cap_rec;
  unsigned long long* compact;   // single-base germline records of the tile kernels (guac_compact_record, include/guac.h)
  uint32_t cap_compact;
  uint8_t* pool;
  uint32_t cap_pool;
  SlowLocus* slow;
  uint32_t cap_slow;
  uint32_t slow_ctr;             // counters of this launch's segment of the call: deferred loci (8 + segment) ...
  uint32_t compact_ctr;          // ... and compact records (12 + segment)
  uint32_t* tile_base;           // per tile of the call: where its (sorted) compact records start in the segment's buffer ...
  uint32_t* tile_n;              // ... and how many there are: the egress kernels put the tiles' records in tile order
  uint32_t tile0;                // index of this launch's first tile in those arrays
  // counters: [0] general records, [1] pool bytes, [3] visited loci, [4] tie loci, [5] counter overflow, [7] a granule held
  // too many records for the device-side ordering (the host finishes the sort); per segment of a call (a call over many
  // tiles runs in up to four segments so that the exact kernel and the record egress of one overlap the tile kernel of the
  // next): [8 + s] deferred loci, [12 + s] compact records
  unsigned long long* counters;
  unsigned long long* work;      // k_exact_loci: per segment a ticket counter and the number of warps that are through (self-resetting)
  DevError* err;
};

struct TileDesc {
  int32_t contig;
  int32_t word0;       // first word of the tile (contig-relative)
  int32_t locus_begin; // requested loci of this tile: [locus_begin, locus_end)
  int32_t locus_end;
  uint32_t gran;       // germline tiles: the granule's global index (gs_hdr, gs_dd, gran_first) ...
  uint32_t trk_word;   // ... the global index of its first track word ...
  int32_t n_words;     // ... and how many of its 32 words lie inside the contig: no ContigInfo load on the hot path
  int32_t pad_;
};

// Records and deferred loci of one granule are staged in shared memory and leave with ONE global atomic per warp and kind
// (the counters are single addresses: a quarter of a million returning atomics on them serialise in L2).
constexpr int kStageRecords = 32, kStageSlow = 16;
struct EmitStage {
  unsigned long long rec[kStageRecords];
  unsigned long long slow[kStageSlow];   // contig << 32 | locus
  uint32_t n_rec, n_slow;
  uint32_t pad_[2];
};

struct CallParams {
  int32_t mode;             // 0 germline threshold, 1 per-locus counts, 2 per-allele counts (exact kernel only)
  int32_t threshold_percent;
  int32_t emit_ref;
  int32_t emit_no_call;
  int32_t skip_empty;
  int32_t sample;
};

// ---- compact records + the SNV part of callVariantsAtLocus, shared by both tile kernels --------------------------------------
// guac_compact_record (include/guac.h): contig 63..48 | start 47..16 | alt 15..13 (0 "<ALT>", 1..4 A C G T) | ref code 12..11 |
// gt0 10..9 | gt1 8..7 | tie 6.  Numeric order = canonical (contig, start, ref, alt) order.
__device__ __forceinline__ unsigned long long compact_record(int contig, int locus, int rcode, int alt, uint32_t g0, uint32_t g1, uint32_t tie) {
  return ((unsigned long long)(uint32_t)contig << 48) | ((unsigned long long)(uint32_t)locus << 16) | ((unsigned long long)alt << 13) |
         ((unsigned long long)rcode << 11) | ((unsigned long long)g0 << 9) | ((unsigned long long)g1 << 7) | ((unsigned long long)tie << 6);
}

__device__ __forceinline__ void defer_locus(DevOut& out, int contig, int locus, EmitStage* stage = nullptr) {
  if (stage) {
    const uint32_t k = atomicAdd(&stage->n_slow, 1u);
    if (k < (uint32_t)kStageSlow) {
      stage->slow[k] = ((unsigned long long)(uint32_t)contig << 32) | (uint32_t)locus;
      return;
    }
  }
  const uint32_t s = (uint32_t)atomicAdd(&out.counters[out.slow_ctr], 1ull);
  if (s < out.cap_slow) out.slow[s] = SlowLocus{contig, locus};
}

__device__ __forceinline__ void emit_compact(DevOut& out, unsigned long long rec, EmitStage* stage) {
  if (stage) {
    const uint32_t k = atomicAdd(&stage->n_rec, 1u);
    if (k < (uint32_t)kStageRecords) {
      stage->rec[k] = rec;
      return;
    }
  }
  const uint32_t s = (uint32_t)atomicAdd(&out.counters[out.compact_ctr], 1ull);
  if (s < out.cap_compact) out.compact[s] = rec;
}

// end of a granule (whole warp): the staged records and deferred loci leave with one global atomic each.  The records of
// the tile are put in canonical order on the way (a compact record's value orders like (contig, start, ref, alt); a tile
// holds a handful: every lane counts the records below its own) and the tile's slice is noted for the egress kernels, which
// lay the tiles' slices out in tile order: no sort of the whole output.
__device__ __forceinline__ void flush_stage(DevOut& out, EmitStage* stage, uint32_t tile) {
  __syncwarp();
  const int lane = threadIdx.x & 31;
  const uint32_t n_all = stage->n_rec;
  const uint32_t nr = min(n_all, (uint32_t)kStageRecords), ns = min(stage->n_slow, (uint32_t)kStageSlow);
  if (n_all > (uint32_t)kStageRecords && lane == 0) out.counters[7] = 1ull;  // the surplus left unordered: the host finishes
  if (nr) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&out.counters[out.compact_ctr], (unsigned long long)nr);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if ((uint32_t)lane < nr) {
      const unsigned long long v = stage->rec[lane];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < nr; ++j) {
        const unsigned long long w = stage->rec[j];
        rank += (w < v || (w == v && j < (uint32_t)lane)) ? 1u : 0u;
      }
      if (base + rank < out.cap_compact) out.compact[base + rank] = v;
    }
    if (lane == 0 && out.tile_n) out.tile_base[out.tile0 + tile] = (uint32_t)base;
  }
  if (lane == 0 && out.tile_n) out.tile_n[out.tile0 + tile] = nr;  // (every tile, also 0: the array is not cleared between calls)
  if (ns) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&out.counters[out.slow_ctr], (unsigned long long)ns);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if ((uint32_t)lane < ns && base + lane < out.cap_slow) {
      const unsigned long long v = stage->slow[lane];
      out.slow[base + lane] = SlowLocus{(int32_t)(v >> 32), (int32_t)(uint32_t)v};
    }
  }
}

// GermlineThreshold.Caller.callVariantsAtLocus (commands/GermlineThresholdCaller.scala:90-179) on the A/C/G/T counts of one
// locus: `total` elements, of which `o` are not plain bases and m1..m3 mismatch the reference base (code rcode) by class
// (read code ^ reference code).  Loci the counts cannot decide exactly go to the exact kernel.
__device__ __forceinline__ void call_snv_locus(const CallParams& prm, DevOut& out, int contig, int locus, int total, int o, int m1, int m2,
                                               int m3, int rcode, bool std_ref, bool every_covered, bool pure = false, EmitStage* stage = nullptr) {
  // count * 100 / total > threshold  <=>  count * 100 >= (threshold + 1) * total   (integers, no division)
  const long long bar = (long long)(prm.threshold_percent + 1) * total;
  auto passes = [&](int count) { return (long long)count * 100 >= bar; };
  // any allele made of "other" elements has count <= o: if o cannot pass the threshold the SNV counts decide alone.  `pure`:
  // every "other" element of the locus is a MidDeletion element carrying the reference base (k_expand) — ONE allele,
  // (reference base, ""), which precedes every single-base allele in Allele.compare order (equal ref, shorter alt).
  const bool del = passes(o);
  if (!std_ref || (del && !(pure && o > 0))) {
    defer_locus(out, contig, locus, stage);
    return;
  }
  const int mref = total - o - m1 - m2 - m3;
  if (!del && !every_covered && !passes(max(m1, max(m2, m3)))) return;  // no alternate allele passes
  // alleles in Allele.compare order (= the deletion allele, then base code order), stable-sorted by descending count: keep
  // the best three
  constexpr int kDel = 4;
  int c0 = -1, c1 = -1, c2 = -1, b0 = 0, b1 = 0, n = 0;
  if (del) { c0 = o; b0 = kDel; n = 1; }
#pragma unroll
  for (int code = 0; code < 4; ++code) {
    const int cls = code ^ rcode;
    const int cntv = cls == 0 ? mref : cls == 1 ? m1 : cls == 2 ? m2 : m3;
    if (cntv > 0 && passes(cntv)) {
      ++n;
      if (cntv > c0) { c2 = c1; c1 = c0; b1 = b0; c0 = cntv; b0 = code; }
      else if (cntv > c1) { c2 = c1; c1 = cntv; b1 = code; }
      else if (cntv > c2) { c2 = cntv; }
    }
  }
  const uint32_t tie = (n >= 3 && c1 == c2) ? 1u : 0u;
  if (del && (b0 == kDel || (n >= 2 && b1 == kDel))) {
    // the deletion allele is one of the two leading alleles.  Next to the reference allele: "heterozygous deletion", no
    // genotype (GermlineThresholdCaller.scala:146-149) — by far the commonest case.  Alone or next to an alternate base
    // its record carries an empty allele string: the exact kernel writes it.
    if (n >= 2 && (b0 == kDel ? b1 : b0) == rcode) {
      if (tie) atomicAdd(&out.counters[4], 1ull);
      return;
    }
    defer_locus(out, contig, locus, stage);
    return;
  }
  if (tie) atomicAdd(&out.counters[4], 1ull);
  int ne = 0, alt0 = 0, alt1 = 0;  // alt: 0 = "<ALT>", 1 + base code otherwise
  uint32_t g0 = 0, g1 = 0;
  if (n == 0) {
    if (prm.emit_no_call) { ne = 1; g0 = g1 = GUAC_GT_NO_CALL; }
  } else if (n == 1) {
    if (b0 == rcode) { if (prm.emit_ref) { ne = 1; g0 = g1 = GUAC_GT_REF; } }
    else { ne = 1; alt0 = 1 + b0; g0 = g1 = GUAC_GT_ALT; }
  } else {
    const bool v1 = b0 != rcode, v2 = b1 != rcode;
    if (v1 != v2) { ne = 1; alt0 = 1 + (v1 ? b0 : b1); g0 = GUAC_GT_REF; g1 = GUAC_GT_ALT; }
    else { ne = 2; alt0 = 1 + b0; alt1 = 1 + b1; g0 = GUAC_GT_ALT; g1 = GUAC_GT_OTHER_ALT; }
  }
  for (int k = 0; k < ne; ++k) emit_compact(out, compact_record(contig, locus, rcode, k == 0 ? alt0 : alt1, g0, g1, tie), stage);
}

// ---- K_tile ---------------------------------------------------------------------------------------------------------------
// One WARP owns one granule of 1024 loci and everything about it — its slice of shared memory, its reads, its scan, its
// calls — so the kernel has no block-wide barrier at all: warps of a CTA (and of the other resident CTAs) progress
// independently and hide one another's memory latency.  Reads crossing a granule boundary are handled by both owners
// (for 150 bp reads: 15 % more read work than one 4096-loci tile per CTA, bought back by the missing barrier stalls).
//
// Per-locus counter word: four fields of FB bits — [0] "other" elements, [1..3] mismatches by class (lo ^ ref_lo) |
// (hi ^ ref_hi) << 1, i.e. read base code = reference code ^ class.  CntT = uint32_t (8-bit fields, pileups < 256 deep) or
// uint64_t (16-bit fields, < 65536 deep).  Depth (and, in counts mode, positive-strand depth) come from difference arrays.
constexpr int kWarpWords = kGranuleLoci / 32;   // 32 words = 1024 loci per warp
constexpr int kWarpLoci = kGranuleLoci;
constexpr int kWarpsPerCta = 4;
constexpr int kTileThreads = kWarpsPerCta * 32;
constexpr int kListCap = 64;                    // reads of one granule that need the CIGAR walk (overflow handled inline)
// Depth difference array of one granule.  PACKED (8-bit counter fields, < 2048 reads per granule): two loci per 32-bit word,
// each half biased by 0x4000 so that +1 / -1 never carry into the neighbour — half the shared memory, hence more resident
// warps.  Otherwise one 32-bit word per locus.  One pad word per lane's 32 loci keeps the scan conflict-free.
template <bool PACKED>
struct CovArray {
  static constexpr int kWords = PACKED ? (kWarpLoci / 2 + kWarpLoci / 32 + 4) : (kWarpLoci + kWarpLoci / 32 + 36);
  alignas(16) uint32_t w[kWords];
  static __device__ __forceinline__ int index(int i) { return PACKED ? (i >> 1) + (i >> 5) : i + (i >> 5); }
  __device__ __forceinline__ void clear(int lane) {
    static_assert(kWords % 4 == 0, "cleared with 16-byte stores");
    const uint32_t v = PACKED ? 0x40004000u : 0u;
    uint4* w4 = reinterpret_cast<uint4*>(w);
    for (int i = lane; i < kWords / 4; i += 32) w4[i] = make_uint4(v, v, v, v);
  }
  __device__ __forceinline__ void start(int i) { atomicAdd(&w[index(i)], PACKED ? (1u << (16 * (i & 1))) : 1u); }
  __device__ __forceinline__ void end(int i) { atomicSub(&w[index(i)], PACKED ? (1u << (16 * (i & 1))) : 1u); }
  __device__ __forceinline__ int get(int i) const { return PACKED ? (int)((w[index(i)] >> (16 * (i & 1))) & 0xFFFFu) : (int)w[index(i)]; }
  // inclusive scan over the granule (lane owns loci [32 lane, 32 lane + 32)); returns the bit mask of the lane's loci with
  // depth > 0 and ORs `over` when a depth exceeds `limit`
  // `min_covered`: smallest non-zero depth among the lane's 32 loci (0xFFFFFFFF if none is covered)
  __device__ __forceinline__ uint32_t scan(int lane, uint32_t limit, bool& over, uint32_t* min_covered = nullptr) {
    const int base = PACKED ? lane * 17 : lane * 33;
    uint32_t mn = 0xFFFFFFFFu;
    int run = 0;
    if (PACKED) {
#pragma unroll
      for (int k = 0; k < 16; ++k) run += (int)(w[base + k] & 0xFFFFu) + (int)(w[base + k] >> 16) - 0x8000;
    } else {
#pragma unroll 8
      for (int k = 0; k < 32; ++k) run += (int)w[base + k];
    }
    int incl = run;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += t;
    }
    int acc = incl - run;
    uint32_t covered = 0;
    if (PACKED) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint32_t v = w[base + k];
        acc += (int)(v & 0xFFFFu) - 0x4000;
        const uint32_t d0 = (uint32_t)acc;
        acc += (int)(v >> 16) - 0x4000;
        const uint32_t d1 = (uint32_t)acc;
        w[base + k] = d0 | (d1 << 16);
        covered |= (d0 != 0u ? 1u : 0u) << (2 * k);
        covered |= (d1 != 0u ? 1u : 0u) << (2 * k + 1);
        over |= d0 > limit || d1 > limit;
        mn = min(mn, min(d0 - 1u, d1 - 1u));  // (0 - 1 wraps to the maximum: uncovered loci do not count)
      }
    } else {
#pragma unroll 8
      for (int k = 0; k < 32; ++k) {
        acc += (int)w[base + k];
        w[base + k] = (uint32_t)acc;
        covered |= (acc != 0 ? 1u : 0u) << k;
        over |= (uint32_t)acc > limit;
        mn = min(mn, (uint32_t)acc - 1u);
      }
    }
    if (min_covered) *min_covered = mn == 0xFFFFFFFFu ? mn : mn + 1u;
    return covered;
  }
};

struct NoCovArray {
  __device__ __forceinline__ void clear(int) {}
  __device__ __forceinline__ void start(int) {}
  __device__ __forceinline__ void end(int) {}
  __device__ __forceinline__ int get(int) const { return 0; }
  __device__ __forceinline__ uint32_t scan(int, uint32_t, bool&, uint32_t* = nullptr) { return 0u; }
};

template <typename CntT, int MODE>
struct WarpSmem {
  uint32_t ref_lo[kWarpWords], ref_hi[kWarpWords], ref_std[kWarpWords];   // ref_hi must follow ref_lo
  CovArray<sizeof(CntT) == 4> cov;
  typename std::conditional<MODE == 1, CovArray<sizeof(CntT) == 4>, NoCovArray>::type pos;  // positive-strand depth (counts mode)
  alignas(16) CntT cnt[kWarpLoci];  // (phase 3 reads four counter words per lane)
  uint32_t list[kListCap];
  uint32_t n_list;
  uint32_t pad_[3];
};

// same, callable from divergent code (no reconvergence point inside)
template <typename CntT>
__device__ __forceinline__ void count_bits_nosync(CntT* cnt_word, uint32_t bits, int cls) {
  constexpr int FB = sizeof(CntT) * 2;
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    if constexpr (sizeof(CntT) == 8)  // (two native 32-bit halves: a 64-bit shared-memory atomic is a CAS loop)
      atomicAdd(reinterpret_cast<uint32_t*>(cnt_word + b) + (cls >> 1), 1u << (FB * (cls & 1)));
    else
      atomicAdd(cnt_word + b, (CntT)1 << (FB * cls));
  }
}

// one shared-memory atomic per set bit; lanes loop independently and the warp reconverges right after
template <typename CntT>
__device__ __forceinline__ void count_bits(CntT* cnt_word, uint32_t bits, uint32_t x, uint32_t y) {
  constexpr int FB = sizeof(CntT) * 2;
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    const int cls = (int)((x >> b) & 1u) | ((int)((y >> b) & 1u) << 1);
    if constexpr (sizeof(CntT) == 8)
      atomicAdd(reinterpret_cast<uint32_t*>(cnt_word + b) + (cls >> 1), 1u << (FB * (cls & 1)));
    else
      atomicAdd(cnt_word + b, (CntT)1 << (FB * cls));
  }
  __syncwarp();
}

template <typename CntT, int MODE>
__global__ void __launch_bounds__(kTileThreads) k_pileup_tile(DevReads R, const TileDesc* __restrict__ tiles, uint32_t n_tiles, CallParams prm, DevOut out) {
  constexpr int FB = sizeof(CntT) * 2;
  constexpr uint32_t FMASK = (1u << FB) - 1u;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const uint32_t tile = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (tile >= n_tiles) return;  // whole warp
  WarpSmem<CntT, MODE>& S = reinterpret_cast<WarpSmem<CntT, MODE>*>(smem_raw)[threadIdx.x >> 5];
  const TileDesc td = tiles[tile];
  const ContigInfo ci = R.contigs[td.contig];
  const int tile_lo = td.word0 << 5;
  const int tile_hi = min(tile_lo + kWarpLoci, ci.n_words << 5);

  // ---- phase 0: clear the counter slice, stage the reference planes, candidate reads of the granule
  {  // 16-byte stores (cnt is 16-byte aligned)
    uint4* c4 = reinterpret_cast<uint4*>(S.cnt);
    for (int i = lane; i < (int)(kWarpLoci * sizeof(CntT) / 16); i += 32) c4[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  S.cov.clear(lane);
  if (MODE == 1) S.pos.clear(lane);
  {
    const int w = td.word0 + lane;
    const bool in = w < ci.n_words;
    S.ref_lo[lane] = in ? R.trk_lo[ci.word_off + w] : 0u;
    S.ref_hi[lane] = in ? R.trk_hi[ci.word_off + w] : 0u;
    S.ref_std[lane] = in ? R.trk_std[ci.word_off + w] : 0u;
  }
  if (lane == 0) S.n_list = 0;
  uint32_t first = 0xFFFFFFFFu, last = 0;
  {
    const int g = tile_lo >> kGranuleShift;
    if (g < ci.n_grans) {
      first = R.gran_first[ci.gran_off + g];
      last = R.gran_last[ci.gran_off + g];
    }
    if (first == 0xFFFFFFFFu) last = 0;
  }
  __syncwarp();

  // General path (reads with insertions / deletions / skips, or with non-ACGT bases): the run-length CIGAR is walked ONCE,
  // op by op, in lockstep across the warp.  Every M/=/X run is a segment that slides over the reference words exactly like
  // the fast path; insertion / deletion anchors, deleted and skipped loci and non-ACGT bases are "other" elements.
  auto mark_other = [&](int lo, int hi) {  // loci [lo, hi) of this lane's read that are not plain elements (divergent-safe)
    lo = max(lo, tile_lo);
    hi = min(hi, tile_hi);
    for (int w = (lo - tile_lo) >> 5; lo < hi && w <= (hi - 1 - tile_lo) >> 5; ++w) {
      const int wbase = tile_lo + (w << 5);
      count_bits_nosync<CntT>(S.cnt + (w << 5), bit_range(lo - wbase, hi - wbase), 0);
    }
  };
  auto general_read = [&](uint32_t r, const ReadRec rec, bool active) {
    uint32_t c = active ? R.cig_off[r] : 0u;
    const uint32_t c_end = active ? R.cig_off[r + 1] : 0u;
    const bool has_exc = (rec.info & kInfoHasExc) != 0;
    int ref_pos = rec.start, read_pos = 0;
    bool skip_first = false;  // contig-start insertion: the element at locus 0 is the insertion, not a plain base
    const int n_ops = (int)__reduce_max_sync(0xFFFFFFFFu, c_end - c);
    for (int oi = 0; oi < n_ops; ++oi) {
      int seg_ref = 0, seg_read = 0, seg_len = 0;
      if (c < c_end) {
        const uint32_t v = R.cigar[c];
        const uint32_t op = v & 0xF, next_op = (c + 1 < c_end) ? (R.cigar[c + 1] & 0xF) : 0xFFu;
        const int len = (int)(v >> 4);
        ++c;
        if (op_is_match_like(op)) {
          seg_ref = ref_pos;
          seg_read = read_pos;
          seg_len = len;
          if (skip_first) {
            mark_other(ref_pos, ref_pos + 1);
            ++seg_ref; ++seg_read; --seg_len;
            skip_first = false;
          }
          // (M|=, I) and (M|=|X, D): the run's last base is the insertion / deletion anchor (PileupElement.scala:93, 109)
          const bool anchor = (next_op == GUAC_CIGAR_I && (op == GUAC_CIGAR_M || op == GUAC_CIGAR_EQ)) || next_op == GUAC_CIGAR_D;
          if (anchor && seg_len > 0) {
            mark_other(ref_pos + len - 1, ref_pos + len);
            --seg_len;
          }
          ref_pos += len;
          read_pos += len;
        } else if (op == GUAC_CIGAR_D || op == GUAC_CIGAR_N) {
          mark_other(ref_pos, ref_pos + len);  // mid-deletion / skipped loci
          ref_pos += len;
        } else if (op == GUAC_CIGAR_I) {
          if (ref_pos == 0 && rec.start == 0) skip_first = true;
          read_pos += len;
        } else if (op == GUAC_CIGAR_S) {
          read_pos += len;
        }
      }
      __syncwarp();
      // the plain segment, clipped to the granule
      const int s = max(seg_ref, tile_lo), e = min(seg_ref + seg_len, tile_hi);
      const bool on_seg = seg_len > 0 && s < e;
      const int w0 = on_seg ? (s - tile_lo) >> 5 : 0, w1 = on_seg ? (e - 1 - tile_lo) >> 5 : -1;
      const int q0 = seg_read + (tile_lo + (w0 << 5) - seg_ref);
      const int sh = q0 & 31;
      const uint2* __restrict__ P = R.pairs + rec.pair_off + (q0 >> 5);
      const uint32_t* __restrict__ X = R.xmask + rec.pair_off + (q0 >> 5);
      uint2 pa = make_uint2(0u, 0u), pb = make_uint2(0u, 0u);
      uint32_t xa = 0, xb = 0;
      if (on_seg) {
        if (q0 >= 0) { pa = __ldg(P); if (has_exc) xa = __ldg(X); }
        pb = __ldg(P + 1);
        if (has_exc) xb = __ldg(X + 1);
      }
      const uint32_t first_mask = bit_range(seg_ref - (tile_lo + (w0 << 5)), 32);
      const uint32_t last_mask = bit_range(0, seg_ref + seg_len - (tile_lo + (w1 << 5)));
      const int nw = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)(w1 - w0 + 1));
      for (int k = 0; k < nw; ++k) {
        const int w = w0 + k;
        uint32_t x = 0, y = 0, oth = 0;
        if (on_seg && w <= w1) {
          uint32_t valid = k == 0 ? first_mask : 0xFFFFFFFFu;
          if (w == w1) valid &= last_mask;
          oth = __funnelshift_r(xa, xb, sh) & valid;  // non-ACGT bases
          valid &= ~oth;
          x = (__funnelshift_r(pa.x, pb.x, sh) ^ S.ref_lo[w]) & valid;
          y = (__funnelshift_r(pa.y, pb.y, sh) ^ S.ref_hi[w]) & valid;
          pa = pb;
          xa = xb;
          if (w < w1) { pb = __ldg(P + k + 2); if (has_exc) xb = __ldg(X + k + 2); }
        }
        const int ws = (on_seg && w <= w1) ? w : 0;
        count_bits<CntT>(S.cnt + (ws << 5), x | y, x, y);
        if (__any_sync(0xFFFFFFFFu, oth != 0)) count_bits<CntT>(S.cnt + (ws << 5), oth, 0u, 0u);
      }
    }
  };

  // ---- phase 1: one read per lane: two depth updates, then the read's planes / CIGAR walk.  The records of the next batch
  // are fetched before the current one is worked on.
  ReadRec rec_next{0, 0, 0, 0};
  if (first + lane < last) rec_next = R.rec[first + lane];
  for (uint32_t base = first; base < last; base += 32) {  // warp-uniform
    const uint32_t r = base + lane;
    const ReadRec rec = rec_next;
    rec_next = ReadRec{0, 0, 0, 0};
    if (r + 32 < last) rec_next = R.rec[r + 32];
    const bool active = r < last && rec.end > tile_lo && rec.start < tile_hi && rec.end > rec.start;
    if (active) {
      const int s = max(rec.start, tile_lo) - tile_lo, e = min(rec.end, tile_hi) - tile_lo;
      S.cov.start(s);
      S.cov.end(e);
      if (MODE == 1 && (rec.info & kInfoPositive)) {
        S.pos.start(s);
        S.pos.end(e);
      }
    }
    general_read(r, rec, active);  // planes / CIGAR walk, in lockstep across the warp
    __syncwarp();
  }
  __syncwarp();

  // ---- phase 2: inclusive scan of the difference array(s) -> depth (and positive-strand depth); visited loci counted here
  uint32_t n_visited = 0, word_min_depth = 0;  // (smallest non-zero depth of the lane's word: phase 3's first reject)
  bool overflow = false;
  {
    const uint32_t covered = S.cov.scan(lane, FMASK, overflow, &word_min_depth);  // a counter field may have wrapped: the host widens and reruns
    if (MODE == 1) {
      bool ignore = false;
      S.pos.scan(lane, 0xFFFFFFFFu, ignore);
    }
    const int l0 = tile_lo + (lane << 5);
    const uint32_t in_range = bit_range(td.locus_begin - l0, td.locus_end - l0);
    n_visited = __popc((prm.skip_empty ? covered : 0xFFFFFFFFu) & in_range);
  }
  __syncwarp();

  // ---- phase 3: the caller, one locus per lane per step (stride 32: conflict-free shared memory)
  const bool all_loci = MODE == 1 ? !prm.skip_empty : false;          // rows for empty pileups (counts mode only)
  const bool every_covered = MODE == 1 || prm.emit_ref || prm.emit_no_call;
  const int thr_plus_1 = prm.threshold_percent + 1;
  const int td_contig = td.contig, td_begin = td.locus_begin, td_end = td.locus_end;
  // the per-locus work past the cheap reject: the counts row (counts mode) or callVariantsAtLocus on the SNV alleles
  auto call_locus = [&](const int x) {
    const CntT c = S.cnt[x];
    const int w = x >> 5, b = x & 31;
    const bool std_ref = (S.ref_std[w] >> b) & 1u;
    const int locus = tile_lo + x;
    const int total = S.cov.get(x);
    if (total == 0 && !all_loci) return;  // callVariantsAtLocus returns nothing on an empty pileup
    const int o = (int)((uint32_t)c & FMASK);
    const int m1 = (int)((uint32_t)(c >> FB) & FMASK), m2 = (int)((uint32_t)(c >> (2 * FB)) & FMASK), m3 = (int)((uint32_t)(c >> (3 * FB)) & FMASK);
    const int rcode = (int)(((S.ref_lo[w] >> b) & 1u) | (((S.ref_hi[w] >> b) & 1u) << 1));
    const uint8_t rbase = code_base(rcode);
    if (MODE == 1) {
      if (!std_ref && total > 0) {
        defer_locus(out, td_contig, locus);
        return;
      }
      uint32_t s = (uint32_t)atomicAdd(&out.counters[0], 1ull);
      if (s < out.cap_rec) {
        guac_locus_counts g;
        g.locus = locus;
        g.contig = td_contig;
        g.depth = total;
        g.positive_depth = S.pos.get(x);
        g.reference_depth = std_ref ? total - o - m1 - m2 - m3 : 0;
        g.base_count[rcode] = total - o - m1 - m2 - m3;
        g.base_count[rcode ^ 1] = m1;
        g.base_count[rcode ^ 2] = m2;
        g.base_count[rcode ^ 3] = m3;
        g.other_count = o;
        g.reference_base = std_ref ? rbase : (uint8_t)'N';
        g.pad_[0] = g.pad_[1] = g.pad_[2] = 0;
        out.crec[s] = g;
      }
      return;
    }
    call_snv_locus(prm, out, td_contig, locus, total, o, m1, m2, m3, rcode, std_ref, every_covered);
  };
  // Sparse calls: the few loci of a granule that survive the reject (about three at a 1 % error rate) are parked in the
  // (now idle) read list and called afterwards side by side, instead of one lane at a time while 31 lanes wait.
  const bool park = MODE == 0 && !every_covered;
  if (park) {
    if (lane == 0) S.n_list = 0;
    __syncwarp();
  }
  bool sparse_done = false;
  if constexpr (sizeof(CntT) == 4) {
    if (park) {
      // sparse calls over 8-bit counter fields: four consecutive loci per lane and step (one 16-byte load), clean
      // quadruples skipped at once
      sparse_done = true;
      for (int it = 0; it < kWarpLoci / 128; ++it) {
        const int x0 = (it << 7) + (lane << 2);
        const uint4 c4 = *reinterpret_cast<const uint4*>(&S.cnt[x0]);
        const int w = x0 >> 5;
        const uint32_t std4 = (S.ref_std[w] >> (x0 & 31)) & 0xFu;
        const uint32_t wmin = __shfl_sync(0xFFFFFFFFu, word_min_depth, w);
        const uint32_t thr_word = wmin > 0x00FFFFFFu ? 0u : (uint32_t)thr_plus_1 * wmin;  // (no covered locus: no shortcut)
        if ((c4.x | c4.y | c4.z | c4.w) == 0u && std4 == 0xFu) continue;
        const uint32_t cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t c = cc[k];
          const bool std_ref = (std4 >> k) & 1u;
          if (c == 0u && std_ref) continue;
          const uint32_t differing = (c * 0x01010101u) >> 24;
          if (std_ref && differing * 100u < thr_word) continue;
          const int x = x0 + k, locus = tile_lo + x;
          if (locus < td_begin || locus >= td_end) continue;
          const int total = S.cov.get(x);
          if (std_ref && differing * 100u < (uint32_t)thr_plus_1 * (uint32_t)total) continue;
          const uint32_t slot = atomicAdd(&S.n_list, 1u);
          if (slot < (uint32_t)kListCap) S.list[slot] = (uint32_t)x;
          else call_locus(x);
        }
      }
    }
  }
  for (int x = lane; x < kWarpLoci && !sparse_done; x += 32) {
    const CntT c = S.cnt[x];
    const int w = x >> 5, b = x & 31;   // w is warp-uniform
    const bool std_ref = (S.ref_std[w] >> b) & 1u;
    const uint32_t wmin = __shfl_sync(0xFFFFFFFFu, word_min_depth, w);
    const uint32_t thr_word = wmin > 0x00FFFFFFu ? 0u : (uint32_t)thr_plus_1 * wmin;  // (no covered locus: no shortcut)
    if (c == 0 && std_ref && !every_covered) continue;  // every element matches the reference: nothing to call
    // first reject without touching the depth array: a dirty locus is covered, so its depth is at least the smallest
    // non-zero depth of its word (kept by the lane that scanned the word)
    uint32_t differing;
    if constexpr (sizeof(CntT) == 4) differing = ((uint32_t)c * 0x01010101u) >> 24;
    else differing = (uint32_t)(((unsigned long long)c * 0x0001000100010001ull) >> 48);
    if (std_ref && !every_covered && differing * 100u < thr_word) continue;
    const int locus = tile_lo + x;
    if (locus < td_begin || locus >= td_end) continue;
    const int total = S.cov.get(x);
    // cheap reject of most dirty loci: the four counter fields together (every element that differs from the reference;
    // one multiply sums them: a read adds at most one, so the sum cannot carry) are too few for any of them to pass the
    // threshold, so the only allele that can pass is the reference one and nothing is emitted.  (The exact per-class
    // tests follow for the loci that survive.)
    if (std_ref && !every_covered && differing * 100u < (uint32_t)thr_plus_1 * (uint32_t)total) continue;
    if (park) {
      const uint32_t slot = atomicAdd(&S.n_list, 1u);
      if (slot < (uint32_t)kListCap) { S.list[slot] = (uint32_t)x; continue; }
    }
    call_locus(x);
  }
  if (park) {
    __syncwarp();
    const uint32_t n_parked = min(S.n_list, (uint32_t)kListCap);
    for (uint32_t i = lane; i < n_parked; i += 32) call_locus((int)S.list[i]);
  }
  // one atomic per warp for the visited-loci counter
  for (int o = 16; o; o >>= 1) n_visited += __shfl_xor_sync(0xFFFFFFFFu, n_visited, o);
  if (lane == 0 && n_visited) atomicAdd(&out.counters[3], (unsigned long long)n_visited);
  if (overflow) atomicAdd(&out.counters[5], 1ull);
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
  {  // the read's first deletion was located at pack time (k_md_track)
    const int d0 = R.del_start[r];
    if (d0 < 0) return -1;
    if (pos >= d0 && pos < d0 + (int)R.del_len[r]) return (long)(R.md_off[r] + R.del_md[r] + (uint32_t)(pos - d0));
  }
  // any further deletion: walk CIGAR and MD together; deleted bases of one D op are contiguous in the tag
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

// Everything classify() needs to know about a read before it can look at a base: one round of independent loads (the exact
// kernels are chains of dependent DRAM round trips, ~1.5 us each; a warp decides one locus, so the length of the chain is the
// kernel's duration).
struct ReadMeta {
  ReadRec rec;
  uint64_t seq0, seq1;      // seq_off[r], seq_off[r + 1]
  uint32_t cig0, cig1;      // cig_off[r], cig_off[r + 1]
  uint32_t op[4];           // the first four CIGAR operators (whatever follows a short CIGAR's last one: never looked at)
  int32_t del_start;        // the read's first deletion (k_md_track): -1 = none
  uint32_t del_len, del_md, md0, md1;
};

__device__ __forceinline__ ReadMeta load_read_meta(const DevReads& R, const uint64_t r) {
  ReadMeta M;
  const uint4 rc = __ldg(reinterpret_cast<const uint4*>(R.rec + r));
  M.seq0 = __ldg(R.seq_off + r);
  M.seq1 = __ldg(R.seq_off + r + 1);
  M.cig0 = __ldg(R.cig_off + r);
  M.cig1 = __ldg(R.cig_off + r + 1);
  M.del_start = __ldg(R.del_start + r);
  M.del_len = __ldg(R.del_len + r);
  M.del_md = __ldg(R.del_md + r);
  M.md0 = __ldg(R.md_off + r);
  M.md1 = __ldg(R.md_off + r + 1);
  M.rec.start = (int32_t)rc.x;
  M.rec.end = (int32_t)rc.y;
  M.rec.pair_off = rc.z;
  M.rec.info = rc.w;
  const uint32_t n = M.cig1 - M.cig0;  // (second round, still before any decision)
#pragma unroll
  for (int i = 0; i < 4; ++i) M.op[i] = (uint32_t)i < n ? __ldg(R.cigar + M.cig0 + i) : 0u;
  return M;
}

// offset into R.md of the deleted base at reference position `pos`: the first deletion from the preloaded cache, any other
// through the walk
__device__ __forceinline__ long md_deleted_offset_meta(const DevReads& R, const uint64_t r, const ReadMeta& M, const int pos) {
  if (M.del_start < 0) return -1;
  if (pos >= M.del_start && pos < M.del_start + (int)M.del_len) return (long)(M.md0 + M.del_md + (uint32_t)(pos - M.del_start));
  return md_deleted_offset(R, r, pos);
}

// PileupElement(read, locus, referenceBase) + alignment + qualityScore  (pileup/PileupElement.scala:68-171, 220-274)
__device__ __forceinline__ int classify_meta(const DevReads& R, const uint64_t r, const ReadMeta& M, const int locus, const uint8_t ref_base, Elem& e) {
  const ReadRec rec = M.rec;
  const uint8_t* seq = R.seq + M.seq0;
  const uint8_t* qual = R.qual ? R.qual + M.seq0 : nullptr;  // absent when packed without qualities
  e.kind = kNone;
  if ((rec.info & kInfoSimple) && locus >= rec.start && locus < rec.end) {
    // one M/=/X run between clips: the element is a plain base, no CIGAR walk needed
    const int rp = (int)(rec.info & kInfoLeadMask) + (locus - rec.start);
    e.base = __ldg(seq + rp);
    e.qual = qual ? (int)(int8_t)__ldg(qual + rp) : 0;
    e.kind = (e.base == ref_base) ? kMatch : kMismatch;
    e.len = 1;
    return 0;
  }
  const uint32_t c0 = M.cig0, c1 = M.cig1;
  const int read_len = (int)(M.seq1 - M.seq0);
  const int mapq = (int)(rec.info >> kInfoMapqShift);
  int ref_pos = rec.start, read_pos = 0;
  auto op_at = [&](const uint32_t c) -> uint32_t {
    const uint32_t i = c - c0;
    return i == 0u ? M.op[0] : i == 1u ? M.op[1] : i == 2u ? M.op[2] : i == 3u ? M.op[3] : __ldg(R.cigar + c);
  };
  for (uint32_t c = c0; c < c1; ++c) {
    const uint32_t word = op_at(c);
    const uint32_t op = word & 0xF;
    const int len = (int)(word >> 4);
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
    const uint32_t next_word = has_next ? op_at(c + 1) : 0u;
    const uint32_t next_op = is_final ? (has_next ? (next_word & 0xF) : 0xFFu) : op;
    const int next_len = has_next ? (int)(next_word >> 4) : 0;
    auto insertion = [&](int ins_len) {
      int from = min(max(rp, 0), read_len), until = min(rp + ins_len + 1, read_len);
      if (until <= from) return (int)GUAC_ERR_INVALID_CIGAR;
      e.kind = kInsertion;
      e.len = until - from;
      e.ptr = M.seq0 + from;
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
      long off = md_deleted_offset_meta(R, r, M, locus + 1);
      if (off < 0 || rp < 0 || rp >= read_len) return GUAC_ERR_MISSING_MD;
      // all next_len deleted bases must be present in the tag: they follow contiguously (the tag was upper-cased at pack
      // time; '^' or a digit in between is what a second walk to the last deleted base would also trip over)
      if ((uint64_t)off + (uint64_t)next_len > (uint64_t)M.md1) return GUAC_ERR_MISSING_MD;
      for (int i = 1; i < next_len; i += 8) {  // (eight bytes per round trip)
        bool ok = true;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const char ch = i + j < next_len ? __ldg(R.md + off + i + j) : 'A';
          ok = ok && ch >= 'A' && ch <= 'Z';
        }
        if (!ok) return GUAC_ERR_MISSING_MD;
      }
      e.kind = kDeletion;
      e.len = next_len;
      e.ptr = (uint64_t)off;
      e.qual = qual ? (int)(int8_t)qual[rp] : 0;
      e.base = ref_base;
      return 0;
    }
    if (op == GUAC_CIGAR_D) {
      long off = md_deleted_offset_meta(R, r, M, locus);
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

// The likelihood kernels' copy (one per kernel, called from several places): they classify ninety reads per locus and are
// bound by the number of loads, not by their latency — every load only where it is needed.
// PileupElement(read, locus, referenceBase) + alignment + qualityScore  (pileup/PileupElement.scala:68-171, 220-274)
__device__ __noinline__ int classify(const DevReads& R, uint64_t r, int locus, uint8_t ref_base, Elem& e) {
  const ReadRec rec = R.rec[r];
  const uint8_t* seq = R.seq + R.seq_off[r];
  const uint8_t* qual = R.qual ? R.qual + R.seq_off[r] : nullptr;  // absent when packed without qualities
  e.kind = kNone;
  if ((rec.info & kInfoSimple) && locus >= rec.start && locus < rec.end) {
    // one M/=/X run between clips: the element is a plain base, no CIGAR walk needed
    const int rp = (int)(rec.info & kInfoLeadMask) + (locus - rec.start);
    e.base = seq[rp];
    e.qual = qual ? (int)(int8_t)qual[rp] : 0;
    e.kind = (e.base == ref_base) ? kMatch : kMismatch;
    e.len = 1;
    return 0;
  }
  const uint32_t c0 = R.cig_off[r], c1 = R.cig_off[r + 1];
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
      // all next_len deleted bases must be present in the tag: they follow contiguously (the tag was upper-cased at pack
      // time; '^' or a digit in between is what a second walk to the last deleted base would also trip over)
      if ((uint64_t)off + (uint64_t)next_len > (uint64_t)R.md_off[r + 1]) return GUAC_ERR_MISSING_MD;
      for (int i = 1; i < next_len; ++i) {
        const char ch = R.md[off + i];
        if (ch < 'A' || ch > 'Z') return GUAC_ERR_MISSING_MD;
      }
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
  bool batched = false;  // same(): eight bytes per round trip (the germline exact kernel, where a locus' latency is what counts)
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
    if (batched) {
      if (ek == 2) return bytes_equal(R.seq + a.ptr, R.seq + e.ptr, a.len);
      return bytes_equal(reinterpret_cast<const uint8_t*>(R.md) + a.ptr, reinterpret_cast<const uint8_t*>(R.md) + e.ptr, a.len);
    }
    if (ek == 2) { for (int i = 0; i < a.len; ++i) if (R.seq[a.ptr + i] != R.seq[e.ptr + i]) return false; return true; }
    for (int i = 0; i < a.len; ++i) if (R.md[a.ptr + i] != R.md[e.ptr + i]) return false;
    return true;
  }
  // eight bytes of each side per round trip (a byte-by-byte loop with an early exit is one dependent load per byte)
  __device__ static bool bytes_equal(const uint8_t* __restrict__ x, const uint8_t* __restrict__ y, const int n) {
    for (int i = 0; i < n; i += 8) {
      uint8_t a[8], b[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool in = i + j < n;
        a[j] = in ? __ldg(x + i + j) : (uint8_t)0;
        b[j] = in ? __ldg(y + i + j) : (uint8_t)0;
      }
      bool eq = true;
#pragma unroll
      for (int j = 0; j < 8; ++j) eq = eq && a[j] == b[j];
      if (!eq) return false;
    }
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

// Walks the reads that overlap one locus, 32 at a time with all lanes busy: the granule's candidate range (mostly reads
// that do NOT reach the locus) is scanned with one record test per lane and the hits are compacted, in read order, into
// a 64-entry ring in shared memory.
struct OverlapWalker {
  uint32_t* ring;   // 64 entries, per warp
  uint32_t src, last;
  uint32_t head, tail;
  int locus;
  __device__ OverlapWalker(const DevReads& R, uint32_t* r, uint32_t first, uint32_t last_, int l, bool ahead = false)
      : ring(r), src(first), last(last_), head(0), tail(0), locus(l) {
    if (first == 0xFFFFFFFFu) src = last = 0;
    else narrow_candidates(R, src, last, l, l + 1);
    // the scan below is one dependent load per 32 candidates: ask for the second and third round's records now
    // (the germline exact kernel only: a warp decides one locus there and the chain's length is the kernel's duration; the
    // likelihood kernels keep dozens of loci per warp in flight and are slower with the extra requests)
    const uint32_t lane = threadIdx.x & 31u;
    if (ahead && (lane & 1u) == 0u) {  // (two 16-byte records per sector)
      if (src + 32u + lane < last) asm volatile("prefetch.global.L2 [%0];" ::"l"(R.rec + src + 32u + lane));
      if (src + 64u + lane < last) asm volatile("prefetch.global.L2 [%0];" ::"l"(R.rec + src + 64u + lane));
    }
  }
  // warp-uniform; false when no read is left.  Lanes with valid == true hold one overlapping read each.
  __device__ bool next(const DevReads& R, uint32_t& r, ReadRec& rec, bool& valid) {
    const int lane = threadIdx.x & 31;
    while (tail - head < 32u && src < last) {
      const uint32_t rr = src + lane;
      bool hit = false;
      if (rr < last) {
        const ReadRec c = R.rec[rr];
        hit = c.start <= locus && c.end > locus;
      }
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
      if (hit) ring[(tail + __popc(m & ((1u << lane) - 1u))) & 63u] = rr;
      tail += __popc(m);
      src += 32;
      __syncwarp();
    }
    if (tail == head) return false;
    const uint32_t idx = head + lane;
    valid = idx < tail;
    r = valid ? ring[idx & 63u] : 0u;
    rec = ReadRec{0, 0, 0, 0};
    if (valid) rec = R.rec[r];
    head = min(head + 32u, tail);
    __syncwarp();
    return true;
  }
};

// ---- K_exact: one warp per locus --------------------------------------------------------------------------------------------
constexpr int kExactWarps = 4;

__device__ void exact_locus(const DevReads& R, const int contig, const int locus, const CallParams& prm, DevOut& out, AlleleEntry* tab, uint32_t* ring) {
  const int lane = threadIdx.x & 31;
  const ContigInfo ci = R.contigs[contig];
  // the track word and the granule's candidate range: five independent loads, one round trip
  const uint32_t tw = ci.word_off + (uint32_t)(locus >> 5), gi = ci.gran_off + (uint32_t)(locus >> kGranuleShift);
  const uint32_t t_std = __ldg(R.trk_std + tw), t_lo = __ldg(R.trk_lo + tw), t_hi = __ldg(R.trk_hi + tw);
  const uint32_t first = __ldg(R.gran_first + gi), last = __ldg(R.gran_last + gi);
  const int tb = locus & 31;
  uint8_t ref_base = code_base(((t_lo >> tb) & 1u) | (((t_hi >> tb) & 1u) << 1));
  if (!((t_std >> tb) & 1u)) {  // reference_base_of: no read offers a standard base here
    bool std_ref;
    ref_base = reference_base_of(R, ci, contig, locus, &std_ref);
  }
  AlleleView av{R, ref_base, /*batched=*/true};
  int na = 0, total = 0, pos_depth = 0, ref_depth = 0;
  int bc[4] = {0, 0, 0, 0};
  if (first == 0xFFFFFFFFu) return;
  OverlapWalker walk(R, ring, first, last, locus, /*ahead=*/true);
  uint32_t r;
  ReadRec rec;
  bool valid;
  while (walk.next(R, r, rec, valid)) {
    Elem e;
    e.kind = kNone;
    e.base = 0;
    e.len = 0;
    e.ptr = 0;
    int rc = 0;
    if (valid) {
      const ReadMeta M = load_read_meta(R, r);
      rc = classify_meta(R, r, M, locus, ref_base, e);
      if (rc == 0 && e.kind == kNone) rc = GUAC_ERR_INVALID_CIGAR;
    }
    if (__any_sync(0xFFFFFFFFu, rc != 0)) {
      if (rc) report_error(out.err, rc, r);
      return;
    }
    const bool snv = valid && (e.kind == kMatch || e.kind == kMismatch) && is_std_base(e.base);
    if (valid) {
      ++total;
      if (rec.info & kInfoPositive) ++pos_depth;
      if (e.kind == kMatch) ++ref_depth;
      if (snv) ++bc[base_code(e.base)];
    }
    if (prm.mode == 1) continue;
    // everything that is not an A/C/G/T match or mismatch goes through the allele table, one DISTINCT allele per round:
    // the first waiting lane's element is looked up (or entered) by lane 0, then every waiting lane compares its own
    // element with that entry in parallel and the whole group is counted at once
    uint32_t mask = __ballot_sync(0xFFFFFFFFu, valid && !snv);
    while (mask) {
      const int src = __ffs(mask) - 1;
      Elem s;
      s.kind = __shfl_sync(0xFFFFFFFFu, e.kind, src);
      s.len = __shfl_sync(0xFFFFFFFFu, e.len, src);
      s.base = (uint8_t)__shfl_sync(0xFFFFFFFFu, (int)e.base, src);
      s.ptr = ((uint64_t)__shfl_sync(0xFFFFFFFFu, (uint32_t)(e.ptr >> 32), src) << 32) | __shfl_sync(0xFFFFFFFFu, (uint32_t)e.ptr, src);
      s.qual = 0;
      int k = -1, created = 0;
      if (lane == 0) {
        for (int i = 0; i < na; ++i)
          if (av.same(tab[i], s)) { k = i; break; }
        if (k < 0 && na < kMaxAlleles) {
          k = na;
          created = 1;
          tab[k].kind = (s.kind == kMatch || s.kind == kMismatch) ? 0 : s.kind;
          tab[k].len = s.len;
          tab[k].ptr = s.ptr;
          tab[k].base = s.base;
          tab[k].count = 0;
        }
      }
      k = __shfl_sync(0xFFFFFFFFu, k, 0);
      created = __shfl_sync(0xFFFFFFFFu, created, 0);
      if (k < 0) {
        if (lane == 0) report_error(out.err, GUAC_ERR_UNSUPPORTED, ((unsigned long long)contig << 32) | (uint32_t)locus);
        return;
      }
      na += created;
      __syncwarp();
      const bool same = ((mask >> lane) & 1u) && (lane == src || av.same(tab[k], e));
      const uint32_t grp = __ballot_sync(0xFFFFFFFFu, same);
      if (lane == 0) tab[k].count += __popc(grp);
      mask &= ~grp;
      __syncwarp();
    }
  }
  for (int o = 16; o; o >>= 1) {
    total += __shfl_xor_sync(0xFFFFFFFFu, total, o);
    pos_depth += __shfl_xor_sync(0xFFFFFFFFu, pos_depth, o);
    ref_depth += __shfl_xor_sync(0xFFFFFFFFu, ref_depth, o);
#pragma unroll
    for (int k = 0; k < 4; ++k) bc[k] += __shfl_xor_sync(0xFFFFFFFFu, bc[k], o);
  }
  if (total == 0 || lane != 0) return;
  if (prm.mode == 1) {
    uint32_t s = (uint32_t)atomicAdd(&out.counters[0], 1ull);
    if (s < out.cap_rec) {
      guac_locus_counts c;
      c.locus = locus; c.contig = contig; c.depth = total; c.positive_depth = pos_depth; c.reference_depth = ref_depth;
      c.base_count[0] = bc[0]; c.base_count[1] = bc[1]; c.base_count[2] = bc[2]; c.base_count[3] = bc[3];
      c.other_count = total - bc[0] - bc[1] - bc[2] - bc[3]; c.reference_base = ref_base; c.pad_[0] = c.pad_[1] = c.pad_[2] = 0;
      out.crec[s] = c;
    }
    return;
  }
  // the A/C/G/T match / mismatch alleles join the table from their counters
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (bc[k] > 0) {
      if (na == kMaxAlleles) { report_error(out.err, GUAC_ERR_UNSUPPORTED, ((unsigned long long)contig << 32) | (uint32_t)locus); return; }
      tab[na].kind = 0; tab[na].len = 1; tab[na].ptr = 0; tab[na].base = code_base(k); tab[na].count = bc[k];
      ++na;
    }
  if (prm.mode == 2) {  // VariantSupport.Caller.pileupToAlleleCounts (commands/VariantSupport.scala:110-118)
    for (int k = 0; k < na; ++k) {
      const uint32_t s = (uint32_t)atomicAdd(&out.counters[0], 1ull);
      guac_allele_count rcd;
      rcd.start = locus; rcd.contig = contig; rcd.sample = prm.sample; rcd.count = tab[k].count;
      const int rl = av.ref_len(tab[k]), al = av.alt_len(tab[k]);
      const uint32_t o = pool_alloc(out, (uint32_t)(rl + al));
      if ((unsigned long long)o + rl + al <= out.cap_pool) {
        for (int i = 0; i < rl; ++i) out.pool[o + i] = av.ref_at(tab[k], i);
        for (int i = 0; i < al; ++i) out.pool[o + rl + i] = av.alt_at(tab[k], i);
      }
      rcd.ref_off = o; rcd.ref_len = (uint16_t)rl; rcd.alt_off = o + rl; rcd.alt_len = (uint16_t)al;
      if (s < out.cap_rec) reinterpret_cast<guac_allele_count*>(out.trec)[s] = rcd;
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
        for (int i = 0; i < rl + al; i += 8) {  // (the loads of eight bytes before their stores: one round trip per eight)
          uint8_t v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = i + j < rl ? av.ref_at(*a, i + j) : i + j < rl + al ? av.alt_at(*a, i + j - rl) : (uint8_t)0;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (i + j < rl + al) out.pool[o + i + j] = v[j];
        }
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


// The loci the tile kernel deferred; their number is read from the device counter (no host round trip).  out.trec / out.pool
// are device buffers in every mode: in germline mode k_general_to_host copies the few general records and their allele bytes
// to the pinned host block of the result afterwards.
__global__ void __launch_bounds__(kExactWarps * 32) k_exact_loci(DevReads R, const SlowLocus* __restrict__ loci, CallParams prm, DevOut out) {
  __shared__ AlleleEntry tabs[kExactWarps][kMaxAlleles];
  __shared__ uint32_t rings[kExactWarps][64];
  const uint32_t n_loci = (uint32_t)min(out.counters[out.slow_ctr], (unsigned long long)out.cap_slow);
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  // A locus is a chain of dependent round trips: 15 us for a plain one, four times that where a dozen reads carry a long
  // insertion or deletion, and a warp decides one or two of them.  The first comes by warp index, every further one from a
  // ticket counter, so that the warps that drew short loci take the rest (the counter and the count of warps that are
  // through reset themselves: the last warp out zeroes both).
  unsigned long long* ticket = out.work + 2 * (out.slow_ctr - 8u);
  for (uint32_t t = warp; t < n_loci;) {
    exact_locus(R, loci[t].contig, loci[t].locus, prm, out, tabs[threadIdx.x >> 5], rings[threadIdx.x >> 5]);
    __syncwarp();
    unsigned long long k = 0;
    if ((threadIdx.x & 31) == 0) k = atomicAdd(ticket, 1ull);
    t = n_warps + (uint32_t)__shfl_sync(0xFFFFFFFFu, (uint32_t)k, 0);
  }
  if ((threadIdx.x & 31) == 0 && atomicAdd(ticket + 1, 1ull) + 1ull == (unsigned long long)n_warps) {  // the last warp out
    ticket[0] = 0ull;
    ticket[1] = 0ull;
  }
}

// ---- record egress: the tiles' sorted slices laid out in tile order (compact germline records) ------------------------------
// One block per 256 tiles of a segment: the block sums tile_n[] of every earlier tile itself (a few hundred kilobytes from L2;
// no scan kernels in front of it: the whole egress is two small launches that fit next to the exact kernel; beyond 64 K
// tiles the host runs the three scan kernels first and passes the prefix) and scans its own 256 counts; thread per tile: its records go behind those of the earlier tiles and segments (whose counts are final: the
// tile kernels of a call run in order), to the contiguous device copy that k_rec_to_host and the NCCL gather send from.
__global__ void __launch_bounds__(256) k_rec_gather(const unsigned long long* __restrict__ rec, const uint32_t* __restrict__ tile_base,
                                                    const uint32_t* __restrict__ tile_n, const uint32_t* __restrict__ block_before, uint32_t n_tiles,
                                                    const unsigned long long* counters, uint32_t seg, uint32_t cap_seg,
                                                    unsigned long long* __restrict__ dev_rec, unsigned long long cap_total) {
  __shared__ unsigned long long warp_sum[8];
  __shared__ uint32_t warp_n[8];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  unsigned long long seg_base = 0;
  for (uint32_t j = 0; j < seg; ++j) seg_base += min(counters[12 + j], (unsigned long long)cap_seg);
  const uint32_t t0 = blockIdx.x * 256u;  // (a multiple of 4: tile_n is read 16 bytes at a time; the host aligns it)
  unsigned long long before = 0;
  if (block_before) {  // calls over very many tiles (a whole-genome shard): the per-block sums were scanned beforehand
    before = block_before[t0];
  } else {
    const uint4* n4 = reinterpret_cast<const uint4*>(tile_n);
    for (uint32_t i = threadIdx.x; i < t0 / 4u; i += 256u) {
      const uint4 v = __ldg(n4 + i);
      before += (unsigned long long)v.x + v.y + v.z + v.w;
    }
    for (int o = 16; o; o >>= 1) before += __shfl_xor_sync(0xFFFFFFFFu, before, o);
  }
  const uint32_t t = t0 + threadIdx.x;
  const uint32_t n = t < n_tiles ? tile_n[t] : 0u;
  uint32_t incl = n;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= (uint32_t)o) incl += u;
  }
  if (lane == 31u) warp_n[warp] = incl;
  if (lane == 0u) warp_sum[warp] = block_before ? (warp == 0u ? before : 0ull) : before;
  __syncthreads();
  unsigned long long dst = seg_base + (incl - n);
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    dst += warp_sum[w];
    if ((uint32_t)w < warp) dst += warp_n[w];
  }
  if (!n) return;
  const uint32_t src = tile_base[t];
  for (uint32_t i = 0; i < n; ++i) {
    if (dst + i >= cap_total || src + i >= cap_seg) break;
    dev_rec[dst + i] = rec[src + i];
  }
}

// coalesced copy of one segment's ordered records from the contiguous device copy to the pinned host block
__global__ void __launch_bounds__(256) k_rec_to_host(const unsigned long long* __restrict__ dev_rec, const unsigned long long* counters, uint32_t seg,
                                                     uint32_t cap_seg, unsigned long long* __restrict__ host_rec, unsigned long long cap_total) {
  unsigned long long base = 0;
  for (uint32_t j = 0; j < seg; ++j) base += min(counters[12 + j], (unsigned long long)cap_seg);
  const unsigned long long n = min(counters[12 + seg], (unsigned long long)cap_seg);
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n && base + i < cap_total; i += (unsigned long long)gridDim.x * blockDim.x)
    host_rec[base + i] = dev_rec[base + i];
}

// the exact kernel's general records and allele bytes, HBM -> the pinned block of the result in 16-byte stores.  (The exact
// kernel used to write them to host memory itself: one PCIe write per allele byte from a single lane, which is what a locus
// with a long insertion spent most of its time on, and which slowed k_rec_to_host next to it to half its speed.)
__global__ void __launch_bounds__(256) k_general_to_host(const unsigned long long* counters, const uint4* __restrict__ d_rec, const uint8_t* __restrict__ d_pool,
                                                         uint4* __restrict__ h_rec, uint8_t* __restrict__ h_pool, uint32_t cap_rec, uint32_t cap_pool) {
  static_assert(sizeof(guac_threshold_record) == 32, "two 16-byte stores per general record");
  static_assert(kPoolDynOff % 8 == 0, "the dynamic part of the pool is copied in 8-byte words");
  const unsigned long long n16 = 2ull * min(counters[0], (unsigned long long)cap_rec);
  const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x, nth = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = tid; i < n16; i += nth) h_rec[i] = d_rec[i];
  const unsigned long long pool_end = min((unsigned long long)kPoolDynOff + counters[1], (unsigned long long)cap_pool);
  const unsigned long long n8 = (pool_end - kPoolDynOff + 7ull) / 8ull;  // (both buffers are padded past cap_pool)
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(d_pool + kPoolDynOff);
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(h_pool + kPoolDynOff);
  for (unsigned long long i = tid; i < n8; i += nth) dst[i] = src[i];
}

// plain copy of one segment's records (no ordering asked for, or a kernel that does not note its tiles' slices)
__global__ void __launch_bounds__(256) k_rec_flush(const unsigned long long* __restrict__ rec, const unsigned long long* counters, uint32_t seg,
                                                   uint32_t cap_seg, unsigned long long* __restrict__ host_rec, unsigned long long* __restrict__ dev_rec,
                                                   unsigned long long cap_total) {
  unsigned long long base = 0;
  for (uint32_t j = 0; j < seg; ++j) base += min(counters[12 + j], (unsigned long long)cap_seg);
  const unsigned long long n = min(counters[12 + seg], (unsigned long long)cap_seg);
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    if (base + i >= cap_total) break;
    const unsigned long long v = rec[i];
    host_rec[base + i] = v;
    dev_rec[base + i] = v;
  }
}

// per-allele counts: one warp per requested locus of the ranges (prefix[i] = loci before range i)
__global__ void __launch_bounds__(kExactWarps * 32) k_allele_counts(DevReads R, const guac_locus_range* __restrict__ ranges,
                                                                    const unsigned long long* __restrict__ prefix, uint32_t n_ranges,
                                                                    CallParams prm, DevOut out) {
  __shared__ AlleleEntry tabs[kExactWarps][kMaxAlleles];
  __shared__ uint32_t rings[kExactWarps][64];
  const unsigned long long n_loci = prefix[n_ranges];
  const unsigned long long warp = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * (unsigned long long)blockDim.x) >> 5;
  for (unsigned long long t = warp; t < n_loci; t += n_warps) {
    uint32_t lo = 0, hi = n_ranges - 1;  // the range holding locus number t
    while (lo < hi) {
      const uint32_t mid = (lo + hi + 1) >> 1;
      if (prefix[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const int contig = ranges[lo].contig;
    const long long locus = ranges[lo].start + (long long)(t - prefix[lo]);
    if (locus < R.contigs[contig].length) exact_locus(R, contig, (int)locus, prm, out, tabs[threadIdx.x >> 5], rings[threadIdx.x >> 5]);
    __syncwarp();
  }
}

}  // namespace guac
