// guac_tile.cuh — the germline pileup path over per-granule difference streams.
//
// K_expand (pack time, once per read set): one WARP per granule of 1024 loci.  The CIGAR / MD expansion of the reference
//           (PileupElement.advanceToLocus / alignment, pileup/PileupElement.scala:68-248) done once, bit-parallel: every read
//           overlapping the granule slides its 2-bit base planes over the final reference track (funnel shift + XOR, 32 loci
//           per operation) and leaves
//             * one 16-bit entry per element that DIFFERS from the reference — (locus, class): class 1..3 = read base code ^
//               reference base code (a mismatch), class 0 = an element that is not a plain base (insertion / deletion anchor,
//               deleted or skipped locus, non-ACGT base), stored as the counter word and field K_call adds it to
//               (CntLayout::slot_code) — in the granule's contiguous difference stream;
//             * +1 in the per-locus "reads starting here" / "reads ending here" counters (all reads, positive strand), stored
//               as nibbles (wide stores: 16-bit fields): the depth of a locus is their running sum.
//           This is reference-based encoding, as CRAM does it, laid out by locus instead of by read.
// K_call   (k_call_tile, every call): one WARP per granule, LOCUS-centric, no per-read work at all: the difference stream is
//           replayed into the per-locus counter tile in shared memory (one atomic per difference, coalesced 16-byte loads),
//           then every lane scans the depth of its own 32 consecutive loci from the start / end nibbles (one warp scan per
//           granule) and, in the same pass, rejects clean loci four at a time, parks the few survivors and calls them side
//           by side:  GermlineThreshold.Caller.callVariantsAtLocus      commands/GermlineThresholdCaller.scala:90-179
//                     Pileup.depth / positiveDepth / referenceDepth       pileup/Pileup.scala:76-91
//           With GUAC_OPT_DIFFERENCE_LISTS = 0 the store is packed without streams and k_pileup_tile (guac_pileup.cuh) walks
//           planes and CIGARs inside the call instead: same results, the cross-check every parity test runs.
#pragma once

#include "guac_pack.cuh"
#include "guac_pileup.cuh"

namespace guac {

constexpr int kExpandWarps = 4;
constexpr int kStageCap = 2048;           // difference entries staged per granule in shared memory (deeper granules: second walk)
constexpr int kCntPad = 4;                // pad words per 32 loci: lane L's quad j sits at word 36 L + 4 j (16-byte aligned,
                                          // conflict-free for 16-byte loads by a quarter warp)
constexpr int kCntWords = kGranuleLoci + kCntPad * (kGranuleLoci / 32);

__device__ __forceinline__ int cnt_pos(int x) { return x + kCntPad * (x >> 5); }

// ---- counter tile layout shared by k_expand (which stores, per difference, the counter word and field it lands in) and k_call
template <bool WIDE>
struct CntLayout {
  static constexpr int kFieldBits = WIDE ? 16 : 8;
  static constexpr int kLociPerWord = WIDE ? 2 : 4;
  static constexpr int kLaneWords = (32 / kLociPerWord) + 4;          // a lane's words of one plane + 4 pad words
  static constexpr int kPlaneWords = 32 * kLaneWords;                 // 384 (narrow) / 640 (wide)
  static constexpr int kWords = 4 * kPlaneWords;                      // + one dummy word behind them for padding entries
  static constexpr int kSlotBits = WIDE ? 12 : 11;                    // bits of a word index; the field index sits above them
  static constexpr uint32_t kSlotMask = (1u << kSlotBits) - 1u;
  static constexpr uint32_t kPadCode = (uint32_t)kWords;              // entry that counts into the dummy word
  __host__ __device__ static constexpr int word_of(int x, int cls) {
    return cls * kPlaneWords + (x >> 5) * kLaneWords + ((x & 31) / kLociPerWord);
  }
  __host__ __device__ static constexpr uint32_t slot_code(int x, int cls) {
    return (uint32_t)word_of(x, cls) | ((uint32_t)(x & (kLociPerWord - 1)) << kSlotBits);
  }
  __device__ static __forceinline__ uint32_t field(const uint32_t* cnt, int x, int cls) {
    return (cnt[word_of(x, cls)] >> (kFieldBits * (x & (kLociPerWord - 1)))) & (WIDE ? 0xFFFFu : 0xFFu);
  }
};
static_assert(CntLayout<false>::kWords + 4 <= (1 << CntLayout<false>::kSlotBits), "narrow slot codes fit 11 bits");
static_assert(CntLayout<true>::kWords + 4 <= (1 << CntLayout<true>::kSlotBits), "wide slot codes fit 12 bits");

struct ExpandArgs {
  DevReads R;
  GranHdr* hdr_w;
  uint16_t* diffs_w;
  uint8_t* dd_w;
  uint8_t* dp_w;
  uint32_t* imp_w;
  unsigned long long cap_diffs;    // entries gs_diffs can hold
  uint32_t g_begin, g_end;         // granules (global index) this launch covers
  uint32_t n_contigs;
  int32_t wide;
  // counters: [2] difference entries reserved (multiple of 8 per granule), [3] a start / end count did not fit its field
  unsigned long long* counters;
  DevError* err;
};

template <bool HUGE>
struct ExpandSmem {
  // per locus: reads starting there (low half: all reads, high half: positive strand) and ending there, padded like cnt
  using Word = typename std::conditional<HUGE, unsigned long long, uint32_t>::type;
  Word st[kCntWords], en[kCntWords];
  uint32_t ref_lo[kWarpWords], ref_hi[kWarpWords], ref_std[kWarpWords];
  uint32_t imp[kWarpWords];  // loci with an "other" element that is not a mid-deletion element carrying the track's base
  alignas(16) uint16_t stage[kStageCap];
  uint32_t cursor;
  uint32_t pad_[3];
};

// HUGE: more than 65,535 reads over one granule (the two 16-bit halves of a counter word could carry): 32-bit halves.
template <bool HUGE>
__global__ void __launch_bounds__(kExpandWarps * 32) k_expand(ExpandArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using SM = ExpandSmem<HUGE>;
  using Word = typename SM::Word;
  constexpr int HB = HUGE ? 32 : 16;
  const int lane = threadIdx.x & 31;
  const DevReads& R = A.R;
  SM& S = reinterpret_cast<SM*>(smem_raw)[threadIdx.x >> 5];
  const uint32_t g = A.g_begin + blockIdx.x * kExpandWarps + (threadIdx.x >> 5);
  if (g >= A.g_end) return;  // whole warp
  // the contig of this granule: contigs are few, their granule offsets ascend
  uint32_t c = 0;
  {
    uint32_t lo = 0, hi = A.n_contigs - 1;
    while (lo < hi) {
      const uint32_t mid = (lo + hi + 1) >> 1;
      if (R.contigs[mid].gran_off <= g) lo = mid; else hi = mid - 1;
    }
    c = lo;  // (an empty contig shares its successor's offset: the last contig at or below g is the one that holds it)
  }
  const ContigInfo ci = R.contigs[c];
  const int tile_lo = (int)(g - ci.gran_off) << kGranuleShift;
  const int tile_hi = min(tile_lo + kGranuleLoci, ci.n_words << 5);
  for (int i = lane; i < kCntWords; i += 32) { S.st[i] = 0; S.en[i] = 0; }
  {
    const int w = (tile_lo >> 5) + lane;
    const bool in = w < ci.n_words;
    S.ref_lo[lane] = in ? R.trk_lo[ci.word_off + w] : 0u;
    S.ref_hi[lane] = in ? R.trk_hi[ci.word_off + w] : 0u;
    S.ref_std[lane] = in ? R.trk_std[ci.word_off + w] : 0u;
    S.imp[lane] = 0u;
  }
  if (lane == 0) S.cursor = 0;
  uint32_t first = R.gran_first[g], last = R.gran_last[g];
  if (first == 0xFFFFFFFFu) first = last = 0;
  __syncwarp();

  uint16_t* direct = nullptr;  // second walk of a granule whose entries did not fit the stage: straight to global memory
  const bool wide = A.wide != 0;
  auto append = [&](int x, uint32_t cls) {  // (divergent-safe)
    const uint32_t slot = atomicAdd(&S.cursor, 1u);
    const uint16_t e = (uint16_t)(wide ? CntLayout<true>::slot_code(x, (int)cls) : CntLayout<false>::slot_code(x, (int)cls));
    if (direct) direct[slot] = e;
    else if (slot < (uint32_t)kStageCap) S.stage[slot] = e;
  };
  auto impure = [&](int x) {  // (idempotent: the second walk of a deep granule sets the same bits)
    const uint32_t bit = 1u << (x & 31);
    if (!(S.imp[x >> 5] & bit)) atomicOr(&S.imp[x >> 5], bit);
  };
  auto other = [&](int lo, int hi) {  // reference positions [lo, hi) of this lane's read hold elements that are not plain bases
    lo = max(lo, tile_lo);
    hi = min(hi, tile_hi);
    for (int p = lo; p < hi; ++p) {
      append(p - tile_lo, 0u);
      impure(p - tile_lo);
    }
  };
  // deleted loci [lo, hi) of read r: MidDeletion elements, allele (deleted base, "") (PileupElement.scala:121-123).  Where the
  // deleted base of the MD tag is the track's base, all such elements of the locus are ONE allele and k_call_tile decides the
  // locus from its counters; anything else (no cached tag offset, a non-standard or disagreeing base) marks the locus impure.
  auto deleted = [&](uint32_t r, int lo, int hi) {
    lo = max(lo, tile_lo);
    hi = min(hi, tile_hi);
    if (lo >= hi) return;
    const int d0 = R.del_start[r], dn = (int)R.del_len[r];
    const char* md = R.md + R.md_off[r] + R.del_md[r];
    for (int p = lo; p < hi; ++p) {
      const int x = p - tile_lo;
      append(x, 0u);
      bool same = false;
      if (d0 >= 0 && p >= d0 && p < d0 + dn) {
        const uint8_t ch = (uint8_t)md[p - d0];
        const uint32_t w = (uint32_t)x >> 5, b = (uint32_t)x & 31u;
        same = is_std_base(ch) && ((S.ref_std[w] >> b) & 1u) &&
               base_code(ch) == (((S.ref_lo[w] >> b) & 1u) | (((S.ref_hi[w] >> b) & 1u) << 1));
      }
      if (!same) impure(x);
    }
  };
  // one plain M/=/X run [seg_ref, seg_ref + seg_len) whose first base is read base seg_read
  auto segment = [&](const ReadRec& rec, int seg_ref, int seg_read, int seg_len, bool has_exc) {
    const int s = max(seg_ref, tile_lo), e = min(seg_ref + seg_len, tile_hi);
    if (s >= e) return;
    const uint2* __restrict__ P = R.pairs + rec.pair_off;
    const uint32_t* __restrict__ X = R.xmask + rec.pair_off;
    const int w0 = (s - tile_lo) >> 5, w1 = (e - 1 - tile_lo) >> 5;
    int q0 = seg_read + (tile_lo + (w0 << 5) - seg_ref);  // read base under bit 0 of word w0 (> -32)
    uint2 pa = make_uint2(0u, 0u);
    uint32_t xa = 0;
    if ((q0 >> 5) >= 0) {
      pa = __ldg(P + (q0 >> 5));
      if (has_exc) xa = __ldg(X + (q0 >> 5));
    }
    for (int w = w0; w <= w1; ++w, q0 += 32) {
      const int j = q0 >> 5, sh = q0 & 31;  // arithmetic shift: floor
      const uint2 pb = __ldg(P + j + 1);
      const uint32_t xb = has_exc ? __ldg(X + j + 1) : 0u;
      const int wbase = tile_lo + (w << 5);
      uint32_t valid = bit_range(seg_ref - wbase, seg_ref + seg_len - wbase);
      const uint32_t oth = __funnelshift_r(xa, xb, sh) & valid;  // non-ACGT bases
      valid &= ~oth;
      const uint32_t std_m = S.ref_std[w];  // mismatches over a reference base that is not A/C/G/T are left out: the callers
      const uint32_t x = (__funnelshift_r(pa.x, pb.x, sh) ^ S.ref_lo[w]) & valid & std_m;  // hand those loci to the exact path
      const uint32_t y = (__funnelshift_r(pa.y, pb.y, sh) ^ S.ref_hi[w]) & valid & std_m;
      pa = pb;
      xa = xb;
      uint32_t d = x | y | oth;
      if (oth && (S.imp[w] & oth) != oth) atomicOr(&S.imp[w], oth);  // a non-ACGT read base is its own allele
      while (d) {
        const int b = __ffs(d) - 1;
        d &= d - 1;
        const uint32_t cls = ((oth >> b) & 1u) ? 0u : (((x >> b) & 1u) | (((y >> b) & 1u) << 1));
        append((w << 5) + b, cls);
      }
    }
  };
  // the CIGAR walk (PileupElement.scala:68-135): every M/=/X run is a segment; insertion / deletion anchors, deleted and
  // skipped loci are "other" elements
  auto walk = [&](uint32_t r, const ReadRec& rec) {
    const bool has_exc = (rec.info & kInfoHasExc) != 0;
    if (rec.info & kInfoSimple) {
      segment(rec, rec.start, (int)(rec.info & kInfoLeadMask), rec.end - rec.start, has_exc);
      return;
    }
    int ref_pos = rec.start, read_pos = 0;
    bool skip_first = false;  // contig-start insertion: the element at locus 0 is the insertion, not a plain base
    const uint32_t c0 = R.cig_off[r], c1 = R.cig_off[r + 1];
    for (uint32_t k = c0; k < c1 && ref_pos < tile_hi; ++k) {
      const uint32_t v = R.cigar[k];
      const uint32_t op = v & 0xF, next_op = (k + 1 < c1) ? (R.cigar[k + 1] & 0xF) : 0xFFu;
      const int len = (int)(v >> 4);
      if (op_is_match_like(op)) {
        int seg_ref = ref_pos, seg_read = read_pos, seg_len = len;
        if (skip_first) {
          other(ref_pos, ref_pos + 1);
          ++seg_ref; ++seg_read; --seg_len;
          skip_first = false;
        }
        // (M|=, I) and (M|=|X, D): the run's last base is the insertion / deletion anchor (PileupElement.scala:93, 109)
        const bool anchor = (next_op == GUAC_CIGAR_I && (op == GUAC_CIGAR_M || op == GUAC_CIGAR_EQ)) || next_op == GUAC_CIGAR_D;
        if (anchor && seg_len > 0) {
          other(ref_pos + len - 1, ref_pos + len);
          --seg_len;
        }
        segment(rec, seg_ref, seg_read, seg_len, has_exc);
        ref_pos += len;
        read_pos += len;
      } else if (op == GUAC_CIGAR_D) {
        deleted(r, ref_pos, ref_pos + len);
        ref_pos += len;
      } else if (op == GUAC_CIGAR_N) {
        other(ref_pos, ref_pos + len);  // skipped loci
        ref_pos += len;
      } else if (op == GUAC_CIGAR_I) {
        if (ref_pos == 0 && rec.start == 0) skip_first = true;
        read_pos += len;
      } else if (op == GUAC_CIGAR_S) {
        read_pos += len;
      }
    }
  };

  const uint16_t pad_code = (uint16_t)(wide ? CntLayout<true>::kPadCode : CntLayout<false>::kPadCode);  // counts into the dummy word
  uint32_t depth_in = 0, pos_in = 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (uint32_t base = first; base < last; base += 32) {
      const uint32_t r = base + lane;
      ReadRec rec{0, 0, 0, 0};
      if (r < last) rec = R.rec[r];
      const bool active = r < last && rec.end > tile_lo && rec.start < tile_hi && rec.end > rec.start;
      if (!active) continue;
      if (pass == 0) {
        const Word one = (Word)1 | ((rec.info & kInfoPositive) ? ((Word)1 << HB) : (Word)0);
        if (rec.start < tile_lo) {
          depth_in += 1;
          pos_in += (rec.info & kInfoPositive) ? 1u : 0u;
        } else {
          atomicAdd(&S.st[cnt_pos(rec.start - tile_lo)], one);
        }
        if (rec.end < tile_lo + kGranuleLoci) atomicAdd(&S.en[cnt_pos(rec.end - tile_lo)], one);
      }
      walk(r, rec);
    }
    __syncwarp();
    if (pass == 1) break;
    // reserve the granule's slice of the stream (a multiple of 8 entries = 16 bytes)
    const uint32_t total = S.cursor;
    const uint32_t padded = (total + 7u) & ~7u;
    unsigned long long goff = 0;
    if (lane == 0) goff = atomicAdd(&A.counters[2], (unsigned long long)padded);
    goff = __shfl_sync(0xFFFFFFFFu, goff, 0);
    for (int o = 16; o; o >>= 1) {
      depth_in += __shfl_xor_sync(0xFFFFFFFFu, depth_in, o);
      pos_in += __shfl_xor_sync(0xFFFFFFFFu, pos_in, o);
    }
    if (lane == 0) A.hdr_w[g] = GranHdr{(uint32_t)(goff >> 3), total, depth_in, pos_in};
    const bool fits_global = goff + padded <= A.cap_diffs;  // (else the host grows the buffer and packs the streams again)
    if (!fits_global || total == 0) break;
    uint16_t* dst = A.diffs_w + goff;
    if (total <= (uint32_t)kStageCap) {
      for (uint32_t i = total + lane; i < padded; i += 32) S.stage[i] = pad_code;
      __syncwarp();
      const uint4* src4 = reinterpret_cast<const uint4*>(S.stage);
      uint4* dst4 = reinterpret_cast<uint4*>(dst);
      for (uint32_t i = lane; i < padded / 8; i += 32) dst4[i] = src4[i];
      break;
    }
    for (uint32_t i = total + lane; i < padded; i += 32) dst[i] = pad_code;
    direct = dst;
    if (lane == 0) S.cursor = 0;
    __syncwarp();
  }
  __syncwarp();

  // ---- start / end counts of the lane's 32 loci -> nibbles (or 16-bit fields), coalesced 16-byte stores
  constexpr Word HMASK = ((Word)1 << HB) - 1;
  bool overflow = false;
  const size_t locus0 = (size_t)g * kGranuleLoci + (size_t)lane * 32;
  A.imp_w[(size_t)g * kWarpWords + lane] = S.imp[lane];
  if (!wide) {
    uint32_t wd[8], wp[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t vd = 0, vp = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int x = lane * 32 + j * 4 + k;
        const Word s = S.st[cnt_pos(x)], e = S.en[cnt_pos(x)];
        const uint32_t sa = (uint32_t)(s & HMASK), sp = (uint32_t)(s >> HB), ea = (uint32_t)(e & HMASK), ep = (uint32_t)(e >> HB);
        overflow = overflow || sa > 15u || ea > 15u;
        vd |= ((sa & 15u) | ((ea & 15u) << 4)) << (8 * k);
        vp |= ((sp & 15u) | ((ep & 15u) << 4)) << (8 * k);
      }
      wd[j] = vd;
      wp[j] = vp;
    }
    uint4* dd4 = reinterpret_cast<uint4*>(A.dd_w + locus0);
    uint4* dp4 = reinterpret_cast<uint4*>(A.dp_w + locus0);
    dd4[0] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    dd4[1] = make_uint4(wd[4], wd[5], wd[6], wd[7]);
    dp4[0] = make_uint4(wp[0], wp[1], wp[2], wp[3]);
    dp4[1] = make_uint4(wp[4], wp[5], wp[6], wp[7]);
  } else {
    uint4* dd4 = reinterpret_cast<uint4*>(A.dd_w + locus0 * 4);
    uint4* dp4 = reinterpret_cast<uint4*>(A.dp_w + locus0 * 4);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t vd[4], vp[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int x = lane * 32 + j * 4 + k;
        const Word s = S.st[cnt_pos(x)], e = S.en[cnt_pos(x)];
        const uint32_t sa = (uint32_t)(s & HMASK), sp = (uint32_t)(s >> HB), ea = (uint32_t)(e & HMASK), ep = (uint32_t)(e >> HB);
        overflow = overflow || sa > 0xFFFFu || ea > 0xFFFFu;
        vd[k] = (sa & 0xFFFFu) | (ea << 16);
        vp[k] = (sp & 0xFFFFu) | (ep << 16);
      }
      dd4[j] = make_uint4(vd[0], vd[1], vd[2], vd[3]);
      dp4[j] = make_uint4(vp[0], vp[1], vp[2], vp[3]);
    }
  }
  if (overflow) atomicAdd(&A.counters[3], 1ull);
}

// ---- K_call ---------------------------------------------------------------------------------------------------------------------
// Counter tile of one granule: four planes, one per counter class — [0] "other" elements, [1..3] mismatches by class — so that
// four consecutive loci of one class share a 32-bit word (narrow stores: 8-bit fields, pileups < 256 deep) and the sum over the
// classes of four loci is three 32-bit additions.  Wide stores: 16-bit fields, two loci per word (< 65,536 deep).  Every update
// is a native 32-bit shared-memory atomic.  A lane owns 32 consecutive loci = 8 (16) consecutive words of every plane; 4 pad
// words per lane keep its 16-byte loads conflict-free.  The difference stream stores, per entry, the counter word and the
// field inside it (slot_code): the replay is one shift and one atomic per entry.
template <bool WIDE>
struct CallSmem {
  alignas(16) uint32_t cnt[CntLayout<WIDE>::kWords + 4];  // (+ the dummy word the stream's padding entries point at)
  EmitStage stage;
};

// the per-locus work past the cheap reject: the counts row (counts mode) or callVariantsAtLocus on the SNV alleles
template <bool WIDE, int MODE>
__device__ __forceinline__ void tile_call_locus(const uint32_t* cnt, const DevReads& R, const TileDesc& td, const CallParams& prm,
                                                DevOut& out, const int x, const int total, const int pos_total, const bool every_covered,
                                                const bool all_loci, const bool pure, EmitStage* stage) {
  if (total == 0 && !all_loci) return;  // callVariantsAtLocus returns nothing on an empty pileup
  using L = CntLayout<WIDE>;
  const int o = (int)L::field(cnt, x, 0), m1 = (int)L::field(cnt, x, 1), m2 = (int)L::field(cnt, x, 2), m3 = (int)L::field(cnt, x, 3);
  const int w = x >> 5, b = x & 31;
  uint32_t wl = 0, wh = 0, ws = 0;
  if (w < td.n_words) { wl = R.trk_lo[td.trk_word + w]; wh = R.trk_hi[td.trk_word + w]; ws = R.trk_std[td.trk_word + w]; }
  const bool std_ref = (ws >> b) & 1u;
  const int rcode = (int)(((wl >> b) & 1u) | (((wh >> b) & 1u) << 1));
  const int locus = (td.word0 << 5) + x;
  if (MODE == 1) {
    if (!std_ref && total > 0) {
      defer_locus(out, td.contig, locus, stage);
      return;
    }
    const uint32_t s = (uint32_t)atomicAdd(&out.counters[0], 1ull);
    if (s < out.cap_rec) {
      guac_locus_counts gc;
      gc.locus = locus;
      gc.contig = td.contig;
      gc.depth = total;
      gc.positive_depth = pos_total;
      gc.reference_depth = std_ref ? total - o - m1 - m2 - m3 : 0;
      gc.base_count[rcode] = total - o - m1 - m2 - m3;
      gc.base_count[rcode ^ 1] = m1;
      gc.base_count[rcode ^ 2] = m2;
      gc.base_count[rcode ^ 3] = m3;
      gc.other_count = o;
      gc.reference_base = std_ref ? code_base(rcode) : (uint8_t)'N';
      gc.pad_[0] = gc.pad_[1] = gc.pad_[2] = 0;
      out.crec[s] = gc;
    }
    return;
  }
  call_snv_locus(prm, out, td.contig, locus, total, o, m1, m2, m3, rcode, std_ref, every_covered, pure, stage);
}

__device__ __forceinline__ uint32_t pick8(const uint32_t (&v)[8], int i) {  // v[i] without dynamic register indexing
  uint32_t r = v[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) r = i == k ? v[k] : r;
  return r;
}

template <bool WIDE, int MODE>
__global__ void __launch_bounds__(kTileThreads, 8) k_call_tile(DevReads R, const TileDesc* __restrict__ tiles, uint32_t n_tiles, CallParams prm, DevOut out) {
  using L = CntLayout<WIDE>;
  constexpr uint32_t FMASK = WIDE ? 0xFFFFu : 0xFFu;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const uint32_t tile = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (tile >= n_tiles) return;  // whole warp
  CallSmem<WIDE>& S = reinterpret_cast<CallSmem<WIDE>*>(smem_raw)[threadIdx.x >> 5];
  const TileDesc td = tiles[tile];
  const int tile_lo = td.word0 << 5;
  const uint32_t g = td.gran;
  const GranHdr hdr = R.gs_hdr[g];

  // ---- phase 0: clear the counter tile; the lane's own track word and start / end fields are requested right away
  {
    uint4* c4 = reinterpret_cast<uint4*>(S.cnt);
    for (int i = lane; i < (int)(sizeof(S.cnt) / 16); i += 32) c4[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (lane == 0) { S.stage.n_rec = 0; S.stage.n_slow = 0; }
  uint32_t my_std = 0;  // the lane owns loci [32 lane, 32 lane + 32) of the granule = one track word
  if (lane < td.n_words) my_std = R.trk_std[td.trk_word + lane];
  constexpr uint32_t kOnes = 0x01010101u;
  const size_t locus0 = (size_t)g * kGranuleLoci + (size_t)lane * 32;
  // narrow stores: the lane's 32 start / end bytes (and, in counts mode, the positive-strand ones) stay in registers;
  // wide stores (deep pileups) read their 32-bit fields again where they are needed
  uint32_t dd[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const uint32_t* dd_wide = reinterpret_cast<const uint32_t*>(R.gs_dd) + locus0;
  const uint32_t* dp_wide = reinterpret_cast<const uint32_t*>(R.gs_dp) + locus0;
  int run = 0, prun = 0;
  if (WIDE) {
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(dd_wide) + j);
      run += (int)((v.x & 0xFFFFu) + (v.y & 0xFFFFu) + (v.z & 0xFFFFu) + (v.w & 0xFFFFu)) - (int)((v.x >> 16) + (v.y >> 16) + (v.z >> 16) + (v.w >> 16));
      if (MODE == 1) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(dp_wide) + j);
        prun += (int)((u.x & 0xFFFFu) + (u.y & 0xFFFFu) + (u.z & 0xFFFFu) + (u.w & 0xFFFFu)) - (int)((u.x >> 16) + (u.y >> 16) + (u.z >> 16) + (u.w >> 16));
      }
    }
  } else {
    const uint4* p = reinterpret_cast<const uint4*>(R.gs_dd + locus0);
    const uint4 v0 = __ldg(p), v1 = __ldg(p + 1);
    dd[0] = v0.x; dd[1] = v0.y; dd[2] = v0.z; dd[3] = v0.w; dd[4] = v1.x; dd[5] = v1.y; dd[6] = v1.z; dd[7] = v1.w;
    if (MODE == 1) {
      const uint4* q = reinterpret_cast<const uint4*>(R.gs_dp + locus0);
      const uint4 u0 = __ldg(q), u1 = __ldg(q + 1);
      dp[0] = u0.x; dp[1] = u0.y; dp[2] = u0.z; dp[3] = u0.w; dp[4] = u1.x; dp[5] = u1.y; dp[6] = u1.z; dp[7] = u1.w;
    }
  }
  __syncwarp();

  // ---- phase 1: replay the granule's difference stream into the counter tile (8 entries per 16-byte load)
  {
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(R.gs_diffs) + hdr.df_off;
    const uint32_t n16 = (hdr.n_df + 7u) >> 3;
    for (uint32_t i = lane; i < n16; i += 32) {
      const uint4 v = __ldg(src + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t e = (k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xFFFFu);
        atomicAdd(&S.cnt[e & L::kSlotMask], 1u << ((e >> L::kSlotBits) * L::kFieldBits));
      }
    }
  }

  // ---- phase 2 + 3 fused: depth scan of the lane's 32 loci and, in the same pass, the caller
  const bool all_loci = MODE == 1 ? !prm.skip_empty : false;          // rows for empty pileups (counts mode only)
  const bool every_covered = MODE == 1 || prm.emit_ref || prm.emit_no_call;
  const bool dense = every_covered || all_loci;
  const uint32_t thr_plus_1 = (uint32_t)(prm.threshold_percent + 1);
  if (!WIDE) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // nibble sums of four loci by one dot product each
      run += (int)__dp4a(dd[j] & 0x0F0F0F0Fu, kOnes, 0u) - (int)__dp4a((dd[j] >> 4) & 0x0F0F0F0Fu, kOnes, 0u);
      if (MODE == 1) prun += (int)__dp4a(dp[j] & 0x0F0F0F0Fu, kOnes, 0u) - (int)__dp4a((dp[j] >> 4) & 0x0F0F0F0Fu, kOnes, 0u);
    }
  }
  int dep, pdep;
  {
    int incl = run, pincl = prun;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += t;
      if (MODE == 1) {
        const int u = __shfl_up_sync(0xFFFFFFFFu, pincl, o);
        if (lane >= o) pincl += u;
      }
    }
    dep = (int)hdr.depth_in + incl - run;
    pdep = (int)hdr.pos_in + pincl - prun;
  }
  __syncwarp();  // (the counter tile is complete)
  const int l0 = tile_lo + (lane << 5);
  const uint32_t in_range = bit_range(td.locus_begin - l0, td.locus_end - l0);
  uint32_t n_visited = 0;
  bool overflow = false;
  // the lane's word of the impure mask, read only where a locus gets that far (nothing of it is kept live across the scan)
  auto imp_word = [&]() { return __ldg(R.gs_imp + (size_t)__ldg(&tiles[tile].gran) * kWarpWords + lane); };
  if (WIDE || dense) {
    const uint32_t my_imp = MODE == 0 ? imp_word() : 0xFFFFFFFFu;
    // the general loop, one locus at a time: deep pileups (wide stores) and the dense outputs (counts, emit-ref / emit-no-call)
#pragma unroll 1
    for (int kk = 0; kk < 32; ++kk) {
      uint32_t s, e, ps = 0, pe = 0;
      if (WIDE) {
        const uint32_t v = __ldg(dd_wide + kk);
        s = v & 0xFFFFu; e = v >> 16;
        if (MODE == 1) { const uint32_t u = __ldg(dp_wide + kk); ps = u & 0xFFFFu; pe = u >> 16; }
      } else {
        const uint32_t v = pick8(dd, kk >> 2) >> (8 * (kk & 3));
        s = v & 15u; e = (v >> 4) & 15u;
        if (MODE == 1) { const uint32_t u = pick8(dp, kk >> 2) >> (8 * (kk & 3)); ps = u & 15u; pe = (u >> 4) & 15u; }
      }
      dep += (int)s - (int)e;
      pdep += (int)ps - (int)pe;
      if (!((in_range >> kk) & 1u)) continue;
      if (dep != 0 || !prm.skip_empty) ++n_visited;
      const int x = (lane << 5) + kk;
      const uint32_t differing = L::field(S.cnt, x, 0) + L::field(S.cnt, x, 1) + L::field(S.cnt, x, 2) + L::field(S.cnt, x, 3);
      if (differing != 0u && (uint32_t)dep > FMASK) overflow = true;  // a counter field may have wrapped: the host widens the store
      const bool std_ref = (my_std >> kk) & 1u;
      if (!dense && differing == 0u && std_ref) continue;  // every element matches the reference: nothing to call
      if (!dense && std_ref && (unsigned long long)differing * 100ull < (unsigned long long)thr_plus_1 * (unsigned long long)(uint32_t)dep) continue;
      tile_call_locus<WIDE, MODE>(S.cnt, R, td, prm, out, x, dep, pdep, every_covered, all_loci, !((my_imp >> kk) & 1u), &S.stage);
    }
  } else {
    // sparse calls over 8-bit counter fields, four consecutive loci per step.  The few loci that survive the reject are
    // remembered (one bit per locus of the lane) and called afterwards, the lanes that hold one side by side.
    const int dep_lane = dep;
    uint32_t survivors = 0;
    uint4 pl[4][2];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      pl[c][0] = *reinterpret_cast<const uint4*>(&S.cnt[c * L::kPlaneWords + L::kLaneWords * lane]);
      pl[c][1] = *reinterpret_cast<const uint4*>(&S.cnt[c * L::kPlaneWords + L::kLaneWords * lane + 4]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      auto word = [&](int c) -> uint32_t {
        const uint4 v = pl[c][j >> 2];
        return (j & 3) == 0 ? v.x : (j & 3) == 1 ? v.y : (j & 3) == 2 ? v.z : v.w;
      };
      // per-locus sums of the four classes, byte-wise (no carry between loci while the depth is below 256)
      const uint32_t w0 = word(0), w1 = word(1), w2 = word(2), w3 = word(3);
      const uint32_t sum4 = w0 + w1 + w2 + w3;
      const uint32_t st_sum = __dp4a(dd[j] & 0x0F0F0F0Fu, kOnes, 0u), en_sum = __dp4a((dd[j] >> 4) & 0x0F0F0F0Fu, kOnes, 0u);
      const uint32_t std4 = (my_std >> (4 * j)) & 0xFu, inr4 = (in_range >> (4 * j)) & 0xFu;
      // The whole quadruple at once: every locus stays covered by at least dmin reads, so an allele seen fewer than
      // tq = ceil((threshold + 1) * dmin / 100) times cannot pass the threshold there.  Byte-wise "count >= tq" over the
      // three mismatch classes and the "other" elements (the only alleles besides the reference one): a locus with such a
      // count is remembered — its exact depth and the full rule follow below —, every other locus yields nothing.
      const uint32_t dmin = (uint32_t)dep - en_sum;
      const uint32_t tq = (thr_plus_1 * dmin + 99u) / 100u;
      if (std4 == 0xFu && inr4 == 0xFu && (uint32_t)dep > en_sum && (uint32_t)dep + st_sum <= 255u && tq - 1u < 128u) {
        const uint32_t bias = (128u - tq) * kOnes;
        const uint32_t hit = ((((w0 & 0x7F7F7F7Fu) + bias) | w0) | (((w1 & 0x7F7F7F7Fu) + bias) | w1) | (((w2 & 0x7F7F7F7Fu) + bias) | w2) |
                              (((w3 & 0x7F7F7F7Fu) + bias) | w3)) & 0x80808080u;
        // bits 7, 15, 23, 31 -> bits 0..3
        if (hit) survivors |= ((hit * 0x00204081u) >> 28) << (4 * j);
        dep += (int)st_sum - (int)en_sum;
        n_visited += 4;
        continue;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // one locus at a time, exactly
        dep += (int)((dd[j] >> (8 * k)) & 15u) - (int)((dd[j] >> (8 * k + 4)) & 15u);
        if (!((inr4 >> k) & 1u)) continue;
        if (dep != 0 || !prm.skip_empty) ++n_visited;
        const uint32_t differing = (sum4 >> (8 * k)) & 0xFFu;
        const bool std_ref = (std4 >> k) & 1u;
        if (differing != 0u && (uint32_t)dep > FMASK) overflow = true;
        if (differing == 0u && std_ref) continue;
        if (std_ref && differing * 100u < thr_plus_1 * (uint32_t)dep) continue;
        survivors |= 1u << (4 * j + k);
      }
    }
    const uint32_t my_imp = survivors ? imp_word() : 0u;
    while (survivors) {
      const int kk = __ffs(survivors) - 1;
      survivors &= survivors - 1;
      int d = dep_lane;  // the depth at locus kk again: the lane's start / end nibbles up to and including kk
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t m = j < (kk >> 2) ? 0x0F0F0F0Fu : j == (kk >> 2) ? (0x0F0F0F0Fu >> (8 * (3 - (kk & 3)))) : 0u;
        d += (int)__dp4a(dd[j] & m, kOnes, 0u) - (int)__dp4a((dd[j] >> 4) & m, kOnes, 0u);
      }
      tile_call_locus<WIDE, MODE>(S.cnt, R, td, prm, out, (lane << 5) + kk, d, 0, every_covered, all_loci, !((my_imp >> kk) & 1u), &S.stage);
    }
  }
  flush_stage(out, &S.stage, tile);
  // one atomic per warp for the visited-loci counter
  n_visited = __reduce_add_sync(0xFFFFFFFFu, n_visited);
  if (lane == 0 && n_visited) atomicAdd(&out.counters[3], (unsigned long long)n_visited);
  if (overflow) atomicAdd(&out.counters[5], 1ull);
}

}  // namespace guac
