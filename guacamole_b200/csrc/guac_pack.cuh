// guac_pack.cuh — device side of guac_reads_pack: raw host columns (already copied to HBM) -> the packed read store.
//
// Replaces, for the hot path, the lazy per-read work the reference does on the JVM:
//   MDTagUtils.getReference / MappedRead.mdTagReferenceBases   (reads/MDTagUtils.scala:23-78, reads/MappedRead.scala:57-76)
//   Pileup.referenceBaseAtLocus                                 (pileup/Pileup.scala:157-165)
//   windowTaskFlatMapMultipleRDDs' read -> task expansion       (DistributedUtil.scala:585-597) -> granule index
#pragma once

#include "guac_device.cuh"

namespace guac {

struct PackArgs {
  DevReads R;            // const view
  uint64_t r_begin, r_end;  // reads this launch handles (k_pack_bases, k_md_track: launched per chunk of the host -> device copy)
  ReadRec* rec_w;        // writable aliases
  uint2* pairs_w;
  uint32_t* xmask_w;
  uint8_t* qc_w;         // (quality | base code << 6) per base, nullptr when qualities are not packed
  uint16_t* nm_w;
  int32_t* del_start_w;
  uint32_t* del_md_w;
  uint16_t* del_len_w;
  char* md_w;
  uint32_t* trk_lo_w;
  uint32_t* trk_hi_w;
  uint32_t* trk_std_w;
  uint32_t* conflict_w;  // per track word: loci where reads' MD-derived reference bases disagree
  uint32_t* gran_first_w;
  uint32_t* gran_last_w;
  uint32_t* gran_count_w;
  const uint32_t* read_contig;  // [n] contig index per read (scratch)
  DevError* err;
  unsigned long long* counters;  // [0] order-sensitive loci resolved, [1] max reads per granule
};

// ---- K_header: thread per read; the per-read checks and derived columns (what SlidingWindow / MappedRead do lazily) ------------
// Validates sortedness (windowing/SlidingWindow.scala:56, DistributedUtil.scala:662-664), CIGAR / read-length consistency
// (reads/MappedRead.scala:87, pileup/PileupElement.scala:106) and the presence of MD tags; derives [start, end), the SIMPLE
// flag + leading clip, the number of 32-base plane words and the per-contig read ranges.  The first failing read (lowest
// index) wins: errors are folded with atomicMin over (read index << 8 | status).
struct HeaderArgs {
  uint64_t n;
  uint32_t n_contigs;
  const int32_t* contig;
  const int64_t* start;
  const uint64_t* cigar_off;
  const uint32_t* cigar;
  const uint64_t* seq_off;
  const uint8_t* mapq;
  const uint8_t* flags;
  const int32_t* sample;            // may be null
  const uint64_t* md_off;
  const int64_t* contig_length;     // device copy, or null: no upper bound checked
  ReadRec* rec;                     // pair_off is filled in after the scan of n_pairs
  uint32_t* cig_off32;
  uint32_t* md_off32;
  uint32_t* read_contig;
  uint32_t* n_pairs;
  unsigned long long* contig_first; // [n_contigs], initialised to ~0
  unsigned long long* contig_last;  // [n_contigs], 0
  long long* contig_end;            // [n_contigs], 0: largest read end
  unsigned long long* summary;      // [0] first error (~0 = none), [1] longest reference span, [2] sample of read 0
  uint32_t* mapq_mask;              // [8]: bit m set = some read has mapping quality m (sizes the somatic kernel's table)
};

__global__ void __launch_bounds__(256) k_header(HeaderArgs H) {
  const int lane = threadIdx.x & 31;
  uint32_t mask_sent[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // (lane 0: bits of mapq_mask this warp has already set)
  // Largest end per contig and longest span: running maxima of the warp over its iterations, sent when the contig changes and
  // at the end (one same-address atomic per warp and iteration was most of this kernel's 1.5 ms on a chr20 read set)
  unsigned span_sent = 0, run_end = 0;
  int32_t run_contig = -1;
  for (uint64_t i0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) & ~31ull; i0 < H.n; i0 += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = i0 + lane;
    int status = 0;
    int64_t end = 0, ref_len = 0;
    int32_t c = 0;
    if (i < H.n) {
      c = H.contig[i];
      const int64_t start = H.start[i];
      const int32_t sample0 = H.sample ? H.sample[0] : 0;
      if (i == 0) H.summary[2] = (unsigned long long)(uint32_t)sample0;
      const uint64_t c0 = H.cigar_off[i], c1 = H.cigar_off[i + 1];
      const uint64_t read_len = H.seq_off[i + 1] - H.seq_off[i];
      const bool in_range = c >= 0 && (uint32_t)c < H.n_contigs;
      const bool run_start = i == 0 || H.contig[i - 1] != c;
      if (!in_range) status = GUAC_ERR_INVALID_ARGUMENT;
      else if (!run_start && start < H.start[i - 1]) status = GUAC_ERR_UNSORTED_READS;
      else if (start < 0) status = GUAC_ERR_INVALID_ARGUMENT;
      else if (H.sample && H.sample[i] != sample0) status = GUAC_ERR_UNSUPPORTED;
      else if (!(H.flags[i] & GUAC_READ_HAS_MD)) status = GUAC_ERR_MISSING_MD;
      else if (read_len > (uint64_t)kMaxReadLen) status = GUAC_ERR_UNSUPPORTED;
      int64_t consumed = 0, lead = 0;
      int phase = 0;  // 0 leading clips, 1 inside the aligned run, 2 trailing clips
      bool simple = true;
      for (uint64_t k = c0; k < c1 && status == 0; ++k) {
        const uint32_t op = H.cigar[k] & 0xF, len = H.cigar[k] >> 4;
        if (op > 8 || op == GUAC_CIGAR_P || len == 0) { status = GUAC_ERR_INVALID_CIGAR; break; }
        const bool m = op == GUAC_CIGAR_M || op == GUAC_CIGAR_EQ || op == GUAC_CIGAR_X;
        if (m || op == GUAC_CIGAR_D || op == GUAC_CIGAR_N) ref_len += len;
        if (m || op == GUAC_CIGAR_I || op == GUAC_CIGAR_S) consumed += len;
        if (m) {
          if (phase == 2) simple = false;
          phase = 1;
        } else if (op == GUAC_CIGAR_S || op == GUAC_CIGAR_H) {
          if (phase == 0) {
            if (op == GUAC_CIGAR_S) lead += len;
          } else
            phase = 2;
        } else {
          simple = false;
        }
      }
      if (c1 == c0) simple = false;
      end = start + ref_len;
      if (status == 0) {
        if ((uint64_t)consumed != read_len) status = GUAC_ERR_INVALID_CIGAR;
        else if (end > 0x7FFFFF00ll) status = GUAC_ERR_UNSUPPORTED;
        else if (H.contig_length && end > H.contig_length[c]) status = GUAC_ERR_INVALID_ARGUMENT;
      }
      if (status == 0) {
        if (simple && lead > 0xFFFF) simple = false;
        const uint32_t info = (uint32_t)(simple ? (lead & 0xFFFF) : 0) | (simple ? kInfoSimple : 0) |
                              ((H.flags[i] & GUAC_READ_POSITIVE_STRAND) ? kInfoPositive : 0) | (ref_len == 0 ? kInfoEmpty : 0) |
                              ((uint32_t)H.mapq[i] << kInfoMapqShift);
        H.rec[i] = ReadRec{(int32_t)start, (int32_t)end, 0u, info};
        H.cig_off32[i] = (uint32_t)c0;
        H.md_off32[i] = (uint32_t)H.md_off[i];
        H.read_contig[i] = (uint32_t)c;
        H.n_pairs[i] = (uint32_t)((read_len + 31) / 32);
        if (run_start) {  // "Regions are not sorted by contig": a contig's reads must form one run
          if (atomicExch(&H.contig_first[c], (unsigned long long)i) != ~0ull) status = GUAC_ERR_CONTIG_ORDER;
          if (i > 0) {
            const int32_t pc = H.contig[i - 1];
            if (pc >= 0 && (uint32_t)pc < H.n_contigs) H.contig_last[pc] = i;
          }
        }
        if (i + 1 == H.n) H.contig_last[c] = H.n;
      } else {
        H.n_pairs[i] = 0;
      }
      if (status) atomicMin(&H.summary[0], ((unsigned long long)i << 8) | (unsigned long long)status);
    }
    {  // mapping qualities present: one OR per warp and 32-bit word, only for bits the warp has not set before
      const uint32_t mq = i < H.n ? (uint32_t)H.mapq[i] : 0xFFFFFFFFu;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const uint32_t bits = __reduce_or_sync(0xFFFFFFFFu, (mq >> 5) == (uint32_t)w ? (1u << (mq & 31u)) : 0u);
        if (lane == 0 && (bits & ~mask_sent[w])) {
          atomicOr(&H.mapq_mask[w], bits);
          mask_sent[w] |= bits;
        }
      }
    }
    // largest end per contig / longest span: one atomic per warp when its reads share a contig (they are sorted by contig)
    const bool ok = i < H.n && status == 0;
    const int32_t c_lead = __shfl_sync(0xFFFFFFFFu, c, 0);
    const bool uniform = __all_sync(0xFFFFFFFFu, !ok || c == c_lead) && __shfl_sync(0xFFFFFFFFu, (int)ok, 0);
    const unsigned span = ok ? (unsigned)ref_len : 0u;
    const unsigned max_span = __reduce_max_sync(0xFFFFFFFFu, span);
    if (lane == 0 && max_span > span_sent) {
      atomicMax(&H.summary[1], (unsigned long long)max_span);
      span_sent = max_span;
    }
    if (uniform) {
      const unsigned max_end = __reduce_max_sync(0xFFFFFFFFu, ok ? (unsigned)end : 0u);
      if (lane == 0) {
        if (run_contig != c_lead) {
          if (run_contig >= 0) atomicMax(&H.contig_end[run_contig], (long long)run_end);
          run_contig = c_lead;
          run_end = 0;
        }
        run_end = max(run_end, max_end);
      }
    } else if (ok) {
      atomicMax(&H.contig_end[c], (long long)end);
    }
  }
  if (lane == 0 && run_contig >= 0) atomicMax(&H.contig_end[run_contig], (long long)run_end);
}

// rec[i].pair_off from the scan of n_pairs; rec[n] = the sentinel
__global__ void __launch_bounds__(256) k_header_finish(ReadRec* __restrict__ rec, const uint32_t* __restrict__ pair_off, uint32_t* __restrict__ cig_off32,
                                                       uint32_t* __restrict__ md_off32, const uint64_t* __restrict__ cigar_off, const uint64_t* __restrict__ md_off,
                                                       uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i <= n; i += (uint64_t)gridDim.x * blockDim.x) {
    if (i < n) rec[i].pair_off = pair_off[i];
    else {
      rec[n] = ReadRec{0x7FFFFFFF, 0x7FFFFFFF, pair_off[n], 0u};
      cig_off32[n] = (uint32_t)cigar_off[n];
      md_off32[n] = (uint32_t)md_off[n];
    }
  }
}

// ---- K_pack_bases: ASCII bases -> (lo, hi) bit-plane pairs + non-ACGT mask (+ the qc bytes) -------------------------------
// A CTA takes 64 consecutive reads at a time.  Their bases (and qualities) are one contiguous byte range of the raw columns:
// it is staged into shared memory with one TMA bulk copy per column (cp.async.bulk, completion on an mbarrier).  Then
//   pass A (qualities packed): the qc bytes are a function of (base, quality) alone, so the staged range is converted 16
//     bytes per thread, SIMD within 32-bit words, and stored with aligned 16-byte stores;
//   pass B: eight lanes per read, one 32-base word per lane and step: nine aligned shared-memory words realigned with byte
//     permutes, then per 4 bases the bit of interest of every byte is gathered into a nibble with one multiply
//     (hi = bit 2 of the ASCII code, lo = bit 1 ^ bit 2: A=0 C=1 G=2 T=3), non-ACGT bytes found with byte-wise compares.
constexpr int kPackReads = 64;
constexpr int kPackStageBytes = 20 * 1024;
constexpr int kPackLanesPerRead = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {  // 16 B aligned, 16 B multiple
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(phase)
                 : "memory");
  } while (!done);
}

// bit i of the result = bit 0 of byte i of t (t holds 0 / 1 bytes)
__device__ __forceinline__ uint32_t gather_nibble(uint32_t t) { return (t * 0x01020408u) >> 24; }

__device__ __forceinline__ uint32_t std_bytes_mask(uint32_t w) {  // 0xFF in every byte that is 'A', 'C', 'G' or 'T'
  return __vcmpeq4(w, 0x41414141u) | __vcmpeq4(w, 0x43434343u) | __vcmpeq4(w, 0x47474747u) | __vcmpeq4(w, 0x54545454u);
}

// the read (index in [r0, r1)) that owns byte b of the raw columns
__device__ __forceinline__ uint64_t read_of_byte(const uint64_t* __restrict__ seq_off, uint64_t r0, uint64_t r1, uint64_t b) {
  uint64_t lo = r0, hi = r1 - 1;
  while (lo < hi) {
    const uint64_t mid = (lo + hi + 1) >> 1;
    if (seq_off[mid] <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) k_pack_bases(PackArgs A) {
  __shared__ __align__(128) uint8_t s_seq[kPackStageBytes + 64];   // (pass B reads whole words a little past the last read)
  __shared__ __align__(128) uint8_t s_qual[kPackStageBytes + 64];
  __shared__ __align__(8) uint64_t bar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  uint32_t phase = 0;
  const bool have_qual = A.R.qual != nullptr;
  for (uint64_t r0 = A.r_begin + (uint64_t)blockIdx.x * kPackReads; r0 < A.r_end; r0 += (uint64_t)gridDim.x * kPackReads) {
    const uint64_t r1 = min(r0 + (uint64_t)kPackReads, A.r_end);
    const uint64_t b0 = A.R.seq_off[r0], b1 = A.R.seq_off[r1];
    const uint64_t a0 = b0 & ~(uint64_t)15;                      // align the source down to 16 bytes
    const uint32_t bytes = (uint32_t)(((b1 - a0) + 15) & ~(uint64_t)15);
    const bool staged = b1 > b0 && (b1 - a0) + 16 <= (uint64_t)kPackStageBytes;  // (the raw columns carry 64 bytes of padding)
    if (staged) {
      if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, have_qual ? 2 * bytes : bytes);
        bulk_copy_g2s(s_seq, A.R.seq + a0, bytes, &bar);
        if (have_qual) bulk_copy_g2s(s_qual, A.R.qual + a0, bytes, &bar);
      }
      mbar_wait(&bar, phase);
      phase ^= 1u;
      // ---- pass A: qc = (quality & 63) | base code << 6, 16 bytes per thread.  The 16-byte chunks at both ends also hold
      // bytes of the neighbouring blocks of reads: those CTAs store the very same values there.
      if (have_qual) {
        for (uint32_t c = threadIdx.x * 16u; c < bytes; c += 256u * 16u) {
          const uint4 sv = *reinterpret_cast<const uint4*>(s_seq + c), qv = *reinterpret_cast<const uint4*>(s_qual + c);
          const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w}, qw[4] = {qv.x, qv.y, qv.z, qv.w};
          uint32_t o[4], wide = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t hi = (sw[k] >> 2) & 0x01010101u, lo = ((sw[k] >> 1) ^ (sw[k] >> 2)) & 0x01010101u;
            o[k] = (qw[k] & 0x3F3F3F3Fu) | (lo << 6) | (hi << 7);
            wide |= qw[k] & 0xC0C0C0C0u;
          }
          *reinterpret_cast<uint4*>(A.qc_w + a0 + c) = make_uint4(o[0], o[1], o[2], o[3]);
          if (wide) {  // rare: a quality > 63 (the read leaves the one-byte path) or > 127 (rejected)
            for (uint32_t i = 0; i < 16; ++i) {
              const uint64_t bpos = a0 + c + i;
              const uint32_t q = s_qual[c + i];
              if (q > 63u && bpos >= b0 && bpos < b1) {
                const uint64_t r = read_of_byte(A.R.seq_off, r0, r1, bpos);
                atomicOr(&A.rec_w[r].info, kInfoWideQ);
                if (q > 127u) report_error(A.err, GUAC_ERR_BAD_QUALITY, r);
              }
            }
          }
        }
      }
      // ---- pass B: planes, eight lanes per read
      const int grp = lane / kPackLanesPerRead, gl = lane % kPackLanesPerRead;
      for (uint64_t r = r0 + (uint64_t)warp * (32 / kPackLanesPerRead) + grp; r < r1; r += 8 * (32 / kPackLanesPerRead)) {
        const uint64_t s0 = A.R.seq_off[r];
        const int len = (int)(A.R.seq_off[r + 1] - s0);
        const uint32_t p0 = A.R.rec[r].pair_off;
        const uint32_t off = (uint32_t)(s0 - a0);
        bool any_exc = false;
        for (int wd = gl; wd * 32 < len; wd += kPackLanesPerRead) {
          const uint32_t byte0 = off + (uint32_t)wd * 32u;          // first byte of this word in the stage
          const uint32_t* __restrict__ src = reinterpret_cast<const uint32_t*>(s_seq + (byte0 & ~3u));
          const uint32_t sel = 0x3210u + 0x1111u * (byte0 & 3u);    // byte permute: 4 bytes starting at (byte0 & 3)
          const int nb = min(32, len - wd * 32);                    // bases of the read in this word
          uint32_t lo = 0, hi = 0, x = 0;
          uint32_t prev = src[0];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t next = src[k + 1];
            const uint32_t w4 = __byte_perm(prev, next, sel);
            prev = next;
            const uint32_t h = (w4 >> 2) & 0x01010101u, l = ((w4 >> 1) ^ (w4 >> 2)) & 0x01010101u;
            const uint32_t ok = std_bytes_mask(w4) & 0x01010101u;
            hi |= gather_nibble(h & ok) << (4 * k);
            lo |= gather_nibble(l & ok) << (4 * k);
            x |= gather_nibble(ok ^ 0x01010101u) << (4 * k);
          }
          const uint32_t valid = nb >= 32 ? 0xFFFFFFFFu : ((1u << nb) - 1u);
          lo &= valid; hi &= valid; x &= valid;
          A.pairs_w[p0 + wd] = make_uint2(lo, hi);
          A.xmask_w[p0 + wd] = x;
          any_exc = any_exc || x != 0u;
        }
        if (any_exc) atomicOr(&A.rec_w[r].info, kInfoHasExc);
        // upper-case the MD tag in place (ADAM MdTag upper-cases before parsing)
        const uint32_t m0 = A.R.md_off[r], m1 = A.R.md_off[r + 1];
        for (uint32_t i = m0 + gl; i < m1; i += kPackLanesPerRead) {
          const char ch = A.md_w[i];
          if (ch >= 'a' && ch <= 'z') A.md_w[i] = (char)(ch - 32);
        }
      }
    } else {
      // ---- a block of reads larger than the stage (long reads): warp per read straight from global memory
      for (uint64_t r = r0 + warp; r < r1; r += 8) {
        const uint64_t s0 = A.R.seq_off[r], s1 = A.R.seq_off[r + 1];
        const int len = (int)(s1 - s0);
        const uint32_t p0 = A.R.rec[r].pair_off;
        uint32_t any_exc = 0;
        bool bad_q = false, wide_q = false;
        for (int base = 0; base < len; base += 32) {
          const int i = base + lane;
          uint8_t b = 'A';
          uint32_t q = 0;
          if (i < len) {
            b = A.R.seq[s0 + i];
            if (have_qual) q = A.R.qual[s0 + i];
          }
          bad_q = bad_q || q > 127;
          wide_q = wide_q || q > 63;
          const uint32_t code = base_code(b);
          if (have_qual && i < len) A.qc_w[s0 + i] = (uint8_t)((q & 63u) | (code << 6));
          const bool exc = i < len && !is_std_base(b);
          const uint32_t lo = __ballot_sync(0xFFFFFFFFu, (code & 1u) && !exc);
          const uint32_t hi = __ballot_sync(0xFFFFFFFFu, (code & 2u) && !exc);
          const uint32_t x = __ballot_sync(0xFFFFFFFFu, exc);
          any_exc |= x;
          if (lane == 0) {
            A.pairs_w[p0 + (base >> 5)] = make_uint2(lo, hi);
            A.xmask_w[p0 + (base >> 5)] = x;
          }
        }
        if (__any_sync(0xFFFFFFFFu, bad_q) && lane == 0) report_error(A.err, GUAC_ERR_BAD_QUALITY, r);
        const bool any_wide = __any_sync(0xFFFFFFFFu, wide_q);
        if (lane == 0 && (any_exc || any_wide)) atomicOr(&A.rec_w[r].info, (any_exc ? kInfoHasExc : 0u) | (any_wide ? kInfoWideQ : 0u));
        const uint32_t m0 = A.R.md_off[r], m1 = A.R.md_off[r + 1];
        for (uint32_t i = m0 + lane; i < m1; i += 32) {
          const char c = A.md_w[i];
          if (c >= 'a' && c <= 'z') A.md_w[i] = (char)(c - 32);
        }
      }
    }
    __syncthreads();  // the stage buffers are reused by the next block of reads
  }
}

// ---- K_granule_index: thread per read; which reads can overlap each 1024-loci granule ---------------------------------
__global__ void __launch_bounds__(256) k_granule_index(PackArgs A) {
  for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < A.R.n; r += (uint64_t)gridDim.x * blockDim.x) {
    const ReadRec rec = A.R.rec[r];
    if (rec.end <= rec.start) continue;
    const ContigInfo ci = A.R.contigs[A.read_contig[r]];
    int g0 = rec.start >> kGranuleShift, g1 = (rec.end - 1) >> kGranuleShift;
    for (int g = g0; g <= g1; ++g) {
      atomicMin(&A.gran_first_w[ci.gran_off + g], (uint32_t)r);
      atomicMax(&A.gran_last_w[ci.gran_off + g], (uint32_t)r + 1u);
      atomicAdd(&A.gran_count_w[ci.gran_off + g], 1u);
    }
  }
}

__global__ void k_granule_max(PackArgs A, uint32_t n_grans) {
  uint32_t m = 0;
  for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n_grans; g += gridDim.x * blockDim.x) m = max(m, A.gran_count_w[g]);
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(&A.counters[1], (unsigned long long)m);
}

// ---- K_md_track: thread per read; MD-derived reference bases -> the per-contig reference track ---------------------------
// Each bit of a locus' 2-bit base code is OR-merged into one of TWO planes — "some read says 1" (trk_lo / trk_hi) and "some read
// says 0" (kept in trk_std / conflict until k_track_finish) — so that one walk both builds the track and finds the loci where
// the reads' MD tags disagree (a bit that is set in both planes of a pair); k_track_finish derives the final planes.
struct TrackVisitor {
  const PackArgs& A;
  const ContigInfo& ci;
  uint64_t r;
  uint32_t pair_off;
  bool has_exc;
  int n_mismatch = 0;
  int del_first = -1, del_md_pos = 0, del_n = 0;
  __device__ TrackVisitor(const PackArgs& a, const ContigInfo& c, uint64_t read, uint32_t po, bool exc)
      : A(a), ci(c), r(read), pair_off(po), has_exc(exc) {}

  __device__ __forceinline__ void put_base(int ref_pos, uint8_t ch) {
    if (ref_pos < 0 || ref_pos >= ci.length || !is_std_base(ch)) return;
    const uint32_t w = ci.word_off + (uint32_t)(ref_pos >> 5), bit = 1u << (ref_pos & 31);
    const uint32_t code = base_code(ch);
    uint32_t* lo_plane = (code & 1u) ? A.trk_lo_w : A.trk_std_w;
    uint32_t* hi_plane = (code & 2u) ? A.trk_hi_w : A.conflict_w;
    if (!(lo_plane[w] & bit)) atomicOr(&lo_plane[w], bit);
    if (!(hi_plane[w] & bit)) atomicOr(&hi_plane[w], bit);
  }
  __device__ bool run(int ref_pos, int read_pos, int k) {
    // k aligned bases whose reference base equals the read base: word-parallel merge of the read's planes
    const uint2* __restrict__ P = A.R.pairs + pair_off;
    const uint32_t* __restrict__ X = A.R.xmask + pair_off;
    const int w0 = max(ref_pos >> 5, 0), w1 = min((ref_pos + k - 1) >> 5, ci.n_words - 1);
    int q0 = read_pos + ((w0 << 5) - ref_pos);  // read base under bit 0 of word w0
    uint2 pa = make_uint2(0u, 0u);
    uint32_t xa = 0;
    if (w0 <= w1 && (q0 >> 5) >= 0) {  // the rolling window: one 8-byte plane-pair load per word
      pa = P[q0 >> 5];
      if (has_exc) xa = X[q0 >> 5];
    }
    for (int w = w0; w <= w1; ++w, q0 += 32) {
      const int j = q0 >> 5, sh = q0 & 31;  // arithmetic shift: floor
      const uint2 pb = P[j + 1];
      const uint32_t xb = has_exc ? X[j + 1] : 0u;
      const int wbase = w << 5;
      uint32_t bits = bit_range(ref_pos - wbase, ref_pos + k - wbase) & ~__funnelshift_r(xa, xb, sh);
      if (wbase + 32 > ci.length) bits &= bit_range(0, ci.length - wbase);
      const uint32_t lo = __funnelshift_r(pa.x, pb.x, sh) & bits, hi = __funnelshift_r(pa.y, pb.y, sh) & bits;
      const uint32_t lo0 = bits & ~lo, hi0 = bits & ~hi;
      pa = pb;
      xa = xb;
      const uint32_t gw = ci.word_off + (uint32_t)w;
      if ((A.trk_lo_w[gw] & lo) != lo) atomicOr(&A.trk_lo_w[gw], lo);
      if ((A.trk_hi_w[gw] & hi) != hi) atomicOr(&A.trk_hi_w[gw], hi);
      if ((A.trk_std_w[gw] & lo0) != lo0) atomicOr(&A.trk_std_w[gw], lo0);
      if ((A.conflict_w[gw] & hi0) != hi0) atomicOr(&A.conflict_w[gw], hi0);
    }
    return true;
  }
  __device__ bool mismatch(int ref_pos, int, uint8_t ch) {
    ++n_mismatch;
    put_base(ref_pos, ch);
    return true;
  }
  __device__ bool deleted(int ref_pos, uint8_t ch, int md_pos) {
    put_base(ref_pos, ch);
    // remember the read's first deletion for the exact per-locus path
    if (del_first < 0) { del_first = ref_pos; del_md_pos = md_pos; del_n = 1; }
    else if (ref_pos == del_first + del_n && md_pos == del_md_pos + del_n) ++del_n;
    return true;
  }
  __device__ bool skipped(int, int) { return true; }
};

__global__ void __launch_bounds__(128) k_md_track(PackArgs A) {
  for (uint64_t r = A.r_begin + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < A.r_end; r += (uint64_t)gridDim.x * blockDim.x) {
    const ContigInfo ci = A.R.contigs[A.read_contig[r]];
    const ReadRec rec = A.R.rec[r];
    TrackVisitor v(A, ci, r, rec.pair_off, (rec.info & kInfoHasExc) != 0);
    int rc = md_walk(A.R, r, v);
    if (rc) report_error(A.err, rc, r);
    A.nm_w[r] = (uint16_t)min(v.n_mismatch, 65535);
    A.del_start_w[r] = v.del_first;
    A.del_md_w[r] = (uint32_t)v.del_md_pos;
    A.del_len_w[r] = (uint16_t)min(v.del_n, 65535);
  }
}

// thread per track word: the "says 0" planes become the standard-base plane and the plane of disagreeing loci
__global__ void __launch_bounds__(256) k_track_finish(PackArgs A, uint32_t w_begin, uint32_t w_end) {
  for (uint32_t w = w_begin + blockIdx.x * blockDim.x + threadIdx.x; w < w_end; w += gridDim.x * blockDim.x) {
    const uint32_t lo1 = A.trk_lo_w[w], hi1 = A.trk_hi_w[w], lo0 = A.trk_std_w[w], hi0 = A.conflict_w[w];
    A.trk_std_w[w] = lo1 | lo0;
    A.conflict_w[w] = (lo1 & lo0) | (hi1 & hi0);
  }
}

// MD-derived reference base of one read at one locus (MappedRead.getReferenceBaseAtLocus); 0 if the read has none there
struct BaseAtVisitor {
  const DevReads& R;
  uint64_t r;
  int locus;
  uint8_t out = 0;
  __device__ BaseAtVisitor(const DevReads& rr, uint64_t read, int l) : R(rr), r(read), locus(l) {}
  __device__ bool run(int ref_pos, int read_pos, int k) {
    if (locus >= ref_pos && locus < ref_pos + k) {
      out = R.seq[R.seq_off[r] + read_pos + (locus - ref_pos)];
      return false;
    }
    return locus >= ref_pos + k;
  }
  __device__ bool mismatch(int ref_pos, int, uint8_t ch) {
    if (ref_pos == locus) {
      out = ch;
      return false;
    }
    return locus > ref_pos;
  }
  __device__ bool deleted(int ref_pos, uint8_t ch, int) {
    if (ref_pos == locus) {
      out = ch;
      return false;
    }
    return locus > ref_pos;
  }
  __device__ bool skipped(int ref_pos, int len) {
    if (locus >= ref_pos && locus < ref_pos + len) {
      out = 'N';
      return false;
    }
    return locus >= ref_pos + len;
  }
};

// ---- K_resolve_conflicts: thread per track word.  Canonical rule for loci where the reads' MD tags disagree (SURVEY
// H1a): the standard base given by the overlapping read with the smallest end, earliest read on ties — what
// Pileup.referenceBaseAtLocus sees first in the sliding window's heap order on a single-task run. ---------------------
__global__ void __launch_bounds__(128) k_resolve_conflicts(PackArgs A, uint32_t n_contigs, uint32_t w_begin, uint32_t w_end) {  // global words [w_begin, w_end)
  for (uint32_t c = 0; c < n_contigs; ++c) {
    const ContigInfo ci = A.R.contigs[c];
    if (ci.word_off >= w_end || ci.word_off + (uint32_t)ci.n_words <= w_begin) continue;
    const int w_first = (int)(max(ci.word_off, w_begin) - ci.word_off), w_last = (int)(min(ci.word_off + (uint32_t)ci.n_words, w_end) - ci.word_off);
    for (int w = w_first + blockIdx.x * blockDim.x + threadIdx.x; w < w_last; w += gridDim.x * blockDim.x) {
      uint32_t conf = A.conflict_w[ci.word_off + w];
      if (!conf) continue;
      uint32_t lo = A.trk_lo_w[ci.word_off + w], hi = A.trk_hi_w[ci.word_off + w];
      while (conf) {
        int b = __ffs(conf) - 1;
        conf &= conf - 1;
        int locus = (w << 5) + b;
        int g = locus >> kGranuleShift;
        uint32_t first = A.gran_first_w[ci.gran_off + g], last = A.gran_last_w[ci.gran_off + g];
        int best_end = 0x7FFFFFFF;
        uint8_t best = 0;
        for (uint32_t r = first; r < last && first != 0xFFFFFFFFu; ++r) {
          const ReadRec rec = A.R.rec[r];
          if (rec.start > locus || rec.end <= locus || rec.end >= best_end) continue;
          BaseAtVisitor v(A.R, r, locus);
          md_walk(A.R, r, v);
          if (is_std_base(v.out)) {
            best_end = rec.end;
            best = v.out;
          }
        }
        if (best) {
          uint32_t code = base_code(best), bit = 1u << b;
          lo = (lo & ~bit) | ((code & 1u) ? bit : 0u);
          hi = (hi & ~bit) | ((code & 2u) ? bit : 0u);
        }
        atomicAdd(&A.counters[0], 1ull);
      }
      A.trk_lo_w[ci.word_off + w] = lo;
      A.trk_hi_w[ci.word_off + w] = hi;
    }
  }
}

// ---- K_micro_counts: partitionLociByApproximateDepth step (2) (DistributedUtil.scala:182-195) ---------------------------------
// thread per read of one contig: +1 for every micro partition that has a locus inside the read's [start, end).  `ranges` =
// the micro-partition ranges of this contig sorted by start (disjoint, their micro index `task` non-decreasing); a micro
// partition split into several ranges by gaps in the loci counts a read once (LociMap.getAll returns a Set).
__global__ void __launch_bounds__(256) k_micro_counts(DevReads R, uint64_t r_begin, uint64_t r_end, const guac_locus_range* __restrict__ ranges,
                                                      uint32_t n_ranges, unsigned long long* __restrict__ counts) {
  for (uint64_t r = r_begin + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < r_end; r += (uint64_t)gridDim.x * blockDim.x) {
    const ReadRec rec = R.rec[r];
    if (rec.end <= rec.start) continue;
    uint32_t lo = 0, hi = n_ranges;  // first range that ends past the read's start
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (ranges[mid].end > (int64_t)rec.start) hi = mid; else lo = mid + 1;
    }
    int32_t last = -1;
    for (uint32_t k = lo; k < n_ranges && ranges[k].start < (int64_t)rec.end; ++k) {
      if (ranges[k].task != last) atomicAdd(&counts[ranges[k].task], 1ull);
      last = ranges[k].task;
    }
  }
}

// ---- K_fasta_track: FASTA mode (ReferenceGenome.getReferenceBase, DistributedUtil.scala:266): track = given bases ----------
__global__ void __launch_bounds__(256) k_fasta_track(PackArgs A, uint32_t n_contigs) {
  for (uint32_t c = 0; c < n_contigs; ++c) {
    const ContigInfo ci = A.R.contigs[c];
    const uint8_t* f = A.R.fasta + A.R.fasta_off[c];
    const int64_t flen = (int64_t)(A.R.fasta_off[c + 1] - A.R.fasta_off[c]);
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < ci.n_words; w += gridDim.x * blockDim.x) {
      uint32_t lo = 0, hi = 0, st = 0;
      for (int b = 0; b < 32; ++b) {
        int64_t x = ((int64_t)w << 5) + b;
        if (x >= flen || x >= ci.length) break;
        uint8_t ch = f[x];
        if (is_std_base(ch)) {
          uint32_t code = base_code(ch);
          lo |= (code & 1u) << b;
          hi |= ((code >> 1) & 1u) << b;
          st |= 1u << b;
        }
      }
      A.trk_lo_w[ci.word_off + w] = lo;
      A.trk_hi_w[ci.word_off + w] = hi;
      A.trk_std_w[ci.word_off + w] = st;
    }
  }
}

}  // namespace guac
