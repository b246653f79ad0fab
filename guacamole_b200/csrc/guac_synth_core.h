// guac_synth_core.h — the synthetic read generator of the benchmark shapes, shared by the host build
// (guac_synth.cpp: the batches the CPU oracle and the tests consume) and the device build (guac_synth_device.cuh: reads
// generated straight into HBM, per shard, for the whole-genome shape).  Integer arithmetic only, so that both builds produce
// the same bytes:
//   reference   counter-hashed i.i.d. ACGT, one 100-base N run per 100 kb (0.1 %); like an aligner, no read is placed on one
//   germline    one SNV per 1,000 loci (2/3 het, 1/3 hom), one 1-10 bp indel per 10,000 loci (het)
//   somatic     tumor sample only: one SNV per 100,000 loci at VAF U(0.1, 0.5)
//   read starts a Poisson process: the number of reads starting at locus p is Poisson(depth / read length), a pure function of
//               (seed, sample, contig, p) — what "uniform starts, then sorted" converges to, and it makes every read a pure
//               function of (seed, sample, contig, start, rank among the reads of that start): a shard of the genome can be
//               generated alone and holds exactly the reads a whole-genome run would place there
//   reads       fixed length; 78 % all-M, 20 % one soft clip of 5-50 bases, 0.9 % one insertion, 0.9 % one deletion,
//               0.2 % both; base quality 70 % Q37-41 / 20 % Q25-36 / 10 % Q2-24 with errors at 10^(-q/10);
//               MAPQ 90 % 60 / 10 % U{0..59}; strand 50/50
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define GS_HD __host__ __device__ __forceinline__
#else
#define GS_HD inline
#endif

namespace gsynth {

constexpr int kPad = 40;            // a read's reference span never exceeds read_length + kPad
constexpr int kMaxPoisson = 1024;   // entries of the Poisson table (depth / read length up to ~800 reads per locus)
constexpr int kMaxOps = 24;         // CIGAR operators of one read (upper bound)

GS_HD uint64_t mix(uint64_t x) {  // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
GS_HD uint64_t h2(uint64_t a, uint64_t b) { return mix(a ^ mix(b)); }
GS_HD uint64_t h3(uint64_t a, uint64_t b, uint64_t c) { return mix(a ^ mix(b ^ mix(c))); }

struct Rng {
  uint64_t s;
  GS_HD explicit Rng(uint64_t seed) : s(seed) {}
  GS_HD uint64_t next() { return s = mix(s); }
  GS_HD uint32_t u32() { return (uint32_t)(next() >> 32); }
  GS_HD uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};

// Everything that involves floating point is tabulated once on the host (guac_synth_tables) and used as integers.
struct Tables {
  uint64_t seed;
  int32_t sample;           // 1 = tumor carries the somatic SNVs
  int32_t read_length;
  uint32_t thr_clip, thr_ins, thr_del, thr_both;   // cumulative class thresholds over a 32-bit uniform
  uint32_t q_error[64];                            // P(error | quality q) * 2^32
  int32_t n_poisson;                               // entries used in poisson_cdf
  int32_t pad_;
  uint64_t poisson_cdf[kMaxPoisson];               // P(K <= k) * 2^53, non-decreasing; K = first k with u53 < cdf[k]
};

GS_HD int base_index(char r) { return r == 'A' ? 0 : r == 'C' ? 1 : r == 'G' ? 2 : 3; }
GS_HD char base_char(int i) { return i == 0 ? 'A' : i == 1 ? 'C' : i == 2 ? 'G' : 'T'; }

struct Indel {
  int64_t pos;
  int len;
  bool is_del;
  int hap;
};

struct Genome {
  uint64_t seed;
  int sample;
  GS_HD int64_t n_run_start(int contig, int64_t blk) const { return blk * 100000 + (int64_t)(h3(seed, 0x4E00 + (uint64_t)contig, (uint64_t)blk) % 99900); }
  GS_HD char ref(int contig, int64_t p) const {
    const int64_t n0 = n_run_start(contig, p / 100000);
    if (p >= n0 && p < n0 + 100) return 'N';
    return base_char((int)(h3(seed, 0x1000 + (uint64_t)contig, (uint64_t)p) & 3));
  }
  // germline indel of the 10 kb block: position = first deleted base / base before which bases are inserted
  GS_HD Indel indel(int contig, int64_t blk) const {
    const uint64_t h = h3(seed, 0x2000 + (uint64_t)contig, (uint64_t)blk);
    Indel d;
    d.pos = blk * 10000 + 50 + (int64_t)(h % 9900);
    d.len = 1 + (int)((h >> 20) % 10);
    d.is_del = ((h >> 40) & 1) != 0;
    d.hap = (int)((h >> 41) & 1);
    return d;
  }
  GS_HD bool near_indel(int contig, int64_t p) const {
    const Indel d = indel(contig, p / 10000);
    return p >= d.pos - 2 && p <= d.pos + d.len + 2;
  }
  // base carried by haplotype `hap` at p for a read whose somatic draw is u (tumor SNVs are carried when u < VAF * 2^32)
  GS_HD char hap_base(int contig, int64_t p, int hap, uint32_t u) const {
    const char r = ref(contig, p);
    if (r == 'N' || near_indel(contig, p)) return r;
    const int64_t blk = p / 1000;
    const uint64_t h = h3(seed, 0x3000 + (uint64_t)contig, (uint64_t)blk);
    if (p == blk * 1000 + (int64_t)(h % 1000)) {
      const bool hom = ((h >> 20) % 3) == 0;
      const int vh = (int)((h >> 24) & 1);
      if (hom || vh == hap) return base_char((base_index(r) + 1 + (int)((h >> 28) % 3)) & 3);
      return r;
    }
    if (sample == 1) {
      const int64_t sb = p / 100000;
      const uint64_t hs = h3(seed, 0x5000 + (uint64_t)contig, (uint64_t)sb);
      if (p == sb * 100000 + (int64_t)(hs % 100000)) {
        // VAF = 0.1 + 0.4 * t / 65536 as a 32-bit threshold
        const uint64_t t = (hs >> 20) & 0xFFFF;
        const uint32_t vaf = (uint32_t)(429496730ull + ((t * 1717986918ull) >> 16));
        if (u < vaf) return base_char((base_index(r) + 1 + (int)((hs >> 40) % 3)) & 3);
      }
    }
    return r;
  }
};

// May a read start at p?  It must stay inside the contig and clear of the reference's N runs (aligners place none there).
GS_HD bool start_allowed(const Genome& G, int contig, int64_t p, int64_t contig_length, int read_length) {
  const int64_t span = read_length + kPad;
  if (p < 0 || p + span > contig_length) return false;
  const int64_t blk0 = p / 100000, blk1 = (p + span) / 100000;
  for (int64_t blk = blk0; blk <= blk1; ++blk) {
    const int64_t n0 = G.n_run_start(contig, blk);
    if (p < n0 + 100 && p + span > n0) return false;
  }
  return true;
}

// number of reads starting at locus p: Poisson by inversion over the host-built table
GS_HD uint32_t reads_starting_at(const Tables& T, int contig, int64_t p) {
  const uint64_t u = h3(T.seed ^ 0xABCDEF12345ull, 0x8000 + (uint64_t)contig + 977ull * (uint64_t)T.sample, (uint64_t)p) >> 11;
  uint32_t k = 0;
  while ((int)k < T.n_poisson - 1 && u >= T.poisson_cdf[k]) ++k;
  return k;
}

GS_HD int draw_quality(Rng& rng) {
  const uint32_t t = rng.u32();
  if (t < 3006477107u) return 37 + (int)rng.below(5);    // 70 %
  if (t < 3865470566u) return 25 + (int)rng.below(12);   // 20 %
  return 2 + (int)rng.below(23);                          // 10 %
}

// One read.  Sink: op(op, len) — consecutive equal operators already merged —, base(b, q), md_char(c), md_number(n),
// finish(mapq, flags, ref_len).
template <class Sink>
GS_HD void make_read(const Tables& T, int contig, int64_t start, uint32_t rank, Sink& out) {
  const Genome G{T.seed, T.sample};
  Rng rng(h3(T.seed, 0x7000 + (uint64_t)T.sample + 16ull * (uint64_t)contig, (uint64_t)start * 4096ull + rank));
  const int L = T.read_length;
  const int hap = (int)(rng.next() & 1);
  const uint32_t u_som = rng.u32();
  // read class
  const uint32_t cls = rng.u32();
  int lead = 0, trail = 0;
  bool seq_ins = false, seq_del = false;
  if (cls < T.thr_clip) {
    int clip = 5 + (int)rng.below(46);
    clip = clip < L / 3 ? clip : L / 3;
    if (rng.next() & 1) lead = clip; else trail = clip;
  } else if (cls < T.thr_ins) seq_ins = true;
  else if (cls < T.thr_del) seq_del = true;
  else if (cls < T.thr_both) seq_ins = seq_del = true;
  const int target = L - lead - trail;  // read bases in the aligned part
  int ins_at = -1, del_at = -1;
  if (target > 40) {
    if (seq_ins) ins_at = 10 + (int)rng.below((uint32_t)(target - 30));
    if (seq_del) del_at = 10 + (int)rng.below((uint32_t)(target - 30));
    if (seq_ins && seq_del && (ins_at - del_at < 8 && del_at - ins_at < 8)) del_at = -1;
  }
  uint32_t cur_op = 0xFFu, cur_len = 0;
  auto push_op = [&](uint32_t op, uint32_t len) {
    if (len == 0) return;
    if (op == cur_op) { cur_len += len; return; }
    if (cur_len) out.op(cur_op, cur_len);
    cur_op = op;
    cur_len = len;
  };
  for (int i = 0; i < lead; ++i) { const int q = draw_quality(rng); out.base(base_char((int)rng.below(4)), (uint8_t)q); }
  push_op(4u /* S */, (uint32_t)lead);
  int64_t pos = start;
  int produced = 0;  // aligned-part read bases so far
  long match_run = 0;
  bool last_was_indel = true;  // no indel before the first aligned base
  while (produced < target) {
    const int remaining = target - produced;
    const Indel gi = G.indel(contig, pos / 10000);
    const bool can_indel = !last_was_indel && remaining > 6 && produced > 5;
    if (can_indel && gi.hap == hap && gi.pos == pos) {
      if (gi.is_del) {
        out.md_number(match_run);
        match_run = 0;
        out.md_char('^');
        for (int k = 0; k < gi.len; ++k) out.md_char(G.ref(contig, pos + k));
        push_op(2u /* D */, (uint32_t)gi.len);
        pos += gi.len;
        last_was_indel = true;
        continue;
      } else if (remaining > gi.len + 6) {
        for (int k = 0; k < gi.len; ++k) {
          const int q = draw_quality(rng);
          out.base(base_char((int)(h3(T.seed, 0x6000 + (uint64_t)contig, (uint64_t)(gi.pos * 16 + k)) & 3)), (uint8_t)q);
        }
        push_op(1u /* I */, (uint32_t)gi.len);
        produced += gi.len;
        last_was_indel = true;
        continue;
      }
    }
    if (can_indel && produced == ins_at) {
      const int n = 1 + (int)rng.below(3);
      if (remaining > n + 6) {
        for (int k = 0; k < n; ++k) { const int q = draw_quality(rng); out.base(base_char((int)rng.below(4)), (uint8_t)q); }
        push_op(1u, (uint32_t)n);
        produced += n;
        last_was_indel = true;
        ins_at = -1;
        continue;
      }
    }
    if (can_indel && produced == del_at) {
      const int n = 1 + (int)rng.below(3);
      out.md_number(match_run);
      match_run = 0;
      out.md_char('^');
      for (int k = 0; k < n; ++k) out.md_char(G.ref(contig, pos + k));
      push_op(2u, (uint32_t)n);
      pos += n;
      last_was_indel = true;
      del_at = -1;
      continue;
    }
    const char r = G.ref(contig, pos);
    char b = G.hap_base(contig, pos, hap, u_som);
    if (b == 'N') b = base_char((int)rng.below(4));
    const int q = draw_quality(rng);
    if (rng.u32() < T.q_error[q & 63]) b = base_char((base_index(b) + 1 + (int)rng.below(3)) & 3);
    out.base(b, (uint8_t)q);
    if (b == r) {
      ++match_run;
    } else {
      out.md_number(match_run);
      match_run = 0;
      out.md_char(r);
    }
    push_op(0u /* M */, 1u);
    ++pos;
    ++produced;
    last_was_indel = false;
  }
  out.md_number(match_run);
  for (int i = 0; i < trail; ++i) { const int q = draw_quality(rng); out.base(base_char((int)rng.below(4)), (uint8_t)q); }
  push_op(4u, (uint32_t)trail);
  if (cur_len) out.op(cur_op, cur_len);
  const uint8_t mapq = (uint8_t)(rng.u32() < 3865470566u ? 60 : rng.below(60));  // 90 % 60
  const uint8_t flags = (uint8_t)(0x08u /* GUAC_READ_HAS_MD */ | ((rng.next() & 1) ? 0x01u /* POSITIVE_STRAND */ : 0u));
  out.finish(mapq, flags, (int64_t)(pos - start));
}

GS_HD int decimal_digits(long n) {
  int d = 1;
  while (n >= 10) { n /= 10; ++d; }
  return d;
}

// sizes only
struct CountSink {
  uint32_t n_ops = 0, md_len = 0, n_bases = 0;
  int64_t ref_len = 0;
  GS_HD void op(uint32_t, uint32_t) { ++n_ops; }
  GS_HD void base(char, uint8_t) { ++n_bases; }
  GS_HD void md_char(char) { ++md_len; }
  GS_HD void md_number(long n) { md_len += (uint32_t)decimal_digits(n); }
  GS_HD void finish(uint8_t, uint8_t, int64_t r) { ref_len = r; }
};

// into the columns of a guac_read_batch
struct WriteSink {
  uint32_t* cigar;
  uint8_t* seq;
  uint8_t* qual;   // may be null
  char* md;
  uint8_t* mapq;
  uint8_t* flags;
  GS_HD void op(uint32_t o, uint32_t len) { *cigar++ = (len << 4) | o; }
  GS_HD void base(char b, uint8_t q) {
    *seq++ = (uint8_t)b;
    if (qual) *qual++ = q;
  }
  GS_HD void md_char(char c) { *md++ = c; }
  GS_HD void md_number(long n) {
    const int d = decimal_digits(n);
    for (int i = d - 1; i >= 0; --i) { md[i] = (char)('0' + n % 10); n /= 10; }
    md += d;
  }
  GS_HD void finish(uint8_t q, uint8_t f, int64_t) { *mapq = q; *flags = f; }
};

}  // namespace gsynth
