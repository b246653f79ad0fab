// guac_synth.cpp — deterministic synthetic read generator for the benchmark shapes of BASELINE.json (host code, no CUDA).
//
// Produces a guac_read_batch (include/guac.h) of start-sorted mapped reads with consistent CIGAR + MD tags:
//   reference   counter-hashed i.i.d. ACGT, one 100-base N run per 100 kb (0.1 %); like an aligner, no read is placed on one
//   germline    one SNV per 1,000 loci (2/3 het, 1/3 hom), one 1-10 bp indel per 10,000 loci (het)
//   somatic     tumor sample only: one SNV per 100,000 loci at VAF U(0.1, 0.5)
//   reads       fixed length; 78 % all-M, 20 % one soft clip of 5-50 bases, 0.9 % one insertion, 0.9 % one deletion,
//               0.2 % both; base quality 70 % Q37-41 / 20 % Q25-36 / 10 % Q2-24 with errors at 10^(-q/10);
//               MAPQ 90 % 60 / 10 % U{0..59}; strand 50/50                                    
// Every read is a pure function of (seed, read index), so the output does not depend on the thread count.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/guac.h"
#include "../../include/guac_synth.h"

namespace {

inline uint64_t mix(uint64_t x) {  // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline uint64_t h2(uint64_t a, uint64_t b) { return mix(a ^ mix(b)); }
inline uint64_t h3(uint64_t a, uint64_t b, uint64_t c) { return mix(a ^ mix(b ^ mix(c))); }

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() { return s = mix(s); }
  double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  uint32_t below(uint32_t n) { return (uint32_t)((next() >> 32) * (uint64_t)n >> 32); }
};

const char kBases[4] = {'A', 'C', 'G', 'T'};

struct Genome {
  uint64_t seed;
  int sample;  // 1 = tumor carries the somatic SNVs
  char ref(int contig, int64_t p) const {
    int64_t blk = p / 100000;
    int64_t n0 = (int64_t)(h3(seed, 0x4E00 + contig, (uint64_t)blk) % 99900);
    int64_t off = p - blk * 100000;
    if (off >= n0 && off < n0 + 100) return 'N';
    return kBases[h3(seed, 0x1000 + contig, (uint64_t)p) & 3];
  }
  // germline indel of the 10 kb block: position v (first deleted base / base before which bases are inserted)
  struct Indel {
    int64_t pos;
    int len;
    bool is_del;
    int hap;
  };
  Indel indel(int contig, int64_t blk) const {
    uint64_t h = h3(seed, 0x2000 + contig, (uint64_t)blk);
    Indel d;
    d.pos = blk * 10000 + 50 + (int64_t)(h % 9900);
    d.len = 1 + (int)((h >> 20) % 10);
    d.is_del = (h >> 40) & 1;
    d.hap = (h >> 41) & 1;
    return d;
  }
  bool near_indel(int contig, int64_t p) const {
    Indel d = indel(contig, p / 10000);
    return p >= d.pos - 2 && p <= d.pos + d.len + 2;
  }
  // base carried by haplotype `hap` at p for a read with somatic draw u (tumor SNVs are carried when u < VAF)
  char hap_base(int contig, int64_t p, int hap, double u) const {
    char r = ref(contig, p);
    if (r == 'N' || near_indel(contig, p)) return r;
    int64_t blk = p / 1000;
    uint64_t h = h3(seed, 0x3000 + contig, (uint64_t)blk);
    if (p == blk * 1000 + (int64_t)(h % 1000)) {
      bool hom = ((h >> 20) % 3) == 0;
      int vh = (h >> 24) & 1;
      if (hom || vh == hap) return kBases[(((r == 'A' ? 0 : r == 'C' ? 1 : r == 'G' ? 2 : 3)) + 1 + (int)((h >> 28) % 3)) & 3];
      return r;
    }
    if (sample == 1) {
      int64_t sb = p / 100000;
      uint64_t hs = h3(seed, 0x5000 + contig, (uint64_t)sb);
      if (p == sb * 100000 + (int64_t)(hs % 100000)) {
        double vaf = 0.1 + 0.4 * (double)((hs >> 20) & 0xFFFF) / 65536.0;
        if (u < vaf) return kBases[(((r == 'A' ? 0 : r == 'C' ? 1 : r == 'G' ? 2 : 3)) + 1 + (int)((hs >> 40) % 3)) & 3];
      }
    }
    return r;
  }
};

struct OneRead {
  std::vector<uint32_t> cigar;
  std::string seq, md;
  std::vector<uint8_t> qual;
  uint8_t mapq, flags;
};

inline void push_op(std::vector<uint32_t>& c, uint32_t op, uint32_t len) {
  if (len == 0) return;
  if (!c.empty() && (c.back() & 0xF) == op)
    c.back() += len << 4;
  else
    c.push_back((len << 4) | op);
}

void make_read(const Genome& G, const guac_synth_params& P, int contig, int64_t start, uint64_t idx, OneRead& out) {
  Rng rng(h3(G.seed, 0x7000 + (uint64_t)G.sample, idx));
  const int L = P.read_length;
  out.cigar.clear();
  out.seq.clear();
  out.md.clear();
  out.qual.clear();
  const int hap = (int)(rng.next() & 1);
  const double u_som = rng.unit();
  // read class
  double cls = rng.unit();
  int lead = 0, trail = 0;
  bool seq_ins = false, seq_del = false;
  if (cls < P.frac_clip) {
    int clip = 5 + (int)rng.below(46);
    clip = std::min(clip, L / 3);
    if (rng.next() & 1) lead = clip; else trail = clip;
  } else if (cls < P.frac_clip + P.frac_ins) seq_ins = true;
  else if (cls < P.frac_clip + P.frac_ins + P.frac_del) seq_del = true;
  else if (cls < P.frac_clip + P.frac_ins + P.frac_del + P.frac_both) seq_ins = seq_del = true;
  const int target = L - lead - trail;  // read bases in the aligned part
  int ins_at = -1, del_at = -1;
  if (target > 40) {
    if (seq_ins) ins_at = 10 + (int)rng.below((uint32_t)(target - 30));
    if (seq_del) del_at = 10 + (int)rng.below((uint32_t)(target - 30));
    if (seq_ins && seq_del && std::abs(ins_at - del_at) < 8) del_at = -1;
  }
  auto draw_q = [&]() -> int {
    double t = rng.unit();
    if (t < 0.70) return 37 + (int)rng.below(5);
    if (t < 0.90) return 25 + (int)rng.below(12);
    return 2 + (int)rng.below(23);
  };
  auto put = [&](char b, int q) {
    out.seq.push_back(b);
    out.qual.push_back((uint8_t)q);
  };
  for (int i = 0; i < lead; ++i) put(kBases[rng.below(4)], draw_q());
  push_op(out.cigar, GUAC_CIGAR_S, (uint32_t)lead);
  int64_t pos = start;
  int produced = 0;  // aligned-part read bases so far
  long match_run = 0;
  bool last_was_indel = true;  // no indel before the first aligned base
  bool md_after_del = false;
  while (produced < target) {
    const int remaining = target - produced;
    const Genome::Indel gi = G.indel(contig, pos / 10000);
    const bool can_indel = !last_was_indel && remaining > 6 && produced > 5;
    if (can_indel && gi.hap == hap && gi.pos == pos) {
      if (gi.is_del) {
        out.md += std::to_string(match_run);
        match_run = 0;
        out.md.push_back('^');
        for (int k = 0; k < gi.len; ++k) out.md.push_back(G.ref(contig, pos + k));
        md_after_del = true;
        push_op(out.cigar, GUAC_CIGAR_D, (uint32_t)gi.len);
        pos += gi.len;
        last_was_indel = true;
        continue;
      } else if (remaining > gi.len + 6) {
        for (int k = 0; k < gi.len; ++k) put(kBases[h3(G.seed, 0x6000 + (uint64_t)contig, (uint64_t)(gi.pos * 16 + k)) & 3], draw_q());
        push_op(out.cigar, GUAC_CIGAR_I, (uint32_t)gi.len);
        produced += gi.len;
        last_was_indel = true;
        continue;
      }
    }
    if (can_indel && produced == ins_at) {
      int n = 1 + (int)rng.below(3);
      if (remaining > n + 6) {
        for (int k = 0; k < n; ++k) put(kBases[rng.below(4)], draw_q());
        push_op(out.cigar, GUAC_CIGAR_I, (uint32_t)n);
        produced += n;
        last_was_indel = true;
        ins_at = -1;
        continue;
      }
    }
    if (can_indel && produced == del_at) {
      int n = 1 + (int)rng.below(3);
      out.md += std::to_string(match_run);
      match_run = 0;
      out.md.push_back('^');
      for (int k = 0; k < n; ++k) out.md.push_back(G.ref(contig, pos + k));
      md_after_del = true;
      push_op(out.cigar, GUAC_CIGAR_D, (uint32_t)n);
      pos += n;
      last_was_indel = true;
      del_at = -1;
      continue;
    }
    const char r = G.ref(contig, pos);
    char b = G.hap_base(contig, pos, hap, u_som);
    if (b == 'N') b = kBases[rng.below(4)];
    const int q = draw_q();
    if (rng.unit() < std::pow(10.0, -q / 10.0)) {
      int code = (b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : 3);
      b = kBases[(code + 1 + (int)rng.below(3)) & 3];
    }
    put(b, q);
    if (b == r) {
      ++match_run;
    } else {
      out.md += std::to_string(match_run);
      match_run = 0;
      out.md.push_back(r);
    }
    md_after_del = false;
    push_op(out.cigar, GUAC_CIGAR_M, 1);
    ++pos;
    ++produced;
    last_was_indel = false;
  }
  (void)md_after_del;
  out.md += std::to_string(match_run);
  for (int i = 0; i < trail; ++i) put(kBases[rng.below(4)], draw_q());
  push_op(out.cigar, GUAC_CIGAR_S, (uint32_t)trail);
  out.mapq = (uint8_t)(rng.unit() < 0.9 ? 60 : rng.below(60));
  out.flags = (uint8_t)(GUAC_READ_HAS_MD | ((rng.next() & 1) ? GUAC_READ_POSITIVE_STRAND : 0));
}

}  // namespace

struct guac_synth_batch {
  std::vector<int64_t> contig_length, start;
  std::vector<int32_t> contig, sample;
  std::vector<uint64_t> cigar_off, seq_off, md_off;
  std::vector<uint32_t> cigar;
  std::vector<uint8_t> seq, qual, mapq, flags;
  std::vector<char> md;
  guac_read_batch view;
};

extern "C" {

void guac_synth_default_params(guac_synth_params* p) {
  std::memset(p, 0, sizeof *p);
  p->seed = 20261018;
  p->read_length = 150;
  p->frac_clip = 0.20;
  p->frac_ins = 0.009;
  p->frac_del = 0.009;
  p->frac_both = 0.002;
  p->n_threads = 0;
}

int guac_synth_generate(const guac_synth_params* P, guac_synth_batch** out) {
  if (!P || !out || !P->contig_length || P->n_contigs == 0 || P->read_length < 20 || P->read_length > 60000) return GUAC_ERR_INVALID_ARGUMENT;
  try {
    std::unique_ptr<guac_synth_batch> B(new guac_synth_batch());
    const uint64_t n = P->n_reads;
    B->contig_length.assign(P->contig_length, P->contig_length + P->n_contigs);
    // reads per contig proportional to the loci available, inside the optional window of contig `window_contig`
    std::vector<int64_t> lo(P->n_contigs), hi(P->n_contigs);
    double total = 0;
    for (uint32_t c = 0; c < P->n_contigs; ++c) {
      lo[c] = 0;
      hi[c] = P->contig_length[c];
      if (P->window_end > P->window_start) {
        if ((int32_t)c == P->window_contig) {
          lo[c] = std::max<int64_t>(0, P->window_start);
          hi[c] = std::min<int64_t>(hi[c], P->window_end);
        } else {
          hi[c] = 0;
        }
      }
      hi[c] = std::max<int64_t>(lo[c], hi[c] - (P->read_length + 40));  // reads stay inside the contig
      total += (double)(hi[c] - lo[c]);
    }
    if (total <= 0 && n) return GUAC_ERR_INVALID_ARGUMENT;
    std::vector<uint64_t> first(P->n_contigs + 1, 0);
    double acc = 0;
    for (uint32_t c = 0; c < P->n_contigs; ++c) {
      acc += (double)(hi[c] - lo[c]);
      first[c + 1] = (c + 1 == P->n_contigs) ? n : (uint64_t)std::llround((double)n * acc / total);
    }
    B->start.resize(n);
    B->contig.resize(n);
    B->sample.assign(n, P->sample);
    for (uint32_t c = 0; c < P->n_contigs; ++c) {
      Rng rng(h2(P->seed ^ 0xABCDEFull, c + 977ull * (uint64_t)P->sample));
      const uint64_t span = (uint64_t)(hi[c] - lo[c]);
      Genome Gc{P->seed, P->sample};
      for (uint64_t i = first[c]; i < first[c + 1]; ++i) {
        // aligners place no reads inside the reference's N runs: redraw starts whose read would touch one
        int64_t st = 0;
        for (int tries = 0; tries < 16; ++tries) {
          st = lo[c] + (int64_t)(rng.next() % std::max<uint64_t>(span, 1));
          const int64_t blk0 = st / 100000, blk1 = (st + P->read_length + 40) / 100000;
          bool hit = false;
          for (int64_t blk = blk0; blk <= blk1 && !hit; ++blk) {
            const int64_t n0 = blk * 100000 + (int64_t)(h3(P->seed, 0x4E00 + c, (uint64_t)blk) % 99900);
            hit = st < n0 + 100 && st + P->read_length + 40 > n0;
          }
          if (!hit) break;
        }
        B->start[i] = st;
        B->contig[i] = (int32_t)c;
      }
      (void)Gc;
      std::sort(B->start.begin() + first[c], B->start.begin() + first[c + 1]);
    }
    int nt = P->n_threads > 0 ? P->n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    nt = (int)std::min<uint64_t>((uint64_t)nt, std::max<uint64_t>(1, n / 4096));
    struct Part {
      std::vector<uint32_t> cigar, n_ops, md_len;
      std::vector<uint8_t> seq, qual, mapq, flags;
      std::vector<char> md;
    };
    std::vector<Part> parts(nt);
    Genome G{P->seed, P->sample};
    auto work = [&](int t) {
      const uint64_t a = n * (uint64_t)t / nt, b = n * (uint64_t)(t + 1) / nt;
      Part& pt = parts[t];
      pt.seq.reserve((b - a) * P->read_length);
      pt.qual.reserve((b - a) * P->read_length);
      OneRead r;
      for (uint64_t i = a; i < b; ++i) {
        make_read(G, *P, B->contig[i], B->start[i], i, r);
        pt.n_ops.push_back((uint32_t)r.cigar.size());
        pt.cigar.insert(pt.cigar.end(), r.cigar.begin(), r.cigar.end());
        pt.seq.insert(pt.seq.end(), r.seq.begin(), r.seq.end());
        pt.qual.insert(pt.qual.end(), r.qual.begin(), r.qual.end());
        pt.md_len.push_back((uint32_t)r.md.size());
        pt.md.insert(pt.md.end(), r.md.begin(), r.md.end());
        pt.mapq.push_back(r.mapq);
        pt.flags.push_back(r.flags);
      }
    };
    {
      std::vector<std::thread> th;
      for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
      work(0);
      for (auto& x : th) x.join();
    }
    B->cigar_off.assign(1, 0);
    B->seq_off.assign(1, 0);
    B->md_off.assign(1, 0);
    B->cigar_off.reserve(n + 1);
    B->seq_off.reserve(n + 1);
    B->md_off.reserve(n + 1);
    for (auto& pt : parts) {
      for (size_t k = 0; k < pt.n_ops.size(); ++k) {
        B->cigar_off.push_back(B->cigar_off.back() + pt.n_ops[k]);
        B->seq_off.push_back(B->seq_off.back() + (uint64_t)P->read_length);
        B->md_off.push_back(B->md_off.back() + pt.md_len[k]);
      }
      B->cigar.insert(B->cigar.end(), pt.cigar.begin(), pt.cigar.end());
      B->seq.insert(B->seq.end(), pt.seq.begin(), pt.seq.end());
      B->qual.insert(B->qual.end(), pt.qual.begin(), pt.qual.end());
      B->md.insert(B->md.end(), pt.md.begin(), pt.md.end());
      B->mapq.insert(B->mapq.end(), pt.mapq.begin(), pt.mapq.end());
      B->flags.insert(B->flags.end(), pt.flags.begin(), pt.flags.end());
      pt = Part();
    }
    if (B->cigar.empty()) B->cigar.push_back(0);
    if (B->md.empty()) B->md.push_back(0);
    if (B->seq.empty()) { B->seq.push_back(0); B->qual.push_back(0); }
    guac_read_batch& v = B->view;
    v.n_reads = n;
    v.n_contigs = P->n_contigs;
    v.contig_length = B->contig_length.data();
    v.contig = B->contig.data();
    v.start = B->start.data();
    v.cigar_off = B->cigar_off.data();
    v.cigar = B->cigar.data();
    v.seq_off = B->seq_off.data();
    v.seq = B->seq.data();
    v.qual = B->qual.data();
    v.mapq = B->mapq.data();
    v.flags = B->flags.data();
    v.sample = B->sample.data();
    v.md_off = B->md_off.data();
    v.md = B->md.data();
    *out = B.release();
    return GUAC_OK;
  } catch (const std::bad_alloc&) {
    return GUAC_ERR_OOM;
  }
}

const guac_read_batch* guac_synth_batch_view(const guac_synth_batch* b) { return b ? &b->view : nullptr; }
void guac_synth_batch_free(guac_synth_batch* b) { delete b; }

}  // extern "C"
