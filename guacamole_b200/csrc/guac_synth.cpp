// guac_synth.cpp — host build of the synthetic read generator (guac_synth_core.h): a guac_read_batch in host memory for the
// CPU oracle, the tests and the end-to-end bench leg.  No CUDA.  The output does not depend on the thread count.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "../../include/guac.h"
#include "../../include/guac_synth.h"
#include "guac_synth_core.h"
#include "guac_synth_tables.h"

struct guac_synth_batch {
  std::vector<int64_t> contig_length, start;
  std::vector<int32_t> contig, sample;
  std::vector<uint64_t> cigar_off, seq_off, md_off;
  std::vector<uint32_t> cigar;
  std::vector<uint8_t> seq, qual, mapq, flags;
  std::vector<char> md;
  guac_read_batch view;
};

namespace {

template <class F>
void parallel_for(uint64_t n, int nt, F f) {  // f(begin, end, part)
  nt = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)nt, n / 4096 + 1));
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(f, n * (uint64_t)t / nt, n * (uint64_t)(t + 1) / nt, t);
  f(0, n / nt, 0);
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

void guac_synth_default_params(guac_synth_params* p) {
  std::memset(p, 0, sizeof *p);
  p->seed = 20261018;
  p->read_length = 150;
  p->reads_per_locus = 30.0 / 150.0;
  p->frac_clip = 0.20;
  p->frac_ins = 0.009;
  p->frac_del = 0.009;
  p->frac_both = 0.002;
  p->with_qualities = 1;
}

int guac_synth_generate(const guac_synth_params* P, guac_synth_batch** out) {
  if (!P || !out || !P->contig_length || P->n_contigs == 0 || (P->n_windows && !P->windows)) return GUAC_ERR_INVALID_ARGUMENT;
  try {
    gsynth::Tables T;
    if (!gsynth::build_tables(*P, T)) return GUAC_ERR_INVALID_ARGUMENT;
    std::unique_ptr<guac_synth_batch> B(new guac_synth_batch());
    B->contig_length.assign(P->contig_length, P->contig_length + P->n_contigs);
    const std::vector<guac_locus_range> windows = gsynth::start_windows(*P);
    const int nt = P->n_threads > 0 ? P->n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    const gsynth::Genome G{T.seed, T.sample};
    // ---- read starts: the Poisson process over the loci of every window, in chunks of loci
    struct Chunk { int32_t contig; int64_t lo, hi; std::vector<int64_t> start; std::vector<uint32_t> rank; };
    std::vector<Chunk> chunks;
    for (const guac_locus_range& w : windows)
      for (int64_t lo = w.start; lo < w.end; lo += 65536) chunks.push_back(Chunk{w.contig, lo, std::min<int64_t>(w.end, lo + 65536), {}, {}});
    parallel_for(chunks.size(), nt, [&](uint64_t a, uint64_t b, int) {
      for (uint64_t i = a; i < b; ++i) {
        Chunk& c = chunks[i];
        for (int64_t p = c.lo; p < c.hi; ++p) {
          if (!gsynth::start_allowed(G, c.contig, p, P->contig_length[c.contig], T.read_length)) continue;
          const uint32_t k = gsynth::reads_starting_at(T, c.contig, p);
          for (uint32_t j = 0; j < k; ++j) { c.start.push_back(p); c.rank.push_back(j); }
        }
      }
    });
    uint64_t n = 0;
    for (auto& c : chunks) n += c.start.size();
    B->start.reserve(n);
    B->contig.reserve(n);
    std::vector<uint32_t> rank;
    rank.reserve(n);
    for (auto& c : chunks) {
      B->start.insert(B->start.end(), c.start.begin(), c.start.end());
      B->contig.insert(B->contig.end(), c.start.size(), c.contig);
      rank.insert(rank.end(), c.rank.begin(), c.rank.end());
      c = Chunk();
    }
    B->sample.assign(n, P->sample);
    // ---- sizes, offsets, content
    B->cigar_off.assign(n + 1, 0);
    B->md_off.assign(n + 1, 0);
    B->seq_off.assign(n + 1, 0);
    parallel_for(n, nt, [&](uint64_t a, uint64_t b, int) {
      for (uint64_t i = a; i < b; ++i) {
        gsynth::CountSink cs;
        gsynth::make_read(T, B->contig[i], B->start[i], rank[i], cs);
        B->cigar_off[i + 1] = cs.n_ops;
        B->md_off[i + 1] = cs.md_len;
        B->seq_off[i + 1] = cs.n_bases;
      }
    });
    for (uint64_t i = 0; i < n; ++i) {
      B->cigar_off[i + 1] += B->cigar_off[i];
      B->md_off[i + 1] += B->md_off[i];
      B->seq_off[i + 1] += B->seq_off[i];
    }
    B->cigar.resize(std::max<uint64_t>(1, B->cigar_off[n]));
    B->md.resize(std::max<uint64_t>(1, B->md_off[n]));
    B->seq.resize(std::max<uint64_t>(1, B->seq_off[n]));
    B->qual.resize(std::max<uint64_t>(1, B->seq_off[n]));
    B->mapq.resize(std::max<uint64_t>(1, n));
    B->flags.resize(std::max<uint64_t>(1, n));
    parallel_for(n, nt, [&](uint64_t a, uint64_t b, int) {
      for (uint64_t i = a; i < b; ++i) {
        gsynth::WriteSink ws{B->cigar.data() + B->cigar_off[i], B->seq.data() + B->seq_off[i], B->qual.data() + B->seq_off[i],
                             B->md.data() + B->md_off[i], B->mapq.data() + i, B->flags.data() + i};
        gsynth::make_read(T, B->contig[i], B->start[i], rank[i], ws);
      }
    });
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
