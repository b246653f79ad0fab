// guac_synth_tables.h — host-side set-up shared by both builds of the generator: the integer tables of guac_synth_core.h
// (everything that needs floating point happens here, once) and the list of start windows.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/guac_synth.h"
#include "guac_synth_core.h"

namespace gsynth {

inline uint32_t to_u32_threshold(double p) {
  if (!(p > 0)) return 0u;
  if (p >= 1.0) return 0xFFFFFFFFu;
  return (uint32_t)std::floor(p * 4294967296.0);
}

inline bool build_tables(const guac_synth_params& P, Tables& T) {
  if (P.read_length < 20 || P.read_length > 60000 || !(P.reads_per_locus >= 0) || P.reads_per_locus > 800.0) return false;
  std::memset(&T, 0, sizeof T);
  T.seed = P.seed;
  T.sample = P.sample;
  T.read_length = P.read_length;
  T.thr_clip = to_u32_threshold(P.frac_clip);
  T.thr_ins = to_u32_threshold(P.frac_clip + P.frac_ins);
  T.thr_del = to_u32_threshold(P.frac_clip + P.frac_ins + P.frac_del);
  T.thr_both = to_u32_threshold(P.frac_clip + P.frac_ins + P.frac_del + P.frac_both);
  for (int q = 0; q < 64; ++q) T.q_error[q] = to_u32_threshold(std::pow(10.0, -q / 10.0));
  // Poisson CDF scaled to 2^53 (u = 53-bit uniform; K = first k with u < cdf[k]); the last entry is forced to 2^53
  const double lambda = P.reads_per_locus;
  long double term = std::exp(-(long double)lambda), cum = 0;
  int n = 0;
  for (int k = 0; k < kMaxPoisson; ++k) {
    cum += term;
    const long double scaled = cum * 9007199254740992.0L;
    T.poisson_cdf[k] = scaled >= 9007199254740992.0L ? 9007199254740992ull : (uint64_t)scaled;
    n = k + 1;
    if ((k > lambda && term < 1e-22L) || T.poisson_cdf[k] >= 9007199254740992ull) break;
    term *= (long double)lambda / (long double)(k + 1);
  }
  T.poisson_cdf[n - 1] = 9007199254740992ull;
  T.n_poisson = n;
  return true;
}

// the windows of read starts: as given, or every contig whole; clipped to the loci where a read fits its contig
inline std::vector<guac_locus_range> start_windows(const guac_synth_params& P) {
  std::vector<guac_locus_range> w;
  auto add = [&](int32_t c, int64_t s, int64_t e) {
    if (c < 0 || (uint32_t)c >= P.n_contigs) return;
    s = std::max<int64_t>(s, 0);
    e = std::min<int64_t>(e, P.contig_length[c] - (P.read_length + kPad) + 1);
    if (e > s) w.push_back(guac_locus_range{c, 0, s, e});
  };
  if (P.n_windows == 0)
    for (uint32_t c = 0; c < P.n_contigs; ++c) add((int32_t)c, 0, P.contig_length[c]);
  else
    for (uint32_t i = 0; i < P.n_windows; ++i) add(P.windows[i].contig, P.windows[i].start, P.windows[i].end);
  return w;
}

}  // namespace gsynth
