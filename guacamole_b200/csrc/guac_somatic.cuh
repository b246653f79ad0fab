// guac_somatic.cuh — somatic-standard caller: tumor / normal genotype likelihoods per locus, fused with the pileup.
//
//   PileupFilter.apply (mapq >= min, optional multi-allelic)              filters/PileupFilter.scala:69-89
//   Likelihood.likelihoodsOfAllPossibleGenotypesFromPileup / OfGenotypes   likelihood/Likelihood.scala:99-113, 149-201
//   SomaticStandard.Caller.findPotentialVariantAtLocus                     commands/SomaticStandardCaller.scala:162-245
//   AlleleEvidence.apply                                                   variants/AlleleEvidence.scala:58-101
//
// K_somatic       one warp per 32 loci, lane = locus (gather): both samples' overlapping reads are walked in lockstep by the
//                 warp (record loads are broadcast, base-quality loads are coalesced: consecutive lanes = consecutive bases
//                 of the same read).  Every element adds table values log(s + s), log((1-s) + (1-s)) into per-allele fp64
//                 sums; loci whose elements are all A/C/G/T matches / mismatches are decided right there.
// K_somatic_exact one warp per locus holding an insertion / deletion / clipped / non-ACGT element: literal per-element walk
//                 with a small allele table.
// K_evidence      one warp per emitted record: AlleleEvidence of both samples (depths, strand depths, mean / median MQ, BQ,
//                 median mismatches).
//
// With P[a][e] = s_e if element e carries allele a else 1 - s_e, the reference's
//   logL(a1, a2) = sum_e log(P[a1][e] + P[a2][e]) - log(2) * depth
// depends only on per-allele sums:  S1[a] = sum_{e in a} log(s + s),  S0[a] = sum_{e in a} log((1-s) + (1-s)),
// T0 = sum_all log((1-s) + (1-s)):   hom(a) = S1[a] + T0 - S0[a],  het(a, b) = T0 - S0[a] - S0[b]  (+ sum log(s + (1-s)),
// which is 0 up to one ulp per element).  Summation order differs from colt's last-to-first walk: results agree with the
// reference to ~1e-13 relative, inside the 1e-9 tolerance the north star states.
#pragma once

#include "guac_host.cuh"
#include "guac_pileup.cuh"
#include "guac_rows.cuh"
#include "guac_order.cuh"

namespace guac {

constexpr int kSomMaxAlleles = 32;                                           // alleles entering the genotype enumeration
constexpr int kSomMaxGenotypes = kSomMaxAlleles * (kSomMaxAlleles + 1) / 2;
#ifndef GUAC_LEAN_UNROLL
#define GUAC_LEAN_UNROLL 3
#endif
#ifndef GUAC_SOM_MINB
#define GUAC_SOM_MINB 3
#endif
constexpr int kLeanUnroll = GUAC_LEAN_UNROLL;                                               // reads whose element loads fly together
constexpr int kSomTab = 64;                                                  // distinct alleles kept per sample and locus

// d_tables layout (doubles): succ[256] | normal (l1, l0)[256] | tumor (l1, l0)[256 mapq][256 quality] — pairs are read as double2
constexpr int kTabSucc = 0, kTabN = 256, kTabT = 768, kTabTotal = 768 + 2 * 65536;

struct SomParams {
  int32_t odds_threshold, min_mapq, filter_multi_allelic, max_read_depth, skip_empty, tumor_sample;
};

struct SomOut {
  guac_somatic_record* rec;
  uint32_t cap_rec;
  uint8_t* pool;
  uint32_t cap_pool;
  SlowLocus* slow;
  uint32_t cap_slow;
  unsigned long long* counters;  // [0] records [1] pool bytes [2] slow loci [3] visited loci
  DevError* err;
};

// ---- scala.math.round(Double).toInt and ADAM PhredUtils.successProbabilityToPhred (see oracle/guac_oracle.cpp) ------------
__device__ inline int java_round_to_int(double x) {
  long long r;
  if (isnan(x)) r = 0;
  else {
    double f = floor(x + 0.5);
    if (f >= 9.2233720368547758e18) r = 0x7FFFFFFFFFFFFFFFll;
    else if (f <= -9.2233720368547758e18) r = (long long)0x8000000000000000ull;
    else r = (long long)f;
  }
  return (int)(unsigned int)(unsigned long long)r;
}
__device__ inline int success_probability_to_phred(double p) { return java_round_to_int(-10.0 * log10(1.0 - p)); }

// ---- per-sample allele statistics at one locus -------------------------------------------------------------------------------
struct SomAllele {
  int kind;      // AlleleEntry kinds: 0 SNV, 2 insertion, 3 deletion, 4 mid-deletion, 5 clipped
  int len;
  uint64_t ptr;
  uint8_t base;
  int count;
  int n0inf;     // elements whose l0 = log((1-s) + (1-s)) is -inf (s == 1: a deletion element of a mapq-255 read), kept out of s0
  double s1, s0;
};

struct SampleStats {
  int depth;      // filtered elements
  int ref_depth;  // Match elements among them
  int n_alleles;  // entries in tab (all distinct alleles of the filtered pileup)
  int distinct_unfiltered;  // distinct alleles before the mapq filter (multi-allelic filter), capped
  int t0inf;                // elements with l0 = -inf, kept out of t0 (T0 - S0[a] must not become inf - inf)
  double t0;
};

__device__ inline AlleleEntry as_entry(const SomAllele& a) {
  AlleleEntry e;
  e.kind = a.kind; e.len = a.len; e.ptr = a.ptr; e.base = a.base; e.count = a.count;
  return e;
}

// per-warp scratch of the general genotype enumeration (shared memory: 528 genotypes would not fit a thread's stack)
struct GenotypeScratch {
  double lk[kSomMaxGenotypes];
  uint8_t gi[kSomMaxGenotypes], gj[kSomMaxGenotypes];
};

// Likelihood.likelihoodsOfAllPossibleGenotypesFromPileup(pileup, probabilityCorrect, normalize = true), plain probabilities.
// tab[0..n) must be sorted by Allele.compare.  Returns the number of genotypes; lk[] in the reference's (i <= j) order.
__device__ int genotype_likelihoods(const AlleleView& av, const SomAllele* tab, int n_tab, const SampleStats& st, uint8_t* gi, uint8_t* gj,
                                    double* lk, bool log_space = false) {
  int idx[kSomMaxAlleles], n = 0;
  for (int k = 0; k < n_tab; ++k) {  // alleles whose alternate bases are all standard (an empty alternate passes)
    const AlleleEntry e = as_entry(tab[k]);
    bool ok = true;
    const int al = av.alt_len(e);
    for (int i = 0; i < al && ok; ++i) ok = is_std_base(av.alt_at(e, i));
    if (ok) {
      if (n == kSomMaxAlleles) return -1;
      idx[n++] = k;
    }
  }
  int ng = 0;
  const double nlog2 = log(2.0) * (double)st.depth;
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) {
      const SomAllele& a = tab[idx[i]];
      const SomAllele& b = tab[idx[j]];
      // sum over the elements outside the genotype's alleles of l0: -inf as soon as one of them has l0 = -inf
      const int outside_inf = st.t0inf - a.n0inf - (i == j ? 0 : b.n0inf);
      const double outside = outside_inf > 0 ? -1.0 / 0.0 : ((i == j) ? (st.t0 - a.s0) : (st.t0 - a.s0 - b.s0));
      const double agg = (i == j) ? (a.s1 + outside) : outside;
      gi[ng] = (uint8_t)idx[i];
      gj[ng] = (uint8_t)idx[j];
      lk[ng] = agg + 0.0 - nlog2;
      ++ng;
    }
  double total = 0.0;
  for (int g = 0; g < ng; ++g) total += exp(lk[g]);
  const double log_total = log(total);  // naive normalisation on purpose: Inf / NaN flow like the reference
  for (int g = 0; g < ng; ++g) lk[g] = log_space ? lk[g] - log_total : exp(lk[g] - log_total);
  return ng;
}

// findPotentialVariantAtLocus, tumor half: early outs + the most likely tumor genotype.  Returns true when that genotype
// holds a variant allele (only then does the reference look at the normal sample); a1 / a2 / tumor_l describe it.
__device__ bool tumor_most_likely(const AlleleView& avT, const SomAllele* tabT, const SampleStats& sT, int contig, int locus,
                                  const SomParams& prm, SomOut& out, GenotypeScratch& G, AlleleEntry* a1, AlleleEntry* a2, double* tumor_l) {
  if (sT.depth == 0 || sT.depth > prm.max_read_depth || sT.ref_depth == sT.depth) return false;
  uint8_t* gi = G.gi;
  uint8_t* gj = G.gj;
  double* lk = G.lk;
  const int ng = genotype_likelihoods(avT, tabT, sT.n_alleles, sT, gi, gj, lk);
  if (ng < 0) { report_error(out.err, GUAC_ERR_UNSUPPORTED, ((unsigned long long)contig << 32) | (uint32_t)locus); return false; }
  if (ng == 0) return false;
  int best = 0;  // maxBy = reduceLeft((x, y) => if (f(x) >= f(y)) x else y)
  for (int g = 1; g < ng; ++g)
    if (!(lk[best] >= lk[g])) best = g;
  *a1 = as_entry(tabT[gi[best]]);
  *a2 = as_entry(tabT[gj[best]]);
  *tumor_l = lk[best];
  return avT.is_variant(*a1) || avT.is_variant(*a2);
}

// ... normal half, last step: somatic odds from the normal sample's variant-genotype mass, then the record
__device__ void emit_somatic(const AlleleView& avT, const AlleleEntry& a1, const AlleleEntry& a2, double tumor_l,
                             double normal_variants_total, int contig, int locus, const SomParams& prm, SomOut& out) {
  const double somatic_odds = tumor_l / normal_variants_total;
  if (!(somatic_odds * 100 >= (double)prm.odds_threshold)) return;
  // first non-reference allele of the genotype whose alternate is not empty
  const bool v1 = avT.is_variant(a1), v2 = avT.is_variant(a2);
  const AlleleEntry* allele = nullptr;
  if (v1 && !avT.alt_empty(a1)) allele = &a1;
  else if (v2 && !avT.alt_empty(a2)) allele = &a2;
  if (!allele) return;
  guac_somatic_record r;
  memset(&r, 0, sizeof r);
  r.start = locus;
  r.contig = contig;
  r.sample = prm.tumor_sample;
  const int rl = avT.ref_len(*allele), al = avT.alt_len(*allele);
  const uint32_t o = kPoolDynOff + (uint32_t)atomicAdd(&out.counters[1], (unsigned long long)(rl + al));
  if ((unsigned long long)o + rl + al <= out.cap_pool) {
    for (int i = 0; i < rl; ++i) out.pool[o + i] = avT.ref_at(*allele, i);
    for (int i = 0; i < al; ++i) out.pool[o + rl + i] = avT.alt_at(*allele, i);
  }
  r.ref_off = o;
  r.ref_len = (uint16_t)rl;
  r.alt_off = o + rl;
  r.alt_len = (uint16_t)al;
  r.somatic_log_odds = log(somatic_odds);
  r.tumor.likelihood = tumor_l;
  r.normal.likelihood = 1 - normal_variants_total;
  r.phred_scaled_somatic_likelihood = success_probability_to_phred(r.tumor.likelihood * r.normal.likelihood - 1e-10);
  const uint32_t s = (uint32_t)atomicAdd(&out.counters[0], 1ull);
  if (s < out.cap_rec) out.rec[s] = r;
}

// ... normal half: the normal sample's genotype likelihoods, summed over the genotypes holding a variant allele
__device__ void somatic_against_normal(const AlleleView& avT, const AlleleView& avN, const AlleleEntry& a1, const AlleleEntry& a2,
                                       double tumor_l, const SomAllele* tabN, const SampleStats& sN, int contig, int locus,
                                       const SomParams& prm, SomOut& out, GenotypeScratch& G) {
  if (sN.depth == 0 || sN.depth > prm.max_read_depth) return;
  uint8_t* gi = G.gi;
  uint8_t* gj = G.gj;
  double* lk = G.lk;
  const int ngn = genotype_likelihoods(avN, tabN, sN.n_alleles, sN, gi, gj, lk);
  if (ngn < 0) { report_error(out.err, GUAC_ERR_UNSUPPORTED, ((unsigned long long)contig << 32) | (uint32_t)locus); return; }
  double normal_variants_total = 0.0;
  for (int g = 0; g < ngn; ++g)
    if (avN.is_variant(as_entry(tabN[gi[g]])) || avN.is_variant(as_entry(tabN[gj[g]]))) normal_variants_total += lk[g];
  emit_somatic(avT, a1, a2, tumor_l, normal_variants_total, contig, locus, prm, out);
}

// findPotentialVariantAtLocus once both filtered pileups are summarised.  tabT / tabN sorted by Allele.compare.
__device__ void decide_somatic(const DevReads& RT, const AlleleView& avT, const AlleleView& avN, const SomAllele* tabT,
                               const SampleStats& sT, const SomAllele* tabN, const SampleStats& sN, int contig, int locus,
                               const SomParams& prm, SomOut& out, GenotypeScratch& G) {
  (void)RT;
  if (sT.depth == 0 || sN.depth == 0 || sT.depth > prm.max_read_depth || sN.depth > prm.max_read_depth) return;
  AlleleEntry a1, a2;
  double tumor_l;
  if (!tumor_most_likely(avT, tabT, sT, contig, locus, prm, out, G, &a1, &a2, &tumor_l)) return;
  somatic_against_normal(avT, avN, a1, a2, tumor_l, tabN, sN, contig, locus, prm, out, G);
}

// sort a small allele table by Allele.compare (insertion sort)
__device__ void sort_alleles(const AlleleView& av, SomAllele* tab, int n) {
  for (int i = 1; i < n; ++i) {
    SomAllele x = tab[i];
    int j = i;
    while (j > 0 && av.compare(as_entry(tab[j - 1]), as_entry(x)) > 0) {
      tab[j] = tab[j - 1];
      --j;
    }
    tab[j] = x;
  }
}

__device__ inline int lower_bound_start(const ReadRec* rec, uint64_t lo, uint64_t hi, int value) {  // first i with start >= value
  while (lo < hi) {
    uint64_t mid = (lo + hi) >> 1;
    if (rec[mid].start >= value) hi = mid; else lo = mid + 1;
  }
  return (int)lo;
}

// ---- K_somatic: warp per 32 loci, lane per locus ---------------------------------------------------------------------------------
struct LaneAcc {
  int depth, ref_depth, other;   // filtered depth, Match elements, elements that are not A/C/G/T matches / mismatches
  int hard;                      // elements only the exact kernel can handle (it reports their error)
  int any;                       // overlapping reads before any filter (decides whether the locus is visited)
  int cnt[4];                    // filtered elements by base code
  uint32_t seen;                 // base codes seen before the mapq filter (multi-allelic filter)
  double t0, s1[4], s0[4];
  // over the kept "other" elements (l1, l0 = their table values): sum l0, sum max(0, l0), sum max(l1, l0)
  double o0, ohet, ohom;
};

__device__ __forceinline__ void acc_clear(LaneAcc& a) {
  a.depth = a.ref_depth = a.other = a.hard = a.any = 0;
  a.seen = 0;
  a.t0 = a.o0 = a.ohet = a.ohom = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) { a.cnt[k] = 0; a.s1[k] = 0.0; a.s0[k] = 0.0; }
}

// ---- loci whose elements are all A/C/G/T matches / mismatches: at most 4 alleles and 10 genotypes, kept in registers -------
// (the general routines above index small local arrays, which the compiler places in local memory: at 768 threads per SM
// that traffic alone thrashes L1.)  Alleles in base-code order = Allele.compare order, since they share the reference base.
struct SnvAlleles {
  int n, depth;
  int code[4];
  double s1[4], s0[4], t0;
};

__device__ __forceinline__ void snv_compact(const LaneAcc& A, SnvAlleles& S) {
  S.n = 0;
  S.depth = A.depth;
  S.t0 = A.t0;
#pragma unroll
  for (int p = 0; p < 4; ++p) { S.code[p] = 0; S.s1[p] = 0.0; S.s0[p] = 0.0; }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool have = A.cnt[k] > 0;
#pragma unroll
    for (int p = 0; p <= k; ++p) {
      const bool here = have && p == S.n;
      S.code[p] = here ? k : S.code[p];
      S.s1[p] = here ? A.s1[k] : S.s1[p];
      S.s0[p] = here ? A.s0[k] : S.s0[p];
    }
    S.n += have ? 1 : 0;
  }
}

// un-normalised log likelihoods of the genotypes (i <= j < n) in the static slot order (0,0) (0,1) (0,2) (0,3) (1,1) ... (3,3),
// which restricted to the valid slots is the reference's enumeration order.  Same expression as genotype_likelihoods().
__device__ __forceinline__ void snv_log_likelihoods(const SnvAlleles& S, double (&lk)[10]) {
  const double nlog2 = log(2.0) * (double)S.depth;
  int g = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) {
      const double agg = (i == j) ? (S.s1[i] + (S.t0 - S.s0[i])) : (S.t0 - S.s0[i] - S.s0[j]);
      lk[g++] = agg + 0.0 - nlog2;
    }
}

// Likelihood.likelihoodsOfAllPossibleGenotypesFromPileup(..., normalize = true) on the valid slots, in place
template <bool LOG_SPACE = false>
__device__ __forceinline__ void snv_normalize(const SnvAlleles& S, double (&lk)[10]) {
  double total = 0.0;
  int g = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j, ++g)
      if (j < S.n) total += exp(lk[g]);
  const double log_total = log(total);  // naive normalisation on purpose
  g = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j, ++g)
      if (j < S.n) lk[g] = LOG_SPACE ? lk[g] - log_total : exp(lk[g] - log_total);
}

// tumor half of findPotentialVariantAtLocus (tumor_most_likely() above) for such a locus over the standard reference base
// code `rc`.  *c1 / *c2 = base codes of the most likely genotype.
__device__ __forceinline__ bool snv_tumor_most_likely(const SnvAlleles& S, int ref_depth, int rc, const SomParams& prm, int* c1, int* c2,
                                                      double* tumor_l) {
  if (S.depth == 0 || S.depth > prm.max_read_depth || ref_depth == S.depth) return false;
  double lk[10];
  snv_log_likelihoods(S, lk);
  // Exact early out.  If the homozygous-reference genotype leads every other genotype by a clear margin in log space (and
  // is far from underflow, and nothing is NaN), then it is also the maximum after exp / normalisation, which are monotone
  // and accurate to an ulp: the most likely genotype holds no variant allele and nothing else about it is needed.
  {
    double ref_lk = -1.0 / 0.0;
    int g = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = i; j < 4; ++j, ++g)
        if (i == j && i < S.n && S.code[i] == rc) ref_lk = lk[g];
    bool lead = ref_lk > -700.0;
    g = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = i; j < 4; ++j, ++g)
        if (j < S.n && !(i == j && S.code[i] == rc)) lead = lead && (lk[g] < ref_lk - 1e-6);
    if (lead) return false;
  }
  snv_normalize(S, lk);
  // maxBy = reduceLeft((x, y) => if (f(x) >= f(y)) x else y)
  double best = 0.0;
  int bi = 0, bj = 0, g = 0;
  bool first = true;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j, ++g)
      if (j < S.n) {
        if (first || !(best >= lk[g])) { best = lk[g]; bi = i; bj = j; }
        first = false;
      }
  int ci = 0, cj = 0;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    ci = bi == p ? S.code[p] : ci;
    cj = bj == p ? S.code[p] : cj;
  }
  *c1 = ci;
  *c2 = cj;
  *tumor_l = best;
  return ci != rc || cj != rc;
}

// normal half: likelihood mass of the normal genotypes holding a variant allele (reference base code `rc`)
__device__ __forceinline__ double snv_normal_variants_total(const SnvAlleles& S, int rc) {
  double lk[10];
  snv_log_likelihoods(S, lk);
  snv_normalize(S, lk);
  double total = 0.0;
  int g = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j, ++g)
      if (j < S.n && (S.code[i] != rc || S.code[j] != rc)) total += lk[g];
  return total;
}

// A locus that also holds insertion / deletion / clipped / non-ACGT elements (set O, kept by the mapq filter).  Which
// alleles they form is unknown here, but every genotype has an upper bound: an element of O contributes at most
// max(l1, l0) to a homozygous genotype and at most max(0, l0) to a heterozygous one, an A/C/G/T element contributes what it
// always does.  If the homozygous-reference genotype (whose value is exact) leads all of these bounds by a clear margin,
// it is the most likely genotype, tumor_most_likely() would return "no variant allele" and the locus yields nothing:
// no exact pass needed.  (Alleles the reference drops from the enumeration only shrink the candidate set.)
__device__ __forceinline__ bool ref_leads_despite_others(const LaneAcc& A, int rc) {
  double s1r = 0.0, s0r = 0.0;
  int cr = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    s1r = k == rc ? A.s1[k] : s1r;
    s0r = k == rc ? A.s0[k] : s0r;
    cr = k == rc ? A.cnt[k] : cr;
  }
  if (cr == 0) return false;
  const double ref_lk = s1r + (A.t0 - s0r) + A.o0;
  bool lead = ref_lk - log(2.0) * (double)(A.depth + A.other) > -700.0;  // (false for NaN)
  const double bar = ref_lk - 1e-6;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool have = k != rc && A.cnt[k] > 0;
    lead = lead && (!have || (A.s1[k] + (A.t0 - A.s0[k]) + A.o0 < bar));  // hom(k)
    lead = lead && (!have || (A.t0 - s0r - A.s0[k] + A.o0 < bar));        // het(ref, k)
    lead = lead && (!have || ((A.t0 - A.s0[k]) + A.ohet < bar));          // het(k, O*)
#pragma unroll
    for (int m = k + 1; m < 4; ++m) {
      const bool both = have && m != rc && A.cnt[m] > 0;
      lead = lead && (!both || (A.t0 - A.s0[k] - A.s0[m] + A.o0 < bar));  // het(k, m)
    }
  }
  lead = lead && ((A.t0 - s0r) + A.ohet < bar);  // het(ref, O*)
  lead = lead && (A.t0 + A.ohet < bar);          // het(O*, O*)
  lead = lead && (A.t0 + A.ohom < bar);          // hom(O*)
  return lead;
}

__device__ __forceinline__ AlleleEntry snv_entry(int code, int count) {
  AlleleEntry e;
  e.kind = 0; e.len = 1; e.ptr = 0; e.base = code_base(code); e.count = count;
  return e;
}

template <bool TUMOR>
__device__ void gather_sample(const DevReads& R, int contig, int span_lo, int x, int rcode, bool std_ref,
                              const SomParams& prm, const double* __restrict__ tables, LaneAcc& A) {
  const int lane = threadIdx.x & 31;
  const ContigInfo ci = R.contigs[contig];
  acc_clear(A);
  if (span_lo >= ci.length) return;
  // candidate reads of the word: the granule's read range (no dependent binary-search chain); every lane tests one record
  // and the warp then walks only the reads that really overlap the 32 loci
  const int g = span_lo >> kGranuleShift;
  uint32_t first = R.gran_first[ci.gran_off + g], last = R.gran_last[ci.gran_off + g];
  if (first == 0xFFFFFFFFu) return;
  narrow_candidates(R, first, last, span_lo, span_lo + 32);
  const double2* __restrict__ tab = reinterpret_cast<const double2*>(tables + (TUMOR ? kTabT : kTabN));
  const uint8_t ref_base = std_ref ? code_base(rcode) : (uint8_t)'N';
  // the reference class is the common one: it keeps one running sum (its S0 is T0 minus the other classes' S0 at the
  // end) and a plain counter; the mismatch classes are touched only when some lane mismatches
  double sr1 = 0.0;
  int n_ref = 0;
  unsigned long long cnt_packed = 0;  // mismatching elements: four 16-bit fields, one per base code
  uint32_t n_word_reads = 0;          // reads overlapping the word (bounds every counter field)
  const uint32_t rc_eff = std_ref ? (uint32_t)rcode : 4u;
  const bool fma = prm.filter_multi_allelic != 0;
  const int min_mapq = prm.min_mapq;
  constexpr uint32_t kLeanMask = kInfoSimple | kInfoHasExc | kInfoWideQ;
  for (uint32_t base = first; base < last; base += 32) {
    const uint32_t mine = base + lane;
    ReadRec my{0, 0, 0, 0};
    if (mine < last) my = R.rec[mine];
    const bool overlaps = mine < last && my.start < span_lo + 32 && my.end > span_lo && my.end > my.start;
    const bool keep_mine = !(min_mapq > 0) || (int)(my.info >> kInfoMapqShift) >= min_mapq;
    // kept SIMPLE reads of plain A/C/G/T bases with 6-bit qualities take the lean loop: one byte (quality | base code << 6)
    // per pileup element.  Everything else (CIGAR walk, mapq-dropped reads that only count towards the visited loci / the
    // multi-allelic filter) takes the general loop below.
    const bool lean_mine = overlaps && keep_mine && (my.info & kLeanMask) == kInfoSimple;
    uint32_t ov = __ballot_sync(0xFFFFFFFFu, lean_mine);
    uint32_t ov_general = __ballot_sync(0xFFFFFFFFu, overlaps && !lean_mine);
    n_word_reads += (uint32_t)(__popc(ov) + __popc(ov_general));
    // Every lane owns one read of the batch here: it forms the address its read's byte for locus 0 would have (so the
    // per-read loop adds just the lane's locus), packs (span, mapq << 8) into one word and pulls the bytes this word's loci
    // need into the cache now, so that the loop below does not serialise one memory round trip after another.
    uint64_t my_qa = 0;
    uint32_t my_lm = 0;
    if (lean_mine) {
      my_qa = (uint64_t)(uintptr_t)R.qc + R.seq_off[mine] + (uint64_t)(my.info & kInfoLeadMask) - (uint64_t)(int64_t)my.start;
      my_lm = (uint32_t)(my.end - my.start) | ((my.info >> kInfoMapqShift) << 24);
      const int x0 = max(span_lo, my.start);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(my_qa + (uint64_t)(int64_t)x0));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(my_qa + (uint64_t)(int64_t)min(span_lo + 31, my.end - 1)));
    }
    while (ov) {  // warp-uniform, lean: kLeanUnroll reads per round, so that their element loads are in flight together
      bool inr[kLeanUnroll];
      uint64_t ea[kLeanUnroll];
      uint32_t lmu[kLeanUnroll];
#pragma unroll
      for (int u = 0; u < kLeanUnroll; ++u) {
        const bool act = ov != 0;
        const int j = __ffs(ov) - 1;  // (-1 once the batch is exhausted: the shuffles then read lane 31, unused)
        ov &= ov - 1;
        const int start = __shfl_sync(0xFFFFFFFFu, my.start, j);
        lmu[u] = __shfl_sync(0xFFFFFFFFu, my_lm, j);
        const uint64_t qa = ((uint64_t)__shfl_sync(0xFFFFFFFFu, (uint32_t)(my_qa >> 32), j) << 32) |
                            __shfl_sync(0xFFFFFFFFu, (uint32_t)my_qa, j);
        inr[u] = act && (unsigned)(x - start) < (lmu[u] & 0xFFFFu);  // is this lane's locus inside the read?
        ea[u] = qa + (uint64_t)(int64_t)x;
      }
      uint32_t bq[kLeanUnroll];
#pragma unroll
      for (int u = 0; u < kLeanUnroll; ++u) bq[u] = inr[u] ? (uint32_t)__ldg(reinterpret_cast<const uint8_t*>((uintptr_t)ea[u])) : 0u;
      double2 lq[kLeanUnroll];
#pragma unroll
      for (int u = 0; u < kLeanUnroll; ++u)  // (log(s + s), log((1-s) + (1-s)))
        lq[u] = inr[u] ? __ldg(&tab[TUMOR ? ((lmu[u] >> 16) & 0xFF00u) + (bq[u] & 63u) : (bq[u] & 63u)]) : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < kLeanUnroll; ++u) {  // accumulate in read order
        if (!inr[u]) continue;
        const double2 l = lq[u];
        const uint32_t b = bq[u];
        A.t0 += l.y;
        if ((b >> 6) == rc_eff) {
          sr1 += l.x;
          n_ref += 1;
        } else {
          const uint32_t code = b >> 6;
          cnt_packed += 1ull << (16 * code);
          if (code == 0u) { A.s1[0] += l.x; A.s0[0] += l.y; }
          else if (code == 1u) { A.s1[1] += l.x; A.s0[1] += l.y; }
          else if (code == 2u) { A.s1[2] += l.x; A.s0[2] += l.y; }
          else { A.s1[3] += l.x; A.s0[3] += l.y; }
        }
      }
    }
    while (ov_general) {  // warp-uniform, rare
      const int j = __ffs(ov_general) - 1;
      ov_general &= ov_general - 1;
      const ReadRec rec = R.rec[base + j];
      if (!(rec.start <= x && x < rec.end)) continue;
      A.any += 1;
      const int mapq = (int)(rec.info >> kInfoMapqShift);
      const bool keep = !(min_mapq > 0) || mapq >= min_mapq;
      if (!keep && !fma) continue;
      Elem e;
      const int rc = classify(R, (uint64_t)(base + j), x, ref_base, e);
      if (rc || e.kind == kNone) {  // an error the exact kernel reports
        A.other += 1;
        A.hard += 1;
        continue;
      }
      if (!((e.kind == kMatch || e.kind == kMismatch) && is_std_base(e.base))) {
        A.other += 1;  // insertion / deletion / clipped / non-ACGT element: which allele it carries is the exact kernel's job,
        if (keep) {    // but its likelihood terms bound every genotype it can be part of (ref_leads_despite_others)
          const double2 l = __ldg(&tab[TUMOR ? (mapq << 8) + (e.qual & 255) : (e.qual & 255)]);
          A.o0 += l.y;
          A.ohet += fmax(0.0, l.y);
          A.ohom += fmax(l.x, l.y);
        }
        continue;
      }
      const int code = (int)base_code(e.base);
      A.seen |= 1u << code;
      if (!keep) continue;
      const double2 l = __ldg(&tab[TUMOR ? (mapq << 8) + (e.qual & 255) : (e.qual & 255)]);
      A.t0 += l.y;
      if ((uint32_t)code == rc_eff) {
        sr1 += l.x;
        n_ref += 1;
      } else {
        cnt_packed += 1ull << (16 * code);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool is = code == k;
          A.s1[k] += is ? l.x : 0.0;
          A.s0[k] += is ? l.y : 0.0;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    A.cnt[k] = (int)((cnt_packed >> (16 * k)) & 0xFFFFu) + ((uint32_t)k == rc_eff ? n_ref : 0);
    A.depth += A.cnt[k];
  }
  A.any += A.depth;  // (only any > 0 matters: the lean loop's elements all count towards the depth)
  A.ref_depth = !std_ref ? 0 : rcode == 0 ? A.cnt[0] : rcode == 1 ? A.cnt[1] : rcode == 2 ? A.cnt[2] : A.cnt[3];  // (no dynamic index: A stays in registers)
#pragma unroll
  for (int k = 0; k < 4; ++k) A.seen |= A.cnt[k] > 0 ? (1u << k) : 0u;
  if (n_word_reads > 0xFFFFu) {  // the packed counters hold 16 bits and may have wrapped: the exact kernel decides this word's
    A.other += 1;                // loci, whatever the bounds computed from the wrapped counts would say
    A.hard += 1;
  }
  // S0 of the reference class = T0 - the other classes' S0 (every kept plain element is in exactly one class)
  double sr0 = A.t0;
#pragma unroll
  for (int k = 0; k < 4; ++k) sr0 -= A.s0[k];
  // fold the reference-class sums into their base code
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool is = std_ref && k == rcode;
    A.s1[k] += is ? sr1 : 0.0;
    A.s0[k] += is ? sr0 : 0.0;
  }
}

// ---- the same sums from the word's transposed pileup (guac_rows.cuh): no per-read work, the table in shared memory -------------
// Shared-memory table of one CTA.  Tumor: 64 rows of 64 (log(s + s), log((1-s) + (1-s))) pairs, s = success(quality) *
// success(mapq) (probabilityCorrectIncludingAlignment, likelihood/Likelihood.scala:58-62): row r belongs to the r-th largest
// mapping quality present in the tumor sample; rows of mapping qualities below --min-mapq, the unused rows and row 63 (the
// sentinel's) are ZERO, so dropped reads and the padding of short columns add nothing — no test per element.  Normal:
// the one row of probabilityCorrectIgnoringAlignment (:48-50), and a zero row for its dropped reads.
constexpr int kSomTabRows = 64;
struct RowTables {
  uint8_t row_mapq[kSomTabRows];  // mapping quality of the tumor table's row r
  int32_t n_rows;                 // rows in use (< kMaxRank)
  int32_t n_keep_tumor;           // ranks below this pass the mapq filter ...
  int32_t n_keep_normal;          // ... same for the normal sample's ranks
  int32_t pad_;
};
struct SomSmem {
  double2 tumor[kSomTabRows * 64];
  double2 normal[2 * 64];
  uint4 hdr[kTileWords];   // the tumor sample's word headers of this CTA's tile
};

__device__ __forceinline__ double2 lds_double2(uint32_t shared_addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(shared_addr));
  return v;
}

// L2 prefetch of everything gather_rows streams for one word (its header is known already): the kernel handles a word in about
// the time of two DRAM round trips, so the word after this one is requested while this one is summed
__device__ __forceinline__ void prefetch_rows(const DevReads& R, const uint32_t word, const uint4 wh) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t blocks = ((wh.y & 0xFFFFFFu) + 7u) >> 3, groups = ((wh.w & 0xFFFFFFu) + 3u) >> 2;
  const uint4* cp = reinterpret_cast<const uint4*>(R.q_cols) + (size_t)wh.x * 32 + lane;
  if ((lane & 1u) == 0u)  // (one request per 32-byte sector)
    for (uint32_t b = 0; b < blocks; ++b) asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + (size_t)b * 32));
  if (lane < 2u) asm volatile("prefetch.global.L2 [%0];" ::"l"(R.q_depth + (size_t)word * 32 + lane * 16));
  if (lane < groups) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(R.q_groups + wh.z + lane));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(R.q_rows + ((size_t)wh.z + lane) * 32));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(R.q_rows + ((size_t)wh.z + lane) * 32 + 16));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(R.q_rows + ((size_t)wh.z + lane) * 32 + 24));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(R.q_rows + ((size_t)wh.z + lane) * 32 + 8));
  }
}

template <bool TUMOR>
__device__ __forceinline__ void gather_rows(const DevReads& R, const uint32_t word, const uint4 wh, const int rcode, const bool std_ref, const SomParams& prm,
                                            const uint32_t smem_table /* shared address of SomSmem */, const RowTables& rt,
                                            const double* __restrict__ tables, LaneAcc& A) {
  const uint32_t lane = threadIdx.x & 31u;
  acc_clear(A);
  const uint32_t n_cols = wh.y & 0xFFFFFFu, n_rows = wh.w & 0xFFFFFFu, max_rank = wh.w >> 24;
  const uint32_t depth_all = R.q_depth[(size_t)word * 32 + lane];
  const double2* __restrict__ gtab = reinterpret_cast<const double2*>(tables + (TUMOR ? kTabT : kTabN));
  const uint32_t tab = TUMOR ? smem_table : smem_table + (uint32_t)(kSomTabRows * 64 * sizeof(double2));
  const uint32_t n_keep = (uint32_t)(TUMOR ? rt.n_keep_tumor : rt.n_keep_normal);
  const uint32_t min_mapq = prm.min_mapq > 0 ? (uint32_t)prm.min_mapq : 0u;
  const bool fma = prm.filter_multi_allelic != 0;
  double sr1 = 0.0;
  int any = depth_all != 0u ? 1 : 0;
  uint32_t dropped = 0;
  unsigned long long cnt_packed = 0;  // mismatching elements: four 16-bit fields, one per base code
  const uint32_t rc_eff = std_ref ? (uint32_t)rcode : 4u;
  // an element that does not carry the reference base (rare)
  auto non_ref = [&](const uint32_t cls, const uint32_t rank, const double2 l) {
    if (rank >= n_keep) return;  // dropped by the mapq filter (its table row is zero: nothing was added to T0 either)
    const uint32_t code = cls ^ (uint32_t)rcode;
    cnt_packed += 1ull << (16 * code);
    if (code == 0u) { A.s1[0] += l.x; A.s0[0] += l.y; }
    else if (code == 1u) { A.s1[1] += l.x; A.s0[1] += l.y; }
    else if (code == 2u) { A.s1[2] += l.x; A.s0[2] += l.y; }
    else { A.s1[3] += l.x; A.s0[3] += l.y; }
  };
  if (n_cols) {
    const uint32_t blocks = (n_cols + 7u) >> 3;  // eight columns = 16 bytes per locus per load
    const uint4* __restrict__ cp = reinterpret_cast<const uint4*>(R.q_cols) + (size_t)wh.x * 32 + lane;
    const bool check = max_rank >= n_keep;  // (warp-uniform) some read of this word fails the mapq filter
    const uint32_t nostd = std_ref ? 0u : 1u;  // (no reference class at a locus whose reference base is not A/C/G/T)
    // byte offset of an element's table entry: bits 15..4 of the element are (rank, quality).  The normal sample's table has
    // two rows: the probabilities ignoring the mapping quality, and zeros for dropped reads and the sentinel.
    auto entry = [&](const uint32_t e) -> uint32_t {
      return TUMOR ? (e & 0xFFF0u) : (e & 0x03F0u) + (((e >> 10) & 63u) < n_keep ? 0u : 1024u);
    };
    // Elements of a mismatch class sit in the word's last blocks only (k_expand_rows fills a column with reference-class
    // elements from the front, the others from the back): every block before `first_rare` is summed without a class test.
    const uint32_t rare_blocks = wh.y >> 24;
    const uint32_t first_rare = rare_blocks >= kRareAll ? 0u : blocks - min(rare_blocks, blocks);
    uint4 v = __ldg(cp), v_next = v;
    if (blocks > 1) v_next = __ldg(cp + 32);
    for (uint32_t p = 0; p < blocks; ++p) {
      const uint4 v_now = v;
      v = v_next;
      if (p + 2 < blocks) v_next = __ldg(cp + (size_t)(p + 2) * 32);  // two blocks are on their way while this one is summed
      const uint32_t w4[4] = {v_now.x, v_now.y, v_now.z, v_now.w};
      const bool tail = p >= first_rare;  // (warp-uniform)
      if (nostd) {  // (rare lanes) every element is of a mismatch class
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
          const uint32_t wj = j < 2 ? w4[0] : j < 4 ? w4[1] : j < 6 ? w4[2] : w4[3];
          const uint32_t e = ((j & 1) ? (wj >> 16) : wj) & 0xFFFFu;
          const double2 l = lds_double2(tab + entry(e));
          A.t0 += l.y;
          non_ref(e & 3u, e >> 10, l);
        }
      } else if (!tail) {
        // the hot loop: per element one table look-up and two additions (every element here carries the reference base or is
        // the sentinel, whose table row is zero)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t e = (j & 1) ? (w4[j >> 1] >> 16) : w4[j >> 1];  // (even elements: the upper half word is masked off below)
          const double2 l = lds_double2(tab + entry(e));
          A.t0 += l.y;
          sr1 += l.x;
        }
      } else {
        // the word's last block(s): the second addition only for elements of class 0, the others are handled below
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t e = (j & 1) ? (w4[j >> 1] >> 16) : w4[j >> 1];
          const double2 l = lds_double2(tab + entry(e));
          A.t0 += l.y;
          asm("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %2, 0;\n\t@p add.f64 %0, %0, %1;\n\t}" : "+d"(sr1) : "d"(l.x), "r"(e & 3u));  // a predicated add, no select
        }
        if (((w4[0] | w4[1]) | (w4[2] | w4[3])) & 0x00030003u) {  // (divergent) some element of this lane is not of class 0
          uint32_t rare = 0;  // bit j: element j's class is not 0
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t m = (w4[k] | (w4[k] >> 1)) & 0x00010001u;
            rare |= ((m | (m >> 15)) & 3u) << (2 * k);
          }
          while (rare) {
            const int j = __ffs(rare) - 1;
            rare &= rare - 1;
            const uint32_t wj = j < 2 ? w4[0] : j < 4 ? w4[1] : j < 6 ? w4[2] : w4[3];
            const uint32_t e = ((j & 1) ? (wj >> 16) : wj) & 0xFFFFu;
            non_ref(e & 3u, e >> 10, lds_double2(tab + entry(e)));
          }
        }
      }
      if (check) {  // (rare words) elements dropped by the mapq filter: n_keep <= rank < the sentinel's.  Two elements per word:
        // bit 6 of (rank + 64 - n_keep) says rank >= n_keep, bit 6 of (rank + 1) says rank == 63
        const uint32_t ge = (64u - n_keep) * 0x00010001u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t r2 = (w4[k] >> 10) & 0x003F003Fu;
          dropped += __popc((r2 + ge) & ~(r2 + 0x00010001u) & 0x00400040u);
        }
      }
    }
  }
  // the general rows (reads with insertions / deletions / skips / non-ACGT bases / wide qualities): a class row + a quality row each
  if (n_rows) {
    const uint32_t groups = (n_rows + 3u) >> 2;
    const uint4* __restrict__ gh = R.q_groups + wh.z;
    const uint32_t* __restrict__ rows = R.q_rows + (size_t)wh.z * 32 + lane;
    for (uint32_t g = 0; g < groups; ++g) {
      const uint4 hd = __ldg(gh + g);
      const uint32_t v = __ldg(rows + (size_t)g * 32);
      const uint32_t hs[4] = {hd.x, hd.y, hd.z, hd.w};
#pragma unroll
      for (int k = 0; k < 4; k += 2) {
        const uint32_t h = hs[k];
        if (((h >> 8) & 3u) != kRowGeneral) continue;  // (padding of the last group)
        const uint32_t mapq = h & 0xFFu, b = (v >> (8 * k)) & 0xFFu, q = (v >> (8 * (k + 1))) & 0xFFu;
        const bool keep = mapq >= min_mapq;
        if (b == kElemNone) continue;
        any = 1;
        if (!keep && !fma) continue;
        if (b == kElemHard) {
          A.other += 1;
          A.hard += 1;
        } else if (b == kElemOther) {
          A.other += 1;  // insertion / deletion / clipped / non-ACGT element: which allele it carries is the exact kernel's
          if (keep) {    // job, but its likelihood terms bound every genotype it can be part of (ref_leads_despite_others)
            const double2 l = __ldg(&gtab[TUMOR ? (mapq << 8) + q : q]);
            A.o0 += l.y;
            A.ohet += fmax(0.0, l.y);
            A.ohom += fmax(l.x, l.y);
          }
        } else if (keep) {  // a plain base of a general read
          const uint32_t code = b & 3u;
          const double2 l = __ldg(&gtab[TUMOR ? (mapq << 8) + q : q]);
          A.t0 += l.y;
          if (code == rc_eff) {
            sr1 += l.x;
            dropped -= 1u;  // (counted as a kept element: the reference count below is kept elements - the other classes)
          } else {
            cnt_packed += 1ull << (16 * code);
            if (code == 0u) { A.s1[0] += l.x; A.s0[0] += l.y; }
            else if (code == 1u) { A.s1[1] += l.x; A.s0[1] += l.y; }
            else if (code == 2u) { A.s1[2] += l.x; A.s0[2] += l.y; }
            else { A.s1[3] += l.x; A.s0[3] += l.y; }
            dropped -= 1u;
          }
        }
      }
    }
  }
  // kept plain elements = all of the locus - the dropped ones (+ the general reads' plain bases, entered as negative drops)
  const int kept = (int)depth_all - (int)dropped;
  int others = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) others += (int)((cnt_packed >> (16 * k)) & 0xFFFFu);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    A.cnt[k] = (int)((cnt_packed >> (16 * k)) & 0xFFFFu) + ((uint32_t)k == rc_eff ? kept - others : 0);
    A.depth += A.cnt[k];
  }
  if (!std_ref) A.depth = kept;  // (no reference class: every element sits in a mismatch class; only the totals matter, the locus is deferred)
  A.any = any;
  A.ref_depth = !std_ref ? 0 : rcode == 0 ? A.cnt[0] : rcode == 1 ? A.cnt[1] : rcode == 2 ? A.cnt[2] : A.cnt[3];  // (no dynamic index: A stays in registers)
#pragma unroll
  for (int k = 0; k < 4; ++k) A.seen |= A.cnt[k] > 0 ? (1u << k) : 0u;
  if (n_cols + n_rows > 0xFFFFu) {  // the packed counters (and the u16 depths) hold 16 bits: the exact kernel decides this
    A.other += 1;                   // word's loci, whatever the bounds computed from wrapped counts would say
    A.hard += 1;
  }
  double sr0 = A.t0;  // S0 of the reference class = T0 - the other classes' S0
#pragma unroll
  for (int k = 0; k < 4; ++k) sr0 -= A.s0[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool is = std_ref && k == rcode;
    A.s1[k] += is ? sr1 : 0.0;
    A.s0[k] += is ? sr0 : 0.0;
  }
}

constexpr int kSomThreads = 256;

// ROWS: the samples' pileup elements come from their row stores (guac_rows.cuh); otherwise every word walks its candidate
// reads (gather_sample) — the path of stores packed without rows, and the cross-check the parity tests run.
template <bool ROWS>
__global__ void __launch_bounds__(kSomThreads, GUAC_SOM_MINB) k_somatic(DevReads RT, DevReads RN, const TileDesc* __restrict__ tiles, uint32_t n_tiles,
                                                        SomParams prm, const double* __restrict__ tables, RowTables rt, SomOut out) {
  extern __shared__ __align__(16) unsigned char som_smem_raw[];
  const uint32_t smem_table = (uint32_t)__cvta_generic_to_shared(som_smem_raw);
  if (ROWS) {  // the CTA's copy of the table: kept mapping qualities' rows, zero rows for everything else
    SomSmem& Tw = *reinterpret_cast<SomSmem*>(som_smem_raw);
    const double2* gt = reinterpret_cast<const double2*>(tables + kTabT);
    const double2* gn = reinterpret_cast<const double2*>(tables + kTabN);
    const int live = min(rt.n_rows, rt.n_keep_tumor);
    for (int i = threadIdx.x; i < kSomTabRows * 64; i += kSomThreads)
      Tw.tumor[i] = (i >> 6) < live ? gt[((int)rt.row_mapq[i >> 6] << 8) + (i & 63)] : make_double2(0.0, 0.0);
    for (int i = threadIdx.x; i < 128; i += kSomThreads) Tw.normal[i] = i < 64 ? gn[i] : make_double2(0.0, 0.0);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint4* const tile_hdr = reinterpret_cast<const SomSmem*>(som_smem_raw)->hdr;
  uint32_t n_visited = 0;
  // a CTA stays for many tiles (the grid is what fits the device at once): its copy of the table is built once
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
  const TileDesc td = tiles[tile];
  const ContigInfo ciT = RT.contigs[td.contig], ciN = RN.contigs[td.contig];
  if (ROWS) {
    SomSmem& Tw = *reinterpret_cast<SomSmem*>(som_smem_raw);
    __syncthreads();  // (every warp is done with the previous tile's headers; the first time: the table is complete)
    for (int i = threadIdx.x; i < kTileWords; i += kSomThreads)
      Tw.hdr[i] = td.word0 + i < ciT.n_words ? RT.q_hdr[ciT.word_off + (uint32_t)(td.word0 + i)] : make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
  }
  for (int wi = warp; wi < kTileWords; wi += kSomThreads / 32) {
    const int w = td.word0 + wi;
    const int span_lo = w << 5, x = span_lo + lane;
    if (span_lo >= td.locus_end || span_lo + 32 <= td.locus_begin) continue;
    const bool in_req = x >= td.locus_begin && x < td.locus_end;
    // each sample has its own MD-derived reference track (its pileup's referenceBase comes from its own reads)
    uint32_t tl = 0, th = 0, ts = 0, nl = 0, nh = 0, ns = 0;
    if (w < ciT.n_words) { tl = RT.trk_lo[ciT.word_off + w]; th = RT.trk_hi[ciT.word_off + w]; ts = RT.trk_std[ciT.word_off + w]; }
    if (w < ciN.n_words) { nl = RN.trk_lo[ciN.word_off + w]; nh = RN.trk_hi[ciN.word_off + w]; ns = RN.trk_std[ciN.word_off + w]; }
    const int rcT = (int)((tl >> lane) & 1u) | ((int)((th >> lane) & 1u) << 1), rcN = (int)((nl >> lane) & 1u) | ((int)((nh >> lane) & 1u) << 1);
    const bool stdT = (ts >> lane) & 1u, stdN = (ns >> lane) & 1u;
    LaneAcc AT, AN;
    if (ROWS) {
      const int wn = wi + kSomThreads / 32;  // this warp's next word
      if (wn < kTileWords && td.word0 + wn < ciT.n_words) prefetch_rows(RT, ciT.word_off + (uint32_t)(td.word0 + wn), tile_hdr[wn]);
      if (w < ciT.n_words) gather_rows<true>(RT, ciT.word_off + (uint32_t)w, tile_hdr[wi], rcT, stdT, prm, smem_table, rt, tables, AT);
      else acc_clear(AT);
    } else {
      gather_sample<true>(RT, td.contig, span_lo, x, rcT, stdT, prm, tables, AT);
    }
    // The reference looks at the normal sample only where the tumor's most likely genotype holds a variant allele: decide
    // the tumor half per lane first and walk the normal reads of this word only if some lane still needs them.
    const int covT = AT.depth + AT.other;  // (filtered when no multi-allelic filter is on)
    bool tumor_exact = AT.hard > 0 || !stdT || prm.filter_multi_allelic;
    const bool tumor_all_match = AT.ref_depth == AT.depth && AT.other == 0 && stdT && !prm.filter_multi_allelic;
    bool need_normal = false, variant = false;
    int c1 = 0, c2 = 0;
    double tumor_l = 0.0;
    AlleleView avT{RT, code_base(rcT)};
    if (in_req) {
      if (AT.any == 0) need_normal = true;  // visited iff the normal sample has reads here
      else if (covT > 0 && !tumor_all_match) {
        if (!tumor_exact && AT.other > 0) tumor_exact = !ref_leads_despite_others(AT, rcT);
        if (tumor_exact) need_normal = true;
        else if (AT.other == 0 && !ref_leads_despite_others(AT, rcT)) {
          // (the same exact bound first, on the raw sums: at most loci with a mismatching read or two the homozygous-reference
          // genotype leads every other one clearly, and the ten likelihoods below are never formed)
          SnvAlleles sT;
          snv_compact(AT, sT);
          variant = snv_tumor_most_likely(sT, AT.ref_depth, rcT, prm, &c1, &c2, &tumor_l);
          need_normal = variant;
        }
      }
    }
    if (in_req && AT.any > 0) ++n_visited;
    if (!__any_sync(0xFFFFFFFFu, need_normal)) continue;
    if (ROWS) {
      if (w < ciN.n_words) gather_rows<false>(RN, ciN.word_off + (uint32_t)w, RN.q_hdr[ciN.word_off + (uint32_t)w], rcN, stdN, prm, smem_table, rt, tables, AN);
      else acc_clear(AN);
    } else {
      gather_sample<false>(RN, td.contig, span_lo, x, rcN, stdN, prm, tables, AN);
    }
    if (!need_normal) continue;
    if (AT.any == 0) {
      if (AN.any > 0 || !prm.skip_empty) ++n_visited;
      continue;
    }
    const int covN = AN.depth + AN.other;
    if (covN == 0) continue;
    // anything but A/C/G/T matches / mismatches over a standard reference base goes to the exact kernel
    if (tumor_exact || AN.other > 0 || !stdN) {
      const uint32_t s = (uint32_t)atomicAdd(&out.counters[2], 1ull);
      if (s < out.cap_slow) out.slow[s] = SlowLocus{td.contig, x};
      continue;
    }
    if (AN.depth == 0 || AN.depth > prm.max_read_depth) continue;
    SnvAlleles sN;
    snv_compact(AN, sN);
    const double normal_variants_total = snv_normal_variants_total(sN, rcN);
    emit_somatic(avT, snv_entry(c1, 0), snv_entry(c2, 0), tumor_l, normal_variants_total, td.contig, x, prm, out);
  }
  }  // tiles
  for (int o = 16; o; o >>= 1) n_visited += __shfl_xor_sync(0xFFFFFFFFu, n_visited, o);
  if (lane == 0 && n_visited) atomicAdd(&out.counters[3], (unsigned long long)n_visited);
}

// ---- K_somatic_exact: warp per locus ----------------------------------------------------------------------------------------------
struct ExactSmem {
  SomAllele tab[2][kSomTab];
  GenotypeScratch G;
  uint32_t ring[64];
};

template <bool TUMOR>
__device__ bool exact_sample(const DevReads& R, int contig, int locus, const SomParams& prm, const double* __restrict__ tables,
                             SomAllele* tab, uint32_t* ring, SampleStats& st, uint8_t* ref_base_out, DevError* err) {
  const int lane = threadIdx.x & 31;
  const ContigInfo ci = R.contigs[contig];
  bool std_ref = false;
  uint8_t ref_base = 'N';
  if (locus < ci.length) ref_base = reference_base_of(R, ci, contig, locus, &std_ref);
  *ref_base_out = ref_base;
  AlleleView av{R, ref_base};
  int na = 0, n_unf = 0;  // table entries; entries [0, na) carry filtered sums, distinct_unfiltered counts all alleles seen
  int depth = 0, ref_depth = 0, t0inf = 0;
  double t0 = 0.0, s1c[4] = {0, 0, 0, 0}, s0c[4] = {0, 0, 0, 0};
  int cntc[4] = {0, 0, 0, 0};
  uint32_t seen = 0;
  const double2* __restrict__ lut = reinterpret_cast<const double2*>(tables + (TUMOR ? kTabT : kTabN));
  uint32_t first = 0xFFFFFFFFu, last = 0;
  if (locus < ci.length) {
    const int g = locus >> kGranuleShift;
    first = R.gran_first[ci.gran_off + g];
    last = R.gran_last[ci.gran_off + g];
  }
  OverlapWalker walk(R, ring, first, last, locus);
  uint32_t r;
  ReadRec rec;
  bool valid;
  while (walk.next(R, r, rec, valid)) {
    const int mapq = (int)(rec.info >> kInfoMapqShift);
    const bool keep = valid && (!(prm.min_mapq > 0) || mapq >= prm.min_mapq);
    Elem e;
    e.kind = kNone; e.base = 0; e.len = 0; e.ptr = 0; e.qual = 0;
    int rc = 0;
    if (valid && (keep || prm.filter_multi_allelic)) {
      rc = classify(R, r, locus, ref_base, e);
      if (rc == 0 && e.kind == kNone) rc = GUAC_ERR_INVALID_CIGAR;
    }
    if (__any_sync(0xFFFFFFFFu, rc != 0)) {
      if (rc) report_error(err, rc, r);
      return false;
    }
    const bool have = e.kind != kNone;
    const bool snv = have && (e.kind == kMatch || e.kind == kMismatch) && is_std_base(e.base);
    double l1 = 0.0, l0 = 0.0;
    if (have && keep) {
      const int q = e.qual & 255;
      const double2 l = __ldg(&lut[TUMOR ? (mapq << 8) + q : q]);
      l1 = l.x;
      l0 = l.y;
      depth += 1;
      ref_depth += e.kind == kMatch ? 1 : 0;
      if (l0 == -1.0 / 0.0) t0inf += 1;
      else t0 += l0;
    }
    if (snv) {
      const int code = (int)base_code(e.base);
      seen |= 1u << code;
      if (keep) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool is = code == k;
          cntc[k] += is ? 1 : 0;
          s1c[k] += is ? l1 : 0.0;
          s0c[k] += is ? l0 : 0.0;
        }
      }
    }
    uint32_t mask = __ballot_sync(0xFFFFFFFFu, have && !snv);
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      Elem s;
      s.kind = __shfl_sync(0xFFFFFFFFu, e.kind, src);
      s.len = __shfl_sync(0xFFFFFFFFu, e.len, src);
      s.base = (uint8_t)__shfl_sync(0xFFFFFFFFu, (int)e.base, src);
      s.ptr = ((uint64_t)__shfl_sync(0xFFFFFFFFu, (uint32_t)(e.ptr >> 32), src) << 32) | __shfl_sync(0xFFFFFFFFu, (uint32_t)e.ptr, src);
      s.qual = 0;
      const bool s_keep = __shfl_sync(0xFFFFFFFFu, (int)keep, src) != 0;
      const double sl1 = __shfl_sync(0xFFFFFFFFu, l1, src), sl0 = __shfl_sync(0xFFFFFFFFu, l0, src);
      int status = 0;  // 0 updated, 1 inserted, 2 table full
      if (lane == 0) {
        int found = -1;
        for (int k = 0; k < na; ++k)
          if (av.same(as_entry(tab[k]), s)) { found = k; break; }
        if (found < 0) {
          if (na == kSomTab) status = 2;
          else {
            found = na;
            tab[na].kind = (s.kind == kMatch || s.kind == kMismatch) ? 0 : s.kind;
            tab[na].len = s.len; tab[na].ptr = s.ptr; tab[na].base = s.base; tab[na].count = 0; tab[na].n0inf = 0;
            tab[na].s1 = 0.0; tab[na].s0 = 0.0;
            status = 1;
          }
        }
        if (status != 2 && s_keep) {
          tab[found].count += 1;
          tab[found].s1 += sl1;
          if (sl0 == -1.0 / 0.0) tab[found].n0inf += 1;
          else tab[found].s0 += sl0;
        }
      }
      status = __shfl_sync(0xFFFFFFFFu, status, 0);
      if (status == 2) {
        if (lane == 0) report_error(err, GUAC_ERR_UNSUPPORTED, ((unsigned long long)contig << 32) | (uint32_t)locus);
        return false;
      }
      if (status == 1) { ++na; ++n_unf; }
      __syncwarp();
    }
  }
  for (int o = 16; o; o >>= 1) {
    depth += __shfl_xor_sync(0xFFFFFFFFu, depth, o);
    ref_depth += __shfl_xor_sync(0xFFFFFFFFu, ref_depth, o);
    t0inf += __shfl_xor_sync(0xFFFFFFFFu, t0inf, o);
    t0 += __shfl_xor_sync(0xFFFFFFFFu, t0, o);
    seen |= __shfl_xor_sync(0xFFFFFFFFu, seen, o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      cntc[k] += __shfl_xor_sync(0xFFFFFFFFu, cntc[k], o);
      s1c[k] += __shfl_xor_sync(0xFFFFFFFFu, s1c[k], o);
      s0c[k] += __shfl_xor_sync(0xFFFFFFFFu, s0c[k], o);
    }
  }
  bool ok = true;
  if (lane == 0) {
    // drop entries that only unfiltered (low-mapq) elements carried, then add the A/C/G/T alleles from their counters
    int m = 0;
    for (int k = 0; k < na; ++k)
      if (tab[k].count > 0) tab[m++] = tab[k];
    na = m;
    for (int k = 0; k < 4; ++k)
      if (cntc[k] > 0) {
        if (na == kSomTab) { ok = false; break; }
        tab[na].kind = 0; tab[na].len = 1; tab[na].ptr = 0; tab[na].base = code_base(k); tab[na].count = cntc[k];
        tab[na].n0inf = 0;  // (an A/C/G/T element's quality is a base quality <= 127: s < 1)
        tab[na].s1 = s1c[k]; tab[na].s0 = s0c[k];
        ++na;
      }
    if (ok) sort_alleles(av, tab, na);
    else report_error(err, GUAC_ERR_UNSUPPORTED, ((unsigned long long)contig << 32) | (uint32_t)locus);
  }
  ok = __shfl_sync(0xFFFFFFFFu, (int)ok, 0) != 0;
  st.depth = depth;
  st.ref_depth = ref_depth;
  st.n_alleles = __shfl_sync(0xFFFFFFFFu, na, 0);
  st.distinct_unfiltered = n_unf + __popc(seen);
  st.t0inf = t0inf;
  st.t0 = t0;
  __syncwarp();
  return ok;
}

constexpr int kSomExactWarps = 2;

__global__ void __launch_bounds__(kSomExactWarps * 32) k_somatic_exact(DevReads RT, DevReads RN, const SlowLocus* __restrict__ loci, SomParams prm,
                                                                       const double* __restrict__ tables, SomOut out) {
  __shared__ ExactSmem sm[kSomExactWarps];
  const int lane = threadIdx.x & 31;
  ExactSmem& S = sm[threadIdx.x >> 5];
  const uint32_t n_loci = (uint32_t)min(out.counters[2], (unsigned long long)out.cap_slow);
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t t = warp; t < n_loci; t += n_warps) {
    const int contig = loci[t].contig, locus = loci[t].locus;
    SampleStats sT, sN;
    uint8_t refT, refN;
    bool ok = exact_sample<true>(RT, contig, locus, prm, tables, S.tab[0], S.ring, sT, &refT, out.err);
    ok = ok && exact_sample<false>(RN, contig, locus, prm, tables, S.tab[1], S.ring, sN, &refN, out.err);
    if (ok && lane == 0) {
      // MultiAllelicPileupFilter (before the mapq filter): more than two distinct alleles empty the pileup
      if (prm.filter_multi_allelic) {
        if (sT.distinct_unfiltered > 2) sT.depth = 0;
        if (sN.distinct_unfiltered > 2) sN.depth = 0;
      }
      AlleleView avT{RT, refT}, avN{RN, refN};
      decide_somatic(RT, avT, avN, S.tab[0], sT, S.tab[1], sN, contig, locus, prm, out, S.G);
    }
    __syncwarp();
  }
}

// ---- K_evidence: warp per record ------------------------------------------------------------------------------------------------------
// does element e of read set R carry the allele (ref bytes, alt bytes)?
__device__ bool elem_is_allele(const DevReads& R, const Elem& e, uint8_t ref_base, const uint8_t* ref, int ref_len, const uint8_t* alt, int alt_len) {
  AlleleView av{R, ref_base};
  AlleleEntry a;
  a.kind = (e.kind == kMatch || e.kind == kMismatch) ? 0 : e.kind;
  a.len = e.len; a.ptr = e.ptr; a.base = e.base; a.count = 1;
  if (av.ref_len(a) != ref_len || av.alt_len(a) != alt_len) return false;
  for (int i = 0; i < ref_len; ++i) if (av.ref_at(a, i) != ref[i]) return false;
  for (int i = 0; i < alt_len; ++i) if (av.alt_at(a, i) != alt[i]) return false;
  return true;
}

// The statistics AlleleEvidence needs are over small integers (mapping quality 0..255, element quality 0..255, mismatches per
// read): one histogram each in shared memory gives exact medians at any depth (breeze median = sort + middle element(s))
// and the means as integer sums (breeze's running mean agrees with sum / n to an ulp).
constexpr int kNmBins = 512;   // reads with more MD mismatches than this are not supported by K_evidence (small histograms:
                               // more warps per SM, and the kernel lives on memory latency)

struct EvidenceSmem {
  int hmq[256], hbq[256], hnm[kNmBins];
  uint32_t ring[64];
};

// value of the rank-th smallest element (0-based) of a histogram; warp-cooperative, nbins a multiple of 32
__device__ int hist_select(const int* h, int nbins, int rank) {
  const int lane = threadIdx.x & 31, chunk = nbins >> 5;
  int s = 0;
  for (int k = 0; k < chunk; ++k) s += h[lane * chunk + k];
  int incl = s;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += t;
  }
  const int excl = incl - s;
  int v = 0;
  const bool mine = excl <= rank && rank < incl;
  if (mine) {
    int acc = excl;
    for (int k = 0; k < chunk; ++k) {
      acc += h[lane * chunk + k];
      if (acc > rank) { v = lane * chunk + k; break; }
    }
  }
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, mine);
  return __shfl_sync(0xFFFFFFFFu, v, m ? __ffs(m) - 1 : 0);
}

__device__ void evidence_sample(const DevReads& R, int contig, int locus, const SomParams& prm, const uint8_t* ref, int ref_len,
                                const uint8_t* alt, int alt_len, guac_allele_evidence& ev, EvidenceSmem& S, DevError* err) {
  const int lane = threadIdx.x & 31;
  const ContigInfo ci = R.contigs[contig];
  bool std_ref = false;
  uint8_t ref_base = 'N';
  if (locus < ci.length) ref_base = reference_base_of(R, ci, contig, locus, &std_ref);
  for (int i = lane; i < 256; i += 32) { S.hmq[i] = 0; S.hbq[i] = 0; }
  for (int i = lane; i < kNmBins; i += 32) S.hnm[i] = 0;
  __syncwarp();
  int depth = 0, fwd = 0, adepth = 0, afwd = 0;
  uint32_t first = 0xFFFFFFFFu, last = 0;
  if (locus < ci.length) {
    const int g = locus >> kGranuleShift;
    first = R.gran_first[ci.gran_off + g];
    last = R.gran_last[ci.gran_off + g];
  }
  OverlapWalker walk(R, S.ring, first, last, locus);
  uint32_t r;
  ReadRec rec;
  bool valid;
  while (walk.next(R, r, rec, valid)) {
    const int mapq = (int)(rec.info >> kInfoMapqShift);
    valid = valid && (!(prm.min_mapq > 0) || mapq >= prm.min_mapq);
    Elem e;
    e.kind = kNone; e.base = 0; e.len = 0; e.ptr = 0; e.qual = 0;
    if (valid && classify(R, r, locus, ref_base, e) != 0) valid = false;
    if (valid && e.kind != kNone) {
      ++depth;
      if (rec.info & kInfoPositive) ++fwd;
      if (elem_is_allele(R, e, ref_base, ref, ref_len, alt, alt_len)) {
        ++adepth;
        if (rec.info & kInfoPositive) ++afwd;
        const int nm = (int)R.nm[r];
        if (nm >= kNmBins) report_error(err, GUAC_ERR_UNSUPPORTED, r);
        atomicAdd(&S.hmq[mapq & 255], 1);
        atomicAdd(&S.hbq[e.qual & 255], 1);
        atomicAdd(&S.hnm[min(nm, kNmBins - 1)], 1);
      }
    }
  }
  for (int o = 16; o; o >>= 1) {
    depth += __shfl_xor_sync(0xFFFFFFFFu, depth, o);
    fwd += __shfl_xor_sync(0xFFFFFFFFu, fwd, o);
    adepth += __shfl_xor_sync(0xFFFFFFFFu, adepth, o);
    afwd += __shfl_xor_sync(0xFFFFFFFFu, afwd, o);
  }
  __syncwarp();
  const int n = adepth;
  double mean_mq = 0, mean_bq = 0, med_mq = 0, med_bq = 0, med_nm = 0;
  if (n > 0) {
    long long sm = 0, sb = 0;
    for (int i = lane; i < 256; i += 32) { sm += (long long)i * S.hmq[i]; sb += (long long)i * S.hbq[i]; }
    for (int o = 16; o; o >>= 1) { sm += __shfl_xor_sync(0xFFFFFFFFu, sm, o); sb += __shfl_xor_sync(0xFFFFFFFFu, sb, o); }
    mean_mq = (double)sm / (double)n;
    mean_bq = (double)sb / (double)n;
    if (n % 2 == 1) {
      med_mq = (double)hist_select(S.hmq, 256, (n - 1) / 2);
      med_bq = (double)hist_select(S.hbq, 256, (n - 1) / 2);
      med_nm = (double)hist_select(S.hnm, kNmBins, (n - 1) / 2);
    } else {  // mean of the middle two; DenseVector[Int] (mismatches) in integer arithmetic
      med_mq = ((double)hist_select(S.hmq, 256, n / 2 - 1) + (double)hist_select(S.hmq, 256, n / 2)) / 2;
      med_bq = ((double)hist_select(S.hbq, 256, n / 2 - 1) + (double)hist_select(S.hbq, 256, n / 2)) / 2;
      med_nm = (double)((hist_select(S.hnm, kNmBins, n / 2 - 1) + hist_select(S.hnm, kNmBins, n / 2)) / 2);
    }
  }
  if (lane == 0) {
    ev.read_depth = depth;
    ev.allele_read_depth = adepth;
    ev.forward_depth = fwd;
    ev.allele_forward_depth = afwd;
    if (n == 0) {
      const double nan = __longlong_as_double(0x7FF8000000000000ll);
      ev.mean_mapping_quality = ev.median_mapping_quality = ev.mean_base_quality = ev.median_base_quality = ev.median_mismatches_per_read = nan;
    } else {
      ev.mean_mapping_quality = mean_mq;
      ev.mean_base_quality = mean_bq;
      ev.median_mapping_quality = med_mq;
      ev.median_base_quality = med_bq;
      ev.median_mismatches_per_read = med_nm;
    }
  }
  __syncwarp();
}

constexpr int kEvidenceWarps = 4;

__global__ void __launch_bounds__(kEvidenceWarps * 32, 8) k_evidence(DevReads RT, DevReads RN, SomParams prm, SomOut out) {
  __shared__ EvidenceSmem sm[kEvidenceWarps];
  EvidenceSmem& S = sm[threadIdx.x >> 5];
  const uint32_t n_rec = (uint32_t)min(out.counters[0], (unsigned long long)out.cap_rec);
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t t = warp; t < n_rec; t += n_warps) {
    guac_somatic_record& r = out.rec[t];
    if ((unsigned long long)r.alt_off + r.alt_len > out.cap_pool) continue;  // pool overflow: the host reruns
    const uint8_t* ref = out.pool + r.ref_off;
    const uint8_t* alt = out.pool + r.alt_off;
    guac_allele_evidence evT = r.tumor, evN = r.normal;
    // tumorVariantEvidence = AlleleEvidence(L, allele, filteredTumor); normalReferenceEvidence = AlleleEvidence(1 - total,
    // Allele(allele.refBases, allele.refBases), filteredNormal)
    evidence_sample(RT, r.contig, (int)r.start, prm, ref, r.ref_len, alt, r.alt_len, evT, S, out.err);
    evidence_sample(RN, r.contig, (int)r.start, prm, ref, r.ref_len, ref, r.ref_len, evN, S, out.err);
    if ((threadIdx.x & 31) == 0) {
      r.tumor = evT;
      r.normal = evN;
    }
    __syncwarp();
  }
}

}  // namespace guac

// ---- post-call filters in the epilogue: SomaticGenotypeFilter over the finished records, on the device, so that
// filtered records never cross PCIe (filters/SomaticGenotypeFilter.scala:30-335, called from SomaticStandardCaller.scala:125-151)
namespace guac {

// SomaticReadDepthFilter (:69-75, upper bound exclusive, filters/GenotypeFilter.scala:63), SomaticAlternateReadDepthFilter
// (:107-110), SomaticVAFFilter (:142-145, Float VAF), SomaticMinimumLikelihoodFilter (:38-41), and in the full overload
// SomaticLogOddsFilter (:176-179), SomaticAverageMappingQualityFilter (:210-214), SomaticAverageBaseQualityFilter (:194-198 —
// compares the MAPPING quality, kept as is), SomaticMedianMismatchFilter (:228-231)
__host__ __device__ inline bool somatic_filter_keep(const guac_somatic_record& g, const guac_somatic_filter_params& p) {
  const guac_allele_evidence& t = g.tumor;
  const guac_allele_evidence& nrm = g.normal;
  bool ok = t.read_depth >= p.min_tumor_read_depth && t.read_depth < p.max_tumor_read_depth && nrm.read_depth >= p.min_normal_read_depth &&
            nrm.read_depth < 0x7FFFFFFF;
  if (p.min_tumor_alternate_read_depth > 0) ok = ok && t.allele_read_depth >= p.min_tumor_alternate_read_depth;
  const float vaf = (float)t.allele_read_depth / (float)t.read_depth;  // AlleleEvidence.variantAlleleFrequency (Float)
  ok = ok && ((double)vaf * 100.0 > (double)p.min_vaf);
  ok = ok && g.phred_scaled_somatic_likelihood >= p.min_likelihood;
  if (!p.seq_overload) {
    ok = ok && g.somatic_log_odds > (double)p.min_lod;
    ok = ok && t.mean_mapping_quality >= (double)p.min_average_mapping_quality && nrm.mean_mapping_quality >= (double)p.min_average_mapping_quality;
    ok = ok && t.mean_mapping_quality >= (double)p.min_average_base_quality && nrm.mean_mapping_quality >= (double)p.min_average_base_quality;
    ok = ok && t.median_mismatches_per_read <= (double)p.max_median_mismatches;
  }
  return ok;
}

// thread per finished record: the ones that pass every filter are compacted into `kept` (counters[6])
__global__ void __launch_bounds__(256) k_somatic_filter(SomOut out, guac_somatic_filter_params p, guac_somatic_record* __restrict__ kept) {
  const uint32_t n = (uint32_t)min(out.counters[0], (unsigned long long)out.cap_rec);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const guac_somatic_record r = out.rec[i];
    if (somatic_filter_keep(r, p)) kept[atomicAdd(&out.counters[6], 1ull)] = r;
  }
}

}  // namespace guac

// ---- host side ------------------------------------------------------------------------------------------------------------------------
namespace {

// Few tiles (a short contig, an amplicon) leave most SMs idle: split every 4096-loci tile into CTAs that own 256 loci each
// (the kernels skip the words of the tile outside [locus_begin, locus_end), so a narrower range is all it takes).
inline void split_tiles_for_occupancy(std::vector<TileDesc>& tiles, int sm_count) {
  if (tiles.empty() || tiles.size() >= (size_t)sm_count * 4) return;
  constexpr int kSub = 256;
  std::vector<TileDesc> fine;
  for (const TileDesc& t : tiles)
    for (int lo = (t.locus_begin / kSub) * kSub; lo < t.locus_end; lo += kSub) {
      TileDesc s = t;
      s.locus_begin = std::max(t.locus_begin, lo);
      s.locus_end = std::min(t.locus_end, lo + kSub);
      if (s.locus_end > s.locus_begin) fine.push_back(s);
    }
  tiles.swap(fine);
}

void somatic_init_tables(guac_ctx* ctx) {
  std::vector<double> t(kTabTotal);
  for (int p = 0; p < 256; ++p) t[kTabSucc + p] = 1.0 - std::pow(10.0, -p / 10.0);  // ADAM PhredUtils.phredToSuccessProbability
  for (int q = 0; q < 256; ++q) {
    const double s = t[kTabSucc + q];  // probabilityCorrectIgnoringAlignment
    t[kTabN + 2 * q] = std::log(s + s);
    t[kTabN + 2 * q + 1] = std::log((1 - s) + (1 - s));
    for (int m = 0; m < 256; ++m) {
      const double sm = t[kTabSucc + q] * t[kTabSucc + m];  // probabilityCorrectIncludingAlignment
      t[kTabT + 2 * (m * 256 + q)] = std::log(sm + sm);
      t[kTabT + 2 * (m * 256 + q) + 1] = std::log((1 - sm) + (1 - sm));
    }
  }
  CUDA_OK(cudaMalloc((void**)&ctx->d_tables, kTabTotal * sizeof(double)));
  CUDA_OK(cudaMemcpy(ctx->d_tables, t.data(), kTabTotal * sizeof(double), cudaMemcpyHostToDevice));
}

void run_somatic(guac_ctx* ctx, const guac_reads& tumor, const guac_reads& normal, const guac_locus_range* ranges, size_t n_ranges,
                 const guac_somatic_params& p, guac_result& res, const guac_somatic_filter_params* filters = nullptr) {
  if (tumor.n_contigs != normal.n_contigs) fail(GUAC_ERR_INVALID_ARGUMENT, "tumor and normal samples have different sequence dictionaries");
  if (!tumor.has_qualities || !normal.has_qualities) fail(GUAC_ERR_UNSUPPORTED, "somatic-standard needs reads packed with base qualities");
  cudaStream_t st = ctx->stream;
  Trace tr("somatic");
  // tiles over the union of both tracks: a locus past one sample's track simply holds no reads of that sample
  check_ranges_disjoint(ranges, n_ranges);
  std::vector<TileDesc> tiles;
  uint64_t requested = 0;
  for (size_t i = 0; i < n_ranges; ++i) {
    const guac_locus_range& r = ranges[i];
    if (r.contig < 0 || (uint32_t)r.contig >= tumor.n_contigs) fail(GUAC_ERR_INVALID_ARGUMENT, "locus range %zu: contig out of range", i);
    if (r.start < 0 || r.end < r.start) fail(GUAC_ERR_INVALID_ARGUMENT, "locus range %zu: bad bounds", i);
    requested += (uint64_t)(r.end - r.start);
    const int64_t len = std::max(tumor.contigs[r.contig].length, normal.contigs[r.contig].length);
    const int64_t s = r.start, e = std::min<int64_t>(r.end, len);
    for (int64_t t = s / kTileLoci; t * kTileLoci < e; ++t) {
      TileDesc td{r.contig, (int32_t)(t * kTileWords), (int32_t)std::max<int64_t>(s, t * kTileLoci), (int32_t)std::min<int64_t>(e, (t + 1) * kTileLoci)};
      if (td.locus_end > td.locus_begin) tiles.push_back(td);
    }
  }
  res.stats.reads_total = tumor.n + normal.n;
  res.stats.loci_requested = requested;
  res.stats.order_sensitive_loci = tumor.order_sensitive_loci + normal.order_sensitive_loci;
  if (tiles.empty()) return;
  split_tiles_for_occupancy(tiles, ctx->sm_count);
  uint64_t tile_loci = 0;
  for (auto& t : tiles) tile_loci += (uint64_t)(t.locus_end - t.locus_begin);
  DevBuf<TileDesc> d_tiles;
  h2d(ctx, d_tiles, tiles.data(), tiles.size());
  res.stats.h2d_bytes = d_tiles.bytes();
  tr.lap("tiles");
  uint64_t cap_rec = std::max<uint64_t>(4096, tile_loci / 256), cap_slow = std::max<uint64_t>(4096, tile_loci / 8);
  uint64_t cap_pool = kPoolDynOff + std::max<uint64_t>(65536, tile_loci / 64);
  SomParams prm{p.odds_threshold, p.min_alignment_quality, p.filter_multi_allelic, p.max_read_depth, p.skip_empty, tumor.sample};
  for (int attempt = 0; attempt < 6; ++attempt) {
    if (cap_rec >= 0xFFFFFFF0ull || cap_slow >= 0xFFFFFFF0ull || cap_pool >= 0xFFFFFFF0ull)
      fail(GUAC_ERR_UNSUPPORTED, "too many output records for one call: split the loci ranges");
    ctx->out_rec.ensure(cap_rec * sizeof(guac_somatic_record));
    ctx->out_slow.ensure(cap_slow * sizeof(SlowLocus));
    if (ctx->out_pool.ensure(cap_pool)) ctx->pool_head_ready = false;
    CUDA_OK(cudaMemsetAsync(ctx->d_counters, 0, 16 * sizeof(unsigned long long), st));
    SomOut out;
    out.rec = (guac_somatic_record*)ctx->out_rec.p;
    out.cap_rec = (uint32_t)cap_rec;
    out.pool = ctx->out_pool.p;
    out.cap_pool = (uint32_t)cap_pool;
    out.slow = (SlowLocus*)ctx->out_slow.p;
    out.cap_slow = (uint32_t)cap_slow;
    out.counters = ctx->d_counters;
    out.err = ctx->d_err;
    const DevReads RT = tumor.view(), RN = normal.view();
    CUDA_OK(cudaEventRecord(ctx->ev[0], st));
    if (tumor.q_hdr.n && normal.q_hdr.n) {
      RowTables rt;  // the table's row r belongs to the r-th largest mapping quality present in the tumor sample
      memset(&rt, 0, sizeof rt);
      const int min_q = std::max(0, p.min_alignment_quality);
      for (int m = 255; m >= 0; --m) {
        if ((tumor.mapq_mask[m >> 5] >> (m & 31)) & 1u) {
          if (rt.n_rows < (int)kMaxRank) rt.row_mapq[rt.n_rows++] = (uint8_t)m;
          if (m >= min_q) rt.n_keep_tumor += 1;
        }
        if (((normal.mapq_mask[m >> 5] >> (m & 31)) & 1u) && m >= min_q) rt.n_keep_normal += 1;
      }
      if (!ctx->som_attr_done) {
        CUDA_OK(cudaFuncSetAttribute(k_somatic<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SomSmem)));
        ctx->som_attr_done = true;
      }
      const int grid = (int)std::min<size_t>(tiles.size(), (size_t)ctx->sm_count * GUAC_SOM_MINB);
      k_somatic<true><<<grid, kSomThreads, sizeof(SomSmem), st>>>(RT, RN, d_tiles.p, (uint32_t)tiles.size(), prm, ctx->d_tables, rt, out);
    } else {
      RowTables rt{};
      k_somatic<false><<<(int)tiles.size(), kSomThreads, 0, st>>>(RT, RN, d_tiles.p, (uint32_t)tiles.size(), prm, ctx->d_tables, rt, out);
    }
    CUDA_OK(cudaEventRecord(ctx->ev[1], st));
    k_somatic_exact<<<ctx->sm_count * 32, kSomExactWarps * 32, 0, st>>>(RT, RN, out.slow, prm, ctx->d_tables, out);
    k_evidence<<<ctx->sm_count * 12, kEvidenceWarps * 32, 0, st>>>(RT, RN, prm, out);
    const guac_somatic_record* d_final = out.rec;
    if (filters) {  // the post-call genotype filters, before anything crosses PCIe
      ctx->sort_rec.ensure(cap_rec * sizeof(guac_somatic_record));
      k_somatic_filter<<<ctx->sm_count * 2, 256, 0, st>>>(out, *filters, (guac_somatic_record*)ctx->sort_rec.p);
      d_final = (const guac_somatic_record*)ctx->sort_rec.p;
      res.stats.kernel_launches += 1;
    }
    CUDA_OK(cudaEventRecord(ctx->ev[2], st));
    CUDA_OK(cudaGetLastError());
    unsigned long long* c = ctx->h_counters;
    CUDA_OK(cudaMemcpyAsync(c, ctx->d_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    check_device_error(ctx, "somatic-standard");
    tr.lap("kernels");
    float ms0 = 0, ms1 = 0;
    CUDA_OK(cudaEventElapsedTime(&ms0, ctx->ev[0], ctx->ev[1]));
    CUDA_OK(cudaEventElapsedTime(&ms1, ctx->ev[1], ctx->ev[2]));
    res.stats.tile_kernel_ms += ms0;
    res.stats.exact_kernel_ms += ms1;
    res.stats.kernel_launches += 3;
    if (c[2] > cap_slow || c[0] > cap_rec || kPoolDynOff + c[1] > cap_pool) {
      cap_slow = std::max<uint64_t>(cap_slow, c[2] + c[2] / 8 + 16);
      cap_rec = std::max<uint64_t>(cap_rec, c[0] + c[0] / 8 + 16);
      cap_pool = std::max<uint64_t>(cap_pool, kPoolDynOff + c[1] + c[1] / 8 + 16);
      continue;
    }
    const uint64_t n_rec = filters ? c[6] : c[0];
    const size_t pool_bytes = (size_t)(kPoolDynOff + c[1]), rec_bytes = (size_t)(n_rec * sizeof(guac_somatic_record));
    const size_t rec_at = (pool_bytes + 63) & ~(size_t)63;
    res.pool = ctx->pinned;
    res.block = ctx->pinned->take(rec_at + rec_bytes + 64, &res.block_bytes);
    if (!res.block) fail(GUAC_ERR_OOM, "pinned host allocation of %zu bytes failed", rec_at + rec_bytes + 64);
    unsigned char* hs = (unsigned char*)res.block;
    unsigned char* hrec = hs + rec_at;
    // canonical order on the device (guac_order.cuh): the records cross PCIe in the order the caller sees them
    const unsigned char* d_sorted = (const unsigned char*)d_final;
    if (ctx->sort_records) d_sorted = device_order_records(ctx, tumor, d_sorted, (uint32_t)sizeof(guac_somatic_record), n_rec, ctx->out_pool.p);
    const bool device_sorted = d_sorted != (const unsigned char*)d_final || n_rec < 2;
    if (d_sorted != (const unsigned char*)d_final) res.stats.kernel_launches += 7;
    CUDA_OK(cudaMemcpyAsync(hs, ctx->out_pool.p, pool_bytes, cudaMemcpyDeviceToHost, st));
    if (n_rec) CUDA_OK(cudaMemcpyAsync(hrec, d_sorted, rec_bytes, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    tr.lap("d2h");
    res.stats.d2h_bytes = pool_bytes + rec_bytes + 64;
    res.records = hrec;
    res.n_records = (size_t)n_rec;
    res.bytes = hs;
    res.n_bytes = pool_bytes;
    if (ctx->sort_records && !device_sorted) {
      const uint8_t* pool = hs;
      guac_somatic_record* first = (guac_somatic_record*)hrec;
      sort_records_canonical(first, (size_t)n_rec, [pool](const guac_somatic_record& a, const guac_somatic_record& b) {
        int c = memcmp(pool + a.ref_off, pool + b.ref_off, std::min(a.ref_len, b.ref_len));
        if (c != 0) return c < 0;
        if (a.ref_len != b.ref_len) return a.ref_len < b.ref_len;
        c = memcmp(pool + a.alt_off, pool + b.alt_off, std::min(a.alt_len, b.alt_len));
        if (c != 0) return c < 0;
        return a.alt_len < b.alt_len;
      });
    }
    tr.lap("sort");
    res.stats.loci_visited = c[3];
    res.stats.records = n_rec;
    res.stats.exact_loci = c[2];
    res.stats.kernel_ms = res.stats.tile_kernel_ms + res.stats.exact_kernel_ms;
    return;
  }
  fail(GUAC_ERR_CUDA, "output buffers did not converge");
}

}  // namespace
