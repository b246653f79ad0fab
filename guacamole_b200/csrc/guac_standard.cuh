// guac_standard.cuh — germline-standard caller: genotype likelihoods of ONE sample per locus, fused with the pileup.
//
//   GermlineStandard.Caller.callVariantsAtLocus                           commands/GermlineStandardCaller.scala:90-124
//   QualityAlignedReadsFilter                                              filters/PileupElementsFilter.scala:48-50
//   Likelihood.likelihoodsOfAllPossibleGenotypesFromPileup(logSpace, normalize), probabilityCorrectIgnoringAlignment
//                                                                          likelihood/Likelihood.scala:48-50, 99-113, 149-201
//   Genotype.getNonReferenceAlleles                                        variants/Genotype.scala:46-48
//   AlleleEvidence.apply over the UNFILTERED sample pileup (:120)          variants/AlleleEvidence.scala:58-101
//
// The machinery is the somatic caller's (guac_somatic.cuh): K_standard gathers one sample exactly like K_somatic gathers the
// normal sample (base quality only), decides A/C/G/T-only loci in registers and rejects, with the same exact bound, the
// loci whose homozygous-reference genotype leads; K_standard_exact walks the remaining loci element by element;
// K_standard_evidence fills the AlleleEvidence of every emitted record.
#pragma once

#include "guac_somatic.cuh"

namespace guac {

struct StdOut {
  guac_called_allele* rec;
  uint32_t cap_rec;
  uint8_t* pool;
  uint32_t cap_pool;
  SlowLocus* slow;
  uint32_t cap_slow;
  unsigned long long* counters;  // [0] records [1] pool bytes [2] slow loci [3] visited loci
  DevError* err;
};

// one CalledAllele; its evidence (all but the likelihood) is filled by K_standard_evidence
__device__ void emit_called(const AlleleView& av, const AlleleEntry& a, double probability, int contig, int locus, int sample, StdOut& out) {
  guac_called_allele r;
  memset(&r, 0, sizeof r);
  r.start = locus;
  r.contig = contig;
  r.sample = sample;
  const int rl = av.ref_len(a), al = av.alt_len(a);
  const uint32_t o = kPoolDynOff + (uint32_t)atomicAdd(&out.counters[1], (unsigned long long)(rl + al));
  if ((unsigned long long)o + rl + al <= out.cap_pool) {
    for (int i = 0; i < rl; ++i) out.pool[o + i] = av.ref_at(a, i);
    for (int i = 0; i < al; ++i) out.pool[o + rl + i] = av.alt_at(a, i);
  }
  r.ref_off = o;
  r.ref_len = (uint16_t)rl;
  r.alt_off = o + rl;
  r.alt_len = (uint16_t)al;
  r.evidence.likelihood = probability;
  const uint32_t s = (uint32_t)atomicAdd(&out.counters[0], 1ull);
  if (s < out.cap_rec) out.rec[s] = r;
}

// Most likely genotype of an A/C/G/T-only locus, log space, normalised (maxBy: the first maximum wins).  false = the
// homozygous-reference genotype (no non-reference allele: nothing to emit), decided where possible without exp / log by
// the lead test of snv_tumor_most_likely().
__device__ __forceinline__ bool snv_most_likely_log(const SnvAlleles& S, int rc, int* c1, int* c2, double* best_log) {
  double lk[10];
  snv_log_likelihoods(S, lk);
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
  snv_normalize<true>(S, lk);
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
  *best_log = best;
  return ci != rc || cj != rc;
}

__global__ void __launch_bounds__(kSomThreads, GUAC_SOM_MINB) k_standard(DevReads R, const TileDesc* __restrict__ tiles, SomParams prm,
                                                                         const double* __restrict__ tables, StdOut out) {
  const TileDesc td = tiles[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const ContigInfo ci = R.contigs[td.contig];
  uint32_t n_visited = 0;
  for (int wi = warp; wi < kTileWords; wi += kSomThreads / 32) {
    const int w = td.word0 + wi;
    const int span_lo = w << 5, x = span_lo + lane;
    if (span_lo >= td.locus_end || span_lo + 32 <= td.locus_begin) continue;
    const bool in_req = x >= td.locus_begin && x < td.locus_end;
    uint32_t tl = 0, th = 0, ts = 0;
    if (w < ci.n_words) { tl = R.trk_lo[ci.word_off + w]; th = R.trk_hi[ci.word_off + w]; ts = R.trk_std[ci.word_off + w]; }
    const int rc = (int)((tl >> lane) & 1u) | ((int)((th >> lane) & 1u) << 1);
    const bool std_ref = (ts >> lane) & 1u;
    LaneAcc A;
    gather_sample<false>(R, td.contig, span_lo, x, rc, std_ref, prm, tables, A);  // base quality only (IgnoringAlignment)
    if (!in_req) continue;  // (no warp-collective operation below this line)
    if (A.any == 0) {
      if (!prm.skip_empty) ++n_visited;
      continue;
    }
    ++n_visited;
    if (A.depth + A.other == 0) continue;                               // every element dropped by the mapq filter
    if (A.ref_depth == A.depth && A.other == 0 && std_ref) continue;    // only the reference allele: hom-ref
    bool exact = A.hard > 0 || !std_ref;
    if (!exact && A.other > 0) {
      if (ref_leads_despite_others(A, rc)) continue;                    // hom-ref is the most likely genotype
      exact = true;
    }
    if (exact) {
      const uint32_t s = (uint32_t)atomicAdd(&out.counters[2], 1ull);
      if (s < out.cap_slow) out.slow[s] = SlowLocus{td.contig, x};
      continue;
    }
    SnvAlleles S;
    snv_compact(A, S);
    int c1 = 0, c2 = 0;
    double best_log = 0.0;
    if (!snv_most_likely_log(S, rc, &c1, &c2, &best_log)) continue;
    const double probability = exp(best_log);
    AlleleView av{R, code_base(rc)};
    if (c1 != rc) emit_called(av, snv_entry(c1, 0), probability, td.contig, x, prm.tumor_sample, out);
    if (c2 != rc) emit_called(av, snv_entry(c2, 0), probability, td.contig, x, prm.tumor_sample, out);
  }
  for (int o = 16; o; o >>= 1) n_visited += __shfl_xor_sync(0xFFFFFFFFu, n_visited, o);
  if (lane == 0 && n_visited) atomicAdd(&out.counters[3], (unsigned long long)n_visited);
}

__global__ void __launch_bounds__(kSomExactWarps * 32) k_standard_exact(DevReads R, const SlowLocus* __restrict__ loci, SomParams prm,
                                                                        const double* __restrict__ tables, StdOut out) {
  __shared__ ExactSmem sm[kSomExactWarps];
  const int lane = threadIdx.x & 31;
  ExactSmem& S = sm[threadIdx.x >> 5];
  const uint32_t n_loci = (uint32_t)min(out.counters[2], (unsigned long long)out.cap_slow);
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t t = warp; t < n_loci; t += n_warps) {
    const int contig = loci[t].contig, locus = loci[t].locus;
    SampleStats st;
    uint8_t ref_base;
    const bool ok = exact_sample<false>(R, contig, locus, prm, tables, S.tab[0], S.ring, st, &ref_base, out.err);
    if (ok && lane == 0 && st.depth > 0) {
      AlleleView av{R, ref_base};
      uint8_t* gi = S.G.gi;
      uint8_t* gj = S.G.gj;
      double* lk = S.G.lk;
      const int ng = genotype_likelihoods(av, S.tab[0], st.n_alleles, st, gi, gj, lk, /*log_space=*/true);
      if (ng < 0) report_error(out.err, GUAC_ERR_UNSUPPORTED, ((unsigned long long)contig << 32) | (uint32_t)locus);
      if (ng > 0) {
        int best = 0;  // maxBy = reduceLeft((x, y) => if (f(x) >= f(y)) x else y)
        for (int g = 1; g < ng; ++g)
          if (!(lk[best] >= lk[g])) best = g;
        const double probability = exp(lk[best]);
        const AlleleEntry a1 = as_entry(S.tab[0][gi[best]]), a2 = as_entry(S.tab[0][gj[best]]);
        if (av.is_variant(a1)) emit_called(av, a1, probability, contig, locus, prm.tumor_sample, out);
        if (av.is_variant(a2)) emit_called(av, a2, probability, contig, locus, prm.tumor_sample, out);
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kEvidenceWarps * 32) k_standard_evidence(DevReads R, SomParams prm_unfiltered, StdOut out) {
  __shared__ EvidenceSmem sm[kEvidenceWarps];
  EvidenceSmem& S = sm[threadIdx.x >> 5];
  const uint32_t n_rec = (uint32_t)min(out.counters[0], (unsigned long long)out.cap_rec);
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t t = warp; t < n_rec; t += n_warps) {
    guac_called_allele& r = out.rec[t];
    if ((unsigned long long)r.alt_off + r.alt_len > out.cap_pool) continue;  // pool overflow: the host reruns
    guac_allele_evidence ev = r.evidence;
    evidence_sample(R, r.contig, (int)r.start, prm_unfiltered, out.pool + r.ref_off, r.ref_len, out.pool + r.alt_off, r.alt_len, ev, S, out.err);
    if ((threadIdx.x & 31) == 0) {
      r.evidence = ev;
      r.phred_scaled_likelihood = success_probability_to_phred(ev.likelihood - 1e-10);  // AlleleEvidence.scala:52
    }
    __syncwarp();
  }
}

}  // namespace guac

// ---- host side ------------------------------------------------------------------------------------------------------------------------
namespace {

void run_standard(guac_ctx* ctx, const guac_reads& reads, const guac_locus_range* ranges, size_t n_ranges, const guac_standard_params& p,
                  guac_result& res) {
  if (!reads.has_qualities) fail(GUAC_ERR_UNSUPPORTED, "germline-standard needs reads packed with base qualities");
  cudaStream_t st = ctx->stream;
  check_ranges_disjoint(ranges, n_ranges);
  std::vector<TileDesc> tiles;
  uint64_t requested = 0;
  for (size_t i = 0; i < n_ranges; ++i) {
    const guac_locus_range& r = ranges[i];
    if (r.contig < 0 || (uint32_t)r.contig >= reads.n_contigs) fail(GUAC_ERR_INVALID_ARGUMENT, "locus range %zu: contig out of range", i);
    if (r.start < 0 || r.end < r.start) fail(GUAC_ERR_INVALID_ARGUMENT, "locus range %zu: bad bounds", i);
    requested += (uint64_t)(r.end - r.start);
    const int64_t s = r.start, e = std::min<int64_t>(r.end, reads.contigs[r.contig].length);
    for (int64_t t = s / kTileLoci; t * kTileLoci < e; ++t) {
      TileDesc td{r.contig, (int32_t)(t * kTileWords), (int32_t)std::max<int64_t>(s, t * kTileLoci), (int32_t)std::min<int64_t>(e, (t + 1) * kTileLoci)};
      if (td.locus_end > td.locus_begin) tiles.push_back(td);
    }
  }
  res.stats.reads_total = reads.n;
  res.stats.loci_requested = requested;
  res.stats.order_sensitive_loci = reads.order_sensitive_loci;
  if (tiles.empty()) {
    res.stats.loci_visited = p.skip_empty ? 0 : requested;
    return;
  }
  split_tiles_for_occupancy(tiles, ctx->sm_count);
  uint64_t tile_loci = 0;
  for (auto& t : tiles) tile_loci += (uint64_t)(t.locus_end - t.locus_begin);
  DevBuf<TileDesc> d_tiles;
  h2d(ctx, d_tiles, tiles.data(), tiles.size());
  res.stats.h2d_bytes = d_tiles.bytes();
  uint64_t cap_rec = std::max<uint64_t>(4096, tile_loci / 128), cap_slow = std::max<uint64_t>(4096, tile_loci / 8);
  uint64_t cap_pool = kPoolDynOff + std::max<uint64_t>(65536, tile_loci / 64);
  SomParams prm{0, p.min_alignment_quality, 0, 0x7FFFFFFF, p.skip_empty, reads.sample};
  SomParams prm_unfiltered = prm;
  prm_unfiltered.min_mapq = 0;
  for (int attempt = 0; attempt < 6; ++attempt) {
    if (cap_rec >= 0xFFFFFFF0ull || cap_slow >= 0xFFFFFFF0ull || cap_pool >= 0xFFFFFFF0ull)
      fail(GUAC_ERR_UNSUPPORTED, "too many output records for one call: split the loci ranges");
    ctx->out_rec.ensure(cap_rec * sizeof(guac_called_allele));
    ctx->out_slow.ensure(cap_slow * sizeof(SlowLocus));
    if (ctx->out_pool.ensure(cap_pool)) ctx->pool_head_ready = false;
    CUDA_OK(cudaMemsetAsync(ctx->d_counters, 0, 16 * sizeof(unsigned long long), st));
    StdOut out;
    out.rec = (guac_called_allele*)ctx->out_rec.p;
    out.cap_rec = (uint32_t)cap_rec;
    out.pool = ctx->out_pool.p;
    out.cap_pool = (uint32_t)cap_pool;
    out.slow = (SlowLocus*)ctx->out_slow.p;
    out.cap_slow = (uint32_t)cap_slow;
    out.counters = ctx->d_counters;
    out.err = ctx->d_err;
    const DevReads R = reads.view();
    CUDA_OK(cudaEventRecord(ctx->ev[0], st));
    k_standard<<<(int)tiles.size(), kSomThreads, 0, st>>>(R, d_tiles.p, prm, ctx->d_tables, out);
    CUDA_OK(cudaEventRecord(ctx->ev[1], st));
    k_standard_exact<<<ctx->sm_count * 32, kSomExactWarps * 32, 0, st>>>(R, out.slow, prm, ctx->d_tables, out);
    k_standard_evidence<<<ctx->sm_count * 12, kEvidenceWarps * 32, 0, st>>>(R, prm_unfiltered, out);
    CUDA_OK(cudaEventRecord(ctx->ev[2], st));
    CUDA_OK(cudaGetLastError());
    unsigned long long* c = ctx->h_counters;
    CUDA_OK(cudaMemcpyAsync(c, ctx->d_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    check_device_error(ctx, "germline-standard");
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
    const uint64_t n_rec = c[0];
    const size_t pool_bytes = (size_t)(kPoolDynOff + c[1]), rec_bytes = (size_t)(n_rec * sizeof(guac_called_allele));
    const size_t rec_at = (pool_bytes + 63) & ~(size_t)63;
    res.pool = ctx->pinned;
    res.block = ctx->pinned->take(rec_at + rec_bytes + 64, &res.block_bytes);
    if (!res.block) fail(GUAC_ERR_OOM, "pinned host allocation of %zu bytes failed", rec_at + rec_bytes + 64);
    unsigned char* hs = (unsigned char*)res.block;
    unsigned char* hrec = hs + rec_at;
    const unsigned char* d_sorted = ctx->out_rec.p;  // canonical order on the device (guac_order.cuh)
    if (ctx->sort_records) d_sorted = device_order_records(ctx, reads, d_sorted, (uint32_t)sizeof(guac_called_allele), n_rec, ctx->out_pool.p);
    const bool device_sorted = d_sorted != ctx->out_rec.p || n_rec < 2;
    if (d_sorted != ctx->out_rec.p) res.stats.kernel_launches += 7;
    CUDA_OK(cudaMemcpyAsync(hs, ctx->out_pool.p, pool_bytes, cudaMemcpyDeviceToHost, st));
    if (n_rec) CUDA_OK(cudaMemcpyAsync(hrec, d_sorted, rec_bytes, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    res.stats.d2h_bytes = pool_bytes + rec_bytes + 64;
    res.records = hrec;
    res.n_records = (size_t)n_rec;
    res.bytes = hs;
    res.n_bytes = pool_bytes;
    if (ctx->sort_records && !device_sorted) {
      const uint8_t* pool = hs;
      guac_called_allele* first = (guac_called_allele*)hrec;
      sort_records_canonical(first, (size_t)n_rec, [pool](const guac_called_allele& a, const guac_called_allele& b) {
        int c = memcmp(pool + a.ref_off, pool + b.ref_off, std::min(a.ref_len, b.ref_len));
        if (c != 0) return c < 0;
        if (a.ref_len != b.ref_len) return a.ref_len < b.ref_len;
        c = memcmp(pool + a.alt_off, pool + b.alt_off, std::min(a.alt_len, b.alt_len));
        if (c != 0) return c < 0;
        return a.alt_len < b.alt_len;
      });
    }
    res.stats.loci_visited = c[3] + (p.skip_empty ? 0 : requested - tile_loci);
    res.stats.records = n_rec;
    res.stats.exact_loci = c[2];
    res.stats.kernel_ms = res.stats.tile_kernel_ms + res.stats.exact_kernel_ms;
    return;
  }
  fail(GUAC_ERR_CUDA, "output buffers did not converge");
}

}  // namespace
