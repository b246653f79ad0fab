// guac_api.cu — host side of libguac_b200.so: the C ABI of include/guac.h over the CUDA kernels (sm_100a).
// There is NO CPU fallback: without a CUDA device every entry point fails with GUAC_ERR_NO_DEVICE.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <exception>
#include <memory>
#include <new>
#include <string>
#include <tuple>
#include <vector>

#include <functional>
#include <thread>

#include "guac_host.cuh"
#include "guac_pack.cuh"
#include "guac_pileup.cuh"
#include "guac_tile.cuh"
#include "guac_somatic.cuh"
#include "guac_standard.cuh"
#include "guac_synth_device.cuh"
#include "guac_comm.cuh"
#include "guac_batch2.cuh"
#include "guac_bam.cuh"
#include "../../include/guac_synth.h"

namespace {


// ---- per-granule difference streams (k_expand, guac_tile.cuh) -------------------------------------------------------------------
// One launch builds every granule's stream.  The stream buffer is sized from an estimate and the start / end counters are
// stored as nibbles first: a launch that ran out of either (counters[2] / [3]) is repeated with what it asked for.
void alloc_streams(guac_reads& rd, uint64_t cap_entries) {
  const uint64_t grans = rd.total_grans;
  cap_entries = (cap_entries + 7) & ~7ull;
  rd.gs_hdr.alloc(grans + 1);
  rd.gs_diffs.alloc(cap_entries + 8);
  rd.gs_dd.alloc(grans * kGranuleLoci * (rd.gs_wide ? 4 : 1) + 16);
  rd.gs_dp.alloc(grans * kGranuleLoci * (rd.gs_wide ? 4 : 1) + 16);
  rd.gs_imp.alloc(grans * (kGranuleLoci / 32) + 32);
  rd.gs_entries = cap_entries;
}

// [g_begin, g_end): the granules of this launch (guac_reads_pack expands the granules behind a copy chunk's last read while
// the next chunk is still on the bus); `first` / `last`: clears the counters / closes the timing of a series of launches
void launch_expand(guac_ctx* ctx, guac_reads& rd, bool huge, uint64_t cap_entries, bool allocated = false, uint64_t g_begin = 0,
                   uint64_t g_end = ~0ull, bool first = true, bool last = true) {
  cudaStream_t st = ctx->stream;
  const uint64_t grans = rd.total_grans;
  if (!allocated) alloc_streams(rd, cap_entries);
  cap_entries = rd.gs_entries;
  if (first) CUDA_OK(cudaMemsetAsync(ctx->d_counters + 2, 0, 2 * sizeof(unsigned long long), st));
  if (!grans) return;
  g_end = std::min<uint64_t>(g_end, grans);
  ExpandArgs E;
  E.R = rd.view();
  E.hdr_w = rd.gs_hdr.p;
  E.diffs_w = rd.gs_diffs.p;
  E.dd_w = rd.gs_dd.p;
  E.dp_w = rd.gs_dp.p;
  E.imp_w = rd.gs_imp.p;
  E.cap_diffs = cap_entries;
  E.g_begin = (uint32_t)g_begin;
  E.g_end = (uint32_t)g_end;
  E.n_contigs = rd.n_contigs;
  E.wide = rd.gs_wide ? 1 : 0;
  E.counters = ctx->d_counters;
  E.err = ctx->d_err;
  const int ctas = (int)((g_end - std::min(g_begin, g_end) + kExpandWarps - 1) / kExpandWarps);
  if (!ctx->expand_attrs_done) {
    CUDA_OK(cudaFuncSetAttribute(k_expand<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kExpandWarps * sizeof(ExpandSmem<false>))));
    CUDA_OK(cudaFuncSetAttribute(k_expand<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kExpandWarps * sizeof(ExpandSmem<true>))));
    ctx->expand_attrs_done = true;
  }
  if (first) CUDA_OK(cudaEventRecord(ctx->ev[2], st));
  if (ctas > 0) {
    if (huge) k_expand<true><<<ctas, kExpandWarps * 32, kExpandWarps * sizeof(ExpandSmem<true>), st>>>(E);
    else k_expand<false><<<ctas, kExpandWarps * 32, kExpandWarps * sizeof(ExpandSmem<false>), st>>>(E);
  }
  if (last) CUDA_OK(cudaEventRecord(ctx->ev[3], st));
  CUDA_OK(cudaGetLastError());
}

// After a launch: `entries` reserved, `field_overflow` = a start / end count did not fit.  Repeats the launch until it fits.
void settle_streams(guac_ctx* ctx, guac_reads& rd, uint64_t entries, bool field_overflow, bool was_huge) {
  cudaStream_t st = ctx->stream;
  for (int attempt = 0; attempt < 6; ++attempt) {
    const bool need_huge = rd.max_reads_per_granule > 65535 && !was_huge;
    const bool need_wide = !rd.gs_wide && (field_overflow || rd.max_reads_per_granule >= 2048);
    if (field_overflow && rd.gs_wide && !need_huge) fail(GUAC_ERR_UNSUPPORTED, "more than 65535 reads start or end at one locus");
    if (!need_huge && !need_wide && entries <= rd.gs_entries) {
      float ms = 0;
      CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
      rd.expand_ms = ms;
      return;
    }
    if (need_wide) rd.gs_wide = true;
    was_huge = was_huge || need_huge;
    launch_expand(ctx, rd, was_huge, std::max<uint64_t>(entries + entries / 16 + 64, rd.gs_entries));
    rd.pack_launches += 1;
    unsigned long long c[2];
    CUDA_OK(cudaMemcpyAsync(c, ctx->d_counters + 2, sizeof c, cudaMemcpyDeviceToHost, st));
    check_device_error(ctx, "guac_reads_pack (difference streams)");
    entries = c[0];
    field_overflow = c[1] != 0;
  }
  fail(GUAC_ERR_CUDA, "difference streams did not converge");
}

// ---- per-word rows of the likelihood callers (k_expand_rows, guac_rows.cuh) -----------------------------------------------------
void alloc_rows(guac_reads& rd, uint64_t cap_pairs, uint64_t cap_groups) {
  if (cap_pairs >= 0xFFFFFFF0ull || cap_groups >= 0xFFFFFFF0ull) fail(GUAC_ERR_UNSUPPORTED, "row store beyond 2^32 entries: shard the read set");
  rd.q_hdr.alloc(rd.total_words + 1);
  rd.q_depth.alloc(rd.total_words * 32 + 32);
  rd.q_cols.alloc((cap_pairs + 1) * 32 * 4);  // (blocks of eight columns: 16 bytes per locus)
  rd.q_groups.alloc(cap_groups + 1);
  rd.q_rows.alloc((cap_groups + 1) * 32);
  rd.q_cap_pairs = cap_pairs;
  rd.q_cap_groups = cap_groups;
}

void launch_rows(guac_ctx* ctx, guac_reads& rd, uint64_t w_begin = 0, uint64_t w_end = ~0ull, bool first = true, bool last = true) {
  cudaStream_t st = ctx->stream;
  if (first) CUDA_OK(cudaMemsetAsync(ctx->d_counters + 4, 0, 2 * sizeof(unsigned long long), st));
  if (!rd.total_words) return;
  w_end = std::min<uint64_t>(w_end, rd.total_words);
  w_begin = std::min(w_begin, w_end);
  RowsArgs A;
  A.R = rd.view();
  A.hdr_w = rd.q_hdr.p;
  A.depth_w = rd.q_depth.p;
  A.cols_w = rd.q_cols.p;
  A.groups_w = rd.q_groups.p;
  A.rows_w = rd.q_rows.p;
  A.cap_pairs = rd.q_cap_pairs;
  A.cap_groups = rd.q_cap_groups;
  A.w_begin = (uint32_t)w_begin;
  A.w_end = (uint32_t)w_end;
  A.n_contigs = rd.n_contigs;
  A.pad_ = 0;
  A.counters = ctx->d_counters;
  A.err = ctx->d_err;
  memcpy(A.mapq_mask, rd.mapq_mask, sizeof A.mapq_mask);
  if (first) CUDA_OK(cudaEventRecord(ctx->ev_rows[0], st));
  if (w_end > w_begin) k_expand_rows<<<(unsigned)((w_end - w_begin + kRowsWarps - 1) / kRowsWarps), kRowsWarps * 32, 0, st>>>(A);
  if (last) CUDA_OK(cudaEventRecord(ctx->ev_rows[1], st));
  CUDA_OK(cudaGetLastError());
}

void settle_rows(guac_ctx* ctx, guac_reads& rd, uint64_t pairs, uint64_t groups) {
  cudaStream_t st = ctx->stream;
  for (int attempt = 0; attempt < 4; ++attempt) {
    if (pairs <= rd.q_cap_pairs && groups <= rd.q_cap_groups) {
      float ms = 0;
      CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev_rows[0], ctx->ev_rows[1]));
      rd.rows_ms = ms;
      return;
    }
    alloc_rows(rd, std::max<uint64_t>(rd.q_cap_pairs, pairs + pairs / 16 + 64), std::max<uint64_t>(rd.q_cap_groups, groups + groups / 16 + 64));
    launch_rows(ctx, rd);
    rd.pack_launches += 1;
    unsigned long long c[2] = {0, 0};
    CUDA_OK(cudaMemcpyAsync(c, ctx->d_counters + 4, sizeof c, cudaMemcpyDeviceToHost, st));
    check_device_error(ctx, "guac_reads_pack (rows)");
    pairs = c[0];
    groups = c[1];
  }
  fail(GUAC_ERR_CUDA, "row store did not converge");
}

// `on_device`: the batch's column pointers are device memory of ctx's device (guac_reads_pack_device): no host -> device copies.
// `adopt`: a generated device batch whose large columns (bases, qualities, CIGARs, MD tags, base offsets) the store takes
// over instead of copying them (guac_reads_pack_synth): the whole-genome shards would not fit twice.
// `b2`: the compact host batch (guac_reads_pack_v2): its columns are copied as they are and widened on the device, on the copy
// stream (guac_batch2.cuh); `b` is then ignored.
void pack_reads(guac_ctx* ctx, const guac_read_batch* b, const guac_reference* ref, guac_reads& out, bool on_device,
                guac_synth_device_batch* adopt = nullptr, const guac_read_batch_v2* b2 = nullptr) {
  guac_read_batch shell{};
  if (b2) {  // (the column pointers of `shell` stay null: every use below is behind !b2)
    shell.n_reads = b2->n_reads;
    shell.n_contigs = b2->n_contigs;
    shell.contig_length = b2->contig_length;
    b = &shell;
    on_device = false;
  }
  if (!b) fail(GUAC_ERR_INVALID_ARGUMENT, "null batch");
  const uint64_t n = b->n_reads;
  if (n >= 0xFFFFFFF0ull) fail(GUAC_ERR_UNSUPPORTED, "more than 2^32 reads in one read set: shard it");
  const bool need_qual = ctx->pack_qualities != 0;
  if (b2) {
    if (n && (!b2->contig_read_off || !b2->start || !b2->cigar_off || (!b2->seq_off && !b2->read_length) || !b2->seq4 || (need_qual && !b2->qual) ||
              !b2->mapq || !b2->flags || !b2->md_off))
      fail(GUAC_ERR_INVALID_ARGUMENT, "null column in read batch");
    if (n) {
      if (b2->contig_read_off[0] != 0 || b2->contig_read_off[b2->n_contigs] != n) fail(GUAC_ERR_INVALID_ARGUMENT, "contig_read_off does not span the batch");
      for (uint32_t c = 0; c < b2->n_contigs; ++c)
        if (b2->contig_read_off[c] > b2->contig_read_off[c + 1]) fail(GUAC_ERR_CONTIG_ORDER, "contig_read_off descends at contig %u", c);
      if (b2->read_length && (uint64_t)b2->read_length * n >= 0xFFFFFFFF00ull) fail(GUAC_ERR_UNSUPPORTED, "more than 2^37 bases in one read set: shard it");
    }
  } else if (n && (!b->contig || !b->start || !b->cigar_off || !b->seq_off || !b->seq || (need_qual && !b->qual) || !b->mapq || !b->flags || !b->md_off))
    fail(GUAC_ERR_INVALID_ARGUMENT, "null column in read batch");
  out.ctx = ctx;
  out.device = ctx->device;
  out.n = n;
  out.n_contigs = b->n_contigs;
  if (b->n_contigs == 0 && n) fail(GUAC_ERR_INVALID_ARGUMENT, "reads without contigs");
  if (ref && ref->n_contigs != b->n_contigs) fail(GUAC_ERR_INVALID_ARGUMENT, "reference and batch disagree on the number of contigs");

  Trace tr("pack");
  nvtx_push("guac pack: copies + header kernel");
  cudaStream_t cs = ctx->copy_stream, st = ctx->stream;
  CUDA_OK(cudaEventRecord(ctx->copy_ev[9], st));  // buffers handed back by earlier calls may still be in use there
  CUDA_OK(cudaStreamWaitEvent(cs, ctx->copy_ev[9], 0));
  struct DrainOnError {  // a failed pack must not leave copies from the caller's buffers in flight
    guac_ctx* c;
    int pending = std::uncaught_exceptions();
    ~DrainOnError() {
      if (std::uncaught_exceptions() > pending) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamSynchronize(c->stream);
        nvtx_pop();
      }
    }
  } drain{ctx};
  out.has_qualities = need_qual;
  // totals of the variable-length columns (the last offsets)
  uint64_t n_ops = 0, n_md = 0, n_bases = 0;
  if (n) {
    if (on_device) {
      CUDA_OK(cudaMemcpyAsync(&n_ops, b->cigar_off + n, 8, cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaMemcpyAsync(&n_md, b->md_off + n, 8, cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaMemcpyAsync(&n_bases, b->seq_off + n, 8, cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaStreamSynchronize(st));
    } else if (b2) {
      n_ops = b2->cigar_off[n]; n_md = b2->md_off[n];
      n_bases = b2->read_length ? (uint64_t)b2->read_length * n : (uint64_t)b2->seq_off[n];
    } else {
      n_ops = b->cigar_off[n]; n_md = b->md_off[n]; n_bases = b->seq_off[n];
    }
  }
  if (n_ops >= 0xFFFFFFFFull || n_md >= 0xFFFFFFFFull) fail(GUAC_ERR_UNSUPPORTED, "cigar / MD columns too large: shard the read set");
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  auto bring = [&](auto& dst, const auto* src, size_t count, size_t extra) {  // a column the store keeps: copied either way
    dst.alloc(count + extra);
    if (extra) CUDA_OK(cudaMemsetAsync(dst.p + count, 0, extra * sizeof(*dst.p), cs));
    if (count) CUDA_OK(cudaMemcpyAsync(dst.p, src, count * sizeof(*dst.p), kind, cs));
  };
  // ---- the small columns first (copy stream), then the bases / qualities in chunks of reads: their copies run underneath
  // the header kernel and the pack kernels of the earlier chunks
  DevBuf<int32_t> t_contig, t_sample;
  DevBuf<int64_t> t_start, d_contig_length;
  DevBuf<uint64_t> t_cigar_off, t_md_off;
  DevBuf<uint8_t> t_mapq, t_flags;
  HeaderArgs H{};
  H.n = n;
  H.n_contigs = b->n_contigs;
  DevBuf<int32_t> t2_start;
  DevBuf<uint32_t> t2_cigar_off, t2_seq_off, t2_md_off;
  DevBuf<unsigned long long> t2_contig_off;
  DevBuf<uint8_t> t2_seq4;
  if (b2) {
    bring(out.cigar, b2->cigar, (size_t)n_ops, 1);
    bring(out.md, b2->md, (size_t)n_md, 16);
    out.seq_off.alloc(n + 1);
    if (!n) CUDA_OK(cudaMemsetAsync(out.seq_off.p, 0, sizeof(uint64_t), cs));
  } else if (adopt && n) {
    out.cigar.adopt(adopt->cigar);
    out.seq_off.adopt(adopt->seq_off);
    out.md.adopt(adopt->md);
  } else {
    bring(out.cigar, b->cigar, (size_t)n_ops, 1);
    bring(out.seq_off, b->seq_off, n ? n + 1 : 0, n ? 0 : 1);
    bring(out.md, b->md, (size_t)n_md, 16);
  }
  if (on_device) {
    H.contig = b->contig; H.start = b->start; H.cigar_off = b->cigar_off; H.mapq = b->mapq; H.flags = b->flags;
    H.sample = b->sample; H.md_off = b->md_off;
  } else if (b2 && n) {
    bring(t2_start, b2->start, n, 0);
    bring(t2_cigar_off, b2->cigar_off, n + 1, 0);
    bring(t2_md_off, b2->md_off, n + 1, 0);
    if (!b2->read_length) bring(t2_seq_off, b2->seq_off, n + 1, 0);
    bring(t2_contig_off, reinterpret_cast<const unsigned long long*>(b2->contig_read_off), (size_t)b2->n_contigs + 1, 0);
    bring(t_mapq, b2->mapq, n, 0);
    bring(t_flags, b2->flags, n, 0);
    t_contig.alloc(n);
    t_start.alloc(n);
    t_cigar_off.alloc(n + 1);
    t_md_off.alloc(n + 1);
    WidenArgs W{};
    W.n = n;
    W.n_contigs = b2->n_contigs;
    W.read_length = b2->read_length;
    W.contig_read_off = t2_contig_off.p;
    W.start = t2_start.p;
    W.cigar_off = t2_cigar_off.p;
    W.seq_off = b2->read_length ? nullptr : t2_seq_off.p;
    W.md_off = t2_md_off.p;
    W.contig_w = t_contig.p;
    W.start_w = reinterpret_cast<long long*>(t_start.p);
    W.cigar_off_w = reinterpret_cast<unsigned long long*>(t_cigar_off.p);
    W.seq_off_w = reinterpret_cast<unsigned long long*>(out.seq_off.p);
    W.md_off_w = reinterpret_cast<unsigned long long*>(t_md_off.p);
    k_widen_header<<<grid_for(n + 1, 256, ctx->sm_count), 256, 0, cs>>>(W);
    CUDA_OK(cudaGetLastError());
    H.contig = t_contig.p; H.start = t_start.p; H.cigar_off = t_cigar_off.p; H.mapq = t_mapq.p; H.flags = t_flags.p;
    H.sample = nullptr; H.md_off = t_md_off.p;
  } else if (n) {
    bring(t_contig, b->contig, n, 0);
    bring(t_start, b->start, n, 0);
    bring(t_cigar_off, b->cigar_off, n + 1, 0);
    bring(t_md_off, b->md_off, n + 1, 0);
    bring(t_mapq, b->mapq, n, 0);
    bring(t_flags, b->flags, n, 0);
    if (b->sample) bring(t_sample, b->sample, n, 0);
    H.contig = t_contig.p; H.start = t_start.p; H.cigar_off = t_cigar_off.p; H.mapq = t_mapq.p; H.flags = t_flags.p;
    H.sample = b->sample ? t_sample.p : nullptr; H.md_off = t_md_off.p;
  }
  H.cigar = out.cigar.p;
  H.seq_off = out.seq_off.p;
  if (b->contig_length) {
    d_contig_length.alloc(b->n_contigs + 1);
    CUDA_OK(cudaMemcpyAsync(d_contig_length.p, b->contig_length, b->n_contigs * sizeof(int64_t), cudaMemcpyHostToDevice, cs));
    H.contig_length = d_contig_length.p;
  }
  CUDA_OK(cudaEventRecord(ctx->copy_ev[8], cs));  // the small columns are on the device
  const bool adopt_bases = adopt && n;
  if (adopt_bases) {
    if (need_qual && !adopt->qual.n) fail(GUAC_ERR_INVALID_ARGUMENT, "the generated batch holds no base qualities");
    out.seq.adopt(adopt->seq);
    if (need_qual) out.qual.adopt(adopt->qual);
  } else {
    out.seq.alloc((size_t)n_bases + 64);
    CUDA_OK(cudaMemsetAsync(out.seq.p + n_bases, 0, 64, cs));
    if (need_qual) {
      out.qual.alloc((size_t)n_bases + 64);
      CUDA_OK(cudaMemsetAsync(out.qual.p + n_bases, 0, 64, cs));
    }
    if (b2) t2_seq4.alloc((size_t)(n_bases + 1) / 2 + 64);
  }
  constexpr int kMaxCopyChunks = 8;
  const int n_copy_chunks = n >= 1000000 ? kMaxCopyChunks : 1;
  uint64_t chunk_read[kMaxCopyChunks + 1], chunk_byte[kMaxCopyChunks + 1];
  for (int k = 0; k <= n_copy_chunks; ++k) chunk_read[k] = n * (uint64_t)k / (uint64_t)n_copy_chunks;
  if (n) {
    if (on_device) {  // (fixed offsets cannot be assumed: fetch the few chunk boundaries)
      for (int k = 0; k <= n_copy_chunks; ++k) CUDA_OK(cudaMemcpyAsync(&chunk_byte[k], b->seq_off + chunk_read[k], 8, cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaStreamSynchronize(st));
    } else if (b2) {
      for (int k = 0; k <= n_copy_chunks; ++k)
        chunk_byte[k] = b2->read_length ? chunk_read[k] * (uint64_t)b2->read_length : (uint64_t)b2->seq_off[chunk_read[k]];
    } else {
      for (int k = 0; k <= n_copy_chunks; ++k) chunk_byte[k] = b->seq_off[chunk_read[k]];
    }
    for (int k = 0; k < n_copy_chunks; ++k) {
      const uint64_t o0 = chunk_byte[k], o1 = chunk_byte[k + 1];
      if (o1 > o0 && b2) {  // the chunk's nibbles (whole bytes: a byte shared with a neighbour travels twice), widened where they land
        const uint64_t q0 = o0 >> 1, q1 = (o1 + 1) >> 1;
        CUDA_OK(cudaMemcpyAsync(t2_seq4.p + q0, b2->seq4 + q0, q1 - q0, kind, cs));
        k_unpack_bases<<<(unsigned)std::min<uint64_t>(((o1 - o0) / 32 + 256) / 256, (uint64_t)ctx->sm_count * 8), 256, 0, cs>>>(t2_seq4.p, out.seq.p, o0, o1);
        if (need_qual) CUDA_OK(cudaMemcpyAsync(out.qual.p + o0, b2->qual + o0, o1 - o0, kind, cs));
      } else if (o1 > o0 && !adopt_bases) {
        CUDA_OK(cudaMemcpyAsync(out.seq.p + o0, b->seq + o0, o1 - o0, kind, cs));
        if (need_qual) CUDA_OK(cudaMemcpyAsync(out.qual.p + o0, b->qual + o0, o1 - o0, kind, cs));
      }
      CUDA_OK(cudaEventRecord(ctx->copy_ev[k], cs));
    }
  }
  // ---- header kernel: per-read checks + derived columns, the plane offsets by a scan
  DevBuf<uint32_t> d_read_contig, d_n_pairs, d_pair_off, conflict, gran_count;
  DevBuf<unsigned long long> d_summary;  // [0..2] summary, then contig_first / contig_last / contig_end
  const size_t nc = b->n_contigs;
  d_summary.alloc(8 + 3 * nc);
  {
    std::vector<unsigned long long> init(8 + 3 * nc, 0ull);
    init[0] = ~0ull;
    for (size_t c = 0; c < nc; ++c) init[8 + c] = ~0ull;
    CUDA_OK(cudaMemcpyAsync(d_summary.p, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaStreamSynchronize(st));  // (init goes out of scope)
  }
  out.rec.alloc(n + 1);
  out.cig_off.alloc(n + 1);
  out.md_off.alloc(n + 1);
  d_read_contig.alloc(n + 1);
  d_n_pairs.alloc(n + 1);
  d_pair_off.alloc(n + 2);
  H.rec = out.rec.p;
  H.cig_off32 = out.cig_off.p;
  H.md_off32 = out.md_off.p;
  H.read_contig = d_read_contig.p;
  H.n_pairs = d_n_pairs.p;
  H.summary = d_summary.p;
  H.mapq_mask = reinterpret_cast<uint32_t*>(d_summary.p + 4);
  H.contig_first = d_summary.p + 8;
  H.contig_last = d_summary.p + 8 + nc;
  H.contig_end = reinterpret_cast<long long*>(d_summary.p + 8 + 2 * nc);
  CUDA_OK(cudaStreamWaitEvent(st, ctx->copy_ev[8], 0));
  CUDA_OK(cudaEventRecord(ctx->ev[0], st));
  uint64_t pair_total = 0;
  float header_ms = 0;
  if (n) {
    k_header<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>(H);
    pair_total = device_exclusive_scan<uint32_t>(ctx, d_n_pairs.p, n, d_pair_off.p);
    CUDA_OK(cudaEventRecord(ctx->ev[1], st));
    CUDA_OK(cudaEventSynchronize(ctx->ev[1]));
    CUDA_OK(cudaEventElapsedTime(&header_ms, ctx->ev[0], ctx->ev[1]));
    if (pair_total >= 0xFFFFFF00ull) fail(GUAC_ERR_UNSUPPORTED, "more than 2^37 bases in one read set: shard it");
    k_header_finish<<<grid_for(n + 1, 256, ctx->sm_count), 256, 0, st>>>(out.rec.p, d_pair_off.p, out.cig_off.p, out.md_off.p, H.cigar_off, H.md_off, n);
  } else {
    const ReadRec sentinel{0x7FFFFFFF, 0x7FFFFFFF, 0u, 0u};
    CUDA_OK(cudaMemcpyAsync(out.rec.p, &sentinel, sizeof sentinel, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemsetAsync(out.cig_off.p, 0, 4, st));
    CUDA_OK(cudaMemsetAsync(out.md_off.p, 0, 4, st));
  }
  std::vector<unsigned long long> summary(8 + 3 * nc);
  CUDA_OK(cudaMemcpyAsync(summary.data(), d_summary.p, summary.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  if (n && summary[0] != ~0ull) {
    const unsigned long long read = summary[0] >> 8;
    const guac_status code = (guac_status)(summary[0] & 0xFF);
    const char* why = code == GUAC_ERR_UNSORTED_READS ? "Regions must be sorted by start locus"
                      : code == GUAC_ERR_CONTIG_ORDER ? "Regions are not sorted by contig"
                      : code == GUAC_ERR_MISSING_MD   ? "the read has no MD tag (the callers load reads with hasMdTag = true)"
                      : code == GUAC_ERR_INVALID_CIGAR ? "unsupported or inconsistent CIGAR (operator, zero length, or bases consumed != read length)"
                      : code == GUAC_ERR_UNSUPPORTED  ? "unsupported read (several samples in one read set, read too long, or coordinates beyond 2^31)"
                                                      : "contig index or coordinates out of range";
    fail(code, "read %llu: %s", read, why);
  }
  out.max_ref_span = (int64_t)summary[1];
  memcpy(out.mapq_mask, &summary[4], sizeof out.mapq_mask);
  out.sample = b2 ? b2->sample : n ? (int32_t)(uint32_t)summary[2] : 0;
  std::vector<int64_t> contig_end(nc, 0);
  std::vector<uint64_t> contig_first(nc, ~0ull), contig_last(nc, 0);
  for (size_t c = 0; c < nc; ++c) {
    contig_first[c] = summary[8 + c];
    contig_last[c] = summary[8 + nc + c];
    contig_end[c] = (int64_t)summary[8 + 2 * nc + c];
  }
  nvtx_pop();
  nvtx_push("guac pack: kernels");

  tr.lap("header pass");
  // ---- contig geometry
  out.contigs.resize(b->n_contigs);
  uint64_t word_off = 0, gran_off = 0;
  for (uint32_t c = 0; c < b->n_contigs; ++c) {
    ContigInfo& ci = out.contigs[c];
    int64_t len = b->contig_length ? b->contig_length[c] : contig_end[c];
    if (ref) len = std::max<int64_t>(len, (int64_t)(ref->base_off[c + 1] - ref->base_off[c]));
    if (len > 0x7FFFFF00ll) fail(GUAC_ERR_UNSUPPORTED, "contig longer than 2^31");
    // Without a FASTA reference nothing is known about the loci behind a contig's last read (no pileup, reference base N:
    // exactly what the loci past the track yield): the track, the granule index and the by-locus stores end with the reads,
    // so a shard of a whole-genome dictionary pays for the loci it covers, not for 3 G loci of empty arrays.
    if (!ref) len = std::min<int64_t>(len, (std::max<int64_t>(contig_end[c], 0) + kGranuleLoci - 1) / kGranuleLoci * kGranuleLoci);
    ci.read_begin = contig_first[c] == ~0ull ? 0 : contig_first[c];
    ci.read_end = contig_first[c] == ~0ull ? 0 : contig_last[c];
    ci.length = (int32_t)len;
    ci.n_words = (int32_t)((len + 31) / 32);
    ci.n_grans = (int32_t)((len + kGranuleLoci - 1) / kGranuleLoci);
    ci.word_off = (uint32_t)word_off;
    ci.gran_off = (uint32_t)gran_off;
    ci.pad_ = 0;
    word_off += (uint64_t)ci.n_words;
    gran_off += (uint64_t)ci.n_grans;
    if (word_off >= 0xFFFFFF00ull) fail(GUAC_ERR_UNSUPPORTED, "reference track too large");
  }
  out.total_words = word_off;
  out.total_grans = gran_off;

  out.pairs.alloc(pair_total + 8);
  out.xmask.alloc(pair_total + 8);
  out.nm.alloc(n);
  out.del_start.alloc(n);
  out.del_md.alloc(n);
  out.del_len.alloc(n);
  if (out.has_qualities) out.qc.alloc((size_t)n_bases + 64);
  out.trk_lo.alloc(word_off + 1);
  out.trk_hi.alloc(word_off + 1);
  out.trk_std.alloc(word_off + 1);
  conflict.alloc(word_off + 1);
  out.gran_first.alloc(gran_off + 1);
  out.gran_last.alloc(gran_off + 1);
  gran_count.alloc(gran_off + 1);
  if (n && ctx->difference_lists) {
    out.gs_wide = false;
    alloc_streams(out, (uint64_t)n * 3 + gran_off * 8 + 4096);
  }
  const bool want_rows = n && ctx->difference_lists && out.has_qualities;
  if (want_rows) {  // column pairs: ~1.3 slots per base (a word's columns run to its deepest locus); rows: the few general reads
    out.total_words = word_off;
    // (at most one partly filled block per word that holds reads: a sparse read set over a large dictionary pays for its reads)
    alloc_rows(out, (uint64_t)((double)n_bases * 1.3 / 256.0) + std::min<uint64_t>(word_off, n * 8) + 1024, n / 4 + 1024);
  }
  // (everything is allocated: from here to the last pack kernel the device works without waiting for the host)
  CUDA_OK(cudaEventRecord(ctx->ev[0], st));
  CUDA_OK(cudaMemsetAsync(out.pairs.p, 0, out.pairs.bytes(), st));
  CUDA_OK(cudaMemsetAsync(out.xmask.p, 0, out.xmask.bytes(), st));
  CUDA_OK(cudaMemsetAsync(out.trk_lo.p, 0, out.trk_lo.bytes(), st));
  CUDA_OK(cudaMemsetAsync(out.trk_hi.p, 0, out.trk_hi.bytes(), st));
  CUDA_OK(cudaMemsetAsync(out.trk_std.p, 0, out.trk_std.bytes(), st));
  CUDA_OK(cudaMemsetAsync(conflict.p, 0, conflict.bytes(), st));
  CUDA_OK(cudaMemsetAsync(out.gran_first.p, 0xFF, out.gran_first.bytes(), st));
  CUDA_OK(cudaMemsetAsync(out.gran_last.p, 0, out.gran_last.bytes(), st));
  CUDA_OK(cudaMemsetAsync(gran_count.p, 0, gran_count.bytes(), st));
  if (ref) {
    h2d(ctx, out.fasta_off, ref->base_off, (size_t)ref->n_contigs + 1);
    h2d(ctx, out.fasta, ref->bases, (size_t)ref->base_off[ref->n_contigs], 16);
  }
  h2d(ctx, out.d_contigs, out.contigs.data(), out.contigs.size());
  CUDA_OK(cudaMemsetAsync(ctx->d_counters, 0, 16 * sizeof(unsigned long long), st));
  out.h2d_bytes = on_device ? out.d_contigs.bytes() + out.fasta.bytes()
                  : b2      ? n * (4 + 4 + 4 + (b2->read_length ? 0 : 4) + 1 + 1) + out.cigar.bytes() + (n_bases + 1) / 2 + out.qual.bytes() + out.md.bytes() +
                                  out.d_contigs.bytes() + out.fasta.bytes()
                            : n * (4 + 8 + 8 + 8 + 1 + 1 + (b->sample ? 4 : 0) + 8) + out.cigar.bytes() + out.seq.bytes() + out.qual.bytes() +
                                  out.md.bytes() + out.d_contigs.bytes() + out.fasta.bytes();

  if (tr.on) { cudaStreamSynchronize(cs); cudaStreamSynchronize(st); }
  tr.lap("alloc + h2d");
  PackArgs A;
  A.R = out.view();
  A.rec_w = out.rec.p;
  A.pairs_w = out.pairs.p;
  A.xmask_w = out.xmask.p;
  A.qc_w = out.has_qualities ? out.qc.p : nullptr;
  A.nm_w = out.nm.p;
  A.del_start_w = out.del_start.p;
  A.del_md_w = out.del_md.p;
  A.del_len_w = out.del_len.p;
  A.md_w = out.md.p;
  A.trk_lo_w = out.trk_lo.p;
  A.trk_hi_w = out.trk_hi.p;
  A.trk_std_w = out.trk_std.p;
  A.conflict_w = conflict.p;
  A.gran_first_w = out.gran_first.p;
  A.gran_last_w = out.gran_last.p;
  A.gran_count_w = gran_count.p;
  A.read_contig = d_read_contig.p;
  A.err = ctx->d_err;
  A.counters = ctx->d_counters;

  out.pack_launches = n ? 5 : 0;  // header kernel + scan
  bool finished = false;  // the by-locus stores were built chunk by chunk
  if (n) {
    A.r_begin = 0;
    A.r_end = n;
    k_granule_index<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>(A);
    k_granule_max<<<grid_for(gran_off, 256, ctx->sm_count), 256, 0, st>>>(A, (uint32_t)gran_off);
    out.pack_launches += 2;
    // Reads are sorted by start: once a copy chunk's reads went through the MD walk, the track in front of the next chunk's
    // first read is final, and so are its granules' streams and its words' rows.  They are built right away, underneath the
    // copy of the following chunks; what remains after the last chunk has landed is an eighth of the work.
    const bool finish_by_chunk = !ref && !on_device && n_copy_chunks > 1 && ctx->pack_overlap;
    uint64_t g_done = 0, w_done = 0;
    bool first_finish = true;
    for (int k = 0; k < n_copy_chunks; ++k) {  // bases -> planes and the MD walk, chunk by chunk as the copies land
      A.r_begin = chunk_read[k];
      A.r_end = chunk_read[k + 1];
      if (A.r_end == A.r_begin) continue;
      const uint64_t nr = A.r_end - A.r_begin;
      CUDA_OK(cudaStreamWaitEvent(st, ctx->copy_ev[k], 0));
      k_pack_bases<<<(int)std::min<uint64_t>((nr + kPackReads - 1) / kPackReads, (uint64_t)ctx->sm_count * 8), 256, 0, st>>>(A);
      k_md_track<<<grid_for(nr, 128, ctx->sm_count), 128, 0, st>>>(A);
      out.pack_launches += 2;
      if (!finish_by_chunk) continue;
      const bool last_chunk = chunk_read[k + 1] >= n;
      uint64_t g_end = gran_off, w_end = word_off;
      if (!last_chunk) {  // the granule of the next chunk's first read
        const uint64_t r = chunk_read[k + 1];
        uint32_t c = 0;
        int64_t s0 = 0;
        if (b2) {
          c = (uint32_t)(std::upper_bound(b2->contig_read_off, b2->contig_read_off + b2->n_contigs + 1, r) - b2->contig_read_off - 1);
          s0 = b2->start[r];
        } else {
          c = (uint32_t)b->contig[r];
          s0 = b->start[r];
        }
        const ContigInfo& ci = out.contigs[std::min<uint32_t>(c, b->n_contigs - 1)];
        const uint64_t lg = std::min<uint64_t>((uint64_t)std::max<int64_t>(s0, 0) >> kGranuleShift, (uint64_t)ci.n_grans);
        g_end = (uint64_t)ci.gran_off + lg;
        w_end = (uint64_t)ci.word_off + std::min<uint64_t>(lg * (kGranuleLoci / 32), (uint64_t)ci.n_words);
      }
      if (g_end <= g_done && !last_chunk) continue;
      k_track_finish<<<grid_for(w_end - w_done + 1, 256, ctx->sm_count), 256, 0, st>>>(A, (uint32_t)w_done, (uint32_t)w_end);
      k_resolve_conflicts<<<grid_for(w_end - w_done + 1, 128, ctx->sm_count), 128, 0, st>>>(A, b->n_contigs, (uint32_t)w_done, (uint32_t)w_end);
      out.pack_launches += 2;
      if (ctx->difference_lists) {
        launch_expand(ctx, out, /*huge=*/false, out.gs_entries, /*allocated=*/true, g_done, g_end, first_finish, last_chunk);
        out.pack_launches += 1;
      }
      if (want_rows) {
        launch_rows(ctx, out, w_done, w_end, first_finish, last_chunk);
        out.pack_launches += 1;
      }
      first_finish = false;
      g_done = g_end;
      w_done = w_end;
    }
    A.r_begin = 0;
    A.r_end = n;
    if (!ref && !finish_by_chunk) {
      k_track_finish<<<grid_for(word_off, 256, ctx->sm_count), 256, 0, st>>>(A, 0u, (uint32_t)word_off);
      k_resolve_conflicts<<<grid_for(word_off, 128, ctx->sm_count), 128, 0, st>>>(A, b->n_contigs, 0u, (uint32_t)word_off);
      out.pack_launches += 2;
    }
    finished = finish_by_chunk;
  }
  if (ref) {
    k_fasta_track<<<grid_for(word_off, 256, ctx->sm_count), 256, 0, st>>>(A, b->n_contigs);
    out.pack_launches += 1;
  }
  if (n && ctx->difference_lists && !finished) {  // the track is final: every granule's reads as their differences against it
    launch_expand(ctx, out, /*huge=*/false, out.gs_entries, /*allocated=*/true);
    out.pack_launches += 1;
  }
  if (want_rows && !finished) {  // the likelihood callers' rows (the MD walk has located every read's first deletion: classify() uses it)
    launch_rows(ctx, out);
    out.pack_launches += 1;
  }
  CUDA_OK(cudaEventRecord(ctx->ev[1], st));
  CUDA_OK(cudaGetLastError());
  unsigned long long counters[6];
  CUDA_OK(cudaMemcpyAsync(counters, ctx->d_counters, sizeof counters, cudaMemcpyDeviceToHost, st));
  check_device_error(ctx, "guac_reads_pack");
  float ms = 0;
  CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
  out.pack_kernel_ms = ms + header_ms;  // header kernel + scan, then first memset to last pack kernel (waits for copy chunks included)
  out.order_sensitive_loci = counters[0];
  out.max_reads_per_granule = counters[1];
  float rows_ms = 0;
  if (want_rows) CUDA_OK(cudaEventElapsedTime(&rows_ms, ctx->ev_rows[0], ctx->ev_rows[1]));
  if (n && ctx->difference_lists) settle_streams(ctx, out, counters[2], counters[3] != 0, /*was_huge=*/false);
  if (want_rows) {
    if (counters[4] > out.q_cap_pairs || counters[5] > out.q_cap_groups) settle_rows(ctx, out, counters[4], counters[5]);
    else out.rows_ms = rows_ms;
  }
  tr.lap("kernels");
  nvtx_pop();
}

// ---- tiles over the requested loci -------------------------------------------------------------------------------------------
uint64_t build_tiles(const guac_reads& reads, const guac_locus_range* ranges, size_t n_ranges, std::vector<TileDesc>& tiles) {
  uint64_t requested = 0;
  for (size_t i = 0; i < n_ranges; ++i) {
    const guac_locus_range& r = ranges[i];
    if (r.contig < 0 || (uint32_t)r.contig >= reads.n_contigs) fail(GUAC_ERR_INVALID_ARGUMENT, "locus range %zu: contig out of range", i);
    if (r.start < 0 || r.end < r.start) fail(GUAC_ERR_INVALID_ARGUMENT, "locus range %zu: bad bounds", i);
    requested += (uint64_t)(r.end - r.start);
    const ContigInfo& ci = reads.contigs[r.contig];
    int64_t s = r.start, e = std::min<int64_t>(r.end, ci.length);  // loci past the track hold no reads
    for (int64_t t = s / kWarpLoci; t * kWarpLoci < e; ++t) {  // one descriptor per granule = per warp
      TileDesc td;
      td.contig = r.contig;
      td.word0 = (int32_t)(t * kWarpWords);
      td.locus_begin = (int32_t)std::max<int64_t>(s, t * kWarpLoci);
      td.locus_end = (int32_t)std::min<int64_t>(e, (t + 1) * kWarpLoci);
      td.gran = ci.gran_off + (uint32_t)t;
      td.trk_word = ci.word_off + (uint32_t)td.word0;
      td.n_words = std::min<int32_t>(kWarpWords, ci.n_words - td.word0);
      td.pad_ = 0;
      if (td.locus_end > td.locus_begin) tiles.push_back(td);
    }
  }
  return requested;
}

template <int MODE>
void launch_tile(bool streams, bool wide, int grid, cudaStream_t st, const DevReads& R, const TileDesc* tiles, const CallParams& prm, const DevOut& out) {
  const uint32_t n = (uint32_t)grid;  // descriptors (one per warp)
  const int ctas = (int)((n + kWarpsPerCta - 1) / kWarpsPerCta);
  if (streams) {
    if (wide) k_call_tile<true, MODE><<<ctas, kTileThreads, kWarpsPerCta * sizeof(CallSmem<true>), st>>>(R, tiles, n, prm, out);
    else k_call_tile<false, MODE><<<ctas, kTileThreads, kWarpsPerCta * sizeof(CallSmem<false>), st>>>(R, tiles, n, prm, out);
  } else {  // GUAC_OPT_DIFFERENCE_LISTS = 0: planes / CIGAR walk inside the call
    if (wide) k_pileup_tile<uint64_t, MODE><<<ctas, kTileThreads, kWarpsPerCta * sizeof(WarpSmem<uint64_t, MODE>), st>>>(R, tiles, n, prm, out);
    else k_pileup_tile<uint32_t, MODE><<<ctas, kTileThreads, kWarpsPerCta * sizeof(WarpSmem<uint32_t, MODE>), st>>>(R, tiles, n, prm, out);
  }
}

template <typename CntT, int MODE>
void set_smem_attr() {
  CUDA_OK(cudaFuncSetAttribute(k_pileup_tile<CntT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kWarpsPerCta * sizeof(WarpSmem<CntT, MODE>))));
}
void set_all_smem_attrs() {
  set_smem_attr<uint32_t, 0>(); set_smem_attr<uint64_t, 0>(); set_smem_attr<uint32_t, 1>(); set_smem_attr<uint64_t, 1>();
  CUDA_OK(cudaFuncSetAttribute(k_call_tile<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kWarpsPerCta * sizeof(CallSmem<true>))));
  CUDA_OK(cudaFuncSetAttribute(k_call_tile<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kWarpsPerCta * sizeof(CallSmem<true>))));
}

// the tile list of (reads, ranges) is cached in the context: a repeated call does not rebuild or re-upload it
uint64_t prepare_tiles(guac_ctx* ctx, const guac_reads& reads, const guac_locus_range* ranges, size_t n_ranges, uint64_t* tile_loci) {
  bool same = ctx->tiles_key_reads == reads.id && ctx->tiles_key_ranges.size() == n_ranges;
  for (size_t i = 0; same && i < n_ranges; ++i)
    same = ctx->tiles_key_ranges[i].contig == ranges[i].contig && ctx->tiles_key_ranges[i].start == ranges[i].start &&
           ctx->tiles_key_ranges[i].end == ranges[i].end;
  uint64_t requested = 0;
  for (size_t i = 0; i < n_ranges; ++i) requested += (uint64_t)std::max<int64_t>(0, ranges[i].end - ranges[i].start);
  if (!same) {
    check_ranges_disjoint(ranges, n_ranges);
    std::vector<TileDesc> tiles;
    build_tiles(reads, ranges, n_ranges, tiles);
    uint64_t tl = 0;
    for (auto& t : tiles) tl += (uint64_t)(t.locus_end - t.locus_begin);
    ctx->tiles.ensure(std::max<size_t>(tiles.size(), 1) * sizeof(TileDesc));
    if (!tiles.empty())
      CUDA_OK(cudaMemcpyAsync(ctx->tiles.p, tiles.data(), tiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    ctx->n_tiles = tiles.size();
    ctx->tiles_in_order = true;  // are the tiles in canonical (contig, start) order?  (then so are the records laid out tile by tile)
    for (size_t i = 1; i < tiles.size() && ctx->tiles_in_order; ++i)
      ctx->tiles_in_order = std::make_pair(tiles[i - 1].contig, tiles[i - 1].locus_begin) < std::make_pair(tiles[i].contig, tiles[i].locus_begin);
    ctx->tiles_key_loci = tl;
    ctx->tiles_key_reads = reads.id;
    ctx->tiles_key_ranges.assign(ranges, ranges + n_ranges);
  }
  *tile_loci = ctx->tiles_key_loci;
  return requested;
}

// runs the tile kernel (+ the exact kernel on the loci it defers) for one read set
void run_pileup(guac_ctx* ctx, const guac_reads& reads, const guac_locus_range* ranges, size_t n_ranges, CallParams prm,
                guac_result& res) {
  uint64_t tile_loci = 0;
  const uint64_t requested = prepare_tiles(ctx, reads, ranges, n_ranges, &tile_loci);
  res.stats.reads_total = reads.n;
  res.stats.loci_requested = requested;
  res.stats.order_sensitive_loci = reads.order_sensitive_loci;
  res.sample = reads.sample;
  res.owner = ctx;
  res.generation = ++ctx->generation;
  auto append_rows_past_track = [&] {  // requested loci past the end of the track hold no reads: empty pileups, reference base N
    if (prm.mode != 1 || prm.skip_empty) return;
    for (size_t i = 0; i < n_ranges; ++i) {
      const ContigInfo& ci = reads.contigs[ranges[i].contig];
      for (int64_t x = std::max<int64_t>(ranges[i].start, ci.length); x < ranges[i].end; ++x) {
        guac_locus_counts z{};
        z.locus = x;
        z.contig = ranges[i].contig;
        z.reference_base = 'N';
        res.counts.push_back(z);
      }
    }
  };
  if (ctx->n_tiles == 0) {
    append_rows_past_track();
    res.stats.records = res.counts.size();
    res.stats.loci_visited = prm.skip_empty ? 0 : requested;
    return;
  }
  if (reads.n_contigs > 65535) fail(GUAC_ERR_UNSUPPORTED, "more than 65535 contigs");
  if (!ctx->smem_attrs_done) {  // per context: function attributes belong to the context's device
    set_all_smem_attrs();
    ctx->smem_attrs_done = true;
  }
  cudaStream_t st = ctx->stream, st2 = ctx->stream2, st3 = ctx->stream3;
  const bool counts_mode = prm.mode == 1;
  const bool dense = counts_mode || prm.emit_ref || prm.emit_no_call;
  // counts mode: rows (guac_locus_counts) in HBM, copied afterwards.  Germline mode: the tile kernel's single-base records
  // are compact (8 bytes) in HBM, ordered on the device and streamed to the pinned host block of the result; the exact
  // kernel's few general records and their allele bytes are staged in HBM as well and copied there by k_general_to_host.
  uint64_t cap_rec = counts_mode ? tile_loci + 16 : std::max<uint64_t>(4096, tile_loci / 256);
  uint64_t cap_compact = counts_mode ? 8 : (dense ? tile_loci + tile_loci / 8 + 16 : std::max<uint64_t>(4096, tile_loci / 64));
  uint64_t cap_slow = std::max<uint64_t>(4096, tile_loci / 32);
  uint64_t cap_pool = kPoolDynOff + std::max<uint64_t>(65536, tile_loci / 64);
  if (counts_mode) cap_rec = std::max<uint64_t>(cap_rec, ctx->out_rec.n / sizeof(guac_locus_counts));
  cap_compact = std::max<uint64_t>(cap_compact, ctx->out_compact.n / 8 / 4);  // (up to four segments share the buffers)
  cap_slow = std::max<uint64_t>(cap_slow, ctx->out_slow.n / sizeof(SlowLocus) / 4);
  if (counts_mode) cap_pool = std::max<uint64_t>(cap_pool, ctx->out_pool.n);
  const bool streams = reads.gs_hdr.n != 0;
  // narrow counter fields (8 bits) unless the store is wide; the tile kernel reports a possible overflow and we widen
  bool wide = streams ? reads.gs_wide : reads.max_reads_per_granule >= 2048;
  double tile_ms = 0, exact_ms = 0;
  int launches = 0;
  static const bool x_tl = getenv("GUAC_TRACE") != nullptr;
  cudaEvent_t* xe = ctx->trace_ev;
  if (x_tl && !xe[0])
    for (int i = 0; i < 5; ++i) CUDA_OK(cudaEventCreate(&xe[i]));
  for (int attempt = 0; attempt < 8; ++attempt) {
    if (cap_rec >= 0xFFFFFFF0ull || cap_compact >= 0xFFFFFFF0ull || cap_slow >= 0xFFFFFFF0ull || cap_pool >= 0xFFFFFFF0ull)
      fail(GUAC_ERR_UNSUPPORTED, "too many output records for one call: split the loci ranges");
    const size_t full_at = ((size_t)cap_pool + 63) & ~(size_t)63;
    const size_t compact_at = (full_at + (size_t)cap_rec * sizeof(guac_threshold_record) + 63) & ~(size_t)63;
    if (!counts_mode) {
      const size_t want = compact_at + (size_t)cap_compact * 8 + 64;
      if (res.block && res.block_bytes < want) {
        ctx->pinned->give(res.block, res.block_bytes);
        res.block = nullptr;
      }
      if (!res.block) {
        res.pool = ctx->pinned;
        res.block = ctx->pinned->take(want, &res.block_bytes);
        if (!res.block) fail(GUAC_ERR_OOM, "pinned host allocation of %zu bytes failed", want);
      }
      unsigned char* hs = (unsigned char*)res.block;  // the static head of the allele pool: "<ALT>" and the 256 single bytes
      memset(hs, 0, kPoolDynOff);
      memcpy(hs, "<ALT>", 5);
      for (int v = 0; v < 256; ++v) hs[kPoolByteOff + v] = (uint8_t)v;
      ctx->out_compact.ensure(cap_compact * 8 + 16);
      ctx->out_rec.ensure(cap_rec * sizeof(guac_threshold_record) + 16);  // the exact kernel's general records and allele bytes:
      if (ctx->out_pool.ensure(cap_pool + 16)) ctx->pool_head_ready = false;  // HBM first, one coalesced copy to the block
    } else {
      ctx->out_rec.ensure(cap_rec * sizeof(guac_locus_counts));
      if (ctx->out_pool.ensure(cap_pool)) ctx->pool_head_ready = false;
    }
    // After the tile kernel, the exact kernel (second stream) and the egress of the compact records (third stream) run side by
    // side.  A call may also be cut into up to four segments of tiles (GUAC_OPT_SEGMENTS) whose side work overlaps the next
    // segment's tile kernel; measured on B200 that interleaving costs the tile kernel more than it hides, so the default is 1.
    const int n_seg = counts_mode ? 1 : ctx->segments;
    const uint64_t cap_seg = counts_mode ? 8 : cap_compact;   // compact records one segment may hold (each gets the full allowance)
    ctx->out_slow.ensure((size_t)n_seg * cap_slow * sizeof(SlowLocus));
    if (!counts_mode) {
      ctx->out_compact.ensure((size_t)n_seg * cap_seg * 8 + 16);
      ctx->sort_rec.ensure((size_t)(n_seg + 1) * cap_seg * 8 + 16);   // per-segment grouped records, then the contiguous copy
    }
    CUDA_OK(cudaMemsetAsync(ctx->d_counters, 0, 16 * sizeof(unsigned long long), st));
    DevOut out;
    out.trec = (guac_threshold_record*)ctx->out_rec.p;
    out.crec = (guac_locus_counts*)ctx->out_rec.p;
    out.cap_rec = (uint32_t)cap_rec;
    out.compact = (unsigned long long*)ctx->out_compact.p;
    out.cap_compact = counts_mode ? 0u : (uint32_t)cap_seg;
    out.pool = ctx->out_pool.p;
    out.cap_pool = (uint32_t)cap_pool;
    out.slow = (SlowLocus*)ctx->out_slow.p;
    out.cap_slow = (uint32_t)cap_slow;
    out.slow_ctr = 8;
    out.compact_ctr = 12;
    out.counters = ctx->d_counters;
    out.err = ctx->d_err;
    out.work = ctx->d_counters + kStatusBytes / sizeof(unsigned long long);  // (zero at rest: the exact kernel resets its tickets itself)
    const DevReads R = reads.view();
    const TileDesc* d_tiles = (const TileDesc*)ctx->tiles.p;
    unsigned long long* h_compact = counts_mode ? nullptr : (unsigned long long*)((unsigned char*)res.block + compact_at);
    unsigned long long* d_contig = counts_mode ? nullptr : (unsigned long long*)ctx->sort_rec.p + (size_t)n_seg * cap_seg;
    // canonical order on the device: every tile sorts its own few records, the egress kernels lay the tiles out in order
    const bool device_sort = !counts_mode && ctx->sort_records && !dense && streams;
    const uint64_t nt_all = ctx->n_tiles;
    const uint64_t nt_pad = (nt_all + 11) & ~3ull;  // (tile_n is read 16 bytes at a time; the segments start at multiples of 4)
    if (device_sort) {
      ctx->sort_bins.ensure(3 * nt_pad);  // tile_base | tile_n | prefix (calls over more than 64 K tiles)
      ctx->scan_totals.ensure((size_t)n_seg * ((nt_all + kScanChunk - 1) / kScanChunk + 2));
      // (tile_n needs no clearing: every tile of the call writes its count, k_call_tile's flush_stage)
      out.tile_base = ctx->sort_bins.p;
      out.tile_n = ctx->sort_bins.p + nt_pad;
    } else {
      out.tile_base = nullptr;
      out.tile_n = nullptr;
    }
    out.tile0 = 0;
    nvtx_push("guac tile + exact kernels + record egress");
    CUDA_OK(cudaEventRecord(ctx->ev[0], st));
    CUDA_OK(cudaStreamWaitEvent(st3, ctx->ev[0], 0));  // (the counters are cleared)
    for (int seg = 0; seg < n_seg; ++seg) {
      const uint64_t t0 = ((ctx->n_tiles * (uint64_t)seg / n_seg) + 3) & ~3ull, t1 = seg + 1 == n_seg ? ctx->n_tiles : ((ctx->n_tiles * (uint64_t)(seg + 1) / n_seg) + 3) & ~3ull;
      DevOut so = out;
      so.tile0 = (uint32_t)t0;
      so.slow = out.slow + (size_t)seg * cap_slow;
      so.slow_ctr = 8u + (uint32_t)seg;
      so.compact = out.compact + (size_t)seg * cap_seg;
      so.compact_ctr = 12u + (uint32_t)seg;
      const int nt = (int)(t1 - t0);
      if (nt > 0) {
        if (counts_mode) launch_tile<1>(streams, wide, nt, st, R, d_tiles + t0, prm, so);
        else launch_tile<0>(streams, wide, nt, st, R, d_tiles + t0, prm, so);
        launches += 1;
      }
      // the side streams read the segment's counters from device memory: no host round trip in between
      cudaEvent_t done = seg + 1 == n_seg ? ctx->ev[1] : ctx->seg_ev[seg];
      CUDA_OK(cudaEventRecord(done, st));
      CUDA_OK(cudaStreamWaitEvent(st2, done, 0));
      launches += 1;
      if (!counts_mode) {
        CUDA_OK(cudaStreamWaitEvent(st3, done, 0));
        if (device_sort && nt > 0) {
          const uint32_t* block_before = nullptr;
          if ((uint64_t)nt > 65536) {  // (the in-kernel sums are quadratic in the number of tiles)
            const uint64_t n_chunks = ((uint64_t)nt + kScanChunk - 1) / kScanChunk;
            uint64_t* totals = ctx->scan_totals.p + (size_t)seg * ((nt_all + kScanChunk - 1) / kScanChunk + 2);
            uint32_t* prefix = ctx->sort_bins.p + 2 * nt_pad + t0;
            k_scan_totals<<<(unsigned)n_chunks, 256, 0, st3>>>(out.tile_n + t0, (uint64_t)nt, totals);
            k_scan_chunks<<<1, 1024, 0, st3>>>(totals, n_chunks);
            k_scan_final<uint32_t><<<(unsigned)n_chunks, 256, 0, st3>>>(out.tile_n + t0, (uint64_t)nt, totals, prefix);
            block_before = prefix;
            launches += 3;
          }
          if (x_tl) cudaEventRecord(xe[2], st3);
          k_rec_gather<<<(unsigned)(((uint64_t)nt + 255) / 256), 256, 0, st3>>>(so.compact, out.tile_base + t0, out.tile_n + t0, block_before, (uint32_t)nt,
                                                                               ctx->d_counters, (uint32_t)seg, so.cap_compact, d_contig, cap_compact);
          if (x_tl) cudaEventRecord(xe[3], st3);
          k_rec_to_host<<<ctx->sm_count, 256, 0, st3>>>(d_contig, ctx->d_counters, (uint32_t)seg, so.cap_compact, h_compact, cap_compact);
          if (x_tl) cudaEventRecord(xe[4], st3);
          launches += 2;
        } else {
          k_rec_flush<<<ctx->sm_count, 256, 0, st3>>>(so.compact, ctx->d_counters, (uint32_t)seg, so.cap_compact, h_compact, d_contig, cap_compact);
          launches += 1;
        }
      }
      // (after the egress kernels in launch order, and 8 CTAs of 48 registers per SM: the egress kernels' 256-thread blocks find
      // room next to them.  At 56 registers the register file was full and the record gather waited for the exact kernel to
      // drain — 90 us instead of 16; from 10 CTAs per SM on k_rec_to_host is starved the same way.)
      if (x_tl) cudaEventRecord(xe[0], st2);
      k_exact_loci<<<ctx->sm_count * 8, kExactWarps * 32, 0, st2>>>(R, so.slow, prm, so);
      if (x_tl) cudaEventRecord(xe[1], st2);
      if (!counts_mode && seg + 1 == n_seg) {  // (the segments' exact kernels run in order on st2 and share the buffers)
        k_general_to_host<<<8, 256, 0, st2>>>(ctx->d_counters, (const uint4*)ctx->out_rec.p, ctx->out_pool.p, (uint4*)((unsigned char*)res.block + full_at),
                                              (uint8_t*)res.block, (uint32_t)cap_rec, (uint32_t)cap_pool);
        launches += 1;
      }
    }
    CUDA_OK(cudaEventRecord(ctx->join_ev, st2));
    CUDA_OK(cudaStreamWaitEvent(st, ctx->join_ev, 0));
    if (!counts_mode) {
      CUDA_OK(cudaEventRecord(ctx->join3_ev, st3));
      CUDA_OK(cudaStreamWaitEvent(st, ctx->join3_ev, 0));
    }
    nvtx_pop();
    nvtx_push("guac status copy");
    const bool device_sorted = device_sort;
    CUDA_OK(cudaEventRecord(ctx->ev[2], st));
    CUDA_OK(cudaGetLastError());
    // counters and the device error word come back in one copy, one synchronisation per call
    unsigned long long* c = ctx->h_counters;
    CUDA_OK(cudaMemcpyAsync(c, ctx->d_counters, kStatusBytes, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    nvtx_pop();
    raise_device_error(ctx, "pileup");
    if (x_tl && !counts_mode && device_sort) {  // (GUAC_TRACE) the side streams' kernels relative to the start of the call
      float t[8] = {0};
      cudaEventElapsedTime(&t[0], ctx->ev[0], ctx->ev[1]);
      for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&t[1 + i], ctx->ev[0], xe[i]);
      cudaEventElapsedTime(&t[6], ctx->ev[0], ctx->ev[2]);
      fprintf(stderr, "[guac timeline] tile end %.1f us | exact %.1f..%.1f | gather %.1f..%.1f | to_host ..%.1f | joined %.1f\n", t[0] * 1e3, t[1] * 1e3,
              t[2] * 1e3, t[3] * 1e3, t[4] * 1e3, t[5] * 1e3, t[6] * 1e3);
    }
    if (device_sort && c[7]) {
      // a tile held more records than it orders in shared memory: its surplus is not part of any tile's slice.  Copy every
      // segment's records as they are instead (still on the device); the host orders them when the view is built.
      for (int seg = 0; seg < n_seg; ++seg)
        k_rec_flush<<<ctx->sm_count, 256, 0, st>>>(out.compact + (size_t)seg * cap_seg, ctx->d_counters, (uint32_t)seg, (uint32_t)cap_seg, h_compact, d_contig, cap_compact);
      CUDA_OK(cudaGetLastError());
      CUDA_OK(cudaStreamSynchronize(st));
      launches += n_seg;
    }
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    tile_ms += ms;
    CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]));
    exact_ms += ms;
    if (c[5]) {  // a counter field may have wrapped: widen and rerun
      if (wide) fail(GUAC_ERR_UNSUPPORTED, "pileup deeper than 65535 reads");
      wide = true;
      if (streams) {  // the store's start / end fields widen with the counters: its streams are built again
        guac_reads& rw = const_cast<guac_reads&>(reads);
        rw.gs_wide = true;
        launch_expand(ctx, rw, rw.max_reads_per_granule > 65535, rw.gs_entries);
        unsigned long long e[2];
        CUDA_OK(cudaMemcpyAsync(e, ctx->d_counters + 2, sizeof e, cudaMemcpyDeviceToHost, st));
        check_device_error(ctx, "difference streams");
        settle_streams(ctx, rw, e[0], e[1] != 0, rw.max_reads_per_granule > 65535);
      }
      continue;
    }
    unsigned long long max_slow = 0, max_seg = 0, n_compact_total = 0, n_slow_total = 0;
    for (int sg = 0; sg < 4; ++sg) {
      max_slow = std::max(max_slow, c[8 + sg]);
      max_seg = std::max(max_seg, c[12 + sg]);
      n_compact_total += c[12 + sg];
      n_slow_total += c[8 + sg];
    }
    if (max_slow > cap_slow || c[0] > cap_rec || max_seg > cap_seg || n_compact_total > cap_compact || kPoolDynOff + c[1] > cap_pool) {
      cap_slow = std::max<uint64_t>(cap_slow, max_slow + max_slow / 8 + 16);
      cap_rec = std::max<uint64_t>(cap_rec, c[0] + c[0] / 8 + 16);
      cap_compact = std::max<uint64_t>(cap_compact, n_compact_total + n_compact_total / 8 + 16);
      cap_pool = std::max<uint64_t>(cap_pool, kPoolDynOff + c[1] + c[1] / 8 + 16);
      continue;
    }
    const uint64_t n_rec = c[0];
    const size_t pool_bytes = (size_t)(kPoolDynOff + c[1]);
    if (counts_mode) {
      const size_t rec_bytes = (size_t)(n_rec * sizeof(guac_locus_counts));
      res.counts.resize((size_t)n_rec);
      if (n_rec) CUDA_OK(cudaMemcpyAsync(res.counts.data(), ctx->out_rec.p, rec_bytes, cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaStreamSynchronize(st));
      res.stats.d2h_bytes = rec_bytes + kStatusBytes;
      append_rows_past_track();
      if (ctx->sort_records)
        std::sort(res.counts.begin(), res.counts.end(), [](const guac_locus_counts& a, const guac_locus_counts& b) {
          return std::make_pair(a.contig, a.locus) < std::make_pair(b.contig, b.locus);
        });
    } else {
      res.general = (guac_threshold_record*)((unsigned char*)res.block + full_at);
      res.n_general = (size_t)n_rec;
      res.compact = h_compact;
      res.d_compact = d_contig;
      res.n_compact = (size_t)n_compact_total;
      res.n_records = res.n_general + res.n_compact;
      res.bytes = (const uint8_t*)res.block;
      res.n_bytes = pool_bytes;
      res.want_sorted = ctx->sort_records != 0;
      res.compact_sorted = device_sorted && c[7] == 0 && ctx->tiles_in_order;
      res.stats.d2h_bytes = (pool_bytes - kPoolDynOff) + n_rec * sizeof(guac_threshold_record) + res.n_compact * 8 + kStatusBytes;
    }
    res.stats.loci_visited = c[3] + (prm.skip_empty ? 0 : requested - tile_loci);  // loci past the track: empty pileups
    res.stats.tie_loci = c[4];
    res.stats.records = counts_mode ? res.counts.size() : res.n_records;
    res.stats.kernel_ms = tile_ms + exact_ms;
    res.stats.tile_kernel_ms = tile_ms;
    res.stats.exact_kernel_ms = exact_ms;
    res.stats.kernel_launches = (uint64_t)launches;
    res.stats.exact_loci = n_slow_total;
    return;
  }
  fail(GUAC_ERR_CUDA, "output buffers did not converge");
}

// guac_threshold_record view of a germline result: the compact records expanded and merged with the exact kernel's general
// records, in canonical order when the context asked for it.  Built on first use.
void expand_threshold_records(guac_result& r) {
  if (r.expanded_ready) return;
  r.expanded_ready = true;
  if (r.kind != 0 || r.n_records == 0) return;
  const uint8_t* pool = r.bytes;
  std::vector<unsigned long long> sorted_compact;
  const unsigned long long* cr = r.compact;
  if (r.want_sorted && !r.compact_sorted && r.n_compact > 1) {
    sorted_compact.assign(r.compact, r.compact + r.n_compact);
    std::sort(sorted_compact.begin(), sorted_compact.end());
    cr = sorted_compact.data();
  }
  std::vector<guac_threshold_record> gen(r.general, r.general + r.n_general);
  auto less_general = [pool](const guac_threshold_record& a, const guac_threshold_record& b) {
    if (a.contig != b.contig) return a.contig < b.contig;
    if (a.start != b.start) return a.start < b.start;
    if (a.sample != b.sample) return a.sample < b.sample;
    int c = memcmp(pool + a.ref_off, pool + b.ref_off, std::min(a.ref_len, b.ref_len));
    if (c != 0) return c < 0;
    if (a.ref_len != b.ref_len) return a.ref_len < b.ref_len;
    c = memcmp(pool + a.alt_off, pool + b.alt_off, std::min(a.alt_len, b.alt_len));
    if (c != 0) return c < 0;
    return a.alt_len < b.alt_len;
  };
  if (r.want_sorted && gen.size() > 1) {
    // (contig, start) as one integer first — the exact kernel hands its records out in ticket order, and a comparator that
    // looks at allele bytes for every pair was half of this function's time —, the full order only inside a locus' own records
    std::vector<std::pair<unsigned long long, uint32_t>> keyed(gen.size());
    for (size_t q = 0; q < gen.size(); ++q)
      keyed[q] = {(((unsigned long long)(uint32_t)gen[q].contig) << 32) | (unsigned long long)(uint32_t)gen[q].start, (uint32_t)q};
    std::sort(keyed.begin(), keyed.end());
    std::vector<guac_threshold_record> ordered(gen.size());
    for (size_t q = 0; q < gen.size(); ++q) ordered[q] = gen[keyed[q].second];
    for (size_t a = 0; a < ordered.size();) {
      size_t b = a + 1;
      while (b < ordered.size() && keyed[b].first == keyed[a].first) ++b;
      if (b - a > 1) std::sort(ordered.begin() + (ptrdiff_t)a, ordered.begin() + (ptrdiff_t)b, less_general);
      a = b;
    }
    gen.swap(ordered);
  }
  r.expanded.resize(r.n_records);
  auto widen = [&](unsigned long long v) {
    guac_threshold_record t;
    t.start = (int64_t)(uint32_t)(v >> 16);
    t.contig = (int32_t)(v >> 48);
    t.sample = r.sample;
    const uint32_t alt = (uint32_t)(v >> 13) & 7u, rcode = (uint32_t)(v >> 11) & 3u;
    t.ref_off = kPoolByteOff + (uint32_t)"ACGT"[rcode];
    t.ref_len = 1;
    t.alt_off = alt == 0 ? kPoolAltOff : kPoolByteOff + (uint32_t)"ACGT"[alt - 1];
    t.alt_len = alt == 0 ? 5 : 1;
    t.gt[0] = (uint8_t)((v >> 9) & 3u);
    t.gt[1] = (uint8_t)((v >> 7) & 3u);
    t.tie = (uint8_t)((v >> 6) & 1u);
    t.pad_ = 0;
    return t;
  };
  size_t i = 0, j = 0, k = 0;
  if (r.want_sorted) {  // merge by (contig, start): a locus is decided by one kernel or the other, never both
    while (i < r.n_compact && j < gen.size()) {
      const unsigned long long key = (((unsigned long long)(uint32_t)gen[j].contig) << 32) | (unsigned long long)(uint32_t)gen[j].start;
      if ((cr[i] >> 16) <= key) r.expanded[k++] = widen(cr[i++]);
      else r.expanded[k++] = gen[j++];
    }
  }
  while (i < r.n_compact) r.expanded[k++] = widen(cr[i++]);
  while (j < gen.size()) r.expanded[k++] = gen[j++];
}

}  // namespace

// ---- exported C ABI -------------------------------------------------------------------------------------------------------
extern "C" {

int guac_abi_version(void) { return GUAC_ABI_VERSION; }

const char* guac_status_string(guac_status s) {
  switch (s) {
    case GUAC_OK: return "GUAC_OK";
    case GUAC_ERR_INVALID_ARGUMENT: return "GUAC_ERR_INVALID_ARGUMENT";
    case GUAC_ERR_UNSORTED_READS: return "GUAC_ERR_UNSORTED_READS";
    case GUAC_ERR_CONTIG_ORDER: return "GUAC_ERR_CONTIG_ORDER";
    case GUAC_ERR_INVALID_CIGAR: return "GUAC_ERR_INVALID_CIGAR";
    case GUAC_ERR_MISSING_MD: return "GUAC_ERR_MISSING_MD";
    case GUAC_ERR_MULTIPLE_REFERENCE_BASES: return "GUAC_ERR_MULTIPLE_REFERENCE_BASES";
    case GUAC_ERR_BAD_QUALITY: return "GUAC_ERR_BAD_QUALITY";
    case GUAC_ERR_CUDA: return "GUAC_ERR_CUDA";
    case GUAC_ERR_OOM: return "GUAC_ERR_OOM";
    case GUAC_ERR_NO_DEVICE: return "GUAC_ERR_NO_DEVICE";
    case GUAC_ERR_UNSUPPORTED: return "GUAC_ERR_UNSUPPORTED";
  }
  return "GUAC_ERR_?";
}

guac_status guac_ctx_create(int device, guac_ctx** out) {
  if (!out) return GUAC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return GUAC_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GUAC_ERR_NO_DEVICE;
  if (prop.major != 10) return GUAC_ERR_NO_DEVICE;  // kernels are built for sm_100a only
  guac_ctx* ctx = new (std::nothrow) guac_ctx();
  if (!ctx) return GUAC_ERR_OOM;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  guac_status s = guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(device));
    CUDA_OK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&ctx->stream3, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&ctx->join_ev, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&ctx->join3_ev, cudaEventDisableTiming));
    for (auto& e : ctx->seg_ev) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUDA_OK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (auto& e : ctx->copy_ev) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));

    CUDA_OK(cudaMalloc((void**)&ctx->d_counters, kStatusBytes + 8 * sizeof(unsigned long long)));  // status block + the exact kernel's tickets
    CUDA_OK(cudaMemset(ctx->d_counters, 0, kStatusBytes + 8 * sizeof(unsigned long long)));
    ctx->d_err = reinterpret_cast<DevError*>(ctx->d_counters + 16);
    CUDA_OK(cudaMemset(ctx->d_counters, 0, kStatusBytes));
    CUDA_OK(cudaMallocHost((void**)&ctx->h_counters, kStatusBytes));
    for (auto& e : ctx->ev) CUDA_OK(cudaEventCreate(&e));
    for (auto& e : ctx->ev_rows) CUDA_OK(cudaEventCreate(&e));
    somatic_init_tables(ctx);
  });
  if (s != GUAC_OK) {
    guac_ctx_destroy(ctx);
    return s;
  }
  *out = ctx;
  return GUAC_OK;
}

void guac_ctx_destroy(guac_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->d_counters) cudaFree(ctx->d_counters);
  if (ctx->d_tables) cudaFree(ctx->d_tables);
  if (ctx->d_hist) cudaFree(ctx->d_hist);
  if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->h_pack) cudaFreeHost(ctx->h_pack);
  for (auto& e : ctx->ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : ctx->ev_rows)
    if (e) cudaEventDestroy(e);
  ctx->out_rec.release();
  ctx->out_pool.release();
  ctx->out_slow.release();
  ctx->out_compact.release();
  ctx->sort_rec.release();
  ctx->sort_bins.release();
  ctx->tiles.release();
  tl_dev_cache.trim();
  for (auto& e : ctx->copy_ev)
    if (e) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->join_ev) cudaEventDestroy(ctx->join_ev);
  if (ctx->join3_ev) cudaEventDestroy(ctx->join3_ev);
  for (auto& e : ctx->trace_ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : ctx->seg_ev)
    if (e) cudaEventDestroy(e);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->stream3) cudaStreamDestroy(ctx->stream3);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

guac_status guac_ctx_set_option(guac_ctx* ctx, int option, int64_t value) {
  if (!ctx) return GUAC_ERR_INVALID_ARGUMENT;
  switch (option) {
    case GUAC_OPT_SORT_RECORDS: ctx->sort_records = value != 0; return GUAC_OK;
    case GUAC_OPT_PACK_QUALITIES: ctx->pack_qualities = value != 0; return GUAC_OK;
    case GUAC_OPT_HOST_THREADS: ctx->host_threads = value > 0 ? (int)value : 0; return GUAC_OK;
    case GUAC_OPT_DIFFERENCE_LISTS: ctx->difference_lists = value != 0; return GUAC_OK;
    case GUAC_OPT_SEGMENTS: ctx->segments = value < 1 ? 1 : value > 4 ? 4 : (int)value; return GUAC_OK;
    case GUAC_OPT_PACK_OVERLAP: ctx->pack_overlap = value != 0; return GUAC_OK;
    case GUAC_OPT_TRIM_CACHE:
      cudaSetDevice(ctx->device);
      cudaStreamSynchronize(ctx->stream);
      tl_dev_cache.trim();
      return GUAC_OK;
  }
  ctx->last_error = "unknown option";
  return GUAC_ERR_INVALID_ARGUMENT;
}

guac_status guac_ctx_timer_start(guac_ctx* ctx) {
  if (!ctx) return GUAC_ERR_INVALID_ARGUMENT;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    CUDA_OK(cudaEventRecord(ctx->ev[4], ctx->stream));
  });
}
guac_status guac_ctx_timer_stop(guac_ctx* ctx, double* elapsed_ms) {
  if (!ctx || !elapsed_ms) return GUAC_ERR_INVALID_ARGUMENT;
  return guarded(ctx, [&] {
    CUDA_OK(cudaEventRecord(ctx->ev[5], ctx->stream));
    CUDA_OK(cudaEventSynchronize(ctx->ev[5]));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
    *elapsed_ms = ms;
  });
}

guac_status guac_host_register(void* ptr, size_t bytes) {
  if (!ptr || !bytes) return GUAC_ERR_INVALID_ARGUMENT;
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return GUAC_OK; }
  return e == cudaSuccess ? GUAC_OK : GUAC_ERR_CUDA;
}
guac_status guac_host_unregister(void* ptr) {
  if (!ptr) return GUAC_ERR_INVALID_ARGUMENT;
  cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) cudaGetLastError();
  return e == cudaSuccess ? GUAC_OK : GUAC_ERR_CUDA;
}

const char* guac_last_error(const guac_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "no context"; }

guac_status guac_reads_pack(guac_ctx* ctx, const guac_read_batch* batch, const guac_reference* ref, guac_reads** out) {
  if (!ctx || !out) return GUAC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    std::unique_ptr<guac_reads> r(new guac_reads());
    try {
      pack_reads(ctx, batch, ref, *r, false);
    } catch (...) {
      cudaStreamSynchronize(ctx->stream);  // copies from the caller's buffers may still be in flight
      throw;
    }
    *out = r.release();
  });
}

guac_status guac_reads_pack_v2(guac_ctx* ctx, const guac_read_batch_v2* batch, const guac_reference* ref, guac_reads** out) {
  if (!ctx || !out || !batch) return GUAC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    std::unique_ptr<guac_reads> r(new guac_reads());
    try {
      pack_reads(ctx, nullptr, ref, *r, false, nullptr, batch);
    } catch (...) {
      cudaStreamSynchronize(ctx->copy_stream);  // copies from the caller's buffers may still be in flight
      cudaStreamSynchronize(ctx->stream);
      throw;
    }
    *out = r.release();
  });
}

guac_status guac_read_batch_compact(const guac_read_batch* batch, int pinned, int fixed_length, guac_host_batch_v2** out) {
  if (!batch || !out) return GUAC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded(nullptr, [&] {
    std::unique_ptr<guac_host_batch_v2> h(new guac_host_batch_v2());
    h->pinned = pinned != 0;
    compact_batch(*batch, fixed_length != 0, *h);
    *out = h.release();
  });
}

const guac_read_batch_v2* guac_host_batch_v2_view(const guac_host_batch_v2* b) { return b ? &b->view : nullptr; }
uint64_t guac_host_batch_v2_bytes(const guac_host_batch_v2* b) { return b ? b->bytes : 0; }
void guac_host_batch_v2_free(guac_host_batch_v2* b) { delete b; }

static thread_local std::string tl_bam_error;
guac_status guac_bam_load(const char* path, const guac_bam_options* options, guac_host_batch_v2** out) {
  if (!path || !out) return GUAC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  guac_bam_options opt{};
  if (options) opt = *options;
  try {
    std::unique_ptr<guac_host_batch_v2> h(new guac_host_batch_v2());
    h->pinned = opt.pinned != 0;
    bam_load(path, opt, *h);
    *out = h.release();
    return GUAC_OK;
  } catch (const StatusError& e) {
    tl_bam_error = e.msg;
    return e.code;
  } catch (const std::bad_alloc&) {
    tl_bam_error = "host allocation failed";
    return GUAC_ERR_OOM;
  }
}
const char* guac_bam_last_error(void) { return tl_bam_error.c_str(); }
const char* guac_host_batch_v2_contig_name(const guac_host_batch_v2* b, uint32_t contig) {
  return b && contig < b->contig_names.size() ? b->contig_names[contig].c_str() : "";
}
const char* guac_host_batch_v2_sample_name(const guac_host_batch_v2* b) { return b ? b->sample_name.c_str() : ""; }
double guac_host_batch_v2_decode_stats(const guac_host_batch_v2* b, uint64_t stats[4]) {
  if (!b) return 0;
  if (stats) { stats[0] = b->file_bytes; stats[1] = b->inflated_bytes; stats[2] = b->records_in_file; stats[3] = b->view.n_reads; }
  return b->decode_ms;
}

guac_status guac_reads_pack_device(guac_ctx* ctx, const guac_read_batch* device_batch, const guac_reference* ref, guac_reads** out) {
  if (!ctx || !out) return GUAC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    std::unique_ptr<guac_reads> r(new guac_reads());
    try {
      pack_reads(ctx, device_batch, ref, *r, true);
    } catch (...) {
      cudaStreamSynchronize(ctx->stream);
      throw;
    }
    *out = r.release();
  });
}

guac_status guac_reads_pack_synth(guac_ctx* ctx, guac_synth_device_batch* batch, const guac_reference* ref, guac_reads** out) {
  if (!ctx || !out || !batch) return GUAC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    if (batch->ctx != ctx) fail(GUAC_ERR_INVALID_ARGUMENT, "the batch was generated on another context");
    if (batch->consumed) fail(GUAC_ERR_INVALID_ARGUMENT, "the batch was already packed (its columns moved into that read set)");
    std::unique_ptr<guac_reads> r(new guac_reads());
    try {
      pack_reads(ctx, &batch->view, ref, *r, true, batch);
    } catch (...) {
      cudaStreamSynchronize(ctx->stream);
      throw;
    }
    batch->consumed = true;
    *out = r.release();
  });
}
double guac_reads_expand_kernel_ms(const guac_reads* reads) { return reads ? reads->expand_ms : 0.0; }

// ---- device build of the synthetic read generator (include/guac_synth.h) ---------------------------------------------------
guac_status guac_synth_generate_device(guac_ctx* ctx, const guac_synth_params* p, guac_synth_device_batch** out) {
  if (!ctx || !p || !out) return GUAC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    std::unique_ptr<guac_synth_device_batch> b(new guac_synth_device_batch());
    nvtx_push("guac synth (device)");
    try {
      synth_generate_device(ctx, *p, *b);
    } catch (...) {
      nvtx_pop();
      throw;
    }
    nvtx_pop();
    *out = b.release();
  });
}
const guac_read_batch* guac_synth_device_batch_view(const guac_synth_device_batch* b) { return b ? &b->view : nullptr; }
double guac_synth_device_batch_ms(const guac_synth_device_batch* b) { return b ? b->kernel_ms : 0.0; }
void guac_synth_device_batch_totals(const guac_synth_device_batch* b, uint64_t* n_cigar_ops, uint64_t* n_md_bytes, uint64_t* n_bases) {
  if (n_cigar_ops) *n_cigar_ops = b ? b->n_ops : 0;
  if (n_md_bytes) *n_md_bytes = b ? b->n_md : 0;
  if (n_bases) *n_bases = b ? b->n_bases : 0;
}
void guac_synth_device_batch_free(guac_synth_device_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  delete b;
}
guac_status guac_synth_device_batch_download(guac_ctx* ctx, const guac_synth_device_batch* b, int pinned, guac_synth_host_batch** out) {
  if (!ctx || !b || !out) return GUAC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    if (b->consumed) fail(GUAC_ERR_INVALID_ARGUMENT, "the batch was packed with guac_reads_pack_synth: its columns moved into that read set");
    std::unique_ptr<guac_synth_host_batch> h(new guac_synth_host_batch());
    synth_download(ctx, *b, pinned != 0, *h);
    *out = h.release();
  });
}
const guac_read_batch* guac_synth_host_batch_view(const guac_synth_host_batch* b) { return b ? &b->view : nullptr; }
void guac_synth_host_batch_free(guac_synth_host_batch* b) { delete b; }

void guac_reads_free(guac_reads* reads) {
  if (!reads) return;
  cudaSetDevice(reads->device);  // (ids are never reused: a tile list cached for this read set can never match another one)
  delete reads;
}
uint64_t guac_reads_count(const guac_reads* reads) { return reads ? reads->n : 0; }
uint64_t guac_reads_device_bytes(const guac_reads* reads) { return reads ? reads->device_bytes() : 0; }
uint64_t guac_reads_order_sensitive_loci(const guac_reads* reads) { return reads ? reads->order_sensitive_loci : 0; }
uint64_t guac_reads_h2d_bytes(const guac_reads* reads) { return reads ? reads->h2d_bytes : 0; }
double guac_reads_pack_kernel_ms(const guac_reads* reads) { return reads ? reads->pack_kernel_ms : 0.0; }

guac_status guac_germline_threshold(guac_ctx* ctx, const guac_reads* reads, const guac_locus_range* ranges, size_t n_ranges,
                                    const guac_threshold_params* params, guac_result** out) {
  if (!ctx || !reads || !params || !out || (n_ranges && !ranges)) return GUAC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    std::unique_ptr<guac_result> res(new guac_result());
    res->kind = 0;
    CallParams prm{0, params->threshold_percent, params->emit_ref, params->emit_no_call, params->skip_empty, reads->sample};
    run_pileup(ctx, *reads, ranges, n_ranges, prm, *res);
    *out = res.release();
  });
}

guac_status guac_pileup_counts(guac_ctx* ctx, const guac_reads* reads, const guac_locus_range* ranges, size_t n_ranges,
                               int skip_empty, guac_result** out) {
  if (!ctx || !reads || !out || (n_ranges && !ranges)) return GUAC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    std::unique_ptr<guac_result> res(new guac_result());
    res->kind = 2;
    CallParams prm{1, 0, 0, 0, skip_empty, reads->sample};
    run_pileup(ctx, *reads, ranges, n_ranges, prm, *res);
    *out = res.release();
  });
}

guac_status guac_somatic_standard(guac_ctx* ctx, const guac_reads* tumor, const guac_reads* normal, const guac_locus_range* ranges,
                                  size_t n_ranges, const guac_somatic_params* params, guac_result** out) {
  if (!ctx || !tumor || !normal || !params || !out || (n_ranges && !ranges)) return GUAC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    std::unique_ptr<guac_result> res(new guac_result());
    res->kind = 1;
    run_somatic(ctx, *tumor, *normal, ranges, n_ranges, *params, *res);
    *out = res.release();
  });
}

guac_status guac_somatic_standard_filtered(guac_ctx* ctx, const guac_reads* tumor, const guac_reads* normal, const guac_locus_range* ranges,
                                           size_t n_ranges, const guac_somatic_params* params, const guac_somatic_filter_params* filters,
                                           guac_result** out) {
  if (!ctx || !tumor || !normal || !params || !filters || !out || (n_ranges && !ranges)) return GUAC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    std::unique_ptr<guac_result> res(new guac_result());
    res->kind = 1;
    run_somatic(ctx, *tumor, *normal, ranges, n_ranges, *params, *res, filters);
    *out = res.release();
  });
}

// pileupFlatMap(reads, ranges, true, pileupToAlleleCounts): the exact per-element walk over every requested locus
static void run_allele_counts(guac_ctx* ctx, const guac_reads& reads, const guac_locus_range* ranges, size_t n_ranges, guac_result& res) {
  cudaStream_t st = ctx->stream;
  std::vector<unsigned long long> prefix(n_ranges + 1, 0);
  for (size_t i = 0; i < n_ranges; ++i) {
    const guac_locus_range& r = ranges[i];
    if (r.contig < 0 || (uint32_t)r.contig >= reads.n_contigs) fail(GUAC_ERR_INVALID_ARGUMENT, "locus range %zu: contig out of range", i);
    if (r.start < 0 || r.end < r.start) fail(GUAC_ERR_INVALID_ARGUMENT, "locus range %zu: bad bounds", i);
    prefix[i + 1] = prefix[i] + (unsigned long long)(r.end - r.start);
  }
  check_ranges_disjoint(ranges, n_ranges);
  const uint64_t requested = prefix[n_ranges];
  res.stats.reads_total = reads.n;
  res.stats.loci_requested = requested;
  res.stats.order_sensitive_loci = reads.order_sensitive_loci;
  if (requested == 0 || reads.n == 0) return;
  DevBuf<guac_locus_range> d_ranges;
  DevBuf<unsigned long long> d_prefix;
  h2d(ctx, d_ranges, ranges, n_ranges);
  h2d(ctx, d_prefix, prefix.data(), prefix.size());
  res.stats.h2d_bytes = d_ranges.bytes() + d_prefix.bytes();
  uint64_t cap_rec = std::max<uint64_t>(4096, 2 * requested), cap_pool = kPoolDynOff + std::max<uint64_t>(65536, 8 * requested);
  const CallParams prm{2, 0, 0, 0, 1, reads.sample};
  for (int attempt = 0; attempt < 6; ++attempt) {
    if (cap_rec >= 0xFFFFFFF0ull || cap_pool >= 0xFFFFFFF0ull) fail(GUAC_ERR_UNSUPPORTED, "too many output records for one call: split the loci ranges");
    ctx->out_rec.ensure(cap_rec * sizeof(guac_allele_count));
    if (ctx->out_pool.ensure(cap_pool)) ctx->pool_head_ready = false;
    CUDA_OK(cudaMemsetAsync(ctx->d_counters, 0, 16 * sizeof(unsigned long long), st));
    DevOut out;
    out.trec = (guac_threshold_record*)ctx->out_rec.p;
    out.crec = (guac_locus_counts*)ctx->out_rec.p;
    out.cap_rec = (uint32_t)cap_rec;
    out.pool = ctx->out_pool.p;
    out.cap_pool = (uint32_t)cap_pool;
    out.slow = nullptr;
    out.cap_slow = 0;
    out.slow_ctr = 8;
    out.compact_ctr = 12;
    out.compact = nullptr;
    out.cap_compact = 0;
    out.counters = ctx->d_counters;
    out.err = ctx->d_err;
    CUDA_OK(cudaEventRecord(ctx->ev[0], st));
    const uint64_t warps_needed = requested;
    const int grid = (int)std::min<uint64_t>((warps_needed + kExactWarps - 1) / kExactWarps, (uint64_t)ctx->sm_count * 16);
    k_allele_counts<<<grid, kExactWarps * 32, 0, st>>>(reads.view(), d_ranges.p, d_prefix.p, (uint32_t)n_ranges, prm, out);
    CUDA_OK(cudaEventRecord(ctx->ev[1], st));
    CUDA_OK(cudaGetLastError());
    unsigned long long* c = ctx->h_counters;
    CUDA_OK(cudaMemcpyAsync(c, ctx->d_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    check_device_error(ctx, "allele counts");
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    res.stats.exact_kernel_ms += ms;
    res.stats.kernel_launches += 1;
    if (c[0] > cap_rec || kPoolDynOff + c[1] > cap_pool) {
      cap_rec = std::max<uint64_t>(cap_rec, c[0] + c[0] / 8 + 16);
      cap_pool = std::max<uint64_t>(cap_pool, kPoolDynOff + c[1] + c[1] / 8 + 16);
      continue;
    }
    const uint64_t n_rec = c[0];
    const size_t pool_bytes = (size_t)(kPoolDynOff + c[1]), rec_bytes = (size_t)(n_rec * sizeof(guac_allele_count));
    const size_t rec_at = (pool_bytes + 63) & ~(size_t)63;
    res.pool = ctx->pinned;
    res.block = ctx->pinned->take(rec_at + rec_bytes + 64, &res.block_bytes);
    if (!res.block) fail(GUAC_ERR_OOM, "pinned host allocation of %zu bytes failed", rec_at + rec_bytes + 64);
    unsigned char* hs = (unsigned char*)res.block;
    unsigned char* hrec = hs + rec_at;
    CUDA_OK(cudaMemcpyAsync(hs, ctx->out_pool.p, pool_bytes, cudaMemcpyDeviceToHost, st));
    if (n_rec) CUDA_OK(cudaMemcpyAsync(hrec, ctx->out_rec.p, rec_bytes, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    res.stats.d2h_bytes = pool_bytes + rec_bytes + 64;
    res.records = hrec;
    res.n_records = (size_t)n_rec;
    res.bytes = hs;
    res.n_bytes = pool_bytes;
    if (ctx->sort_records) {
      const uint8_t* pool = hs;
      sort_records_canonical((guac_allele_count*)hrec, (size_t)n_rec, [pool](const guac_allele_count& a, const guac_allele_count& b) {
        int c = memcmp(pool + a.ref_off, pool + b.ref_off, std::min(a.ref_len, b.ref_len));
        if (c != 0) return c < 0;
        if (a.ref_len != b.ref_len) return a.ref_len < b.ref_len;
        c = memcmp(pool + a.alt_off, pool + b.alt_off, std::min(a.alt_len, b.alt_len));
        if (c != 0) return c < 0;
        return a.alt_len < b.alt_len;
      });
    }
    res.stats.records = n_rec;
    res.stats.exact_loci = requested;
    res.stats.kernel_ms = res.stats.exact_kernel_ms;
    return;
  }
  fail(GUAC_ERR_CUDA, "output buffers did not converge");
}

guac_status guac_allele_counts(guac_ctx* ctx, const guac_reads* reads, const guac_locus_range* ranges, size_t n_ranges, guac_result** out) {
  if (!ctx || !reads || !out || (n_ranges && !ranges)) return GUAC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    std::unique_ptr<guac_result> res(new guac_result());
    res->kind = 4;
    run_allele_counts(ctx, *reads, ranges, n_ranges, *res);
    *out = res.release();
  });
}

guac_status guac_germline_standard(guac_ctx* ctx, const guac_reads* reads, const guac_locus_range* ranges, size_t n_ranges,
                                   const guac_standard_params* params, guac_result** out) {
  if (!ctx || !reads || !params || !out || (n_ranges && !ranges)) return GUAC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    std::unique_ptr<guac_result> res(new guac_result());
    res->kind = 3;
    run_standard(ctx, *reads, ranges, n_ranges, *params, *res);
    *out = res.release();
  });
}

size_t guac_result_n(const guac_result* r) {
  if (!r) return 0;
  return r->kind == 2 ? r->counts.size() : r->n_records;
}
const guac_threshold_record* guac_result_threshold_records(const guac_result* r) {
  if (!r || r->kind != 0) return nullptr;
  guac_result* w = const_cast<guac_result*>(r);  // (the expanded view is a cache; a result belongs to one thread)
  try {
    expand_threshold_records(*w);
  } catch (const std::bad_alloc&) {
    return nullptr;
  }
  return w->expanded.data();
}
size_t guac_result_compact_records(const guac_result* r, const guac_compact_record** compact, const guac_threshold_record** general,
                                   size_t* n_general, int32_t* sample) {
  if (!r || r->kind != 0) return 0;
  if (compact) *compact = (const guac_compact_record*)r->compact;
  if (general) *general = r->general;
  if (n_general) *n_general = r->n_general;
  if (sample) *sample = r->sample;
  return r->n_compact;
}
const guac_somatic_record* guac_result_somatic_records(const guac_result* r) { return (r && r->kind == 1) ? (const guac_somatic_record*)r->records : nullptr; }
const guac_locus_counts* guac_result_counts(const guac_result* r) { return (r && r->kind == 2) ? r->counts.data() : nullptr; }
const guac_allele_count* guac_result_allele_counts(const guac_result* r) { return (r && r->kind == 4) ? (const guac_allele_count*)r->records : nullptr; }
const guac_called_allele* guac_result_called_alleles(const guac_result* r) { return (r && r->kind == 3) ? (const guac_called_allele*)r->records : nullptr; }
const uint8_t* guac_result_bytes(const guac_result* r, size_t* n_bytes) {
  if (n_bytes) *n_bytes = r ? r->n_bytes : 0;
  return r ? r->bytes : nullptr;
}
const guac_stats* guac_result_stats(const guac_result* r) { return r ? &r->stats : nullptr; }
void guac_result_free(guac_result* r) { delete r; }

// SomaticGenotypeFilter (filters/SomaticGenotypeFilter.scala): cheap predicates on the emitted records
size_t guac_somatic_genotype_filter(const guac_somatic_record* g, size_t n, const guac_somatic_filter_params* p, uint8_t* keep) {
  if (!g || !p || !keep) return 0;
  size_t kept = 0;
  for (size_t i = 0; i < n; ++i) {
    const bool ok = somatic_filter_keep(g[i], *p);
    keep[i] = ok ? 1 : 0;
    kept += ok ? 1 : 0;
  }
  return kept;
}

// partitionLociUniformly (DistributedUtil.scala:83-108): host-side LociPartitioning
guac_status guac_partition_loci_uniformly(int64_t tasks, const guac_locus_range* loci, size_t n_loci, guac_locus_range* out,
                                          size_t max_out, size_t* n_out) {
  if (tasks < 1 || (!loci && n_loci) || !n_out) return GUAC_ERR_INVALID_ARGUMENT;
  int64_t count = 0;
  for (size_t i = 0; i < n_loci; ++i) count += loci[i].end - loci[i].start;
  const double per_task = std::max(1.0, (double)count / (double)tasks);
  int64_t assigned = 0, task = 0;
  auto remaining = [&] { return (int64_t)std::floor((double)(task + 1) * per_task - (double)assigned + 0.5); };  // math.round
  size_t n = 0;
  guac_locus_range last{};
  bool have_last = false;
  auto flush = [&] {
    if (have_last) {
      if (n < max_out && out) out[n] = last;
      ++n;
    }
  };
  for (size_t i = 0; i < n_loci; ++i) {
    int64_t start = loci[i].start;
    const int64_t end = loci[i].end;
    while (start < end) {
      const int64_t length = std::min(remaining(), end - start);
      if (length <= 0) return GUAC_ERR_INVALID_ARGUMENT;
      if (have_last && last.contig == loci[i].contig && last.task == (int32_t)task && last.end == start) {
        last.end = start + length;
      } else {
        flush();
        last = guac_locus_range{loci[i].contig, (int32_t)task, start, start + length};
        have_last = true;
      }
      start += length;
      assigned += length;
      if (remaining() == 0) ++task;
    }
  }
  flush();
  *n_out = n;
  return n > max_out && out ? GUAC_ERR_INVALID_ARGUMENT : GUAC_OK;
}

// ---- depth histogram + the NCCL exchange (guac_comm.cuh) --------------------------------------------------------------------
guac_status guac_depth_histogram(guac_ctx* ctx, const guac_reads* reads, const guac_locus_range* ranges, size_t n_ranges, uint64_t* hist) {
  if (!ctx || !reads || (n_ranges && !ranges)) return GUAC_ERR_INVALID_ARGUMENT;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    if (!reads->gs_hdr.n && reads->n) fail(GUAC_ERR_UNSUPPORTED, "the depth histogram reads the difference streams (GUAC_OPT_DIFFERENCE_LISTS = 1)");
    cudaStream_t st = ctx->stream;
    if (!ctx->d_hist) CUDA_OK(cudaMalloc((void**)&ctx->d_hist, kDepthBins * sizeof(unsigned long long)));
    CUDA_OK(cudaMemsetAsync(ctx->d_hist, 0, kDepthBins * sizeof(unsigned long long), st));
    uint64_t tile_loci = 0;
    const uint64_t requested = prepare_tiles(ctx, *reads, ranges, n_ranges, &tile_loci);
    if (ctx->n_tiles) {
      const size_t smem = (size_t)kHistWarps * kDepthBins * 32 * sizeof(uint32_t);
      if (!ctx->hist_attr_done) {
        CUDA_OK(cudaFuncSetAttribute(k_depth_histogram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->hist_attr_done = true;
      }
      const int grid = (int)std::min<uint64_t>((ctx->n_tiles + kHistWarps - 1) / kHistWarps, (uint64_t)ctx->sm_count);
      k_depth_histogram<<<grid, kHistWarps * 32, smem, st>>>(reads->view(), (const TileDesc*)ctx->tiles.p, (uint32_t)ctx->n_tiles, ctx->d_hist);
      CUDA_OK(cudaGetLastError());
    }
    if (hist) {
      CUDA_OK(cudaMemcpyAsync(hist, ctx->d_hist, kDepthBins * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaStreamSynchronize(st));
      hist[0] += requested - tile_loci;  // requested loci past the track hold no reads
    } else if (requested > tile_loci) {
      const unsigned long long extra = requested - tile_loci;
      unsigned long long h0 = 0;
      CUDA_OK(cudaMemcpyAsync(&h0, ctx->d_hist, 8, cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaStreamSynchronize(st));
      h0 += extra;
      CUDA_OK(cudaMemcpyAsync(ctx->d_hist, &h0, 8, cudaMemcpyHostToDevice, st));
      CUDA_OK(cudaStreamSynchronize(st));
    }
  });
}

guac_status guac_comm_unique_id(uint8_t* id) {
  if (!id) return GUAC_ERR_INVALID_ARGUMENT;
  static_assert(sizeof(ncclUniqueId) <= GUAC_COMM_ID_BYTES, "ncclUniqueId fits the id buffer");
  ncclUniqueId u;
  if (ncclGetUniqueId(&u) != ncclSuccess) return GUAC_ERR_CUDA;
  memset(id, 0, GUAC_COMM_ID_BYTES);
  memcpy(id, &u, sizeof u);
  return GUAC_OK;
}

guac_status guac_comm_create(guac_ctx* ctx, const uint8_t* id, int rank, int world, guac_comm** out) {
  if (!ctx || !id || !out || world < 1 || rank < 0 || rank >= world) return GUAC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    std::unique_ptr<guac_comm> c(new guac_comm());
    c->ctx = ctx;
    c->device = ctx->device;
    c->rank = rank;
    c->world = world;
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    NCCL_OK(ncclCommInitRank(&c->comm, world, u, rank));
    c->d_sizes.alloc((size_t)6 * (world + 1));
    CUDA_OK(cudaMallocHost((void**)&c->h_sizes, (size_t)6 * (world + 1) * sizeof(unsigned long long)));
    *out = c.release();
  });
}

void guac_comm_destroy(guac_comm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->comm) ncclCommDestroy(c->comm);
  if (c->h_sizes) cudaFreeHost(c->h_sizes);
  delete c;
}

guac_status guac_result_gather(guac_comm* comm, const guac_result* local, int root, guac_result** out) {
  if (!comm || !local || !out || root < 0 || root >= comm->world) return GUAC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  guac_ctx* ctx = comm->ctx;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    if (local->kind != 0) fail(GUAC_ERR_UNSUPPORTED, "guac_result_gather takes germline-threshold results");
    if (local->n_compact && (local->owner != ctx || local->generation != ctx->generation))
      fail(GUAC_ERR_INVALID_ARGUMENT, "gather the result before the context runs another call (its records are sent from device memory)");
    nvtx_push("guac gather (NCCL)");
    cudaStream_t st = ctx->stream;
    const int W = comm->world, me = comm->rank;
    const size_t pool_dyn = local->n_bytes > kPoolDynOff ? local->n_bytes - kPoolDynOff : 0;
    // 1. counts
    unsigned long long* mine = comm->h_sizes + (size_t)6 * W;
    mine[0] = local->n_compact; mine[1] = local->n_general; mine[2] = pool_dyn;
    mine[3] = local->stats.loci_visited; mine[4] = local->stats.tie_loci; mine[5] = local->stats.exact_loci;
    CUDA_OK(cudaMemcpyAsync(comm->d_sizes.p + (size_t)6 * W, mine, 6 * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    NCCL_OK(ncclAllGather(comm->d_sizes.p + (size_t)6 * W, comm->d_sizes.p, 6, ncclUint64, comm->comm, st));
    CUDA_OK(cudaMemcpyAsync(comm->h_sizes, comm->d_sizes.p, (size_t)6 * W * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    // the local general records and allele bytes were written to pinned host memory by the exact kernel: stage them in HBM
    const size_t gen_bytes = local->n_general * sizeof(guac_threshold_record);
    comm->d_stage.ensure(gen_bytes + pool_dyn + 64);
    if (gen_bytes) CUDA_OK(cudaMemcpyAsync(comm->d_stage.p, local->general, gen_bytes, cudaMemcpyHostToDevice, st));
    if (pool_dyn) CUDA_OK(cudaMemcpyAsync(comm->d_stage.p + gen_bytes, local->bytes + kPoolDynOff, pool_dyn, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaStreamSynchronize(st));
    const unsigned long long* S = comm->h_sizes;
    std::unique_ptr<guac_result> res(new guac_result());
    res->kind = 0;
    res->sample = local->sample;
    res->want_sorted = local->want_sorted;
    res->stats = local->stats;
    unsigned long long tot_c = 0, tot_g = 0, tot_p = 0, visited = 0, ties = 0, exact = 0;
    for (int r = 0; r < W; ++r) { tot_c += S[6 * r]; tot_g += S[6 * r + 1]; tot_p += S[6 * r + 2]; visited += S[6 * r + 3]; ties += S[6 * r + 4]; exact += S[6 * r + 5]; }
    // 2. records rank -> root, from device memory
    unsigned char* recv = nullptr;
    size_t at_c = 0, at_g = 0, at_p = 0;
    if (me == root) {
      at_g = (size_t)tot_c * 8;
      at_p = at_g + (size_t)tot_g * sizeof(guac_threshold_record);
      comm->d_recv.ensure(at_p + (size_t)tot_p + 64);
      recv = comm->d_recv.p;
    }
    NCCL_OK(ncclGroupStart());
    if (me == root) {
      size_t oc = 0, og = 0, op = 0;
      for (int r = 0; r < W; ++r) {
        const size_t bc = (size_t)S[6 * r] * 8, bg = (size_t)S[6 * r + 1] * sizeof(guac_threshold_record), bp = (size_t)S[6 * r + 2];
        if (r == me) {
          if (bc) CUDA_OK(cudaMemcpyAsync(recv + at_c + oc, local->d_compact, bc, cudaMemcpyDeviceToDevice, st));
          if (bg) CUDA_OK(cudaMemcpyAsync(recv + at_g + og, comm->d_stage.p, bg, cudaMemcpyDeviceToDevice, st));
          if (bp) CUDA_OK(cudaMemcpyAsync(recv + at_p + op, comm->d_stage.p + gen_bytes, bp, cudaMemcpyDeviceToDevice, st));
        } else {
          if (bc) NCCL_OK(ncclRecv(recv + at_c + oc, bc, ncclUint8, r, comm->comm, st));
          if (bg) NCCL_OK(ncclRecv(recv + at_g + og, bg, ncclUint8, r, comm->comm, st));
          if (bp) NCCL_OK(ncclRecv(recv + at_p + op, bp, ncclUint8, r, comm->comm, st));
        }
        oc += bc; og += bg; op += bp;
      }
    } else {
      if (local->n_compact) NCCL_OK(ncclSend(local->d_compact, local->n_compact * 8, ncclUint8, root, comm->comm, st));
      if (gen_bytes) NCCL_OK(ncclSend(comm->d_stage.p, gen_bytes, ncclUint8, root, comm->comm, st));
      if (pool_dyn) NCCL_OK(ncclSend(comm->d_stage.p + gen_bytes, pool_dyn, ncclUint8, root, comm->comm, st));
    }
    NCCL_OK(ncclGroupEnd());
    if (me == root) {
      // 3. one copy to the pinned block of the merged result: [allele pool][general records][compact records]
      const size_t pool_bytes = kPoolDynOff + (size_t)tot_p;
      const size_t full_at = (pool_bytes + 63) & ~(size_t)63;
      const size_t compact_at = (full_at + (size_t)tot_g * sizeof(guac_threshold_record) + 63) & ~(size_t)63;
      res->pool = ctx->pinned;
      res->block = ctx->pinned->take(compact_at + (size_t)tot_c * 8 + 64, &res->block_bytes);
      if (!res->block) fail(GUAC_ERR_OOM, "pinned host allocation failed");
      unsigned char* hs = (unsigned char*)res->block;
      memset(hs, 0, kPoolDynOff);
      memcpy(hs, "<ALT>", 5);
      for (int v = 0; v < 256; ++v) hs[kPoolByteOff + v] = (uint8_t)v;
      if (tot_p) CUDA_OK(cudaMemcpyAsync(hs + kPoolDynOff, recv + at_p, (size_t)tot_p, cudaMemcpyDeviceToHost, st));
      if (tot_g) CUDA_OK(cudaMemcpyAsync(hs + full_at, recv + at_g, (size_t)tot_g * sizeof(guac_threshold_record), cudaMemcpyDeviceToHost, st));
      if (tot_c) CUDA_OK(cudaMemcpyAsync(hs + compact_at, recv + at_c, (size_t)tot_c * 8, cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaStreamSynchronize(st));
      // general records keep offsets into their own rank's pool: re-base them onto the merged pool
      guac_threshold_record* g = (guac_threshold_record*)(hs + full_at);
      size_t og = 0, op = 0;
      for (int r = 0; r < W; ++r) {
        for (size_t i = 0; i < (size_t)S[6 * r + 1]; ++i) {
          guac_threshold_record& t = g[og + i];
          if (t.ref_off >= kPoolDynOff) t.ref_off += (uint32_t)op;
          if (t.alt_off >= kPoolDynOff) t.alt_off += (uint32_t)op;
        }
        og += (size_t)S[6 * r + 1];
        op += (size_t)S[6 * r + 2];
      }
      if (kPoolDynOff + tot_p >= 0xFFFFFFF0ull) fail(GUAC_ERR_UNSUPPORTED, "merged allele pool beyond 4 GiB");
      res->general = g;
      res->n_general = (size_t)tot_g;
      res->compact = (const unsigned long long*)(hs + compact_at);
      res->n_compact = (size_t)tot_c;
      res->n_records = res->n_general + res->n_compact;
      res->bytes = hs;
      res->n_bytes = pool_bytes;
      // shards are contiguous loci ranges in rank order: the concatenation of sorted shards is sorted iff every shard was
      bool sorted = local->compact_sorted;
      for (size_t i = 1; sorted && i < res->n_compact; ++i) sorted = res->compact[i - 1] <= res->compact[i];
      res->compact_sorted = sorted;
      res->stats.records = res->n_records;
      res->stats.loci_visited = visited;
      res->stats.tie_loci = ties;
      res->stats.exact_loci = exact;
      res->stats.d2h_bytes = (size_t)tot_p + (size_t)tot_g * sizeof(guac_threshold_record) + (size_t)tot_c * 8;
    } else {
      CUDA_OK(cudaStreamSynchronize(st));
      res->stats.records = 0;
    }
    nvtx_pop();
    *out = res.release();
  });
}

guac_status guac_comm_reduce_depth_histogram(guac_comm* comm, int root, uint64_t* hist) {
  if (!comm || root < 0 || root >= comm->world) return GUAC_ERR_INVALID_ARGUMENT;
  guac_ctx* ctx = comm->ctx;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    if (!ctx->d_hist) fail(GUAC_ERR_INVALID_ARGUMENT, "run guac_depth_histogram on this context first");
    cudaStream_t st = ctx->stream;
    NCCL_OK(ncclReduce(ctx->d_hist, ctx->d_hist, kDepthBins, ncclUint64, ncclSum, root, comm->comm, st));
    if (comm->rank == root && hist) CUDA_OK(cudaMemcpyAsync(hist, ctx->d_hist, kDepthBins * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
  });
}

// partitionLociByApproximateDepth (DistributedUtil.scala:162-251): micro partitions by the uniform rule, regions per micro
// partition counted on the device from the packed read sets, then the greedy assignment on the host in the reference's
// own Double arithmetic.
guac_status guac_partition_loci_by_approximate_depth(guac_ctx* ctx, int64_t tasks, const guac_locus_range* loci, size_t n_loci,
                                                     int64_t accuracy, const guac_reads* const* read_sets, size_t n_read_sets,
                                                     guac_locus_range* out, size_t max_out, size_t* n_out) {
  if (!ctx || tasks < 1 || accuracy < 1 || (!loci && n_loci) || !read_sets || n_read_sets < 1 || !n_out) return GUAC_ERR_INVALID_ARGUMENT;
  return guarded(ctx, [&] {
    CUDA_OK(cudaSetDevice(ctx->device));
    int64_t count = 0;
    for (size_t i = 0; i < n_loci; ++i) count += loci[i].end - loci[i].start;
    if (count <= 0) fail(GUAC_ERR_INVALID_ARGUMENT, "assumption failed: lociUsed.count > 0");
    // step (1): micro partitions
    const int64_t n_micro = (accuracy * tasks < count) ? accuracy * tasks : count;
    if (n_micro > (int64_t)0x7FFFFFF0) fail(GUAC_ERR_UNSUPPORTED, "too many micro partitions");
    std::vector<guac_locus_range> micro((size_t)n_micro + n_loci + 8);
    size_t n_ranges = 0;
    if (guac_partition_loci_uniformly(n_micro, loci, n_loci, micro.data(), micro.size(), &n_ranges) != GUAC_OK)
      fail(GUAC_ERR_INVALID_ARGUMENT, "partitionLociUniformly failed");
    micro.resize(n_ranges);
    // step (2): region counts, on the device.  The kernel wants each contig's ranges sorted by start.
    std::vector<guac_locus_range> by_contig(micro);
    std::stable_sort(by_contig.begin(), by_contig.end(), [](const guac_locus_range& a, const guac_locus_range& b) {
      return a.contig != b.contig ? a.contig < b.contig : a.start < b.start;
    });
    cudaStream_t st = ctx->stream;
    DevBuf<guac_locus_range> d_ranges;
    DevBuf<unsigned long long> d_counts;
    h2d(ctx, d_ranges, by_contig.data(), by_contig.size());
    d_counts.alloc((size_t)n_micro);
    CUDA_OK(cudaMemsetAsync(d_counts.p, 0, (size_t)n_micro * sizeof(unsigned long long), st));
    for (size_t s = 0; s < n_read_sets; ++s) {
      const guac_reads* rd = read_sets[s];
      if (!rd) fail(GUAC_ERR_INVALID_ARGUMENT, "null read set");
      const DevReads R = rd->view();
      size_t k0 = 0;
      while (k0 < by_contig.size()) {
        size_t k1 = k0;
        while (k1 < by_contig.size() && by_contig[k1].contig == by_contig[k0].contig) ++k1;
        const int32_t c = by_contig[k0].contig;
        if (c >= 0 && (uint32_t)c < rd->n_contigs) {
          const ContigInfo& ci = rd->contigs[c];
          if (ci.read_end > ci.read_begin)
            k_micro_counts<<<grid_for(ci.read_end - ci.read_begin, 256, ctx->sm_count), 256, 0, st>>>(R, ci.read_begin, ci.read_end, d_ranges.p + k0,
                                                                                                    (uint32_t)(k1 - k0), d_counts.p);
        }
        k0 = k1;
      }
    }
    CUDA_OK(cudaGetLastError());
    std::vector<unsigned long long> counts((size_t)n_micro);
    CUDA_OK(cudaMemcpyAsync(counts.data(), d_counts.p, counts.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    // step (3): the greedy assignment (:204-247)
    auto round_double = [](double x) -> long long {  // java.lang.Math.round(double)
      if (std::isnan(x)) return 0;
      const double f = std::floor(x + 0.5);
      if (f >= 9.2233720368547758e18) return INT64_MAX;
      if (f <= -9.2233720368547758e18) return INT64_MIN;
      return (long long)f;
    };
    auto round_long = [](long long v) -> long long {  // scala.math.round(Long) in Scala 2.10 = Math.round(v.toFloat): Int
      const float f = std::floor((float)v + 0.5f);
      if (f >= 2147483648.0f) return INT32_MAX;
      if (f <= -2147483648.0f) return INT32_MIN;
      return (long long)(int)f;
    };
    long long total = 0;
    for (unsigned long long c : counts) total += (long long)c;
    const double per_task = std::max(1.0, (double)total / (double)tasks);
    double assigned = 0.0;
    long long task = 0;
    size_t n = 0;
    guac_locus_range last{};
    bool have_last = false;
    auto put = [&](int32_t contig, int64_t start, int64_t end) {  // LociMap.Builder.put: adjacent ranges of one task coalesce
      if (end <= start) return;
      if (have_last && last.contig == contig && last.task == (int32_t)task && last.end == start) {
        last.end = end;
        return;
      }
      if (have_last) {
        if (n < max_out && out) out[n] = last;
        ++n;
      }
      last = guac_locus_range{contig, (int32_t)task, start, end};
      have_last = true;
    };
    auto remaining = [&] { return round_double((double)(task + 1) * per_task - assigned); };
    size_t k = 0;
    for (long long mt = 0; mt < n_micro; ++mt) {
      size_t k_end = k;
      long long set_count = 0;
      for (; k_end < micro.size() && micro[k_end].task == (int32_t)mt; ++k_end) set_count += micro[k_end].end - micro[k_end].start;
      long long in_set = (long long)counts[(size_t)mt];
      int64_t cursor = k < k_end ? micro[k].start : 0;  // first locus of the set not yet assigned
      while (set_count > 0) {
        long long take = set_count;
        if (in_set != 0) {
          if (remaining() == 0) task += 1;
          if (!(remaining() > 0) || !(task < tasks)) fail(GUAC_ERR_INVALID_ARGUMENT, "partitionLociByApproximateDepth: assertion failed");
          const double fraction = std::min(1.0, (double)remaining() / (double)in_set);
          take = std::max<long long>(1, (long long)(fraction * (double)set_count));
          const long long regions = round_long((long long)(fraction * (double)in_set));
          assigned += (double)regions;
          in_set -= regions;
        }
        set_count -= take;
        while (take > 0) {  // set.take(lociToTake): the first `take` loci of the set in range order
          const long long avail = micro[k].end - cursor;
          const long long now = std::min(avail, take);
          put(micro[k].contig, cursor, cursor + now);
          cursor += now;
          take -= now;
          if (cursor == micro[k].end && ++k < k_end) cursor = micro[k].start;
        }
      }
      k = k_end;
    }
    if (have_last) {
      if (n < max_out && out) out[n] = last;
      ++n;
    }
    *n_out = n;
    if (n > max_out && out) fail(GUAC_ERR_INVALID_ARGUMENT, "output array too small: %zu ranges", n);
  });
}

}  // extern "C"
