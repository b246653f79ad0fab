// guac_order.cuh — canonical order of the likelihood callers' records on the device.
//
// The callers' kernels hand out record slots with one atomic, so a result leaves the kernels in no particular order; the
// reference's records are compared as a set, and the C ABI promises (contig, start, ref, alt) order when GUAC_OPT_SORT_RECORDS
// is on.  All record types (guac_somatic_record, guac_called_allele, guac_allele_count) start with the same head, so one
// counting sort serves them: bucket = the granule (1024 loci) of the record's locus — a handful of records at most —, an
// exclusive scan of the bucket counts, a scatter of record indices, an insertion sort inside the few buckets that hold more
// than one record (start, then reference allele, then alternate allele bytes: the host comparator's order), and a gather of
// whole records into the buffer the device -> host copy reads.  Replaces a host-side radix sort + permutation of the pinned
// records (2.4 ms per chr20 somatic call) by ~40 us of small kernels.
#pragma once

#include "guac_synth_device.cuh"

namespace guac {

struct RecHead {  // the leading fields every record type shares (include/guac.h)
  int64_t start;
  int32_t contig;
  int32_t sample;
  uint32_t ref_off, alt_off;
  uint16_t ref_len, alt_len;
};

__device__ __forceinline__ const RecHead& rec_head(const unsigned char* recs, uint32_t rec_bytes, uint32_t i) {
  return *reinterpret_cast<const RecHead*>(recs + (size_t)i * rec_bytes);
}

__device__ __forceinline__ uint32_t rec_bucket(const RecHead& h, const ContigInfo* __restrict__ contigs, uint32_t n_contigs, uint32_t n_buckets) {
  const uint32_t c = min((uint32_t)max(h.contig, 0), n_contigs - 1u);
  const unsigned long long b = (unsigned long long)contigs[c].gran_off + (unsigned long long)(max(h.start, (int64_t)0) >> kGranuleShift);
  return (uint32_t)min(b, (unsigned long long)(n_buckets - 1u));
}

__global__ void __launch_bounds__(256) k_ord_count(const unsigned char* __restrict__ recs, uint32_t rec_bytes, uint32_t n, const ContigInfo* __restrict__ contigs,
                                                   uint32_t n_contigs, uint32_t n_buckets, uint32_t* __restrict__ cnt) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    atomicAdd(&cnt[rec_bucket(rec_head(recs, rec_bytes, i), contigs, n_contigs, n_buckets)], 1u);
}

// cnt[] is counted down again: a bucket's records take its slots from the back
__global__ void __launch_bounds__(256) k_ord_scatter(const unsigned char* __restrict__ recs, uint32_t rec_bytes, uint32_t n, const ContigInfo* __restrict__ contigs,
                                                     uint32_t n_contigs, uint32_t n_buckets, uint32_t* __restrict__ cnt, const uint32_t* __restrict__ base,
                                                     uint32_t* __restrict__ order) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t b = rec_bucket(rec_head(recs, rec_bytes, i), contigs, n_contigs, n_buckets);
    order[base[b] + atomicSub(&cnt[b], 1u) - 1u] = i;
  }
}

__device__ inline int bytes_compare(const uint8_t* a, uint32_t la, const uint8_t* b, uint32_t lb) {  // memcmp, then length
  const uint32_t n = min(la, lb);
  for (uint32_t i = 0; i < n; ++i)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return la == lb ? 0 : (la < lb ? -1 : 1);
}

// thread per bucket with more than one record: insertion sort of its slice of `order`
__global__ void __launch_bounds__(256) k_ord_buckets(const unsigned char* __restrict__ recs, uint32_t rec_bytes, const uint8_t* __restrict__ pool,
                                                     uint32_t n_buckets, const uint32_t* __restrict__ base, uint32_t* __restrict__ order) {
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < n_buckets; b += gridDim.x * blockDim.x) {
    const uint32_t lo = base[b], hi = base[b + 1];
    if (hi - lo < 2u) continue;
    auto less = [&](uint32_t x, uint32_t y) {
      const RecHead& p = rec_head(recs, rec_bytes, x);
      const RecHead& q = rec_head(recs, rec_bytes, y);
      if (p.contig != q.contig) return p.contig < q.contig;
      if (p.start != q.start) return p.start < q.start;
      int c = bytes_compare(pool + p.ref_off, p.ref_len, pool + q.ref_off, q.ref_len);
      if (c == 0) c = bytes_compare(pool + p.alt_off, p.alt_len, pool + q.alt_off, q.alt_len);
      return c != 0 ? c < 0 : x < y;  // (equal records: by slot, whatever that was)
    };
    for (uint32_t i = lo + 1; i < hi; ++i) {
      const uint32_t v = order[i];
      uint32_t j = i;
      while (j > lo && less(v, order[j - 1])) {
        order[j] = order[j - 1];
        --j;
      }
      order[j] = v;
    }
  }
}

// out[i] = recs[order[i]], eight bytes per thread and step (record sizes are multiples of 8)
__global__ void __launch_bounds__(256) k_ord_gather(const unsigned long long* __restrict__ recs, uint32_t rec_words, uint32_t n, const uint32_t* __restrict__ order,
                                                    unsigned long long* __restrict__ out) {
  const unsigned long long total = (unsigned long long)n * rec_words;
  for (unsigned long long w = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; w < total; w += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t i = (uint32_t)(w / rec_words), k = (uint32_t)(w % rec_words);
    out[w] = recs[(size_t)order[i] * rec_words + k];
  }
}

}  // namespace guac

namespace {

// Orders n records (device memory, `rec_bytes` each, allele bytes in `pool` on the device) canonically; returns the device
// buffer that holds them in order (ctx scratch, valid until the next call).  Asynchronous on the context's stream.
const unsigned char* device_order_records(guac_ctx* ctx, const guac_reads& reads, const unsigned char* recs, uint32_t rec_bytes, uint64_t n,
                                          const uint8_t* pool) {
  if (n < 2 || n >= 0xFFFFFFF0ull || rec_bytes % 8 != 0 || reads.n_contigs == 0) return recs;
  cudaStream_t st = ctx->stream;
  const uint32_t n_buckets = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(reads.total_grans, 1), 0xFFFFFFF0ull);
  const uint64_t n_chunks = ((uint64_t)n_buckets + kScanChunk - 1) / kScanChunk;
  ctx->ord_u32.ensure(2 * ((size_t)n_buckets + 4) + (size_t)n + 4);
  ctx->scan_totals.ensure(n_chunks + 2);
  ctx->ord_out.ensure((size_t)n * rec_bytes + 16);
  uint32_t* cnt = ctx->ord_u32.p;
  uint32_t* base = cnt + ((size_t)n_buckets + 4);
  uint32_t* order = base + ((size_t)n_buckets + 4);
  CUDA_OK(cudaMemsetAsync(cnt, 0, (size_t)n_buckets * sizeof(uint32_t), st));
  const int grid = grid_for(n, 256, ctx->sm_count);
  const ContigInfo* contigs = reads.d_contigs.p;
  k_ord_count<<<grid, 256, 0, st>>>(recs, rec_bytes, (uint32_t)n, contigs, reads.n_contigs, n_buckets, cnt);
  k_scan_totals<<<(unsigned)n_chunks, 256, 0, st>>>(cnt, n_buckets, ctx->scan_totals.p);
  k_scan_chunks<<<1, 1024, 0, st>>>(ctx->scan_totals.p, n_chunks);
  k_scan_final<uint32_t><<<(unsigned)n_chunks, 256, 0, st>>>(cnt, n_buckets, ctx->scan_totals.p, base);
  k_ord_scatter<<<grid, 256, 0, st>>>(recs, rec_bytes, (uint32_t)n, contigs, reads.n_contigs, n_buckets, cnt, base, order);
  k_ord_buckets<<<grid_for(n_buckets, 256, ctx->sm_count), 256, 0, st>>>(recs, rec_bytes, pool, n_buckets, base, order);
  k_ord_gather<<<grid_for(n * (rec_bytes / 8), 256, ctx->sm_count), 256, 0, st>>>(reinterpret_cast<const unsigned long long*>(recs), rec_bytes / 8, (uint32_t)n, order,
                                                                                reinterpret_cast<unsigned long long*>(ctx->ord_out.p));
  CUDA_OK(cudaGetLastError());
  return ctx->ord_out.p;
}

}  // namespace
