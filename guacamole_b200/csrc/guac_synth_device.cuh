// guac_synth_device.cuh — device build of the synthetic read generator (guac_synth_core.h): the columns of a guac_read_batch
// generated straight into HBM.  The whole-genome shape (619 M reads, 124 GB of raw columns) would otherwise be bound by PCIe;
// every rank generates the reads of its own loci shard from the seed (include/guac_synth.h).
//   k_synth_count   CTA per 1024 loci of the start windows: reads starting at every locus (Poisson table), block totals
//   k_synth_starts  same blocks again after the scan of the totals: start / contig / rank of every read, in (contig, start) order
//   k_synth_sizes   thread per read: CIGAR operators and MD characters the read will need (the generator run into a counting sink)
//   k_synth_write   thread per read: the read itself
#pragma once

#include "guac_host.cuh"
#include "guac_synth_core.h"
#include "guac_synth_tables.h"

namespace guac {

struct SynthWin {
  int32_t contig;
  int32_t pad_;
  int64_t start, end;     // loci where reads may start
  int64_t contig_length;
  uint64_t block0;        // first 1024-loci block of this window in the concatenated block list
};

constexpr int kSynthBlock = 1024;

__device__ __forceinline__ uint32_t synth_locus_count(const gsynth::Tables& T, const SynthWin& w, int64_t p) {
  const gsynth::Genome G{T.seed, T.sample};
  if (p >= w.end || !gsynth::start_allowed(G, w.contig, p, w.contig_length, T.read_length)) return 0u;
  return gsynth::reads_starting_at(T, w.contig, p);
}

__device__ __forceinline__ uint32_t synth_window_of(const SynthWin* wins, uint32_t n_wins, uint64_t block) {
  uint32_t lo = 0, hi = n_wins - 1;
  while (lo < hi) {
    const uint32_t mid = (lo + hi + 1) >> 1;
    if (wins[mid].block0 <= block) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// exclusive scan inside a CTA of 256 threads; returns the thread's exclusive prefix, *total = CTA total (valid in all threads)
__device__ __forceinline__ uint32_t cta_exclusive_scan_256(uint32_t v, uint32_t* total) {
  __shared__ uint32_t warp_sum[8];
  __shared__ uint32_t cta_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();  // (warp_sum may still be read by a previous call)
  if (lane == 31) warp_sum[warp] = incl;
  __syncthreads();
  uint32_t base = 0;
  for (int w = 0; w < warp; ++w) base += warp_sum[w];
  if (threadIdx.x == 255) cta_total = base + incl;
  __syncthreads();
  *total = cta_total;
  return base + incl - v;
}

__global__ void __launch_bounds__(256) k_synth_count(const gsynth::Tables* __restrict__ T, const SynthWin* __restrict__ wins, uint32_t n_wins,
                                                     uint64_t n_blocks, uint32_t* __restrict__ block_count) {
  for (uint64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
    const SynthWin w = wins[synth_window_of(wins, n_wins, b)];
    const int64_t p0 = w.start + (int64_t)(b - w.block0) * kSynthBlock + (int64_t)threadIdx.x * 4;
    uint32_t c = 0;
    for (int k = 0; k < 4; ++k) c += synth_locus_count(*T, w, p0 + k);
    uint32_t total;
    cta_exclusive_scan_256(c, &total);
    if (threadIdx.x == 0) block_count[b] = total;
  }
}

__global__ void __launch_bounds__(256) k_synth_starts(const gsynth::Tables* __restrict__ T, const SynthWin* __restrict__ wins, uint32_t n_wins,
                                                      uint64_t n_blocks, const uint64_t* __restrict__ block_off, int64_t* __restrict__ start,
                                                      int32_t* __restrict__ contig, int32_t* __restrict__ sample, uint16_t* __restrict__ rank) {
  for (uint64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
    const SynthWin w = wins[synth_window_of(wins, n_wins, b)];
    const int64_t p0 = w.start + (int64_t)(b - w.block0) * kSynthBlock + (int64_t)threadIdx.x * 4;
    uint32_t k4[4], c = 0;
    for (int k = 0; k < 4; ++k) { k4[k] = synth_locus_count(*T, w, p0 + k); c += k4[k]; }
    uint32_t total;
    uint64_t r = block_off[b] + cta_exclusive_scan_256(c, &total);
    for (int k = 0; k < 4; ++k)
      for (uint32_t j = 0; j < k4[k]; ++j, ++r) {
        start[r] = p0 + k;
        contig[r] = w.contig;
        sample[r] = T->sample;
        rank[r] = (uint16_t)j;
      }
  }
}

__global__ void __launch_bounds__(128) k_synth_sizes(const gsynth::Tables* __restrict__ T, uint64_t n, const int64_t* __restrict__ start,
                                                     const int32_t* __restrict__ contig, const uint16_t* __restrict__ rank, uint32_t* __restrict__ n_ops,
                                                     uint32_t* __restrict__ md_len) {
  for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
    gsynth::CountSink cs;
    gsynth::make_read(*T, contig[r], start[r], rank[r], cs);
    n_ops[r] = cs.n_ops;
    md_len[r] = cs.md_len;
  }
}

__global__ void __launch_bounds__(128) k_synth_write(const gsynth::Tables* __restrict__ T, uint64_t n, const int64_t* __restrict__ start,
                                                     const int32_t* __restrict__ contig, const uint16_t* __restrict__ rank,
                                                     const uint64_t* __restrict__ cigar_off, const uint64_t* __restrict__ md_off, uint64_t* __restrict__ seq_off,
                                                     uint32_t* __restrict__ cigar, uint8_t* __restrict__ seq, uint8_t* __restrict__ qual, char* __restrict__ md,
                                                     uint8_t* __restrict__ mapq, uint8_t* __restrict__ flags) {
  const uint64_t L = (uint64_t)T->read_length;
  for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
    seq_off[r] = r * L;
    if (r + 1 == n) seq_off[n] = n * L;
    gsynth::WriteSink ws{cigar + cigar_off[r], seq + r * L, qual ? qual + r * L : nullptr, md + md_off[r], mapq + r, flags + r};
    gsynth::make_read(*T, contig[r], start[r], rank[r], ws);
  }
}

// ---- exclusive scan of a u32 array into u64 / u32 offsets (out[n] = total): chunk totals, scan of the totals, final pass -------
constexpr int kScanChunk = 2048;  // elements per CTA of 256 threads

__global__ void __launch_bounds__(256) k_scan_totals(const uint32_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ chunk_total) {
  const uint64_t base = (uint64_t)blockIdx.x * kScanChunk;
  uint32_t s = 0;
  for (int k = 0; k < kScanChunk / 256; ++k) {
    const uint64_t i = base + (uint64_t)k * 256 + threadIdx.x;
    if (i < n) s += in[i];
  }
  uint32_t total;
  cta_exclusive_scan_256(s, &total);
  if (threadIdx.x == 0) chunk_total[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_chunks(uint64_t* __restrict__ chunk_total, uint64_t n_chunks) {  // in place, exclusive; [n_chunks] = total
  __shared__ uint64_t warp_sum[32];
  __shared__ uint64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint64_t base = 0; base <= n_chunks; base += 1024) {
    const uint64_t i = base + threadIdx.x;
    const uint64_t v = i < n_chunks ? chunk_total[i] : 0ull;
    uint64_t incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const uint64_t w = warp_sum[lane];
      uint64_t wi = w;
      for (int o = 1; o < 32; o <<= 1) {
        const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_sum[lane] = wi - w;
    }
    __syncthreads();
    const uint64_t excl = carry + warp_sum[warp] + incl - v;
    if (i <= n_chunks) chunk_total[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
}

template <typename OutT>
__global__ void __launch_bounds__(256) k_scan_final(const uint32_t* __restrict__ in, uint64_t n, const uint64_t* __restrict__ chunk_base, OutT* __restrict__ out) {
  const uint64_t base = (uint64_t)blockIdx.x * kScanChunk;
  uint64_t run = chunk_base[blockIdx.x];
  for (int k = 0; k < kScanChunk / 256; ++k) {
    const uint64_t i = base + (uint64_t)k * 256 + threadIdx.x;
    const uint32_t v = i < n ? in[i] : 0u;
    uint32_t total;
    const uint32_t excl = cta_exclusive_scan_256(v, &total);
    if (i < n) out[i] = (OutT)(run + excl);
    run += total;
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = (OutT)chunk_base[gridDim.x];
}

}  // namespace guac

namespace {

// out[0..n] = exclusive prefix sums of in[0..n) (out[n] = total); returns the total (synchronises the stream)
template <typename OutT>
uint64_t device_exclusive_scan(guac_ctx* ctx, const uint32_t* in, uint64_t n, OutT* out) {
  cudaStream_t st = ctx->stream;
  if (n == 0) {
    CUDA_OK(cudaMemsetAsync(out, 0, sizeof(OutT), st));
    return 0;
  }
  const uint64_t n_chunks = (n + kScanChunk - 1) / kScanChunk;
  DevBuf<uint64_t> totals;
  totals.alloc(n_chunks + 1);
  k_scan_totals<<<(unsigned)n_chunks, 256, 0, st>>>(in, n, totals.p);
  k_scan_chunks<<<1, 1024, 0, st>>>(totals.p, n_chunks);
  k_scan_final<OutT><<<(unsigned)n_chunks, 256, 0, st>>>(in, n, totals.p, out);
  CUDA_OK(cudaGetLastError());
  uint64_t total = 0;
  CUDA_OK(cudaMemcpyAsync(&total, totals.p + n_chunks, sizeof total, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  return total;
}

}  // namespace

struct guac_synth_device_batch {
  guac_ctx* ctx = nullptr;
  int device = 0;
  std::vector<int64_t> contig_length;
  DevBuf<int64_t> start;
  DevBuf<int32_t> contig, sample;
  DevBuf<uint64_t> cigar_off, seq_off, md_off;
  DevBuf<uint32_t> cigar;
  DevBuf<uint8_t> seq, qual, mapq, flags;
  DevBuf<char> md;
  uint64_t n = 0, n_ops = 0, n_md = 0, n_bases = 0;
  double kernel_ms = 0;
  bool consumed = false;  // guac_reads_pack_synth moved the large columns into a read set
  guac_read_batch view{};
};

struct guac_synth_host_batch {
  std::vector<int64_t> contig_length;
  std::vector<void*> blocks;
  bool pinned = false;
  guac_read_batch view{};
  ~guac_synth_host_batch() {
    for (void* p : blocks) {
      if (pinned) cudaFreeHost(p);
      else free(p);
    }
  }
};

namespace {

void synth_generate_device(guac_ctx* ctx, const guac_synth_params& P, guac_synth_device_batch& B) {
  gsynth::Tables T;
  if (!P.contig_length || P.n_contigs == 0 || (P.n_windows && !P.windows) || !gsynth::build_tables(P, T)) fail(GUAC_ERR_INVALID_ARGUMENT, "bad generator parameters");
  cudaStream_t st = ctx->stream;
  B.ctx = ctx;
  B.device = ctx->device;
  B.contig_length.assign(P.contig_length, P.contig_length + P.n_contigs);
  const std::vector<guac_locus_range> windows = gsynth::start_windows(P);
  std::vector<SynthWin> wins;
  uint64_t n_blocks = 0;
  for (const guac_locus_range& w : windows) {
    wins.push_back(SynthWin{w.contig, 0, w.start, w.end, P.contig_length[w.contig], n_blocks});
    n_blocks += (uint64_t)((w.end - w.start + kSynthBlock - 1) / kSynthBlock);
  }
  DevBuf<gsynth::Tables> d_T;
  DevBuf<SynthWin> d_wins;
  h2d(ctx, d_T, &T, 1);
  h2d(ctx, d_wins, wins.data(), wins.size());
  CUDA_OK(cudaEventRecord(ctx->ev[2], st));
  uint64_t n = 0;
  DevBuf<uint16_t> rank;
  if (n_blocks) {
    DevBuf<uint32_t> block_count;
    DevBuf<uint64_t> block_off;
    block_count.alloc(n_blocks);
    block_off.alloc(n_blocks + 1);
    const unsigned grid = (unsigned)std::min<uint64_t>(n_blocks, (uint64_t)ctx->sm_count * 64);
    k_synth_count<<<grid, 256, 0, st>>>(d_T.p, d_wins.p, (uint32_t)wins.size(), n_blocks, block_count.p);
    n = device_exclusive_scan<uint64_t>(ctx, block_count.p, n_blocks, block_off.p);
    if (n >= 0xFFFFFFF0ull) fail(GUAC_ERR_UNSUPPORTED, "more than 2^32 reads in one generated batch: shard it");
    B.start.alloc(n + 1);
    B.contig.alloc(n + 1);
    B.sample.alloc(n + 1);
    rank.alloc(n + 1);
    if (n) k_synth_starts<<<grid, 256, 0, st>>>(d_T.p, d_wins.p, (uint32_t)wins.size(), n_blocks, block_off.p, B.start.p, B.contig.p, B.sample.p, rank.p);
  } else {
    B.start.alloc(1); B.contig.alloc(1); B.sample.alloc(1); rank.alloc(1);
  }
  B.n = n;
  B.cigar_off.alloc(n + 1);
  B.md_off.alloc(n + 1);
  B.seq_off.alloc(n + 1);
  B.mapq.alloc(n + 1);
  B.flags.alloc(n + 1);
  {
    DevBuf<uint32_t> n_ops, md_len;
    n_ops.alloc(n + 1);
    md_len.alloc(n + 1);
    if (n) k_synth_sizes<<<grid_for(n, 128, ctx->sm_count), 128, 0, st>>>(d_T.p, n, B.start.p, B.contig.p, rank.p, n_ops.p, md_len.p);
    B.n_ops = device_exclusive_scan<uint64_t>(ctx, n_ops.p, n, B.cigar_off.p);
    B.n_md = device_exclusive_scan<uint64_t>(ctx, md_len.p, n, B.md_off.p);
  }
  B.n_bases = n * (uint64_t)P.read_length;
  B.cigar.alloc(B.n_ops + 4);
  B.md.alloc(B.n_md + 16);
  B.seq.alloc(B.n_bases + 64);
  CUDA_OK(cudaMemsetAsync(B.cigar.p + B.n_ops, 0, 4 * sizeof(uint32_t), st));
  CUDA_OK(cudaMemsetAsync(B.md.p + B.n_md, 0, 16, st));
  CUDA_OK(cudaMemsetAsync(B.seq.p + B.n_bases, 0, 64, st));
  if (P.with_qualities) {
    B.qual.alloc(B.n_bases + 64);
    CUDA_OK(cudaMemsetAsync(B.qual.p + B.n_bases, 0, 64, st));
  }
  if (n) {
    k_synth_write<<<grid_for(n, 128, ctx->sm_count), 128, 0, st>>>(d_T.p, n, B.start.p, B.contig.p, rank.p, B.cigar_off.p, B.md_off.p, B.seq_off.p, B.cigar.p,
                                                                   B.seq.p, P.with_qualities ? B.qual.p : nullptr, B.md.p, B.mapq.p, B.flags.p);
  } else {
    CUDA_OK(cudaMemsetAsync(B.seq_off.p, 0, sizeof(uint64_t), st));
  }
  CUDA_OK(cudaEventRecord(ctx->ev[3], st));
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(st));
  float ms = 0;
  CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
  B.kernel_ms = ms;
  guac_read_batch& v = B.view;
  v.n_reads = n;
  v.n_contigs = P.n_contigs;
  v.contig_length = B.contig_length.data();  // host
  v.contig = B.contig.p;
  v.start = B.start.p;
  v.cigar_off = B.cigar_off.p;
  v.cigar = B.cigar.p;
  v.seq_off = B.seq_off.p;
  v.seq = B.seq.p;
  v.qual = P.with_qualities ? B.qual.p : nullptr;
  v.mapq = B.mapq.p;
  v.flags = B.flags.p;
  v.sample = B.sample.p;
  v.md_off = B.md_off.p;
  v.md = B.md.p;
}

void synth_download(guac_ctx* ctx, const guac_synth_device_batch& D, bool pinned, guac_synth_host_batch& H) {
  H.pinned = pinned;
  H.contig_length = D.contig_length;
  auto fetch = [&](const void* src, size_t bytes) -> void* {
    void* p = nullptr;
    const size_t want = std::max<size_t>(bytes, 64);
    if (pinned) CUDA_OK(cudaMallocHost(&p, want));
    else if (!(p = malloc(want))) fail(GUAC_ERR_OOM, "host allocation of %zu bytes failed", want);
    H.blocks.push_back(p);
    if (bytes && src) CUDA_OK(cudaMemcpyAsync(p, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    else memset(p, 0, want);
    return p;
  };
  const uint64_t n = D.n;
  guac_read_batch& v = H.view;
  v.n_reads = n;
  v.n_contigs = (uint32_t)D.contig_length.size();
  v.contig_length = H.contig_length.data();
  v.contig = (const int32_t*)fetch(D.contig.p, n * 4);
  v.start = (const int64_t*)fetch(D.start.p, n * 8);
  v.cigar_off = (const uint64_t*)fetch(D.cigar_off.p, (n + 1) * 8);
  v.cigar = (const uint32_t*)fetch(D.cigar.p, D.n_ops * 4);
  v.seq_off = (const uint64_t*)fetch(D.seq_off.p, (n + 1) * 8);
  v.seq = (const uint8_t*)fetch(D.seq.p, D.n_bases);
  v.qual = (const uint8_t*)fetch(D.qual.n ? D.qual.p : nullptr, D.n_bases);
  v.mapq = (const uint8_t*)fetch(D.mapq.p, n);
  v.flags = (const uint8_t*)fetch(D.flags.p, n);
  v.sample = (const int32_t*)fetch(D.sample.p, n * 4);
  v.md_off = (const uint64_t*)fetch(D.md_off.p, (n + 1) * 8);
  v.md = (const char*)fetch(D.md.p, D.n_md);
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
}

}  // namespace
