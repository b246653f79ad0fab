// guac_batch2.cuh — the compact read batch (guac_read_batch_v2, include/guac.h): the columns as a BAM record holds them.
//
// Read.fromSAMRecord (reads/Read.scala:217-291) decodes a record's 4-bit bases to one byte each and widens every field before
// the reads reach the callers; guac_reads_pack takes that widened form (199 B per 150 bp read across PCIe).  The compact batch
// keeps the fields at their BAM width (97 B per read) and widens them HERE, on the device: k_widen_header turns the 32-bit
// columns into the ones k_header reads, k_unpack_bases turns each chunk of nibbles into the ASCII bases the store keeps —
// both on the copy stream, between one chunk's copy and the next.  Everything downstream is the same code as for
// guac_read_batch, so the packed store and every record are identical.
#pragma once

#include <atomic>
#include <climits>
#include <thread>

#include "guac_host.cuh"

namespace guac {

struct WidenArgs {
  uint64_t n;
  uint32_t n_contigs, read_length;
  const unsigned long long* contig_read_off;  // [n_contigs + 1]
  const int32_t* start;
  const uint32_t* cigar_off;
  const uint32_t* seq_off;                    // null with read_length
  const uint32_t* md_off;
  int32_t* contig_w;
  long long* start_w;
  unsigned long long* cigar_off_w;
  unsigned long long* seq_off_w;
  unsigned long long* md_off_w;
};

__global__ void __launch_bounds__(256) k_widen_header(WidenArgs A) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i <= A.n; i += (uint64_t)gridDim.x * blockDim.x) {
    A.cigar_off_w[i] = A.cigar_off[i];
    A.md_off_w[i] = A.md_off[i];
    A.seq_off_w[i] = A.read_length ? i * (unsigned long long)A.read_length : (unsigned long long)A.seq_off[i];
    if (i < A.n) {
      A.start_w[i] = A.start[i];
      uint32_t lo = 0, hi = A.n_contigs - 1;  // the last contig whose first read is at or below i (empty contigs share offsets)
      while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (A.contig_read_off[mid] <= i) lo = mid; else hi = mid - 1;
      }
      A.contig_w[i] = (int32_t)lo;
    }
  }
}

// bases [g_begin, g_end) of the batch: nibbles -> ASCII.  One thread per 32 bases (16 bytes in, 32 out); the threads at the two
// ends of the range go base by base, so that neighbouring chunks never write each other's bytes.
__global__ void __launch_bounds__(256) k_unpack_bases(const uint8_t* __restrict__ seq4, uint8_t* __restrict__ seq, unsigned long long g_begin,
                                                      unsigned long long g_end) {
  __shared__ uint16_t two[256];  // a byte of nibbles -> its two letters (the high nibble's first: the lower address)
  {
    const unsigned long long lo = 0x565352474D43413Dull, hi = 0x4E42444B48595754ull;  // "=ACMGRSV", "TWYHKDBN", first letter lowest
    const uint32_t h = threadIdx.x >> 4, l = threadIdx.x & 15u;
    const uint32_t ch = (uint32_t)((h < 8 ? lo : hi) >> (8 * (h & 7u))) & 0xFFu, cl = (uint32_t)((l < 8 ? lo : hi) >> (8 * (l & 7u))) & 0xFFu;
    two[threadIdx.x] = (uint16_t)(ch | cl << 8);
  }
  __syncthreads();
  const unsigned long long t_begin = g_begin >> 5, t_end = (g_end + 31) >> 5;
  for (unsigned long long t = t_begin + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; t < t_end; t += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long g0 = t << 5;
    if (g0 >= g_begin && g0 + 32 <= g_end) {
      const uint4 in = __ldg(reinterpret_cast<const uint4*>(seq4 + (g0 >> 1)));
      const uint32_t w[4] = {in.x, in.y, in.z, in.w};
      uint32_t o[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        o[2 * k] = (uint32_t)two[w[k] & 0xFFu] | (uint32_t)two[(w[k] >> 8) & 0xFFu] << 16;
        o[2 * k + 1] = (uint32_t)two[(w[k] >> 16) & 0xFFu] | (uint32_t)two[w[k] >> 24] << 16;
      }
      uint4* out = reinterpret_cast<uint4*>(seq + g0);
      out[0] = make_uint4(o[0], o[1], o[2], o[3]);
      out[1] = make_uint4(o[4], o[5], o[6], o[7]);
    } else {
      const unsigned long long lo = g0 > g_begin ? g0 : g_begin, hi = g0 + 32 < g_end ? g0 + 32 : g_end;
      for (unsigned long long g = lo; g < hi; ++g) {
        const uint32_t pair = two[seq4[g >> 1]];
        seq[g] = (uint8_t)((g & 1ull) ? pair >> 8 : pair);
      }
    }
  }
}

}  // namespace guac

// ---- host side: guac_read_batch -> guac_read_batch_v2 ------------------------------------------------------------------------------
struct guac_host_batch_v2 {
  std::vector<int64_t> contig_length;
  std::vector<uint64_t> contig_read_off;
  std::vector<void*> blocks;
  bool pinned = false;
  uint64_t bytes = 0;
  guac_read_batch_v2 view{};
  // guac_bam_load: the file's sequence dictionary, the sample of the kept reads, and what the decode cost
  std::vector<std::string> contig_names;
  std::string sample_name = "default";
  uint64_t file_bytes = 0, inflated_bytes = 0, records_in_file = 0;
  double decode_ms = 0;
  ~guac_host_batch_v2() {
    for (void* p : blocks) {
      if (pinned) cudaFreeHost(p);
      else free(p);
    }
  }
  template <typename T>
  T* take(size_t count) {
    void* p = nullptr;
    const size_t sz = std::max<size_t>(count * sizeof(T), 64);
    if (pinned) {
      if (cudaMallocHost(&p, sz) != cudaSuccess) { cudaGetLastError(); p = nullptr; }
    } else {
      p = malloc(sz);
    }
    if (!p) fail(GUAC_ERR_OOM, "host buffer of %zu bytes", sz);
    blocks.push_back(p);
    return static_cast<T*>(p);
  }
};

namespace {

void compact_batch(const guac_read_batch& b, bool fixed_length, guac_host_batch_v2& H) {
  const uint64_t n = b.n_reads;
  if (n && (!b.contig || !b.start || !b.cigar_off || !b.seq_off || !b.seq || !b.mapq || !b.flags || !b.md_off))
    fail(GUAC_ERR_INVALID_ARGUMENT, "null column in read batch");
  const uint64_t n_ops = n ? b.cigar_off[n] : 0, n_md = n ? b.md_off[n] : 0, n_bases = n ? b.seq_off[n] : 0;
  if (n_ops >= 0xFFFFFFFFull || n_md >= 0xFFFFFFFFull || n_bases >= 0xFFFFFFFFull)
    fail(GUAC_ERR_UNSUPPORTED, "the compact batch holds fewer than 2^32 bases, CIGAR ops and MD bytes: split the batch");
  guac_read_batch_v2& V = H.view;
  V.n_reads = n;
  V.n_contigs = b.n_contigs;
  if (b.contig_length) {
    H.contig_length.assign(b.contig_length, b.contig_length + b.n_contigs);
    V.contig_length = H.contig_length.data();
  }
  H.contig_read_off.assign((size_t)b.n_contigs + 1, n);
  {
    int64_t prev = -1;
    for (uint64_t i = 0; i < n; ++i) {
      const int64_t c = b.contig[i];
      if (c < 0 || c >= (int64_t)b.n_contigs) fail(GUAC_ERR_INVALID_ARGUMENT, "read %llu: contig index out of range", (unsigned long long)i);
      if (c < prev) fail(GUAC_ERR_INVALID_ARGUMENT, "read %llu: the compact batch needs the reads grouped by ascending contig", (unsigned long long)i);
      for (int64_t k = prev + 1; k <= c; ++k) H.contig_read_off[(size_t)k] = i;
      prev = c;
      if (b.sample && b.sample[i] != b.sample[0]) fail(GUAC_ERR_INVALID_ARGUMENT, "read %llu: the compact batch holds one sample", (unsigned long long)i);
      if (b.start[i] < INT32_MIN || b.start[i] > INT32_MAX) fail(GUAC_ERR_UNSUPPORTED, "read %llu: start beyond 32 bits", (unsigned long long)i);
    }
    H.contig_read_off[0] = 0;
  }
  V.contig_read_off = H.contig_read_off.data();
  V.sample = (n && b.sample) ? b.sample[0] : 0;
  bool fixed = fixed_length && n > 0;
  const uint64_t L = n ? b.seq_off[1] - b.seq_off[0] : 0;
  for (uint64_t i = 0; fixed && i < n; ++i) fixed = b.seq_off[i + 1] - b.seq_off[i] == L;
  fixed = fixed && L > 0 && L < 0xFFFFFFFFull;
  int32_t* start = H.take<int32_t>(n);
  uint32_t* cigar_off = H.take<uint32_t>(n + 1);
  uint32_t* md_off = H.take<uint32_t>(n + 1);
  uint32_t* seq_off = fixed ? nullptr : H.take<uint32_t>(n + 1);
  uint32_t* cigar = H.take<uint32_t>((size_t)n_ops);
  uint8_t* seq4 = H.take<uint8_t>((size_t)(n_bases + 1) / 2 + 16);
  uint8_t* qual = b.qual ? H.take<uint8_t>((size_t)n_bases) : nullptr;
  uint8_t* mapq = H.take<uint8_t>(n);
  uint8_t* flags = H.take<uint8_t>(n);
  char* md = H.take<char>((size_t)n_md);
  for (uint64_t i = 0; i < n; ++i) start[i] = (int32_t)b.start[i];
  for (uint64_t i = 0; i <= n && n; ++i) {
    cigar_off[i] = (uint32_t)b.cigar_off[i];
    md_off[i] = (uint32_t)b.md_off[i];
    if (seq_off) seq_off[i] = (uint32_t)b.seq_off[i];
  }
  if (!n) { cigar_off[0] = md_off[0] = 0; if (seq_off) seq_off[0] = 0; }
  if (n_ops) memcpy(cigar, b.cigar, (size_t)n_ops * 4);
  if (n_md) memcpy(md, b.md, (size_t)n_md);
  if (n) { memcpy(mapq, b.mapq, n); memcpy(flags, b.flags, n); }
  if (qual) memcpy(qual, b.qual, (size_t)n_bases);
  uint8_t code[256];
  memset(code, 0xFF, sizeof code);
  {
    const char* letters = "=ACMGRSVTWYHKDBN";
    for (int k = 0; k < 16; ++k) code[(uint8_t)letters[k]] = (uint8_t)k;
  }
  memset(seq4, 0, (size_t)(n_bases + 1) / 2 + 16);
  const unsigned n_thr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  std::vector<std::thread> pool;
  std::atomic<uint64_t> bad{~0ull};
  const uint64_t n_bytes = (n_bases + 1) / 2;
  for (unsigned t = 0; t < n_thr; ++t)
    pool.emplace_back([&, t] {
      const uint64_t b0 = n_bytes * t / n_thr, b1 = n_bytes * (t + 1) / n_thr;
      for (uint64_t k = b0; k < b1; ++k) {
        const uint8_t hi = code[b.seq[2 * k]], lo = 2 * k + 1 < n_bases ? code[b.seq[2 * k + 1]] : 0;
        if ((hi | lo) & 0xF0) {
          uint64_t seen = bad.load();
          const uint64_t at = (hi & 0xF0) ? 2 * k : 2 * k + 1;
          while (at < seen && !bad.compare_exchange_weak(seen, at)) {}
          continue;
        }
        seq4[k] = (uint8_t)(hi << 4 | lo);
      }
    });
  for (std::thread& th : pool) th.join();
  if (bad.load() != ~0ull) fail(GUAC_ERR_INVALID_ARGUMENT, "base %llu of the batch is not one of \"=ACMGRSVTWYHKDBN\": use guac_read_batch", (unsigned long long)bad.load());
  V.read_length = fixed ? (uint32_t)L : 0u;
  V.start = start;
  V.cigar_off = cigar_off;
  V.cigar = cigar;
  V.seq_off = seq_off;
  V.seq4 = seq4;
  V.qual = qual;
  V.mapq = mapq;
  V.flags = flags;
  V.md_off = md_off;
  V.md = md;
  H.bytes = n * (4 + 4 + 4 + (fixed ? 0 : 4) + 1 + 1) + n_ops * 4 + (n_bases + 1) / 2 + (qual ? n_bases : 0) + n_md;
}

}  // namespace
