// guac_host.cuh — host-side state shared by the C-ABI translation unit: context, device buffers, packed read store,
// results.  (Internal; the public contract is include/guac.h.)
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "guac_device.cuh"

using namespace guac;

// ---- errors / device buffers -----------------------------------------------------------------------------------------------
struct guac_ctx;

namespace {

struct StatusError {
  guac_status code;
  std::string msg;
};

[[noreturn]] void fail(guac_status code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw StatusError{code, buf};
}

#define CUDA_OK(expr)                                                                                          \
  do {                                                                                                         \
    cudaError_t e_ = (expr);                                                                                   \
    if (e_ != cudaSuccess)                                                                                     \
      fail(e_ == cudaErrorMemoryAllocation ? GUAC_ERR_OOM : GUAC_ERR_CUDA, "%s: %s (%s:%d)", #expr,            \
           cudaGetErrorString(e_), __FILE__, __LINE__);                                                        \
  } while (0)

// Device buffers come from a small exact-size cache kept per host thread (i.e. per context user): repeated pack / call
// cycles over batches of the same shape reuse their allocations instead of paying cudaMalloc / cudaFree every time.
struct DevCache {
  std::multimap<std::pair<int, size_t>, void*> free_blocks;  // (device, bytes) -> block
  size_t cached_bytes = 0;
  static constexpr size_t kMaxCachedBytes = 64ull << 30;
  static int device() {
    int d = 0;
    cudaGetDevice(&d);
    return d;
  }
  void* take(size_t bytes) {
    auto it = free_blocks.find({device(), bytes});
    if (it == free_blocks.end()) return nullptr;
    void* p = it->second;
    free_blocks.erase(it);
    cached_bytes -= bytes;
    return p;
  }
  void give(void* p, size_t bytes) {
    if (cached_bytes + bytes > kMaxCachedBytes || free_blocks.size() >= 512) {
      cudaFree(p);
      return;
    }
    free_blocks.emplace(std::make_pair(device(), bytes), p);
    cached_bytes += bytes;
  }
  void trim() {
    for (auto& kv : free_blocks) cudaFree(kv.second);
    free_blocks.clear();
    cached_bytes = 0;
  }
  ~DevCache() { trim(); }
};
static thread_local DevCache tl_dev_cache;

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  size_t alloc_bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) tl_dev_cache.give(p, alloc_bytes);
    p = nullptr;
    n = 0;
    alloc_bytes = 0;
  }
  void alloc(size_t count) {
    release();
    n = count;
    alloc_bytes = ((std::max<size_t>(count, 1) * sizeof(T)) + 255) & ~(size_t)255;
    p = (T*)tl_dev_cache.take(alloc_bytes);
    if (!p) {
      cudaError_t e = cudaMalloc((void**)&p, alloc_bytes);
      if (e == cudaErrorMemoryAllocation) {  // give the cache back and retry once
        cudaGetLastError();
        tl_dev_cache.trim();
        e = cudaMalloc((void**)&p, alloc_bytes);
      }
      if (e != cudaSuccess) {
        p = nullptr;
        n = 0;
        fail(e == cudaErrorMemoryAllocation ? GUAC_ERR_OOM : GUAC_ERR_CUDA, "cudaMalloc(%zu bytes): %s", alloc_bytes, cudaGetErrorString(e));
      }
    }
  }
  void adopt(DevBuf& o) {  // takes o's allocation (o is left empty)
    release();
    p = o.p; n = o.n; alloc_bytes = o.alloc_bytes;
    o.p = nullptr; o.n = 0; o.alloc_bytes = 0;
  }
  bool ensure(size_t count) {  // grow-only; true if (re)allocated
    if (p && n >= count) return false;
    alloc(count);
    return true;
  }
  size_t bytes() const { return n * sizeof(T); }
};

}  // namespace

// Page-locked host blocks for results: a call downloads its records straight into the block its guac_result will own
// (no staging copy); guac_result_free hands the block back for the next call.  Shared between the context and its
// results so that either may be destroyed first.
struct PinnedPool {
  std::mutex mu;
  std::multimap<size_t, void*> free_blocks;
  void* take(size_t bytes, size_t* got) {
    {
      std::lock_guard<std::mutex> g(mu);
      auto it = free_blocks.lower_bound(bytes);
      if (it != free_blocks.end() && it->first <= bytes * 2 + (1 << 20)) {
        void* p = it->second;
        *got = it->first;
        free_blocks.erase(it);
        return p;
      }
    }
    void* p = nullptr;
    const size_t want = bytes + bytes / 4 + 4096;
    if (cudaMallocHost(&p, want) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    *got = want;
    return p;
  }
  void give(void* p, size_t bytes) {
    std::lock_guard<std::mutex> g(mu);
    if (free_blocks.size() >= 8) {
      cudaFreeHost(p);
      return;
    }
    free_blocks.emplace(bytes, p);
  }
  ~PinnedPool() {
    for (auto& kv : free_blocks) cudaFreeHost(kv.second);
  }
};

// ---- context ----------------------------------------------------------------------------------------------------------------
// device status block: 16 counters followed by the device error word, fetched in one copy
constexpr size_t kStatusBytes = 16 * sizeof(unsigned long long) + sizeof(DevError);

// NVTX ranges around pack / copies / kernel families (visible in Nsight Systems; no cost without a profiler attached)
inline void nvtx_push(const char* name) { nvtxRangePushA(name); }
inline void nvtx_pop() { nvtxRangePop(); }

struct guac_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;       // the exact per-locus kernel of a germline call, next to the record egress on `stream`
  cudaStream_t stream3 = nullptr;       // ... and the ordering + egress of the compact records
  cudaEvent_t join_ev = nullptr, join3_ev = nullptr, seg_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t trace_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // GUAC_TRACE: the side streams' kernels of a germline call
  cudaStream_t copy_stream = nullptr;   // guac_reads_pack: host -> device copies, overlapped with the pack kernels
  cudaEvent_t copy_ev[10] = {};  // [0..7] chunks of bases on the device, [8] small columns there, [9] compute stream caught up
  std::string last_error;
  unsigned long long* d_counters = nullptr;  // 16 counters, then the DevError (kStatusBytes)
  DevError* d_err = nullptr;                 // = d_counters + 16
  unsigned long long* h_counters = nullptr;  // pinned mirror of the whole status block
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // [4], [5]: user stopwatch
  cudaEvent_t ev_rows[2] = {nullptr, nullptr};
  int sm_count = 148;
  // options
  int sort_records = 1;
  int pack_qualities = 1;
  int host_threads = 0;
  int difference_lists = 1;
  int segments = 1;
  bool pack_overlap = true;
  bool smem_attrs_done = false, expand_attrs_done = false;
  // scratch kept across calls so that a repeated call neither allocates nor rebuilds its tile list
  DevBuf<unsigned char> out_rec, out_pool, out_slow, out_compact, tiles, sort_rec;
  DevBuf<uint32_t> sort_bins;
  DevBuf<uint64_t> scan_totals;
  DevBuf<uint32_t> ord_u32;             // device-side canonical ordering of the likelihood callers' records (guac_order.cuh)
  DevBuf<unsigned char> ord_out;
  std::vector<guac_locus_range> tiles_key_ranges;
  uint64_t tiles_key_reads = 0;  // guac_reads::id of the cached tile list (0 = none)
  uint64_t tiles_key_loci = 0, n_tiles = 0;
  bool pool_head_ready = false, tiles_in_order = true;
  // pinned host staging for result downloads (grow-only)
  unsigned char* h_stage = nullptr;
  size_t h_stage_bytes = 0;
  // pinned host arena for the per-read header columns built by guac_reads_pack (grow-only)
  unsigned char* h_pack = nullptr;
  size_t h_pack_bytes = 0;
  std::shared_ptr<PinnedPool> pinned = std::make_shared<PinnedPool>();
  // somatic tables (device): see guac_somatic.cuh
  double* d_tables = nullptr;
  uint64_t generation = 0;                   // bumped by every call that reuses the output scratch
  unsigned long long* d_hist = nullptr;      // depth histogram of the last guac_depth_histogram (GUAC_DEPTH_BINS bins)
  bool hist_attr_done = false, som_attr_done = false, rows_attr_done = false;
};

namespace {

// GUAC_DEBUG_STALE=1: report a CUDA error some call left pending (cudaGetLastError is otherwise only consulted after launches)
// An error some earlier runtime call of this thread left pending (ours or another library's: the runtime's error state is per
// thread, not per library) must not be blamed on the first launch that asks: it is cleared on entry to every API call.
inline void report_stale_error(const char* when) {
  static const bool on = getenv("GUAC_DEBUG_STALE") != nullptr;
  const cudaError_t e = cudaGetLastError();
  if (on && e != cudaSuccess) fprintf(stderr, "[guac] pending CUDA error %s an API call: %s\n", when, cudaGetErrorString(e));
}

template <typename F>
guac_status guarded(guac_ctx* ctx, F&& f) {
  try {
    report_stale_error("before");
    f();
    return GUAC_OK;
  } catch (const StatusError& e) {
    if (ctx) ctx->last_error = e.msg;
    return e.code;
  } catch (const std::bad_alloc&) {
    if (ctx) ctx->last_error = "host allocation failed";
    return GUAC_ERR_OOM;
  }
}

void check_device_error(guac_ctx* ctx, const char* what) {
  DevError e;
  CUDA_OK(cudaMemcpyAsync(&e, ctx->d_err, sizeof e, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (e.code) {
    CUDA_OK(cudaMemsetAsync(ctx->d_err, 0, sizeof(DevError), ctx->stream));
    fail((guac_status)e.code, "%s: %s at read/locus %llu", what, guac_status_string((guac_status)e.code), e.where);
  }
}

// after a copy of the status block into ctx->h_counters has completed: throws the device-side error, if any
void raise_device_error(guac_ctx* ctx, const char* what) {
  DevError e;
  memcpy(&e, ctx->h_counters + 16, sizeof e);
  if (e.code) {
    CUDA_OK(cudaMemsetAsync(ctx->d_err, 0, sizeof(DevError), ctx->stream));
    fail((guac_status)e.code, "%s: %s at read/locus %llu", what, guac_status_string((guac_status)e.code), e.where);
  }
}

template <typename T>
void h2d_on(cudaStream_t s, DevBuf<T>& dst, const T* src, size_t n, size_t extra = 0) {
  dst.alloc(n + extra);
  if (extra) CUDA_OK(cudaMemsetAsync(dst.p + n, 0, extra * sizeof(T), s));
  if (n) CUDA_OK(cudaMemcpyAsync(dst.p, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
}
template <typename T>
void h2d(guac_ctx* ctx, DevBuf<T>& dst, const T* src, size_t n, size_t extra = 0) {
  h2d_on(ctx->stream, dst, src, n, extra);
}

unsigned char* stage(guac_ctx* ctx, size_t bytes) {
  if (ctx->h_stage_bytes < bytes) {
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    ctx->h_stage = nullptr;
    ctx->h_stage_bytes = 0;
    size_t want = bytes + bytes / 4 + 4096;
    CUDA_OK(cudaMallocHost((void**)&ctx->h_stage, want));
    ctx->h_stage_bytes = want;
  }
  return ctx->h_stage;
}

unsigned char* pack_arena(guac_ctx* ctx, size_t bytes) {
  if (ctx->h_pack_bytes < bytes) {
    if (ctx->h_pack) cudaFreeHost(ctx->h_pack);
    ctx->h_pack = nullptr;
    ctx->h_pack_bytes = 0;
    size_t want = bytes + bytes / 8 + 4096;
    CUDA_OK(cudaMallocHost((void**)&ctx->h_pack, want));
    ctx->h_pack_bytes = want;
  }
  return ctx->h_pack;
}

struct Trace {  // GUAC_TRACE=1 prints host-side phase times (diagnostics only)
  bool on;
  std::chrono::steady_clock::time_point t0;
  const char* what;
  Trace(const char* w) : on(getenv("GUAC_TRACE") != nullptr), t0(std::chrono::steady_clock::now()), what(w) {}
  void lap(const char* phase) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[guac %s] %-18s %8.3f ms\n", what, phase, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

// Canonical record order (contig, start, then `tie_less` among equal loci): LSD radix sort of (key, index) words on the
// bits of the (contig, start) key that actually vary, 13 bits per pass (one contig of up to 67 M loci: two passes over
// 8-byte words), one gather of the records, groups of equal keys finished with the full comparator.
// f(begin, end) over contiguous slices of [0, n) on up to `max_threads` host threads (the caller's included); slices shorter
// than `grain` are not worth a thread
template <typename F>
void host_parallel(size_t n, size_t grain, unsigned max_threads, F&& f) {
  const unsigned n_thr = (unsigned)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(max_threads, std::max(1u, std::thread::hardware_concurrency())), n / std::max<size_t>(grain, 1)));
  if (n_thr <= 1) { f((size_t)0, n); return; }
  std::vector<std::thread> pool;
  for (unsigned t = 1; t < n_thr; ++t) pool.emplace_back([&f, n, n_thr, t] { f(n * t / n_thr, n * (t + 1) / n_thr); });
  f((size_t)0, n / n_thr);
  for (std::thread& th : pool) th.join();
}

template <typename Rec, typename TieLess>
void sort_records_canonical(Rec* recs, size_t n, TieLess tie_less) {
  if (n < 2) return;
  constexpr unsigned kThreads = 8;
  const size_t grain = std::max<size_t>(1, (size_t)(1u << 20) / sizeof(Rec));  // a megabyte of records per thread at least
  int ib = 1;  // bits of a record index
  while (((n - 1) >> ib) != 0) ++ib;
  std::mutex mu;
  bool fits = true;
  uint64_t kmin = ~0ull, kmax = 0;
  host_parallel(n, grain, kThreads, [&](size_t b, size_t e) {
    bool ok = true;
    uint64_t lo = ~0ull, hi = 0;
    for (size_t i = b; i < e && ok; ++i) {
      ok = recs[i].contig >= 0 && recs[i].contig < 65536 && recs[i].start >= 0 && recs[i].start < (1ll << 32);
      const uint64_t k = ((uint64_t)recs[i].contig << 32) | (uint64_t)recs[i].start;
      lo = std::min(lo, k);
      hi = std::max(hi, k);
    }
    std::lock_guard<std::mutex> lk(mu);
    fits = fits && ok;
    kmin = std::min(kmin, lo);
    kmax = std::max(kmax, hi);
  });
  int bits = 0;
  while (fits && bits < 64 && ((kmax - kmin) >> bits) != 0) ++bits;
  if (!fits || bits + ib > 64) {  // (key - kmin) and the record index must share one 64-bit word
    std::sort(recs, recs + n, [&](const Rec& a, const Rec& b) {
      if (a.contig != b.contig) return a.contig < b.contig;
      if (a.start != b.start) return a.start < b.start;
      return tie_less(a, b);
    });
    return;
  }
  static thread_local std::vector<uint64_t> v, v2;  // scratch kept between calls: no page faults on the hot path
  static thread_local std::vector<unsigned char> tmp_bytes;
  if (v.size() < n) { v.resize(n); v2.resize(n); }
  {
    uint64_t* keys = v.data();  // (a thread_local is per thread: the workers get the caller's array through a plain pointer)
    host_parallel(n, grain, kThreads, [&, keys](size_t b, size_t e) {
      for (size_t i = b; i < e; ++i) keys[i] = (((((uint64_t)recs[i].contig << 32) | (uint64_t)recs[i].start) - kmin) << ib) | (uint64_t)i;
    });
  }
  bool sorted = true;  // device order is often already canonical for small outputs
  for (size_t i = 1; i < n && sorted; ++i) sorted = (v[i] >> ib) >= (v[i - 1] >> ib);
  if (!sorted) {
    constexpr int kDigit = 13;
    std::vector<uint32_t> count((1u << kDigit) + 1);
    for (int shift = ib; shift < ib + bits; shift += kDigit) {
      std::fill(count.begin(), count.end(), 0u);
      for (size_t i = 0; i < n; ++i) ++count[((v[i] >> shift) & ((1u << kDigit) - 1)) + 1];
      for (uint32_t d = 0; d < (1u << kDigit); ++d) count[d + 1] += count[d];
      for (size_t i = 0; i < n; ++i) v2[count[(v[i] >> shift) & ((1u << kDigit) - 1)]++] = v[i];
      v.swap(v2);
    }
  }
  if (tmp_bytes.size() < n * sizeof(Rec)) tmp_bytes.resize(n * sizeof(Rec));
  Rec* tmp = reinterpret_cast<Rec*>(tmp_bytes.data());
  const uint64_t imask = (1ull << ib) - 1;
  const uint64_t* order = v.data();
  if (!sorted) host_parallel(n, grain, kThreads, [&, order](size_t b, size_t e) { for (size_t i = b; i < e; ++i) tmp[i] = recs[order[i] & imask]; });
  Rec* cur = sorted ? recs : tmp;
  for (size_t i = 0; i < n;) {  // records of one locus: the full comparator
    size_t j = i + 1;
    while (j < n && (v[j] >> ib) == (v[i] >> ib)) ++j;
    if (j - i > 1) std::sort(cur + i, cur + j, tie_less);
    i = j;
  }
  if (!sorted) host_parallel(n, grain, kThreads, [&](size_t b, size_t e) { memcpy(recs + b, tmp + b, (e - b) * sizeof(Rec)); });
  if (n > (1u << 22)) {  // dense outputs: do not keep gigabytes of scratch around
    std::vector<uint64_t>().swap(v);
    std::vector<uint64_t>().swap(v2);
    std::vector<unsigned char>().swap(tmp_bytes);
  }
}

// The reference's LociSet merges overlapping ranges; here they are the caller's to merge: overlapping input would visit loci
// twice and return duplicate records, so it is refused (GUAC_ERR_INVALID_ARGUMENT, as include/guac.h says).
void check_ranges_disjoint(const guac_locus_range* ranges, size_t n_ranges) {
  std::vector<std::tuple<int32_t, int64_t, int64_t>> v;
  v.reserve(n_ranges);
  for (size_t i = 0; i < n_ranges; ++i)
    if (ranges[i].end > ranges[i].start) v.emplace_back(ranges[i].contig, ranges[i].start, ranges[i].end);
  bool sorted = true;
  for (size_t i = 1; i < v.size() && sorted; ++i) sorted = v[i - 1] <= v[i];
  if (!sorted) std::sort(v.begin(), v.end());
  for (size_t i = 1; i < v.size(); ++i)
    if (std::get<0>(v[i]) == std::get<0>(v[i - 1]) && std::get<1>(v[i]) < std::get<2>(v[i - 1]))
      fail(GUAC_ERR_INVALID_ARGUMENT, "loci ranges overlap (contig %d near locus %lld): merge them first", std::get<0>(v[i]), (long long)std::get<1>(v[i]));
}

int grid_for(uint64_t n, int block, int sm_count) {
  uint64_t g = (n + block - 1) / block;
  uint64_t cap = (uint64_t)sm_count * 32;
  return (int)std::max<uint64_t>(1, std::min(g, cap));
}

}  // namespace

// ---- packed read store ---------------------------------------------------------------------------------------------------
inline uint64_t next_object_id() {
  static std::atomic<uint64_t> counter{0};
  return ++counter;
}

struct guac_reads {
  guac_ctx* ctx = nullptr;      // (never dereferenced when the read set is freed: a context may be destroyed first)
  int device = 0;
  uint64_t id = next_object_id();
  uint64_t n = 0;
  uint32_t n_contigs = 0;
  int32_t sample = 0;
  std::vector<ContigInfo> contigs;
  DevBuf<ReadRec> rec;
  DevBuf<uint32_t> cig_off, cigar, xmask, md_off, trk_lo, trk_hi, trk_std, gran_first, gran_last;
  DevBuf<uint2> pairs;
  DevBuf<GranHdr> gs_hdr;             // per-granule difference streams (k_expand); empty when packed without them
  DevBuf<uint16_t> gs_diffs;
  DevBuf<uint8_t> gs_dd, gs_dp;
  DevBuf<uint32_t> gs_imp;
  bool gs_wide = false;
  uint64_t gs_entries = 0;
  // per 32-locus word, the overlapping reads as ROWS of (quality | base code << 6) bytes, one byte per locus of the word
  // (k_expand_rows, guac_rows.cuh): what the likelihood kernels stream.  Empty when packed without qualities / streams.
  DevBuf<uint4> q_hdr, q_groups;
  DevBuf<uint16_t> q_depth;
  DevBuf<uint32_t> q_cols, q_rows;
  uint64_t q_cap_pairs = 0, q_cap_groups = 0;
  double rows_ms = 0;
  uint32_t mapq_mask[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double expand_ms = 0;
  DevBuf<uint64_t> seq_off, fasta_off;
  DevBuf<uint8_t> seq, qual, qc, fasta;
  DevBuf<char> md;
  DevBuf<uint16_t> nm, del_len;
  DevBuf<int32_t> del_start;
  DevBuf<uint32_t> del_md;
  DevBuf<ContigInfo> d_contigs;
  uint64_t order_sensitive_loci = 0;
  uint64_t max_reads_per_granule = 0;
  int64_t max_ref_span = 0;  // longest reference span of a read (look-back of the start-sorted binary searches)
  uint64_t total_words = 0, total_grans = 0;
  double pack_kernel_ms = 0;
  int pack_launches = 0;
  uint64_t h2d_bytes = 0;
  bool has_qualities = true;

  DevReads view() const {
    DevReads R;
    R.n = n;
    R.max_ref_span = (int32_t)std::min<int64_t>(max_ref_span, 0x7FFFFFFF);
    R.pad_ = 0;
    R.rec = rec.p;
    R.cig_off = cig_off.p;
    R.cigar = cigar.p;
    R.pairs = pairs.p;
    R.xmask = xmask.p;
    R.gs_hdr = gs_hdr.n ? gs_hdr.p : nullptr;
    R.gs_diffs = gs_diffs.p;
    R.gs_dd = gs_dd.p;
    R.gs_dp = gs_dp.p;
    R.gs_imp = gs_imp.p;
    R.gs_wide = gs_wide ? 1 : 0;
    R.pad2_ = 0;
    R.q_hdr = q_hdr.n ? q_hdr.p : nullptr;
    R.q_depth = q_depth.p;
    R.q_cols = q_cols.p;
    R.q_groups = q_groups.p;
    R.q_rows = q_rows.p;
    R.seq_off = seq_off.p;
    R.seq = seq.p;
    R.qual = qual.p;
    R.qc = qc.p;
    R.md_off = md_off.p;
    R.md = md.p;
    R.nm = nm.p;
    R.del_start = del_start.p;
    R.del_md = del_md.p;
    R.del_len = del_len.p;
    R.contigs = d_contigs.p;
    R.trk_lo = trk_lo.p;
    R.trk_hi = trk_hi.p;
    R.trk_std = trk_std.p;
    R.fasta = fasta.n ? fasta.p : nullptr;
    R.fasta_off = fasta.n ? fasta_off.p : nullptr;
    R.gran_first = gran_first.p;
    R.gran_last = gran_last.p;
    return R;
  }
  uint64_t device_bytes() const {
    return rec.bytes() + cig_off.bytes() + cigar.bytes() + xmask.bytes() + md_off.bytes() + trk_lo.bytes() * 3 +
           gran_first.bytes() * 2 + pairs.bytes() + gs_hdr.bytes() + gs_diffs.bytes() + gs_dd.bytes() + gs_dp.bytes() + gs_imp.bytes() + q_hdr.bytes() + q_depth.bytes() + q_cols.bytes() + q_groups.bytes() + q_rows.bytes() + seq_off.bytes() + seq.bytes() + qual.bytes() + qc.bytes() + md.bytes() + nm.bytes() + del_start.bytes() + del_md.bytes() + del_len.bytes() +
           fasta.bytes();
  }
};

struct guac_result {
  int kind = 0;  // 0 threshold, 1 somatic, 2 counts, 3 called alleles (germline-standard), 4 allele counts
  // germline-threshold: compact single-base records + the exact kernel's general records + the allele byte pool in one pinned
  // block; the guac_threshold_record view is built on first use (expand_threshold_records)
  int32_t sample = 0;
  const guac_threshold_record* general = nullptr;
  size_t n_general = 0;
  const unsigned long long* compact = nullptr;
  size_t n_compact = 0;
  bool want_sorted = true, compact_sorted = false, expanded_ready = false;
  // the compact records as the call left them in HBM (guac_result_gather sends them from there); valid while the context
  // has not run another call (ctx->generation)
  const unsigned long long* d_compact = nullptr;
  guac_ctx* owner = nullptr;
  uint64_t generation = 0;
  std::vector<guac_threshold_record> expanded;
  // somatic / called-allele / allele-count records and the allele byte pool live in one pinned block (downloaded in place)
  std::shared_ptr<PinnedPool> pool;
  void* block = nullptr;
  size_t block_bytes = 0;
  size_t n_records = 0;
  const void* records = nullptr;
  const uint8_t* bytes = nullptr;
  size_t n_bytes = 0;
  std::vector<guac_locus_counts> counts;  // counts mode (rows for empty loci are appended on the host)
  guac_stats stats{};
  ~guac_result() {
    if (block && pool) pool->give(block, block_bytes);
  }
};

