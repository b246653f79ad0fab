// guac_bam.cuh — BAM file -> compact read batch (guac_read_batch_v2) on host threads.
//
// The reference reads BAM through htsjdk / hadoop-bam and Read.fromSAMRecord (reads/Read.scala:217-291, 368-451), one object per
// record, with Read.InputFilters (reads/Read.scala:88-136) applied afterwards.  Here the BGZF members (independent raw-deflate
// streams of at most 64 KB, SAM spec 4.1) are inflated in parallel, the records are located with one pointer walk, filtered
// and counted in parallel, and written in parallel straight into the columns guac_reads_pack_v2 copies to the device: the 4-bit
// bases and the CIGAR words are the BAM record's own bytes (shifted by a nibble where a read starts at an odd base offset).
// Host code only (zlib + std::thread); nothing here touches the GPU.
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <numeric>
#include <thread>

#include "guac_batch2.cuh"
#include "guac_inflate.h"

namespace {

struct BamRec {            // what pass A learns about one record
  uint64_t off;            // of the record body (behind block_size) in the inflated stream
  uint32_t md_off, md_len; // MD:Z value relative to `off` (md_len = 0 and md_off = 0: no tag)
  int32_t ref_id, pos;
  uint32_t l_seq, n_cigar;
  uint16_t flag;
  uint8_t mapq, keep;
  int32_t sample;          // index into the header's distinct @RG SM values, -1 = none ("default")
};

inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }

template <typename F>
void parallel_for(unsigned n_thr, uint64_t n, F&& f) {  // f(thread, begin, end) over contiguous slices
  n_thr = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(n_thr, n));
  std::vector<std::thread> pool;
  for (unsigned t = 1; t < n_thr; ++t) pool.emplace_back([&, t] { f(t, n * t / n_thr, n * (t + 1) / n_thr); });
  f(0u, (uint64_t)0, n / n_thr);
  for (std::thread& th : pool) th.join();
}

struct MappedFile {
  const uint8_t* p = nullptr;
  size_t n = 0;
  int fd = -1;
  ~MappedFile() {
    if (p && n) munmap(const_cast<uint8_t*>(p), n);
    if (fd >= 0) close(fd);
  }
};

void bam_load(const char* path, const guac_bam_options& opt, guac_host_batch_v2& H) {
  const auto t_begin = std::chrono::steady_clock::now();
  const bool trace = getenv("GUAC_TRACE") != nullptr;
  auto lap = [&, last = t_begin](const char* what) mutable {
    const auto now = std::chrono::steady_clock::now();
    if (trace) fprintf(stderr, "[guac bam] %-22s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
  };
  const unsigned n_thr = opt.n_threads > 0 ? (unsigned)opt.n_threads : std::max(1u, std::thread::hardware_concurrency());
  MappedFile file;
  file.fd = open(path, O_RDONLY);
  if (file.fd < 0) fail(GUAC_ERR_INVALID_ARGUMENT, "cannot open %s", path);
  struct stat sb;
  if (fstat(file.fd, &sb) != 0 || sb.st_size < 28) fail(GUAC_ERR_INVALID_ARGUMENT, "%s: not a BAM file", path);
  file.n = (size_t)sb.st_size;
  void* m = mmap(nullptr, file.n, PROT_READ, MAP_PRIVATE, file.fd, 0);
  if (m == MAP_FAILED) { file.n = 0; fail(GUAC_ERR_OOM, "mmap(%s) failed", path); }
  file.p = static_cast<const uint8_t*>(m);

  // ---- the BGZF members
  struct Member { uint64_t c_off; uint32_t c_len, u_len; uint64_t u_off; };
  std::vector<Member> members;
  uint64_t total = 0;
  for (uint64_t p = 0; p + 18 <= file.n;) {
    const uint8_t* h = file.p + p;
    if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) fail(GUAC_ERR_INVALID_ARGUMENT, "%s: not BGZF at byte %llu", path, (unsigned long long)p);
    const uint32_t xlen = rd16(h + 10);
    uint32_t bsize = 0;
    for (uint32_t x = 0; x + 4 <= xlen;) {
      const uint8_t* sf = h + 12 + x;
      const uint32_t slen = rd16(sf + 2);
      if (sf[0] == 'B' && sf[1] == 'C' && slen == 2) bsize = (uint32_t)rd16(sf + 4) + 1;
      x += 4 + slen;
    }
    if (bsize < 12 + xlen + 8 || p + bsize > file.n) fail(GUAC_ERR_INVALID_ARGUMENT, "%s: truncated BGZF member at byte %llu", path, (unsigned long long)p);
    Member mb;
    mb.c_off = p + 12 + xlen;
    mb.c_len = bsize - xlen - 20;
    mb.u_len = rd32(h + bsize - 4);
    mb.u_off = total;
    total += mb.u_len;
    if (mb.u_len) members.push_back(mb);
    p += bsize;
  }
  std::unique_ptr<uint8_t[]> raw(new uint8_t[total + 16]);
  memset(raw.get() + total, 0, 16);
  const uint8_t* D = raw.get();
  // The members are inflated by the worker threads in file order (one shared counter), and THIS thread follows right behind
  // them: it parses the header and walks the records (a chain of dependent 4-byte loads, one cache miss per record — serial by
  // nature) as soon as the member holding the next length word is done, so the walk hides underneath the inflate.
  std::unique_ptr<std::atomic<uint8_t>[]> done(new std::atomic<uint8_t>[members.size() + 1]);
  for (size_t k = 0; k <= members.size(); ++k) done[k].store(0, std::memory_order_relaxed);
  std::atomic<size_t> next{0}, n_fast{0};
  std::atomic<int> bad{0};
  const bool use_fast_inflate = getenv("GUAC_BAM_ZLIB_ONLY") == nullptr;  // (A/B and the fallback's test)
  std::vector<std::thread> inflaters;
  for (unsigned t = 0; t < n_thr; ++t)
    inflaters.emplace_back([&] {
      z_stream zs;
      memset(&zs, 0, sizeof zs);
      const bool ok = inflateInit2(&zs, -15) == Z_OK;
      std::unique_ptr<guac_inflate::Tables> tables(new guac_inflate::Tables);
      for (size_t k; (k = next.fetch_add(1)) < members.size();) {
        const Member& mb = members[k];
        // the member's own decoder first (guac_inflate.h: 2 - 3 x zlib on BGZF members); whatever it refuses goes through zlib
        if (use_fast_inflate && guac_inflate::inflate_member(file.p + mb.c_off, mb.c_len, raw.get() + mb.u_off, mb.u_len, *tables)) {
          n_fast.fetch_add(1, std::memory_order_relaxed);
        } else if (ok) {
          inflateReset(&zs);
          zs.next_in = const_cast<Bytef*>(file.p + mb.c_off);
          zs.avail_in = mb.c_len;
          zs.next_out = raw.get() + mb.u_off;
          zs.avail_out = mb.u_len;
          const int rc = inflate(&zs, Z_FINISH);
          if (rc != Z_STREAM_END || zs.avail_out != 0) bad = 1;
        } else {
          bad = 1;
        }
        done[k].store(1, std::memory_order_release);
      }
      if (ok) inflateEnd(&zs);
    });
  struct JoinAll {
    std::vector<std::thread>& v;
    ~JoinAll() { for (std::thread& th : v) if (th.joinable()) th.join(); }
  } join_all{inflaters};
  size_t ready_member = 0;   // members [0, ready_member) are known to be inflated
  uint64_t ready_bytes = 0;  // ... which is the inflated stream up to here
  auto need = [&](uint64_t upto) {  // block until bytes [0, upto) of the inflated stream are there (false: past the end)
    if (upto > total) return false;
    while (ready_bytes < upto) {
      while (!done[ready_member].load(std::memory_order_acquire)) std::this_thread::yield();
      ready_bytes = members[ready_member].u_off + members[ready_member].u_len;
      ++ready_member;
    }
    return true;
  };

  // ---- header: text (@RG ID -> SM), reference names and lengths
  if (!need(12) || memcmp(D, "BAM\1", 4) != 0) fail(GUAC_ERR_INVALID_ARGUMENT, "%s: no BAM magic", path);
  uint64_t p = 4;
  const uint32_t l_text = rd32(D + p);
  p += 4;
  if (!need(p + l_text + 4)) fail(GUAC_ERR_INVALID_ARGUMENT, "%s: truncated header", path);
  std::string text(reinterpret_cast<const char*>(D + p), strnlen(reinterpret_cast<const char*>(D + p), l_text));
  p += l_text;
  const uint32_t n_ref = rd32(D + p);
  p += 4;
  for (uint32_t r = 0; r < n_ref; ++r) {
    if (!need(p + 4)) fail(GUAC_ERR_INVALID_ARGUMENT, "%s: truncated reference list", path);
    const uint32_t l_name = rd32(D + p);
    p += 4;
    if (l_name == 0 || !need(p + l_name + 4)) fail(GUAC_ERR_INVALID_ARGUMENT, "%s: truncated reference list", path);
    H.contig_names.emplace_back(reinterpret_cast<const char*>(D + p), l_name - 1);
    p += l_name;
    H.contig_length.push_back((int64_t)(int32_t)rd32(D + p));
    p += 4;
  }
  std::vector<std::pair<std::string, int>> rg_sample;  // read group id -> sample index
  std::vector<std::string> samples;
  for (size_t a = 0; a < text.size();) {
    size_t e = text.find('\n', a);
    if (e == std::string::npos) e = text.size();
    if (text.compare(a, 3, "@RG") == 0) {
      std::string id, sm;
      for (size_t f = a; f < e;) {
        size_t g = text.find('\t', f);
        if (g == std::string::npos || g > e) g = e;
        if (g - f > 3 && text.compare(f, 3, "ID:") == 0) id = text.substr(f + 3, g - f - 3);
        if (g - f > 3 && text.compare(f, 3, "SM:") == 0) sm = text.substr(f + 3, g - f - 3);
        f = g + 1;
      }
      while (!sm.empty() && (sm.back() == '\r')) sm.pop_back();
      if (!id.empty() && !sm.empty()) {
        auto it = std::find(samples.begin(), samples.end(), sm);
        if (it == samples.end()) { samples.push_back(sm); it = samples.end() - 1; }
        rg_sample.emplace_back(id, (int)(it - samples.begin()));
      }
    }
    a = e + 1;
  }

  // ---- record offsets: one pointer walk, right behind the inflaters
  std::vector<BamRec> recs;
  recs.reserve((size_t)((total - p) / 200 + 16));
  while (p + 4 <= total) {
    need(p + 4);
    const uint32_t bs = rd32(D + p);
    if (bs < 32 || p + 4 + bs > total) fail(GUAC_ERR_INVALID_ARGUMENT, "%s: truncated record at inflated byte %llu", path, (unsigned long long)p);
    BamRec r{};
    r.off = p + 4;
    r.md_len = bs;  // (pass A replaces it; until then: the record's size)
    recs.push_back(r);
    p += 4 + bs;
  }
  for (std::thread& th : inflaters) th.join();
  if (bad) fail(GUAC_ERR_INVALID_ARGUMENT, "%s: corrupt BGZF member", path);
  const uint64_t n_rec = recs.size();
  lap("inflate + record walk");
  if (trace) fprintf(stderr, "[guac bam] members %zu, of which %zu through guac_inflate\n", members.size(), n_fast.load());

  // ---- pass A: fields, tags, filters
  std::atomic<int> bad_record{0};
  parallel_for(n_thr, n_rec, [&](unsigned, uint64_t b, uint64_t e) {
    for (uint64_t i = b; i < e; ++i) {
      BamRec& r = recs[i];
      const uint8_t* R = D + r.off;
      const uint32_t size = r.md_len;
      r.md_len = 0;
      r.ref_id = (int32_t)rd32(R);
      r.pos = (int32_t)rd32(R + 4);
      const uint32_t l_name = R[8];
      r.mapq = R[9];
      r.n_cigar = rd16(R + 12);
      r.flag = rd16(R + 14);
      r.l_seq = rd32(R + 16);
      r.sample = -1;
      uint64_t q = 32ull + l_name + 4ull * r.n_cigar + (r.l_seq + 1) / 2 + r.l_seq;
      if (q > size) { bad_record = 1; r.keep = 0; continue; }
      while (q + 3 <= size) {  // auxiliary fields
        const uint8_t t0 = R[q], t1 = R[q + 1], ty = R[q + 2];
        q += 3;
        if (ty == 'Z' || ty == 'H') {
          const void* z = memchr(R + q, 0, size - q);
          if (!z) { bad_record = 1; break; }
          const uint32_t len = (uint32_t)(static_cast<const uint8_t*>(z) - (R + q));
          if (t0 == 'M' && t1 == 'D' && ty == 'Z') { r.md_off = (uint32_t)q; r.md_len = len; if (!len) r.md_off = 0xFFFFFFFFu; }
          if (t0 == 'R' && t1 == 'G' && ty == 'Z') {
            for (const auto& kv : rg_sample)
              if (kv.first.size() == len && memcmp(kv.first.data(), R + q, len) == 0) { r.sample = kv.second; break; }
          }
          q += len + 1;
        } else if (ty == 'A' || ty == 'c' || ty == 'C') q += 1;
        else if (ty == 's' || ty == 'S') q += 2;
        else if (ty == 'i' || ty == 'I' || ty == 'f') q += 4;
        else if (ty == 'B') {
          if (q + 5 > size) { bad_record = 1; break; }
          const uint8_t sub = R[q];
          const uint32_t cnt = rd32(R + q + 1);
          const uint32_t w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
          q += 5 + (uint64_t)cnt * w;
        } else { bad_record = 1; break; }
      }
      const bool has_md = r.md_len != 0 || r.md_off == 0xFFFFFFFFu;
      bool keep = !(r.flag & 0x4) && r.ref_id >= 0 && (uint32_t)r.ref_id < n_ref && r.pos >= 0 && r.n_cigar > 0;  // mapped reads only
      if (opt.non_duplicate && (r.flag & 0x400)) keep = false;
      if (opt.passed_qc && (r.flag & 0x200)) keep = false;
      if (opt.has_md_tag && !has_md) keep = false;
      if (opt.is_paired && !(r.flag & 0x1)) keep = false;
      r.keep = keep ? (has_md ? 3 : 1) : 0;
    }
  });
  if (bad_record) fail(GUAC_ERR_INVALID_ARGUMENT, "%s: malformed record", path);
  lap("pass A (fields, tags)");

  // ---- the sample: reads of one sample per batch (the callers pack per sample); `opt.sample` selects one
  int want_sample = -2;  // -2: whatever the kept reads carry, if it is one
  if (opt.sample) {
    auto it = std::find(samples.begin(), samples.end(), std::string(opt.sample));
    want_sample = it == samples.end() ? (std::string(opt.sample) == "default" ? -1 : -3) : (int)(it - samples.begin());
  }
  // Selection and offsets, slice by slice on all threads (a serial pass over the records is a cache miss per record: a third of
  // a second for a chr20 at 30x): per slice the kept records, their sizes, whether they are in (contig, start) order and of one
  // sample; the slices' sums are scanned, then every slice writes its part of `kept` and of the offset columns.
  if (n_rec >= 0xFFFFFFFFull) fail(GUAC_ERR_UNSUPPORTED, "%s: more than 2^32 records", path);
  struct Slice {
    uint64_t kept = 0, cig = 0, seq = 0, md = 0;
    int sample = -2;          // of its kept reads (-2: none kept)
    bool mixed = false, sorted = true;
    int32_t first_ref = 0, first_pos = 0, last_ref = 0, last_pos = 0;
  };
  const unsigned n_sl = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(n_thr, n_rec));
  std::vector<Slice> sl(n_sl);
  parallel_for(n_thr, n_rec, [&](unsigned t, uint64_t b, uint64_t e) {
    Slice z;
    for (uint64_t i = b; i < e; ++i) {
      BamRec& r = recs[i];
      if (!r.keep) continue;
      if (want_sample != -2 && r.sample != want_sample) { r.keep = 0; continue; }
      if (z.sample == -2) z.sample = r.sample;
      else if (z.sample != r.sample) z.mixed = true;
      if (z.kept == 0) { z.first_ref = r.ref_id; z.first_pos = r.pos; }
      else if (z.last_ref > r.ref_id || (z.last_ref == r.ref_id && z.last_pos > r.pos)) z.sorted = false;
      z.last_ref = r.ref_id;
      z.last_pos = r.pos;
      z.kept += 1;
      z.cig += r.n_cigar;
      z.seq += r.l_seq;
      z.md += r.md_len;
    }
    sl[t] = z;
  });
  int seen_sample = -2;
  bool sorted = true;
  std::vector<uint64_t> base_kept(n_sl + 1, 0), base_cig(n_sl + 1, 0), base_seq(n_sl + 1, 0), base_md(n_sl + 1, 0);
  {
    const Slice* before = nullptr;
    for (unsigned t = 0; t < n_sl; ++t) {
      const Slice& z = sl[t];
      base_kept[t + 1] = base_kept[t] + z.kept;
      base_cig[t + 1] = base_cig[t] + z.cig;
      base_seq[t + 1] = base_seq[t] + z.seq;
      base_md[t + 1] = base_md[t] + z.md;
      if (!z.kept) continue;
      if (z.mixed || (seen_sample != -2 && seen_sample != z.sample))
        fail(GUAC_ERR_INVALID_ARGUMENT, "%s holds reads of several samples: name one in guac_bam_options.sample", path);
      seen_sample = z.sample;
      sorted = sorted && z.sorted;
      if (before && (before->last_ref > z.first_ref || (before->last_ref == z.first_ref && before->last_pos > z.first_pos))) sorted = false;
      before = &z;
    }
  }
  const uint64_t n = base_kept[n_sl];
  if (base_cig[n_sl] >= 0xFFFFFFFFull || base_seq[n_sl] >= 0xFFFFFFFFull || base_md[n_sl] >= 0xFFFFFFFFull)
    fail(GUAC_ERR_UNSUPPORTED, "%s: the compact batch holds fewer than 2^32 bases, CIGAR ops and MD bytes: load it by region / sample", path);
  std::vector<uint32_t> kept(n);
  std::vector<uint64_t> o_cig(n + 1, 0), o_seq(n + 1, 0), o_md(n + 1, 0);
  parallel_for(n_thr, n_rec, [&](unsigned t, uint64_t b, uint64_t e) {
    uint64_t k = base_kept[t], c = base_cig[t], q = base_seq[t], m = base_md[t];
    for (uint64_t i = b; i < e; ++i) {
      const BamRec& r = recs[i];
      if (!r.keep) continue;
      kept[k] = (uint32_t)i;
      o_cig[k] = c;
      o_seq[k] = q;
      o_md[k] = m;
      c += r.n_cigar;
      q += r.l_seq;
      m += r.md_len;
      ++k;
    }
  });
  o_cig[n] = base_cig[n_sl];
  o_seq[n] = base_seq[n_sl];
  o_md[n] = base_md[n_sl];
  if (!sorted) {  // (contig index, start), file order among equals: what sortBy(start) inside a task gives the reference
    std::stable_sort(kept.begin(), kept.end(), [&](uint32_t a, uint32_t b) {
      return recs[a].ref_id != recs[b].ref_id ? recs[a].ref_id < recs[b].ref_id : recs[a].pos < recs[b].pos;
    });
    for (uint64_t k = 0; k < n; ++k) {  // (the offsets follow the new order)
      const BamRec& r = recs[kept[k]];
      o_cig[k + 1] = o_cig[k] + r.n_cigar;
      o_seq[k + 1] = o_seq[k] + r.l_seq;
      o_md[k + 1] = o_md[k] + r.md_len;
    }
  }
  H.sample_name = seen_sample >= 0 ? samples[(size_t)seen_sample] : std::string("default");
  const uint64_t n_bases = o_seq[n];
  lap("select + offsets");
  bool fixed = n > 0;
  const uint32_t L0 = n ? recs[kept[0]].l_seq : 0;
  for (uint64_t k = 0; fixed && k < n; ++k) fixed = recs[kept[k]].l_seq == L0;
  fixed = fixed && L0 > 0;
  const bool with_qual = opt.with_qualities != 0;

  guac_read_batch_v2& V = H.view;
  V.n_reads = n;
  V.n_contigs = n_ref;
  V.contig_length = H.contig_length.data();
  H.contig_read_off.assign((size_t)n_ref + 1, n);
  {
    int64_t prev = -1;
    for (uint64_t k = 0; k < n; ++k) {
      const int64_t c = recs[kept[k]].ref_id;
      for (int64_t j = prev + 1; j <= c; ++j) H.contig_read_off[(size_t)j] = k;
      prev = std::max(prev, c);
    }
    if (n_ref) H.contig_read_off[0] = 0;
  }
  V.contig_read_off = H.contig_read_off.data();
  int32_t* start = H.take<int32_t>(n);
  uint32_t* cigar_off = H.take<uint32_t>(n + 1);
  uint32_t* md_off = H.take<uint32_t>(n + 1);
  uint32_t* seq_off = fixed ? nullptr : H.take<uint32_t>(n + 1);
  uint32_t* cigar = H.take<uint32_t>((size_t)o_cig[n]);
  uint8_t* seq4 = H.take<uint8_t>((size_t)(n_bases + 1) / 2 + 16);
  uint8_t* qual = with_qual ? H.take<uint8_t>((size_t)n_bases) : nullptr;
  uint8_t* mapq = H.take<uint8_t>(n);
  uint8_t* flags = H.take<uint8_t>(n);
  char* md = H.take<char>((size_t)o_md[n]);
  memset(seq4 + (n_bases + 1) / 2, 0, 16);
  cigar_off[n] = (uint32_t)o_cig[n];
  md_off[n] = (uint32_t)o_md[n];
  if (seq_off) seq_off[n] = (uint32_t)n_bases;

  // ---- pass B: the columns.  A byte of seq4 is written by the read that owns its HIGH nibble; when that read ends there, the
  // low nibble is the first base of the next read that has bases (pulled from that read's record).
  parallel_for(n_thr, n, [&](unsigned, uint64_t b, uint64_t e) {
    for (uint64_t k = b; k < e; ++k) {
      const BamRec& r = recs[kept[k]];
      const uint8_t* R = D + r.off;
      start[k] = r.pos;
      cigar_off[k] = (uint32_t)o_cig[k];
      md_off[k] = (uint32_t)o_md[k];
      if (seq_off) seq_off[k] = (uint32_t)o_seq[k];
      mapq[k] = r.mapq;
      flags[k] = (uint8_t)(((r.flag & 0x10) ? 0u : GUAC_READ_POSITIVE_STRAND) | ((r.flag & 0x400) ? GUAC_READ_DUPLICATE : 0u) |
                           ((r.flag & 0x200) ? GUAC_READ_FAILED_QC : 0u) | ((r.keep & 2) ? GUAC_READ_HAS_MD : 0u) | ((r.flag & 0x1) ? GUAC_READ_PAIRED : 0u));
      const uint8_t* C = R + 32 + R[8];
      memcpy(cigar + o_cig[k], C, 4ull * r.n_cigar);
      if (r.md_len) memcpy(md + o_md[k], R + r.md_off, r.md_len);
      const uint8_t* S = C + 4ull * r.n_cigar;
      const uint32_t L = r.l_seq;
      if (qual && L) {
        const uint8_t* Q = S + (L + 1) / 2;
        if (Q[0] == 0xFF) memset(qual + o_seq[k], 0, L);  // qualities absent
        else memcpy(qual + o_seq[k], Q, L);
      }
      if (!L) continue;
      const uint64_t o = o_seq[k];
      auto next_first = [&]() -> uint32_t {  // first base of the next read that has one
        for (uint64_t j = k + 1; j < n; ++j) {
          const BamRec& x = recs[kept[j]];
          if (x.l_seq) return (uint32_t)(D + x.off)[32 + (D + x.off)[8] + 4ull * x.n_cigar] >> 4;
        }
        return 0u;
      };
      if (!(o & 1)) {
        memcpy(seq4 + o / 2, S, L / 2);
        if (L & 1) seq4[o / 2 + L / 2] = (uint8_t)((S[L / 2] & 0xF0u) | next_first());
      } else {  // the first base sits in the previous read's last byte; bases 1.. start on a byte boundary
        uint8_t* dst = seq4 + (o + 1) / 2;
        const uint32_t M = L - 1;
        for (uint32_t t = 0; t < M / 2; ++t) dst[t] = (uint8_t)((S[t] << 4) | (S[t + 1] >> 4));
        if (M & 1) {  // base L-1 is a high nibble; L is even here, so it is the LOW nibble of S[(L-1)/2]
          dst[M / 2] = (uint8_t)(((S[(L - 1) / 2] & 0x0Fu) << 4) | next_first());
        }
      }
    }
  });
  lap("alloc + pass B (columns)");
  V.read_length = fixed ? L0 : 0u;
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
  V.sample = 0;
  H.bytes = n * (4 + 4 + 4 + (fixed ? 0 : 4) + 1 + 1) + o_cig[n] * 4 + (n_bases + 1) / 2 + (qual ? n_bases : 0) + o_md[n];
  H.file_bytes = file.n;
  H.inflated_bytes = total;
  H.records_in_file = n_rec;
  H.decode_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
}

}  // namespace
