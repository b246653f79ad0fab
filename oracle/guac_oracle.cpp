// guac_oracle.cpp — CPU ORACLE: a literal restatement of the reference's pileup-and-call path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under guacamole_b200/ (the product) may link, load or call this file; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / "--impl reference" legs do.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/src/main/scala/org/hammerlab/guacamole/).  It deliberately keeps the reference's algorithmic
// SHAPE (sliding window with a binary heap, one cursor object per (read, locus), per-locus allele grouping, dense
// alleles x depth probability matrix) so that it doubles as the CPU baseline of bench.py.
//
// Third-party behaviour that is not under /root/reference is restated from the libraries' published algorithms:
//   ADAM 0.18.1  MdTag / PhredUtils        (adam-core_2.10, pom.xml:19,277-286)
//   colt 1.2.0   DoubleMatrix1D.aggregate  (sums from the LAST element to the first)
//   breeze 0.11.2 mean / median
//   scala-library 2.10.3 mutable.PriorityQueue (1-based array heap; iteration = array order)
// Parity pinning: the golden vectors of PileupSuite, MDTagUtilsSuite, SlidingWindowSuite, DistributedUtilSuite,
// GermlineThresholdCallerSuite, LikelihoodSuite, SomaticStandardCallerSuite, AlleleEvidenceSuite and
// VariantSupportSuite are replayed by tests/test_oracle_*.py.  End-to-end output on chrM.sorted.bam is NOT pinned
// by the reference (it ships no golden VCF): "parity unpinned" for that config beyond the unit vectors.

#include "guac_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_last_error;

struct OrcError {
  int code;
  std::string msg;
};

[[noreturn]] void fail(int code, const std::string& msg) { throw OrcError{code, msg}; }

// ---- htsjdk CigarOperator (CigarUtils.scala:30-42) ---------------------------------------------------------
inline bool consumes_read(int op) {
  return op == GUAC_CIGAR_M || op == GUAC_CIGAR_I || op == GUAC_CIGAR_S || op == GUAC_CIGAR_EQ || op == GUAC_CIGAR_X;
}
inline bool consumes_ref(int op) {
  return op == GUAC_CIGAR_M || op == GUAC_CIGAR_D || op == GUAC_CIGAR_N || op == GUAC_CIGAR_EQ || op == GUAC_CIGAR_X;
}
inline bool is_standard_base(uint8_t b) { return b == 'A' || b == 'C' || b == 'G' || b == 'T'; }  // Bases.scala:66-68

struct CigarElem {
  int op;
  int len;
};

// ---- ADAM PhredUtils (third party; restated) -----------------------------------------------------------------
struct PhredTables {
  double error[256], success[256];
  PhredTables() {
    for (int p = 0; p < 256; ++p) {
      error[p] = std::pow(10.0, -p / 10.0);
      success[p] = 1.0 - error[p];
    }
  }
};
const PhredTables& phred_tables() {
  static PhredTables t;
  return t;
}
double phred_to_success(int q) {
  if (q < 0 || q > 255) fail(GUAC_ERR_BAD_QUALITY, "phred score outside PhredUtils' 0..255 table");
  return phred_tables().success[q];
}
// scala.math.round(Double): Long == java.lang.Math.round: floor(x + 0.5), NaN -> 0, saturating; then Long.toInt.
int java_round_to_int(double x) {
  long long r;
  if (std::isnan(x))
    r = 0;
  else {
    double f = std::floor(x + 0.5);
    if (f >= 9.2233720368547758e18)
      r = std::numeric_limits<long long>::max();
    else if (f <= -9.2233720368547758e18)
      r = std::numeric_limits<long long>::min();
    else
      r = (long long)f;
  }
  return (int)(unsigned int)(unsigned long long)r;  // Long.toInt keeps the low 32 bits
}
int success_probability_to_phred(double p) { return java_round_to_int(-10.0 * std::log10(1.0 - p)); }

// ---- MappedRead (reads/MappedRead.scala:35-101) ------------------------------------------------------------
struct Read {
  int64_t idx = 0;
  int32_t contig = 0;
  int64_t start = 0, end = 0;
  std::vector<CigarElem> cigar;
  const uint8_t* seq = nullptr;
  const uint8_t* qual = nullptr;
  int len = 0;
  int mapq = 0;
  bool positive = true;
  bool has_md = false;
  int32_t sample = 0;
  std::string md;
  // ADAM MdTag, parsed lazily
  mutable bool md_parsed = false;
  mutable std::unordered_map<int64_t, char> mismatches, deletions;
  mutable bool md_ref_built = false;
  mutable std::vector<uint8_t> md_ref;

  bool overlaps_locus(int64_t locus) const { return start <= locus && end > locus; }  // HasReferenceRegion.scala:51-53

  // ADAM MdTag(mdString, referenceStart, cigar): CIGAR-driven walk of the MD tokens.
  void parse_md() const {
    if (md_parsed) return;
    md_parsed = true;
    if (!has_md) fail(GUAC_ERR_MISSING_MD, "read " + std::to_string(idx) + " has no MD tag");  // mdTagOpt.get
    if (md == "0" || md.empty()) return;
    std::string up = md;
    for (auto& c : up) c = (char)std::toupper((unsigned char)c);
    size_t pos = 0;
    int64_t ref_pos = start;
    int64_t pending = 0;  // matching bases announced by the last number and not yet used
    auto bad = [&](const char* why) {
      fail(GUAC_ERR_MISSING_MD, std::string("MD tag '") + md + "' inconsistent with CIGAR of read " + std::to_string(idx) + ": " + why);
    };
    auto take_number = [&]() {
      int64_t n = 0;
      while (pos < up.size() && std::isdigit((unsigned char)up[pos])) n = n * 10 + (up[pos++] - '0');
      return n;
    };
    if (!std::isdigit((unsigned char)up[0])) bad("does not start with a digit");
    for (const auto& ce : cigar) {
      if (ce.op == GUAC_CIGAR_M || ce.op == GUAC_CIGAR_EQ || ce.op == GUAC_CIGAR_X) {
        int64_t remaining = ce.len;
        while (remaining > 0) {
          if (pending > 0) {
            int64_t k = std::min(remaining, pending);
            ref_pos += k;
            remaining -= k;
            pending -= k;
          } else if (pos >= up.size()) {
            bad("tag ends before the alignment does");
          } else if (std::isdigit((unsigned char)up[pos])) {
            pending = take_number();
          } else if (up[pos] == '^') {
            bad("deletion marker inside a match element");
          } else {
            mismatches[ref_pos] = up[pos++];
            ++ref_pos;
            --remaining;
          }
        }
      } else if (ce.op == GUAC_CIGAR_D) {
        int64_t remaining = ce.len;
        bool seen_caret = false;
        while (remaining > 0) {
          if (pending > 0) bad("found matching bases in deletion");
          if (pos >= up.size()) bad("tag ends inside a deletion");
          if (std::isdigit((unsigned char)up[pos])) {
            pending = take_number();  // a "0" separator is skipped
          } else if (up[pos] == '^') {
            ++pos;
            seen_caret = true;
          } else {
            (void)seen_caret;
            deletions[ref_pos] = up[pos++];
            ++ref_pos;
            --remaining;
          }
        }
      } else if (ce.op == GUAC_CIGAR_N) {
        ref_pos += ce.len;
      }
    }
  }

  int count_of_mismatches() const {
    parse_md();
    return (int)mismatches.size();
  }

  // MDTagUtils.getReference(mdTag, seq, cigar, allowNBase = true)  reads/MDTagUtils.scala:23-78
  const std::vector<uint8_t>& md_reference() const {
    if (md_ref_built) return md_ref;
    if (!has_md) fail(GUAC_ERR_MISSING_MD, "ReferenceWithoutMDTagException: read " + std::to_string(idx));  // MappedRead.scala:58-59
    parse_md();
    int64_t ref_pos = start;
    int read_pos = 0;
    std::vector<uint8_t> out;
    for (const auto& ce : cigar) {
      if (ce.op == GUAC_CIGAR_M || ce.op == GUAC_CIGAR_EQ || ce.op == GUAC_CIGAR_X) {
        for (int i = 0; i < ce.len; ++i) {
          auto it = mismatches.find(ref_pos);
          if (it != mismatches.end())
            out.push_back((uint8_t)it->second);
          else {
            if (read_pos >= len) fail(GUAC_ERR_INVALID_CIGAR, "CIGAR consumes more bases than the read has");
            out.push_back(seq[read_pos]);
          }
          ++read_pos;
          ++ref_pos;
        }
      } else if (ce.op == GUAC_CIGAR_N) {
        ref_pos += ce.len;
        out.insert(out.end(), (size_t)ce.len, (uint8_t)'N');
      } else if (ce.op == GUAC_CIGAR_D) {
        for (int i = 0; i < ce.len; ++i) {
          auto it = deletions.find(ref_pos);
          if (it == deletions.end())
            fail(GUAC_ERR_MISSING_MD, "CigarMDTagMismatchException: could not find deleted base at cigar offset " + std::to_string(i));
          out.push_back((uint8_t)it->second);
          ++ref_pos;
        }
      } else {
        if (consumes_read(ce.op)) read_pos += ce.len;
        if (consumes_ref(ce.op)) fail(GUAC_ERR_INVALID_CIGAR, "Cannot handle operator");
      }
    }
    md_ref = std::move(out);
    md_ref_built = true;
    return md_ref;
  }

  uint8_t reference_base_at(int64_t locus) const {  // MappedRead.getReferenceBaseAtLocus :73-76
    if (!(locus >= start && locus < end)) fail(GUAC_ERR_INVALID_ARGUMENT, "assumption failed: locus outside read");
    const auto& r = md_reference();
    size_t i = (size_t)(locus - start);
    if (i >= r.size()) fail(GUAC_ERR_INVALID_CIGAR, "MD-derived reference shorter than the alignment");
    return r[i];
  }
};

struct ReadSet {
  std::vector<Read> reads;
  uint32_t n_contigs = 0;
};

void load_batch(const guac_read_batch* b, ReadSet& rs) {
  if (!b) fail(GUAC_ERR_INVALID_ARGUMENT, "null batch");
  rs.n_contigs = b->n_contigs;
  rs.reads.resize(b->n_reads);
  for (uint64_t i = 0; i < b->n_reads; ++i) {
    Read& r = rs.reads[i];
    r.idx = (int64_t)i;
    r.contig = b->contig[i];
    r.start = b->start[i];
    int64_t ref_len = 0;
    for (uint64_t k = b->cigar_off[i]; k < b->cigar_off[i + 1]; ++k) {
      int op = (int)(b->cigar[k] & 0xF), len = (int)(b->cigar[k] >> 4);
      if (op > 8) fail(GUAC_ERR_INVALID_CIGAR, "unknown CIGAR operator");
      if (op == GUAC_CIGAR_P) fail(GUAC_ERR_INVALID_CIGAR, "P operator is not supported (SURVEY 8a quirks)");
      r.cigar.push_back({op, len});
      if (consumes_ref(op)) ref_len += len;  // Cigar.getPaddedReferenceLength without P
    }
    r.end = r.start + ref_len;  // MappedRead.scala:87
    r.seq = b->seq + b->seq_off[i];
    r.qual = b->qual + b->seq_off[i];
    r.len = (int)(b->seq_off[i + 1] - b->seq_off[i]);
    r.mapq = b->mapq[i];
    r.positive = (b->flags[i] & GUAC_READ_POSITIVE_STRAND) != 0;
    r.has_md = (b->flags[i] & GUAC_READ_HAS_MD) != 0;
    r.sample = b->sample ? b->sample[i] : 0;
    if (b->md_off) r.md.assign(b->md + b->md_off[i], b->md + b->md_off[i + 1]);
    for (int k = 0; k < r.len; ++k)
      if (r.qual[k] > 127) fail(GUAC_ERR_BAD_QUALITY, "base quality > 127");
  }
}

// ---- Allele (variants/Allele.scala:26-37, Bases.scala:45-49) ----------------------------------------------------
struct Allele {
  std::string ref, alt;
  bool is_variant() const { return ref != alt; }
  bool operator==(const Allele& o) const { return ref == o.ref && alt == o.alt; }
};
// java.lang.String.compareTo over chars made by Byte.toChar (sign-extending)
int bases_compare(const std::string& x, const std::string& y) {
  size_t n = std::min(x.size(), y.size());
  for (size_t i = 0; i < n; ++i) {
    int a = (uint16_t)(int16_t)(int8_t)x[i], b = (uint16_t)(int16_t)(int8_t)y[i];
    if (a != b) return a - b;
  }
  return (int)x.size() - (int)y.size();
}
int allele_compare(const Allele& a, const Allele& b) {
  int c = bases_compare(a.ref, b.ref);
  return c != 0 ? c : bases_compare(a.alt, b.alt);
}
bool allele_less(const Allele& a, const Allele& b) { return allele_compare(a, b) < 0; }

// ---- PileupElement (pileup/PileupElement.scala) ------------------------------------------------------------------
struct Element {
  const Read* read = nullptr;
  int64_t locus = 0;
  uint8_t reference_base = 'N';
  int read_position = 0;
  int cigar_index = 0;
  int64_t cigar_locus = 0;
  int index_within = 0;
  // lazy val alignment
  int kind = -1;
  std::string ref_bases, seq_bases;
  int quality = 0;

  const CigarElem& cigar_element() const {
    if (cigar_index >= (int)read->cigar.size()) fail(GUAC_ERR_INVALID_CIGAR, "advanced past the last cigar element");
    return read->cigar[cigar_index];
  }

  // PileupElement.advanceToLocus :220-248 (+ advanceToNextCigarElement :176-198, currentCigarElementContainsLocus :205-207)
  void advance_to_locus(int64_t new_locus, uint8_t new_reference_base) {
    if (!(new_locus >= locus)) fail(GUAC_ERR_INVALID_ARGUMENT, "assumption failed: Pileups only advance");
    if (!(new_locus < read->end)) fail(GUAC_ERR_INVALID_ARGUMENT, "assumption failed: can't advance past the end of the read");
    for (;;) {
      const CigarElem& ce = cigar_element();
      int64_t ref_len = consumes_ref(ce.op) ? ce.len : 0;
      if (cigar_locus <= new_locus && new_locus < cigar_locus + ref_len) {
        int off = consumes_read(ce.op) ? (int)(new_locus - cigar_locus - index_within) : 0;
        locus = new_locus;
        reference_base = new_reference_base;
        read_position += off;
        index_within = (int)(new_locus - cigar_locus);
        kind = -1;
        return;
      } else if (new_locus == 0 && ce.op == GUAC_CIGAR_I) {
        return;  // insertion at the start of a contig: stay on the I element
      } else {
        int rp_off = consumes_read(ce.op) ? ce.len - index_within : 0;
        locus = locus + (ref_len - index_within);
        reference_base = 'N';
        read_position += rp_off;
        cigar_index += 1;
        cigar_locus += ref_len;
        index_within = 0;
        kind = -1;
      }
    }
  }

  static Element create(const Read* r, int64_t locus, uint8_t reference_base) {  // PileupElement.apply :264-274
    Element e;
    e.read = r;
    e.locus = r->start;
    e.reference_base = 'N';
    e.cigar_locus = r->start;
    if (!(locus >= r->start && locus < r->end)) fail(GUAC_ERR_INVALID_ARGUMENT, "assumption failed: locus outside read");
    e.advance_to_locus(locus, reference_base);
    return e;
  }

  // lazy val alignment :68-135, allele bases Alignment.scala:44-94, qualityScore :166-171
  void evaluate() {
    if (kind >= 0) return;
    const CigarElem& ce = cigar_element();
    const Read& r = *read;
    bool is_final = index_within == ce.len - 1;
    int next_op = -1;  // None
    const CigarElem* next_elem = (cigar_index + 1 < (int)r.cigar.size()) ? &r.cigar[cigar_index + 1] : nullptr;
    if (is_final) {
      if (next_elem) next_op = next_elem->op;
    } else {
      next_op = ce.op;
    }
    auto make_insertion = [&](const CigarElem& ins) {
      int from = std::min(std::max(read_position, 0), r.len);
      int until = std::min(read_position + (consumes_read(ins.op) ? ins.len : 0) + 1, r.len);
      if (until < from) until = from;
      seq_bases.assign((const char*)r.seq + from, (const char*)r.seq + until);
      ref_bases = seq_bases.empty() ? std::string() : std::string(1, seq_bases[0]);
      if (from == until) fail(GUAC_ERR_INVALID_CIGAR, "empty insertion (min of empty quality sequence)");
      int q = 255;
      for (int k = from; k < until; ++k) q = std::min(q, (int)(int8_t)r.qual[k]);
      quality = q;
      kind = ORC_INSERTION;
    };
    int op = ce.op;
    bool m_like = (op == GUAC_CIGAR_M || op == GUAC_CIGAR_EQ || op == GUAC_CIGAR_X);
    if ((op == GUAC_CIGAR_M || op == GUAC_CIGAR_EQ) && next_op == GUAC_CIGAR_I) {
      make_insertion(*next_elem);
    } else if (op == GUAC_CIGAR_I && next_op >= 0 && cigar_locus == 0) {
      make_insertion(ce);
    } else if (op == GUAC_CIGAR_I) {
      fail(GUAC_ERR_INVALID_CIGAR, "InvalidCigarElementException: PileupElement at non-reference-consuming cigar-operator I");
    } else if (m_like && next_op == GUAC_CIGAR_D) {
      r.parse_md();
      int64_t ref_string_idx = (cigar_locus - r.start) + index_within;
      std::string deleted(1, (char)reference_base);
      for (int64_t off = ref_string_idx + 1; off < ref_string_idx + 1 + next_elem->len; ++off) {
        auto it = r.deletions.find(r.start + off);
        if (it == r.deletions.end()) fail(GUAC_ERR_MISSING_MD, "NoSuchElementException: MD tag has no deleted base at the locus");
        deleted.push_back(it->second);
      }
      if (read_position < 0 || read_position >= r.len) fail(GUAC_ERR_INVALID_CIGAR, "read position outside the read");
      ref_bases = deleted;
      seq_bases = std::string(1, deleted[0]);
      quality = (int8_t)r.qual[read_position];
      kind = ORC_DELETION;
    } else if (op == GUAC_CIGAR_D) {
      r.parse_md();
      auto it = r.deletions.find(locus);
      if (it == r.deletions.end()) fail(GUAC_ERR_MISSING_MD, "NoSuchElementException: MD tag has no deleted base at the locus");
      ref_bases = std::string(1, it->second);
      seq_bases.clear();
      quality = r.mapq;
      kind = ORC_MID_DELETION;
    } else if (next_op == GUAC_CIGAR_D) {
      fail(GUAC_ERR_INVALID_CIGAR, "AssertionError: found deletion preceded by a non-match cigar operator");
    } else if (m_like) {
      if (read_position < 0 || read_position >= r.len) fail(GUAC_ERR_INVALID_CIGAR, "read position outside the read");
      uint8_t base = r.seq[read_position];
      quality = (int8_t)r.qual[read_position];
      seq_bases = std::string(1, (char)base);
      ref_bases = std::string(1, (char)reference_base);
      kind = (base == reference_base) ? ORC_MATCH : ORC_MISMATCH;
    } else if (op == GUAC_CIGAR_S || op == GUAC_CIGAR_N || op == GUAC_CIGAR_H) {
      ref_bases.clear();
      seq_bases.clear();
      quality = r.mapq;
      kind = ORC_CLIPPED;
    } else {
      fail(GUAC_ERR_INVALID_CIGAR, "AssertionError: P CIGAR-op");
    }
  }
  Allele allele() {
    evaluate();
    return Allele{ref_bases, seq_bases};
  }
  bool is_match() {
    evaluate();
    return kind == ORC_MATCH;
  }
  int quality_score() {
    evaluate();
    return quality;
  }
};

// ---- Pileup (pileup/Pileup.scala) ------------------------------------------------------------------------------
struct Pileup {
  int32_t contig = 0;
  int64_t locus = 0;
  uint8_t reference_base = 'N';
  std::vector<Element> elements;

  int depth() const { return (int)elements.size(); }
  int positive_depth() const {
    int n = 0;
    for (const auto& e : elements) n += e.read->positive;
    return n;
  }
  int reference_depth() {
    int n = 0;
    for (auto& e : elements) n += e.is_match();
    return n;
  }
  std::vector<Allele> distinct_alleles() {  // :44 elements.map(_.allele).distinct.sorted
    std::vector<Allele> out;
    for (auto& e : elements) {
      Allele a = e.allele();
      if (std::find(out.begin(), out.end(), a) == out.end()) out.push_back(a);
    }
    std::stable_sort(out.begin(), out.end(), allele_less);
    return out;
  }

  // Pileup.atGreaterLocus :103-132
  Pileup at_greater_locus(int64_t new_locus, uint8_t new_reference_base, const std::vector<const Read*>& new_reads) const {
    if (!(elements.empty() || new_locus > locus)) fail(GUAC_ERR_INVALID_ARGUMENT, "assumption failed: new locus not greater than current locus");
    Pileup p;
    p.contig = contig;
    p.locus = new_locus;
    p.reference_base = new_reference_base;
    if (elements.empty() && new_reads.empty()) return p;
    p.elements.reserve(elements.size());
    for (const auto& e : elements) {
      if (e.read->overlaps_locus(new_locus)) {
        Element c = e;
        c.advance_to_locus(new_locus, new_reference_base);
        p.elements.push_back(std::move(c));
      }
    }
    for (const Read* r : new_reads) p.elements.push_back(Element::create(r, new_locus, new_reference_base));
    return p;
  }
};

// Pileup.referenceBaseAtLocus :157-165
uint8_t reference_base_at_locus(const std::vector<const Read*>& reads, int64_t locus) {
  for (const Read* r : reads) {
    uint8_t b = r->reference_base_at(locus);
    if (is_standard_base(b)) return b;
  }
  return 'N';
}

// Pileup.apply(reads, referenceName, locus, referenceBase) :175-179
Pileup make_pileup(const std::vector<const Read*>& reads, int32_t contig, int64_t locus, uint8_t reference_base) {
  Pileup p;
  p.contig = contig;
  p.locus = locus;
  p.reference_base = reference_base;
  for (const Read* r : reads)
    if (r->overlaps_locus(locus)) p.elements.push_back(Element::create(r, locus, reference_base));
  return p;
}
// Pileup.apply(reads, referenceName, locus) :181-186
Pileup make_pileup(const std::vector<const Read*>& reads, int32_t contig, int64_t locus) {
  std::vector<const Read*> ov;
  for (const Read* r : reads)
    if (r->contig == contig && r->overlaps_locus(locus)) ov.push_back(r);
  uint8_t rb = reference_base_at_locus(ov, locus);
  return make_pileup(ov, contig, locus, rb);
}

// ---- scala.collection.mutable.PriorityQueue (2.10.3) restated: 1-based array heap -------------------------------------
// Ordering from windowing/SlidingWindow.scala:62-68: compare(first, second) = second.end compare first.end, so
// "a < b" iff a.end > b.end and the root is the region with the smallest end.
struct ScalaHeap {
  std::vector<const Read*> a{nullptr};  // slot 0 unused
  static bool lt(const Read* x, const Read* y) { return x->end > y->end; }
  static bool gteq(const Read* x, const Read* y) { return x->end <= y->end; }
  bool empty() const { return a.size() <= 1; }
  size_t size() const { return a.size() - 1; }
  const Read* head() const { return a[1]; }
  void push(const Read* r) {  // += : append then fixUp
    a.push_back(r);
    size_t k = a.size() - 1;
    while (k > 1 && lt(a[k / 2], a[k])) {
      std::swap(a[k], a[k / 2]);
      k /= 2;
    }
  }
  const Read* dequeue() {  // swap(1,last); fixDown(1, last-1)
    size_t last = a.size() - 1;
    std::swap(a[1], a[last]);
    size_t n = last - 1, k = 1;
    while (n >= 2 * k) {
      size_t j = 2 * k;
      if (j < n && lt(a[j], a[j + 1])) ++j;
      if (gteq(a[k], a[j])) break;
      std::swap(a[k], a[j]);
      k = j;
    }
    const Read* out = a[last];
    a.pop_back();
    return out;
  }
  std::vector<const Read*> to_seq() const { return std::vector<const Read*>(a.begin() + 1, a.end()); }
};

// ---- SlidingWindow (windowing/SlidingWindow.scala) ---------------------------------------------------------------
struct SlidingWindow {
  int32_t contig;
  int64_t half_window;
  const std::vector<const Read*>* sorted;  // the raw (supposedly sorted) region iterator
  size_t next = 0;
  int64_t current_locus = -1;
  int64_t most_recent_start = 0;
  std::vector<const Read*> new_regions;
  ScalaHeap queue;

  SlidingWindow(int32_t c, int64_t hw, const std::vector<const Read*>* s) : contig(c), half_window(hw), sorted(s) {}

  bool has_next() const { return next < sorted->size(); }
  const Read* peek() {  // the .map(require...) of the buffered iterator runs when an element is first looked at
    const Read* r = (*sorted)[next];
    if (r->contig != contig) fail(GUAC_ERR_CONTIG_ORDER, "Regions must have the same reference name");
    if (!(r->start >= most_recent_start)) fail(GUAC_ERR_UNSORTED_READS, "Regions must be sorted by start locus");
    most_recent_start = r->start;
    return r;
  }
  static bool overlaps(const Read* r, int64_t locus, int64_t hw) { return r->start - hw <= locus && r->end + hw > locus; }

  std::vector<const Read*> current_regions() const { return queue.to_seq(); }

  void set_current_locus(int64_t locus) {  // :83-110
    if (!(locus >= current_locus)) fail(GUAC_ERR_INVALID_ARGUMENT, "Pileup window can only move forward in locus");
    current_locus = locus;
    while (!queue.empty() && queue.head()->end <= locus - half_window) queue.dequeue();
    new_regions.clear();
    while (has_next() && peek()->start <= locus + half_window) {
      const Read* r = (*sorted)[next++];
      if (overlaps(r, locus, half_window)) new_regions.push_back(r);
    }
    for (const Read* r : new_regions) queue.push(r);
  }
  bool next_locus_with_regions(int64_t* out) {  // :118-128
    for (size_t i = 1; i < queue.a.size(); ++i)
      if (overlaps(queue.a[i], current_locus + 1, half_window)) {
        *out = current_locus + 1;
        return true;
      }
    if (has_next()) {
      int64_t r = std::max<int64_t>(0, peek()->start - half_window);
      if (!(r > current_locus)) fail(GUAC_ERR_UNSORTED_READS, "assertion failed: next region start not past the current locus");
      *out = r;
      return true;
    }
    return false;
  }
};

// LociSet.SingleContig.Iterator (LociSet.scala:287-351)
struct LociIterator {
  std::vector<std::pair<int64_t, int64_t>> ranges;
  size_t ri = 0;
  int64_t head_index = 0;
  bool has_next() const { return ri < ranges.size(); }
  int64_t head() const { return ranges[ri].first + head_index; }
  int64_t next_locus() {
    int64_t l = head();
    ++head_index;
    if (head_index == ranges[ri].second - ranges[ri].first) {
      head_index = 0;
      ++ri;
    }
    return l;
  }
  void skip_to(int64_t locus) {
    while (has_next() && ranges[ri].second <= locus) {
      head_index = 0;
      ++ri;
    }
    if (has_next() && locus >= ranges[ri].first && locus < ranges[ri].second) head_index = locus - ranges[ri].first;
  }
};

// SlidingWindow.advanceMultipleWindows :149-187
bool advance_multiple_windows(std::vector<SlidingWindow>& windows, LociIterator& loci, bool skip_empty, int64_t* out) {
  if (skip_empty) {
    while (loci.has_next()) {
      bool any = false;
      int64_t next_non_empty = 0;
      for (auto& w : windows) {
        int64_t l;
        if (w.next_locus_with_regions(&l)) {
          next_non_empty = any ? std::min(next_non_empty, l) : l;
          any = true;
        }
      }
      if (!any) return false;
      if (next_non_empty <= loci.head()) {
        int64_t next_locus = loci.next_locus();
        for (auto& w : windows) w.set_current_locus(next_locus);
        for (auto& w : windows)
          if (!w.queue.empty()) {
            *out = next_locus;
            return true;
          }
      } else {
        loci.skip_to(next_non_empty);
      }
    }
    return false;
  } else if (loci.has_next()) {
    int64_t next_locus = loci.next_locus();
    for (auto& w : windows) w.set_current_locus(next_locus);
    *out = next_locus;
    return true;
  }
  return false;
}

// ---- output accumulation ------------------------------------------------------------------------------------------
struct Output {
  std::vector<guac_threshold_record> threshold;
  std::vector<guac_somatic_record> somatic;
  std::vector<guac_called_allele> called;
  std::vector<guac_allele_count> allele_counts;
  std::vector<guac_locus_counts> counts;
  std::vector<orc_element> elements;
  std::vector<orc_genotype_likelihood> likelihoods;
  std::vector<uint8_t> bytes;
  guac_stats stats{};
  uint8_t reference_base = 'N';
  int kind = 0;
  uint32_t put(const std::string& s) {
    uint32_t off = (uint32_t)bytes.size();
    bytes.insert(bytes.end(), s.begin(), s.end());
    return off;
  }
};

// ---- GermlineThreshold.Caller.callVariantsAtLocus (commands/GermlineThresholdCaller.scala:90-179) --------------------
void call_variants_at_locus(Pileup& pileup, const guac_threshold_params& p, Output& out) {
  if (pileup.elements.empty()) return;
  // pileup.bySample: groups in Scala Map order (unpinned); canonical here: ascending sample index
  std::map<int32_t, std::vector<Element*>> by_sample;
  for (auto& e : pileup.elements) by_sample[e.read->sample].push_back(&e);
  for (auto& kv : by_sample) {
    int32_t sample = kv.first;
    auto& elems = kv.second;
    int total_reads = (int)elems.size();
    // counts = elements.map(_.allele).groupBy(x => x).mapValues(_.length)
    std::vector<std::pair<Allele, int>> counts;
    for (Element* e : elems) {
      Allele a = e->allele();
      bool found = false;
      for (auto& c : counts)
        if (c.first == a) {
          ++c.second;
          found = true;
          break;
        }
      if (!found) counts.push_back({a, 1});
    }
    // Scala iterates the groupBy map in hash order (unpinned); canonical: Allele.compare order.
    std::stable_sort(counts.begin(), counts.end(), [](const auto& x, const auto& y) { return allele_less(x.first, y.first); });
    std::vector<std::pair<Allele, int>> sorted;
    for (auto& c : counts)
      if ((int)((int64_t)c.second * 100 / total_reads) > p.threshold_percent) sorted.push_back(c);
    std::stable_sort(sorted.begin(), sorted.end(), [](const auto& x, const auto& y) { return x.second > y.second; });
    uint8_t tie = (sorted.size() >= 3 && sorted[1].second == sorted[2].second) ? 1 : 0;

    auto variant = [&](const Allele& a, uint8_t g0, uint8_t g1) {
      guac_threshold_record r{};
      r.start = pileup.locus;
      r.contig = pileup.contig;
      r.sample = sample;
      r.ref_off = out.put(a.ref);
      r.ref_len = (uint16_t)a.ref.size();
      r.alt_off = out.put(a.alt);
      r.alt_len = (uint16_t)a.alt.size();
      r.gt[0] = g0;
      r.gt[1] = g1;
      r.tie = tie;
      out.threshold.push_back(r);
    };
    const std::string ALT = "<ALT>";  // Bases.ALT
    std::string ref1(1, (char)pileup.reference_base);
    if (sorted.empty()) {
      if (p.emit_no_call) variant(Allele{ref1, ALT}, GUAC_GT_NO_CALL, GUAC_GT_NO_CALL);
    } else if (sorted.size() == 1 && !sorted[0].first.is_variant()) {
      if (p.emit_ref) variant(Allele{ref1, ALT}, GUAC_GT_REF, GUAC_GT_REF);
    } else if (sorted.size() == 1) {
      variant(sorted[0].first, GUAC_GT_ALT, GUAC_GT_ALT);
    } else {
      const Allele& a1 = sorted[0].first;
      const Allele& a2 = sorted[1].first;
      if ((!a1.is_variant() || !a2.is_variant()) && (a1.alt.empty() != a2.alt.empty())) {
        // heterozygous deletion: nothing
      } else if (a1.is_variant() != a2.is_variant()) {
        variant(a1.is_variant() ? a1 : a2, GUAC_GT_REF, GUAC_GT_ALT);
      } else if (a1.is_variant() && a2.is_variant()) {
        variant(a1, GUAC_GT_ALT, GUAC_GT_OTHER_ALT);
        variant(a2, GUAC_GT_ALT, GUAC_GT_OTHER_ALT);
      } else {
        if (a1.ref == "N" || a2.ref == "N") {
          const std::string& proper = (a1.ref == "N") ? a2.ref : a1.ref;
          variant(Allele{proper, ALT}, GUAC_GT_REF, GUAC_GT_REF);
        } else {
          fail(GUAC_ERR_MULTIPLE_REFERENCE_BASES, "Multiple reference bases found at locus " + std::to_string(pileup.locus));
        }
      }
    }
    if (tie) ++out.stats.tie_loci;
  }
}

// ---- Likelihood (likelihood/Likelihood.scala) ----------------------------------------------------------------------
struct GenotypeL {
  Allele a1, a2;
  double value;
  bool has_variant_allele() const { return a1.is_variant() || a2.is_variant(); }
};

// likelihoodsOfGenotypes :149-201
std::vector<double> likelihoods_of_genotypes(std::vector<Element*>& elements, const std::vector<std::pair<Allele, Allele>>& genotypes,
                                             bool include_alignment, bool log_space, bool normalize) {
  std::vector<Allele> alleles;
  for (auto& g : genotypes) {
    if (std::find(alleles.begin(), alleles.end(), g.first) == alleles.end()) alleles.push_back(g.first);
    if (std::find(alleles.begin(), alleles.end(), g.second) == alleles.end()) alleles.push_back(g.second);
  }
  std::stable_sort(alleles.begin(), alleles.end(), allele_less);
  size_t depth = elements.size();
  std::vector<double> P(alleles.size() * depth);  // colt DenseDoubleMatrix2D(alleles.size, depth)
  std::vector<Allele> element_alleles(depth);
  std::vector<double> success(depth);
  for (size_t e = 0; e < depth; ++e) {
    element_alleles[e] = elements[e]->allele();
    double s = phred_to_success(elements[e]->quality_score());           // probabilityCorrectIgnoringAlignment :48-50
    if (include_alignment) s = s * phred_to_success(elements[e]->read->mapq);  // IncludingAlignment :60-62
    success[e] = s;
  }
  for (size_t a = 0; a < alleles.size(); ++a)
    for (size_t e = 0; e < depth; ++e) P[a * depth + e] = (alleles[a] == element_alleles[e]) ? success[e] : 1 - success[e];
  auto index_of = [&](const Allele& a) { return (size_t)(std::find(alleles.begin(), alleles.end(), a) - alleles.begin()); };
  std::vector<double> logl;
  for (auto& g : genotypes) {
    const double* r1 = &P[index_of(g.first) * depth];
    const double* r2 = &P[index_of(g.second) * depth];
    // colt aggregate(other, plus, chain(log, plus)): from the last element down to the first; NaN if empty
    double agg;
    if (depth == 0)
      agg = std::numeric_limits<double>::quiet_NaN();
    else {
      agg = std::log(r1[depth - 1] + r2[depth - 1]);
      for (size_t i = depth - 1; i-- > 0;) agg = agg + std::log(r1[i] + r2[i]);
    }
    logl.push_back(agg + std::log(1.0) - std::log(2.0) * (double)depth);
  }
  if (normalize) {
    double total = 0.0;
    for (double l : logl) total += std::exp(l);
    double log_total = std::log(total);
    for (double& l : logl) l = l - log_total;
  }
  if (!log_space)
    for (double& l : logl) l = std::exp(l);
  return logl;
}

// likelihoodsOfAllPossibleGenotypesFromPileup :99-113
std::vector<GenotypeL> likelihoods_of_all_possible_genotypes(Pileup& pileup, bool include_alignment, bool log_space, bool normalize) {
  std::vector<Allele> alleles;
  for (auto& a : pileup.distinct_alleles()) {
    bool ok = true;
    for (char c : a.alt) ok = ok && is_standard_base((uint8_t)c);
    if (ok) alleles.push_back(a);
  }
  std::vector<std::pair<Allele, Allele>> genotypes;
  for (size_t i = 0; i < alleles.size(); ++i)
    for (size_t j = i; j < alleles.size(); ++j) genotypes.push_back({alleles[i], alleles[j]});
  std::vector<Element*> elems;
  for (auto& e : pileup.elements) elems.push_back(&e);
  std::vector<double> l = likelihoods_of_genotypes(elems, genotypes, include_alignment, log_space, normalize);
  std::vector<GenotypeL> out;
  for (size_t i = 0; i < genotypes.size(); ++i) out.push_back({genotypes[i].first, genotypes[i].second, l[i]});
  return out;
}

// ---- breeze mean / median (third party; restated) --------------------------------------------------------------------
double breeze_mean(const std::vector<double>& v) {
  double mu = 0.0;
  long n = 0;
  for (double y : v) {
    n += 1;
    double d = y - mu;
    mu = mu + d / (double)n;
  }
  return mu;
}
double breeze_median(std::vector<double> v) {
  std::sort(v.begin(), v.end());
  size_t n = v.size();
  if (n % 2 == 1) return v[(n - 1) / 2];
  return (v[n / 2 - 1] + v[n / 2]) / 2;
}
double breeze_median_int(std::vector<int> v) {  // DenseVector[Int]: integer arithmetic, widened afterwards
  std::sort(v.begin(), v.end());
  size_t n = v.size();
  if (n % 2 == 1) return (double)v[(n - 1) / 2];
  return (double)((v[n / 2 - 1] + v[n / 2]) / 2);
}

// AlleleEvidence.apply (variants/AlleleEvidence.scala:58-101)
guac_allele_evidence allele_evidence(double likelihood, const Allele& allele, Pileup& pileup) {
  guac_allele_evidence ev{};
  std::vector<double> mq, bq;
  std::vector<int> mm;
  int allele_depth = 0, allele_pos = 0;
  for (auto& e : pileup.elements) {
    if (e.allele() == allele) {  // alleleReadDepthAndPositiveDepth pileup/Pileup.scala:139-145
      ++allele_depth;
      allele_pos += e.read->positive;
      mq.push_back((double)e.read->mapq);
      bq.push_back((double)e.quality_score());
    }
  }
  ev.likelihood = likelihood;
  ev.read_depth = pileup.depth();
  ev.allele_read_depth = allele_depth;
  ev.forward_depth = pileup.positive_depth();
  ev.allele_forward_depth = allele_pos;
  double nan = std::numeric_limits<double>::quiet_NaN();
  if (allele_depth == 0) {
    ev.mean_mapping_quality = ev.median_mapping_quality = ev.mean_base_quality = ev.median_base_quality = ev.median_mismatches_per_read = nan;
  } else {
    for (auto& e : pileup.elements)
      if (e.allele() == allele) mm.push_back(e.read->count_of_mismatches());
    ev.mean_mapping_quality = breeze_mean(mq);
    ev.median_mapping_quality = breeze_median(mq);
    ev.mean_base_quality = breeze_mean(bq);
    ev.median_base_quality = breeze_median(bq);
    ev.median_mismatches_per_read = breeze_median_int(mm);
  }
  return ev;
}

// PileupFilter.apply (filters/PileupFilter.scala:69-89) with minEdgeDistance = 0
Pileup pileup_filter(Pileup& p, bool filter_multi_allelic, int min_alignment_quality) {
  Pileup out;
  out.contig = p.contig;
  out.locus = p.locus;
  out.reference_base = p.reference_base;
  bool drop_all = false;
  if (filter_multi_allelic) {  // MultiAllelicPileupFilter filters/PileupElementsFilter.scala:32-38
    std::vector<Allele> d;
    for (auto& e : p.elements) {
      Allele a = e.allele();
      if (std::find(d.begin(), d.end(), a) == d.end()) d.push_back(a);
    }
    drop_all = d.size() > 2;
  }
  if (!drop_all)
    for (auto& e : p.elements)
      if (!(min_alignment_quality > 0) || e.read->mapq >= min_alignment_quality) out.elements.push_back(e);
  return out;
}

// SomaticStandard.Caller.findPotentialVariantAtLocus (commands/SomaticStandardCaller.scala:162-245)
void find_potential_variant_at_locus(Pileup& tumor, Pileup& normal, const guac_somatic_params& p, Output& out) {
  Pileup fn = pileup_filter(normal, p.filter_multi_allelic != 0, p.min_alignment_quality);
  Pileup ft = pileup_filter(tumor, p.filter_multi_allelic != 0, p.min_alignment_quality);
  if (ft.elements.empty() || fn.elements.empty() || ft.depth() > p.max_read_depth || fn.depth() > p.max_read_depth ||
      ft.reference_depth() == ft.depth())
    return;
  std::vector<GenotypeL> gl = likelihoods_of_all_possible_genotypes(ft, /*include_alignment=*/true, false, true);
  if (gl.empty()) return;
  // maxBy(_._2) = reduceLeft((x, y) => if (f(x) >= f(y)) x else y)
  size_t best = 0;
  for (size_t i = 1; i < gl.size(); ++i)
    if (!(gl[best].value >= gl[i].value)) best = i;
  const GenotypeL& g = gl[best];
  if (!g.has_variant_allele()) return;
  std::vector<GenotypeL> nl = likelihoods_of_all_possible_genotypes(fn, /*include_alignment=*/false, false, true);
  double normal_variants_total = 0.0;  // summed in Scala Map order (unpinned) — genotype order here; fp tolerance
  for (auto& x : nl)
    if (x.has_variant_allele()) normal_variants_total += x.value;
  double somatic_odds = g.value / normal_variants_total;
  if (!(somatic_odds * 100 >= (double)p.odds_threshold)) return;
  const Allele* allele = nullptr;
  if (g.a1.is_variant() && !g.a1.alt.empty())
    allele = &g.a1;
  else if (g.a2.is_variant() && !g.a2.alt.empty())
    allele = &g.a2;
  if (!allele) return;
  guac_somatic_record r{};
  r.start = tumor.locus;
  r.contig = tumor.contig;
  r.sample = tumor.elements.front().read->sample;  // tumorPileup.sampleName = elements.head.read.sampleName
  r.ref_off = out.put(allele->ref);
  r.ref_len = (uint16_t)allele->ref.size();
  r.alt_off = out.put(allele->alt);
  r.alt_len = (uint16_t)allele->alt.size();
  r.somatic_log_odds = std::log(somatic_odds);
  r.tumor = allele_evidence(g.value, *allele, ft);
  r.normal = allele_evidence(1 - normal_variants_total, Allele{allele->ref, allele->ref}, fn);
  // CalledSomaticAllele.phredScaledSomaticLikelihood variants/CalledSomaticAllele.scala:49-50
  r.phred_scaled_somatic_likelihood = success_probability_to_phred(r.tumor.likelihood * r.normal.likelihood - 1e-10);
  out.somatic.push_back(r);
}

// GermlineStandard.Caller.callVariantsAtLocus (commands/GermlineStandardCaller.scala:90-124)
void call_standard_at_locus(Pileup& pileup, const guac_standard_params& p, Output& out) {
  if (pileup.elements.empty()) return;
  // pileup.bySample: groups in Scala Map order (unpinned); canonical here: ascending sample index
  std::map<int32_t, std::vector<const Element*>> by_sample;
  for (auto& e : pileup.elements) by_sample[e.read->sample].push_back(&e);
  for (auto& kv : by_sample) {
    Pileup sample_pileup;  // Pileup(referenceName, locus, referenceBase, the sample's elements)
    sample_pileup.contig = pileup.contig;
    sample_pileup.locus = pileup.locus;
    sample_pileup.reference_base = pileup.reference_base;
    for (const Element* e : kv.second) sample_pileup.elements.push_back(*e);
    Pileup filtered = sample_pileup;  // QualityAlignedReadsFilter (filters/PileupElementsFilter.scala:48-50)
    filtered.elements.clear();
    for (auto& e : sample_pileup.elements)
      if (e.read->mapq >= p.min_alignment_quality) filtered.elements.push_back(e);
    if (filtered.elements.empty()) continue;
    std::vector<GenotypeL> gl = likelihoods_of_all_possible_genotypes(filtered, /*include_alignment=*/false, /*log_space=*/true,
                                                                      /*normalize=*/true);
    if (gl.empty()) continue;  // (Scala's maxBy would throw on an empty list: every allele holds a non-ACGT base)
    size_t best = 0;  // maxBy = reduceLeft((x, y) => if (f(x) >= f(y)) x else y)
    for (size_t i = 1; i < gl.size(); ++i)
      if (!(gl[best].value >= gl[i].value)) best = i;
    const double probability = std::exp(gl[best].value);
    const Allele* pair[2] = {&gl[best].a1, &gl[best].a2};
    for (const Allele* a : pair) {  // genotype.getNonReferenceAlleles: both copies of a homozygous alternate
      if (!a->is_variant()) continue;
      guac_called_allele r{};
      r.start = pileup.locus;
      r.contig = pileup.contig;
      r.sample = kv.first;
      r.ref_off = out.put(a->ref);
      r.ref_len = (uint16_t)a->ref.size();
      r.alt_off = out.put(a->alt);
      r.alt_len = (uint16_t)a->alt.size();
      r.evidence = allele_evidence(probability, *a, sample_pileup);  // over the unfiltered sample pileup (:120)
      r.phred_scaled_likelihood = success_probability_to_phred(r.evidence.likelihood - 1e-10);  // AlleleEvidence.scala:52
      out.called.push_back(r);
    }
  }
}

// VariantSupport.Caller.pileupToAlleleCounts (commands/VariantSupport.scala:110-118): elements.groupBy(_.allele)
void allele_counts_at_locus(Pileup& pileup, Output& out) {
  if (pileup.elements.empty()) return;
  std::vector<std::pair<Allele, int>> counts;  // Scala Map order is unspecified; the records are sorted afterwards
  for (auto& e : pileup.elements) {
    Allele a = e.allele();
    bool found = false;
    for (auto& c : counts)
      if (c.first == a) {
        ++c.second;
        found = true;
        break;
      }
    if (!found) counts.push_back({a, 1});
  }
  for (auto& c : counts) {
    guac_allele_count r{};
    r.start = pileup.locus;
    r.contig = pileup.contig;
    r.sample = pileup.elements.front().read->sample;  // pileup.sampleName = elements.head.read.sampleName
    r.ref_off = out.put(c.first.ref);
    r.ref_len = (uint16_t)c.first.ref.size();
    r.alt_off = out.put(c.first.alt);
    r.alt_len = (uint16_t)c.first.alt.size();
    r.count = c.second;
    out.allele_counts.push_back(r);
  }
}

void counts_at_locus(Pileup& p, Output& out) {
  guac_locus_counts c{};
  c.locus = p.locus;
  c.contig = p.contig;
  c.depth = p.depth();
  c.positive_depth = p.positive_depth();
  c.reference_depth = p.reference_depth();
  c.reference_base = p.reference_base;
  for (auto& e : p.elements) {
    e.evaluate();
    int b = -1;
    if (e.kind == ORC_MATCH || e.kind == ORC_MISMATCH) {
      switch (e.seq_bases[0]) {
        case 'A': b = 0; break;
        case 'C': b = 1; break;
        case 'G': b = 2; break;
        case 'T': b = 3; break;
      }
    }
    if (b >= 0)
      ++c.base_count[b];
    else
      ++c.other_count;
  }
  out.counts.push_back(c);
}

// ---- the engine: windowTaskFlatMapMultipleRDDs + collectByContig + windowFlatMapWithState + initOrMovePileup ----------
//      (DistributedUtil.scala:558-634, 473-486, 388-418, 260-274)
struct Engine {
  std::vector<const ReadSet*> sets;
  const guac_reference* ref = nullptr;
  bool skip_empty = true;

  uint8_t fasta_base(int32_t contig, int64_t locus) const {
    uint64_t lo = ref->base_off[contig], hi = ref->base_off[contig + 1];
    if ((uint64_t)locus >= hi - lo) fail(GUAC_ERR_INVALID_ARGUMENT, "locus beyond the reference contig");
    return ref->bases[lo + locus];
  }

  template <typename F>
  void run_task(const std::vector<guac_locus_range>& task_ranges, Output& out, F&& per_locus) {
    // contigs of the task's loci (the reference walks them in contig-NAME order; output is re-sorted canonically)
    std::map<int32_t, std::vector<std::pair<int64_t, int64_t>>> by_contig;
    for (auto& r : task_ranges)
      if (r.end > r.start) by_contig[r.contig].push_back({r.start, r.end});
    for (auto& kv : by_contig) {
      int32_t contig = kv.first;
      auto& ranges = kv.second;
      std::sort(ranges.begin(), ranges.end());
      // reads of this task on this contig: those overlapping any of the task's loci (:585-597), in input order
      std::vector<std::vector<const Read*>> task_reads(sets.size());
      for (size_t s = 0; s < sets.size(); ++s) {
        for (const Read& r : sets[s]->reads) {
          ++out.stats.reads_total;
          if (r.contig != contig) continue;
          bool hit = false;
          for (auto& rg : ranges)
            if (r.start < rg.second && r.end > rg.first) {
              hit = true;
              break;
            }
          if (hit) {
            task_reads[s].push_back(&r);
            ++out.stats.reads_expanded;
          }
        }
      }
      std::vector<SlidingWindow> windows;
      for (size_t s = 0; s < sets.size(); ++s) windows.emplace_back(contig, 0, &task_reads[s]);
      LociIterator loci;
      loci.ranges = ranges;
      for (auto& rg : ranges) out.stats.loci_requested += (uint64_t)(rg.second - rg.first);
      std::vector<std::unique_ptr<Pileup>> state(sets.size());  // Option[Pileup] per sample, None at task/contig start
      int64_t locus;
      while (advance_multiple_windows(windows, loci, skip_empty, &locus)) {
        ++out.stats.loci_visited;
        for (size_t s = 0; s < sets.size(); ++s) {  // initOrMovePileup :260-274
          SlidingWindow& w = windows[s];
          uint8_t rb = ref ? fasta_base(contig, locus) : reference_base_at_locus(w.current_regions(), locus);
          if (!state[s]) {
            state[s].reset(new Pileup(make_pileup(w.current_regions(), contig, locus, rb)));
          } else {
            Pileup np = state[s]->at_greater_locus(locus, rb, w.new_regions);
            *state[s] = std::move(np);
          }
        }
        per_locus(state, out);
      }
    }
  }

  template <typename F>
  void run(const guac_locus_range* ranges, size_t n_ranges, int n_threads, Output& out, F&& per_locus) {
    std::map<int32_t, std::vector<guac_locus_range>> tasks;
    for (size_t i = 0; i < n_ranges; ++i) tasks[ranges[i].task].push_back(ranges[i]);
    std::vector<std::vector<guac_locus_range>> task_list;
    for (auto& kv : tasks) task_list.push_back(kv.second);
    std::vector<Output> outs(task_list.size());
    std::vector<OrcError> errors(task_list.size(), OrcError{0, ""});
    size_t nt = std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, n_threads), task_list.size()));
    auto worker = [&](size_t tid) {
      for (size_t t = tid; t < task_list.size(); t += nt) {
        try {
          run_task(task_list[t], outs[t], per_locus);
        } catch (const OrcError& e) {
          errors[t] = e;
        }
      }
    };
    if (nt == 1)
      worker(0);
    else {
      std::vector<std::thread> th;
      for (size_t i = 0; i < nt; ++i) th.emplace_back(worker, i);
      for (auto& t : th) t.join();
    }
    for (auto& e : errors)
      if (e.code) throw e;
    // concatenate (task order), re-basing byte-pool offsets
    for (auto& o : outs) {
      uint32_t base = (uint32_t)out.bytes.size();
      out.bytes.insert(out.bytes.end(), o.bytes.begin(), o.bytes.end());
      for (auto r : o.threshold) {
        r.ref_off += base;
        r.alt_off += base;
        out.threshold.push_back(r);
      }
      for (auto r : o.somatic) {
        r.ref_off += base;
        r.alt_off += base;
        out.somatic.push_back(r);
      }
      for (auto r : o.allele_counts) {
        r.ref_off += base;
        r.alt_off += base;
        out.allele_counts.push_back(r);
      }
      for (auto r : o.called) {
        r.ref_off += base;
        r.alt_off += base;
        out.called.push_back(r);
      }
      out.counts.insert(out.counts.end(), o.counts.begin(), o.counts.end());
      out.stats.reads_expanded += o.stats.reads_expanded;
      out.stats.loci_requested += o.stats.loci_requested;
      out.stats.loci_visited += o.stats.loci_visited;
      out.stats.tie_loci += o.stats.tie_loci;
    }
    out.stats.reads_total = 0;
    for (auto* s : sets) out.stats.reads_total += s->reads.size();
  }
};

// canonical output order: (contig, start, sample, ref, alt)   — SURVEY 8b "record layouts"
template <typename R>
void sort_records(std::vector<R>& v, const std::vector<uint8_t>& bytes) {
  auto key = [&](const R& r) {
    return std::make_tuple(r.contig, r.start, r.sample, std::string(bytes.begin() + r.ref_off, bytes.begin() + r.ref_off + r.ref_len),
                           std::string(bytes.begin() + r.alt_off, bytes.begin() + r.alt_off + r.alt_len));
  };
  std::stable_sort(v.begin(), v.end(), [&](const R& a, const R& b) { return key(a) < key(b); });
}

}  // namespace

struct orc_result {
  Output o;
};

template <typename F>
static int guarded(F&& f) {
  try {
    f();
    return GUAC_OK;
  } catch (const OrcError& e) {
    g_last_error = e.msg;
    return e.code;
  } catch (const std::bad_alloc&) {
    g_last_error = "out of memory";
    return GUAC_ERR_OOM;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return GUAC_ERR_INVALID_ARGUMENT;
  }
}

static std::vector<const Read*> all_reads(const ReadSet& rs) {
  std::vector<const Read*> v;
  for (auto& r : rs.reads) v.push_back(&r);
  return v;
}

extern "C" {

const char* orc_last_error(void) { return g_last_error.c_str(); }

int orc_germline_threshold(const guac_read_batch* batch, const guac_reference* ref, const guac_locus_range* ranges,
                           size_t n_ranges, const guac_threshold_params* params, int n_threads, orc_result** out) {
  return guarded([&] {
    ReadSet rs;
    load_batch(batch, rs);
    auto res = std::make_unique<orc_result>();
    Engine eng;
    eng.sets = {&rs};
    eng.ref = ref;
    eng.skip_empty = params->skip_empty != 0;
    guac_threshold_params p = *params;
    eng.run(ranges, n_ranges, n_threads, res->o,
            [p](std::vector<std::unique_ptr<Pileup>>& st, Output& o) { call_variants_at_locus(*st[0], p, o); });
    sort_records(res->o.threshold, res->o.bytes);
    res->o.stats.records = res->o.threshold.size();
    *out = res.release();
  });
}

int orc_somatic_standard(const guac_read_batch* tumor, const guac_read_batch* normal, const guac_reference* ref,
                         const guac_locus_range* ranges, size_t n_ranges, const guac_somatic_params* params,
                         int n_threads, orc_result** out) {
  return guarded([&] {
    ReadSet t, n;
    load_batch(tumor, t);
    load_batch(normal, n);
    auto res = std::make_unique<orc_result>();
    Engine eng;
    eng.sets = {&t, &n};
    eng.ref = ref;
    eng.skip_empty = params->skip_empty != 0;
    guac_somatic_params p = *params;
    eng.run(ranges, n_ranges, n_threads, res->o, [p](std::vector<std::unique_ptr<Pileup>>& st, Output& o) {
      find_potential_variant_at_locus(*st[0], *st[1], p, o);
    });
    sort_records(res->o.somatic, res->o.bytes);
    res->o.stats.records = res->o.somatic.size();
    *out = res.release();
  });
}

int orc_germline_standard(const guac_read_batch* batch, const guac_reference* ref, const guac_locus_range* ranges,
                          size_t n_ranges, const guac_standard_params* params, int n_threads, orc_result** out) {
  return guarded([&] {
    ReadSet rs;
    load_batch(batch, rs);
    auto res = std::make_unique<orc_result>();
    Engine eng;
    eng.sets = {&rs};
    eng.ref = ref;
    eng.skip_empty = params->skip_empty != 0;
    guac_standard_params p = *params;
    eng.run(ranges, n_ranges, n_threads, res->o,
            [p](std::vector<std::unique_ptr<Pileup>>& st, Output& o) { call_standard_at_locus(*st[0], p, o); });
    sort_records(res->o.called, res->o.bytes);
    res->o.stats.records = res->o.called.size();
    *out = res.release();
  });
}

int orc_allele_counts(const guac_read_batch* batch, const guac_reference* ref, const guac_locus_range* ranges,
                      size_t n_ranges, int n_threads, orc_result** out) {
  return guarded([&] {
    ReadSet rs;
    load_batch(batch, rs);
    auto res = std::make_unique<orc_result>();
    Engine eng;
    eng.sets = {&rs};
    eng.ref = ref;
    eng.skip_empty = true;
    eng.run(ranges, n_ranges, n_threads, res->o,
            [](std::vector<std::unique_ptr<Pileup>>& st, Output& o) { allele_counts_at_locus(*st[0], o); });
    sort_records(res->o.allele_counts, res->o.bytes);
    res->o.stats.records = res->o.allele_counts.size();
    *out = res.release();
  });
}

int orc_pileup_counts(const guac_read_batch* batch, const guac_reference* ref, const guac_locus_range* ranges,
                      size_t n_ranges, int skip_empty, int n_threads, orc_result** out) {
  return guarded([&] {
    ReadSet rs;
    load_batch(batch, rs);
    auto res = std::make_unique<orc_result>();
    Engine eng;
    eng.sets = {&rs};
    eng.ref = ref;
    eng.skip_empty = skip_empty != 0;
    eng.run(ranges, n_ranges, n_threads, res->o,
            [](std::vector<std::unique_ptr<Pileup>>& st, Output& o) { counts_at_locus(*st[0], o); });
    std::stable_sort(res->o.counts.begin(), res->o.counts.end(), [](const guac_locus_counts& a, const guac_locus_counts& b) {
      return std::make_pair(a.contig, a.locus) < std::make_pair(b.contig, b.locus);
    });
    res->o.stats.records = res->o.counts.size();
    *out = res.release();
  });
}

size_t orc_result_n(const orc_result* r) {
  const Output& o = r->o;
  return std::max({o.threshold.size(), o.somatic.size(), o.called.size(), o.allele_counts.size(), o.counts.size(), o.elements.size(), o.likelihoods.size()});
}
const guac_threshold_record* orc_result_threshold_records(const orc_result* r) { return r->o.threshold.data(); }
const guac_somatic_record* orc_result_somatic_records(const orc_result* r) { return r->o.somatic.data(); }
const guac_called_allele* orc_result_called_alleles(const orc_result* r) { return r->o.called.data(); }
const guac_allele_count* orc_result_allele_counts(const orc_result* r) { return r->o.allele_counts.data(); }
const guac_locus_counts* orc_result_counts(const orc_result* r) { return r->o.counts.data(); }
const orc_element* orc_result_elements(const orc_result* r) { return r->o.elements.data(); }
const orc_genotype_likelihood* orc_result_likelihoods(const orc_result* r) { return r->o.likelihoods.data(); }
const uint8_t* orc_result_bytes(const orc_result* r, size_t* n) {
  if (n) *n = r->o.bytes.size();
  return r->o.bytes.data();
}
const guac_stats* orc_result_stats(const orc_result* r) { return &r->o.stats; }
uint8_t orc_result_reference_base(const orc_result* r) { return r->o.reference_base; }
void orc_result_free(orc_result* r) { delete r; }

int orc_pileup_at(const guac_read_batch* batch, int32_t contig, int64_t locus, int reference_base, orc_result** out) {
  return guarded([&] {
    ReadSet rs;
    load_batch(batch, rs);
    auto reads = all_reads(rs);
    Pileup p;
    if (reference_base >= 0) {
      std::vector<const Read*> ov;
      for (auto* r : reads)
        if (r->contig == contig) ov.push_back(r);
      p = make_pileup(ov, contig, locus, (uint8_t)reference_base);
    } else {
      p = make_pileup(reads, contig, locus);
    }
    auto res = std::make_unique<orc_result>();
    res->o.reference_base = p.reference_base;
    for (auto& e : p.elements) {
      e.evaluate();
      orc_element oe{};
      oe.read_index = e.read->idx;
      oe.kind = e.kind;
      oe.quality_score = e.quality;
      oe.read_position = e.read_position;
      oe.cigar_element_index = e.cigar_index;
      oe.index_within_cigar_element = e.index_within;
      oe.ref_off = res->o.put(e.ref_bases);
      oe.ref_len = (uint32_t)e.ref_bases.size();
      oe.seq_off = res->o.put(e.seq_bases);
      oe.seq_len = (uint32_t)e.seq_bases.size();
      oe.is_positive_strand = e.read->positive;
      res->o.elements.push_back(oe);
    }
    *out = res.release();
  });
}

int orc_md_reference(const guac_read_batch* batch, uint64_t read_index, uint8_t* outb, size_t max_out, size_t* n_out,
                     int* n_mismatches) {
  return guarded([&] {
    ReadSet rs;
    load_batch(batch, rs);
    if (read_index >= rs.reads.size()) fail(GUAC_ERR_INVALID_ARGUMENT, "read index out of range");
    const auto& v = rs.reads[read_index].md_reference();
    if (n_out) *n_out = v.size();
    if (n_mismatches) *n_mismatches = rs.reads[read_index].count_of_mismatches();
    std::memcpy(outb, v.data(), std::min(max_out, v.size()));
  });
}

int orc_likelihoods_at(const guac_read_batch* batch, int32_t contig, int64_t locus, int include_alignment,
                       int log_space, int normalize, orc_result** out) {
  return guarded([&] {
    ReadSet rs;
    load_batch(batch, rs);
    Pileup p = make_pileup(all_reads(rs), contig, locus);
    auto gl = likelihoods_of_all_possible_genotypes(p, include_alignment != 0, log_space != 0, normalize != 0);
    auto res = std::make_unique<orc_result>();
    for (auto& g : gl) {
      orc_genotype_likelihood o{};
      o.a1_ref_off = res->o.put(g.a1.ref);
      o.a1_ref_len = (uint32_t)g.a1.ref.size();
      o.a1_alt_off = res->o.put(g.a1.alt);
      o.a1_alt_len = (uint32_t)g.a1.alt.size();
      o.a2_ref_off = res->o.put(g.a2.ref);
      o.a2_ref_len = (uint32_t)g.a2.ref.size();
      o.a2_alt_off = res->o.put(g.a2.alt);
      o.a2_alt_len = (uint32_t)g.a2.alt.size();
      o.value = g.value;
      res->o.likelihoods.push_back(o);
    }
    *out = res.release();
  });
}

int orc_somatic_at(const guac_read_batch* tumor, const guac_read_batch* normal, int32_t contig, int64_t locus,
                   const guac_somatic_params* params, orc_result** out) {
  return guarded([&] {
    ReadSet t, n;
    load_batch(tumor, t);
    load_batch(normal, n);
    Pileup pt = make_pileup(all_reads(t), contig, locus);
    Pileup pn = make_pileup(all_reads(n), contig, locus);
    auto res = std::make_unique<orc_result>();
    if (pt.elements.empty()) {  // tumorPileup.sampleName would throw on an empty pileup only if a call is made
      *out = res.release();
      return;
    }
    find_potential_variant_at_locus(pt, pn, *params, res->o);
    *out = res.release();
  });
}

int orc_threshold_at(const guac_read_batch* batch, int32_t contig, int64_t locus, const guac_threshold_params* params,
                     orc_result** out) {
  return guarded([&] {
    ReadSet rs;
    load_batch(batch, rs);
    Pileup p = make_pileup(all_reads(rs), contig, locus);
    auto res = std::make_unique<orc_result>();
    res->o.reference_base = p.reference_base;
    call_variants_at_locus(p, *params, res->o);
    *out = res.release();
  });
}

int orc_allele_evidence_at(const guac_read_batch* batch, int32_t contig, int64_t locus, const uint8_t* ref,
                           size_t ref_len, const uint8_t* alt, size_t alt_len, double likelihood,
                           guac_allele_evidence* out) {
  return guarded([&] {
    ReadSet rs;
    load_batch(batch, rs);
    Pileup p = make_pileup(all_reads(rs), contig, locus);
    Allele a{std::string((const char*)ref, ref_len), std::string((const char*)alt, alt_len)};
    *out = allele_evidence(likelihood, a, p);
  });
}

int orc_visited_loci(const guac_read_batch* a, const guac_read_batch* b, const guac_locus_range* ranges, size_t n_ranges,
                     int skip_empty, int64_t half_window, int64_t* loci_out, int32_t* count_a_out, int32_t* count_b_out,
                     size_t max_out, size_t* n_out) {
  return guarded([&] {
    ReadSet ra, rb;
    load_batch(a, ra);
    if (b) load_batch(b, rb);
    if (n_ranges == 0) fail(GUAC_ERR_INVALID_ARGUMENT, "no ranges");
    int32_t contig = ranges[0].contig;
    std::vector<const Read*> va, vb;
    for (auto& r : ra.reads)
      if (r.contig == contig) va.push_back(&r);
    for (auto& r : rb.reads)
      if (r.contig == contig) vb.push_back(&r);
    std::vector<SlidingWindow> windows;
    windows.emplace_back(contig, half_window, &va);
    if (b) windows.emplace_back(contig, half_window, &vb);
    LociIterator loci;
    for (size_t i = 0; i < n_ranges; ++i) loci.ranges.push_back({ranges[i].start, ranges[i].end});
    size_t n = 0;
    int64_t locus;
    while (advance_multiple_windows(windows, loci, skip_empty != 0, &locus)) {
      if (n < max_out) {
        loci_out[n] = locus;
        if (count_a_out) count_a_out[n] = (int32_t)windows[0].queue.size();
        if (count_b_out) count_b_out[n] = b ? (int32_t)windows[1].queue.size() : 0;
      }
      ++n;
    }
    *n_out = n;
  });
}

// partitionLociUniformly DistributedUtil.scala:83-108
int orc_partition_loci_uniformly(int64_t tasks, const guac_locus_range* loci, size_t n_loci, guac_locus_range* out,
                                 size_t max_out, size_t* n_out) {
  return guarded([&] {
    if (tasks < 1) fail(GUAC_ERR_INVALID_ARGUMENT, "`tasks` (--parallelism) should be >= 1");
    int64_t count = 0;
    for (size_t i = 0; i < n_loci; ++i) count += loci[i].end - loci[i].start;
    double loci_per_task = std::max(1.0, (double)count / (double)tasks);
    int64_t assigned = 0, task = 0;
    auto remaining = [&]() { return (int64_t)std::floor(((double)(task + 1) * loci_per_task) - (double)assigned + 0.5); };
    std::vector<guac_locus_range> v;
    for (size_t i = 0; i < n_loci; ++i) {
      int64_t start = loci[i].start, end = loci[i].end;
      while (start < end) {
        int64_t length = std::min(remaining(), end - start);
        if (length <= 0) fail(GUAC_ERR_INVALID_ARGUMENT, "partitioning made no progress");
        // LociMap.Builder.put coalesces adjacent ranges that carry the same value
        if (!v.empty() && v.back().contig == loci[i].contig && v.back().task == (int32_t)task && v.back().end == start)
          v.back().end = start + length;
        else
          v.push_back(guac_locus_range{loci[i].contig, (int32_t)task, start, start + length});
        start += length;
        assigned += length;
        if (remaining() == 0) task += 1;
      }
    }
    size_t n = v.size();
    for (size_t i = 0; i < n && i < max_out; ++i) out[i] = v[i];
    *n_out = n;
  });
}

// partitionLociByApproximateDepth DistributedUtil.scala:162-251.  `loci` in LociSet order (the order partitionLociUniformly
// walks them); regions = the reads of every batch.  Restated literally: micro partitions = partitionLociUniformly(accuracy *
// tasks), one count per (region, micro partition it overlaps) — LociMap.getAll returns a Set —, then the greedy assignment in
// Double arithmetic.  scala.math.round applied to a Long resolves to round(Float): Int in Scala 2.10 (no Long overload):
// kept, it only matters beyond 2^24 regions in one micro partition.
namespace {
long long scala_round_double(double x) {  // java.lang.Math.round(double)
  if (std::isnan(x)) return 0;
  double f = std::floor(x + 0.5);
  if (f >= 9.2233720368547758e18) return INT64_MAX;
  if (f <= -9.2233720368547758e18) return INT64_MIN;
  return (long long)f;
}
long long scala_round_long(long long v) {  // math.round(v: Long) = java.lang.Math.round(v.toFloat): Int
  float f = std::floor((float)v + 0.5f);
  if (f >= 2147483648.0f) return INT32_MAX;
  if (f <= -2147483648.0f) return INT32_MIN;
  return (long long)(int)f;
}
}  // namespace

int orc_partition_loci_by_approximate_depth(int64_t tasks, const guac_locus_range* loci, size_t n_loci, int64_t accuracy,
                                            const guac_read_batch* const* batches, size_t n_batches, guac_locus_range* out,
                                            size_t max_out, size_t* n_out) {
  return guarded([&] {
    if (tasks < 1 || accuracy < 1 || n_batches < 1) fail(GUAC_ERR_INVALID_ARGUMENT, "tasks, accuracy and the number of region sets must be >= 1");
    int64_t count = 0;
    for (size_t i = 0; i < n_loci; ++i) count += loci[i].end - loci[i].start;
    if (count <= 0) fail(GUAC_ERR_INVALID_ARGUMENT, "assumption failed: lociUsed.count > 0");
    // Step (1): micro partitions
    const int64_t n_micro = (accuracy * tasks < count) ? accuracy * tasks : count;
    std::vector<guac_locus_range> micro((size_t)n_micro + n_loci + 8);
    size_t n_ranges = 0;
    if (orc_partition_loci_uniformly(n_micro, loci, n_loci, micro.data(), micro.size(), &n_ranges) != 0) fail(GUAC_ERR_INVALID_ARGUMENT, g_last_error);
    micro.resize(n_ranges);
    // Step (2): regions overlapping each micro partition
    std::vector<long long> counts((size_t)n_micro, 0);
    for (size_t b = 0; b < n_batches; ++b) {
      ReadSet rs;
      load_batch(batches[b], rs);
      for (const Read& r : rs.reads) {
        long long last = -1;  // getAll(start, end): the SET of micro partitions with a locus in [start, end)
        std::vector<long long> seen;
        for (const guac_locus_range& m : micro)
          if (m.contig == r.contig && m.start < r.end && m.end > r.start && r.end > r.start) {
            if ((long long)m.task != last && std::find(seen.begin(), seen.end(), (long long)m.task) == seen.end()) {
              seen.push_back(m.task);
              counts[(size_t)m.task] += 1;
            }
            last = m.task;
          }
      }
    }
    // Step (3): greedy assignment
    long long total = 0;
    for (long long c : counts) total += c;
    const double regions_per_task = std::max(1.0, (double)total / (double)tasks);
    std::vector<guac_locus_range> v;
    auto put = [&](int32_t contig, int64_t start, int64_t end, long long task) {  // LociMap.Builder.put coalesces
      if (end <= start) return;
      if (!v.empty() && v.back().contig == contig && v.back().task == (int32_t)task && v.back().end == start) v.back().end = end;
      else v.push_back(guac_locus_range{contig, (int32_t)task, start, end});
    };
    double regions_assigned = 0.0;
    long long task = 0;
    auto remaining_for_task = [&]() { return scala_round_double(((double)(task + 1) * regions_per_task) - regions_assigned); };
    size_t next_range = 0;
    for (long long mt = 0; mt < n_micro; ++mt) {
      // microPartitions.asInverseMap(mt): the ranges of this micro partition, in contig order
      std::vector<guac_locus_range> set;
      while (next_range < micro.size() && micro[next_range].task == (int32_t)mt) set.push_back(micro[next_range++]);
      long long regions_in_set = counts[(size_t)mt];
      auto set_count = [&]() { long long c = 0; for (auto& x : set) c += x.end - x.start; return c; };
      while (!set.empty()) {
        if (regions_in_set == 0) {
          for (auto& x : set) put(x.contig, x.start, x.end, task);
          set.clear();
        } else {
          if (remaining_for_task() == 0) task += 1;
          if (!(remaining_for_task() > 0) || !(task < tasks)) fail(GUAC_ERR_INVALID_ARGUMENT, "assertion failed in partitionLociByApproximateDepth");
          const double fraction = std::min(1.0, (double)remaining_for_task() / (double)regions_in_set);
          long long loci_to_take = std::max<long long>(1, (long long)(fraction * (double)set_count()));
          const long long regions_to_take = (long long)(fraction * (double)regions_in_set);
          // set.take(lociToTake)
          std::vector<guac_locus_range> rest;
          for (auto& x : set) {
            const long long len = x.end - x.start;
            if (loci_to_take >= len) { put(x.contig, x.start, x.end, task); loci_to_take -= len; }
            else if (loci_to_take > 0) { put(x.contig, x.start, x.start + loci_to_take, task); rest.push_back(guac_locus_range{x.contig, x.task, x.start + loci_to_take, x.end}); loci_to_take = 0; }
            else rest.push_back(x);
          }
          set.swap(rest);
          regions_assigned += (double)scala_round_long(regions_to_take);
          regions_in_set -= scala_round_long(regions_to_take);
        }
      }
    }
    *n_out = v.size();
    for (size_t i = 0; i < v.size() && i < max_out; ++i) out[i] = v[i];
  });
}

double orc_phred_to_success_probability(int phred) { return phred_tables().success[phred & 255]; }
int orc_success_probability_to_phred(double p) { return success_probability_to_phred(p); }

// SomaticGenotypeFilter.apply(Seq, ...) filters/SomaticGenotypeFilter.scala:310-335
int orc_somatic_genotype_filter(const guac_somatic_record* g, int min_tumor_read_depth, int max_tumor_read_depth,
                                int min_normal_read_depth, int min_tumor_alternate_read_depth, int min_log_odds,
                                int min_vaf, int min_likelihood) {
  (void)min_log_odds;  // the Seq overload does not apply SomaticLogOddsFilter
  // SomaticReadDepthFilter.withinReadDepthRange (upper bound exclusive, filters/GenotypeFilter.scala:63)
  bool ok = g->tumor.read_depth >= min_tumor_read_depth && g->tumor.read_depth < max_tumor_read_depth &&
            g->normal.read_depth >= min_normal_read_depth && g->normal.read_depth < std::numeric_limits<int>::max();
  // SomaticVAFFilter: variantAlleleFrequency (Float) * 100.0 > minVAF
  float vaf = (float)g->tumor.allele_read_depth / (float)g->tumor.read_depth;
  ok = ok && ((double)vaf * 100.0 > (double)min_vaf);
  ok = ok && (g->phred_scaled_somatic_likelihood >= min_likelihood);
  if (min_tumor_alternate_read_depth > 0) ok = ok && (g->tumor.allele_read_depth >= min_tumor_alternate_read_depth);
  return ok ? 1 : 0;
}

}  // extern "C"
