/* A plain-C consumer of include/guac.h that drives the HOT PATH on a GPU: what a JNI / cgo / Panama shim does.
 * usage: gpu_driver <columns.bin>     (written by tests/test_gpu_c_driver.py: one read batch as raw columns)
 * Packs the batch (guac_reads_pack), runs guac_germline_threshold over loci "all" at threshold 8, and prints the records — in
 * canonical order, once from the compact form as it crossed the bus and once from the guac_threshold_record view — plus the
 * depth histogram's total, for the Python test to compare with the oracle.  No Python, no torch in this process. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "guac.h"

#define CHECK(cond)                                                \
  do {                                                             \
    if (!(cond)) {                                                 \
      fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                    \
    }                                                              \
  } while (0)

static void* slurp(FILE* f, uint64_t bytes) {
  void* p = malloc(bytes ? bytes : 1);
  if (p && bytes && fread(p, 1, bytes, f) != bytes) {
    free(p);
    return NULL;
  }
  return p;
}

int main(int argc, char** argv) {
  CHECK(argc == 2);
  FILE* f = fopen(argv[1], "rb");
  CHECK(f != NULL);
  uint64_t hdr[6]; /* n_reads, n_contigs, n_cigar_ops, n_bases, n_md_bytes, has_sample */
  CHECK(fread(hdr, sizeof hdr[0], 6, f) == 6);
  const uint64_t n = hdr[0], nc = hdr[1], n_ops = hdr[2], n_bases = hdr[3], n_md = hdr[4];
  guac_read_batch b;
  memset(&b, 0, sizeof b);
  b.n_reads = n;
  b.n_contigs = (uint32_t)nc;
  CHECK((b.contig_length = slurp(f, nc * 8)) != NULL);
  CHECK((b.contig = slurp(f, n * 4)) != NULL);
  CHECK((b.start = slurp(f, n * 8)) != NULL);
  CHECK((b.cigar_off = slurp(f, (n + 1) * 8)) != NULL);
  CHECK((b.cigar = slurp(f, n_ops * 4)) != NULL);
  CHECK((b.seq_off = slurp(f, (n + 1) * 8)) != NULL);
  CHECK((b.seq = slurp(f, n_bases)) != NULL);
  CHECK((b.qual = slurp(f, n_bases)) != NULL);
  CHECK((b.mapq = slurp(f, n)) != NULL);
  CHECK((b.flags = slurp(f, n)) != NULL);
  CHECK((b.md_off = slurp(f, (n + 1) * 8)) != NULL);
  CHECK((b.md = slurp(f, n_md)) != NULL);
  fclose(f);

  guac_ctx* ctx = NULL;
  CHECK(guac_ctx_create(0, &ctx) == GUAC_OK);
  guac_reads* reads = NULL;
  guac_status s = guac_reads_pack(ctx, &b, NULL, &reads);
  if (s != GUAC_OK) fprintf(stderr, "pack: %s: %s\n", guac_status_string(s), guac_last_error(ctx));
  CHECK(s == GUAC_OK);
  CHECK(guac_reads_count(reads) == n);

  /* LociSet "all": every contig but its last base (LociSet.scala:205-207) */
  guac_locus_range* ranges = calloc(nc ? nc : 1, sizeof *ranges);
  CHECK(ranges != NULL);
  size_t n_ranges = 0;
  for (uint64_t c = 0; c < nc; ++c)
    if (b.contig_length[c] - 1 > 0) {
      ranges[n_ranges].contig = (int32_t)c;
      ranges[n_ranges].start = 0;
      ranges[n_ranges].end = b.contig_length[c] - 1;
      ++n_ranges;
    }
  guac_threshold_params prm = {8, 0, 0, 1};
  guac_result* res = NULL;
  s = guac_germline_threshold(ctx, reads, ranges, n_ranges, &prm, &res);
  if (s != GUAC_OK) fprintf(stderr, "call: %s: %s\n", guac_status_string(s), guac_last_error(ctx));
  CHECK(s == GUAC_OK);

  const guac_compact_record* compact = NULL;
  const guac_threshold_record* general = NULL;
  size_t n_general = 0;
  int32_t sample = -1;
  const size_t n_compact = guac_result_compact_records(res, &compact, &general, &n_general, &sample);
  CHECK(n_compact + n_general == guac_result_n(res));
  size_t n_bytes = 0;
  const uint8_t* pool = guac_result_bytes(res, &n_bytes);
  const guac_threshold_record* rec = guac_result_threshold_records(res);
  const guac_stats* st = guac_result_stats(res);
  printf("records %zu compact %zu general %zu visited %llu launches %llu\n", guac_result_n(res), n_compact, n_general,
         (unsigned long long)st->loci_visited, (unsigned long long)st->kernel_launches);
  for (size_t i = 0; i < guac_result_n(res); ++i) {
    CHECK((size_t)rec[i].ref_off + rec[i].ref_len <= n_bytes && (size_t)rec[i].alt_off + rec[i].alt_len <= n_bytes);
    printf("R %d %lld %.*s %.*s %u %u\n", rec[i].contig, (long long)rec[i].start, (int)rec[i].ref_len, (const char*)pool + rec[i].ref_off,
           (int)rec[i].alt_len, (const char*)pool + rec[i].alt_off, rec[i].gt[0], rec[i].gt[1]);
  }
  for (size_t i = 0; i < n_compact; ++i) { /* the compact form: contig 63..48 | start 47..16 | alt 15..13 | ref 12..11 | gt 10..7 */
    const uint64_t v = compact[i];
    const unsigned alt = (unsigned)(v >> 13) & 7u, ref = (unsigned)(v >> 11) & 3u;
    CHECK(i == 0 || compact[i - 1] <= v); /* canonical order */
    printf("C %u %u %c %s %u %u\n", (unsigned)(v >> 48), (unsigned)(uint32_t)(v >> 16), "ACGT"[ref],
           alt == 0 ? "<ALT>" : alt == 1 ? "A" : alt == 2 ? "C" : alt == 3 ? "G" : "T", (unsigned)(v >> 9) & 3u, (unsigned)(v >> 7) & 3u);
  }
  uint64_t hist[GUAC_DEPTH_BINS], covered = 0, total = 0;
  CHECK(guac_depth_histogram(ctx, reads, ranges, n_ranges, hist) == GUAC_OK);
  for (int d = 0; d < GUAC_DEPTH_BINS; ++d) {
    total += hist[d];
    if (d) covered += hist[d];
  }
  printf("loci %llu covered %llu\n", (unsigned long long)total, (unsigned long long)covered);
  CHECK(covered == st->loci_visited);
  guac_result_free(res);
  guac_reads_free(reads);
  guac_ctx_destroy(ctx);
  printf("gpu_driver ok\n");
  return 0;
}
