/*
 * guac_synth.h — deterministic synthetic read generator for the benchmark shapes of BASELINE.json: start-sorted reads with
 * consistent CIGAR and MD tags.  Measurement utility, two builds of ONE generator (csrc/guac_synth_core.h,
 * integer arithmetic only, same bytes from both):
 *   libguac_synth.so (host only)  guac_synth_generate: a guac_read_batch in host memory — what the CPU oracle, the tests and
 *                                 the end-to-end bench leg consume;
 *   libguac_b200.so (device)      guac_synth_generate_device: the same columns generated straight into HBM (the whole-genome
 *                                 shape would otherwise be bound by PCIe), packed with guac_reads_pack_device.
 * The number of reads starting at a locus is Poisson(reads_per_locus), a pure function of (seed, sample, contig, locus), and a
 * read is a pure function of (seed, sample, contig, start, rank at that start): a shard generated alone holds exactly the
 * reads a whole-genome run would place there.
 */
#ifndef GUAC_SYNTH_H_
#define GUAC_SYNTH_H_

#include "guac.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct guac_synth_params {
  uint64_t seed;
  uint32_t n_contigs;
  int32_t read_length;
  const int64_t* contig_length;     /* [n_contigs] */
  double reads_per_locus;           /* expected reads starting per locus = depth / read_length */
  int32_t sample;                   /* 0 = normal / germline sample, 1 = tumor (carries the somatic SNVs) */
  uint32_t n_windows;               /* 0 = every contig whole */
  const guac_locus_range* windows;  /* [n_windows] (contig, start, end): only reads STARTING there; ascending, disjoint */
  double frac_clip, frac_ins, frac_del, frac_both;
  int32_t n_threads;                /* host build: 0 = all host threads */
  int32_t with_qualities;           /* device build: 0 = do not generate base qualities (germline-threshold reads none) */
} guac_synth_params;

typedef struct guac_synth_batch guac_synth_batch;

void guac_synth_default_params(guac_synth_params* p);
int guac_synth_generate(const guac_synth_params* p, guac_synth_batch** out);
const guac_read_batch* guac_synth_batch_view(const guac_synth_batch* b);
void guac_synth_batch_free(guac_synth_batch* b);

/* ---- device build (exported by libguac_b200.so) ---------------------------------------------------------------------- */
typedef struct guac_synth_device_batch guac_synth_device_batch;
/* Generates the batch into the memory of ctx's device.  The view's column pointers are DEVICE pointers (n_reads, n_contigs and
 * contig_length are host-side): hand it to guac_reads_pack_device. */
guac_status guac_synth_generate_device(guac_ctx* ctx, const guac_synth_params* p, guac_synth_device_batch** out);
const guac_read_batch* guac_synth_device_batch_view(const guac_synth_device_batch* b);
double guac_synth_device_batch_ms(const guac_synth_device_batch* b);   /* device time of the generator kernels */
/* totals of the variable-length columns (the last offsets, which live on the device) */
void guac_synth_device_batch_totals(const guac_synth_device_batch* b, uint64_t* n_cigar_ops, uint64_t* n_md_bytes, uint64_t* n_bases);
/* guac_reads_pack_device that takes the batch's large columns over instead of copying them (the whole-genome shards would not
 * fit twice); afterwards the batch can only be freed. */
guac_status guac_reads_pack_synth(guac_ctx* ctx, guac_synth_device_batch* b, const guac_reference* ref, guac_reads** out);
void guac_synth_device_batch_free(guac_synth_device_batch* b);
/* Copies the columns into host memory (page-locked when `pinned`): the end-to-end bench leg starts there. */
typedef struct guac_synth_host_batch guac_synth_host_batch;
guac_status guac_synth_device_batch_download(guac_ctx* ctx, const guac_synth_device_batch* b, int pinned, guac_synth_host_batch** out);
const guac_read_batch* guac_synth_host_batch_view(const guac_synth_host_batch* b);
void guac_synth_host_batch_free(guac_synth_host_batch* b);

#ifdef __cplusplus
}
#endif
#endif
