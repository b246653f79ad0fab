/*
 * guac_synth.h — deterministic synthetic read generator (libguac_synth.so, host only) for the benchmark shapes of
 * BASELINE.json: produces a guac_read_batch of start-sorted reads with consistent CIGAR and MD tags.
 * Measurement utility: both bench arms (the CUDA engine and the CPU oracle) consume the same generated batch.
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
  const int64_t* contig_length;  /* [n_contigs] */
  uint64_t n_reads;              /* total, spread over the contigs (or the window) in proportion to their loci */
  int32_t sample;                /* 0 = normal / germline sample, 1 = tumor (carries the somatic SNVs) */
  int32_t window_contig;         /* optional window: only reads starting in [window_start, window_end) of this contig */
  int64_t window_start;
  int64_t window_end;            /* window_end <= window_start: no window */
  double frac_clip, frac_ins, frac_del, frac_both;
  int32_t n_threads;             /* 0 = all host threads */
  int32_t pad_;
} guac_synth_params;

typedef struct guac_synth_batch guac_synth_batch;

void guac_synth_default_params(guac_synth_params* p);
int guac_synth_generate(const guac_synth_params* p, guac_synth_batch** out);
const guac_read_batch* guac_synth_batch_view(const guac_synth_batch* b);
void guac_synth_batch_free(guac_synth_batch* b);

#ifdef __cplusplus
}
#endif
#endif
