// guac_somatic.cuh — somatic-standard caller (placeholder until the kernels land)
#pragma once
#include "guac_host.cuh"
namespace {
void somatic_init_tables(guac_ctx*) {}
void run_somatic(guac_ctx*, const guac_reads&, const guac_reads&, const guac_locus_range*, size_t, const guac_somatic_params&, guac_result&) {
  fail(GUAC_ERR_UNSUPPORTED, "somatic-standard kernels not built yet");
}
}  // namespace
