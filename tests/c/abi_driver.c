/* A plain-C consumer of include/guac.h: what a JNI / cgo / Panama shim would compile against.  Built and run by
 * tests/test_abi.py (no GPU needed): the header must be valid C, the record layouts must have the documented sizes, the
 * host-only entry points must work and the engine must refuse to start without a device (no CPU fallback). */
#include <stdio.h>
#include <string.h>

#include "guac.h"

#define CHECK(cond)                                                \
  do {                                                             \
    if (!(cond)) {                                                 \
      fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                    \
    }                                                              \
  } while (0)

int main(void) {
  CHECK(sizeof(guac_threshold_record) == 32);
  CHECK(sizeof(guac_allele_count) == 32);
  CHECK(sizeof(guac_allele_evidence) == 64);
  CHECK(sizeof(guac_somatic_record) == 168);
  CHECK(sizeof(guac_called_allele) == 96);
  CHECK(sizeof(guac_locus_counts) == 48);
  CHECK(sizeof(guac_locus_range) == 24);
  CHECK(sizeof(guac_stats) == 120);
  CHECK(strcmp(guac_status_string(GUAC_OK), "GUAC_OK") == 0);

  /* DistributedUtilSuite.scala:46-50: partitionLociUniformly(4, chrM:0-16571) */
  guac_locus_range loci = {0, 0, 0, 16571}, parts[8];
  size_t n = 0;
  memset(parts, 0, sizeof parts);
  CHECK(guac_partition_loci_uniformly(4, &loci, 1, parts, 8, &n) == GUAC_OK);
  CHECK(n == 4);
  CHECK(parts[0].start == 0 && parts[0].end == 4143 && parts[1].end == 8286 && parts[2].end == 12428 && parts[3].end == 16571);
  CHECK(parts[3].task == 3);

  /* no device in this process (the test hides the GPUs): the engine must say so, not compute on the CPU */
  guac_ctx* ctx = NULL;
  const guac_status s = guac_ctx_create(0, &ctx);
  if (s == GUAC_OK) {  /* a GPU is visible after all: fine, just close it again */
    guac_ctx_destroy(ctx);
    printf("abi_driver ok (device present)\n");
    return 0;
  }
  CHECK(s == GUAC_ERR_NO_DEVICE);
  CHECK(ctx == NULL);
  printf("abi_driver ok\n");
  return 0;
}
