/*
 * guac_oracle.h — C interface of the CPU ORACLE (test infrastructure, NOT the product).
 *
 * The oracle is a literal CPU restatement of the reference's (MartijnAB/guacamole, Scala) pileup-and-call
 * path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * It shares the plain-C data contract of include/guac.h (same input batch, same record structs) so that a
 * parity test is "run both, memcmp the records".
 *
 * Parity pinning: checked against the golden vectors of the reference's own suites (see tests/test_oracle_*.py);
 * end-to-end chrM output is NOT pinned by the reference (no golden VCF exists) — see DESIGN.md.
 */
#ifndef GUAC_ORACLE_H_
#define GUAC_ORACLE_H_

#include "../include/guac.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_result orc_result;

/* Alignment kinds (pileup/Alignment.scala:44-94) */
#define ORC_MATCH 0
#define ORC_MISMATCH 1
#define ORC_INSERTION 2
#define ORC_DELETION 3
#define ORC_MID_DELETION 4
#define ORC_CLIPPED 5

/* One PileupElement (pileup/PileupElement.scala:40-47) with its lazily derived fields evaluated. */
typedef struct orc_element {
  int64_t read_index;
  int32_t kind;
  int32_t quality_score;          /* PileupElement.qualityScore :166-171 */
  int32_t read_position;
  int32_t cigar_element_index;
  int32_t index_within_cigar_element;
  uint32_t ref_off, ref_len;      /* alignment.referenceBases, into the byte pool */
  uint32_t seq_off, seq_len;      /* alignment.sequencedBases */
  uint8_t is_positive_strand;
  uint8_t pad_[3];
} orc_element;

typedef struct orc_genotype_likelihood {
  uint32_t a1_ref_off, a1_ref_len, a1_alt_off, a1_alt_len;
  uint32_t a2_ref_off, a2_ref_len, a2_alt_off, a2_alt_len;
  double value;
} orc_genotype_likelihood;

const char* orc_last_error(void);

/* pileupFlatMap(reads, ranges, skipEmpty, callVariantsAtLocus) — DistributedUtil.scala:288-306 +
 * GermlineThresholdCaller.scala:90-179.  Ranges with different `task` ids are run as independent tasks
 * (fresh windows / pileup state, DistributedUtil.scala:404-414) on up to n_threads threads. */
int orc_germline_threshold(const guac_read_batch* batch, const guac_reference* ref, const guac_locus_range* ranges,
                           size_t n_ranges, const guac_threshold_params* params, int n_threads, orc_result** out);
/* pileupFlatMapTwoRDDs(tumor, normal, ...) + findPotentialVariantAtLocus — SomaticStandardCaller.scala:162-245 */
int orc_somatic_standard(const guac_read_batch* tumor, const guac_read_batch* normal, const guac_reference* ref,
                         const guac_locus_range* ranges, size_t n_ranges, const guac_somatic_params* params,
                         int n_threads, orc_result** out);
int orc_pileup_counts(const guac_read_batch* batch, const guac_reference* ref, const guac_locus_range* ranges,
                      size_t n_ranges, int skip_empty, int n_threads, orc_result** out);

size_t orc_result_n(const orc_result* r);
const guac_threshold_record* orc_result_threshold_records(const orc_result* r);
const guac_somatic_record* orc_result_somatic_records(const orc_result* r);
/* pileupFlatMap(reads, ...) + GermlineStandard.Caller.callVariantsAtLocus — GermlineStandardCaller.scala:66-70, 90-124.
 * PARITY UNPINNED end to end: the reference has no test of this caller; its parts (likelihoods, AlleleEvidence, pileup)
 * are pinned by LikelihoodSuite / AlleleEvidenceSuite / PileupSuite. */
int orc_germline_standard(const guac_read_batch* batch, const guac_reference* ref, const guac_locus_range* ranges,
                          size_t n_ranges, const guac_standard_params* params, int n_threads, orc_result** out);
const guac_called_allele* orc_result_called_alleles(const orc_result* r);
/* pileupFlatMap(reads, ..., true, VariantSupport.Caller.pileupToAlleleCounts) — VariantSupport.scala:93-98, 110-118 */
int orc_allele_counts(const guac_read_batch* batch, const guac_reference* ref, const guac_locus_range* ranges,
                      size_t n_ranges, int n_threads, orc_result** out);
const guac_allele_count* orc_result_allele_counts(const orc_result* r);
const guac_locus_counts* orc_result_counts(const orc_result* r);
const orc_element* orc_result_elements(const orc_result* r);
const orc_genotype_likelihood* orc_result_likelihoods(const orc_result* r);
const uint8_t* orc_result_bytes(const orc_result* r, size_t* n_bytes);
const guac_stats* orc_result_stats(const orc_result* r);
uint8_t orc_result_reference_base(const orc_result* r);
void orc_result_free(orc_result* r);

/* ---- unit-level probes used by the golden-vector tests -------------------------------------------------- */
/* Pileup(reads, contig, locus) (pileup/Pileup.scala:181-186): reads in batch order, overlapping ones kept,
 * reference base = referenceBaseAtLocus(overlapping reads) unless reference_base >= 0 is forced. */
int orc_pileup_at(const guac_read_batch* batch, int32_t contig, int64_t locus, int reference_base, orc_result** out);
/* MDTagUtils.getReference(read, allowNBase=true) reads/MDTagUtils.scala:23-78 */
int orc_md_reference(const guac_read_batch* batch, uint64_t read_index, uint8_t* out, size_t max_out, size_t* n_out,
                     int* n_mismatches);
/* Likelihood.likelihoodsOfAllPossibleGenotypesFromPileup on Pileup(reads, contig, locus)
 * (likelihood/Likelihood.scala:99-113) */
int orc_likelihoods_at(const guac_read_batch* batch, int32_t contig, int64_t locus, int include_alignment,
                       int log_space, int normalize, orc_result** out);
/* findPotentialVariantAtLocus(Pileup(tumor,contig,locus), Pileup(normal,contig,locus), ...) */
int orc_somatic_at(const guac_read_batch* tumor, const guac_read_batch* normal, int32_t contig, int64_t locus,
                   const guac_somatic_params* params, orc_result** out);
/* callVariantsAtLocus(Pileup(reads, contig, locus), ...) */
int orc_threshold_at(const guac_read_batch* batch, int32_t contig, int64_t locus, const guac_threshold_params* params,
                     orc_result** out);
/* AlleleEvidence(likelihood, Allele(ref, alt), Pileup(reads, contig, locus)) variants/AlleleEvidence.scala:58-101 */
int orc_allele_evidence_at(const guac_read_batch* batch, int32_t contig, int64_t locus, const uint8_t* ref,
                           size_t ref_len, const uint8_t* alt, size_t alt_len, double likelihood,
                           guac_allele_evidence* out);
/* SlidingWindow + advanceMultipleWindows over up to 2 read sets: writes the visited loci
 * (windowing/SlidingWindow.scala:149-187) and, per visited locus, the number of current regions per window. */
int orc_visited_loci(const guac_read_batch* a, const guac_read_batch* b, const guac_locus_range* ranges, size_t n_ranges,
                     int skip_empty, int64_t half_window, int64_t* loci_out, int32_t* count_a_out, int32_t* count_b_out,
                     size_t max_out, size_t* n_out);
/* partitionLociUniformly DistributedUtil.scala:83-108 */
int orc_partition_loci_uniformly(int64_t tasks, const guac_locus_range* loci, size_t n_loci, guac_locus_range* out,
                                 size_t max_out, size_t* n_out);
/* partitionLociByApproximateDepth DistributedUtil.scala:162-251 (loci in LociSet order; regions = the reads of the batches) */
int orc_partition_loci_by_approximate_depth(int64_t tasks, const guac_locus_range* loci, size_t n_loci, int64_t accuracy,
                                            const guac_read_batch* const* batches, size_t n_batches, guac_locus_range* out,
                                            size_t max_out, size_t* n_out);
/* ADAM PhredUtils (third party, restated): */
double orc_phred_to_success_probability(int phred);
int orc_success_probability_to_phred(double p);
/* SomaticGenotypeFilter.apply(Seq, ...) filters/SomaticGenotypeFilter.scala:310-335: returns 1 if the record passes */
int orc_somatic_genotype_filter(const guac_somatic_record* rec, int min_tumor_read_depth, int max_tumor_read_depth,
                                int min_normal_read_depth, int min_tumor_alternate_read_depth, int min_log_odds,
                                int min_vaf, int min_likelihood);

#ifdef __cplusplus
}
#endif
#endif
