/*
 * guac.h — C ABI of the B200 pileup-and-call engine (libguac_b200.so).
 *
 * This is the drop-in boundary for ONE path of the reference (MartijnAB/guacamole, Scala/Spark):
 * "pileupFlatMap ∘ caller closure".  The reference has no FFI for this path; its boundary is the Scala
 * closure API
 *     DistributedUtil.pileupFlatMap[T](reads, lociPartitions, skipEmpty, Pileup => Iterator[T], reference)
 *         (src/main/scala/org/hammerlab/guacamole/DistributedUtil.scala:288-306)
 *     DistributedUtil.pileupFlatMapTwoRDDs[T](reads1, reads2, lociPartitions, skipEmpty, (Pileup,Pileup) => Iterator[T], ref)
 *         (DistributedUtil.scala:316-335)
 * A JVM closure cannot run on a GPU, so the ABI is cut one level up: one fused entry point per caller
 * closure that the reference ships on this path:
 *     GermlineThreshold.Caller.callVariantsAtLocus   (commands/GermlineThresholdCaller.scala:90-179)
 *     SomaticStandard.Caller.findPotentialVariantAtLocus (commands/SomaticStandardCaller.scala:162-245)
 * plus a generic per-locus histogram entry point for closures that only need counts
 * (Pileup.depth / positiveDepth / referenceDepth, pileup/Pileup.scala:76-91).
 *
 * All functions are extern "C", take plain pointers and sizes, never call back into the host, never abort():
 * they return a guac_status and leave a message retrievable with guac_last_error().
 * INTEGRATION.md shows the Scala-side (Panama / JNI) binding a maintainer would add.
 */
#ifndef GUAC_H_
#define GUAC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GUAC_ABI_VERSION 2

/* ---- status codes.  Each maps to the Scala exception the reference throws on this path. -------------- */
typedef enum guac_status {
  GUAC_OK = 0,
  GUAC_ERR_INVALID_ARGUMENT = 1,        /* null pointer, bad sizes, unsorted/overlapping loci ranges            */
  GUAC_ERR_UNSORTED_READS = 2,          /* "Regions must be sorted by start locus" windowing/SlidingWindow.scala:56 */
  GUAC_ERR_CONTIG_ORDER = 3,            /* "Regions are not sorted by contig" DistributedUtil.scala:662-664      */
  GUAC_ERR_INVALID_CIGAR = 4,           /* InvalidCigarElementException pileup/PileupElement.scala:106,277;
                                           deletion preceded by non-M/=/X (AssertionError :118-122); P operator  */
  GUAC_ERR_MISSING_MD = 5,              /* ReferenceWithoutMDTagException reads/MappedRead.scala:58-59;
                                           NoSuchElementException from mdTag.deletions(..) PileupElement.scala:112,117;
                                           CigarMDTagMismatchException MappedRead.scala:63 */
  GUAC_ERR_MULTIPLE_REFERENCE_BASES = 6,/* IllegalArgumentException commands/GermlineThresholdCaller.scala:171-174 */
  GUAC_ERR_BAD_QUALITY = 7,             /* quality byte > 127 would index PhredUtils' table negatively            */
  GUAC_ERR_CUDA = 8,                    /* any CUDA runtime error (message has the cudaError string)              */
  GUAC_ERR_OOM = 9,                     /* host or device allocation failed                                       */
  GUAC_ERR_NO_DEVICE = 10,              /* no CUDA device / kernels not built for it: there is NO CPU fallback    */
  GUAC_ERR_UNSUPPORTED = 11             /* e.g. more than GUAC_MAX_SAMPLES samples in one read set                */
} guac_status;

/* ---- input: one columnar batch of MappedReads (reads/MappedRead.scala:35-48), caller-owned ------------- */
/* flags bits (reads/Read.scala:227-232, 253-267) */
#define GUAC_READ_POSITIVE_STRAND 0x01u  /* isPositiveStrand = !SAM flag 0x10 */
#define GUAC_READ_DUPLICATE       0x02u  /* isDuplicate (SAM 0x400); informational — input filters run host-side */
#define GUAC_READ_FAILED_QC       0x04u  /* failedVendorQualityChecks (SAM 0x200) */
#define GUAC_READ_HAS_MD          0x08u  /* mdTagOpt.isDefined */
#define GUAC_READ_PAIRED          0x10u

/* CIGAR ops use the BAM encoding (len << 4 | op), op index into "MIDNSHP=X" (htsjdk CigarOperator). */
#define GUAC_CIGAR_M 0u
#define GUAC_CIGAR_I 1u
#define GUAC_CIGAR_D 2u
#define GUAC_CIGAR_N 3u
#define GUAC_CIGAR_S 4u
#define GUAC_CIGAR_H 5u
#define GUAC_CIGAR_P 6u
#define GUAC_CIGAR_EQ 7u
#define GUAC_CIGAR_X 8u

typedef struct guac_read_batch {
  uint64_t n_reads;
  uint32_t n_contigs;            /* contig indices are 0..n_contigs-1                                           */
  const int64_t* contig_length;  /* [n_contigs] (may be NULL: then no upper bound is checked)                    */
  const int32_t* contig;         /* [n] contig index; reads must be sorted by (contig, start)                    */
  const int64_t* start;          /* [n] 0-based inclusive (MappedRead.start)                                     */
  const uint64_t* cigar_off;     /* [n+1] offsets into cigar[]                                                   */
  const uint32_t* cigar;         /* BAM-encoded ops                                                              */
  const uint64_t* seq_off;       /* [n+1] offsets into seq[] and qual[] (sequence.length == baseQualities.length)*/
  const uint8_t* seq;            /* ASCII bases, any byte allowed (Bases.scala)                                  */
  const uint8_t* qual;           /* numeric phred (no +33), each <= 127                                          */
  const uint8_t* mapq;           /* [n] alignmentQuality 0..255                                                  */
  const uint8_t* flags;          /* [n] GUAC_READ_* bits                                                         */
  const int32_t* sample;         /* [n] sample index (sampleName interned by the shim), or NULL = all sample 0   */
  const uint64_t* md_off;        /* [n+1] offsets into md[]; empty string + !HAS_MD = no MD tag                  */
  const char* md;                /* MD tag strings, not NUL-terminated                                           */
} guac_read_batch;

/* The same reads in the compact form they have in a BAM record (what Read.fromSAMRecord, reads/Read.scala:217-291, decodes
 * FROM): half the bytes of guac_read_batch cross PCIe.  Bases are 4-bit codes into "=ACMGRSVTWYHKDBN" (SAM spec 4.2.3),
 * packed back to back over the whole batch: base g (= seq_off[i] + j for base j of read i) is the HIGH nibble of
 * seq4[g / 2] when g is even, the low nibble when odd.  Starts are 32-bit, offsets 32-bit (a batch holds fewer than 2^32
 * bases, CIGAR ops and MD bytes: split larger ones), reads of one contig are one run of the batch (contig_read_off) and the
 * batch holds one sample.  Bases outside the sixteen letters (lower case, '.', ...) need guac_read_batch. */
typedef struct guac_read_batch_v2 {
  uint64_t n_reads;
  uint32_t n_contigs;
  uint32_t read_length;            /* != 0: every read holds exactly this many bases; seq_off is then ignored (may be NULL) */
  const int64_t* contig_length;    /* [n_contigs] or NULL                                                           */
  const uint64_t* contig_read_off; /* [n_contigs + 1] reads of contig c are [off[c], off[c+1]); off[n_contigs] = n  */
  const int32_t* start;            /* [n] 0-based inclusive, ascending within a contig                              */
  const uint32_t* cigar_off;       /* [n+1]                                                                         */
  const uint32_t* cigar;           /* BAM-encoded ops                                                               */
  const uint32_t* seq_off;         /* [n+1] in bases (or NULL with read_length)                                     */
  const uint8_t* seq4;             /* (bases + 1) / 2 bytes                                                         */
  const uint8_t* qual;             /* one byte per base, numeric phred; NULL allowed when GUAC_OPT_PACK_QUALITIES=0 */
  const uint8_t* mapq;             /* [n]                                                                           */
  const uint8_t* flags;            /* [n] GUAC_READ_* bits                                                          */
  const uint32_t* md_off;          /* [n+1]                                                                         */
  const char* md;
  int32_t sample;                  /* the batch's sample index                                                      */
  int32_t reserved;
} guac_read_batch_v2;

/* optional FASTA-derived reference (ReferenceGenome.getReferenceBase, DistributedUtil.scala:266) */
typedef struct guac_reference {
  uint32_t n_contigs;
  const uint64_t* base_off;      /* [n_contigs+1] offsets into bases[] */
  const uint8_t* bases;          /* ASCII, upper-case */
} guac_reference;

typedef struct guac_locus_range {
  int32_t contig;
  int32_t task;                  /* LociMap[Long] value = task / GPU shard id (informational on one GPU)        */
  int64_t start;                 /* inclusive */
  int64_t end;                   /* exclusive */
} guac_locus_range;

/* ---- parameters: 1:1 with the reference's args4j flags ------------------------------------------------- */
typedef struct guac_threshold_params {   /* commands/GermlineThresholdCaller.scala:42-51 */
  int32_t threshold_percent;     /* --threshold, default 8 */
  int32_t emit_ref;              /* --emit-ref */
  int32_t emit_no_call;          /* --emit-no-call */
  int32_t skip_empty;            /* pileupFlatMap skipEmpty (the caller passes true) */
} guac_threshold_params;

typedef struct guac_somatic_params {     /* commands/SomaticStandardCaller.scala:49-60, filters/PileupFilter.scala:48-59 */
  int32_t odds_threshold;        /* --odds, default 20 */
  int32_t min_alignment_quality; /* --min-mapq, default 1 */
  int32_t filter_multi_allelic;  /* --filter-multi-allelic */
  int32_t max_read_depth;        /* --max-tumor-read-depth, default INT32_MAX */
  int32_t skip_empty;
} guac_somatic_params;

typedef struct guac_standard_params {    /* commands/GermlineStandardCaller.scala:64-70, 90-92 */
  int32_t min_alignment_quality; /* --min-mapq (PileupFilterArguments, default 1): elements of reads below it do not enter
                                    the likelihoods (QualityAlignedReadsFilter, filters/PileupElementsFilter.scala:48-50);
                                    the AlleleEvidence of a call is still taken over the unfiltered pileup (:120) */
  int32_t skip_empty;            /* the caller passes skipEmpty = true (:68) */
} guac_standard_params;

/* ---- output records ------------------------------------------------------------------------------------ */
/* bdg-formats GenotypeAllele */
#define GUAC_GT_REF 0u
#define GUAC_GT_ALT 1u
#define GUAC_GT_OTHER_ALT 2u
#define GUAC_GT_NO_CALL 3u

/* One bdg-formats Genotype as built by GermlineThresholdCaller.scala:106-117.
 * ref/alt bytes live in the result's byte pool at [ref_off, ref_off+ref_len) / [alt_off, ...). */
typedef struct guac_threshold_record {
  int64_t start;                 /* Variant.start = pileup.locus */
  int32_t contig;
  int32_t sample;
  uint32_t ref_off;
  uint32_t alt_off;
  uint16_t ref_len;
  uint16_t alt_len;
  uint8_t gt[2];
  uint8_t tie;                   /* 1 if an equal-count tie made the allele choice depend on Scala HashMap order */
  uint8_t pad_;
} guac_threshold_record;

/* Compact form of a guac_threshold_record whose reference and alternate are single bases (or the symbolic "<ALT>"): what the
 * tile kernel emits, 8 bytes instead of 32 across PCIe / NVLink.  Bits: contig 63..48 | start 47..16 | alt 15..13 (0 = "<ALT>",
 * 1..4 = A C G T) | ref 12..11 (0..3 = A C G T) | gt[0] 10..9 | gt[1] 8..7 | tie 6.  The numeric order of the values is the
 * canonical (contig, start, ref, alt) order; the sample is the read set's. */
typedef uint64_t guac_compact_record;

/* variants/AlleleEvidence.scala:41-50 */
typedef struct guac_allele_evidence {
  double likelihood;
  double mean_mapping_quality;
  double median_mapping_quality;
  double mean_base_quality;
  double median_base_quality;
  double median_mismatches_per_read;
  int32_t read_depth;
  int32_t allele_read_depth;
  int32_t forward_depth;
  int32_t allele_forward_depth;
} guac_allele_evidence;

/* variants/CalledSomaticAllele.scala:36-51 + the fields AlleleConversions.scala:47-62 derives from it */
typedef struct guac_somatic_record {
  int64_t start;
  int32_t contig;
  int32_t sample;                /* tumorPileup.sampleName */
  uint32_t ref_off;
  uint32_t alt_off;
  uint16_t ref_len;
  uint16_t alt_len;
  int32_t phred_scaled_somatic_likelihood;  /* genotypeQuality */
  double somatic_log_odds;
  guac_allele_evidence tumor;    /* tumorVariantEvidence */
  guac_allele_evidence normal;   /* normalReferenceEvidence */
} guac_somatic_record;

/* variants/CalledAllele.scala:34-43 as built by GermlineStandard.Caller.callVariantsAtLocus
 * (commands/GermlineStandardCaller.scala:113-121) + genotypeQuality of AlleleConversions.scala:30-45.  One record per
 * non-reference allele of the most likely genotype: a homozygous alternate call yields two equal records, as the reference
 * does (Genotype.getNonReferenceAlleles keeps both copies, variants/Genotype.scala:46-48). */
typedef struct guac_called_allele {
  int64_t start;
  int32_t contig;
  int32_t sample;
  uint32_t ref_off;
  uint32_t alt_off;
  uint16_t ref_len;
  uint16_t alt_len;
  int32_t phred_scaled_likelihood;  /* AlleleEvidence.phredScaledLikelihood (variants/AlleleEvidence.scala:52) */
  guac_allele_evidence evidence;    /* likelihood = exp(normalised log likelihood of the most likely genotype) */
} guac_called_allele;

/* commands/VariantSupport.scala:36-41 AlleleCount as built by pileupToAlleleCounts (:110-118): one row per distinct allele
 * of a non-empty pileup, count = elements carrying it.  Same layout as guac_threshold_record up to the last 4 bytes. */
typedef struct guac_allele_count {
  int64_t start;
  int32_t contig;
  int32_t sample;
  uint32_t ref_off;
  uint32_t alt_off;
  uint16_t ref_len;
  uint16_t alt_len;
  int32_t count;
} guac_allele_count;

/* per-locus histogram (guac_pileup_counts): one row per visited locus */
typedef struct guac_locus_counts {
  int64_t locus;
  int32_t contig;
  int32_t depth;                 /* Pileup.depth */
  int32_t positive_depth;        /* Pileup.positiveDepth */
  int32_t reference_depth;       /* Pileup.referenceDepth (Match elements only) */
  int32_t base_count[4];         /* Match/Mismatch elements whose sequenced base is A,C,G,T */
  int32_t other_count;           /* every other element: insertion / deletion / mid-deletion / clipped / non-ACGT base */
  uint8_t reference_base;        /* Pileup.referenceBase (ASCII) */
  uint8_t pad_[3];
} guac_locus_counts;

/* counters the reference keeps in Spark accumulators (DistributedUtil.scala:573-618) */
typedef struct guac_stats {
  uint64_t reads_total;
  uint64_t reads_relevant;       /* overlap >= 1 requested locus */
  uint64_t reads_expanded;       /* after duplication across task boundaries */
  uint64_t loci_requested;
  uint64_t loci_visited;         /* non-empty pileups when skip_empty */
  uint64_t records;
  uint64_t tie_loci;             /* SURVEY H1b */
  uint64_t order_sensitive_loci; /* MD-derived reference bases disagree */
  double kernel_ms;              /* device time of the pileup+call kernels of the last call (CUDA events) */
  uint64_t kernel_launches;      /* number of kernels launched by the last call */
  double tile_kernel_ms;         /* ... of which the bit-sliced tile kernel (K_tile) */
  double exact_kernel_ms;        /* ... of which the exact per-element kernel(s) (K_exact / likelihoods) */
  uint64_t exact_loci;           /* loci decided by the exact per-element kernel */
  uint64_t h2d_bytes;            /* bytes copied host -> device by the call */
  uint64_t d2h_bytes;            /* bytes copied device -> host by the call */
} guac_stats;

typedef struct guac_ctx guac_ctx;
typedef struct guac_reads guac_reads;
typedef struct guac_result guac_result;

/* ---- lifecycle ------------------------------------------------------------------------------------------ */
int guac_abi_version(void);
guac_status guac_ctx_create(int device, guac_ctx** out);
void guac_ctx_destroy(guac_ctx* ctx);
const char* guac_last_error(const guac_ctx* ctx);   /* valid until the next call on ctx */
const char* guac_status_string(guac_status s);

/* per-context options (defaults in brackets) */
#define GUAC_OPT_SORT_RECORDS 1    /* [1] return records in canonical (contig, start, sample, ref, alt) order; 0 = device
                                      order (the reference's own order is unspecified: coalesce(1, shuffle = true),
                                      Common.scala:293) */
#define GUAC_OPT_PACK_QUALITIES 2  /* [1] copy base qualities to the device at pack time; germline-threshold never reads
                                      them, somatic-standard needs them */
#define GUAC_OPT_HOST_THREADS 3     /* [0 = all] host threads guac_reads_pack may use for its header pass (set it to
                                      cores / ranks when several ranks share one box) */
#define GUAC_OPT_DIFFERENCE_LISTS 4  /* [1] guac_reads_pack also expands the reads into per-granule difference streams
                                      (every element that differs from the reference track, plus per-locus start / end
                                      counts): the germline kernels then do no per-read work.  0 = the call walks base planes
                                      and CIGARs itself (same results; the cross-check the parity tests run) */
#define GUAC_OPT_SEGMENTS 5          /* [1] 1..4: a germline call runs in this many segments of tiles, the exact kernel and the
                                      record egress of one overlapping the tile kernel of the next (same results) */
#define GUAC_OPT_TRIM_CACHE 6        /* (an action, any value) hand the device buffers cached from freed read sets / results back
                                      * to the driver now: between workloads of very different shapes on one context */
#define GUAC_OPT_PACK_OVERLAP 7      /* [1] guac_reads_pack of a large host batch finishes the reference track and the by-locus
                                      stores copy chunk by copy chunk (reads are start-sorted: everything in front of the next
                                      chunk's first read is final), underneath the copies still on the bus.  0 = after the last
                                      chunk, in one launch each (same store) */
guac_status guac_ctx_set_option(guac_ctx* ctx, int option, int64_t value);

/* Device-side stopwatch on the context's stream (CUDA events): start, run any number of calls, stop -> elapsed ms. */
guac_status guac_ctx_timer_start(guac_ctx* ctx);
guac_status guac_ctx_timer_stop(guac_ctx* ctx, double* elapsed_ms);

/* Page-lock caller-owned host buffers (e.g. direct ByteBuffers) so that guac_reads_pack copies at full PCIe speed. */
guac_status guac_host_register(void* ptr, size_t bytes);
guac_status guac_host_unregister(void* ptr);

/* Pack a host batch into the device SoA (copies; caller keeps ownership of `batch`).  Validates sortedness,
 * CIGAR/MD consistency and quality range — the checks SlidingWindow / MappedRead do lazily. `ref` may be NULL
 * (reference bases then come from MD tags, Pileup.referenceBaseAtLocus pileup/Pileup.scala:157-165). */
guac_status guac_reads_pack(guac_ctx* ctx, const guac_read_batch* batch, const guac_reference* ref, guac_reads** out);
/* Same, for a batch whose column pointers are DEVICE memory of ctx's device (n_reads, n_contigs and contig_length stay on the
 * host): no host -> device copies; the per-read checks and derived columns are computed by the same kernels either way. */
guac_status guac_reads_pack_device(guac_ctx* ctx, const guac_read_batch* device_batch, const guac_reference* ref, guac_reads** out);
/* Same as guac_reads_pack for the compact batch: the columns cross PCIe as they are and are widened on the device (the
 * nibbles to the ASCII bases the store keeps, chunk by chunk underneath the copies); identical store, identical records. */
guac_status guac_reads_pack_v2(guac_ctx* ctx, const guac_read_batch_v2* batch, const guac_reference* ref, guac_reads** out);
/* Host-side conversion guac_read_batch -> guac_read_batch_v2 (what a JVM shim does straight from SAMRecord; here for tests,
 * the C driver and the bench).  `pinned` != 0 page-locks the buffers.  GUAC_ERR_INVALID_ARGUMENT when a base is outside
 * "=ACMGRSVTWYHKDBN", the reads hold several samples or are not grouped by contig; GUAC_ERR_UNSUPPORTED when an offset or a
 * start needs more than 32 bits.  `fixed_length` != 0 stores read_length instead of seq_off when every read has that many bases. */
typedef struct guac_host_batch_v2 guac_host_batch_v2;
guac_status guac_read_batch_compact(const guac_read_batch* batch, int pinned, int fixed_length, guac_host_batch_v2** out);
const guac_read_batch_v2* guac_host_batch_v2_view(const guac_host_batch_v2* b);
uint64_t guac_host_batch_v2_bytes(const guac_host_batch_v2* b);   /* bytes guac_reads_pack_v2 will copy for it (with qualities) */
void guac_host_batch_v2_free(guac_host_batch_v2* b);

/* BAM file -> compact batch on host threads: the BGZF members are inflated in parallel, the records filtered (mapped reads
 * only, then Read.InputFilters reads/Read.scala:88-136) and written in parallel straight into the columns guac_reads_pack_v2
 * copies — the 4-bit bases and the CIGAR words are the record's own bytes.  Replaces Read.fromSAMRecord / loadReadsFromBAM
 * (reads/Read.scala:217-291, 368-451) for this path.  Reads come out sorted by (contig index, start) (file order among equals;
 * a coordinate-sorted file is taken as it is).  One sample per batch: `sample` selects the reads whose @RG SM equals it
 * ("default" = reads without a read group); NULL requires the kept reads to share one.  No GPU needed. */
typedef struct guac_bam_options {
  int32_t n_threads;      /* 0 = all hardware threads */
  int32_t non_duplicate;  /* drop SAM flag 0x400 */
  int32_t passed_qc;      /* drop SAM flag 0x200 */
  int32_t has_md_tag;     /* drop reads without an MD tag */
  int32_t is_paired;      /* drop reads without SAM flag 0x1 */
  int32_t with_qualities; /* 0: the qual column stays NULL (germline-threshold does not read it) */
  int32_t pinned;         /* page-lock the columns */
  int32_t reserved;
  const char* sample;
} guac_bam_options;
guac_status guac_bam_load(const char* path, const guac_bam_options* options, guac_host_batch_v2** out);
const char* guac_bam_last_error(void);                                  /* message of this thread's last failed guac_bam_load */
const char* guac_host_batch_v2_contig_name(const guac_host_batch_v2* b, uint32_t contig);
const char* guac_host_batch_v2_sample_name(const guac_host_batch_v2* b);
/* [0] file bytes, [1] inflated bytes, [2] records in the file, [3] reads kept; returns the decode time in milliseconds */
double guac_host_batch_v2_decode_stats(const guac_host_batch_v2* b, uint64_t stats[4]);
void guac_reads_free(guac_reads* reads);
uint64_t guac_reads_count(const guac_reads* reads);
uint64_t guac_reads_device_bytes(const guac_reads* reads);
/* loci where the reads' MD-derived reference bases disagreed and the canonical rule chose one */
uint64_t guac_reads_order_sensitive_loci(const guac_reads* reads);
uint64_t guac_reads_h2d_bytes(const guac_reads* reads);      /* bytes guac_reads_pack copied host -> device */
double guac_reads_expand_kernel_ms(const guac_reads* reads); /* ... of which the expansion into difference streams (k_expand) */
double guac_reads_pack_kernel_ms(const guac_reads* reads);   /* first to last pack kernel on the compute stream (CUDA events);
                                                                 includes waiting for the chunked host -> device copies */

/* ---- the hot path ---------------------------------------------------------------------------------------- */
/* pileupFlatMap(reads, ranges, skip_empty, callVariantsAtLocus(_, threshold, emitRef, emitNoCall)) */
guac_status guac_germline_threshold(guac_ctx* ctx, const guac_reads* reads, const guac_locus_range* ranges,
                                    size_t n_ranges, const guac_threshold_params* params, guac_result** out);
/* pileupFlatMapTwoRDDs(tumor, normal, ranges, skip_empty, findPotentialVariantAtLocus(...)) */
guac_status guac_somatic_standard(guac_ctx* ctx, const guac_reads* tumor, const guac_reads* normal,
                                  const guac_locus_range* ranges, size_t n_ranges,
                                  const guac_somatic_params* params, guac_result** out);
/* pileupFlatMap(reads, ranges, skip_empty, callVariantsAtLocus(_, minAlignmentQuality)) of GermlineStandard
 * (commands/GermlineStandardCaller.scala:66-70, 90-124).  Needs reads packed with base qualities. */
guac_status guac_germline_standard(guac_ctx* ctx, const guac_reads* reads, const guac_locus_range* ranges,
                                   size_t n_ranges, const guac_standard_params* params, guac_result** out);
/* pileupFlatMap(reads, ranges, skip_empty, p => (depth, positiveDepth, referenceDepth, base histogram)) */
guac_status guac_pileup_counts(guac_ctx* ctx, const guac_reads* reads, const guac_locus_range* ranges,
                               size_t n_ranges, int skip_empty, guac_result** out);

/* pileupFlatMap(reads, ranges, true, VariantSupport.Caller.pileupToAlleleCounts) (commands/VariantSupport.scala:93-98,
 * 110-118): every distinct allele of every non-empty pileup with its read count, variable-length alleles
 * included.  One warp per requested locus (the exact per-element walk): meant for the loci of a variant list. */
guac_status guac_allele_counts(guac_ctx* ctx, const guac_reads* reads, const guac_locus_range* ranges,
                               size_t n_ranges, guac_result** out);

/* ---- results (library-owned; pointers valid until guac_result_free) -------------------------------------- */
size_t guac_result_n(const guac_result* r);
const guac_threshold_record* guac_result_threshold_records(const guac_result* r); /* NULL if other kind.  Built on first use
                                                                                    from the compact + general records below */
/* The records of a germline-threshold result as they crossed the bus: `*compact` (return value = their number, in canonical
 * order when GUAC_OPT_SORT_RECORDS is on) + `*general` (`*n_general` records with variable-length alleles, from the exact
 * per-locus kernel; unordered).  guac_result_n = the sum.  Any out pointer may be NULL. */
size_t guac_result_compact_records(const guac_result* r, const guac_compact_record** compact,
                                   const guac_threshold_record** general, size_t* n_general, int32_t* sample);
const guac_somatic_record* guac_result_somatic_records(const guac_result* r);
const guac_locus_counts* guac_result_counts(const guac_result* r);
const guac_called_allele* guac_result_called_alleles(const guac_result* r);
const guac_allele_count* guac_result_allele_counts(const guac_result* r);
const uint8_t* guac_result_bytes(const guac_result* r, size_t* n_bytes);            /* allele byte pool */
const guac_stats* guac_result_stats(const guac_result* r);
void guac_result_free(guac_result* r);

/* ---- post-call genotype filters (host-side predicates on emitted records) -------------------------------
 * filters/SomaticGenotypeFilter.scala: SomaticReadDepthFilter (:69-75, upper bound exclusive, filters/GenotypeFilter.scala:63),
 * SomaticAlternateReadDepthFilter (:107-110), SomaticLogOddsFilter (:176-179), SomaticMinimumLikelihoodFilter (:38-41),
 * SomaticVAFFilter (:142-145, Float VAF), SomaticAverageMappingQualityFilter (:210-214), SomaticAverageBaseQualityFilter
 * (:194-198 — tests MAPPING quality, kept as is), SomaticMedianMismatchFilter (:228-231). */
typedef struct guac_somatic_filter_params {   /* SomaticGenotypeFilterArguments :245-280; defaults in brackets */
  int32_t min_tumor_read_depth;            /* --min-tumor-read-depth [0] */
  int32_t max_tumor_read_depth;            /* --max-tumor-read-depth [INT32_MAX] */
  int32_t min_normal_read_depth;           /* --min-normal-read-depth [0] */
  int32_t min_tumor_alternate_read_depth;  /* --min-tumor-alternate-read-depth [0 = off] */
  int32_t min_lod;                         /* --min-lod [0] */
  int32_t min_likelihood;                  /* --min-likelihood [0] */
  int32_t min_vaf;                         /* --min-vaf [0] */
  int32_t min_average_mapping_quality;     /* --min-average-mapping-quality [0] */
  int32_t min_average_base_quality;        /* --min-average-base-quality [0] */
  int32_t max_median_mismatches;           /* --max-median-mismatches [INT32_MAX] */
  int32_t seq_overload;                    /* 1: the Seq[...] overload (:310-335: depth, VAF, likelihood, alternate depth only) */
  int32_t pad_;
} guac_somatic_filter_params;
/* pileupFlatMapTwoRDDs(...findPotentialVariantAtLocus...) followed by the genotype filters of SomaticStandardCaller.scala:125-151,
 * applied on the device in the caller kernels' epilogue: records that fail a filter never cross PCIe. */
guac_status guac_somatic_standard_filtered(guac_ctx* ctx, const guac_reads* tumor, const guac_reads* normal,
                                           const guac_locus_range* ranges, size_t n_ranges, const guac_somatic_params* params,
                                           const guac_somatic_filter_params* filters, guac_result** out);
/* The same predicates over records already on the host: keep[i] = 1 if records[i] passes every filter; returns the number kept. */
size_t guac_somatic_genotype_filter(const guac_somatic_record* records, size_t n, const guac_somatic_filter_params* params,
                                    uint8_t* keep);

/* ---- LociPartitioning (host-side; DistributedUtil.scala:83-108). Writes at most max_out ranges (task set),
 * returns the number produced through *n_out. ------------------------------------------------------------- */
guac_status guac_partition_loci_uniformly(int64_t tasks, const guac_locus_range* loci, size_t n_loci,
                                          guac_locus_range* out, size_t max_out, size_t* n_out);


/* ---- loci by depth, and the path's only exchange: records + depth histograms to one rank over NCCL ------------------------
 * Loci shard naturally (DistributedUtil.scala:537-545: a locus depends only on the reads overlapping it), so there is no
 * collective on the data path.  What the reference does with collect / coalesce(1, shuffle = true) (Common.scala:290-293)
 * is guac_result_gather: counts all-gathered (48 bytes per rank), every rank's records sent from device memory to the root
 * with one group of ncclSend / ncclRecv, one device -> host copy on the root. */
#define GUAC_DEPTH_BINS 256
/* hist[d] = requested loci covered by exactly d reads (d < 255); hist[255] = deeper.  `hist` may be NULL: the histogram stays
 * on the device for guac_comm_reduce_depth_histogram.  Needs GUAC_OPT_DIFFERENCE_LISTS = 1 (it reads the streams). */
guac_status guac_depth_histogram(guac_ctx* ctx, const guac_reads* reads, const guac_locus_range* ranges, size_t n_ranges,
                                 uint64_t* hist);
typedef struct guac_comm guac_comm;
#define GUAC_COMM_ID_BYTES 128
/* One rank makes the id, the host side hands it to every rank (Spark broadcast / MPI / a file), every rank creates its
 * communicator over its own context (one process or thread per GPU). */
guac_status guac_comm_unique_id(uint8_t* id /* [GUAC_COMM_ID_BYTES] */);
guac_status guac_comm_create(guac_ctx* ctx, const uint8_t* id, int rank, int world, guac_comm** out);
void guac_comm_destroy(guac_comm* comm);
/* Collective: every rank passes the germline-threshold result of its own loci shard (before its context runs another call);
 * `*out` on the root holds all records in rank order — the order of the partitions —, an empty result elsewhere. */
guac_status guac_result_gather(guac_comm* comm, const guac_result* local, int root, guac_result** out);
/* Collective: ncclReduce(sum) of the device-resident histogram of each rank's last guac_depth_histogram; `hist` (root only) */
guac_status guac_comm_reduce_depth_histogram(guac_comm* comm, int root, uint64_t* hist);

/* partitionLociByApproximateDepth (DistributedUtil.scala:162-251; the reference's default at --partition-accuracy 250):
 * `accuracy * tasks` uniform micro partitions, the reads of `read_sets` overlapping each counted on the device, loci then
 * assigned to tasks so that every task sees about the same number of reads.  `loci` in LociSet order, as above. */
guac_status guac_partition_loci_by_approximate_depth(guac_ctx* ctx, int64_t tasks, const guac_locus_range* loci, size_t n_loci,
                                                     int64_t accuracy, const guac_reads* const* read_sets, size_t n_read_sets,
                                                     guac_locus_range* out, size_t max_out, size_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* GUAC_H_ */
