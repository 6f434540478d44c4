/* ORACLE — TEST INFRASTRUCTURE ONLY. C API of the CPU restatement (liboracle.so), loaded with ctypes from tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs. Never linked into or called from libpisces_b200.so. */
#ifndef PO_CAPI_H
#define PO_CAPI_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct po_config {
    int32_t min_base_call_quality, min_map_quality, remove_duplicates, only_proper_pairs;
    float min_frequency, min_frequency_filter, target_lod_frequency;
    int32_t max_vq, min_vq, vq_filter, max_gq, min_gq, low_gq_filter, min_coverage, low_depth_filter, indel_repeat_filter;
    int32_t rmxn_max_repeat_len, rmxn_min_repetitions;
    float rmxn_freq_limit;
    int32_t ploidy, forced_noise_level, noise_model;
    float sb_acceptance;
    int32_t sb_model, filter_single_strand;
    float no_call_filter;
    int32_t call_mnvs, max_size_mnv, max_gap_mnv, collapse;
    float collapse_freq_threshold, collapse_freq_ratio_threshold;
    int32_t exclude_mnvs_from_collapsing, tracked_anchor_size, output_gvcf, source_is_stitched, source_is_collapsed;
    int32_t apply_validation;   /* 1: VariantCallingParameters.Validate() derived values (what Program.Main does); 0: raw options as some reference tests build them */
    float diploid_minor_vf, diploid_major_vf, diploid_sum_vf_multiallelic;   /* DiploidSNVThresholdingParameters (VariantCallingParameters.cs:84) */
    int32_t is_male;            /* bool? IsMale: -1 = null (GenotypeCreator.GetPloidyForThisChr) */
    float amplicon_bias_filter; /* float? AmpliconBiasFilterThreshold (-abfilter): < 0 = null; amplicon tracking follows it (Factory.ShouldTrackAmpliconCounts) */
} po_config;

typedef struct po_read {
    int32_t pos0;              /* 0-based BAM position */
    int32_t flag;              /* SAM flag bits */
    int32_t mapq;
    int32_t n_cigar;
    const uint32_t* cigar;     /* BAM encoding: len<<4 | op (MIDNSHP=X) */
    int32_t l_seq;
    const char* seq;           /* upper-case ASCII bases */
    const uint8_t* qual;
    int32_t has_tags;
    const char* xd;            /* NULL if absent */
    const char* xr;            /* NULL if absent */
    int32_t has_xv, xv, has_xw, xw;
} po_read;

typedef struct po_record {
    int32_t pos, type, genotype, gq, vq;
    uint32_t filter_mask;      /* bit i = FilterType i */
    int32_t n_filters, filters[8]; /* in first-seen order */
    int32_t noise_level, total_coverage;
    double sum_base_quality;
    int32_t cov[3], support[3], well_anchored[3];
    int32_t allele_support, ref_support, num_no_calls;
    float fraction_no_calls, frequency;
    double bias_score, gatk_bias_score;
    int32_t bias_acceptable, var_both_strands, cov_both_strands, forced;
    int32_t collapsed_mut[8], collapsed_total[8];
    int32_t ref_len, alt_len;
    int32_t has_amplicon_bias, amplicon_bias_detected;   /* AmpliconBiasResults != null, .BiasDetected */
    int32_t n_amp_support, amp_support_names[6], amp_support_counts[6];      /* SupportByAmplicon: filled slots in slot order (-1 arrays null) */
    int32_t n_amp_coverage, amp_coverage_names[6], amp_coverage_counts[6];   /* CoverageByAmplicon */
} po_record;

void po_default_config(po_config* c);

/* full per-chromosome caller (SmallVariantCaller.Execute restated) */
void* po_caller_create(const po_config* c, const char* chr_name, const char* seq, int64_t seq_len, const int32_t* iv_start, const int32_t* iv_end, int32_t n_iv);
void po_caller_destroy(void* h);
const char* po_last_error(void);
void po_caller_add_forced(void* h, int32_t pos, const char* ref, const char* alt);
int po_caller_add_read(void* h, const po_read* r);           /* find candidates + counts + call(upTo) */
int po_caller_add_read_counts_only(void* h, const po_read* r); /* RegionStateManager.AddAlleleCounts only */
int po_caller_add_read_candidates_only(void* h, const po_read* r); /* FindCandidates + AddCandidates only */
int po_caller_add_reads_soa(void* h, int32_t n, const int32_t* pos0, const uint16_t* flag, const int64_t* cigar_off, const uint32_t* cigar, const int64_t* seq_off,
                            const uint8_t* bases, const uint8_t* quals, const uint8_t* collapsed /* or NULL */, const int32_t* xd_runs /* [n][3] or NULL */);
int po_caller_add_reads_soa_counts_only(void* h, int32_t n, const int32_t* pos0, const uint16_t* flag, const int64_t* cigar_off, const uint32_t* cigar, const int64_t* seq_off,
                                        const uint8_t* bases, const uint8_t* quals, const uint8_t* collapsed, const int32_t* xd_runs);
/* as po_caller_add_reads_soa, plus the reads' amplicon names (XN tag) as ids, -1 = no tag */
int po_caller_add_reads_soa_amplicons(void* h, int32_t n, const int32_t* pos0, const uint16_t* flag, const int64_t* cigar_off, const uint32_t* cigar, const int64_t* seq_off,
                                      const uint8_t* bases, const uint8_t* quals, const uint8_t* collapsed, const int32_t* xd_runs, const int32_t* amplicon);
int po_caller_add_pileup(void* h, int64_t n_loci, int32_t first_pos, const int64_t* off, const uint8_t* code, const uint8_t* qual, const uint8_t* anch, int32_t call_every);
/* IAlleleSource.AddCandidates with one hand-built candidate (explicit candidates of the locus-major path) */
int po_caller_add_candidate(void* h, int32_t type, int32_t pos, const char* ref, const char* alt, const int32_t support[3], const int32_t well_anchored[3],
                            int32_t open_left, int32_t open_right, const int32_t collapsed_mut[8]);
int po_caller_finish(void* h);
int32_t po_caller_num_records(void* h);
int po_caller_get_record(void* h, int32_t i, po_record* out);
const char* po_caller_record_ref(void* h, int32_t i);
const char* po_caller_record_alt(void* h, int32_t i);
int32_t po_caller_num_write_batches(void* h);
void po_caller_write_batch(void* h, int32_t i, int32_t* begin, int32_t* end);
int32_t po_caller_total_called(void* h);
int32_t po_caller_total_collapsed(void* h);

/* state access (IAlleleSource) */
int32_t po_get_allele_count(void* h, int32_t pos, int32_t allele, int32_t dir, int32_t min_anchor, int32_t max_anchor /* -1 = null */, int32_t from_end, int32_t symmetric);
double po_get_sum_base_quality(void* h, int32_t pos, int32_t allele, int32_t dir, int32_t min_anchor, int32_t max_anchor, int32_t from_end);
int32_t po_get_collapsed_count(void* h, int32_t pos, int32_t type);
void po_set_allele_count(void* h, int32_t pos, int32_t allele, int32_t dir, int32_t anchor, int32_t value); /* mock source for KATs */
void po_add_gapped_ref_count(void* h, int32_t pos, int32_t count);
int po_dump_counts(void* h, int32_t pos0, int32_t n, int32_t* out /* [n][6][3][2K+1] */);
/* candidates currently held by the state manager at a position */
int32_t po_num_candidates_at(void* h, int32_t pos);
int po_get_candidate_at(void* h, int32_t pos, int32_t i, int32_t* type, int32_t support[3], int32_t well_anchored[3], int32_t* open_left, int32_t* open_right,
                        char* ref_buf, char* alt_buf, int32_t buf_len, int32_t collapsed_mut[8]);
/* run CoverageCalculator (+ optionally the rest of ProcessVariant) on one allele against the handle's state */
int po_process_allele(void* h, int32_t type, int32_t pos, const char* ref, const char* alt, const int32_t support[3], const int32_t well_anchored[3],
                      int32_t coverage_only, po_record* out);

/* scalar KAT entry points */
double po_raw_vq(int32_t call_count, int32_t coverage, int32_t nl);
int32_t po_vq(int32_t call_count, int32_t coverage, int32_t nl, int32_t max_q);
double po_pvalue(int32_t call_count, int32_t coverage, int32_t nl);
double po_poisson_cdf(double k, double lambda);
double po_mathnet_gamma_lower_regularized(double a, double x);
double po_mathnet_gamma_ln(double z);
void po_strand_bias(const int32_t cov[3], const int32_t sup[3], int32_t q_noise, double min_vf, double acceptance, int32_t model,
                    double* out /* [0]=bias [1]=gatk [2]=acceptable [3]=varBoth [4]=covBoth, then 4 stats x {FN,FP,VG,cov,freq,sup}: overall,fwd,rev,stitched */);
int32_t po_somatic_gq(int32_t type, int32_t genotype, int32_t vq, int32_t total_coverage, int32_t allele_support, float target_lod, int32_t min_gq, int32_t max_gq);
int32_t po_somatic_genotype(int32_t type, int32_t total_coverage, int32_t allele_support, int32_t ref_support, float min_freq_filter, int32_t min_depth);
/* germline genotypers (po_genotype.hpp) */
double po_mathnet_binomial_cdf(double p, int32_t n, double x);
/* ExactCoverageCalculator (ExactCoverageCalculator.cs:18-109), one spanning read summary: cigar as "5M4I4M", directions as "2F:9S:2R" (DirectionInfo.cs:15-33).
 * Returns the DirectionType the read's coverage goes to, -1 if it does not contribute, -2 for single-point allele types, -3 where the reference throws. */
int32_t po_exact_spanning_read_direction(int32_t allele_type, int32_t reference_position, int32_t allele_length, int32_t clip_adjusted_start,
                                         int32_t clip_adjusted_end, const char* cigar, const char* direction_string);
/* AmpliconBiasCalculator.CalculateAmpliconBias (AmpliconBiasCalculator.cs:45-133); names are ints, -1 = null, n_support < 0 = null array. Returns -1 for a
 * null result, else the number of amplicons; per_amp[i] = {name, frequency, coverage, observedSupport, expectedSupport, chanceItsReal, qScore, biasDetected} */
int32_t po_amplicon_bias(const int32_t* support_names, const int32_t* support_counts, int32_t n_support, const int32_t* coverage_names,
                         const int32_t* coverage_counts, int32_t n_coverage, float acceptance, int32_t max_qscore, int32_t* bias_detected,
                         int32_t* artifact_amplicon, double* per_amp /* [n_coverage][8] */);
double po_mathnet_binomial_probability_ln(double p, int32_t n, int32_t k);
int32_t po_diploid_gq(int32_t genotype, int32_t total_coverage, int32_t allele_support, int32_t min_gq, int32_t max_gq);
int32_t po_haploid_gq(int32_t genotype, int32_t total_coverage, int32_t allele_support, int32_t min_gq, int32_t max_gq);
/* ploidy: 1 = DiploidThresholdingGenotyper.SetGenotypes, 3 = HaploidGenotyper.SetGenotypes over n alleles of one locus (types[] AlleleCategory;
 * ref_alt: n strings "REF>ALT" or NULL). Returns the locus genotype; pruned[i] = 1 for allelesToPrune; *multiallelic = MultiAllelicSite filter set. */
int32_t po_genotype_locus(int32_t ploidy, int32_t n, const int32_t* types, const int32_t* allele_support, const int32_t* total_coverage, const int32_t* ref_support,
                          const char* const* ref_alt, int32_t min_depth, float minor_vf, float major_vf, float sum_vf, int32_t* pruned, int32_t* multiallelic,
                          int32_t* gq_out);
int32_t po_ploidy_for_chr(int32_t sample_ploidy, int32_t is_male, const char* chr_name);
int32_t po_anchor_adjusted_count(const int32_t* bins, int32_t k, int32_t min_anchor, int32_t max_anchor, int32_t from_end, int32_t symmetric);

#ifdef __cplusplus
}
#endif
#endif
