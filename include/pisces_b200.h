/* pisces_b200 — C ABI of the B200-native per-locus variant-calling hot path.
 *
 * Drop-in boundary for Illumina/Pisces (reference @ becd35f). A .NET host P/Invokes these entry points
 * (CallingConvention.Cdecl, the same convention as the reference's only native binding,
 * src/lib/Common.IO/FileCompression.cs:10-35) from implementations of the reference's own seams:
 *
 *   IStateManager.AddAlleleCounts(Read)            src/lib/Pisces.Processing/Interfaces/IStateManager.cs:8-15     -> pb2_push_reads
 *   IStateManager.AddCandidates(candidates)        src/lib/Pisces.Domain/Interfaces/IAlleleSource.cs:10            -> pb2_push_candidates
 *   ICandidateVariantFinder.FindCandidates(Read..) src/lib/Pisces.Domain/Interfaces/ICandidateAlleleFinder.cs:7-10    -> pb2_push_reads (found on the device)
 *   IVariantCollapser.Collapse / MnvReallocator    src/exe/Pisces/Logic/VariantCalling/VariantCollapser.cs:31-113, MnvReallocator.cs:12-98 -> inside pb2_flush
 *   IStateManager.GetCandidatesToProcess/DoneProcessing + IAlleleCaller.Call(batch, source)
 *                                                  src/exe/Pisces/Interfaces/IAlleleCaller.cs:8-13                 -> pb2_flush
 *   IAlleleSource.GetAlleleCount(pos, allele, dir, ...)   IAlleleSource.cs:12                                      -> pb2_get_counts
 *   Factory.CreateStateManager / CreateVariantCaller      src/exe/Pisces/Logic/Factory.cs:128,209                  -> pb2_create (+ pb2_config)
 *   ChrReference / ChrIntervalSet handed to the caller    Factory.cs:253-269                                       -> pb2_set_reference / pb2_set_intervals
 *
 * Conventions: every call returns 0 on success, <0 on error (message: pb2_last_error); nothing throws across the
 * boundary. The caller owns all input buffers (they are consumed before the call returns). Output buffers belong to the
 * handle and stay valid until the next pb2_flush / pb2_destroy on it. One handle per (BAM, chromosome) job, one CUDA stream per
 * handle, no global mutable state: N concurrent handles are safe. There is NO CPU fallback: without a CUDA device every
 * compute entry point fails with PB2_ERR_CUDA.
 *
 * Enum values are the reference's: AlleleType A=0,G=1,C=2,T=3,N=4,Deletion=5 (Types/AlleleType.cs:5-10);
 * DirectionType Forward=0,Reverse=1,Stitched=2; AlleleCategory Snv=0,Insertion=1,Deletion=2,Mnv=3,Reference=4;
 * FilterType / Genotype ordinals as in Types/FilterType.cs, Types/Genotype.cs.
 */
#ifndef PISCES_B200_H
#define PISCES_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PB2_OK 0
#define PB2_ERR_ARG (-1)
#define PB2_ERR_CUDA (-2)
#define PB2_ERR_STATE (-3)
#define PB2_ERR_NOMEM (-4)
#define PB2_ERR_UNSUPPORTED (-5)

typedef struct pb2_handle pb2_handle;

/* The complete parameter set of the path: VariantCallerConfig (src/exe/Pisces/Logic/VariantCalling/AlleleCaller.cs:266-291, filled at
 * Factory.cs:149-179) + the ctor arguments of RegionStateManager (RegionStateManager.cs:36-53) and CandidateVariantFinder
 * (CandidateVariantFinder.cs:20-29). Nullable C# options use -1 for null. Defaults: pb2_default_config. */
typedef struct pb2_config {
    int32_t device;                      /* CUDA device ordinal */
    int32_t min_base_call_quality;       /* BamFilterParameters.MinimumBaseCallQuality (20) */
    float   min_frequency;               /* VariantCallingParameters.MinimumFrequency (0.01) */
    float   min_frequency_filter;        /* MinimumFrequencyFilter (-1 -> raised to min_frequency) */
    float   target_lod_frequency;        /* TargetLODFrequency (-1 -> raised to min_frequency_filter) */
    int32_t max_variant_qscore;          /* 100 */
    int32_t min_variant_qscore;          /* 20 */
    int32_t variant_qscore_filter;       /* MinimumVariantQScoreFilter 30 */
    int32_t max_genotype_qscore;         /* 100 */
    int32_t min_genotype_qscore;         /* 0 */
    int32_t low_genotype_quality_filter; /* null */
    int32_t min_coverage;                /* 10 */
    int32_t low_depth_filter;            /* null -> min_coverage */
    int32_t rmxn_max_repeat_len;         /* 5 */
    int32_t rmxn_min_repetitions;        /* 9 */
    float   rmxn_frequency_limit;        /* 0.35 */
    int32_t forced_noise_level;          /* -1 -> noise level = min_base_call_quality (VariantCallingParameters.cs:109-118) */
    int32_t noise_model;                 /* 0 Flat, 1 Window */
    float   strand_bias_acceptance;      /* 0.5 */
    int32_t strand_bias_model;           /* 0 Poisson, 1 Extended (default), 2 Diploid (StrandBiasCalculator.PopulateDiploidStats :150-173) */
    int32_t filter_single_strand;        /* FilterOutVariantsPresentOnlyOneStrand */
    float   no_call_filter;              /* 0.6 */
    int32_t ploidy;                      /* PloidyModel of the sample: 0 Somatic, 1 DiploidByThresholding, 3 Haploid (2 DiploidByAdaptiveGT -> PB2_ERR_UNSUPPORTED).
                                            The genotyper of a chromosome follows GenotypeCreator.GetPloidyForThisChr (chrM -> somatic, sex chromosomes
                                            of a male sample -> haploid) from the name given to pb2_set_reference */
    int32_t tracked_anchor_size;         /* TrackedAnchorSize 5 -> 11 anchor bins; only 5 is built */
    int32_t output_gvcf;                 /* VcfWritingParameters.OutputGvcfFile (1) = IncludeReferenceCalls */
    int32_t expect_stitched;             /* IAlleleSource.ExpectStitchedReads */
    int32_t expect_collapsed;            /* CollapsedRegionStateManager in use */
    int32_t want_sum_base_quality;       /* fill pb2_call_record.sum_base_quality (always on when noise_model == Window) */
    int32_t collapse;                    /* PiscesApplicationOptions.Collapse (1): candidates are tracked per open-end state (trackOpenEnded) */
    int32_t call_mnvs;                   /* CallMNVs (0): SNV/MNV candidates then come from the candidate finder, not from the counts */
    int32_t indel_repeat_filter;         /* IndelRepeatFilter (null = -1) */
    int32_t max_size_mnv;                /* MaxSizeMNV (3) */
    int32_t max_gap_mnv;                 /* MaxGapBetweenMNV (1) */
    float   collapse_freq_threshold;     /* CollapseFreqThreshold (0) */
    float   collapse_freq_ratio_threshold; /* CollapseFreqRatioThreshold (0.5) */
    int32_t exclude_mnvs_from_collapsing;  /* ExcludeMNVsFromCollapsing (0) */
    int32_t skip_validation;             /* 1: take the option values as given, without the adjustments of VariantCallingParameters.Validate (:137-155) that
                                            Program.Main applies (filters raised to their minimum values); the reference's own functional tests build
                                            options by hand this way (SomaticVariantCallerFunctionalTests.cs:683-758) */
    float   diploid_minor_vf;            /* DiploidSNVThresholdingParameters.MinorVF (0.20): also MinVarFrequency of the germline genotypers */
    float   diploid_major_vf;            /* .MajorVF (0.70) */
    float   diploid_sum_vf_multiallelic; /* .SumVFforMultiAllelicSite (0.80) */
    int32_t is_male;                     /* VariantCallingParameters.IsMale: -1 null, 0 false, 1 true */
    float   amplicon_bias_filter;        /* VariantCallingParameters.AmpliconBiasFilterThreshold (-abfilter): < 0 = null (default). With a threshold the reads'
                                            amplicon names are tracked (Factory.ShouldTrackAmpliconCounts, Factory.cs:51-54) and SNVs get the AmpliconBias filter */
    int32_t reserved[2];                 /* tuning knobs of bench.py; 0 in production */
} pb2_config;

/* Reads for IStateManager.AddAlleleCounts(Read) / ICandidateVariantFinder.FindCandidates(Read, ...), as a struct of arrays. Only
 * reads that passed AlignmentSource.ShouldSkipRead (src/exe/Pisces/Logic/Alignment/AlignmentsSource.cs:84-92) are pushed; they must
 * arrive in position order like the BAM. Bases are upper-case ASCII (BamReader.cs:185-201). cigar_off / seq_off need not start at 0 (a batch may be
 * a window of larger arrays). base_dirs / collapsed may be given for some batches and not for others: reads pushed without them get the values
 * their flag implies. The arrays are copied to the device with cudaMemcpyAsync before the call returns: page-locked (pinned) arrays are copied at the
 * full rate of the link. */
typedef struct pb2_read_batch {
    int32_t n_reads;
    const int32_t*  pos0;        /* [n] BamAlignment.Position (0-based) */
    const uint16_t* flag;        /* [n] SAM flag (0x10 reverse, 0x2 proper pair, 0x40 first mate) */
    const int64_t*  cigar_off;   /* [n+1] into cigar */
    const uint32_t* cigar;       /* BAM encoding len<<4|op, ops MIDNSHP=X */
    const int64_t*  seq_off;     /* [n+1] into bases / quals / base_dirs */
    const uint8_t*  bases;
    const uint8_t*  quals;
    const uint8_t*  base_dirs;   /* optional: Read.SequencedBaseDirectionMap per base (XD tag projected, Read.cs:390-421,664-682); NULL -> from flag 0x10 */
    const uint8_t*  collapsed;   /* optional [n]: bit0 IsCollapsedRead (XV/XW present), bit1 IsDuplex, bits2-3 ReadPairDirection 1=FR 2=RF 0=other (Read.cs:17-71,311-349) */
    const int32_t*  amplicon;    /* optional [n]: Read.GetAmpliconNameIfExists (the XN tag, Read.cs:479-486) as an id >= 0 of the host's name dictionary
                                    (pb2_bam_batch_amplicons hands these out), -1 = no tag. Read only when amplicon_bias_filter >= 0 */
} pb2_read_batch;

/* The same reads for hosts behind a PCIe link: one byte per base instead of two. seq[i] = allele2 << 6 | quality with allele2 = A 0, G 1, C 2, T 3
 * (the reference's AlleleType order) and quality <= 63. A base that is not A/C/G/T, a quality above 63, and an A of quality 0 (whose byte would be 0) are
 * EXCEPTIONS: their seq byte is 0 and exc_index[] (indices into seq, increasing) / exc_base[] (ASCII) / exc_qual[] carry the real values, so the form is
 * lossless. pb2_pack_reads builds it from bases + qualities. The device unpacks into the same read store pb2_push_reads fills. */
typedef struct pb2_packed_read_batch {
    int32_t n_reads;
    const int32_t*  pos0;
    const uint16_t* flag;
    const int64_t*  cigar_off;
    const uint32_t* cigar;
    const int64_t*  seq_off;
    const uint8_t*  seq;
    int64_t n_exceptions;
    const int64_t*  exc_index;
    const uint8_t*  exc_base;
    const uint8_t*  exc_qual;
    const uint8_t*  base_dirs;   /* optional, one byte per base as in pb2_read_batch */
    const uint8_t*  collapsed;   /* optional */
    const int32_t*  amplicon;    /* optional, as in pb2_read_batch */
    /* Compact offsets (optional; 11 instead of 26 bytes of metadata per single-operation read): with cigar_off == NULL and seq_off == NULL, cigar_ops[i]
     * is the number of CIGAR operations of read i (<= 255), cigar[] and seq[] hold exactly this batch's reads back to back (n_cigar_total operations,
     * n_seq_total bases; exception indices address seq[]), and a read's length is the read span of its CIGAR. The offsets are built on the device. */
    const uint8_t*  cigar_ops;
    int64_t n_cigar_total, n_seq_total;
} pb2_packed_read_batch;

/* Locus-major pileup ("pileup columns") in CSR form: locus i covers reference position first_position + i (or positions[i]),
 * its entries are [offsets[i], offsets[i+1]) in the three byte planes. One entry = one call to RegionState.AddAlleleCount
 * (RegionStateManager.cs:155,174,183,206) before minimum-base-quality is applied:
 *   code  : bits 0-2 AlleleType of the read base (Deletion for gap entries), bits 3-4 DirectionType,
 *           bit 5 PB2_ENTRY_OPEN_LEFT, bit 6 PB2_ENTRY_OPEN_RIGHT (CandidateAllele.OpenOnLeft/Right of the SNV this base would
 *           raise, CandidateVariantFinder.cs:112-117,496-553), bit 7 PB2_ENTRY_NO_CANDIDATE (base sits in an '='/'X' op or
 *           outside the chromosome: counted, but never an SNV candidate, CandidateVariantFinder.cs:46-64,102-103)
 *   qual  : base-call quality byte (for Deletion entries: min of the two flanking qualities, CandidateVariantFinder.cs:294-320)
 *   anchor: bits 0-3 anchor bin 0..10 (RegionStateManager.GetAnchorType :83-116), bits 4-7 ReadCollapsedType+1 (0 = none). */
#define PB2_ENTRY_OPEN_LEFT 0x20
#define PB2_ENTRY_OPEN_RIGHT 0x40
#define PB2_ENTRY_NO_CANDIDATE 0x80
typedef struct pb2_pileup_csr {
    int64_t n_loci;
    int32_t first_position;      /* 1-based; used when positions == NULL */
    const int32_t* positions;    /* optional [n_loci], strictly increasing */
    const int64_t* offsets;      /* [n_loci + 1] */
    const uint8_t* code;         /* [offsets[n_loci]] */
    const uint8_t* qual;
    const uint8_t* anchor;
    const uint8_t* ref_bases;    /* optional [n_loci] ASCII; NULL -> taken from pb2_set_reference */
    /* PB2_LAYOUT_PLANES (0): the three planes above. PB2_LAYOUT_PACKED2 (1): two bytes per entry for hosts behind a PCIe link — `anchor` is NULL, the
     * anchor bin rides in the spare bits (code = AlleleType | DirectionType << 3 | (bin & 7) << 5, qual = quality | (bin >> 3) << 7), and the rare
     * candidate flags come as a sparse list: flag_index[n_flags] (entry indices, increasing) with flag_bits[n_flags] (PB2_ENTRY_* bits). Not available
     * with expect_collapsed (the collapsed-read type needs the third byte). */
    int32_t layout;
    int32_t reserved;
    int64_t n_flags;
    const int64_t* flag_index;
    const uint8_t* flag_bits;
} pb2_pileup_csr;
#define PB2_LAYOUT_PLANES 0
#define PB2_LAYOUT_PACKED2 1

/* One candidate allele: the POD image of CandidateAllele (src/lib/Pisces.Domain/Models/Alleles/CandidateAllele.cs:8-125) for
 * IAlleleSource.AddCandidates. Insertions, deletions and MNVs (and, with CallMNVs, SNVs) are explicit candidates: their support is the number
 * of reads in which CandidateVariantFinder raised them, which the pileup counts cannot express. 72 bytes. */
typedef struct pb2_candidate {
    int32_t  position;                 /* ReferencePosition (1-based; the base before an insertion / deletion) */
    uint8_t  type;                     /* AlleleCategory: 0 Snv, 1 Insertion, 2 Deletion, 3 Mnv */
    uint8_t  open_flags;               /* bit0 OpenOnLeft, bit1 OpenOnRight */
    uint16_t ref_len;
    uint16_t alt_len;
    uint16_t reserved;
    uint32_t allele_offset;            /* ReferenceAllele then AlternateAllele (ASCII) at this offset of the arena */
    int32_t  support[3];               /* SupportByDirection */
    int32_t  well_anchored[3];         /* WellAnchoredSupportByDirection */
    int32_t  collapsed_mut[8];         /* ReadCollapsedCountsMut */
} pb2_candidate;

/* One called allele: the POD image of CalledAllele (src/lib/Pisces.Domain/Models/Alleles/CalledAllele.cs:7-140). 96 bytes. */
typedef struct pb2_call_record {
    int32_t  position;                 /* ReferencePosition, 1-based */
    uint8_t  type;                     /* AlleleCategory */
    uint8_t  genotype;                 /* Genotype */
    uint8_t  sb_flags;                 /* bit0 BiasAcceptable, bit1 VarPresentOnBothStrands, bit2 CovPresentOnBothStrands, bit3 IsForcedToReport */
    uint8_t  open_flags;               /* bit0 OpenOnLeft, bit1 OpenOnRight of the source candidate (diagnostic) */
    uint16_t filters;                  /* bit i = FilterType i */
    uint16_t noise_level;              /* NoiseLevelApplied */
    int32_t  variant_qscore;
    int32_t  genotype_qscore;
    int32_t  total_coverage;
    int32_t  coverage_by_direction[3]; /* EstimatedCoverageByDirection */
    int32_t  support_by_direction[3];
    int32_t  allele_support;
    int32_t  reference_support;
    int32_t  num_no_calls;
    float    fraction_no_calls;
    uint32_t allele_bytes;             /* ref_len + alt_len <= 4: the ASCII bases inline (ref then alt, little-endian); else offset into pb2_allele_arena */
    uint16_t ref_len, alt_len;
    double   sum_base_quality;
    double   bias_score;               /* StrandBiasResults.BiasScore */
    double   gatk_bias_score;          /* StrandBiasResults.GATKBiasScore */
} pb2_call_record;

/* The rest of CalledAllele that does not fit the 96-byte record (CalledAllele.cs:7-140): one per record of the last pb2_flush, same order. 80 bytes.
 * collapsed_mut is the candidate's ReadCollapsedCountsMut (explicit candidates; zero for count-based SNVs and reference alleles), collapsed_total is
 * ReadCollapsedCountTotal (CollapsedCoverageCalculator.cs:18-37: the counts at the allele's position / start point; zero unless expect_collapsed). */
typedef struct pb2_call_record_ext {
    int32_t collapsed_mut[8];
    int32_t collapsed_total[8];
    int32_t well_anchored_support[3];  /* WellAnchoredSupportByDirection */
    int32_t phase_set_index;           /* CalledAllele.PhaseSetIndex as the diploid genotyper sets it (DiploidThresholdingGenotyper.cs:57-72): 0 for the
                                        * reference allele, 1, 2, ... for the variant alleles of a locus in their order; 0 where no genotyper set it */
} pb2_call_record_ext;

void pb2_default_config(pb2_config* cfg);
int pb2_create(const pb2_config* cfg, pb2_handle** out);
void pb2_destroy(pb2_handle* h);
const char* pb2_last_error(pb2_handle* h);      /* h may be NULL: last error of a failed pb2_create on this thread */
int pb2_device_count(void);

int pb2_set_reference(pb2_handle* h, const char* chr_name, const uint8_t* seq, int64_t len);
int pb2_set_intervals(pb2_handle* h, const int32_t* start, const int32_t* end, int32_t n);

/* Stage a locus-major pileup. Host pointers: copied through pinned staging with cudaMemcpyAsync. */
int pb2_push_pileup(pb2_handle* h, const pb2_pileup_csr* p);
/* Same, but every pointer in *p is a device pointer (data already resident in HBM). */
int pb2_push_pileup_device(pb2_handle* h, const pb2_pileup_csr* p);

/* IStateManager.AddAlleleCounts + FindCandidates for a batch of reads: the reads are kept by the handle until a pb2_flush clears the
 * positions they cover; the pileup is built on the device (read expansion, bucketing to loci, tile interleave). */
int pb2_push_reads(pb2_handle* h, const pb2_read_batch* batch);

int pb2_push_reads_packed(pb2_handle* h, const pb2_packed_read_batch* batch);
/* Host helper (no device work): packs n bases + qualities into seq[n] and lists the exceptions. Returns the number of exceptions; when it exceeds
 * exc_capacity only the first exc_capacity were stored (call again with larger arrays). */
int64_t pb2_pack_reads(const uint8_t* bases, const uint8_t* quals, int64_t n, uint8_t* seq, int64_t* exc_index, uint8_t* exc_base, uint8_t* exc_qual, int64_t exc_capacity);
/* Host helper of the locus-major path (no device work): code / qual / anchor planes of a pb2_pileup_csr -> PB2_LAYOUT_PACKED2 (pcode, pqual: two bytes per
 * entry) + the sparse list of flagged entries (flag_index increasing, flag_bits = code & 0xe0). offsets [n_loci + 1] + ref_bases [n_loci] are optional:
 * with them, flags on entries that show the reference base of their locus are dropped (they raise no SNV candidate, CandidateVariantFinder.cs:112-141).
 * Returns the number of flagged entries (> flag_capacity: call again with larger arrays), PB2_ERR_UNSUPPORTED when an anchor byte carries a collapsed-read
 * type (bits 4-7). */
int64_t pb2_pack_pileup(const uint8_t* code, const uint8_t* qual, const uint8_t* anchor, int64_t n_entries, const int64_t* offsets, const uint8_t* ref_bases, int64_t n_loci,
                        uint8_t* pcode, uint8_t* pqual, int64_t* flag_index, uint8_t* flag_bits, int64_t flag_capacity);
/* Stages everything pushed through pb2_push_reads so far as one device-resident segment (the pileup is built on the device from the reads the handle
 * keeps there), so that pb2_call_resident / pb2_resident_results run on it: the whole-chromosome form of IStateManager for hosts that push all reads
 * first (bench.py, multi-GPU shards). The reads stay staged; a later pb2_flush re-stages what it needs and supersedes this segment. */
int pb2_stage_reads(pb2_handle* h);

/* IAlleleSource.AddCandidates: explicit candidates for the staged positions (the locus-major path has no reads to find them in; a host
 * that keeps its own ICandidateVariantFinder uses this too). Candidates equal in (position, type, ref, alt[, open ends when Collapse is on])
 * are merged by summing their counts (RegionState.AddCandidate, RegionState.cs:94-174). */
int pb2_push_candidates(pb2_handle* h, const pb2_candidate* cands, int32_t n, const uint8_t* allele_arena, int64_t arena_len);
/* Forced-genotyping alleles of this chromosome (the forcedGtAlleles argument of Factory.CreateSomaticVariantCaller -> SmallVariantCaller's
 * constructor, src/exe/Pisces/Logic/SmallVariantCaller.cs:48-77, and AlleleCaller.AddForcedGtAlleles, AlleleCaller.cs:63-66). Only position,
 * ref_len, alt_len and allele_offset of each pb2_candidate are read. Alleles outside the interval set are dropped; the rest become zero-support
 * candidates once calling passes their position (AddForcedAlleleAsCandidate :118-155), put a reference candidate at their position
 * (RegionState.GetAllCandidates :383-453) and are reported by pb2_flush even when not callable: sb_flags bit3 (IsForcedToReport) and the
 * ForcedReport filter set, genotype / GQ left at a new CalledAllele's values (AlleleCaller.cs:98-118,143-150). They stay set across pb2_reset;
 * n = 0 clears them. pb2_call_resident refuses to run while any are set (their records only come through pb2_flush). */
int pb2_set_forced_alleles(pb2_handle* h, const pb2_candidate* alleles, int32_t n, const uint8_t* allele_arena, int64_t arena_len);
/* The byte arena that pb2_call_record.allele_bytes of the last pb2_flush / pb2_call_resident points into for alleles longer than 4 bases. */
int pb2_allele_arena(pb2_handle* h, const uint8_t** arena, int64_t* len);

/* Count + score everything staged; records stay on the device (bench / multi-GPU gather use this). */
int pb2_call_resident(pb2_handle* h, int64_t* n_records);
/* Device pointers to the results of the last pb2_call_resident: the dense per-locus reference stream (gVCF), its validity bytes,
 * and the compacted variant stream. Any may be NULL. */
int pb2_resident_results(pb2_handle* h, const pb2_call_record** ref_records, const uint8_t** ref_valid, int64_t* n_loci,
                         const pb2_call_record** variant_records, int64_t* n_variants);
/* The job's record sink for the end-of-job gather (one all-gather of per-interval call records): a caller-provided DEVICE buffer of
 * 8 * n_slots + n_slots * slot_records * 96 bytes, [n_slots int64 record counts | n_slots blocks of slot_records pb2_call_record]. Every resident step from
 * now on copies its variant stream (at most slot_records records) and their count into slot (step number % n_slots) on the handle's stream; NULL unsets.
 * pb2_call_resident_async enqueues one step without synchronising the host (after one synchronous pb2_call_resident built the plan); pb2_resident_sync
 * waits for everything enqueued; pb2_sink_sort orders every slot's records by position on the device, ready to be gathered in rank order. */
int pb2_set_resident_sink(pb2_handle* h, void* device_buffer, int64_t slot_records, int32_t n_slots);
int pb2_call_resident_async(pb2_handle* h);
int pb2_resident_sync(pb2_handle* h, int64_t* n_records_last);
int pb2_sink_sort(pb2_handle* h);
/* IAlleleCaller.Call for everything staged up to up_to_position (-1 = all): runs pb2_call_resident if needed, copies the
 * records to the host ordered by (position, ref, alt) as AlleleCaller.cs:96-140,172-176 orders them. */
int pb2_flush(pb2_handle* h, int32_t up_to_position, const pb2_call_record** out, int64_t* n);
/* pb2_flush(-1) for everything pushed through pb2_push_reads with the reads kept on the device: candidates are found again in the stored reads, the
 * batches replayed, nothing is consumed - the whole job from device-resident reads, repeatable. */
int pb2_flush_resident(pb2_handle* h, const pb2_call_record** out, int64_t* n);
/* The pb2_call_record_ext rows of the records the last pb2_flush returned (valid until the next flush). */
int pb2_flush_ext(pb2_handle* h, const pb2_call_record_ext** out, int64_t* n);
/* Parity hook = IAlleleSource.GetAlleleCount over a position range: out[n][6][3][11] int32 (RegionState._alleleCounts). */
int pb2_get_counts(pb2_handle* h, int32_t position0, int32_t n, int32_t* out);
/* Drop staged pileups and results (IStateManager.DoneProcessing). */
int pb2_reset(pb2_handle* h);

/* VCF record lines for called alleles, as the reference's writer prints them: every allele on its own line (AllowMultipleVcfLinesPerLoci, the
 * somatic default) or, with options->crushed, the alleles of a position merged into one line (REF / ALT of MergeCrushedReferenceAndAlt, the smallest
 * QUAL and GQ, merged filters, AD / VF of the 1/2 genotypes; the germline default), optionally with RegionMapper's padding of uncovered interval
 * positions: VcfFileWriter.WriteListOfColocatedAlleles (src/lib/Pisces.IO/VcfFileWriter.cs:206-260) with VcfFormatter
 * (src/lib/Pisces.IO/VcfFormatter.cs:52-71,143-251,283-420). CHROM is the name given to pb2_set_reference; alleles longer than 4 bases are read from
 * the arena of the last pb2_flush. `ext` (pb2_flush_ext) feeds the US tag and may be NULL. The text (one '\n'-terminated line per record; no header)
 * is owned by the handle until the next call. */
typedef struct pb2_vcf_options {
    int32_t debug_mode;          /* PiscesApplicationOptions.DebugMode */
    int32_t output_bias_files;   /* OutputBiasFiles */
    int32_t report_rc_counts;    /* VcfWritingParameters.ReportRcCounts: the US tag */
    int32_t report_ts_counts;    /* ReportTsCounts */
    int32_t crushed;             /* !VcfWritingParameters.AllowMultipleVcfLinesPerLoci: one line per position (GroupsAllelesThenWrite, VcfFileWriter.cs:177-204;
                                  * the default of the diploid / haploid ploidy models, VcfWritingParameters.cs:18-40) */
    int32_t report_no_calls;     /* VcfWritingParameters.ReportNoCalls: the NC tag */
    int32_t pad_intervals;       /* RegionMapper padding (RegionMapper.cs:31-84) with the handle's intervals and reference: 1 = the uncovered interval positions
                                  * before each written position (PadIfNeeded), 2 = also those after the last one (WriteRemaining). One call is one writer
                                  * pass over a chromosome: the padding state starts fresh on every call */
    int32_t reserved_[1];
} pb2_vcf_options;
int pb2_vcf_format(pb2_handle* h, const pb2_call_record* records, const pb2_call_record_ext* ext, int64_t n, const pb2_vcf_options* options, const char** text,
                   int64_t* len);

/* BAM -> pb2_read_batch stager (host code; rows a1-a4 of the path): BGZF inflate + record decode (src/lib/Alignment.IO/BamReader.cs:137-224), the read
 * filter AlignmentSource.ShouldSkipRead (src/exe/Pisces/Logic/Alignment/AlignmentsSource.cs:84-92), Read.SequencedBaseDirectionMap from the XD tag
 * (Read.cs:390-421,664-682) and the collapsed-read summary from XV / XW / XR or the pair flags (Read.cs:66-71,311-349). pb2_bam_next_batch hands out up
 * to max_reads kept reads of ONE reference sequence in file order (*ref_id; n_reads == 0 at the end of the file) as a pb2_read_batch whose arrays the
 * reader owns until the next call: push it with pb2_push_reads. is_stitched / is_collapsed are read off the @PG header lines
 * (BamFileAlignmentExtractor.cs:111-153) and belong in pb2_config.expect_stitched / expect_collapsed. */
typedef struct pb2_bam_reader pb2_bam_reader;
typedef struct pb2_bam_filter {
    int32_t min_map_quality;     /* BamFilterParameters.MinimumMapQuality (1) */
    int32_t remove_duplicates;   /* RemoveDuplicates (1) */
    int32_t only_proper_pairs;   /* OnlyUseProperPairs (0) */
} pb2_bam_filter;
int pb2_bam_open(const char* path, pb2_bam_reader** out);
void pb2_bam_close(pb2_bam_reader* r);
const char* pb2_bam_last_error(pb2_bam_reader* r);
int pb2_bam_header(pb2_bam_reader* r, int32_t* n_refs, const char* const** names, const int32_t** lengths, int32_t* is_stitched, int32_t* is_collapsed);
int pb2_bam_next_batch(pb2_bam_reader* r, const pb2_bam_filter* filter, int32_t max_reads, pb2_read_batch* batch, int32_t* ref_id, int64_t* n_skipped);
/* The same, as the packed batch of pb2_push_reads_packed (one byte per base + exceptions, compact offsets, per-base directions / collapsed summaries only
 * when a read of the batch carries the tags): what a host behind a PCIe link should push. Same lifetime rules as pb2_bam_next_batch. */
int pb2_bam_next_batch_packed(pb2_bam_reader* r, const pb2_bam_filter* filter, int32_t max_reads, pb2_packed_read_batch* batch, int32_t* ref_id, int64_t* n_skipped);
/* Amplicon names of the reads: Read.GetAmpliconNameIfExists (src/lib/Pisces.Domain/Models/Read.cs:479-486), the XN tag through TagUtils.GetStringTag
 * (src/lib/Alignment.Domain/BamCommon.cs:1182-1216). amplicon_id[i] of read i of the batch handed out last indexes the reader's name dictionary
 * (first-seen order over the kept reads of the file so far), -1 without the tag. Host-side input of the amplicon-bias filter (whose
 * device side is not built): RegionState.AddAmpliconCount keeps per-position slots in exactly this first-seen order (RegionState.cs:269-307). */
int pb2_bam_batch_amplicons(pb2_bam_reader* r, const int32_t** amplicon_id, int32_t* n_reads);
int pb2_bam_amplicon_names(pb2_bam_reader* r, int32_t* n, const char* const** names);

/* Interval sharding of one chromosome across handles / GPUs (the reference shards by chromosome and concatenates in genome order:
 * src/lib/Pisces.Processing/Logic/BaseGenomeProcessor.cs:60-72, src/exe/Pisces/Logic/Processing/GenomeProcessor.cs:156-186). pb2_shard_plan cuts the positions
 * [first_position, last_position] into n_shards runs of 1000-bp blocks, balanced by the position-sorted reads' starts (pos0, may be NULL: balanced by
 * positions). Shard i EMITS the positions [own_lo, own_hi] and must SEE the reads [read_first, read_end) (those that can touch [stage_lo, stage_hi]: its own
 * positions plus a halo of two blocks and max_read_span on either side, which holds everything that reaches across a cut: far end points of spanning
 * alleles, MNV leftovers moving into the next block (AlleleCaller.cs:91-92), collapsable candidates pulled from the following block
 * (RegionStateManager.cs:441-457), gapped-MNV reference counts (AlleleCaller.cs:94)). A shard's handle gets pb2_set_owned_range(own_lo, own_hi) and its
 * reads; the records of the shards, concatenated in shard order, are the records of the unsharded chromosome. */
typedef struct pb2_shard {
    int32_t own_lo, own_hi;
    int32_t stage_lo, stage_hi;
    int64_t read_first, read_end;
} pb2_shard;
int pb2_shard_plan(const int32_t* pos0, int64_t n_reads, int32_t first_position, int32_t last_position, int32_t max_read_span, int32_t n_shards, pb2_shard* out);
/* Only positions in [own_lo, own_hi] are emitted by pb2_flush / pb2_call_resident from now on (0, 0: everything). */
int pb2_set_owned_range(pb2_handle* h, int32_t own_lo, int32_t own_hi);

/* IAlleleCaller.TotalNumCollapsed (src/exe/Pisces/Interfaces/IAlleleCaller.cs:11): candidates merged by the collapser since pb2_create. */
int pb2_totals(pb2_handle* h, int64_t* total_collapsed);

/* Timing/diagnostics for bench.py: kernel launches and device milliseconds of the hot kernel since the last call, measured with
 * CUDA events on the handle's stream. */
int pb2_stats(pb2_handle* h, int64_t* hot_kernel_launches, double* hot_kernel_ms, int64_t* total_kernel_launches);
/* The last reads -> device pileup staging (pb2_stage_reads / pb2_flush): bytes of the staged form the hot kernel reads, its rows, and the device time of
 * the staging kernels (CUDA events on the handle's stream). */
int pb2_stage_stats(pb2_handle* h, int64_t* staged_bytes, int64_t* rows, double* stage_ms);
/* The handle's cudaStream_t (as void*) so the host can order its own work against it. */
void* pb2_stream(pb2_handle* h);

#ifdef __cplusplus
}
#endif
#endif
