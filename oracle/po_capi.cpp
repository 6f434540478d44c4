// ORACLE — TEST INFRASTRUCTURE ONLY (see po_math.hpp header). C API over the restatement for ctypes.
#include "po_capi.h"
#include <cstring>
#include "po_caller.hpp"
#include "po_exact_coverage.hpp"

using namespace po;

namespace po {
double MathNetBinomialCdf(double p, int n, double x) { return mathnet::BinomialCdf(p, n, x); }   // Distributions.Binomial(p, n).CumulativeDistribution(x)
}

static thread_local std::string g_err;
const char* po_last_error(void) { return g_err.c_str(); }

void po_default_config(po_config* c) {
    Config d;
    c->min_base_call_quality = d.MinimumBaseCallQuality; c->min_map_quality = d.MinimumMapQuality; c->remove_duplicates = d.RemoveDuplicates;
    c->only_proper_pairs = d.OnlyUseProperPairs; c->min_frequency = d.MinimumFrequency; c->min_frequency_filter = d.MinimumFrequencyFilter;
    c->target_lod_frequency = d.TargetLODFrequency; c->max_vq = d.MaximumVariantQScore; c->min_vq = d.MinimumVariantQScore; c->vq_filter = d.MinimumVariantQScoreFilter;
    c->max_gq = d.MaximumGenotypeQScore; c->min_gq = d.MinimumGenotypeQScore; c->low_gq_filter = d.LowGenotypeQualityFilter; c->min_coverage = d.MinimumCoverage;
    c->low_depth_filter = d.LowDepthFilter; c->indel_repeat_filter = d.IndelRepeatFilter; c->rmxn_max_repeat_len = d.RMxNFilterMaxLengthRepeat;
    c->rmxn_min_repetitions = d.RMxNFilterMinRepetitions; c->rmxn_freq_limit = d.RMxNFilterFrequencyLimit; c->ploidy = d.ploidy; c->forced_noise_level = d.ForcedNoiseLevel;
    c->noise_model = d.noiseModel; c->sb_acceptance = d.StrandBiasAcceptanceCriteria; c->sb_model = d.strandBiasModel; c->filter_single_strand = d.FilterOutVariantsPresentOnlyOneStrand;
    c->no_call_filter = d.NoCallFilterThreshold; c->call_mnvs = d.CallMNVs; c->max_size_mnv = d.MaxSizeMNV; c->max_gap_mnv = d.MaxGapBetweenMNV; c->collapse = d.Collapse;
    c->collapse_freq_threshold = d.CollapseFreqThreshold; c->collapse_freq_ratio_threshold = d.CollapseFreqRatioThreshold;
    c->exclude_mnvs_from_collapsing = d.ExcludeMNVsFromCollapsing; c->tracked_anchor_size = d.TrackedAnchorSize; c->output_gvcf = d.OutputGvcfFile;
    c->source_is_stitched = d.SourceIsStitched; c->source_is_collapsed = d.SourceIsCollapsed; c->apply_validation = 1;
    c->diploid_minor_vf = d.DiploidMinorVF; c->diploid_major_vf = d.DiploidMajorVF; c->diploid_sum_vf_multiallelic = d.DiploidSumVFforMultiAllelicSite; c->is_male = d.IsMale;
    c->amplicon_bias_filter = d.AmpliconBiasFilterThreshold;
}
static Config FromC(const po_config* c) {
    Config d;
    d.MinimumBaseCallQuality = c->min_base_call_quality; d.MinimumMapQuality = c->min_map_quality; d.RemoveDuplicates = c->remove_duplicates;
    d.OnlyUseProperPairs = c->only_proper_pairs; d.MinimumFrequency = c->min_frequency; d.MinimumFrequencyFilter = c->min_frequency_filter;
    d.TargetLODFrequency = c->target_lod_frequency; d.MaximumVariantQScore = c->max_vq; d.MinimumVariantQScore = c->min_vq; d.MinimumVariantQScoreFilter = c->vq_filter;
    d.MaximumGenotypeQScore = c->max_gq; d.MinimumGenotypeQScore = c->min_gq; d.LowGenotypeQualityFilter = c->low_gq_filter; d.MinimumCoverage = c->min_coverage;
    d.LowDepthFilter = c->low_depth_filter; d.IndelRepeatFilter = c->indel_repeat_filter; d.RMxNFilterMaxLengthRepeat = c->rmxn_max_repeat_len;
    d.RMxNFilterMinRepetitions = c->rmxn_min_repetitions; d.RMxNFilterFrequencyLimit = c->rmxn_freq_limit; d.ploidy = c->ploidy; d.ForcedNoiseLevel = c->forced_noise_level;
    d.noiseModel = c->noise_model; d.StrandBiasAcceptanceCriteria = c->sb_acceptance; d.strandBiasModel = c->sb_model; d.FilterOutVariantsPresentOnlyOneStrand = c->filter_single_strand;
    d.NoCallFilterThreshold = c->no_call_filter; d.CallMNVs = c->call_mnvs; d.MaxSizeMNV = c->max_size_mnv; d.MaxGapBetweenMNV = c->max_gap_mnv; d.Collapse = c->collapse;
    d.CollapseFreqThreshold = c->collapse_freq_threshold; d.CollapseFreqRatioThreshold = c->collapse_freq_ratio_threshold;
    d.ExcludeMNVsFromCollapsing = c->exclude_mnvs_from_collapsing; d.TrackedAnchorSize = c->tracked_anchor_size; d.OutputGvcfFile = c->output_gvcf;
    d.SourceIsStitched = c->source_is_stitched; d.SourceIsCollapsed = c->source_is_collapsed; d.ApplyValidation = c->apply_validation != 0;
    d.DiploidMinorVF = c->diploid_minor_vf; d.DiploidMajorVF = c->diploid_major_vf; d.DiploidSumVFforMultiAllelicSite = c->diploid_sum_vf_multiallelic; d.IsMale = c->is_male;
    d.AmpliconBiasFilterThreshold = c->amplicon_bias_filter;
    return d;
}
static Read ToRead(const po_read* r) {
    static const char ops[] = "MIDNSHP=X";
    Read o;
    o.BamPosition = r->pos0;
    o.MapQuality = (uint32_t)r->mapq;
    for (int i = 0; i < r->n_cigar; i++) {
        uint32_t c = r->cigar[i];
        uint32_t op = c & 0xf;
        if (op > 8) throw std::runtime_error("bad cigar op");
        o.CigarData.push_back(CigarOp{ops[op], c >> 4});
    }
    o.Sequence.assign(r->seq, (size_t)r->l_seq);
    o.Qualities.assign(r->qual, r->qual + r->l_seq);
    int f = r->flag;  // BamCommon.cs flag accessors
    o.IsMapped = !(f & 0x4); o.IsPrimaryAlignment = !(f & 0x100); o.IsPcrDuplicate = (f & 0x400) != 0; o.IsProperPair = (f & 0x2) != 0;
    o.IsReverseStrand = (f & 0x10) != 0; o.IsFirstMate = (f & 0x40) != 0;
    o.hasTagData = r->has_tags != 0;
    if (r->xd) o.XD = std::string(r->xd);
    if (r->xr) o.XR = std::string(r->xr);
    if (r->has_xv) o.XV = r->xv;
    if (r->has_xw) o.XW = r->xw;
    return o;
}
#define GUARD(expr) try { expr; return 0; } catch (const std::exception& e) { g_err = e.what(); return -1; }

void* po_caller_create(const po_config* c, const char* chr_name, const char* seq, int64_t seq_len, const int32_t* iv_start, const int32_t* iv_end, int32_t n_iv) {
    try {
        std::vector<Region> ivs;
        for (int i = 0; i < n_iv; i++) ivs.push_back(Region{iv_start[i], iv_end[i]});
        return new SmallVariantCaller(FromC(c), chr_name, std::string(seq, (size_t)seq_len), n_iv >= 0 && iv_start ? &ivs : nullptr);
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void po_caller_destroy(void* h) { delete (SmallVariantCaller*)h; }
void po_caller_add_forced(void* h, int32_t pos, const char* ref, const char* alt) { ((SmallVariantCaller*)h)->AddForcedAllele(pos, ref, alt); }
int po_caller_add_read(void* h, const po_read* r) { GUARD(((SmallVariantCaller*)h)->ProcessRead(ToRead(r))) }
int po_caller_add_read_counts_only(void* h, const po_read* r) { GUARD(((SmallVariantCaller*)h)->state->AddAlleleCounts(ToRead(r))) }
int po_caller_add_read_candidates_only(void* h, const po_read* r) {
    auto* s = (SmallVariantCaller*)h;
    GUARD(s->state->AddCandidates(s->finder->FindCandidates(ToRead(r), s->chrSeq, s->chrName)))
}
// A struct of arrays of reads (the layout of pb2_read_batch) through SmallVariantCaller.Execute's per-read loop (SmallVariantCaller.cs:88-104): the bulk form
// of po_caller_add_read for large synthetic read sets (bench.py's CPU baseline / reference arm, full-size parity tests). collapsed: optional per-read
// summary byte (bit0 XV/XW present, bit1 duplex, bits 2-3 pair direction 1 FR / 2 RF / 0 other) turned back into the tags the reference reads;
// xd_runs: optional [n][3] lengths of the F / S / R runs of the XD direction string over the CIGAR-expanded alignment (all zero: no XD tag).
static int add_reads_soa_impl(void* h, int32_t n, const int32_t* pos0, const uint16_t* flag, const int64_t* cigar_off, const uint32_t* cigar, const int64_t* seq_off,
                              const uint8_t* bases, const uint8_t* quals, const uint8_t* collapsed, const int32_t* xd_runs, bool counts_only, const int32_t* amplicon);
int po_caller_add_reads_soa(void* h, int32_t n, const int32_t* pos0, const uint16_t* flag, const int64_t* cigar_off, const uint32_t* cigar, const int64_t* seq_off,
                            const uint8_t* bases, const uint8_t* quals, const uint8_t* collapsed, const int32_t* xd_runs) {
    return add_reads_soa_impl(h, n, pos0, flag, cigar_off, cigar, seq_off, bases, quals, collapsed, xd_runs, false, nullptr);
}
int po_caller_add_reads_soa_amplicons(void* h, int32_t n, const int32_t* pos0, const uint16_t* flag, const int64_t* cigar_off, const uint32_t* cigar, const int64_t* seq_off,
                                      const uint8_t* bases, const uint8_t* quals, const uint8_t* collapsed, const int32_t* xd_runs, const int32_t* amplicon) {
    return add_reads_soa_impl(h, n, pos0, flag, cigar_off, cigar, seq_off, bases, quals, collapsed, xd_runs, false, amplicon);
}
// RegionStateManager.AddAlleleCounts only (nothing is called, no block is cleared): for po_dump_counts
int po_caller_add_reads_soa_counts_only(void* h, int32_t n, const int32_t* pos0, const uint16_t* flag, const int64_t* cigar_off, const uint32_t* cigar, const int64_t* seq_off,
                                        const uint8_t* bases, const uint8_t* quals, const uint8_t* collapsed, const int32_t* xd_runs) {
    return add_reads_soa_impl(h, n, pos0, flag, cigar_off, cigar, seq_off, bases, quals, collapsed, xd_runs, true, nullptr);
}
static int add_reads_soa_impl(void* h, int32_t n, const int32_t* pos0, const uint16_t* flag, const int64_t* cigar_off, const uint32_t* cigar, const int64_t* seq_off,
                              const uint8_t* bases, const uint8_t* quals, const uint8_t* collapsed, const int32_t* xd_runs, bool counts_only, const int32_t* amplicon) {
    auto* s = (SmallVariantCaller*)h;
    try {
        for (int32_t i = 0; i < n; i++) {
            po_read r;
            memset(&r, 0, sizeof(r));
            r.pos0 = pos0[i]; r.flag = flag[i]; r.mapq = 60;
            r.n_cigar = (int32_t)(cigar_off[i + 1] - cigar_off[i]); r.cigar = cigar + cigar_off[i];
            r.l_seq = (int32_t)(seq_off[i + 1] - seq_off[i]); r.seq = (const char*)bases + seq_off[i]; r.qual = quals + seq_off[i];
            std::string xd;
            if (xd_runs && (xd_runs[3 * i] | xd_runs[3 * i + 1] | xd_runs[3 * i + 2])) {
                static const char d[3] = {'F', 'S', 'R'};
                for (int k = 0; k < 3; k++) if (xd_runs[3 * i + k] > 0) xd += std::to_string(xd_runs[3 * i + k]) + d[k];
                r.has_tags = 1; r.xd = xd.c_str();
            }
            if (collapsed && (collapsed[i] & 1)) {
                r.has_tags = 1; r.has_xv = 1; r.xv = 1; r.has_xw = 1; r.xw = (collapsed[i] & 2) ? 1 : 0;
                const int pd = (collapsed[i] >> 2) & 3;
                r.xr = pd == 1 ? "FR" : (pd == 2 ? "RF" : "FF");
            }
            Read rd = ToRead(&r);
            if (amplicon) rd.AmpliconName = amplicon[i];
            if (counts_only) s->state->AddAlleleCounts(rd); else s->ProcessRead(rd);
        }
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// Locus-major pileup entries (the staging format of include/pisces_b200.h, pb2_pileup_csr) pushed through the SAME per-base operations
// the reference performs in RegionStateManager.AddAlleleCounts (:161-192) and CandidateVariantFinder (SNV candidates, CallMNVs=false),
// followed by Call(upTo) as positions clear. Used for parity of the locus-major path and as the CPU baseline of bench.py.
int po_caller_add_pileup(void* h, int64_t n_loci, int32_t first_pos, const int64_t* off, const uint8_t* code, const uint8_t* qual, const uint8_t* anch, int32_t call_every) {
    auto* s = (SmallVariantCaller*)h;
    try {
        auto& st = *s->state;
        const int minBQ = s->cfg.MinimumBaseCallQuality;
        for (int64_t i = 0; i < n_loci; i++) {
            const int p = first_pos + (int)i;
            const char refc = (p >= 1 && p <= (int)s->chrSeq.size()) ? s->chrSeq[(size_t)p - 1] : 'N';
            const AlleleType refA = GetAlleleType(refc);
            std::vector<CandPtr> cands;
            for (int64_t e = off[i]; e < off[i + 1]; e++) {
                const int allele = code[e] & 7, dir = (code[e] >> 3) & 3, q = qual[e], anchor = anch[e] & 15, ct = anch[e] >> 4;
                if (allele == AT_Del) {
                    if (q >= minBQ) {
                        st.GetBlock(p)->AddAlleleCount(p, AT_Del, (DirectionType)dir, anchor);
                        if (ct && st.expectCollapsed) st.GetBlock(p)->AddCollapsedReadCount(p, (ReadCollapsedType)(ct - 1));
                    }
                    continue;
                }
                AlleleType a2 = q < minBQ ? AT_N : (AlleleType)allele;
                st.GetBlock(p)->AddAlleleCount(p, a2, (DirectionType)dir, anchor);
                if (a2 != AT_N && ct && st.expectCollapsed) st.GetBlock(p)->AddCollapsedReadCount(p, (ReadCollapsedType)(ct - 1));
                float expo = (float)(-1 * q) / 10.0f;
                st.GetBlock(p)->AddBaseQualites(p, a2, (DirectionType)dir, std::pow(10.0, (double)expo), anchor);
                if (a2 != AT_N && !(code[e] & 0x80) && refA != AT_N && a2 != refA) {
                    static const char b[4] = {'A', 'G', 'C', 'T'};
                    auto c = std::make_shared<CandidateAllele>(s->chrName, p, std::string(1, refc), std::string(1, b[a2]), Snv);
                    c->SupportByDirection[dir] = 1;
                    c->OpenOnLeft = (code[e] & 0x20) != 0;
                    c->OpenOnRight = (code[e] & 0x40) != 0;
                    cands.push_back(c);
                }
            }
            st.AddCandidates(cands);
            if (call_every > 0 && (i % call_every) == (call_every - 1)) s->Call(p);
        }
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int po_caller_add_candidate(void* h, int32_t type, int32_t pos, const char* ref, const char* alt, const int32_t support[3], const int32_t wa[3], int32_t open_left,
                            int32_t open_right, const int32_t collapsed_mut[8]) {
    try {
        auto* s = (SmallVariantCaller*)h;
        auto c = std::make_shared<CandidateAllele>(s->chrName, pos, ref, alt, (AlleleCategory)type);
        for (int k = 0; k < 3; k++) { c->SupportByDirection[k] = support[k]; c->WellAnchoredSupportByDirection[k] = wa[k]; }
        if (collapsed_mut) for (int k = 0; k < 8; k++) c->ReadCollapsedCountsMut[k] = collapsed_mut[k];
        c->OpenOnLeft = open_left != 0; c->OpenOnRight = open_right != 0;
        s->state->AddCandidates({c});
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int po_caller_finish(void* h) { GUARD(((SmallVariantCaller*)h)->Finish()) }
int32_t po_caller_num_records(void* h) { return (int32_t)((SmallVariantCaller*)h)->output.size(); }

static void Fill(const CalledAllele& a, po_record* o) {
    memset(o, 0, sizeof(*o));
    o->pos = a.ReferencePosition; o->type = a.Type; o->genotype = a.genotype; o->gq = a.GenotypeQscore; o->vq = a.VariantQscore;
    o->n_filters = (int)std::min<size_t>(8, a.Filters.size());
    for (size_t i = 0; i < a.Filters.size(); i++) { o->filter_mask |= 1u << a.Filters[i]; if (i < 8) o->filters[i] = a.Filters[i]; }
    o->noise_level = a.NoiseLevelApplied; o->total_coverage = a.TotalCoverage; o->sum_base_quality = a.SumOfBaseQuality;
    for (int i = 0; i < 3; i++) { o->cov[i] = a.EstimatedCoverageByDirection[i]; o->support[i] = a.SupportByDirection[i]; o->well_anchored[i] = a.WellAnchoredSupportByDirection[i]; }
    o->allele_support = a.AlleleSupport; o->ref_support = a.ReferenceSupport; o->num_no_calls = a.NumNoCalls; o->fraction_no_calls = a.FractionNoCalls;
    o->frequency = a.Frequency(); o->bias_score = a.StrandBiasResults.BiasScore; o->gatk_bias_score = a.StrandBiasResults.GATKBiasScore;
    o->bias_acceptable = a.StrandBiasResults.BiasAcceptable; o->var_both_strands = a.StrandBiasResults.VarPresentOnBothStrands;
    o->cov_both_strands = a.StrandBiasResults.CovPresentOnBothStrands; o->forced = a.IsForcedToReport;
    for (int i = 0; i < 8; i++) { o->collapsed_mut[i] = a.ReadCollapsedCountsMut[i]; o->collapsed_total[i] = a.ReadCollapsedCountTotal[i]; }
    o->ref_len = (int)a.ReferenceAllele.size(); o->alt_len = (int)a.AlternateAllele.size();
    o->has_amplicon_bias = a.HasAmpliconBiasResults; o->amplicon_bias_detected = a.AmpliconBiasDetected;
    auto amp = [](const AmpliconCounts& c, int32_t* n, int32_t* names, int32_t* counts) {
        *n = c.isNull ? -1 : 0;
        for (int i = 0; i < MaxNumOverlappingAmplicons && !c.isNull; i++)
            if (c.names[(size_t)i] >= 0) { names[*n] = c.names[(size_t)i]; counts[*n] = c.counts[(size_t)i]; (*n)++; }
    };
    amp(a.SupportByAmplicon, &o->n_amp_support, o->amp_support_names, o->amp_support_counts);
    amp(a.CoverageByAmplicon, &o->n_amp_coverage, o->amp_coverage_names, o->amp_coverage_counts);
}
int po_caller_get_record(void* h, int32_t i, po_record* out) { GUARD(Fill(*((SmallVariantCaller*)h)->output.at((size_t)i), out)) }
const char* po_caller_record_ref(void* h, int32_t i) { return ((SmallVariantCaller*)h)->output.at((size_t)i)->ReferenceAllele.c_str(); }
const char* po_caller_record_alt(void* h, int32_t i) { return ((SmallVariantCaller*)h)->output.at((size_t)i)->AlternateAllele.c_str(); }
int32_t po_caller_num_write_batches(void* h) { return (int32_t)((SmallVariantCaller*)h)->writeBatches.size(); }
void po_caller_write_batch(void* h, int32_t i, int32_t* b, int32_t* e) { auto& w = ((SmallVariantCaller*)h)->writeBatches.at((size_t)i); *b = w.first; *e = w.second; }
int32_t po_caller_total_called(void* h) { return ((SmallVariantCaller*)h)->caller->TotalNumCalled; }
int32_t po_caller_total_collapsed(void* h) { return ((SmallVariantCaller*)h)->caller->TotalNumCollapsed(); }

int32_t po_get_allele_count(void* h, int32_t pos, int32_t allele, int32_t dir, int32_t min_anchor, int32_t max_anchor, int32_t from_end, int32_t symmetric) {
    return ((SmallVariantCaller*)h)->state->GetAlleleCount(pos, (AlleleType)allele, (DirectionType)dir, min_anchor,
                                                           max_anchor < 0 ? std::nullopt : std::optional<int>(max_anchor), from_end != 0, symmetric != 0);
}
double po_get_sum_base_quality(void* h, int32_t pos, int32_t allele, int32_t dir, int32_t min_anchor, int32_t max_anchor, int32_t from_end) {
    return ((SmallVariantCaller*)h)->state->GetSumOfAlleleBaseQualities(pos, (AlleleType)allele, (DirectionType)dir, min_anchor,
                                                                       max_anchor < 0 ? std::nullopt : std::optional<int>(max_anchor), from_end != 0);
}
int32_t po_get_collapsed_count(void* h, int32_t pos, int32_t type) { return ((SmallVariantCaller*)h)->state->GetCollapsedReadCount(pos, (ReadCollapsedType)type); }
void po_set_allele_count(void* h, int32_t pos, int32_t allele, int32_t dir, int32_t anchor, int32_t value) {
    auto* b = ((SmallVariantCaller*)h)->state->GetBlock(pos);
    b->counts[b->Idx(pos, allele, dir, anchor)] = value;
}
void po_add_gapped_ref_count(void* h, int32_t pos, int32_t count) { ((SmallVariantCaller*)h)->state->AddGappedMnvRefCount({{pos, count}}); }
int po_dump_counts(void* h, int32_t pos0, int32_t n, int32_t* out) {
    auto* st = ((SmallVariantCaller*)h)->state.get();
    int NA = st->NumAnchorIndexes();
    try {
        for (int i = 0; i < n; i++) {
            auto* b = st->GetBlock(pos0 + i, false);
            for (int c = 0; c < 18 * NA; c++) out[(size_t)i * 18 * NA + c] = b ? b->counts[b->Idx(pos0 + i, 0, 0, 0) + c] : 0;
        }
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int32_t po_num_candidates_at(void* h, int32_t pos) {
    auto* b = ((SmallVariantCaller*)h)->state->GetBlock(pos, false);
    return b ? (int32_t)b->cands[pos - b->StartPosition].size() : 0;
}
int po_get_candidate_at(void* h, int32_t pos, int32_t i, int32_t* type, int32_t support[3], int32_t wa[3], int32_t* ol, int32_t* orr, char* ref_buf, char* alt_buf,
                        int32_t buf_len, int32_t collapsed_mut[8]) {
    try {
        auto* b = ((SmallVariantCaller*)h)->state->GetBlock(pos, false);
        if (!b) throw std::runtime_error("no block");
        auto& c = *b->cands[pos - b->StartPosition].at((size_t)i);
        *type = c.Type; *ol = c.OpenOnLeft; *orr = c.OpenOnRight;
        for (int k = 0; k < 3; k++) { support[k] = c.SupportByDirection[k]; wa[k] = c.WellAnchoredSupportByDirection[k]; }
        for (int k = 0; k < 8; k++) collapsed_mut[k] = c.ReadCollapsedCountsMut[k];
        snprintf(ref_buf, (size_t)buf_len, "%s", c.ReferenceAllele.c_str());
        snprintf(alt_buf, (size_t)buf_len, "%s", c.AlternateAllele.c_str());
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int po_process_allele(void* h, int32_t type, int32_t pos, const char* ref, const char* alt, const int32_t support[3], const int32_t wa[3], int32_t coverage_only, po_record* out) {
    try {
        auto* s = (SmallVariantCaller*)h;
        CandidateAllele c(s->chrName, pos, ref, alt, (AlleleCategory)type);
        for (int k = 0; k < 3; k++) { c.SupportByDirection[k] = support[k]; c.WellAnchoredSupportByDirection[k] = wa[k]; }
        CalledAllele a = MapToCalled(c);
        if (coverage_only) s->caller->coverage.Compute(a, *s->state);
        else {
            s->caller->ProcessVariant(*s->state, a);
            std::vector<CalledPtr> at{std::make_shared<CalledAllele>(a)};
            s->caller->ComputeGenotypeAndFilterAllele(at);
            a = *at[0];
        }
        Fill(a, out);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

double po_raw_vq(int32_t k, int32_t n, int32_t nl) { return AssignRawPoissonQScore(k, n, nl); }
int32_t po_vq(int32_t k, int32_t n, int32_t nl, int32_t max_q) { return AssignPoissonQScore(k, n, nl, max_q); }
double po_pvalue(int32_t k, int32_t n, int32_t nl) { return AssignPValue(k, n, nl); }
double po_poisson_cdf(double k, double lambda) { return pisces_poisson::Cdf(k, lambda); }
int32_t po_exact_spanning_read_direction(int32_t type, int32_t pos, int32_t len, int32_t start, int32_t end, const char* cigar, const char* dirs) {
    std::vector<CigarOp> ops;
    std::vector<std::pair<int, DirectionType>> d;
    int num = 0;
    for (const char* c = cigar; *c; c++) {
        if (*c >= '0' && *c <= '9') num = num * 10 + (*c - '0');
        else { ops.push_back(CigarOp{*c, (uint32_t)num}); num = 0; }
    }
    num = 0;
    for (const char* c = dirs; *c; c++) {
        if (*c >= '0' && *c <= '9') num = num * 10 + (*c - '0');
        else if (*c != ':') { d.push_back({num, *c == 'F' ? Forward : *c == 'R' ? Reverse : Stitched}); num = 0; }
    }
    return ExactSpanningReadDirection((AlleleCategory)type, pos, len, start, end, ops, d);
}
int32_t po_amplicon_bias(const int32_t* sn, const int32_t* sc, int32_t ns, const int32_t* cn, const int32_t* cc, int32_t nc, float acceptance, int32_t max_q,
                         int32_t* bias_detected, int32_t* artifact, double* per_amp) {
    const AmpliconBiasResults r = CalculateAmpliconBias(sn, sc, ns, cn, cc, nc, acceptance, max_q);
    if (r.isNull) return -1;
    if (bias_detected) *bias_detected = r.biasDetected ? 1 : 0;
    if (artifact) *artifact = r.ampliconWithCandidateArtifact;
    for (size_t i = 0; per_amp && i < r.results.size(); i++) {
        const AmpliconBiasEntry& e = r.results[i];
        double* o = per_amp + 8 * i;
        o[0] = e.name; o[1] = e.frequency; o[2] = e.coverage; o[3] = e.observedSupport; o[4] = e.expectedSupport; o[5] = e.chanceItsReal;
        o[6] = e.confidenceQScore; o[7] = e.biasDetected ? 1 : 0;
    }
    return (int32_t)r.results.size();
}
double po_mathnet_gamma_lower_regularized(double a, double x) { return mathnet::GammaLowerRegularized(a, x); }
double po_mathnet_gamma_ln(double z) { return mathnet::GammaLn(z); }
void po_strand_bias(const int32_t cov[3], const int32_t sup[3], int32_t q, double min_vf, double acc, int32_t model, double* out) {
    BiasResults r = CalculateStrandBiasResults(cov, sup, q, min_vf, acc, model);
    out[0] = r.BiasScore; out[1] = r.GATKBiasScore; out[2] = r.BiasAcceptable; out[3] = r.VarPresentOnBothStrands; out[4] = r.CovPresentOnBothStrands;
    const StrandBiasStats* st[4] = {&r.OverallStats, &r.ForwardStats, &r.ReverseStats, &r.StitchedStats};
    for (int i = 0; i < 4; i++) {
        double* o = out + 5 + i * 6;
        o[0] = st[i]->ChanceFalseNeg; o[1] = st[i]->ChanceFalsePos; o[2] = st[i]->ChanceVarFreqGreaterThanZero; o[3] = st[i]->Coverage; o[4] = st[i]->Frequency; o[5] = st[i]->Support;
    }
}
int32_t po_somatic_gq(int32_t type, int32_t genotype, int32_t vq, int32_t total_coverage, int32_t allele_support, float target_lod, int32_t min_gq, int32_t max_gq) {
    CalledAllele a((AlleleCategory)type);
    a.genotype = (Genotype)genotype; a.VariantQscore = vq; a.TotalCoverage = total_coverage; a.AlleleSupport = allele_support;
    return SomaticGenotypeQuality(a, target_lod, min_gq, max_gq);
}
int32_t po_somatic_genotype(int32_t type, int32_t total_coverage, int32_t allele_support, int32_t ref_support, float min_freq_filter, int32_t min_depth) {
    CalledAllele a((AlleleCategory)type);
    a.TotalCoverage = total_coverage; a.AlleleSupport = allele_support; a.ReferenceSupport = ref_support;
    return CalculateSomaticGenotype(a, min_freq_filter, min_depth);
}
int32_t po_anchor_adjusted_count(const int32_t* bins, int32_t k, int32_t min_anchor, int32_t max_anchor, int32_t from_end, int32_t symmetric) {
    return AnchorAdjusted<int>(min_anchor, from_end != 0, k, 2 * k + 1, bins, max_anchor < 0 ? std::nullopt : std::optional<int>(max_anchor), symmetric != 0);
}

// ------------------------------------------------------------------ germline genotypers (po_genotype.hpp)
double po_mathnet_binomial_cdf(double p, int32_t n, double x) { return mathnet::BinomialCdf(p, n, x); }
double po_mathnet_binomial_probability_ln(double p, int32_t n, int32_t k) { return mathnet::BinomialProbabilityLn(p, n, k); }
int32_t po_diploid_gq(int32_t genotype, int32_t total_coverage, int32_t allele_support, int32_t min_gq, int32_t max_gq) {
    CalledAllele a; a.genotype = (Genotype)genotype; a.TotalCoverage = total_coverage; a.AlleleSupport = allele_support;
    return DiploidGenotypeQuality(a, min_gq, max_gq);
}
int32_t po_haploid_gq(int32_t genotype, int32_t total_coverage, int32_t allele_support, int32_t min_gq, int32_t max_gq) {
    CalledAllele a; a.genotype = (Genotype)genotype; a.TotalCoverage = total_coverage; a.AlleleSupport = allele_support;
    return HaploidGenotypeQuality(a, min_gq, max_gq);
}
int32_t po_genotype_locus(int32_t ploidy, int32_t n, const int32_t* types, const int32_t* allele_support, const int32_t* total_coverage, const int32_t* ref_support,
                          const char* const* ref_alt, int32_t min_depth, float minor_vf, float major_vf, float sum_vf, int32_t* pruned, int32_t* multiallelic,
                          int32_t* gq_out) {
    std::vector<CalledPtr> alleles;
    for (int i = 0; i < n; i++) {
        auto a = std::make_shared<CalledAllele>((AlleleCategory)types[i]);
        a->Chromosome = "chr1"; a->ReferencePosition = 1;
        a->AlleleSupport = allele_support[i]; a->TotalCoverage = total_coverage[i]; a->ReferenceSupport = ref_support[i];
        if (ref_alt && ref_alt[i]) { std::string s(ref_alt[i]); auto k = s.find('>'); a->ReferenceAllele = s.substr(0, k); a->AlternateAllele = s.substr(k + 1); }
        alleles.push_back(a);
    }
    DiploidThresholdingParameters snv; snv.MinorVF = minor_vf; snv.MajorVF = major_vf; snv.SumVFforMultiAllelicSite = sum_vf;
    auto prune = ploidy == PM_Haploid ? HaploidSetGenotypes(alleles, min_depth, minor_vf, major_vf, 0, 100) : DiploidSetGenotypes(alleles, min_depth, snv, snv, 0, 100);
    *multiallelic = 0;
    for (int i = 0; i < n; i++) {
        pruned[i] = std::find(prune.begin(), prune.end(), alleles[(size_t)i]) != prune.end();
        for (auto f : alleles[(size_t)i]->Filters) if (f == F_MultiAllelicSite) *multiallelic = 1;
        if (gq_out) gq_out[i] = alleles[(size_t)i]->GenotypeQscore;
    }
    return n ? (int32_t)alleles[0]->genotype : -1;
}
int32_t po_ploidy_for_chr(int32_t sample_ploidy, int32_t is_male, const char* chr_name) { return GetPloidyForThisChr(sample_ploidy, is_male, chr_name); }
