// Explicit-candidate path (insertions, deletions, MNVs, and SNVs when CallMNVs is on): per-locus 198-bin count gather + spanning
// coverage + scoring on the device; the candidate table, VariantCollapser and MnvReallocator run on the host around it (pb2_explicit.cu).
#pragma once
#include "pb2_kernels.cuh"

namespace pb2 {

// One candidate as the scoring kernel sees it. 64 bytes.
struct DevCand {
    int32_t position;          // ReferencePosition
    uint8_t type;              // AlleleCategory
    uint8_t flags;             // kCandAltHasN | kCandReportable
    uint16_t pad0;
    int32_t ref_len, alt_len;
    uint32_t allele_off;       // ref bytes then alt bytes in the device arena
    int32_t support[3];        // SupportByDirection
    int32_t well_anchored;     // Σ WellAnchoredSupportByDirection
    int32_t req_start, req_end;  // rows of the gathered count table for the start / end point position (-1: not staged -> all zero)
    int32_t gapped_ref;        // RegionState._gappedMnvReferenceCounts at position (point alleles)
    int32_t locus;             // locus index of `position` in the segment (ref_valid is cleared there when the allele is called), or -1
    int32_t pad1[3];
};
static_assert(sizeof(DevCand) == 64, "DevCand layout");
constexpr uint8_t kCandAltHasN = 1;       // AlternateAllele contains 'N' (stitched-source strand-bias filter, AlleleProcessor.cs:66-69)
constexpr uint8_t kCandReportable = 2;    // AlleleCaller.ShouldReport: position inside the interval set (or no intervals)
constexpr uint8_t kCandForced = 4;        // AlleleCaller.IsForcedAllele: reported even when not callable, with the ForcedReport filter (:108-118)

// What CoverageCalculator.CalculateSpanning (CoverageCalculator.cs:162-321) reads from the counts; independent of the candidate's support, so the
// host collapser can re-derive a candidate's coverage (-> CandidateAllele.Frequency) after every merge without another launch.
struct SpanIngredients {
    int32_t sp[3], ep[3];          // start / end point coverage by direction (anchor-selected for insertions)
    int32_t spu[3], epu[3];        // the unanchored remainder (insertions)
    int32_t conf_l, conf_r, susp_l, susp_r;
};
struct SpanCoverage { int32_t cov[3]; int32_t total; float weight; };

// The arithmetic tail of CalculateSpanning (:255-320): unanchored-coverage weighting for insertions, stitched redistribution (:324-331),
// per-direction (start+end)/2f or min, truncations. Float arithmetic as in the reference.
__host__ __device__ inline SpanCoverage spanning_tail(const SpanIngredients& g, bool picky, bool presume_anchored, int allele_support, int well_anchored) {
    int sp[3], ep[3];
    for (int d = 0; d < 3; d++) { sp[d] = g.sp[d]; ep[d] = g.ep[d]; }
    SpanCoverage out;
    out.weight = 0.0f;
    if (picky) {
        const int unanchored_support = allele_support - well_anchored;
        const bool use_un = unanchored_support > 0;   // the unanchored sums are only collected then (:222)
        const int susp_l = use_un ? g.susp_l : 0, susp_r = use_un ? g.susp_r : 0;
        const float truly = (((g.conf_l - susp_r) + (g.conf_r - susp_l)) / 2.0f);
        const float anchored_vf = truly <= 0 ? 0.0f : (float)well_anchored / truly;
        const int total_susp = susp_l + susp_r;
        const float unanchored_vf = total_susp == 0 ? 0.0f : (float)unanchored_support / ((float)total_susp);
        float w = anchored_vf == 0 ? 1.0f : (unanchored_vf / anchored_vf < 1.0f ? unanchored_vf / anchored_vf : 1.0f);
        if (!(w > 0.0f)) w = 0.0f;
        out.weight = w;
        for (int d = 0; d < 3; d++) {
            sp[d] += (int)((float)(use_un ? g.spu[d] : 0) * w);
            ep[d] += (int)((float)(use_un ? g.epu[d] : 0) * w);
        }
    }
    for (int k = 0; k < 2; k++) {   // RedistributeStitchedCoverage
        int* dp = k == 0 ? sp : ep;
        const int stitched = dp[2];
        dp[0] += (stitched + 1) / 2;   // (int)Math.Ceiling((float)stitched / 2), stitched >= 0
        dp[1] += stitched / 2;         // (int)Math.Floor((float)stitched / 2)
        dp[2] = 0;
    }
    float exact_total = 0.0f;
    for (int d = 0; d < 2; d++) {
        const float exact = presume_anchored ? ((sp[d] + ep[d])) / 2.0f : (float)(sp[d] < ep[d] ? sp[d] : ep[d]);
        out.cov[d] = (int)exact;
        exact_total += exact;
    }
    out.cov[2] = 0;
    out.total = (int)exact_total;
    return out;
}

struct CandScoreArgs {
    const DevCand* cands;
    int32_t n;
    const int32_t* counts;        // [n_req][198] RegionState order [allele][direction][anchor]
    const int32_t* collapsed;     // [n_req][8] or nullptr
    const double* qsum;           // [n_req][198] or nullptr
    const uint8_t* arena;
    const uint8_t* chr_seq;
    int64_t chr_len;
    const double* q_to_p_table;
    int q_table_max;
    pb2_call_record* out_dense;   // [n]: every candidate's record (host-orchestrated passes), or nullptr
    uint8_t* out_callable;        // [n] bit0 AlleleCaller.IsCallable, bit1 && ShouldReport, bit2 forced to report; or nullptr
    SpanIngredients* out_ingredients;  // [n] (spanning alleles) or nullptr
    pb2_call_record* var_records; // append mode: callable alleles go to the variant stream ...
    unsigned long long* var_count;
    int64_t var_capacity;
    uint8_t* ref_valid;           // ... and the reference record of their position is pruned (AlleleCaller.cs:146-147)
    int indel_repeat_filter;
};

cudaError_t launch_gather_locus_counts(const TilePileup& in, const int32_t* req_locus, int32_t n_req, int32_t* out_counts, int32_t* out_collapsed,
                                       double* out_qsum, int min_bq, cudaStream_t stream);
cudaError_t launch_prune_ref_valid(const DevCand* cands, const uint8_t* flags, int32_t n, uint8_t* ref_valid, cudaStream_t stream);
cudaError_t launch_score_candidates(const CandScoreArgs& args, const DeviceConfig& cfg, cudaStream_t stream);

}  // namespace pb2
