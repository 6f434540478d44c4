// ORACLE — TEST INFRASTRUCTURE ONLY (see po_math.hpp header). Types mirroring the reference's domain model.
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

namespace po {

// src/lib/Pisces.Domain/Types/AlleleType.cs:5-10 — NOTE order A,G,C,T,N,Del
enum AlleleType : int { AT_A = 0, AT_G = 1, AT_C = 2, AT_T = 3, AT_N = 4, AT_Del = 5 };
constexpr int NumAlleleTypes = 6;      // Constants.cs:18-26
constexpr int NumDirectionTypes = 3;   // Constants.cs:28-35
constexpr int NumReadCollapsedTypes = 8;
// Constants.cs:40-43 CoverageContributingAlleles = {A, C, G, T, Deletion} (iteration order matters for double sums)
static const AlleleType CoverageContributingAlleles[5] = {AT_A, AT_C, AT_G, AT_T, AT_Del};

enum DirectionType : int { Forward = 0, Reverse = 1, Stitched = 2 };  // Types/DirectionType.cs
// Types/AlleleCategory.cs
enum AlleleCategory : int { Snv = 0, Insertion = 1, Deletion = 2, Mnv = 3, Reference = 4, NonReference = 5, Unsupported = 6 };
// Types/FilterType.cs
enum FilterType : int { F_StrandBias = 0, F_PoolBias, F_AmpliconBias, F_LowVariantQscore, F_LowDepth, F_LowVariantFrequency,
                        F_LowGenotypeQuality, F_IndelRepeatLength, F_MultiAllelicSite, F_RMxN, F_ForcedReport, F_OffTarget,
                        F_NoCall, F_Unknown };
// Types/Genotype.cs
enum Genotype : int { HeterozygousAlt1Alt2 = 0, Alt12LikeNoCall, HeterozygousAltRef, HomozygousAlt, HomozygousRef, RefLikeNoCall,
                      AltLikeNoCall, RefAndNoCall, AltAndNoCall, HemizygousRef, HemizygousAlt, HemizygousNoCall, Others };
// Types/ReadCollapsedType.cs
enum ReadCollapsedType : int { DuplexStitched = 0, DuplexNonStitched = 1, SimplexStitched = 2, SimplexNonStitched = 3,
                               SimplexForwardStitched = 4, SimplexForwardNonStitched = 5, SimplexReverseStitched = 6,
                               SimplexReverseNonStitched = 7 };
enum StrandBiasModel : int { SBM_Poisson = 0, SBM_Extended = 1, SBM_Diploid = 2 };
enum NoiseModel : int { NM_Flat = 0, NM_Window = 1 };
enum PloidyModel : int { PM_Somatic = 0, PM_DiploidByThresholding = 1, PM_DiploidByAdaptiveGT = 2, PM_Haploid = 3 };

// Utility/AlleleHelper.cs:13-32
inline AlleleType GetAlleleType(char c) {
    switch (c) {
        case 'A': return AT_A;
        case 'C': return AT_C;
        case 'G': return AT_G;
        case 'T': return AT_T;
        default: return AT_N;
    }
}

// StrandBiasStats.cs
struct StrandBiasStats {
    double ChanceFalseNeg = 0, ChanceFalsePos = 0, ChanceVarFreqGreaterThanZero = 0, Coverage = 0, Frequency = 0, Support = 0;
    StrandBiasStats() {}
    StrandBiasStats(double support, double coverage) {
        Frequency = support / coverage;
        Support = support;
        Coverage = coverage;
        if (coverage == 0) Frequency = 0;
    }
};
struct BiasResults {
    bool BiasAcceptable = false;
    double BiasScore = 0, GATKBiasScore = 0;
    bool VarPresentOnBothStrands = false, CovPresentOnBothStrands = false;
    StrandBiasStats ForwardStats, OverallStats, ReverseStats, StitchedStats;
};

// Models/Alleles/CandidateAllele.cs
// AmpliconCounts (src/lib/Pisces.Domain/Models/AmpliconCounts.cs:34-110): names are small integers (the XN dictionary), -1 = null slot; isNull: the
// AmpliconNames array itself is null. Slots are filled in first-seen order, Constants.MaxNumOverlappingAmplicons = 6 of them (Constants.cs:54-62).
constexpr int MaxNumOverlappingAmplicons = 6;
struct AmpliconCounts {
    bool isNull = true;
    std::array<int, MaxNumOverlappingAmplicons> names{{-1, -1, -1, -1, -1, -1}};
    std::array<int, MaxNumOverlappingAmplicons> counts{{0, 0, 0, 0, 0, 0}};
    static AmpliconCounts Empty() { AmpliconCounts a; a.isNull = false; return a; }
    // GetAmpliconNameIndex (:48-66): (index of the name, first empty slot), -1 where there is none
    std::pair<int, int> Index(int name) const {
        int firstEmpty = -1;
        for (int i = 0; i < MaxNumOverlappingAmplicons; i++) {
            if (names[(size_t)i] == name) return {i, -1};
            if (names[(size_t)i] < 0 && firstEmpty == -1) firstEmpty = i;
        }
        return {-1, firstEmpty};
    }
    void Add(int name, int count) {   // the merge of RegionState.AddCandidate (:138-170) / CandidateAllele.AddSupport (:82-103) / RegionState.AddAmpliconCount (:269-307)
        auto ix = Index(name);
        if (ix.first > -1) counts[(size_t)ix.first] += count;
        else {
            if (ix.second < 0) throw std::out_of_range("Index was outside the bounds of the array.");   // more than 6 overlapping amplicons: the reference throws
            names[(size_t)ix.second] = name; counts[(size_t)ix.second] += count;
        }
    }
};

struct CandidateAllele {
    std::string Chromosome;
    int ReferencePosition = 0;
    std::string ReferenceAllele, AlternateAllele;
    AlleleCategory Type = Reference;
    std::array<int, 3> SupportByDirection{{0, 0, 0}};
    std::array<int, 3> WellAnchoredSupportByDirection{{0, 0, 0}};
    std::array<int, 8> ReadCollapsedCountsMut{{0, 0, 0, 0, 0, 0, 0, 0}};
    bool OpenOnRight = false, OpenOnLeft = false, IsKnown = false, IsForcedAllele = false;
    float Frequency = 0;
    AmpliconCounts SupportByAmplicon;   // CandidateAllele.cs:23
    CandidateAllele() {}
    CandidateAllele(const std::string& chr, int coord, const std::string& ref, const std::string& alt, AlleleCategory t)
        : Chromosome(chr), ReferencePosition(coord), ReferenceAllele(ref), AlternateAllele(alt), Type(t) {
        if (chr.empty()) throw std::invalid_argument("Chromosome is empty.");
        if (coord < 0) throw std::invalid_argument("Coordinate is invalid.");
        if (ref.empty()) throw std::invalid_argument("Reference is empty.");
        if (alt.empty()) throw std::invalid_argument("Alternate is empty.");
    }
    int Support() const { return SupportByDirection[0] + SupportByDirection[1] + SupportByDirection[2]; }
    int WellAnchoredSupport() const { return WellAnchoredSupportByDirection[0] + WellAnchoredSupportByDirection[1] + WellAnchoredSupportByDirection[2]; }
    bool FullyAnchored() const { return !OpenOnLeft && !OpenOnRight; }
    // CandidateAllele.Equals :56-66
    bool Equals(const CandidateAllele& o) const {
        return o.ReferencePosition == ReferencePosition && o.AlternateAllele == AlternateAllele && o.Type == Type &&
               o.Chromosome == Chromosome && o.ReferenceAllele == ReferenceAllele;
    }
    int Length() const;  // BaseAllele.Length
    void AddSupport(const CandidateAllele& from) {  // :74-106
        for (int i = 0; i < 3; i++) SupportByDirection[i] += from.SupportByDirection[i];
        for (int i = 0; i < 3; i++) WellAnchoredSupportByDirection[i] += from.WellAnchoredSupportByDirection[i];
        if (!from.SupportByAmplicon.isNull) {
            if (SupportByAmplicon.isNull) SupportByAmplicon = AmpliconCounts::Empty();
            // (:89-103 walks all six slots of the source, null names included: a null name "matches" the first empty slot of the target and adds 0)
            for (int i = 0; i < MaxNumOverlappingAmplicons; i++)
                if (from.SupportByAmplicon.names[(size_t)i] >= 0) SupportByAmplicon.Add(from.SupportByAmplicon.names[(size_t)i], from.SupportByAmplicon.counts[(size_t)i]);
        }
    }
};
inline int AlleleLength(AlleleCategory t, const std::string& ref, const std::string& alt) {  // BaseAllele.cs:24-43
    switch (t) {
        case Mnv: case Snv: return (int)alt.size();
        case Insertion: return (int)alt.size() - 1;
        case Deletion: return (int)ref.size() - 1;
        case Reference: return (int)ref.size();
        default: throw std::invalid_argument("Unrecognized allele type");
    }
}
inline int CandidateAllele::Length() const { return AlleleLength(Type, ReferenceAllele, AlternateAllele); }

// Models/Alleles/CalledAllele.cs
struct CalledAllele {
    std::string Chromosome;
    int ReferencePosition = 0;
    std::string ReferenceAllele, AlternateAllele;
    AlleleCategory Type = Reference;
    Genotype genotype = HomozygousRef;
    int GenotypeQscore = 0, VariantQscore = 0;
    std::vector<FilterType> Filters;
    BiasResults StrandBiasResults;
    int NoiseLevelApplied = 0;
    int TotalCoverage = 0;
    double SumOfBaseQuality = 0;
    std::array<int, 3> EstimatedCoverageByDirection{{0, 0, 0}};
    std::array<int, 8> ReadCollapsedCountTotal{{0, 0, 0, 0, 0, 0, 0, 0}};
    std::array<int, 8> ReadCollapsedCountsMut{{0, 0, 0, 0, 0, 0, 0, 0}};
    std::array<int, 3> SupportByDirection{{0, 0, 0}};
    std::array<int, 3> WellAnchoredSupportByDirection{{0, 0, 0}};
    int AlleleSupport = 0, NumNoCalls = 0;
    float FractionNoCalls = 0;
    bool IsForcedToReport = false;
    int ConfidentCoverageStart = 0, SuspiciousCoverageStart = 0, ConfidentCoverageEnd = 0, SuspiciousCoverageEnd = 0;
    int WellAnchoredSupport = 0;
    double UnanchoredCoverageWeight = 0;
    int ReferenceSupport = 0;
    AmpliconCounts SupportByAmplicon, CoverageByAmplicon;   // CalledAllele.cs:35-36
    bool AmpliconBiasDetected = false, HasAmpliconBiasResults = false;   // AmpliconBiasResults != null / .BiasDetected (CalledAllele.cs:20)

    CalledAllele() {}
    explicit CalledAllele(AlleleCategory t) : Type(t) { genotype = (t == Reference) ? HomozygousRef : HeterozygousAltRef; }  // :148-161

    float Frequency() const {  // :49-52
        if (TotalCoverage == 0) return 0.0f;
        float f = (float)AlleleSupport / (float)TotalCoverage;
        return f < 1.0f ? f : 1.0f;
    }
    float RefFrequency() const {  // :123-126
        if (TotalCoverage == 0) return 0.0f;
        float f = (float)ReferenceSupport / (float)TotalCoverage;
        return f < 1.0f ? f : 1.0f;
    }
    bool IsNocall() const { return genotype == Alt12LikeNoCall || genotype == AltLikeNoCall || genotype == HemizygousNoCall || genotype == RefLikeNoCall; }
    void SetFractionNoCalls() {  // :107-114
        float allReads = (float)(TotalCoverage + NumNoCalls);
        if (allReads == 0) FractionNoCalls = 0;
        else FractionNoCalls = ((float)NumNoCalls / allReads);
    }
    void AddFilter(FilterType f) {  // :116-119
        for (auto x : Filters) if (x == f) return;
        Filters.push_back(f);
    }
    int Length() const { return AlleleLength(Type, ReferenceAllele, AlternateAllele); }
};

// AlleleHelper.Map  (Utility/AlleleHelper.cs:34-85)
inline CalledAllele MapToCalled(const CandidateAllele& c) {
    CalledAllele a(c.Type);
    a.AlternateAllele = c.AlternateAllele;
    a.ReferenceAllele = c.ReferenceAllele;
    a.Chromosome = c.Chromosome;
    a.ReferencePosition = c.ReferencePosition;
    a.AlleleSupport = c.Support();
    a.WellAnchoredSupport = c.WellAnchoredSupport();
    a.IsForcedToReport = c.IsForcedAllele;
    a.SupportByDirection = c.SupportByDirection;
    a.WellAnchoredSupportByDirection = c.WellAnchoredSupportByDirection;
    if (c.Type != Reference) a.ReadCollapsedCountsMut = c.ReadCollapsedCountsMut;
    if (!c.SupportByAmplicon.isNull) a.SupportByAmplicon = c.SupportByAmplicon;   // AlleleHelper.Map :72-82
    return a;
}
inline CandidateAllele MapToCandidate(const CalledAllele& a) {
    CandidateAllele c(a.Chromosome, a.ReferencePosition, a.ReferenceAllele, a.AlternateAllele, a.Type);
    c.SupportByDirection = a.SupportByDirection;
    c.WellAnchoredSupportByDirection = a.WellAnchoredSupportByDirection;
    if (a.Type != Reference) c.ReadCollapsedCountsMut = a.ReadCollapsedCountsMut;
    return c;
}

struct Region { int StartPosition = 0, EndPosition = 0; };

// Subset of PiscesApplicationOptions / VariantCallingParameters / BamFilterParameters / VcfWritingParameters that the hot path
// reads (Factory.cs:123-227; defaults from Options/*.cs cited in SURVEY.md §5).
struct Config {
    // BamFilterParameters.cs:7-11
    int MinimumBaseCallQuality = 20;
    int MinimumMapQuality = 1;
    bool RemoveDuplicates = true;
    bool OnlyUseProperPairs = false;
    // VariantCallingParameters.cs:59-107 (+ derived :109-178)
    float MinimumFrequency = 0.01f;
    float MinimumFrequencyFilter = -1;  // raised to MinimumFrequency if below (Validate :144-147)
    float TargetLODFrequency = -1;      // raised to MinimumFrequencyFilter if below (:152-155)
    int MaximumVariantQScore = 100, MinimumVariantQScore = 20, MinimumVariantQScoreFilter = 30;
    int MaximumGenotypeQScore = 100, MinimumGenotypeQScore = 0;
    int LowGenotypeQualityFilter = -1;  // <0 = null
    int MinimumCoverage = 10;
    int LowDepthFilter = -1;            // <0 = null -> set to MinimumCoverage by Validate
    int IndelRepeatFilter = -1;         // <0 = null
    int RMxNFilterMaxLengthRepeat = 5, RMxNFilterMinRepetitions = 9;  // <0 = null
    float RMxNFilterFrequencyLimit = 0.35f;
    int ploidy = PM_Somatic;
    float DiploidMinorVF = 0.20f, DiploidMajorVF = 0.70f, DiploidSumVFforMultiAllelicSite = 0.80f;   // DiploidSNVThresholdingParameters (:84)
    int IsMale = -1;                    // bool? IsMale: -1 = null
    float AmpliconBiasFilterThreshold = -1;   // float? (VariantCallingParameters.cs:102-107): < 0 = null; tracking follows it (Factory.ShouldTrackAmpliconCounts :51-54)
    bool TrackAmpliconCounts() const { return AmpliconBiasFilterThreshold >= 0; }
    int ForcedNoiseLevel = -1;          // NL = ForcedNoiseLevel == -1 ? MinimumBaseCallQuality : ForcedNoiseLevel  (:109-118)
    int noiseModel = NM_Flat;
    float StrandBiasAcceptanceCriteria = 0.5f;
    int strandBiasModel = SBM_Extended;
    bool FilterOutVariantsPresentOnlyOneStrand = false;
    float NoCallFilterThreshold = 0.6f;  // <0 = null
    // PiscesApplicationOptions.cs:43-66
    bool CallMNVs = false;
    int MaxSizeMNV = 3, MaxGapBetweenMNV = 1;
    bool Collapse = true;
    float CollapseFreqThreshold = 0.0f, CollapseFreqRatioThreshold = 0.5f;
    bool ExcludeMNVsFromCollapsing = false;
    int TrackedAnchorSize = 5;
    // VcfWritingParameters.cs:7
    bool OutputGvcfFile = true;
    // source properties (BamFileAlignmentExtractor.cs:111-153)
    bool SourceIsStitched = false, SourceIsCollapsed = false;
    bool ApplyValidation = true;   // false: options object built by hand without Validate() (e.g. SomaticVariantCallerFunctionalTests.cs:683-758)

    int NoiseLevelUsedForQScoring() const { return ForcedNoiseLevel == -1 ? MinimumBaseCallQuality : ForcedNoiseLevel; }
    void Validate() {
        if (!ApplyValidation) return;
        if (MinimumFrequencyFilter < MinimumFrequency) MinimumFrequencyFilter = MinimumFrequency;
        if (TargetLODFrequency < MinimumFrequencyFilter) TargetLODFrequency = MinimumFrequencyFilter;
        if (LowDepthFilter < MinimumCoverage) LowDepthFilter = MinimumCoverage;  // checked against VariantCallingParameters.Validate below
    }
};

}  // namespace po
