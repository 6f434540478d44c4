// ORACLE — TEST INFRASTRUCTURE ONLY (see po_math.hpp header).
// Calculators: src/lib/Pisces.Calculators/{CoverageCalculator,CollapsedCoverageCalculator,VariantQualityCalculator,StrandBiasCalculator,
// RMxNCalculator}.cs ; src/exe/Pisces/Logic/VariantCalling/AlleleProcessor.cs ; src/lib/Pisces.Genotyping/Somatic/*.cs
#pragma once
#include "po_math.hpp"
#include "po_state.hpp"

namespace po {

// ------------------------------------------------------------------ CoverageCalculator.cs
struct CoverageCalculator {
    bool considerAnchorInformation;
    bool collapsedCalculator;  // CollapsedCoverageCalculator (Factory.cs:193-199)
    bool trackAmpliconCoverage = false;   // the reference asks the source for CoverageByAmplicon unconditionally (:52); only a tracking run ever reads the answer
    explicit CoverageCalculator(bool considerAnchor = false, bool collapsed = false) : considerAnchorInformation(considerAnchor), collapsedCalculator(collapsed) {}

    void Compute(CalledAllele& allele, IAlleleSource& src) const {  // :19-47 (+ CollapsedCoverageCalculator.cs:18-37)
        if (allele.Type == Reference) CalculateSinglePoint(allele, src);
        else {
            switch (allele.Type) {
                case Deletion: CalculateSpanning(allele, src, allele.ReferencePosition + 1, allele.ReferencePosition + allele.Length(), true); break;
                case Mnv: CalculateSpanning(allele, src, allele.ReferencePosition, allele.ReferencePosition + allele.Length() - 1, true); break;
                case Insertion: CalculateSpanning(allele, src, allele.ReferencePosition, allele.ReferencePosition + 1, src.ExpectStitchedReads()); break;
                default: CalculateSinglePoint(allele, src); break;
            }
        }
    }
    void CalculateSinglePoint(CalledAllele& allele, IAlleleSource& src) const {  // :49-98
        if (trackAmpliconCoverage) allele.CoverageByAmplicon = src.GetCoverageByAmplicon(allele.ReferencePosition);   // :52, :334-337
        if (collapsedCalculator) {  // CollapsedCoverageCalculator.CalculateSinglePoint :23-31
            for (int t = 0; t < NumReadCollapsedTypes; t++) allele.ReadCollapsedCountTotal[t] += src.GetCollapsedReadCount(allele.ReferencePosition, (ReadCollapsedType)t);
        }
        AlleleType refType = GetAlleleType(allele.ReferenceAllele.size() == 1 ? allele.ReferenceAllele[0] : '?');
        for (int direction = 0; direction < NumDirectionTypes; direction++) {
            for (AlleleType alleleType : CoverageContributingAlleles) {
                allele.EstimatedCoverageByDirection[direction] += src.GetAlleleCount(allele.ReferencePosition, alleleType, (DirectionType)direction);
                allele.SumOfBaseQuality += src.GetSumOfAlleleBaseQualities(allele.ReferencePosition, alleleType, (DirectionType)direction);
                if (alleleType != refType) continue;
                allele.ReferenceSupport += src.GetAlleleCount(allele.ReferencePosition, alleleType, (DirectionType)direction);
            }
            allele.TotalCoverage += allele.EstimatedCoverageByDirection[direction];
            allele.ConfidentCoverageStart += allele.EstimatedCoverageByDirection[direction];
            allele.ConfidentCoverageEnd += allele.EstimatedCoverageByDirection[direction];
            allele.NumNoCalls += src.GetAlleleCount(allele.ReferencePosition, AT_N, (DirectionType)direction);
        }
        int gappedRefCounts = src.GetGappedMnvRefCount(allele.ReferencePosition);
        if (allele.Type == Snv) allele.ReferenceSupport = std::max(0, allele.ReferenceSupport - gappedRefCounts);
        else if (allele.Type == Reference) allele.AlleleSupport = std::max(0, allele.AlleleSupport - gappedRefCounts);
    }
    static void RedistributeStitchedCoverage(int* dp) {  // :324-331
        int stitched = dp[Stitched];
        dp[Forward] += (int)std::ceil((float)stitched / 2);
        dp[Reverse] += (int)std::floor((float)stitched / 2);
        dp[Stitched] = 0;
    }
    void CalculateSpanning(CalledAllele& allele, IAlleleSource& src, int startPointPosition, int endPointPosition, bool presumeAnchoredForExactCov) const {  // :162-321
        if (collapsedCalculator) {  // CollapsedCoverageCalculator.CalculateSpanning :28-37 — always the start point; += (order vs base call is immaterial)
            for (int t = 0; t < NumReadCollapsedTypes; t++) allele.ReadCollapsedCountTotal[t] += src.GetCollapsedReadCount(startPointPosition, (ReadCollapsedType)t);
        }
        int startPointCoverage[3] = {0, 0, 0}, endPointCoverage[3] = {0, 0, 0};
        float exactTotalCoverage = 0.0f;
        int confidentCoverageLeft = 0, confidentCoverageRight = 0, suspiciousCoverageLeft = 0, suspiciousCoverageRight = 0;
        AlleleType firstBase = AT_N, lastBase = AT_N;
        bool bePickyAboutAnchors = considerAnchorInformation && allele.Type == Insertion;
        if (bePickyAboutAnchors) {
            firstBase = GetAlleleType(allele.AlternateAllele[1]);
            lastBase = GetAlleleType(allele.AlternateAllele[allele.AlternateAllele.size() - 1]);
        }
        int startPointCoverageUnanchored[3] = {0, 0, 0}, endPointCoverageUnanchored[3] = {0, 0, 0};
        double unanchoredCoverageStartQuality = 0, unanchoredCoverageEndQuality = 0;
        int unanchoredSupport = allele.AlleleSupport - allele.WellAnchoredSupport;
        for (int d = 0; d < NumDirectionTypes; d++) {
            for (AlleleType alleleType : CoverageContributingAlleles) {
                bool anchoredCoverageOnlyEnd = bePickyAboutAnchors && alleleType == firstBase;
                bool anchoredCoverageOnlyStart = bePickyAboutAnchors && alleleType == lastBase;
                int minAnchorEnd = anchoredCoverageOnlyEnd ? allele.Length() : 0;
                int minAnchorStart = anchoredCoverageOnlyStart ? allele.Length() : 0;
                int s = src.GetAlleleCount(startPointPosition, alleleType, (DirectionType)d, minAnchorStart);
                startPointCoverage[d] += s;
                int e = src.GetAlleleCount(endPointPosition, alleleType, (DirectionType)d, minAnchorEnd, std::nullopt, true);
                endPointCoverage[d] += e;
                confidentCoverageLeft += s;
                confidentCoverageRight += e;
                allele.SumOfBaseQuality += src.GetSumOfAlleleBaseQualities(startPointPosition, alleleType, (DirectionType)d, minAnchorStart);
                allele.SumOfBaseQuality += src.GetSumOfAlleleBaseQualities(endPointPosition, alleleType, (DirectionType)d, minAnchorEnd, std::nullopt, true);
                if (bePickyAboutAnchors && unanchoredSupport > 0) {
                    if (minAnchorStart > 0) {
                        int c = src.GetAlleleCount(startPointPosition, alleleType, (DirectionType)d, 0, minAnchorStart - 1);
                        startPointCoverageUnanchored[d] += c;
                        suspiciousCoverageLeft += c;
                        unanchoredCoverageStartQuality += src.GetSumOfAlleleBaseQualities(startPointPosition, alleleType, (DirectionType)d, 0, minAnchorStart - 1);
                    }
                    if (minAnchorEnd > 0) {
                        int c = src.GetAlleleCount(endPointPosition, alleleType, (DirectionType)d, 0, minAnchorEnd - 1, true);
                        endPointCoverageUnanchored[d] += c;
                        suspiciousCoverageRight += c;
                        // NB reads startPointPosition (reference quirk, :253)
                        unanchoredCoverageEndQuality += src.GetSumOfAlleleBaseQualities(startPointPosition, alleleType, (DirectionType)d, 0, minAnchorEnd - 1, true);
                    }
                }
            }
        }
        if (bePickyAboutAnchors) {
            float trulyAnchoredCoverage = (((confidentCoverageLeft - suspiciousCoverageRight) + (confidentCoverageRight - suspiciousCoverageLeft)) / 2.0f);
            float anchoredVariantFreq = trulyAnchoredCoverage <= 0 ? 0.0f : (float)allele.WellAnchoredSupport / trulyAnchoredCoverage;
            int totalSuspiciousCoverage = suspiciousCoverageLeft + suspiciousCoverageRight;
            float unanchoredVariantFreq = totalSuspiciousCoverage == 0 ? 0.0f : (float)unanchoredSupport / ((float)totalSuspiciousCoverage);
            float w = std::max(0.0f, anchoredVariantFreq == 0 ? 1.0f : std::min(1.0f, unanchoredVariantFreq / anchoredVariantFreq));
            allele.UnanchoredCoverageWeight = w;
            for (int d = 0; d < NumDirectionTypes; d++) {
                startPointCoverage[d] += (int)((float)startPointCoverageUnanchored[d] * w);
                endPointCoverage[d] += (int)((float)endPointCoverageUnanchored[d] * w);
                allele.SumOfBaseQuality += unanchoredCoverageStartQuality * (double)w;
                allele.SumOfBaseQuality += unanchoredCoverageEndQuality * (double)w;
            }
        }
        RedistributeStitchedCoverage(startPointCoverage);
        RedistributeStitchedCoverage(endPointCoverage);
        for (int d = 0; d < 2; d++) {
            float exact = presumeAnchoredForExactCov ? ((startPointCoverage[d] + endPointCoverage[d])) / 2.0f
                                                     : (float)std::min(startPointCoverage[d], endPointCoverage[d]);
            allele.EstimatedCoverageByDirection[d] = (int)exact;
            exactTotalCoverage += exact;
        }
        allele.TotalCoverage = (int)exactTotalCoverage;
        allele.ReferenceSupport = std::max(0, allele.TotalCoverage - allele.AlleleSupport);
        allele.SuspiciousCoverageStart = suspiciousCoverageLeft;
        allele.ConfidentCoverageStart = confidentCoverageLeft;
        allele.SuspiciousCoverageEnd = suspiciousCoverageRight;
        allele.ConfidentCoverageEnd = confidentCoverageRight;
    }
};

// ------------------------------------------------------------------ VariantQualityCalculator.cs
inline double AssignRawPoissonQScore(int callCount, int coverage, int estimatedBaseCallQuality) {  // :27-52
    double errorRate = QtoP(estimatedBaseCallQuality);
    double callCountMinusOne = callCount - 1;
    double callCountDouble = callCount;
    double lambda = errorRate * coverage;
    double pValue = 1 - mathnet::PoissonCumulativeDistribution(lambda, callCountMinusOne);
    if (pValue > 0) return PtoQ(pValue);
    double A = mathnet::PoissonProbabilityLn(lambda, (int)callCountMinusOne);
    double correction = (callCountDouble - lambda) / callCountDouble;
    return -10.0 * (A - std::log(2.0 * correction)) / std::log(10.0);
}
inline int AssignPoissonQScore(int callCount, int coverage, int estimatedBaseCallQuality, int maxQScore) {  // :54-65
    if ((callCount <= 0) || (coverage <= 0)) return 0;
    double rawQ = AssignRawPoissonQScore(callCount, coverage, estimatedBaseCallQuality);
    double q = std::min((double)maxQScore, rawQ);
    q = std::max(q, 0.0);
    return (int)std::nearbyint(q);  // Math.Round = half-to-even
}
inline double AssignPValue(int observedCallCount, int coverage, int estimatedBaseCallQuality) {  // :67-74
    double errorRate = QtoP(estimatedBaseCallQuality);
    if (observedCallCount == 0) return 1.0;
    return (1 - pisces_poisson::Cdf(observedCallCount - 1.0, coverage * errorRate));
}
inline void VariantQualityCompute(CalledAllele& a, int maxQScore, int estimatedBaseCallQuality) {  // :11-24
    a.NoiseLevelApplied = estimatedBaseCallQuality;
    if (a.TotalCoverage == 0) a.VariantQscore = 0;
    else a.VariantQscore = AssignPoissonQScore(a.AlleleSupport, a.TotalCoverage, estimatedBaseCallQuality, maxQScore);
}

// ------------------------------------------------------------------ StrandBiasCalculator.cs
// MathNet Binomial(p, n).CumulativeDistribution — only the Diploid SB model needs it; restated in po_diploid.hpp when built.
double MathNetBinomialCdf(double p, int n, double x);

inline void PopulateStats(StrandBiasStats& st, double noiseFreq, double minDetectableSNP, int model) {  // :175-231
    if (st.Support == 0) {
        if (model == SBM_Poisson) { st.ChanceFalsePos = 1; st.ChanceVarFreqGreaterThanZero = 0; st.ChanceFalseNeg = 0; }
        else {
            st.ChanceVarFreqGreaterThanZero = std::pow(1 - minDetectableSNP, st.Coverage);
            st.ChanceFalsePos = 1 - st.ChanceVarFreqGreaterThanZero;
            st.ChanceFalseNeg = st.ChanceVarFreqGreaterThanZero;
        }
    } else if (model == SBM_Diploid) {  // PopulateDiploidStats :150-173
        if (st.Frequency >= minDetectableSNP) { st.ChanceFalseNeg = 1; st.ChanceFalsePos = 0; st.ChanceVarFreqGreaterThanZero = 1; return; }
        st.ChanceFalseNeg = std::max(MathNetBinomialCdf(minDetectableSNP, (int)st.Coverage, st.Support), 0.0);
        st.ChanceFalsePos = std::max(0.0, 1 - pisces_poisson::Cdf(st.Support, st.Coverage * 0.1));
        st.ChanceVarFreqGreaterThanZero = st.ChanceFalseNeg;
    } else {
        st.ChanceVarFreqGreaterThanZero = std::max(0.0, pisces_poisson::Cdf(st.Support - 1, st.Coverage * noiseFreq));
        st.ChanceFalsePos = std::max(0.0, 1 - st.ChanceVarFreqGreaterThanZero);
        st.ChanceFalseNeg = std::max(0.0, pisces_poisson::Cdf(st.Support, st.Coverage * minDetectableSNP));
    }
}
inline StrandBiasStats CreateStats(double support, double coverage, double noiseFreq, double minDetectableSNP, int model) {  // :137-148
    if (model != SBM_Diploid) minDetectableSNP = noiseFreq;
    StrandBiasStats st(support, coverage);
    PopulateStats(st, noiseFreq, minDetectableSNP, model);
    return st;
}
inline BiasResults CalculateStrandBiasResults(const int* cov, const int* sup, int qNoise, double minVariantFreq, double acceptanceCriteria, int model) {  // :21-72
    int fS = sup[Forward], fC = cov[Forward], rS = sup[Reverse], rC = cov[Reverse], sS = sup[Stitched], sC = cov[Stitched];
    double errorRate = std::pow(10.0, (double)((float)(-1 * qNoise) / 10.0f));  // :32 float exponent
    BiasResults r;
    r.OverallStats = CreateStats(fS + rS + sS, fC + rC + sC, errorRate, minVariantFreq, model);
    r.ForwardStats = CreateStats(fS + sS / 2, fC + sC / 2, errorRate, minVariantFreq, model);
    r.ReverseStats = CreateStats(rS + sS / 2, rC + sC / 2, errorRate, minVariantFreq, model);
    r.StitchedStats = CreateStats(sS, sC, errorRate, minVariantFreq, model);
    // AssignBiasScore :89-105
    double forwardBias = (r.ForwardStats.ChanceVarFreqGreaterThanZero * r.ReverseStats.ChanceFalsePos) / r.OverallStats.ChanceVarFreqGreaterThanZero;
    double reverseBias = (r.ReverseStats.ChanceVarFreqGreaterThanZero * r.ForwardStats.ChanceFalsePos) / r.OverallStats.ChanceVarFreqGreaterThanZero;
    if (r.OverallStats.ChanceVarFreqGreaterThanZero == 0) { forwardBias = 1; reverseBias = 1; }
    // Math.Max(double,double): returns NaN if either is NaN
    double p = (std::isnan(forwardBias) || std::isnan(reverseBias)) ? std::numeric_limits<double>::quiet_NaN() : std::max(forwardBias, reverseBias);
    r.BiasScore = p;
    r.GATKBiasScore = PtoGATKBiasScale(p);
    r.CovPresentOnBothStrands = ((r.ForwardStats.Coverage > 0) && (r.ReverseStats.Coverage > 0));
    r.VarPresentOnBothStrands = ((r.ForwardStats.Support > 0) && (r.ReverseStats.Support > 0));
    if (!r.CovPresentOnBothStrands) { r.BiasScore = 0; r.GATKBiasScore = -std::numeric_limits<double>::infinity(); }
    r.BiasAcceptable = (r.BiasScore < acceptanceCriteria);
    return r;
}

// ------------------------------------------------------------------ RMxNCalculator.cs
inline bool CompareSubstring(const std::string& needle, const std::string& hay, int start) {  // Pisces.IO VcfVariantUtilities.CompareSubstring
    if (start < 0 || start + (int)needle.size() > (int)hay.size()) return false;
    return hay.compare((size_t)start, needle.size(), needle) == 0;
}
inline int ComputeRMxNLengthForIndel(int variantPosition, const std::string& variantBases, const std::string& referenceBases, int maxRepeatUnitLength) {  // :49-95
    int maxRepeatsFound = 0;
    std::vector<std::string> bookends;
    int length = (int)variantBases.size();
    std::vector<std::string> prefixes, suffixes;
    for (int i = length - std::min(maxRepeatUnitLength, length); i < length; i++) {
        prefixes.push_back(variantBases.substr(0, length - i));
        suffixes.push_back(variantBases.substr(i, length - i));
    }
    bookends = prefixes;
    bookends.insert(bookends.end(), suffixes.begin(), suffixes.end());
    for (auto& bookend : bookends) {
        int backPeekPosition = variantPosition;
        while (true) {
            int newBack = backPeekPosition - (int)bookend.size();
            if (newBack < 0) break;
            if (!CompareSubstring(bookend, referenceBases, newBack)) break;
            backPeekPosition = newBack;
        }
        int repeatCount = 0, currentPosition = backPeekPosition;
        while (true) {
            if (currentPosition + (int)bookend.size() > (int)referenceBases.size()) break;
            if (!CompareSubstring(bookend, referenceBases, currentPosition)) break;
            repeatCount++;
            currentPosition += (int)bookend.size();
        }
        if (repeatCount > maxRepeatsFound) maxRepeatsFound = repeatCount;
    }
    return maxRepeatsFound;
}
inline std::pair<int, int> ComputeComponentRMxNLengths(const CalledAllele& a, const std::string& referenceBases, int maxRepeatUnitLength) {  // :104-133
    int component1 = 0, component2 = INT32_MAX;
    std::string variantBases = (a.Type == Mnv || a.Type == Snv) ? a.AlternateAllele : a.Type == Insertion ? a.AlternateAllele.substr(1) : a.ReferenceAllele.substr(1);
    if (a.Type == Insertion || a.Type == Deletion) {
        component1 = ComputeRMxNLengthForIndel(a.ReferencePosition, variantBases, referenceBases, maxRepeatUnitLength);
    } else {
        component1 = ComputeRMxNLengthForIndel(a.ReferencePosition - 1, a.ReferenceAllele, referenceBases, maxRepeatUnitLength);
        int c1 = ComputeRMxNLengthForIndel(a.ReferencePosition + (int)a.ReferenceAllele.size() - 1, variantBases, referenceBases, maxRepeatUnitLength);
        int c2 = ComputeRMxNLengthForIndel(a.ReferencePosition - 1, variantBases, referenceBases, maxRepeatUnitLength);
        component2 = std::max(c1, c2);
    }
    return {component1, component2};
}
inline bool RMxNShouldFilter(const CalledAllele& a, const Config& cfg, const std::string& referenceSequence) {  // :19-38
    if (a.Frequency() >= cfg.RMxNFilterFrequencyLimit) return false;
    if (cfg.RMxNFilterMaxLengthRepeat >= 0 && cfg.RMxNFilterMinRepetitions >= 0) {
        auto m = ComputeComponentRMxNLengths(a, referenceSequence, cfg.RMxNFilterMaxLengthRepeat);
        if (std::min(m.first, m.second) >= cfg.RMxNFilterMinRepetitions) return true;
    }
    return false;
}

// ------------------------------------------------------------------ AlleleProcessor.cs
int ComputeIndelRepeatLength(const CalledAllele& allele, const std::string& referenceBases);  // :80-213, po_caller.hpp

// ---------------------------------------------------------------- AmpliconBiasCalculator.cs (src/lib/Pisces.Calculators/AmpliconBiasCalculator.cs)
// Amplicon names are small integers here (a host-side dictionary of the XN tag strings, Read.cs:479-495); -1 stands for a null name.
struct AmpliconBiasEntry {   // AmpliconBiasResult (Models/AmpliconCounts.cs:13-23)
    int name = -1;
    double frequency = 0, coverage = 0, observedSupport = 0, expectedSupport = 0, chanceItsReal = 0;
    int confidenceQScore = 0;
    bool biasDetected = false;
};
struct AmpliconBiasResults {  // BiasResultsAcrossAmplicons (:6-11); isNull: CalculateAmpliconBias returned null
    bool isNull = true, biasDetected = false;
    int ampliconWithCandidateArtifact = -1;
    std::vector<AmpliconBiasEntry> results;   // Dictionary enumeration order = insertion order
};
inline int CsDoubleToInt(double v) { return (v > -2147483649.0 && v < 2147483648.0) ? (int)v : (int)0x80000000; }   // (int)double of the CLR on x64
// CalculateAmpliconBias (:45-133). Arrays may be shorter than Constants.MaxNumOverlappingAmplicons (RegionState.GetCountsByAmpliconForPosition trims
// them, RegionState.cs:325-352); nSupport < 0 stands for a null AmpliconNames array.
inline AmpliconBiasResults CalculateAmpliconBias(const int* supportNames, const int* supportCounts, int nSupport, const int* coverageNames, const int* coverageCounts,
                                                 int nCoverage, float acceptanceCriteria, int maxQScore) {
    const int MinNumObservations = 5;             // Constants (:15-19)
    const double FreePassObservationFreq = 0.1;
    AmpliconBiasResults out;
    if (nSupport <= 0 || supportNames[0] < 0) return out;   // :50-54
    if (nCoverage < 2) return out;                           // :57-59
    out.isNull = false;
    double maxFreq = 0.0;
    for (int i = 0; i < nCoverage; i++) {                    // :64-82
        const int name = coverageNames[i];
        if (name < 0) break;
        double support = 0;                                  // AmpliconCounts.GetCountsForAmplicon (AmpliconCounts.cs:74-81): first equal name, else 0
        for (int k = 0; k < nSupport; k++) if (supportNames[k] == name) { support = supportCounts[k]; break; }
        const double coverage = coverageCounts[i];
        const double freq = (coverage > 0) ? support / coverage : 0;
        if (freq >= maxFreq) { out.ampliconWithCandidateArtifact = name; maxFreq = freq; }
        AmpliconBiasEntry e;
        e.name = name; e.frequency = freq; e.observedSupport = support; e.coverage = coverage;
        out.results.push_back(e);
    }
    bool shouldFailVariant = false;
    for (auto& e : out.results) {                            // :84-129
        int qScore = 0;
        bool biasDetected = false;
        const float allowableProb = acceptanceCriteria;
        const double expectedNumObservationsOfVariant = maxFreq * e.coverage;
        double pChanceItsReal = 1.0;
        if (expectedNumObservationsOfVariant < MinNumObservations) qScore = maxQScore;
        else if ((expectedNumObservationsOfVariant <= e.observedSupport) || (e.frequency > FreePassObservationFreq)) qScore = maxQScore;
        else {
            pChanceItsReal = std::max(0.0, pisces_poisson::Cdf(e.observedSupport, expectedNumObservationsOfVariant));
            qScore = CsDoubleToInt(PtoQ(1.0 - pChanceItsReal));
        }
        if (pChanceItsReal < (double)allowableProb) { biasDetected = true; shouldFailVariant = true; }
        e.chanceItsReal = pChanceItsReal; e.confidenceQScore = qScore; e.biasDetected = biasDetected; e.expectedSupport = expectedNumObservationsOfVariant;
        out.biasDetected = shouldFailVariant;
    }
    return out;
}

}  // namespace po
