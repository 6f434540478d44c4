// ORACLE — TEST INFRASTRUCTURE ONLY. Not shipped, not linked into libpisces_b200.so, never on the product path.
//
// CPU restatement of the germline genotypers of the per-locus calling path, reference @ /root/reference:
//   GenotypeCalculatorUtilities            src/lib/Pisces.Genotyping/GenotypeCalculatorUtilities.cs:11-237
//   DiploidThresholdingGenotyper           src/lib/Pisces.Genotyping/Thresholding/DiploidThresholdingGenotyper.cs:53-141
//   DiploidGenotypeQualityCalculator       src/lib/Pisces.Genotyping/Thresholding/DiploidGenotypeQualityCalculator.cs:17-103
//   HaploidGenotyper                       src/lib/Pisces.Genotyping/Haploid/HaploidGenotyper.cs:38-84
//   HaploidGenotypeQualityCalculator       src/lib/Pisces.Genotyping/Haploid/HaploidGenotypeQualityCalculator.cs:12-59
//   GenotypeCreator.GetPloidyForThisChr    src/lib/Pisces.Genotyping/GenotypeCreator.cs:39-68
//   DiploidLocusProcessor                  src/exe/Pisces/Logic/VariantCalling/DiploidLocusProcessor.cs:13-51
// MathNet.Numerics 4.5.1 Binomial::PMFLn / ::CDF, SpecialFunctions::BinomialLn / ::BetaRegularized are restated from the IL of the
// shipped MathNet.Numerics.dll (oracle/tools/il_dump.py), like the rest of po_math.hpp.
// Parity pin: tests/test_oracle_kats.py re-asserts DiploidGenotypeQualityCalculatorTests.cs:15-109, GenotypeCalculatorTest.cs:27-88,
// HaploidGenotypeCalculatorTests.cs:55-83, StrandBiasCalculatorTests.cs:157-174 (Binomial(0.2,100) CDF).
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <memory>
#include <vector>
#include "po_math.hpp"
#include "po_types.hpp"

namespace po {
using CalledPtr = std::shared_ptr<CalledAllele>;

namespace mathnet {
// SpecialFunctions::BinomialLn(n, k)
inline double BinomialLn(int n, int k) {
    if (k < 0 || n < 0 || k > n) return -std::numeric_limits<double>::infinity();
    return FactorialLn(n) - FactorialLn(k) - FactorialLn(n - k);
}
// Distributions.Binomial::PMFLn(p, n, k)  (= Binomial(p, n).ProbabilityLn(k))
inline double BinomialProbabilityLn(double p, int n, int k) {
    if (k < 0 || k > n) return -std::numeric_limits<double>::infinity();
    if (p == 0.0) return k == 0 ? 0.0 : -std::numeric_limits<double>::infinity();
    if (p == 1.0) return k == n ? 0.0 : -std::numeric_limits<double>::infinity();
    return BinomialLn(n, k) + (double)k * std::log(p) + (double)(n - k) * std::log(1.0 - p);
}
// SpecialFunctions::BetaRegularized(a, b, x): Lentz continued fraction; DoublePrecision = 2^-53, Precision.Increment(0.0, 1) = 4.94e-324
inline double BetaRegularized(double a, double b, double x) {
    const double bt = (x == 0.0 || x == 1.0) ? 0.0 : std::exp(GammaLn(a + b) - GammaLn(a) - GammaLn(b) + a * std::log(x) + b * std::log(1.0 - x));
    const bool symmetryTransformation = x >= (a + 1.0) / (a + b + 2.0);
    const double eps = 1.1102230246251565e-16;      // Precision.DoublePrecision = 2^-53
    const double fpmin = 4.9406564584124654e-324 / eps;   // 0.0.Increment() / eps
    if (symmetryTransformation) { x = 1.0 - x; std::swap(a, b); }
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0;
    double d = 1.0 - qab * x / qap;
    if (std::fabs(d) < fpmin) d = fpmin;
    d = 1.0 / d;
    double h = d;
    for (int m = 1, m2 = 2; m <= 50000; m++, m2 += 2) {
        double aa = (double)m * (b - (double)m) * x / ((qam + (double)m2) * (a + (double)m2));
        d = 1.0 + aa * d;
        if (std::fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c;
        if (std::fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + (double)m) * (qab + (double)m) * x / ((a + (double)m2) * (qap + (double)m2));
        d = 1.0 + aa * d;
        if (std::fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c;
        if (std::fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (std::fabs(del - 1.0) <= eps) return symmetryTransformation ? 1.0 - bt * h / a : bt * h / a;
    }
    return symmetryTransformation ? 1.0 - bt * h / a : bt * h / a;
}
// Distributions.Binomial::CDF(p, n, x)
inline double BinomialCdf(double p, int n, double x) {
    if (x < 0.0) return 0.0;
    if (x > (double)n) return 1.0;
    const double k = std::floor(x);
    return BetaRegularized((double)n - k, k + 1.0, 1.0 - p);
}
}  // namespace mathnet

// (int) of a double in C# (unchecked, x64): NaN and out-of-range values become int.MinValue
inline int CsIntCast(double v) {
    if (std::isnan(v) || v >= 2147483648.0 || v <= -2147483649.0) return INT_MIN;
    return (int)v;
}

struct DiploidThresholdingParameters { float MinorVF = 0.20f, MajorVF = 0.70f, SumVFforMultiAllelicSite = 0.80f; };
enum SimplifiedDiploidGenotype { SDG_HomozygousRef, SDG_HeterozygousAltRef, SDG_HomozygousAlt };

// GenotypeCreator.GetPloidyForThisChr (:39-68); isMale: -1 = null
inline int GetPloidyForThisChr(int samplePloidy, int isMale, const std::string& refName) {
    if (samplePloidy == PM_Somatic || refName == "chrM" || refName == "M") return PM_Somatic;
    if (samplePloidy == PM_Haploid) return PM_Haploid;
    if (isMale < 0) return samplePloidy;
    if (isMale > 0 && (refName == "chrY" || refName == "chrX" || refName == "Y" || refName == "X")) return PM_Haploid;
    if (isMale == 0 && (refName == "chrY" || refName == "Y")) return PM_Haploid;
    return samplePloidy;
}

namespace gtutil {
// GetAllelesToPruneBasedOnGTCall (:11-46)
inline void GetAllelesToPruneBasedOnGTCall(Genotype gt, const std::vector<CalledPtr>& orderedVariants, std::vector<CalledPtr>& allelesToPrune) {
    int allowed = 0;
    switch (gt) {
        case AltAndNoCall: case AltLikeNoCall: case HomozygousAlt: case HeterozygousAltRef: case HemizygousAlt: allowed = 1; break;
        case Alt12LikeNoCall: case HeterozygousAlt1Alt2: allowed = 2; break;
        default: allowed = 0; break;
    }
    for (int i = 0; i < (int)orderedVariants.size(); i++) if (i >= allowed) allelesToPrune.push_back(orderedVariants[(size_t)i]);
}
inline bool CheckForDepthIssue(const std::vector<CalledPtr>& alleles, int minDepthToEmit) {  // :48-58
    for (auto& a : alleles) if (a->TotalCoverage < minDepthToEmit) return true;
    return false;
}
// FilterAndOrderAllelesByFrequency (:60-82): OrderByDescending(Frequency).ThenBy(AlleleCompareByLociAndAllele) — a stable sort
inline std::vector<CalledPtr> FilterAndOrderAllelesByFrequency(const std::vector<CalledPtr>& alleles, std::vector<CalledPtr>& allelesToPrune, double minFreqThreshold) {
    std::vector<CalledPtr> variantAlleles;
    for (auto& a : alleles) {
        if (a->Type == Reference) continue;
        if ((double)a->Frequency() >= minFreqThreshold) variantAlleles.push_back(a);
        else allelesToPrune.push_back(a);
    }
    std::stable_sort(variantAlleles.begin(), variantAlleles.end(), [](const CalledPtr& x, const CalledPtr& y) {
        if (x->Frequency() != y->Frequency()) return x->Frequency() > y->Frequency();
        if (x->ReferencePosition != y->ReferencePosition) return x->ReferencePosition < y->ReferencePosition;
        if (x->ReferenceAllele != y->ReferenceAllele) return x->ReferenceAllele < y->ReferenceAllele;
        return x->AlternateAllele < y->AlternateAllele;
    });
    return variantAlleles;
}
inline double GetReferenceFrequency(const std::vector<CalledPtr>& alleles, double /*minorVF*/) {  // :85-130
    double altFrequencyCount = 0, refFrequencyCountBySNP = 0, indelFrequencyCount = 0;
    if (alleles.empty()) return 0;
    if (alleles.size() == 1) return alleles.front()->RefFrequency();
    for (auto& a : alleles) {
        if (a->Type == Reference) return a->Frequency();
        altFrequencyCount += a->Frequency();
        if (a->Type == Snv) refFrequencyCountBySNP = a->RefFrequency();
        else indelFrequencyCount += a->Frequency();
    }
    return std::max(refFrequencyCountBySNP - indelFrequencyCount, 0.0);
}
inline bool CheckForTriAllelicIssue(bool hasReference, double referenceFreq, const std::vector<CalledPtr>& variantAlleles, float threshold) {  // :133-150
    if (variantAlleles.back()->Type != Snv) return false;
    if (hasReference && (((double)variantAlleles[0]->Frequency() + referenceFreq) < (double)threshold)) return true;
    return ((variantAlleles[0]->Frequency() + variantAlleles[1]->Frequency()) < threshold);
}
inline Genotype ConvertSimpleGenotypeToComplexGenotype(const std::vector<CalledPtr>& alleles, const std::vector<CalledPtr>& orderedVariants, double referenceFrequency,
                                                       bool refExists, bool depthIssue, bool refCall, float minVarFrequency, float sumVFforMultiAllelicSite,
                                                       SimplifiedDiploidGenotype preliminaryGenotype) {  // :160-233
    if (depthIssue) return refCall ? RefLikeNoCall : AltLikeNoCall;
    switch (preliminaryGenotype) {
        case SDG_HomozygousRef: {
            if (!refExists) return RefLikeNoCall;
            auto& first = alleles.front();
            if (first->Type == Reference && ((1 - first->Frequency()) > minVarFrequency)) return RefAndNoCall;
            return HomozygousRef;
        }
        case SDG_HeterozygousAltRef:
            if (orderedVariants.size() == 1) return refExists ? HeterozygousAltRef : AltAndNoCall;
            if (CheckForTriAllelicIssue(refExists, referenceFrequency, orderedVariants, sumVFforMultiAllelicSite)) {
                for (auto& a : alleles) a->Filters.push_back(F_MultiAllelicSite);   // SetMultiAllelicFilter: Filters.Add (no duplicate check)
                return refExists ? AltLikeNoCall : Alt12LikeNoCall;
            }
            return refExists ? HeterozygousAltRef : HeterozygousAlt1Alt2;
        default:
            return HomozygousAlt;
    }
}
}  // namespace gtutil

// DiploidGenotypeQualityCalculator.Compute (:17-103)
inline int DiploidGenotypeQuality(const CalledAllele& allele, int minQScore, int maxQScore) {
    if (allele.TotalCoverage == 0) return minQScore;
    const float noiseHomRef = 0.05f, noiseHomAlt = 0.075f, noiseHetAlt = 0.10f, expectedHetFreq = 0.40f;
    const float depth = (float)allele.TotalCoverage;
    const double lamHomRef = (double)(noiseHomRef * depth), lamHomAlt = (double)(noiseHomAlt * depth);
    const double pHetExpected = (double)expectedHetFreq, pHomRefNoise = (double)noiseHetAlt, pHomAltNoise = (double)(1 - noiseHetAlt);
    const int n = allele.TotalCoverage;
    const int nonAlleleCalls = std::max(allele.TotalCoverage - allele.AlleleSupport, 0);
    double LnPofH0GT = 0, LnPofH1GT = 0;
    switch (allele.genotype) {
        case HomozygousRef:
            LnPofH0GT = mathnet::PoissonProbabilityLn(lamHomRef, nonAlleleCalls);
            LnPofH1GT = mathnet::BinomialProbabilityLn(pHetExpected, n, nonAlleleCalls);
            break;
        case HomozygousAlt:
            LnPofH0GT = mathnet::PoissonProbabilityLn(lamHomAlt, nonAlleleCalls);
            LnPofH1GT = mathnet::BinomialProbabilityLn(pHetExpected, n, allele.AlleleSupport);
            break;
        case HeterozygousAlt1Alt2:
        case HeterozygousAltRef: {
            const int k = (int)(depth * allele.Frequency());
            LnPofH0GT = mathnet::BinomialProbabilityLn(pHetExpected, n, k);
            LnPofH1GT = ((double)allele.Frequency() >= 0.50) ? mathnet::BinomialProbabilityLn(pHomAltNoise, n, k) : mathnet::BinomialProbabilityLn(pHomRefNoise, n, k);
            break;
        }
        default:
            return minQScore;
    }
    const int qScore = CsIntCast(std::floor(10.0 * std::log10(M_E) * (LnPofH0GT - LnPofH1GT)));
    if ((LnPofH1GT <= (double)INT_MIN) && (LnPofH0GT > LnPofH1GT)) return maxQScore;
    if ((LnPofH0GT <= (double)INT_MIN) && (LnPofH0GT < LnPofH1GT)) return minQScore;
    return std::max(std::min(qScore, maxQScore), minQScore);
}
// HaploidGenotypeQualityCalculator.Compute (:12-59)
inline int HaploidGenotypeQuality(const CalledAllele& allele, int minQScore, int maxQScore) {
    if (allele.TotalCoverage == 0) return minQScore;
    const float noiseHomRef = 0.05f, noiseHomAlt = 0.075f, expectedHetFreq = 0.40f;
    const float depth = (float)allele.TotalCoverage;
    const double lamHomRef = (double)(noiseHomRef * depth), lamHomAlt = (double)(noiseHomAlt * depth);
    const int n = allele.TotalCoverage;
    const int nonAlleleCalls = std::max(allele.TotalCoverage - allele.AlleleSupport, 0);
    double LnPofH0GT = 0, LnPofH1GT = 0;
    switch (allele.genotype) {
        case HemizygousRef:
            LnPofH0GT = mathnet::PoissonProbabilityLn(lamHomRef, nonAlleleCalls);
            LnPofH1GT = mathnet::BinomialProbabilityLn((double)expectedHetFreq, n, nonAlleleCalls);
            break;
        case HemizygousAlt:
            LnPofH0GT = mathnet::PoissonProbabilityLn(lamHomAlt, nonAlleleCalls);
            LnPofH1GT = mathnet::BinomialProbabilityLn((double)expectedHetFreq, n, allele.AlleleSupport);
            break;
        default:
            return minQScore;
    }
    const int qScore = CsIntCast(std::floor(10.0 * std::log10(M_E) * (LnPofH0GT - LnPofH1GT)));
    return std::max(std::min(qScore, maxQScore), minQScore);
}

// DiploidThresholdingGenotyper.SetGenotypes (:53-75) with CalculateDiploidGenotype (:77-100); GenotypeCreator passes the SNV parameters for
// indels too (GenotypeCreator.cs:28). Returns the alleles to prune.
inline std::vector<CalledPtr> DiploidSetGenotypes(const std::vector<CalledPtr>& alleles, int minDepthToGenotype, const DiploidThresholdingParameters& snv,
                                                  const DiploidThresholdingParameters& indel, int minGQ, int maxGQ) {
    std::vector<CalledPtr> allelesToPrune;
    auto orderedVariants = gtutil::FilterAndOrderAllelesByFrequency(alleles, allelesToPrune, (double)snv.MinorVF);
    const double referenceFrequency = gtutil::GetReferenceFrequency(alleles, (double)snv.MinorVF);
    const bool refExists = referenceFrequency >= (double)snv.MinorVF;
    const bool depthIssue = gtutil::CheckForDepthIssue(alleles, minDepthToGenotype);
    const bool refCall = orderedVariants.empty() || (orderedVariants[0]->Frequency() < snv.MinorVF);
    const DiploidThresholdingParameters& parameters = (refCall || orderedVariants.front()->Type == Snv) ? snv : indel;   // SelectParameters (:127-139)
    SimplifiedDiploidGenotype prelim;   // GetPreliminaryGenotype (:102-125)
    if (refCall) prelim = SDG_HomozygousRef;
    else if (orderedVariants[0]->Frequency() >= parameters.MinorVF && orderedVariants[0]->Frequency() <= parameters.MajorVF) prelim = SDG_HeterozygousAltRef;
    else if (orderedVariants[0]->Frequency() > parameters.MajorVF) prelim = SDG_HomozygousAlt;
    else prelim = SDG_HomozygousRef;
    const Genotype gt = gtutil::ConvertSimpleGenotypeToComplexGenotype(alleles, orderedVariants, referenceFrequency, refExists, depthIssue, refCall, parameters.MinorVF,
                                                                       parameters.SumVFforMultiAllelicSite, prelim);
    gtutil::GetAllelesToPruneBasedOnGTCall(gt, orderedVariants, allelesToPrune);
    for (auto& a : alleles) {
        a->genotype = gt;
        a->GenotypeQscore = DiploidGenotypeQuality(*a, minGQ, maxGQ);
    }
    return allelesToPrune;
}
// HaploidGenotyper.SetGenotypes (:38-52) with CalculateHaploidGenotype (:54-82)
inline std::vector<CalledPtr> HaploidSetGenotypes(const std::vector<CalledPtr>& alleles, int minDepthToGenotype, float minorVF, float majorVF, int minGQ, int maxGQ) {
    std::vector<CalledPtr> allelesToPrune;
    Genotype gt = HemizygousNoCall;
    auto orderedVariants = gtutil::FilterAndOrderAllelesByFrequency(alleles, allelesToPrune, (double)minorVF);
    const double referenceFrequency = gtutil::GetReferenceFrequency(alleles, (double)minorVF);
    const bool refExists = referenceFrequency >= (double)minorVF;
    const bool depthIssue = gtutil::CheckForDepthIssue(alleles, minDepthToGenotype);
    const bool refCall = orderedVariants.empty() || (orderedVariants[0]->Frequency() < minorVF);
    if (!depthIssue && refCall && refExists && referenceFrequency > (double)majorVF) gt = HemizygousRef;
    if (!depthIssue && !refCall && !refExists && orderedVariants[0]->Frequency() > majorVF) gt = HemizygousAlt;
    gtutil::GetAllelesToPruneBasedOnGTCall(gt, orderedVariants, allelesToPrune);
    for (auto& a : alleles) {
        a->genotype = gt;
        a->GenotypeQscore = HaploidGenotypeQuality(*a, minGQ, maxGQ);
    }
    return allelesToPrune;
}

// DiploidLocusProcessor.Process (:13-51)
inline void DiploidLocusProcess(std::vector<CalledPtr>& at) {
    std::vector<CalledPtr> forced, nonForced;
    for (auto& a : at) {
        bool isForced = false;
        for (auto f : a->Filters) if (f == F_ForcedReport) isForced = true;
        (isForced ? forced : nonForced).push_back(a);
    }
    if (forced.empty()) return;
    bool isRef = false, anyNoCall = false;
    for (auto& v : nonForced) { isRef |= v->Type == Reference; anyNoCall |= v->IsNocall(); }
    const bool isNoCall = nonForced.empty() || anyNoCall;
    const Genotype gt = isNoCall ? AltLikeNoCall : (isRef ? HomozygousRef : Others);
    for (auto& f : forced) f->genotype = gt;
    int minGQ = 0;
    if (!nonForced.empty()) { minGQ = nonForced.front()->GenotypeQscore; for (auto& v : nonForced) minGQ = std::min(minGQ, v->GenotypeQscore); }
    for (auto& a : at) a->GenotypeQscore = minGQ;
}

}  // namespace po
