// ORACLE — TEST INFRASTRUCTURE ONLY (see po_math.hpp header).
// Caller: src/exe/Pisces/Logic/VariantCalling/{AlleleCaller,AlleleProcessor,VariantCollapser,MnvReallocator,SomaticLocusProcessor}.cs,
// src/exe/Pisces/Logic/SmallVariantCaller.cs, src/lib/Pisces.Genotyping/Somatic/{SomaticGenotyper,SomaticGenotypeQualityCalculator}.cs
#pragma once
#include "po_calc.hpp"
#include "po_finder.hpp"
#include "po_genotype.hpp"

namespace po {

// ------------------------------------------------------------------ AlleleProcessor.cs :80-213 (indel repeat filter; off by default)
inline std::string SimplifyRepeatUnit(const std::string& repeatUnit) {  // :140-155
    if (repeatUnit.empty()) return "";
    std::string sb = repeatUnit.substr(0, 1);
    for (size_t i = 1; i < repeatUnit.size(); i++) {
        // repeatUnit.Split(sb).Length - 1 == number of non-overlapping occurrences scanning left to right
        size_t occurrences = 0, pos = 0;
        while ((pos = repeatUnit.find(sb, pos)) != std::string::npos) { occurrences++; pos += sb.size(); }
        if (repeatUnit.size() == occurrences * sb.size()) break;
        sb.push_back(repeatUnit[i]);
    }
    return sb;
}
inline int GetRepeatLength(const std::string& bases, int currentPos, const std::string& repeatUnit) {  // :160-213
    int numRepeatBases = (int)repeatUnit.size();
    if (numRepeatBases == 0) return 0;
    int lastPosition = (int)bases.size() - numRepeatBases - 1;
    int requiredLength = currentPos + numRepeatBases + 1;
    if (requiredLength > (int)bases.size()) return 1;
    int previousPos = currentPos;
    while (currentPos > 0) {
        bool match = true;
        for (int index = 0; index < numRepeatBases; index++) if (bases[currentPos + index] != repeatUnit[index]) { match = false; break; }
        if (!match) break;
        previousPos = currentPos;
        currentPos -= numRepeatBases;
    }
    currentPos = previousPos;
    int repeatLength = 0;
    while (currentPos <= lastPosition) {
        bool match = true;
        for (int index = 0; index < numRepeatBases; index++) if (bases[currentPos + index] != repeatUnit[index]) { match = false; break; }
        if (!match) break;
        currentPos += numRepeatBases;
        repeatLength++;
    }
    return repeatLength;
}
inline int ComputeIndelRepeatLength(const CalledAllele& allele, const std::string& referenceBases) {  // :80-135
    const int FlankingBaseCount = 50;
    if (referenceBases.empty()) return 0;
    if (allele.Type != Insertion && allele.Type != Deletion && allele.Type != Snv) return 0;
    int stringPos = allele.ReferencePosition - 1;
    int upstreamBegin = stringPos - FlankingBaseCount, upstreamEnd = stringPos - 1, downstreamBegin = stringPos, downstreamEnd = stringPos + FlankingBaseCount - 1;
    if (upstreamBegin < 0) upstreamBegin = 0;
    if (downstreamBegin < 0) downstreamBegin = 0;
    if (downstreamEnd >= (int)referenceBases.size()) downstreamEnd = (int)referenceBases.size() - 1;
    if (upstreamEnd >= (int)referenceBases.size()) upstreamEnd = (int)referenceBases.size() - 1;
    std::string upstream;
    if (upstreamEnd >= 0) upstream = referenceBases.substr(upstreamBegin, upstreamEnd - upstreamBegin + 1);
    std::string downstream = referenceBases.substr(downstreamBegin, downstreamEnd - downstreamBegin + 1);
    int currentPosition = (int)upstream.size();
    std::string variantBases;
    if (allele.Type == Insertion) { variantBases = allele.AlternateAllele.substr(1); currentPosition++; }
    if (allele.Type == Deletion) { variantBases = allele.ReferenceAllele.substr(1); currentPosition++; }
    return GetRepeatLength(upstream + downstream, currentPosition, SimplifyRepeatUnit(variantBases));
}

// AlleleProcessor.Process / ApplyFilters :16-71
inline void AlleleProcessorProcess(CalledAllele& a, const Config& cfg, const std::string& chrSeq, bool isStitchedSource, float variantFreqFilter) {
    a.SetFractionNoCalls();
    a.Filters.clear();
    if (cfg.LowDepthFilter >= 0 && a.TotalCoverage < cfg.LowDepthFilter) a.AddFilter(F_LowDepth);
    if (a.VariantQscore < cfg.MinimumVariantQScoreFilter && a.TotalCoverage != 0) a.AddFilter(F_LowVariantQscore);
    if (a.Type != Reference) {
        if (cfg.NoCallFilterThreshold >= 0 && a.FractionNoCalls > cfg.NoCallFilterThreshold) a.AddFilter(F_NoCall);
        if (!a.StrandBiasResults.BiasAcceptable || (cfg.FilterOutVariantsPresentOnlyOneStrand && !a.StrandBiasResults.VarPresentOnBothStrands)) a.AddFilter(F_StrandBias);
        if (a.HasAmpliconBiasResults && a.AmpliconBiasDetected && cfg.AmpliconBiasFilterThreshold >= 0) a.AddFilter(F_AmpliconBias);   // :49-50
        if (cfg.IndelRepeatFilter > 0) {
            if (cfg.IndelRepeatFilter <= ComputeIndelRepeatLength(a, chrSeq)) a.AddFilter(F_IndelRepeatLength);
        }
        if (RMxNShouldFilter(a, cfg, chrSeq)) a.AddFilter(F_RMxN);
        // VariantFreqFilter = genotypeCalculator.MinVarFrequencyFilter (Factory.cs:166) = max(MinimumFrequencyFilter, MinVarFrequency) (SetMinFreqFilter)
        if (a.Frequency() < variantFreqFilter) a.AddFilter(F_LowVariantFrequency);
        if (isStitchedSource && a.AlternateAllele.find('N') != std::string::npos) a.AddFilter(F_StrandBias);
    }
}

// ------------------------------------------------------------------ SomaticGenotyper.cs :51-100, SomaticGenotypeQualityCalculator.cs :10-48
inline Genotype CalculateSomaticGenotype(const CalledAllele& a, float minFrequencyFilter, int minDepthToGenotype) {
    if (a.TotalCoverage < minDepthToGenotype) return (a.Type == Reference) ? RefLikeNoCall : AltLikeNoCall;
    if (a.Type != Reference) {
        if (a.RefFrequency() < minFrequencyFilter) {
            if ((1 - a.Frequency()) > minFrequencyFilter) return AltAndNoCall;
            return HomozygousAlt;
        }
        return HeterozygousAltRef;
    }
    if (a.Frequency() < minFrequencyFilter) return RefLikeNoCall;
    if ((1 - a.Frequency()) > minFrequencyFilter) return RefAndNoCall;
    return HomozygousRef;
}
inline int SomaticGenotypeQuality(const CalledAllele& a, float targetLimitOfDetectionVF, int minGTQScore, int maxGTQScore) {
    double rawQ = a.VariantQscore;
    if ((a.TotalCoverage == 0) || a.IsNocall()) return minGTQScore;
    if ((a.genotype == HomozygousRef) || (a.genotype == HomozygousAlt)) {
        double p1 = QtoP(a.VariantQscore);
        float nonAlleleObservationsF = (1.0f - a.Frequency()) * (float)a.TotalCoverage;
        float expectedNonAllelObservationsF = targetLimitOfDetectionVF * (float)a.TotalCoverage;
        if (nonAlleleObservationsF >= expectedNonAllelObservationsF) return minGTQScore;
        double p2 = pisces_poisson::Cdf(nonAlleleObservationsF, expectedNonAllelObservationsF);
        rawQ = PtoQ(p1 + p2);
    }
    double q = std::min((double)maxGTQScore, rawQ);
    q = std::max(q, (double)minGTQScore);
    return (int)std::nearbyint(q);
}

// ------------------------------------------------------------------ VariantCollapser.cs
struct VariantCollapser {
    int TotalNumCollapsed = 0;
    const CoverageCalculator* cov;
    float freqThreshold, freqRatioThreshold;
    bool excludeMNVs;
    VariantCollapser(const CoverageCalculator* c, float ft, float frt, bool excl) : cov(c), freqThreshold(ft), freqRatioThreshold(frt), excludeMNVs(excl) {}

    static bool CanCollapse(const CandidateAllele& toCollapse, const CandidateAllele& pm) {  // :125-175
        if ((toCollapse.Type == Insertion && pm.Type != Insertion) || (toCollapse.Type != Insertion && pm.Type == Insertion) ||
            (toCollapse.Type == Deletion && pm.Type != Deletion) || (toCollapse.Type != Deletion && pm.Type == Deletion) ||
            toCollapse.Length() > pm.Length() || (toCollapse.FullyAnchored() && !pm.FullyAnchored()))
            return false;
        const std::string& tcb = toCollapse.Type == Deletion ? toCollapse.ReferenceAllele : toCollapse.AlternateAllele;
        const std::string& pmb = pm.Type == Deletion ? pm.ReferenceAllele : pm.AlternateAllele;
        if (toCollapse.FullyAnchored() && pm.FullyAnchored()) return toCollapse.Equals(pm);
        if (toCollapse.Type == Deletion) {
            if (toCollapse.OpenOnRight) return pm.ReferencePosition + 1 == toCollapse.ReferencePosition + 1;
            return pm.ReferencePosition + (int)pmb.size() - 1 == toCollapse.ReferencePosition + (int)tcb.size() - 1;
        }
        if (toCollapse.OpenOnRight) return pm.ReferencePosition == toCollapse.ReferencePosition && pmb.substr(0, tcb.size()) == tcb;
        if (toCollapse.Type == Insertion)
            return pm.ReferencePosition + 1 == toCollapse.ReferencePosition + 1 && pmb.substr(pmb.size() - tcb.size() + 1) == tcb.substr(1);
        return pm.ReferencePosition + (int)pm.AlternateAllele.size() - 1 == toCollapse.ReferencePosition + (int)toCollapse.AlternateAllele.size() - 1 &&
               pm.AlternateAllele.substr(pm.AlternateAllele.size() - toCollapse.AlternateAllele.size()) == toCollapse.AlternateAllele;
    }
    static int Compare(const CandidateAllele& first, const CandidateAllele& second) {  // :221-245
        if (first.IsKnown && !second.IsKnown) return -1;
        if (!first.IsKnown && second.IsKnown) return 1;
        if (first.FullyAnchored() && !second.FullyAnchored()) return -1;
        if (!first.FullyAnchored() && second.FullyAnchored()) return 1;
        if (first.Length() != second.Length()) return first.Length() < second.Length() ? 1 : -1;
        if (std::fabs(first.Frequency - second.Frequency) > 0.0f) return first.Frequency < second.Frequency ? 1 : -1;
        if (first.ReferencePosition != second.ReferencePosition) return first.ReferencePosition < second.ReferencePosition ? -1 : 1;
        int c = first.AlternateAllele.compare(second.AlternateAllele);
        return c < 0 ? -1 : (c > 0 ? 1 : 0);
    }
    CandPtr GetMatches(const CandPtr& toCollapse, const std::vector<CandPtr>& targets, IAlleleSource& source) {  // :191-219
        std::vector<CandPtr> pms;
        for (auto& c : targets) if (CanCollapse(*toCollapse, *c) && c != toCollapse) pms.push_back(c);
        if (pms.empty()) return nullptr;
        for (auto& v : pms) {
            CalledAllele cv = MapToCalled(*v);
            cov->Compute(cv, source);
            v->Frequency = cv.Frequency();
        }
        CalledAllele tc = MapToCalled(*toCollapse);
        cov->Compute(tc, source);
        // List.Sort(IComparer) is an unstable introsort in .NET; elements comparing 0 here are same pos/alt/length/freq — stable order kept
        std::stable_sort(pms.begin(), pms.end(), [](const CandPtr& a, const CandPtr& b) { return Compare(*a, *b) < 0; });
        for (auto& m : pms) if (m->Equals(*toCollapse) && !m->OpenOnLeft && !m->OpenOnRight) return m;
        float tcf = tc.Frequency();
        for (auto& m : pms) if (m->Frequency >= freqThreshold && m->Frequency / tcf > freqRatioThreshold) return m;
        return nullptr;
    }
    // Collapse :31-113  (candidates list is modified in place and returned)
    void Collapse(std::vector<CandPtr>& candidates, IAlleleSource& source, std::optional<int> maxClearedPosition) {
        std::vector<CandPtr> targetVariants;
        for (auto& c : candidates) if (!excludeMNVs || c->Type != Mnv) targetVariants.push_back(c);
        // AnnotateKnown: priors out of scope (knownVariants == null)
        std::vector<CandPtr> toCollapse;
        for (auto& v : targetVariants) if (v->OpenOnLeft || v->OpenOnRight) toCollapse.push_back(v);
        std::stable_sort(toCollapse.begin(), toCollapse.end(), [](const CandPtr& a, const CandPtr& b) {  // LINQ OrderBy chain (stable) :42-47
            if (a->Length() != b->Length()) return a->Length() > b->Length();
            bool ab = a->OpenOnLeft && a->OpenOnRight, bb = b->OpenOnLeft && b->OpenOnRight;
            if (ab != bb) return ab;
            bool ao = a->OpenOnLeft || a->OpenOnRight, bo = b->OpenOnLeft || b->OpenOnRight;
            if (ao != bo) return ao;
            if (a->ReferenceAllele != b->ReferenceAllele) return a->ReferenceAllele < b->ReferenceAllele;
            if (a->AlternateAllele != b->AlternateAllele) return a->AlternateAllele < b->AlternateAllele;
            if (a->Support() != b->Support()) return a->Support() < b->Support();
            if (a->OpenOnRight != b->OpenOnRight) return !a->OpenOnRight;
            if (a->OpenOnLeft != b->OpenOnLeft) return !a->OpenOnLeft;
            return false;
        });
        for (size_t i = 0; i < toCollapse.size(); i++) {
            auto v = toCollapse[i];
            auto match = GetMatches(v, targetVariants, source);
            if (match) {
                TotalNumCollapsed++;
                match->AddSupport(*v);  // :115-123
                match->OpenOnLeft = match->OpenOnLeft && v->OpenOnLeft;
                match->OpenOnRight = match->OpenOnRight && v->OpenOnRight;
                for (int k = 0; k < 8; k++) match->ReadCollapsedCountsMut[k] += v->ReadCollapsedCountsMut[k];
                targetVariants.erase(std::remove(targetVariants.begin(), targetVariants.end(), v), targetVariants.end());  // reference equality
                candidates.erase(std::remove(candidates.begin(), candidates.end(), v), candidates.end());
            }
        }
        if (maxClearedPosition.has_value()) {
            std::vector<CandPtr> notCleared;
            for (auto& c : candidates) if (c->ReferencePosition > *maxClearedPosition && c->Type != Reference) notCleared.push_back(c);
            for (auto& nc : notCleared) {  // List.Remove -> first element that Equals() (value equality)
                for (auto it = candidates.begin(); it != candidates.end(); ++it) if ((*it)->Equals(*nc)) { candidates.erase(it); break; }
            }
            source.AddCandidates(notCleared);
        }
    }
};

// ------------------------------------------------------------------ MnvReallocator.cs
namespace mnv {
inline CalledPtr CreateVariant(const std::string& chr, int coordinate, int alleleSupport, const std::string& alternate, const std::string& reference,
                               const std::array<int, 3>* supportByDirection = nullptr) {  // :156-173
    CalledPtr a = (alternate == reference) ? std::make_shared<CalledAllele>() : std::make_shared<CalledAllele>(alternate.size() > 1 ? Mnv : Snv);
    a->Chromosome = chr; a->ReferencePosition = coordinate; a->AlleleSupport = alleleSupport; a->AlternateAllele = alternate; a->ReferenceAllele = reference;
    if (supportByDirection) a->SupportByDirection = *supportByDirection;
    return a;
}
inline std::vector<CalledPtr> BreakOffEdgeReferences(const CalledPtr& allele) {  // :215-246
    if (allele->Type != Mnv) return {allele};
    int leftAdjust = 0, rightAdjust = 0, n = (int)allele->ReferenceAllele.size();
    for (int i = 0; i < n; i++) { if (allele->ReferenceAllele[i] != allele->AlternateAllele[i]) break; leftAdjust++; }
    for (int i = 0; i < n; i++) { int k = n - 1 - i; if (allele->ReferenceAllele[k] != allele->AlternateAllele[k]) break; rightAdjust++; }
    return {CreateVariant(allele->Chromosome, allele->ReferencePosition + leftAdjust, allele->AlleleSupport,
                          allele->AlternateAllele.substr(leftAdjust, allele->AlternateAllele.size() - (leftAdjust + rightAdjust)),
                          allele->ReferenceAllele.substr(leftAdjust, allele->ReferenceAllele.size() - (leftAdjust + rightAdjust)), &allele->SupportByDirection)};
}
inline std::vector<CalledPtr> CreateAllelesFromRemainder(const CalledAllele& overlap, const CalledAllele& r) {  // :175-213
    std::vector<CalledPtr> remainders;
    int overlapIndexInFailedMnv = overlap.ReferencePosition - r.ReferencePosition;
    int overlapAlleleLength = (int)overlap.AlternateAllele.size();
    int rightSideOverlap = overlapIndexInFailedMnv + overlapAlleleLength;
    int altLen = (int)r.AlternateAllele.size();
    if (altLen - rightSideOverlap > 0 && rightSideOverlap <= r.ReferencePosition + altLen) {
        auto rr = CreateVariant(r.Chromosome, r.ReferencePosition + rightSideOverlap, r.AlleleSupport, r.AlternateAllele.substr(rightSideOverlap, altLen - rightSideOverlap),
                                r.ReferenceAllele.substr(rightSideOverlap, altLen - rightSideOverlap), &r.SupportByDirection);
        if (rr->Type != Reference) remainders.push_back(rr);
    }
    if (overlapIndexInFailedMnv > 0) {
        auto lr = CreateVariant(r.Chromosome, r.ReferencePosition, r.AlleleSupport, r.AlternateAllele.substr(0, overlapIndexInFailedMnv),
                                r.ReferenceAllele.substr(0, overlapIndexInFailedMnv), &r.SupportByDirection);
        if (lr->Type != Reference) remainders.push_back(lr);
    }
    std::vector<CalledPtr> out;
    for (auto& rem : remainders) { auto b = BreakOffEdgeReferences(rem); out.insert(out.end(), b.begin(), b.end()); }
    return out;
}
inline bool IsPotentialOverlap(const CalledAllele& c, const CalledAllele& f) {  // :255-265
    int fEnd = f.ReferencePosition + (int)f.AlternateAllele.size();
    return c.ReferencePosition >= f.ReferencePosition && c.Chromosome == f.Chromosome && c.ReferencePosition <= fEnd &&
           c.AlternateAllele.size() <= f.AlternateAllele.size() && c.ReferencePosition + (int)c.AlternateAllele.size() <= fEnd &&
           (c.Type == Mnv || c.Type == Snv || c.Type == Reference);
}
inline bool OverlapMatches(const CalledAllele& overlap, const CalledAllele& r) {  // :248-253
    int idx = overlap.ReferencePosition - r.ReferencePosition;
    return overlap.AlternateAllele == r.AlternateAllele.substr(idx, overlap.AlternateAllele.size());
}
inline void RemoveRef(std::vector<CalledPtr>& v, const CalledPtr& x) { auto it = std::find(v.begin(), v.end(), x); if (it != v.end()) v.erase(it); }
inline void ProcessOverlap(std::optional<int> blockMaxPos, const CalledPtr& overlap, const CalledPtr& alleleToReassign, std::vector<CalledPtr>& remainderAlleles,
                           std::vector<CalledPtr>& outsideThisBlock) {  // :100-137
    overlap->AlleleSupport += alleleToReassign->AlleleSupport;
    for (int i = 0; i < 3; i++) overlap->SupportByDirection[i] += alleleToReassign->SupportByDirection[i];
    RemoveRef(remainderAlleles, alleleToReassign);
    auto remainders = CreateAllelesFromRemainder(*overlap, *alleleToReassign);
    if (blockMaxPos.has_value()) {
        if (overlap->ReferencePosition > *blockMaxPos) { RemoveRef(remainderAlleles, overlap); outsideThisBlock.push_back(overlap); }
        for (auto& rem : remainders) {
            if (rem->ReferencePosition <= *blockMaxPos) remainderAlleles.push_back(rem);
            else outsideThisBlock.push_back(rem);
        }
    } else remainderAlleles.insert(remainderAlleles.end(), remainders.begin(), remainders.end());
}
inline bool OrderKey(const CalledPtr& a, const CalledPtr& b) {  // OrderByDescending(alt len).ThenByDescending(support).ThenBy(alt).ThenBy(ref)
    if (a->AlternateAllele.size() != b->AlternateAllele.size()) return a->AlternateAllele.size() > b->AlternateAllele.size();
    if (a->AlleleSupport != b->AlleleSupport) return a->AlleleSupport > b->AlleleSupport;
    if (a->AlternateAllele != b->AlternateAllele) return a->AlternateAllele < b->AlternateAllele;
    return a->ReferenceAllele < b->ReferenceAllele;
}
inline std::vector<CalledPtr> ReallocateFailedMnvs(std::vector<CalledPtr>& failedMnvs, std::vector<CalledPtr>& callableAlleles, std::optional<int> blockMaxPos) {  // :12-98
    std::vector<CalledPtr> outsideThisBlock;
    std::vector<CalledPtr> ordered = failedMnvs;
    std::stable_sort(ordered.begin(), ordered.end(), [](const CalledPtr& a, const CalledPtr& b) {
        if (a->ReferencePosition != b->ReferencePosition) return a->ReferencePosition < b->ReferencePosition;
        return OrderKey(a, b);
    });
    for (auto& failedMnv : ordered) {
        std::vector<CalledPtr> remainderAlleles{failedMnv};
        while (!remainderAlleles.empty()) {
            CalledPtr alleleToReassign = remainderAlleles.front();
            std::vector<CalledPtr> orderedOverlaps;
            for (auto& a : callableAlleles) if (IsPotentialOverlap(*a, *alleleToReassign)) orderedOverlaps.push_back(a);
            std::stable_sort(orderedOverlaps.begin(), orderedOverlaps.end(), OrderKey);
            bool reallocated = false;
            std::vector<CalledPtr> matchingOverlaps;
            for (auto& o : orderedOverlaps) if (OverlapMatches(*o, *alleleToReassign)) matchingOverlaps.push_back(o);
            if (blockMaxPos.has_value()) {
                int distanceIntoNextBlock = (int)(alleleToReassign->ReferencePosition + ((int)alleleToReassign->AlternateAllele.size() - 1) - *blockMaxPos);
                bool anyLong = false;
                for (auto& o : matchingOverlaps) if (o->AlternateAllele.size() > 1) anyLong = true;
                if (distanceIntoNextBlock > 0 && !anyLong) {
                    if (alleleToReassign->ReferencePosition <= *blockMaxPos) {
                        int coordinate = *blockMaxPos + 1;
                        int originalAlleleLength = (int)alleleToReassign->ReferenceAllele.size();
                        auto nextBlockVariant = CreateVariant(alleleToReassign->Chromosome, coordinate, 0,
                                                              alleleToReassign->AlternateAllele.substr(originalAlleleLength - distanceIntoNextBlock, distanceIntoNextBlock),
                                                              alleleToReassign->ReferenceAllele.substr(originalAlleleLength - distanceIntoNextBlock, distanceIntoNextBlock));
                        auto nextBlockVariants = BreakOffEdgeReferences(nextBlockVariant);
                        ProcessOverlap(blockMaxPos, nextBlockVariants.front(), alleleToReassign, remainderAlleles, outsideThisBlock);
                    } else {
                        RemoveRef(remainderAlleles, alleleToReassign);
                        outsideThisBlock.push_back(alleleToReassign);
                    }
                    reallocated = true;
                }
            }
            if (!reallocated && !matchingOverlaps.empty()) {
                ProcessOverlap(blockMaxPos, matchingOverlaps.front(), alleleToReassign, remainderAlleles, outsideThisBlock);
                reallocated = true;
            }
            if (!reallocated) {
                // BreakDownToSingleNucCalls :139-154
                for (size_t i = 0; i < alleleToReassign->AlternateAllele.size(); i++) {
                    auto sn = CreateVariant(alleleToReassign->Chromosome, alleleToReassign->ReferencePosition + (int)i, alleleToReassign->AlleleSupport,
                                            alleleToReassign->AlternateAllele.substr(i, 1), alleleToReassign->ReferenceAllele.substr(i, 1), &alleleToReassign->SupportByDirection);
                    if (sn->Type == Reference) continue;
                    if (blockMaxPos.has_value() && !(sn->ReferencePosition <= *blockMaxPos)) outsideThisBlock.push_back(sn);
                    else callableAlleles.push_back(sn);
                }
                RemoveRef(remainderAlleles, alleleToReassign);
            }
        }
    }
    return outsideThisBlock;
}
}  // namespace mnv

// ------------------------------------------------------------------ AlleleCaller.cs
using ForcedAllele = std::tuple<std::string, int, std::string, std::string>;

struct AlleleCaller {
    Config cfg;
    std::string chrName;
    const std::string* chrSeq;
    ChrIntervalSet* intervalSet;
    CoverageCalculator coverage;
    std::unique_ptr<VariantCollapser> collapser;
    std::set<ForcedAllele> ForcedGtAlleles;
    int TotalNumCalled = 0;
    // GenotypeCreator.CreateGenotypeCalculator (GenotypeCreator.cs:10-37) for this chromosome; VariantCallerConfig.MinFrequency = MinVarFrequency,
    // .VariantFreqFilter = MinVarFrequencyFilter (Factory.cs:160,166); the locus processor follows the SAMPLE ploidy (Factory.cs:145-147)
    int chrPloidy = PM_Somatic;
    float minFrequency = 0, variantFreqFilter = 0;

    AlleleCaller(const Config& c, const std::string& chr, const std::string* seq, ChrIntervalSet* iv)
        : cfg(c), chrName(chr), chrSeq(seq), intervalSet(iv),
          coverage(c.TrackedAnchorSize > 0, c.SourceIsCollapsed && c.SourceIsStitched) {  // Factory.cs:193-199
        coverage.trackAmpliconCoverage = c.TrackAmpliconCounts();
        if (cfg.Collapse) collapser = std::make_unique<VariantCollapser>(&coverage, cfg.CollapseFreqThreshold, cfg.CollapseFreqRatioThreshold, cfg.ExcludeMNVsFromCollapsing);
        chrPloidy = GetPloidyForThisChr(cfg.ploidy, cfg.IsMale, chrName);
        if (chrPloidy == PM_DiploidByAdaptiveGT) throw std::runtime_error("DiploidByAdaptiveGT genotyper is out of scope (SURVEY 8f rank 4)");
        minFrequency = chrPloidy == PM_Somatic ? cfg.MinimumFrequency : cfg.DiploidMinorVF;
        variantFreqFilter = cfg.MinimumFrequencyFilter > minFrequency ? cfg.MinimumFrequencyFilter : minFrequency;
    }
    int TotalNumCollapsed() const { return collapser ? collapser->TotalNumCollapsed : 0; }
    bool IsForcedAllele(const CalledAllele& a) const { return ForcedGtAlleles.count(ForcedAllele{a.Chromosome, a.ReferencePosition, a.ReferenceAllele, a.AlternateAllele}) > 0; }

    void ProcessVariant(IAlleleSource& source, CalledAllele& v) {  // :208-234
        coverage.Compute(v, source);
        if (v.AlleleSupport > 0) {
            int NL = cfg.NoiseLevelUsedForQScoring();
            if (cfg.noiseModel == NM_Window) VariantQualityCompute(v, cfg.MaximumVariantQScore, (int)PtoQ(v.SumOfBaseQuality / v.TotalCoverage));
            else VariantQualityCompute(v, cfg.MaximumVariantQScore, NL);
            // _config.MinFrequency = genotypeCalculator.MinVarFrequency (Factory.cs:140,160)
            v.StrandBiasResults = CalculateStrandBiasResults(v.EstimatedCoverageByDirection.data(), v.SupportByDirection.data(), NL, minFrequency,
                                                             cfg.StrandBiasAcceptanceCriteria, cfg.strandBiasModel);
            // AmpliconBiasCalculator.Compute (AmpliconBiasCalculator.cs:20-31): SNVs only, only with a threshold
            if (cfg.AmpliconBiasFilterThreshold >= 0 && v.Type == Snv) {
                int sn[MaxNumOverlappingAmplicons], sc[MaxNumOverlappingAmplicons], cn[MaxNumOverlappingAmplicons], cc[MaxNumOverlappingAmplicons];
                int nc = 0;
                for (int i = 0; i < MaxNumOverlappingAmplicons; i++) {
                    sn[i] = v.SupportByAmplicon.names[(size_t)i]; sc[i] = v.SupportByAmplicon.counts[(size_t)i];
                    if (!v.CoverageByAmplicon.isNull && v.CoverageByAmplicon.names[(size_t)i] >= 0) { cn[nc] = v.CoverageByAmplicon.names[(size_t)i]; cc[nc] = v.CoverageByAmplicon.counts[(size_t)i]; nc++; }
                }
                const AmpliconBiasResults r = CalculateAmpliconBias(sn, sc, v.SupportByAmplicon.isNull ? -1 : MaxNumOverlappingAmplicons, cn, cc, nc,
                                                                    cfg.AmpliconBiasFilterThreshold, cfg.MaximumVariantQScore);
                v.HasAmpliconBiasResults = !r.isNull;
                v.AmpliconBiasDetected = !r.isNull && r.biasDetected;
            }
        }
        AlleleProcessorProcess(v, cfg, *chrSeq, source.ExpectStitchedReads(), variantFreqFilter);
    }
    bool IsCallable(const CalledAllele& a) {  // :236-258
        if (a.Type == Reference) { TotalNumCalled++; return true; }
        if (a.TotalCoverage < cfg.MinimumCoverage && !cfg.OutputGvcfFile) return false;
        if (a.TotalCoverage != 0 && a.Frequency() < minFrequency) return false;
        if (a.VariantQscore < cfg.MinimumVariantQScore) return false;
        TotalNumCalled++;
        return true;
    }
    bool ShouldReport(const CalledAllele& a) const { return intervalSet == nullptr ? true : intervalSet->ContainsPosition(a.ReferencePosition); }  // :260-263

    static std::map<int, int> GetRefSupportFromGappedMnvs(const std::vector<CalledPtr>& callable) {  // :186-206
        std::map<int, int> taken;
        for (auto& a : callable) {
            if (a->Type != Mnv) continue;
            for (size_t i = 0; i < a->ReferenceAllele.size(); i++) {
                if (a->ReferenceAllele[i] != a->AlternateAllele[i]) continue;
                taken[a->ReferencePosition + (int)i] += a->AlleleSupport;
            }
        }
        return taken;
    }
    void ComputeGenotypeAndFilterAllele(std::vector<CalledPtr>& at) {  // :143-177
        bool anyVar = false;
        for (auto& v : at) if (v->Type != Reference && !v->IsForcedToReport) anyVar = true;
        if (anyVar) at.erase(std::remove_if(at.begin(), at.end(), [](const CalledPtr& v) { return v->Type == Reference; }), at.end());
        std::vector<CalledPtr> toGenotype;
        for (auto& a : at) if (!a->IsForcedToReport) toGenotype.push_back(a);
        std::vector<CalledPtr> toPrune;
        if (chrPloidy == PM_Somatic) {
            // SomaticGenotyper.SetGenotypes :51-63 ; ctor args Factory.cs:131-141: minVariantFrequencyFilter = MinimumFrequencyFilter,
            // MinDepthToGenotype = MinimumCoverage, targetLOD = TargetLODFrequency; nothing is pruned
            for (auto& a : toGenotype) {
                a->genotype = CalculateSomaticGenotype(*a, cfg.MinimumFrequencyFilter, cfg.MinimumCoverage);
                a->GenotypeQscore = SomaticGenotypeQuality(*a, cfg.TargetLODFrequency, cfg.MinimumGenotypeQScore, cfg.MaximumGenotypeQScore);
            }
        } else {
            DiploidThresholdingParameters snv;
            snv.MinorVF = cfg.DiploidMinorVF; snv.MajorVF = cfg.DiploidMajorVF; snv.SumVFforMultiAllelicSite = cfg.DiploidSumVFforMultiAllelicSite;
            toPrune = chrPloidy == PM_Haploid
                          ? HaploidSetGenotypes(toGenotype, cfg.MinimumCoverage, snv.MinorVF, snv.MajorVF, cfg.MinimumGenotypeQScore, cfg.MaximumGenotypeQScore)
                          : DiploidSetGenotypes(toGenotype, cfg.MinimumCoverage, snv, snv, cfg.MinimumGenotypeQScore, cfg.MaximumGenotypeQScore);
        }
        for (auto& p : toPrune) {   // :153-162: pruned unless it is a forced allele; List.Remove = first element that Equals (reference equality for CalledAllele)
            if (IsForcedAllele(*p)) continue;
            auto it = std::find(at.begin(), at.end(), p);
            if (it != at.end()) at.erase(it);
        }
        for (auto& a : at)
            if (cfg.LowGenotypeQualityFilter >= 0 && (float)a->GenotypeQscore < (float)cfg.LowGenotypeQualityFilter) a->AddFilter(F_LowGenotypeQuality);
        std::stable_sort(at.begin(), at.end(), [](const CalledPtr& a, const CalledPtr& b) {
            if (a->ReferenceAllele != b->ReferenceAllele) return a->ReferenceAllele < b->ReferenceAllele;
            return a->AlternateAllele < b->AlternateAllele;
        });
    }
    // Call / CallForPositions :50-141
    std::map<int, std::vector<CalledPtr>> Call(CandidateBatch& batch, IAlleleSource& source) {
        std::vector<CandPtr> candidates = batch.candidates;
        std::optional<int> maxPosition = batch.MaxClearedPosition;
        std::vector<CalledPtr> failedMnvs, callableAlleles;
        if (collapser) collapser->Collapse(candidates, source, maxPosition);
        for (auto& c : candidates) {
            auto variant = std::make_shared<CalledAllele>(MapToCalled(*c));
            if (variant->Type == Mnv) {
                ProcessVariant(source, *variant);
                if (IsCallable(*variant)) callableAlleles.push_back(variant);
                else failedMnvs.push_back(variant);
            } else callableAlleles.push_back(variant);
        }
        auto leftovers = mnv::ReallocateFailedMnvs(failedMnvs, callableAlleles, maxPosition);
        std::vector<CandPtr> leftoverCands;
        for (auto& l : leftovers) leftoverCands.push_back(std::make_shared<CandidateAllele>(MapToCandidate(*l)));
        source.AddCandidates(leftoverCands);
        source.AddGappedMnvRefCount(GetRefSupportFromGappedMnvs(callableAlleles));
        std::map<int, std::vector<CalledPtr>> byPos;
        for (auto& f : failedMnvs) if (IsForcedAllele(*f)) callableAlleles.push_back(f);
        for (auto& a : callableAlleles) {
            ProcessVariant(source, *a);
            if (IsForcedAllele(*a) && !(IsCallable(*a) && ShouldReport(*a))) { a->IsForcedToReport = true; a->AddFilter(F_ForcedReport); }
            if ((IsCallable(*a) && ShouldReport(*a)) || IsForcedAllele(*a)) byPos[a->ReferencePosition].push_back(a);
        }
        for (auto& kv : byPos) {
            ComputeGenotypeAndFilterAllele(kv.second);
            if (cfg.ploidy == PM_DiploidByThresholding) DiploidLocusProcess(kv.second);   // Factory.cs:145-147; SomaticLocusProcessor.Process is a no-op
        }
        return byPos;
    }
};

// ------------------------------------------------------------------ SmallVariantCaller.cs :79-189 + AlignmentsSource.cs :57-97
struct SmallVariantCaller {
    Config cfg;
    std::string chrName, chrSeq;
    std::unique_ptr<ChrIntervalSet> intervals;
    std::unique_ptr<RegionStateManager> state;
    std::unique_ptr<CandidateVariantFinder> finder;
    std::unique_ptr<AlleleCaller> caller;
    std::map<int, std::vector<std::pair<std::string, std::string>>> unProcessedForcedAllelesByPos;  // SortedList
    std::vector<int> forcedPositions;
    std::vector<CalledPtr> output;                 // in _vcfWriter.Write order
    std::vector<std::pair<int, int>> writeBatches;  // [begin,end) into output per Write() call, for the VCF writer restatement
    long long totalReadsReturned = 0, totalSkipped = 0;
    int lastReadPosition = 0;

    SmallVariantCaller(Config c, const std::string& name, const std::string& seq, const std::vector<Region>* ivs) : cfg(c), chrName(name), chrSeq(seq) {
        cfg.Validate();
        if (ivs) { intervals = std::make_unique<ChrIntervalSet>(); intervals->Intervals = *ivs; intervals->SortAndCollapse(); }  // Factory.cs:229-245
        // Factory.CreateStateManager :209-227
        state = std::make_unique<RegionStateManager>(cfg.OutputGvcfFile, cfg.MinimumBaseCallQuality, cfg.SourceIsStitched, intervals.get(), 1000, cfg.Collapse,
                                                     cfg.TrackedAnchorSize, cfg.SourceIsStitched && cfg.SourceIsCollapsed, cfg.TrackAmpliconCounts());
        finder = std::make_unique<CandidateVariantFinder>(cfg.MinimumBaseCallQuality, cfg.MaxSizeMNV, cfg.MaxGapBetweenMNV, cfg.CallMNVs, 5, cfg.TrackAmpliconCounts());  // Factory.cs:123-126
        caller = std::make_unique<AlleleCaller>(cfg, chrName, &chrSeq, intervals.get());
    }
    void AddForcedAllele(int pos, const std::string& ref, const std::string& alt) {  // SmallVariantCaller.cs:48-77 + Factory.SelectForcedAllele
        if (intervals && !intervals->ContainsPosition(pos)) return;
        caller->ForcedGtAlleles.insert(ForcedAllele{chrName, pos, ref, alt});
        unProcessedForcedAllelesByPos[pos].push_back({ref, alt});
        forcedPositions.push_back(pos);
    }
    bool ShouldSkipRead(const Read& r) const {  // AlignmentsSource.cs:84-92
        return (!r.IsMapped || !r.IsPrimaryAlignment || (cfg.OnlyUseProperPairs && !r.IsProperPair) || (cfg.RemoveDuplicates && r.IsPcrDuplicate) ||
                (int)r.MapQuality < cfg.MinimumMapQuality || !r.HasCigar());
    }
    void AddForcedAlleleAsCandidate(std::optional<int> upTo) {  // :118-155
        while (!unProcessedForcedAllelesByPos.empty()) {
            auto it = unProcessedForcedAllelesByPos.begin();
            if (upTo.has_value() && it->first > *upTo) break;
            std::vector<CandPtr> cs;
            for (auto& ra : it->second) {
                AlleleCategory cat = (ra.first.size() == 1 && ra.second.size() == 1) ? Snv : ra.first.size() == ra.second.size() ? Mnv : ra.first.size() > ra.second.size() ? Deletion : Insertion;
                cs.push_back(std::make_shared<CandidateAllele>(chrName, it->first, ra.first, ra.second, cat));
            }
            state->AddCandidates(cs);
            unProcessedForcedAllelesByPos.erase(it);
        }
    }
    void Call(std::optional<int> upTo) {  // :157-189
        auto batch = state->GetCandidatesToProcess(upTo, chrName, chrSeq, &forcedPositions);
        if (!batch) return;
        if (batch->HasCandidates()) {
            auto byPos = caller->Call(*batch, *state);
            int b = (int)output.size();
            for (auto& kv : byPos) for (auto& a : kv.second) output.push_back(a);
            writeBatches.push_back({b, (int)output.size()});
        }
        if (intervals && batch->MaxClearedPosition.has_value()) intervals->SetCleared(*batch->MaxClearedPosition);
        state->DoneProcessing(*batch);
    }
    // Execute() :79-116 split so reads can be streamed through the C API
    void ProcessRead(const Read& r) {
        lastReadPosition = r.Position();
        if (ShouldSkipRead(r)) { totalSkipped++; return; }
        totalReadsReturned++;
        auto cands = finder->FindCandidates(r, chrSeq, chrName);
        state->AddCandidates(cands);
        state->AddAlleleCounts(r);
        AddForcedAlleleAsCandidate(lastReadPosition - 1);
        Call(lastReadPosition - 1);
    }
    void Finish() {
        AddForcedAlleleAsCandidate(std::nullopt);
        Call(std::nullopt);
    }
};

}  // namespace po
