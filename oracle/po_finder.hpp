// ORACLE — TEST INFRASTRUCTURE ONLY (see po_math.hpp header).
// Candidate discovery: src/lib/Pisces.Domain/Logic/CandidateVariantFinder.cs
#pragma once
#include "po_state.hpp"

namespace po {

struct CandidateVariantFinder {
    int minimumBaseCallQuality, maxLengthMnv, maxLengthInterveningRef;
    bool callMnvs;
    int wellAnchoredAnchorSize;
    bool trackAmpliconCounts;   // :19-28
    CandidateVariantFinder(int qualityCutoff, int maxMnv, int maxGap, bool mnvs, int wellAnchored = 5, bool trackAmplicons = false)
        : minimumBaseCallQuality(qualityCutoff), maxLengthMnv(maxMnv), maxLengthInterveningRef(maxGap), callMnvs(mnvs), wellAnchoredAnchorSize(wellAnchored),
          trackAmpliconCounts(trackAmplicons) {}

    // GetSupportDirection :396-445 (+ GetDeletionDirectionForStitchedRead :462-488)
    static DirectionType GetSupportDirection(const CandidateAllele& c, const Read& r, int startIndexInRead) {
        auto dirMap = r.SequencedBaseDirectionMap();
        if (c.Type == Snv || c.Type == Reference) return dirMap.at(startIndexInRead);
        int leftAnchorIndex = startIndexInRead - 1;
        int rightAnchorIndex = c.Type == Deletion ? startIndexInRead : startIndexInRead + c.Length();
        int lastIndex = (int)r.Sequence.size() - 1;
        if (rightAnchorIndex == 0) return dirMap.at(rightAnchorIndex);
        if (leftAnchorIndex == lastIndex) return dirMap.at(lastIndex);
        if (leftAnchorIndex == rightAnchorIndex - 1) {  // deletions
            if (r.HasCigarDirections()) {
                auto expanded = r.ExpandedBaseDirectionMap();
                // SequencedIndexesToExpandedIndexes({left, right})  Read.cs:432-476
                int want[2] = {leftAnchorIndex, rightAnchorIndex}, got[2] = {-1, -1};
                int which = 0, seqIdx = 0, ext = 0;
                for (auto& op : r.CigarData) {
                    for (uint32_t k = 0; k < op.Length && which < 2; k++, ext++) {
                        if (op.IsReadSpan()) {
                            if (seqIdx == want[which]) { got[which] = ext; which++; }
                            seqIdx++;
                        }
                    }
                }
                int first = got[0] + 1, last = got[1] - 1;
                if (got[0] >= 0 && got[1] >= 0 && first >= 0 && first < (int)expanded.size() && last >= 0 && last < (int)expanded.size()) {
                    DirectionType s = expanded[first], e = expanded[last];
                    return (s == Stitched) ? e : s;
                }
                throw std::runtime_error("Unable to find direction info for deletion.");
            }
            DirectionType s = dirMap.at(leftAnchorIndex), e = dirMap.at(rightAnchorIndex);
            return s == Stitched ? e : s;
        }
        DirectionType direction = Forward;
        for (int i = leftAnchorIndex + 1; i < rightAnchorIndex; i++) {
            direction = dirMap.at(i);
            if (direction == Stitched) return Stitched;
        }
        if (dirMap.at(leftAnchorIndex + 1) != dirMap.at(rightAnchorIndex - 1))
            throw std::runtime_error("Alignment error: Found change in direction without encountering stitched direction");
        return direction;
    }
    // Create :334-387
    static CandPtr Create(AlleleCategory type, const std::string& chr, int coordinate, const std::string& ref, const std::string& alt, const Read& r,
                          int startIndexInRead, int wellAnchoredAnchorSize) {
        auto c = std::make_shared<CandidateAllele>(chr, coordinate, ref, alt, type);
        DirectionType d = GetSupportDirection(*c, r, startIndexInRead);
        c->SupportByDirection[d]++;
        int anchor = std::min(coordinate - r.Position(), r.EndPosition() - coordinate);
        if (anchor > std::min(wellAnchoredAnchorSize - 1, (int)alt.size() - 1)) c->WellAnchoredSupportByDirection[d]++;
        if (r.IsCollapsedRead()) {
            auto t = r.GetReadCollapsedType(d);
            if (t.has_value()) {
                switch (*t) {
                    case DuplexNonStitched: c->ReadCollapsedCountsMut[DuplexNonStitched]++; break;
                    case DuplexStitched: c->ReadCollapsedCountsMut[DuplexStitched]++; break;
                    case SimplexStitched: c->ReadCollapsedCountsMut[SimplexStitched]++; break;
                    case SimplexReverseStitched: c->ReadCollapsedCountsMut[SimplexStitched]++; c->ReadCollapsedCountsMut[SimplexReverseStitched]++; break;
                    case SimplexForwardStitched: c->ReadCollapsedCountsMut[SimplexStitched]++; c->ReadCollapsedCountsMut[SimplexForwardStitched]++; break;
                    case SimplexNonStitched: c->ReadCollapsedCountsMut[SimplexNonStitched]++; break;
                    case SimplexReverseNonStitched: c->ReadCollapsedCountsMut[SimplexNonStitched]++; c->ReadCollapsedCountsMut[SimplexReverseNonStitched]++; break;
                    case SimplexForwardNonStitched: c->ReadCollapsedCountsMut[SimplexNonStitched]++; c->ReadCollapsedCountsMut[SimplexForwardNonStitched]++; break;
                }
            }
        }
        return c;
    }
    bool ShouldBuildUpMNV(int mnvLengthSoFar, int interveningRefLengthSoFar, bool refCallNext) const {  // :170-181
        if (!callMnvs) return false;
        if (refCallNext && mnvLengthSoFar == 0) return false;
        if ((mnvLengthSoFar + 1) > maxLengthMnv) return false;
        if ((interveningRefLengthSoFar + (refCallNext ? 1 : 0)) > maxLengthInterveningRef) return false;
        return true;
    }
    void FlushVariant(const Read& r, const std::string& refChr, int variantStartIndexInRead, int variantStartIndexInReference, const std::string& chrName,
                      int variantLengthSoFar, int interveningRefLengthSoFar, std::vector<CandPtr>& out, bool openLeft, bool openRight) const {  // :183-203
        if (interveningRefLengthSoFar >= 1) { variantLengthSoFar -= interveningRefLengthSoFar; openRight = false; }
        if (variantLengthSoFar >= 1) {
            std::string referenceBases = refChr.substr(variantStartIndexInReference, variantLengthSoFar);
            std::string readBases = r.Sequence.substr(variantStartIndexInRead, variantLengthSoFar);
            auto c = Create(referenceBases.size() > 1 ? Mnv : Snv, chrName, variantStartIndexInReference + 1, referenceBases, readBases, r,
                            variantStartIndexInRead, wellAnchoredAnchorSize);
            if (trackAmpliconCounts && r.AmpliconName >= 0) {   // CreateMnvSnv -> SetAmpliconName :205-232 (SNVs and MNVs only; the indel calls are commented out :256,288)
                c->SupportByAmplicon = AmpliconCounts::Empty();
                c->SupportByAmplicon.names[0] = r.AmpliconName;
                c->SupportByAmplicon.counts[0] = 1;
            }
            c->OpenOnLeft = openLeft;
            c->OpenOnRight = openRight;
            out.push_back(c);
        }
    }
    void ExtractSnvsFromOperation(const Read& r, const std::string& refChr, int opStartIndexInRead, uint32_t operationLength, int opStartIndexInReference,
                                  const std::string& chrName, std::vector<CandPtr>& out) const {  // :90-168
        int variantLengthSoFar = 0, interveningRefLengthSoFar = 0;
        bool openLeft = false;
        for (int i = 0; i < (int)operationLength; i++) {
            bool qualityGoodEnough = r.Qualities[opStartIndexInRead + i] >= minimumBaseCallQuality;
            char readBase = r.Sequence[opStartIndexInRead + i];
            if (opStartIndexInReference + i >= (int)refChr.size()) break;
            char refBase = refChr[opStartIndexInReference + i];
            bool atEndOfOperation = i == ((int)operationLength - 1);
            bool startingMnvAtEndOfOperation = (atEndOfOperation && variantLengthSoFar == 0);
            if ((GetAlleleType(readBase) == AT_N) || (GetAlleleType(refBase) == AT_N) || !qualityGoodEnough) {
                FlushVariant(r, refChr, opStartIndexInRead + i - variantLengthSoFar, opStartIndexInReference + i - variantLengthSoFar, chrName,
                             variantLengthSoFar, interveningRefLengthSoFar, out, openLeft, true);
                variantLengthSoFar = 0; interveningRefLengthSoFar = 0; openLeft = true;
            } else if (refBase == readBase) {
                if (ShouldBuildUpMNV(variantLengthSoFar, interveningRefLengthSoFar, true) && !startingMnvAtEndOfOperation) {
                    variantLengthSoFar++; interveningRefLengthSoFar++;
                } else {
                    FlushVariant(r, refChr, opStartIndexInRead + i - variantLengthSoFar, opStartIndexInReference + i - variantLengthSoFar, chrName,
                                 variantLengthSoFar, interveningRefLengthSoFar, out, openLeft, false);
                    variantLengthSoFar = 0; interveningRefLengthSoFar = 0; openLeft = false;
                }
            } else {
                if (ShouldBuildUpMNV(variantLengthSoFar, interveningRefLengthSoFar, false) && !startingMnvAtEndOfOperation) {
                    variantLengthSoFar++; interveningRefLengthSoFar = 0;
                } else {
                    FlushVariant(r, refChr, opStartIndexInRead + i - variantLengthSoFar, opStartIndexInReference + i - variantLengthSoFar, chrName,
                                 variantLengthSoFar, interveningRefLengthSoFar, out, openLeft, false);
                    variantLengthSoFar = 1; interveningRefLengthSoFar = 0; openLeft = false;
                }
            }
        }
        FlushVariant(r, refChr, opStartIndexInRead + (int)operationLength - variantLengthSoFar, opStartIndexInReference + (int)operationLength - variantLengthSoFar,
                     chrName, variantLengthSoFar, interveningRefLengthSoFar, out, openLeft, false);
    }
    // FindCandidates / ProcessCigarOps :31-83
    std::vector<CandPtr> FindCandidates(const Read& r, const std::string& refChr, const std::string& chrName) const {
        std::vector<CandPtr> candidates;
        int startIndexInRead = 0, startIndexInReference = r.Position() - 1;
        for (auto& op : r.CigarData) {
            switch (op.Type) {
                case 'S': break;
                case 'M': ExtractSnvsFromOperation(r, refChr, startIndexInRead, op.Length, startIndexInReference, chrName, candidates); break;
                case 'I': {  // :234-260
                    if (startIndexInReference - 1 >= (int)refChr.size() || startIndexInReference == 0) break;
                    std::string referenceBases = refChr.substr(startIndexInReference - 1, 1);
                    std::string addedBases = r.Sequence.substr(startIndexInRead, op.Length);
                    if (!(r.Qualities[startIndexInRead] >= minimumBaseCallQuality)) break;
                    candidates.push_back(Create(Insertion, chrName, startIndexInReference, referenceBases, referenceBases + addedBases, r, startIndexInRead, wellAnchoredAnchorSize));
                    break;
                }
                case 'D': {  // :262-292
                    if (startIndexInReference + (long long)op.Length >= (long long)refChr.size()) break;
                    std::string referenceBases = refChr.substr(startIndexInReference - 1, op.Length + 1);
                    std::string readBases = refChr.substr(startIndexInReference - 1, 1);
                    if (!CheckDeletionQuality(r, startIndexInRead, minimumBaseCallQuality)) break;
                    candidates.push_back(Create(Deletion, chrName, startIndexInReference, referenceBases, readBases, r, startIndexInRead, wellAnchoredAnchorSize));
                    break;
                }
                default: break;
            }
            if (op.IsReadSpan()) startIndexInRead += (int)op.Length;
            if (op.IsReferenceSpan()) startIndexInReference += (int)op.Length;
        }
        Annotate(candidates, r);
        return candidates;
    }
    static void Annotate(std::vector<CandPtr>& candidates, const Read& read) {  // :496-553
        if (candidates.empty()) return;
        CigarOp firstOperation = read.CigarData.front(), lastOperation = read.CigarData.back();
        if (firstOperation.Type == 'S') firstOperation = read.CigarData.at(1);
        if (lastOperation.Type == 'S') lastOperation = read.CigarData.at(read.CigarData.size() - 2);
        int maxPosition = read.MaxPosition();
        if (maxPosition == -1) maxPosition = read.Position() - 1;
        for (auto& c : candidates) {
            switch (firstOperation.Type) {
                case 'M': if (c->ReferencePosition == read.Position() && (c->Type == Mnv || c->Type == Snv)) c->OpenOnLeft = true; break;
                case 'I': if (c->ReferencePosition == read.Position() - 1 && c->Type == Insertion) c->OpenOnLeft = true; break;
                case 'D': if (c->ReferencePosition == read.Position() - 1 && c->Type == Deletion) c->OpenOnLeft = true; break;
                default: break;
            }
        }
        for (auto& c : candidates) {
            switch (lastOperation.Type) {
                case 'M': if (c->ReferencePosition + (int)c->AlternateAllele.size() - 1 == maxPosition && (c->Type == Mnv || c->Type == Snv)) c->OpenOnRight = true; break;
                case 'I': if (c->ReferencePosition == maxPosition && c->Type == Insertion) c->OpenOnRight = true; break;
                case 'D': if (c->ReferencePosition == maxPosition && c->Type == Deletion) c->OpenOnRight = true; break;
                default: break;
            }
        }
    }
};

}  // namespace po
