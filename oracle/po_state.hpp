// ORACLE — TEST INFRASTRUCTURE ONLY (see po_math.hpp header).
// Pileup state: src/lib/Pisces.Processing/RegionState/{RegionState,RegionStateManager,AlleleCountHelper,CollapsedRegionState,
// CollapedRegionStateManager}.cs ; intervals: src/lib/Pisces.Domain/Models/IntervalSet.cs
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>
#include "po_read.hpp"

namespace po {

using CandPtr = std::shared_ptr<CandidateAllele>;
using CalledPtr = std::shared_ptr<CalledAllele>;

// IntervalSet.cs (ChrIntervalSet) — only what the hot path uses
struct ChrIntervalSet {
    std::vector<Region> Intervals;
    int lastIndexCleared = -1;
    static bool Overlaps(const Region& a, const Region& b) { return a.StartPosition <= b.EndPosition && b.StartPosition <= a.EndPosition; }
    void SortAndCollapse() {  // :38-75 (equivalent result: union of overlapping intervals, sorted by start)
        std::sort(Intervals.begin(), Intervals.end(), [](const Region& a, const Region& b) { return a.StartPosition < b.StartPosition; });
        std::vector<Region> out;
        for (auto& r : Intervals) {
            if (!out.empty() && Overlaps(out.back(), r)) out.back().EndPosition = std::max(out.back().EndPosition, r.EndPosition);
            else out.push_back(r);
        }
        Intervals = out;
    }
    std::vector<Region> GetClipped(const Region& clip) const {  // :77-106
        std::vector<Region> out;
        for (int i = lastIndexCleared + 1; i < (int)Intervals.size(); i++) {
            auto& iv = Intervals[i];
            if (iv.StartPosition > clip.EndPosition) break;
            if (!Overlaps(clip, iv)) continue;
            out.push_back(Region{std::max(clip.StartPosition, iv.StartPosition), std::min(clip.EndPosition, iv.EndPosition)});
        }
        return out;
    }
    bool ContainsPosition(int p) const {  // :167-180
        for (int i = lastIndexCleared + 1; i < (int)Intervals.size(); i++) {
            auto& iv = Intervals[i];
            if (iv.StartPosition > p) break;
            if (p >= iv.StartPosition && p <= iv.EndPosition) return true;
        }
        return false;
    }
    void SetCleared(int position) {  // :204-216
        for (int i = lastIndexCleared + 1; i < (int)Intervals.size(); i++) {
            auto& iv = Intervals[i];
            if (iv.EndPosition <= position) lastIndexCleared = i;
            if (iv.StartPosition > position) break;
        }
    }
    std::vector<Region> GetIntervals(int endPosition) const {  // :218-233
        std::vector<Region> out;
        for (int i = lastIndexCleared + 1; i < (int)Intervals.size(); i++) {
            if (Intervals[i].StartPosition > endPosition) break;
            out.push_back(Intervals[i]);
        }
        return out;
    }
};

// AlleleCountHelper.GetAnchorAdjustedAlleleCount / ...TotalQuality  AlleleCountHelper.cs:21-85,102-166
template <class T>
inline T AnchorAdjusted(int minAnchor, bool fromEnd, int wellAnchoredIndex, int numAnchorIndexes, const T* bins /*[numAnchorIndexes]*/,
                        std::optional<int> maxAnchor, bool symmetric) {
    int trueMinAnchor = std::min(wellAnchoredIndex, minAnchor);
    int initialMaxAnchor = wellAnchoredIndex;
    if (maxAnchor.has_value()) {
        if (*maxAnchor >= wellAnchoredIndex) initialMaxAnchor = wellAnchoredIndex - 1;
        if (*maxAnchor < wellAnchoredIndex) initialMaxAnchor = *maxAnchor;
    }
    T tot = 0;
    if (fromEnd) {
        for (int i = trueMinAnchor; i <= initialMaxAnchor; i++) tot += bins[numAnchorIndexes - i - 1];
        if (!maxAnchor.has_value())
            for (int i = symmetric ? trueMinAnchor : 0; i < initialMaxAnchor; i++) tot += bins[i];
    } else {
        for (int i = trueMinAnchor; i <= initialMaxAnchor; i++) tot += bins[i];
        if (!maxAnchor.has_value())
            for (int i = initialMaxAnchor + 1; i < (symmetric ? numAnchorIndexes - trueMinAnchor : numAnchorIndexes); i++) tot += bins[i];
    }
    return tot;
}

// IAlleleSource  src/lib/Pisces.Domain/Interfaces/IAlleleSource.cs:8-26
struct IAlleleSource {
    virtual ~IAlleleSource() {}
    virtual void AddCandidates(const std::vector<CandPtr>& c) = 0;
    virtual int GetAlleleCount(int position, AlleleType a, DirectionType d, int minAnchor = 0, std::optional<int> maxAnchor = std::nullopt,
                               bool fromEnd = false, bool symmetric = false) = 0;
    virtual double GetSumOfAlleleBaseQualities(int position, AlleleType a, DirectionType d, int minAnchor = 0,
                                               std::optional<int> maxAnchor = std::nullopt, bool fromEnd = false, bool symmetric = false) = 0;
    virtual int GetCollapsedReadCount(int position, ReadCollapsedType t) = 0;
    virtual void AddGappedMnvRefCount(const std::map<int, int>& lookup) = 0;
    virtual int GetGappedMnvRefCount(int position) = 0;
    virtual bool ExpectStitchedReads() const = 0;
    virtual AmpliconCounts GetCoverageByAmplicon(int position) = 0;   // IAlleleSource.cs:23
};

// RegionState.cs
struct RegionState {
    int StartPosition, EndPosition, K;  // K = _numAnchorTypes
    int NA;                             // NumAnchorIndexes = 2K+1
    std::vector<int> counts;            // [size][6][3][NA]
    std::vector<double> qsum;           // [size][6][3][NA]
    std::vector<int> gapped;            // [size]
    std::vector<std::vector<CandPtr>> cands;  // [size]
    std::vector<int> collapsed;         // [size][8]  (CollapsedRegionState)
    std::vector<AmpliconCounts> amplicons;   // [size] _ampliconNamesPerPos / _ampliconCountsPerPos (:269-307)
    int MaxAlleleEndpoint = 0;          // survives Reset() — Initialize() never clears it (RegionState.cs:31,54-78)

    RegionState(int s, int e, int k) : StartPosition(s), EndPosition(e), K(k), NA(2 * k + 1) { Initialize(); }
    int Size() const { return EndPosition - StartPosition + 1; }
    void Initialize() {  // :54-66
        size_t n = (size_t)Size();
        counts.assign(n * 6 * 3 * NA, 0);
        qsum.assign(n * 6 * 3 * NA, 0.0);
        gapped.assign(n, 0);
        cands.assign(n, {});
        collapsed.assign(n * 8, 0);
        amplicons.assign(n, AmpliconCounts());
    }
    void AddAmpliconCount(int p, int name) {   // :269-307
        if (name < 0 || !IsPositionInRegion(p)) return;
        AmpliconCounts& a = amplicons[(size_t)(p - StartPosition)];
        if (a.isNull) a = AmpliconCounts::Empty();
        a.Add(name, 1);
    }
    AmpliconCounts GetCountsByAmpliconForPosition(int p) const {   // :325-352: the filled slots, in slot order
        if (!IsPositionInRegion(p)) throw std::invalid_argument("Position is not in region");
        const AmpliconCounts& a = amplicons[(size_t)(p - StartPosition)];
        AmpliconCounts out = AmpliconCounts::Empty();
        int k = 0;
        for (int i = 0; i < MaxNumOverlappingAmplicons; i++)
            if (!a.isNull && a.names[(size_t)i] >= 0) { out.names[(size_t)k] = a.names[(size_t)i]; out.counts[(size_t)k] = a.counts[(size_t)i]; k++; }
        return out;
    }
    void Reset(int s, int e) { StartPosition = s; EndPosition = e; Initialize(); }  // :73-78
    bool IsPositionInRegion(int p) const { return p >= StartPosition && p <= EndPosition; }
    size_t Idx(int p, int a, int d, int anchor) const { return (((size_t)(p - StartPosition) * 6 + a) * 3 + d) * NA + anchor; }

    void AddAlleleCount(int p, AlleleType a, DirectionType d, int anchor) { if (IsPositionInRegion(p)) counts[Idx(p, a, d, anchor)]++; }  // :225-231
    void AddBaseQualites(int p, AlleleType a, DirectionType d, double bq, int anchor) { if (IsPositionInRegion(p)) qsum[Idx(p, a, d, anchor)] += bq; }
    void AddGappedMnvRefCount(int p, int c) { if (IsPositionInRegion(p)) gapped[p - StartPosition] += c; }  // :86-92
    void AddCollapsedReadCount(int p, ReadCollapsedType t) {  // CollapsedRegionState.cs:28-44
        if (!IsPositionInRegion(p)) return;
        int* row = &collapsed[(size_t)(p - StartPosition) * 8];
        row[t]++;
        if (t == SimplexReverseStitched || t == SimplexForwardStitched) row[SimplexStitched]++;
        else if (t == SimplexForwardNonStitched || t == SimplexReverseNonStitched) row[SimplexNonStitched]++;
    }
    int GetAlleleCount(int p, AlleleType a, DirectionType d, int minAnchor, std::optional<int> maxAnchor, bool fromEnd, bool symmetric) const {
        if (!IsPositionInRegion(p)) throw std::invalid_argument("Position is not in region");
        return AnchorAdjusted<int>(minAnchor, fromEnd, K, NA, &counts[Idx(p, a, d, 0)], maxAnchor, symmetric);  // :313-323
    }
    double GetSumOfAlleleBaseQualites(int p, AlleleType a, DirectionType d, int minAnchor, std::optional<int> maxAnchor, bool fromEnd, bool symmetric) const {
        if (!IsPositionInRegion(p)) throw std::invalid_argument("Position is not in region");
        return AnchorAdjusted<double>(minAnchor, fromEnd, K, NA, &qsum[Idx(p, a, d, 0)], maxAnchor, symmetric);  // :357-365
    }

    void UpdateMaxPosition(const CandidateAllele& c) {  // :204-223
        int otherEnd = 0;
        switch (c.Type) {
            case Deletion: otherEnd = c.ReferencePosition + (int)c.ReferenceAllele.size(); break;
            case Insertion: otherEnd = c.ReferencePosition + 1; break;
            case Mnv: otherEnd = c.ReferencePosition + (int)c.ReferenceAllele.size() - 1; break;
            default: break;
        }
        if (otherEnd > MaxAlleleEndpoint) MaxAlleleEndpoint = otherEnd;
    }
    void AddCandidate(const CandPtr& nc, bool trackOpenEnded, bool trackAmplicon = false) {  // :94-174
        if (nc->Type == Reference) throw std::invalid_argument("reference candidates are not tracked");
        if (!IsPositionInRegion(nc->ReferencePosition)) throw std::invalid_argument("Unable to add candidate to region");
        auto& existing = cands[nc->ReferencePosition - StartPosition];
        CandidateAllele* found = nullptr;
        for (auto& c : existing) {
            if (c->Equals(*nc) && (!trackOpenEnded || (c->OpenOnLeft == nc->OpenOnLeft && c->OpenOnRight == nc->OpenOnRight))) { found = c.get(); break; }
        }
        if (!found) existing.push_back(nc);
        else {
            for (int i = 0; i < 3; i++) found->SupportByDirection[i] += nc->SupportByDirection[i];
            for (int i = 0; i < 3; i++) found->WellAnchoredSupportByDirection[i] += nc->WellAnchoredSupportByDirection[i];
            for (int i = 0; i < 8; i++) found->ReadCollapsedCountsMut[i] += nc->ReadCollapsedCountsMut[i];
            if (trackAmplicon && !nc->SupportByAmplicon.isNull) {   // :138-170 (every merged occurrence counts 1, whatever the new candidate's own count says)
                for (int i = 0; i < MaxNumOverlappingAmplicons; i++) {
                    const int name = nc->SupportByAmplicon.names[(size_t)i];
                    if (name < 0) continue;
                    if (found->SupportByAmplicon.isNull) found->SupportByAmplicon = AmpliconCounts::Empty();
                    auto ix = found->SupportByAmplicon.Index(name);
                    if (ix.first == -1) {
                        if (ix.second < 0) throw std::out_of_range("Index was outside the bounds of the array.");
                        found->SupportByAmplicon.names[(size_t)ix.second] = name; found->SupportByAmplicon.counts[(size_t)ix.second] = 1;
                    } else found->SupportByAmplicon.counts[(size_t)ix.first]++;
                }
            }
        }
        UpdateMaxPosition(*nc);
    }
    // :383-453  forcedGtPositions == CreateIntervalsFromAllels(chrReference, forcesGtAlleles) when !includeRefAlleles
    std::vector<CandPtr> GetAllCandidates(bool includeRefAlleles, const std::string& chrName, const std::string& chrSeq,
                                          const ChrIntervalSet* intervals, const std::vector<int>* forcedGtPositions) const {
        std::vector<CandPtr> alleles;
        for (auto& l : cands) for (auto& c : l) alleles.push_back(c);
        ChrIntervalSet forcedSet;
        const ChrIntervalSet* inUse = nullptr;
        if (includeRefAlleles) inUse = intervals;
        else if (forcedGtPositions && !forcedGtPositions->empty()) {
            for (int p : *forcedGtPositions) forcedSet.Intervals.push_back(Region{p, p});
            inUse = &forcedSet;
        }
        bool haveForced = forcedGtPositions != nullptr && !forcedGtPositions->empty();
        if (includeRefAlleles || haveForced) {
            std::vector<Region> regionsToFetch = inUse == nullptr ? std::vector<Region>{Region{StartPosition, EndPosition}}
                                                                  : inUse->GetClipped(Region{StartPosition, EndPosition});
            for (auto& ci : regionsToFetch) {
                for (int position = ci.StartPosition; position <= ci.EndPosition; position++) {
                    if (position > (int)chrSeq.size()) break;
                    std::string refBase(1, chrSeq[position - 1]);
                    int refBaseIndex = (int)GetAlleleType(refBase[0]);
                    auto refAllele = std::make_shared<CandidateAllele>(chrName, position, refBase, refBase, Reference);
                    int totalSupport = 0;
                    for (int a = 0; a < NumAlleleTypes; a++)
                        for (int d = 0; d < NumDirectionTypes; d++) {
                            int count = 0;
                            for (int an = 0; an < NA; an++) count += counts[Idx(position, a, d, an)];
                            if (a == refBaseIndex) refAllele->SupportByDirection[d] = count;
                            totalSupport += count;
                        }
                    if (inUse != nullptr || totalSupport > 0) alleles.push_back(refAllele);
                }
            }
        }
        return alleles;
    }
    std::vector<CandPtr> ExtractCollapsable(int upToPosition) {  // :470-490
        std::vector<CandPtr> all;
        for (auto& lookup : cands) {
            std::vector<CandPtr> collapsables;
            for (auto& c : lookup)
                if (c->ReferencePosition + (int)c->AlternateAllele.size() - 1 <= upToPosition && !c->OpenOnRight && (c->Type == Mnv || c->Type == Snv))
                    collapsables.push_back(c);
            for (auto& c : collapsables) {
                all.push_back(c);
                // List.Remove uses CandidateAllele.Equals (value equality): removes the FIRST element equal to it
                for (auto it = lookup.begin(); it != lookup.end(); ++it) if ((*it)->Equals(*c)) { lookup.erase(it); break; }
            }
        }
        return all;
    }
};

// Models/CandidateBatch.cs
struct CandidateBatch {
    std::vector<CandPtr> candidates;
    std::optional<int> MaxClearedPosition;
    std::vector<int> BlockKeys;
    bool HasCandidates() const { return !candidates.empty(); }
};

// RegionStateManager.cs (+ CollapsedRegionStateManager)
struct RegionStateManager : IAlleleSource {
    std::map<int, std::shared_ptr<RegionState>> regionLookup;
    int regionSize = 1000;
    int minBasecallQuality;
    RegionState* lastAccessedBlock = nullptr;
    std::vector<std::shared_ptr<RegionState>> reusableBlocks;  // Stack
    int lastUpToBlockKey = 0;
    bool includeRefAlleles;
    ChrIntervalSet* intervalSet;
    bool trackOpenEnded;
    int numAnchorTypes;
    bool expectStitched, expectCollapsed;
    bool trackAmpliconCounts = false;   // RegionStateManager.cs:25,40,52

    RegionStateManager(bool includeRef, int minBQ, bool expectStitchedReads, ChrIntervalSet* intervals, int blockSize, bool trackOpen,
                       int anchorTypes, bool expectCollapsedReads, bool trackAmplicons = false)
        : regionSize(blockSize), minBasecallQuality(minBQ), includeRefAlleles(includeRef), intervalSet(intervals), trackOpenEnded(trackOpen),
          numAnchorTypes(anchorTypes), expectStitched(expectStitchedReads), expectCollapsed(expectCollapsedReads), trackAmpliconCounts(trackAmplicons) {}
    AmpliconCounts GetCoverageByAmplicon(int position) override {   // :228-232 (a position without a block dereferences null in the reference)
        RegionState* region = GetBlock(position, false);
        if (!region) throw std::runtime_error("Object reference not set to an instance of an object.");
        return region->GetCountsByAmpliconForPosition(position);
    }
    int WellAnchoredIndex() const { return numAnchorTypes; }
    int NumAnchorIndexes() const { return numAnchorTypes * 2 + 1; }
    bool ExpectStitchedReads() const override { return expectStitched; }

    int GetBlockKey(int position) const { return (int)std::ceil((double)position / regionSize); }  // :385-391
    RegionState* GetBlock(int position, bool addIfMissing = true) {  // :361-383
        if (position <= 0) throw std::invalid_argument("Position must be greater than 0.");
        if (lastAccessedBlock && lastAccessedBlock->IsPositionInRegion(position)) return lastAccessedBlock;
        int key = GetBlockKey(position);
        auto it = regionLookup.find(key);
        if (it == regionLookup.end()) {
            if (!addIfMissing) return nullptr;
            std::shared_ptr<RegionState> b;
            int s = (key - 1) * regionSize + 1, e = key * regionSize;
            if (!reusableBlocks.empty()) { b = reusableBlocks.back(); reusableBlocks.pop_back(); b->Reset(s, e); }  // :425-439
            else b = std::make_shared<RegionState>(s, e, numAnchorTypes);
            it = regionLookup.emplace(key, b).first;
        }
        lastAccessedBlock = it->second.get();
        return lastAccessedBlock;
    }
    void AddCandidates(const std::vector<CandPtr>& cs) override {  // :55-66
        for (auto& c : cs) GetBlock(c->ReferencePosition)->AddCandidate(c, trackOpenEnded, trackAmpliconCounts);
    }
    int GetAnchorType(int alignmentEndPosition, int basePosition, int alignmentStartPosition) const {  // :83-116
        int leftAnchor = basePosition - alignmentStartPosition;
        int rightAnchor = alignmentEndPosition - basePosition;
        int minAnchor;
        if (leftAnchor >= rightAnchor) {
            if (rightAnchor >= numAnchorTypes) return WellAnchoredIndex();
            minAnchor = NumAnchorIndexes() - rightAnchor - 1;
        } else {
            if (leftAnchor >= numAnchorTypes) return WellAnchoredIndex();
            minAnchor = leftAnchor;
        }
        if (minAnchor < 0) throw std::invalid_argument("Base position does not appear to be mapped in read");
        return minAnchor;
    }
    void AddCollapsedReadCount(int position, const Read& r, DirectionType d) {  // CollapedRegionStateManager.cs:40-52
        if (!expectCollapsed) return;  // RegionStateManager base: no-op (:265-268)
        if (!r.IsCollapsedRead()) throw std::runtime_error("The input is collapsed BAM, but read is not a collapsed read.");
        auto t = r.GetReadCollapsedType(d);
        if (t.has_value()) GetBlock(position)->AddCollapsedReadCount(position, *t);
    }
    void AddAlleleCounts(const Read& alignment) {  // :118-220
        int lastPosition = alignment.Position() - 1;
        int deletionLength = 0;
        int lengthBeforeDeletion = alignment.ReadLength();
        bool endsInDeletion = alignment.HasOperationAtOpIndex(0, 'D', true);
        bool endsInDeletionBeforeSoftclip = alignment.HasOperationAtOpIndex(1, 'D', true) && alignment.HasOperationAtOpIndex(0, 'S', true);
        auto& cigar = alignment.CigarData;
        if (endsInDeletion || endsInDeletionBeforeSoftclip) {
            deletionLength = (int)(endsInDeletionBeforeSoftclip ? cigar[cigar.size() - 2].Length : cigar[cigar.size() - 1].Length);
            lengthBeforeDeletion = (int)(endsInDeletionBeforeSoftclip ? alignment.ReadLength() - cigar[cigar.size() - 1].Length : alignment.ReadLength());
        }
        auto positionMap = alignment.PositionMap();
        auto dirMap = alignment.SequencedBaseDirectionMap();
        int alignmentEndPosition = alignment.EndPosition();
        int alignmentStartPosition = alignment.Position();
        for (int i = 0; i < (int)positionMap.size(); i++) {
            DirectionType directionType = dirMap[i];
            if (endsInDeletionBeforeSoftclip && i == lengthBeforeDeletion) {
                if (CheckDeletionQuality(alignment, i, minBasecallQuality)) {
                    for (int j = 1; j < deletionLength + 1; j++) {
                        int anchorIndex = NumAnchorIndexes() - 1;
                        GetBlock(j + lastPosition)->AddAlleleCount(j + lastPosition, AT_Del, directionType, anchorIndex);
                        AddCollapsedReadCount(j + lastPosition, alignment, directionType);
                    }
                }
            }
            int position = positionMap[i];
            if (position == -1) continue;
            int anchorType = GetAnchorType(alignmentEndPosition, position, alignmentStartPosition);
            if (CheckDeletionQuality(alignment, i, minBasecallQuality)) {
                for (int j = lastPosition + 1; j < position; j++) {
                    GetBlock(j)->AddAlleleCount(j, AT_Del, directionType, anchorType);
                    AddCollapsedReadCount(j, alignment, directionType);
                }
            }
            AlleleType alleleType = GetAlleleType(alignment.Sequence[i]);
            if (alignment.Qualities[i] < minBasecallQuality) alleleType = AT_N;
            GetBlock(position)->AddAlleleCount(position, alleleType, directionType, anchorType);
            if (alleleType != AT_N) {
                AddCollapsedReadCount(position, alignment, directionType);
                if (trackAmpliconCounts) GetBlock(position)->AddAmpliconCount(position, alignment.AmpliconName);   // :188, :407-416
            }
            // Math.Pow(10, -1 * (int)q / 10f): int / float -> float exponent (:191)
            float expo = (float)(-1 * (int)alignment.Qualities[i]) / 10.0f;
            GetBlock(position)->AddBaseQualites(position, alleleType, directionType, std::pow(10.0, (double)expo), anchorType);
            lastPosition = position;
        }
        if (endsInDeletion) {
            int lastIdx = (int)dirMap.size() - 1;
            if (CheckDeletionQuality(alignment, lastIdx, minBasecallQuality)) {
                for (int j = 1; j < deletionLength + 1; j++) {
                    DirectionType directionType = dirMap[lastIdx];
                    int anchorIndex = NumAnchorIndexes() - 1;
                    GetBlock(j + lastPosition)->AddAlleleCount(j + lastPosition, AT_Del, directionType, anchorIndex);
                    AddCollapsedReadCount(j + lastPosition, alignment, directionType);
                }
            }
        }
    }
    int GetAlleleCount(int position, AlleleType a, DirectionType d, int minAnchor = 0, std::optional<int> maxAnchor = std::nullopt,
                       bool fromEnd = false, bool symmetric = false) override {  // :222-226
        auto* region = GetBlock(position, false);
        return region == nullptr ? 0 : region->GetAlleleCount(position, a, d, minAnchor, maxAnchor, fromEnd, symmetric);
    }
    // NB the reference drops `symmetric` here (:68-72)
    double GetSumOfAlleleBaseQualities(int position, AlleleType a, DirectionType d, int minAnchor = 0, std::optional<int> maxAnchor = std::nullopt,
                                       bool fromEnd = false, bool /*symmetric*/ = false) override {
        auto* region = GetBlock(position, false);
        return region == nullptr ? 0 : region->GetSumOfAlleleBaseQualites(position, a, d, minAnchor, maxAnchor, fromEnd, false);
    }
    int GetCollapsedReadCount(int position, ReadCollapsedType t) override {  // :270-273 / CollapedRegionStateManager.cs:34-38
        if (!expectCollapsed) return 0;
        auto* region = GetBlock(position, false);
        return region == nullptr ? 0 : region->collapsed[(size_t)(position - region->StartPosition) * 8 + t];
    }
    void AddGappedMnvRefCount(const std::map<int, int>& lookup) override {  // :74-81
        for (auto& kv : lookup) GetBlock(kv.first)->AddGappedMnvRefCount(kv.first, kv.second);
    }
    int GetGappedMnvRefCount(int position) override {  // :256-261
        auto* region = GetBlock(position, false);
        return region == nullptr ? 0 : region->gapped[position - region->StartPosition];
    }
    // :283-334. Returns nullptr for "null batch".
    std::unique_ptr<CandidateBatch> GetCandidatesToProcess(std::optional<int> upToPosition, const std::string& chrName, const std::string& chrSeq,
                                                           const std::vector<int>* forcedGtPositions) {
        struct Finally { RegionStateManager* m; std::optional<int> up; ~Finally() { m->lastUpToBlockKey = up.has_value() ? m->GetBlockKey(*up) : -1; } } fin{this, upToPosition};
        if (upToPosition.has_value() && GetBlockKey(*upToPosition) == lastUpToBlockKey) return nullptr;
        auto batch = std::make_unique<CandidateBatch>();
        if (upToPosition.has_value()) batch->MaxClearedPosition = -1;
        std::vector<RegionState*> blocks;
        for (auto& kv : regionLookup) {  // std::map iterates keys in sorted order (Array.Sort(blockKeys))
            int key = kv.first;
            if (upToPosition.has_value() && !((long long)key * regionSize <= *upToPosition)) continue;
            auto* block = kv.second.get();
            if (upToPosition.has_value() && block->MaxAlleleEndpoint > *upToPosition) break;
            auto c = block->GetAllCandidates(includeRefAlleles, chrName, chrSeq, intervalSet, forcedGtPositions);
            batch->candidates.insert(batch->candidates.end(), c.begin(), c.end());
            batch->BlockKeys.push_back(key);
            blocks.push_back(block);
        }
        if (!blocks.empty()) {
            int maxEnd = INT32_MIN, maxEndpoint = INT32_MIN;
            for (auto* b : blocks) { maxEnd = std::max(maxEnd, b->EndPosition); maxEndpoint = std::max(maxEndpoint, b->MaxAlleleEndpoint); }
            batch->MaxClearedPosition = maxEnd;
            if (upToPosition.has_value() && maxEndpoint > maxEnd && trackOpenEnded) {  // AddCollapsableFromOtherBlocks :441-457
                for (auto& kv : regionLookup) {  // NB Dictionary enumeration order in the reference; insertion order == key order for sorted input
                    auto* block = kv.second.get();
                    if (block->StartPosition > maxEnd && block->StartPosition <= *upToPosition) {
                        auto cv = block->ExtractCollapsable(*upToPosition);
                        batch->candidates.insert(batch->candidates.end(), cv.begin(), cv.end());
                    }
                }
            }
        }
        return batch;
    }
    void DoneProcessing(const CandidateBatch& batch) {  // :336-353
        for (int key : batch.BlockKeys) {
            auto it = regionLookup.find(key);
            if (it == regionLookup.end()) continue;
            reusableBlocks.push_back(it->second);
            if (lastAccessedBlock == it->second.get()) lastAccessedBlock = nullptr;
            regionLookup.erase(it);
        }
    }
};

}  // namespace po
