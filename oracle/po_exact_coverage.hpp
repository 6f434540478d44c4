// ORACLE - TEST INFRASTRUCTURE ONLY. ExactCoverageCalculator's per-read decision (src/lib/Pisces.Calculators/ExactCoverageCalculator.cs): does a read
// summary contribute to the spanning coverage of an insertion / deletion / MNV, and in which direction. The product does not build this calculator yet
// (SURVEY 8f rank 4); this pins the restatement on the reference's ExactCoverageCalculatorTests.
#pragma once
#include <utility>
#include <vector>
#include "po_read.hpp"
#include "po_types.hpp"

namespace po {

// Read.UpdatePositionMap with differentiateSoftClip (Read.cs:535-562): -1 for inserted bases, -2 for soft-clipped ones
inline std::vector<int> PositionMapDifferentiatingSoftClips(int position, const std::vector<CigarOp>& cigar) {
    std::vector<int> map;
    int referencePosition = position;
    for (const CigarOp& op : cigar)
        for (uint32_t k = 0; k < op.Length; k++) {
            if (op.IsReadSpan()) map.push_back(op.IsReferenceSpan() ? referencePosition++ : (op.Type == 'S' ? -2 : -1));
            else if (op.IsReferenceSpan()) referencePosition++;
        }
    return map;
}
inline bool HasOperationAtOpIndex(const std::vector<CigarOp>& cigar, int index, char type, bool fromEnd = false) {   // CigarExtensions.cs:38-44
    const int opIndex = fromEnd ? (int)cigar.size() - index - 1 : index;
    return (int)cigar.size() > opIndex && opIndex >= 0 && cigar[(size_t)opIndex].Type == type;
}
inline uint32_t GetPrefixClip(const std::vector<CigarOp>& cigar) {   // BamCommon.cs:787-802
    uint32_t length = 0;
    for (const CigarOp& op : cigar) {
        if (op.Type == 'S') length += op.Length;
        else if (op.Type != 'H') break;
    }
    return length;
}
// GetIndexBoundaries (:146-187)
inline std::pair<int, int> ExactIndexBoundaries(int startPosition, int endPosition, const std::vector<int>& positionMap) {
    int startIndex = -1, endIndex = -1;
    bool hasStart = false, hasEnd = false;
    const int n = (int)positionMap.size();
    for (int i = 0; i < n; i++) {
        const int positionAtIndex = positionMap[(size_t)i];
        if (positionAtIndex >= 0 && positionAtIndex <= startPosition) { startIndex = i; hasStart = true; }
        if (!hasEnd && positionMap[(size_t)i] >= endPosition) { endIndex = i; hasEnd = true; }
    }
    if (hasStart && !hasEnd && n > 0 && positionMap[(size_t)n - 1] == -2)
        for (int i = startIndex + 1; i < n; i++) if (positionMap[(size_t)i] == -2) { endIndex = i; hasEnd = true; break; }
    if (hasEnd && !hasStart && n > 0 && positionMap[0] == -2)
        for (int i = endIndex - 1; i >= 0; i--) if (positionMap[(size_t)i] == -2) { startIndex = i; hasStart = true; break; }
    return {hasStart ? startIndex : -1, hasEnd ? endIndex : -1};
}
// GetDirection (:111-144); -3 where the reference throws (both indices -1)
inline int ExactDirection(int precedingIndex, int trailingIndex, const std::vector<DirectionType>& directionMap) {
    DirectionType direction = Forward;
    if (precedingIndex == -1 && trailingIndex == -1) return -3;
    if (trailingIndex == precedingIndex + 1) {
        if (precedingIndex == -1) direction = directionMap[(size_t)trailingIndex];
        else if (trailingIndex == -1) direction = directionMap[(size_t)precedingIndex];
        else {
            direction = directionMap[(size_t)precedingIndex];
            if (direction == Stitched) direction = directionMap[(size_t)trailingIndex];
        }
    } else {
        if (trailingIndex == -1) trailingIndex = (int)directionMap.size();
        for (int i = precedingIndex + 1; i <= trailingIndex - 1; i++) {
            direction = directionMap[(size_t)i];
            if (direction == Stitched) break;
        }
    }
    return (int)direction;
}
// One spanning read of CalculateSpanning (:45-109) for the allele of Compute (:18-43). Returns the direction the read's coverage goes to, -1 when the
// read does not contribute, -2 for allele types that take the single-point path, -3 where the reference throws.
inline int ExactSpanningReadDirection(AlleleCategory type, int referencePosition, int alleleLength, int clipAdjustedStart, int clipAdjustedEnd,
                                      const std::vector<CigarOp>& cigar, const std::vector<std::pair<int, DirectionType>>& directions) {
    int precedingPosition, trailingPosition;
    switch (type) {
        case Deletion: precedingPosition = referencePosition; trailingPosition = referencePosition + alleleLength + 1; break;
        case Mnv: precedingPosition = referencePosition - 1; trailingPosition = referencePosition + alleleLength; break;
        case Insertion: precedingPosition = referencePosition; trailingPosition = referencePosition + 1; break;
        default: return -2;
    }
    if ((clipAdjustedEnd < precedingPosition || clipAdjustedStart > trailingPosition) ||
        (clipAdjustedEnd == precedingPosition && !HasOperationAtOpIndex(cigar, 0, 'I', true)) ||
        (clipAdjustedStart == trailingPosition && !HasOperationAtOpIndex(cigar, 0, 'I')))
        return -1;
    if (directions.size() == 1) return (int)directions[0].second;
    uint32_t readLength = 0;
    for (const CigarOp& op : cigar) if (op.IsReadSpan()) readLength += op.Length;
    std::vector<DirectionType> directionMap(readLength, Forward);   // Read.UpdateDirectionMap (:521-533)
    size_t mapIndex = 0;
    for (auto& d : directions) {
        for (int i = 0; i < d.first; i++) { if (mapIndex + (size_t)i >= directionMap.size()) return -3; directionMap[mapIndex + (size_t)i] = d.second; }
        mapIndex += (size_t)d.first;
    }
    const std::vector<int> positionMap = PositionMapDifferentiatingSoftClips(clipAdjustedStart - (int)GetPrefixClip(cigar), cigar);
    const std::pair<int, int> b = ExactIndexBoundaries(precedingPosition, trailingPosition, positionMap);
    return ExactDirection(b.first, b.second, directionMap);
}

}  // namespace po
