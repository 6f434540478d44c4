// ORACLE — TEST INFRASTRUCTURE ONLY (see po_math.hpp header).
// Read model: src/lib/Pisces.Domain/Models/Read.cs, PositionMap.cs, CigarDirection.cs; Alignment.Domain/BamCommon.cs (CigarOp).
#pragma once
#include "po_types.hpp"

namespace po {

struct CigarOp {
    char Type;
    uint32_t Length;
    bool IsReferenceSpan() const {  // BamCommon.cs:560-573
        switch (Type) { case 'M': case 'D': case 'N': case '=': case 'X': return true; default: return false; }
    }
    bool IsReadSpan() const {  // BamCommon.cs:575-588
        switch (Type) { case 'M': case 'I': case 'S': case '=': case 'X': return true; default: return false; }
    }
};

struct Read {
    std::string Name;
    int BamPosition = 0;  // 0-based (BamAlignment.Position)
    std::vector<CigarOp> CigarData;
    std::string Sequence;
    std::vector<uint8_t> Qualities;
    uint32_t MapQuality = 0;
    // flag-derived
    bool IsMapped = true, IsPrimaryAlignment = true, IsPcrDuplicate = false, IsProperPair = false, IsReverseStrand = false,
         IsFirstMate = false;
    // tags (TagData present iff hasTagData)
    bool hasTagData = false;
    std::optional<std::string> XD, XR;
    std::optional<int> XV, XW;
    int AmpliconName = -1;   // Read.GetAmpliconNameIfExists (Read.cs:483-486): the XN tag as an index into the caller's name dictionary, -1 = no tag

    int Position() const { return BamPosition + 1; }  // Read.cs:81
    uint32_t ReferenceSpan() const { uint32_t l = 0; for (auto& o : CigarData) if (o.IsReferenceSpan()) l += o.Length; return l; }
    uint32_t ReadSpan() const { uint32_t l = 0; for (auto& o : CigarData) if (o.IsReadSpan()) l += o.Length; return l; }
    // BamAlignment.EndPosition = Position + refSpan - 1 (BamCommon.cs:119); Read.EndPosition = that + 1 (Read.cs:88-91)
    int EndPosition() const { return (BamPosition + (int)ReferenceSpan() - 1) + 1; }
    int ReadLength() const { return (int)Sequence.size(); }
    bool HasCigar() const { return !CigarData.empty(); }
    bool HasOperationAtOpIndex(int index, char type, bool fromEnd) const {  // Utility/CigarExtensions.cs:38-44
        int count = (int)CigarData.size();
        int opIndex = fromEnd ? count - index - 1 : index;
        return count > opIndex && opIndex >= 0 && CigarData[opIndex].Type == type;
    }

    // Read.UpdatePositionMap  Read.cs:535-562 ; -1 for bases not mapped to reference
    std::vector<int> PositionMap() const {
        std::vector<int> map((size_t)ReadLength(), -1);
        if (CigarData.empty()) return map;
        if ((int)ReadSpan() != ReadLength()) throw std::runtime_error("Invalid cigar: does not match length of read");  // :603-605
        int readIndex = 0, referencePosition = Position();
        for (auto& op : CigarData) {
            bool readSpan = op.IsReadSpan(), refSpan = op.IsReferenceSpan();
            for (uint32_t i = 0; i < op.Length; i++) {
                if (readSpan) { map[readIndex] = refSpan ? referencePosition++ : -1; readIndex++; }
                else if (refSpan) referencePosition++;
            }
        }
        return map;
    }
    static DirectionType ParseDir(char c) {
        switch (c) { case 'F': return Forward; case 'R': return Reverse; case 'S': return Stitched; default: throw std::runtime_error("bad direction char"); }
    }
    // CigarDirection(string) + Expand()  CigarDirection.cs:19-83
    static std::vector<DirectionType> ExpandDirectionString(const std::string& s) {
        std::vector<DirectionType> out;
        size_t head = 0;
        for (size_t i = 0; i < s.size(); ++i) {
            if (s[i] >= '0' && s[i] <= '9') continue;
            DirectionType d = ParseDir(s[i]);
            int len = std::stoi(s.substr(head, i - head));
            for (int k = 0; k < len; k++) out.push_back(d);
            head = i + 1;
        }
        if (head != s.size()) throw std::runtime_error("Unexpected format in direction string");
        return out;
    }
    bool HasCigarDirections() const { return hasTagData && XD.has_value(); }  // Read.cs:351-362
    std::vector<DirectionType> ExpandedBaseDirectionMap() const { return ExpandDirectionString(*XD); }
    // Read.SetSequencedBaseDirectionMapFromBam  Read.cs:390-421 ; CreateSequencedBaseDirectionMap :664-682
    std::vector<DirectionType> SequencedBaseDirectionMap() const {
        std::vector<DirectionType> m((size_t)ReadLength(), Forward);
        if (HasCigarDirections() && !XD->empty()) {
            auto expanded = ExpandedBaseDirectionMap();
            std::vector<const CigarOp*> ops;  // CigarData.Expand(): one entry per op unit, every op type
            for (auto& op : CigarData) for (uint32_t i = 0; i < op.Length; i++) ops.push_back(&op);
            std::vector<DirectionType> seq((size_t)ReadSpan(), Forward);
            size_t si = 0;
            for (size_t ci = 0; ci < expanded.size(); ci++) {
                if (ci >= ops.size()) throw std::runtime_error("direction map longer than cigar");
                if (ops[ci]->IsReadSpan()) { seq.at(si) = expanded[ci]; si++; }
            }
            return seq;
        }
        for (auto& d : m) d = IsReverseStrand ? Reverse : Forward;
        return m;
    }
    // ReadExtentions.IsCollapsedRead  Read.cs:66-71
    bool IsCollapsedRead() const { return hasTagData && (XV.has_value() || XW.has_value()); }
    bool IsDuplex() const {  // Read.cs:311-331
        if (hasTagData) {
            if (!XV.has_value() || *XV == 0) return false;
            if (!XW.has_value() || *XW == 0) return false;
            return true;
        }
        return false;
    }
    std::optional<std::string> ReadPairDirection() const {  // Read.cs:333-349
        std::optional<std::string> xr;
        if (hasTagData) xr = XR;
        if (!xr.has_value() && IsProperPair) {
            char dir = IsReverseStrand ? 'R' : 'F';
            char dirmate = (dir == 'F') ? 'R' : 'F';
            xr = IsFirstMate ? std::string{dir, dirmate} : std::string{dirmate, dir};
        }
        return xr;
    }
    std::optional<ReadCollapsedType> GetReadCollapsedType(DirectionType d) const {  // Read.cs:17-64
        auto rpd = ReadPairDirection();
        if (IsDuplex()) return d == Stitched ? DuplexStitched : DuplexNonStitched;
        if (d == Stitched) {
            if (rpd && *rpd == "FR") return SimplexForwardStitched;
            if (rpd && *rpd == "RF") return SimplexReverseStitched;
            return std::nullopt;
        }
        if (rpd && *rpd == "FR") return SimplexForwardNonStitched;
        if (rpd && *rpd == "RF") return SimplexReverseNonStitched;
        return std::nullopt;
    }
    int MaxPosition() const { int m = INT32_MIN; for (int p : PositionMap()) if (p > m) m = p; return m; }
};

// CandidateVariantFinder.CheckDeletionQuality  CandidateVariantFinder.cs:294-320
inline bool CheckDeletionQuality(const Read& r, int opStartIndexInRead, int minBQ) {
    if (r.Qualities.empty()) return false;
    int n = (int)r.Qualities.size();
    int after = (opStartIndexInRead < n) ? r.Qualities[opStartIndexInRead] : r.Qualities[opStartIndexInRead - 1];
    int before = after;
    if (opStartIndexInRead > 0) before = r.Qualities[opStartIndexInRead - 1];
    return (before >= minBQ) && (after >= minBQ);
}

}  // namespace po
