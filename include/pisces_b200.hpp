// Header-only C++ mirror of the reference's plug-in interfaces for the hot path, over the C ABI of pisces_b200.h.
//   pb2::GpuStateManager  ~ IStateManager : IAlleleSource   (src/lib/Pisces.Processing/Interfaces/IStateManager.cs:8-15,
//                                                            src/lib/Pisces.Domain/Interfaces/IAlleleSource.cs:8-26)
//   pb2::GpuAlleleCaller  ~ IAlleleCaller                   (src/exe/Pisces/Interfaces/IAlleleCaller.cs:8-13)
// Same method names and argument meaning as the reference; errors become exceptions here exactly as the C# shim turns non-zero return
// codes into exceptions. All compute is in libpisces_b200.so on the GPU.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>
#include "pisces_b200.h"

namespace pb2 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("pisces_b200 error " + std::to_string(c) + ": " + m), code(c) {}
};

// The fields of Pisces.Domain.Models.Read (src/lib/Pisces.Domain/Models/Read.cs) the path consumes.
struct Read {
    int32_t Position = 0;                  // 1-based
    uint16_t Flag = 0;
    std::vector<uint32_t> CigarData;       // BAM encoding len<<4|op
    std::string Sequence;
    std::vector<uint8_t> Qualities;
    std::vector<uint8_t> SequencedBaseDirectionMap;   // optional (stitched reads)
    int CollapsedSummary = -1;             // optional, see pb2_read_batch.collapsed
    int32_t AmpliconId = -1;               // optional: Read.GetAmpliconNameIfExists as an id of the host's dictionary (pb2_read_batch.amplicon), -1 = no XN tag
};

class GpuStateManager {
public:
    explicit GpuStateManager(const pb2_config& cfg, const std::string& chrName = "chr1", const std::string& chrSequence = std::string()) : cfg_(cfg) {
        const int rc = pb2_create(&cfg_, &h_);
        if (rc != PB2_OK) throw Error(rc, pb2_last_error(nullptr));
        if (!chrSequence.empty()) check(pb2_set_reference(h_, chrName.c_str(), reinterpret_cast<const uint8_t*>(chrSequence.data()), (int64_t)chrSequence.size()));
    }
    ~GpuStateManager() { pb2_destroy(h_); }
    GpuStateManager(const GpuStateManager&) = delete;
    GpuStateManager& operator=(const GpuStateManager&) = delete;

    bool ExpectStitchedReads() const { return cfg_.expect_stitched != 0; }
    void SetIntervals(const std::vector<int32_t>& start, const std::vector<int32_t>& end) { check(pb2_set_intervals(h_, start.data(), end.data(), (int32_t)start.size())); }

    // forcedGtAlleles of Factory.CreateSomaticVariantCaller (Factory.cs:253) for this chromosome: (position, ref, alt)
    void SetForcedAlleles(const std::vector<std::tuple<int32_t, std::string, std::string>>& alleles) {
        std::vector<pb2_candidate> cs(alleles.size());
        std::vector<uint8_t> arena;
        for (size_t i = 0; i < alleles.size(); i++) {
            pb2_candidate c{};
            c.position = std::get<0>(alleles[i]);
            const std::string &ref = std::get<1>(alleles[i]), &alt = std::get<2>(alleles[i]);
            c.ref_len = (uint16_t)ref.size(); c.alt_len = (uint16_t)alt.size(); c.allele_offset = (uint32_t)arena.size();
            arena.insert(arena.end(), ref.begin(), ref.end());
            arena.insert(arena.end(), alt.begin(), alt.end());
            cs[i] = c;
        }
        arena.push_back(0);
        check(pb2_set_forced_alleles(h_, cs.data(), (int32_t)cs.size(), arena.data(), (int64_t)arena.size()));
    }

    // IStateManager.AddAlleleCounts(Read) (+ the SNV part of ICandidateVariantFinder.FindCandidates): buffered, expanded on the device at the next Call
    void AddAlleleCounts(const Read& r) { buffer_.push_back(r); if (buffer_.size() >= 65536) PushBuffered(); }
    // IAlleleSource.GetAlleleCount over a window: int32 [n][6][3][11]
    std::vector<int32_t> GetAlleleCounts(int32_t position0, int32_t n) {
        PushBuffered();
        std::vector<int32_t> out((size_t)n * 198);
        check(pb2_get_counts(h_, position0, n, out.data()));
        return out;
    }
    void DoneProcessing() { check(pb2_reset(h_)); }
    // VcfFileWriter's record lines for called alleles (src/lib/Pisces.IO/VcfFileWriter.cs:177-260): options.crushed = !AllowMultipleVcfLinesPerLoci,
    // options.pad_intervals replays RegionMapper. `ext` (pb2_flush_ext) may be null.
    std::string FormatVcf(const pb2_call_record* records, const pb2_call_record_ext* ext, int64_t n, const pb2_vcf_options& options) {
        const char* text = nullptr;
        int64_t len = 0;
        check(pb2_vcf_format(h_, records, ext, n, &options, &text, &len));
        return std::string(text, (size_t)len);
    }
    pb2_handle* handle() { return h_; }

    void PushBuffered() {
        if (buffer_.empty()) return;
        std::vector<int32_t> pos0; std::vector<uint16_t> flag; std::vector<int64_t> coff{0}, soff{0};
        std::vector<uint32_t> cigar; std::vector<uint8_t> bases, quals, dirs, coll;
        std::vector<int32_t> amp;
        bool anyDirs = false, anyColl = false, anyAmp = false;
        for (auto& r : buffer_) { anyDirs |= !r.SequencedBaseDirectionMap.empty(); anyColl |= r.CollapsedSummary >= 0; }
        for (auto& r : buffer_) {
            pos0.push_back(r.Position - 1); flag.push_back(r.Flag);
            cigar.insert(cigar.end(), r.CigarData.begin(), r.CigarData.end()); coff.push_back((int64_t)cigar.size());
            bases.insert(bases.end(), r.Sequence.begin(), r.Sequence.end());
            quals.insert(quals.end(), r.Qualities.begin(), r.Qualities.end());
            if (anyDirs) {
                if (!r.SequencedBaseDirectionMap.empty()) dirs.insert(dirs.end(), r.SequencedBaseDirectionMap.begin(), r.SequencedBaseDirectionMap.end());
                else dirs.insert(dirs.end(), r.Sequence.size(), (uint8_t)((r.Flag & 0x10) ? 1 : 0));
            }
            if (anyColl) coll.push_back((uint8_t)(r.CollapsedSummary < 0 ? 0 : r.CollapsedSummary));
            amp.push_back(r.AmpliconId);
            anyAmp = anyAmp || r.AmpliconId >= 0;
            soff.push_back((int64_t)bases.size());
        }
        pb2_read_batch b{(int32_t)buffer_.size(), pos0.data(), flag.data(), coff.data(), cigar.data(), soff.data(), bases.data(), quals.data(),
                         anyDirs ? dirs.data() : nullptr, anyColl ? coll.data() : nullptr, anyAmp ? amp.data() : nullptr};
        check(pb2_push_reads(h_, &b));
        buffer_.clear();
    }
    void check(int rc) { if (rc != PB2_OK) throw Error(rc, pb2_last_error(h_)); }

private:
    pb2_config cfg_;
    pb2_handle* h_ = nullptr;
    std::vector<Read> buffer_;
};

class GpuAlleleCaller {
public:
    int TotalNumCalled = 0;
    int TotalNumCollapsed = 0;
    // IAlleleCaller.Call(batch, source): SortedList<int, List<CalledAllele>> as position -> records, ordered by (ref, alt) within a position
    std::map<int32_t, std::vector<pb2_call_record>> Call(GpuStateManager& source, int32_t upToPosition = -1) {
        source.PushBuffered();
        const pb2_call_record* recs = nullptr;
        int64_t n = 0;
        source.check(pb2_flush(source.handle(), upToPosition, &recs, &n));
        std::map<int32_t, std::vector<pb2_call_record>> out;
        for (int64_t i = 0; i < n; i++) out[recs[i].position].push_back(recs[i]);
        TotalNumCalled += (int)n;
        return out;
    }
};

}  // namespace pb2
