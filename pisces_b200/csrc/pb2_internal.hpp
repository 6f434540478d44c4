// Internal to libpisces_b200.so: the handle, a staged segment, the read buffer and the explicit-candidate table (host side).
#pragma once
#include <map>
#include <set>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>
#include "pb2_candidates.cuh"
#include "pb2_kernels.cuh"
#include "pb2_math.cuh"

namespace pb2 {
struct ReadsView {
    int32_t n_reads;
    const int32_t* pos0;
    const uint16_t* flag;
    const int64_t* cigar_off;
    const uint32_t* cigar;
    const int64_t* seq_off;
    const uint8_t* bases;
    const uint8_t* quals;
    const uint8_t* base_dirs;
    const uint8_t* collapsed;
    const uint8_t* slots = nullptr;   // PVERT slot byte of every read base (pb2_pvert.cuh), derived at ingest; 16 bytes of slack on either side
    const int32_t* amplicon = nullptr;   // [n] amplicon name id of a read (-1 none); only with amplicon tracking
};
struct RegionView {
    int32_t lo, hi;
    const int32_t* index_of_pos;   // [hi - lo + 1] locus of a position, -1 = not staged; nullptr: locus = position - lo
    const uint8_t* chr;
    int64_t chr_len;
    int min_bq;
    int expect_collapsed;
    const int32_t* index_ge = nullptr;    // [hi - lo + 2] first locus at or after a position (with index_of_pos; n_loci past the last)
    const int32_t* positions = nullptr;   // [n_loci] position of a locus (with index_of_pos)
    int64_t n_loci = 0;
};
cudaError_t launch_reads_count(const ReadsView& rv, const RegionView& rg, unsigned int* depth, cudaStream_t st);
cudaError_t launch_reads_emit(const ReadsView& rv, const RegionView& rg, const int64_t* offsets, unsigned int* cursor, uint8_t* code, uint8_t* qual, uint8_t* anch,
                              cudaStream_t st);
cudaError_t launch_depth_to_i64(const unsigned int* depth, int64_t* out, int64_t n, cudaStream_t st);

// One candidate found in one read by reads_candidates_kernel. 32 bytes.
struct RawCand {
    int32_t read;            // index of the read in the pushed buffer
    int32_t order;           // (CIGAR operation index << 16) | offset in the operation: FindCandidates' emission order inside the read
    int32_t position;        // ReferencePosition
    int32_t start_in_read;   // first inserted / variant base in the read (insertions: the alt allele is the reference base + bases[start, start+len))
    uint16_t ref_len, alt_len;
    uint8_t type, dir;       // AlleleCategory, DirectionType of the support
    uint8_t flags;           // bit0 OpenOnLeft, bit1 OpenOnRight, bit2 well anchored
    uint8_t collapsed;       // ReadCollapsedType + 1, or 0
    uint8_t read_bases[8];   // the first (up to 8) read bases of the allele, bases[start_in_read ...]: the reads themselves stay on the device
};
static_assert(sizeof(RawCand) == 32, "RawCand layout");
cudaError_t launch_reads_candidates(const ReadsView& rv, int32_t first_read, const uint8_t* chr, int64_t chr_len, int min_bq, int call_mnvs, int max_mnv, int max_gap,
                                    int expect_collapsed, RawCand* out, unsigned long long* count, int64_t capacity, int32_t pos_lo, int32_t pos_hi, int snv_only,
                                    const int32_t* end_pos, cudaStream_t st);

// One distinct candidate of a read set after the device-side reduce by key (RegionState.AddCandidate): the counts of all its occurrences and where it
// was raised first.
struct CandGroup {
    int32_t support[3], well_anchored[3], collapsed_mut[8];
    uint32_t first_index;            // index of (one of) its raw candidates
    uint32_t pad_;
    unsigned long long first_seen;   // (read << 32) | order of its first occurrence: FindCandidates' emission order
};
cudaError_t cand_reduce_temp_bytes(int64_t n, size_t* bytes);
cudaError_t launch_cand_group(const RawCand* raw, int64_t n, unsigned long long* keys_in, unsigned long long* keys_out, uint32_t* idx_in, uint32_t* idx_out, int32_t* head,
                              int32_t* group_of, void* temp, size_t temp_bytes, int32_t* flags, cudaStream_t st);
cudaError_t launch_cand_reduce(const RawCand* raw, const uint32_t* idx, const int32_t* group_of, int64_t n, CandGroup* groups, int64_t n_groups, cudaStream_t st);

// pb2_push_reads on the device: the new reads [first, n) of the store are validated (Read.cs:603-605, RegionStateManager.cs:363-364), their offsets
// rebased, Read.EndPosition computed, and the positions where SmallVariantCaller.Execute would have called a batch collected
// (SmallVariantCaller.cs:99-104: Call(read.Position - 1) whenever that enters a new 1000-bp block key).
struct IngestStatus {
    int32_t error;            // 0 ok, 1 bad CIGAR operation, 2 CIGAR does not match the read length, 3 negative position, 4 offsets not monotone,
                              // 5 a read without XV / XW in a collapsed BAM
    int32_t error_read;
    int32_t n_triggers;
    int32_t min_start, max_end;   // 1-based first / last reference position covered by the new reads
    int32_t pad_[3];
};
cudaError_t launch_reads_ingest(const ReadsView& rv, int32_t first_read, int64_t cigar_base, int64_t seq_base, int64_t* cigar_off, int64_t* seq_off, int32_t* end_pos,
                                int32_t prev_key, int2* triggers, int32_t trigger_capacity, IngestStatus* status, int expect_collapsed, cudaStream_t st);
// keep[i] = read i ends after `cleared_to`; compaction of the store after a partial flush (exclusive scans of the flags / lengths by the caller)
cudaError_t launch_reads_keep_flags(const int32_t* end_pos, int64_t n, int32_t cleared_to, const int64_t* cigar_off, const int64_t* seq_off, int64_t* keep_reads,
                                    int64_t* keep_cigar, int64_t* keep_seq, cudaStream_t st);
struct ReadsCompactArgs {
    int64_t n;
    const int64_t *new_index, *new_cigar, *new_seq;   // exclusive scans, [n + 1]
    const int32_t* pos0; const int32_t* end_pos; const uint16_t* flag; const int64_t* cigar_off; const uint32_t* cigar; const int64_t* seq_off;
    const uint8_t *bases, *quals, *base_dirs, *collapsed;
    const int32_t* amplicon = nullptr; int32_t* o_amplicon = nullptr;
    int32_t* o_pos0; int32_t* o_end_pos; uint16_t* o_flag; int64_t* o_cigar_off; uint32_t* o_cigar; int64_t* o_seq_off;
    uint8_t *o_bases, *o_quals, *o_base_dirs, *o_collapsed;
    const uint8_t* slots; uint8_t* o_slots;   // (already offset by the leading slack)
};
// slot bytes (pb2_pvert.cuh) of the bases [first, first + n) of the store
cudaError_t launch_reads_slots(const uint8_t* bases, const uint8_t* quals, int64_t n, uint8_t* slots, cudaStream_t st);
cudaError_t launch_reads_compact(const ReadsCompactArgs& a, cudaStream_t st);
// 1000-bp blocks touched by the reads at positions > cleared_through: bit (key - key0) of the bitmap
cudaError_t launch_reads_block_bitmap(const int32_t* pos0, const int32_t* end_pos, int64_t n, int32_t cleared_through, int32_t key0, int32_t n_keys, uint32_t* bitmap,
                                      cudaStream_t st);
}  // namespace pb2
#include "pb2_pvert.cuh"

// Reads staged by pb2_push_reads: a struct of arrays in device memory (the handle's pool), kept until a flush clears the positions they cover. The host
// keeps no per-read state: what the host-side replay of SmallVariantCaller.Execute needs (batch triggers, touched blocks, extent) is computed by kernels.
// std::allocator whose value-less construct() default-initialises (no zero fill for PODs)
template <class T>
struct DefaultInitAllocator : std::allocator<T> {
    template <class U> struct rebind { using other = DefaultInitAllocator<U>; };
    DefaultInitAllocator() = default;
    template <class U> DefaultInitAllocator(const DefaultInitAllocator<U>&) {}
    template <class U, class... Args>
    void construct(U* p, Args&&... args) {
        if constexpr (sizeof...(Args) == 0) ::new ((void*)p) U;
        else ::new ((void*)p) U(std::forward<Args>(args)...);
    }
};
template <class T>
struct GrowBuf {
    T* p = nullptr;
    size_t cap = 0;
};
struct DeviceReads {
    int64_t n = 0, n_cigar = 0, n_seq = 0;
    GrowBuf<int32_t> pos0, end_pos;
    GrowBuf<uint16_t> flag;
    GrowBuf<int64_t> cigar_off, seq_off;   // [n + 1]
    GrowBuf<uint32_t> cigar;
    GrowBuf<uint8_t> bases, quals, base_dirs, collapsed;
    GrowBuf<uint8_t> slots;   // [16 + n_seq + 16]
    GrowBuf<int32_t> amplicon;   // [n] with amplicon tracking (pb2_config.amplicon_bias_filter >= 0), else unused
    bool has_dirs = false, has_collapsed = false, has_amplicon = false;
    int32_t min_start = INT32_MAX, max_end = 0;   // extent of the stored reads (1-based positions)
    int32_t last_pos0 = -1;                       // Position of the read pushed last (-1: none yet)
    size_t size() const { return (size_t)n; }
    pb2::ReadsView view() const {
        return pb2::ReadsView{(int32_t)n, pos0.p, flag.p, cigar_off.p, cigar.p, seq_off.p, bases.p, quals.p, has_dirs ? base_dirs.p : nullptr,
                              has_collapsed ? collapsed.p : nullptr, slots.p ? slots.p + 16 : nullptr, has_amplicon ? amplicon.p : nullptr};
    }
};

// CandidateAllele (src/lib/Pisces.Domain/Models/Alleles/CandidateAllele.cs:8-125) on the host: one row of the explicit-candidate table.
struct HostCand {
    int32_t position = 0;
    uint8_t type = 0;              // AlleleCategory
    bool open_left = false, open_right = false;
    std::string ref, alt;
    int32_t support[3] = {0, 0, 0}, well_anchored[3] = {0, 0, 0}, collapsed_mut[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float frequency = 0;           // CandidateAllele.Frequency, set by the collapser
    bool alive = true;
    int Support() const { return support[0] + support[1] + support[2]; }
    int WellAnchored() const { return well_anchored[0] + well_anchored[1] + well_anchored[2]; }
    bool FullyAnchored() const { return !open_left && !open_right; }
    int Length() const {           // BaseAllele.Length (BaseAllele.cs:24-43)
        switch (type) { case 1: return (int)alt.size() - 1; case 2: return (int)ref.size() - 1; case 4: return (int)ref.size(); default: return (int)alt.size(); }
    }
    bool Equals(const HostCand& o) const { return o.position == position && o.type == type && o.alt == alt && o.ref == ref; }   // CandidateAllele.Equals (:56-66)
};

struct Segment {   // one staged pileup (pb2_push_pileup*)
    int64_t n_loci = 0;
    int32_t n_tiles = 0;
    int32_t first_position = 0;
    bool has_positions = false;
    int64_t plane_bytes = 0;
    int64_t n_entries = 0;
    int32_t max_depth = 0;
    size_t alloc_plane = 0, alloc_ref = 0, alloc_var = 0, alloc_pending = 0;
    // device
    int32_t* depth = nullptr;
    int32_t* pad = nullptr;
    int64_t* tile_base = nullptr;
    uint8_t *code = nullptr, *qual = nullptr, *anch = nullptr, *ref_base = nullptr;
    int32_t* positions = nullptr;
    // PVERT form (pb2_pvert.cuh; segments built from pushed reads): pv_data != nullptr, and none of the PTILE32 / PNIB16 planes exist
    uint8_t* pv_data = nullptr;
    int2* pv_row_meta = nullptr;
    int32_t* pv_row_amp = nullptr;   // [n_rows] amplicon id of the row's read (base rows; -1 none), with amplicon tracking only
    int64_t* pv_tile_row0 = nullptr;
    int32_t* pv_cls_end = nullptr;
    int32_t pv_classes = 0;
    int64_t pv_rows = 0;
    // PNIB16 form of the same pileup (pileup_nib_score_kernel); nib == nullptr: not staged (stitched directions, collapsed reads, very deep loci)
    uint8_t* nib = nullptr;
    int64_t* nib_tile_base = nullptr;
    int32_t *nib_store = nullptr, *nib_depth = nullptr;
    int32_t n_nib_tiles = 0, nib_max_store = 0;
    int64_t nib_bytes = 0;
    pb2_call_record* ref_records = nullptr;
    uint8_t* ref_valid = nullptr;
    pb2_call_record* var_records = nullptr;
    int64_t var_capacity = 0;
    uint32_t* exc_entries = nullptr;
    int64_t exc_capacity = 0;
    pb2::PendingLocus* pending = nullptr;
    int64_t pending_capacity = 0;
    unsigned long long* counters = nullptr;   // [0] var_count, [1] exc_count, [2] pending_count
    // host copies needed for ordering / lookups
    std::vector<int32_t> h_positions;
    bool called = false;
    bool temporary = false;   // built inside pb2_flush from staged reads; freed when the flush returns
    bool from_reads = false;  // built by pb2_stage_reads: replaced by the next pb2_stage_reads, ignored by pb2_flush (which stages the reads itself)
    unsigned long long h_var_count = 0, h_exc_count = 0;
};

struct DevBlock { void* p; size_t bytes; bool used; };
struct pb2_handle {
    pb2_config cfg;
    std::vector<DevBlock> blocks;   // device blocks handed out by pb2_dev_alloc: in use, or idle and waiting for the next request of their size
    size_t cached_free_bytes = 0;
    pb2::DeviceConfig dcfg;
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_scattered[2] = {nullptr, nullptr};
    cudaEvent_t ev_piece[4] = {nullptr, nullptr, nullptr, nullptr};   // pacing of large host-to-device copies (h2d_in_pieces)
    cudaMemPool_t pool = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_stage0 = nullptr, ev_stage1 = nullptr;   // around the last reads -> PVERT staging
    bool have_stage_events = false;
    int64_t last_stage_bytes = 0, last_stage_rows = 0;
    std::string error;
    std::string chr_name;
    uint8_t* d_chr = nullptr;
    int gq_capped_vq = -1;          // the variant q-score the finished-GQ half of d_gq_tail was filled for
    double* d_gq_tail = nullptr;    // somatic-GQ Poisson tails [kGqTailMaxCov][kGqTailMaxA] (pb2_math.cuh:somatic_gq)
    double* d_q_to_p = nullptr;     // QtoP(q) for q = 0..max_variant_qscore (capped at 1024 entries)
    int q_table_max = -1;
    int64_t chr_len = 0;
    std::vector<uint8_t> h_chr;
    std::vector<int32_t> iv_start, iv_end;
    bool have_intervals = false;
    std::vector<Segment> segs;
    DeviceReads reads;
    int32_t own_lo = 0, own_hi = 0;   // pb2_set_owned_range (0, 0: everything)
    void* sink = nullptr;             // pb2_set_resident_sink: device buffer of the job's records, filled slot by slot by the resident steps
    int64_t sink_slot_records = 0, sink_next = 0;
    int32_t sink_slots = 0;
    int32_t cleared_through = 0;   // positions <= this were called by an earlier pb2_flush(up_to >= 0)
    int* d_tile_counter = nullptr;
    // (resize() of these does not zero-fill: a gVCF flush sizes them for a million records and then writes every byte from several threads)
    std::vector<pb2_call_record, DefaultInitAllocator<pb2_call_record>> h_out;
    std::vector<pb2_call_record_ext, DefaultInitAllocator<pb2_call_record_ext>> h_out_ext;   // parallel to h_out
    int64_t hot_launches = 0, total_launches = 0;
    double hot_ms = 0;
    // ---- explicit candidates (pb2_explicit.cu)
    std::vector<HostCand> cands;                      // uncleared candidates in RegionState.AddCandidate order
    std::unordered_map<int32_t, std::vector<uint32_t>> cand_by_pos;   // position -> indices into cands, in insertion order (explicit_reindex after a compaction)
    std::map<int32_t, int32_t> block_max_endpoint;    // RegionState.MaxAlleleEndpoint per 1000-bp block key
    std::map<int32_t, int32_t> gapped_ref;            // RegionState._gappedMnvReferenceCounts of uncleared positions
    std::vector<int32_t> triggers;                    // upTo values at which SmallVariantCaller.Execute would have called a batch
    int32_t last_trigger_key = 0;                     // RegionStateManager._lastUpToBlockKey
    int32_t push_last_key = 0;                        // the same, as seen while reads are pushed (which read positions open a batch)
    void* resident_explicit = nullptr;                // cached device plan of pb2_call_resident's explicit-candidate pass
    cudaGraphExec_t resident_graph = nullptr;         // the whole resident step (both streams) as one CUDA graph, replayed by pb2_call_resident
    bool resident_graph_failed = false;
    int64_t resident_graph_launches = 0;              // kernels inside the graph
    unsigned long long* h_counters = nullptr;         // pinned: the segment's counters after a step
    void* pin_refs = nullptr; size_t pin_refs_bytes = 0;     // pinned landing area of a flush's dense reference stream (gVCF: 96 B per locus) ...
    void* pin_valid = nullptr; size_t pin_valid_bytes = 0;   // ... and its validity bytes; grow-only, released by pb2_destroy
    std::vector<uint8_t> arena;                       // allele bytes the last flush's records point into
    std::vector<std::pair<int32_t, int32_t>> snv_explicit_ranges;   // (lo, hi] positions whose SNV candidates were made explicit (explicit_materialize_snvs)
    // ---- forced-genotyping alleles (pb2_set_forced_alleles)
    std::set<std::tuple<int32_t, std::string, std::string>> forced;                             // AlleleCaller.ForcedGtAlleles (position, ref, alt)
    std::map<int32_t, std::vector<std::pair<std::string, std::string>>> forced_pending;         // SmallVariantCaller._unProcessedForcedAllelesByPos
    std::vector<std::tuple<int32_t, std::string, std::string>> forced_order;                    // the HashSet's enumeration (= insertion) order
    std::vector<int32_t> forced_positions;   // one per forced allele inside the intervals, in that order: RegionState.CreateIntervalsFromAllels (:455-468)
    int64_t total_collapsed = 0;
    std::string vcf_text;                             // pb2_vcf_format's output
};

constexpr int32_t kCollapsedTotalTwice = INT32_MIN;   // marker in pb2_call_record_ext.collapsed_total[0] between explicit_call_batch and pb2_flush
// pb2_api.cu: device memory of a handle (pool + block cache, stream-ordered on the handle's stream)
cudaError_t pb2_dev_alloc(pb2_handle* h, void** p, size_t bytes);
void pb2_dev_free(pb2_handle* h, void* p);
// pb2_explicit.cu
int pb2_fail(pb2_handle* h, int code, const std::string& msg);
// RegionState.AddCandidate (:94-174): merge into the table (summing counts) or append; tracks MaxAlleleEndpoint of the block.
void explicit_add_candidate(pb2_handle* h, const HostCand& c);
void explicit_reindex(pb2_handle* h);
// The explicit-candidate part of AlleleCaller.Call for one batch of candidates (indices into h->cands, in batch order): VariantCollapser, MNV scoring
// + MnvReallocator, gapped-MNV reference counts, final ProcessVariant of every callable allele on the device. Called alleles (IsCallable &&
// ShouldReport) are appended to `called`; candidates that go back to the state (not cleared / MNV leftovers) are re-added to h->cands.
// max_cleared < 0 = null (everything is cleared). (ref_lo, ref_hi] are the positions of the batch's blocks: with forced alleles and reference calls off,
// RegionState.GetAllCandidates (:383-453) adds a Reference candidate at every forced position in them.
// `kill`: the candidates that leave the state when the batch is formed (default: the batch itself). RegionState.ExtractCollapsable (:470-490) removes
// with List.Remove — the first candidate that Equals (open ends ignored) — so the two can differ.
int explicit_call_batch(pb2_handle* h, const std::vector<size_t>& batch, int32_t max_cleared, int32_t ref_lo, int32_t ref_hi, std::vector<pb2_call_record>& called,
                        std::vector<pb2_call_record_ext>& called_ext, const std::vector<size_t>* kill = nullptr);
int explicit_materialize_snvs(pb2_handle* h, int32_t lo, int32_t hi);
// SmallVariantCaller.AddForcedAlleleAsCandidate (:118-155): forced alleles at positions <= up_to (< 0: all) become zero-support candidates.
int explicit_add_forced_candidates(pb2_handle* h, int32_t up_to);
// Finds the candidates of the stored reads [first_read, end) on the device (CandidateVariantFinder.FindCandidates) and adds them to the table. host_bases:
// the caller's bases array of the batch being pushed (read first_read starts at offset 0 of it), or nullptr (alleles longer than RawCand::read_bases
// are then fetched from the device store).
struct BatchHostView { const uint8_t* bases; const int64_t* seq_off; int64_t seq_lo; };   // bases points at the first base of the batch's first read
int explicit_find_candidates(pb2_handle* h, size_t first_read, const BatchHostView* host);
pb2::PvertPileup pvert_view(const Segment& s);
// pb2_call_resident: gather + score + append on the device, no host round trip; only for candidates that need neither the collapser nor the MNV logic.
int explicit_call_resident(pb2_handle* h, Segment& seg, cudaStream_t side);
bool explicit_resident_ready(pb2_handle* h);
int explicit_prune_resident(pb2_handle* h, Segment& seg);
void explicit_release_resident(pb2_handle* h);

